python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1_click or c2_tracks or intermediates or key_path or sample_rates or ragged or golden or stft_geometry or key_stft" > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02t_tests.log; tail -3 gpurun_out/r02t_tests.log
python bench.py --tracks 512 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02t_new.json 2>gpurun_out/r02t_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02t_new.json").read().strip().splitlines()[-1])
s=d["stages_ms_per_step"]
print("new", round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features","multires_features","stft_2048_hop512","stft_multires")})
PY

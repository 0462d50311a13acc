python -m pytest tests/test_gpu_parity.py tests/test_gpu_long.py -m gpu -x -q -k "c1_click or c2_tracks or intermediates or key_path or sample_rates or ragged or golden or c4_sixty or c5_ragged" > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02x_tests.log; tail -3 gpurun_out/r02x_tests.log
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_c4.json 2>gpurun_out/r02x_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02x_c4.json").read().strip().splitlines()[-1])
s=d["stages_ms_per_step"]
print("c4", round(d["ms_per_step"],1), "e2e", d.get("e2e",{}).get("value"), {k:round(v,1) for k,v in sorted(s.items(), key=lambda kv:-kv[1])[:12]})
PY
python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02x_new.json 2>gpurun_out/r02x_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02x_new.json").read().strip().splitlines()[-1])
s=d["stages_ms_per_step"]
print("c2/512", round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features","multires_features","stft_2048_hop512","stft_multires")})
PY

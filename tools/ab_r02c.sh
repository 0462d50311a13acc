set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "divisions or stft_bit or c1_click or c2_tracks or intermediates or accepted_config" > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log; tail -3 gpurun_out/r02c_tests.log
B="python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r02c_new.json 2>gpurun_out/r02c_new.err
STRATUM_B200_KEY_COMPACT=0 $B > gpurun_out/r02c_nocompact.json 2>/dev/null
python - <<'PY'
import json
for n in ("new","nocompact"):
    try:
        d=json.loads(open(f"gpurun_out/r02c_{n}.json").read().strip().splitlines()[-1])
        s=d["stages_ms_per_step"]
        print(n, round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features","multires_features","stft_2048_hop512")})
    except Exception as e: print(n, "failed", e)
PY

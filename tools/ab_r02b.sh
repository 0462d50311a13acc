set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft or c1_click or c2_tracks or intermediates or other_sample_rates or accepted_config" > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b_tests.log; tail -3 gpurun_out/r02b_tests.log
B="python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r02b_new.json 2>gpurun_out/r02b_new.err
STRATUM_B200_KEY_STFT_LEGACY=1 $B > gpurun_out/r02b_legacy_stft.json 2>/dev/null
STRATUM_B200_KEY_COMPACT=0 $B > gpurun_out/r02b_nocompact.json 2>/dev/null
python - <<'PY'
import json
for n in ("new","legacy_stft","nocompact"):
    try:
        d=json.loads(open(f"gpurun_out/r02b_{n}.json").read().strip().splitlines()[-1])
        s=d["stages_ms_per_step"]
        print(n, round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features","stft_2048_hop512")})
    except Exception as e: print(n, "failed", e)
PY

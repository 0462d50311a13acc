for g in 3 4; do
  STRATUM_B200_K12_GROUPS=$g python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02u_g$g.json 2>/dev/null
  python - "$g" <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02u_g{w}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("k12 groups", w, "value", round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features")})
PY
done
STRATUM_B200_K12_GROUPS=4 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft or c1_click or key_path" 2>&1 | tail -2

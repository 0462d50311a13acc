for w in 103 128 171 205 256; do
  STRATUM_B200_WAVE_MAX_TRACKS=$w python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02p_w$w.json 2>/dev/null
  python - "$w" <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02p_w{w}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("wave", w, "value", round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","key_mask","key_hpcp","spec_features","stft_2048_hop512")})
PY
done

python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft or c1_click or c2_tracks or intermediates or segmented or ragged or escalation" 2>&1 | tail -3
for v in new legacy; do
  if [ $v = legacy ]; then export STRATUM_B200_HOP_STFT_LEGACY=1; fi
  python bench.py --tracks 512 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02z_$v.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02z_{w}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("hop stft", w, "value", round(d["value"],1), {k:round(s[k],1) for k in ("stft_2048_hop512","stft_multires","spec_features","stft_8192_key","key_mask")})
PY
done

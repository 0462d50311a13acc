python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft or c1_click or c2_tracks or intermediates or key_path" 2>&1 | tail -2
python bench.py --tracks 512 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/quick_new.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/quick_new.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("value", round(d["value"],1), {k:round(s[k],1) for k in ("stft_8192_key","stft_2048_hop512","stft_multires","spec_features","key_mask","key_hpcp")})
PY

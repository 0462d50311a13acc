python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft or c1_click or c2_tracks or intermediates or escalation or ragged or golden or trap or multi_res or accepted_config" 2>&1 | tail -2
for v in share noshare; do
  if [ $v = noshare ]; then export STRATUM_B200_MULTIRES_NO_SHARE=1; fi
  python bench.py --tracks 512 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02P_$v.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02P_{w}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("multires", w, "value", round(d["value"],1), {k:round(s[k],1) for k in ("stft_2048_hop512","stft_multires","multires_features","spec_features","stft_8192_key")}, "esc", d["escalated_tracks"], "ok", d["results_ok"], d["bpm_within_2_or_octave"])
PY
done

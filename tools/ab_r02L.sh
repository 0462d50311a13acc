for r in "2,2" "1,2" "1,3" "2,3"; do
  STRATUM_B200_STAGE_RAMP=$r python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02L_$r.json 2>/dev/null
  python - "$r" <<'PY'
import json,sys
r=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02L_{r}.json").read().strip().splitlines()[-1])
print("ramp", r, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pcm16", round(d["e2e"]["pcm16_value"],1), "ratio", round(d["e2e"]["value"]/d["value"],3))
PY
done

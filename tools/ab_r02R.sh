for mb in 3876 4845 6238; do
  STRATUM_B200_STAGE_MB=$mb python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02R_$mb.json 2>/dev/null
  python - "$mb" <<'PY'
import json,sys
r=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02R_{r}.json").read().strip().splitlines()[-1])
print("stage MB", r, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pcm16", round(d["e2e"]["pcm16_value"],1), "ratio", round(d["e2e"]["value"]/d["value"],3))
PY
done

B="python bench.py --tracks 384 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
for m in 0 1 2 3; do STRATUM_B200_PARFEAT=$m $B > gpurun_out/r02h_m$m.json 2>/dev/null; done
python - <<'PY'
import json
for m in range(4):
    d=json.loads(open(f"gpurun_out/r02h_m{m}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
    print("mode",m, round(d["value"],1), {k:round(s[k],1) for k in ("spec_features","multires_features")})
PY

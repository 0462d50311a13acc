for v in 0 1 2 3 4 5; do
  STRATUM_B200_PAR_VARIANT=$v python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02s_v$v.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02s_v{w}.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("par variant", w, "value", round(d["value"],1), {k:round(s[k],1) for k in ("spec_features","multires_features","key_mask","key_hpcp")})
PY
done
STRATUM_B200_PAR_VARIANT=0 STRATUM_B200_SEQ_UNROLL8=1 python bench.py --tracks 512 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02s_seq8.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02s_seq8.json").read().strip().splitlines()[-1]); s=d["stages_ms_per_step"]
print("par 0 + seq8", "value", round(d["value"],1), {k:round(s[k],1) for k in ("spec_features","multires_features")})
PY

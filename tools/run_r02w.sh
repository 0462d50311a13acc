python -m pytest tests/test_gpu_long.py -m gpu -x -q > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log; tail -3 gpurun_out/r02w_tests.log
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_c4.json 2>gpurun_out/r02w_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w_c4.json").read().strip().splitlines()[-1])
s=d["stages_ms_per_step"]
print("c4", round(d["ms_per_step"],1), "e2e", d.get("e2e",{}).get("value"), {k:round(v,1) for k,v in sorted(s.items(), key=lambda kv:-kv[1])[:12]})
PY

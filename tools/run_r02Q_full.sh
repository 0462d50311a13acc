python -m pytest tests -m gpu -x -q > gpurun_out/r02Q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02Q_tests.log; tail -4 gpurun_out/r02Q_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02Q_bench.json 2> gpurun_out/r02Q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02Q_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pcm16", round(d["e2e"]["pcm16_value"],1))
print({k:round(v,1) for k,v in d["stages_ms_per_step"].items()})
PY

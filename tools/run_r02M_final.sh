python bench.py > gpurun_out/r02M_bench_1024tracks.json 2> gpurun_out/r02M_bench.err; echo "bench rc=$?"
python bench.py --workload c5 > gpurun_out/r02M_bench_c5_256tracks.json 2> gpurun_out/r02M_c5.err; echo "c5 rc=$?"
python bench.py --workload c4 > gpurun_out/r02M_bench_c4_60min.json 2> gpurun_out/r02M_c4.err; echo "c4 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02M_bench_reference_arm.json 2> gpurun_out/r02M_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("r02M_bench_1024tracks","r02M_bench_c5_256tracks","r02M_bench_c4_60min","r02M_bench_reference_arm"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    e=d.get("e2e",{})
    print(f, "value", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],1), "e2e", round(e.get("value",0),2), "pcm16", e.get("pcm16_value"), "roofline", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("single_worker",{}).get("value"))
PY

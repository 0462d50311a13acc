nvidia-smi -L | wc -l; free -g | head -2; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02n_bench_8gpu.json 2> gpurun_out/r02n_bench_8gpu.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r02n_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench_8gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pcm16", d["e2e"].get("pcm16_value"), "per_rank", round(d["e2e"]["per_rank"]["value"],1), d["e2e"]["per_rank"].get("pcm16_value"), d["e2e"].get("tracks_per_gpu"))
PY

set -x
nvidia-smi -L
free -g | head -2
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or repeated_device or concurrent" -rs > gpurun_out/r02e_2gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_2gpu_tests.log; tail -4 gpurun_out/r02e_2gpu_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02e_bench_2gpu.err

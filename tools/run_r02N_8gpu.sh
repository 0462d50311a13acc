set -x
nvidia-smi -L | wc -l
free -g | head -2
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or repeated_device or concurrent" -rs > gpurun_out/r02N_8gpu_multi_device_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02N_8gpu_multi_device_tests.log; tail -4 gpurun_out/r02N_8gpu_multi_device_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02N_bench_8gpu_one_call.json 2> gpurun_out/r02N_bench_8gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02N_bench_8gpu.err

set -x
python bench.py --tracks 256 --steps 2 --warmup 1 > gpurun_out/r02d_c2_256.json 2> gpurun_out/r02d_c2_256.err; echo "c2 rc=$?"
python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/r02d_c5.json 2> gpurun_out/r02d_c5.err; echo "c5 rc=$?"
python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/r02d_c4.json 2> gpurun_out/r02d_c4.err; echo "c4 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02d_ref.json 2> gpurun_out/r02d_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/r02d_*.err

set -x
python tools/prof_cmd.py 64 1 > gpurun_out/r02J_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stft_key12_kernel|mask_kernel|hpcp_kernel|par_feat_kernel|seq_feat_kernel|stft_hop10_kernel' -c 12 -o gpurun_out/r02J_prof -f python tools/prof_cmd.py 64 1 > gpurun_out/r02J_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02J_ncu.log
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02J_bench_for_launches.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02J_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02J_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02J_launches_default_bench.csv

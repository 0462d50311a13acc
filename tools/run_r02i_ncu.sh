set -x
python tools/prof_cmd.py 64 1 > gpurun_out/r02i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stft_key12_kernel|mask_kernel|hpcp_kernel' -c 3 -o gpurun_out/r02i_prof -f python tools/prof_cmd.py 64 1 > gpurun_out/r02i_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02i_ncu.log

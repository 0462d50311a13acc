set -x
python tools/prof_cmd.py 64 1 > gpurun_out/r02q_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stft_key12_kernel|mask_kernel|hpcp_kernel|par_feat_kernel|seq_feat_kernel|stft_tracks_kernel|energy_rms|silence_rms' -c 14 -o gpurun_out/r02q_prof -f python tools/prof_cmd.py 64 1 > gpurun_out/r02q_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r02q_ncu.log; ls -la gpurun_out/r02q_prof.ncu-rep

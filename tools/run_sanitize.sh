python tools/sanitize_case.py > gpurun_out/r02y_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/r02y_memcheck.log 2>&1
echo "sanitizer rc=$?"; tail -15 gpurun_out/r02y_memcheck.log

python -m pytest tests -m gpu -x -q > gpurun_out/r02T_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02T_tests.log; tail -4 gpurun_out/r02T_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

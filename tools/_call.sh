#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2c29_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2c29_sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2c29_sanitizer_racecheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nmf_gpu.py -q -m gpu -x -k "tensor_core_error_pass or row_sharded_iterations or any_feature" > gpurun_out/r2c29_sanitizer_memcheck_nmf_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -3 gpurun_out/r2c29_sanitizer_memcheck_nmf_tests.log

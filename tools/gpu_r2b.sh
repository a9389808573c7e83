#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_ilv_small.py 4 > gpurun_out/r2b_small.log 2>&1; echo "small exit $?" >> gpurun_out/r2b_small.log; cat gpurun_out/r2b_small.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_ilv_small.py 4 > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2b_memcheck.log; tail -25 gpurun_out/r2b_memcheck.log
timeout 900 python -m pytest tests/test_gpu_ilv.py -q -s > gpurun_out/r2b_ilv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_ilv.log; tail -40 gpurun_out/r2b_ilv.log
timeout 900 python tools/gpu_ilv_ab.py 1024 8 > gpurun_out/r2b_ab.log 2>&1; cat gpurun_out/r2b_ab.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_ilv.py > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log; tail -60 gpurun_out/r2b_pytest.log

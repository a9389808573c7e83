#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_c2000.py 12 > gpurun_out/r2c_c2000.log 2>&1; echo "exit $?" >> gpurun_out/r2c_c2000.log; tail -40 gpurun_out/r2c_c2000.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_closed_loop.py -q -s -k "case2000 or generic_lane or jump_model or box_fallback or error_paths or empty_rows" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log; tail -30 gpurun_out/r2c_pytest.log

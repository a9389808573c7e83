#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_ring_ab.py 1024 6 "default" "default unfused" 2>&1 | tail -2
timeout 600 python tools/gpu_c2000.py 6 2>&1 | tail -3
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ring.py -q -x ) > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log; tail -5 gpurun_out/r2l_pytest.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ring.py -x -q -s > gpurun_out/r2h_ring_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_ring_tests.log; tail -6 gpurun_out/r2h_ring_tests.log
SQPQP_PROF=1 timeout 300 python tools/gpu_ring_ab.py 148 2 2>&1 | grep -E "kcycles|ms/round" | tail -3
timeout 600 python tools/gpu_ring_ab.py 1024 5 2>&1 | tail -3
timeout 600 python tools/gpu_ring_ab.py 128 5 2>&1 | tail -3

#!/bin/bash
# ring (TMA-streamed index programs) first light: parity tests, then A/B with the in-kernel profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ring.py -x -q -s > gpurun_out/r2f_ring_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_ring_tests.log; tail -25 gpurun_out/r2f_ring_tests.log
timeout 600 python tools/gpu_ring_ab.py 1024 6 > gpurun_out/r2f_ab1024.log 2>&1; cat gpurun_out/r2f_ab1024.log
timeout 600 python tools/gpu_ring_ab.py 128 6 > gpurun_out/r2f_ab128.log 2>&1; cat gpurun_out/r2f_ab128.log
SQPQP_PROF=1 timeout 600 python tools/gpu_ring_ab.py 148 3 > gpurun_out/r2f_prof148.log 2>&1; grep -E "kcycles|ms/round" gpurun_out/r2f_prof148.log | tail -12

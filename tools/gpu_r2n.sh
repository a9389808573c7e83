#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ring.py -q -x -s ) > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log; tail -12 gpurun_out/r2n_pytest.log
timeout 900 python tools/gpu_ring_ab.py 1024 60 "no handoff" "handoff 32" "handoff 40" "default" "handoff 56" "handoff 64" 2>&1 | tail -6
timeout 900 python tools/gpu_ring_ab.py 256 60 "no handoff" "handoff 40" "default" "handoff 64" 2>&1 | tail -4

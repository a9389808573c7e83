#!/bin/bash
# round-2 GPU call A: all GPU tests (old + closed-loop), trajectories, baseline bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python tools/gpu_traj.py > gpurun_out/r2a_traj.log 2>&1; tail -30 gpurun_out/r2a_traj.log
timeout 600 python bench.py --no-spmv --no-device-eval > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 3000 gpurun_out/r2a_bench.json

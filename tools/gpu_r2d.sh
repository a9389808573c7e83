#!/bin/bash
# round-2 GPU call D (re-entry baseline): all GPU tests, bench line, in-kernel phase profile, 2000-bus timing, ncu launch list
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -30 gpurun_out/r2d_pytest.log
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"; tail -c 1500 gpurun_out/r2d_bench.json
SQPQP_PROF=1 timeout 600 python tools/gpu_prof.py 1024 8 > gpurun_out/r2d_prof.log 2>&1; tail -12 gpurun_out/r2d_prof.log
timeout 900 python tools/gpu_c2000.py 8 > gpurun_out/r2d_c2000.log 2>&1; tail -14 gpurun_out/r2d_c2000.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-spmv --no-device-eval --no-cpu-baseline --sqp-max-iter 16 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu exit $?"

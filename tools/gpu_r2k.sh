#!/bin/bash
# round-2 evidence run: bench line, ncu launch list, ncu --set full of the three solve kernels
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-spmv --no-device-eval --no-cpu-baseline --no-single2000 --sqp-max-iter 16 > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 4 -c 1 -o gpurun_out/r02_batch1024 -f python tools/gpu_ring_one.py 1024 3 > gpurun_out/r02_ncu_batch.log 2>&1; echo "ncu batch exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 2 -c 1 -o gpurun_out/r02_ring148 -f python tools/gpu_ring_one.py 148 2 > gpurun_out/r02_ncu_ring.log 2>&1; echo "ncu ring exit $?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_solve_grid -s 1 -c 1 -o gpurun_out/r02_grid2000 -f python tools/gpu_c2000.py 2 > gpurun_out/r02_ncu_grid.log 2>&1; echo "ncu grid exit $?"
ls -la gpurun_out | tail -12

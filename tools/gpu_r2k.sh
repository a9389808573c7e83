#!/bin/bash
# round-2 evidence run: bench line, ncu launch list, ncu --set full of the three solve kernels (reports are reduced to CSV /
# JSON summaries on the box: gpurun copies at most 64 MiB back)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-spmv --no-device-eval --no-cpu-baseline --no-single2000 --sqp-max-iter 16 > $O/r02_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 4 -c 1 -o $O/r02_batch1024 -f python tools/gpu_ring_one.py 1024 3 > $O/r02_ncu_batch.log 2>&1; echo "ncu batch exit $?"
python tools/ncu_traffic.py $O/r02_batch1024.ncu-rep $O/r02_traffic.json "k_solve_cta<384,2,1>" 1024 batch118 > $O/r02_ncu_batch_summary.txt 2>&1
ncu -i $O/r02_batch1024.ncu-rep --page source --csv > $O/r02_ncu_batch_source.csv 2>/dev/null; rm -f $O/r02_batch1024.ncu-rep
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 2 -c 1 -o $O/r02_ring148 -f python tools/gpu_ring_one.py 148 2 > $O/r02_ncu_ring.log 2>&1; echo "ncu ring exit $?"
python tools/ncu_traffic.py $O/r02_ring148.ncu-rep $O/r02_traffic_ring148.json "k_solve_cta<512,1,1>+ring" 148 batch118 > $O/r02_ncu_ring_summary.txt 2>&1
ncu -i $O/r02_ring148.ncu-rep --page source --csv > $O/r02_ncu_ring_source.csv 2>/dev/null; rm -f $O/r02_ring148.ncu-rep
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_solve_grid -s 1 -c 1 -o $O/r02_grid2000 -f python tools/gpu_c2000.py 2 > $O/r02_ncu_grid.log 2>&1; echo "ncu grid exit $?"
python tools/ncu_traffic.py $O/r02_grid2000.ncu-rep $O/r02_traffic_2000.json "k_solve_grid" 1 single2000 > $O/r02_ncu_grid_summary.txt 2>&1
ncu -i $O/r02_grid2000.ncu-rep --page source --csv > $O/r02_ncu_grid_source.csv 2>/dev/null; rm -f $O/r02_grid2000.ncu-rep
du -sh $O; ls -la $O | tail -16

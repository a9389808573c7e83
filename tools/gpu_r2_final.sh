#!/bin/bash
# final evidence run of round 2: GPU test suite, bench line, ncu launch list, ncu --set full of one launch of the dominant kernel
# in the shape the grouped bench launches it (256 instances), reduced to CSV / JSON summaries on the box
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -q -m gpu -x) > $O/r02_pytest_final.log 2>&1; echo "pytest exit $?"; tail -n 4 $O/r02_pytest_final.log
timeout 600 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-spmv --no-device-eval --no-cpu-baseline --no-single2000 --sqp-max-iter 16 > $O/r02_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 4 -c 1 -o $O/r02_batch256 -f python tools/gpu_ring_one.py 256 3 handoff=0 > $O/r02_ncu_batch.log 2>&1; echo "ncu batch exit $?"
ALG=$(grep "^round 2:" $O/r02_ncu_batch.log | sed 's/.*algorithmic bytes \([0-9.e+]*\).*/\1/')
python tools/ncu_traffic.py $O/r02_batch256.ncu-rep $O/r02_traffic.json "k_solve_cta<384,2,1>" 256 batch118 $ALG > $O/r02_ncu_batch_summary.txt 2>&1
ncu -i $O/r02_batch256.ncu-rep --page source --csv > $O/r02_ncu_batch_source.csv 2>/dev/null; rm -f $O/r02_batch256.ncu-rep
grep "^round" $O/r02_ncu_batch.log; head -n 16 $O/r02_ncu_batch_summary.txt
du -sh $O

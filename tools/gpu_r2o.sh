#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "device_acopf or device_evaluator or merit" ) 2>&1 | tail -4
C2000_ITERS=6 timeout 600 python tests/devtools/gpu_sqp.py toy,readme,case9,case9_soc,c118 --no-oracle 2>&1 | grep DEVICE | cut -c1-200
timeout 900 python bench.py --no-spmv --no-cpu-baseline --no-single2000 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])
print(d['full_sqp_solve']); print(d.get('full_sqp_solve_device_evaluator'))
PY

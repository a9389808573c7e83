#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_closed_loop.py tests/test_gpu_parity.py -q -x -k "case2000 or jump_model or shared_subsolver or registered_host or caller_owned" ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -8 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','single_instance_2000','solver'):
    v=d.get(k)
    if isinstance(v,dict): v={a:b for a,b in v.items() if a not in('note','sample','unit','phases')}
    print(k, v)
PY

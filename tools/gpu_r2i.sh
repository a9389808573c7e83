#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log

#!/bin/bash
# E1: in-kernel phase profile of the resident (one CTA per SM) configuration on the big batch; PCIe link check
mkdir -p gpurun_out
for cfg in "occupancy=1 threads=512" "occupancy=1 threads=384" "occupancy=1 threads=256"; do
  echo "=== $cfg" >> gpurun_out/r2e_prof.log
  SQPQP_PROF=1 timeout 300 python tools/gpu_prof.py 1024 3 $cfg >> gpurun_out/r2e_prof.log 2>&1
done
echo "=== B=148 default" >> gpurun_out/r2e_prof.log
SQPQP_PROF=1 timeout 300 python tools/gpu_prof.py 148 3 >> gpurun_out/r2e_prof.log 2>&1
grep -E "===|round   [23]|kcycles" gpurun_out/r2e_prof.log | tail -40
timeout 120 python tools/gpu_pcie.py > gpurun_out/r2e_pcie.log 2>&1; cat gpurun_out/r2e_pcie.log

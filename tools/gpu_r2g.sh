#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_solve_cta -s 2 -c 1 -o gpurun_out/r2g_ring148 -f python tools/gpu_ring_one.py 148 2 > gpurun_out/r2g_ncu.log 2>&1; tail -3 gpurun_out/r2g_ncu.log
ls -la gpurun_out/

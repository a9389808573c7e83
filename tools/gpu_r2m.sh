#!/bin/bash
for B in 256 512 1024; do echo "== B=$B"; timeout 900 python tools/gpu_ring_ab.py $B 60 "default" "occ1 ring" 2>&1 | tail -2; done

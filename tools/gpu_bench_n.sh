#!/bin/bash
# usage: bash tools/gpu_bench_n.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 16 --warmup 4 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "exit $?"
tail -c 1800 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err

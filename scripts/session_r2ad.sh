#!/bin/bash
# round-2 session AD (8 GPUs): the driver's 8-GPU bench command with the final library
N=${1:-8}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2ad_bench_n${N}.json 2> gpurun_out/r2ad_bench_n${N}.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/r2ad_bench_n${N}.json; tail -2 gpurun_out/r2ad_bench_n${N}.err

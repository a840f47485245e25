#!/bin/bash
# round-2 session X (2 GPUs): distributed parity and the driver's 2-GPU bench command with the final library
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/r2x_dist${N}.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2x_dist${N}.log
timeout 300 $TR --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x_bench_n${N}.json 2> gpurun_out/r2x_bench_n${N}.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2x_bench_n${N}.json; tail -2 gpurun_out/r2x_bench_n${N}.err
timeout 120 $TR --master-port 29533 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x_bench_ref_n${N}.json 2> gpurun_out/r2x_bench_ref_n${N}.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2x_bench_ref_n${N}.json

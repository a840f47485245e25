#!/bin/bash
# One 8-GPU session: distributed parity (both protocols), the bench line (66^3 block + config 5), the NCCL A/B line,
# and BASELINE config 3 as a transient on 2x2x2 boxes.  Usage: gpurun --gpus 8 -- bash scripts/run_8gpu.sh [N]
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/r2_dist${N}.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2_dist${N}.log
$TR --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_n${N}.json
$TR --master-port 29522 bench.py --gpus $N --steps 100 --warmup 5 --halo 0 --allreduce 0 --config5-n 0 > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "bench nccl rc=$?"; cat gpurun_out/r2_bench_n${N}_nccl.json
$TR --master-port 29523 bench.py --gpus $N --steps 100 --warmup 5 --overlap 0 --config5-n 0 > gpurun_out/r2_bench_n${N}_nooverlap.json 2> gpurun_out/r2_bench_n${N}_nooverlap.err; echo "bench no-overlap rc=$?"; cat gpurun_out/r2_bench_n${N}_nooverlap.json
$TR --master-port 29524 scripts/transient_bench.py --elems 156 --steps 3 > gpurun_out/r2_transient_n${N}.json 2> gpurun_out/r2_transient_n${N}.err; echo "transient rc=$?"; cat gpurun_out/r2_transient_n${N}.json; tail -2 gpurun_out/r2_transient_n${N}.err
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1

#!/bin/bash
# Final 8-GPU session of round 2: host-link probe with 1 / 2 / 4 / 8 ranks active, distributed parity of the final
# library, the bench line.  Usage: gpurun --gpus 8 -- bash scripts/run_8gpu_final.sh
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29531 scripts/pcie_multi_probe.py > gpurun_out/r2_pcie_n${N}.jsonl 2> gpurun_out/r2_pcie_n${N}.err; echo "pcie rc=$?"; cat gpurun_out/r2_pcie_n${N}.jsonl
python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/r2_dist${N}_final.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2_dist${N}_final.log
$TR --master-port 29532 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_n${N}_final.json; tail -2 gpurun_out/r2_bench_n${N}_final.err

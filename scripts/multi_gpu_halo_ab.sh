#!/bin/bash
# NCCL send/recv vs peer-memory halo exchange, same box, N GPUs (run through `gpurun --gpus N`):
#   scripts/multi_gpu_halo_ab.sh N [steps]
# 1. parity of both paths against the un-partitioned oracle (tests/dist_check.py --p2p),
# 2. bench.py with --halo 0 / 1, serial (--overlap 0) and default schedule.
# Every command is wrapped in `timeout`: a rank that never posts must not wedge the box.
set -u
N=${1:-2}
STEPS=${2:-100}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
echo "== parity (order 3 and 2)"
for o in 3 2; do
   run 2951$o tests/dist_check.py --mode gpu --order $o --mesh 8 6 6 --p2p > gpurun_out/halo_ab_parity_o$o.log 2>&1
   echo "order $o rc=$?"; grep -o "\[rank 0\] [^\[]*" gpurun_out/halo_ab_parity_o$o.log | grep "peer-memory\|gmres" | cut -c1-120
done
echo "== bench"
for ov in 0 1; do
   for h in 0 1; do
      run 2952$ov$h bench.py --gpus "$N" --steps "$STEPS" --warmup 10 --no-cpu --overlap $ov --halo $h --krylov-iters 30 2>/dev/null |
         python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap=$ov', d['config']['halo'], 'GDOF/s', round(d['value'],3), 'ms/apply', round(d['ms_per_step'],4), 'ms/krylov-iter', round(d['krylov']['ms_per_iter'],4))"
   done
done

#!/bin/bash
# One profiling session on one B200 (gpurun -- bash scripts/profile_session.sh):
#   1. plain runs (must exit 0 before anything runs under ncu)
#   2. per-launch device times of the default bench.py line (shares of the step)
#   3. ncu --set full of one launch of every kernel of the path (NVTX-filtered)
set -x
mkdir -p gpurun_out
python scripts/profile_targets.py > gpurun_out/r2_targets_plain.log 2>&1 || { tail -20 gpurun_out/r2_targets_plain.log; exit 1; }
python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 30 --config5-n 0 > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err || { tail -20 gpurun_out/r2_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 30 --config5-n 0 > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "prof/" -o gpurun_out/r2_prof_targets -f \
    python scripts/profile_targets.py > gpurun_out/r2_ncu_targets.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -5 gpurun_out/r2_ncu_targets.log

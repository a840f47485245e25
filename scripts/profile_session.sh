#!/bin/bash
# One profiling session on one B200 (gpurun -- bash scripts/profile_session.sh):
#   1. plain runs (must exit 0 before anything runs under ncu)
#   2. per-launch device times of the default bench.py line (shares of the step)
#   3. one launch of every kernel of the path (NVTX-filtered) with the sections the summaries use
#   4. ncu --set full + source of the headline kernel only
# Reports are converted to CSV on the box and the big .ncu-rep files removed (gpurun_out/ is limited to 64 MiB).
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 30 --config5-n 0"
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --section SchedulerStats"
python scripts/profile_targets.py > gpurun_out/r2_targets_plain.log 2>&1 || { tail -20 gpurun_out/r2_targets_plain.log; exit 1; }
$BENCH > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err || { tail -20 gpurun_out/r2_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv $BENCH > gpurun_out/r2_ncu_launches.log 2>&1
ncu $SECT --clock-control none --nvtx --nvtx-include "prof/" -o /tmp/r2_prof_targets -f python scripts/profile_targets.py > gpurun_out/r2_ncu_targets.log 2>&1
ncu -i /tmp/r2_prof_targets.ncu-rep --page raw --csv > gpurun_out/r2_prof_targets_raw.csv 2> gpurun_out/r2_prof_export.err
ncu --set full --clock-control none --import-source on -k regex:k_apply3d_warp_bg -s 4 -c 1 -o gpurun_out/r2_prof_p3 -f $BENCH > gpurun_out/r2_ncu_p3.log 2>&1
ncu -i gpurun_out/r2_prof_p3.ncu-rep --page raw --csv > gpurun_out/r2_prof_p3_raw.csv 2>> gpurun_out/r2_prof_export.err
ls -la gpurun_out/ /tmp/*.ncu-rep
du -sh gpurun_out
tail -3 gpurun_out/r2_ncu_targets.log

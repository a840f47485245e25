#!/bin/bash
# round-2 session AB: ncu --set full of the final high-order kernels at 8 M dofs (one launch each) and of the headline kernel at
# config 2 (refreshes profiles/traffic.json); un-profiled runs first
mkdir -p gpurun_out
for p in 4 5 6; do
  python scripts/sweep.py --dofs 8e6 --orders $p --steps 2 > gpurun_out/r2ab_plain_p$p.jsonl 2> gpurun_out/r2ab_plain_p$p.err || { tail -5 gpurun_out/r2ab_plain_p$p.err; exit 1; }
  ncu --set full --clock-control none -k regex:k_apply3d_group -s 3 -c 1 -o /tmp/r2ab_p$p -f python scripts/sweep.py --dofs 8e6 --orders $p --steps 2 > gpurun_out/r2ab_ncu_p$p.log 2>&1
  ncu -i /tmp/r2ab_p$p.ncu-rep --page raw --csv > gpurun_out/r2ab_p${p}_raw.csv 2>> gpurun_out/r2ab_export.err
  cat gpurun_out/r2ab_plain_p$p.jsonl
done
python scripts/sweep.py --dofs 7.88e6 --orders 3 --steps 2 > gpurun_out/r2ab_plain_p3.jsonl 2> gpurun_out/r2ab_plain_p3.err; cat gpurun_out/r2ab_plain_p3.jsonl
ncu --set full --clock-control none -k regex:k_apply3d_warp_bg -s 3 -c 1 -o /tmp/r2ab_p3 -f python scripts/sweep.py --dofs 7.88e6 --orders 3 --steps 2 > gpurun_out/r2ab_ncu_p3.log 2>&1
ncu -i /tmp/r2ab_p3.ncu-rep --page raw --csv > gpurun_out/r2ab_p3_raw.csv 2>> gpurun_out/r2ab_export.err
ls -la gpurun_out/r2ab_*raw.csv

#!/bin/bash
# round-2 session O: ncu --set full of the apply kernels at ~50 M dofs (config 4), one launch each: where the gap
# between 8 M and 50 M dofs comes from (DRAM traffic vs algorithmic bytes, L2 hit rates, stalls)
mkdir -p gpurun_out
for p in 3 6 5 4; do
  python scripts/sweep.py --dofs 5e7 --orders $p --steps 2 > gpurun_out/r2o_plain_p$p.jsonl 2> gpurun_out/r2o_plain_p$p.err || { tail -5 gpurun_out/r2o_plain_p$p.err; exit 1; }
  ncu --set full --clock-control none -k regex:k_apply3d -s 3 -c 1 -o /tmp/r2o_p$p -f python scripts/sweep.py --dofs 5e7 --orders $p --steps 2 > gpurun_out/r2o_ncu_p$p.log 2>&1
  ncu -i /tmp/r2o_p$p.ncu-rep --page raw --csv > gpurun_out/r2o_p${p}_raw.csv 2>> gpurun_out/r2o_export.err
  cat gpurun_out/r2o_plain_p$p.jsonl
done
ls -la gpurun_out/r2o_* | head -20

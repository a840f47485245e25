#!/bin/bash
# round-2 session D: full-set ncu captures (with source) of the high-order group kernel and the 2D thread kernel
mkdir -p gpurun_out
python scripts/profile_targets.py --only apply3d_p4,apply3d_p5,apply3d_p6,apply2d_p2,apply2d_p3 --gmres-steps 0 > gpurun_out/r2d_plain.log 2>&1 || { tail gpurun_out/r2d_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "prof/" -k regex:"k_apply3d_group|k_apply2d_thread" -o gpurun_out/r2d_group -f python scripts/profile_targets.py --only apply3d_p4,apply3d_p5,apply3d_p6,apply2d_p2,apply2d_p3 --gmres-steps 0 > gpurun_out/r2d_ncu.log 2>&1
tail -5 gpurun_out/r2d_ncu.log
ls -la gpurun_out/r2d_group.ncu-rep

#!/bin/bash
# round-2 session C: 2D kernel variants, config-2 steady solve
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "2d" 2>&1 | tail -3
for v in base 2dv1 2dv2; do
  if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$PWD/continuum-mechanics-mfem_b200/libcdm_b200_$v.so; fi
  python scripts/sweep.py --dim 2 --dofs 8e6 --orders 1 2 3 4 --steps 20 > gpurun_out/r2_sweep2d_8M_$v.jsonl 2> gpurun_out/r2_sweep2d_$v.err
  echo "== $v"; python - <<PY
import json
for l in open("gpurun_out/r2_sweep2d_8M_$v.jsonl"):
    r=json.loads(l); print(r["order"], round(r["kernel_ms"],4), round(r["roofline_frac"],3))
PY
  tail -2 gpurun_out/r2_sweep2d_$v.err
done
unset CDM_B200_LIB
timeout 600 python scripts/steady_solve.py --n 66 --skip-assembled > gpurun_out/r2_steady_n66.jsonl 2>gpurun_out/r2_steady_n66.err; cat gpurun_out/r2_steady_n66.jsonl; tail -3 gpurun_out/r2_steady_n66.err

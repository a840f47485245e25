#!/bin/bash
# round-2 session G: 2D thread kernel, two-level gather prefetch A/B
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "2d or dim2 or quad" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
for rep in 1 2; do
  for v in base pf1 pf3; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== $v 8e6 rep$rep"; python scripts/sweep.py --dim 2 --dofs 8e6 --orders 1 2 3 4 --steps 20 > gpurun_out/r2g_sweep2d_8e6_${v}_$rep.jsonl 2>> gpurun_out/r2g_err.log; show gpurun_out/r2g_sweep2d_8e6_${v}_$rep.jsonl
  done
done
unset CDM_B200_LIB
echo "== base 5e7"; python scripts/sweep.py --dim 2 --dofs 5e7 --orders 1 2 3 4 --steps 20 > gpurun_out/r2g_sweep2d_5e7_base.jsonl 2>> gpurun_out/r2g_err.log; show gpurun_out/r2g_sweep2d_5e7_base.jsonl
tail -3 gpurun_out/r2g_err.log

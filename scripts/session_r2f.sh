#!/bin/bash
# round-2 session F: in-session A/B of the scatter fix (sub kernel p=1,2; group kernel p=4,5,6) and the barrier form
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
for rep in 1 2; do
for d in 8e6 5e7; do
  for v in base old regbar; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    orders="1 2 4 5 6"; [ $v = regbar ] && orders="5 6"
    echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders $orders --steps 20 > gpurun_out/r2f_sweep_${d}_${v}_$rep.jsonl 2>> gpurun_out/r2f_err.log; show gpurun_out/r2f_sweep_${d}_${v}_$rep.jsonl
  done
done
done
unset CDM_B200_LIB
tail -3 gpurun_out/r2f_err.log

#!/bin/bash
# round-2 session M: why the aliased exchange buffers lose at p = 5 (resident blocks?), wait-all in the order-3 kernel
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
export CDM_CFG_DEBUG=1
for v in base a5m4 noalias_waitall; do
  if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
  echo "== $v 8e6"; python scripts/sweep.py --dofs 8e6 --orders 5 6 --steps 20 > gpurun_out/r2m_sweep_${v}.jsonl 2> gpurun_out/r2m_err_$v.log; show gpurun_out/r2m_sweep_${v}.jsonl; grep "cdm\]" gpurun_out/r2m_err_$v.log | grep group
done
CDM_B200_LIB=$L/libcdm_b200_p3wait.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "config2_apply or capped" 2>&1 | tail -2
for rep in 1 2; do for d in 8e6 5e7; do for v in base p3wait; do
  if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
  echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders 3 --steps 20 > gpurun_out/r2m_p3_${d}_${v}_$rep.jsonl 2>> gpurun_out/r2m_err.log; show gpurun_out/r2m_p3_${d}_${v}_$rep.jsonl
done; done; done

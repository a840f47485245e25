#!/bin/bash
# round-2 session Y: 2D order 4 with the dof indices re-read before the scatter (A/B); new edge-case / odd-count tests
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; tail -2 gpurun_out/r2y_pytest.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["dim"], r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
for rep in 1 2; do
  for v in base nogreload; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== 2D $v 8e6 rep$rep"; python scripts/sweep.py --dim 2 --dofs 8e6 --orders 4 --steps 20 > gpurun_out/r2y_sweep2d_${v}_$rep.jsonl 2> gpurun_out/r2y_err_${v}.log; show gpurun_out/r2y_sweep2d_${v}_$rep.jsonl
  done
done

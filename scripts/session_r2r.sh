#!/bin/bash
# round-2 session R: plain stores instead of red.add for the dofs inside an element (they belong to one element), A/B
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; tail -2 gpurun_out/r2r_pytest.log
CDM_B200_LIB=$L/libcdm_b200_p3inner.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "config2 or midsize or capped" > gpurun_out/r2r_pytest_p3inner.log 2>&1; tail -2 gpurun_out/r2r_pytest_p3inner.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
for rep in 1 2; do
  for v in base noinner p3inner; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    o="3 4 5 6"; if [ $v = p3inner ]; then o="3"; fi; if [ $v = noinner ]; then o="4 5 6"; fi
    echo "== $v 8e6 rep$rep"; python scripts/sweep.py --dofs 8e6 --orders $o --steps 20 > gpurun_out/r2r_sweep_${v}_$rep.jsonl 2> gpurun_out/r2r_err_${v}.log; show gpurun_out/r2r_sweep_${v}_$rep.jsonl
  done
done

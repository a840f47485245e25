#!/bin/bash
# round-2 session AA: order 5 with 32-bit indices + wait-all (164 registers: six groups still fit), A/B; repetition test; order 4
# after the ballot-mask reconvergence change
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1; tail -2 gpurun_out/r2aa_pytest.log
CDM_B200_LIB=$L/libcdm_b200_g5iw.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "capped or midsize or repeated" > gpurun_out/r2aa_pytest_g5iw.log 2>&1; tail -1 gpurun_out/r2aa_pytest_g5iw.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["dim"], r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
export CDM_CFG_DEBUG=1
for rep in 1 2; do
  for v in base g5iw; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    o="4 5"; if [ $v = g5iw ]; then o="5"; fi
    echo "== $v 8e6 rep$rep"; python scripts/sweep.py --dofs 8e6 --orders $o --steps 20 > gpurun_out/r2aa_sweep_${v}_$rep.jsonl 2> gpurun_out/r2aa_err_${v}.log; show gpurun_out/r2aa_sweep_${v}_$rep.jsonl
    if [ $rep = 1 ]; then grep "cdm\]" gpurun_out/r2aa_err_$v.log | grep group; fi
  done
done
for v in base g5iw; do
  if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
  echo "== $v 5e7 burst"; python scripts/sweep.py --dofs 5e7 --orders 5 --steps 2 > gpurun_out/r2aa_sweep50_${v}.jsonl 2>> gpurun_out/r2aa_err_${v}.log; show gpurun_out/r2aa_sweep50_${v}.jsonl
done

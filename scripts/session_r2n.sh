#!/bin/bash
# round-2 session N: in-place y contractions at p = 5 (6 resident groups, no extra barrier), p = 6 in-place vs aliased
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "midsize or capped or group or config2_apply" 2>&1 | tail -2
CDM_B200_LIB=$L/libcdm_b200_g6inplace.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "capped or midsize" 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
export CDM_CFG_DEBUG=1
for rep in 1 2; do
for d in 8e6 5e7; do
  for v in base g6inplace g5wait old; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders 4 5 6 --steps 20 > gpurun_out/r2n_sweep_${d}_${v}_$rep.jsonl 2> gpurun_out/r2n_err_${v}.log; show gpurun_out/r2n_sweep_${d}_${v}_$rep.jsonl
    if [ $rep = 1 ] && [ $d = 8e6 ]; then grep "cdm\]" gpurun_out/r2n_err_$v.log | grep group; fi
  done
done
done

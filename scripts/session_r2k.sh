#!/bin/bash
# round-2 session K: order-4 group kernel, both warps in the x / y exchange stages (A/B against the one-warp stages)
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "midsize or capped or group" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
for rep in 1 2; do
for d in 8e6 5e7; do
  for v in base nobal; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders 4 --steps 20 > gpurun_out/r2k_sweep_${d}_${v}_$rep.jsonl 2>> gpurun_out/r2k_err.log; show gpurun_out/r2k_sweep_${d}_${v}_$rep.jsonl
  done
done
done
unset CDM_B200_LIB
for sc in 0; do python scripts/sweep.py --dofs 8e6 --orders 4 --steps 20 --scatter 0 2>> gpurun_out/r2k_err.log | python -c "import json,sys; r=json.loads(sys.stdin.read()); print('scatter0', r['kernel_ms'], r['ms_per_apply'])"; done
tail -3 gpurun_out/r2k_err.log

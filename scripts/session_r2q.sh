#!/bin/bash
# round-2 session Q: z contraction + point-wise D + transposed z contraction fused per quadrature level, A/B per order;
# deterministic scatter at p = 4 with the fused form (the separate stages spilled ~500 B)
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "midsize or capped or group" > gpurun_out/r2q_pytest.log 2>&1; tail -2 gpurun_out/r2q_pytest.log
CDM_B200_LIB=$L/libcdm_b200_fuse.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "midsize or capped" > gpurun_out/r2q_pytest_fuse.log 2>&1; tail -2 gpurun_out/r2q_pytest_fuse.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
export CDM_CFG_DEBUG=1
for rep in 1 2; do
for d in 8e6; do
  for v in base fuse f5w; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    o="4 5 6"; if [ $v = f5w ]; then o="5"; fi
    echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders $o --steps 20 > gpurun_out/r2q_sweep_${d}_${v}_$rep.jsonl 2> gpurun_out/r2q_err_${v}.log; show gpurun_out/r2q_sweep_${d}_${v}_$rep.jsonl
    if [ $rep = 1 ]; then grep "cdm\]" gpurun_out/r2q_err_$v.log | grep group; fi
  done
done
done
unset CDM_B200_LIB
echo "== base scatter 0"; python scripts/sweep.py --dofs 8e6 --orders 3 4 5 6 --steps 20 --scatter 0 > gpurun_out/r2q_sweep_scatter0.jsonl 2>> gpurun_out/r2q_err.log; show gpurun_out/r2q_sweep_scatter0.jsonl

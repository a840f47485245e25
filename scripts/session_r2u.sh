#!/bin/bash
# round-2 session U: order 5 with five warps for three elements, 32-bit element indices in the group kernel, A/B
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; tail -2 gpurun_out/r2u_pytest.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
export CDM_CFG_DEBUG=1
for rep in 1 2; do
  for v in base nopenta; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== $v 8e6 rep$rep"; python scripts/sweep.py --dofs 8e6 --orders 4 5 6 --steps 20 > gpurun_out/r2u_sweep_${v}_$rep.jsonl 2> gpurun_out/r2u_err_${v}.log; show gpurun_out/r2u_sweep_${v}_$rep.jsonl
    if [ $rep = 1 ]; then grep "cdm\]" gpurun_out/r2u_err_$v.log | grep group; fi
  done
done
unset CDM_B200_LIB
echo "== base 5e7 burst"; python scripts/sweep.py --dofs 5e7 --orders 4 5 6 --steps 2 > gpurun_out/r2u_sweep50_base.jsonl 2>> gpurun_out/r2u_err_base.log; show gpurun_out/r2u_sweep50_base.jsonl
echo "== base scatter 0"; python scripts/sweep.py --dofs 8e6 --orders 4 5 6 --steps 20 --scatter 0 > gpurun_out/r2u_sweep_scatter0.jsonl 2>> gpurun_out/r2u_err.log; show gpurun_out/r2u_sweep_scatter0.jsonl

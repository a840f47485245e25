#!/bin/bash
# round-2 session W: 32-bit element indices in the sub-warp kernel (3D orders 1, 2) and in the 2D kernel at orders 1, 2, A/B
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
CDM_B200_LIB=$L/libcdm_b200_subidx.so python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2w_pytest_subidx.log 2>&1; tail -1 gpurun_out/r2w_pytest_subidx.log
CDM_B200_LIB=$L/libcdm_b200_2didx1.so python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -x -k "2d" > gpurun_out/r2w_pytest_2didx1.log 2>&1; tail -1 gpurun_out/r2w_pytest_2didx1.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["dim"], r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
for rep in 1 2; do
  for v in base subidx; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== 3D $v 8e6 rep$rep"; python scripts/sweep.py --dofs 8e6 --orders 1 2 --steps 20 > gpurun_out/r2w_sweep_${v}_$rep.jsonl 2> gpurun_out/r2w_err_${v}.log; show gpurun_out/r2w_sweep_${v}_$rep.jsonl
  done
  for v in base 2didx1; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== 2D $v 8e6 rep$rep"; python scripts/sweep.py --dim 2 --dofs 8e6 --orders 1 2 --steps 20 > gpurun_out/r2w_sweep2d_${v}_$rep.jsonl 2>> gpurun_out/r2w_err_${v}.log; show gpurun_out/r2w_sweep2d_${v}_$rep.jsonl
  done
done
for v in base subidx; do
  if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
  echo "== 3D $v 3e7 p2 (config 3 size) burst"; python scripts/sweep.py --dofs 3.008e7 --orders 2 --steps 3 > gpurun_out/r2w_sweep30_${v}.jsonl 2>> gpurun_out/r2w_err_${v}.log; show gpurun_out/r2w_sweep30_${v}.jsonl
done

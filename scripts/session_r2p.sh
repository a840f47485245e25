#!/bin/bash
# round-2 session P: padded half-warp lane maps of the x / y exchange roles at p = 5, 6 (ideal wavefront counts), A/B
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "midsize or capped or group" > gpurun_out/r2p_pytest.log 2>&1; tail -2 gpurun_out/r2p_pytest.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
for rep in 1 2; do
for d in 8e6 5e7; do
  for v in base nolanemap; do
    if [ $v = base ]; then unset CDM_B200_LIB; else export CDM_B200_LIB=$L/libcdm_b200_$v.so; fi
    echo "== $v $d rep$rep"; python scripts/sweep.py --dofs $d --orders 5 6 --steps 20 > gpurun_out/r2p_sweep_${d}_${v}_$rep.jsonl 2> gpurun_out/r2p_err_${v}.log; show gpurun_out/r2p_sweep_${d}_${v}_$rep.jsonl
  done
done
done

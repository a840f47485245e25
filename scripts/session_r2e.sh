#!/bin/bash
# round-2 session E: group / sub kernel scatter fix, p=4 block shape variants, p3 split accumulators, ILU value-as-flag sweeps
mkdir -p gpurun_out
L=$PWD/continuum-mechanics-mfem_b200
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "midsize or capped or group or subwarp" 2>&1 | tail -3
python -m pytest tests/test_gpu_simplex_ilu.py -m gpu -q -x -k ilu0 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2))
PY
}
for d in 8e6 5e7; do
  echo "== base $d"; python scripts/sweep.py --dofs $d --orders 1 2 3 4 5 6 --steps 20 > gpurun_out/r2e_sweep_${d}_base.jsonl 2> gpurun_out/r2e_err.log; show gpurun_out/r2e_sweep_${d}_base.jsonl
  for v in g4m7 g4two; do
    echo "== $v $d"; CDM_B200_LIB=$L/libcdm_b200_$v.so python scripts/sweep.py --dofs $d --orders 4 --steps 20 > gpurun_out/r2e_sweep_${d}_$v.jsonl 2>> gpurun_out/r2e_err.log; show gpurun_out/r2e_sweep_${d}_$v.jsonl
  done
  echo "== p3split $d"; CDM_B200_LIB=$L/libcdm_b200_p3split.so python scripts/sweep.py --dofs $d --orders 3 --steps 20 > gpurun_out/r2e_sweep_${d}_p3split.jsonl 2>> gpurun_out/r2e_err.log; show gpurun_out/r2e_sweep_${d}_p3split.jsonl
done
tail -3 gpurun_out/r2e_err.log
timeout 300 python scripts/steady_solve.py --n 32 --skip-levels > gpurun_out/r2e_steady_n32.jsonl 2>gpurun_out/r2e_steady_n32.err; cat gpurun_out/r2e_steady_n32.jsonl | cut -c1-400; tail -3 gpurun_out/r2e_steady_n32.err

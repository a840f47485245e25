#!/bin/bash
# round-2 session AC: final confirmation with the library as committed -- full GPU suite, smoke, the driver's bench commands
# (both arms), order sweeps (8 M 20-apply, 50 M 2-apply, 2D)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2ac_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2ac_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2ac_smoke.log 2>&1; tail -1 gpurun_out/r2ac_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ac_bench_reference.json 2> gpurun_out/r2ac_bench_reference.err; cut -c1-200 gpurun_out/r2ac_bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ac_bench_n1.json 2> gpurun_out/r2ac_bench_n1.err; cut -c1-1300 gpurun_out/r2ac_bench_n1.json; tail -2 gpurun_out/r2ac_bench_n1.err
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["dim"], r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
python scripts/sweep.py --dofs 8e6 --steps 20 > gpurun_out/r2ac_sweep_8M.jsonl 2> gpurun_out/r2ac_err.log; show gpurun_out/r2ac_sweep_8M.jsonl
python scripts/sweep.py --dofs 5e7 --steps 2 > gpurun_out/r2ac_sweep_50M_burst.jsonl 2>> gpurun_out/r2ac_err.log; show gpurun_out/r2ac_sweep_50M_burst.jsonl
python scripts/sweep.py --dim 2 --dofs 8e6 --steps 20 --orders 1 2 3 4 > gpurun_out/r2ac_sweep2d_8M.jsonl 2>> gpurun_out/r2ac_err.log; show gpurun_out/r2ac_sweep2d_8M.jsonl
tail -3 gpurun_out/r2ac_err.log

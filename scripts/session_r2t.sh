#!/bin/bash
# round-2 session T: evidence refresh with the final kernels -- full GPU test suite, smoke, the driver's bench commands
# (both arms), order sweeps (8 M, 50 M sustained and burst, 2D), per-launch list of the bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2t_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; tail -2 gpurun_out/r2t_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t_bench_reference.json 2> gpurun_out/r2t_bench_reference.err; cut -c1-400 gpurun_out/r2t_bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; cut -c1-1200 gpurun_out/r2t_bench_n1.json; tail -2 gpurun_out/r2t_bench_n1.err
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(r["dim"], r["order"], r["dofs"], round(r["kernel_ms"],4), round(r["roofline_frac"],3), round(r["gdofs"],2), round(r["ms_per_apply"],4))
PY
}
python scripts/sweep.py --dofs 8e6 --steps 20 > gpurun_out/r2t_sweep_8M.jsonl 2> gpurun_out/r2t_err.log; show gpurun_out/r2t_sweep_8M.jsonl
python scripts/sweep.py --dofs 5e7 --steps 20 > gpurun_out/r2t_sweep_50M.jsonl 2>> gpurun_out/r2t_err.log; show gpurun_out/r2t_sweep_50M.jsonl
python scripts/sweep.py --dofs 5e7 --steps 2 > gpurun_out/r2t_sweep_50M_burst.jsonl 2>> gpurun_out/r2t_err.log; show gpurun_out/r2t_sweep_50M_burst.jsonl
python scripts/sweep.py --dim 2 --dofs 8e6 --steps 20 --orders 1 2 3 4 > gpurun_out/r2t_sweep2d_8M.jsonl 2>> gpurun_out/r2t_err.log; show gpurun_out/r2t_sweep2d_8M.jsonl
BENCH="python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 30 --config5-n 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2t_launches_bench.csv $BENCH > gpurun_out/r2t_ncu_launches.log 2>&1
tail -2 gpurun_out/r2t_ncu_launches.log; tail -3 gpurun_out/r2t_err.log

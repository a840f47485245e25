#!/bin/bash
# round-2 session B: FP64 pipe probe, ILU sweeps, steady solves, 2D thread-per-element kernel
mkdir -p gpurun_out
scripts/build/fp64_pipe_probe > gpurun_out/r2_fp64_probe.json 2>&1; cat gpurun_out/r2_fp64_probe.json
python -m pytest tests/test_gpu_simplex_ilu.py -m gpu -q -x -k ilu0 2>&1 | tail -5
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py -m gpu -q -x -k "2d or 2D or dim2 or quad" 2>&1 | tail -8
python scripts/sweep.py --dim 2 --dofs 8e6 --orders 1 2 3 4 --steps 20 > gpurun_out/r2_sweep2d_8M_thread.jsonl 2> gpurun_out/r2_sweep2d.err; cat gpurun_out/r2_sweep2d_8M_thread.jsonl; tail -3 gpurun_out/r2_sweep2d.err
python scripts/sweep.py --dim 2 --dofs 8e6 --orders 1 2 3 4 --steps 20 --kernel 0 > gpurun_out/r2_sweep2d_8M_block.jsonl 2>> gpurun_out/r2_sweep2d.err; cat gpurun_out/r2_sweep2d_8M_block.jsonl
python scripts/sweep.py --dim 2 --dofs 5e7 --orders 1 2 3 4 --steps 20 > gpurun_out/r2_sweep2d_50M_thread.jsonl 2>> gpurun_out/r2_sweep2d.err; cat gpurun_out/r2_sweep2d_50M_thread.jsonl
timeout 300 python scripts/steady_solve.py --n 16 > gpurun_out/r2_steady_n16.jsonl 2>gpurun_out/r2_steady_n16.err; cat gpurun_out/r2_steady_n16.jsonl; tail -3 gpurun_out/r2_steady_n16.err
timeout 500 python scripts/steady_solve.py --n 32 > gpurun_out/r2_steady_n32.jsonl 2>gpurun_out/r2_steady_n32.err; cat gpurun_out/r2_steady_n32.jsonl; tail -3 gpurun_out/r2_steady_n32.err

#!/bin/bash
# round-2 session H: full GPU suite after the helper refactor, host-link probe on one GPU, e2e stage timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
python scripts/pcie_multi_probe.py > gpurun_out/r2h_pcie_n1.jsonl 2>gpurun_out/r2h_pcie.err; cat gpurun_out/r2h_pcie_n1.jsonl; tail -2 gpurun_out/r2h_pcie.err
python scripts/pcie_overlap_probe.py 2>&1 | tail -4
CDM_PIPE_DEBUG=1 python scripts/e2e_probe.py > gpurun_out/r2h_e2e.log 2>&1; grep -v "cdm pipe" gpurun_out/r2h_e2e.log; grep "cdm pipe" gpurun_out/r2h_e2e.log | tail -3

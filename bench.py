#!/usr/bin/env python
"""bench.py -- headline benchmark of BASELINE.json: operator-apply GDOF/s (and Krylov
iteration time) of the 3D hex order-3 convection-diffusion operator.

  python bench.py --gpus N --steps K --warmup W            (N=1: BASELINE config 2)
  torchrun ... bench.py --gpus N ...                       (N>1: same block per GPU, weak scaling)
  python bench.py --impl reference ...                     (CPU arm: the reference's path on host cores)

A "step" is one constrained operator apply  y = A x  (mfem::Operator::Mult inside the
Krylov loop, linear_convection_diffusion_2D.cpp:368-370) over this job's whole mesh.
Rank 0 prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KAPPA, VEL, MASS = 0.1, (1.0, -2.0, 0.5), 1.0          # Input/input_2d.yaml:7-10 (+ z component)
PERTURB = 0.1


def algorithmic_bytes(ndof, ne, p, dim=3):
    """read x + write y once, every stored D value once, the int32 gather map once"""
    d1d, q1d = p + 1, (p + 2 if dim == 3 else p + 1)
    n_d = 10 if dim == 3 else 6
    return 16 * ndof + ne * (8 * n_d * q1d ** dim + 4 * d1d ** dim)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = str(index), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.index], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(order, n_csr, n_pa, steps, warmup, budget_s=25.0):
    """The reference's CPU path on this box's host cores, from the oracle port:
    (i) assembled CSR SpMV = what the app executes (full assembly -> ParCSR -> PETSc MatMult),
    (ii) CPU sum-factorised PA apply.  Bounded sample of the same operator."""
    from oracle import pyoracle as orc
    cores = orc.num_threads()
    out = {"cores": cores, "kind": "port"}
    P = orc.Problem(3, order, n_csr, perturb=PERTURB, kappa=KAPPA, vel=VEL, mass=MASS)
    x = np.sin(1.0 + 0.37 * np.arange(P.ndof))
    A = P.csr()
    for _ in range(max(1, warmup)):
        A.spmv(x)
    t0, k = time.perf_counter(), 0
    while k < steps and (k < 3 or time.perf_counter() - t0 < budget_s / 2):
        A.spmv(x)
        k += 1
    t_csr = (time.perf_counter() - t0) / k
    out.update(csr_spmv_gdofs=P.ndof / t_csr / 1e9, csr_ms=t_csr * 1e3, csr_steps=k, csr_ndof=P.ndof)
    P2 = orc.Problem(3, order, n_pa, perturb=PERTURB, kappa=KAPPA, vel=VEL, mass=MASS)
    x2 = np.sin(1.0 + 0.37 * np.arange(P2.ndof))
    y2 = np.zeros(P2.ndof)
    P2.pa_apply_fast(x2, y2)
    t0, k2 = time.perf_counter(), 0
    while k2 < steps and (k2 < 3 or time.perf_counter() - t0 < budget_s / 2):
        P2.pa_apply_fast(x2, y2)          # fused, order-specialised CPU partial-assembly apply
        k2 += 1
    t_pa = (time.perf_counter() - t0) / k2
    out.update(pa_apply_gdofs=P2.ndof / t_pa / 1e9, pa_ms=t_pa * 1e3, pa_steps=k2, pa_ndof=P2.ndof)
    out["value"] = out["csr_spmv_gdofs"]
    out["unit"] = "GDOF/s"
    out["sample"] = (f"assembled CSR SpMV on {n_csr}^3 hex p={order} ({P.ndof} dofs, {len(A.vals)} nnz), {k} applies; "
                     f"CPU PA apply on {n_pa}^3 ({P2.ndof} dofs), {k2} applies; OpenMP over rows/elements")
    return out


def parts_for(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n, (n, 1, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=66, help="elements per axis per GPU (66 -> 7 880 599 dofs, config 2)")
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--kernel", type=int, default=-1)
    ap.add_argument("--scatter", type=int, default=-1)
    ap.add_argument("--overlap", type=int, default=1, help="multi-GPU: overlap the halo exchange with interior elements")
    ap.add_argument("--halo", type=int, default=0, help="multi-GPU: 0 NCCL send/recv, 1 peer-memory stores + flags (experimental)")
    ap.add_argument("--krylov-iters", type=int, default=60)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    workload = (f"3D hex H1 order {args.order}, {args.n}^3 perturbed elements per GPU, Diffusion+Convection+Mass "
                f"(kappa={KAPPA}, c={VEL}, s={MASS}), all-Dirichlet constrained apply")

    if args.impl == "reference":
        if rank != 0:
            return 0
        ref = cpu_reference(args.order, 24, 32, K, W)
        t_ms = ref["csr_ms"]
        line = {"impl": "reference", "metric": "operator_apply_gdofs", "value": ref["value"], "unit": "GDOF/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": t_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "reference_path": "assembled CSR SpMV (what the app executes), CPU oracle port; "
                           "MFEM/hypre/PETSc cannot be built here"},
                "cpu_baseline": ref,
                "e2e": {"value": ref["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if world > 1:
        # the halo messages (<= 0.7 MB) and the Krylov all-reduces (<= 31 doubles) are latency-bound:
        # NCCL's LL protocol is 5.6 % faster per Krylov iteration at 8 GPUs (profiles/r01_scaling.md)
        os.environ.setdefault("NCCL_PROTO", "LL")
    # stdout carries exactly one JSON line: anything a library writes to fd 1 (NCCL prints its version
    # banner there on some boxes) is sent to stderr, and the result goes to a private copy of the real stdout
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    cdm = importlib.import_module("continuum-mechanics-mfem_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = cdm.Context(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(cdm.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    parts = parts_for(world)
    gmesh = cdm.Mesh.cartesian(ctx, 3, [args.n * parts[0], args.n * parts[1], args.n * parts[2]], perturb=PERTURB)
    mesh = gmesh.partition_box(parts, rank) if world > 1 else gmesh
    sp = cdm.H1Space(mesh, args.order)
    ess = sp.essential_dofs(np.ones(6, np.int32))
    op = cdm.ConvectionDiffusionOperator(sp, kappa=KAPPA, vel=VEL, mass=MASS, ess_dofs=ess)
    if args.kernel >= 0:
        op.set_option("kernel", args.kernel)
    if args.scatter >= 0:
        op.set_option("scatter", args.scatter)
    op.set_option("overlap", args.overlap)
    if world > 1 and args.halo:
        op.set_option("halo", args.halo)
    op.set_option("tail", 1)       # x, y are allocated with the local (L-vector) size: no T<->L copies
    n_true = sp.ntrue
    tot = torch.tensor([n_true, sp.ne], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    n_global, ne_global = int(tot[0]), int(tot[1])
    ld = sp.ndof
    x = torch.sin(1.0 + 0.37 * torch.arange(ld, dtype=torch.float64, device="cuda"))
    y = torch.zeros(ld, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident apply: K steps, CUDA events on the launching stream, max over ranks
    for _ in range(W):
        op.Mult(x, y)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        op.Mult(x, y)
    e1.record(stream)
    barrier()
    launches = ctx.launches - l0
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms[0]) / K
    gdofs = n_global / (ms_step * 1e-3) / 1e9

    # ---- kernel-only time of the dominant kernel (roofline)
    k_ms = op.time_kernel(x, y, reps=min(K, 20), constrained=True)
    bytes_launch = algorithmic_bytes(sp.ndof, sp.ne, args.order)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_launch / (k_ms * 1e-3) / 1e9
    traffic = None                     # DRAM bytes per launch from the committed ncu --set full capture (same config only)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        if args.n == 66 and args.order == 3 and world == 1:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": bytes_launch,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}

    # ---- setup kernels (a.Assemble(): quadrature data; Jacobi diagonal), device time of one call each
    dvec = torch.zeros(ld, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    op.update(kappa=KAPPA, vel=VEL, mass=MASS)
    op.AssembleDiagonal(dvec)
    ctx.sync()
    s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s0.record(stream)
    op.update(kappa=KAPPA, vel=VEL, mass=MASS)
    s1.record(stream)
    op.AssembleDiagonal(dvec)
    s2.record(stream)
    ctx.sync()
    setup = {"qdata_ms": s0.elapsed_time(s1), "diag_ms": s1.elapsed_time(s2)}

    # ---- end to end through the host-buffer entry point (mfem::Vector under Device("cpu"))
    xh = torch.empty(n_true, dtype=torch.float64).pin_memory()
    yh = torch.empty(n_true, dtype=torch.float64).pin_memory()
    xh.copy_(x[:n_true].cpu())
    xh_np, yh_np = xh.numpy(), yh.numpy()
    K2 = max(3, min(K, 10))
    for _ in range(2):
        op.mult_host(xh_np, yh_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K2):
        op.mult_host(xh_np, yh_np)
    barrier()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / K2], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": n_global / float(t_e2e[0]) / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * n_true,
           "d2h_bytes_per_step": 8 * n_true, "ms_per_step": float(t_e2e[0]) * 1e3, "steps": K2}

    # ---- Krylov iteration time: GMRES(30)/CGS + Jacobi (Input/petsc.opts:2-6), fixed iteration count
    b = torch.sin(0.5 + 0.11 * torch.arange(n_true, dtype=torch.float64, device="cuda"))
    xs = torch.zeros(ld, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    s = cdm.GMRESSolver(cdm.GMRES_PETSC, 0, args.krylov_iters, 0.0, 0.0, jacobi=True)
    s.SetOperator(op)
    s.Mult(b, xs)            # warm-up (allocates the basis)
    barrier()
    l1 = ctx.launches
    s.Mult(b, xs)
    barrier()
    kry_launches = ctx.launches - l1
    t_k = torch.tensor([s.res.seconds], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_k, op=dist.ReduceOp.MAX)
    krylov = {"iters": s.GetNumIterations(), "ms_per_iter": float(t_k[0]) * 1e3 / max(1, s.GetNumIterations()),
              "launches": kry_launches, "config": "GMRES(30) classical Gram-Schmidt, left Jacobi, x0=0"}
    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(args.order, 24, 32, 50, 2)

    if rank == 0:
        line = {"metric": "operator_apply_gdofs", "value": gdofs, "unit": "GDOF/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "global_dofs": n_global, "global_elements": ne_global,
                           "partition": "x".join(map(str, parts)),
                           "nccl_proto": os.environ.get("NCCL_PROTO") if world > 1 else None,
                           "halo": ("peer memory" if args.halo else "nccl send/recv") if world > 1 else None, "l2_policy": "inputs larger than L2 "
                           f"({bytes_launch / 1e9:.2f} GB streamed per apply per GPU vs 126 MB L2)",
                           "scatter": "fp64 red.add" if op_scatter(op, args) == 1 else "E-vector + gather transpose"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "krylov": krylov, "setup": setup, "clocks": clocks}
        print(json.dumps(line), file=result_out)
        result_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def op_scatter(op, args):
    return 1 if args.scatter < 0 else args.scatter


if __name__ == "__main__":
    sys.exit(main())

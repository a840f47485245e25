#!/usr/bin/env python
"""Burst vs sustained behaviour of the apply kernel on one B200: per-chunk kernel times of a long back-to-back run
next to nvidia-smi clocks / power / temperature sampled every 20 ms, for several kernel options and for a plain
device-to-device copy of the same byte count (the roofline denominator under the same conditions).

    python scripts/sustained_probe.py [--n 66] [--order 3] [--seconds 2.5]
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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Smi:
    Q = "timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"

    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                                  stdout=subprocess.PIPE, text=True)
        self.t = threading.Thread(target=self._rd, daemon=True)
        self.t.start()

    def _rd(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        r = [x for t, x in self.rows if t0 <= t <= t1 and len(x) >= 8]
        if not r:
            return None
        f = lambda i: [float(x[i]) for x in r if x[i].replace(".", "").isdigit()]
        sm, mem, pw, tc = f(1), f(2), f(3), f(4)
        return {"samples": len(r), "sm_mhz_min": min(sm), "sm_mhz_median": float(np.median(sm)), "mem_mhz_min": min(mem),
                "power_w_median": float(np.median(pw)), "power_w_max": max(pw), "temp_c_max": max(tc),
                "sw_power_cap_frac": sum(x[5].lower().startswith("active") for x in r) / len(r),
                "hw_slowdown": any(x[6].lower().startswith("active") for x in r),
                "sw_thermal": any(x[7].lower().startswith("active") for x in r)}

    def stop(self):
        self.p.terminate()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=66)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=2.5)
    ap.add_argument("--kernels", type=int, nargs="*", default=None)
    args = ap.parse_args()
    import torch
    cdm = importlib.import_module("continuum-mechanics-mfem_b200")
    ctx = cdm.Context(0)
    mesh = cdm.Mesh.cartesian(ctx, 3, args.n, perturb=0.1)
    sp = cdm.H1Space(mesh, args.order)
    ess = sp.essential_dofs(np.ones(6, np.int32))
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
    x = torch.sin(1.0 + 0.37 * torch.arange(sp.ndof, dtype=torch.float64, device="cuda"))
    y = torch.zeros_like(x)
    stream = torch.cuda.ExternalStream(ctx.stream)
    d1d, q1d = args.order + 1, args.order + 2
    nbytes = 16 * sp.ndof + sp.ne * (80 * q1d ** 3 + 4 * d1d ** 3)
    smi = Smi()
    time.sleep(0.5)
    out = {"n": args.n, "order": args.order, "algorithmic_bytes": nbytes, "phases": []}
    default = op.get_option("kernel")
    kernels = args.kernels if args.kernels is not None else [default, 0, 4]

    def run_phase(name, fn, chunk):
        for _ in range(3):
            fn()
        ctx.sync(); torch.cuda.synchronize()
        time.sleep(1.0)                                   # cool-down: every phase starts from idle
        evs = []
        t0 = time.time()
        tw0 = time.perf_counter()
        while time.perf_counter() - tw0 < args.seconds:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(chunk):
                fn()
            e1.record(stream)
            e1.synchronize()
            evs.append(e0.elapsed_time(e1) / chunk)
        t1 = time.time()
        ms = np.array(evs)
        k = max(1, len(ms) // 10)
        ph = {"phase": name, "chunks": len(ms), "chunk_applies": chunk, "ms_first": float(ms[:k].mean()), "ms_last": float(ms[-k:].mean()),
              "ms_min": float(ms.min()), "ms_max": float(ms.max()), "gbs_first": nbytes / ms[:k].mean() / 1e6, "gbs_last": nbytes / ms[-k:].mean() / 1e6,
              "smi": smi.window(t0 + 0.3, t1)}
        out["phases"].append(ph)
        print(json.dumps(ph), flush=True)

    for kv in kernels:
        op.set_option("kernel", kv)
        run_phase(f"apply kernel={kv}", lambda: op.Mult(x, y), 20)
    op.set_option("kernel", default)
    # plain copy of the same number of bytes (read n/2, write n/2), on the same stream
    a = torch.empty(nbytes // 16, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    with torch.cuda.stream(stream):
        run_phase("d2d copy, same bytes", lambda: b.copy_(a), 20)
    smi.stop()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

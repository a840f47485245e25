#!/usr/bin/env python
"""Host-link ceiling of the end-to-end (host-vector) path at N GPUs: every active rank copies one config-2 vector
(63 MB, pinned) host->device and one device->host, (a) one direction at a time, (b) both directions at once on two
streams, with k = 1, 2, 4, ... N ranks active at the same time (the others wait at the barrier).  Reports GB/s per rank
and in aggregate, i.e. what `cdm_operator_mult_host` can at best achieve per step when N processes share the host's
memory system and PCIe root complexes.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/pcie_multi_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 7880599
REPS = 20
xh = torch.empty(N, dtype=torch.float64).pin_memory()
yh = torch.empty(N, dtype=torch.float64).pin_memory()
xh.fill_(1.0)
xd = torch.empty(N, dtype=torch.float64, device="cuda")
yd = torch.ones(N, dtype=torch.float64, device="cuda")
su, sd = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = N * 8


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, active):
    barrier()
    if active:
        fn()
    barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(REPS):
            fn()
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / REPS
    t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def h2d():
    with torch.cuda.stream(su):
        xd.copy_(xh, non_blocking=True)
    su.synchronize()


def d2h():
    with torch.cuda.stream(sd):
        yh.copy_(yd, non_blocking=True)
    sd.synchronize()


SPLIT = [torch.cuda.Stream() for _ in range(4)]


def h2d_split(k):
    """the same upload cut into k pieces on k streams (k copy engines reading host memory at once)"""
    def fn():
        for i in range(k):
            a, b = N * i // k, N * (i + 1) // k
            with torch.cuda.stream(SPLIT[i]):
                xd[a:b].copy_(xh[a:b], non_blocking=True)
        for i in range(k):
            SPLIT[i].synchronize()
    return fn


def both():
    with torch.cuda.stream(su):
        xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(sd):
        yh.copy_(yd, non_blocking=True)
    su.synchronize(); sd.synchronize()


k = 1
rows = []
while k <= world:
    act = rank < k
    r = {"active_ranks": k}
    for name, fn, vol in (("h2d", h2d, nbytes), ("d2h", d2h, nbytes), ("both", both, 2 * nbytes), ("h2d_again", h2d, nbytes),
                          ("h2d_2streams", h2d_split(2), nbytes), ("h2d_4streams", h2d_split(4), nbytes)):
        t = timed(fn, act)
        r[name + "_ms"] = round(t * 1e3, 4)
        r[name + "_GBs_per_rank"] = round(vol / t / 1e9, 2)
        r[name + "_GBs_aggregate"] = round(k * vol / t / 1e9, 2)
    rows.append(r)
    if rank == 0:
        print(json.dumps(r), flush=True)
    k *= 2
if world > 1:
    dist.destroy_process_group()

"""Distributed parity check of the mesh-partitioned mode, run with one process per rank:

  CPU / gloo  (host logic only: partition plan + P / P^T semantics; local element work is
               done by the ORACLE, so this runs without a GPU):
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
             tests/dist_check.py --mode cpu
  GPU / NCCL  (the product path: CUDA apply + NCCL halo exchange + all-reduced Krylov dots):
      python -m torch.distributed.run --nproc-per-node N ... tests/dist_check.py --mode gpu

Every rank compares its owned part of  y = A x  (and of the GMRES solution / residual
history) with the oracle's result on the un-partitioned mesh.  Exit code 0 = parity.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cdm_b200 as cdm  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1), 6: (3, 2, 1), 8: (2, 2, 2)}


def lattice_keys(P):
    """global lattice key of every dof of an un-partitioned oracle problem (Cartesian, lexicographic elements)"""
    p, d1d, dim = P.p, P.p + 1, P.dim
    n = list(P.n) + [0] * (3 - dim)
    N0, N1 = p * n[0] + 1, p * n[1] + 1
    keys = np.zeros(P.ndof, np.int64)
    e = np.arange(P.ne)
    ei, ej, ek = e % n[0], (e // n[0]) % n[1], (e // (n[0] * n[1]) if dim == 3 else 0 * e)
    for l in range(P.nd):
        lx, ly, lz = l % d1d, (l // d1d) % d1d, (l // (d1d * d1d) if dim == 3 else 0)
        keys[P.elem_dof[:, l]] = (p * ei + lx) + N0 * ((p * ej + ly) + N1 * (p * ek + lz))
    return keys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu", choices=["cpu", "gpu"])
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--mesh", type=int, nargs=3, default=[6, 5, 4])
    ap.add_argument("--p2p", action="store_true", help="also check the peer-memory P / P^T exchange (option halo=1)")
    ap.add_argument("--partition", default="box", choices=["box", "checker", "random", "slabs"],
                    help="box: cdm_mesh_partition_box; the others: cdm_mesh_partition_elements with a per-element rank array "
                    "(checker: (i+j+k) mod ranks, every element face is a partition boundary; random: seeded; slabs: z-slabs)")
    ap.add_argument("--halo", type=int, default=2, help="shared-dof protocol of the main checks: 2 symmetric peer-memory (default), 0 NCCL")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    gpu = args.mode == "gpu"
    if gpu:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx = cdm.Context(local_rank)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(cdm.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, uid.cpu().numpy().tobytes())
    else:
        dist.init_process_group("gloo")
        ctx = cdm.Context(host_only=True)
    p, parts = args.order, PARTS[world]
    kap, vel, mass = 0.1, (1.0, -2.0, 0.5), 1.0

    # ---- oracle on the un-partitioned mesh
    P = orc.Problem(3, p, args.mesh, perturb=0.1, kappa=kap, vel=vel, mass=mass)
    gkeys = lattice_keys(P)
    inv = np.zeros(gkeys.max() + 1, np.int64)
    inv[gkeys] = np.arange(P.ndof)
    xg = np.sin(1.0 + 0.37 * gkeys)
    yg = P.pa_op(True).mult(xg)                                   # constrained apply, all-Dirichlet
    rng = np.random.default_rng(5)
    bg = rng.uniform(-1, 1, P.ndof)[np.argsort(np.argsort(gkeys))]  # any fixed vector keyed by lattice position
    bg = np.where(P.ess_mark, 0.0, bg)
    dg = np.where(P.ess_mark, 1.0, P.pa_diag())
    xs_g, info = P.pa_op(True).gmres(bg, dinv=1 / dg, variant=0, rtol=1e-10, atol=1e-12, max_it=500)

    # ---- this rank's part
    gm = cdm.Mesh.cartesian(ctx, 3, args.mesh, perturb=0.1)
    if args.partition == "box":
        lm = gm.partition_box(parts, rank)
    else:
        # METIS-style input: one rank id per element of the replicated parent mesh (ParMesh(comm, mesh), :300)
        e = np.arange(P.ne)
        n0, n1 = args.mesh[0], args.mesh[1]
        ei, ej, ek = e % n0, (e // n0) % n1, e // (n0 * n1)
        if args.partition == "checker":
            er = (ei + ej + ek) % world
        elif args.partition == "slabs":
            er = np.minimum(ek * world // args.mesh[2], world - 1)
        else:
            er = np.random.default_rng(7).integers(0, world, P.ne)
            er[:world] = np.arange(world)                         # every rank owns something
        lm = gm.partition_elements(er, world, rank)
        assert lm.ne == int((er == rank).sum())
    sp = cdm.H1Space(lm, p)
    keys = sp.dof_global()
    nt = sp.ntrue
    # box parts key their dofs by lattice position; element-wise parts by the global dof id of the parent space, which is the
    # oracle's own numbering (bit-exact, tests/test_host_abi.py)
    mine = inv[keys] if args.partition == "box" else keys        # local dof -> global oracle dof
    assert np.abs(sp.dof_coords() - P.coords()[mine]).max() < 1e-13
    tot = torch.tensor([nt], dtype=torch.int64, device="cuda" if gpu else "cpu")
    dist.all_reduce(tot)
    assert int(tot[0]) == P.ndof, (int(tot[0]), P.ndof)
    ess = sp.essential_dofs(np.ones(6, np.int32))
    assert np.all(P.ess_mark[mine[ess]] == 1) and P.ess_mark[mine].sum() == len(ess)

    ok = True
    if gpu:
        op = cdm.ConvectionDiffusionOperator(sp, kappa=kap, vel=vel, mass=mass, ess_dofs=ess)
        op.set_option("halo", args.halo)
        print(f"[rank {rank}] shared-dof protocol requested: {args.halo}", flush=True)
        xd = torch.from_numpy(xg[mine[:nt]]).cuda()
        yd = torch.zeros(nt, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        # overlap: 0 serial halo exchange, 1 automatic (by neighbour count), 2 forced overlapped schedule
        for kernel, scatter, overlap in ((0, 0, 0), (0, 1, 0), (1, 1, 0), (3, 0, 0), (3, 1, 0), (3, 1, 1), (3, 1, 2), (4, 1, 2),
                                         (5, 0, 0), (5, 1, 1), (5, 1, 2)):       # 5: sub-warp kernel (orders 1-2; else block kernel)
            op.set_option("kernel", kernel)
            op.set_option("scatter", scatter)
            op.set_option("overlap", overlap)
            for rep in range(3):                                  # repeated: the overlapped schedule reuses events / buffers
                op.Mult(xd, yd)
            ctx.sync()
            err = np.linalg.norm(yd.cpu().numpy() - yg[mine[:nt]]) / np.linalg.norm(yg)
            ok &= err < 1e-12
            print(f"[rank {rank}] apply kernel={kernel} scatter={scatter} overlap={overlap} rel err {err:.2e}", flush=True)
        op.set_option("kernel", 3 if p == 3 else (5 if p < 3 else 4))     # the defaults
        op.set_option("overlap", 1)
        used = op.get_option("halo")
        print(f"[rank {rank}] shared-dof protocol in use: {used}, peer-memory all-reduce: {op.get_option('allreduce')}", flush=True)
        ok &= used == args.halo                                   # no silent fall-back on a box with NVLink peers
        if used == 2:
            # ghost consistency: local-size vectors ("tail"), the result must equal the oracle on EVERY local dof, and
            # feeding y back as x without a P exchange ("ghost_in") must reproduce A (A x)
            op.set_option("tail", 1)
            xl, yl, zl = (torch.zeros(sp.ndof, dtype=torch.float64, device="cuda") for _ in range(3))
            xl[:nt] = xd
            torch.cuda.synchronize()
            for scatter in (0, 1):
                op.set_option("scatter", scatter)
                op.Mult(xl, yl)                                   # P on x, then one symmetric exchange
                ctx.sync()
                err = np.linalg.norm(yl.cpu().numpy() - yg[mine]) / np.linalg.norm(yg)
                ok &= err < 1e-12
                op.set_option("ghost_in", 1)
                for rep in range(4):                              # repeated: double-buffered receive areas, epochs
                    op.Mult(yl, zl)
                ctx.sync()
                op.set_option("ghost_in", 0)
                zz = P.pa_op(True).mult(yg)
                err2 = np.linalg.norm(zl.cpu().numpy() - zz[mine]) / np.linalg.norm(zz)
                ok &= err2 < 1e-12
                print(f"[rank {rank}] symmetric exchange scatter={scatter}: whole local vector err {err:.2e}, chained apply without P {err2:.2e}", flush=True)
            op.set_option("scatter", 1)
            op.set_option("tail", 0)
            # RecoverFEMSolution: u_L = P x_T
            ul = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
            sp.prolongate(xd, ul)
            ctx.sync()
            ok &= np.array_equal(ul.cpu().numpy(), xg[mine])
        # Jacobi diagonal (P^T-summed) and distributed GMRES
        dd = torch.zeros(nt, dtype=torch.float64, device="cuda")
        op.AssembleDiagonal(dd)
        ctx.sync()
        err = np.linalg.norm(dd.cpu().numpy() - dg[mine[:nt]]) / np.linalg.norm(dg)
        ok &= err < 1e-12
        bd = torch.from_numpy(bg[mine[:nt]]).cuda()
        xs = torch.zeros(nt, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        s = cdm.GMRESSolver(cdm.GMRES_PETSC, 0, 500, 1e-10, 1e-12, jacobi=True)
        s.SetOperator(op)
        s.Mult(bd, xs)
        ctx.sync()
        herr = np.max(np.abs(s.history - info["hist"][:len(s.history)]) / info["hist"][0]) if len(s.history) == len(info["hist"]) else 1.0
        serr = np.linalg.norm(xs.cpu().numpy() - xs_g[mine[:nt]]) / np.linalg.norm(xs_g)
        ok &= s.GetConverged() and s.GetNumIterations() == info["iters"] and herr < 1e-10 and serr < 1e-9
        print(f"[rank {rank}] diag err {err:.2e} gmres iters {s.GetNumIterations()}/{info['iters']} hist err {herr:.2e} sol err {serr:.2e}", flush=True)
        nrm = ctx.norm2(xd)
        ok &= abs(nrm - np.linalg.norm(xg)) < 1e-12 * np.linalg.norm(xg)      # all-reduced dot over T-dofs
        # same solve with ncclAllReduce for the Krylov scalars, and with a non-zero initial guess
        op.set_option("allreduce", 0)
        xs2 = torch.zeros(nt, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        s.Mult(bd, xs2)
        ctx.sync()
        ok &= s.GetNumIterations() == info["iters"] and np.linalg.norm((xs2 - xs).cpu().numpy()) <= 1e-9 * np.linalg.norm(xs_g)
        op.set_option("allreduce", 1)
        x0g = 0.5 * xs_g
        xr_g, info0 = P.pa_op(True).gmres(bg, dinv=1 / dg, x0=x0g, variant=0, rtol=1e-10, atol=1e-12, max_it=500)
        xs2.copy_(torch.from_numpy(x0g[mine[:nt]]).cuda())
        torch.cuda.synchronize()
        s.iterative_mode = True
        s.Mult(bd, xs2)
        ctx.sync()
        s.iterative_mode = False
        h0 = np.max(np.abs(s.history - info0["hist"][:len(s.history)]) / info0["hist"][0]) if len(s.history) == len(info0["hist"]) else 1.0
        ok &= s.GetNumIterations() == info0["iters"] and h0 < 1e-10
        print(f"[rank {rank}] gmres with non-zero initial guess: iters {s.GetNumIterations()}/{info0['iters']} hist err {h0:.2e}", flush=True)
        # CG on the symmetric part (mfem::CGSolver, mesh_recession_handler.cpp:270-276)
        Pk = orc.Problem(3, p, args.mesh, perturb=0.1, kappa=1.0, vel=None, mass=None)
        xk_g, infok = Pk.pa_op(True).cg(bg, rtol=1e-12, atol=0.0, max_it=500)
        opk = cdm.ConvectionDiffusionOperator(sp, kappa=1.0, ess_dofs=ess)
        opk.set_option("halo", args.halo)
        cg = cdm.CGSolver(500, 1e-12, 0.0, jacobi=False)
        cg.SetOperator(opk)
        xk = torch.zeros(nt, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        cg.Mult(bd, xk)
        ctx.sync()
        mk = min(len(cg.history), len(infok["hist"]))
        hk = np.max(np.abs(cg.history[:mk] - infok["hist"][:mk]) / infok["hist"][0])
        kerr = np.linalg.norm(xk.cpu().numpy() - xk_g[mine[:nt]]) / np.linalg.norm(xk_g)
        ok &= cg.GetConverged() and abs(cg.GetNumIterations() - infok["iters"]) <= 1 and hk < 1e-10 and kerr < 1e-9
        print(f"[rank {rank}] cg iters {cg.GetNumIterations()}/{infok['iters']} hist err {hk:.2e} sol err {kerr:.2e}", flush=True)
        del opk
        # linear form (P^T-assembled) and L2 error (ghost values via P, all-reduced) on the partitioned space
        fn = lambda x: 1.0 + np.sin(2.3 * x[..., 0]) * np.cos(1.7 * x[..., 1]) + 0.5 * x[..., 2] ** 2
        lf_g = P.domain_lf(fn(P.rule_coords(p + 1)))
        lf = torch.zeros(nt, dtype=torch.float64, device="cuda")
        sp.domain_lf(fn(sp.rule_coords(p + 1)), lf)
        ctx.sync()
        lerr = np.linalg.norm(lf.cpu().numpy() - lf_g[mine[:nt]]) / np.linalg.norm(lf_g)
        e_g = P.l2_error(xg, fn(P.rule_coords(p + 2)))
        e_l = sp.l2_error(xd, fn(sp.rule_coords(p + 2)))
        ok &= lerr < 1e-12 and abs(e_l - e_g) < 1e-12 * e_g
        print(f"[rank {rank}] linear form err {lerr:.2e}  L2 error {e_l:.12e} vs {e_g:.12e}", flush=True)
        # backward-Euler time steps (BASELINE config 3 semantics, linear_convection_diffusion_1D.cpp:537-572) on the
        # partitioned mesh: mass apply, boundary values on the x faces, RHS elimination, GMRES -- against the oracle
        # doing the same on the un-partitioned mesh
        dt, pe = 1e-2, 10.0
        Pm = orc.Problem(3, p, args.mesh, perturb=0.1, kappa=None, vel=None, mass=1.0, ess_attrs=(3, 5))
        Pb = orc.Problem(3, p, args.mesh, perturb=0.1, kappa=dt / pe, vel=(1.0, 0.0, 0.0), alpha=dt, mass=1.0, ess_attrs=(3, 5))
        mk = np.zeros(6, np.int32); mk[[2, 4]] = 1
        ess2 = sp.essential_dofs(mk)
        assert np.all(Pb.ess_mark[mine[ess2]] == 1) and Pb.ess_mark[mine].sum() == len(ess2)
        mform = cdm.ConvectionDiffusionOperator(sp, mass=1.0)
        bform = cdm.ConvectionDiffusionOperator(sp, kappa=dt / pe, vel=(1.0, 0.0, 0.0), alpha=dt, mass=1.0, ess_dofs=ess2)
        for o in (mform, bform):
            o.set_option("halo", args.halo)
            o.set_option("tail", 1)
        bs = cdm.GMRESSolver(cdm.GMRES_PETSC, 0, 500, 1e-10, 1e-12, jacobi=True)
        bs.SetOperator(bform)
        ref_b = Pb.pa_op(True)
        db = np.where(Pb.ess_mark, 1.0, Pb.pa_diag())
        c_g = np.zeros(P.ndof)
        c_l, rhs_l, sol_l, g_l = (torch.zeros(sp.ndof, dtype=torch.float64, device="cuda") for _ in range(4))
        essd2 = torch.from_numpy(ess2).cuda()
        Xg = P.coords()
        torch.cuda.synchronize()
        for step in (1, 2, 3):
            gb = np.where(Pb.ess_mark, np.cos(3.0 * Xg[:, 1] + step) * (1.0 - Xg[:, 0]), 0.0)      # time-dependent Dirichlet data
            rhs_g = Pm.pa_apply(c_g)
            ref_b.eliminate_rhs(gb, rhs_g)
            c_new, ib = ref_b.gmres(rhs_g, dinv=1 / db, variant=0, rtol=1e-10, atol=1e-12, max_it=500)
            mform.MultUnconstrained(c_l, rhs_l)
            sp.project_dofs(essd2, gb[mine[ess2]], g_l)
            bform.EliminateRHS(g_l, rhs_l)
            bs.Mult(rhs_l, sol_l)
            ctx.sync()
            c_l.copy_(sol_l)
            torch.cuda.synchronize()
            c_g = c_new
            hb = np.max(np.abs(bs.history - ib["hist"][:len(bs.history)]) / ib["hist"][0]) if len(bs.history) == len(ib["hist"]) else 1.0
            eb = np.linalg.norm(sol_l[:nt].cpu().numpy() - c_new[mine[:nt]]) / np.linalg.norm(c_new)
            ok &= bs.GetConverged() and bs.GetNumIterations() == ib["iters"] and hb < 1e-10 and eb < 1e-9
            print(f"[rank {rank}] backward-Euler step {step}: gmres iters {bs.GetNumIterations()}/{ib['iters']} hist err {hb:.2e} sol err {eb:.2e}", flush=True)
        del mform, bform
        if args.p2p:
            # the same apply / GMRES with the shared dofs exchanged through peer memory (NVLink stores + flags)
            op.set_option("halo", 1)
            for overlap in (0, 2):
                op.set_option("overlap", overlap)
                for rep in range(6):
                    op.Mult(xd, yd)
                ctx.sync()
                err = np.linalg.norm(yd.cpu().numpy() - yg[mine[:nt]]) / np.linalg.norm(yg)
                ok &= err < 1e-12
                print(f"[rank {rank}] peer-memory halo: apply overlap={overlap} rel err {err:.2e}", flush=True)
            op.set_option("overlap", 1)
            xs.zero_()
            torch.cuda.synchronize()
            s.Mult(bd, xs)
            ctx.sync()
            herr = np.max(np.abs(s.history - info["hist"][:len(s.history)]) / info["hist"][0]) if len(s.history) == len(info["hist"]) else 1.0
            ok &= s.GetConverged() and s.GetNumIterations() == info["iters"] and herr < 1e-10
            print(f"[rank {rank}] peer-memory halo: gmres iters {s.GetNumIterations()}/{info['iters']} hist err {herr:.2e}", flush=True)
            op.set_option("halo", args.halo)
    else:
        # host emulation of cdm_halo_P / cdm_halo_PT with gloo, element work by the oracle
        lvx, lev, lbv, lba = lm.arrays()
        perm, nbdr = sp.elem_perm()                               # the space orders boundary elements first
        lev = np.ascontiguousarray(lev[perm])
        assert 0 < nbdr <= sp.ne and np.array_equal(np.sort(perm), np.arange(sp.ne))
        g, o, i = sp.maps()
        shared = np.zeros(sp.ndof, bool)
        for _, own, ghost in sp.halo():
            shared[own] = True
            shared[ghost] = True
        touches = shared[g].any(axis=1)
        assert touches[:nbdr].all() and not touches[nbdr:].any()
        L = orc.lib()
        nsym = 6
        Dd = np.zeros((sp.ne, nsym, sp.nq)); Dc = np.zeros((sp.ne, 3, sp.nq)); Dm = np.zeros((sp.ne, sp.nq))
        ck, cv, cm = np.array([kap]), np.array(vel), np.array([mass])
        L.orc_qdata(3, p, sp.ne, lev, lvx, 1, 1, orc._opt(ck), 1, orc._opt(cv), 1.0, 1, orc._opt(cm),
                    orc._opt(Dd), orc._opt(Dc), orc._opt(Dm))
        plan = sp.halo()

        def exchange(vec, forward):
            reqs, bufs = [], []
            for peer, own, ghost in plan:
                snd, rcv = (own, ghost) if forward else (ghost, own)
                if len(snd):
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec[snd])), peer))
                if len(rcv):
                    t = torch.zeros(len(rcv), dtype=torch.float64)
                    reqs.append(dist.irecv(t, peer))
                    bufs.append((rcv, t))
            for r in reqs:
                r.wait()
            for idx, t in bufs:
                if forward:
                    vec[idx] = t.numpy()
                else:
                    vec[idx] += t.numpy()

        essm = np.zeros(sp.ndof, bool)
        essm[ess] = True
        xL = np.zeros(sp.ndof)
        xL[:nt] = xg[mine[:nt]]
        exchange(xL, True)                                        # P
        assert np.array_equal(xL, xg[mine])                       # ghosts received the owners' values
        z = np.where(essm, 0.0, xL)
        yL = np.zeros(sp.ndof)
        L.orc_pa_apply(3, p, sp.ne, sp.ndof, g, o, i, orc._opt(Dd), orc._opt(Dc), orc._opt(Dm), z, yL)
        exchange(yL, False)                                       # P^T
        yL[essm] = xL[essm]
        err = np.linalg.norm(yL[:nt] - yg[mine[:nt]]) / np.linalg.norm(yg)
        ok &= err < 1e-12
        print(f"[rank {rank}] host-emulated partitioned apply rel err {err:.2e} (ntrue {nt}, ghosts {sp.ndof - nt})", flush=True)
        # the symmetric exchange (default GPU protocol): every sharer sends its partial sums to every other sharer and
        # adds all contributions in rank order -> the WHOLE local vector (ghosts included) equals the oracle's result,
        # bitwise identical on all sharers
        speers, sdof, soff, ssrc = sp.sym_plan()
        assert sorted(r for r, _, _ in speers) == [r for r, _, _ in speers] and rank not in [r for r, _, _ in speers]
        assert np.array_equal(np.unique(sdof), np.flatnonzero(shared))          # exactly the shared dofs, each once
        yS = np.zeros(sp.ndof)
        L.orc_pa_apply(3, p, sp.ne, sp.ndof, g, o, i, orc._opt(Dd), orc._opt(Dc), orc._opt(Dm), z, yS)
        total = sum(len(idx) for _, _, idx in speers)
        recv = np.zeros(total)
        reqs, bufs = [], []
        for peer, off, idx in speers:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(yS[idx])), peer))
            t = torch.zeros(len(idx), dtype=torch.float64)
            reqs.append(dist.irecv(t, peer))
            bufs.append((off, t))
        for r in reqs:
            r.wait()
        for off, t in bufs:
            recv[off:off + len(t)] = t.numpy()
        own = yS.copy()
        for k in range(len(sdof)):
            acc = 0.0
            for j in range(soff[k], soff[k + 1]):
                acc += own[sdof[k]] if ssrc[j] < 0 else recv[ssrc[j]]
            yS[sdof[k]] = acc
        yS[essm] = xL[essm]
        err = np.linalg.norm(yS - yg[mine]) / np.linalg.norm(yg)
        ok &= err < 1e-12
        # bitwise agreement between sharers: compare my shared values with each peer's copy
        reqs, bufs = [], []
        for peer, off, idx in speers:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(yS[idx])), peer))
            t = torch.zeros(len(idx), dtype=torch.float64)
            reqs.append(dist.irecv(t, peer))
            bufs.append((idx, t))
        for r in reqs:
            r.wait()
        ok &= all(np.array_equal(yS[idx], t.numpy()) for idx, t in bufs)
        print(f"[rank {rank}] host-emulated symmetric exchange rel err {err:.2e} ({len(speers)} peers, {len(sdof)} shared dofs)", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda" if gpu else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()

"""The shared-memory layouts of the high-order apply kernel (`csrc/kernels_apply_warp.cu`) are chosen with the wavefront
model of `scripts/bank_model.py`.  These CPU tests pin the two claims DESIGN.md makes about them: the lane maps used at
orders 5 and 6 reach the minimum wavefront count for their lane numbers, and the strides in the kernel source are the ones
the model was run with (in-place layout: injective, so an (qx,dz) thread only ever touches its own entries)."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("bank_model", os.path.join(ROOT, "scripts", "bank_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _kernel_source():
    with open(os.path.join(ROOT, "continuum-mechanics-mfem_b200", "csrc", "kernels_apply_warp.cu")) as fh:
        return fh.read()


def test_padded_lane_maps_reach_the_ideal_wavefront_count():
    m = _model()
    # p=5: 12 x-lines and 14 (qx,dz) pairs per half-warp, RSTR 56
    P, D, Q = 5, 6, 7
    l1 = [(16 * h + r, r % D, 2 * h + r // D) for h in range(3) for r in range(12)]
    l2 = [(16 * h + r, r % Q, 2 * h + r // Q) for h in range(3) for r in range(14)]
    assert sorted((a, b) for _, a, b in l1) == [(a, b) for a in range(D) for b in range(D)]
    assert sorted((a, b) for _, a, b in l2) == [(a, b) for a in range(Q) for b in range(D)]
    tot, _ = m.cost(P, l1, l2, lambda dy, dz: Q * dy + 56 * dz, Q, 56)
    assert tot == m.ideal(P) == 426
    dense1 = [(t, t % D, t // D) for t in range(D * D)]
    dense2 = [(t, t // D, t % D) for t in range(Q * D)]
    assert m.cost(P, dense1, dense2, lambda dy, dz: Q * dy + 58 * dz, Q, 58)[0] == 558
    # p=6: 14 x-lines per half-warp (dy pairs), dense (qx,dz) role, RSTR 71
    P, D, Q = 6, 7, 8
    l1 = [(16 * h + r, 2 * h + r // D, r % D) for h in range(4) for r in range(14) if 2 * h + r // D < D]
    l2 = [(t, t // D, t % D) for t in range(Q * D)]
    assert sorted((a, b) for _, a, b in l1) == [(a, b) for a in range(D) for b in range(D)]
    assert max(lane for lane, _, _ in l1) < 64
    tot, _ = m.cost(P, l1, l2, lambda dy, dz: Q * dy + 71 * dz, Q, 71)
    assert tot == m.ideal(P) == 600
    # p=4 keeps separate P buffers: already ideal
    P, D, Q = 4, 5, 6
    l1 = [(t, t % D, t // D) for t in range(D * D)]
    l2 = [(t, t // D, t % D) for t in range(Q * D)]
    assert m.cost(P, l1, l2, lambda dy, dz: 9 * (dy + D * dz), Q, 45)[0] == m.ideal(P) == 250


def test_kernel_source_uses_the_modelled_strides():
    src = _kernel_source()
    m = re.search(r"static constexpr int RSTR = \(P == 3\) \? 28 : \(P == 4 \? 45 : \(P == 6 \? 71 : \(INPLACE \? \(L2MAP \? 56 : 58\)", src)
    assert m, "RSTR table of GroupCfg changed: re-run scripts/bank_model.py and update this test"
    assert "static constexpr int PST = (P == 4) ? 9 : (P == 6 ? 17 : Q);" in src
    assert "PSY = INPLACE ? Q : PST, PSZ = INPLACE ? RSTR : PST * D" in src


def test_in_place_layout_is_injective():
    """P(q; dy,dz) = q + Q dy + RSTR dz and R(qx,qy,dz) = qx + Q qy + RSTR dz share one array: distinct (qx, j, dz) must be
    distinct addresses, otherwise the in-place y contraction of one thread would overwrite another thread's inputs."""
    for p, rstr in ((5, 56), (5, 58), (6, 71)):
        q = p + 2
        d = p + 1
        seen = set()
        for qx in range(q):
            for j in range(q):
                for dz in range(d):
                    a = qx + q * j + rstr * dz
                    assert a not in seen
                    seen.add(a)
        # buffer size the kernel allocates per array: RS = ((D-1) RSTR + Q^2 + 1) & ~1
        assert max(seen) < (((d - 1) * rstr + q * q + 1) & ~1)

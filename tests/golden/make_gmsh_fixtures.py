#!/usr/bin/env python
"""Writes the Gmsh 2.2 ASCII fixtures tests use for the triangle path (the GPU box has no /root/reference, so the
reference's own Mesh/*.msh files cannot be read there; these files follow the same conventions: physical curves
bottom=1, right=2, top=3, left=4 and a physical surface, Mesh/unit_square.geo:18-22; a single boundary attribute for the
disk, Mesh/unit_circle.geo).  Delaunay triangulations of seeded point sets:
    python tests/golden/make_gmsh_fixtures.py   ->  tests/golden/square_tri.msh, tests/golden/disk_tri.msh
"""
import os

import numpy as np
from scipy.spatial import Delaunay

HERE = os.path.dirname(os.path.abspath(__file__))


def write_msh(path, pts, tris, lines, names):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % len(names))
        for dim, tag, name in names:
            f.write('%d %d "%s"\n' % (dim, tag, name))
        f.write("$EndPhysicalNames\n$Nodes\n%d\n" % len(pts))
        for i, (x, y) in enumerate(pts):
            f.write("%d %.16g %.16g 0\n" % (i + 1, x, y))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(lines) + len(tris)))
        k = 1
        for a, b, tag in lines:
            f.write("%d 1 2 %d %d %d %d\n" % (k, tag, tag, a + 1, b + 1)); k += 1
        for a, b, c in tris:
            f.write("%d 2 2 1 1 %d %d %d\n" % (k, a + 1, b + 1, c + 1)); k += 1
        f.write("$EndElements\n")


def square(n=9, seed=1):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1, n + 1)
    X, Y = np.meshgrid(t, t, indexing="xy")
    pts = np.stack([X.ravel(), Y.ravel()], 1)
    interior = (pts[:, 0] > 0) & (pts[:, 0] < 1) & (pts[:, 1] > 0) & (pts[:, 1] < 1)
    pts[interior] += rng.uniform(-0.3, 0.3, (interior.sum(), 2)) / n
    tri = Delaunay(pts).simplices
    lines = []
    idx = lambda i, j: i + (n + 1) * j
    for i in range(n):
        lines.append((idx(i, 0), idx(i + 1, 0), 1))          # bottom
        lines.append((idx(n, i), idx(n, i + 1), 2))          # right
        lines.append((idx(i + 1, n), idx(i, n), 3))          # top
        lines.append((idx(0, i + 1), idx(0, i), 4))          # left
    return pts, tri, lines, [(1, 1, "bottom"), (1, 2, "right"), (1, 3, "top"), (1, 4, "left"), (2, 1, "domain")]


def disk(nr=6, seed=2):
    rng = np.random.default_rng(seed)
    pts = [(0.0, 0.0)]
    for k in range(1, nr + 1):
        m = 6 * k
        r = k / nr
        th = 2 * np.pi * (np.arange(m) + (0.0 if k == nr else rng.uniform(-0.15, 0.15, m))) / m
        pts += list(zip(r * np.cos(th), r * np.sin(th)))
    pts = np.array(pts)
    tri = Delaunay(pts).simplices
    nb = 6 * nr
    first = len(pts) - nb
    lines = [(first + i, first + (i + 1) % nb, 1) for i in range(nb)]
    return pts, tri, lines, [(1, 1, "boundary"), (2, 1, "domain")]


if __name__ == "__main__":
    write_msh(os.path.join(HERE, "square_tri.msh"), *square())
    write_msh(os.path.join(HERE, "disk_tri.msh"), *disk())
    print("written")

"""CPU ORACLE for Mesh::UniformRefinement (test infrastructure, numpy only; PARITY UNPINNED like the rest of oracle/: MFEM is
not vendored).  Restates, independently of the C++ host code (dictionary-based entity tables), what the reference's drivers do
with serial_ref_levels / par_ref_levels (linear_convection_diffusion_2D.cpp:295-304; Input/input_diffusion_mms.yaml refines
Mesh/unit_square.msh once).  Conventions of MFEM's UniformRefinement2D_base / UniformRefinement3D_base [MFEM-upstream, from
memory]."""
import numpy as np

HEX_E = ((0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7))
HEX_F = ((3, 2, 1, 0), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7))


def uniform_refine_2d(vx, ev, bv, battr):
    """Mesh::UniformRefinement() of a conforming 2D triangle or quadrilateral mesh, restated independently of the C++ host
    code (dictionary-based edge table, numpy arrays) -- linear_convection_diffusion_2D.cpp:295-298; conventions of MFEM's
    UniformRefinement2D_base [MFEM-upstream, from memory]: edges numbered in first-encounter order over the elements, new
    vertices = old, edge midpoints, (quadrilateral centres); children of element i listed consecutively,
    triangle: (v0,e0,e2) (e1,e2,e0) (e0,v1,e1) (e2,e1,v2); quad: (v0,e0,c,e3) (e0,v1,e1,c) (c,e1,v2,e2) (e3,c,e2,v3);
    boundary segment (a,b) -> (a,m), (m,b)."""
    vx, ev, bv, battr = np.asarray(vx, float), np.asarray(ev), np.asarray(bv), np.asarray(battr)
    nv, (ne, k) = len(vx), ev.shape
    edges = {}
    el_edge = np.zeros((ne, k), np.int64)
    for e in range(ne):
        for j in range(k):
            a, b = int(ev[e, j]), int(ev[e, (j + 1) % k])
            el_edge[e, j] = edges.setdefault((min(a, b), max(a, b)), len(edges))
    nedges = len(edges)
    out_v = np.zeros((nv + nedges + (ne if k == 4 else 0), 2))
    out_v[:nv] = vx
    out_e = np.zeros((4 * ne, k), np.int32)
    for e in range(ne):
        v = [int(t) for t in ev[e]]
        m = [nv + int(el_edge[e, j]) for j in range(k)]
        for j in range(k):
            out_v[m[j]] = (0.0 + vx[v[j]] + vx[v[(j + 1) % k]]) * 0.5
        if k == 3:
            out_e[4 * e:4 * e + 4] = [[v[0], m[0], m[2]], [m[1], m[2], m[0]], [m[0], v[1], m[1]], [m[2], m[1], v[2]]]
        else:
            c = nv + nedges + e
            out_v[c] = (((0.0 + vx[v[0]]) + vx[v[1]]) + vx[v[2]] + vx[v[3]]) * 0.25
            out_e[4 * e:4 * e + 4] = [[v[0], m[0], c, m[3]], [m[0], v[1], m[1], c], [c, m[1], v[2], m[2]], [m[3], c, m[2], v[3]]]
    out_b = np.zeros((2 * len(bv), 2), np.int32)
    out_a = np.repeat(battr, 2).astype(np.int32)
    for b in range(len(bv)):
        a0, a1 = int(bv[b, 0]), int(bv[b, 1])
        mid = nv + edges[(min(a0, a1), max(a0, a1))]
        out_b[2 * b], out_b[2 * b + 1] = (a0, mid), (mid, a1)
    return out_v, out_e, out_b, out_a


def uniform_refine_3d(vx, ev, bv, battr):
    """hexahedral meshes: edges and faces numbered in first-encounter order over the elements (all edges first, then all faces),
    new vertices = old, edge midpoints, face centres, element centres; child k at corner k of its parent, boundary quadrilaterals
    split into four.  Averages are recomputed by every element that visits an entity (Mesh::AverageVertices): the last visitor's
    summation order stays."""
    vx, ev, bv, battr = np.asarray(vx, float), np.asarray(ev), np.asarray(bv), np.asarray(battr)
    nv, ne = len(vx), len(ev)
    edges, faces = {}, {}
    el_edge = np.zeros((ne, 12), np.int64)
    el_face = np.zeros((ne, 6), np.int64)
    for e in range(ne):
        for k, (a, b) in enumerate(HEX_E):
            va, vb = int(ev[e, a]), int(ev[e, b])
            el_edge[e, k] = edges.setdefault((min(va, vb), max(va, vb)), len(edges))
    for e in range(ne):
        for k, f in enumerate(HEX_F):
            el_face[e, k] = faces.setdefault(tuple(sorted(int(ev[e, i]) for i in f)), len(faces))
    oedge, oface = nv, nv + len(edges)
    oelem = oface + len(faces)
    out_v = np.zeros((oelem + ne, 3))
    out_v[:nv] = vx

    def avg(idx):
        s = np.zeros(3)
        for i in idx:
            s = s + out_v[i]
        return s * (1.0 / len(idx))

    out_e = np.zeros((8 * ne, 8), np.int32)
    for i in range(ne):
        v = [int(t) for t in ev[i]]
        c = oelem + i
        out_v[c] = avg(v)
        f = [oface + int(el_face[i, k]) for k in range(6)]
        for k in range(6):
            out_v[f[k]] = avg([v[j] for j in HEX_F[k]])
        e = [oedge + int(el_edge[i, k]) for k in range(12)]
        for k in range(12):
            out_v[e[k]] = avg([v[HEX_E[k][0]], v[HEX_E[k][1]]])
        out_e[8 * i:8 * i + 8] = [
            [v[0], e[0], f[0], e[3], e[8], f[1], c, f[4]],
            [e[0], v[1], e[1], f[0], f[1], e[9], f[2], c],
            [f[0], e[1], v[2], e[2], c, f[2], e[10], f[3]],
            [e[3], f[0], e[2], v[3], f[4], c, f[3], e[11]],
            [e[8], f[1], c, f[4], v[4], e[4], f[5], e[7]],
            [f[1], e[9], f[2], c, e[4], v[5], e[5], f[5]],
            [c, f[2], e[10], f[3], f[5], e[5], v[6], e[6]],
            [f[4], c, f[3], e[11], e[7], f[5], e[6], v[7]]]
    out_b = np.zeros((4 * len(bv), 4), np.int32)
    out_a = np.repeat(battr, 4).astype(np.int32)
    for b in range(len(bv)):
        v = [int(t) for t in bv[b]]
        e = [oedge + edges[(min(v[k], v[(k + 1) % 4]), max(v[k], v[(k + 1) % 4]))] for k in range(4)]
        q = oface + faces[tuple(sorted(v))]
        out_b[4 * b:4 * b + 4] = [[v[0], e[0], q, e[3]], [e[0], v[1], e[1], q], [q, e[1], v[2], e[2]], [e[3], q, e[2], v[3]]]
    return out_v, out_e, out_b, out_a

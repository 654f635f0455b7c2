"""Independent golden vectors for the Kantorovich evaluation — generated WITHOUT the oracle and without the engine.

The reference ships no golden vectors, and oracle/ is this repo's own restatement of it.  These
fixtures pin both the oracle and the CUDA engine against a third, deliberately different computation:

  * exact rational arithmetic (fractions.Fraction of the binary64 inputs) for every geometric decision and every
    coordinate: the Laguerre cell of site i is the mesh box clipped by the bisector with EVERY other site (no neighbour
    search at all), each cell is clipped by every triangle of the mesh (no traversal), Sutherland–Hodgman with exact
    signs, ties outside (predicates.hpp:85-86);
  * integrals by rules other than the reference's: the piece is fanned from its first vertex and each fan triangle
    integrated with the vertex / mid-edge / centroid rule (weights 1/20, 2/15, 9/20), exact for cubics — the reference
    uses the centroid rule and Albrecht–Collatz (quadrature.hpp:24-42,69-77);
  * the Hessian entry H_ij = -(integral of rho over the common edge) / (2 |y_i - y_j|) (kantorovich.hpp:117-121) from the exact
    end points of the edge pieces; only the final square roots are taken in floating point;
  * SciPy/Qhull's lower convex hull of the lifted sites as a cross-check of the adjacency (every Hessian edge must be
    an edge of the regular triangulation).

    python tests/golden/make_independent.py        # rewrites tests/golden/indep_*.npz (a few seconds each)
"""
import math
import os
from fractions import Fraction as Fr

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def F(v):
    return Fr(float(v))


def clip(poly, tags, a, b, c, newtag):
    """Keep { a x + b y + c > 0 } (strict).  poly: list of exact points, tags[k] = label of the edge leaving vertex k."""
    n = len(poly)
    out, ot = [], []
    for k in range(n):
        p0, p1 = poly[k], poly[(k + 1) % n]
        s0 = a * p0[0] + b * p0[1] + c
        s1 = a * p1[0] + b * p1[1] + c
        if s0 > 0:
            out.append(p0)
            if s1 > 0:
                ot.append(tags[k])
            else:
                t = s0 / (s0 - s1)
                ot.append(tags[k])
                out.append((p0[0] + t * (p1[0] - p0[0]), p0[1] + t * (p1[1] - p0[1])))
                ot.append(newtag)
        elif s1 > 0:
            t = s0 / (s0 - s1)
            out.append((p0[0] + t * (p1[0] - p0[0]), p0[1] + t * (p1[1] - p0[1])))
            ot.append(tags[k])
    return out, ot


def tri_cubic(q, p0, p1, p2):
    """Exact integral of a polynomial q of degree <= 3 over the triangle (signed area)."""
    A = ((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1])) / 2
    mid = lambda u, v: ((u[0] + v[0]) / 2, (u[1] + v[1]) / 2)
    cen = ((p0[0] + p1[0] + p2[0]) / 3, (p0[1] + p1[1] + p2[1]) / 3)
    s = Fr(1, 20) * (q(p0) + q(p1) + q(p2)) + Fr(2, 15) * (q(mid(p0, p1)) + q(mid(p1, p2)) + q(mid(p2, p0))) + Fr(9, 20) * q(cen)
    return A * s


def evaluate(vx, vy, tri, rho, X, w):
    """-> f, g (N), H (N x N dense), pieces count; everything exact up to the final float conversions."""
    N = len(X)
    P = [(F(X[i, 0]), F(X[i, 1])) for i in range(N)]
    W = [F(w[i]) for i in range(N)]
    V = [(F(vx[k]), F(vy[k])) for k in range(len(vx))]
    R = [F(r) for r in rho]
    x0, x1 = min(v[0] for v in V), max(v[0] for v in V)
    y0, y1 = min(v[1] for v in V), max(v[1] for v in V)
    g = np.zeros(N)
    H = np.zeros((N, N))
    f = Fr(0)
    npieces = 0
    for i in range(N):
        xi, yi = P[i]
        poly = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
        tags = [-1, -1, -1, -1]
        for j in range(N):
            if j == i or not poly:
                continue
            xj, yj = P[j]
            if (xj, yj) == (xi, yi):  # coincident sites: the heavier, then the earlier one keeps the cell
                if W[j] > W[i] or (W[j] == W[i] and j < i):
                    poly = []
                continue
            # pow_i(x) < pow_j(x)  <=>  2 x.(y_i - y_j) - |y_i|^2 + |y_j|^2 + w_i - w_j > 0
            a, b = 2 * (xi - xj), 2 * (yi - yj)
            c = -(xi * xi + yi * yi) + (xj * xj + yj * yj) + W[i] - W[j]
            poly, tags = clip(poly, tags, a, b, c, j)
        if len(poly) < 3:
            continue
        mass = Fr(0)
        cost = Fr(0)
        for t in range(len(tri)):
            ia, ib, ic = (int(v) for v in tri[t])
            T = (V[ia], V[ib], V[ic])
            pc, pt = poly, tags
            for e in range(3):
                p, q = T[e], T[(e + 1) % 3]
                aa, bb = -(q[1] - p[1]), (q[0] - p[0])  # left of p -> q
                pc, pt = clip(pc, pt, aa, bb, -(aa * p[0] + bb * p[1]), -2)
                if len(pc) < 3:
                    break
            if len(pc) < 3:
                continue
            npieces += 1
            (ax, ay), (bx, by), (cx, cy) = T
            det = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay)
            A_ = ((R[ib] - R[ia]) * (cy - ay) - (R[ic] - R[ia]) * (by - ay)) / det
            B_ = ((bx - ax) * (R[ic] - R[ia]) - (cx - ax) * (R[ib] - R[ia])) / det
            C_ = R[ia] - A_ * ax - B_ * ay
            dens = lambda p: A_ * p[0] + B_ * p[1] + C_
            qcost = lambda p: dens(p) * ((p[0] - xi) ** 2 + (p[1] - yi) ** 2)
            for k in range(1, len(pc) - 1):
                mass += tri_cubic(dens, pc[0], pc[k], pc[k + 1])
                cost += tri_cubic(qcost, pc[0], pc[k], pc[k + 1])
            for k in range(len(pc)):
                j = pt[k]
                if j is None or j < 0:
                    continue
                p0, p1 = pc[k], pc[(k + 1) % len(pc)]
                mid = ((p0[0] + p1[0]) / 2, (p0[1] + p1[1]) / 2)
                L = math.sqrt(float((p1[0] - p0[0]) ** 2 + (p1[1] - p0[1]) ** 2))
                d = 2 * math.sqrt(float((xi - P[j][0]) ** 2 + (yi - P[j][1]) ** 2))
                r = L * float(dens(mid)) / d
                H[i, j] -= r
                H[i, i] += r
        g[i] = float(mass)
        f += mass * W[i] - cost
    return float(f), g, H, npieces


def qhull_edges(X, w):
    from scipy.spatial import ConvexHull
    lift = np.c_[X, (X ** 2).sum(1) - w]
    hull = ConvexHull(lift)
    low = hull.simplices[hull.equations[:, 2] < -1e-12]
    E = set()
    for a, b, c in low:
        E |= {(min(a, b), max(a, b)), (min(b, c), max(b, c)), (min(a, c), max(a, c))}
    return E


def grid_mesh(n, m, diag=None):
    xs = np.linspace(-1, 1, n)
    ys = np.linspace(-1, 1, m)
    vx = np.repeat(xs, m)
    vy = np.tile(ys, n)
    tri = []
    for i in range(n - 1):
        for j in range(m - 1):
            v00, v10, v11, v01 = i * m + j, (i + 1) * m + j, (i + 1) * m + j + 1, i * m + j + 1
            if diag is not None and diag[i, j]:
                tri += [[v00, v10, v01], [v10, v11, v01]]
            else:
                tri += [[v00, v10, v11], [v00, v11, v01]]
    return vx, vy, np.array(tri, np.int32)


CASES = {
    # name: (mesh, N, seed, weight scale in cell areas)
    "indep_square_n40": ("square", 40, 11, 0.3),
    "indep_grid5x4_n60": ("grid5x4", 60, 12, 0.4),
    "indep_grid9_n30_w0": ("grid9", 30, 13, 0.0),
    # the squares split along either diagonal at random (what a Delaunay triangulation of the pixel grid gives):
    # pins the per-square-diagonal form of the boundary-integration kernel (k_seg<VD>) and ma_set_mesh's grid recognition
    "indep_grid6x5_altdiag_n50": ("grid6x5alt", 50, 14, 0.3),
}


def build(name):
    kind, N, seed, ws = CASES[name]
    rng = np.random.default_rng(seed)
    if kind == "square":
        vx = np.array([0.0, 1.0, 1.0, 0.0]); vy = np.array([0.0, 0.0, 1.0, 1.0])
        tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
        rho = np.array([1.0, 2.5, 0.5, 1.5])
        X = rng.uniform(0.02, 0.98, (N, 2))
        n = m = 0
        area = 1.0
    else:
        n, m = {"grid5x4": (5, 4), "grid9": (9, 9), "grid6x5alt": (6, 5)}[kind]
        diag = rng.integers(0, 2, (n - 1, m - 1)) if kind.endswith("alt") else None
        vx, vy, tri = grid_mesh(n, m, diag)
        rho = 0.2 + rng.random(n * m) + np.exp(-((vx - 0.3) ** 2 + (vy + 0.2) ** 2) / 0.1)
        X = rng.uniform(-0.97, 0.97, (N, 2))
        area = 4.0
    w = rng.normal(0.0, ws * area / N, N) if ws else np.zeros(N)
    f, g, H, npieces = evaluate(vx, vy, tri, rho, X, w)
    E = qhull_edges(X, w)
    iu, ju = np.nonzero(np.triu(H, 1))
    assert {(int(a), int(b)) for a, b in zip(iu, ju)} <= E, "a Hessian edge that is not in the regular triangulation"
    # mass conservation against the exact total mass of the PL density
    tot = 0.0
    for t in tri:
        a, b, c = (int(v) for v in t)
        ar = ((vx[b] - vx[a]) * (vy[c] - vy[a]) - (vx[c] - vx[a]) * (vy[b] - vy[a])) / 2
        tot += ar * (rho[a] + rho[b] + rho[c]) / 3
    assert abs(g.sum() - tot) < 1e-13 * tot
    extra = dict(diag=diag) if kind != "square" and diag is not None else {}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, n=n, m=m, vx=vx, vy=vy, tri=tri, rho=rho, X=X, w=w,
                        f=f, g=g, H=H, npieces=npieces, **extra)
    print(name, "N", N, "pieces", npieces, "nnz", int((H != 0).sum()), "f", f)


if __name__ == "__main__":
    for name in CASES:
        build(name)

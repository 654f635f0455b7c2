"""Pins the CPU oracle (oracle/ma_oracle.cpp): closed-form known answers, the invariants the reference's
drivers print, an independent Qhull adjacency cross-check, and the committed golden
fixtures.  The reference ships no golden vectors, so these are the pins ("parity unpinned" otherwise)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from mongeampere_b200 import inputs
from tests import common

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def unit_square_oracle(O, rho=None):
    vx, vy, tri = inputs.unit_square_mesh()
    rho = np.ones(4) if rho is None else rho
    abc = O.linear_functions(vx, vy, rho, tri)
    return O.Oracle(vx, vy, tri, abc)


def test_linear_function_matches_plane_fit(oracle_mod):
    rng = np.random.default_rng(1)
    vx, vy = inputs.grid_vertices(5, 4)
    tri = inputs.grid_triangles(5, 4)
    rho = rng.random(20) + 0.1
    a1 = oracle_mod.linear_functions(vx, vy, rho, tri)
    a2 = inputs.pl_coefficients(vx, vy, rho, tri)
    assert np.abs(a1 - a2).max() < 1e-12
    # interpolation property at the three vertices
    for k in range(3):
        v = tri[:, k]
        assert np.abs(a1[:, 0] * vx[v] + a1[:, 1] * vy[v] + a1[:, 2] - rho[v]).max() < 1e-12


def test_single_dirac_closed_form(oracle_mod):
    orc = unit_square_oracle(oracle_mod)
    a, b = 0.3, 0.6
    orc.set_points(np.array([[a, b]]))
    f, g, H = orc.kantorovich(np.array([0.25]))
    cost = ((1 - a) ** 3 + a ** 3) / 3 + ((1 - b) ** 3 + b ** 3) / 3
    assert abs(g[0] - 1.0) < 1e-15
    assert abs(f - (0.25 * 1.0 - cost)) < 1e-15
    assert H.nnz == 0


def test_two_diracs_closed_form(oracle_mod):
    orc = unit_square_oracle(oracle_mod)
    orc.set_points(np.array([[0.25, 0.5], [0.75, 0.5]]))
    f, g, H = orc.kantorovich(np.zeros(2))
    assert np.allclose(g, [0.5, 0.5], atol=1e-15)
    assert np.allclose(H.toarray(), [[1, -1], [-1, 1]], atol=1e-15)  # H01 = -len/(2 d) = -1/(2*0.5)
    assert abs(f + 2 * (0.5 * 0.25 ** 2 / 3 + 0.5 / 12)) < 1e-15
    # weighted: bisector moves to x = 0.5 + (w0 - w1)/(2 d), d = 0.5
    w = np.array([0.05, -0.03])
    f, g, H = orc.kantorovich(w)
    xb = 0.5 + (w[0] - w[1]) / (2 * 0.5)
    assert np.allclose(g, [xb, 1 - xb], atol=1e-15)
    assert np.allclose(H.toarray(), [[1, -1], [-1, 1]], atol=1e-14)


def test_two_diracs_linear_density(oracle_mod):
    # rho(x, y) = 1 + x on the unit square; cell 0 = [0, .5] x [0, 1]
    vx, vy, tri = inputs.unit_square_mesh()
    orc = unit_square_oracle(oracle_mod, rho=1 + vx)
    orc.set_points(np.array([[0.25, 0.5], [0.75, 0.5]]))
    f, g, H = orc.kantorovich(np.zeros(2))
    assert np.allclose(g, [0.5 + 0.125, 0.5 + 0.375], atol=1e-15)
    assert abs(H[0, 1] + 1.5 / (2 * 0.5)) < 1e-15  # ∫_edge rho = 1 * (1 + 0.5)


def test_lattice_is_five_point_laplacian(oracle_mod):
    n = 6
    orc = unit_square_oracle(oracle_mod)
    c = (np.arange(n) + 0.5) / n
    X = np.stack(np.meshgrid(c, c, indexing="ij"), -1).reshape(-1, 2)
    orc.set_points(X)
    f, g, H = orc.kantorovich(np.zeros(n * n))
    assert np.allclose(g, 1.0 / n ** 2, atol=1e-15)
    Hd = H.toarray()
    for i in range(n):
        for j in range(n):
            k = i * n + j
            nb = [(i + 1, j), (i - 1, j), (i, j + 1), (i, j - 1)]
            nb = [a * n + b for a, b in nb if 0 <= a < n and 0 <= b < n]
            assert abs(Hd[k, k] - 0.5 * len(nb)) < 1e-13
            for q in nb:
                assert abs(Hd[k, q] + 0.5) < 1e-13
            others = np.setdiff1d(np.arange(n * n), nb + [k])
            assert np.abs(Hd[k, others]).max() < 1e-13  # diagonal (degenerate) neighbours carry zero


@pytest.mark.parametrize("name,scale,weights", [("c1", 0.05, "zero"), ("c1", 0.05, "0.4"), ("c2", 0.005, "0.4"),
                                                ("c1r", 0.03, "0.2")])
def test_invariants(oracle_mod, name, scale, weights):
    case = common.make_case(name, scale, weights)
    orc = common.oracle_for(oracle_mod, case)
    f, g, H = orc.kantorovich(case["w"])
    cfg = case["cfg"]
    tm = inputs.total_mass(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"])
    assert abs(g.sum() - tm) < 1e-12 * tm                      # mass conservation (test_quantization.cpp:78-79)
    d = np.abs(H.diagonal()).max()
    assert np.abs(np.asarray(H.sum(axis=1))).max() < 1e-11 * d   # Laplacian rows sum to zero
    assert abs(H - H.T).max() < 1e-8 * d                         # symmetric up to rounding
    # area conservation (test_voronoi_tri.cpp:67) from the recorded pieces
    orc.kantorovich(case["w"], mode=oracle_mod.MODE_RECORD | orc.default_mode())
    cell, face, ptr, tag, xy = orc.pieces()
    area = 0.0
    for p in range(len(cell)):
        P = xy[ptr[p]:ptr[p + 1]]
        area += 0.5 * np.sum(P[:, 0] * np.roll(P[:, 1], -1) - np.roll(P[:, 0], -1) * P[:, 1])
    ext = (cfg["vx"].max() - cfg["vx"].min()) * (cfg["vy"].max() - cfg["vy"].min())
    assert abs(area - ext) < 1e-12 * ext


def test_finite_differences(oracle_mod):
    """tests/test_opttransport_eval.cpp:20-55: g = df/dw and H = dg/dw by central differences (N = 50)."""
    case = common.make_case("c2", 0.0005, "0.3")
    assert case["N"] == 50
    orc = common.oracle_for(oracle_mod, case)
    w = case["w"]
    f, g, H = orc.kantorovich(w)
    eps = 1e-6
    gfd = np.zeros(50)
    Hfd = np.zeros((50, 50))
    for i in range(50):
        e = np.zeros(50); e[i] = eps
        fp, gp, _ = orc.kantorovich(w + e)
        fm, gm, _ = orc.kantorovich(w - e)
        gfd[i] = (fp - fm) / (2 * eps)
        Hfd[:, i] = (gp - gm) / (2 * eps)
    assert np.abs(gfd - g).max() < 1e-7 * np.abs(g).max()
    assert np.abs(Hfd - H.toarray()).max() < 1e-5 * np.abs(H.diagonal()).max()


def test_modes_agree(oracle_mod):
    """brute-force vs quadtree neighbour search, reference BFS vs per-cell OpenMP enumeration."""
    case = common.make_case("c2", 0.01, "0.5")
    orc = common.oracle_for(oracle_mod, case, nthreads=2)
    ref = orc.kantorovich(case["w"], mode=1)
    cref = orc.counters()
    for mode in (0, 2, 3):
        f, g, H = orc.kantorovich(case["w"], mode=mode)
        assert abs(f - ref[0]) < 1e-13 * abs(ref[0])
        assert np.abs(g - ref[1]).max() < 1e-13 * ref[1].max()
        assert common.same_pattern(ref[2], H)
        c = orc.counters()
        for k in ("pieces", "piece_vertices", "laguerre_edges", "sum_k"):
            assert c[k] == cref[k]


def test_adjacency_against_qhull(oracle_mod):
    """Independent check of the power-diagram neighbour sets: lower hull of the lifted points
    (x, y, x²+y²-w) computed by Qhull."""
    from scipy.spatial import ConvexHull
    case = common.make_case("c1", 0.1, "0.3")
    X, w, N = case["X"], case["w"], case["N"]
    orc = common.oracle_for(oracle_mod, case)
    f, g, H = orc.kantorovich(w, mode=0)
    lifted = np.c_[X, (X ** 2).sum(1) - w]
    hull = ConvexHull(lifted)
    lower = hull.simplices[hull.equations[:, 2] < -1e-12]
    E = set()
    for s in lower:
        for a, b in ((0, 1), (1, 2), (0, 2)):
            E.add((min(s[a], s[b]), max(s[a], s[b])))
    C = sp.coo_matrix(H)
    P = {(min(i, j), max(i, j)) for i, j in zip(C.row, C.col) if i != j}
    assert P <= E, "oracle found a Laguerre edge that is not in the regular triangulation"
    # RT edges missing from the Hessian pattern must be dual to edges outside the domain, i.e. join
    # sites close to the boundary
    h = 3.0 / np.sqrt(N)
    for i, j in E - P:
        if g[i] == 0 or g[j] == 0:
            continue
        di = min(X[i, 0], 1 - X[i, 0], X[i, 1], 1 - X[i, 1])
        dj = min(X[j, 0], 1 - X[j, 0], X[j, 1], 1 - X[j, 1])
        assert min(di, dj) < h, (i, j)


def test_hidden_dirac_has_zero_mass_and_row(oracle_mod):
    """SURVEY App. B T2 / §7.3-6."""
    orc = unit_square_oracle(oracle_mod)
    X = np.array([[0.3, 0.5], [0.7, 0.5], [0.5, 0.5]])
    orc.set_points(X)
    f, g, H = orc.kantorovich(np.array([0.0, 0.0, -1.0]))
    assert g[2] == 0 and abs(g.sum() - 1) < 1e-15
    assert H.getrow(2).nnz == 0


def test_moments_closed_form(oracle_mod):
    orc = unit_square_oracle(oracle_mod)
    orc.set_points(np.array([[0.25, 0.5], [0.75, 0.5]]))
    mom = orc.moments(np.zeros(2), order=2)
    # cell 0 = [0,.5]x[0,1], uniform density
    assert np.allclose(mom[0], [0.5, 0.5 * 0.25, 0.5 * 0.5, 0.5 ** 3 / 3, 0.5 / 3, 0.125 * 0.5], atol=1e-15)
    m, c = orc.lloyd(np.zeros(2))
    assert np.allclose(c, [[0.25, 0.5], [0.75, 0.5]], atol=1e-15)


def test_newton_converges_like_reference(oracle_mod):
    """optimal_transport.hpp:150: ||m - nu||_2 < eps_g at exit; direct and CG solves agree."""
    case = common.make_case("c2", 0.003, "zero")
    orc = common.oracle_for(oracle_mod, case)
    cfg = case["cfg"]
    tm = inputs.total_mass(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"])
    nu = np.full(case["N"], tm / case["N"])
    x, st, tr = oracle_mod.ot_solve(orc, nu, eps_g=1e-9)
    assert st["status"] == "ok"
    f, g, H = orc.kantorovich(x)
    assert np.linalg.norm(g - nu) < 1e-9
    x2, st2, _ = oracle_mod.ot_solve(orc, nu, eps_g=1e-9, direct=False)
    assert st2["niter"] == st["niter"] and st2["neval"] == st["neval"]
    assert np.abs((x - x[-1]) - (x2 - x2[-1])).max() < 1e-8


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("indep_")) if os.path.isdir(GOLDEN) else [])
def test_golden_fixtures(oracle_mod, fname):
    z = np.load(os.path.join(GOLDEN, fname))
    orc = oracle_mod.Oracle(z["vx"], z["vy"], z["tri"], z["abc"])
    orc.set_points(z["X"])
    f, g, H = orc.kantorovich(z["w"])
    assert abs(f - float(z["f"])) <= 1e-13 * abs(float(z["f"]))
    assert np.abs(g - z["g"]).max() <= 1e-13 * np.abs(z["g"]).max()
    H0 = sp.csr_matrix((z["H_data"], z["H_indices"], z["H_indptr"]), shape=H.shape)
    assert common.same_pattern(H0, H)
    assert abs(H - H0).max() <= 1e-12 * np.abs(H0.diagonal()).max()

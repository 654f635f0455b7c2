"""CPU emulation of the CUDA kernels' per-thread / per-lane code (tests/emu, same __host__ __device__
functions the GPU runs) against the oracle.  Covers the geometry logic of K2/K3 without a GPU."""
import numpy as np
import pytest

from mongeampere_b200 import inputs
from tests import common

CASES = [("c1", 0.05, "zero"), ("c1", 0.05, "0.4"), ("c1r", 0.03, "0.2"), ("c2", 0.004, "zero"), ("c2", 0.004, "0.5"),
         ("c3", 0.0005, "0.3"), ("c5", 0.0001, "0.2")]


def check(r, ref, orc_counters, tol=1e-10):
    f0, g0, H0 = ref
    assert r["flags"] == 0
    assert abs(r["f"] - f0) <= tol * abs(f0)
    assert np.abs(r["g"] - g0).max() <= tol * np.abs(g0).max()
    assert common.same_pattern(H0, r["H"])
    assert abs(H0 - r["H"]).max() <= tol * np.abs(H0.diagonal()).max()
    for k in ("pieces", "piece_vertices", "new_vertices", "laguerre_edges", "sum_k", "sum_k_np"):
        assert orc_counters[k] == r["counters"][k], k


@pytest.mark.parametrize("name,scale,weights", CASES)
def test_emulated_kernels_match_oracle(oracle_mod, emu_mod, name, scale, weights):
    case = common.make_case(name, scale, weights)
    orc = common.oracle_for(oracle_mod, case)
    ref = orc.kantorovich(case["w"])
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"])
    check(r, ref, orc.counters())


def test_general_mesh_enumeration(oracle_mod, emu_mod):
    case = common.make_case("c2", 0.004, "0.5")
    cfg = case["cfg"]
    orc = common.oracle_for(oracle_mod, case)
    ref = orc.kantorovich(case["w"])
    mesh = dict(kind="mesh", vx=cfg["vx"], vy=cfg["vy"], tri=cfg["tri"], abc=case["abc"])
    check(emu_mod.evaluate(mesh, case["X"], case["w"]), ref, orc.counters())


def test_double_double_fallback_agrees(oracle_mod, emu_mod):
    """Force EVERY sign through the double-double predicates (filter_tol = inf)."""
    case = common.make_case("c2", 0.004, "0.5")
    orc = common.oracle_for(oracle_mod, case)
    ref = orc.kantorovich(case["w"])
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], filter_tol=1e300)
    assert r["counters"]["fallbacks"] > 1000
    check(r, ref, orc.counters())


@pytest.mark.parametrize("nlanes,bin_target", [(1, 1), (8, 4), (32, 16)])
def test_lane_count_and_bin_size_do_not_matter(oracle_mod, emu_mod, nlanes, bin_target):
    case = common.make_case("c3", 0.0005, "0.3")
    orc = common.oracle_for(oracle_mod, case)
    ref = orc.kantorovich(case["w"])
    check(emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], nlanes=nlanes, bin_target=bin_target), ref,
          orc.counters())


def test_moments(oracle_mod, emu_mod):
    case = common.make_case("c2", 0.004, "0.3")
    orc = common.oracle_for(oracle_mod, case)
    for order, mode in ((1, 1), (2, 2)):
        ref = orc.moments(case["w"], order)
        r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], mode=mode)
        ncol = 3 if order == 1 else 6
        scale = np.abs(ref[:, :ncol]).max(axis=0)
        assert (np.abs(r["mom"][:, :ncol] - ref[:, :ncol]).max(axis=0) <= 1e-11 * scale).all()


def test_lattice_degenerate_input(oracle_mod, emu_mod):
    """Exactly co-circular sites on the mesh diagonal: ties everywhere.  Values must still be right."""
    n = 6
    vx, vy, tri = inputs.unit_square_mesh()
    abc = inputs.pl_coefficients(vx, vy, np.ones(4), tri)
    c = (np.arange(n) + 0.5) / n
    X = np.stack(np.meshgrid(c, c, indexing="ij"), -1).reshape(-1, 2)
    r = emu_mod.evaluate(dict(kind="mesh", vx=vx, vy=vy, tri=tri, abc=abc), X, np.zeros(n * n))
    assert np.allclose(r["g"], 1.0 / n ** 2, atol=1e-14)
    Hd = r["H"].toarray()
    assert np.allclose(np.diag(Hd)[[n + 1, 2 * n + 2]], 2.0, atol=1e-12)
    assert np.abs(Hd.sum(1)).max() < 1e-12


def test_hidden_and_coincident_diracs(emu_mod):
    vx, vy, tri = inputs.unit_square_mesh()
    abc = inputs.pl_coefficients(vx, vy, np.ones(4), tri)
    mesh = dict(kind="mesh", vx=vx, vy=vy, tri=tri, abc=abc)
    X = np.array([[0.3, 0.5], [0.7, 0.5], [0.5, 0.5]])
    r = emu_mod.evaluate(mesh, X, np.array([0.0, 0.0, -1.0]))
    assert r["g"][2] == 0 and abs(r["g"].sum() - 1) < 1e-15 and r["H"].getrow(2).nnz == 0
    X = np.array([[0.3, 0.5], [0.7, 0.5], [0.7, 0.5]])  # duplicate site: exactly one of the twins keeps the cell
    r = emu_mod.evaluate(mesh, X, np.zeros(3))
    assert abs(r["g"].sum() - 1) < 1e-15 and sorted(r["g"][1:])[0] == 0


def test_capacity_overflow_is_flagged(emu_mod):
    """A cell with more neighbours than kmax must raise the overflow flag (the host then escalates)."""
    vx, vy, tri = inputs.unit_square_mesh()
    abc = inputs.pl_coefficients(vx, vy, np.ones(4), tri)
    t = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    X = np.r_[[[0.5, 0.5]], 0.5 + 0.3 * np.c_[np.cos(t), np.sin(t)]]
    r = emu_mod.evaluate(dict(kind="mesh", vx=vx, vy=vy, tri=tri, abc=abc), X, np.zeros(len(X)), kmax=16, maxv_piece=12)
    assert r["flags"] != 0
    r = emu_mod.evaluate(dict(kind="mesh", vx=vx, vy=vy, tri=tri, abc=abc), X, np.zeros(len(X)), kmax=64, maxv_piece=70)
    assert r["flags"] == 0 and abs(r["g"].sum() - 1) < 1e-14 and len(r["adjacency"][0]) == 40


@pytest.mark.parametrize("name,scale,kind", [("c2", 0.02, "lin"), ("c2", 0.02, "quad"), ("c1", 0.2, "bump")])
@pytest.mark.parametrize("seg", [False, True])
def test_weights_with_a_gradient(oracle_mod, emu_mod, name, scale, kind, seg):
    """Weights with a strong gradient displace every cell far from its Dirac (and hide some Diracs):
    K2 then leaves the ring walk for the quadtree walk with per-node supporting planes, and prunes with
    the disk around the polygon instead of the one around y_i."""
    case = common.make_case(name, scale, "0.2")
    X = case["X"]
    if kind == "lin":
        w = case["w"] + 0.3 * X[:, 0] - 0.1 * X[:, 1]
    elif kind == "quad":
        w = case["w"] + 0.25 * (X ** 2).sum(1)
    else:
        w = case["w"] + 0.05 * np.exp(-((X - X.mean(0)) ** 2).sum(1) / 0.1)
    orc = common.oracle_for(oracle_mod, case)
    f0, g0, H0 = orc.kantorovich(w)
    assert (g0 == 0).sum() > 50  # hidden Diracs (trap T2)
    r = emu_mod.evaluate(case["emu_mesh"], X, w, seg=seg)
    assert r["flags"] == 0
    assert abs(r["f"] - f0) <= 1e-10 * abs(f0)
    assert np.abs(r["g"] - g0).max() <= 1e-10 * np.abs(g0).max()
    assert common.same_pattern(H0, r["H"])
    # the oracle follows the reference's global-coordinate constructions: good to ~2e-10 of the row diagonal here
    assert common.hessian_rel_err(H0, r["H"]) <= 5e-10


@pytest.mark.parametrize("name,scale,weights", [("c2", 0.004, "0.5"), ("c3", 0.0005, "0.3"), ("c4", 0.004, "0.3")])
def test_boundary_segment_path_matches_oracle(oracle_mod, emu_mod, name, scale, weights):
    """ma_seg.cuh (what k_seg runs on grid meshes): kantorovich and both moment orders."""
    case = common.make_case(name, scale, weights)
    orc = common.oracle_for(oracle_mod, case)
    f0, g0, H0 = orc.kantorovich(case["w"])
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=True)
    assert r["flags"] == 0
    assert abs(r["f"] - f0) <= 1e-10 * abs(f0)
    assert np.abs(r["g"] - g0).max() <= 1e-10 * np.abs(g0).max()
    assert common.same_pattern(H0, r["H"])
    assert common.hessian_rel_err(H0, r["H"]) <= 1e-10
    for order in (1, 2):
        ref = orc.moments(case["w"], order)
        m = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=True, mode=order)["mom"]
        k = 3 if order == 1 else 6
        assert np.abs(m[:, :k] - ref[:, :k]).max() <= 1e-10 * np.abs(ref[:, :k]).max()


@pytest.mark.parametrize("kmax", [32, 64])
def test_array_polygon_classes(oracle_mod, emu_mod, kmax):
    """kmax = 32 / 64 are the capacity classes a cell with more than 16 vertices escalates to: polygons as
    shifted arrays (clip_rebuild) instead of the packed-order fast path (clip_packed)."""
    case = common.make_case("c3", 0.0005, "0.3")
    orc = common.oracle_for(oracle_mod, case)
    ref = orc.kantorovich(case["w"])
    check(emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], kmax=kmax, maxv_piece=20 if kmax == 32 else 40), ref,
          orc.counters())
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], kmax=kmax, seg=True)
    assert r["flags"] == 0 and common.same_pattern(ref[2], r["H"])
    assert np.abs(r["g"] - ref[1]).max() <= 1e-10 * np.abs(ref[1]).max()


@pytest.mark.parametrize("name,scale,weights", [("c1", 0.2, "zero"), ("c1", 0.2, "0.3"), ("c2", 0.02, "0.5"),
                                                ("c3", 0.003, "0.3"), ("c5", 0.0005, "zero"), ("c1r", 0.1, "zero")])
def test_block_kernel_lane_code(oracle_mod, emu_mod, name, scale, weights):
    """K2's fast path (ma_block.cuh, what k_cells_block runs): block of radius 2, then 3, then CellSearch for the
    cells neither certifies — same neighbours, masses and Hessian as the oracle, and most cells certified."""
    case = common.make_case(name, scale, weights)
    orc = common.oracle_for(oracle_mod, case)
    f0, g0, H0 = orc.kantorovich(case["w"])
    emu_mod.set_lean(True)
    try:
        r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=case["emu_mesh"]["kind"] == "grid")
        c2, c3, rest = emu_mod.lean_counts()
    finally:
        emu_mod.set_lean(False)
    assert r["flags"] == 0
    assert c2 + c3 + rest == case["N"] and c2 + c3 >= 0.9 * case["N"]
    assert abs(r["f"] - f0) <= 1e-10 * abs(f0)
    assert np.abs(r["g"] - g0).max() <= 1e-10 * np.abs(g0).max()
    assert common.same_pattern(H0, r["H"])
    assert abs(H0 - r["H"]).max() <= 1e-10 * np.abs(H0.diagonal()).max()


def test_block_kernel_degenerate_and_forced_exact(oracle_mod, emu_mod):
    """Lattice (every Voronoi vertex a co-circular quadruple), a jittered grid as in tests/bench_opttransport.cpp:45-60,
    and every sign forced through the double-double stage."""
    emu_mod.set_lean(True)
    try:
        n = 12
        vx, vy, tri = inputs.unit_square_mesh()
        abc = inputs.pl_coefficients(vx, vy, np.ones(4), tri)
        mesh = dict(kind="mesh", vx=vx, vy=vy, tri=tri, abc=abc)
        c = (np.arange(n) + 0.5) / n
        X = np.stack(np.meshgrid(c, c, indexing="ij"), -1).reshape(-1, 2)
        r = emu_mod.evaluate(mesh, X, np.zeros(n * n))
        assert np.allclose(r["g"], 1.0 / n ** 2, atol=1e-14) and np.abs(r["H"].sum(1)).max() < 1e-12
        Hd = r["H"].toarray()
        assert np.allclose(np.diag(Hd)[[n + 1, 2 * n + 2]], 2.0, atol=1e-12)
        # jittered grid on the 2 x 2 image mesh
        case = common.make_case("c1r", 0.001, "zero")
        X = np.clip(inputs.jittered_grid_points(30, 11), -1 + 1e-9, 1 - 1e-9)
        cfg = case["cfg"]
        orc = oracle_mod.Oracle(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"])
        orc.set_points(X)
        f0, g0, H0 = orc.kantorovich(np.zeros(len(X)))
        for tol in (1e-11, 1e300):
            r = emu_mod.evaluate(case["emu_mesh"], X, np.zeros(len(X)), seg=True, filter_tol=tol)
            assert sum(emu_mod.lean_counts()[:2]) > 0.9 * len(X)
            assert abs(r["f"] - f0) <= 1e-10 * abs(f0) and np.abs(r["g"] - g0).max() <= 1e-10 * g0.max()
            assert common.same_pattern(H0, r["H"])
    finally:
        emu_mod.set_lean(False)


def _grid_case(n, m, seed=0):
    vx, vy = inputs.grid_vertices(n, m)
    rng = np.random.default_rng(seed)
    rho = inputs.gaussian_mixture_density(vx, vy) + 0.05 * rng.random(n * m)  # rough: every mesh edge has a kink
    tri = inputs.grid_triangles(n, m)
    abc = inputs.pl_coefficients(vx, vy, rho, tri)
    return vx, vy, tri, abc, dict(kind="grid", n=n, m=m, abc=abc, rho=rho)


@pytest.mark.parametrize("layout", ["pixel_centres", "pixel_corners", "every_second_corner", "quarter_shift", "rows_on_lines"])
def test_segment_path_on_pixel_aligned_lattices(oracle_mod, emu_mod, layout):
    """Cells whose edges lie ON mesh lines and whose vertices sit ON mesh vertices (Diracs on a lattice aligned with the
    image grid, rough density): the sign rule of ma_seg.cuh attributes such an edge to one side and the line's chord
    supplies the difference.  (The walk-along-the-edges formulation of round 1 was off by 4e-4 here.)  Checked against
    the oracle's per-cell mode; its BFS mode, like vti.hpp:219-313, loses the traversal on exact ties."""
    n, m = 17, 13
    vx, vy, tri, abc, mesh = _grid_case(n, m, 3)
    hx, hy = 2.0 / (n - 1), 2.0 / (m - 1)
    if layout == "pixel_centres":
        pts = [(-1 + hx * (a + 0.5), -1 + hy * (b + 0.5)) for a in range(n - 1) for b in range(m - 1)]
    elif layout == "pixel_corners":
        pts = [(-1 + hx * a, -1 + hy * b) for a in range(1, n - 1) for b in range(1, m - 1)]
    elif layout == "every_second_corner":
        pts = [(-1 + hx * a, -1 + hy * b) for a in range(1, n - 1, 2) for b in range(1, m - 1, 2)]
    elif layout == "quarter_shift":
        pts = [(-1 + hx * (a + 0.25), -1 + hy * (b + 0.75)) for a in range(n - 1) for b in range(m - 1)]
    else:  # Diracs between the rows: horizontal cell edges on y-lines, vertical ones generic
        pts = [(-1 + hx * (a + 0.37), -1 + hy * (b + 0.5)) for a in range(n - 1) for b in range(m - 1)]
    X = np.array(pts)
    orc = oracle_mod.Oracle(vx, vy, tri, abc)
    orc.set_points(X)
    for w in (np.zeros(len(X)), 1e-3 * np.sin(7 * X[:, 0]) * np.cos(5 * X[:, 1])):
        f0, g0, H0 = orc.kantorovich(w, mode=oracle_mod.MODE_PER_CELL | 1)
        r = emu_mod.evaluate(mesh, X, w, seg=True)
        assert r["flags"] == 0
        assert abs(g0.sum() - inputs.total_mass(vx, vy, tri, abc)) <= 1e-13
        assert abs(r["f"] - f0) <= 1e-11 * abs(f0)
        assert np.abs(r["g"] - g0).max() <= 1e-11 * np.abs(g0).max()
        d = np.abs(H0.diagonal()).max()
        assert abs(H0 - r["H"]).max() <= 1e-11 * d
        mom0 = orc.moments(w, 2, mode=oracle_mod.MODE_PER_CELL | 1)
        mom = emu_mod.evaluate(mesh, X, w, seg=True, mode=2)["mom"]
        assert (np.abs(mom - mom0).max(axis=0) <= 1e-11 * np.abs(mom0).max(axis=0)).all()


def test_segment_path_many_boundary_cells(oracle_mod, emu_mod):
    """Few Diracs on a fine non-square grid: every cell spans many squares and most touch the domain boundary."""
    n, m = 41, 29
    vx, vy, tri, abc, mesh = _grid_case(n, m, 5)
    rng = np.random.default_rng(1)
    for N in (3, 17, 120):
        X = rng.uniform(-0.999, 0.999, (N, 2))
        orc = oracle_mod.Oracle(vx, vy, tri, abc)
        orc.set_points(X)
        w = rng.normal(0, 0.02 / N, N)
        f0, g0, H0 = orc.kantorovich(w)
        r = emu_mod.evaluate(mesh, X, w, seg=True)
        assert abs(r["f"] - f0) <= 1e-11 * abs(f0)
        assert np.abs(r["g"] - g0).max() <= 1e-11 * np.abs(g0).max()
        assert common.same_pattern(H0, r["H"])
        assert abs(H0 - r["H"]).max() <= 1e-11 * np.abs(H0.diagonal()).max()


# ------------------------------------------------------------------------------------------------
# K2 warm path (ma_warm.cuh): cells from the adjacency of an earlier evaluation + the ring-match certificate
# ------------------------------------------------------------------------------------------------
def _graded(case, amp, seed):
    """Smooth weights with a large gradient (cells displaced from their Diracs) plus noise, the Newton-iterate regime."""
    X = case["X"]
    ext = X.max(0) - X.min(0)
    u = (X - X.min(0)) / ext
    cell = ext.prod() / len(X)
    rng = np.random.default_rng(seed)
    return amp * ext.prod() * (0.3 * np.sin(2.1 * u[:, 0] + 0.3) * np.cos(1.7 * u[:, 1]) + 0.1 * u[:, 0] * u[:, 1]) + 0.03 * cell * rng.normal(size=len(X))


def _same_raw(a, b):
    """Same cells: identical neighbour sets (the polygons may start at another vertex, so slot order and the last bits of
    the sums may differ), values equal to rounding."""
    assert a["flags"] == b["flags"] == 0
    assert a["adjacency"] == b["adjacency"]
    assert np.array_equal(a["raw"]["nbr_cnt"], b["raw"]["nbr_cnt"])
    assert np.abs(a["g"] - b["g"]).max() <= 1e-14 * np.abs(a["g"]).max()
    assert abs(a["f"] - b["f"]) <= 1e-13 * abs(a["f"])
    assert common.same_pattern(a["H"], b["H"])
    assert abs(a["H"] - b["H"]).max() <= 1e-13 * np.abs(a["H"].diagonal()).max()


@pytest.mark.parametrize("name,scale,seg", [("c3", 0.002, True), ("c2", 0.02, True), ("c1", 0.2, False)])
@pytest.mark.parametrize("step", [0.02, 0.3])
def test_warm_path_gives_the_cells_of_the_cold_one(emu_mod, name, scale, seg, step):
    """Seeds from an evaluation at w1; the evaluation at w2 = w1 + step * (another graded field) through the warm path must
    give the same cells as without seeds: identical adjacency, masses and Hessian equal to rounding.  A small step rebuilds a few cells,
    a large one a good part of them — the certificate decides, never the answer."""
    case = common.make_case(name, scale, "zero")
    w1 = _graded(case, 0.05, 1)
    w2 = w1 + step * _graded(case, 0.05, 2)
    first = emu_mod.evaluate(case["emu_mesh"], case["X"], w1, seg=seg)
    cold = emu_mod.evaluate(case["emu_mesh"], case["X"], w2, seg=seg)
    warm = emu_mod.evaluate(case["emu_mesh"], case["X"], w2, seg=seg, seeds=first["seeds"])
    used, r0, r1, f0, f1, f2 = emu_mod.warm_counts()
    assert used == 1 and f2 == 0
    N = case["N"]
    assert r0 + r1 < (0.25 if step < 0.1 else 0.95) * N, (r0, r1, N)
    _same_raw(cold, warm)


def test_warm_path_with_identical_weights_rebuilds_nothing(emu_mod):
    case = common.make_case("c3", 0.002, "zero")
    w = _graded(case, 0.05, 3)
    first = emu_mod.evaluate(case["emu_mesh"], case["X"], w, seg=True)
    warm = emu_mod.evaluate(case["emu_mesh"], case["X"], w, seg=True, seeds=first["seeds"])
    hidden = int((first["raw"]["nbr_cnt"] < 0).sum())  # no seed row: CellSearch has to look (the cell may have come back)
    assert emu_mod.warm_counts() == (1, hidden, 0, hidden, 0, 0)
    _same_raw(first, warm)


def test_ring_match_holds_on_every_cold_diagram(emu_mod):
    """The certificate itself: with seeds = the TRUE adjacency the match must pass at every vertex — interior vertices
    (three cells), vertices on the sides of the box (two cells) and the corners — on uniform and on graded weights."""
    for name, scale, weights in (("c2", 0.01, "zero"), ("c1", 0.1, "0.4"), ("c5", 0.0003, "0.2")):
        case = common.make_case(name, scale, weights)
        first = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"])
        again = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seeds=first["seeds"])
        hidden = int((first["raw"]["nbr_cnt"] < 0).sum())
        assert emu_mod.warm_counts() == (1, hidden, 0, hidden, 0, 0), (name, emu_mod.warm_counts())
        _same_raw(first, again)


def test_warm_path_with_hidden_and_emptied_cells(emu_mod):
    """A Dirac that is hidden in the seed evaluation has no seed row (rebuilt by CellSearch); one that becomes hidden at
    the new weights is found empty on its superset.  Either way the answer is the cold one."""
    case = common.make_case("c2", 0.01, "zero")
    N = case["N"]
    w1 = np.zeros(N)
    w1[5] = -1.0  # hidden in the seed evaluation
    first = emu_mod.evaluate(case["emu_mesh"], case["X"], w1, seg=True)
    w2 = np.zeros(N)
    w2[7] = -1.0  # hidden now
    cold = emu_mod.evaluate(case["emu_mesh"], case["X"], w2, seg=True)
    warm = emu_mod.evaluate(case["emu_mesh"], case["X"], w2, seg=True, seeds=first["seeds"])
    assert emu_mod.warm_counts()[0] == 1
    assert cold["g"][7] == 0.0 and cold["g"][5] > 0.0
    _same_raw(cold, warm)


def test_warm_path_gives_up_on_degenerate_lattices(emu_mod):
    """Co-circular quadruples everywhere: with ties outside the rings of the four cells round a vertex do not pair up, the
    last match still reports failures and the evaluation is redone without seeds — same answer, no certificate claimed."""
    n = 12
    g = (np.arange(n) + 0.5) / n * 2 - 1
    X = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    case = common.make_case("c2", 0.004, "zero")
    w = np.zeros(len(X))
    first = emu_mod.evaluate(case["emu_mesh"], X, w, seg=True)
    warm = emu_mod.evaluate(case["emu_mesh"], X, w, seg=True, seeds=first["seeds"])
    used = emu_mod.warm_counts()[0]
    _same_raw(first, warm)
    assert used in (0, 1)


# ------------------------------------------------------------------------------------------------
# K3 on grids whose squares are split along EITHER diagonal (what CGAL's Delaunay of the pixel grid produces)
# ------------------------------------------------------------------------------------------------
def _random_diagonal_case(name, scale, weights, seed, X=None):
    case = common.make_case(name, scale, weights)
    cfg = case["cfg"]
    n, m = cfg["n"], cfg["m"]
    diag = np.random.default_rng(seed).integers(0, 2, (n - 1, m - 1))
    tri = inputs.grid_triangles_diag(n, m, diag)
    cfg = dict(cfg, tri=tri)
    abc = inputs.pl_coefficients(cfg["vx"], cfg["vy"], cfg["rho"], tri)
    case = dict(case, cfg=cfg, abc=abc, emu_mesh=dict(case["emu_mesh"], abc=abc, diag=diag))
    if X is not None:
        case["X"] = X
        case["N"] = len(X)
        case["w"] = np.zeros(len(X))
    return case


@pytest.mark.parametrize("name,scale,weights", [("c2", 0.004, "zero"), ("c2", 0.004, "0.5"), ("c3", 0.0005, "0.3"), ("c4", 0.002, "0.2")])
def test_segment_path_with_random_diagonals(oracle_mod, emu_mod, name, scale, weights):
    """The oracle sees an explicit triangulation (general-mesh enumeration), the line-major K3 the grid + one bit per square."""
    case = _random_diagonal_case(name, scale, weights, seed=5)
    orc = common.oracle_for(oracle_mod, case)
    f0, g0, H0 = orc.kantorovich(case["w"], mode=oracle_mod.MODE_PER_CELL)
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=True)
    assert r["flags"] == 0
    assert abs(r["f"] - f0) <= 1e-10 * abs(f0)
    assert np.abs(r["g"] - g0).max() <= 1e-10 * np.abs(g0).max()
    assert common.same_pattern(H0, r["H"])
    assert abs(H0 - r["H"]).max() <= 1e-10 * np.abs(H0.diagonal()).max()
    # moments through the same kernel
    m0 = orc.moments(case["w"], 2) if hasattr(orc, "moments") else None
    if m0 is not None:
        rm = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=True, mode=2)
        ref = np.column_stack([m0[0], m0[1], m0[2]]) if isinstance(m0, tuple) else np.asarray(m0)
        assert np.abs(rm["mom"] - ref).max() <= 1e-10 * np.abs(ref).max()


def test_random_diagonals_on_pixel_aligned_lattice(oracle_mod, emu_mod):
    """Cell edges ON grid lines and cell vertices ON grid vertices, every second square split the other way."""
    case0 = common.make_case("c2", 0.004, "zero")
    n = case0["cfg"]["n"]
    k = (n - 1) // 2
    g = -1.0 + (2.0 * np.arange(k) + 1.0) * (2.0 / (n - 1))  # Diracs at the odd grid vertices: cells = 2 x 2 blocks of squares
    X = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    case = _random_diagonal_case("c2", 0.004, "zero", seed=9, X=X)
    orc = common.oracle_for(oracle_mod, case)
    f0, g0, H0 = orc.kantorovich(case["w"], mode=oracle_mod.MODE_PER_CELL)
    r = emu_mod.evaluate(case["emu_mesh"], case["X"], case["w"], seg=True)
    assert np.abs(r["g"] - g0).max() <= 1e-10 * np.abs(g0).max()
    assert abs(r["f"] - f0) <= 1e-10 * abs(f0)

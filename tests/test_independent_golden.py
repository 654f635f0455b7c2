"""The independent golden vectors (tests/golden/make_independent.py: exact rational arithmetic, every site against every
other site, every cell against every triangle, different quadrature rules, no oracle and no engine involved) against
  * the oracle                      (CPU: this is what pins the oracle to something other than itself),
  * the emulated kernel lane code   (CPU),
  * the CUDA engine through the C-ABI (-m gpu)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from mongeampere_b200 import inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["indep_square_n40", "indep_grid5x4_n60", "indep_grid9_n30_w0", "indep_grid6x5_altdiag_n50"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    abc = inputs.pl_coefficients(z["vx"], z["vy"], z["rho"], z["tri"])
    return z, abc


def check(z, f, g, H, tol=1e-12):
    H = sp.csr_matrix(H).toarray()
    assert abs(f - float(z["f"])) <= tol * abs(float(z["f"]))
    assert np.abs(g - z["g"]).max() <= tol * np.abs(z["g"]).max()
    d = np.abs(np.diag(z["H"])).max()
    assert np.abs(H - z["H"]).max() <= tol * d
    # pattern: identical up to entries that are zero to rounding (edges of zero length)
    assert np.array_equal(np.abs(H) > 1e-12 * d, np.abs(z["H"]) > 1e-12 * d)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_independent_vectors(oracle_mod, name):
    z, abc = load(name)
    orc = oracle_mod.Oracle(z["vx"], z["vy"], z["tri"], abc)
    orc.set_points(z["X"])
    for mode in (oracle_mod.MODE_BRUTE, oracle_mod.MODE_BRUTE | oracle_mod.MODE_PER_CELL, 0):
        f, g, H = orc.kantorovich(z["w"], mode=mode)
        check(z, f, g, H)
    assert orc.counters()["pieces"] == int(z["npieces"])


@pytest.mark.parametrize("name", NAMES)
def test_emulated_kernels_match_independent_vectors(emu_mod, name):
    z, abc = load(name)
    grid = str(z["kind"]) != "square"
    alt = "diag" in z.files  # squares split along either diagonal: explicit mesh for the piece kernel, grid + bits for k_seg
    gmesh = dict(kind="grid", n=int(z["n"]), m=int(z["m"]), abc=abc, rho=z["rho"], diag=z["diag"] if alt else None)
    emesh = dict(kind="mesh", vx=z["vx"], vy=z["vy"], tri=z["tri"], abc=abc)
    for lean in (False, True):
        emu_mod.set_lean(lean)
        try:
            for seg in ((False, True) if grid else (False,)):
                mesh = gmesh if grid and (seg or not alt) else emesh
                r = emu_mod.evaluate(mesh, z["X"], z["w"], seg=seg, maxv_piece=24)
                assert r["flags"] == 0
                check(z, r["f"], r["g"], r["H"])
        finally:
            emu_mod.set_lean(False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_engine_matches_independent_vectors(gpu_ctx, name):
    z, abc = load(name)
    if str(z["kind"]) == "square":
        gpu_ctx.set_mesh(z["vx"], z["vy"], z["tri"], abc)
        variants = [0]
    elif "diag" in z.files:  # an explicit triangulation that ma_set_mesh recognises as a grid (either diagonal per square)
        gpu_ctx.set_mesh(z["vx"], z["vy"], z["tri"], abc)
        assert gpu_ctx.info("grid_overlay") == 1
        variants = [0, 2]  # k_seg with the diagonal bits, and the general-mesh piece kernel
    else:
        gpu_ctx.set_grid(int(z["n"]), int(z["m"]), z["rho"])
        variants = [0, 2]  # boundary-integration K3 and piece-clipping K3
    gpu_ctx.set_points(z["X"])
    try:
        for strat in variants:
            gpu_ctx.set_option("strategy", strat)
            for lean in (1, 0):
                gpu_ctx.set_option("lean", lean)
                f, g, H = gpu_ctx.kantorovich(z["w"])
                check(z, f, g, H)
    finally:
        gpu_ctx.set_option("strategy", 0)
        gpu_ctx.set_option("lean", 1)

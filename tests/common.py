"""Shared helpers of the parity tests: build the same inputs for the oracle, the emulation harness and
the CUDA engine."""
import numpy as np

from mongeampere_b200 import inputs


def make_case(name, scale, weights="zero", seed=0):
    """-> dict(cfg, abc, X, w, emu_mesh)."""
    cfg = inputs.config(name, scale)
    abc = inputs.pl_coefficients(cfg["vx"], cfg["vy"], cfg["rho"], cfg["tri"])
    X = cfg["X"]
    N = len(X)
    if weights == "zero":
        w = np.zeros(N)
    else:
        # random weights small enough to keep (almost) every cell non-empty
        ext = max(cfg["vx"].max() - cfg["vx"].min(), cfg["vy"].max() - cfg["vy"].min())
        cell = ext * ext / N
        w = np.random.default_rng(seed).normal(0.0, float(weights) * cell, N)
    if cfg["kind"] == "grid":
        emu_mesh = dict(kind="grid", n=cfg["n"], m=cfg["m"], abc=abc, rho=cfg["rho"])
    else:
        emu_mesh = dict(kind="mesh", vx=cfg["vx"], vy=cfg["vy"], tri=cfg["tri"], abc=abc)
    return dict(cfg=cfg, abc=abc, X=X, w=w, emu_mesh=emu_mesh, N=N)


def oracle_for(O, case, nthreads=1):
    cfg = case["cfg"]
    orc = O.Oracle(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"], nthreads=nthreads)
    orc.set_points(case["X"])
    return orc


def load_engine(ctx, case, as_general_mesh=False):
    cfg = case["cfg"]
    if cfg["kind"] == "grid" and not as_general_mesh:
        ctx.set_grid(cfg["n"], cfg["m"], cfg["rho"])
    else:
        ctx.set_mesh(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"])
    ctx.set_points(case["X"])


def hessian_rel_err(H_ref, H):
    """max |H - H_ref| relative to the row diagonal."""
    import scipy.sparse as sp
    D = sp.csr_matrix(H_ref - H)
    if D.nnz == 0:
        return 0.0
    diag = np.maximum(np.abs(H_ref.diagonal()), 1e-300)
    rows = np.repeat(np.arange(D.shape[0]), np.diff(D.indptr))
    return float(np.max(np.abs(D.data) / diag[rows]))


def same_pattern(H_ref, H):
    A = H_ref.copy(); A.sort_indices()
    B = H.copy(); B.sort_indices()
    return A.nnz == B.nnz and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)

"""The BASELINE.json workloads as ready-to-load cases (one description of the inputs for the engine and for the test harnesses).
Used by bench.py, scripts/ and tests/; numpy only."""
import numpy as np

from . import inputs


def make_case(name, scale, weights="zero", seed=0):
    """-> dict(cfg, abc, X, w, emu_mesh, N).  weights: 'zero' or a float = std of i.i.d. normal weights in cell areas."""
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


def load_engine(ctx, case, as_general_mesh=False):
    """Mesh + Diracs of a case into a capi.Context (grid meshes through ma_set_grid unless asked otherwise)."""
    cfg = case["cfg"]
    if cfg["kind"] == "grid" and not as_general_mesh:
        ctx.set_grid(cfg["n"], cfg["m"], cfg["rho"])
    else:
        ctx.set_mesh(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"])
    ctx.set_points(case["X"])

"""Runs every BASELINE.json config once at full size on one GPU (short versions of the long ones) and prints a
JSON summary: a smoke test of the whole path at the named sizes plus the timings quoted in DESIGN.md §5."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common

out = {}
only = sys.argv[1:] or ["c1", "c3", "c4", "c5"]


def total_mass(ctx, N):
    tm = getattr(ctx, "total_mass", None)
    return tm if tm else float(ctx.kantorovich(np.zeros(N), hessian=False)[1].sum())


if "c1" in only:  # configs[0]: 10k Diracs, uniform density on the unit square, full damped-Newton solve
    case = common.make_case("c1", 1.0, "zero")
    ctx = capi.Context(0); common.load_engine(ctx, case)
    tm = total_mass(ctx, case["N"])
    t = time.perf_counter()
    w, st, rc = ctx.ot_solve(np.full(case["N"], tm / case["N"]), eps_g=1e-7, maxiter=100)
    out["c1_newton"] = dict(N=case["N"], seconds=time.perf_counter() - t, status=capi.STATUS_NAMES[rc], niter=st["niter"],
                            neval=st["neval"], cg_iters=st["cg_iters"], final_norm=st["final_norm"])
    ctx.close()
if "c3" in only:  # configs[2]: 1M Diracs on the 2048^2 image triangulation: evaluation + the first Newton iterations
    case = common.make_case("c3", 1.0, "zero")
    ctx = capi.Context(0); common.load_engine(ctx, case)
    nu = np.full(case["N"], ctx.total_mass / case["N"])
    t = time.perf_counter()
    w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=3)
    dt = time.perf_counter() - t
    f, g, H = ctx.kantorovich(w)
    out["c3_newton_first_iterations"] = dict(N=case["N"], seconds=dt, status=capi.STATUS_NAMES[rc], niter=st["niter"],
                                             neval=st["neval"], cg_iters=st["cg_iters"], norm=st["final_norm"],
                                             mass_err=abs(g.sum() - ctx.total_mass) / ctx.total_mass, kmax=int(ctx.info("kmax")))
    ctx.close()
if "c4" in only:  # configs[3]: Lloyd quantization, 250k points, exact centroids (tests/test_lloyd.cpp:52-56), 10 of the 50 iterations
    case = common.make_case("c4", 1.0, "zero")
    ctx = capi.Context(0); common.load_engine(ctx, case)
    X = case["X"].copy()
    t = time.perf_counter()
    move = []
    for it in range(10):
        ctx.set_points(X)
        m, c = ctx.lloyd(np.zeros(len(X)))
        move.append(float(np.abs(c - X).max()))
        X = c
    out["c4_lloyd_10_iterations"] = dict(N=len(X), seconds=time.perf_counter() - t, first_move=move[0], last_move=move[-1],
                                         mass_err=abs(m.sum() - ctx.total_mass) / ctx.total_mass)
    ctx.close()
if "c5" in only:  # configs[4]: 4M Diracs, uniform density on 2 triangles: evaluation + first Newton iterations
    case = common.make_case("c5", 1.0, "zero")
    ctx = capi.Context(0); common.load_engine(ctx, case)
    tm = total_mass(ctx, case["N"])
    t = time.perf_counter()
    w, st, rc = ctx.ot_solve(np.full(case["N"], tm / case["N"]), eps_g=1e-7, maxiter=1)
    out["c5_newton_first_iterations"] = dict(N=case["N"], seconds=time.perf_counter() - t, status=capi.STATUS_NAMES[rc],
                                             niter=st["niter"], neval=st["neval"], cg_iters=st["cg_iters"], norm=st["final_norm"])
    ctx.close()
print(json.dumps(out))

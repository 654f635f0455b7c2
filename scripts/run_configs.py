"""The BASELINE.json configs as the reference's drivers state them, at full size on the GPU(s):
  c1  tests/test_opttransport.cpp            10 k Diracs, uniform density on the unit square: damped-Newton solve
  c2  bench_opttransport.cpp-style           100 k Diracs, 512^2 PL mixture: damped-Newton solve
  c3  image triangulation                    1 M Diracs, 2048^2 image: damped-Newton solve
  c4  tests/test_lloyd.cpp:48-56             250 k points, 50 Lloyd iterations X <- centroids
  c5  tests/test_zeldovich.cpp:101-120       4 M Diracs on uniform density: outer loop { ot_solve from w = 0;
                                             barycentres of the Laguerre cells; X += 0.03 (X - bary) }
    python scripts/run_configs.py c4 c5                                  # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/run_configs.py c5    # NCCL inside the engine
Prints one JSON object (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")  # out-of-band channel for the NCCL id only
only = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2", "c3", "c4", "c5"]
outer = int(os.environ.get("ZELDOVICH_OUTER", "3"))
scale = float(os.environ.get("SCALE", "1.0"))
out = {"gpus": world}


def progress(*a):
    if rank == 0:
        print("[run_configs]", *a, file=sys.stderr, flush=True)


def context(case):
    ctx = capi.Context(local)
    common.load_engine(ctx, case)
    if world > 1:
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])
    return ctx


for name in ("c1", "c2", "c3"):
    if name not in only:
        continue
    progress(name, "building the case")
    case = common.make_case(name, scale, "zero")
    ctx = context(case)
    progress(name, "solving")
    nu = np.full(case["N"], ctx.total_mass / case["N"])
    t = time.perf_counter()
    w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=3000)
    out[name + "_newton"] = dict(N=case["N"], seconds=time.perf_counter() - t, status=capi.STATUS_NAMES[rc], niter=st["niter"],
                                 neval=st["neval"], cg_iters=st["cg_iters"], final_norm=st["final_norm"])
    progress(name, out[name + "_newton"])
    ctx.close()
if "c4" in only:  # tests/test_lloyd.cpp:48-56
    case = common.make_case("c4", scale, "zero")
    ctx = context(case)
    X = case["X"].copy()
    move = []
    t = time.perf_counter()
    for it in range(50):
        ctx.set_points(X)
        m, c = ctx.lloyd(np.zeros(len(X)))
        move.append(float(np.abs(c - X).max()))
        if it % 10 == 0:
            progress("c4 iteration", it, "move", move[-1], "t", time.perf_counter() - t)
        X = c
    dt = time.perf_counter() - t
    out["c4_lloyd_50_iterations"] = dict(N=len(X), seconds=dt, ms_per_iteration=1e3 * dt / 50, first_move=move[0], last_move=move[-1],
                                         mass_err=abs(m.sum() - ctx.total_mass) / ctx.total_mass)
    ctx.close()
if "c5" in only:  # tests/test_zeldovich.cpp:101-120
    case = common.make_case("c5", scale, "zero")
    ctx = context(case)
    X = case["X"].copy()
    N = len(X)
    nu = np.full(N, ctx.total_mass / N)
    log = []
    t0 = time.perf_counter()
    for it in range(outer):
        progress("c5 outer", it, "set_points")
        ctx.set_points(X)
        progress("c5 outer", it, "ot_solve")
        t = time.perf_counter()
        w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=100)           # weights start from zero every time (:103,106)
        t_solve = time.perf_counter() - t
        m, bary = ctx.lloyd(w)                                           # uniform density: rho-centroid = area centroid (:58-74)
        X = X + 0.03 * (X - bary)                                        # :117-118
        log.append(dict(outer=it, solve_seconds=t_solve, status=capi.STATUS_NAMES[rc], niter=st["niter"], neval=st["neval"],
                        cg_iters=st["cg_iters"], final_norm=st["final_norm"], max_push=float(np.abs(0.03 * (X - bary)).max())))
        progress(log[-1])
    out["c5_zeldovich"] = dict(N=N, outer_iterations=outer, seconds=time.perf_counter() - t0, log=log)
    ctx.close()
if rank == 0:
    print(json.dumps(out))
if dist is not None:
    dist.destroy_process_group()

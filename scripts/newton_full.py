"""Full damped-Newton solve of one BASELINE config on one GPU (metric 2), with the engine's own split of the wall time
(MA_TRACE=1 prints per-evaluation stage times).  usage: newton_full.py c3 [scale] [maxiter] [out.npy]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
case = common.make_case(name, scale, "zero")
ctx = capi.Context(0)
common.load_engine(ctx, case)
for kv in os.environ.get("MA_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
N = case["N"]
nu = np.full(N, ctx.total_mass / N)
t = time.perf_counter()
w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=maxiter, verbose=bool(int(os.environ.get("VERBOSE", "0"))))
dt = time.perf_counter() - t
f, g, H = ctx.kantorovich(w)
print(json.dumps(dict(workload=name, scale=scale, N=N, seconds=dt, status=capi.STATUS_NAMES[rc], niter=st["niter"],
                      neval=st["neval"], cg_iters=st["cg_iters"], final_norm=st["final_norm"],
                      check_norm=float(np.linalg.norm(g - nu)), mass_err=abs(g.sum() - ctx.total_mass) / ctx.total_mass)))
if len(sys.argv) > 4:
    np.save(sys.argv[4], w)
ctx.close()

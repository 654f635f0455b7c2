"""Newton solve with per-evaluation stage timings (MA_TRACE=1) — diagnostic."""
import os, sys, time
os.environ.setdefault("MA_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 3
case = common.make_case(name, scale, "zero")
ctx = capi.Context(0)
for kv in filter(None, os.environ.get("MA_OPTS", "").split(",")):  # before the points are binned (bin_target)
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
common.load_engine(ctx, case)
tm = ctx.total_mass if getattr(ctx, "total_mass", None) else float(ctx.kantorovich(np.zeros(case["N"]))[1].sum())
nu = np.full(case["N"], tm / case["N"])
t = time.time()
w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=maxiter, verbose=True)
print("rc", rc, st, "wall", time.time() - t)

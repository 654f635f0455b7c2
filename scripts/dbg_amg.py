"""AMG-preconditioned PCG vs Jacobi PCG inside the Newton solve (same trajectory, fewer iterations)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
res = {}
for amg in (1, 0):
    case = common.make_case(name, scale, "zero")
    ctx = capi.Context(0)
    ctx.set_option("amg", amg)
    for kv in filter(None, os.environ.get("MA_OPTS", "").split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    common.load_engine(ctx, case)
    N = case["N"]
    nu = np.full(N, ctx.total_mass / N)
    t = time.time()
    w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=3000)
    res[amg] = (w, st)
    print(name, scale, "amg", amg, "rc", rc, st, round(time.time() - t, 2), flush=True)
    ctx.close()
w1, w0 = res[1][0], res[0][0]
print("weights amg vs jacobi", np.abs((w1 - w1[-1]) - (w0 - w0[-1])).max())

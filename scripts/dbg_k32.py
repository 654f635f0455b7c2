"""K2 at the larger capacity classes on the first Newton iterate.  usage: dbg_k32.py scale kmax persist [maxiter]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common

scale = float(sys.argv[1]); kmax = int(sys.argv[2]); persist = int(sys.argv[3])
maxiter = int(sys.argv[4]) if len(sys.argv) > 4 else 1
case = common.make_case("c3", scale, "zero")
ctx = capi.Context(0)
common.load_engine(ctx, case)
N = case["N"]
nu = np.full(N, ctx.total_mass / N)
w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=maxiter)
cell = 4.0 / N
print("newton:", capi.STATUS_NAMES[rc], st["niter"], st["neval"], "w range in cell areas", (w.max() - w.min()) / cell, flush=True)
ctx.set_option("persist", persist)
ctx.set_option("kmax", kmax)
for kv in os.environ.get("MA_OPTS2", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
for ws, tag in ((np.zeros(N), "w=0"), (w, "w1")):
    print(f"kmax {kmax} persist {persist} {tag} ...", file=sys.stderr, flush=True)
    t = time.perf_counter()
    f, g, H = ctx.kantorovich(ws)
    t2 = time.perf_counter() - t
    print(f"kmax {kmax} persist {persist} {tag}: kantorovich {t2*1e3:.1f} ms  f {f:.12g} min g {g.min():.3g}", flush=True)
ctx.close()

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from oracle import oracle as O
from tests import common
O.build()
NT = os.cpu_count()
# 1. c1 Newton with lean on / off
for lean in (1, 0):
    case = common.make_case("c1", 1.0, "zero")
    ctx = capi.Context(0)
    ctx.set_option("lean", lean)
    common.load_engine(ctx, case)
    N = case["N"]
    f, g, H = ctx.kantorovich(np.zeros(N))
    nu = np.full(N, g.sum() / N)
    t = time.time()
    w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=12, verbose=(lean == 1))
    print("c1 lean", lean, "rc", rc, st, round(time.time() - t, 2), flush=True)
    ctx.close()
# 2. converged c2 x 0.3: which K3 deviates from the oracle
case = common.make_case("c2", 0.3, "zero")
ctx = capi.Context(0)
common.load_engine(ctx, case)
N = case["N"]
nu = np.full(N, ctx.total_mass / N)
w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=2000)
print("c2x0.3 newton", rc, st, flush=True)
orc = common.oracle_for(O, case, nthreads=NT)
f0, g0, H0 = orc.kantorovich(w, mode=O.MODE_PER_CELL)
for strat in (0, 2):
    ctx.set_option("strategy", strat)
    f1, g1, H1 = ctx.kantorovich(w)
    d = np.abs(g1 - g0) / np.abs(g0).max()
    k = np.argsort(-d)[:5]
    print("strategy", strat, "df", abs(f1 - f0) / abs(f0), "dg max", d.max(), "worst cells", k, d[k], "H", abs(H0 - H1).max() / np.abs(H0.diagonal()).max(), flush=True)
    if strat == 0: gs = g1
    else: print("   seg vs pieces dg", np.abs(gs - g1).max() / np.abs(g0).max())
np.save("gpurun_out/r2c_c2x03_w.npy", w)
ctx.close()

"""Small end-to-end check on the GPU (used under compute-sanitizer and for quick timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi
from oracle import oracle as O
from tests import common

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.003
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c1", "c2"]
ctx = capi.Context(0)
for name in names:
    case = common.make_case(name, scale, "0.4")
    common.load_engine(ctx, case)
    ctx.set_stats(True)
    t = time.time(); f1, g1, H1 = ctx.kantorovich(case["w"]); t1 = time.time() - t
    orc = common.oracle_for(O, case)
    t = time.time(); f0, g0, H0 = orc.kantorovich(case["w"]); t0 = time.time() - t
    print(name, "N", case["N"], "f", f0, f1, "df", abs(f1 - f0), "dg", np.abs(g1 - g0).max() / np.abs(g0).max(),
          "pattern", common.same_pattern(H0, H1), "dH", common.hessian_rel_err(H0, H1) if common.same_pattern(H0, H1) else None,
          "nnz", H0.nnz, H1.nnz, "t_gpu", round(t1, 4), "t_cpu", round(t0, 4))
    print("  ", orc.counters()); print("  ", ctx.counters())
    nu = np.full(case["N"], ctx.total_mass / case["N"]) if ctx.total_mass else np.full(case["N"], g0.sum() / case["N"])
    w, st, rc = ctx.ot_solve(nu, eps_g=1e-8, verbose=False)
    print("   ot_solve rc", rc, st)
    m, cen = ctx.lloyd()
    mo = orc.lloyd(np.zeros(case["N"]))
    print("   lloyd dm", np.abs(m - mo[0]).max(), "dc", np.nanmax(np.abs(cen - mo[1])))
ctx.close()
print("done")

"""Minimal driver for ncu captures: loads a workload and runs a few device-resident evaluations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
nev = int(sys.argv[3]) if len(sys.argv) > 3 else 3
weights = sys.argv[4] if len(sys.argv) > 4 else "zero"
wfile = None
if weights.startswith("file:"):
    wfile, weights = weights[5:], "zero"
if weights.startswith("newton:"):  # weights after that many Newton iterations from w = 0
    wfile, nit, weights = "", int(weights[7:]), "zero"
case = common.make_case(name, scale, weights)
ctx = capi.Context(0)
for kv in filter(None, os.environ.get("MA_OPTS", "").split(",")):  # e.g. MA_OPTS=persist=0,bin_target=2
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
common.load_engine(ctx, case)
if wfile is not None:
    import numpy as np
    if wfile == "":
        nu = np.full(case["N"], ctx.total_mass / case["N"])
        case["w"], st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=nit, verbose=False)
        print("newton warm-up", st)
    else:
        case["w"] = np.load(wfile)
ctx.set_weights(case["w"])
if os.environ.get("MA_PART"):  # e.g. MA_PART=0,8: evaluate only Morton tile 0 of 8 (what one rank of an 8-GPU run does)
    r, n = map(int, os.environ["MA_PART"].split(","))
    ctx.set_partition(r, n)
if os.environ.get("MA_PROFILER_START"):  # with `ncu --profile-from-start off`: capture only the evaluations below
    import ctypes, glob
    libs = sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
    ctypes.CDLL(libs[0]).cudaProfilerStart()
for _ in range(nev):
    ctx.evaluate(True)
print("nnz", ctx.info("nnz"), "mass_sum", ctx.info("mass_sum"))
ctx.close()

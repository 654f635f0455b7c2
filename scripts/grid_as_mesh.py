"""c3 handed over as an EXPLICIT triangulation with random per-square diagonals (what CGAL's Delaunay of the pixel grid gives):
ma_set_mesh recognises the grid; per-evaluation times on the recognised path, on k_pieces, and on the ma_set_grid path."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mongeampere_b200 import capi, inputs
from mongeampere_b200 import workloads as common

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
case = common.make_case("c3", scale, "zero")
cfg = case["cfg"]
n, m = cfg["n"], cfg["m"]
diag = np.random.default_rng(3).integers(0, 2, (n - 1, m - 1))
tri = inputs.grid_triangles_diag(n, m, diag)
abc = inputs.pl_coefficients(cfg["vx"], cfg["vy"], cfg["rho"], tri)
ctx = capi.Context(0)
out = {"N": case["N"], "faces": int(len(tri))}


def timed(tag, steps=10):
    ctx.set_weights(case["w"])
    for _ in range(3):
        ctx.evaluate(True)
    ctx.timer_start()
    for _ in range(steps):
        ctx.evaluate(True)
    out[tag + "_ms"] = ctx.timer_stop() / steps
    out[tag + "_f"] = ctx.info("fval")


t = time.perf_counter()
ctx.set_grid(n, m, cfg["rho"]); ctx.set_points(case["X"])
timed("set_grid_fixed_diagonal")
t = time.perf_counter()
ctx.set_mesh(cfg["vx"], cfg["vy"], tri, abc)
out["set_mesh_seconds"] = time.perf_counter() - t
out["grid_overlay"] = ctx.info("grid_overlay")
timed("set_mesh_random_diagonals_recognised")
ctx.set_option("strategy", 2)
timed("set_mesh_random_diagonals_k_pieces", steps=3)
print(json.dumps(out))
ctx.close()

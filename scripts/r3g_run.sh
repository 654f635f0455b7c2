mkdir -p gpurun_out
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) > gpurun_out/r3g_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3g_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r3g_bench.log | head -1
( MA_PART=0,8 timeout 120 python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from mongeampere_b200 import capi, workloads
case = workloads.make_case("c3", 1.0, "zero")
ctx = capi.Context(0); workloads.load_engine(ctx, case); ctx.set_weights(case["w"]); ctx.set_partition(0, 8)
for _ in range(5): ctx.evaluate(True)
ctx.set_profiling(True)
acc = {}
for _ in range(10):
    ctx.evaluate(True)
    for k, v in ctx.timings().items(): acc[k] = acc.get(k, 0) + v / 10
print("tile 0 of 8:", {k: round(v, 4) for k, v in acc.items()})
ctx.set_profiling(False)
ctx.timer_start()
for _ in range(20): ctx.evaluate(True)
print("ms per evaluation (graph):", ctx.timer_stop() / 20)
PY
) 2>&1 | tail -3
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r3g_gpu_tests.log 2>&1; grep -E 'passed|failed' gpurun_out/r3g_gpu_tests.log | tail -2
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -2

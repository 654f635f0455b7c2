mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/qc_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/qc_gpu_tests.log | tail -6
( timeout 200 python - <<'PY'
import sys
sys.path.insert(0, '.')
from mongeampere_b200 import capi, workloads
case = workloads.make_case("c3", 1.0, "zero")
ctx = capi.Context(0); workloads.load_engine(ctx, case); ctx.set_weights(case["w"])
for part in ((0, 1), (0, 8)):
    ctx.set_partition(*part)
    for _ in range(5): ctx.evaluate(True)
    ctx.timer_start()
    for _ in range(40): ctx.evaluate(True)
    a = ctx.timer_stop() / 40
    ctx.timer_start()
    for _ in range(40): ctx.evaluate_async(True)
    b = ctx.timer_stop() / 40
    print("tile %d of %d: blocking %.4f ms, queued %.4f ms per evaluation; nnz %d" % (part[0], part[1], a, b, ctx.info("nnz")))
PY
) 2>&1 | tail -2
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) 2>&1 | grep -o '"value": [0-9.]*' | head -1

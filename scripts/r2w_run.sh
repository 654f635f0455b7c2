mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2w_gpu_tests.log 2>&1; tail -5 gpurun_out/r2w_gpu_tests.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu ) > gpurun_out/r2w_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2w_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2w_bench.log | head -1; grep -o '"newton": {[^}]*}' gpurun_out/r2w_bench.log | cut -c1-300
for o in cg_rtol=1e-10 cg_rtol=1e-9; do
echo "== $o"; ( MA_OPTS=$o timeout 300 python scripts/newton_full.py c3 ) 2>&1 | head -1 | cut -c1-330
( MA_OPTS=$o timeout 300 python scripts/newton_full.py c2 ) 2>&1 | head -1 | cut -c1-330
done

mkdir -p gpurun_out
echo "== default (mb5)"; ( timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2h_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2h_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2h_bench.log | head -1; grep -o '"e2e": {[^}]*}' gpurun_out/r2h_bench.log | cut -c1-120
for mb in 4 6; do
  echo "== mb$mb"; ( MA_B200_LIB=mongeampere_b200/variants/libma_b200_mb$mb.so timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2h_bench_mb$mb.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2h_bench_mb$mb.log; grep -o '"value": [0-9.]*' gpurun_out/r2h_bench_mb$mb.log | head -1
done
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2h_gpu_tests.log 2>&1; tail -5 gpurun_out/r2h_gpu_tests.log
( MA_TRACE=0 timeout 600 python scripts/newton_full.py c3 1.0 3000 ) > gpurun_out/r2h_newton_c3.log 2>&1; tail -1 gpurun_out/r2h_newton_c3.log | cut -c1-300

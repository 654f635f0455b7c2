mkdir -p gpurun_out
( MA_TRACE=1 ZELDOVICH_OUTER=1 timeout 150 python scripts/run_configs.py c5 ) > gpurun_out/r2k_c5_trace.log 2>&1; grep -c 'eval kmax' gpurun_out/r2k_c5_trace.log; tail -12 gpurun_out/r2k_c5_trace.log | cut -c1-400
echo "== default"; ( timeout 300 python bench.py --steps 20 --warmup 5 --no-newton --no-cpu ) > gpurun_out/r2k_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2k_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2k_bench.log | head -1; grep -o '"e2e": {[^}]*' gpurun_out/r2k_bench.log | cut -c1-200; tail -3 gpurun_out/r2k_bench.log | cut -c1-300
echo "== graph=0"; ( timeout 300 python bench.py --steps 20 --warmup 5 --no-newton --no-cpu --opt graph=0 ) > gpurun_out/r2k_bench_nograph.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2k_bench_nograph.log; grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_nograph.log | head -1
for v in qi8 qi3 qi6c12 mb5; do
  echo "== $v"; ( MA_B200_LIB=mongeampere_b200/variants/libma_b200_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-newton --no-cpu ) > gpurun_out/r2k_bench_$v.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2k_bench_$v.log; grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_$v.log | head -1
done
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2k_gpu_tests.log 2>&1; tail -8 gpurun_out/r2k_gpu_tests.log
( time timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2k_newton_c3.log 2>&1; tail -4 gpurun_out/r2k_newton_c3.log | cut -c1-600

mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2i_gpu_tests.log 2>&1; tail -12 gpurun_out/r2i_gpu_tests.log
for mb in 3; do
  echo "== K2B mb$mb"; ( MA_B200_LIB=mongeampere_b200/variants/libma_b200_k2b$mb.so timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2i_bench_k2b$mb.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2i_bench_k2b$mb.log
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_full.log 2>&1; tail -2 gpurun_out/r2i_bench_full.log | cut -c1-200; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2i_bench_full.log; grep -o '"cpu_baseline": {.*' gpurun_out/r2i_bench_full.log | cut -c1-600
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_ref.log 2>&1; tail -3 gpurun_out/r2i_bench_ref.log | cut -c1-400
( time timeout 900 python scripts/run_configs.py c1 c2 c4 c5 ) > gpurun_out/r2i_configs.log 2>&1; tail -2 gpurun_out/r2i_configs.log | cut -c1-2000

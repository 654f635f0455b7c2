mkdir -p gpurun_out
for v in base "" nor2; do
echo "== variant '$v'"; ( MA_B200_LIB=${v:+mongeampere_b200/variants/libma_b200_$v.so} timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) > gpurun_out/r3a_bench_$v.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3a_bench_$v.log; grep -o '"value": [0-9.]*' gpurun_out/r3a_bench_$v.log | head -1
done
( time timeout 300 python scripts/newton_full.py c3 ) 2>&1 | head -1 | cut -c1-330
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r3a_gpu_tests.log 2>&1; tail -4 gpurun_out/r3a_gpu_tests.log

mkdir -p gpurun_out
for v in "" k3mb3 k3mb4 k3mb6; do
echo "== variant '$v'"; ( MA_B200_LIB=${v:+mongeampere_b200/variants/libma_b200_$v.so} timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) > gpurun_out/r3i_bench_$v.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3i_bench_$v.log; grep -o '"value": [0-9.]*' gpurun_out/r3i_bench_$v.log | head -1
done

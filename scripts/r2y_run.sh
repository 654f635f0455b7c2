mkdir -p gpurun_out
for o in "k3_overlap=1" "k3_overlap=1 --opt graph=0" "k3_overlap=0"; do
echo "== $o"; ( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton --opt $o ) > gpurun_out/r2y_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2y_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2y_bench.log | head -1
done

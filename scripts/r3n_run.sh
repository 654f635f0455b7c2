mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r3n_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/r3n_gpu_tests.log | tail -6
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) > gpurun_out/r3n_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3n_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r3n_bench.log | head -1; grep -o '"gpu_launches": [0-9]*' gpurun_out/r3n_bench.log; grep -o '"e2e": {[^}]*}' gpurun_out/r3n_bench.log | cut -c1-150

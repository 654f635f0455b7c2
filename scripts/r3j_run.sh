mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r3j_gpu_tests.log 2>&1; grep -E 'passed|failed|Error|assert' gpurun_out/r3j_gpu_tests.log | tail -6
( timeout 300 python scripts/grid_as_mesh.py 1.0 ) 2>&1 | tail -2 | cut -c1-900
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) > gpurun_out/r3j_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3j_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r3j_bench.log | head -1

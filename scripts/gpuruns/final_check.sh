mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/fc_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/fc_gpu_tests.log | tail -4
( timeout 100 python scripts/run_configs.py c4 ) 2>/dev/null | cut -c1-300
( timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu --no-newton ) 2>&1 | grep -o '"value": [0-9.]*' | head -1

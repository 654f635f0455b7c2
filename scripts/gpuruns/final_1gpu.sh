mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r3e_gpu_tests.log 2>&1; grep -E 'passed|failed' gpurun_out/r3e_gpu_tests.log | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r3e_bench_full.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r3e_bench_full.log; grep -o '"value": [0-9.]*' gpurun_out/r3e_bench_full.log | head -1; grep -o '"newton": {[^}]*}' gpurun_out/r3e_bench_full.log | cut -c1-300; grep -o '"e2e": {[^}]*}' gpurun_out/r3e_bench_full.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r3e_bench_ref.log 2>&1; grep '^{' gpurun_out/r3e_bench_ref.log | cut -c1-300
( time timeout 300 python scripts/run_configs.py c1 c2 c4 ) > gpurun_out/r3e_cfg_c124.log 2>&1; grep '^{' gpurun_out/r3e_cfg_c124.log | cut -c1-1000
( time ZELDOVICH_OUTER=3 timeout 400 python scripts/run_configs.py c5 ) > gpurun_out/r3e_cfg_c5.log 2>&1; grep '^{' gpurun_out/r3e_cfg_c5.log | cut -c1-1200

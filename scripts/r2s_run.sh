mkdir -p gpurun_out
( time ZELDOVICH_OUTER=2 timeout 400 python scripts/run_configs.py c5 ) > gpurun_out/r2s_cfg_c5.log 2>&1; tail -6 gpurun_out/r2s_cfg_c5.log | cut -c1-900
( time timeout 200 python scripts/run_configs.py c1 c2 c4 ) > gpurun_out/r2s_cfg_c124.log 2>&1; tail -2 gpurun_out/r2s_cfg_c124.log | cut -c1-1200
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2s_bench_full.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2s_bench_full.log; grep -o '"value": [0-9.]*' gpurun_out/r2s_bench_full.log | head -1; grep -o '"newton": {[^}]*}' gpurun_out/r2s_bench_full.log | cut -c1-500; grep -o '"roofline": {[^}]*}' gpurun_out/r2s_bench_full.log | cut -c1-400

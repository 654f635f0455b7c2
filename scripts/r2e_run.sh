mkdir -p gpurun_out
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2e_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2e_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2e_bench.log | head -1
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2e_gpu_tests.log 2>&1; tail -12 gpurun_out/r2e_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-newton > gpurun_out/r2e_bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2e_launches.csv 2>/dev/null | head -14
( MA_TRACE=1 timeout 600 python scripts/newton_full.py c3 1.0 3000 ) > gpurun_out/r2e_newton_c3_trace.log 2>&1; grep -c "eval aborted" gpurun_out/r2e_newton_c3_trace.log; grep "ot_solve:" gpurun_out/r2e_newton_c3_trace.log; tail -1 gpurun_out/r2e_newton_c3_trace.log | cut -c1-300

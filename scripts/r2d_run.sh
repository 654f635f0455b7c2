mkdir -p gpurun_out
( timeout 300 python scripts/dbg_amg.py c2 1.0 ) > gpurun_out/r2d_amg_c2.log 2>&1; tail -4 gpurun_out/r2d_amg_c2.log
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2d_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2d_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2d_bench.log | head -1
( time timeout 600 python scripts/newton_full.py c3 1.0 3000 gpurun_out/r2d_c3_w.npy ) > gpurun_out/r2d_newton_c3.log 2>&1; tail -3 gpurun_out/r2d_newton_c3.log
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_gpu_tests.log 2>&1; tail -15 gpurun_out/r2d_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-newton > gpurun_out/r2d_bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2d_launches.csv 2>/dev/null | head -14

mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity_size.py -m gpu -q -x ) > gpurun_out/r2l_gpu_tests_size.log 2>&1; tail -15 gpurun_out/r2l_gpu_tests_size.log
( MA_TRACE=1 timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2l_newton_c3_trace.log 2>&1; grep -c 'warm K2' gpurun_out/r2l_newton_c3_trace.log; grep 'warm K2' gpurun_out/r2l_newton_c3_trace.log | head -30; grep 'eval kmax' gpurun_out/r2l_newton_c3_trace.log | tail -5; tail -2 gpurun_out/r2l_newton_c3_trace.log | cut -c1-600
( time timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2l_newton_c3.log 2>&1; tail -4 gpurun_out/r2l_newton_c3.log | cut -c1-600
( time timeout 100 python scripts/run_configs.py c2 ) > gpurun_out/r2l_cfg_c2.log 2>&1; tail -4 gpurun_out/r2l_cfg_c2.log | cut -c1-600
echo "== kmax=32 graded at 1M (c3, 2 Newton iterations)"; ( MA_OPTS=kmax=32 MA_TRACE=1 timeout 120 python scripts/newton_full.py c3 1.0 2 ) > gpurun_out/r2l_newton_c3_k32.log 2>&1; grep 'eval kmax' gpurun_out/r2l_newton_c3_k32.log | head -8 | cut -c1-250; tail -2 gpurun_out/r2l_newton_c3_k32.log | cut -c1-300
echo "== bench"; ( timeout 300 python bench.py --steps 20 --warmup 5 --no-newton --no-cpu ) > gpurun_out/r2l_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2l_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2l_bench.log | head -1
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2l_gpu_tests.log 2>&1; tail -8 gpurun_out/r2l_gpu_tests.log

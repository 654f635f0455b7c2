mkdir -p gpurun_out
( time timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2o_newton_c3.log 2>&1; head -1 gpurun_out/r2o_newton_c3.log | cut -c1-400
( MA_TRACE=1 timeout 300 python scripts/newton_full.py c3 ) 2>&1 | grep 'ot_solve:' | tail -1
( time timeout 100 python scripts/run_configs.py c2 ) > gpurun_out/r2o_cfg_c2.log 2>&1; grep -o '"c2_newton.*' gpurun_out/r2o_cfg_c2.log | cut -c1-300
( time timeout 900 python -m pytest tests/test_gpu_parity_size.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r2o_gpu_tests.log 2>&1; tail -6 gpurun_out/r2o_gpu_tests.log
echo "== kmax=32 from the start, 1 M, Newton (the stall)"; ( MA_OPTS=kmax=32 MA_TRACE=2 timeout 60 python scripts/newton_full.py c3 1.0 2 ) > gpurun_out/r2o_newton_c3_k32.log 2>&1; grep -v 'quick empty' gpurun_out/r2o_newton_c3_k32.log | tail -6 | cut -c1-250
echo "== same, quick_reject=0"; ( MA_OPTS=kmax=32,quick_reject=0 MA_TRACE=2 timeout 60 python scripts/newton_full.py c3 1.0 2 ) > gpurun_out/r2o_newton_c3_k32_noqr.log 2>&1; tail -6 gpurun_out/r2o_newton_c3_k32_noqr.log | cut -c1-250

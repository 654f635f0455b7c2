mkdir -p gpurun_out
( time timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2j_newton_c3.log 2>&1; tail -4 gpurun_out/r2j_newton_c3.log | cut -c1-600
for c in c1 c2 c4; do
  ( time timeout 200 python scripts/run_configs.py $c ) > gpurun_out/r2j_cfg_$c.log 2>&1; tail -6 gpurun_out/r2j_cfg_$c.log | cut -c1-800
done
( time SCALE=0.25 ZELDOVICH_OUTER=2 timeout 300 python scripts/run_configs.py c5 ) > gpurun_out/r2j_cfg_c5q.log 2>&1; tail -8 gpurun_out/r2j_cfg_c5q.log | cut -c1-800
( time ZELDOVICH_OUTER=2 timeout 500 python scripts/run_configs.py c5 ) > gpurun_out/r2j_cfg_c5.log 2>&1; tail -8 gpurun_out/r2j_cfg_c5.log | cut -c1-800
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j_gpu_tests.log 2>&1; tail -8 gpurun_out/r2j_gpu_tests.log

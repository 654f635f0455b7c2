mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/qc_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/qc_gpu_tests.log | tail -6
( time timeout 300 python scripts/run_configs.py c4 ) > gpurun_out/qc_cfg_c4.log 2>&1; grep '^{' gpurun_out/qc_cfg_c4.log | cut -c1-500

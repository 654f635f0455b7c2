mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r3k_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/r3k_gpu_tests.log | tail -6

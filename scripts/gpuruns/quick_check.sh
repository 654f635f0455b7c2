# the round-end checks on a GPU box: the -m gpu suite (through the C-ABI and the C++ drop-in driver) and smoke()
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/qc_gpu_tests.log 2>&1; grep -E 'passed|failed|^E ' gpurun_out/qc_gpu_tests.log | tail -6
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1

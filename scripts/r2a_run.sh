set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2a_bench.log 2>&1; tail -3 gpurun_out/r2a_bench.log
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu --opt lean=0 ) > gpurun_out/r2a_bench_nolean.log 2>&1; tail -3 gpurun_out/r2a_bench_nolean.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_gpu_tests.log 2>&1; tail -15 gpurun_out/r2a_gpu_tests.log
( time MA_TRACE=0 timeout 900 python scripts/newton_full.py c3 1.0 3000 gpurun_out/r2a_c3_w.npy ) > gpurun_out/r2a_newton_c3.log 2>&1; tail -5 gpurun_out/r2a_newton_c3.log

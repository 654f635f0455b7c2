set -x
mkdir -p gpurun_out
nproc
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu ) > gpurun_out/r2b_bench.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2b_bench.log; grep -o '"value": [0-9.]*' gpurun_out/r2b_bench.log | head -1
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_gpu_tests.log 2>&1; tail -25 gpurun_out/r2b_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-newton > gpurun_out/r2b_bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches.csv | head -30
MA_PROFILER_START=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_cells_block -c 2 -o gpurun_out/r2b_block -f python scripts/prof_eval.py c3 1.0 2 > gpurun_out/r2b_ncu_block.log 2>&1
MA_PROFILER_START=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_seg -c 1 -o gpurun_out/r2b_seg -f python scripts/prof_eval.py c3 1.0 2 > gpurun_out/r2b_ncu_seg.log 2>&1
MA_PROFILER_START=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_cells_persist -c 1 -o gpurun_out/r2b_persist -f python scripts/prof_eval.py c3 1.0 2 > gpurun_out/r2b_ncu_persist.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5

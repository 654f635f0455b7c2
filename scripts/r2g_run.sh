mkdir -p gpurun_out
for opt in "block_target=1.0" "block_target=0.8" "block_target=1.25" "clip_a=1" "clip_a=4" "clip_b=2"; do
  echo "== $opt"; ( timeout 300 python bench.py --steps 10 --warmup 3 --no-newton --no-cpu --opt $opt ) > gpurun_out/r2g_bench_$opt.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/r2g_bench_$opt.log; grep -o '"value": [0-9.]*' gpurun_out/r2g_bench_$opt.log | head -1
done
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2g_gpu_tests.log 2>&1; tail -8 gpurun_out/r2g_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-newton > gpurun_out/r2g_bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2g_launches.csv 2>/dev/null | head -16
MA_PROFILER_START=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_cells_block -c 1 -o gpurun_out/r2g_block -f python scripts/prof_eval.py c3 1.0 2 > gpurun_out/r2g_ncu_block.log 2>&1

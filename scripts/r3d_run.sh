mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r3d_launches_newton_c3.csv python scripts/newton_full.py c3 1.0 2 > gpurun_out/r3d_newton_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r3d_launches_newton_c3.csv | head -12

mkdir -p gpurun_out
( MA_TRACE=2 timeout 60 python scripts/dbg_k32.py 0.2 64 ) > gpurun_out/r2n_k64_200k.log 2>&1; grep -v 'warm K2\|quick empty' gpurun_out/r2n_k64_200k.log | tail -12 | cut -c1-250
( MA_TRACE=2 timeout 60 python scripts/dbg_k32.py 0.2 64 0 ) > gpurun_out/r2n_k64_200k_nopersist.log 2>&1; grep -v 'warm K2\|quick empty' gpurun_out/r2n_k64_200k_nopersist.log | tail -8 | cut -c1-250
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2n_launches_newton.csv python scripts/newton_full.py c3 1.0 2 > gpurun_out/r2n_newton_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2n_launches_newton.csv > gpurun_out/r2n_launches_newton_summary.txt 2>&1; head -40 gpurun_out/r2n_launches_newton_summary.txt

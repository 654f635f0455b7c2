mkdir -p gpurun_out
timeout 600 python scripts/dbg_r2c.py > gpurun_out/r2c_dbg.log 2>&1; tail -60 gpurun_out/r2c_dbg.log

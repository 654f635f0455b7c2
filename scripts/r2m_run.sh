mkdir -p gpurun_out
( timeout 200 python scripts/dbg_k32.py 0.05 ) > gpurun_out/r2m_k32_50k.log 2>&1; cat gpurun_out/r2m_k32_50k.log | cut -c1-300
( timeout 300 python scripts/dbg_k32.py 0.2 ) > gpurun_out/r2m_k32_200k.log 2>&1; cat gpurun_out/r2m_k32_200k.log | cut -c1-300
( time timeout 300 python scripts/newton_full.py c3 ) > gpurun_out/r2m_newton_c3.log 2>&1; tail -4 gpurun_out/r2m_newton_c3.log | cut -c1-600
( time timeout 100 python scripts/run_configs.py c2 ) > gpurun_out/r2m_cfg_c2.log 2>&1; tail -4 gpurun_out/r2m_cfg_c2.log | cut -c1-600

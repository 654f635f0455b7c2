mkdir -p gpurun_out
( MA_OPTS=kmax=32,quick_reject=0 MA_TRACE=2 timeout 45 python scripts/newton_full.py c3 1.0 2 ) > gpurun_out/r2p_newton_c3_k32_noqr.log 2>&1; grep 'stage\|eval' gpurun_out/r2p_newton_c3_k32_noqr.log | head -24 | cut -c1-200

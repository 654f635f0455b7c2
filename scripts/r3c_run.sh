mkdir -p gpurun_out
for v in "" tail16k tail64k; do
echo "== variant '$v'"
( MA_B200_LIB=${v:+mongeampere_b200/variants/libma_b200_$v.so} MA_TRACE=1 timeout 300 python scripts/newton_full.py c3 ) 2>&1 | grep 'ot_solve:\|^{' | cut -c1-330
( MA_B200_LIB=${v:+mongeampere_b200/variants/libma_b200_$v.so} timeout 300 python scripts/newton_full.py c2 ) 2>&1 | head -1 | cut -c1-330
done
( time timeout 900 python -m pytest tests/test_gpu_parity_size.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r3c_gpu_tests.log 2>&1; tail -4 gpurun_out/r3c_gpu_tests.log

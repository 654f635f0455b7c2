mkdir -p gpurun_out
N=${NGPU:-8}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu ) > gpurun_out/r2v_bench_${N}gpu.log 2>&1
grep -o '"stages_ms": {[^}]*}' gpurun_out/r2v_bench_${N}gpu.log; grep -o '"value": [0-9.]*' gpurun_out/r2v_bench_${N}gpu.log | head -1; grep -o '"e2e": {[^}]*}' gpurun_out/r2v_bench_${N}gpu.log | cut -c1-200; grep -o '"newton": {[^}]*}' gpurun_out/r2v_bench_${N}gpu.log | cut -c1-400; tail -3 gpurun_out/r2v_bench_${N}gpu.log | cut -c1-200

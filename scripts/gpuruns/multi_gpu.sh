mkdir -p gpurun_out
N=${NGPU:-2}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu ) > gpurun_out/r3f_bench_${N}gpu.log 2>&1
grep -o '"stages_ms": {[^}]*}' gpurun_out/r3f_bench_${N}gpu.log; grep -o '"value": [0-9.]*' gpurun_out/r3f_bench_${N}gpu.log | head -1; grep -o '"e2e": {[^}]*}' gpurun_out/r3f_bench_${N}gpu.log | cut -c1-120; grep -o '"newton": {[^}]*}' gpurun_out/r3f_bench_${N}gpu.log | cut -c1-200
if [ "$N" = "2" ]; then
for sh in 0 1; do
( NEWTON_SHARDED=$sh timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/dist_newton_nccl.py c2 0.3 ) > gpurun_out/r3f_dist_newton_c2_sh${sh}.log 2>&1; grep '^{' gpurun_out/r3f_dist_newton_c2_sh${sh}.log | cut -c1-700; tail -2 gpurun_out/r3f_dist_newton_c2_sh${sh}.log | cut -c1-200
done
( NEWTON_SHARDED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/dist_newton_nccl.py c3 1.0 ) > gpurun_out/r3f_dist_newton_c3_sh1.log 2>&1; grep '^{' gpurun_out/r3f_dist_newton_c3_sh1.log | cut -c1-700
fi

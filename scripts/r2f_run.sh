mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_newton_nccl.py c2 0.3 ) > gpurun_out/r2f_dist_c2.log 2>&1; tail -6 gpurun_out/r2f_dist_c2.log | cut -c1-900
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu ) > gpurun_out/r2f_bench2.log 2>&1; tail -4 gpurun_out/r2f_bench2.log | cut -c1-3000

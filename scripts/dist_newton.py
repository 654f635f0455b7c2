"""Distributed damped Newton on N GPUs (one process per GPU, NCCL), checked against the single-GPU ma_ot_solve.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_newton.py [workload] [scale]
Every rank evaluates its Morton tile (ma_set_partition), f / g are all-reduced, the Hessian rows gathered, the grounded solve
replicated (mongeampere_b200/distributed.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mongeampere_b200 import capi
from mongeampere_b200.distributed import ContextTile, DistributedKantorovich
from mongeampere_b200 import workloads as common

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
case = common.make_case(name, scale, "zero")
ctx = capi.Context(local)
common.load_engine(ctx, case)
N = case["N"]
nu = np.full(N, ctx.total_mass / N)
ref = None
if rank == 0:  # single-GPU reference, before the context is restricted to a tile
    t = time.perf_counter()
    w0, st0, rc0 = ctx.ot_solve(nu, eps_g=1e-8, maxiter=1000)
    ref = dict(seconds=time.perf_counter() - t, niter=st0["niter"], neval=st0["neval"], rc=rc0)
print(f"[rank {rank}] reference done: {ref}", flush=True)
dk = DistributedKantorovich(ContextTile(ctx, rank, world), device=f"cuda:{local}")
dist.barrier()
print(f"[rank {rank}] barrier passed", flush=True)
t = time.perf_counter()
w, st = dk.ot_solve(nu, eps_g=1e-8, maxiter=1000, verbose=bool(os.environ.get("MA_DIST_VERBOSE")))
dt = time.perf_counter() - t
# every rank must hold the same weights
wt = torch.from_numpy(w).cuda()
lo, hi = wt.clone(), wt.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool((lo == hi).all().item())
if rank == 0:
    print(json.dumps(dict(workload=name, N=N, gpus=world, distributed=dict(seconds=dt, **{k: st[k] for k in ("niter", "neval", "status", "final_norm")}),
                          single_gpu=ref, max_weight_diff=float(np.abs(w - w0).max()), identical_on_all_ranks=same)))
    assert same and st["status"] == "ok" and (st["niter"], st["neval"]) == (ref["niter"], ref["neval"]) and np.abs(w - w0).max() <= 1e-8 * max(np.abs(w0).max(), 1e-300)
ctx.close()
dist.destroy_process_group()

"""Distributed damped Newton with NCCL INSIDE the engine (ma_comm_init + ma_ot_solve on every rank), checked against the
single-GPU solve.  One process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/dist_newton_nccl.py [workload] [scale]
torch.distributed (gloo) is only the out-of-band channel that ships the 128-byte NCCL id to the other ranks."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mongeampere_b200 import capi
from mongeampere_b200 import workloads as common

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
case = common.make_case(name, scale, "zero")
N = case["N"]
ref = None
if rank == 0:  # single-GPU reference on its own context
    c0 = capi.Context(local)
    common.load_engine(c0, case)
    nu0 = np.full(N, c0.total_mass / N)
    t = time.perf_counter()
    w0, st0, rc0 = c0.ot_solve(nu0, eps_g=1e-7, maxiter=3000)
    ref = dict(seconds=time.perf_counter() - t, niter=st0["niter"], neval=st0["neval"], cg_iters=st0["cg_iters"], rc=rc0)
    f0, g0, H0 = c0.kantorovich(w0)
    c0.close()
ids = [capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx = capi.Context(local)
common.load_engine(ctx, case)
ctx.comm_init(rank, world, ids[0])
sharded = int(os.environ.get("NEWTON_SHARDED", "0"))
ctx.set_option("newton_sharded", sharded)  # 1: evaluations sharded, collectives per trial; 0 (default): the loop replicated
nu = np.full(N, ctx.total_mass / N)
dist.barrier()
t = time.perf_counter()
w, st, rc = ctx.ot_solve(nu, eps_g=1e-7, maxiter=3000)
dt = time.perf_counter() - t
f, g, H = ctx.kantorovich(w)  # with a communicator: the whole problem's g and h on every rank
allw = [None] * world
dist.all_gather_object(allw, w.tobytes())
same = all(b == allw[0] for b in allw)
if rank == 0:
    out = dict(workload=name, scale=scale, N=N, gpus=world, newton_sharded=sharded,
               distributed=dict(seconds=dt, rc=rc, niter=st["niter"], neval=st["neval"], cg_iters=st["cg_iters"], final_norm=st["final_norm"]),
               single_gpu=ref, max_weight_diff=float(np.abs((w - w[-1]) - (w0 - w0[-1])).max()), identical_on_all_ranks=same,
               g_diff=float(np.abs(g - g0).max() / np.abs(g0).max()), H_nnz=(int(H.nnz), int(H0.nnz)),
               H_diff=float(abs(H - H0).max() / np.abs(H0.diagonal()).max()))
    print(json.dumps(out), flush=True)
    assert same and rc == 0 and (st["niter"], st["neval"]) == (ref["niter"], ref["neval"])
    assert out["max_weight_diff"] <= 1e-8 and out["g_diff"] <= 1e-9 and H.nnz == H0.nnz
ctx.close()
dist.destroy_process_group()

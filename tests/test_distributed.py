"""N > 1 host logic on CPU: world_size-2 gloo processes, each owning one tile of the Diracs.  The tile
evaluator here is the ORACLE restricted to a tile (test infrastructure); on GPUs it is a capi.Context
after ma_set_partition (tests/test_gpu_parity.py covers that the tiles of one GPU add up)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleTile:
    def __init__(self, orc, oracle_mod, rank, world):
        self.orc, self.O, self.N = orc, oracle_mod, orc.N
        self.lo, self.hi = self.N * rank // world, self.N * (rank + 1) // world

    def kantorovich(self, w):
        import scipy.sparse as sp
        f, g, H = self.orc.kantorovich(w)
        mask = np.zeros(self.N, bool)
        mask[self.lo:self.hi] = True
        # this tile's share of f = sum over its cells of (w_i m_i - cost_i); the oracle only returns the
        # total, so rank 0 carries f and the others 0 — the all-reduce must still sum to f
        f_part = f if self.lo == 0 else 0.0
        D = sp.diags(mask.astype(np.float64))
        return f_part, np.where(mask, g, 0.0), sp.csr_matrix(D @ H)

    def has_empty_cell(self, w):
        g = self.orc.kantorovich(w)[1]
        return bool((g[self.lo:self.hi] <= 0).any())

    def solve_laplacian_matrix(self, H, g):
        return self.O.solve_laplacian_matrix(H, g)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mongeampere_b200.distributed import DistributedKantorovich
    from oracle import oracle as O
    from tests import common
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = common.make_case("c2", 0.004, "0.3")
        orc = common.oracle_for(O, case)
        dk = DistributedKantorovich(OracleTile(orc, O, rank, world))
        f, g, H = dk.kantorovich(case["w"])
        f0, g0, H0 = orc.kantorovich(case["w"])
        assert abs(f - f0) <= 1e-15 * abs(f0) and np.array_equal(g, g0)
        assert abs(H - H0).max() == 0 and H.nnz == H0.nnz
        nu = np.full(case["N"], g0.sum() / case["N"])
        x, st = dk.ot_solve(nu, eps_g=1e-9)
        x0, st0, _ = O.ot_solve(orc, nu, eps_g=1e-9)
        q.put((rank, st["niter"], st["neval"], st0["niter"], st0["neval"], float(np.abs(x - x0).max()), st["final_norm"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_newton_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ni, ne, ni0, ne0, dx, fn in res:
        assert (ni, ne) == (ni0, ne0)
        assert dx <= 1e-12 and fn < 1e-9
    # both ranks hold the same weights (direction broadcast from rank 0)
    assert res[0][1:] == res[1][1:]


@pytest.mark.gpu
def test_tiles_of_one_gpu_add_up(gpu_ctx):
    """ma_set_partition: the tiles' partial f / g / H rows sum to the unpartitioned evaluation."""
    from tests import common
    case = common.make_case("c3", 0.003, "0.3")
    common.load_engine(gpu_ctx, case)
    f0, g0, H0 = gpu_ctx.kantorovich(case["w"])
    f, g, H = 0.0, np.zeros_like(g0), None
    for r in range(3):
        gpu_ctx.set_partition(r, 3)
        fr, gr, Hr = gpu_ctx.kantorovich(case["w"])
        f += fr; g += gr; H = Hr if H is None else H + Hr
    gpu_ctx.set_partition(0, 1)
    assert abs(f - f0) <= 1e-13 * abs(f0)
    assert np.array_equal(g, g0)
    assert abs(H - H0).max() == 0

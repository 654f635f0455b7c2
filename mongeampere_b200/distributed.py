"""Multi-GPU host side: one process per GPU, Morton tiles of Diracs.

Every Laguerre cell is independent once points, weights and mesh are replicated (kantorovich.hpp:87-136
writes only g[idv], row idv of h and a scalar), so rank r evaluates only its tile
(`ma_set_partition(ctx, r, G)`) and the evaluation itself needs no data-path collective.  What IS
exchanged, through torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests):

  * per evaluation: all-reduce (sum) of f and of g — the tiles' supports are disjoint, so the sum of
    the zero-padded gradients is the gathered gradient — and the rows of H of every tile;
  * per Newton iteration: nothing more — the grounded Laplacian solve is replicated on every rank
    (it is latency-bound, SURVEY.md §7.3-5: a single B200 does one SpMV of the 1 M-Dirac Hessian in
    ~20 µs, less than one NVLink all-gather round trip), and rank 0's direction is broadcast so that
    all ranks hold bit-identical weights.

`TileEvaluator` is anything with `.N`, `.kantorovich(w) -> (f_part, g_part, H_part)` (g_part zero and
H_part empty outside the tile), `.has_empty_cell(w) -> bool` (cheap: neighbour search only) and
`.solve_laplacian_matrix(H, g) -> d`; `mongeampere_b200.capi.Context`
after `set_partition` is one.  The Newton loop restates optimal_transport.hpp:89-193 (same conditions,
same `niter++ <= maxiter` quirk) around the distributed evaluation.
"""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


class DistributedKantorovich:
    def __init__(self, tile, group=None, device=None):
        self.tile = tile
        self.group = group
        self.device = device  # torch device of the collectives' tensors ("cuda:k" for NCCL, None/cpu for gloo)
        d = _dist()
        self.rank = d.get_rank(group) if d else 0
        self.world = d.get_world_size(group) if d else 1
        self.N = tile.N

    # ---- collectives on host arrays ----
    def _allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        d = _dist()
        if d is None or self.world == 1:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        if self.device is not None:
            t = t.to(self.device)
        d.all_reduce(t, op=d.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _allgather_rows(self, H_part):
        """H = the disjoint union of the tiles' rows."""
        import scipy.sparse as sp
        H_part = sp.csr_matrix(H_part)
        d = _dist()
        if d is None or self.world == 1:
            return H_part
        import torch
        dev = self.device
        coo = H_part.tocoo()
        nnz = torch.tensor([coo.nnz], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        d.all_gather(sizes, nnz, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        cap = max(max(sizes), 1)
        pack = torch.zeros(3, cap, dtype=torch.float64, device=dev)  # row, col, value (indices < 2^53 are exact)
        if coo.nnz:
            pack[0, :coo.nnz] = torch.from_numpy(coo.row.astype(np.float64)).to(pack.device)
            pack[1, :coo.nnz] = torch.from_numpy(coo.col.astype(np.float64)).to(pack.device)
            pack[2, :coo.nnz] = torch.from_numpy(coo.data.astype(np.float64)).to(pack.device)
        parts = [torch.empty_like(pack) for _ in range(self.world)]
        d.all_gather(parts, pack, group=self.group)
        rows, cols, vals = [], [], []
        for p, s in zip(parts, sizes):
            p = p.cpu().numpy()
            rows.append(p[0, :s].astype(np.int64)); cols.append(p[1, :s].astype(np.int64)); vals.append(p[2, :s])
        H = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=H_part.shape)
        H.sort_indices()
        return H

    def _bcast(self, a: np.ndarray) -> np.ndarray:
        d = _dist()
        if d is None or self.world == 1:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        if self.device is not None:
            t = t.to(self.device)
        src = d.get_global_rank(self.group, 0) if self.group is not None else 0
        d.broadcast(t, src=src, group=self.group)
        return t.cpu().numpy()

    # ---- the reference's entry points, distributed ----
    def kantorovich(self, w, hessian=True):
        """kantorovich.hpp:35-42 over all tiles -> (fval, g, H | None), identical on every rank."""
        f_part, g_part, H_part = self.tile.kantorovich(w)
        red = self._allreduce_sum(np.concatenate(([f_part], g_part)))
        H = self._allgather_rows(H_part) if hessian else None
        return float(red[0]), red[1:], H

    def ot_solve(self, masses, x=None, eps_g=1e-7, maxiter=100, verbose=False):
        """optimal_transport.hpp:89-193 -> (x, stats).  stats['status'] in ok / empty_cell / not_converged."""
        N = self.N
        masses = np.asarray(masses, np.float64)
        stats = {"niter": 0, "neval": 0, "status": "ok"}

        def f(xx):  # :110-120
            stats["neval"] += 1
            r, g, h = self.kantorovich(xx)
            return r - masses.dot(xx), g.copy(), g - masses, h

        x = np.zeros(N) if x is None or len(x) != N else np.array(x, np.float64)  # :125-128
        fx, m, g, h = f(x)
        eps0 = min(m.min(), masses.min()) / 2  # :137-138
        if eps0 <= 0:  # :139-148
            stats["status"] = "empty_cell"
            return x, stats
        niter = 0
        while np.linalg.norm(g) >= eps_g:  # :150-151, with the niter++ <= maxiter quirk (App. B T6)
            ok = niter <= maxiter
            niter += 1
            if not ok:
                stats["status"] = "not_converged"
                break
            d = self._bcast(-np.asarray(self.tile.solve_laplacian_matrix(h, g)))  # :153, replicated solve
            alpha, x0, n0 = 1.0, x.copy(), np.linalg.norm(g)
            while True:  # :163-176
                x = x0 + alpha * d
                # a trial point that hides a Dirac anywhere has min m = 0 < eps0 and is rejected whatever the
                # rest of the evaluation says (:167): find that out from the neighbour search alone, on every
                # tile, before paying for the integrals of a wildly distorted diagram
                if self._allreduce_sum(np.array([1.0 if self.tile.has_empty_cell(x) else 0.0]))[0] > 0:
                    stats["neval"] += 1  # the reference evaluates it (App. B T5)
                    m, g = np.zeros(N), np.full(N, np.inf)
                else:
                    fx, m, g, h = f(x)
                if m.min() >= eps0 and np.linalg.norm(g) <= (1 - alpha / 2) * n0:
                    break
                alpha *= 0.5
                if alpha < 1e-30:
                    stats["status"] = "linesearch_failed"
                    break
            if verbose and self.rank == 0:
                print(f"it {niter}: f={fx} |df|={np.linalg.norm(g)} tau = {alpha} eval = {stats['neval']}")
            if stats["status"] != "ok":
                break
        stats["niter"] = niter
        stats["final_norm"] = float(np.linalg.norm(g))
        stats["fval"] = fx
        return x, stats


class ContextTile:
    """A capi.Context restricted to this rank's Morton tile, as a TileEvaluator."""

    def __init__(self, ctx, rank, world):
        self.ctx = ctx
        ctx.set_partition(rank, world)
        self.N = ctx.N

    def kantorovich(self, w):
        return self.ctx.kantorovich(w)

    def has_empty_cell(self, w):
        return self.ctx.has_empty_cell(w)

    def solve_laplacian_matrix(self, H, g):
        d, _ = self.ctx.solve_laplacian_matrix(H, g)
        return d

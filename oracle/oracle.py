"""ctypes front-end of the CPU oracle (oracle/ma_oracle.cpp) + the reference's Newton loop.

TEST INFRASTRUCTURE ONLY (parity unpinned, see ma_oracle.cpp header): importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the
product package.

`ot_solve` below restates /root/reference/include/MA/optimal_transport.hpp:89-193 (control flow,
including the `niter++ <= maxiter` quirk, App. B T6) around the oracle's kantorovich; the grounded
linear solve (optimal_transport.hpp:41-87, Eigen::SimplicialLLT — un-vendored Eigen 3.2.1) is done
with SciPy's direct sparse factorisation of the same (N-1)x(N-1) block.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libma_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

MODE_BRUTE = 1      # O(N^2) neighbour search
MODE_PER_CELL = 2   # per-cell OpenMP enumeration instead of the reference's global BFS
MODE_RECORD = 4     # keep the pieces


def _cpu_key() -> str:
    """Identifies the host CPU's instruction set: the library is built -march=native, so a copy
    built on another machine (it travels to the GPU box with the snapshot) is rebuilt there."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile the oracle with g++ -O3 -march=native (no external dependencies)."""
    src = os.path.join(_HERE, "ma_oracle.cpp")
    stamp = _LIB_PATH + ".cpu"
    key = _cpu_key()
    try:
        same_cpu = open(stamp).read().strip() == key
    except OSError:
        same_cpu = False
    if (not force and same_cpu and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    tmp = _LIB_PATH + f".{os.getpid()}.tmp"
    cmd = ["g++", "-O3", "-march=native", "-std=c++17", "-fopenmp", "-shared", "-fPIC", src, "-o", tmp, "-lquadmath"]
    subprocess.check_call(cmd)
    os.replace(tmp, _LIB_PATH)  # atomic: concurrent ranks / xdist workers may build at the same time
    with open(stamp, "w") as fh:
        fh.write(key)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()  # no-op when the library is current for this source and this CPU
    L = C.CDLL(_LIB_PATH)
    L.mao_create.restype = C.c_void_p
    L.mao_destroy.argtypes = [C.c_void_p]
    L.mao_linear_functions.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, _ip, _dp]
    L.mao_set_mesh.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _ip, _dp]
    L.mao_set_points.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.mao_set_cell_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mao_kantorovich.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int]
    L.mao_moments.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, _dp]
    L.mao_fval.argtypes = [C.c_void_p]
    L.mao_fval.restype = C.c_double
    L.mao_get_g.argtypes = [C.c_void_p, _dp]
    L.mao_nnz.argtypes = [C.c_void_p]
    L.mao_get_csr.argtypes = [C.c_void_p, _ip, _ip, _dp]
    L.mao_num_neighbors.argtypes = [C.c_void_p]
    L.mao_get_neighbors.argtypes = [C.c_void_p, _ip, _ip]
    L.mao_get_counters.argtypes = [C.c_void_p, _lp]
    L.mao_num_pieces.argtypes = [C.c_void_p]
    L.mao_num_piece_vertices.argtypes = [C.c_void_p]
    L.mao_get_pieces.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _dp]
    L.mao_solve_laplacian.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, C.c_int]
    _lib = L
    return L


def linear_functions(vx, vy, rho, tri):
    """Per-face (a, b, c) of the PL density (functions.hpp:55-80)."""
    vx = np.ascontiguousarray(vx, np.float64)
    vy = np.ascontiguousarray(vy, np.float64)
    rho = np.ascontiguousarray(rho, np.float64)
    tri = np.ascontiguousarray(tri, np.int32).reshape(-1, 3)
    abc = np.empty((tri.shape[0], 3), np.float64)
    lib().mao_linear_functions(len(vx), vx, vy, rho, tri.shape[0], tri.reshape(-1), abc.reshape(-1))
    return abc


COUNTER_NAMES = ("pieces", "piece_vertices", "new_vertices", "laguerre_edges", "sum_k", "sum_k_np",
                 "fallbacks", "ties")


def algorithmic_flops(c: dict) -> float:
    """F_alg of SURVEY.md §8(d) from the piece combinatorics."""
    return (13.0 * c["sum_k"] + 4.0 * c["sum_k_np"] + 11.0 * c["new_vertices"]
            + 153.0 * (c["piece_vertices"] - 2 * c["pieces"]) + 24.0 * c["laguerre_edges"])


class Oracle:
    def __init__(self, vx, vy, tri, abc, nthreads: int = 1):
        self.L = lib()
        self.h = C.c_void_p(self.L.mao_create())
        self.vx = np.ascontiguousarray(vx, np.float64)
        self.vy = np.ascontiguousarray(vy, np.float64)
        self.tri = np.ascontiguousarray(tri, np.int32).reshape(-1, 3)
        self.abc = np.ascontiguousarray(abc, np.float64).reshape(-1, 3)
        rc = self.L.mao_set_mesh(self.h, len(self.vx), self.vx, self.vy, self.tri.shape[0],
                                 self.tri.reshape(-1), self.abc.reshape(-1))
        if rc != 0:
            raise ValueError("mesh faces must be counter-clockwise")
        self.N = 0
        self.nthreads = nthreads

    def __del__(self):
        try:
            self.L.mao_destroy(self.h)
        except Exception:
            pass

    def set_points(self, X):
        X = np.asarray(X, np.float64)
        self.x = np.ascontiguousarray(X[:, 0])
        self.y = np.ascontiguousarray(X[:, 1])
        self.N = len(self.x)
        self.L.mao_set_points(self.h, self.N, self.x, self.y)

    def set_cell_range(self, lo=0, hi=-1):
        """bench.py's CPU legs: evaluate only the cells [lo, hi) of this problem (per-cell mode); hi < 0 = all."""
        self.L.mao_set_cell_range(self.h, int(lo), int(hi))

    def default_mode(self):
        return MODE_BRUTE if self.N <= 2000 else 0

    def kantorovich(self, w, mode=None):
        """-> (fval, g, H) with H a scipy.sparse.csr_matrix (kantorovich.hpp:37-141)."""
        import scipy.sparse as sp
        if mode is None:
            mode = self.default_mode()
        w = np.ascontiguousarray(w, np.float64)
        assert w.shape == (self.N,)
        self.L.mao_kantorovich(self.h, w, mode, self.nthreads)
        g = np.empty(self.N)
        self.L.mao_get_g(self.h, g)
        nnz = self.L.mao_nnz(self.h)
        ptr = np.empty(self.N + 1, np.int32)
        col = np.empty(max(nnz, 1), np.int32)
        val = np.empty(max(nnz, 1), np.float64)
        self.L.mao_get_csr(self.h, ptr, col, val)
        H = sp.csr_matrix((val[:nnz], col[:nnz], ptr), shape=(self.N, self.N))
        return self.L.mao_fval(self.h), g, H

    def counters(self):
        c = np.zeros(8, np.int64)
        self.L.mao_get_counters(self.h, c)
        return dict(zip(COUNTER_NAMES, (int(v) for v in c)))

    def neighbors(self):
        n = self.L.mao_num_neighbors(self.h)
        ptr = np.empty(self.N + 1, np.int32)
        idx = np.empty(max(n, 1), np.int32)
        self.L.mao_get_neighbors(self.h, ptr, idx)
        return ptr, idx[:n]

    def pieces(self):
        P = self.L.mao_num_pieces(self.h)
        nv = self.L.mao_num_piece_vertices(self.h)
        cell = np.empty(max(P, 1), np.int32)
        face = np.empty(max(P, 1), np.int32)
        ptr = np.zeros(P + 1, np.int32)
        tag = np.empty(max(nv, 1), np.int32)
        xy = np.empty(max(2 * nv, 2), np.float64)
        if P:
            self.L.mao_get_pieces(self.h, cell, face, ptr, tag, xy)
        return cell[:P], face[:P], ptr, tag[:nv], xy[:2 * nv].reshape(-1, 2)

    def moments(self, w, order=1, mode=None):
        """-> N x 6 array (mass, ∫ρx, ∫ρy, ∫ρx², ∫ρy², ∫ρxy) (lloyd.hpp:30-123)."""
        if mode is None:
            mode = self.default_mode()
        w = np.ascontiguousarray(w, np.float64)
        mom = np.zeros((self.N, 6))
        self.L.mao_moments(self.h, w, order, mode, self.nthreads, mom.reshape(-1))
        return mom

    def lloyd(self, w, mode=None):
        """-> (masses, centroids) (lloyd.hpp:126-144)."""
        mom = self.moments(w, 1, mode)
        with np.errstate(divide="ignore", invalid="ignore"):
            return mom[:, 0].copy(), mom[:, 1:3] / mom[:, 0:1]


def solve_laplacian_matrix(H, g, direct=True):
    """optimal_transport.hpp:41-87: ground the last index, solve the leading block, d[N-1] = 0."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    N = H.shape[0]
    diag = H.diagonal()
    if diag.min() == 0:
        print("Error: hessian of Kantorovich's functional is not invertible", file=sys.stderr)
    d = np.zeros(N)
    if N == 1:
        return d
    if direct:
        hs = sp.csc_matrix(H[: N - 1, : N - 1])
        d[: N - 1] = spla.splu(hs).solve(np.asarray(g[: N - 1], np.float64))
    else:
        Hc = sp.csr_matrix(H)
        Hc.sort_indices()
        lib().mao_solve_laplacian(N, Hc.indptr.astype(np.int32), Hc.indices.astype(np.int32),
                                  np.ascontiguousarray(Hc.data), np.ascontiguousarray(g, np.float64), d,
                                  1e-14, 100000)
    return d


def ot_solve(orc: Oracle, masses, x=None, eps_g=1e-7, maxiter=100, verbose=False, mode=None, direct=True):
    """Damped Newton (optimal_transport.hpp:89-193).  Returns (x, stats, trace)."""
    N = orc.N
    masses = np.asarray(masses, np.float64)
    stats = {"niter": 0, "neval": 0, "status": "ok"}
    trace = []

    def f(xx):  # :110-120
        stats["neval"] += 1
        r, g, h = orc.kantorovich(xx, mode)
        m = g.copy()
        return r - masses.dot(xx), m, g - masses, h

    if x is None or len(x) != N:  # :125-128
        x = np.zeros(N)
    x = np.array(x, np.float64)
    fx, m, g, h = f(x)
    eps0 = min(m.min(), masses.min()) / 2  # :137-138
    if eps0 <= 0:  # :139-148
        stats["status"] = "empty_cell"
        stats["empty_cell"] = int(np.argmin(m))
        return x, stats, trace
    niter = 0
    while True:  # while (g.norm() >= eps_g && niter++ <= maxiter)  :150-151
        if not (np.linalg.norm(g) >= eps_g):
            break
        ok = niter <= maxiter
        niter += 1
        if not ok:
            break
        d = -solve_laplacian_matrix(h, g, direct)  # :153
        alpha = 1.0
        x0 = x.copy()
        n0 = np.linalg.norm(g)
        while True:  # :163-176
            x = x0 + alpha * d
            fx, m, g, h = f(x)
            if m.min() >= eps0 and np.linalg.norm(g) <= (1 - alpha / 2) * n0:
                break
            alpha *= 0.5
            if alpha < 1e-20:
                stats["status"] = "linesearch_failed"
                break
        trace.append((niter, fx, float(np.linalg.norm(g)), alpha, stats["neval"]))
        if verbose:
            print(f"it {niter}: f={fx} |df|={np.linalg.norm(g)} tau = {alpha} eval = {stats['neval']}",
                  file=sys.stderr)
        if stats["status"] != "ok":
            break
    stats["niter"] = niter
    return x, stats, trace

"""ctypes front-end of tests/emu/emu_harness.cu (CPU emulation of the CUDA per-lane code; TEST ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "_build", "libma_emu.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "emu_harness.cu")
    deps = [src] + [os.path.join(_ROOT, "mongeampere_b200", "csrc", f) for f in ("ma_cell.cuh", "ma_geom.cuh", "ma_seg.cuh", "ma_block.cuh", "ma_warm.cuh")]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(d) for d in deps):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-extended-lambda", "--expt-relaxed-constexpr",
                           "-Xcompiler", "-fPIC", "-shared", src, "-o", _LIB])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def set_lean(on: bool):
    """K2 through the block kernels' lane code (ma_block.cuh) first, CellSearch for what they cannot certify."""
    lib().emu_set_lean(int(on))


def lean_counts():
    """Cells finished by the radius-2 block, the radius-3 block and CellSearch in the last evaluation."""
    out = (C.c_int * 4)()
    lib().emu_get_lean(out)
    return tuple(out[1:4])


def warm_counts():
    """Warm path of K2 in the last evaluation: (certified, cells rebuilt in round 0, in round 1, failing cells per match x 3)."""
    out = (C.c_int * 6)()
    lib().emu_get_warm(out)
    return tuple(out)


def pack_diag(diag):
    """diag[i, j] in {0, 1} for the real squares (i < n-1, j < m-1): 1 = split along (i+1,j)-(i,j+1).  -> bit array over the
    padded squares i in [-1, n-1], j in [-1, m-1] as ma_seg.cuh reads it (the fictitious layer is 0)."""
    diag = np.asarray(diag)
    nx, ny = diag.shape
    pad = np.zeros((nx + 2, ny + 2), np.uint8)
    pad[1:-1, 1:-1] = diag
    flat = pad.reshape(-1)
    nwords = (len(flat) + 31) // 32
    bits = np.zeros(nwords * 32, np.uint8)
    bits[:len(flat)] = flat
    return np.ascontiguousarray((bits.reshape(nwords, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32))


def evaluate(mesh, X, w, kmax=16, maxv_piece=12, mode=0, filter_tol=1e-11, bin_target=2, nlanes=32, seg=False, seeds=None):
    """mesh: dict(kind='grid', n, m, x0, y0, x1, y1, abc) or dict(kind='mesh', vx, vy, tri, abc).
    seeds: the `seeds` entry of an earlier result for the same X (K2 then takes the warm path, ma_warm.cuh).
    Returns dict(f, g, H, mom, counters, flags, adjacency, seeds)."""
    X = np.asarray(X, np.float64)
    N = len(X)
    x = np.ascontiguousarray(X[:, 0]); y = np.ascontiguousarray(X[:, 1])
    w = np.ascontiguousarray(w, np.float64)
    abc = np.ascontiguousarray(mesh["abc"], np.float64).reshape(-1)
    perm = np.zeros(N, np.int32); mass = np.zeros(N); fcell = np.zeros(N)
    nbr_cnt = np.zeros(N, np.int32); nbr = np.zeros(N * kmax, np.int32); hslot = np.zeros(N * kmax)
    touched = np.zeros(N, np.uint64); mom = np.zeros(N * 6); counters = np.zeros(8, np.int64)
    flags = C.c_int(0)
    if mesh["kind"] == "grid":
        n, m = mesh["n"], mesh["m"]
        x0, y0, x1, y1 = mesh.get("x0", -1.0), mesh.get("y0", -1.0), mesh.get("x1", 1.0), mesh.get("y1", 1.0)
        gargs = (2, n, m, C.c_double(x0), C.c_double(y0), C.c_double((x1 - x0) / (n - 1)), C.c_double((y1 - y0) / (m - 1)),
                 0, None, None, 0, None)
    else:
        vx = np.ascontiguousarray(mesh["vx"], np.float64); vy = np.ascontiguousarray(mesh["vy"], np.float64)
        tri = np.ascontiguousarray(mesh["tri"], np.int32).reshape(-1)
        gargs = (1, 0, 0, C.c_double(0), C.c_double(0), C.c_double(1), C.c_double(1),
                 len(vx), vx.ctypes.data_as(C.c_void_p), vy.ctypes.data_as(C.c_void_p), len(tri) // 3,
                 tri.ctypes.data_as(C.c_void_p))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rho = np.ascontiguousarray(mesh["rho"], np.float64).reshape(-1) if mesh.get("rho") is not None else None
    if seg and mesh["kind"] == "grid":
        assert rho is not None, "the segment path reads the vertex densities"
    dbits = None
    if seg and mesh["kind"] == "grid" and mesh.get("diag") is not None:
        dbits = pack_diag(mesh["diag"])
        lib().emu_set_diag(p(dbits))
    else:
        lib().emu_set_diag(None)
    if seeds is not None:
        assert kmax == 16
        snbr = np.ascontiguousarray(seeds[0], np.int32); scnt = np.ascontiguousarray(seeds[1], np.int32)
        lib().emu_set_seeds(p(snbr), p(scnt))
    else:
        lib().emu_set_seeds(None, None)
    rc = lib().emu_eval(*gargs, p(abc), p(rho) if rho is not None else None, N, p(x), p(y), p(w), kmax, maxv_piece, mode, C.c_double(filter_tol),
                        bin_target, nlanes, (int(seg) if mesh["kind"] == "grid" else 0), p(perm), p(mass), p(fcell), p(nbr_cnt), p(nbr), p(hslot), p(touched),
                        p(mom), p(counters), C.byref(flags))
    assert rc == 0
    lib().emu_set_seeds(None, None)
    lib().emu_set_diag(None)
    raw = dict(mass=mass.copy(), nbr=nbr.copy(), nbr_cnt=nbr_cnt.copy(), hslot=hslot.copy(), fcell=fcell.copy())
    g = np.zeros(N); g[perm] = mass
    momc = np.zeros((N, 6)); momc[perm] = mom.reshape(N, 6)
    # CSR from slots (what k_csr_fill does)
    rows, cols, vals = [], [], []
    nbr = nbr.reshape(N, kmax); hslot = hslot.reshape(N, kmax)
    adj = [None] * N
    for k in range(N):
        i = perm[k]
        adj[i] = sorted(int(perm[j]) for j in nbr[k, :max(nbr_cnt[k], 0)])
        t = int(touched[k])
        if not t:
            continue
        d = 0.0
        for s in range(kmax):
            if (t >> s) & 1:
                rows.append(i); cols.append(perm[nbr[k, s]]); vals.append(-hslot[k, s]); d += hslot[k, s]
        rows.append(i); cols.append(i); vals.append(d)
    H = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    names = ("pieces", "piece_vertices", "new_vertices", "laguerre_edges", "sum_k", "sum_k_np", "fallbacks", "candidates")
    return dict(f=float(fcell.sum()), g=g, H=H, mom=momc, counters=dict(zip(names, map(int, counters))),
                flags=flags.value, adjacency=adj, seeds=(raw["nbr"].reshape(N, kmax), raw["nbr_cnt"]), raw=raw)

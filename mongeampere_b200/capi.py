"""ctypes binding of libma_b200.so (include/ma_b200.h) and a thin Python mirror of the reference's
entry points (kantorovich / ot_solve / lloyd / first_moment / second_moment /
voronoi_triangulation_intersection).

There is deliberately NO fallback: if the shared library is missing, or no CUDA device is
present, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MA_OK, MA_EMPTY_CELL, MA_SINGULAR_HESSIAN, MA_LINSOLVE_RESIDUAL, MA_CUDA_ERROR, MA_INVALID, MA_NOT_CONVERGED = range(7)
STATUS_NAMES = ("MA_OK", "MA_EMPTY_CELL", "MA_SINGULAR_HESSIAN", "MA_LINSOLVE_RESIDUAL", "MA_CUDA_ERROR",
                "MA_INVALID", "MA_NOT_CONVERGED")
TIMING_NAMES = ("total", "prep", "cells", "pieces", "csr", "reduce", "_6", "_7")
COUNTER_NAMES = ("pieces", "piece_vertices", "new_vertices", "laguerre_edges", "sum_k", "sum_k_np", "fallbacks",
                 "candidates")

# every symbol include/ma_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = (
    "ma_create", "ma_destroy", "ma_last_error", "ma_abi_version", "ma_set_mesh", "ma_set_mesh_pl", "ma_set_grid",
    "ma_set_image", "ma_set_points", "ma_kantorovich", "ma_get_hessian_csr", "ma_moments", "ma_lloyd",
    "ma_solve_laplacian", "ma_ot_solve", "ma_pieces_build", "ma_pieces_get", "ma_cells_build", "ma_cells_get",
    "ma_set_weights", "ma_evaluate", "ma_evaluate_async", "ma_sync",
    "ma_get_adjacency", "ma_set_profiling", "ma_get_timings", "ma_set_stats", "ma_get_counters", "ma_flush_l2",
    "ma_measure_fp64_peak", "ma_set_option", "ma_get_info", "ma_set_partition", "ma_timer_start", "ma_timer_stop",
    "ma_comm_unique_id", "ma_comm_init", "ma_comm_destroy", "ma_get_tile_rows", "ma_draw_laguerre_diagram",
)


class MAError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}: {message}")
        self.status = status


class Statistics(C.Structure):  # struct ma_statistics
    _fields_ = [("niter", C.c_size_t), ("neval", C.c_size_t), ("cg_iters", C.c_size_t),
                ("final_norm", C.c_double), ("fval", C.c_double), ("seconds", C.c_double)]


_lib = None


def load_library(path: str | None = None):
    """Loads (never builds) the CUDA library; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("MA_B200_LIB") or _build.LIB_PATH  # MA_B200_LIB: experiments with build variants
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.ma_create.argtypes = [C.POINTER(vp), C.c_int]
    L.ma_destroy.argtypes = [vp]
    L.ma_destroy.restype = None
    L.ma_last_error.argtypes = [vp]
    L.ma_last_error.restype = C.c_char_p
    L.ma_set_mesh.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
    L.ma_set_mesh_pl.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, dp]
    L.ma_set_grid.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, vp, dp]
    L.ma_set_image.argtypes = [vp, C.c_int, C.c_int, vp, dp]
    L.ma_set_points.argtypes = [vp, C.c_int, vp, vp]
    L.ma_kantorovich.argtypes = [vp, vp, dp, vp, ip]
    L.ma_get_hessian_csr.argtypes = [vp, vp, vp, vp]
    L.ma_moments.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.ma_lloyd.argtypes = [vp, vp, vp, vp]
    L.ma_solve_laplacian.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, ip]
    L.ma_ot_solve.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_size_t, C.c_int, C.POINTER(Statistics)]
    L.ma_pieces_build.argtypes = [vp, vp, ip, ip]
    L.ma_pieces_get.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ma_cells_build.argtypes = [vp, vp, ip]
    L.ma_cells_get.argtypes = [vp, vp, vp, vp]
    L.ma_set_weights.argtypes = [vp, vp]
    L.ma_evaluate.argtypes = [vp, C.c_int]
    L.ma_evaluate_async.argtypes = [vp, C.c_int]
    L.ma_sync.argtypes = [vp]
    L.ma_get_adjacency.argtypes = [vp, vp, vp, C.c_int]
    L.ma_set_profiling.argtypes = [vp, C.c_int]
    L.ma_get_timings.argtypes = [vp, vp]
    L.ma_set_stats.argtypes = [vp, C.c_int]
    L.ma_get_counters.argtypes = [vp, vp]
    L.ma_flush_l2.argtypes = [vp, C.c_size_t]
    L.ma_measure_fp64_peak.argtypes = [vp, dp]
    L.ma_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.ma_get_info.argtypes = [vp, C.c_char_p]
    L.ma_get_info.restype = C.c_double
    L.ma_set_partition.argtypes = [vp, C.c_int, C.c_int]
    L.ma_timer_start.argtypes = [vp]
    L.ma_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.ma_get_tile_rows.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp]
    L.ma_draw_laguerre_diagram.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vp]
    L.ma_comm_unique_id.argtypes = [vp]
    L.ma_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.ma_comm_destroy.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


class Context:
    """One engine context on one GPU (wraps ma_ctx)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.ma_create(C.byref(self.h), device)
        if rc != MA_OK:
            msg = self.L.ma_last_error(self.h).decode() if self.h else "ma_create failed"
            if self.h:
                self.L.ma_destroy(self.h)
                self.h = None
            raise MAError(rc, msg)
        self.N = 0
        self.total_mass = None

    def close(self):
        if getattr(self, "h", None):
            self.L.ma_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow=()):
        if rc != MA_OK and rc not in allow:
            raise MAError(rc, self.L.ma_last_error(self.h).decode())
        return rc

    # ---- inputs ----
    def set_mesh(self, vx, vy, tri, abc):
        vx, vy, abc = _f64(vx), _f64(vy), _f64(abc).reshape(-1)
        tri = np.ascontiguousarray(tri, np.int32).reshape(-1)
        self._ck(self.L.ma_set_mesh(self.h, len(vx), _ptr(vx), _ptr(vy), len(tri) // 3, _ptr(tri), _ptr(abc)))
        # sum_f area_f * rho_f(centroid_f) (functions.hpp:117), so that total_mass never describes an earlier mesh
        t = tri.reshape(-1, 3)
        a = abc.reshape(-1, 3)
        ax, ay, bx, by, cx, cy = vx[t[:, 0]], vy[t[:, 0]], vx[t[:, 1]], vy[t[:, 1]], vx[t[:, 2]], vy[t[:, 2]]
        area = ((bx - ax) * (cy - ay) - (cx - ax) * (by - ay)) / 2
        self.total_mass = float(np.sum(area * (a[:, 0] * (ax + bx + cx) / 3 + a[:, 1] * (ay + by + cy) / 3 + a[:, 2])))

    def set_mesh_pl(self, vx, vy, rho, tri):
        vx, vy, rho = _f64(vx), _f64(vy), _f64(rho)
        tri = np.ascontiguousarray(tri, np.int32).reshape(-1)
        tm = C.c_double()
        self._ck(self.L.ma_set_mesh_pl(self.h, len(vx), _ptr(vx), _ptr(vy), _ptr(rho), len(tri) // 3, _ptr(tri),
                                       C.byref(tm)))
        self.total_mass = tm.value
        return tm.value

    def set_grid(self, n, m, rho_v, x0=-1.0, y0=-1.0, x1=1.0, y1=1.0):
        rho_v = _f64(rho_v).reshape(-1)
        assert rho_v.size == n * m
        tm = C.c_double()
        self._ck(self.L.ma_set_grid(self.h, n, m, x0, y0, x1, y1, _ptr(rho_v), C.byref(tm)))
        self.total_mass = tm.value
        return tm.value

    def set_image(self, image):
        """image[i, j] = CImg image(i, j) -> image_to_pl_function (functions.hpp:82-120)."""
        image = np.asarray(image, np.float64)
        n, m = image.shape
        pix = np.ascontiguousarray(image.T)  # pixels[j*n + i]
        tm = C.c_double()
        self._ck(self.L.ma_set_image(self.h, n, m, _ptr(pix), C.byref(tm)))
        self.total_mass = tm.value
        return tm.value

    def set_points(self, X):
        X = np.asarray(X, np.float64)
        x, y = _f64(X[:, 0]), _f64(X[:, 1])
        self.N = len(x)
        self._ck(self.L.ma_set_points(self.h, self.N, _ptr(x), _ptr(y)))

    # ---- the reference's entry points ----
    def kantorovich(self, w, hessian=True):
        """kantorovich.hpp:35-42 -> (fval, g, H csr_matrix | None)."""
        import scipy.sparse as sp
        w = _f64(w)
        assert w.shape == (self.N,)
        f, nnz = C.c_double(), C.c_int()
        g = np.empty(self.N)
        self._ck(self.L.ma_kantorovich(self.h, _ptr(w), C.byref(f), _ptr(g), C.byref(nnz)))
        if not hessian:
            return f.value, g, None
        ptr = np.empty(self.N + 1, np.int32)
        col = np.empty(max(nnz.value, 1), np.int32)
        val = np.empty(max(nnz.value, 1), np.float64)
        self._ck(self.L.ma_get_hessian_csr(self.h, _ptr(ptr), _ptr(col), _ptr(val)))
        H = sp.csr_matrix((val[:nnz.value], col[:nnz.value], ptr), shape=(self.N, self.N))
        return f.value, g, H

    def kantorovich_into(self, w, g, rowptr, col, val):
        """Same call with caller-owned (e.g. pinned) host buffers; col/val must hold >= nnz entries.
        -> (fval, nnz)."""
        f, nnz = C.c_double(), C.c_int()
        self._ck(self.L.ma_kantorovich(self.h, _ptr(w), C.byref(f), _ptr(g), C.byref(nnz)))
        if nnz.value > len(col):
            raise MAError(MA_INVALID, f"Hessian has {nnz.value} entries, buffers hold {len(col)}")
        self._ck(self.L.ma_get_hessian_csr(self.h, _ptr(rowptr), _ptr(col), _ptr(val)))
        return f.value, nnz.value

    def moments(self, w, order=1):
        """first_moment / second_moment (lloyd.hpp:30-123) -> masses, m1 (N,2)[, m2 (N,3)]."""
        w = _f64(w)
        masses = np.empty(self.N)
        m1 = np.empty((2, self.N))
        m2 = np.empty((3, self.N)) if order == 2 else None
        self._ck(self.L.ma_moments(self.h, _ptr(w), order, _ptr(masses), _ptr(m1), _ptr(m2)))
        return (masses, m1.T.copy()) if order == 1 else (masses, m1.T.copy(), m2.T.copy())

    def lloyd(self, w=None):
        """lloyd (lloyd.hpp:126-144) -> masses, centroids (N,2)."""
        w = np.zeros(self.N) if w is None else _f64(w)
        masses = np.empty(self.N)
        cen = np.empty((2, self.N))
        self._ck(self.L.ma_lloyd(self.h, _ptr(w), _ptr(masses), _ptr(cen)))
        return masses, cen.T.copy()

    def solve_laplacian_matrix(self, H, g):
        """solve_laplacian_matrix (optimal_transport.hpp:41-87) -> d, iterations."""
        import scipy.sparse as sp
        H = sp.csr_matrix(H)
        H.sort_indices()
        N = H.shape[0]
        ptr = np.ascontiguousarray(H.indptr, np.int32)
        col = np.ascontiguousarray(H.indices, np.int32)
        val = _f64(H.data)
        g = _f64(g)
        d = np.empty(N)
        it = C.c_int()
        self._ck(self.L.ma_solve_laplacian(self.h, N, _ptr(ptr), _ptr(col), _ptr(val), _ptr(g), _ptr(d), C.byref(it)),
                 allow=(MA_SINGULAR_HESSIAN, MA_LINSOLVE_RESIDUAL))
        return d, it.value

    def ot_solve(self, masses, x=None, eps_g=1e-7, maxiter=100, verbose=False):
        """ot_solve (optimal_transport.hpp:89-193) -> (x, stats dict, status)."""
        nu = _f64(masses)
        have = x is not None and len(x) == self.N
        w = _f64(x).copy() if have else np.zeros(self.N)
        st = Statistics()
        rc = self.L.ma_ot_solve(self.h, _ptr(nu), _ptr(w), int(have), eps_g, maxiter, int(verbose), C.byref(st))
        self._ck(rc, allow=(MA_EMPTY_CELL, MA_NOT_CONVERGED, MA_SINGULAR_HESSIAN))
        stats = dict(niter=st.niter, neval=st.neval, cg_iters=st.cg_iters, final_norm=st.final_norm, fval=st.fval,
                     seconds=st.seconds)
        return w, stats, rc

    def pieces(self, w):
        """voronoi_triangulation_intersection (vti.hpp:315-343) -> cell, face, ptr, tag, xy."""
        w = _f64(w)
        np_, nv = C.c_int(), C.c_int()
        self._ck(self.L.ma_pieces_build(self.h, _ptr(w), C.byref(np_), C.byref(nv)))
        P, V = np_.value, nv.value
        cell = np.empty(max(P, 1), np.int32)
        face = np.empty(max(P, 1), np.int32)
        ptr = np.zeros(P + 1, np.int32)
        tag = np.empty(max(V, 1), np.int32)
        xy = np.empty((max(V, 1), 2))
        self._ck(self.L.ma_pieces_get(self.h, _ptr(cell), _ptr(face), _ptr(ptr), _ptr(tag), _ptr(xy)))
        return cell[:P], face[:P], ptr, tag[:V], xy[:V]

    def cells(self, w=None):
        """Laguerre cells clipped to the mesh bounding box (voronoi_polygon_intersection.hpp:153-188 with
        P = the box) -> ptr (N+1), xy (nv, 2), tag (nv): neighbour index or -1..-4 for the box sides."""
        w = np.zeros(self.N) if w is None else _f64(w)
        nv = C.c_int()
        self._ck(self.L.ma_cells_build(self.h, _ptr(w), C.byref(nv)))
        V = nv.value
        ptr = np.zeros(self.N + 1, np.int32)
        xy = np.empty((max(V, 1), 2))
        tag = np.empty(max(V, 1), np.int32)
        self._ck(self.L.ma_cells_get(self.h, _ptr(ptr), _ptr(xy), _ptr(tag)))
        return ptr, xy[:V], tag[:V]

    def draw_laguerre_diagram(self, w, colors, box, width, height):
        """draw_laguerre_diagram (rasterization.hpp:512-547) -> image[height, width] (row y, column x)."""
        w, colors = _f64(w), _f64(colors)
        img = np.zeros((height, width))
        self._ck(self.L.ma_draw_laguerre_diagram(self.h, _ptr(w), _ptr(colors), box[0], box[1], box[2], box[3], width, height, _ptr(img)))
        return img

    def has_empty_cell(self, w):
        """True iff some Dirac of this context's tile has an empty Laguerre cell at weights w (K1 + K2 only, stops at
        the first one): the line search rejects such a trial point (optimal_transport.hpp:167)."""
        w = _f64(w)
        nv = C.c_int()
        self.set_option("abort_on_empty", 1)
        try:
            self._ck(self.L.ma_cells_build(self.h, _ptr(w), C.byref(nv)))
            return bool(self.info("aborted"))
        finally:
            self.set_option("abort_on_empty", 0)

    # ---- device-resident path / instrumentation ----
    def set_weights(self, w):
        w = _f64(w)
        self._ck(self.L.ma_set_weights(self.h, _ptr(w)))

    def evaluate(self, hessian=True):
        self._ck(self.L.ma_evaluate(self.h, int(hessian)))

    def evaluate_async(self, hessian=True):
        """Queues the evaluation and returns; the next call on the context (or sync()) completes it."""
        self._ck(self.L.ma_evaluate_async(self.h, int(hessian)))

    def sync(self):
        self._ck(self.L.ma_sync(self.h))

    def set_partition(self, rank, nranks):
        self._ck(self.L.ma_set_partition(self.h, rank, nranks))

    def tile_sizes(self):
        nt, nz = C.c_int(), C.c_int()
        self._ck(self.L.ma_get_tile_rows(self.h, C.byref(nt), C.byref(nz), None, None, None, None, None))
        return nt.value, nz.value

    def tile_rows_into(self, row_ids, g, rowptr, col, val):
        """ma_get_tile_rows into caller-owned (e.g. pinned) buffers -> (ntile, nnz_tile)."""
        nt, nz = C.c_int(), C.c_int()
        self._ck(self.L.ma_get_tile_rows(self.h, C.byref(nt), C.byref(nz), _ptr(row_ids), _ptr(g), _ptr(rowptr), _ptr(col), _ptr(val)))
        return nt.value, nz.value

    def tile_rows(self):
        """-> row_ids, g, H_tile (csr_matrix of shape (ntile, N), caller column indices)."""
        import scipy.sparse as sp
        nt, nz = self.tile_sizes()
        ids = np.empty(nt, np.int32); g = np.empty(nt); ptr = np.empty(nt + 1, np.int32)
        col = np.empty(max(nz, 1), np.int32); val = np.empty(max(nz, 1))
        self.tile_rows_into(ids, g, ptr, col, val)
        return ids, g, sp.csr_matrix((val[:nz], col[:nz], ptr), shape=(nt, self.N))

    def comm_init(self, rank, nranks, unique_id: bytes):
        """NCCL communicator of the ranks sharing this problem (ma_comm_init); unique_id from comm_unique_id() on one rank."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.L.ma_comm_init(self.h, rank, nranks, buf))

    def comm_destroy(self):
        self._ck(self.L.ma_comm_destroy(self.h))

    def timer_start(self):
        self._ck(self.L.ma_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.ma_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def adjacency(self):
        ptr = np.empty(self.N + 1, np.int32)
        self._ck(self.L.ma_get_adjacency(self.h, _ptr(ptr), None, 0))
        idx = np.empty(max(int(ptr[-1]), 1), np.int32)
        self._ck(self.L.ma_get_adjacency(self.h, _ptr(ptr), _ptr(idx), int(ptr[-1])))
        return ptr, idx[: ptr[-1]]

    def set_profiling(self, on=True):
        self._ck(self.L.ma_set_profiling(self.h, int(on)))

    def timings(self):
        t = np.zeros(8, np.float32)
        self._ck(self.L.ma_get_timings(self.h, _ptr(t)))
        return dict(zip(TIMING_NAMES, (float(v) for v in t)))

    def set_stats(self, on=True):
        self._ck(self.L.ma_set_stats(self.h, int(on)))

    def counters(self):
        c = np.zeros(8, np.int64)
        self._ck(self.L.ma_get_counters(self.h, _ptr(c)))
        return dict(zip(COUNTER_NAMES, (int(v) for v in c)))

    def flush_l2(self, nbytes=256 << 20):
        self._ck(self.L.ma_flush_l2(self.h, nbytes))

    def fp64_peak(self):
        v = C.c_double()
        self._ck(self.L.ma_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def set_option(self, name, value):
        self._ck(self.L.ma_set_option(self.h, name.encode(), float(value)))

    def info(self, name):
        return self.L.ma_get_info(self.h, name.encode())


def comm_unique_id() -> bytes:
    """128 bytes identifying a new NCCL communicator (ncclGetUniqueId): create on one rank, ship to the others."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.ma_comm_unique_id(buf) != MA_OK:
        raise MAError(MA_CUDA_ERROR, "libnccl.so.2 could not be loaded")
    return buf.raw


def algorithmic_flops(c: dict) -> float:
    """F_alg of SURVEY.md §8(d) from the piece combinatorics of one evaluation."""
    return (13.0 * c["sum_k"] + 4.0 * c["sum_k_np"] + 11.0 * c["new_vertices"]
            + 153.0 * (c["piece_vertices"] - 2 * c["pieces"]) + 24.0 * c["laguerre_edges"])

/* ma_b200.h — C-ABI of libma_b200.so, the CUDA (sm_100a) engine behind the MongeAmpere++ API.
 *
 * The reference has no FFI: its API is the header templates in include/MA.
 * The drop-in headers in this repo's include/MA/ keep those templates' names and argument order
 * and forward to the entry points below.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions: plain pointers and sizes only; host buffers are caller-owned, device buffers are
 * library-owned; every call blocks until its results are in the caller's buffers; one context
 * per host thread and per GPU.  All reals are IEEE fp64, all indices int32.
 * Every function returns an MA_* status; ma_last_error() gives the message for the last failure.
 * There is no CPU fallback: without a CUDA device ma_create() fails with MA_CUDA_ERROR.
 */
#ifndef MA_B200_H
#define MA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ma_ctx ma_ctx;

enum {
  MA_OK = 0,
  MA_EMPTY_CELL = 1,        /* optimal_transport.hpp:139-148: a Laguerre cell is empty at the initial point */
  MA_SINGULAR_HESSIAN = 2,  /* optimal_transport.hpp:49-58: zero on the Hessian diagonal */
  MA_LINSOLVE_RESIDUAL = 3, /* optimal_transport.hpp:68-71: linear solve residual above 1e-7 */
  MA_CUDA_ERROR = 4,
  MA_INVALID = 5,           /* bad argument / call order (the reference's shape asserts, kantorovich.hpp:60-62) */
  MA_NOT_CONVERGED = 6      /* Newton loop hit maxiter (optimal_transport.hpp:150-151) */
};

/* ---- context ---------------------------------------------------------------------------- */
int ma_create(ma_ctx **out, int device);
void ma_destroy(ma_ctx *ctx);
const char *ma_last_error(const ma_ctx *ctx);
/* ABI version, bumped on any signature change. */
int ma_abi_version(void);   /* 3: + ma_evaluate_async / ma_sync;  2: ma_comm_*, options "lean" / "block_target" / "amg*" */

/* ---- source density: a triangulation with one linear function per face --------------------
 * Replaces the (T densityT, Functions densityF) pair of kantorovich.hpp:37-39 / lloyd.hpp:31-33:
 * T is given as vertices + CCW index triples (the input format of
 * include/CGAL/Triangulation_incremental_builder_2.h:25-81, tests/test_triangulation.cpp:12-34),
 * densityF as abc[3 f + (0,1,2)] with rho_f(x,y) = a x + b y + c (functions.hpp:55-80).
 * A triangulation that is a regular grid with each square split along one of its diagonals (the Delaunay triangulation of
 * an image's pixel grid) and whose face functions agree at the shared vertices is recognised: the integrating calls then
 * run on the grid kernel (ma_get_info "grid_overlay" = 1; option "detect_grid" = 0 turns the recognition off). */
int ma_set_mesh(ma_ctx *ctx, int nV, const double *vx, const double *vy, int nF, const int *tri,
                const double *abc);
/* Same, building the per-face functions from per-vertex values as MA::Linear_function's
 * constructor does (functions.hpp:64-71).  total_mass (may be NULL) receives
 * sum_f area_f * rho_f(centroid_f) (functions.hpp:117). */
int ma_set_mesh_pl(ma_ctx *ctx, int nV, const double *vx, const double *vy, const double *rho_v, int nF,
                   const int *tri, double *total_mass);
/* Structured n x m vertex grid on [x0,x1] x [y0,y1]: vertex (i,j) has index i*m+j and density
 * rho_v[i*m+j]; square (i,j) is split along (i,j)-(i+1,j+1) into faces 2*(i*(m-1)+j) and +1.
 * Replaces the triangulation built by image_to_pl_function (functions.hpp:82-120) with a fixed
 * diagonal rule (CGAL's choice is not reproducible). */
int ma_set_grid(ma_ctx *ctx, int n, int m, double x0, double y0, double x1, double y1, const double *rho_v,
                double *total_mass);
/* image_to_pl_function (functions.hpp:82-120): pixels[j*n + i] = image(i,j) (CImg layout), mapped to
 * [-1,1]^2 with rho = image(i, m-j-1)/255 + 1e-3. */
int ma_set_image(ma_ctx *ctx, int n, int m, const double *pixels, double *total_mass);

/* ---- Diracs ------------------------------------------------------------------------------
 * X of kantorovich.hpp:40 / optimal_transport.hpp:93: an Eigen column-major N x 2 matrix maps to
 * (x = X.data(), y = X.data() + N) without a copy.  Sorts the points into Morton-ordered bins. */
int ma_set_points(ma_ctx *ctx, int N, const double *x, const double *y);

/* ---- one evaluation of Kantorovich's functional -------------------------------------------
 * double kantorovich(densityT, densityF, X, weights, g, h)   kantorovich.hpp:35-42.
 * g[N] receives the cell masses (= gradient); the Hessian is kept in the context: *nnz receives its
 * number of stored entries and ma_get_hessian_csr copies it out as CSR in the caller's ordering
 * (rowptr[N+1], col[nnz] ascending within a row, val[nnz]).  Row i is what the reference's triplets
 * (i, *) sum to (kantorovich.hpp:120-121), i.e. the CSR is the row-major form of the reference's h.
 * g, nnz may be NULL. */
int ma_kantorovich(ma_ctx *ctx, const double *weights, double *fval, double *g, int *nnz);
int ma_get_hessian_csr(ma_ctx *ctx, int *rowptr, int *col, double *val);

/* first_moment / second_moment / lloyd   lloyd.hpp:30-144.
 * order 1: masses[N], m1[2N] = (∫ρx [N], ∫ρy [N]) (column-major N x 2 like Eigen).
 * order 2: additionally m2[3N] = (∫ρx², ∫ρy², ∫ρxy) (column-major N x 3).  m2 may be NULL for order 1. */
int ma_moments(ma_ctx *ctx, const double *weights, int order, double *masses, double *m1, double *m2);
/* lloyd(): centroids = m1 / masses (lloyd.hpp:139-143); centroids is column-major N x 2. */
int ma_lloyd(ma_ctx *ctx, const double *weights, double *masses, double *centroids);

/* ---- Newton solver -------------------------------------------------------------------------
 * Vector solve_laplacian_matrix(h, g)   optimal_transport.hpp:41-87: grounds the LAST index, solves
 * the leading (N-1)x(N-1) block of the CSR matrix by Jacobi-preconditioned CG on the device,
 * d[N-1] = 0.  iters (may be NULL) receives the CG iteration count. */
int ma_solve_laplacian(ma_ctx *ctx, int N, const int *rowptr, const int *col, const double *val, const double *g,
                       double *d, int *iters);

typedef struct ma_statistics { /* struct Statistics, optimal_transport.hpp:35-39 (+ extras) */
  size_t niter;
  size_t neval;
  size_t cg_iters;     /* total CG iterations */
  double final_norm;   /* ||m - nu||_2 at exit */
  double fval;         /* f(w) - nu.w at exit */
  double seconds;      /* wall time of the call */
} ma_statistics;

/* void ot_solve(t, functions, X, masses, x, eps_g, maxiter, verbose, stats)  optimal_transport.hpp:89-99.
 * nu[N]: target masses; w[N]: in = initial guess (used iff have_initial != 0, else zeros, :125-128),
 * out = result.  Returns MA_EMPTY_CELL and leaves w = initial guess when a cell is empty at the
 * start (:139-148). */
int ma_ot_solve(ma_ctx *ctx, const double *nu, double *w, int have_initial, double eps_g, size_t maxiter,
                int verbose, ma_statistics *stats);

/* ---- pieces (cell ∩ face polygons) ---------------------------------------------------------
 * voronoi_triangulation_intersection(t, dt, out)  voronoi_triangulation_intersection.hpp:315-343:
 * enumerates every non-empty piece.  Two-call protocol: ma_pieces_build computes them on the device
 * and returns the counts, ma_pieces_get copies them out: piece p belongs to cell[p] and face[p] and
 * has vertices xy[2*ptr[p] .. 2*ptr[p+1]) (CCW); tag[k] names the edge starting at vertex k: the Laguerre
 * neighbour across it (>= 0: an EDGE_DT of vti.hpp:54-96), or -1 / -2 / -3 for the edge (0,1) / (1,2) / (2,0)
 * of the source triangle tri[3 f ..] (an EDGE_T). */
int ma_pieces_build(ma_ctx *ctx, const double *weights, int *npieces, int *nvertices);
int ma_pieces_get(ma_ctx *ctx, int *cell, int *face, int *ptr, int *tag, double *xy);

/* draw_laguerre_diagram(densityT, densityF, X, weights, colors, x0, y0, x1, y1, w, h, put_pixel)  rasterization.hpp:512-547:
 * every piece (cell ∩ face) drawn into a w x h image over the box [x0,x1] x [y0,y1] with exact pixel coverage;
 * image[y * w + x] = sum over the pieces of coverage(piece, pixel) * (mean density at the piece's vertices) * colors[cell]
 * (what the reference's put_pixel callback accumulates).  colors[N] in the caller's ordering; image is overwritten. */
int ma_draw_laguerre_diagram(ma_ctx *ctx, const double *weights, const double *colors, double x0, double y0, double x1, double y1,
                             int w, int h, double *image);

/* ---- Laguerre cells as polygons ---------------------------------------------------------------
 * The power cell of every Dirac clipped to the bounding box of the source mesh, i.e. what
 * voronoi_polygon_intersection(P, dt, v) (voronoi_polygon_intersection.hpp:153-188) returns for P = that box
 * (tests/test_voronoi.cpp:41-48, tests/test_power.cpp:43-50 measure these cells); a convex P is then one
 * host-side clip away (include/MA/voronoi_polygon_intersection.hpp).  Two calls: ma_cells_build runs the
 * neighbour search (K1 + K2) and returns the total vertex count, ma_cells_get copies out, in the caller's
 * ordering: cell i has vertices xy[2*ptr[i] .. 2*ptr[i+1]) (CCW, empty for a hidden Dirac) and tag[k] names
 * the line supporting the edge that STARTS at vertex k: the Laguerre neighbour's index, or -1 / -2 / -3 / -4
 * for the bottom / right / top / left side of the box. */
int ma_cells_build(ma_ctx *ctx, const double *weights, int *nvertices);
/* With ma_set_option(ctx, "abort_on_empty", 1) ma_cells_build stops as soon as a cell of this context's tile is found
 * empty and ma_get_info(ctx, "aborted") returns 1 (no polygons then): the cheap "does this trial point hide a Dirac?"
 * test of the line search (optimal_transport.hpp:167), used by the multi-GPU Newton loop before it pays for a full
 * evaluation. */
int ma_cells_get(ma_ctx *ctx, int *ptr /* N+1 */, double *xy /* 2*nvertices */, int *tag /* nvertices */);

/* ---- device-resident evaluation (what bench.py times as `value`) -----------------------------
 * ma_set_weights uploads w (caller order); ma_evaluate runs K1(per-eval part)+K2+K3+K4 on the
 * device-resident state and leaves masses / Hessian on the device (internal Morton order). */
int ma_set_weights(ma_ctx *ctx, const double *weights);
int ma_evaluate(ma_ctx *ctx, int with_hessian);
/* The same without waiting for the device: the evaluation is queued on the context's stream and the call returns; the next
 * call on the context other than ma_set_weights / ma_evaluate_async (any call; ma_sync does nothing else) waits for it,
 * reads its scalars and, in the rare case that it needs a larger capacity class, repeats it.  A further ma_evaluate_async
 * is queued behind the one in flight and supersedes it (the results live in the same device buffers).  For back-to-back evaluations (a parameter sweep, a benchmark loop): the
 * host round trip between two evaluations — 30 us alone, several times that with 8 processes on one box — leaves the
 * pipeline.  Evaluations that need the host in the middle (line-search trials, warm path, communicator, profiling) are
 * carried out synchronously by this call as well. */
int ma_evaluate_async(ma_ctx *ctx, int with_hessian);
int ma_sync(ma_ctx *ctx);

/* Multi-GPU: this context evaluates only the Morton tile `rank` of `nranks` of the Diracs (every cell is
 * independent once points and weights are replicated, kantorovich.hpp:87-136 writes only g[idv] and
 * row idv of h).  Masses and Hessian rows of the other tiles stay zero/empty in this context; fval,
 * mass_sum and mass_min are the tile's partial values, to be combined by the caller (sum, sum, min). */
int ma_set_partition(ma_ctx *ctx, int rank, int nranks);

/* The rows this context owns after ma_evaluate / ma_kantorovich (all rows without a partition), in the engine's
 * internal (Morton) order: row k is Dirac row_ids[k] of the caller's ordering, g[k] its mass, and
 * col[rowptr[k] .. rowptr[k+1]) / val its Hessian row with CALLER column indices (ordered along the Morton curve, not
 * ascending).  What a multi-GPU caller reads back per rank: the transfer shrinks with the tile instead of staying at N.
 * Call with all array pointers NULL to get the sizes. */
int ma_get_tile_rows(ma_ctx *ctx, int *ntile, int *nnz_tile, int *row_ids, double *g, int *rowptr, int *col, double *val);

/* Multi-GPU with NCCL inside the engine (one process per GPU).  ma_comm_unique_id fills 128
 * bytes on ONE rank (ncclGetUniqueId); the caller ships them to the other ranks by any means (MPI, a file,
 * torch.distributed, a socket) and every rank calls ma_comm_init, which creates the NCCL communicator on the context's
 * device and makes the context evaluate Morton tile `rank` of `nranks` (as ma_set_partition does).  With a communicator
 *   - ma_kantorovich returns the WHOLE problem's f, g and Hessian on every rank (tiles evaluated in parallel, slices
 *     gathered over NVLink);
 *   - ma_ot_solve runs the damped Newton loop of optimal_transport.hpp:89-193 collectively: evaluations are sharded,
 *     per trial point 6 integers and 3 scalars are all-reduced, per accepted point the gradient / Hessian slices are
 *     gathered and the grounded solve is replicated; all ranks return bit-identical weights;
 *   - ma_evaluate keeps its results distributed (no collective except the flag / scalar all-reduces).
 * All ranks must make the same sequence of calls.  libnccl.so.2 is loaded at the first ma_comm_* call; without it
 * these functions return MA_CUDA_ERROR and everything else works. */
int ma_comm_unique_id(void *id128);
int ma_comm_init(ma_ctx *ctx, int rank, int nranks, const void *id128);
int ma_comm_destroy(ma_ctx *ctx);

/* Laguerre adjacency found by the last evaluation (caller ordering, CSR): the neighbours whose
 * bisector supports an edge of (cell ∩ mesh bounding box). */
int ma_get_adjacency(ma_ctx *ctx, int *ptr /* N+1 */, int *idx /* >= ptr[N] */, int capacity);

/* ---- instrumentation ------------------------------------------------------------------------ */
enum {
  MA_T_TOTAL = 0,   /* whole evaluation */
  MA_T_PREP = 1,    /* K1 per-eval: weight gather + max-weight pyramid */
  MA_T_CELLS = 2,   /* K2: Laguerre-cell construction / neighbour search */
  MA_T_PIECES = 3,  /* K3: clipping + integration */
  MA_T_CSR = 4,     /* K4: scan + CSR fill */
  MA_T_REDUCE = 5,  /* scalar reductions */
  MA_T_COUNT = 8
};
/* CUDA-event times (ms) of the stages of the last ma_evaluate (profiling must be enabled). */
int ma_set_profiling(ma_ctx *ctx, int on);
int ma_get_timings(ma_ctx *ctx, float *ms /* MA_T_COUNT */);
/* CUDA-event stopwatch on the context's own stream (torch.cuda.Event would not see it). */
int ma_timer_start(ma_ctx *ctx);
int ma_timer_stop(ma_ctx *ctx, float *ms);
/* Counters of the last evaluation when stats are enabled: pieces, piece vertices, new vertices,
 * Laguerre edges, sum k_i, sum k_i*n_p, robust-predicate fallbacks, candidate triangles. */
int ma_set_stats(ma_ctx *ctx, int on);
int ma_get_counters(ma_ctx *ctx, int64_t *c /* 8 */);
/* Writes `bytes` of device memory (to evict L2 between timed steps). */
int ma_flush_l2(ma_ctx *ctx, size_t bytes);
/* DFMA throughput micro-benchmark (FLOP/s), the fp64 roofline denominator "of measured". */
int ma_measure_fp64_peak(ma_ctx *ctx, double *flops_per_s);
/* Option knobs (ma_set_option; defaults in parentheses):
 *   "kmax" (16)            capacity class every evaluation starts from: 16 (packed-order polygons), 32, 64; escalated
 *                          automatically when a cell has more vertices
 *   "strategy" (0)         0: grid meshes use the boundary-segment kernel k_seg, general meshes k_pieces; 2: k_pieces always
 *   "bin_target" (1)       average Diracs per leaf bin (upper bound), takes effect at the next ma_set_points
 *   "rmax" (6)             rings of leaf bins K2 walks before it turns to the quadtree
 *   "rtree" (4)            rings it walks in any case, even when the ring certificate cannot succeed (weights with a gradient)
 *   "lean" (1)             K2 on weights within a few bin areas runs the lock-step block kernels (k_cells_block 5x5 -> 7x7 ->
 *                          warp-per-cell 11x11 -> CellSearch for the rest); 0: CellSearch for every cell
 *   "block_target" (1)     average Diracs per bin of the block grid, takes effect at the next ma_set_points
 *   "persist" (1)          CellSearch with persistent lanes (k_cells_persist) or one cell per lane (0)
 *   "persist_waves", "persist_min_chunk", "clip_a", "clip_b", "refill_at"   scheduling of k_cells_persist
 *   "cg_rtol" (1e-10), "cg_maxit" (200000)   PCG stopping rule, relative to |g|
 *   "amg" (1), "amg_omega" (0.7), "amg_alpha" (1.5)   quadtree-aggregation multigrid preconditioner of the Newton solves
 *                          (0: Jacobi), its smoother damping and the scaling of the coarse correction
 *   "quick_reject" (1)     ma_ot_solve tests a trial point for an empty cell against the adjacency of the last accepted
 *                          point before it evaluates it (same outcome, optimal_transport.hpp:167)
 *   "warm" (1)             K2 inside ma_ot_solve starts from the adjacency of the last accepted point and certifies the cells by
 *                          ring matching (ma_warm.cuh); 2: every evaluation seeds the next one of the same point set; 0: off
 *   "newton_sharded" (0)   ma_ot_solve on a communicator: 1 = evaluations sharded over the ranks (collectives per trial),
 *                          0 = every rank runs the whole loop (faster at <= a few million Diracs; same bits either way)
 *   "graph" (1)            evaluations are replayed as CUDA graphs;  "k3_overlap" (0)  K3 of the cells the first block kernel
 *                          certifies runs on a side stream under the tail of K2 (measured: no gain)
 *   "filter_tol" (1e-11)   relative threshold below which a sign is re-decided in double-double (K2 and k_pieces);
 *                          1e300 sends every decision through the exact stage
 *   "abort_on_empty" (0)   see ma_cells_build
 * Read-outs (ma_get_info): "N", "nF", "nnz", "kmax", "levels", "mesh_kind", "strategy", "sm_count", "fval", "mass_sum",
 *   "mass_min", "cg_iters", "launches", "cell_lo", "cell_hi", "aborted", "cell_fallbacks" (sign decisions of the last
 *   evaluation's K2 that went through the exact stage), "warm_evals" / "warm_rebuilt" / "warm_failed" (evaluations certified
 *   by the warm path, cells CellSearch rebuilt in them, attempts that had to be redone without seeds), "cg_rtol". */
int ma_set_option(ma_ctx *ctx, const char *name, double value);
double ma_get_info(ma_ctx *ctx, const char *name);

#ifdef __cplusplus
}
#endif
#endif /* MA_B200_H */

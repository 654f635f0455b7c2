// ma_b200.cu — host side of libma_b200.so: context, device buffers, kernel orchestration and the
// C-ABI declared in include/ma_b200.h.  No CPU fallback anywhere: every entry point needs a CUDA
// device and fails with MA_CUDA_ERROR otherwise.
#include "../../include/ma_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl is loaded at run time (ma_comm_init), single-GPU use never needs it

#include "ma_pcg.cuh"

using namespace ma;

namespace {

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct ma_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // copy stream of ma_get_hessian_csr
  cudaStream_t side = nullptr;      // evaluation: work that is off the critical path (supporting planes, scalar reductions)
  cudaEvent_t ev_fork[6] = {};
  std::function<int()> k3_early;  // set by the evaluation: launches k_seg's first pass (launch_cells_lean calls it after the first block kernel)
  bool k3_split = false;          // ... and it did: K3 proper then only takes the rest
  bool planes_pending = false;      // the side stream is building the supporting planes: join before CellSearch runs
  int k3_overlap = 0;               // 1: K3 of the cells the first block kernel certifies runs on the side stream under K2's tail.
                                    // Measured (profiles/r02y): K2 + K3 2.18 ms against 2.23 ms in sequence, and slower when replayed
                                    // as a graph (stream priorities are not captured) — the tail is not idle time, so off by default.
  bool pending = false, pending_hess = false;  // ma_evaluate_async: an evaluation is in flight, its scalars not yet read
  bool want_async = false;
  int use_graph = 1;                // evaluations are replayed as CUDA graphs (captured per distinct launch sequence)
  struct EvalGraph { std::string key; cudaGraphExec_t exec = nullptr; int launches = 0; unsigned long long stamp = 0; };
  std::vector<EvalGraph> graphs;
  std::string graph_candidate;      // key of the last configuration that ran without a graph
  unsigned long long graph_clock = 0;
  cudaEvent_t ev_chunk[8] = {};
  std::string err;
  int sm_count = 148;

  // options
  int kmax = 16;       // current capacity class (16 / 32 / 64 neighbours per cell)
  int kmax_base = 16;  // class every evaluation starts from
  bool capacity_hit = false;  // last evaluation failed because a cell exceeded the largest class
  int bin_target = 1;  // average Diracs per leaf bin (upper bound); ~1 per bin + rings 0..6 measured best (profiles/r01m)
  double cg_rtol = 1e-10;  // relative residual of the Newton solves: below what a sparse Cholesky leaves (cond * eps ~ 1e-10 on these
                           // Hessians); 1e-12, 1e-10 and 1e-9 give the same Newton trajectories on c2 / c3 (profiles/r02w)
  int cg_maxit = 200000;
  double filter_tol = 1e-11;
  int profiling = 0, stats = 0, trace = 0;
  int clip_a = 2, clip_b = 1, refill_at = 8, rmax = 6, rtree = 4;
  int persist = 1, persist_waves = 3, persist_min_chunk = 32;  // K2 with persistent lanes (k_cells_persist)
  int strategy = 0;  // 0 auto (grid mesh: fused segment kernel, general mesh: pieces), 2: pieces always
  int part_rank = 0, part_n = 1;  // Morton tile of the Diracs this context evaluates
  void *comm = nullptr;           // ncclComm_t of the ranks that share the problem (ma_comm_init), else null
  Buf dist_buf, rowptr_g, col_g, val_g;  // collectives' staging, the gathered (global) Hessian of the distributed Newton loop
  int nnz_g = 0;
  long long launches = 0;         // kernels launched by this context (bench.py's gpu_launches)
  long long cell_fallbacks = 0;   // K2: vertices whose in/out sign went to the exact stage in the last evaluation

  // mesh
  int mesh_kind = MESH_NONE;
  int nV = 0, nF = 0;
  double bb[4] = {0, 0, 0, 0};
  Buf vx, vy, tri, abc, rho_v, rho_p, tbin_ptr, tbin_face;
  int tg = 1;
  double tinvx = 1, tinvy = 1;
  int gn = 0, gm = 0;
  double gx0 = 0, gy0 = 0, gdx = 1, gdy = 1;

  // points
  int N = 0;
  Buf x, y, xs, ys, perm, pos, code, bin_count, bin_start, bin_rm, wmax;
  Buf code_s, pre0, pre1, fs_tiles, nodeG, nodeA, wstat;  // per-node supporting planes (ma_geom.cuh)
  Buf xr, yr, wr, rm2s, s2rm, rm_start, blk_cnt;          // the sites in row-major bin order (ma_block.cuh)
  int bG = 1;
  double bph = 1, binv = 1, block_target = 1.0;           // block grid: bG x bG bins of side bph, ~block_target Diracs each
  Buf hard1, hard2, hard3, hard_n;                               // cells the block kernels pass on (lists + 2 counters)
  Buf nbr_prev, cnt_prev;                                 // adjacency at the last accepted Newton point (quick reject of trials)
  int prev_stride = 0;
  Buf diag;                   // general mesh that IS a regular grid (either diagonal per square): the bits k_seg<VD> reads
  bool grid_overlay = false;  // ... found by ma_set_mesh: the integrating modes then run on the grid path
  int detect_grid = 1;
  double mesh_mass = 0.0;  // integral of the density over the mesh (the warm path's sheet count)
  bool in_newton = false;
  int newton_sharded = 0;     // ma_ot_solve on a communicator: 1 = shard the evaluations (collectives per trial), 0 = every rank runs the whole loop
  bool seeds_global = false;  // multi-GPU: nbr_prev holds the rows of ALL tiles (gathered, or left by a warm evaluation)
  bool k2_all = false;        // the last evaluation's K2 covered all cells (warm path on a partitioned context)
  // warm path of K2 (ma_warm.cuh): cells from that adjacency + the ring-match certificate
  Buf ring, ring_n, cstate;
  int warm = 1;            // 0: off; 1: inside ma_ot_solve; 2: every evaluation seeds the next one
  int warm_skip = 0;       // evaluations to run cold after a warm attempt that could not be certified
  int warm_penalty = 4;
  bool warm_now = false;   // the evaluation being enqueued uses the warm path
  long long warm_evals = 0, warm_rebuilt = 0, warm_failed = 0;  // statistics: warm evaluations, cells CellSearch rebuilt in them, attempts redone cold
  int quick_reject = 1;
  int lean = 1;                                           // K2: block kernels first (0: CellSearch for every cell)
  bool abort_on_empty = false, aborted = false;
  bool probe_empty = false;  // option "abort_on_empty": ma_cells_build stops at the first empty cell and reports it
  int L = 0;
  double px0 = 0, py0 = 0, ph = 1;

  // evaluation state
  Buf poly_x, poly_y, poly_t, poly_n;
  Buf w, ws, nbr, nbr_cnt, cell_bb, mass, fcell, hslot, touched, rowcnt, rowptr, col, val, mom;
  Buf scan_tmp, red_partial, red_partial2, red_out, counters, flags, scratch_d, scratch_i;
  Buf cptr, ccol, cval, cg_out;  // caller-order copies
  int nnz = 0;
  bool have_eval = false, have_hessian = false;
  double fval = 0, mass_sum = 0, mass_min = 0;

  // pieces
  Buf pc_count, pc_off, pc_cell, pc_face, pc_ptr, pc_tag, pc_xy;
  int pc_np = 0, pc_nv = 0;
  int cells_nv = -1;  // ma_cells_build

  // pcg
  Buf dinv, cgx, cgr, cgz, cgp0, cgp1, cgq, part_pq, part_rz, part_rr, scal, cgflag;
  int cg_warm = 1;    // Newton: start each PCG from (1 - tau) times the previous direction
  // aggregation multigrid preconditioner (ma_amg.cuh): hierarchy per point set, matrices per solve
  struct AmgHost {
    int nlev = 0, n[AMG_MAX_LEVELS] = {};
    bool ready = false;
    AmgLevel lev[AMG_MAX_LEVELS];
    Buf agg[AMG_MAX_LEVELS], cstart[AMG_MAX_LEVELS], code[AMG_MAX_LEVELS + 1], rowptr[AMG_MAX_LEVELS], col[AMG_MAX_LEVELS],
        val[AMG_MAX_LEVELS], dinv[AMG_MAX_LEVELS], x[AMG_MAX_LEVELS], r[AMG_MAX_LEVELS], t[AMG_MAX_LEVELS], x2[AMG_MAX_LEVELS];
    Buf excl, W, Ainv, flag;
  } amg;
  int amg_on = 1;
  bool amg_stale = true;  // the hierarchy does not belong to the current point set yet
  double amg_omega = 0.7, amg_alpha = 1.5;  // Jacobi damping, over-correction of the flat prolongation (tuned on the c2 Hessian)
  Buf nu_s, x0_s, d_s, g_s;
  size_t last_cg_iters = 0;

  // L2 flush
  Buf flush;
  // pinned host staging for the scalars read back every evaluation (pageable copies are staged and slow)
  struct HostScalars { int flags, abort_, cell_fallbacks, warm_fail, nnz, pad2[3]; double red[8]; int warm[8]; } *hs = nullptr;  // flags..pad mirror the device flags[4]

  // timing
  cudaEvent_t ev[MA_T_COUNT + 2] = {};
  cudaEvent_t ev_user[2] = {};
  float t_ms[MA_T_COUNT] = {};
  int64_t host_counters[CNT_N] = {};
};

namespace {

typedef ma_ctx::AmgHost AmgHost;

int fail(ma_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(c, MA_CUDA_ERROR, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)
#define CKR(call)                     \
  do {                                \
    int r_ = (call);                  \
    if (r_ != MA_OK) return r_;       \
  } while (0)

// (re)allocates; `headroom` > 1 over-allocates when growing, for buffers whose size drifts from call to call
// (cudaFree / cudaMalloc synchronise the device and take milliseconds)
int ensure(ma_ctx *c, Buf &b, size_t bytes, double headroom = 1.0) {
  if (bytes <= b.cap && b.p) return MA_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = std::max<size_t>((size_t)(bytes * headroom), 256);
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return MA_OK;
}
void release(Buf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// exclusive scan of n ints (device) into out[0..n], out[n] = total
int scan_i32(ma_ctx *c, const int *in, int *out, int n) {
  int nt = std::max(1, cdiv(n, SCAN_TILE));
  CKR(ensure(c, c->scan_tmp, (size_t)(nt + 1) * sizeof(int)));
  int *ts = c->scan_tmp.as<int>();
  k_scan_tiles<<<nt, SCAN_NT, 0, c->stream>>>(in, out, ts, n);
  k_scan_sums<<<1, SCAN_NT, 0, c->stream>>>(ts, nt, ts + nt);
  k_scan_add<<<nt, SCAN_NT, 0, c->stream>>>(out, ts, n, ts + nt);
  c->launches += 3;
  CK(cudaGetLastError());
  return MA_OK;
}

// out (device, 4 doubles) = { sum a, sum a*b (a*a if b null), min a, max a }
int reduce4(ma_ctx *c, const double *a, const double *b, int n, double *out_dev) {
  CKR(ensure(c, c->red_partial, (size_t)RED_BLOCKS * 8 * sizeof(double)));
  int nb = std::min(RED_BLOCKS, std::max(1, cdiv(n, RED_NT)));
  k_reduce_stage1<<<nb, RED_NT, 0, c->stream>>>(a, b, n, c->red_partial.as<double>());
  k_reduce_stage2<<<1, RED_NT, 0, c->stream>>>(c->red_partial.as<double>(), nb, out_dev);
  c->launches += 2;
  CK(cudaGetLastError());
  return MA_OK;
}
// the same for two arrays in one pair of launches: out[0..4) from a0, out[4..8) from a1 (same reduction trees as reduce4)
int reduce4x2(ma_ctx *c, const double *a0, const double *a1, int n, double *out_dev, cudaStream_t st) {
  CKR(ensure(c, c->red_partial2, (size_t)RED_BLOCKS * 8 * sizeof(double)));
  int nb = std::min(RED_BLOCKS, std::max(1, cdiv(n, RED_NT)));
  k_reduce2_stage1<<<dim3(nb, 2), RED_NT, 0, st>>>(a0, a1, n, c->red_partial2.as<double>());
  k_reduce2_stage2<<<2, RED_NT, 0, st>>>(c->red_partial2.as<double>(), nb, out_dev);
  c->launches += 2;
  CK(cudaGetLastError());
  return MA_OK;
}

// prefix sums of the moment terms of SET over the Morton-sorted sites into out[K][n+1]
template <int SET> int moment_scan(ma_ctx *c, double *out, PlaneGate gate = PlaneGate{nullptr, 0.0}, cudaStream_t st = nullptr) {
  if (!st) st = c->stream;
  constexpr int K = MomentTerms<SET>::K;
  const int n = c->N, nt = std::max(1, cdiv(n, FS_TILE));
  CKR(ensure(c, c->fs_tiles, (size_t)K * nt * 8));
  const double cx = c->px0 + 0.5 * c->ph * (1 << c->L), cy = c->py0 + 0.5 * c->ph * (1 << c->L);
  k_moment_scan_tiles<SET><<<nt, FS_NT, 0, st>>>(c->xs.as<double>(), c->ys.as<double>(), c->ws.as<double>(), cx, cy,
                                                 n, out, c->fs_tiles.as<double>(), gate);
  k_moment_scan_sums<<<K, FS_NT, 0, st>>>(c->fs_tiles.as<double>(), nt, n, out, gate);
  k_moment_scan_add<<<nt, FS_NT, 0, st>>>(out, c->fs_tiles.as<double>(), n, K, gate);
  c->launches += 3;
  CK(cudaGetLastError());
  return MA_OK;
}

// ---------------------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------------------
double host_total_mass(int nF, const int *tri, const double *vx, const double *vy, const double *abc) {
  double tot = 0;
  for (int f = 0; f < nF; ++f) {
    int a = tri[3 * f], b = tri[3 * f + 1], cc = tri[3 * f + 2];
    double area = ((vx[b] - vx[a]) * (vy[cc] - vy[a]) - (vx[cc] - vx[a]) * (vy[b] - vy[a])) / 2;
    double gx = (vx[a] + vx[b] + vx[cc]) / 3, gy = (vy[a] + vy[b] + vy[cc]) / 3;
    tot += area * (abc[3 * f] * gx + abc[3 * f + 1] * gy + abc[3 * f + 2]);
  }
  return tot;
}

// per-face plane through the three vertex values: what MA::Linear_function represents
// (functions.hpp:55-80), obtained from the 2x2 system instead of three barycentric extrapolations
void host_pl_coefficients(int nF, const int *tri, const double *vx, const double *vy, const double *rho,
                          double *abc) {
  for (int f = 0; f < nF; ++f) {
    int ia = tri[3 * f], ib = tri[3 * f + 1], ic = tri[3 * f + 2];
    double ax = vx[ia], ay = vy[ia], e1x = vx[ib] - ax, e1y = vy[ib] - ay, e2x = vx[ic] - ax, e2y = vy[ic] - ay;
    double f1 = rho[ib] - rho[ia], f2 = rho[ic] - rho[ia];
    double det = e1x * e2y - e2x * e1y;
    double a = (f1 * e2y - f2 * e1y) / det, b = (e1x * f2 - e2x * f1) / det;
    abc[3 * f] = a;
    abc[3 * f + 1] = b;
    abc[3 * f + 2] = rho[ia] - a * ax - b * ay;
  }
}

int upload(ma_ctx *c, Buf &b, const void *src, size_t bytes) {
  CKR(ensure(c, b, bytes));
  if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return MA_OK;
}

void invalidate_eval(ma_ctx *c) {
  c->cells_nv = -1;
  c->have_eval = false;
  c->have_hessian = false;
}

}  // namespace

// =============================================================================================
// context
// =============================================================================================
extern "C" int ma_abi_version(void) { return 3; }

extern "C" int ma_create(ma_ctx **out, int device) {
  if (!out) return MA_INVALID;
  *out = nullptr;
  ma_ctx *c = new ma_ctx();
  c->device = device;
  if (const char *t = getenv("MA_TRACE")) c->trace = atoi(t);
  if (c->trace) c->profiling = 1;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    // no silent CPU path: the context is returned so that the message can be read, but is unusable
    fail(c, MA_CUDA_ERROR, "no usable CUDA device %d (%s); libma_b200 has no CPU fallback", device,
         e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
    *out = c;
    return MA_CUDA_ERROR;
  }
  *out = c;
  CK(cudaSetDevice(device));
  {
    // the side stream carries work that only fills idle SMs (K3 under the tail of K2, the planes, the scalar reductions):
    // lowest priority, so that blocks of the main stream's kernels are scheduled first whenever a slot frees up
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest));
    CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, least));
  }
  for (auto &ev : c->ev_fork) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->ev_chunk) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CK(cudaHostAlloc((void **)&c->hs, sizeof(*c->hs), cudaHostAllocDefault));
  for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
  for (auto &ev : c->ev_user) CK(cudaEventCreate(&ev));
  return MA_OK;
}

extern "C" void ma_destroy(ma_ctx *c) {
  if (!c) return;
  if (c->stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Buf *all[] = {&c->vx, &c->vy, &c->tri, &c->abc, &c->tbin_ptr, &c->tbin_face, &c->x, &c->y, &c->xs, &c->ys,
                  &c->perm, &c->pos, &c->code, &c->bin_count, &c->bin_start, &c->wmax, &c->w, &c->ws, &c->nbr,
                  &c->nbr_cnt, &c->cell_bb, &c->mass, &c->fcell, &c->hslot, &c->touched, &c->rowcnt, &c->rowptr,
                  &c->col, &c->val, &c->mom, &c->scan_tmp, &c->red_partial, &c->red_partial2, &c->red_out, &c->counters, &c->flags,
                  &c->scratch_d, &c->scratch_i, &c->cptr, &c->ccol, &c->cval, &c->cg_out, &c->pc_count, &c->pc_off,
                  &c->pc_cell, &c->pc_face, &c->pc_ptr, &c->pc_tag, &c->pc_xy, &c->dinv, &c->cgx, &c->cgr, &c->cgz,
                  &c->cgp0, &c->cgp1, &c->cgq, &c->part_pq, &c->part_rz, &c->part_rr, &c->scal, &c->cgflag,
                  &c->nu_s, &c->x0_s, &c->d_s, &c->g_s, &c->flush, &c->code_s, &c->pre0, &c->pre1, &c->fs_tiles,
                  &c->nodeG, &c->nodeA, &c->poly_x, &c->poly_y, &c->poly_t, &c->poly_n, &c->wstat, &c->rho_v, &c->rho_p, &c->bin_rm,
                  &c->xr, &c->yr, &c->wr, &c->rm2s, &c->s2rm, &c->rm_start, &c->blk_cnt, &c->diag, &c->nbr_prev, &c->cnt_prev, &c->ring, &c->ring_n, &c->cstate, &c->dist_buf, &c->rowptr_g, &c->col_g, &c->val_g, &c->hard1, &c->hard2, &c->hard3, &c->hard_n};
    for (Buf *b : all) release(*b);
    for (int l = 0; l < AMG_MAX_LEVELS; ++l) {
      Buf *lv[] = {&c->amg.agg[l], &c->amg.cstart[l], &c->amg.code[l], &c->amg.rowptr[l], &c->amg.col[l], &c->amg.val[l],
                   &c->amg.dinv[l], &c->amg.x[l], &c->amg.r[l], &c->amg.t[l], &c->amg.x2[l]};
      for (Buf *b : lv) release(*b);
    }
    release(c->amg.code[AMG_MAX_LEVELS]); release(c->amg.excl); release(c->amg.W); release(c->amg.Ainv); release(c->amg.flag);
    for (auto &ev : c->ev)
      if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev_user)
      if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev_chunk)
      if (ev) cudaEventDestroy(ev);
    ma_comm_destroy(c);
    for (auto &g : c->graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto &ev : c->ev_fork)
      if (ev) cudaEventDestroy(ev);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream);
    if (c->hs) cudaFreeHost(c->hs);
  }
  delete c;
}

extern "C" const char *ma_last_error(const ma_ctx *c) { return c ? c->err.c_str() : "null context"; }

namespace { int finish_pending(ma_ctx *c); }
// (the same without completing an evaluation that is still in flight: calls that only append to the stream)
#define NEED_CTX_NOWAIT()                                                           \
  do {                                                                              \
    if (!c) return MA_INVALID;                                                      \
    if (!c->stream) return fail(c, MA_CUDA_ERROR, "context has no CUDA device");    \
    cudaSetDevice(c->device);                                                       \
  } while (0)
#define NEED_CTX()                                                                  \
  do {                                                                              \
    if (!c) return MA_INVALID;                                                      \
    if (!c->stream) return fail(c, MA_CUDA_ERROR, "context has no CUDA device");    \
    cudaSetDevice(c->device);                                                       \
    if (c->pending) {                                                               \
      const int r_pending_ = finish_pending(c);                                     \
      if (r_pending_ != MA_OK) return r_pending_;                                   \
    }                                                                               \
  } while (0)

extern "C" int ma_set_option(ma_ctx *c, const char *name, double value) {
  if (!c || !name) return MA_INVALID;
  std::string n(name);
  if (n == "kmax") {
    int k = (int)value;
    if (k != 16 && k != 32 && k != 64) return fail(c, MA_INVALID, "kmax must be 16, 32 or 64");
    c->kmax = c->kmax_base = k;
    invalidate_eval(c);
  } else if (n == "bin_target") c->bin_target = std::max(1, (int)value);
  else if (n == "cg_rtol") c->cg_rtol = value;
  else if (n == "cg_maxit") c->cg_maxit = (int)value;
  else if (n == "cg_warm") c->cg_warm = (int)value;
  else if (n == "amg") c->amg_on = (int)value;
  else if (n == "amg_omega") c->amg_omega = value;
  else if (n == "amg_alpha") c->amg_alpha = value;
  else if (n == "filter_tol") c->filter_tol = value;
  else if (n == "persist") c->persist = (int)value;
  else if (n == "lean") c->lean = (int)value;
  else if (n == "graph") c->use_graph = (int)value;
  else if (n == "k3_overlap") c->k3_overlap = (int)value;
  else if (n == "warm") c->warm = (int)value;
  else if (n == "detect_grid") c->detect_grid = (int)value;
  else if (n == "newton_sharded") c->newton_sharded = (int)value;
  else if (n == "quick_reject") c->quick_reject = (int)value;
  else if (n == "block_target") c->block_target = std::max(0.05, value);
  else if (n == "abort_on_empty") c->probe_empty = value != 0;
  else if (n == "rmax") c->rmax = std::min(MA_RING_TABLE_RMAX, std::max(1, (int)value));
  else if (n == "rtree") c->rtree = std::min(MA_RING_TABLE_RMAX, std::max(0, (int)value));
  else if (n == "clip_a") c->clip_a = std::max(1, (int)value);
  else if (n == "clip_b") c->clip_b = std::max(1, (int)value);
  else if (n == "refill_at") c->refill_at = std::min(32, std::max(1, (int)value));
  else if (n == "persist_waves") c->persist_waves = std::max(1, (int)value);
  else if (n == "persist_min_chunk") c->persist_min_chunk = std::max(32, (int)value);
  else if (n == "strategy") {
    int v = (int)value;
    if (v != 0 && v != 2) return fail(c, MA_INVALID, "strategy must be 0 (auto) or 2 (pieces)");
    c->strategy = v;
    invalidate_eval(c);
  }
  else return fail(c, MA_INVALID, "unknown option %s", name);
  return MA_OK;
}

extern "C" double ma_get_info(ma_ctx *c, const char *name) {
  if (c && c->stream && c->pending) { cudaSetDevice(c->device); finish_pending(c); }  // (an evaluation in flight: its scalars first)
  if (!c || !name) return -1;
  std::string n(name);
  if (n == "kmax") return c->kmax;
  if (n == "levels") return c->L;
  if (n == "nnz") return c->nnz;
  if (n == "N") return c->N;
  if (n == "nF") return c->nF;
  if (n == "sm_count") return c->sm_count;
  if (n == "mass_sum") return c->mass_sum;
  if (n == "mass_min") return c->mass_min;
  if (n == "cg_iters") return (double)c->last_cg_iters;
  if (n == "mesh_kind") return c->mesh_kind;
  if (n == "launches") return (double)c->launches;
  if (n == "strategy") return c->strategy;
  if (n == "aborted") return c->aborted ? 1 : 0;
  if (n == "fval") return c->fval;
  if (n == "warm_evals") return (double)c->warm_evals;      // evaluations certified by the warm path of K2 ...
  if (n == "warm_rebuilt") return (double)c->warm_rebuilt;  // ... cells CellSearch rebuilt in them ...
  if (n == "warm_failed") return (double)c->warm_failed;    // ... and attempts that had to be redone cold
  if (n == "cg_rtol") return c->cg_rtol;
  if (n == "grid_overlay") return c->grid_overlay ? 1.0 : 0.0;  // ma_set_mesh recognised a regular grid: K3 runs on k_seg
  if (n == "cell_fallbacks") return (double)c->cell_fallbacks;  // K2 sign decisions that went to the exact stage (last evaluation)
  if (n == "cell_lo") return (double)((long long)c->N * c->part_rank / c->part_n);
  if (n == "cell_hi") return (double)((long long)c->N * (c->part_rank + 1) / c->part_n);
  return -1;
}

// =============================================================================================
// mesh
// =============================================================================================
namespace {
// Is this explicit triangulation a regular n x m grid whose squares are each split along one of their diagonals, with a
// density that is continuous across the faces (the values the face functions take at a shared vertex agree)?  That is what
// image_to_pl_function + a Delaunay triangulation give (functions.hpp:82-120; the diagonal of a square is CGAL's to choose,
// SURVEY App. B T1).  Then: rho[i * m + j] = the vertex densities, bits = one bit per PADDED square (ma_seg.cuh).
bool detect_regular_grid(int nV, const double *vx, const double *vy, int nF, const int *tri, const double *abc, int &n, int &m,
                         double box[4], std::vector<double> &rho, std::vector<unsigned> &bits) {
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
  for (int v = 0; v < nV; ++v) {
    x0 = std::min(x0, vx[v]); x1 = std::max(x1, vx[v]);
    y0 = std::min(y0, vy[v]); y1 = std::max(y1, vy[v]);
  }
  if (!(x1 > x0) || !(y1 > y0)) return false;
  const double tolx = 1e-9 * (x1 - x0), toly = 1e-9 * (y1 - y0);
  long long mcount = 0;
  for (int v = 0; v < nV; ++v) mcount += std::fabs(vx[v] - x0) <= tolx;  // vertices of the first column
  if (mcount < 2 || nV % mcount) return false;
  m = (int)mcount;
  n = nV / m;
  if (n < 2 || (long long)nF != 2ll * (n - 1) * (m - 1)) return false;
  const double dx = (x1 - x0) / (n - 1), dy = (y1 - y0) / (m - 1);
  std::vector<int> vi(nV), vj(nV);
  std::vector<char> seen((size_t)n * m, 0);
  for (int v = 0; v < nV; ++v) {
    const double fi = (vx[v] - x0) / dx, fj = (vy[v] - y0) / dy;
    const long long i = std::llround(fi), j = std::llround(fj);
    if (std::fabs(fi - (double)i) > 1e-7 || std::fabs(fj - (double)j) > 1e-7 || i < 0 || i >= n || j < 0 || j >= m) return false;
    if (seen[(size_t)i * m + j]) return false;
    seen[(size_t)i * m + j] = 1;
    vi[v] = (int)i; vj[v] = (int)j;
  }
  rho.assign((size_t)n * m, 0.0);
  std::vector<char> have((size_t)n * m, 0);
  std::vector<unsigned char> sq((size_t)(n - 1) * (m - 1), 0);  // bits 0-1: halves seen, bit 2: kind, bit 3: kind known
  double scale = 0.0;
  for (int f = 0; f < nF; ++f) {
    const int *t = tri + 3 * f;
    const int i0 = std::min(vi[t[0]], std::min(vi[t[1]], vi[t[2]])), j0 = std::min(vj[t[0]], std::min(vj[t[1]], vj[t[2]]));
    int mask = 0;
    for (int k = 0; k < 3; ++k) {
      const int di = vi[t[k]] - i0, dj = vj[t[k]] - j0;
      if (di > 1 || dj > 1) return false;
      mask |= 1 << (di + 2 * dj);  // corners 00, 10, 01, 11
    }
    if (i0 >= n - 1 || j0 >= m - 1) return false;
    int kind, half;
    switch (mask) {
      case 0b1011: kind = 0; half = 0; break;  // 00 10 11: below the "+" diagonal
      case 0b1101: kind = 0; half = 1; break;  // 00 01 11: above it
      case 0b0111: kind = 1; half = 0; break;  // 00 10 01: below the "-" diagonal
      case 0b1110: kind = 1; half = 1; break;  // 10 01 11: above it
      default: return false;
    }
    unsigned char &q = sq[(size_t)i0 * (m - 1) + j0];
    if ((q & 8) && ((q >> 2) & 1) != kind) return false;
    if (q & (1 << half)) return false;
    q |= (unsigned char)(8 | (kind << 2) | (1 << half));
    for (int k = 0; k < 3; ++k) {  // continuity of the density at the shared vertices
      const int v = t[k];
      const double val = abc[3 * f] * vx[v] + abc[3 * f + 1] * vy[v] + abc[3 * f + 2];
      const size_t id = (size_t)vi[v] * m + vj[v];
      scale = std::max(scale, std::fabs(val));
      if (!have[id]) { have[id] = 1; rho[id] = val; }
      else if (std::fabs(val - rho[id]) > 1e-10 * std::max(scale, 1e-300)) return false;
    }
  }
  for (unsigned char q : sq)
    if ((q & 3) != 3) return false;
  const size_t npad = (size_t)(n + 1) * (m + 1);
  bits.assign((npad + 31) / 32, 0u);
  for (int i = 0; i < n - 1; ++i)
    for (int j = 0; j < m - 1; ++j)
      if ((sq[(size_t)i * (m - 1) + j] >> 2) & 1) {
        const size_t q = (size_t)(i + 1) * (m + 1) + (j + 1);
        bits[q >> 5] |= 1u << (q & 31);
      }
  box[0] = x0; box[1] = y0; box[2] = x1; box[3] = y1;
  return true;
}
}  // namespace

extern "C" int ma_set_mesh(ma_ctx *c, int nV, const double *vx, const double *vy, int nF, const int *tri,
                           const double *abc) {
  NEED_CTX();
  if (nV < 3 || nF < 1 || !vx || !vy || !tri || !abc) return fail(c, MA_INVALID, "ma_set_mesh: bad arguments");
  for (int f = 0; f < nF; ++f) {
    int a = tri[3 * f], b = tri[3 * f + 1], cc = tri[3 * f + 2];
    if (a < 0 || b < 0 || cc < 0 || a >= nV || b >= nV || cc >= nV)
      return fail(c, MA_INVALID, "ma_set_mesh: face %d has an out-of-range vertex", f);
    double area2 = (vx[b] - vx[a]) * (vy[cc] - vy[a]) - (vx[cc] - vx[a]) * (vy[b] - vy[a]);
    if (!(area2 > 0)) return fail(c, MA_INVALID, "ma_set_mesh: face %d is not counter-clockwise", f);
  }
  c->mesh_kind = MESH_GENERAL;
  c->nV = nV;
  c->nF = nF;
  c->bb[0] = c->bb[1] = 1e300;
  c->bb[2] = c->bb[3] = -1e300;
  for (int i = 0; i < nV; ++i) {
    c->bb[0] = std::min(c->bb[0], vx[i]); c->bb[2] = std::max(c->bb[2], vx[i]);
    c->bb[1] = std::min(c->bb[1], vy[i]); c->bb[3] = std::max(c->bb[3], vy[i]);
  }
  // bins of faces (a face is listed in every bin its bounding box overlaps)
  int g = (int)std::ceil(std::sqrt(std::max(1.0, nF / 2.0)));
  g = std::min(g, 2048);
  c->tg = g;
  c->tinvx = g / std::max(c->bb[2] - c->bb[0], 1e-300);
  c->tinvy = g / std::max(c->bb[3] - c->bb[1], 1e-300);
  auto range = [&](int f, int &i0, int &i1, int &j0, int &j1) {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int k = 0; k < 3; ++k) {
      int v = tri[3 * f + k];
      x0 = std::min(x0, vx[v]); x1 = std::max(x1, vx[v]);
      y0 = std::min(y0, vy[v]); y1 = std::max(y1, vy[v]);
    }
    i0 = std::min(std::max((int)std::floor((x0 - c->bb[0]) * c->tinvx), 0), g - 1);
    i1 = std::min(std::max((int)std::floor((x1 - c->bb[0]) * c->tinvx), 0), g - 1);
    j0 = std::min(std::max((int)std::floor((y0 - c->bb[1]) * c->tinvy), 0), g - 1);
    j1 = std::min(std::max((int)std::floor((y1 - c->bb[1]) * c->tinvy), 0), g - 1);
  };
  std::vector<int> ptr((size_t)g * g + 1, 0);
  for (int f = 0; f < nF; ++f) {
    int i0, i1, j0, j1;
    range(f, i0, i1, j0, j1);
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) ptr[(size_t)j * g + i + 1]++;
  }
  for (size_t b = 0; b < (size_t)g * g; ++b) ptr[b + 1] += ptr[b];
  std::vector<int> faces(ptr.back()), fill(ptr.begin(), ptr.end() - 1);
  for (int f = 0; f < nF; ++f) {
    int i0, i1, j0, j1;
    range(f, i0, i1, j0, j1);
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) faces[fill[(size_t)j * g + i]++] = f;
  }
  CKR(upload(c, c->vx, vx, (size_t)nV * 8));
  CKR(upload(c, c->vy, vy, (size_t)nV * 8));
  CKR(upload(c, c->tri, tri, (size_t)nF * 12));
  CKR(upload(c, c->abc, abc, (size_t)nF * 24));
  CKR(upload(c, c->tbin_ptr, ptr.data(), ptr.size() * 4));
  CKR(upload(c, c->tbin_face, faces.data(), faces.size() * 4));
  CK(cudaStreamSynchronize(c->stream));
  c->mesh_mass = host_total_mass(nF, tri, vx, vy, abc);
  // A regular grid in disguise (the triangulation of an image, whichever diagonals it has)?  Then the integrating modes run
  // on the boundary-segment kernel; the face arrays above stay what the pieces / raster / counter paths use.
  c->grid_overlay = false;
  if (c->detect_grid) {
    int n = 0, m = 0;
    double box[4];
    std::vector<double> rho;
    std::vector<unsigned> bits;
    if (detect_regular_grid(nV, vx, vy, nF, tri, abc, n, m, box, rho, bits)) {
      c->gn = n; c->gm = m;
      c->gx0 = box[0]; c->gy0 = box[1];
      c->gdx = (box[2] - box[0]) / (n - 1); c->gdy = (box[3] - box[1]) / (m - 1);
      std::vector<double> pad((size_t)(n + 2) * (m + 2));
      for (int a = -1; a <= n; ++a)
        for (int b = -1; b <= m; ++b)
          pad[(size_t)(a + 1) * (m + 2) + (b + 1)] = rho[(size_t)std::min(std::max(a, 0), n - 1) * m + std::min(std::max(b, 0), m - 1)];
      CKR(upload(c, c->rho_p, pad.data(), pad.size() * 8));
      CKR(upload(c, c->diag, bits.data(), bits.size() * 4));
      CK(cudaStreamSynchronize(c->stream));
      c->grid_overlay = true;
    }
  }
  invalidate_eval(c);
  return MA_OK;
}

extern "C" int ma_set_mesh_pl(ma_ctx *c, int nV, const double *vx, const double *vy, const double *rho_v, int nF,
                              const int *tri, double *total_mass) {
  NEED_CTX();
  if (nV < 3 || nF < 1 || !vx || !vy || !tri || !rho_v) return fail(c, MA_INVALID, "ma_set_mesh_pl: bad arguments");
  for (int f = 0; f < 3 * nF; ++f)
    if (tri[f] < 0 || tri[f] >= nV) return fail(c, MA_INVALID, "ma_set_mesh_pl: vertex index out of range");
  std::vector<double> abc((size_t)3 * nF);
  host_pl_coefficients(nF, tri, vx, vy, rho_v, abc.data());
  if (total_mass) *total_mass = host_total_mass(nF, tri, vx, vy, abc.data());
  return ma_set_mesh(c, nV, vx, vy, nF, tri, abc.data());
}

extern "C" int ma_set_grid(ma_ctx *c, int n, int m, double x0, double y0, double x1, double y1, const double *rho_v,
                           double *total_mass) {
  NEED_CTX();
  if (n < 2 || m < 2 || !rho_v || !(x1 > x0) || !(y1 > y0)) return fail(c, MA_INVALID, "ma_set_grid: bad arguments");
  const double dx = (x1 - x0) / double(n - 1), dy = (y1 - y0) / double(m - 1);
  const size_t nF = (size_t)2 * (n - 1) * (m - 1);
  if (nF > 0x7fffffffull / 3) return fail(c, MA_INVALID, "ma_set_grid: grid too large for int32 face ids");
  std::vector<double> abc(3 * nF);
  double tot = 0;
  for (int i = 0; i + 1 < n; ++i)
    for (int j = 0; j + 1 < m; ++j) {
      // vertex coordinates exactly as functions.hpp:96-99 computes them
      double X0 = x0 + i * dx, X1 = x0 + (i + 1) * dx, Y0 = y0 + j * dy, Y1 = y0 + (j + 1) * dy;
      double r00 = rho_v[(size_t)i * m + j], r10 = rho_v[(size_t)(i + 1) * m + j];
      double r11 = rho_v[(size_t)(i + 1) * m + j + 1], r01 = rho_v[(size_t)i * m + j + 1];
      size_t f = 2 * ((size_t)i * (m - 1) + j);
      double ex = X1 - X0, ey = Y1 - Y0;
      {  // face {(i,j),(i+1,j),(i+1,j+1)}
        double a = (r10 - r00) / ex, b = (r11 - r10) / ey;
        abc[3 * f] = a; abc[3 * f + 1] = b; abc[3 * f + 2] = r00 - a * X0 - b * Y0;
        double gx = (X0 + X1 + X1) / 3, gy = (Y0 + Y0 + Y1) / 3;
        tot += 0.5 * ex * ey * (a * gx + b * gy + abc[3 * f + 2]);
      }
      {  // face {(i,j),(i+1,j+1),(i,j+1)}
        double a = (r11 - r01) / ex, b = (r01 - r00) / ey;
        abc[3 * f + 3] = a; abc[3 * f + 4] = b; abc[3 * f + 5] = r00 - a * X0 - b * Y0;
        double gx = (X0 + X1 + X0) / 3, gy = (Y0 + Y1 + Y1) / 3;
        tot += 0.5 * ex * ey * (a * gx + b * gy + abc[3 * f + 5]);
      }
    }
  if (total_mass) *total_mass = tot;
  c->mesh_kind = MESH_GRID;
  c->grid_overlay = false;
  c->gn = n; c->gm = m;
  c->gx0 = x0; c->gy0 = y0; c->gdx = dx; c->gdy = dy;
  c->nV = n * m;
  c->nF = (int)nF;
  c->bb[0] = x0; c->bb[1] = y0;
  c->bb[2] = x0 + (n - 1) * dx; c->bb[3] = y0 + (m - 1) * dy;
  CKR(upload(c, c->abc, abc.data(), abc.size() * 8));
  CKR(upload(c, c->rho_v, rho_v, (size_t)n * m * 8));
  {  // one replicated layer around the grid for the line-major K3 (ma_seg.cuh: seg_rv)
    std::vector<double> pad((size_t)(n + 2) * (m + 2));
    for (int a = -1; a <= n; ++a) {
      const double *src = rho_v + (size_t)std::min(std::max(a, 0), n - 1) * m;
      double *dst = pad.data() + (size_t)(a + 1) * (m + 2);
      dst[0] = src[0];
      memcpy(dst + 1, src, (size_t)m * 8);
      dst[m + 1] = src[m - 1];
    }
    CKR(upload(c, c->rho_p, pad.data(), pad.size() * 8));
    CK(cudaStreamSynchronize(c->stream));  // pad is freed at the end of this scope
  }
  CK(cudaStreamSynchronize(c->stream));
  invalidate_eval(c);
  return MA_OK;
}

extern "C" int ma_set_image(ma_ctx *c, int n, int m, const double *pixels, double *total_mass) {
  NEED_CTX();
  if (n < 2 || m < 2 || !pixels) return fail(c, MA_INVALID, "ma_set_image: bad arguments");
  std::vector<double> rho((size_t)n * m);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j)  // fgrid[p] = image(i, m-j-1)/255 + 1e-3   functions.hpp:102
      rho[(size_t)i * m + j] = pixels[(size_t)(m - j - 1) * n + i] / double(255) + 1e-3;
  return ma_set_grid(c, n, m, -1.0, -1.0, 1.0, 1.0, rho.data(), total_mass);
}

// =============================================================================================
// Diracs (K1, once per point set)
// =============================================================================================
namespace { int amg_build_hierarchy(ma_ctx *c); }

extern "C" int ma_set_points(ma_ctx *c, int N, const double *x, const double *y) {
  NEED_CTX();
  if (N < 1 || !x || !y) return fail(c, MA_INVALID, "ma_set_points: bad arguments");
  c->N = N;
  invalidate_eval(c);
  c->prev_stride = 0;  // the saved adjacency (seeds of the warm path / quick test) belongs to the previous point set
  c->seeds_global = false;
  c->warm_skip = 0;
  CKR(upload(c, c->x, x, (size_t)N * 8));
  CKR(upload(c, c->y, y, (size_t)N * 8));
  CKR(ensure(c, c->red_out, 16 * sizeof(double)));
  // bounding box of the points: {sum, sum of squares, min, max} of x and of y (one block over 1 M points took 0.3 ms)
  CKR(reduce4x2(c, c->x.as<double>(), c->y.as<double>(), N, c->red_out.as<double>(), c->stream));
  double rb[8];
  CK(cudaMemcpyAsync(rb, c->red_out.p, sizeof rb, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const double pb[4] = {rb[2], rb[6], rb[3], rb[7]};
  if (!(std::isfinite(pb[0]) && std::isfinite(pb[1]) && std::isfinite(pb[2]) && std::isfinite(pb[3]) && std::isfinite(rb[0]) && std::isfinite(rb[4])))  // (the sums catch a NaN, which min / max skip)
    return fail(c, MA_INVALID, "ma_set_points: non-finite coordinates");
  double ext = std::max(std::max(pb[2] - pb[0], pb[3] - pb[1]), 1e-300) * (1 + 1e-9);
  int L = 0;
  while (((long long)1 << (2 * L)) * c->bin_target < N && L < 12) ++L;
  const int G = 1 << L;
  c->L = L;
  c->px0 = pb[0]; c->py0 = pb[1];
  c->ph = ext / G;
  const double pinv = G / ext;
  const size_t nb = (size_t)G * G;
  CKR(ensure(c, c->code, (size_t)N * 4));
  CKR(ensure(c, c->bin_count, nb * 4));
  CKR(ensure(c, c->bin_start, (nb + 1) * 4));
  CKR(ensure(c, c->perm, (size_t)N * 4));
  CKR(ensure(c, c->pos, (size_t)N * 4));
  CKR(ensure(c, c->xs, (size_t)N * 8));
  CKR(ensure(c, c->ys, (size_t)N * 8));
  CKR(ensure(c, c->ws, (size_t)N * 8));
  CKR(ensure(c, c->w, (size_t)N * 8));
  CKR(ensure(c, c->wmax, ((4 * nb - 1) / 3) * 8));
  CK(cudaMemsetAsync(c->bin_count.p, 0, nb * 4, c->stream));
  k_bin_count<<<cdiv(N, 256), 256, 0, c->stream>>>(c->x.as<double>(), c->y.as<double>(), N, c->px0, c->py0, pinv, G,
                                                   c->code.as<unsigned>(), c->bin_count.as<int>());
  CKR(scan_i32(c, c->bin_count.as<int>(), c->bin_start.as<int>(), (int)nb));
  CK(cudaMemsetAsync(c->bin_count.p, 0, nb * 4, c->stream));
  k_bin_scatter<<<cdiv(N, 256), 256, 0, c->stream>>>(c->code.as<unsigned>(), N, c->bin_start.as<int>(),
                                                     c->bin_count.as<int>(), c->perm.as<int>());
  k_bin_sort<<<cdiv((long long)nb, 256), 256, 0, c->stream>>>(c->bin_start.as<int>(), (int)nb, c->perm.as<int>());
  CKR(ensure(c, c->bin_rm, nb * 8));
  k_bin_rowmajor<<<cdiv((long long)nb, 256), 256, 0, c->stream>>>(G, c->bin_start.as<int>(), c->bin_rm.as<int>());
  k_gather_points<<<cdiv(N, 256), 256, 0, c->stream>>>(c->x.as<double>(), c->y.as<double>(), c->perm.as<int>(), N,
                                                       c->xs.as<double>(), c->ys.as<double>(), c->pos.as<int>());
  {  // the sites once more, filed under the row-major bins of the block grid (about block_target Diracs per bin)
    int bG = (int)std::ceil(std::sqrt((double)N / c->block_target));
    bG = std::max(1, std::min(bG, 8192));
    c->bG = bG;
    c->bph = ext / bG;
    c->binv = bG / ext;
    const size_t nbb = (size_t)bG * bG;
    CKR(ensure(c, c->xr, (size_t)N * 8)); CKR(ensure(c, c->yr, (size_t)N * 8)); CKR(ensure(c, c->wr, (size_t)N * 8));
    CKR(ensure(c, c->rm2s, (size_t)N * 4)); CKR(ensure(c, c->s2rm, (size_t)N * 4));
    CKR(ensure(c, c->rm_start, (nbb + 1) * 4)); CKR(ensure(c, c->blk_cnt, nbb * 4)); CKR(ensure(c, c->scratch_i, (size_t)N * 4));
    CKR(ensure(c, c->hard1, (size_t)N * 4)); CKR(ensure(c, c->hard2, (size_t)N * 4)); CKR(ensure(c, c->hard3, (size_t)N * 4)); CKR(ensure(c, c->hard_n, 64));
    CK(cudaMemsetAsync(c->blk_cnt.p, 0, nbb * 4, c->stream));
    k_blk_count<<<cdiv(N, 256), 256, 0, c->stream>>>(c->xs.as<double>(), c->ys.as<double>(), N, c->px0, c->py0, c->binv, bG,
                                                     c->scratch_i.as<int>(), c->blk_cnt.as<int>());
    CKR(scan_i32(c, c->blk_cnt.as<int>(), c->rm_start.as<int>(), (int)nbb));
    CK(cudaMemsetAsync(c->blk_cnt.p, 0, nbb * 4, c->stream));
    k_blk_scatter<<<cdiv(N, 256), 256, 0, c->stream>>>(c->scratch_i.as<int>(), N, c->rm_start.as<int>(), c->blk_cnt.as<int>(),
                                                       c->rm2s.as<int>());
    k_blk_fill<<<cdiv((long long)nbb, 256), 256, 0, c->stream>>>((int)nbb, c->rm_start.as<int>(), c->rm2s.as<int>(),
                                                                c->xs.as<double>(), c->ys.as<double>(), c->xr.as<double>(),
                                                                c->yr.as<double>(), c->s2rm.as<int>());
    CK(cudaMemsetAsync(c->wr.p, 0, (size_t)N * 8, c->stream));
  }
  CK(cudaMemsetAsync(c->w.p, 0, (size_t)N * 8, c->stream));
  CK(cudaGetLastError());
  // per-node planes: sorted leaf codes + the geometric half of the moment prefix sums
  const size_t nnodes = (4 * nb - 1) / 3;
  CKR(ensure(c, c->code_s, (size_t)N * 4));
  CKR(ensure(c, c->pre0, (size_t)5 * (N + 1) * 8));
  CKR(ensure(c, c->pre1, (size_t)3 * (N + 1) * 8));
  CKR(ensure(c, c->nodeG, nnodes * 16));
  CKR(ensure(c, c->nodeA, nnodes * 8));
  k_gather_u32<<<cdiv(N, 256), 256, 0, c->stream>>>(c->code.as<unsigned>(), c->perm.as<int>(), N, c->code_s.as<unsigned>());
  CKR(moment_scan<0>(c, c->pre0.as<double>()));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  // (the multigrid hierarchy of this point set is built by the first linear solve that wants it: Lloyd iterations move the
  // points every time and never solve)
  c->amg.ready = false;
  c->amg.nlev = 0;
  c->amg_stale = true;
  return MA_OK;
}

// =============================================================================================
// multi-GPU: NCCL inside the engine.  One process per GPU; every rank holds the points,
// the weights and the mesh and evaluates its Morton tile of the Diracs; what crosses NVLink is
//   * per evaluation: 6 integers (capacity / abort flags, so that all ranks take the same branch) and 3 scalars
//     (f, sum of masses: sum; min mass: min);
//   * per ACCEPTED Newton point: the tiles' slices of m - nu, of the row counts and of the Hessian's CSR
//     (grouped ncclBroadcast, one per tile: the tiles differ in size by one row at most, but in nnz freely);
// the grounded Laplacian solve is then replicated on every rank — with the multigrid preconditioner it is ~10 ms at
// 1 M rows, less than a distributed CG would spend in its per-iteration all-gathers — and all ranks hold bit-identical
// weights without ever exchanging them (every kernel of the solve reduces in a fixed order).
// libnccl is dlopen()ed on first use so that libma_b200.so loads without it, and so that a process which already
// has torch's NCCL mapped binds to that very copy.
// =============================================================================================
namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi &nccl_api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) return a;
#define MA_NCCL_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, sym))
  MA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); MA_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  MA_NCCL_SYM(CommDestroy, "ncclCommDestroy"); MA_NCCL_SYM(AllReduce, "ncclAllReduce");
  MA_NCCL_SYM(Broadcast, "ncclBroadcast"); MA_NCCL_SYM(GroupStart, "ncclGroupStart");
  MA_NCCL_SYM(GroupEnd, "ncclGroupEnd"); MA_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef MA_NCCL_SYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast && a.GroupStart && a.GroupEnd;
  return a;
}
#define NCK(call)                                                                                            \
  do {                                                                                                       \
    ncclResult_t r_ = (call);                                                                                \
    if (r_ != ncclSuccess)                                                                                   \
      return fail(c, MA_CUDA_ERROR, "%s failed: %s", #call, nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "?"); \
  } while (0)

inline bool is_dist(const ma_ctx *c) { return c->comm != nullptr && c->part_n > 1; }
inline int tile_lo(const ma_ctx *c, int r) { return (int)((long long)c->N * r / c->part_n); }

// flags[0] bit field + flags[1] abort + flags[2] exact-stage count  <->  one integer per item, so that MAX / SUM apply
__global__ void k_flags_unpack(const int *__restrict__ flags, int *__restrict__ out) {
  if (threadIdx.x < 4) out[threadIdx.x] = (flags[0] >> threadIdx.x) & 1;
  if (threadIdx.x == 4) out[4] = flags[1] != 0;
  if (threadIdx.x == 5) out[5] = flags[2] > 0;
  if (threadIdx.x == 6) out[6] = flags[3] != 0;  // warm path: a vertex still failed the last ring match
}
__global__ void k_flags_pack(const int *__restrict__ in, int *__restrict__ flags) {
  if (threadIdx.x == 0) {
    flags[0] = (in[0] ? 1 : 0) | (in[1] ? 2 : 0) | (in[2] ? 4 : 0) | (in[3] ? 8 : 0);
    flags[1] = in[4];
    flags[2] = max(flags[2], in[5]);
    flags[3] = in[6];
  }
}
// every rank leaves with the union of all ranks' flags (they then take the same branch: escalate / abort / go on)
int dist_sync_flags(ma_ctx *c) {
  if (!is_dist(c)) return MA_OK;
  CKR(ensure(c, c->dist_buf, 256));
  int *b = c->dist_buf.as<int>();
  k_flags_unpack<<<1, 32, 0, c->stream>>>(c->flags.as<int>(), b);
  NCK(nccl_api().AllReduce(b, b, 7, ncclInt32, ncclMax, (ncclComm_t)c->comm, c->stream));
  k_flags_pack<<<1, 32, 0, c->stream>>>(b, c->flags.as<int>());
  c->launches += 2;
  return MA_OK;
}
// red = {sum f, -, -, -, sum m, -, min m, -} of the tile (reduce4 layout)  ->  of all tiles
__global__ void k_red_pack(const double *__restrict__ red, double *__restrict__ out) {
  if (threadIdx.x == 0) { out[0] = red[0]; out[1] = red[4]; out[2] = red[6]; }
}
__global__ void k_red_unpack(const double *__restrict__ in, double *__restrict__ red) {
  if (threadIdx.x == 0) { red[0] = in[0]; red[4] = in[1]; red[6] = in[2]; }
}
int dist_reduce_eval(ma_ctx *c) {
  if (!is_dist(c)) return MA_OK;
  CKR(ensure(c, c->dist_buf, 256));
  double *b = reinterpret_cast<double *>(c->dist_buf.as<char>() + 64);
  k_red_pack<<<1, 32, 0, c->stream>>>(c->red_out.as<double>(), b);
  NCK(nccl_api().AllReduce(b, b, 2, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  NCK(nccl_api().AllReduce(b + 2, b + 2, 1, ncclDouble, ncclMin, (ncclComm_t)c->comm, c->stream));
  k_red_unpack<<<1, 32, 0, c->stream>>>(b, c->red_out.as<double>());
  c->launches += 2;
  return MA_OK;
}
// in-place gather of the tiles' slices of a full-length array (element size esz): rank r owns [tile_lo(r), tile_lo(r+1))
int dist_gather_slices(ma_ctx *c, void *v, size_t esz) {
  if (!is_dist(c)) return MA_OK;
  NCK(nccl_api().GroupStart());
  for (int r = 0; r < c->part_n; ++r) {
    const int lo = tile_lo(c, r), hi = tile_lo(c, r + 1);
    if (hi > lo) {
      char *ptr = (char *)v + (size_t)lo * esz;
      NCK(nccl_api().Broadcast(ptr, ptr, (size_t)(hi - lo) * esz, ncclChar, r, (ncclComm_t)c->comm, c->stream));
    }
  }
  NCK(nccl_api().GroupEnd());
  return MA_OK;
}
__global__ void k_copy_rows(const int *__restrict__ rowptr_l, const int *__restrict__ rowptr_g, int lo, int hi,
                            const int *__restrict__ col_l, const double *__restrict__ val_l, int *__restrict__ col_g,
                            double *__restrict__ val_g) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int s = rowptr_l[i], e = rowptr_l[i + 1], o = rowptr_g[i];
  for (int q = s; q < e; ++q) { col_g[o + (q - s)] = col_l[q]; val_g[o + (q - s)] = val_l[q]; }
}
// The Hessian of ALL tiles on every rank (internal order): row counts -> global row pointers -> this tile's rows
// copied to their global place -> one broadcast per tile.  Leaves rowptr_g / col_g / val_g / nnz_g.
int dist_gather_hessian(ma_ctx *c) {
  const int N = c->N;
  CKR(dist_gather_slices(c, c->rowcnt.p, 4));
  CKR(ensure(c, c->rowptr_g, ((size_t)N + 1) * 4));
  CKR(scan_i32(c, c->rowcnt.as<int>(), c->rowptr_g.as<int>(), N));
  std::vector<int> off(c->part_n + 1);
  for (int r = 0; r <= c->part_n; ++r)
    CK(cudaMemcpyAsync(&off[r], c->rowptr_g.as<int>() + tile_lo(c, r), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->nnz_g = off[c->part_n];
  CKR(ensure(c, c->col_g, (size_t)std::max(c->nnz_g, 1) * 4, 1.2));
  CKR(ensure(c, c->val_g, (size_t)std::max(c->nnz_g, 1) * 8, 1.2));
  const int lo = tile_lo(c, c->part_rank), hi = tile_lo(c, c->part_rank + 1);
  if (hi > lo)
    k_copy_rows<<<cdiv(hi - lo, 256), 256, 0, c->stream>>>(c->rowptr.as<int>(), c->rowptr_g.as<int>(), lo, hi, c->col.as<int>(),
                                                          c->val.as<double>(), c->col_g.as<int>(), c->val_g.as<double>());
  c->launches++;
  NCK(nccl_api().GroupStart());
  for (int r = 0; r < c->part_n; ++r) {
    const size_t cnt = (size_t)(off[r + 1] - off[r]);
    if (!cnt) continue;
    NCK(nccl_api().Broadcast(c->col_g.as<int>() + off[r], c->col_g.as<int>() + off[r], cnt, ncclInt32, r, (ncclComm_t)c->comm, c->stream));
    NCK(nccl_api().Broadcast(c->val_g.as<double>() + off[r], c->val_g.as<double>() + off[r], cnt, ncclDouble, r, (ncclComm_t)c->comm, c->stream));
  }
  NCK(nccl_api().GroupEnd());
  CK(cudaGetLastError());
  return MA_OK;
}

}  // namespace

extern "C" int ma_comm_unique_id(void *id128) {
  if (!id128) return MA_INVALID;
  NcclApi &a = nccl_api();
  if (!a.ok) return MA_CUDA_ERROR;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != ncclSuccess) return MA_CUDA_ERROR;
  memcpy(id128, &id, 128);
  return MA_OK;
}
extern "C" int ma_comm_init(ma_ctx *c, int rank, int nranks, const void *id128) {
  if (!c) return MA_INVALID;
  if (!c->stream) return fail(c, MA_CUDA_ERROR, "context has no CUDA device");
  cudaSetDevice(c->device);
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, MA_INVALID, "ma_comm_init: bad arguments");
  NcclApi &a = nccl_api();
  if (!a.ok) return fail(c, MA_CUDA_ERROR, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
  if (c->comm) { a.CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  NCK(a.CommInitRank(&comm, nranks, id, rank));
  c->comm = comm;
  c->part_rank = rank;
  c->part_n = nranks;
  c->cells_nv = -1; c->have_eval = false; c->have_hessian = false;
  return MA_OK;
}
extern "C" int ma_comm_destroy(ma_ctx *c) {
  if (!c) return MA_INVALID;
  if (c->comm) { nccl_api().CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
  return MA_OK;
}

// =============================================================================================
// evaluation (K1 per-eval + K2 + K3 + K4)
// =============================================================================================
namespace {

int fill_params(ma_ctx *c, Params &p) {
  memset(&p, 0, sizeof p);
  p.N = c->N;
  p.cell_lo = (int)((long long)c->N * c->part_rank / c->part_n);
  p.cell_hi = (int)((long long)c->N * (c->part_rank + 1) / c->part_n);
  p.xs = c->xs.as<double>(); p.ys = c->ys.as<double>(); p.ws = c->ws.as<double>();
  p.L = c->L; p.px0 = c->px0; p.py0 = c->py0; p.ph = c->ph;
  p.bin_start = c->bin_start.as<int>();
  p.bin_rm = c->bin_rm.as<int>();
  p.wmax = c->wmax.as<double>();
  p.xr = c->xr.as<double>(); p.yr = c->yr.as<double>(); p.wr = c->wr.as<double>();
  p.rm2s = c->rm2s.as<int>(); p.rm_start = c->rm_start.as<int>();
  p.bG = c->bG; p.bph = c->bph; p.binv = c->binv;
  p.wstat = c->wstat.as<double>();
  p.nodeG = c->nodeG.as<double>();
  p.nodeA = c->nodeA.as<unsigned long long>();
  p.abort_flag = c->flags.as<int>() + 1;
  p.abort_on_empty = c->abort_on_empty ? 1 : 0;
  p.clip_a = c->clip_a; p.clip_b = c->clip_b; p.refill_at = c->refill_at; p.rmax = c->rmax; p.rtree = std::min(c->rtree, c->rmax);
  for (int k = 0; k < 4; ++k) p.bb[k] = c->bb[k];
  p.mesh_kind = c->mesh_kind;
  p.nF = c->nF;
  p.abc = c->abc.as<double>();
  p.rho_v = c->rho_v.as<double>();
  p.rho_p = c->rho_p.as<double>();
  p.diag = c->grid_overlay ? c->diag.as<unsigned>() : nullptr;
  p.gn = c->gn; p.gm = c->gm; p.gx0 = c->gx0; p.gy0 = c->gy0; p.gdx = c->gdx; p.gdy = c->gdy;
  p.vx = c->vx.as<double>(); p.vy = c->vy.as<double>(); p.tri = c->tri.as<int>();
  p.tg = c->tg; p.tinvx = c->tinvx; p.tinvy = c->tinvy;
  p.tbin_ptr = c->tbin_ptr.as<int>(); p.tbin_face = c->tbin_face.as<int>();
  p.kmax = c->kmax;
  p.nbr = c->nbr.as<int>(); p.nbr_cnt = c->nbr_cnt.as<int>(); p.cell_bb = c->cell_bb.as<double>();
  p.poly_x = c->poly_x.as<double>(); p.poly_y = c->poly_y.as<double>();
  p.poly_t = c->poly_t.as<int>(); p.poly_n = c->poly_n.as<int>();
  p.ring = c->ring.as<int>(); p.ring_n = c->ring_n.as<int>(); p.cstate = c->cstate.as<int>();
  p.mass = c->mass.as<double>(); p.fcell = c->fcell.as<double>(); p.hslot = c->hslot.as<double>();
  p.touched = c->touched.as<unsigned long long>(); p.rowcnt = c->rowcnt.as<int>();
  p.mom = c->mom.as<double>();
  p.counters = c->counters.as<unsigned long long>();
  p.stats = c->stats;
  p.flags = c->flags.as<int>();
  p.filter_tol = c->filter_tol;
  return MA_OK;
}

// CellSearch is about to run: the supporting planes must be complete
int join_planes(ma_ctx *c) {
  if (c->planes_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_fork[1], 0));
  c->planes_pending = false;
  return MA_OK;
}

constexpr int cells_maxv(int kmax) { return kmax == 16 ? 16 : (kmax == 32 ? 36 : 64); }

// K2 fast path: block kernels of radius 2, then 3, then CellSearch on what is left (lists stay on the device)
template <int NT, bool POLY> int launch_cells_lean(ma_ctx *c, const Params &p) {
  constexpr int MAXV = 16;
  const size_t sm = cells_smem_bytes<MAXV, NT>();
  const int ncells = p.cell_hi - p.cell_lo;
  int *cnt = c->hard_n.as<int>();
  CK(cudaMemsetAsync(cnt, 0, 16, c->stream));
  const size_t smb = sm;
  CK(cudaFuncSetAttribute(k_cells_block<-1, 2, MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
  CK(cudaFuncSetAttribute(k_cells_block<2, 3, MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
  CK(cudaFuncSetAttribute(k_cells_persist<MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int nblk = std::max(1, cdiv(ncells, NT));
  k_cells_block<-1, 2, MAXV, NT, POLY><<<nblk, NT, smb, c->stream>>>(p, nullptr, nullptr, c->hard1.as<int>(), cnt);
  c->k3_split = false;
  if (c->k3_early) {
    // K3 of the cells this kernel certified starts now, on the side stream, under the tail of K2 below
    CK(cudaEventRecord(c->ev_fork[4], c->stream));
    CK(cudaStreamWaitEvent(c->side, c->ev_fork[4], 0));
    CKR(c->k3_early());
    CK(cudaEventRecord(c->ev_fork[5], c->side));
    c->k3_split = true;
  }
  // The later stages see a fraction of the cells (or, with graded weights, all of them: k_cells_persist then ignores the
  // list): the ~15 % the 5 x 5 block cannot certify continue from their polygon with the ring of bins around it, the
  // ~0.4 % left after that get a warp each and the block of radius 5 (one by one through CellSearch those few cells
  // cost 0.6 ms of pure latency, profiles/r02b), CellSearch takes whatever remains.  (hard1 keeps the list of the first
  // kernel: it is K3's second pass.)
  const int nblk2 = std::max(1, std::min(nblk, c->sm_count * 8));
  k_cells_block<2, 3, MAXV, NT, POLY><<<nblk2, NT, smb, c->stream>>>(p, c->hard1.as<int>(), cnt, c->hard2.as<int>(), cnt + 1);
  k_cells_warp<3, 5, POLY><<<c->sm_count * 16, 128, 0, c->stream>>>(p, c->hard2.as<int>(), cnt + 1, c->hard3.as<int>(), cnt + 2);
  // CellSearch: a small grid for the leftovers of the list (a handful of cells, if any) ...
  CKR(join_planes(c));
  k_cells_persist<MAXV, NT, POLY><<<c->sm_count * 2, NT, sm, c->stream>>>(p, 1, c->hard3.as<int>(), cnt + 2);
  // ... and a grid sized for the whole tile that only works when the weights are graded (the regime of the Newton iterates)
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cells_persist<MAXV, NT, POLY>, NT, sm));
  const long long warps_target = (long long)c->sm_count * std::max(per_sm, 1) * (NT / 32) * c->persist_waves;
  const int chunk = (int)std::max<long long>(c->persist_min_chunk, (ncells + warps_target - 1) / warps_target);
  const int nwarps = std::max(1, cdiv(ncells, chunk));
  k_cells_persist<MAXV, NT, POLY><<<cdiv(nwarps, NT / 32), NT, sm, c->stream>>>(p, chunk, nullptr, cnt + 2);
  c->launches += 5;
  CK(cudaGetLastError());
  return MA_OK;
}

// K2 warm path (ma_warm.cuh): seed every cell from the saved adjacency, certify by ring matching, let CellSearch rebuild
// the cells around every failing vertex, twice; the count of the last match goes to the host with the other scalars.
template <int NT, bool POLY> int launch_cells_warm(ma_ctx *c, const Params &tile) {
  constexpr int MAXV = 16;
  const size_t sm = cells_smem_bytes<MAXV, NT>();
  // The certificate is global (every vertex of every cell): with the Diracs partitioned over several GPUs each rank runs
  // this stage on ALL cells — a couple of ms at 1 M Diracs, what the sharded quadtree walk costs on 8 GPUs — and K3 / K4
  // on its own tile; every rank reaches the same verdict from the same inputs.
  Params p = tile;
  p.cell_lo = 0; p.cell_hi = p.N;
  const int ncells = p.N;
  int *cnt = c->hard_n.as<int>();  // [0], [1]: lengths of the two rebuild lists, [5]: of the last match's; [2], [3]: failing cells per match
  CK(cudaMemsetAsync(cnt, 0, 32, c->stream));
  CK(cudaFuncSetAttribute(k_cells_seed<NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  CK(cudaFuncSetAttribute(k_cells_persist<MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int nblk = std::max(1, cdiv(ncells, NT)), nblk_m = std::max(1, cdiv(ncells, 256));
  k_cells_seed<NT, POLY><<<nblk, NT, sm, c->stream>>>(p, c->nbr_prev.as<int>(), c->cnt_prev.as<int>(), c->hard1.as<int>(), cnt);
  k_cells_match<<<nblk_m, 256, 0, c->stream>>>(p, c->hard1.as<int>(), cnt, cnt + 2);
  CKR(join_planes(c));
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cells_persist<MAXV, NT, POLY>, NT, sm));
  k_cells_persist<MAXV, NT, POLY><<<c->sm_count * std::max(per_sm, 1), NT, sm, c->stream>>>(p, 1, c->hard1.as<int>(), cnt, true);
  k_cells_match<<<nblk_m, 256, 0, c->stream>>>(p, c->hard2.as<int>(), cnt + 1, cnt + 3);
  k_cells_persist<MAXV, NT, POLY><<<c->sm_count * 2, NT, sm, c->stream>>>(p, 1, c->hard2.as<int>(), cnt + 1, true);
  k_cells_match<<<nblk_m, 256, 0, c->stream>>>(p, c->hard1.as<int>(), cnt + 5, p.flags + 3);  // the verdict: flags[3] failing cells
  c->launches += 6;
  CK(cudaGetLastError());
  return MA_OK;
}

template <int MAXV, int NT, bool POLY> int launch_cells(ma_ctx *c, const Params &p) {
  size_t sm = cells_smem_bytes<MAXV, NT>();
  const int ncells = p.cell_hi - p.cell_lo;
  if constexpr (MAXV == 16) {
    if (c->warm_now) return launch_cells_warm<NT, POLY>(c, p);
    if (c->lean && c->persist) return launch_cells_lean<NT, POLY>(c, p);
  }
  CKR(join_planes(c));
  // The larger capacity classes (array polygons) run one cell per lane: k_cells_persist<36 / 64> does not finish on graded
  // weights at 1 M Diracs (seen at c5 x 4 M and c3 x 1 M with kmax = 32, profiles/r02q: > 30 s against 47 ms for k_cells on
  // the same input, whatever the scheduling knobs; the 16-vertex instantiation is not affected).  Not understood yet; these
  // classes are the rare fallback (P(n >= 17) ~ 5e-10 per cell on random points), so they take the slower, proven kernel.
  if (c->persist && MAXV == 16) {
    // persistent lanes: each warp owns `chunk` consecutive cells; about 3 waves of resident blocks
    CK(cudaFuncSetAttribute(k_cells_persist<MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cells_persist<MAXV, NT, POLY>, NT, sm));
    const long long warps_target = (long long)c->sm_count * std::max(per_sm, 1) * (NT / 32) * c->persist_waves;
    int chunk = (int)std::max<long long>(c->persist_min_chunk, (ncells + warps_target - 1) / warps_target);
    const int nwarps = std::max(1, cdiv(ncells, chunk));
    k_cells_persist<MAXV, NT, POLY><<<cdiv(nwarps, NT / 32), NT, sm, c->stream>>>(p, chunk, nullptr, nullptr);
  } else {
    CK(cudaFuncSetAttribute(k_cells<MAXV, NT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_cells<MAXV, NT, POLY><<<std::max(1, cdiv(ncells, NT)), NT, sm, c->stream>>>(p);
  }
  c->launches++;
  CK(cudaGetLastError());
  return MA_OK;
}
template <int KMAX, int MAXV, int MODE> int launch_pieces(ma_ctx *c, const Params &p) {
  size_t sm = pieces_warp_bytes<KMAX, MAXV>() * PIECES_WPB;
  CK(cudaFuncSetAttribute(k_pieces<KMAX, MAXV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_pieces<KMAX, MAXV, MODE><<<std::max(1, cdiv(p.cell_hi - p.cell_lo, PIECES_WPB)), PIECES_WPB * 32, sm, c->stream>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  return MA_OK;
}
template <int MODE> int launch_pieces_mode(ma_ctx *c, const Params &p) {
  switch (c->kmax) {
    case 16: return launch_pieces<16, 12, MODE>(c, p);
    case 32: return launch_pieces<32, 20, MODE>(c, p);
    default: return launch_pieces<64, 40, MODE>(c, p);
  }
}
template <bool POLY> int launch_cells_kmax(ma_ctx *c, const Params &p) {
  switch (c->kmax) {
    case 16: return launch_cells<16, 128, POLY>(c, p);
    case 32: return launch_cells<36, 64, POLY>(c, p);
    default: return launch_cells<64, 32, POLY>(c, p);
  }
}
template <int MAXV, int NT, int MODE> int launch_seg(ma_ctx *c, const Params &p, int sel = SEG_ALL, cudaStream_t st = nullptr) {
  size_t sm = seg_smem_bytes<MAXV, NT, MODE>();
  const int nblk = std::max(1, cdiv(p.cell_hi - p.cell_lo, NT));
  if (p.diag) {  // an explicit triangulation recognised as a grid: either diagonal per square
    CK(cudaFuncSetAttribute(k_seg<MAXV, NT, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_seg<MAXV, NT, MODE, true><<<nblk, NT, sm, st ? st : c->stream>>>(p, sel, c->hard1.as<int>(), c->hard_n.as<int>());
  } else {
    CK(cudaFuncSetAttribute(k_seg<MAXV, NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_seg<MAXV, NT, MODE><<<nblk, NT, sm, st ? st : c->stream>>>(p, sel, c->hard1.as<int>(), c->hard_n.as<int>());
  }
  c->launches++;
  CK(cudaGetLastError());
  return MA_OK;
}
template <int MODE> int launch_seg_kmax(ma_ctx *c, const Params &p) {
  switch (c->kmax) {
    case 16: {
      if (!c->k3_split) return launch_seg<16, 128, MODE>(c, p);
      // the second pass (launch_cells_lean started the first one): the cells the first block kernel left to K2's later stages
      CKR((launch_seg<16, 128, MODE>(c, p, SEG_REST)));
      CK(cudaStreamWaitEvent(c->stream, c->ev_fork[5], 0));
      c->k3_split = false;
      return MA_OK;
    }
    case 32: return launch_seg<36, 64, MODE>(c, p);
    default: return launch_seg<64, 32, MODE>(c, p);
  }
}
// the boundary-segment kernel handles grid meshes in the integrating modes
template <int MODE> bool use_seg(const ma_ctx *c) {
  return (c->mesh_kind == MESH_GRID || c->grid_overlay) && c->strategy == 0 && !c->stats &&
         (MODE == MODE_KANTOROVICH || MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2);
}

int alloc_eval(ma_ctx *c) {
  const size_t N = c->N, K = c->kmax;
  CKR(ensure(c, c->nbr, N * K * 4));
  CKR(ensure(c, c->nbr_cnt, N * 4));
  CKR(ensure(c, c->cell_bb, N * 32));
  CKR(ensure(c, c->mass, N * 8));
  CKR(ensure(c, c->fcell, N * 8));
  CKR(ensure(c, c->hslot, N * K * 8));
  CKR(ensure(c, c->touched, N * 8));
  CKR(ensure(c, c->rowcnt, N * 4));
  CKR(ensure(c, c->rowptr, (N + 1) * 4));
  CKR(ensure(c, c->counters, CNT_N * 8));
  CKR(ensure(c, c->flags, 16));
  CKR(ensure(c, c->red_out, 16 * sizeof(double)));
  CKR(ensure(c, c->wstat, 4 * sizeof(double)));
  // scratch of the scans / reductions (no allocation may happen while an evaluation is being captured into a graph)
  CKR(ensure(c, c->scan_tmp, (size_t)(cdiv((int)N, SCAN_TILE) + 2) * sizeof(int)));
  CKR(ensure(c, c->red_partial, (size_t)RED_BLOCKS * 8 * sizeof(double)));
  CKR(ensure(c, c->red_partial2, (size_t)RED_BLOCKS * 8 * sizeof(double)));
  CKR(ensure(c, c->fs_tiles, (size_t)8 * std::max(1, cdiv((int)N, FS_TILE)) * 8));
  // cell polygons: what k_seg integrates, and what the block kernels hand from one pass to the next
  const size_t slots = (size_t)(c->kmax == 16 ? 16 : (c->kmax == 32 ? 36 : 64)) * N;
  CKR(ensure(c, c->poly_x, slots * 8)); CKR(ensure(c, c->poly_y, slots * 8));
  CKR(ensure(c, c->poly_t, slots * 4)); CKR(ensure(c, c->poly_n, N * 4));
  CKR(ensure(c, c->cstate, N * 4));
  if (c->warm && c->prev_stride == RING_STRIDE) { CKR(ensure(c, c->ring, N * RING_STRIDE * 4)); CKR(ensure(c, c->ring_n, N * 4)); }
  return MA_OK;
}

// K1 per-eval part + K2 on the current device weights (c->w, caller order)
template <bool POLY = false> int run_cells(ma_ctx *c, Params &p) {
  const int N = c->N;
  k_gather_w<<<cdiv(N, 256), 256, 0, c->stream>>>(c->w.as<double>(), c->perm.as<int>(), c->s2rm.as<int>(), N,
                                                  c->ws.as<double>(), c->wr.as<double>());
  const size_t nb = (size_t)1 << (2 * c->L);
  CKR(reduce4(c, c->ws.as<double>(), nullptr, N, c->wstat.as<double>()));  // weight range and maximum (the certificates' w_max)
  c->launches += 1;
  // The max-weight pyramid and the supporting planes are read by CellSearch's tree walk only (graded weights, or the few
  // cells the block kernels leave over): they are built on the side stream while the block kernels run.
  const PlaneGate gate{c->wstat.as<double>(), 0.25 * c->ph * c->ph};
  CK(cudaEventRecord(c->ev_fork[0], c->stream));
  CK(cudaStreamWaitEvent(c->side, c->ev_fork[0], 0));
  k_wmax_leaf<<<cdiv((long long)nb, 256), 256, 0, c->side>>>(c->ws.as<double>(), c->bin_start.as<int>(), c->L, c->wmax.as<double>());
  if (c->L >= 5) k_wmax_top<<<1, 1024, 0, c->side>>>(c->L, c->wmax.as<double>());
  c->launches += 1 + (c->L >= 5);
  CKR(moment_scan<1>(c, c->pre1.as<double>(), gate, c->side));
  {
    const size_t nnodes = (4 * nb - 1) / 3;
    k_node_fit<<<cdiv((long long)nnodes, 256), 256, 0, c->side>>>(c->L, c->bin_start.as<int>(), N, c->pre0.as<double>(),
                                                                  c->pre1.as<double>(), c->nodeG.as<double>(),
                                                                  c->nodeA.as<unsigned long long>(), gate);
    k_node_alpha<<<cdiv(N, 256), 256, 0, c->side>>>(N, c->L, c->xs.as<double>(), c->ys.as<double>(), c->ws.as<double>(),
                                                    c->code_s.as<unsigned>(), c->px0, c->py0, c->ph, c->nodeG.as<double>(),
                                                    c->nodeA.as<unsigned long long>(), gate);
    c->launches += 2;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev_fork[1], c->side));
  c->planes_pending = true;
  if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_PREP + 1], c->stream));
  CKR(launch_cells_kmax<POLY>(c, p));
  if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_CELLS + 1], c->stream));
  return MA_OK;
}

// remembers the adjacency of the evaluation just done (an accepted Newton point): the seeds of the warm
// path and of the quick empty-cell test
int save_adjacency(ma_ctx *c) {
  const size_t N = c->N, K = c->kmax;
  CKR(ensure(c, c->nbr_prev, N * K * 4));
  CKR(ensure(c, c->cnt_prev, N * 4));
  CK(cudaMemcpyAsync(c->nbr_prev.p, c->nbr.p, N * K * 4, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->cnt_prev.p, c->nbr_cnt.p, N * 4, cudaMemcpyDeviceToDevice, c->stream));
  c->prev_stride = (int)K;
  c->seeds_global = c->part_n == 1 || c->k2_all;
  if (c->comm && c->part_n > 1 && !c->k2_all && c->warm && K == RING_STRIDE) {
    // a cold evaluation knows its own tile only: the warm path wants the rows of all tiles on every rank
    CKR(dist_gather_slices(c, c->nbr_prev.p, K * 4));
    CKR(dist_gather_slices(c, c->cnt_prev.p, 4));
    c->seeds_global = true;
  }
  return MA_OK;
}

// one evaluation with automatic capacity escalation
template <int MODE> int evaluate_mode(ma_ctx *c, bool with_hessian) {
  if (c->mesh_kind == MESH_NONE) return fail(c, MA_INVALID, "no mesh set");
  if (c->N < 1) return fail(c, MA_INVALID, "no points set");
  c->capacity_hit = false;
  c->kmax = c->kmax_base;
  for (int attempt = 0; attempt < 4; ++attempt) {
    // warm path of K2: the adjacency of an earlier evaluation of these points seeds this one (ma_warm.cuh)
    c->warm_now = false;
    if (c->warm && (c->warm == 2 || c->in_newton) && c->kmax == 16 && (c->part_n == 1 || (c->comm && c->seeds_global)) && c->prev_stride == RING_STRIDE && c->persist &&
        (MODE == MODE_KANTOROVICH || MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2)) {
      if (c->warm_skip > 0) --c->warm_skip;
      else c->warm_now = true;
    }
    CKR(alloc_eval(c));
    if (MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2) CKR(ensure(c, c->mom, (size_t)c->N * 48));
    if (use_seg<MODE>(c)) {
      const size_t slots = (size_t)cells_maxv(c->kmax) * c->N;
      CKR(ensure(c, c->poly_x, slots * 8)); CKR(ensure(c, c->poly_y, slots * 8));
      CKR(ensure(c, c->poly_t, slots * 4)); CKR(ensure(c, c->poly_n, (size_t)c->N * 4));
    }
    Params p;
    fill_params(c, p);
    const int lo = p.cell_lo, nloc = p.cell_hi - p.cell_lo;
    constexpr int SEG_MODE = (MODE == MODE_KANTOROVICH || MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2) ? MODE : 0;
    const bool seg = use_seg<MODE>(c);
    const bool hess = (MODE == MODE_KANTOROVICH) && with_hessian;
    auto csr_fill = [&]() -> int {
      const int cap = (int)std::min<size_t>(c->col.cap / 4, c->val.cap / 8);
      const int *skip = c->abort_on_empty ? c->flags.as<int>() + 1 : nullptr;
      switch (c->kmax) {
        case 16: k_csr_fill<16><<<std::max(1, cdiv(nloc, 32 * csr_wpb<16>())), 32 * csr_wpb<16>(), 0, c->stream>>>(p.cell_lo, p.cell_hi, p.nbr, p.hslot, p.touched, c->rowptr.as<int>(), c->col.as<int>(), c->val.as<double>(), cap, skip); break;
        case 32: k_csr_fill<32><<<std::max(1, cdiv(nloc, 32 * csr_wpb<32>())), 32 * csr_wpb<32>(), 0, c->stream>>>(p.cell_lo, p.cell_hi, p.nbr, p.hslot, p.touched, c->rowptr.as<int>(), c->col.as<int>(), c->val.as<double>(), cap, skip); break;
        default: k_csr_fill<64><<<std::max(1, cdiv(nloc, 32 * csr_wpb<64>())), 32 * csr_wpb<64>(), 0, c->stream>>>(p.cell_lo, p.cell_hi, p.nbr, p.hslot, p.touched, c->rowptr.as<int>(), c->col.as<int>(), c->val.as<double>(), cap, skip); break;
      }
      c->launches++;
      CK(cudaGetLastError());
      return MA_OK;
    };
    if (hess) {
      // K4 is launched BEFORE the host knows nnz (no synchronisation in the middle of an evaluation): col / val keep the
      // size of the largest Hessian seen so far plus headroom (8 entries per row to start with), the kernel drops what
      // does not fit and the fill is repeated in the rare case that it did not
      const size_t want = std::max<size_t>((size_t)8 * nloc, 1024);
      if (c->col.cap / 4 < want || c->val.cap / 8 < want) { CKR(ensure(c, c->col, want * 4)); CKR(ensure(c, c->val, want * 8)); }
    }
    // Everything the evaluation launches, in stream order (no allocation, no synchronisation in here: the sequence is
    // captured once per distinct configuration and replayed as a CUDA graph — ~30 launches of which ~20 are tiny).
    auto stage = [&](const char *what) {  // MA_TRACE=2: synchronise after every stage and say which one just ended
      if (c->trace < 2) return;
      const cudaError_t e = cudaStreamSynchronize(c->stream);
      static auto t_first = std::chrono::steady_clock::now();
      const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_first).count();
      fprintf(stderr, "[ma]   %9.3f s  stage %s done (kmax=%d, abort_on_empty=%d)%s%s\n", t, what, c->kmax, (int)c->abort_on_empty, e == cudaSuccess ? "" : ": ", e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    auto enqueue = [&]() -> int {
      CK(cudaMemsetAsync(c->flags.p, 0, 16, c->stream));
      if (c->stats) CK(cudaMemsetAsync(c->counters.p, 0, CNT_N * 8, c->stream));
      if (c->profiling) CK(cudaEventRecord(c->ev[0], c->stream));
      if (MODE == MODE_KANTOROVICH && c->part_n > 1) {
        // rows outside this context's Morton tile are empty here (another GPU owns them)
        const size_t N = c->N, hi = p.cell_hi;
        auto zero = [&](Buf &b, size_t esz) {
          if (lo) cudaMemsetAsync(b.p, 0, (size_t)lo * esz, c->stream);
          if (hi < N) cudaMemsetAsync((char *)b.p + hi * esz, 0, (N - hi) * esz, c->stream);
        };
        zero(c->mass, 8); zero(c->fcell, 8); zero(c->touched, 8); zero(c->rowcnt, 4);
      }
      c->k3_early = nullptr;
      c->k3_split = false;
      if (seg && c->kmax == 16 && c->lean && c->persist && !c->warm_now && c->k3_overlap)
        c->k3_early = [c, &p]() -> int {
          constexpr int SM = (MODE == MODE_KANTOROVICH || MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2) ? MODE : 0;
          return launch_seg<16, 128, SM>(c, p, SEG_CERTIFIED, c->side);
        };
      const int rc_cells = seg ? run_cells<true>(c, p) : run_cells<false>(c, p);
      c->k3_early = nullptr;
      CKR(rc_cells);
      stage("K1+K2");
      // (line-search trials: once K2 has found an empty cell the point is rejected whatever the rest says,
      // optimal_transport.hpp:167 — K3 / K4 then return at once on the device-side flag, the host learns it at the one
      // synchronisation below)
      if (seg) CKR(launch_seg_kmax<SEG_MODE>(c, p));
      else CKR(launch_pieces_mode<MODE>(c, p));
      if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_PIECES + 1], c->stream));
      stage("K3");
      if (MODE == MODE_KANTOROVICH) {
        // f, sum m, min m on the side stream while the main stream assembles the Hessian
        CK(cudaEventRecord(c->ev_fork[2], c->stream));
        CK(cudaStreamWaitEvent(c->side, c->ev_fork[2], 0));
        CKR(reduce4x2(c, c->fcell.as<double>() + lo, c->mass.as<double>() + lo, nloc, c->red_out.as<double>(), c->side));
        CK(cudaEventRecord(c->ev_fork[3], c->side));
        // row pointers of this context's rows only (rowptr[lo] = 0 ... rowptr[hi] = nnz); rows of other tiles have none
        if (hess) CKR(scan_i32(c, c->rowcnt.as<int>() + lo, c->rowptr.as<int>() + lo, nloc));
      }
      if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_REDUCE + 1], c->stream));
      if (hess) {
        CKR(csr_fill());
        if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_CSR + 1], c->stream));
      }
      if (MODE == MODE_KANTOROVICH) CK(cudaStreamWaitEvent(c->stream, c->ev_fork[3], 0));
      stage("scan + K4 + reductions");
      return MA_OK;
    };
    c->aborted = false;
    if (c->use_graph && !c->profiling && !c->stats) {
      // key: every launch argument and dimension is a function of these bytes
      std::string key((const char *)&p, sizeof p);
      const void *ptrs[] = {c->w.p, c->perm.p, c->s2rm.p, c->pre0.p, c->pre1.p, c->code_s.p, c->fs_tiles.p, c->hard1.p, c->hard2.p, c->hard3.p, c->cstate.p,
                            c->hard_n.p, c->rowptr.p, c->col.p, c->val.p, c->scan_tmp.p, c->red_partial.p, c->red_partial2.p,
                            c->red_out.p, c->mom.p};
      const long long ints[] = {c->k3_overlap, c->warm_now, (long long)(size_t)c->nbr_prev.p, (long long)(size_t)c->cnt_prev.p, MODE, hess, seg, c->lean, c->persist, c->persist_waves, c->persist_min_chunk, c->L, c->sm_count,
                                (long long)c->col.cap, (long long)c->val.cap, c->part_rank, c->part_n, c->abort_on_empty};
      key.append((const char *)ptrs, sizeof ptrs);
      key.append((const char *)ints, sizeof ints);
      ma_ctx::EvalGraph *g = nullptr;
      for (auto &e : c->graphs)
        if (e.key == key) g = &e;
      if (!g && key != c->graph_candidate) {
        // a configuration seen for the first time runs directly; it is captured when it comes back (Lloyd iterations move
        // the points — and with them the bins' origin, part of the key — every time: capturing + instantiating a graph
        // that is launched once costs more than the launches it saves)
        c->graph_candidate = key;
        CKR(enqueue());
      } else if (!g) {
        const long long l0 = c->launches;
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        const int rc_q = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(c->stream, &graph);
        c->planes_pending = false;
        if (rc_q != MA_OK) { if (graph) cudaGraphDestroy(graph); return rc_q; }
        if (e_end != cudaSuccess) return fail(c, MA_CUDA_ERROR, "graph capture of the evaluation failed: %s", cudaGetErrorString(e_end));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e_inst != cudaSuccess) return fail(c, MA_CUDA_ERROR, "cudaGraphInstantiate: %s", cudaGetErrorString(e_inst));
        if (c->graphs.size() >= 8) {  // evict the least recently used
          size_t o = 0;
          for (size_t k = 1; k < c->graphs.size(); ++k)
            if (c->graphs[k].stamp < c->graphs[o].stamp) o = k;
          cudaGraphExecDestroy(c->graphs[o].exec);
          c->graphs.erase(c->graphs.begin() + o);
        }
        c->graphs.push_back(ma_ctx::EvalGraph{key, exec, (int)(c->launches - l0), 0});
        c->launches = l0;
        g = &c->graphs.back();
      }
      if (g) {
        g->stamp = ++c->graph_clock;
        CK(cudaGraphLaunch(g->exec, c->stream));
        c->launches += g->launches;
      }
    } else {
      CKR(enqueue());
    }
    const bool was_warm = c->warm_now;
    c->k2_all = was_warm || c->part_n == 1;
    c->warm_now = false;  // (ma_cells_build / ma_pieces_build share launch_cells: they never take the warm path)
    CKR(dist_sync_flags(c));  // multi-GPU: every rank learns about an overflow / an empty cell of ANY tile
    if (MODE == MODE_KANTOROVICH) CKR(dist_reduce_eval(c));  // f, sum m, min m over all tiles
    c->hs->flags = 0; c->hs->nnz = 0;
    CK(cudaMemcpyAsync(&c->hs->flags, c->flags.p, 16, cudaMemcpyDeviceToHost, c->stream));  // flags, abort, K2 exact-stage count, warm-path failures
    if (was_warm) CK(cudaMemcpyAsync(c->hs->warm, c->hard_n.p, 32, cudaMemcpyDeviceToHost, c->stream));
    if (MODE == MODE_KANTOROVICH) {
      CK(cudaMemcpyAsync(c->hs->red, c->red_out.p, sizeof c->hs->red, cudaMemcpyDeviceToHost, c->stream));
      if (hess) CK(cudaMemcpyAsync(&c->hs->nnz, c->rowptr.as<int>() + p.cell_hi, 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[MA_T_COUNT + 1], c->stream));
    if (c->want_async && attempt == 0 && MODE == MODE_KANTOROVICH && !was_warm && !c->abort_on_empty && !is_dist(c) && !c->profiling && !c->stats) {
      // ma_evaluate_async: the launches and the read-back of the scalars are queued, the host does not wait (finish_pending does)
      c->pending = true;
      c->pending_hess = hess;
      return MA_OK;
    }
    CK(cudaStreamSynchronize(c->stream));
    const int h_flags = c->hs->flags, h_nnz = c->hs->nnz;
    c->cell_fallbacks = c->hs->cell_fallbacks;
    const double *red = c->hs->red;
    if (h_flags & (FLAG_CELL_OVERFLOW | FLAG_PIECE_OVERFLOW | FLAG_KMAX_OVERFLOW)) {
      // (before the abort flag: a cell that outgrew its capacity class is not an empty cell)
      if (c->trace) fprintf(stderr, "[ma] eval kmax=%d overflow flags=%d\n", c->kmax, h_flags);
      if (c->kmax >= 64) {
        c->capacity_hit = true;
        return fail(c, MA_INVALID, "polygon capacity exceeded even at kmax=64 (flags=%d)", h_flags);
      }
      c->kmax *= 2;  // escalate to the next capacity class and redo the evaluation
      continue;
    }
    if (h_flags & FLAG_STACK_OVERFLOW) return fail(c, MA_INVALID, "quadtree stack overflow");
    if (c->abort_on_empty && c->hs->abort_) {
      c->aborted = true;
      c->mass_min = 0.0;
      invalidate_eval(c);
      if (c->trace) fprintf(stderr, "[ma] eval aborted after K2: a cell is empty\n");
      return MA_OK;
    }
    if (was_warm) {
      // certified iff the last ring match found nothing and the cells are ONE sheet over the mesh (ma_warm.cuh)
      const int *wn = c->hs->warm;
      bool ok = c->hs->warm_fail == 0;
      if (ok && MODE == MODE_KANTOROVICH && c->mesh_mass > 0) ok = std::fabs(red[4] - c->mesh_mass) <= 1e-6 * c->mesh_mass;
      if (c->trace) fprintf(stderr, "[ma] warm K2: rebuilt %d + %d cells, failing cells per match %d %d %d%s\n", wn[0], wn[1], wn[2], wn[3], c->hs->warm_fail, ok ? "" : "  -> redone cold");
      if (!ok) {
        c->warm_failed++;
        c->warm_skip = c->warm_penalty;
        c->warm_penalty = std::min(c->warm_penalty * 2, 1024);
        continue;
      }
      c->warm_evals++;
      c->warm_rebuilt += wn[0] + wn[1];
      c->warm_penalty = 4;
    }
    if (MODE == MODE_KANTOROVICH) {
      c->fval = red[0];
      c->mass_sum = red[4];
      c->mass_min = red[6];
      if (hess) {
        c->nnz = h_nnz;
        if ((size_t)h_nnz > std::min<size_t>(c->col.cap / 4, c->val.cap / 8)) {  // did not fit: grow and fill again
          CKR(ensure(c, c->col, (size_t)h_nnz * 4, 1.3));
          CKR(ensure(c, c->val, (size_t)h_nnz * 8, 1.3));
          CKR(csr_fill());
          CK(cudaStreamSynchronize(c->stream));
        }
      }
    }
    if (c->profiling) {
      for (auto &t : c->t_ms) t = 0;
      cudaEventElapsedTime(&c->t_ms[MA_T_TOTAL], c->ev[0], c->ev[MA_T_COUNT + 1]);
      cudaEventElapsedTime(&c->t_ms[MA_T_PREP], c->ev[0], c->ev[MA_T_PREP + 1]);
      cudaEventElapsedTime(&c->t_ms[MA_T_CELLS], c->ev[MA_T_PREP + 1], c->ev[MA_T_CELLS + 1]);
      cudaEventElapsedTime(&c->t_ms[MA_T_PIECES], c->ev[MA_T_CELLS + 1], c->ev[MA_T_PIECES + 1]);
      if (MODE == MODE_KANTOROVICH && with_hessian) {
        cudaEventElapsedTime(&c->t_ms[MA_T_REDUCE], c->ev[MA_T_PIECES + 1], c->ev[MA_T_REDUCE + 1]);
        cudaEventElapsedTime(&c->t_ms[MA_T_CSR], c->ev[MA_T_REDUCE + 1], c->ev[MA_T_CSR + 1]);
      }
    }
    if (c->stats) {
      unsigned long long hc[CNT_N];
      CK(cudaMemcpy(hc, c->counters.p, sizeof hc, cudaMemcpyDeviceToHost));
      for (int k = 0; k < CNT_N; ++k) c->host_counters[k] = (int64_t)hc[k];
    }
    c->have_eval = true;
    c->have_hessian = (MODE == MODE_KANTOROVICH) && with_hessian;
    if (c->warm == 2 && c->kmax == 16 && c->part_n == 1) CKR(save_adjacency(c));
    if (c->trace)
      fprintf(stderr, "[ma] eval kmax=%d attempt=%d total=%.3f prep=%.3f cells=%.3f pieces=%.3f reduce=%.3f csr=%.3f ms  mass_min=%g nnz=%d\n",
              c->kmax, attempt, c->t_ms[MA_T_TOTAL], c->t_ms[MA_T_PREP], c->t_ms[MA_T_CELLS], c->t_ms[MA_T_PIECES],
              c->t_ms[MA_T_REDUCE], c->t_ms[MA_T_CSR], c->mass_min, c->nnz);
    return MA_OK;
  }
  return fail(c, MA_INVALID, "evaluation failed after capacity escalation");
}

}  // namespace

extern "C" int ma_set_partition(ma_ctx *c, int rank, int nranks) {
  if (!c) return MA_INVALID;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, MA_INVALID, "ma_set_partition: bad rank %d of %d", rank, nranks);
  c->part_rank = rank;
  c->part_n = nranks;
  invalidate_eval(c);
  return MA_OK;
}

extern "C" int ma_timer_start(ma_ctx *c) {
  NEED_CTX();
  CK(cudaEventRecord(c->ev_user[0], c->stream));
  return MA_OK;
}
extern "C" int ma_timer_stop(ma_ctx *c, float *ms) {
  NEED_CTX();
  if (!ms) return MA_INVALID;
  CK(cudaEventRecord(c->ev_user[1], c->stream));
  CK(cudaEventSynchronize(c->ev_user[1]));
  CK(cudaEventElapsedTime(ms, c->ev_user[0], c->ev_user[1]));
  return MA_OK;
}

extern "C" int ma_set_weights(ma_ctx *c, const double *w) {
  NEED_CTX_NOWAIT();  // stream-ordered: an evaluation in flight has read its weights by the time this copy runs
  if (!w || c->N < 1) return fail(c, MA_INVALID, "ma_set_weights: bad arguments");
  CK(cudaMemcpyAsync(c->w.p, w, (size_t)c->N * 8, cudaMemcpyHostToDevice, c->stream));
  return MA_OK;
}

extern "C" int ma_evaluate(ma_ctx *c, int with_hessian) {
  NEED_CTX();
  return evaluate_mode<MODE_KANTOROVICH>(c, with_hessian != 0);
}

namespace {
// completes the evaluation ma_evaluate_async left in flight; if its flags ask for anything (a larger capacity class, a
// Hessian that outgrew its arrays) the evaluation is simply done again synchronously
int finish_pending(ma_ctx *c) {
  if (!c->pending) return MA_OK;
  c->pending = false;
  CK(cudaStreamSynchronize(c->stream));
  const int h_flags = c->hs->flags, h_nnz = c->hs->nnz;
  const bool fits = !c->pending_hess || (size_t)h_nnz <= std::min<size_t>(c->col.cap / 4, c->val.cap / 8);
  if ((h_flags & (FLAG_CELL_OVERFLOW | FLAG_PIECE_OVERFLOW | FLAG_KMAX_OVERFLOW | FLAG_STACK_OVERFLOW)) || !fits)
    return evaluate_mode<MODE_KANTOROVICH>(c, c->pending_hess);
  c->cell_fallbacks = c->hs->cell_fallbacks;
  c->fval = c->hs->red[0];
  c->mass_sum = c->hs->red[4];
  c->mass_min = c->hs->red[6];
  if (c->pending_hess) c->nnz = h_nnz;
  c->aborted = false;
  c->have_eval = true;
  c->have_hessian = c->pending_hess;
  return MA_OK;
}
}  // namespace

extern "C" int ma_evaluate_async(ma_ctx *c, int with_hessian) {
  // An evaluation still in flight is NOT waited for: this one is queued behind it on the stream and supersedes it (same
  // buffers, and the scalars the host finally reads are those of the last one queued) — that is what keeps the device
  // busy from one evaluation to the next.
  NEED_CTX_NOWAIT();
  c->want_async = true;
  const int rc = evaluate_mode<MODE_KANTOROVICH>(c, with_hessian != 0);
  c->want_async = false;
  return rc;
}
extern "C" int ma_sync(ma_ctx *c) {
  NEED_CTX();
  return MA_OK;
}

extern "C" int ma_kantorovich(ma_ctx *c, const double *w, double *fval, double *g, int *nnz) {
  NEED_CTX();
  CKR(ma_set_weights(c, w));
  CKR(evaluate_mode<MODE_KANTOROVICH>(c, true));
  if (is_dist(c)) {  // with a communicator the call returns the whole problem's g and h on every rank
    CKR(dist_gather_slices(c, c->mass.p, 8));
    CKR(dist_gather_hessian(c));
  }
  if (fval) *fval = c->fval;
  if (nnz) *nnz = is_dist(c) ? c->nnz_g : c->nnz;
  if (g) {
    CKR(ensure(c, c->cg_out, (size_t)c->N * 8));
    k_scatter_to_caller<<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->mass.as<double>(), c->perm.as<int>(), c->N,
                                                               c->cg_out.as<double>());
    c->launches++;
    CK(cudaMemcpyAsync(g, c->cg_out.p, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return MA_OK;
}

extern "C" int ma_get_hessian_csr(ma_ctx *c, int *rowptr, int *col, double *val) {
  NEED_CTX();
  if (!c->have_hessian) return fail(c, MA_INVALID, "no Hessian: call ma_kantorovich first");
  if (!rowptr || !col || !val) return fail(c, MA_INVALID, "ma_get_hessian_csr: null output");
  const int N = c->N;
  const bool dist = is_dist(c);
  const int nnz_all = dist ? c->nnz_g : c->nnz;
  const int *rowptr_s = dist ? c->rowptr_g.as<int>() : c->rowptr.as<int>();
  const int *col_s = dist ? c->col_g.as<int>() : c->col.as<int>();
  const double *val_s = dist ? c->val_g.as<double>() : c->val.as<double>();
  CKR(ensure(c, c->scratch_i, (size_t)N * 4));
  CKR(ensure(c, c->cptr, (size_t)(N + 1) * 4));
  CKR(ensure(c, c->ccol, (size_t)std::max(nnz_all, 1) * 4));
  CKR(ensure(c, c->cval, (size_t)std::max(nnz_all, 1) * 8));
  k_rowcnt_to_caller<<<cdiv(N, 256), 256, 0, c->stream>>>(c->rowcnt.as<int>(), c->pos.as<int>(), N,
                                                          c->scratch_i.as<int>());
  CKR(scan_i32(c, c->scratch_i.as<int>(), c->cptr.as<int>(), N));
  c->launches += 1;
  // the row pointers first: the host needs them to know which byte ranges of col / val belong to which rows
  CK(cudaMemcpyAsync(rowptr, c->cptr.p, (size_t)(N + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // then chunk by chunk: rows of chunk q are put in caller order on the main stream while the copy engine
  // ships chunk q-1 on the second stream (the transfer is PCIe-bound, the conversion hides behind it)
  const int Q = (nnz_all > (1 << 20)) ? 8 : 1;
  for (int q = 0; q < Q; ++q) {
    const int r0 = (int)((long long)N * q / Q), r1 = (int)((long long)N * (q + 1) / Q);
    if (r1 <= r0) continue;
    k_csr_to_caller<<<cdiv(r1 - r0, 128), 128, 0, c->stream>>>(r0, r1, rowptr_s, col_s,
                                                             val_s, c->pos.as<int>(), c->perm.as<int>(),
                                                             c->cptr.as<int>(), c->ccol.as<int>(), c->cval.as<double>());
    c->launches += 1;
    CK(cudaEventRecord(c->ev_chunk[q], c->stream));
    CK(cudaStreamWaitEvent(c->stream2, c->ev_chunk[q], 0));
    const size_t e0 = (size_t)rowptr[r0], e1 = (size_t)rowptr[r1];
    if (e1 > e0) {
      CK(cudaMemcpyAsync(col + e0, c->ccol.as<int>() + e0, (e1 - e0) * 4, cudaMemcpyDeviceToHost, c->stream2));
      CK(cudaMemcpyAsync(val + e0, c->cval.as<double>() + e0, (e1 - e0) * 8, cudaMemcpyDeviceToHost, c->stream2));
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream2));
  CK(cudaStreamSynchronize(c->stream));
  return MA_OK;
}

// dst[q] = perm[src[q]]
__global__ void k_map_ids(const int *__restrict__ src, const int *__restrict__ perm, int n, int *__restrict__ dst) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) dst[q] = perm[src[q]];
}

extern "C" int ma_get_tile_rows(ma_ctx *c, int *ntile, int *nnz_tile, int *row_ids, double *g, int *rowptr, int *col, double *val) {
  NEED_CTX();
  if (!c->have_eval) return fail(c, MA_INVALID, "no evaluation yet");
  const int lo = (int)((long long)c->N * c->part_rank / c->part_n), hi = (int)((long long)c->N * (c->part_rank + 1) / c->part_n);
  const int nt = hi - lo;
  if (ntile) *ntile = nt;
  if (nnz_tile) *nnz_tile = c->have_hessian ? c->nnz : 0;  // a partitioned context stores exactly its tile's rows
  if (!row_ids && !g && !rowptr && !col && !val) return MA_OK;  // size query
  if (row_ids) CK(cudaMemcpyAsync(row_ids, c->perm.as<int>() + lo, (size_t)nt * 4, cudaMemcpyDeviceToHost, c->stream));
  if (g) CK(cudaMemcpyAsync(g, c->mass.as<double>() + lo, (size_t)nt * 8, cudaMemcpyDeviceToHost, c->stream));
  if (rowptr || col || val) {
    if (!c->have_hessian) return fail(c, MA_INVALID, "no Hessian: evaluate with the Hessian first");
    // rowptr[lo] = 0 on a partitioned context (the rows before the tile are empty): the device array is the answer as it is
    if (rowptr) CK(cudaMemcpyAsync(rowptr, c->rowptr.as<int>() + lo, ((size_t)nt + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    if (col) {
      CKR(ensure(c, c->ccol, (size_t)std::max(c->nnz, 1) * 4));
      k_map_ids<<<cdiv(std::max(c->nnz, 1), 256), 256, 0, c->stream>>>(c->col.as<int>(), c->perm.as<int>(), c->nnz, c->ccol.as<int>());
      c->launches++;
      CK(cudaMemcpyAsync(col, c->ccol.p, (size_t)c->nnz * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (val) CK(cudaMemcpyAsync(val, c->val.p, (size_t)c->nnz * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  if (rowptr && rowptr[0] != 0) {  // part_n == 1 with lo == 0 never gets here; defensive for future layouts
    const int b = rowptr[0];
    for (int k = 0; k <= nt; ++k) rowptr[k] -= b;
  }
  return MA_OK;
}

extern "C" int ma_get_adjacency(ma_ctx *c, int *ptr, int *idx, int capacity) {
  NEED_CTX();
  if (!c->have_eval) return fail(c, MA_INVALID, "no evaluation yet");
  const int N = c->N;
  CKR(ensure(c, c->scratch_i, (size_t)N * 4));
  CKR(ensure(c, c->cptr, (size_t)(N + 1) * 4));
  k_adj_count<<<cdiv(N, 256), 256, 0, c->stream>>>(N, c->nbr_cnt.as<int>(), c->pos.as<int>(), c->scratch_i.as<int>());
  CKR(scan_i32(c, c->scratch_i.as<int>(), c->cptr.as<int>(), N));
  int total = 0;
  CK(cudaMemcpyAsync(&total, c->cptr.as<int>() + N, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(ptr, c->cptr.p, (size_t)(N + 1) * 4, cudaMemcpyDeviceToHost));
  if (!idx) return MA_OK;  // size query: ptr[N] is the number of entries
  if (capacity < total) return fail(c, MA_INVALID, "adjacency needs %d entries", total);
  CKR(ensure(c, c->ccol, (size_t)std::max(total, 1) * 4));
  k_adj_fill<<<cdiv(N, 256), 256, 0, c->stream>>>(N, c->kmax, c->nbr.as<int>(), c->nbr_cnt.as<int>(), c->pos.as<int>(),
                                                  c->perm.as<int>(), c->cptr.as<int>(), c->ccol.as<int>());
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(idx, c->ccol.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MA_OK;
}

// =============================================================================================
// moments / lloyd
// =============================================================================================
namespace {
// moments of all cells in internal order (6 per cell) -> the caller's ordering, column-major like Eigen's N x 2 / N x 3
// (lloyd.hpp:30-123); centroids = true divides the first moments by the mass (lloyd.hpp:139-143)
__global__ void k_moments_to_caller(int N, const int *__restrict__ perm, const double *__restrict__ mom, int order, bool centroids,
                                    double *__restrict__ out_m, double *__restrict__ out_1, double *__restrict__ out_2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = perm[k];
  const double *o = mom + 6 * (size_t)k;
  const double m = o[0];
  out_m[i] = m;
  out_1[i] = centroids ? o[1] / m : o[1];
  out_1[(size_t)N + i] = centroids ? o[2] / m : o[2];
  if (order == 2) { out_2[i] = o[3]; out_2[(size_t)N + i] = o[4]; out_2[2 * (size_t)N + i] = o[5]; }
}

int moments_impl(ma_ctx *c, const double *w, int order, double *masses, double *m1, double *m2, bool centroids) {
  CKR(ma_set_weights(c, w));
  if (order == 1) CKR(evaluate_mode<MODE_MOMENTS1>(c, false));
  else CKR(evaluate_mode<MODE_MOMENTS2>(c, false));
  if (is_dist(c)) CKR(dist_gather_slices(c, c->mom.p, 48));  // every rank returns the moments of ALL cells
  const size_t N = c->N;
  CKR(ensure(c, c->scratch_d, 6 * N * 8));
  double *om = c->scratch_d.as<double>(), *o1 = om + N, *o2 = om + 3 * N;
  k_moments_to_caller<<<cdiv((int)N, 256), 256, 0, c->stream>>>((int)N, c->perm.as<int>(), c->mom.as<double>(), order, centroids, om, o1, o2);
  c->launches++;
  CK(cudaGetLastError());
  if (masses) CK(cudaMemcpyAsync(masses, om, N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (m1) CK(cudaMemcpyAsync(m1, o1, 2 * N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (order == 2 && m2) CK(cudaMemcpyAsync(m2, o2, 3 * N * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MA_OK;
}
}  // namespace

extern "C" int ma_moments(ma_ctx *c, const double *w, int order, double *masses, double *m1, double *m2) {
  NEED_CTX();
  if (order != 1 && order != 2) return fail(c, MA_INVALID, "ma_moments: order must be 1 or 2");
  if (order == 2 && !m2) return fail(c, MA_INVALID, "ma_moments: m2 is required for order 2");
  if (c->part_n > 1 && !c->comm)
    return fail(c, MA_INVALID, "ma_moments on a partitioned context needs a communicator (ma_comm_init): it returns all cells");
  return moments_impl(c, w, order, masses, m1, m2, false);
}

extern "C" int ma_lloyd(ma_ctx *c, const double *w, double *masses, double *centroids) {
  NEED_CTX();
  if (!masses || !centroids) return fail(c, MA_INVALID, "ma_lloyd: null output");
  if (c->part_n > 1 && !c->comm)
    return fail(c, MA_INVALID, "ma_lloyd on a partitioned context needs a communicator (ma_comm_init): it returns all cells");
  return moments_impl(c, w, 1, masses, centroids, nullptr, true);  // lloyd.hpp:139-143: centroid = first moment / mass
}

// =============================================================================================
// pieces
// =============================================================================================
extern "C" int ma_pieces_build(ma_ctx *c, const double *w, int *npieces, int *nvertices) {
  NEED_CTX();
  if (c->mesh_kind == MESH_NONE || c->N < 1) return fail(c, MA_INVALID, "mesh/points not set");
  if (c->part_n > 1) return fail(c, MA_INVALID, "ma_pieces_build: not available on a partitioned context (ma_set_partition)");
  CKR(ma_set_weights(c, w));
  for (int attempt = 0; attempt < 3; ++attempt) {
    CKR(alloc_eval(c));
    const int N = c->N;
    CKR(ensure(c, c->pc_count, (size_t)N * 8));
    CKR(ensure(c, c->pc_off, ((size_t)2 * N + 1) * 4));
    Params p;
    fill_params(c, p);
    p.stats = 0;
    p.pc_count = c->pc_count.as<int>();
    CK(cudaMemsetAsync(c->flags.p, 0, 16, c->stream));
    CKR(run_cells(c, p));
    CKR(launch_pieces_mode<MODE_PIECES_COUNT>(c, p));
    int h_flags = 0;
    CK(cudaMemcpyAsync(&h_flags, c->flags.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h_flags & (FLAG_CELL_OVERFLOW | FLAG_PIECE_OVERFLOW | FLAG_KMAX_OVERFLOW)) {
      if (c->kmax >= 64) return fail(c, MA_INVALID, "polygon capacity exceeded");
      c->kmax *= 2;
      continue;
    }
    // interleaved (pieces, vertices) counts -> two exclusive scans done on the host (output stage)
    std::vector<int> cnt((size_t)2 * N), off((size_t)2 * N);
    CK(cudaMemcpy(cnt.data(), c->pc_count.p, cnt.size() * 4, cudaMemcpyDeviceToHost));
    long long np = 0, nv = 0;
    for (int i = 0; i < N; ++i) {
      off[2 * (size_t)i] = (int)np; off[2 * (size_t)i + 1] = (int)nv;
      np += cnt[2 * (size_t)i]; nv += cnt[2 * (size_t)i + 1];
    }
    if (nv > 0x7fffffffll) return fail(c, MA_INVALID, "too many piece vertices");
    c->pc_np = (int)np; c->pc_nv = (int)nv;
    CK(cudaMemcpy(c->pc_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CKR(ensure(c, c->pc_cell, (size_t)std::max(c->pc_np, 1) * 4));
    CKR(ensure(c, c->pc_face, (size_t)std::max(c->pc_np, 1) * 4));
    CKR(ensure(c, c->pc_ptr, ((size_t)c->pc_np + 1) * 4));
    CKR(ensure(c, c->pc_tag, (size_t)std::max(c->pc_nv, 1) * 4));
    CKR(ensure(c, c->pc_xy, (size_t)std::max(c->pc_nv, 1) * 16));
    p.pc_off = c->pc_off.as<int>();
    p.pc_cell = c->pc_cell.as<int>(); p.pc_face = c->pc_face.as<int>(); p.pc_ptr = c->pc_ptr.as<int>();
    p.pc_tag = c->pc_tag.as<int>(); p.pc_xy = c->pc_xy.as<double>();
    CKR(launch_pieces_mode<MODE_PIECES_FILL>(c, p));
    CK(cudaMemcpyAsync(c->pc_ptr.as<int>() + c->pc_np, &c->pc_nv, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (npieces) *npieces = c->pc_np;
    if (nvertices) *nvertices = c->pc_nv;
    c->have_eval = true;
    return MA_OK;
  }
  return fail(c, MA_INVALID, "pieces failed after capacity escalation");
}

extern "C" int ma_pieces_get(ma_ctx *c, int *cell, int *face, int *ptr, int *tag, double *xy) {
  NEED_CTX();
  const int np = c->pc_np, nv = c->pc_nv;
  std::vector<int> perm(c->N), hc(std::max(np, 1)), ht(std::max(nv, 1));
  CK(cudaMemcpy(perm.data(), c->perm.p, (size_t)c->N * 4, cudaMemcpyDeviceToHost));
  if (np) {
    CK(cudaMemcpy(hc.data(), c->pc_cell.p, (size_t)np * 4, cudaMemcpyDeviceToHost));
    if (face) CK(cudaMemcpy(face, c->pc_face.p, (size_t)np * 4, cudaMemcpyDeviceToHost));
  }
  if (ptr) CK(cudaMemcpy(ptr, c->pc_ptr.p, ((size_t)np + 1) * 4, cudaMemcpyDeviceToHost));
  if (nv) {
    CK(cudaMemcpy(ht.data(), c->pc_tag.p, (size_t)nv * 4, cudaMemcpyDeviceToHost));
    if (xy) CK(cudaMemcpy(xy, c->pc_xy.p, (size_t)nv * 16, cudaMemcpyDeviceToHost));
  }
  if (cell) for (int k = 0; k < np; ++k) cell[k] = perm[hc[k]];
  if (tag) for (int k = 0; k < nv; ++k) tag[k] = ht[k] >= 0 ? perm[ht[k]] : ht[k];
  return MA_OK;
}

// draw_laguerre_diagram (rasterization.hpp:512-547)
extern "C" int ma_draw_laguerre_diagram(ma_ctx *c, const double *weights, const double *colors, double x0, double y0, double x1,
                                        double y1, int w, int h, double *image) {
  NEED_CTX();
  if (!weights || !colors || !image || w < 1 || h < 1 || !(x1 > x0) || !(y1 > y0))
    return fail(c, MA_INVALID, "ma_draw_laguerre_diagram: bad arguments");
  int np = 0, nv = 0;
  CKR(ma_pieces_build(c, weights, &np, &nv));
  Buf b_col, b_img;
  int rc = MA_OK;
  do {
    if ((rc = upload(c, b_col, colors, (size_t)c->N * 8))) break;
    if ((rc = ensure(c, b_img, (size_t)w * h * 8))) break;
    if (cudaMemsetAsync(b_img.p, 0, (size_t)w * h * 8, c->stream) != cudaSuccess) { rc = fail(c, MA_CUDA_ERROR, "memset failed"); break; }
    if (np > 0)
      k_raster_pieces<<<cdiv(np, 128), 128, 0, c->stream>>>(np, c->pc_cell.as<int>(), c->pc_face.as<int>(), c->pc_ptr.as<int>(),
                                                           c->pc_xy.as<double>(), c->abc.as<double>(), c->perm.as<int>(),
                                                           b_col.as<double>(), x0, y0, (double)w / (x1 - x0), (double)h / (y1 - y0), w, h,
                                                           b_img.as<double>());
    c->launches++;
    if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(image, b_img.p, (size_t)w * h * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      rc = fail(c, MA_CUDA_ERROR, "rasterisation failed");
  } while (0);
  release(b_col); release(b_img);
  return rc;
}

// =============================================================================================
// Laguerre cells (cell ∩ mesh bounding box) as polygons: K1 + K2 only
// =============================================================================================
extern "C" int ma_cells_build(ma_ctx *c, const double *w, int *nvertices) {
  NEED_CTX();
  if (c->mesh_kind == MESH_NONE || c->N < 1) return fail(c, MA_INVALID, "mesh/points not set");
  CKR(ma_set_weights(c, w));
  c->kmax = c->kmax_base;
  for (int attempt = 0; attempt < 3; ++attempt) {
    CKR(alloc_eval(c));
    const size_t slots = (size_t)cells_maxv(c->kmax) * c->N;
    CKR(ensure(c, c->poly_x, slots * 8)); CKR(ensure(c, c->poly_y, slots * 8));
    CKR(ensure(c, c->poly_t, slots * 4)); CKR(ensure(c, c->poly_n, (size_t)c->N * 4));
    Params p;
    c->abort_on_empty = c->probe_empty;
    fill_params(c, p);
    c->abort_on_empty = false;
    p.stats = 0;
    CK(cudaMemsetAsync(c->flags.p, 0, 16, c->stream));
    if (c->part_n > 1) CK(cudaMemsetAsync(c->poly_n.p, 0, (size_t)c->N * 4, c->stream));  // other tiles: no polygon here
    CKR(run_cells<true>(c, p));
    CK(cudaMemcpyAsync(&c->hs->flags, c->flags.p, 8, cudaMemcpyDeviceToHost, c->stream));  // flags[0], flags[1]
    CK(cudaStreamSynchronize(c->stream));
    if (c->hs->flags & (FLAG_CELL_OVERFLOW | FLAG_KMAX_OVERFLOW)) {  // before the abort flag: an overflow is not an empty cell
      if (c->kmax >= 64) return fail(c, MA_INVALID, "polygon capacity exceeded");
      c->kmax *= 2;
      continue;
    }
    c->aborted = c->probe_empty && c->hs->abort_ != 0;
    if (c->aborted) {  // a cell of this tile is empty (the caller only wanted to know): no polygons
      invalidate_eval(c);
      if (nvertices) *nvertices = 0;
      return MA_OK;
    }
    if (c->hs->flags & FLAG_STACK_OVERFLOW) return fail(c, MA_INVALID, "quadtree stack overflow");
    const int N = c->N;
    CKR(ensure(c, c->scratch_i, (size_t)N * 4));
    CKR(ensure(c, c->cptr, (size_t)(N + 1) * 4));
    k_cellpoly_count<<<cdiv(N, 256), 256, 0, c->stream>>>(N, c->poly_n.as<int>(), c->pos.as<int>(), c->scratch_i.as<int>());
    CKR(scan_i32(c, c->scratch_i.as<int>(), c->cptr.as<int>(), N));
    CK(cudaMemcpyAsync(&c->hs->nnz, c->cptr.as<int>() + N, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->cells_nv = c->hs->nnz;
    if (nvertices) *nvertices = c->cells_nv;
    c->have_eval = true;
    c->have_hessian = false;
    return MA_OK;
  }
  return fail(c, MA_INVALID, "cells failed after capacity escalation");
}

extern "C" int ma_cells_get(ma_ctx *c, int *ptr, double *xy, int *tag) {
  NEED_CTX();
  if (c->cells_nv < 0) return fail(c, MA_INVALID, "no cells: call ma_cells_build first");
  if (!ptr || !xy || !tag) return fail(c, MA_INVALID, "ma_cells_get: null output");
  const int N = c->N, nv = c->cells_nv;
  CKR(ensure(c, c->pc_xy, (size_t)std::max(nv, 1) * 16));
  CKR(ensure(c, c->pc_tag, (size_t)std::max(nv, 1) * 4));
  k_cellpoly_fill<<<cdiv(N, 128), 128, 0, c->stream>>>(N, c->poly_n.as<int>(), c->poly_x.as<double>(), c->poly_y.as<double>(),
                                                      c->poly_t.as<int>(), c->xs.as<double>(), c->ys.as<double>(),
                                                      c->pos.as<int>(), c->perm.as<int>(), c->cptr.as<int>(),
                                                      c->pc_xy.as<double>(), c->pc_tag.as<int>());
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ptr, c->cptr.p, (size_t)(N + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (nv) {
    CK(cudaMemcpyAsync(xy, c->pc_xy.p, (size_t)nv * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tag, c->pc_tag.p, (size_t)nv * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return MA_OK;
}

// =============================================================================================
// K5: PCG + Newton
// =============================================================================================
namespace {

// ---- aggregation multigrid (ma_amg.cuh) -------------------------------------------------------------------
// hierarchy of aggregates from the Morton-sorted leaf codes: once per point set
int amg_build_hierarchy(ma_ctx *c) {
  AmgHost &A = c->amg;
  A.nlev = 0;
  A.ready = false;
  const int N = c->N;
  if (N < 2 * AMG_DENSE_MAX) return MA_OK;  // small systems: Jacobi PCG is fine
  int n = N;
  const unsigned *code = c->code_s.as<unsigned>();
  for (int l = 0; l < AMG_MAX_LEVELS; ++l) {
    A.n[l] = n;
    A.nlev = l + 1;
    if (n <= AMG_DENSE_MAX) break;
    CKR(ensure(c, A.agg[l], (size_t)n * 4));
    CKR(ensure(c, c->scratch_i, (size_t)n * 4));
    CKR(ensure(c, A.excl, ((size_t)n + 1) * 4));
    k_amg_flags<<<cdiv(n, 256), 256, 0, c->stream>>>(code, n, c->scratch_i.as<int>());
    CKR(scan_i32(c, c->scratch_i.as<int>(), A.excl.as<int>(), n));
    int nc = 0;
    CK(cudaMemcpyAsync(&nc, A.excl.as<int>() + n, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nc >= n || nc < 1) { A.nlev = 0; return MA_OK; }  // no coarsening possible (cannot happen: the root merges everything)
    CKR(ensure(c, A.cstart[l], ((size_t)nc + 1) * 4));
    CKR(ensure(c, A.code[l + 1], (size_t)nc * 4));
    k_amg_index<<<cdiv(n, 256), 256, 0, c->stream>>>(code, c->scratch_i.as<int>(), A.excl.as<int>(), n, A.agg[l].as<int>(),
                                                     A.cstart[l].as<int>(), A.code[l + 1].as<unsigned>());
    CK(cudaGetLastError());
    code = A.code[l + 1].as<unsigned>();
    n = nc;
  }
  if (A.n[A.nlev - 1] > AMG_DENSE_MAX) { A.nlev = 0; return MA_OK; }
  for (int l = 0; l < A.nlev; ++l) {
    const size_t nl = A.n[l];
    if (l > 0) { CKR(ensure(c, A.r[l], nl * 8)); CKR(ensure(c, A.dinv[l], nl * 8)); CKR(ensure(c, A.rowptr[l], (nl + 1) * 4)); }
    CKR(ensure(c, A.x[l], nl * 8)); CKR(ensure(c, A.t[l], nl * 8)); CKR(ensure(c, A.x2[l], nl * 8));
  }
  const size_t nd = A.n[A.nlev - 1];
  CKR(ensure(c, A.W, nd * 2 * nd * 8)); CKR(ensure(c, A.Ainv, nd * nd * 8)); CKR(ensure(c, A.flag, 16));
  CK(cudaStreamSynchronize(c->stream));
  A.ready = true;
  return MA_OK;
}

// coarse matrices of this solve: Galerkin products level by level + the dense inverse of the last one
int amg_setup(ma_ctx *c, const int *rowptr, const int *col, const double *val, int nnz, const double *dinv0, int ground) {
  AmgHost &A = c->amg;
  CK(cudaMemsetAsync(A.flag.p, 0, 16, c->stream));
  A.lev[0].rowptr = rowptr; A.lev[0].col = col; A.lev[0].val = val;
  size_t cap = (size_t)std::max(nnz, 1);
  for (int l = 0; l < A.nlev; ++l) {
    AmgLevel &L = A.lev[l];
    L.n = A.n[l];
    L.agg = l + 1 < A.nlev ? A.agg[l].as<int>() : nullptr;
    L.cstart = l + 1 < A.nlev ? A.cstart[l].as<int>() : nullptr;
    L.dinv = l == 0 ? const_cast<double *>(dinv0) : A.dinv[l].as<double>();
    L.x = A.x[l].as<double>(); L.t = A.t[l].as<double>(); L.x2 = A.x2[l].as<double>();
    L.r = l == 0 ? nullptr : A.r[l].as<double>();
  }
  for (int l = 0; l + 1 < A.nlev; ++l) {
    const AmgLevel &F = A.lev[l];
    AmgLevel &C = A.lev[l + 1];
    const int nc = A.n[l + 1];
    cap = std::min(cap, (size_t)nc * AMG_ROW_CAP);
    CKR(ensure(c, A.col[l + 1], cap * 4, 1.2)); CKR(ensure(c, A.val[l + 1], cap * 8, 1.2));
    CKR(ensure(c, c->scratch_i, (size_t)nc * 4));
    const int g = l == 0 ? ground : -1;
    k_amg_galerkin<false><<<cdiv(nc, 128), 128, 0, c->stream>>>(nc, F.cstart, F.agg, F.rowptr, F.col, F.val, g, c->scratch_i.as<int>(),
                                                                nullptr, nullptr, nullptr, nullptr, A.flag.as<int>());
    CKR(scan_i32(c, c->scratch_i.as<int>(), A.rowptr[l + 1].as<int>(), nc));
    k_amg_galerkin<true><<<cdiv(nc, 128), 128, 0, c->stream>>>(nc, F.cstart, F.agg, F.rowptr, F.col, F.val, g, nullptr,
                                                               A.rowptr[l + 1].as<int>(), A.col[l + 1].as<int>(),
                                                               A.val[l + 1].as<double>(), A.dinv[l + 1].as<double>(), A.flag.as<int>());
    c->launches += 2;
    C.rowptr = A.rowptr[l + 1].as<int>(); C.col = A.col[l + 1].as<int>(); C.val = A.val[l + 1].as<double>();
  }
  const AmgLevel &D = A.lev[A.nlev - 1];
  k_amg_dense_inverse<<<1, 1024, 0, c->stream>>>(D.n, D.rowptr, D.col, D.val, A.W.as<double>(), A.Ainv.as<double>(), A.flag.as<int>());
  c->launches++;
  CK(cudaGetLastError());
  return MA_OK;
}

// z = M^-1 r (one V(1,1) cycle); leaves the partial sums of r.z in part_rz[0 .. nblocks)
int amg_vcycle(ma_ctx *c, const double *r, double *z, double *part_rz, int nblocks, int ground) {
  AmgHost &A = c->amg;
  const double omega = c->amg_omega, alpha = c->amg_alpha;
  auto grid = [&](int n) { return std::max(1, std::min(cdiv(n, 256), c->sm_count * 8)); };
  const int last = A.nlev - 1;
  // the levels from `tail` on run in one block (k_amg_tail); tail == last: only the dense level is left for it
  int tail = last;
  while (tail > 1 && A.n[tail - 1] <= AMG_TAIL_ROWS && last - (tail - 1) < AMG_TAIL_MAX) --tail;
  for (int l = 0; l < tail; ++l) {
    const AmgLevel &L = A.lev[l];
    const double *rl = l == 0 ? r : L.r;
    k_amg_down<<<grid(L.n), 256, 0, c->stream>>>(L.n, L.rowptr, L.col, L.val, L.dinv, rl, omega, l == 0 ? ground : -1, L.x, L.t);
    k_amg_restrict<<<grid(A.n[l + 1]), 256, 0, c->stream>>>(A.n[l + 1], L.cstart, L.t, A.lev[l + 1].r);
  }
  {
    AmgTail T;
    T.nlev = last - tail + 1;
    for (int l = tail; l <= last; ++l) T.lev[l - tail] = A.lev[l];
    T.Ainv = A.Ainv.as<double>();
    k_amg_tail<<<AMG_TAIL_CTAS, 1024, 0, c->stream>>>(T, omega, alpha);  // one cluster
  }
  for (int l = tail - 1; l >= 0; --l) {
    const AmgLevel &L = A.lev[l];
    const double *ec = A.lev[l + 1].x2;
    if (l == 0)
      k_amg_up<true><<<nblocks, 256, 0, c->stream>>>(L.n, L.rowptr, L.col, L.val, L.dinv, r, L.x, L.agg, ec, alpha, omega, ground, z, part_rz);
    else
      k_amg_up<false><<<grid(L.n), 256, 0, c->stream>>>(L.n, L.rowptr, L.col, L.val, L.dinv, L.r, L.x, L.agg, ec, alpha, omega, -1, L.x2, nullptr);
  }
  c->launches += 3 * tail + 1;
  CK(cudaGetLastError());
  return MA_OK;
}

// Solve H d = sign * g on the device (internal order), grounded at `ground`.  d_out may alias nothing.
// use_amg: precondition with the aggregation multigrid of ma_amg.cuh (needs the hierarchy of this context's point set,
// i.e. the matrix must be a Hessian of these Diracs in internal order); otherwise Jacobi.
int pcg_solve(ma_ctx *c, int n, const int *rowptr, const int *col, const double *val, const double *g, double sign,
              int ground, double *d_out, int *iters_out, double *relres_out, const double *x0 = nullptr,
              double x0_scale = 1.0, bool use_amg = false, int nnz = 0) {
  static_assert(PCG_NT == 256, "k_amg_up<true> shares the r.z partials with the PCG kernels: same block size");
  PcgState s;
  if (use_amg && c->amg_on && c->amg_stale) {
    CKR(amg_build_hierarchy(c));
    c->amg_stale = false;
  }
  use_amg = use_amg && c->amg_on && c->amg.ready && c->amg.n[0] == n;
  // one resident wave at most (the kernels are grid-stride loops): a second, partial wave only adds a tail
  int per_sm = 8;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_A, PCG_NT, 0));
  const int nblocks = std::min(std::min(PCG_MAX_BLOCKS, std::max(1, per_sm) * c->sm_count), std::max(1, cdiv(n, PCG_NT)));
  CKR(ensure(c, c->dinv, (size_t)n * 8)); CKR(ensure(c, c->cgx, (size_t)n * 8)); CKR(ensure(c, c->cgr, (size_t)n * 8));
  CKR(ensure(c, c->cgz, (size_t)n * 8)); CKR(ensure(c, c->cgp0, (size_t)n * 8)); CKR(ensure(c, c->cgp1, (size_t)n * 8));
  CKR(ensure(c, c->cgq, (size_t)n * 8));
  CKR(ensure(c, c->part_pq, (size_t)6 * PCG_MAX_BLOCKS * 8)); CKR(ensure(c, c->part_rz, (size_t)2 * nblocks * 8));
  CKR(ensure(c, c->part_rr, (size_t)nblocks * 8)); CKR(ensure(c, c->scal, 64)); CKR(ensure(c, c->cgflag, 16));
  s.n = n; s.ground = ground; s.rowptr = rowptr; s.col = col; s.val = val;
  s.dinv = c->dinv.as<double>(); s.x = c->cgx.as<double>(); s.r = c->cgr.as<double>(); s.z = c->cgz.as<double>();
  s.p[0] = c->cgp0.as<double>(); s.p[1] = c->cgp1.as<double>(); s.q = c->cgq.as<double>();
  s.part_pq = c->part_pq.as<double>(); s.part_rz = c->part_rz.as<double>(); s.part_rr = c->part_rr.as<double>();
  s.scal = c->scal.as<double>(); s.nblocks = nblocks; s.flag = c->cgflag.as<int>();
  double h_scal[4];
  int h_flag = 0, h_amg = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CK(cudaMemsetAsync(c->cgflag.p, 0, 16, c->stream));
    CK(cudaMemsetAsync(c->part_rz.p, 0, (size_t)2 * nblocks * 8, c->stream));
    k_pcg_init<<<nblocks, PCG_NT, 0, c->stream>>>(s, g, sign, x0, x0_scale);
    if (x0) { k_pcg_resid<<<nblocks, PCG_NT, 0, c->stream>>>(s); c->launches++; }
    if (use_amg) {
      CKR(amg_setup(c, rowptr, col, val, nnz, s.dinv, ground));
      CKR(amg_vcycle(c, s.r, s.z, s.part_rz, nblocks, ground));
      CK(cudaMemcpyAsync(&h_amg, c->amg.flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
    }
    k_pcg_init2<<<1, PCG_NT, 0, c->stream>>>(s);
    c->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_scal, c->scal.p, sizeof h_scal, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&h_flag, c->cgflag.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (use_amg && h_amg) {  // a coarse row outgrew its merge buffer / a non-positive pivot: this solve runs with Jacobi
      if (c->trace) fprintf(stderr, "[ma] multigrid setup flag %d: falling back to the Jacobi preconditioner\n", h_amg);
      use_amg = false;
      continue;
    }
    break;
  }
  if (h_flag & 1) {
    fail(c, MA_SINGULAR_HESSIAN, "Error: hessian of Kantorovich's functional is not invertible (zero diagonal)");
    return MA_SINGULAR_HESSIAN;
  }
  const double gg = h_scal[2];
  int it = 0;
  double rr = gg;
  if (gg > 0) {
    const int batch = use_amg ? 8 : 32;  // even, so iteration parity is the same in every graph launch
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const long long l0 = c->launches;
    for (int b = 0; b < batch; ++b) {
      // iteration 0 of the very first batch must use beta = 0: p_old is zero-initialised and
      // part_rz parity 0 is zero, so beta = 0/rz = 0 falls out without a special case
      k_pcg_A<<<nblocks, PCG_NT, 0, c->stream>>>(s, b + 2);
      if (use_amg) {
        k_pcg_B<false><<<nblocks, PCG_NT, 0, c->stream>>>(s, b + 2);
        int rc_v = amg_vcycle(c, s.r, s.z, s.part_rz + (size_t)(((b + 2) & 1) ^ 1) * nblocks, nblocks, ground);
        if (rc_v != MA_OK) { cudaStreamEndCapture(c->stream, &graph); if (graph) cudaGraphDestroy(graph); return rc_v; }
      } else {
        k_pcg_B<true><<<nblocks, PCG_NT, 0, c->stream>>>(s, b + 2);
      }
    }
    k_pcg_rr<<<1, PCG_NT, 0, c->stream>>>(s);
    const long long per_graph = (c->launches - l0) + 2 * batch + 1;
    c->launches = l0;
    CK(cudaStreamEndCapture(c->stream, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    const double tol2 = c->cg_rtol * c->cg_rtol * gg;
    while (it < c->cg_maxit) {
      CK(cudaGraphLaunch(exec, c->stream));
      c->launches += per_graph;
      CK(cudaMemcpyAsync(&rr, c->scal.as<double>() + 3, 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      it += batch;
      if (!(rr > tol2)) break;  // also stops on NaN
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
  }
  if (d_out) CK(cudaMemcpyAsync(d_out, s.x, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = gg > 0 ? std::sqrt(rr / gg) : 0.0;
  c->last_cg_iters = it;
  if (!(rr == rr)) return fail(c, MA_LINSOLVE_RESIDUAL, "PCG produced NaN");
  return MA_OK;
}

}  // namespace

extern "C" int ma_solve_laplacian(ma_ctx *c, int N, const int *rowptr, const int *col, const double *val,
                                  const double *g, double *d, int *iters) {
  NEED_CTX();
  if (N < 1 || !rowptr || !col || !val || !g || !d) return fail(c, MA_INVALID, "ma_solve_laplacian: bad arguments");
  const int nnz = rowptr[N];
  Buf b_ptr, b_col, b_val, b_g, b_d;
  int rc = MA_OK;
  do {
    if ((rc = upload(c, b_ptr, rowptr, (size_t)(N + 1) * 4))) break;
    if ((rc = upload(c, b_col, col, (size_t)std::max(nnz, 1) * 4))) break;
    if ((rc = upload(c, b_val, val, (size_t)std::max(nnz, 1) * 8))) break;
    if ((rc = upload(c, b_g, g, (size_t)N * 8))) break;
    if ((rc = ensure(c, b_d, (size_t)N * 8))) break;
    int it = 0;
    double relres = 0;
    rc = pcg_solve(c, N, b_ptr.as<int>(), b_col.as<int>(), b_val.as<double>(), b_g.as<double>(), 1.0, N - 1,
                   b_d.as<double>(), &it, &relres);
    if (iters) *iters = it;
    if (rc != MA_OK && rc != MA_SINGULAR_HESSIAN) break;
    if (rc == MA_SINGULAR_HESSIAN) { std::fill(d, d + N, 0.0); break; }
    // optimal_transport.hpp:68-79: the reference tests the ABSOLUTE residual |hs ds - gs| > 1e-7 and then turns to
    // a second solver (SPQR).  Here the second stage is a restart of the PCG from the first answer (the residual
    // is recomputed from scratch, which removes the drift of the recurrence), and the warning is issued only if
    // that fails too.
    double gnorm = 0;
    for (int i = 0; i + 1 < N; ++i) gnorm += g[i] * g[i];
    gnorm = std::sqrt(gnorm);
    if (relres * gnorm > 1e-7) {
      Buf b_x0;
      if ((rc = ensure(c, b_x0, (size_t)N * 8))) break;
      cudaMemcpyAsync(b_x0.p, b_d.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream);
      int it2 = 0;
      rc = pcg_solve(c, N, b_ptr.as<int>(), b_col.as<int>(), b_val.as<double>(), b_g.as<double>(), 1.0, N - 1,
                     b_d.as<double>(), &it2, &relres, b_x0.as<double>(), 1.0);
      release(b_x0);
      if (iters) *iters = it + it2;
      if (rc != MA_OK) break;
    }
    if (cudaMemcpyAsync(d, b_d.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
      rc = fail(c, MA_CUDA_ERROR, "copy of the solution failed");
      break;
    }
    if (relres * gnorm > 1e-7 && relres * 0 == 0) {
      rc = fail(c, MA_LINSOLVE_RESIDUAL, "WARNING: in solve_laplacian_matrix: err=%g after %d iterations",
                relres * gnorm, it);
    }
  } while (0);
  release(b_ptr); release(b_col); release(b_val); release(b_g); release(b_d);
  return rc;
}

namespace {
// true in *empty if, at the weights now in c->w, some cell is certainly empty (k_cells_quick_empty)
int quick_empty(ma_ctx *c, bool *empty) {
  *empty = false;
  if (!c->quick_reject || c->prev_stride == 0) return MA_OK;
  CKR(alloc_eval(c));
  Params p;
  fill_params(c, p);
  CK(cudaMemsetAsync(c->flags.p, 0, 16, c->stream));
  k_gather_w<<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->w.as<double>(), c->perm.as<int>(), c->s2rm.as<int>(), c->N,
                                                    c->ws.as<double>(), c->wr.as<double>());
  CKR(reduce4(c, c->ws.as<double>(), nullptr, c->N, c->wstat.as<double>()));  // (CellSearch::init reads the weight range)
  constexpr int NT = 128;
  const size_t sm = cells_smem_bytes<16, NT>();
  CK(cudaFuncSetAttribute(k_cells_quick_empty<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_cells_quick_empty<NT><<<std::max(1, cdiv(p.cell_hi - p.cell_lo, NT)), NT, sm, c->stream>>>(p, c->nbr_prev.as<int>(), c->cnt_prev.as<int>(), c->prev_stride);
  c->launches += 2;
  CK(cudaGetLastError());
  CKR(dist_sync_flags(c));
  CK(cudaMemcpyAsync(&c->hs->flags, c->flags.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *empty = c->hs->abort_ != 0;
  return MA_OK;
}
}  // namespace

extern "C" int ma_ot_solve(ma_ctx *c, const double *nu, double *w, int have_initial, double eps_g, size_t maxiter,
                           int verbose, ma_statistics *stats) {
  NEED_CTX();
  if (!nu || !w) return fail(c, MA_INVALID, "ma_ot_solve: null argument");
  if (c->mesh_kind == MESH_NONE || c->N < 1) return fail(c, MA_INVALID, "mesh/points not set");
  if (c->part_n > 1 && !c->comm)
    return fail(c, MA_INVALID, "ma_ot_solve on a partitioned context needs a communicator (ma_set_comm): the Newton "
                               "loop reduces over all tiles");
  auto t0 = std::chrono::steady_clock::now();
  const int N = c->N;
  size_t neval = 0, niter = 0, cg_total = 0;
  struct NewtonScope {  // evaluations inside this call may take the warm path of K2
    ma_ctx *c;
    int rank, n;
    explicit NewtonScope(ma_ctx *c_) : c(c_), rank(c_->part_rank), n(c_->part_n) {
      c->in_newton = true; c->warm_skip = 0; c->warm_penalty = 4;
      // With a communicator the loop runs REPLICATED on every rank unless "newton_sharded" is set: after the first
      // evaluation K2 takes the warm path (whose certificate covers all cells on every rank anyway) and most trials end at
      // the quick empty-cell test, so what sharding saves is K3 on a few dozen accepted points (~0.1 s at 1 M Diracs) while
      // it adds 4 small collectives per trial and the gather of gradient / Hessian per accepted point (~0.4 s on 2 GPUs,
      // ~1 s on 8, profiles/r02v).  The solve is replicated in both variants; all ranks return the same bits.
      if (c->comm && c->part_n > 1 && !c->newton_sharded) { c->part_rank = 0; c->part_n = 1; }
    }
    ~NewtonScope() {
      c->in_newton = false;
      if (c->part_n != n || c->part_rank != rank) {  // the loop ran replicated: back to this rank's tile, whose results are not the last evaluation's
        c->part_rank = rank; c->part_n = n;
        invalidate_eval(c);
      }
    }
  } newton_scope(c);
  CKR(ensure(c, c->nu_s, (size_t)N * 8)); CKR(ensure(c, c->x0_s, (size_t)N * 8));
  CKR(ensure(c, c->d_s, (size_t)N * 8)); CKR(ensure(c, c->g_s, (size_t)N * 8));
  CKR(ensure(c, c->scratch_d, (size_t)N * 8));
  CKR(ensure(c, c->red_out, 16 * sizeof(double)));
  // nu in caller order on the device (scratch_d), then sorted copy nu_s
  CK(cudaMemcpyAsync(c->scratch_d.p, nu, (size_t)N * 8, cudaMemcpyHostToDevice, c->stream));
  k_gather<<<cdiv(N, 256), 256, 0, c->stream>>>(c->scratch_d.as<double>(), c->perm.as<int>(), N, c->nu_s.as<double>());
  // x: caller-order device weights live in c->w (what the evaluation reads)
  if (have_initial) CK(cudaMemcpyAsync(c->w.p, w, (size_t)N * 8, cudaMemcpyHostToDevice, c->stream));
  else CK(cudaMemsetAsync(c->w.p, 0, (size_t)N * 8, c->stream));  // :125-128
  double nu_red[4], x_dot_nu = 0, gnorm = 0, mmin = 0, fx = 0;
  CKR(reduce4(c, c->nu_s.as<double>(), nullptr, N, c->red_out.as<double>() + 8));
  CK(cudaMemcpyAsync(nu_red, c->red_out.as<double>() + 8, sizeof nu_red, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const double nu_min = nu_red[2];

  // f(x): evaluation + (m, g = m - nu, f - nu.x)   optimal_transport.hpp:110-120
  double t_eval = 0, t_pcg = 0;  // wall seconds spent in evaluations / linear solves (trace only)
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  c->prev_stride = 0;  // no adjacency of an accepted point yet
  c->seeds_global = false;
  auto feval_inner = [&]() -> int {
    ++neval;
    if (neval > 1) {  // a line-search trial: first the cheap "is some cell already empty?" test
      bool empty = false;
      CKR(quick_empty(c, &empty));
      if (empty) {
        if (c->trace) fprintf(stderr, "[ma] trial rejected by the quick empty-cell test\n");
        mmin = 0.0;
        gnorm = 1e300;
        return MA_OK;
      }
    }
    c->abort_on_empty = neval > 1;  // every evaluation after the first is a line-search trial
    int rc_e = evaluate_mode<MODE_KANTOROVICH>(c, true);
    c->abort_on_empty = false;
    if (rc_e == MA_OK && c->aborted) {
      mmin = 0.0;
      gnorm = 1e300;
      return MA_OK;
    }
    if (rc_e != MA_OK && c->capacity_hit && neval > 1) {
      // a trial point so wild that one cell has > 64 Laguerre neighbours: the reference would evaluate it and
      // reject it (hidden neighbours => min m = 0 < eps0, optimal_transport.hpp:167); reject it here too
      mmin = -1.0;
      gnorm = 1e300;
      return MA_OK;
    }
    CKR(rc_e);
    // ws is the sorted copy of the weights used by this evaluation; g = m - nu on this context's tile (multi-GPU: the
    // tiles' |g|^2 are summed over NVLink, the weights are replicated so nu.x needs nothing)
    const int lo = tile_lo(c, c->part_rank), hi = tile_lo(c, c->part_rank + 1);
    k_sub<<<cdiv(hi - lo, 256), 256, 0, c->stream>>>(hi - lo, c->mass.as<double>() + lo, c->nu_s.as<double>() + lo, c->g_s.as<double>() + lo);
    c->launches++;
    CKR(reduce4(c, c->g_s.as<double>() + lo, nullptr, hi - lo, c->red_out.as<double>()));
    if (is_dist(c)) NCK(nccl_api().AllReduce(c->red_out.as<double>() + 1, c->red_out.as<double>() + 1, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
    CKR(reduce4(c, c->ws.as<double>(), c->nu_s.as<double>(), N, c->red_out.as<double>() + 4));
    double red[8];
    CK(cudaMemcpyAsync(red, c->red_out.p, sizeof red, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    gnorm = std::sqrt(red[1]);
    x_dot_nu = red[5];
    mmin = c->mass_min;
    fx = c->fval - x_dot_nu;
    return MA_OK;
  };
  auto feval = [&]() -> int {
    auto t0e = now();
    int rc_f = feval_inner();
    t_eval += secs(t0e, now());
    return rc_f;
  };
  auto finish = [&](int rc) {
    if (c->trace)
      fprintf(stderr, "[ma] ot_solve: niter=%zu neval=%zu cg=%zu  evaluations %.3f s, linear solves %.3f s\n", niter, neval,
              cg_total, t_eval, t_pcg);
    if (stats) {
      stats->niter = niter; stats->neval = neval; stats->cg_iters = cg_total;
      stats->final_norm = gnorm; stats->fval = fx;
      stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return rc;
  };

  CKR(feval());  // :131
  CKR(save_adjacency(c));
  const double eps0 = std::min(mmin, nu_min) / 2;  // :137-138
  if (!(eps0 > 0)) {                              // :139-148
    if (verbose) fprintf(stderr, "Error: computed minimum mass is non-positive\n");
    fail(c, MA_EMPTY_CELL, "computed minimum mass is non-positive: a Laguerre cell is empty at the initial point");
    return finish(MA_EMPTY_CELL);  // w is left untouched (= initial guess)
  }
  const int ground = [&]() {
    int g = 0;
    cudaMemcpy(&g, c->pos.as<int>() + (N - 1), 4, cudaMemcpyDeviceToHost);  // last index of the caller's ordering (T7)
    return g;
  }();
  int rc_final = MA_OK;
  bool have_dir = false;
  double last_alpha = 0.0;
  while (gnorm >= eps_g && niter++ <= maxiter) {  // :150-151 (including the maxiter+1 quirk, T6)
    int it = 0;
    double relres = 0;
    const double n_prev_g = gnorm;
    // d = -solve_laplacian_matrix(h, g)   :153   (internal order)
    auto t0p = now();
    // warm start: after a damped step tau the gradient is ~(1 - tau) g and the Hessian has hardly moved, so the new
    // direction is close to (1 - tau) times the last one (still in d_s)
    const bool warm = c->cg_warm && have_dir;
    if (warm) CK(cudaMemcpyAsync(c->scratch_d.p, c->d_s.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream));
    const bool dist = is_dist(c);
    if (dist) {  // the accepted point's gradient and Hessian of every tile, then the same solve on every rank
      CKR(dist_gather_slices(c, c->g_s.p, 8));
      CKR(dist_gather_hessian(c));
    }
    int rc = pcg_solve(c, N, dist ? c->rowptr_g.as<int>() : c->rowptr.as<int>(), dist ? c->col_g.as<int>() : c->col.as<int>(),
                       dist ? c->val_g.as<double>() : c->val.as<double>(), c->g_s.as<double>(), -1.0,
                       ground, c->d_s.as<double>(), &it, &relres, warm ? c->scratch_d.as<double>() : nullptr,
                       1.0 - last_alpha, /*use_amg=*/true, dist ? c->nnz_g : c->nnz);
    have_dir = true;
    t_pcg += secs(t0p, now());
    cg_total += it;
    if (rc != MA_OK) {
      // the caller keeps the progress made so far: c->w holds the last accepted point (the reference's x is
      // updated in place, optimal_transport.hpp:165, so a failing solve there also leaves the last iterate)
      CK(cudaMemcpyAsync(w, c->w.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      return finish(rc);
    }
    // optimal_transport.hpp:68-71 prints this whatever `verbose` says (absolute residual there)
    if (relres * n_prev_g > 1e-7) fprintf(stderr, "WARNING: in solve_laplacian_matrix: err=%g after %d CG iterations\n", relres * n_prev_g, it);
    double alpha = 1;
    const double n0 = gnorm;
    // x0 (sorted) = ws of the last evaluation
    CK(cudaMemcpyAsync(c->x0_s.p, c->ws.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream));
    size_t nls = 0;
    while (true) {  // :163-176
      // x = x0 + alpha d, written in caller order into c->w
      k_axpy_to<<<cdiv(N, 256), 256, 0, c->stream>>>(N, c->x0_s.as<double>(), alpha, c->d_s.as<double>(),
                                                     c->scratch_d.as<double>());
      k_scatter_to_caller<<<cdiv(N, 256), 256, 0, c->stream>>>(c->scratch_d.as<double>(), c->perm.as<int>(), N,
                                                               c->w.as<double>());
      c->launches += 2;
      CKR(feval());
      if (mmin >= eps0 && gnorm <= (1 - alpha / 2) * n0) { CKR(save_adjacency(c)); break; }
      alpha *= .5;
      if (verbose) fprintf(stderr, "subit %zu.%zu: min(masses)=%g\n", niter, nls++, mmin);
      if (alpha < 1e-30) {
        rc_final = fail(c, MA_NOT_CONVERGED, "line search failed (alpha underflow)");
        break;
      }
    }
    last_alpha = alpha;
    if (verbose)
      fprintf(stderr, "it %zu: f=%.15g |df|=%g min(m)=%g tau = %g eval = %zu cg = %d\n", niter, fx, gnorm, nu_min,
              alpha, neval, it);
    if (rc_final != MA_OK) break;
  }
  CK(cudaMemcpyAsync(w, c->w.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (rc_final == MA_OK && gnorm >= eps_g) rc_final = fail(c, MA_NOT_CONVERGED, "maxiter reached with |g|=%g", gnorm);
  return finish(rc_final);
}

// =============================================================================================
// instrumentation
// =============================================================================================
extern "C" int ma_set_profiling(ma_ctx *c, int on) { if (!c) return MA_INVALID; c->profiling = on; return MA_OK; }
extern "C" int ma_set_stats(ma_ctx *c, int on) { if (!c) return MA_INVALID; c->stats = on; return MA_OK; }
extern "C" int ma_get_timings(ma_ctx *c, float *ms) {
  if (!c || !ms) return MA_INVALID;
  for (int k = 0; k < MA_T_COUNT; ++k) ms[k] = c->t_ms[k];
  return MA_OK;
}
extern "C" int ma_get_counters(ma_ctx *c, int64_t *out) {
  if (!c || !out) return MA_INVALID;
  for (int k = 0; k < CNT_N; ++k) out[k] = c->host_counters[k];
  return MA_OK;
}
extern "C" int ma_flush_l2(ma_ctx *c, size_t bytes) {
  NEED_CTX();
  CKR(ensure(c, c->flush, bytes));
  k_fill_bytes<<<c->sm_count * 4, 256, 0, c->stream>>>(c->flush.as<unsigned long long>(), bytes / 8, 0x0123456789abcdefull);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return MA_OK;
}
extern "C" int ma_measure_fp64_peak(ma_ctx *c, double *flops_per_s) {
  NEED_CTX();
  if (!flops_per_s) return MA_INVALID;
  CKR(ensure(c, c->red_out, 16 * sizeof(double)));
  const int iters = 1 << 15, blocks = c->sm_count * 8;
  cudaEvent_t a = c->ev[0], b = c->ev[1];
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a, c->stream));
    k_dfma_probe<<<blocks, 256, 0, c->stream>>>(c->red_out.as<double>(), iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(b, c->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    double fl = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3);
    if (rep > 0) best = std::max(best, fl);
  }
  *flops_per_s = best;
  return MA_OK;
}

// ma_kernels.cuh — __global__ kernels of the evaluation path (K1 binning, K2 cells, K3 pieces,
// K4 CSR) and small utility kernels (scan, reductions, gathers).  sm_100a, fp64 CUDA-core work.
#pragma once
#include "ma_seg.cuh"

namespace ma {

// ================================================================================================
// utility: exclusive scan of int32 (n inputs -> n+1 outputs, out[n] = total)
// ================================================================================================
constexpr int SCAN_NT = 1024, SCAN_ITEMS = 4, SCAN_TILE = SCAN_NT * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int s = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  *total = warp_sums[nw - 1];
  int res = x - v + (warp > 0 ? warp_sums[warp - 1] : 0);
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_NT) k_scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                         int *__restrict__ tile_sums, int n) {
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int ex = block_exclusive_scan(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place, total to *grand
__global__ void __launch_bounds__(SCAN_NT) k_scan_sums(int *__restrict__ tile_sums, int nt, int *__restrict__ grand) {
  int carry = 0;
  for (int base = 0; base < nt; base += SCAN_NT) {
    int idx = base + threadIdx.x;
    int v = idx < nt ? tile_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    if (idx < nt) tile_sums[idx] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_NT) k_scan_add(int *__restrict__ out, const int *__restrict__ tile_sums, int n,
                                                       const int *__restrict__ grand) {
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int off = tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) out[base + k] += off;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand;
}

// ================================================================================================
// utility: deterministic reductions (fixed tree, no atomics)
// ================================================================================================
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out[0] = sum a, out[1] = sum a*a  (or a*b when b != null), out[2] = min a, out[3] = max a.
// Two stages: RED_BLOCKS partials, then one block.
constexpr int RED_NT = 256, RED_BLOCKS = 296;

__device__ __forceinline__ void block_reduce4(double s, double q, double mn, double mx, double *dst) {
  __shared__ double sh[4][RED_NT / 32];
  s = warp_sum(s); q = warp_sum(q); mn = warp_min(mn); mx = warp_max(mx);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = q; sh[2][warp] = mn; sh[3][warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    constexpr int NW = RED_NT / 32;
    s = lane < NW ? sh[0][lane] : 0.0;
    q = lane < NW ? sh[1][lane] : 0.0;
    mn = lane < NW ? sh[2][lane] : 1e300;
    mx = lane < NW ? sh[3][lane] : -1e300;
    s = warp_sum(s); q = warp_sum(q); mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { dst[0] = s; dst[1] = q; dst[2] = mn; dst[3] = mx; }
  }
}

__global__ void __launch_bounds__(RED_NT) k_reduce_stage1(const double *__restrict__ a, const double *__restrict__ b,
                                                           int n, double *__restrict__ partial) {
  double s = 0, q = 0, mn = 1e300, mx = -1e300;
  for (int i = blockIdx.x * RED_NT + threadIdx.x; i < n; i += gridDim.x * RED_NT) {
    double v = a[i];
    s += v;
    q += b ? v * b[i] : v * v;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  block_reduce4(s, q, mn, mx, partial + 4 * blockIdx.x);
}
__global__ void __launch_bounds__(RED_NT) k_reduce_stage2(const double *__restrict__ partial, int nb,
                                                           double *__restrict__ out) {
  double s = 0, q = 0, mn = 1e300, mx = -1e300;
  for (int i = threadIdx.x; i < nb; i += RED_NT) {
    s += partial[4 * i]; q += partial[4 * i + 1];
    mn = fmin(mn, partial[4 * i + 2]); mx = fmax(mx, partial[4 * i + 3]);
  }
  block_reduce4(s, q, mn, mx, out);
}

// two arrays at once (blockIdx.y / blockIdx.x of stage 2 picks the array); same trees as the kernels above
__global__ void __launch_bounds__(RED_NT) k_reduce2_stage1(const double *__restrict__ a0, const double *__restrict__ a1, int n,
                                                            double *__restrict__ partial) {
  const double *__restrict__ a = blockIdx.y ? a1 : a0;
  double s = 0, q = 0, mn = 1e300, mx = -1e300;
  for (int i = blockIdx.x * RED_NT + threadIdx.x; i < n; i += gridDim.x * RED_NT) {
    double v = a[i];
    s += v;
    q += v * v;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  block_reduce4(s, q, mn, mx, partial + 4 * (blockIdx.y * RED_BLOCKS + blockIdx.x));
}
__global__ void __launch_bounds__(RED_NT) k_reduce2_stage2(const double *__restrict__ partial, int nb, double *__restrict__ out) {
  partial += 4 * RED_BLOCKS * blockIdx.x;
  double s = 0, q = 0, mn = 1e300, mx = -1e300;
  for (int i = threadIdx.x; i < nb; i += RED_NT) {
    s += partial[4 * i]; q += partial[4 * i + 1];
    mn = fmin(mn, partial[4 * i + 2]); mx = fmax(mx, partial[4 * i + 3]);
  }
  block_reduce4(s, q, mn, mx, out + 4 * blockIdx.x);
}

// ================================================================================================
// K1: Morton-ordered uniform bins of the Diracs (once per point set) + per-eval max-weight pyramid
// ================================================================================================
__global__ void k_bin_count(const double *__restrict__ x, const double *__restrict__ y, int n, double px0, double py0,
                            double pinv, int G, unsigned *__restrict__ code, int *__restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int bx = min(max((int)((x[i] - px0) * pinv), 0), G - 1);
  int by = min(max((int)((y[i] - py0) * pinv), 0), G - 1);
  unsigned c = morton2((unsigned)bx, (unsigned)by);
  code[i] = c;
  atomicAdd(&count[c], 1);
}
__global__ void k_bin_scatter(const unsigned *__restrict__ code, int n, const int *__restrict__ bin_start,
                              int *__restrict__ fill, int *__restrict__ perm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned c = code[i];
  int slot = atomicAdd(&fill[c], 1);
  perm[bin_start[c] + slot] = i;
}
// order inside a bin = caller index order, so that the internal ordering (and every sum order that
// follows from it) is reproducible from run to run
__global__ void k_bin_sort(const int *__restrict__ bin_start, int nbins, int *__restrict__ perm) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  int s = bin_start[b], e = bin_start[b + 1];
  for (int i = s + 1; i < e; ++i) {
    int v = perm[i], j = i - 1;
    while (j >= s && perm[j] > v) { perm[j + 1] = perm[j]; --j; }
    perm[j + 1] = v;
  }
}
// row-major (start, end) table of the leaf bins for K2's ring walk
__global__ void k_bin_rowmajor(int G, const int *__restrict__ bin_start, int *__restrict__ bin_rm) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)G * G) return;
  const unsigned cx = (unsigned)(q % G), cy = (unsigned)(q / G);
  const unsigned code = morton2(cx, cy);
  reinterpret_cast<int2 *>(bin_rm)[q] = make_int2(bin_start[code], bin_start[code + 1]);
}
// The block grid of ma_block.cuh: the (Morton-sorted) sites filed a second time under row-major bins of a bG x bG
// grid (count, scan, scatter, sort inside the bin by site index so that the order is reproducible).
__global__ void k_blk_count(const double *__restrict__ xs, const double *__restrict__ ys, int n, double px0, double py0,
                            double binv, int G, int *__restrict__ bin, int *__restrict__ count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int bx = min(max((int)((xs[k] - px0) * binv), 0), G - 1);
  const int by = min(max((int)((ys[k] - py0) * binv), 0), G - 1);
  const int q = by * G + bx;
  bin[k] = q;
  atomicAdd(&count[q], 1);
}
__global__ void k_blk_scatter(const int *__restrict__ bin, int n, const int *__restrict__ start, int *__restrict__ fill,
                              int *__restrict__ rm2s) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int q = bin[k];
  rm2s[start[q] + atomicAdd(&fill[q], 1)] = k;
}
__global__ void k_blk_fill(int nb, const int *__restrict__ start, int *__restrict__ rm2s, const double *__restrict__ xs,
                           const double *__restrict__ ys, double *__restrict__ xr, double *__restrict__ yr,
                           int *__restrict__ s2rm) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb) return;
  const int s = start[q], e = start[q + 1];
  for (int a = s + 1; a < e; ++a) {
    const int v = rm2s[a];
    int b = a - 1;
    while (b >= s && rm2s[b] > v) { rm2s[b + 1] = rm2s[b]; --b; }
    rm2s[b + 1] = v;
  }
  for (int a = s; a < e; ++a) {
    const int k = rm2s[a];
    xr[a] = xs[k]; yr[a] = ys[k]; s2rm[k] = a;
  }
}
// per evaluation: ws[k] = w[perm[k]] and its row-major twin
__global__ void k_gather_w(const double *__restrict__ w, const int *__restrict__ perm, const int *__restrict__ s2rm, int n,
                           double *__restrict__ ws, double *__restrict__ wr) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) { const double v = w[perm[k]]; ws[k] = v; wr[s2rm[k]] = v; }
}
__global__ void k_gather_points(const double *__restrict__ x, const double *__restrict__ y,
                                const int *__restrict__ perm, int n, double *__restrict__ xs, double *__restrict__ ys,
                                int *__restrict__ pos) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int i = perm[k];
  xs[k] = x[i];
  ys[k] = y[i];
  pos[i] = k;
}
__global__ void k_gather_u32(const unsigned *__restrict__ src, const int *__restrict__ perm, int n,
                             unsigned *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[perm[k]];
}
// dst[k] = src[perm[k]]
__global__ void k_gather(const double *__restrict__ src, const int *__restrict__ perm, int n, double *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[perm[k]];
}

// per-eval: leaf maxima of the bins + 4 levels above, one block per 256 consecutive Morton bins
__global__ void __launch_bounds__(256) k_wmax_leaf(const double *__restrict__ ws, const int *__restrict__ bin_start,
                                                    int L, double *__restrict__ wmax) {
  __shared__ double sh[256];
  const size_t nleaf = (size_t)1 << (2 * L);
  size_t b = (size_t)blockIdx.x * 256 + threadIdx.x;
  double m = -1.0 / 0.0;
  if (b < nleaf) {
    int s = bin_start[b], e = bin_start[b + 1];
    for (int k = s; k < e; ++k) m = fmax(m, ws[k]);
    wmax[(nleaf - 1) / 3 + b] = m;
  }
  sh[threadIdx.x] = m;
  __syncthreads();
  int width = 256;
  for (int up = 1; up <= 4 && up <= L; ++up) {
    width >>= 2;
    double v = 0;
    if ((int)threadIdx.x < width) {
      int c = 4 * threadIdx.x;
      v = fmax(fmax(sh[c], sh[c + 1]), fmax(sh[c + 2], sh[c + 3]));
    }
    __syncthreads();
    if ((int)threadIdx.x < width) {
      sh[threadIdx.x] = v;
      size_t nl = (size_t)1 << (2 * (L - up));
      size_t node = (size_t)blockIdx.x * width + threadIdx.x;
      if (node < nl) wmax[(nl - 1) / 3 + node] = v;
    }
    __syncthreads();
  }
}
// remaining top levels (L-5 .. 0), single block
__global__ void __launch_bounds__(1024) k_wmax_top(int L, double *__restrict__ wmax) {
  for (int l = L - 5; l >= 0; --l) {
    size_t nl = (size_t)1 << (2 * l);
    const double *child = wmax + (4 * nl - 1) / 3;
    double *mine = wmax + (nl - 1) / 3;
    for (size_t c = threadIdx.x; c < nl; c += 1024)
      mine[c] = fmax(fmax(child[4 * c], child[4 * c + 1]), fmax(child[4 * c + 2], child[4 * c + 3]));
    __syncthreads();
  }
}

// ================================================================================================
// K1 (continued): per-node supporting planes of the lifted sites (ma_geom.cuh).
//   * inclusive-exclusive prefix sums over the Morton-sorted sites of the moment terms
//       SET 0 (once per point set):  x, y, x^2, y^2, x y        (coordinates relative to (cx, cy))
//       SET 1 (every evaluation):    w, x w, y w
//     out[c * (n + 1) + k] = sum_{j < k} term_c(j); a node's sums are differences of two entries
//     because the sites of a quadtree node are contiguous in Morton order;
//   * k_node_fit: least-squares weight gradient G per node (inherits the nearest valid ancestor's);
//   * k_node_alpha: alpha_B = min_j (|t_j|^2 - w_j + G.t_j) by atomicMin on order-preserving keys.
// ================================================================================================
constexpr int FS_NT = 256, FS_ITEMS = 4, FS_TILE = FS_NT * FS_ITEMS;

// The per-node planes are only consulted when the weights vary enough to displace cells from their
// Diracs (CellSearch::use_m, same threshold): the kernels that build them return at once otherwise.
// gate = {wstat, 0.25 * ph^2}; wstat = {sum, sum of squares, min, max} of the weights (device).
struct PlaneGate {
  const double *wstat;
  double thresh;
  __device__ __forceinline__ bool off() const { return wstat && !((wstat[3] - wstat[2]) > thresh); }
};

template <int SET> struct MomentTerms;
template <> struct MomentTerms<0> { static constexpr int K = 5; };
template <> struct MomentTerms<1> { static constexpr int K = 3; };

template <int SET>
__device__ __forceinline__ void moment_terms(const double *__restrict__ xs, const double *__restrict__ ys,
                                             const double *__restrict__ ws, double cx, double cy, int j, double *v) {
  double x = xs[j] - cx, y = ys[j] - cy;
  if (SET == 0) { v[0] = x; v[1] = y; v[2] = x * x; v[3] = y * y; v[4] = x * y; }
  else { double w = ws[j]; v[0] = w; v[1] = x * w; v[2] = y * w; }
}

// block-wide exclusive scan of one double per thread (FS_NT threads); total to *total
__device__ __forceinline__ double block_exclusive_scan_f64(double v, double *total, double *sh /* FS_NT/32 + 1 */) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    constexpr int NW = FS_NT / 32;
    double s = lane < NW ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < NW) sh[lane] = s;
  }
  __syncthreads();
  *total = sh[FS_NT / 32 - 1];
  double res = x - v + (warp > 0 ? sh[warp - 1] : 0.0);
  __syncthreads();
  return res;
}

template <int SET>
__global__ void __launch_bounds__(FS_NT) k_moment_scan_tiles(const double *__restrict__ xs, const double *__restrict__ ys,
                                                              const double *__restrict__ ws, double cx, double cy, int n,
                                                              double *__restrict__ out, double *__restrict__ tile_sums,
                                                              PlaneGate gate) {
  if (SET == 1 && gate.off()) return;
  constexpr int K = MomentTerms<SET>::K;
  __shared__ double sh[FS_NT / 32 + 1];
  const int base = blockIdx.x * FS_TILE + threadIdx.x * FS_ITEMS;
  double v[FS_ITEMS][K], s[K];
#pragma unroll
  for (int c = 0; c < K; ++c) s[c] = 0.0;
#pragma unroll
  for (int q = 0; q < FS_ITEMS; ++q) {
    if (base + q < n) moment_terms<SET>(xs, ys, ws, cx, cy, base + q, v[q]);
    else
#pragma unroll
      for (int c = 0; c < K; ++c) v[q][c] = 0.0;
#pragma unroll
    for (int c = 0; c < K; ++c) s[c] += v[q][c];
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {
    double total;
    double ex = block_exclusive_scan_f64(s[c], &total, sh);
    double *o = out + (size_t)c * (n + 1);
#pragma unroll
    for (int q = 0; q < FS_ITEMS; ++q) {
      if (base + q < n) o[base + q] = ex;
      ex += v[q][c];
    }
    if (threadIdx.x == 0) tile_sums[(size_t)c * gridDim.x + blockIdx.x] = total;
  }
}
// one block per component: exclusive scan of that component's tile sums in place; grand total to out[c*(n+1)+n]
__global__ void __launch_bounds__(FS_NT) k_moment_scan_sums(double *__restrict__ tile_sums, int nt, int n,
                                                             double *__restrict__ out, PlaneGate gate) {
  if (gate.off()) return;
  __shared__ double sh[FS_NT / 32 + 1];
  double *ts = tile_sums + (size_t)blockIdx.x * nt;
  double carry = 0.0;
  for (int base = 0; base < nt; base += FS_NT) {
    int idx = base + threadIdx.x;
    double v = idx < nt ? ts[idx] : 0.0, total;
    double ex = block_exclusive_scan_f64(v, &total, sh);
    if (idx < nt) ts[idx] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) out[(size_t)blockIdx.x * (n + 1) + n] = carry;
}
__global__ void __launch_bounds__(FS_NT) k_moment_scan_add(double *__restrict__ out, const double *__restrict__ tile_sums,
                                                            int n, int K, PlaneGate gate) {
  if (gate.off()) return;
  const int base = blockIdx.x * FS_TILE + threadIdx.x * FS_ITEMS;
  for (int c = 0; c < K; ++c) {
    double off = tile_sums[(size_t)c * gridDim.x + blockIdx.x];
    double *o = out + (size_t)c * (n + 1);
#pragma unroll
    for (int q = 0; q < FS_ITEMS; ++q)
      if (base + q < n) o[base + q] += off;
  }
}

// gradient of node (l, code) from the prefix sums; false if the node cannot support a fit
__device__ __forceinline__ bool node_fit_one(int L, int l, unsigned code, const int *__restrict__ bin_start, int n,
                                             const double *__restrict__ pre0, const double *__restrict__ pre1,
                                             double &Gx, double &Gy) {
  const int sh = 2 * (L - l);
  const int s = bin_start[(size_t)code << sh], e = bin_start[((size_t)code + 1) << sh];
  if (e - s < 6) return false;
  const size_t st = (size_t)n + 1;
  double m[8];
#pragma unroll
  for (int c = 0; c < 5; ++c) m[c] = pre0[c * st + e] - pre0[c * st + s];
#pragma unroll
  for (int c = 0; c < 3; ++c) m[5 + c] = pre1[c * st + e] - pre1[c * st + s];
  return node_gradient((double)(e - s), m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], Gx, Gy);
}
__global__ void k_node_fit(int L, const int *__restrict__ bin_start, int n, const double *__restrict__ pre0,
                           const double *__restrict__ pre1, double *__restrict__ nodeG,
                           unsigned long long *__restrict__ nodeA, PlaneGate gate) {
  if (gate.off()) return;
  const size_t nnodes = level_offset(L + 1);
  size_t node = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= nnodes) return;
  int l = 0;
  while (level_offset(l + 1) <= node) ++l;
  unsigned code = (unsigned)(node - level_offset(l));
  double Gx = 0.0, Gy = 0.0;
  for (int a = l; a >= 0; --a) {  // own fit, else the nearest ancestor's
    if (node_fit_one(L, a, code >> (2 * (l - a)), bin_start, n, pre0, pre1, Gx, Gy)) break;
    Gx = Gy = 0.0;
  }
  nodeG[2 * node] = Gx;
  nodeG[2 * node + 1] = Gy;
  nodeA[node] = dkey(1.0 / 0.0);
}
// one thread per site: fold the site into alpha of its node at every level
__global__ void k_node_alpha(int n, int L, const double *__restrict__ xs, const double *__restrict__ ys,
                             const double *__restrict__ ws, const unsigned *__restrict__ code_s, double px0, double py0,
                             double ph, const double *__restrict__ nodeG, unsigned long long *__restrict__ nodeA,
                             PlaneGate gate) {
  if (gate.off()) return;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < n;
  const unsigned leaf = live ? code_s[j] : 0u;
  const double x = live ? xs[j] : 0.0, y = live ? ys[j] : 0.0, w = live ? ws[j] : 0.0;
  for (int l = L; l >= 0; --l) {
    const unsigned code = leaf >> (2 * (L - l));
    const size_t node = level_offset(l) + code;
    const double S = ph * (double)(1u << (L - l));
    const double tx = x - (px0 + ((double)morton_compact1(code) + 0.5) * S);
    const double ty = y - (py0 + ((double)morton_compact1(code >> 1) + 0.5) * S);
    double v = 1.0 / 0.0;
    if (live) v = tx * tx + ty * ty - w + nodeG[2 * node] * tx + nodeG[2 * node + 1] * ty;
    // warp-aggregate when the whole warp sits in one node (sites are Morton-contiguous)
    const unsigned c0 = __shfl_sync(0xffffffffu, code, 0);
    const bool uniform = __all_sync(0xffffffffu, !live || code == c0);
    if (uniform) {
      v = warp_min(v);
      if ((threadIdx.x & 31) == 0 && v < 1.0 / 0.0) atomicMin(&nodeA[level_offset(l) + c0], dkey(v));
    } else if (live) {
      atomicMin(&nodeA[node], dkey(v));
    }
  }
}

// The weights vary so much that the block certificate (which only knows the global maximum weight) would certify
// few cells: the block kernels then leave everything to CellSearch (quadtree walk with supporting planes).
__device__ __forceinline__ bool weights_graded(const Params &p) { return (p.wstat[3] - p.wstat[2]) > MA_LEAN_RANGE * p.bph * p.bph; }

// ================================================================================================
// K2: one thread per cell.  POLY: also store the cell polygon for k_seg (grid meshes).
// ================================================================================================
template <int MAXV, int NT, bool POLY> __global__ void __launch_bounds__(NT) k_cells(Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + MAXV * NT;
  int *st = reinterpret_cast<int *>(sy + MAXV * NT);
  int i = p.cell_lo + blockIdx.x * NT + threadIdx.x;
  // cell_build votes across the warp: lanes past the end redo the last cell and write nothing
  const bool valid = i < p.cell_hi;
  if (!valid) i = p.cell_hi - 1;
  PolyRef<NT, (MAXV <= 16)> P{sx + threadIdx.x, sy + threadIdx.x, st + threadIdx.x};
  int fl = 0;
  int n = cell_build(p, i, P, MAXV, &fl);
  MA_WARP_SYNC();
  if (!valid) return;
  if (fl) atomicOr(p.flags, fl);
  // a genuinely empty cell (not a capacity / stack overflow, which the host answers by escalating):
  // the line search rejects this trial point
  if (n == 0 && p.abort_on_empty) p.flags[1] = 1;
  if (n < 0) n = 0;
  cell_emit(p, i, P, n);
  if (POLY) {
    p.poly_n[i] = n;
    for (int k = 0; k < n; ++k) {
      const size_t o = (size_t)k * p.N + i;
      p.poly_x[o] = P.X(k); p.poly_y[o] = P.Y(k); p.poly_t[o] = P.T(k);
    }
  }
}
template <int MAXV, int NT> constexpr size_t cells_smem_bytes() { return (size_t)MAXV * NT * (8 + 8 + 4); }

// K2 with persistent lanes: each warp owns a contiguous chunk of cells and its lanes fetch the next
// cell of the chunk as soon as they finish one (ranks from a ballot, no atomics, so the assignment is
// deterministic), instead of idling until the slowest cell of a fixed group of 32 is done.  Search
// steps, clips and refills are each executed when enough lanes ask for them (warp votes).
#ifndef MA_K2_MINBLOCKS
#define MA_K2_MINBLOCKS 5  // 5 blocks of 128 threads per SM: <= 102 registers, 41 KB of polygons each
#endif
// `list` != null: the kernel handles the cells list[0 .. *list_n) left over by the block kernels (k_cells_block) and
// sizes its chunks itself from the device-side count (the host never reads it); it returns at once when the weights are
// graded.  list == null with list_n != null is the complementary launch: the whole tile, but only when the weights are
// graded (the block kernels then did nothing).
template <int MAXV, int NT, bool POLY> __global__ void __launch_bounds__(NT, (MAXV <= 16 ? MA_K2_MINBLOCKS : 1)) k_cells_persist(Params p, int chunk, const int *__restrict__ list, const int *__restrict__ list_n, bool warm = false) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + MAXV * NT;
  int *st = reinterpret_cast<int *>(sy + MAXV * NT);
  typedef PolyRef<NT, (MAXV <= 16)> Poly;
  Poly P{sx + threadIdx.x, sy + threadIdx.x, st + threadIdx.x};
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  const long long warp_global = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
  int lo = p.cell_lo, hi = p.cell_hi;
  if (list) {
    // the cells the block kernels left over; with graded weights those kernels did nothing and the launch with
    // list == null && list_n != null (a full grid, sized for the whole tile) does the work instead
    // (warm: the cells the ring match sent back, whatever the weights look like)
    if (!warm && weights_graded(p)) return;
    lo = 0; hi = *list_n;
    const long long nwarps = (long long)gridDim.x * (NT / 32);
    chunk = (int)max(1ll, ((long long)(hi - lo) + nwarps - 1) / nwarps);
  } else if (list_n) {
    if (!weights_graded(p)) return;
  }
  long long first = (long long)lo + warp_global * chunk;
  int next = (int)(first < hi ? first : hi);                       // warp-uniform
  const int end = (int)(first + chunk < hi ? first + chunk : hi);  // warp-uniform
  CellSearch<Poly> S;
  unsigned stk[CELL_STACK];
  S.i = -1; S.jc = -1; S.phase = 2; S.n = 0; S.status = 0;  // "done, nothing to emit"
  bool parked = false;  // no more cells for this lane
  for (;;) {
    // line-search trial: as soon as any cell anywhere is known to be empty the trial point is rejected
    // (optimal_transport.hpp:167), so the whole warp drops what it is doing (one lane polls the flag)
    if (p.abort_on_empty) {
      int f = (lane == 0) ? *(volatile const int *)p.abort_flag : 0;
      f = __shfl_sync(0xffffffffu, f, 0);
      if (f) break;
    }
    // ---- refill: lanes whose cell is finished write it out and take the next cell of the chunk ----
    const bool fin = !parked && S.done();
    const unsigned finmask = __ballot_sync(0xffffffffu, fin);
    const int n_fin = __popc(finmask);
    const int n_busy = __popc(__ballot_sync(0xffffffffu, !parked && !S.done()));
    if (n_fin > 0 && (n_fin >= p.refill_at || n_busy == 0)) {
      if (fin) {
        if (S.i >= 0) {
          const int i = S.i;
          int n = S.n;
          // a genuinely empty cell rejects the line-search trial; an overflowed one must NOT (the host
          // escalates the capacity class and evaluates the trial point again)
          if (S.status) { atomicOr(p.flags, S.status); n = 0; }
          else if (n == 0 && p.abort_on_empty) p.flags[1] = 1;
          cell_emit(p, i, P, n);
          if (POLY) {
            p.poly_n[i] = n;
            for (int k = 0; k < n; ++k) {
              const size_t o = (size_t)k * p.N + i;
              p.poly_x[o] = P.X(k); p.poly_y[o] = P.Y(k); p.poly_t[o] = P.T(k);
            }
          }
          if (warm) {  // an exact cell from now on
            ring_store(p.ring, p.ring_n, i, P, n);
            p.cstate[i] = WARM_EXACT;
          }
        }
        const int idx = next + __popc(finmask & lt_mask);
        if (idx < end) S.init(p, list ? list[idx] : idx, P);
        else { parked = true; S.i = -1; }
      }
      next += n_fin;
      __syncwarp();
    }
    if (__all_sync(0xffffffffu, parked)) break;
    // ---- search: while more lanes search than hold a cutting site (and few wait for a refill) ----
    for (;;) {
      const bool se = !parked && S.searching();
      const int n_search = __popc(__ballot_sync(0xffffffffu, se));
      const int n_hold = __popc(__ballot_sync(0xffffffffu, !parked && S.holding()));
      const int n_wait = __popc(__ballot_sync(0xffffffffu, !parked && S.done()));
      if (n_search == 0 || n_hold * p.clip_a >= n_search * p.clip_b || n_wait >= p.refill_at) break;
      if (se) S.search_step(p, P, stk);
      __syncwarp();
    }
    // ---- clip ----
    const bool ho = !parked && S.holding();
    if (__any_sync(0xffffffffu, ho)) {
      if (ho) S.clip(p, P, MAXV);
      __syncwarp();
    }
  }
}

// ================================================================================================
// K2, fast path (ma_block.cuh): one thread per cell, all lanes of a warp walk their candidate lists in lock
// step.  Cells the block of radius R cannot certify (or whose polygon outgrows the 16-vertex class) are
// appended to out_list for the next stage; with graded weights the kernel does nothing at all (CellSearch's
// quadtree walk with supporting planes is the path for that regime).
// ================================================================================================
#ifndef MA_K2B_MINBLOCKS
#define MA_K2B_MINBLOCKS 4  // measured: 128 registers (no spills) at 16 warps per SM beat 96 registers at 20 (1.43 vs 1.54 ms of K2 at c3)
#endif
// R0 >= 0: the pass continues from the polygon the pass of radius R0 stored and looks only at the bins beyond that block.
template <int R0, int R, int MAXV, int NT, bool POLY>
__global__ void __launch_bounds__(NT, MA_K2B_MINBLOCKS) k_cells_block(Params p, const int *__restrict__ in_list, const int *__restrict__ in_n,
                                                                      int *__restrict__ out_list, int *__restrict__ out_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (weights_graded(p)) return;
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + MAXV * NT;
  int *st = reinterpret_cast<int *>(sy + MAXV * NT);
  typedef PolyRef<NT, true> Poly;
  const unsigned lane = threadIdx.x & 31u;
  const int n_in = in_list ? *in_n : p.cell_hi - p.cell_lo;
  for (int base = blockIdx.x * NT; base < n_in; base += gridDim.x * NT) {
    const int idx = base + threadIdx.x;
    if ((idx & ~31) >= n_in) continue;  // warp-uniform (NT is a multiple of 32)
    const bool valid = idx < n_in;
    const int i = valid ? (in_list ? in_list[idx] : p.cell_lo + idx) : (in_list ? in_list[n_in - 1] : p.cell_hi - 1);
    Poly P{sx + threadIdx.x, sy + threadIdx.x, st + threadIdx.x};
    CellSearch<Poly> S;
    S.init(p, i, P);
    bool cert = false, usable = true;
    if (R0 >= 0) usable = block_reload(p, S, P);
    block_search<R0, R>(p, S, P, MAXV, valid && usable, cert);
    cert = cert && usable;
    __syncwarp();
    if (valid && cert) {
      const int n = S.n;
      if (n == 0 && p.abort_on_empty) p.flags[1] = 1;  // a hidden Dirac: the line search rejects this trial point
      cell_emit(p, i, P, n);
    }
    if (valid && usable && (cert ? POLY : true)) {
      // certified: the polygon k_seg integrates; not certified: what the next pass continues from (-1: nothing usable)
      const int n = (S.phase == 0 || cert) ? S.n : -1;
      p.poly_n[i] = n;
      for (int k = 0; k < n; ++k) {
        const size_t o = (size_t)k * p.N + i;
        p.poly_x[o] = P.X(k); p.poly_y[o] = P.Y(k); p.poly_t[o] = P.T(k);
      }
    }
    const bool hard = valid && !cert;
    if (R0 < 0 && valid) p.cstate[i] = hard ? 1 : 0;  // k_seg's first pass integrates the certified cells while the later passes of K2 run
    const unsigned hm = __ballot_sync(0xffffffffu, hard);
    if (hm) {
      int b = 0;
      if (lane == (unsigned)(__ffs(hm) - 1)) b = atomicAdd(out_n, __popc(hm));
      b = __shfl_sync(0xffffffffu, b, __ffs(hm) - 1);
      if (hard) out_list[b + __popc(hm & ((1u << lane) - 1u))] = i;
    }
    __syncwarp();
  }
}

// K2, warm path (ma_warm.cuh), step 1: every cell from the neighbours it had in the seed evaluation.  Cells that cannot
// be seeded (no seed row, more than 16 vertices on the way) go straight to the list of cells CellSearch rebuilds.
template <int NT, bool POLY>
__global__ void __launch_bounds__(NT, 4) k_cells_seed(Params p, const int *__restrict__ seed_nbr, const int *__restrict__ seed_cnt,
                                                      int *__restrict__ out_list, int *__restrict__ out_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + 16 * NT;
  int *st = reinterpret_cast<int *>(sy + 16 * NT);
  typedef PolyRef<NT, true> Poly;
  const unsigned lane = threadIdx.x & 31u;
  const int idx = blockIdx.x * NT + threadIdx.x;
  const int ncell = p.cell_hi - p.cell_lo;
  if ((idx & ~31) >= ncell) return;  // warp-uniform
  if (p.abort_on_empty && *(volatile const int *)p.abort_flag) return;  // line-search trial: an empty cell has already been found
  const bool valid = idx < ncell;
  const int i = p.cell_lo + (valid ? idx : ncell - 1);
  Poly P{sx + threadIdx.x, sy + threadIdx.x, st + threadIdx.x};
  CellSearch<Poly> S;
  S.init(p, i, P);
  const int cnt = valid ? seed_cnt[i] : 0;
  seed_build(p, S, P, 16, valid && cnt > 0, seed_nbr + (size_t)i * RING_STRIDE, cnt);
  __syncwarp();
  // seeded: a polygon (or a proven-empty cell); hard: nothing usable
  const bool hard = valid && (cnt <= 0 || S.status != 0);
  if (valid && !hard) {
    const int n = S.n;
    if (n == 0 && p.abort_on_empty) p.flags[1] = 1;  // a hidden Dirac: the line search rejects this trial point
    cell_emit(p, i, P, n);
    ring_store(p.ring, p.ring_n, i, P, n);
    p.cstate[i] = n == 0 ? WARM_EXACT : WARM_SEEDED;
    if (POLY) {
      p.poly_n[i] = n;
      for (int k = 0; k < n; ++k) {
        const size_t o = (size_t)k * p.N + i;
        p.poly_x[o] = P.X(k); p.poly_y[o] = P.Y(k); p.poly_t[o] = P.T(k);
      }
    }
  }
  if (hard) { p.ring_n[i] = -1; p.cstate[i] = WARM_QUEUED; }
  const unsigned hm = __ballot_sync(0xffffffffu, hard);
  if (hm) {
    int b = 0;
    if (lane == (unsigned)(__ffs(hm) - 1)) b = atomicAdd(out_n, __popc(hm));
    b = __shfl_sync(0xffffffffu, b, __ffs(hm) - 1);
    if (hard) out_list[b + __popc(hm & ((1u << lane) - 1u))] = i;
  }
}

// Step 2: the combinatorial certificate.  The cells of every failing vertex that are not exact yet are appended to
// out_list (once: the state flips SEEDED -> QUEUED under an atomic); *n_fail counts the cells with a failing vertex.
__global__ void __launch_bounds__(256) k_cells_match(Params p, int *__restrict__ out_list, int *__restrict__ out_n, int *__restrict__ n_fail) {
  const int i = p.cell_lo + blockIdx.x * 256 + threadIdx.x;
  if (i >= p.cell_hi) return;
  if (p.abort_on_empty && *(volatile const int *)p.abort_flag) return;
  int *state = p.cstate;
  const bool ok = ring_match(p.ring, p.ring_n, i, [&](int c) {
    if (atomicCAS(state + c, (int)WARM_SEEDED, (int)WARM_QUEUED) == WARM_SEEDED) out_list[atomicAdd(out_n, 1)] = c;
  });
  if (!ok) atomicAdd(n_fail, 1);
}

// Line-search trials (optimal_transport.hpp:163-170): most of them are rejected because some cell has become EMPTY.
// Before a trial point is evaluated this kernel clips every cell against the neighbours it had at the last ACCEPTED
// point only — a superset of the true cell — and raises the abort flag if one of those supersets is already empty:
// then the true cell is empty too and the trial is rejected without the neighbour search (which, with the graded
// weights of the Newton iterates, is the expensive quadtree walk: 20 ms per evaluation at 1 M Diracs against 0.3 ms here).
template <int NT> __global__ void __launch_bounds__(NT, 4) k_cells_quick_empty(Params p, const int *__restrict__ prev_nbr,
                                                                               const int *__restrict__ prev_cnt, int prev_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + 16 * NT;
  int *st = reinterpret_cast<int *>(sy + 16 * NT);
  typedef PolyRef<NT, true> Poly;
  const int idx = blockIdx.x * NT + threadIdx.x;
  const int ncell = p.cell_hi - p.cell_lo;
  if ((idx & ~31) >= ncell) return;  // warp-uniform
  // one empty cell is all the caller wants to know: blocks that start after the flag went up leave at once (a rejected
  // trial usually has thousands of empty cells, so the first wave of blocks settles it)
  if (*(volatile const int *)(p.flags + 1)) return;
  const bool valid = idx < ncell;
  const int i = p.cell_lo + (valid ? idx : ncell - 1);
  Poly P{sx + threadIdx.x, sy + threadIdx.x, st + threadIdx.x};
  CellSearch<Poly> S;
  S.init(p, i, P);
  const int cnt = valid ? max(prev_cnt[i], 0) : 0;
  const int cmax = __reduce_max_sync(0xffffffffu, cnt);
  bool live = valid && cnt > 0;
  for (int s = 0; s < cmax; ++s) {
    const bool act = live && s < cnt;
    const int j = act ? prev_nbr[(size_t)i * prev_stride + s] : i;
    const double Dx = p.xs[j] - S.xi, Dy = p.ys[j] - S.yi, dw = S.wi - p.ws[j];
    const double dd2 = Dx * Dx + Dy * Dy, c = 0.5 * (dd2 + dw);
    bool cut = false;
    if (act && dd2 > 0.0 && !(c >= 0.0 && c * c >= S.R2 * dd2 * (1.0 + 1e-9))) {
      double r2;
      const unsigned long long in = S.sign_mask(p, P, j, Dx, Dy, c, dd2, dw, r2);
      S.R2 = r2;
      if (in == 0ull) { p.flags[1] = 1; live = false; }
      else if (in != lowmask64(S.n)) { S.jc = j; S.cDx = Dx; S.cDy = Dy; S.cc = c; S.cin = in; cut = true; }
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, cut)) {
      if (cut) {
        S.template clip<false>(p, P, 16);
        if (S.phase != 0) live = false;  // more than 16 vertices on the way: no statement about this cell
      }
      __syncwarp();
    }
  }
}

// K2, the tail: ONE WARP per cell for the few cells (0.4 % of a uniform point set) that even the 7 x 7 block cannot
// certify.  Thread-per-cell kernels run them at the latency of a single lane walking 121 candidates; here the 32
// lanes test 32 candidates of the block of radius R at once against the polygon (shared memory), the cutting ones are
// clipped one after the other by their own lane and the survivors re-tested — the mapping of north_star (2).
// R0 >= 0: continues from the polygon the pass of radius R0 stored (only the bins beyond that block are looked at: the polygon
// is tight already, so most candidates fail the disk test and few clips are left to do one after the other).
template <int R0, int R, bool POLY>
__global__ void __launch_bounds__(128) k_cells_warp(Params p, const int *__restrict__ in_list, const int *__restrict__ in_n,
                                                    int *__restrict__ out_list, int *__restrict__ out_n) {
  __shared__ double sx[4][16], sy[4][16];
  __shared__ int st[4][16];
  if (weights_graded(p)) return;
  typedef PolyRef<1, true> Poly;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n_in = *in_n;
  for (int idx = blockIdx.x * 4 + wib; idx < n_in; idx += gridDim.x * 4) {
    const int i = in_list[idx];
    Poly P{sx[wib], sy[wib], st[wib]};
    CellSearch<Poly> S;
    S.init(p, i, P);  // every lane writes the same box into the warp's polygon
    if (R0 >= 0 && !block_reload(p, S, P)) {  // (every lane reloads the same polygon) nothing usable: CellSearch takes the cell
      if (lane == 0) out_list[atomicAdd(out_n, 1)] = i;
      __syncwarp();
      continue;
    }
    __syncwarp();
    const int G = p.bG;
    const int cbx = min(max((int)((S.xi - p.px0) * p.binv), 0), G - 1);
    const int cby = min(max((int)((S.yi - p.py0) * p.binv), 0), G - 1);
    BlockRuns<R0, R> runs;
    runs.build(p, cbx, cby, true);
    const int T = runs.total();
    for (int t0 = 0; t0 < T && S.phase == 0; t0 += 32) {
      const int t = t0 + lane;
      const bool act = t < T;
      const int pos = act ? runs.position(t) : 0;
      const double Dx = p.xr[pos] - S.xi, Dy = p.yr[pos] - S.yi, wj = p.wr[pos];
      const double dd2 = Dx * Dx + Dy * Dy, dw = S.wi - wj, c = 0.5 * (dd2 + dw);
      const int jj = p.rm2s[pos];
      bool kills = act && dd2 == 0.0 && jj != i && (wj > S.wi || (wj == S.wi && jj < i));  // coincident, heavier / earlier
      bool pending = act && dd2 > 0.0;
      for (;;) {
        // every pending lane tests its candidate against the polygon as it is now
        unsigned long long in = 0ull;
        bool cut = false;
        if (pending) {
          if (c >= 0.0 && c * c >= S.R2 * dd2 * (1.0 + 1e-9)) pending = false;  // cannot reach the polygon any more
          else {
            double r2;
            in = S.sign_mask(p, P, jj, Dx, Dy, c, dd2, dw, r2);
            if (in == 0ull) kills = true;
            cut = in != 0ull && in != lowmask64(S.n);
            pending = cut;
          }
        }
        if (__any_sync(0xffffffffu, kills)) { S.n = 0; S.phase = 2; break; }
        const unsigned cm = __ballot_sync(0xffffffffu, cut);
        if (!cm) break;
        const int leader = __ffs(cm) - 1;
        if (lane == leader) {
          S.jc = jj; S.cDx = Dx; S.cDy = Dy; S.cc = c; S.cin = in;
          S.template clip<true>(p, P, 16);
          pending = false;
        }
        __syncwarp();
        // the leader's polygon state to everybody
        S.n = __shfl_sync(0xffffffffu, S.n, leader);
        S.status = __shfl_sync(0xffffffffu, S.status, leader);
        S.phase = __shfl_sync(0xffffffffu, S.phase, leader);
        S.R2 = __shfl_sync(0xffffffffu, S.R2, leader);
        P.ord = __shfl_sync(0xffffffffu, P.ord, leader);
        P.used = __shfl_sync(0xffffffffu, P.used, leader);
        if (S.phase != 0) break;  // the polygon outgrew the 16-vertex class
      }
    }
    __syncwarp();
    bool cert;
    if (S.phase == 0) cert = block_certified<R>(p, S, P, cbx, cby);
    else cert = S.status == 0;
    if (lane == 0) {
      if (cert) {
        const int n = S.n;
        if (n == 0 && p.abort_on_empty) p.flags[1] = 1;
        cell_emit(p, i, P, n);
        if (POLY) {
          p.poly_n[i] = n;
          for (int k = 0; k < n; ++k) {
            const size_t o = (size_t)k * p.N + i;
            p.poly_x[o] = P.X(k); p.poly_y[o] = P.Y(k); p.poly_t[o] = P.T(k);
          }
        }
      } else {
        out_list[atomicAdd(out_n, 1)] = i;
      }
    }
    __syncwarp();
  }
}

// ================================================================================================
// K3 for grid meshes: one thread per cell integrates the cell over its boundary segments
// (ma_seg.cuh); the polygon comes from K2 and is staged in shared memory (the chords of part B
// revisit every vertex).
// ================================================================================================
#ifndef MA_K3_MINBLOCKS
#define MA_K3_MINBLOCKS 5
#endif
// sel = 0: every cell of the tile.  With K2's block kernels the work is split so that it overlaps K2's tail (the passes
// over the ~15 % of the cells the 5 x 5 block cannot certify are latency-bound and leave the GPU mostly idle):
// sel = 1, launched on a side stream right after the first block kernel: the cells that kernel certified (cstate == 0);
// sel = 2, after K2 is complete: the rest (list[0 .. *list_n)).  With graded weights the block kernels did nothing:
// sel = 1 returns at once and sel = 2 takes every cell.
enum { SEG_ALL = 0, SEG_CERTIFIED = 1, SEG_REST = 2 };
// VD: the squares of the grid are split along either diagonal (p.diag; ma_seg.cuh) — explicit triangulations of an image.
template <int MAXV, int NT, int MODE, bool VD = false> __global__ void __launch_bounds__(NT, (MAXV <= 16 ? MA_K3_MINBLOCKS : 1))
k_seg(Params p, int sel, const int *__restrict__ list, const int *__restrict__ list_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + MAXV * NT;
  double *se = sy + MAXV * NT;  // per-edge accumulators of the Hessian's edge integrals (kantorovich mode)
  int i = p.cell_lo + blockIdx.x * NT + threadIdx.x;
  if (sel != SEG_ALL) {
    const bool graded = weights_graded(p);
    if (sel == SEG_CERTIFIED) {
      if (graded || i >= p.cell_hi || p.cstate[i] != 0) return;
    } else if (!graded) {
      const int idx = blockIdx.x * NT + threadIdx.x;
      if (idx >= *list_n) return;
      i = list[idx];
    }
  }
  if (i >= p.cell_hi) return;
  if (p.abort_on_empty && *p.abort_flag) return;  // line-search trial already rejected by K2 (an empty cell)
  PolyRef<NT, false> P{sx + threadIdx.x, sy + threadIdx.x, nullptr};  // the tags stay in global memory
  const int n = p.poly_n[i];
  for (int k = 0; k < n; ++k) {
    const size_t o = (size_t)k * p.N + i;
    P.X(k) = p.poly_x[o]; P.Y(k) = p.poly_y[o];
  }
  SegAcc acc;
  unsigned long long touched = cell_integrate_lines<MODE, NT, VD>(p, i, P, n, acc, p.hslot + (size_t)i * p.kmax, se + threadIdx.x,
                                                              [&](int k) { return p.poly_t[(size_t)k * p.N + i]; });
  if (MODE == MODE_KANTOROVICH) {
    p.mass[i] = acc.mass;
    p.fcell[i] = acc.mass * p.ws[i] - acc.cost;
    p.touched[i] = touched;
    int c = __popcll(touched);
    p.rowcnt[i] = c ? c + 1 : 0;
  } else {
    const double xi = p.xs[i], yi = p.ys[i], mass = acc.mass;
    double *o = p.mom + 6 * (size_t)i;
    o[0] = mass;
    o[1] = acc.m[0] + xi * mass;  // ∫ρ x = ∫ρ (xi + ux)
    o[2] = acc.m[1] + yi * mass;
    o[3] = acc.m[2] + 2 * xi * acc.m[0] + xi * xi * mass;
    o[4] = acc.m[3] + 2 * yi * acc.m[1] + yi * yi * mass;
    o[5] = acc.m[4] + xi * acc.m[1] + yi * acc.m[0] + xi * yi * mass;
  }
}
template <int MAXV, int NT, int MODE> constexpr size_t seg_smem_bytes() {
  return (size_t)MAXV * NT * (8 + 8 + (MODE == MODE_KANTOROVICH ? 8 : 0));
}

// ================================================================================================
// K3: one warp per cell; lanes take candidate faces round-robin
// ================================================================================================
constexpr int PIECES_WPB = 4;  // warps per block
template <int KMAX, int MAXV> constexpr size_t pieces_warp_bytes() {
  return (size_t)(4 * KMAX + 2 * MAXV * 32 + KMAX * 33) * 8 + (size_t)(MAXV * 32 + KMAX) * 4;
}

template <int KMAX, int MAXV, int MODE> __global__ void __launch_bounds__(PIECES_WPB * 32) k_pieces(Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = p.cell_lo + blockIdx.x * PIECES_WPB + warp;
  if (i >= p.cell_hi) return;  // warp-uniform; no block-level barrier below
  if (p.abort_on_empty && *p.abort_flag) return;  // line-search trial already rejected by K2 (an empty cell)
  unsigned char *base = smem_raw + (size_t)warp * pieces_warp_bytes<KMAX, MAXV>();
  double *d = reinterpret_cast<double *>(base);
  CellTable T;
  T.Dx = d; T.Dy = d + KMAX; T.C = d + 2 * KMAX; T.S = d + 3 * KMAX;
  double *px = d + 4 * KMAX, *py = px + MAXV * 32;
  double *hacc = py + MAXV * 32;  // [KMAX][33]
  int *ptag = reinterpret_cast<int *>(hacc + KMAX * 33);
  T.J = ptag + MAXV * 32;
  PolyRef<32> P{px + lane, py + lane, ptag + lane};

  const int k = p.nbr_cnt[i];
  T.k = k < 0 ? 0 : k;
  LaneAcc acc;
  lane_acc_zero(acc);
  if (k >= 0) {
    for (int s = lane; s < k; s += 32) cell_table_fill(p, i, s, T);
    if (MODE == MODE_KANTOROVICH)
      for (int s = 0; s < KMAX; ++s) hacc[s * 33 + lane] = 0.0;
    __syncwarp();
    int pbase = 0, vbase = 0;
    if (MODE == MODE_PIECES_FILL) {
      // per-lane output offsets: cell offset + exclusive prefix over lanes of the counts, recomputed
      // by a dry run (the enumeration is deterministic)
      LaneAcc dry;
      lane_acc_zero(dry);
      Params q = p;
      q.stats = 0;
      lane_pieces<32, MODE_PIECES_COUNT>(q, i, lane, 32, P, MAXV, T, hacc, 33, dry);
      int np = dry.npieces, nv = dry.nverts;
      int ep = np, ev = nv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, ep, o), b = __shfl_up_sync(0xffffffffu, ev, o);
        if (lane >= o) { ep += a; ev += b; }
      }
      pbase = p.pc_off[2 * i] + ep - np;
      vbase = p.pc_off[2 * i + 1] + ev - nv;
      __syncwarp();
    }
    lane_pieces<32, MODE>(p, i, lane, 32, P, MAXV, T, hacc + lane, 33, acc, pbase, vbase);
    __syncwarp();
  }
  // ---- per-cell reductions (fixed order => reproducible) ----
  if (MODE == MODE_KANTOROVICH) {
    double mass = warp_sum(acc.mass), cost = warp_sum(acc.cost);
    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)(acc.touched & 0xffffffffull));
    unsigned hi = KMAX > 32 ? __reduce_or_sync(0xffffffffu, (unsigned)(acc.touched >> 32)) : 0u;
    unsigned long long touched = ((unsigned long long)hi << 32) | lo;
    for (int s = lane; s < KMAX; s += 32) {
      double h = 0.0;
      if (s < T.k && ((touched >> s) & 1ull)) {
        const double *row = hacc + s * 33;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) h += row[l];
      }
      p.hslot[(size_t)i * KMAX + s] = h;
    }
    if (lane == 0) {
      p.mass[i] = mass;
      p.fcell[i] = mass * p.ws[i] - cost;
      p.touched[i] = touched;
      int c = __popcll(touched);
      p.rowcnt[i] = c ? c + 1 : 0;
    }
  } else if (MODE == MODE_MOMENTS1 || MODE == MODE_MOMENTS2) {
    double mass = warp_sum(acc.mass);
    double m0 = warp_sum(acc.m[0]), m1 = warp_sum(acc.m[1]);
    double m2 = 0, m3 = 0, m4 = 0;
    if (MODE == MODE_MOMENTS2) { m2 = warp_sum(acc.m[2]); m3 = warp_sum(acc.m[3]); m4 = warp_sum(acc.m[4]); }
    if (lane == 0) {
      const double xi = p.xs[i], yi = p.ys[i];
      double *o = p.mom + 6 * (size_t)i;
      o[0] = mass;
      o[1] = m0 + xi * mass;  // ∫ρ x = ∫ρ (xi + ux)
      o[2] = m1 + yi * mass;
      o[3] = m2 + 2 * xi * m0 + xi * xi * mass;
      o[4] = m3 + 2 * yi * m1 + yi * yi * mass;
      o[5] = m4 + xi * m1 + yi * m0 + xi * yi * mass;
    }
  } else if (MODE == MODE_PIECES_COUNT) {
    int np = acc.npieces, nv = acc.nverts;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      np += __shfl_xor_sync(0xffffffffu, np, o);
      nv += __shfl_xor_sync(0xffffffffu, nv, o);
    }
    if (lane == 0) { p.pc_count[2 * i] = np; p.pc_count[2 * i + 1] = nv; }
  }
  if (p.stats) {
    if (lane == 0 && k > 0) acc.cnt[CNT_SUMK] += k;
    for (int c = 0; c < CNT_N; ++c) {
      unsigned long long v = acc.cnt[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(&p.counters[c], v);
    }
  }
}

// ================================================================================================
// K4: CSR fill (internal Morton order).  Row i = { (nbr slot s, -hslot) : touched } ∪ { (i, Σ hslot) },
// columns ascending (what Eigen's setFromTriplets yields per row, kantorovich.hpp:120-121,137-139).
// ================================================================================================
// One warp per 32 consecutive rows: the rows' slot tables (hslot, nbr) are read with fully coalesced
// loads into shared memory (row stride KMAX + 1: conflict-free column access), each lane then sorts
// its row and the sorted rows are staged in shared memory again so that the warp writes the whole
// contiguous range [rowptr[first row], rowptr[last row + 1]) with coalesced stores.
// (Fetching only the slots a row uses — lanes of the unused ones sit the load out, a third less DRAM traffic — was
// measured: 0.122 ms against 0.108 ms, the shuffle + predicate per element costs more than the sectors save.)
template <int KMAX> constexpr int csr_wpb() { return KMAX <= 16 ? 4 : (KMAX <= 32 ? 2 : 1); }  // warps per block (48 KB of static shared memory)
template <int KMAX>
__global__ void __launch_bounds__(csr_wpb<KMAX>() * 32) k_csr_fill(int row_lo, int N, const int *__restrict__ nbr, const double *__restrict__ hslot,
                                                          const unsigned long long *__restrict__ touched,
                                                          const int *__restrict__ rowptr, int *__restrict__ col,
                                                          double *__restrict__ val, int cap, const int *__restrict__ skip_flag) {
  // cap: entries col / val can hold (the kernel is launched before the host knows nnz; what does not fit is dropped and
  // the host repeats the fill with larger arrays); skip_flag: the evaluation was abandoned (empty cell in a trial)
  if (skip_flag && *skip_flag) return;
  constexpr int LD = KMAX + 1, CSR_WPB = csr_wpb<KMAX>();
  __shared__ double sh_h[CSR_WPB][32 * LD];
  __shared__ int sh_j[CSR_WPB][32 * LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = row_lo + (blockIdx.x * CSR_WPB + warp) * 32;  // rows [row_lo, N): this context's Morton tile
  if (row0 >= N) return;  // warp-uniform
  const int nrows = min(32, N - row0);
  double *H = sh_h[warp];
  int *J = sh_j[warp];
  {
    const size_t base = (size_t)row0 * KMAX;
    const int cnt = nrows * KMAX;
    for (int q = lane; q < cnt; q += 32) {
      const int r = q / KMAX, s = q - r * KMAX;
      H[r * LD + s] = hslot[base + q];
      J[r * LD + s] = nbr[base + q];
    }
  }
  __syncwarp();
  const int i = row0 + lane;
  const bool live = lane < nrows;
  const unsigned long long t = live ? touched[i] : 0ull;
  const int o0 = rowptr[row0], o1 = rowptr[row0 + nrows];
  // each lane compacts + sorts its row in place (columns ascending, diagonal included)
  int n = 0;
  if (t) {
    double diag = 0.0;
    for (int s = 0; s < KMAX; ++s)
      if ((t >> s) & 1ull) {
        const double h = H[lane * LD + s];
        const int j = J[lane * LD + s];
        diag += h;
        int q = n++;  // insertion into the sorted prefix (q <= s, so unread slots are never overwritten)
        while (q > 0 && J[lane * LD + q - 1] > j) { J[lane * LD + q] = J[lane * LD + q - 1]; H[lane * LD + q] = H[lane * LD + q - 1]; --q; }
        J[lane * LD + q] = j; H[lane * LD + q] = -h;
      }
    int q = n++;
    while (q > 0 && J[lane * LD + q - 1] > i) { J[lane * LD + q] = J[lane * LD + q - 1]; H[lane * LD + q] = H[lane * LD + q - 1]; --q; }
    J[lane * LD + q] = i; H[lane * LD + q] = diag;
  }
  __syncwarp();
  // coalesced write of the warp's contiguous output range: entry e belongs to the LAST row whose
  // start is <= e (empty rows share their start with the row after them); lanes find it by counting
  // row starts broadcast with shuffles
  const int total = o1 - o0;
  const int my_start = live ? rowptr[i] - o0 : 0x7fffffff;
  for (int e0 = 0; e0 < total; e0 += 32) {
    const int e = e0 + lane;
    int r = -1;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) r += (__shfl_sync(0xffffffffu, my_start, rr) <= e) ? 1 : 0;
    r = max(r, 0);
    const int rs = __shfl_sync(0xffffffffu, my_start, r);
    if (e < total && o0 + e < cap) {
      col[o0 + e] = J[r * LD + (e - rs)];
      val[o0 + e] = H[r * LD + (e - rs)];
    }
  }
}

// caller-order views -----------------------------------------------------------------------------
__global__ void k_scatter_to_caller(const double *__restrict__ src_sorted, const int *__restrict__ perm, int n,
                                    double *__restrict__ dst_caller) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst_caller[perm[k]] = src_sorted[k];
}
__global__ void k_rowcnt_to_caller(const int *__restrict__ rowcnt_sorted, const int *__restrict__ pos, int n,
                                   int *__restrict__ cnt_caller) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt_caller[i] = rowcnt_sorted[pos[i]];
}
// caller rows [r0, n)
__global__ void k_csr_to_caller(int r0, int n, const int *__restrict__ rowptr_s, const int *__restrict__ col_s,
                                const double *__restrict__ val_s, const int *__restrict__ pos,
                                const int *__restrict__ perm, const int *__restrict__ rowptr_c, int *__restrict__ col_c,
                                double *__restrict__ val_c) {
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = pos[i];
  const int cnt = rowptr_c[i + 1] - rowptr_c[i];  // rows of other tiles are empty here (and their rowptr_s is not defined)
  if (cnt == 0) return;
  int s = rowptr_s[k], e = s + cnt, o = rowptr_c[i];
  for (int q = s; q < e; ++q) {  // insertion sort by caller column while copying
    int cj = perm[col_s[q]];
    double vj = val_s[q];
    int w = o + (q - s);
    while (w > o && col_c[w - 1] > cj) { col_c[w] = col_c[w - 1]; val_c[w] = val_c[w - 1]; --w; }
    col_c[w] = cj; val_c[w] = vj;
  }
}
// adjacency in caller order
__global__ void k_adj_count(int n, const int *__restrict__ nbr_cnt, const int *__restrict__ pos, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = max(nbr_cnt[pos[i]], 0);
}
__global__ void k_adj_fill(int n, int kmax, const int *__restrict__ nbr, const int *__restrict__ nbr_cnt,
                           const int *__restrict__ pos, const int *__restrict__ perm, const int *__restrict__ ptr,
                           int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = pos[i], c = max(nbr_cnt[k], 0), o = ptr[i];
  for (int s = 0; s < c; ++s) idx[o + s] = perm[nbr[(size_t)k * kmax + s]];
}

// cell polygons in caller order (ma_cells_get)
__global__ void k_cellpoly_count(int n, const int *__restrict__ poly_n, const int *__restrict__ pos, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = poly_n[pos[i]];
}
__global__ void k_cellpoly_fill(int n, const int *__restrict__ poly_n, const double *__restrict__ px,
                                const double *__restrict__ py, const int *__restrict__ pt, const double *__restrict__ xs,
                                const double *__restrict__ ys, const int *__restrict__ pos, const int *__restrict__ perm,
                                const int *__restrict__ ptr, double *__restrict__ xy, int *__restrict__ tag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = pos[i], m = poly_n[k], o = ptr[i];
  const double x0 = xs[k], y0 = ys[k];
  for (int v = 0; v < m; ++v) {
    const size_t s = (size_t)v * n + k;
    xy[2 * (size_t)(o + v)] = px[s] + x0;
    xy[2 * (size_t)(o + v) + 1] = py[s] + y0;
    const int t = pt[s];
    tag[o + v] = t >= 0 ? perm[t] : t;  // -1 bottom, -2 right, -3 top, -4 left side of the mesh box
  }
}

// ================================================================================================
// Rasterised Laguerre diagram (rasterization.hpp:403-547, draw_laguerre_diagram): every piece (cell ∩ face) is drawn
// into a w x h image with EXACT pixel coverage, value = coverage * (mean density at the piece's vertices) * colour of
// the cell.  One thread per piece: the piece, mapped to pixel coordinates, is clipped to every pixel square of its
// bounding box (four Sutherland–Hodgman passes) and the area added to the pixel.  (The reference reaches the same
// coverage with a DDA over the edges plus a per-pixel case analysis, after nudging near-integer coordinates by 1e-6;
// clipping needs neither.)
// ================================================================================================
constexpr int RAST_MAXV = 48;
__global__ void __launch_bounds__(128) k_raster_pieces(int np, const int *__restrict__ pc_cell, const int *__restrict__ pc_face,
                                                       const int *__restrict__ pc_ptr, const double *__restrict__ pc_xy,
                                                       const double *__restrict__ abc, const int *__restrict__ perm,
                                                       const double *__restrict__ colors, double x0, double y0, double sx,
                                                       double sy, int w, int h, double *__restrict__ image) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const int v0 = pc_ptr[p], n = min(pc_ptr[p + 1] - v0, RAST_MAXV - 4);  // (pieces have at most 40 vertices)
  if (n < 3) return;
  const int f = pc_face[p];
  const double a = abc[3 * (size_t)f], b = abc[3 * (size_t)f + 1], c0 = abc[3 * (size_t)f + 2];
  double X[RAST_MAXV], Y[RAST_MAXV], U[RAST_MAXV], V[RAST_MAXV], U2[RAST_MAXV], V2[RAST_MAXV];
  double ff = 0.0, bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
  for (int k = 0; k < n; ++k) {
    const double gx = pc_xy[2 * (size_t)(v0 + k)], gy = pc_xy[2 * (size_t)(v0 + k) + 1];
    ff += a * gx + b * gy + c0;
    X[k] = (gx - x0) * sx; Y[k] = (gy - y0) * sy;
    bx0 = fmin(bx0, X[k]); bx1 = fmax(bx1, X[k]); by0 = fmin(by0, Y[k]); by1 = fmax(by1, Y[k]);
  }
  const double value = ff / (double)n * colors[perm[pc_cell[p]]];
  const int ix0 = max((int)floor(bx0), 0), ix1 = min((int)floor(bx1), w - 1);
  const int iy0 = max((int)floor(by0), 0), iy1 = min((int)floor(by1), h - 1);
  for (int iy = iy0; iy <= iy1; ++iy)
    for (int ix = ix0; ix <= ix1; ++ix) {
      // clip (X, Y)[0..n) to [ix, ix+1] x [iy, iy+1]: keep s >= 0 for s = x - ix, ix+1 - x, y - iy, iy+1 - y
      int m = n;
      const double *sxp = X, *syp = Y;
      double *dxp = U, *dyp = V;
      for (int side = 0; side < 4 && m > 0; ++side) {
        const double A = side == 0 ? 1.0 : (side == 1 ? -1.0 : 0.0), B = side == 2 ? 1.0 : (side == 3 ? -1.0 : 0.0);
        const double C = side == 0 ? -(double)ix : (side == 1 ? (double)(ix + 1) : (side == 2 ? -(double)iy : (double)(iy + 1)));
        int q = 0;
        for (int k = 0; k < m; ++k) {
          const int kk = (k + 1 == m) ? 0 : k + 1;
          const double s0 = A * sxp[k] + B * syp[k] + C, s1 = A * sxp[kk] + B * syp[kk] + C;
          if (s0 >= 0.0) { dxp[q] = sxp[k]; dyp[q] = syp[k]; ++q; }
          if ((s0 >= 0.0) != (s1 >= 0.0)) {
            const double t = s0 / (s0 - s1);
            dxp[q] = sxp[k] + t * (sxp[kk] - sxp[k]); dyp[q] = syp[k] + t * (syp[kk] - syp[k]); ++q;
          }
        }
        m = q;
        sxp = dxp; syp = dyp;  // ping-pong between the two scratch polygons
        dxp = (dxp == U) ? U2 : U; dyp = (dyp == V) ? V2 : V;
      }
      if (m < 3) continue;
      double area = 0.0;
      for (int k = 0; k < m; ++k) {
        const int kk = (k + 1 == m) ? 0 : k + 1;
        area += sxp[k] * syp[kk] - sxp[kk] * syp[k];
      }
      area *= 0.5;
      if (area > 0.0) atomicAdd(&image[(size_t)iy * w + ix], area * value);
    }
}

__global__ void k_fill_bytes(unsigned long long *p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// DFMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;
}

}  // namespace ma

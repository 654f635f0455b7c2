// ma_amg.cuh — K5 preconditioner: aggregation multigrid on the quadtree of the Diracs.
//
// solve_laplacian_matrix (optimal_transport.hpp:41-87) factors the grounded Kantorovich Hessian with a sparse
// Cholesky.  On the device the solve is a conjugate gradient; with the Jacobi preconditioner of round 1 it needed
// ~2 800 iterations per solve at 100 k Diracs and ~8 800 at 1 M (iterations grow like sqrt(N): 31 of the 34 s of the
// 1 M-Dirac Newton solve).  The Hessian is the Laplacian of the Laguerre adjacency graph, whose nodes are already
// sorted along the Morton curve of a quadtree (K1): the children of one quadtree node are a contiguous run of rows.
// That gives a multigrid hierarchy for free:
//   * aggregates of level l = the non-empty quadtree nodes one level up (about 4 rows each), numbered in Morton
//     order; built ONCE per point set from the sorted leaf codes (k_amg_flags / k_amg_index);
//   * P = piecewise constant on the aggregates, coarse matrix = P^T A P assembled row by row (the fine rows of an
//     aggregate are contiguous: count, scan, fill with a small sorted merge per coarse row), every Newton step;
//   * V(1,1) cycle with damped Jacobi (omega), coarse correction scaled by alpha > 1 (the usual remedy for the
//     flat prolongation of unsmoothed aggregation), dense inverse on the last level (<= AMG_DENSE_MAX rows);
//   * symmetric, so plain PCG applies.  Measured (prototype on the c2 Hessian): 57 PCG iterations instead of 2 786
//     at 100 k rows for a relative residual of 1e-12.
// The grounded row (optimal_transport.hpp:62-63: the LAST index of the caller's ordering) is an identity row on
// the finest level; coarse levels see it as +1 on the diagonal of its aggregate.
// All reductions run in a fixed order (no atomics): the solve stays bit-reproducible.
#pragma once
#include <cooperative_groups.h>

#include "ma_kernels.cuh"

namespace ma {

constexpr int AMG_MAX_LEVELS = 16;
constexpr int AMG_DENSE_MAX = 80;    // the last level (64 rows: quadtree level 3) is solved with an explicit dense inverse;
                                     // at 256 rows the one-block Gauss-Jordan cost 6.5 ms per solve and the apply 35 us per
                                     // PCG iteration (profiles/r02n), a fifth of the Newton solve at 1 M Diracs
#ifndef MA_AMG_TAIL_ROWS
#define MA_AMG_TAIL_ROWS 4096
#endif
constexpr int AMG_TAIL_ROWS = MA_AMG_TAIL_ROWS;  // levels with at most this many rows run inside ONE block (k_amg_tail)
constexpr int AMG_TAIL_MAX = 8;
constexpr int AMG_ROW_CAP = 96;      // distinct columns one coarse row may have while it is merged

struct AmgLevel {
  int n = 0;
  const int *agg = nullptr;     // [n]  aggregate (row of the next level) of every row; null on the last level
  const int *cstart = nullptr;  // [n_next + 1] first row of every aggregate
  const int *rowptr = nullptr, *col = nullptr;
  const double *val = nullptr;
  double *dinv = nullptr;       // 1 / diagonal (0 at the grounded row of level 0)
  double *x = nullptr, *r = nullptr, *t = nullptr, *x2 = nullptr;
};

// ---- hierarchy (once per point set) ---------------------------------------------------------------------------
// flag[i] = 1 iff row i starts a new aggregate, i.e. its code's parent differs from the previous row's
__global__ void k_amg_flags(const unsigned *__restrict__ code, int n, int *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || (code[i] >> 2) != (code[i - 1] >> 2)) ? 1 : 0;
}
// excl = exclusive scan of flag: agg[i] = excl[i] + flag[i] - 1; the first row of an aggregate records cstart / code
__global__ void k_amg_index(const unsigned *__restrict__ code, const int *__restrict__ flag, const int *__restrict__ excl,
                            int n, int *__restrict__ agg, int *__restrict__ cstart, unsigned *__restrict__ ccode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = excl[i] + flag[i] - 1;
  agg[i] = a;
  if (flag[i]) { cstart[a] = i; ccode[a] = code[i] >> 2; }
  if (i == n - 1) cstart[a + 1] = n;
}

// ---- Galerkin product (every solve) -----------------------------------------------------------------------------
// One thread per coarse row: merges the mapped columns of its fine rows into a sorted list.  FILL = false counts the
// distinct columns, FILL = true writes them (and 1/diagonal).  `ground` >= 0 only on the finest level.
template <bool FILL>
__global__ void __launch_bounds__(128) k_amg_galerkin(int nc, const int *__restrict__ cstart, const int *__restrict__ agg,
                                                      const int *__restrict__ rowptr, const int *__restrict__ col,
                                                      const double *__restrict__ val, int ground, int *__restrict__ cnt,
                                                      const int *__restrict__ crowptr, int *__restrict__ ccol,
                                                      double *__restrict__ cval, double *__restrict__ cdinv, int *__restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int cj[AMG_ROW_CAP];
  double cv[AMG_ROW_CAP];
  int m = 0;
  bool overflow = false;
  auto add = [&](int j, double v) {
    int q = 0;
    while (q < m && cj[q] < j) ++q;
    if (q < m && cj[q] == j) { cv[q] += v; return; }
    if (m == AMG_ROW_CAP) { overflow = true; return; }
    for (int s = m; s > q; --s) { cj[s] = cj[s - 1]; cv[s] = cv[s - 1]; }
    cj[q] = j; cv[q] = v; ++m;
  };
  for (int i = cstart[c]; i < cstart[c + 1]; ++i) {
    if (i == ground) { add(c, 1.0); continue; }  // identity row
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int j = col[k];
      if (j == ground) continue;  // grounded column removed
      add(agg[j], val[k]);
    }
  }
  if (overflow) atomicOr(flag, 2);
  if (!FILL) { cnt[c] = m; return; }
  const int o = crowptr[c];
  double d = 0.0;
  for (int q = 0; q < m; ++q) {
    ccol[o + q] = cj[q]; cval[o + q] = cv[q];
    if (cj[q] == c) d = cv[q];
  }
  if (!(d > 0.0)) atomicOr(flag, 1);
  cdinv[c] = d > 0.0 ? 1.0 / d : 0.0;
}

// explicit inverse of the last level's matrix (n <= AMG_DENSE_MAX), one block: Gauss–Jordan without pivoting on
// the symmetric positive definite matrix; W is n x 2n scratch in global memory (L2-resident)
__global__ void __launch_bounds__(1024) k_amg_dense_inverse(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                            const double *__restrict__ val, double *__restrict__ W,
                                                            double *__restrict__ Ainv, int *__restrict__ flag) {
  const int w = 2 * n;
  for (int e = threadIdx.x; e < n * w; e += blockDim.x) {
    const int r = e / w, c = e - r * w;
    W[e] = (c == n + r) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < n; r += blockDim.x)
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) W[(size_t)r * w + col[k]] = val[k];
  __syncthreads();
  __shared__ double piv_inv;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      const double pv = W[(size_t)k * w + k];
      if (!(pv > 0.0)) atomicOr(flag, 1);
      piv_inv = pv != 0.0 ? 1.0 / pv : 0.0;
    }
    __syncthreads();
    const double pi = piv_inv;
    // eliminate column k from every other row: row_r -= (W[r][k] / pivot) * row_k, columns k+1 .. n+k (the rest is known)
    const int c0 = k + 1, c1 = n + k + 1, span = c1 - c0;
    for (int e = threadIdx.x; e < n * span; e += blockDim.x) {
      const int r = e / span, c = c0 + (e - r * span);
      if (r != k) W[(size_t)r * w + c] -= W[(size_t)r * w + k] * pi * W[(size_t)k * w + c];
    }
    __syncthreads();
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) W[(size_t)k * w + c] *= pi;
    __syncthreads();
  }
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e - r * n;
    Ainv[e] = W[(size_t)r * w + n + c];
  }
}

// ---- V-cycle ------------------------------------------------------------------------------------------------------
// down, part 1 (one thread per row): x = omega D^-1 r (first Jacobi sweep from 0), t = r - A x
__global__ void __launch_bounds__(256) k_amg_down(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                  const double *__restrict__ val, const double *__restrict__ dinv,
                                                  const double *__restrict__ r, double omega, int ground,
                                                  double *__restrict__ x, double *__restrict__ t) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i == ground) { x[i] = 0.0; t[i] = 0.0; continue; }
    const double ri = r[i];
    double acc = 0.0;
    const int k1 = rowptr[i + 1];
    for (int k = rowptr[i]; k < k1; k += 4) {
      int j[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = k + u < k1;
        j[u] = in ? col[k + u] : i;
        v[u] = in ? val[k + u] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u] * (dinv[j[u]] * r[j[u]]);  // dinv = 0 at the grounded index
    }
    x[i] = omega * dinv[i] * ri;
    t[i] = ri - omega * acc;
  }
}
// down, part 2 (one thread per aggregate): r_coarse = sum of t over the aggregate
__global__ void __launch_bounds__(256) k_amg_restrict(int nc, const int *__restrict__ cstart, const double *__restrict__ t,
                                                      double *__restrict__ rc) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int i = cstart[c]; i < cstart[c + 1]; ++i) s += t[i];
    rc[c] = s;
  }
}
// The coarse end of the V-cycle in ONE launch.  Levels of a few thousand rows are pure launch latency as separate
// kernels (3 per level, ~7 us each against ~1 us of work: 120 of the 280 us of a PCG iteration at 1 M rows,
// profiles/r02n); here one launch walks down from the first level with <= AMG_TAIL_ROWS rows to the dense level and
// back up, a barrier between the phases.  Same arithmetic in the same order as the kernels below.
struct AmgTail {
  int nlev;                    // lev[0] = first level of the tail ... lev[nlev - 1] = the dense level
  AmgLevel lev[AMG_TAIL_MAX];
  const double *Ainv;
};
// One thread-block CLUSTER of AMG_TAIL_CTAS blocks (8 x 1024 threads: a row or two per thread on the largest level of the
// tail, so a phase is about one dependent chain rowptr -> (col, val) -> gather long) with the hardware cluster barrier
// between the phases; one block with __syncthreads measured 60 us per cycle at 1 M Diracs, a quarter of a PCG iteration
// (profiles/r03b).  The vectors the phases hand to each other go through L2 (__ldcg / plain stores: the barrier orders
// them at cluster scope, L1 is bypassed on the reading side); the matrices are read-only.
constexpr int AMG_TAIL_CTAS = 8;
__global__ void __cluster_dims__(AMG_TAIL_CTAS, 1, 1) __launch_bounds__(1024) k_amg_tail(AmgTail T, double omega, double alpha) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = (int)cluster.block_rank() * (int)blockDim.x + (int)threadIdx.x, nt = AMG_TAIL_CTAS * (int)blockDim.x;
  for (int l = 0; l + 1 < T.nlev; ++l) {
    const AmgLevel &L = T.lev[l];
    for (int i = tid; i < L.n; i += nt) {  // k_amg_down
      const double ri = __ldcg(L.r + i);
      double acc = 0.0;
      const int k1 = L.rowptr[i + 1];
      for (int k = L.rowptr[i]; k < k1; k += 4) {
        int j[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = k + u < k1;
          j[u] = in ? L.col[k + u] : i;
          v[u] = in ? L.val[k + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u] * (L.dinv[j[u]] * __ldcg(L.r + j[u]));
      }
      L.x[i] = omega * L.dinv[i] * ri;
      L.t[i] = ri - omega * acc;
    }
    cluster.sync();
    const AmgLevel &C = T.lev[l + 1];
    for (int c = tid; c < C.n; c += nt) {  // k_amg_restrict
      double s = 0.0;
      for (int i = L.cstart[c]; i < L.cstart[c + 1]; ++i) s += __ldcg(L.t + i);
      C.r[c] = s;
    }
    cluster.sync();
  }
  {  // the dense level: e = Ainv r
    const AmgLevel &D = T.lev[T.nlev - 1];
    for (int i = tid; i < D.n; i += nt) {
      const double *row = T.Ainv + (size_t)i;
      double s = 0.0;
      for (int k = 0; k < D.n; ++k) s += row[(size_t)k * D.n] * __ldcg(D.r + k);
      D.x2[i] = s;
    }
    cluster.sync();
  }
  for (int l = T.nlev - 2; l >= 0; --l) {  // k_amg_up<false>
    const AmgLevel &L = T.lev[l];
    const double *ec = T.lev[l + 1].x2;
    for (int i = tid; i < L.n; i += nt) {
      const double ri = __ldcg(L.r + i);
      const double xi = __ldcg(L.x + i) + alpha * __ldcg(ec + L.agg[i]);
      double acc = 0.0;
      const int k1 = L.rowptr[i + 1];
      for (int k = L.rowptr[i]; k < k1; k += 4) {
        int j[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = k + u < k1;
          j[u] = in ? L.col[k + u] : i;
          v[u] = in ? L.val[k + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u] * (__ldcg(L.x + j[u]) + alpha * __ldcg(ec + L.agg[j[u]]));
      }
      L.x2[i] = xi + omega * L.dinv[i] * (ri - acc);
    }
    cluster.sync();
  }
}

// up (one thread per row): x' = x + alpha e_coarse[agg], then one Jacobi sweep x2 = x' + omega D^-1 (r - A x').
// On the finest level x2 is z = M^-1 r and the kernel also leaves the partial sums of r.z in part_rz (fixed grid).
template <bool FINE>
__global__ void __launch_bounds__(256) k_amg_up(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                const double *__restrict__ val, const double *__restrict__ dinv,
                                                const double *__restrict__ r, const double *__restrict__ x,
                                                const int *__restrict__ agg, const double *__restrict__ ec, double alpha,
                                                double omega, int ground, double *__restrict__ x2, double *__restrict__ part_rz) {
  __shared__ double sh[256 / 32];
  double rz = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double out = 0.0;
    if (i != ground) {
      const double ri = r[i];
      const double xi = x[i] + alpha * ec[agg[i]];
      double acc = 0.0;
      const int k1 = rowptr[i + 1];
      for (int k = rowptr[i]; k < k1; k += 4) {
        int j[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = k + u < k1;
          j[u] = in ? col[k + u] : i;
          v[u] = in ? val[k + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u] * ((j[u] == ground) ? 0.0 : x[j[u]] + alpha * ec[agg[j[u]]]);
      }
      out = xi + omega * dinv[i] * (ri - acc);
      rz += ri * out;
    }
    x2[i] = out;
  }
  if (FINE) {
    rz = warp_sum(rz);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = rz;
    __syncthreads();
    if (warp == 0) {
      double v = lane < 256 / 32 ? sh[lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) part_rz[blockIdx.x] = v;
    }
  }
}

}  // namespace ma

// ma_pcg.cuh — K5: Jacobi-preconditioned conjugate gradient on the grounded Kantorovich Hessian
// (replaces Eigen::SimplicialLLT / SPQR in solve_laplacian_matrix, optimal_transport.hpp:41-87).
//
// System: H[0:N-1,0:N-1] d = g[0:N-1] with the grounded index removed (d[ground] = 0).  The grounded
// row/column is handled implicitly: r, z, p, q are forced to 0 there, so column `ground` never
// contributes.  Two kernels per iteration, all scalars stay on the device:
//   A: beta = rz_new / rz_old; p_new = z + beta p_old (own rows, and on the fly for gathered columns);
//      q = H p_new; partial sums of p.q
//   B: alpha = rz / p.q; x += alpha p; r -= alpha q; z = r / diag; partial sums of r.z and r.r
// Dot products are reduced in a fixed order (per-block partials, every block re-sums them), so the
// solve is bit-reproducible.
#pragma once
#include "ma_amg.cuh"

namespace ma {

constexpr int PCG_NT = 256;
constexpr int PCG_MAX_BLOCKS = 1184;  // 8 per SM on 148 SMs

struct PcgState {
  int n, ground;
  const int *rowptr, *col;
  const double *val;
  double *dinv, *x, *r, *z, *p[2], *q;
  double *part_pq;  // [nblocks]
  double *part_rz;  // [2][nblocks]  (parity)
  double *part_rr;  // [nblocks]
  double *scal;     // [0..1] rz by parity, [2] gg (|g|^2), [3] last rr
  int nblocks;
  int *flag;        // bit0: zero diagonal
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < PCG_NT / 32 ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}
// every block sums the same partial array in the same order
__device__ __forceinline__ double sum_partials(const double *part, int nb, double *sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nb; i += PCG_NT) v += part[i];
  return block_sum(v, sh);
}

// dinv, r = g (0 at ground), z = r*dinv, p0 = z, x = 0, partial rz (parity 0), partial gg
// x0 != nullptr: start from x = scale * x0 (the Newton loop passes the previous direction: with a damped step tau
// the next direction is close to (1 - tau) times the last one); k_pcg_resid then turns r into b - A x.
__global__ void __launch_bounds__(PCG_NT) k_pcg_init(PcgState s, const double *__restrict__ g, double sign,
                                                      const double *__restrict__ x0, double scale) {
  __shared__ double sh[PCG_NT / 32];
  double rz = 0.0, gg = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double d = 0.0;
    for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k)
      if (s.col[k] == i) d = s.val[k];
    double ri = (i == s.ground) ? 0.0 : sign * g[i];
    double di = 0.0;
    if (i != s.ground) {
      if (d == 0.0) atomicOr(s.flag, 1);
      else di = 1.0 / d;
    }
    double zi = ri * di;
    s.dinv[i] = di; s.x[i] = (x0 && i != s.ground) ? scale * x0[i] : 0.0; s.r[i] = ri; s.z[i] = zi; s.p[0][i] = 0.0; s.p[1][i] = 0.0;
    rz += ri * zi; gg += ri * ri;
  }
  double a = block_sum(rz, sh), b = block_sum(gg, sh);
  if (threadIdx.x == 0) { s.part_rz[blockIdx.x] = a; s.part_rr[blockIdx.x] = b; }
}
// r = b - A x for a non-zero starting point (same grid as k_pcg_init: it overwrites that kernel's r.z partials)
__global__ void __launch_bounds__(PCG_NT) k_pcg_resid(PcgState s) {
  __shared__ double sh[PCG_NT / 32];
  const double *__restrict__ x = s.x;
  double rz = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double ri = 0.0;
    if (i != s.ground) {
      ri = s.r[i];
      for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k) ri -= s.val[k] * x[s.col[k]];  // x is 0 at the grounded index
    }
    const double zi = ri * s.dinv[i];
    s.r[i] = ri; s.z[i] = zi;
    rz += ri * zi;
  }
  double a = block_sum(rz, sh);
  if (threadIdx.x == 0) s.part_rz[blockIdx.x] = a;
}
__global__ void __launch_bounds__(PCG_NT) k_pcg_init2(PcgState s) {
  __shared__ double sh[PCG_NT / 32];
  double rz = sum_partials(s.part_rz, s.nblocks, sh), gg = sum_partials(s.part_rr, s.nblocks, sh);
  if (threadIdx.x == 0) { s.scal[0] = rz; s.scal[1] = rz; s.scal[2] = gg; s.scal[3] = gg; }
}

// iteration `it` (0-based): parity a = it & 1 reads p[a^1] (old) and writes p[a] (new)
__global__ void __launch_bounds__(PCG_NT) k_pcg_A(PcgState s, int it) {
  __shared__ double sh[PCG_NT / 32];
  const int a = it & 1;
  double beta = 0.0;
  if (it > 0) {
    double rz_new = sum_partials(s.part_rz + a * s.nblocks, s.nblocks, sh);  // written by B of it-1 at parity a
    double rz_old = s.scal[a ^ 1];
    beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) s.scal[a] = rz_new;
  }
  const double *__restrict__ pold = s.p[a ^ 1];
  double *__restrict__ pnew = s.p[a];
  const double *__restrict__ z = s.z;
  double pq = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double pi = 0.0, qi = 0.0;
    if (i != s.ground) {
      pi = z[i] + beta * pold[i];
      // 4 entries per trip, their loads issued together: a row is ~7 entries and the latency of this kernel is the
      // chain rowptr -> (col, val) -> (z, p_old) gathers, not bandwidth
      const int k1 = s.rowptr[i + 1];
      for (int k = s.rowptr[i]; k < k1; k += 4) {
        int j[4];
        double v[4], a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = k + u < k1;
          j[u] = in ? s.col[k + u] : i;
          v[u] = in ? s.val[k + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = z[j[u]] + beta * pold[j[u]];  // z, p_old are 0 at the grounded index
        qi += (v[0] * a[0] + v[1] * a[1]) + (v[2] * a[2] + v[3] * a[3]);
      }
    }
    pnew[i] = pi;
    s.q[i] = qi;
    pq += pi * qi;
  }
  double t = block_sum(pq, sh);
  if (threadIdx.x == 0) s.part_pq[blockIdx.x] = t;
}

// JACOBI = true: z = D^-1 r here; false: the multigrid V-cycle (ma_amg.cuh) that follows computes z and the r.z partials
template <bool JACOBI> __global__ void __launch_bounds__(PCG_NT) k_pcg_B(PcgState s, int it) {
  __shared__ double sh[PCG_NT / 32];
  const int a = it & 1;
  double pq = sum_partials(s.part_pq, s.nblocks, sh);
  // rz of this iteration: scal[a] for it > 0 is written by block 0 of kernel A (same stream, done)
  double rz = s.scal[a];
  double alpha = (pq != 0.0) ? rz / pq : 0.0;
  const double *__restrict__ p = s.p[a];
  double rz2 = 0.0, rr = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double xi = s.x[i] + alpha * p[i];
    double ri = s.r[i] - alpha * s.q[i];
    s.x[i] = xi; s.r[i] = ri;
    if (JACOBI) {
      double zi = ri * s.dinv[i];
      s.z[i] = zi;
      rz2 += ri * zi;
    }
    rr += ri * ri;
  }
  double t2 = block_sum(rr, sh);
  if (JACOBI) {
    double t1 = block_sum(rz2, sh);
    if (threadIdx.x == 0) s.part_rz[(a ^ 1) * s.nblocks + blockIdx.x] = t1;  // read by A of it+1 at parity a^1
  }
  if (threadIdx.x == 0) s.part_rr[blockIdx.x] = t2;
}
__global__ void __launch_bounds__(PCG_NT) k_pcg_rr(PcgState s) {
  __shared__ double sh[PCG_NT / 32];
  double rr = sum_partials(s.part_rr, s.nblocks, sh);
  if (threadIdx.x == 0) s.scal[3] = rr;
}

// vector helpers for the Newton loop
__global__ void k_axpy_to(int n, const double *__restrict__ x0, double alpha, const double *__restrict__ d,
                          double *__restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x0[i] + alpha * d[i];
}
__global__ void k_sub(int n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] - b[i];
}

}  // namespace ma

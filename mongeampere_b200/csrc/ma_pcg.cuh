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
#include "ma_kernels.cuh"

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

__global__ void __launch_bounds__(PCG_NT) k_pcg_B(PcgState s, int it) {
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
    double zi = ri * s.dinv[i];
    s.x[i] = xi; s.r[i] = ri; s.z[i] = zi;
    rz2 += ri * zi; rr += ri * ri;
  }
  double t1 = block_sum(rz2, sh), t2 = block_sum(rr, sh);
  if (threadIdx.x == 0) {
    s.part_rz[(a ^ 1) * s.nblocks + blockIdx.x] = t1;  // read by A of it+1 at parity a^1
    s.part_rr[blockIdx.x] = t2;
  }
}
__global__ void __launch_bounds__(PCG_NT) k_pcg_rr(PcgState s) {
  __shared__ double sh[PCG_NT / 32];
  double rr = sum_partials(s.part_rr, s.nblocks, sh);
  if (threadIdx.x == 0) s.scal[3] = rr;
}

// ------------------------------------------------------------------------------------------------
// The same iteration as ONE persistent kernel: all blocks are co-resident (cooperative launch) and
// separate the two phases with a grid barrier instead of a kernel boundary, and every block tests
// convergence itself (all blocks sum the same partials in the same order, so they agree bit for bit).
// At the sizes of this path one CG iteration is a few microseconds of work, so kernel launches —
// even replayed from a CUDA graph — dominated the two-kernel form.
// Arrays written inside the kernel are read through plain (coherent) loads: no __restrict__/const
// qualifiers on them, the barrier's fences make earlier writes of other blocks visible.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks, unsigned &gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = gen * nblocks;
    while (*(volatile unsigned *)bar < target) { }
    __threadfence();
  }
  __syncthreads();
}

// runs iterations it0 .. it0 + niter - 1 (it0 even), stops early once |r|^2 <= tol2;
// scal[3] = last |r|^2, scal[4] = number of iterations done in this launch
__global__ void __launch_bounds__(PCG_NT) k_pcg_persist(PcgState s, int it0, int niter, double tol2, unsigned *bar) {
  __shared__ double sh[PCG_NT / 32];
  unsigned gen = 0;
  const unsigned nb = gridDim.x;
  const int *__restrict__ rowptr = s.rowptr;
  const int *__restrict__ col = s.col;
  const double *__restrict__ val = s.val;
  const double *__restrict__ dinv = s.dinv;
  int done = 0;
  double rr_last = s.scal[3];
  for (int it = it0; it < it0 + niter; ++it) {
    const int a = it & 1;
    // ---- phase A: beta, p_new = z + beta p_old, q = H p_new, partial p.q ----
    double beta = 0.0;
    if (it > 0) {
      double rz_new = sum_partials(s.part_rz + a * s.nblocks, nb, sh);
      double rz_old = s.scal[a ^ 1];
      beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
      if (blockIdx.x == 0 && threadIdx.x == 0) s.scal[a] = rz_new;
    }
    {
      double *pold = s.p[a ^ 1], *pnew = s.p[a], *z = s.z;
      double pq = 0.0;
      for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += nb * PCG_NT) {
        double pi = 0.0, qi = 0.0;
        if (i != s.ground) {
          pi = z[i] + beta * pold[i];
          for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            int j = col[k];
            qi += val[k] * (z[j] + beta * pold[j]);
          }
        }
        pnew[i] = pi;
        s.q[i] = qi;
        pq += pi * qi;
      }
      double t = block_sum(pq, sh);
      if (threadIdx.x == 0) s.part_pq[blockIdx.x] = t;
    }
    grid_barrier(bar, nb, gen);
    // ---- phase B: alpha, x, r, z, partial r.z and r.r ----
    {
      double pq = sum_partials(s.part_pq, nb, sh);
      double rz = s.scal[a];
      double alpha = (pq != 0.0) ? rz / pq : 0.0;
      double *p = s.p[a];
      double rz2 = 0.0, rr = 0.0;
      for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += nb * PCG_NT) {
        double xi = s.x[i] + alpha * p[i];
        double ri = s.r[i] - alpha * s.q[i];
        double zi = ri * dinv[i];
        s.x[i] = xi; s.r[i] = ri; s.z[i] = zi;
        rz2 += ri * zi; rr += ri * ri;
      }
      double t1 = block_sum(rz2, sh), t2 = block_sum(rr, sh);
      if (threadIdx.x == 0) {
        s.part_rz[(a ^ 1) * s.nblocks + blockIdx.x] = t1;
        s.part_rr[blockIdx.x] = t2;
      }
    }
    grid_barrier(bar, nb, gen);
    ++done;
    rr_last = sum_partials(s.part_rr, nb, sh);  // identical in every block
    if (!(rr_last > tol2)) break;               // also stops on NaN
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { s.scal[3] = rr_last; s.scal[4] = (double)done; }
}

// ------------------------------------------------------------------------------------------------
// Single-reduction CG (Chronopoulos & Gear): ONE kernel per iteration.  With u = M^-1 r, w = A u,
// gamma = (r,u), delta = (w,u):   beta = gamma/gamma_prev,  alpha = gamma / (delta - beta gamma / alpha_prev),
//   p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;
// and both dot products of the NEXT iteration are taken in the same pass.  The SpMV w_new = A u_new
// needs the neighbours' updated u, which are recomputed on the fly from the previous iteration's
// buffers ( u_new_j = (r_j - alpha (w_j + beta s_j)) / diag_j ), so no grid-wide synchronisation sits
// between the vector update and the SpMV; r, s, w are double-buffered by iteration parity.
// One CG iteration of this path is microseconds of work, so halving the launches halves the solve.
// ------------------------------------------------------------------------------------------------
struct Cg1State {
  int n, ground, nblocks;
  const int *rowptr, *col;
  const double *val, *dinv;
  double *x, *p, *r[2], *s[2], *w[2];
  double *part;  // [2][3][nblocks]: gamma, delta, rr partials by parity of the buffers they describe (a kernel
                 // reads one parity at its start and writes the other at its end: blocks of one launch do not
                 // all run at the same time, so the two must not share storage)
  double *scal;  // [2][2]: (gamma, alpha) by parity; [4] = gg; [5] = last rr
};

// r0 = sign*g (0 at ground), s0 = p = x = 0, dinv; then w0 = A u0 and the first partials (second kernel)
__global__ void __launch_bounds__(PCG_NT) k_cg1_init(Cg1State s, const double *__restrict__ g, double sign, int *flag,
                                                      double *dinv_out) {
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double d = 0.0;
    for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k)
      if (s.col[k] == i) d = s.val[k];
    double di = 0.0;
    if (i != s.ground) {
      if (d == 0.0) atomicOr(flag, 1);
      else di = 1.0 / d;
    }
    dinv_out[i] = di;
    s.r[0][i] = (i == s.ground) ? 0.0 : sign * g[i];
    s.s[0][i] = 0.0; s.p[i] = 0.0; s.x[i] = 0.0;
  }
}
__global__ void __launch_bounds__(PCG_NT) k_cg1_init2(Cg1State s) {
  __shared__ double sh[PCG_NT / 32];
  const double *__restrict__ r = s.r[0];
  double ga = 0.0, de = 0.0, rr = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    double wi = 0.0;
    const double ri = r[i], ui = ri * s.dinv[i];
    if (i != s.ground)
      for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k) { int j = s.col[k]; wi += s.val[k] * (r[j] * s.dinv[j]); }
    s.w[0][i] = wi;
    ga += ri * ui; de += wi * ui; rr += ri * ri;
  }
  double a = block_sum(ga, sh), b = block_sum(de, sh), c = block_sum(rr, sh);
  if (threadIdx.x == 0) { s.part[blockIdx.x] = a; s.part[s.nblocks + blockIdx.x] = b; s.part[2 * s.nblocks + blockIdx.x] = c; }  // parity 0
}
// iteration k: reads the buffers of parity k & 1, writes parity (k + 1) & 1
__global__ void __launch_bounds__(PCG_NT) k_cg1_iter(Cg1State s, int k) {
  __shared__ double sh[PCG_NT / 32];
  const int o = k & 1, nw = o ^ 1;
  const double *pin = s.part + (size_t)o * 3 * s.nblocks;
  double *pout = s.part + (size_t)nw * 3 * s.nblocks;
  const double gamma = sum_partials(pin, s.nblocks, sh), delta = sum_partials(pin + s.nblocks, s.nblocks, sh);
  double beta = 0.0, alpha;
  if (k > 0) {
    const double gprev = s.scal[2 * nw], aprev = s.scal[2 * nw + 1];  // written by iteration k-1 at parity (k-1)&1 = nw
    beta = (gprev != 0.0) ? gamma / gprev : 0.0;
    const double den = delta - beta * gamma / aprev;
    alpha = (den != 0.0) ? gamma / den : 0.0;
  } else {
    alpha = (delta != 0.0) ? gamma / delta : 0.0;
  }
  const double *__restrict__ ro = s.r[o];
  const double *__restrict__ so = s.s[o];
  const double *__restrict__ wo = s.w[o];
  const double *__restrict__ dinv = s.dinv;
  double ga = 0.0, de = 0.0, rr = 0.0;
  for (int i = blockIdx.x * PCG_NT + threadIdx.x; i < s.n; i += gridDim.x * PCG_NT) {
    const double di = dinv[i];
    const double ui = ro[i] * di;
    const double pi = ui + beta * s.p[i];
    const double si = wo[i] + beta * so[i];
    const double ri = ro[i] - alpha * si;
    const double un = ri * di;
    double wi = 0.0;
    if (i != s.ground)
      for (int q = s.rowptr[i]; q < s.rowptr[i + 1]; ++q) {
        const int j = s.col[q];
        wi += s.val[q] * ((ro[j] - alpha * (wo[j] + beta * so[j])) * dinv[j]);  // u_new_j, 0 at the grounded index
      }
    s.p[i] = pi;
    s.x[i] += alpha * pi;
    s.s[nw][i] = si; s.r[nw][i] = ri; s.w[nw][i] = wi;
    ga += ri * un; de += wi * un; rr += ri * ri;
  }
  double a = block_sum(ga, sh), b = block_sum(de, sh), c = block_sum(rr, sh);
  if (threadIdx.x == 0) {
    pout[blockIdx.x] = a; pout[s.nblocks + blockIdx.x] = b; pout[2 * s.nblocks + blockIdx.x] = c;
    if (blockIdx.x == 0) { s.scal[2 * o] = gamma; s.scal[2 * o + 1] = alpha; }
  }
}
// |r|^2 of the buffers of parity `par` into scal[slot]
__global__ void __launch_bounds__(PCG_NT) k_cg1_rr(Cg1State s, int par, int slot) {
  __shared__ double sh[PCG_NT / 32];
  double rr = sum_partials(s.part + (size_t)par * 3 * s.nblocks + 2 * s.nblocks, s.nblocks, sh);
  if (threadIdx.x == 0) s.scal[slot] = rr;
}

// y = H x (full matrix, no grounding) — used for the residual check and by tests
__global__ void k_spmv(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                       const double *__restrict__ x, double *__restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
  y[i] = s;
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

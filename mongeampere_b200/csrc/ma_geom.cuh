// ma_geom.cuh — device-side geometry shared by the cell (K2) and piece (K3) kernels.
//
// Conventions
//   * All polygon coordinates are LOCAL to the Dirac y_i that owns the cell (u = x - y_i).  The
//     reference evaluates CGAL::radical_axis in global coordinates (predicates.hpp:46-52), whose
//     constant term cancels catastrophically for small cells; in local coordinates the bisector of
//     (i, j) is { u . D = c } with D = y_j - y_i and c = (|D|^2 + w_i - w_j) / 2, and
//     "strictly closer to i in power distance" (predicates.hpp:85-86) is  c - u . D > 0.
//   * A polygon is a cyclic list of vertices; vertex k is the START of edge k and tag[k] names the
//     line supporting edge k (kantorovich.hpp:95-102 uses the same convention: p[i] = R[i] ∩ R[i-1]).
//     New vertices are always recomputed as line ∩ line of the two ORIGINAL supporting lines
//     (vti.hpp:112-123), never interpolated, so no error accumulates over successive clips.
//   * Storage is column-per-thread in shared memory: element k of thread t lives at base[k * NT + t].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ma {

#define MA_DEV __host__ __device__ __forceinline__
// Warp votes used as explicit reconvergence points in the per-thread loops (the compiler does not
// reliably reconverge lanes that leave a data-dependent loop at different times).  On the host (the
// CPU emulation of tests/emu runs one "lane" at a time) they are the identity.
#ifdef __CUDA_ARCH__
#define MA_WARP_ANY(pred) __any_sync(0xffffffffu, (pred))
#define MA_WARP_COUNT(pred) __popc(__ballot_sync(0xffffffffu, (pred)))
#define MA_WARP_SYNC() __syncwarp()
#else
#define MA_WARP_ANY(pred) (pred)
#define MA_WARP_COUNT(pred) ((pred) ? 1 : 0)
#define MA_WARP_SYNC() ((void)0)
#endif

// PACKED = false: vertex k lives in slot k (a clip shifts the array).
// PACKED = true (capacity <= 16): vertex data never moves; the cyclic order of the slots is a packed
// list of 4-bit slot numbers held in a register, so that a clip is a handful of bit operations with
// no data-dependent control flow (what the SIMT lanes of K2 need to stay converged).
template <int NT, bool PACKED = false> struct PolyRef {
  static constexpr bool PACK = PACKED;
  double *x, *y;  // base pointers already offset by the thread's column
  int *tag;
  unsigned long long ord;  // PACKED: slot of vertex k in bits [4k, 4k+4)
  unsigned used;           // PACKED: bit s set <=> slot s holds a vertex
  MA_DEV int slot(int k) const { return PACKED ? (int)((ord >> (4 * k)) & 15ull) : k; }
  MA_DEV double &X(int k) const { return x[slot(k) * NT]; }
  MA_DEV double &Y(int k) const { return y[slot(k) * NT]; }
  MA_DEV int &T(int k) const { return tag[slot(k) * NT]; }
  MA_DEV double &SX(int s) const { return x[s * NT]; }  // by slot
  MA_DEV double &SY(int s) const { return y[s * NT]; }
  MA_DEV int &ST(int s) const { return tag[s * NT]; }
};

MA_DEV unsigned long long lowmask64(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }
MA_DEV int ctz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

// Solve { u.D = c } ∩ { u.n = cl }.
MA_DEV void line_isect(double Dx, double Dy, double c, double nx, double ny, double cl, double &ux, double &uy) {
  double det = Dx * ny - Dy * nx;
  double inv = 1.0 / det;
  ux = (c * ny - cl * Dy) * inv;
  uy = (Dx * cl - nx * c) * inv;
}

// One Sutherland–Hodgman pass of Tri_intersector::operator() (vti.hpp:168-198) done in place.
//   in   : bit k set  <=>  vertex k is strictly inside the half-plane (decided by the caller)
//   lineof(tag, nx, ny, cl) returns the supporting line { u.n = cl } of an existing edge.
// Returns the new vertex count (0: empty), or -1 on capacity overflow.
template <class Poly, class LineOf>
MA_DEV int clip_rebuild(const Poly &P, int n, int maxv, unsigned long long in, double Dx, double Dy, double c,
                        int newtag, LineOf lineof) {
  // locate the (single, after sanitising) run of outside vertices: [start, start + r)
  int start = -1;
  {
    bool prev = (in >> (n - 1)) & 1ull;
    for (int k = 0; k < n; ++k) {
      bool cur = (in >> k) & 1ull;
      if (!cur && prev) { start = k; break; }
      prev = cur;
    }
  }
  if (start < 0) return n;  // cannot happen when 0 < popcount(in) < n
  int r = 0;
  {
    int k = start;
    while (r < n && !((in >> k) & 1ull)) { ++r; k = (k + 1 == n) ? 0 : k + 1; }
  }
  int A = (start == 0) ? n - 1 : start - 1;  // edge A leaves the half-plane
  int last = start + r - 1;
  if (last >= n) last -= n;                  // edge `last` re-enters it
  int tA = P.T(A), tL = P.T(last);
  double nx, ny, cl, Xx, Xy, Yx, Yy;
  lineof(tA, nx, ny, cl);
  line_isect(Dx, Dy, c, nx, ny, cl, Xx, Xy);
  lineof(tL, nx, ny, cl);
  line_isect(Dx, Dy, c, nx, ny, cl, Yx, Yy);
  if (r == 1) {
    if (n + 1 > maxv) return -1;
    for (int k = n - 1; k > start; --k) { P.X(k + 1) = P.X(k); P.Y(k + 1) = P.Y(k); P.T(k + 1) = P.T(k); }
    P.X(start) = Xx; P.Y(start) = Xy; P.T(start) = newtag;
    P.X(start + 1) = Yx; P.Y(start + 1) = Yy; P.T(start + 1) = tL;
    return n + 1;
  }
  P.X(start) = Xx; P.Y(start) = Xy; P.T(start) = newtag;
  P.X(last) = Yx; P.Y(last) = Yy; P.T(last) = tL;
  if (r == 2) return n;
  int m = 0;
  for (int k = 0; k < n; ++k) {
    int off = k - start;
    if (off < 0) off += n;
    bool del = off >= 1 && off <= r - 2;
    if (!del) {
      if (m != k) { P.X(m) = P.X(k); P.Y(m) = P.Y(k); P.T(m) = P.T(k); }
      ++m;
    }
  }
  return m;
}

// The same pass on a PACKED polygon, without data-dependent branches: the run of outside vertices
// [start, start + r) is located with bit operations on `in`, its first slot is reused for the vertex
// where the polygon leaves the half-plane, its last slot (or a free one when r = 1) for the vertex
// where it re-enters, and the new cyclic order is the inside run followed by those two.
template <class Poly, class LineOf>
MA_DEV int clip_packed(Poly &P, int n, int maxv, unsigned long long in, double Dx, double Dy, double c, int newtag,
                       LineOf lineof) {
  const unsigned long long full = lowmask64(n), out = ~in & full;
  const unsigned long long prev = ((out << 1) | (out >> (n - 1))) & full;  // bit k = out[k-1]
  const unsigned long long starts = out & ~prev;
  if (starts == 0ull) return n;  // cannot happen when 0 < popcount(in) < n
  const int start = ctz64(starts);
  const unsigned long long rot = start ? (((out >> start) | (out << (n - start))) & full) : out;
  const int r = ctz64(~rot);  // length of the outside run
  int A = start - 1; if (A < 0) A += n;
  int last = start + r - 1; if (last >= n) last -= n;
  const int sS = P.slot(start), sL = P.slot(last);
  const int tA = P.T(A), tL = P.ST(sL);
  double nx, ny, cl, Xx, Xy, Yx, Yy;
  lineof(tA, nx, ny, cl);
  line_isect(Dx, Dy, c, nx, ny, cl, Xx, Xy);
  lineof(tL, nx, ny, cl);
  line_isect(Dx, Dy, c, nx, ny, cl, Yx, Yy);
  unsigned used = P.used;
  for (int t = 1; t + 1 < r; ++t) {  // slots strictly inside the run become free
    int k = start + t; if (k >= n) k -= n;
    used &= ~(1u << P.slot(k));
  }
  int sY = sL;
  if (r == 1) {
    const unsigned fr = ~used & (unsigned)lowmask64(maxv);
    if (fr == 0u) return -1;
    sY = ctz64((unsigned long long)fr);
    used |= 1u << sY;
  }
  const int m = n - r;  // inside vertices
  if (m + 2 > maxv) return -1;
  P.SX(sS) = Xx; P.SY(sS) = Xy; P.ST(sS) = newtag;
  P.SX(sY) = Yx; P.SY(sY) = Yy; P.ST(sY) = tL;
  int s0 = last + 1; if (s0 >= n) s0 -= n;  // first inside vertex
  const unsigned long long o = P.ord & lowmask64(4 * n);
  const unsigned long long ro = s0 ? (((o >> (4 * s0)) | (o << (4 * (n - s0)))) & lowmask64(4 * n)) : o;
  P.ord = (ro & lowmask64(4 * m)) | ((unsigned long long)sS << (4 * m)) | ((unsigned long long)sY << (4 * (m + 1)));
  P.used = used;
  return m + 2;
}

// ------------------------------------------------------------------------------------------------
// double-double arithmetic for the robust sign fallback (the role of CGAL::Filtered_predicate's
// exact stage, predicates.hpp:159-167).  Error-free transformations with FMA.
// ------------------------------------------------------------------------------------------------
struct dd { double hi, lo; };
MA_DEV dd dd_from(double a) { return dd{a, 0.0}; }
MA_DEV dd two_sum(double a, double b) {
  double s = a + b, bb = s - a;
  return dd{s, (a - (s - bb)) + (b - bb)};
}
MA_DEV dd two_prod(double a, double b) {
  double p = a * b;
  return dd{p, fma(a, b, -p)};
}
MA_DEV dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = two_sum(s.hi, s.lo);  // fast renormalisation is not valid in general; use two_sum
  s.lo += t.lo;
  return two_sum(s.hi, s.lo);
}
MA_DEV dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
MA_DEV dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
MA_DEV dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return two_sum(p.hi, p.lo);
}
MA_DEV dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return two_sum(p.hi, p.lo);
}
MA_DEV dd dd_sub_dd(double a, double b) { return two_sum(a, -b); }  // exact a - b

// A line { u.n = cl } with double-double coefficients, built from ORIGINAL inputs:
//   bisector of (i, j):  n = y_j - y_i (exact as dd), cl = (|n|^2 + w_i - w_j)/2
//   mesh edge (p, q) in local coordinates: n = (p_y - q_y, q_x - p_x), cl = n . p
struct ddline { dd nx, ny, cl; };

MA_DEV ddline dd_bisector(double xi, double yi, double wi, double xj, double yj, double wj) {
  ddline l;
  l.nx = dd_sub_dd(xj, xi);
  l.ny = dd_sub_dd(yj, yi);
  dd d2 = dd_add(dd_mul(l.nx, l.nx), dd_mul(l.ny, l.ny));
  dd s = dd_add(d2, dd_sub_dd(wi, wj));
  l.cl = dd{0.5 * s.hi, 0.5 * s.lo};
  return l;
}
// mesh edge through global points (px,py)->(qx,qy), expressed in coordinates local to (xi, yi)
MA_DEV ddline dd_mesh_edge(double xi, double yi, double px, double py, double qx, double qy) {
  ddline l;
  l.nx = dd_sub_dd(py, qy);
  l.ny = dd_sub_dd(qx, px);
  dd ux = dd_sub_dd(px, xi), uy = dd_sub_dd(py, yi);
  l.cl = dd_add(dd_mul(l.nx, ux), dd_mul(l.ny, uy));
  return l;
}
// sign of  cT - uT.nT  at the vertex u = A ∩ B:   u = (cA nB.y - cB nA.y, nA.x cB - nB.x cA) / det,
// det = nA.x nB.y - nA.y nB.x, so  sign = sign( (cT det - nT.x X - nT.y Y) ) * sign(det)
// (the 3x3 determinant form of SURVEY App. A.3).  Returns +1 inside, -1 outside, 0 tie/unresolved.
MA_DEV int dd_side(const ddline &A, const ddline &B, const ddline &T) {
  dd det = dd_sub(dd_mul(A.nx, B.ny), dd_mul(A.ny, B.nx));
  dd X = dd_sub(dd_mul(A.cl, B.ny), dd_mul(B.cl, A.ny));
  dd Y = dd_sub(dd_mul(A.nx, B.cl), dd_mul(B.nx, A.cl));
  dd t1 = dd_mul(T.cl, det), t2 = dd_mul(T.nx, X), t3 = dd_mul(T.ny, Y);
  dd v = dd_sub(dd_sub(t1, t2), t3);
  double mag = fabs(t1.hi) + fabs(t2.hi) + fabs(t3.hi);
  if (!(fabs(v.hi) > 1e-26 * mag)) return 0;
  int sv = v.hi > 0 ? 1 : -1;
  int sd = det.hi > 0 ? 1 : (det.hi < 0 ? -1 : 0);
  return sv * sd;
}
// vertex is an original mesh vertex (exact point): sign of cT - u.nT with u = p - y_i
MA_DEV int dd_side_point(double xi, double yi, double px, double py, const ddline &T) {
  dd ux = dd_sub_dd(px, xi), uy = dd_sub_dd(py, yi);
  dd t2 = dd_mul(T.nx, ux), t3 = dd_mul(T.ny, uy);
  dd v = dd_sub(dd_sub(T.cl, t2), t3);
  double mag = fabs(T.cl.hi) + fabs(t2.hi) + fabs(t3.hi);
  if (!(fabs(v.hi) > 1e-28 * mag)) return 0;
  return v.hi > 0 ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------
// Quadratures of quadrature.hpp on one fan triangle (a, b, c), local coordinates.
// ------------------------------------------------------------------------------------------------
// Albrecht–Collatz points (quadrature.hpp:32-40): weights 1/30 at the three edge midpoints and
// 9/30 at (1/6,2/3), (2/3,1/6), (1/6,1/6).
template <class F> MA_DEV void albrecht_collatz_points(double ax, double ay, double bx, double by, double cx, double cy, F f) {
  const double h = 0.5, s = 1.0 / 6.0, t = 2.0 / 3.0;
  double ux = bx - ax, uy = by - ay, vx = cx - ax, vy = cy - ay;
  f(ax + h * ux + h * vx, ay + h * uy + h * vy, 1.0 / 30.0);
  f(ax + h * ux, ay + h * uy, 1.0 / 30.0);
  f(ax + h * vx, ay + h * vy, 1.0 / 30.0);
  f(ax + s * ux + t * vx, ay + s * uy + t * vy, 9.0 / 30.0);
  f(ax + s * vx + t * ux, ay + s * vy + t * uy, 9.0 / 30.0);
  f(ax + s * ux + s * vx, ay + s * uy + s * vy, 9.0 / 30.0);
}

MA_DEV uint32_t morton_compact1(uint32_t v) {
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0f0f0f0fu;
  v = (v | (v >> 4)) & 0x00ff00ffu;
  v = (v | (v >> 8)) & 0x0000ffffu;
  return v;
}
MA_DEV uint32_t morton_part1(uint32_t v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
MA_DEV uint32_t morton2(uint32_t x, uint32_t y) { return morton_part1(x) | (morton_part1(y) << 1); }

// Offsets (ox, oy) of the bins of ring r = 0..8 around a bin, ring after ring, nearest first inside a ring
// (the ring walk of K2 then meets the true neighbours early and wastes fewer clips).  Ring r starts at entry
// (2r-1)^2 for r >= 1 (entry 0 for r = 0) and has 8r entries.
#define MA_RING_TABLE_VALUES \
    0, 0, 0, -1, -1, 0, 1, 0, 0, 1, -1, -1, 1, -1, -1, 1, 1, 1, 0, -2, -2, 0, 2, 0, 0, 2, -1, -2, 1, -2, -2, -1, \
    2, -1, -2, 1, 2, 1, -1, 2, 1, 2, -2, -2, 2, -2, -2, 2, 2, 2, 0, -3, -3, 0, 3, 0, 0, 3, -1, -3, 1, -3, -3, -1, \
    3, -1, -3, 1, 3, 1, -1, 3, 1, 3, -2, -3, 2, -3, -3, -2, 3, -2, -3, 2, 3, 2, -2, 3, 2, 3, -3, -3, 3, -3, -3, 3, \
    3, 3, 0, -4, -4, 0, 4, 0, 0, 4, -1, -4, 1, -4, -4, -1, 4, -1, -4, 1, 4, 1, -1, 4, 1, 4, -2, -4, 2, -4, -4, -2, \
    4, -2, -4, 2, 4, 2, -2, 4, 2, 4, -3, -4, 3, -4, -4, -3, 4, -3, -4, 3, 4, 3, -3, 4, 3, 4, -4, -4, 4, -4, -4, 4, \
    4, 4, 0, -5, -5, 0, 5, 0, 0, 5, -1, -5, 1, -5, -5, -1, 5, -1, -5, 1, 5, 1, -1, 5, 1, 5, -2, -5, 2, -5, -5, -2, \
    5, -2, -5, 2, 5, 2, -2, 5, 2, 5, -3, -5, 3, -5, -5, -3, 5, -3, -5, 3, 5, 3, -3, 5, 3, 5, -4, -5, 4, -5, -5, \
    -4, 5, -4, -5, 4, 5, 4, -4, 5, 4, 5, -5, -5, 5, -5, -5, 5, 5, 5, 0, -6, -6, 0, 6, 0, 0, 6, -1, -6, 1, -6, -6, \
    -1, 6, -1, -6, 1, 6, 1, -1, 6, 1, 6, -2, -6, 2, -6, -6, -2, 6, -2, -6, 2, 6, 2, -2, 6, 2, 6, -3, -6, 3, -6, \
    -6, -3, 6, -3, -6, 3, 6, 3, -3, 6, 3, 6, -4, -6, 4, -6, -6, -4, 6, -4, -6, 4, 6, 4, -4, 6, 4, 6, -5, -6, 5, \
    -6, -6, -5, 6, -5, -6, 5, 6, 5, -5, 6, 5, 6, -6, -6, 6, -6, -6, 6, 6, 6, 0, -7, -7, 0, 7, 0, 0, 7, -1, -7, 1, \
    -7, -7, -1, 7, -1, -7, 1, 7, 1, -1, 7, 1, 7, -2, -7, 2, -7, -7, -2, 7, -2, -7, 2, 7, 2, -2, 7, 2, 7, -3, -7, \
    3, -7, -7, -3, 7, -3, -7, 3, 7, 3, -3, 7, 3, 7, -4, -7, 4, -7, -7, -4, 7, -4, -7, 4, 7, 4, -4, 7, 4, 7, -5, \
    -7, 5, -7, -7, -5, 7, -5, -7, 5, 7, 5, -5, 7, 5, 7, -6, -7, 6, -7, -7, -6, 7, -6, -7, 6, 7, 6, -6, 7, 6, 7, \
    -7, -7, 7, -7, -7, 7, 7, 7, 0, -8, -8, 0, 8, 0, 0, 8, -1, -8, 1, -8, -8, -1, 8, -1, -8, 1, 8, 1, -1, 8, 1, 8, \
    -2, -8, 2, -8, -8, -2, 8, -2, -8, 2, 8, 2, -2, 8, 2, 8, -3, -8, 3, -8, -8, -3, 8, -3, -8, 3, 8, 3, -3, 8, 3, \
    8, -4, -8, 4, -8, -8, -4, 8, -4, -8, 4, 8, 4, -4, 8, 4, 8, -5, -8, 5, -8, -8, -5, 8, -5, -8, 5, 8, 5, -5, 8, \
    5, 8, -6, -8, 6, -8, -8, -6, 8, -6, -8, 6, 8, 6, -6, 8, 6, 8, -7, -8, 7, -8, -8, -7, 8, -7, -8, 7, 8, 7, -7, \
    8, 7, 8, -8, -8, 8, -8, -8, 8, 8, 8
constexpr int MA_RING_TABLE_RMAX = 8;
static const signed char ring_table_host[] = {MA_RING_TABLE_VALUES};
#ifdef __CUDACC__
__device__ const signed char ring_table_dev[] = {MA_RING_TABLE_VALUES};
#endif
MA_DEV void ring_offset(int r, int q, int &ox, int &oy) {
  const int e = 2 * ((r == 0 ? 0 : (2 * r - 1) * (2 * r - 1)) + q);
#ifdef __CUDA_ARCH__
  ox = ring_table_dev[e]; oy = ring_table_dev[e + 1];
#else
  ox = ring_table_host[e]; oy = ring_table_host[e + 1];
#endif
}

// Can a half-plane of a site at squared distance >= d2 whose weight satisfies w_i - w_j >= dw cut a
// polygon contained in the disk of squared radius R2 around y_i?  The bisector's signed distance
// from y_i is t = (d^2 + w_i - w_j) / (2 d); it cannot cut iff t >= R.
MA_DEV bool cannot_cut(double d2, double dw, double R2) {
  double s = d2 + fmin(dw, 0.0);
  return s > 0.0 && s * s >= 4.0 * R2 * d2 * (1.0 + 1e-9);
}


// ------------------------------------------------------------------------------------------------
// Per-node supporting planes of the lifted sites (the "per-bin max-weight bound" made gradient-aware).
// For node B with centre zc and any vector G,   alpha_B = min_{j in B} ( |t_j|^2 - w_j + G . t_j ),
// t_j = y_j - zc, gives for every point p (P = p - zc) and every site j of B
//   pow_j(p) = |P - t_j|^2 - w_j = |P|^2 - (2P + G) . t_j + (|t_j|^2 - w_j + G . t_j)
//            >= |P|^2 + alpha_B - (|2 Px + Gx| + |2 Py + Gy|) * S/2 .
// With G = the least-squares gradient of the weights over B the bound stays tight when the weights
// have a gradient (then a Laguerre cell lies far from its own Dirac and any bound made of
// "max weight + distance to the box" alone keeps a disk of radius ~|grad w| alive).
// ------------------------------------------------------------------------------------------------
union dbits { double d; unsigned long long u; };
// order-preserving map double -> uint64 (so that atomicMin on the key is a min on the double)
MA_DEV unsigned long long dkey(double v) {
  dbits b; b.d = v;
  return (b.u >> 63) ? ~b.u : (b.u | 0x8000000000000000ull);
}
MA_DEV double dkey_inv(unsigned long long k) {
  dbits b;
  b.u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return b.d;
}
// least-squares gradient of w over a node from its moment sums (about any common origin);
// false when the node holds too few / too degenerate sites (the caller then inherits the parent's)
MA_DEV bool node_gradient(double n, double sx, double sy, double sxx, double syy, double sxy, double sw, double sxw,
                          double syw, double &Gx, double &Gy) {
  if (n < 6.0) return false;
  double inv = 1.0 / n;
  double cxx = sxx - sx * sx * inv, cyy = syy - sy * sy * inv, cxy = sxy - sx * sy * inv;
  double cxw = sxw - sx * sw * inv, cyw = syw - sy * sw * inv;
  double det = cxx * cyy - cxy * cxy;
  if (!(cxx > 0.0) || !(cyy > 0.0) || !(det > 1e-2 * cxx * cyy)) return false;
  Gx = (cxw * cyy - cyw * cxy) / det;
  Gy = (cyw * cxx - cxw * cxy) / det;
  return Gx - Gx == 0.0 && Gy - Gy == 0.0;  // finite
}
MA_DEV size_t level_offset(int l) { return (((size_t)1 << (2 * l)) - 1) / 3; }

}  // namespace ma

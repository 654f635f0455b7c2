// ma_seg.cuh — boundary-segment evaluation of one Laguerre cell against a GRID source mesh.
//
// What it replaces.  The reference enumerates the pieces (cell ∩ face), fans each one from its
// vertex 0 and applies a quadrature that is EXACT for the integrand's degree: the centroid rule for
// the linear density (quadrature.hpp:69-77), Albrecht–Collatz for rho(p)|p - y_v|^2 (a cubic,
// quadrature.hpp:24-42, kantorovich.hpp:126-131, lloyd.hpp:57-67,100-121) and the midpoint rule on
// the Laguerre edges (quadrature.hpp:79-85, kantorovich.hpp:110-122).  Because every rule is exact,
// the numbers it produces are the integrals themselves, and any other exact evaluation gives the same
// result to rounding.  Here the same integrals are taken over the BOUNDARY of the pieces with the
// divergence theorem, which needs no polygon clipping at all:
//
//   for a homogeneous polynomial q of degree d (about the Dirac y_i = local origin) and a polygon P,
//        ∫_P q = 1/(d+2) * Σ_{edges e of P} h_e ∫_e q ds ,   h_e = signed distance origin -> line(e).
//
//   On face T the density is rho_T(u) = lam_T(u) + r_T  (lam_T linear homogeneous, r_T = rho_T(y_i)).
//   Summing over the pieces of one cell, the boundary splits into
//   (A) sub-segments of the CELL's own edges, each inside one face: they contribute
//         h_s * [ ∫ lam g /(dg+3) + r ∫ g /(dg+2) ]      for every monomial g of degree dg,
//       and, on a Laguerre edge, the Hessian term ∫ rho ds (kantorovich.hpp:117-121);
//   (B) pieces of interior MESH edges inside the cell, seen once from each side.  rho is continuous,
//       so rho_T1 - rho_T2 = kappa * (n.u - h) with kappa = n.(grad rho_T1 - grad rho_T2), and the two
//       sides collapse to the single term      - kappa h^2 ∫_e g ds / ((dg+2)(dg+3)).
//
//   g runs over {1, |u|^2} for kantorovich (mass, cost) and {1, ux, uy [, ux², uy², ux uy]} for the
//   moments.  All line integrals are of degree <= 3 and are taken with Simpson's rule (exact).
//
// On a grid the families of mesh lines (x = const, y = const, the diagonals of either direction) are enumerated directly:
// no polygon is clipped and no orientation predicate decides anything, so the result depends continuously
// on the input.  One THREAD handles one cell.  SURVEY.md §7.2 "two regimes", DESIGN.md §3.
#pragma once
#include "ma_block.cuh"
#include "ma_warm.cuh"

namespace ma {

struct SegAcc {
  double mass, cost;
  double m[5];  // ∫ρ ux, ∫ρ uy, ∫ρ ux², ∫ρ uy², ∫ρ ux uy   (local coordinates)
};

MA_DEV int seg_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// (A) one sub-segment (ax,ay)->(bx,by) of a cell edge inside the face with rho = a ux + b uy + r;
// wgt = h_e * length of the sub-segment.
template <int MODE>
MA_DEV void seg_item_cell(double ax, double ay, double bx, double by, double a, double b, double r, double wgt,
                          SegAcc &acc) {
  const double lA = a * ax + b * ay, lB = a * bx + b * by, lm = 0.5 * (lA + lB);
  const double mx = 0.5 * (ax + bx), my = 0.5 * (ay + by);
  acc.mass += wgt * (lm * (1.0 / 3.0) + r * 0.5);
  if (MODE == MODE_KANTOROVICH) {
    const double qA = ax * ax + ay * ay, qB = bx * bx + by * by, qm = mx * mx + my * my;
    acc.cost += wgt * ((lA * qA + 4.0 * (lm * qm) + lB * qB) * (1.0 / 30.0) + r * ((qA + 4.0 * qm + qB) * (1.0 / 24.0)));
  } else {
    acc.m[0] += wgt * ((lA * ax + 4.0 * (lm * mx) + lB * bx) * (1.0 / 24.0) + r * (mx * (1.0 / 3.0)));
    acc.m[1] += wgt * ((lA * ay + 4.0 * (lm * my) + lB * by) * (1.0 / 24.0) + r * (my * (1.0 / 3.0)));
    if (MODE == MODE_MOMENTS2) {
      const double xxA = ax * ax, xxB = bx * bx, xxm = mx * mx, yyA = ay * ay, yyB = by * by, yym = my * my;
      const double xyA = ax * ay, xyB = bx * by, xym = mx * my;
      acc.m[2] += wgt * ((lA * xxA + 4.0 * (lm * xxm) + lB * xxB) * (1.0 / 30.0) + r * ((xxA + 4.0 * xxm + xxB) * (1.0 / 24.0)));
      acc.m[3] += wgt * ((lA * yyA + 4.0 * (lm * yym) + lB * yyB) * (1.0 / 30.0) + r * ((yyA + 4.0 * yym + yyB) * (1.0 / 24.0)));
      acc.m[4] += wgt * ((lA * xyA + 4.0 * (lm * xym) + lB * xyB) * (1.0 / 30.0) + r * ((xyA + 4.0 * xym + xyB) * (1.0 / 24.0)));
    }
  }
}

// (B) one piece (ax,ay)->(bx,by) of an interior mesh edge; kh2l = kappa * h^2 * length.
template <int MODE>
MA_DEV void seg_item_mesh(double ax, double ay, double bx, double by, double kh2l, SegAcc &acc) {
  const double mx = 0.5 * (ax + bx), my = 0.5 * (ay + by);
  acc.mass -= kh2l * (1.0 / 6.0);
  if (MODE == MODE_KANTOROVICH) {
    const double qA = ax * ax + ay * ay, qB = bx * bx + by * by, qm = mx * mx + my * my;
    acc.cost -= kh2l * ((qA + 4.0 * qm + qB) * (1.0 / 120.0));
  } else {
    acc.m[0] -= kh2l * (mx * (1.0 / 12.0));
    acc.m[1] -= kh2l * (my * (1.0 / 12.0));
    if (MODE == MODE_MOMENTS2) {
      acc.m[2] -= kh2l * ((ax * ax + 4.0 * (mx * mx) + bx * bx) * (1.0 / 120.0));
      acc.m[3] -= kh2l * ((ay * ay + 4.0 * (my * my) + by * by) * (1.0 / 120.0));
      acc.m[4] -= kh2l * ((ax * ay + 4.0 * (mx * my) + bx * by) * (1.0 / 120.0));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// (A) and (B) with INDEPENDENT items (no walk along the edges).
//
// Along one cell edge u(t) = A + t d the integrand phi_T = lam_T g/(dg+3) + r_T g/(dg+2) changes face at every
// crossing with a mesh line.  Across the mesh edge m on the line { n.u = h } (n = gradient direction of the
// line's level function, kappa_m as in (B) above) phi jumps by
//      after - before = - kappa_m g(u) [ (n_c.u - h_c)/(dg+3) - h_c/((dg+2)(dg+3)) ] ,
// n_c = +-n the normal in the direction of travel, h_c = +-h.  So
//      ∫_0^1 phi_{T(t)} dt = ∫_0^1 phi_{T0} dt + Σ_{crossings c} ∫_{t_c}^1 (jump of c) dt ,
// T0 = the face at the start of the edge: one closed-form term per edge plus one per crossing, each of degree
// <= 3 in t (Simpson, exact), and the Hessian's ∫ rho ds the same way (jump of rho = - kappa_m (n_c.u - h_c)).
// The crossings are enumerated LINE by line together with part (B): a mesh line crosses the convex cell on two
// edges; the two crossings give two jump terms and the end points of the chord whose unit intervals are (B).
// Every item is O(1) and independent of the others; no DDA state, no midpoint classification.
//
// Consistency (what makes the sum exact whatever the alignment of the cell with the mesh): one sign rule,
// g_v = level function at vertex v, "v is on the low side" <=> g_v < 0, decides (a) the face T0 of an edge
// (floor of the start vertex' grid coordinates = the count of lines with g >= 0), (b) which edges a line crosses
// and (c) hence whether its chord takes part in (B).  An edge lying ON a mesh line is then attributed to the high
// side, and the line's chord (which runs along that very edge when the cell is on the low side) supplies the
// difference.  Vertices closer than rounding are made identical first, so that they get identical signs.
// ------------------------------------------------------------------------------------------------
// one jump term: crossing point (cx, cy), end of the edge (bx, by), tau = 1 - t_c, dB = n_c.B - h_c >= 0,
// hc = h_c, khl = kappa_m * (h_e |e|) of the edge
template <int MODE>
MA_DEV void seg_jump_item(double cx, double cy, double bx, double by, double tau, double dB, double hc, double khl, SegAcc &acc) {
  const double mx = 0.5 * (cx + bx), my = 0.5 * (cy + by);
  const double w6 = khl * tau * (1.0 / 6.0);
  // psi(s) = dB(s)/(dg+3) - hc/((dg+2)(dg+3)) at s = 0, 1/2, 1 of [t_c, 1]
  {  // g = 1 (dg = 0)
    const double p0 = -hc * (1.0 / 6.0), pm = dB * (1.0 / 6.0) + p0, p1 = dB * (1.0 / 3.0) + p0;
    acc.mass -= w6 * (p0 + 4.0 * pm + p1);
  }
  if (MODE == MODE_KANTOROVICH) {  // g = |u|^2 (dg = 2)
    const double p0 = -hc * (1.0 / 20.0), pm = dB * (1.0 / 10.0) + p0, p1 = dB * (1.0 / 5.0) + p0;
    const double qC = cx * cx + cy * cy, qM = mx * mx + my * my, qB = bx * bx + by * by;
    acc.cost -= w6 * (qC * p0 + 4.0 * (qM * pm) + qB * p1);
  } else {
    {  // g = ux, uy (dg = 1)
      const double p0 = -hc * (1.0 / 12.0), pm = dB * (1.0 / 8.0) + p0, p1 = dB * (1.0 / 4.0) + p0;
      acc.m[0] -= w6 * (cx * p0 + 4.0 * (mx * pm) + bx * p1);
      acc.m[1] -= w6 * (cy * p0 + 4.0 * (my * pm) + by * p1);
    }
    if (MODE == MODE_MOMENTS2) {  // g = ux^2, uy^2, ux uy (dg = 2)
      const double p0 = -hc * (1.0 / 20.0), pm = dB * (1.0 / 10.0) + p0, p1 = dB * (1.0 / 5.0) + p0;
      acc.m[2] -= w6 * (cx * cx * p0 + 4.0 * (mx * mx * pm) + bx * bx * p1);
      acc.m[3] -= w6 * (cy * cy * p0 + 4.0 * (my * my * pm) + by * by * p1);
      acc.m[4] -= w6 * (cx * cy * p0 + 4.0 * (mx * my * pm) + bx * by * p1);
    }
  }
}

// The line-major formulation reads the vertex densities from rho_p: the grid PADDED by one layer of replicated
// values on every side (vertex (i, j), -1 <= i <= gn, at rho_p[(i + 1) * (gm + 2) + (j + 1)]).  The density is thereby
// extended continuously (and piecewise linearly) one square beyond the domain, and the sign rule below needs no
// special case on the rim: an edge lying ON the domain boundary is attributed to the (fictitious) square outside
// and the boundary line's chord supplies the difference, exactly as for an interior mesh line.
MA_DEV const double *seg_rv(const Params &p, int i, int j) { return p.rho_p + (size_t)(i + 1) * (p.gm + 2) + (j + 1); }

// Variable diagonals (VD): square (i, j) of the padded grid is split along (i,j)-(i+1,j+1) ("+", bit 0 — the only kind
// ma_set_grid makes) or along (i+1,j)-(i,j+1) ("-", bit 1): what a Delaunay triangulation of the pixel grid gives, every
// square being co-circular.  One bit per square, i in [-1, gn-1], j in [-1, gm-1], fictitious layer "+".
MA_DEV bool seg_diag(const Params &p, int i, int j) {
  const size_t q = (size_t)(i + 1) * (p.gm + 1) + (j + 1);
  return (p.diag[q >> 5] >> (q & 31)) & 1u;
}

// kappa of the unit mesh edge mm of line (fam, k): n.(grad rho_low - grad rho_high), n the direction in which the line's
// level function grows.  Families: 0 x = k, 1 y = k, 2 "+" diagonals fx - fy = k, 3 (VD only) "-" diagonals fx + fy = k;
// a diagonal's unit edge exists only in a square of its own kind (kappa = 0 otherwise: rho is smooth across it).
template <bool VD>
MA_DEV double seg_kappa(const Params &p, int fam, int k, int mm, double inv_dx, double inv_dy, double dcoef) {
  const int gs = p.gm + 2;
  if (fam == 0) {  // vertical edge (k,mm)-(k,mm+1): left = square (k-1,mm), right = square (k,mm)
    const double *rv = seg_rv(p, k, mm);
    const bool dl = VD && seg_diag(p, k - 1, mm), dr = VD && seg_diag(p, k, mm);
    const double left = dl ? (rv[1] - rv[1 - gs]) : (rv[0] - rv[-gs]);
    const double right = dr ? (rv[gs] - rv[0]) : (rv[gs + 1] - rv[1]);
    return (left - right) * inv_dx;
  }
  if (fam == 1) {  // horizontal edge (mm,k)-(mm+1,k): below = square (mm,k-1), above = square (mm,k)
    const double *rv = seg_rv(p, mm, k);
    const bool db = VD && seg_diag(p, mm, k - 1), da = VD && seg_diag(p, mm, k);
    const double below = db ? (rv[gs] - rv[gs - 1]) : (rv[0] - rv[-1]);
    const double above = da ? (rv[1] - rv[0]) : (rv[gs + 1] - rv[gs]);
    return (below - above) * inv_dy;
  }
  if (fam == 2) {  // "+" diagonal of square (mm, mm-k)
    if (VD && seg_diag(p, mm, mm - k)) return 0.0;
    const double *rv = seg_rv(p, mm, mm - k);
    return ((rv[0] + rv[gs + 1]) - (rv[1] + rv[gs])) * dcoef;
  }
  // "-" diagonal of square (mm, k-mm-1)
  if (!seg_diag(p, mm, k - mm - 1)) return 0.0;
  const double *rv = seg_rv(p, mm, k - mm - 1);
  return ((rv[1] + rv[gs]) - (rv[0] + rv[gs + 1])) * dcoef;
}

// E: per-edge accumulator of ∫_0^1 rho(u(t)) dt, element k at E[k * ES] (shared memory column of the thread)
// tagof(k): the tag of edge k (only the Hessian slots at the very end need it, so it need not be staged with the vertices)
template <int MODE, int ES, bool VD = false, class Poly, class TagOf>
MA_DEV unsigned long long cell_integrate_lines(const Params &p, int i, const Poly &P, int n, SegAcc &acc,
                                               double *hslot_row, double *E, TagOf tagof) {
  const double xi = p.xs[i], yi = p.ys[i];
  const double inv_dx = 1.0 / p.gdx, inv_dy = 1.0 / p.gdy;
  const double ox = (xi - p.gx0) * inv_dx, oy = (yi - p.gy0) * inv_dy;  // grid coordinate f = u * inv + o
  const int gs = p.gm + 2;
  // squares -1 .. gn-1 x -1 .. gm-1 (one fictitious layer around the gn-1 x gm-1 real ones)
  const int sx_hi = p.gn - 1, sy_hi = p.gm - 1;
  acc.mass = acc.cost = 0.0;
#pragma unroll
  for (int q = 0; q < 5; ++q) acc.m[q] = 0.0;
  unsigned long long touched = 0ull;
  if (n < 3) return touched;
  // Two clean-ups that make coincidences EXACT, so that the sign rule sees them the same way everywhere:
  //  * vertices closer than rounding become identical (degenerate zero-length edges of co-circular sites);
  //  * a coordinate within 1e-11 grid units of a mesh line is moved onto the value that line has in local
  //    coordinates: all vertices on one mesh line (the two ends of an edge along the domain boundary, of an
  //    edge of a pixel-aligned lattice cell) then share one bit pattern, the edge between them does not
  //    "cross" its own line, and every position interpolated along it is that same value.
  {
    double sc = 0.0;
    for (int k = 0; k < n; ++k) {
      double X = P.X(k), Y = P.Y(k);
      const double fx = X * inv_dx + ox, fy = Y * inv_dy + oy;
      const double rx = rint(fx), ry = rint(fy);
      if (fabs(fx - rx) < 1e-11) { X = (rx - ox) * p.gdx; P.X(k) = X; }
      if (fabs(fy - ry) < 1e-11) { Y = (ry - oy) * p.gdy; P.Y(k) = Y; }
      sc = fmax(sc, fmax(fabs(X), fabs(Y)));
    }
    const double eps2 = (1e-13 * sc) * (1e-13 * sc);
    for (int k = 0; k + 1 < n; ++k) {
      const double ex = P.X(k + 1) - P.X(k), ey = P.Y(k + 1) - P.Y(k);
      if (ex * ex + ey * ey <= eps2) { P.X(k + 1) = P.X(k); P.Y(k + 1) = P.Y(k); }
    }
  }
  // ---------------- per edge: the term of the face at its start ----------------
  double fxmin = 1e300, fxmax = -1e300, fymin = 1e300, fymax = -1e300, fdmin = 1e300, fdmax = -1e300, fsmin = 1e300, fsmax = -1e300;
  {
    double Ax = P.X(0), Ay = P.Y(0);
    for (int k = 0; k < n; ++k) {
      const int kk = (k + 1 == n) ? 0 : k + 1;
      const double Bx = P.X(kk), By = P.Y(kk);
      const double dx = Bx - Ax, dy = By - Ay;
      const double hL = dy * Ax - dx * Ay;  // h_e * |e|
      const double fAx = Ax * inv_dx + ox, fAy = Ay * inv_dy + oy, fAd = fAx - fAy;
      fxmin = fmin(fxmin, fAx); fxmax = fmax(fxmax, fAx);
      fymin = fmin(fymin, fAy); fymax = fmax(fymax, fAy);
      fdmin = fmin(fdmin, fAd); fdmax = fmax(fdmax, fAd);
      const double fAs = fAx + fAy;
      if (VD) { fsmin = fmin(fsmin, fAs); fsmax = fmax(fsmax, fAs); }
      const int si = seg_clampi((int)floor(fAx), -1, sx_hi), sj = seg_clampi((int)floor(fAy), -1, sy_hi);
      const double *rv = seg_rv(p, si, sj);
      double a, b, r;
      if (VD && seg_diag(p, si, sj)) {
        // "-" square: low side of fx + fy = si + sj + 1 is the triangle at vertex (si, sj), high side the one at (si+1, sj+1)
        const bool low = (fAs - (double)(si + sj + 1)) < 0.0;
        const double r00 = rv[0], r10 = rv[gs], r01 = rv[1], r11 = rv[gs + 1];
        a = (low ? (r10 - r00) : (r11 - r01)) * inv_dx;
        b = (low ? (r01 - r00) : (r11 - r10)) * inv_dy;
        r = low ? r00 - a * (((double)si - ox) * p.gdx) - b * (((double)sj - oy) * p.gdy)
                : r11 - a * (((double)(si + 1) - ox) * p.gdx) - b * (((double)(sj + 1) - oy) * p.gdy);
      } else {
        const bool upper = (fAd - (double)(si - sj)) < 0.0;  // low side of the square's diagonal = face 1
        const double r00 = rv[0], r11 = rv[gs + 1], rmid = upper ? rv[1] : rv[gs];
        a = (upper ? (r11 - rmid) : (rmid - r00)) * inv_dx;
        b = (upper ? (rmid - r00) : (r11 - rmid)) * inv_dy;
        r = r00 - a * (((double)si - ox) * p.gdx) - b * (((double)sj - oy) * p.gdy);
      }
      seg_item_cell<MODE>(Ax, Ay, Bx, By, a, b, r, hL, acc);
      if (MODE == MODE_KANTOROVICH) E[k * ES] = a * (0.5 * (Ax + Bx)) + b * (0.5 * (Ay + By)) + r;
      Ax = Bx; Ay = By;
    }
  }
  // ---------------- per mesh line: two jump terms + the chord's unit intervals ----------------
  const double ddiag = sqrt(p.gdx * p.gdx + p.gdy * p.gdy);
  const double ninv = 1.0 / sqrt(inv_dx * inv_dx + inv_dy * inv_dy);  // 1 / |(1/dx, -1/dy)|
  const double dcoef = (inv_dx * inv_dx + inv_dy * inv_dy) * ninv;
  // line k of a family has a vertex on either side  <=>  min < k <= max over the vertices; lines of the padded grid only.
  // (One loop per family with the family a compile-time constant: walking the three families as one list with a
  // run-time family evens out the trip counts across a warp but costs more in selects than it saves — measured.)
  const unsigned nmask = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
#pragma unroll
  for (int fam = 0; fam < (VD ? 4 : 3); ++fam) {
    const double lo_f = fam == 0 ? fxmin : (fam == 1 ? fymin : (fam == 2 ? fdmin : fsmin));
    const double hi_f = fam == 0 ? fxmax : (fam == 1 ? fymax : (fam == 2 ? fdmax : fsmax));
    int k0 = (int)floor(lo_f) + 1, k1 = (int)floor(hi_f);
    if (fam == 0) { k0 = max(k0, 0); k1 = min(k1, sx_hi); }
    else if (fam == 1) { k0 = max(k0, 0); k1 = min(k1, sy_hi); }
    else if (fam == 2) { k0 = max(k0, -sy_hi - 1); k1 = min(k1, sx_hi + 1); }
    else { k0 = max(k0, -1); k1 = min(k1, sx_hi + sy_hi + 1); }
    const double fscale = fam == 0 ? p.gdx : (fam == 1 ? p.gdy : ninv);  // level function -> distance
    for (int k = k0; k <= k1; ++k)
    {
      const double lev = (double)k;
      // level function at vertex v (always this very expression: its sign is THE side of v)
      auto gat = [&](int v) {
        const double Qx = P.X(v) * inv_dx + ox, Qy = P.Y(v) * inv_dy + oy;
        return (fam == 0 ? Qx : (fam == 1 ? Qy : (fam == 2 ? Qx - Qy : Qx + Qy))) - lev;
      };
      // the edges with a sign change: one bit per vertex, then bit tricks (cells of more than 32 vertices, which only
      // the largest capacity class can hold, scan edge by edge)
      int e1 = -1, e2 = -1;
      double g1a = 0, g1b = 1, g2a = 0, g2b = 1;
      if (n <= 32) {
        unsigned neg = 0u;
        for (int v = 0; v < n; ++v) neg |= (gat(v) < 0.0 ? 1u : 0u) << v;
        const unsigned nxt = ((neg >> 1) | ((neg & 1u) << (n - 1))) & nmask;  // bit v = side of vertex v+1
        const unsigned cr = (neg ^ nxt) & nmask;
        if (cr & (cr - 1u)) {  // at least two crossings: the first and the last
#ifdef __CUDA_ARCH__
          e1 = __ffs((int)cr) - 1; e2 = 31 - __clz((int)cr);
#else
          e1 = __builtin_ctz(cr); e2 = 31 - __builtin_clz(cr);
#endif
          g1a = gat(e1); g1b = gat(e1 + 1 == n ? 0 : e1 + 1);
          g2a = gat(e2); g2b = gat(e2 + 1 == n ? 0 : e2 + 1);
        }
      } else {
        double gv = gat(0);
        for (int v = 0; v < n; ++v) {
          const double gw = gat(v + 1 == n ? 0 : v + 1);
          const bool cross = (gv < 0.0) != (gw < 0.0);
          const bool first = cross && e1 < 0, second = cross && e1 >= 0;
          if (second) { e2 = v; g2a = gv; g2b = gw; }
          if (first) { e1 = v; g1a = gv; g1b = gw; }
          gv = gw;
        }
      }
      if (e2 < 0) continue;
      // unit mesh edges of this line in the padded grid: mm in [mlo, mhi]
      int mlo, mhi;
      if (fam == 0) { mlo = -1; mhi = sy_hi; }
      else if (fam == 1) { mlo = -1; mhi = sx_hi; }
      else if (fam == 2) { mlo = max(-1, k - 1); mhi = min(sx_hi, sy_hi + k); }
      else { mlo = max(-1, k - 1 - sy_hi); mhi = min(sx_hi, k); }
      if (mhi < mlo) continue;
      double hcoef;  // h of this line (n = the direction in which the level function grows)
      if (fam == 0) hcoef = (lev - ox) * p.gdx;
      else if (fam == 1) hcoef = (lev - oy) * p.gdy;
      else if (fam == 2) hcoef = (lev - (ox - oy)) * ninv;
      else hcoef = (lev - (ox + oy)) * ninv;
      double send[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int e = c == 0 ? e1 : e2;
        const double ga = c == 0 ? g1a : g2a, gb = c == 0 ? g1b : g2b;
        const int ee = (e + 1 == n) ? 0 : e + 1;
        const double Ax = P.X(e), Ay = P.Y(e), Bx = P.X(ee), By = P.Y(ee);
        const double dx = Bx - Ax, dy = By - Ay;
        const double tc = ga / (ga - gb), tau = 1.0 - tc;
        // (exact at both ends: a crossing AT a vertex must get that vertex' own coordinates, whose floor decided
        // the face of the edge that starts there)
        const double cx = tc <= 0.5 ? Ax + tc * dx : Bx - tau * dx, cy = tc <= 0.5 ? Ay + tc * dy : By - tau * dy;
        // Position of the crossing along the line -> the unit mesh edge that is crossed.  Interpolated between the
        // grid coordinates of the edge's end points, the very values whose signs decide all crossings: when the
        // edge runs along another mesh line within rounding, s then falls on the side of that line the edge
        // is attributed to at t_c (recomputing it from (cx, cy) would round it anywhere).
        const double fa = (fam == 0) ? (Ay * inv_dy + oy) : (Ax * inv_dx + ox);
        const double fb = (fam == 0) ? (By * inv_dy + oy) : (Bx * inv_dx + ox);
        const double s = tc <= 0.5 ? fa + tc * (fb - fa) : fb - tau * (fb - fa);
        send[c] = s;
        int mm = (int)floor(s);
        if (fam == 2) {
          // a diagonal's unit edge is square (mm, mm - k): take it from the coordinate that varies LESS along the
          // edge (along a horizontal boundary edge fy is one exact value while fx passes the integers at the very
          // places where the diagonals are crossed)
          const double ga2 = Ay * inv_dy + oy, gb2 = By * inv_dy + oy;
          if (fabs(gb2 - ga2) < fabs(fb - fa)) mm = (int)floor(tc <= 0.5 ? ga2 + tc * (gb2 - ga2) : gb2 - tau * (gb2 - ga2)) + k;
        }
        if (fam == 3) {  // the same for a "-" diagonal: its unit edge is square (mm, k - mm - 1)
          const double ga2 = Ay * inv_dy + oy, gb2 = By * inv_dy + oy;
          if (fabs(gb2 - ga2) < fabs(fb - fa)) mm = k - 1 - (int)floor(tc <= 0.5 ? ga2 + tc * (gb2 - ga2) : gb2 - tau * (gb2 - ga2));
        }
        mm = seg_clampi(mm, mlo, mhi);
        const double kap = seg_kappa<VD>(p, fam, k, mm, inv_dx, inv_dy, dcoef);
        const double sgn = (ga < 0.0) ? 1.0 : -1.0;  // direction of travel across the line
        const double dB = fabs(gb) * fscale;         // n_c.B - h_c
        const double hL = dy * Ax - dx * Ay;
        seg_jump_item<MODE>(cx, cy, Bx, By, tau, dB, sgn * hcoef, kap * hL, acc);
        if (MODE == MODE_KANTOROVICH) E[e * ES] -= kap * dB * tau * 0.5;
      }
      // (B) the chord between the two crossings
      const double lo = fmin(send[0], send[1]), hi = fmax(send[0], send[1]);
      if (!(hi > lo)) continue;
      const int m0 = max((int)floor(lo), mlo), m1 = min((int)floor(hi), mhi);
      const double h2 = hcoef * hcoef;
      for (int mm = m0; mm <= m1; ++mm) {
        const double s0 = fmax(lo, (double)mm), s1 = fmin(hi, (double)(mm + 1));
        if (!(s1 > s0)) continue;
        const double kappa = seg_kappa<VD>(p, fam, k, mm, inv_dx, inv_dy, dcoef);
        if (VD && kappa == 0.0) continue;
        double ax, ay, bx, by, len;
        if (fam == 0) {
          ax = bx = hcoef;
          ay = (s0 - oy) * p.gdy; by = (s1 - oy) * p.gdy;
          len = (s1 - s0) * p.gdy;
        } else if (fam == 1) {
          ay = by = hcoef;
          ax = (s0 - ox) * p.gdx; bx = (s1 - ox) * p.gdx;
          len = (s1 - s0) * p.gdx;
        } else if (fam == 2) {
          ax = (s0 - ox) * p.gdx; bx = (s1 - ox) * p.gdx;
          ay = ((s0 - lev) - oy) * p.gdy; by = ((s1 - lev) - oy) * p.gdy;
          len = (s1 - s0) * ddiag;
        } else {
          ax = (s0 - ox) * p.gdx; bx = (s1 - ox) * p.gdx;
          ay = ((lev - s0) - oy) * p.gdy; by = ((lev - s1) - oy) * p.gdy;
          len = (s1 - s0) * ddiag;
        }
        seg_item_mesh<MODE>(ax, ay, bx, by, kappa * h2 * len, acc);
      }
    }
  }
  // ---------------- Hessian slots ----------------
  if (MODE == MODE_KANTOROVICH) {
    int slot = 0;
    for (int k = 0; k < n; ++k) {
      const int tag = tagof(k);
      if (tag < 0) continue;
      const int kk = (k + 1 == n) ? 0 : k + 1;
      const double dx = P.X(kk) - P.X(k), dy = P.Y(kk) - P.Y(k);
      const double len = sqrt(dx * dx + dy * dy);
      if (len > 0.0 && slot < p.kmax) {
        const double Dx = p.xs[tag] - xi, Dy = p.ys[tag] - yi;
        hslot_row[slot] = E[k * ES] * len * (0.5 / sqrt(Dx * Dx + Dy * Dy));
        touched |= 1ull << slot;
      }
      ++slot;
    }
  }
  return touched;
}

}  // namespace ma

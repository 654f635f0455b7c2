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
// On a grid the sub-segments are found by walking each cell edge through the three families of mesh
// lines (x = const, y = const, diagonals) and each mesh line through its chord of the convex cell;
// no orientation predicate decides anything, so the result depends continuously on the input and
// there is nothing to make robust.  One THREAD handles one cell (the polygon stays in shared memory
// where K2 built it).  SURVEY.md §7.2 "two regimes", DESIGN.md §3.
#pragma once
#include "ma_block.cuh"

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

// next crossing of u(t) = u0 + t*sl with an integer level: initial state of one line family
MA_DEV void seg_family_init(double u0, double sl, int &kn, int &stp, double &inv, double &tn) {
  if (sl > 0.0) { kn = (int)floor(u0) + 1; stp = 1; }
  else if (sl < 0.0) { kn = (int)ceil(u0) - 1; stp = -1; }
  else { kn = 0; stp = 0; inv = 0.0; tn = 1.0 / 0.0; return; }
  inv = 1.0 / sl;
  tn = ((double)kn - u0) * inv;
}

// Integrals of one cell (polygon P of n vertices, local coordinates, vertex k = start of edge k whose
// supporting line is named by tag k: site index >= 0, or < 0 for a side of the mesh bounding box).
// MODE_KANTOROVICH additionally writes, for every Laguerre edge in polygon order (= the order of
// cell_emit's neighbour list), hslot = ∫_edge rho ds / (2 |y_i - y_j|)  and returns the touched mask.
template <int MODE, class Poly>
MA_DEV unsigned long long cell_integrate_grid(const Params &p, int i, const Poly &P, int n, SegAcc &acc,
                                              double *hslot_row) {
  const double xi = p.xs[i], yi = p.ys[i];
  const double inv_dx = 1.0 / p.gdx, inv_dy = 1.0 / p.gdy;
  const double ox = (xi - p.gx0) * inv_dx, oy = (yi - p.gy0) * inv_dy;  // grid coordinate f = u * inv + o
  const int gn2 = p.gn - 2, gm2 = p.gm - 2;
  acc.mass = acc.cost = 0.0;
#pragma unroll
  for (int q = 0; q < 5; ++q) acc.m[q] = 0.0;
  unsigned long long touched = 0ull;
  if (n < 3) return touched;

  // ---------------- (A) the cell's own edges ----------------
  int slot = 0;
  double fxmin = 1e300, fxmax = -1e300, fymin = 1e300, fymax = -1e300, fdmin = 1e300, fdmax = -1e300;
  double Ax = P.X(0), Ay = P.Y(0);
  for (int k = 0; k < n; ++k) {
    const int kk = (k + 1 == n) ? 0 : k + 1;
    const double Bx = P.X(kk), By = P.Y(kk);
    const int tag = P.T(k);
    const double dx = Bx - Ax, dy = By - Ay;
    const double hL = dy * Ax - dx * Ay;  // h_e * |e|  (outward normal (dy,-dx)/|e| of a CCW polygon)
    const double fAx = Ax * inv_dx + ox, fAy = Ay * inv_dy + oy;
    const double sx = dx * inv_dx, sy = dy * inv_dy;
    fxmin = fmin(fxmin, fAx); fxmax = fmax(fxmax, fAx);
    fymin = fmin(fymin, fAy); fymax = fmax(fymax, fAy);
    fdmin = fmin(fdmin, fAx - fAy); fdmax = fmax(fdmax, fAx - fAy);
    int kx, ky, kd, stx, sty, std_;
    double ivx, ivy, ivd, tx, ty, td;
    const double u0d = fAx - fAy, sd = sx - sy;
    seg_family_init(fAx, sx, kx, stx, ivx, tx);
    seg_family_init(fAy, sy, ky, sty, ivy, ty);
    seg_family_init(u0d, sd, kd, std_, ivd, td);
    double tc = 0.0, cx = Ax, cy = Ay, E = 0.0;
    const int cap = 3 * (p.gn + p.gm) + 16;  // more crossings than lines cannot happen
    for (int it = 0; it < cap; ++it) {
      const double t1 = fmin(fmin(tx, ty), fmin(td, 1.0));
      if (t1 > tc) {
        const double ex = Ax + t1 * dx, ey = Ay + t1 * dy;
        const double tm = 0.5 * (tc + t1);
        const double fmx = fAx + tm * sx, fmy = fAy + tm * sy;
        const int si = seg_clampi((int)floor(fmx), 0, gn2), sj = seg_clampi((int)floor(fmy), 0, gm2);
        const bool upper = (fmx - (double)si) < (fmy - (double)sj);  // face 1 of the square (above the diagonal)
        // the face's plane from its three vertex densities: a, b and the value r at y_i
        const double *rv = p.rho_v + (size_t)si * p.gm + sj;
        const double r00 = rv[0], r11 = rv[p.gm + 1], rmid = upper ? rv[1] : rv[p.gm];  // r01 or r10
        const double a = (upper ? (r11 - rmid) : (rmid - r00)) * inv_dx;
        const double b = (upper ? (rmid - r00) : (r11 - rmid)) * inv_dy;
        const double r = r00 - a * (((double)si - ox) * p.gdx) - b * (((double)sj - oy) * p.gdy);
        const double dt = t1 - tc;
        seg_item_cell<MODE>(cx, cy, ex, ey, a, b, r, hL * dt, acc);
        if (MODE == MODE_KANTOROVICH) E += dt * (a * (0.5 * (cx + ex)) + b * (0.5 * (cy + ey)) + r);
        tc = t1; cx = ex; cy = ey;
      }
      if (t1 >= 1.0) break;
      if (tx <= t1) { kx += stx; tx = ((double)kx - fAx) * ivx; }
      if (ty <= t1) { ky += sty; ty = ((double)ky - fAy) * ivy; }
      if (td <= t1) { kd += std_; td = ((double)kd - u0d) * ivd; }
    }
    if (tag >= 0) {
      if (MODE == MODE_KANTOROVICH) {
        const double len = sqrt(dx * dx + dy * dy);
        if (len > 0.0 && slot < p.kmax) {
          const double Dx = p.xs[tag] - xi, Dy = p.ys[tag] - yi;
          hslot_row[slot] = E * len * (0.5 / sqrt(Dx * Dx + Dy * Dy));
          touched |= 1ull << slot;
        }
      }
      ++slot;
    }
    Ax = Bx; Ay = By;
  }

  // ---------------- (B) interior mesh edges inside the cell ----------------
  // chord of the line { fam(u) = level } in the convex polygon, as an interval [lo, hi] of the
  // line's own parameter (fy on x-lines, fx on y-lines and diagonals)
  const double ddiag = sqrt(p.gdx * p.gdx + p.gdy * p.gdy);
  const double ninv = 1.0 / sqrt(inv_dx * inv_dx + inv_dy * inv_dy);  // 1 / |(1/dx, -1/dy)|
  for (int fam = 0; fam < 3; ++fam) {
    const double lo_f = fam == 0 ? fxmin : (fam == 1 ? fymin : fdmin);
    const double hi_f = fam == 0 ? fxmax : (fam == 1 ? fymax : fdmax);
    int k0 = (int)floor(lo_f) + 1, k1 = (int)ceil(hi_f) - 1;
    if (fam == 0) { k0 = max(k0, 1); k1 = min(k1, gn2); }
    else if (fam == 1) { k0 = max(k0, 1); k1 = min(k1, gm2); }
    else { k0 = max(k0, -gm2); k1 = min(k1, gn2); }
    for (int k = k0; k <= k1; ++k) {
      double lo = 1e300, hi = -1e300;
      {
        // a convex polygon crosses the line on (at most) two edges: find them without doing any
        // arithmetic inside the divergent branch, then intersect both after the loop, all lanes together
        const double lev = (double)k;
        double Px = P.X(0) * inv_dx + ox, Py = P.Y(0) * inv_dy + oy;
        double gv = (fam == 0 ? Px : (fam == 1 ? Py : Px - Py)) - lev;
        double sv = fam == 0 ? Py : Px;
        double g1a = 0, g1b = 1, s1a = 0, s1b = 0, g2a = 0, g2b = 1, s2a = 0, s2b = 0;
        int ncross = 0;
        for (int v = 0; v < n; ++v) {
          const int vv = (v + 1 == n) ? 0 : v + 1;
          const double Qx = P.X(vv) * inv_dx + ox, Qy = P.Y(vv) * inv_dy + oy;
          const double gw = (fam == 0 ? Qx : (fam == 1 ? Qy : Qx - Qy)) - lev;
          const double sw = fam == 0 ? Qy : Qx;
          const bool cross = (gv < 0.0) != (gw < 0.0);
          const bool first = cross && ncross == 0, second = cross && ncross > 0;
          g1a = first ? gv : g1a; g1b = first ? gw : g1b; s1a = first ? sv : s1a; s1b = first ? sw : s1b;
          g2a = second ? gv : g2a; g2b = second ? gw : g2b; s2a = second ? sv : s2a; s2b = second ? sw : s2b;
          ncross += cross ? 1 : 0;
          gv = gw; sv = sw;
        }
        if (ncross >= 2) {
          const double c1 = s1a + (g1a / (g1a - g1b)) * (s1b - s1a);
          const double c2 = s2a + (g2a / (g2a - g2b)) * (s2b - s2a);
          lo = fmin(c1, c2); hi = fmax(c1, c2);
        }
      }
      if (!(hi > lo)) continue;
      int m0 = (int)floor(lo), m1 = (int)floor(hi);
      double hcoef;  // h of this line (constant along it)
      if (fam == 0) { m0 = max(m0, 0); m1 = min(m1, gm2); hcoef = ((double)k - ox) * p.gdx; }
      else if (fam == 1) { m0 = max(m0, 0); m1 = min(m1, gn2); hcoef = ((double)k - oy) * p.gdy; }
      else {
        m0 = max(m0, max(0, k)); m1 = min(m1, min(gn2, gm2 + k));
        // n = (1/dx, -1/dy) * ninv; any point of the line has fx - fy = k:  n.u = ((fx-ox) - (fy-oy)) * ninv
        hcoef = ((double)k - (ox - oy)) * ninv;
      }
      const double h2 = hcoef * hcoef;
      for (int mm = m0; mm <= m1; ++mm) {
        const double s0 = fmax(lo, (double)mm), s1 = fmin(hi, (double)(mm + 1));
        if (!(s1 > s0)) continue;
        double ax, ay, bx, by, kappa, len;
        if (fam == 0) {  // vertical edge (k,mm)-(k,mm+1): left = face 0 of square (k-1,mm), right = face 1 of (k,mm)
          const double *rv = p.rho_v + (size_t)k * p.gm + mm;  // a_left - a_right
          kappa = ((rv[0] - rv[-p.gm]) - (rv[p.gm + 1] - rv[1])) * inv_dx;
          ax = bx = hcoef;
          ay = (s0 - oy) * p.gdy; by = (s1 - oy) * p.gdy;
          len = (s1 - s0) * p.gdy;
        } else if (fam == 1) {  // horizontal edge (mm,k)-(mm+1,k): below = face 1 of square (mm,k-1), above = face 0 of (mm,k)
          const double *rv = p.rho_v + (size_t)mm * p.gm + k;  // b_below - b_above
          kappa = ((rv[0] - rv[-1]) - (rv[p.gm + 1] - rv[p.gm])) * inv_dy;
          ay = by = hcoef;
          ax = (s0 - ox) * p.gdx; bx = (s1 - ox) * p.gdx;
          len = (s1 - s0) * p.gdx;
        } else {  // diagonal of square (mm, mm-k): n points to face 0 (lower right), so T1 = face 1
          const double *rv = p.rho_v + (size_t)mm * p.gm + (mm - k);
          // n.(grad rho_1 - grad rho_0) = (r00 + r11 - r01 - r10) (1/dx^2 + 1/dy^2) / |(1/dx, -1/dy)|
          kappa = ((rv[0] + rv[p.gm + 1]) - (rv[1] + rv[p.gm])) * (inv_dx * inv_dx + inv_dy * inv_dy) * ninv;
          ax = (s0 - ox) * p.gdx; bx = (s1 - ox) * p.gdx;
          ay = ((s0 - (double)k) - oy) * p.gdy; by = ((s1 - (double)k) - oy) * p.gdy;
          len = (s1 - s0) * ddiag;
        }
        seg_item_mesh<MODE>(ax, ay, bx, by, kappa * h2 * len, acc);
      }
    }
  }
  return touched;
}

}  // namespace ma

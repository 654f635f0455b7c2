// ma_cell.cuh — per-thread / per-lane work of the evaluation kernels, written as __host__ __device__
// functions so that tests/emu can run the very same code serially on the CPU.
//
//   CellSearch / cell_build   (K2)  replaces CGAL Regular_triangulation_2's neighbour circulator
//                      (kantorovich.hpp:65-72, vti.hpp:265): builds the power cell of Dirac i inside the
//                      mesh bounding box by clipping against candidate sites found by a ring walk over
//                      the leaf bins (security-radius certificate) and, when the weights
//                      have a gradient, a quadtree walk pruned by per-node supporting planes.  A
//                      resumable per-lane state machine (search_step / clip) driven by warp votes.
//   lane_pieces  (K3, general meshes)  replaces the overlay traversal + callbacks (vti.hpp:250-305,
//                      kantorovich.hpp:87-136, lloyd.hpp:48-68,91-122): one lane clips candidate
//                      triangles against the cell's half-plane table and integrates the pieces.
//                      (Grid meshes: ma_seg.cuh.)
#pragma once
#include "ma_geom.cuh"

namespace ma {

enum { MESH_NONE = 0, MESH_GENERAL = 1, MESH_GRID = 2 };
enum { MODE_KANTOROVICH = 0, MODE_MOMENTS1 = 1, MODE_MOMENTS2 = 2, MODE_PIECES_COUNT = 3, MODE_PIECES_FILL = 4 };
enum { FLAG_CELL_OVERFLOW = 1, FLAG_PIECE_OVERFLOW = 2, FLAG_KMAX_OVERFLOW = 4, FLAG_STACK_OVERFLOW = 8 };
enum { CNT_PIECES = 0, CNT_VERTS = 1, CNT_NEWV = 2, CNT_LEDGES = 3, CNT_SUMK = 4, CNT_SUMKNP = 5, CNT_FALLBACK = 6,
       CNT_CAND = 7, CNT_N = 8 };

struct Params {
  // Diracs, Morton-bin order; this context evaluates the cells [cell_lo, cell_hi) (its Morton tile)
  int N;
  int cell_lo, cell_hi;
  const double *xs, *ys, *ws;
  // quadtree of bins over the Diracs: level l has 4^l nodes, node (l, code) covers a square of side
  // ph * 2^(L-l); wmax holds all levels, level l at offset (4^l - 1) / 3.
  int L;
  double px0, py0, ph;
  const int *bin_start;
  const int *bin_rm;  // row-major copy for the ring walk: bin (cx, cy) holds the sites [bin_rm[2q], bin_rm[2q+1]), q = cy * 2^L + cx
  const double *wmax;
  // the sites a second time in ROW-MAJOR bin order (ma_block.cuh): the bins [a, b] of one bin row are the
  // contiguous run [rm_start[row * bG + a], rm_start[row * bG + b + 1]) of (xr, yr, wr); rm2s maps a position
  // of that order to the index in the Morton order (the index cells and tags use)
  // (that "block grid" has its own resolution bG x bG over the same square as the quadtree's leaf bins, about one
  // Dirac per bin whatever N is; bins of side bph = 1 / binv)
  const double *xr, *yr, *wr;
  const int *rm2s, *rm_start;
  int bG;
  double bph, binv;
  const double *wstat;                // {sum, sum of squares, min, max} of the weights of this evaluation
  const double *nodeG;                // 2 per node: least-squares weight gradient (ma_geom.cuh)
  const unsigned long long *nodeA;    // per node: dkey(alpha_B)
  int refill_at;                      // k_cells_persist refills when this many lanes have finished
  int rmax;                           // K2 walks rings 0..rmax of leaf bins before it turns to the quadtree
  int rtree;                          // ... and at least rings 0..rtree even when the ring certificate cannot work
  int clip_a, clip_b;                 // K2 clips when n_hold * clip_a >= n_search * clip_b (warp vote policy)
  const int *abort_flag;              // != 0: another cell is already known to be empty, stop (line search)
  int abort_on_empty;
  // mesh bounding box
  double bb[4];
  // source mesh
  int mesh_kind;
  int nF;
  const double *abc;
  const double *rho_v;  // grid mesh: density at vertex (i, j) = rho_v[i * gm + j] (L2-resident: 8 B/vertex vs 24 B/face)
  const unsigned *diag; // grid meshes with per-square diagonals: one bit per padded square (ma_seg.cuh), else null
  const double *rho_p;  // the same padded by one replicated layer: vertex (i, j), -1 <= i <= gn, at [(i+1)*(gm+2) + j+1]
  // grid mesh
  int gn, gm;
  double gx0, gy0, gdx, gdy;
  // general mesh + its face bins
  const double *vx, *vy;
  const int *tri;
  int tg;  // bins per side
  double tinvx, tinvy;
  const int *tbin_ptr, *tbin_face;
  // K2 outputs
  int kmax;
  int *nbr;         // N * kmax, -1 padded, CCW
  int *nbr_cnt;     // N, -1 = empty cell
  double *cell_bb;  // N * 4 (absolute xmin, ymin, xmax, ymax)
  // cell polygons for the boundary-segment kernel (grid meshes): vertex k of cell i at [k * N + i],
  // local coordinates, tag = site index or < 0 for a side of the mesh box; poly_n[i] vertices
  double *poly_x, *poly_y;
  int *poly_t, *poly_n;
  // warm path (ma_warm.cuh): cyclic edge tags of cell i at ring[16 i ..], ring_n[i] of them; cstate[i] = WARM_*
  int *ring, *ring_n, *cstate;
  // K3 outputs
  double *mass;     // N
  double *fcell;    // N: m_i w_i - cost_i
  double *hslot;    // N * kmax: sum over pieces of (edge integral / (2 |y_i - y_j|)) per neighbour slot
  unsigned long long *touched;  // N: bit s set iff slot s received an edge
  int *rowcnt;      // N
  double *mom;      // N * 6 (moments modes)
  // pieces dump
  int *pc_count;    // N * 2 (pieces, vertices) per cell              (MODE_PIECES_COUNT)
  const int *pc_off;  // N * 2 exclusive offsets                      (MODE_PIECES_FILL)
  int *pc_cell, *pc_face, *pc_ptr, *pc_tag;
  double *pc_xy;
  // instrumentation
  unsigned long long *counters;  // CNT_N, only when stats != 0
  int stats;
  int *flags;
  double filter_tol;  // relative threshold below which a sign goes to the double-double fallback
};

// ------------------------------------------------------------------------------------------------
// K2: power cell of Dirac i in the mesh box, as a resumable per-lane state machine.
//
// The search is written as
//     for (;;) { SEARCH, one small step at a time, until every lane holds a cutting site; CLIP; }
// with warp votes / syncs between the steps: the lanes of a warp find their cutting sites at
// different moments, and the compiler does not reconverge them on its own, so each step is
// straight-line code (no break / continue / early return) and the whole warp moves from step to
// step together.  (On the CPU emulation the votes are the identity.)
//
// SEARCH, phase 0: square rings of leaf bins around the Dirac's own bin.  Every lane walks the same
// ring pattern (consecutive cells share bins), nearest bins first, so the polygon is tight after a
// few dozen sites.  After ring r every site inside the (2r+1)^2 block has been seen, hence every
// site within rho = distance(y_i, block boundary); the search is over when no site at distance
// >= rho can cut, whatever its weight (global maximum weight, SURVEY §7.2 security radius).  If
// that certificate cannot work (weights with a gradient), phase 1 takes over.
//
// SEARCH, phase 1: expanding-radius walk over the quadtree.  A plain nearest-first DFS degenerates
// for a Dirac next to a high-level quadrant boundary: until the polygon is cut on every side its
// security radius is the whole box, so nothing is pruned and the nearest quadrant is searched
// exhaustively.  Instead the tree is walked in passes with a distance cap that doubles: a pass
// handles exactly the sites with prev < |y_j - y_i| <= cap, and the search ends with one uncapped
// pass (pruned by the security tests alone) once the polygon fits in the disk of radius cap/2 or a
// whole pass went by without a cut.  Nodes inside the phase-0 block are skipped (all seen).
// Security tests for a node B, both conservative ("no site of B can take any vertex p of the
// polygon from i", i.e. pow_j(p) >= pow_i(p) for every j in B and every vertex p):
//   (a) disk around y_i with the node's maximum weight;
//   (b) the node's supporting plane (ma_geom.cuh) against every vertex.
// (b) is what keeps the search local once the weights have a gradient: the cell then lies far from
// its own Dirac and (a), which measures from y_i, would keep a disk of radius ~|grad w| alive.
//
// CellSearch holds one lane's state.  A lane is `searching()` (needs search steps), `holding()` (has
// found a cutting site and waits for the clip) or `done()`.  Drivers: cell_build() below (one cell
// per lane, used by k_cells and the CPU emulation) and k_cells_persist (lanes fetch new cells as
// they finish, ma_kernels.cuh).
// ------------------------------------------------------------------------------------------------
constexpr int CELL_STACK = 52;
// host-only work counters (tests/emu): [0] sites looked at, [1] sign loops, [2] clips, [3] tree nodes
#if !defined(__CUDA_ARCH__) && defined(MA_EMU_COUNTERS)
#define MA_COUNT(k) (++ma_emu_counters[k])
extern long long ma_emu_counters[8];
#else
#define MA_COUNT(k) ((void)0)
#endif
template <class Poly> struct CellSearch {
  int i, n, status, phase;   // phase: 0 rings, 1 tree, 2 finished
  double xi, yi, wi, bx0, by0, bx1, by1, R2;
  double mx, my, rc2;        // a disk (centre m, squared radius rc2, local coordinates) containing the polygon
  bool use_m;                // weights vary enough to displace cells from their Diracs: prune with that disk
  int bx, by;                // the Dirac's own leaf bin
  double dw_glob;            // w_i - max weight (<= 0)
  int r, q, nq;              // ring r: next position q of nq
  int j, jend;               // remaining sites of the current leaf bin
  double lo2, hi2;           // only sites with lo2 < |y_j - y_i|^2 <= hi2 are looked at
  int rdone;                 // rings 0..rdone are complete
  int sp, pass;              // tree walk: stack pointer (the stack itself is the caller's, CELL_STACK entries)
  double cap, prev2, cap2;
  bool last, cut_in_pass;
  int jc;                    // the cutting site found by the search (-1: none)
  double cDx, cDy, cc;
  unsigned long long cin;

  MA_DEV bool searching() const { return jc < 0 && phase < 2; }
  MA_DEV bool holding() const { return jc >= 0; }
  MA_DEV bool done() const { return jc < 0 && phase >= 2; }

  MA_DEV void lineof(const Params &p, int tag, double &nx, double &ny, double &cl) const {
    if (tag >= 0) {
      double Dx = p.xs[tag] - xi, Dy = p.ys[tag] - yi;
      nx = Dx; ny = Dy;
      cl = 0.5 * (Dx * Dx + Dy * Dy + (wi - p.ws[tag]));
    } else if (tag == -1) { nx = 0; ny = 1; cl = by0; }
    else if (tag == -2) { nx = 1; ny = 0; cl = bx1; }
    else if (tag == -3) { nx = 0; ny = 1; cl = by1; }
    else { nx = 1; ny = 0; cl = bx0; }
  }

  // the same line with double-double coefficients from the ORIGINAL inputs (exact stage of the sign filter)
  MA_DEV ddline ddlineof(const Params &p, int tag) const {
    if (tag >= 0) return dd_bisector(xi, yi, wi, p.xs[tag], p.ys[tag], p.ws[tag]);
    ddline l;
    const bool horiz = (tag == -1 || tag == -3);
    l.nx = dd_from(horiz ? 0.0 : 1.0);
    l.ny = dd_from(horiz ? 1.0 : 0.0);
    l.cl = tag == -1 ? dd_sub_dd(p.bb[1], yi) : (tag == -2 ? dd_sub_dd(p.bb[2], xi) : (tag == -3 ? dd_sub_dd(p.bb[3], yi) : dd_sub_dd(p.bb[0], xi)));
    return l;
  }
  // Exact stage (the role of CGAL::Filtered_predicate's second stage, predicates.hpp:159-167): is vertex k of the
  // polygon — the intersection of the lines supporting edges k-1 and k — strictly on i's side of the bisector
  // with site jj?  Side2 of predicates.hpp:101-115 (Side1 for a corner of the box); ties are outside (:86, T3).
  MA_DEV bool exact_inside(const Params &p, const Poly &P, int k, int jj) const {
    const int ta = P.T(k == 0 ? n - 1 : k - 1), tb = P.T(k);
    return dd_side(ddlineof(p, ta), ddlineof(p, tb), dd_bisector(xi, yi, wi, p.xs[jj], p.ys[jj], p.ws[jj])) > 0;
  }

  // Bit k set <=> vertex k is strictly on i's side of the bisector { u.D = c } with site jj (pow_i < pow_j there).
  // The sign of c - u.D is taken in fp64 behind a forward-error filter; what the filter cannot decide is
  // re-evaluated in double-double from the original (y, w) (predicates.hpp:101-115,139-168).  The filter costs one
  // |.|-min per vertex: it compares the smallest |value| with a bound on the rounding error that holds for every
  // vertex, |u.D| <= sqrt(R2 dd2), and only then looks at the vertices one by one.
  // (Measured and dropped, profiles/r03a: walking the SLOTS of a packed polygon instead of its vertices — no nibble per
  // vertex, holes zeroed by the clip, the mask brought into vertex order only on a cut — made K2 7 % slower at c3, and
  // leaving the radius to the clips 20 % slower; this loop is 40 % of the block kernel's instructions as it stands.)
  MA_DEV unsigned long long sign_mask(const Params &p, const Poly &P, int jj, double Dx, double Dy, double c, double dd2,
                                      double dw, double &r2_seen) const {
    unsigned long long in = 0ull;
    double amin = 1.0 / 0.0, r2 = 0.0;
    unsigned long long o = P.ord;  // PACKED: the slot of vertex k is the k-th nibble, peeled off one by one
    for (int k = 0; k < n; ++k) {
      const int sl = Poly::PACK ? (int)(o & 15ull) : k;
      o >>= 4;
      const double X = P.SX(sl), Y = P.SY(sl);
      const double val = c - (X * Dx + Y * Dy);
      if (val > 0.0) in |= 1ull << k;
      amin = fmin(amin, fabs(val));
      r2 = fmax(r2, X * X + Y * Y);
    }
    r2_seen = r2;  // the polygon's radius about y_i, for free (the block kernel keeps R2 current with it)
    const double cmag = 0.5 * (dd2 + fabs(dw));
    // (|c| + |ux Dx| + |uy Dy|)^2 <= 2 cmag^2 + 4 R2 dd2
    if (amin * amin <= (p.filter_tol * p.filter_tol) * (2.0 * cmag * cmag + 4.0 * R2 * dd2)) {
      bool any = false;
      for (int k = 0; k < n; ++k) {
        const double tx = P.X(k) * Dx, ty = P.Y(k) * Dy;
        if (fabs(c - (tx + ty)) <= p.filter_tol * (cmag + fabs(tx) + fabs(ty))) {
          any = true;
          if (exact_inside(p, P, k, jj)) in |= 1ull << k;
          else in &= ~(1ull << k);
        }
      }
      if (any) {
#ifdef __CUDA_ARCH__
        atomicAdd(p.flags + 2, 1);
#else
        p.flags[2] += 1;
#endif
      }
    }
    return in;
  }

  MA_DEV void init(const Params &p, int cell, Poly &P) {
    i = cell;
    xi = p.xs[i]; yi = p.ys[i]; wi = p.ws[i];
    bx0 = p.bb[0] - xi; by0 = p.bb[1] - yi; bx1 = p.bb[2] - xi; by1 = p.bb[3] - yi;
    if (Poly::PACK) { P.ord = 0x3210ull; P.used = 0xfu; }
    P.X(0) = bx0; P.Y(0) = by0; P.T(0) = -1;  // bottom
    P.X(1) = bx1; P.Y(1) = by0; P.T(1) = -2;  // right
    P.X(2) = bx1; P.Y(2) = by1; P.T(2) = -3;  // top
    P.X(3) = bx0; P.Y(3) = by1; P.T(3) = -4;  // left
    n = 4;
    R2 = fmax(bx0 * bx0, bx1 * bx1) + fmax(by0 * by0, by1 * by1);
    mx = 0.5 * (bx0 + bx1); my = 0.5 * (by0 + by1);
    rc2 = 0.25 * ((bx1 - bx0) * (bx1 - bx0) + (by1 - by0) * (by1 - by0));
    // a cell sits about |grad w| / 2 away from its Dirac: with a weight range below (ph/2)^2 no cell
    // leaves the neighbourhood of its own bin and the disk around y_i (R2) is the tighter bound
    use_m = (p.wstat[3] - p.wstat[2]) > 0.25 * p.ph * p.ph;
    if (!use_m) { mx = my = 0.0; rc2 = R2; }
    const int G = 1 << p.L;
    const double pinv = 1.0 / p.ph;
    bx = min(max((int)((xi - p.px0) * pinv), 0), G - 1);
    by = min(max((int)((yi - p.py0) * pinv), 0), G - 1);
    dw_glob = wi - p.wstat[3];  // (the global maximum weight; the pyramid p.wmax may still be under construction, it is read by the tree walk only)
    status = 0;
    phase = 0;
    if (p.abort_on_empty && *(volatile const int *)p.abort_flag) { n = 0; phase = 2; }
    r = 0; q = 0; nq = 1;
    j = 0; jend = 0;
    lo2 = -1.0; hi2 = 1.0 / 0.0;
    rdone = -1;
    sp = 0; pass = 0;
    cap = 0.0; prev2 = -1.0; cap2 = 1.0 / 0.0;
    last = false; cut_in_pass = true;
    jc = -1; cDx = cDy = cc = 0.0; cin = 0ull;
  }

  // one search step (requires searching()); straight-line code
  MA_DEV void search_step(const Params &p, const Poly &P, unsigned *stk) {
    const double NEG_INF = -1.0 / 0.0, POS_INF = 1.0 / 0.0;
    const int G = 1 << p.L;
    const int RMAX = p.rmax;
    if (phase == 0 && j >= jend) {
      // ring walk: move on to the next non-empty bin (a few cheap iterations), closing a ring when its last
      // bin is behind us, so that the step below handles a site in (almost) every call
      while (phase == 0 && j >= jend) {
        if (q == nq) {  // ring r is complete: every site inside the (2r+1)^2 block has been seen
          rdone = r;
          // all bins covered?  (tiny grids)
          const bool all = bx - r <= 0 && bx + r >= G - 1 && by - r <= 0 && by + r >= G - 1;
          // distance from y_i to the block boundary (exact for interior bins, a lower bound at the grid's edge)
          const double ex = xi - (p.px0 + (double)bx * p.ph), ey = yi - (p.py0 + (double)by * p.ph);
          const double emin = fmin(fmax(fmin(fmin(ex, p.ph - ex), fmin(ey, p.ph - ey)), 0.0), p.ph);  // to the nearest side of the own bin
          const double rho = (double)r * p.ph + emin, rho2 = rho * rho, rn = rho + p.ph;
          if (all || cannot_cut(rho2, dw_glob, R2)) {
            phase = 2;  // nothing farther than rho can cut
          } else if (r == RMAX || (r >= p.rtree && rn * rn + dw_glob <= 0.0)) {
            // (one more ring could not certify anything if even a tiny polygon fails; rings 0..rtree are always
            // walked: the true neighbours are the nearby Diracs whatever the weights do, and the rings
            // are the cheap way to find them — the quadtree walk prunes badly while the polygon is still big)
            phase = 1;  // first tree pass: everything within rho is done
            prev2 = rho2;
            cap = 2.0 * fmax(rho, p.ph);
            cut_in_pass = true;
            pass = 0;
            sp = -1;  // "begin a pass"
          } else {
            ++r; q = 0; nq = 8 * r;
          }
        } else {
          int ox, oy;  // position q of ring r, nearest bins first (table)
          ring_offset(r, q, ox, oy);
          ++q;
          const int cx = bx + ox, cy = by + oy;
          if (cx >= 0 && cx < G && cy >= 0 && cy < G) {
            const int2 se = reinterpret_cast<const int2 *>(p.bin_rm)[(size_t)cy * G + cx];  // no Morton encoding, one load
            j = se.x; jend = se.y;
          }
        }
      }
    }
    if (j < jend) {
      // ---- step kind 1: the next site of the current bin
      const int jj = j++;
      MA_COUNT(0);
      const double Dx = p.xs[jj] - xi, Dy = p.ys[jj] - yi, wj = p.ws[jj];
      const double dd2 = Dx * Dx + Dy * Dy;
      if (jj != i && dd2 > lo2 && dd2 <= hi2) {  // else: itself / an earlier pass's / a later pass's
        if (dd2 == 0.0) {  // coincident sites: the heavier (then the earlier) one keeps the cell
          if (wj > wi || (wj == wi && jj < i)) { n = 0; phase = 2; }
        } else {
          // the bisector { u.D = c } misses the disk (m, rc) that contains the polygon  =>  it cannot cut.
          // (Measured from the polygon, not from y_i: once the weights have a gradient the cell lies far
          // from its Dirac and a disk around y_i would reject nothing.)
          // (use_m = false: m = 0 and rc2 = R2, the disk around y_i)
          const double dw = wi - wj;
          const double c = 0.5 * (dd2 + dw);
          const double e = c - (mx * Dx + my * Dy);
          // (margin 1e-9: anything closer to tangency than that goes through the filtered sign test below)
          if (!(e >= 0.0 && e * e >= rc2 * dd2 * (1.0 + 1e-9))) {
            MA_COUNT(1);
            double r2_seen;
            const unsigned long long in = sign_mask(p, P, jj, Dx, Dy, c, dd2, dw, r2_seen);
            const unsigned long long full = lowmask64(n);
            if (in == 0ull) { n = 0; phase = 2; }
            else if (in != full) { jc = jj; cDx = Dx; cDy = Dy; cc = c; cin = in; }
          }
        }
      }
    } else if (phase == 1) {
      // ---- step kind 3: the next node of the tree walk
      if (sp <= 0) {
        bool go = true;
        if (sp == 0) {  // the pass is over
          if (last || ++pass >= 64) { phase = 2; go = false; }
          prev2 = cap2;
          cap *= 2.0;
        }
        if (go) {
          // (rc2 = R2 unless the weights displace the cells: then the polygon's own radius is what says
          // whether it is already small compared with the distances this pass looks at)
          last = !(4.0 * rc2 > cap * cap) || !cut_in_pass;
          cut_in_pass = false;
          cap2 = last ? POS_INF : cap * cap;
          lo2 = prev2; hi2 = cap2;
          sp = 0;
          stk[sp++] = 0u;
        }
      }
      if (phase == 1) {
        MA_COUNT(3);
        const unsigned e = stk[--sp];
        const int l = (int)(e >> 26);
        const unsigned code = e & 0x3ffffffu;
        const double wm = p.wmax[(((size_t)1 << (2 * l)) - 1) / 3 + code];
        bool alive = wm != NEG_INF;
        {  // node inside the phase-0 block: all its sites are done
          const int sh = p.L - l;
          const int X0 = (int)morton_compact1(code) << sh, Y0 = (int)morton_compact1(code >> 1) << sh, W = 1 << sh;
          if (X0 >= bx - rdone && X0 + W - 1 <= bx + rdone && Y0 >= by - rdone && Y0 + W - 1 <= by + rdone) alive = false;
        }
        const double S = p.ph * (double)(1u << (p.L - l));
        const double ox = p.px0 + (double)morton_compact1(code) * S - xi;
        const double oy = p.py0 + (double)morton_compact1(code >> 1) * S - yi;
        const double dx = fmax(fmax(ox, -(ox + S)), 0.0), dy = fmax(fmax(oy, -(oy + S)), 0.0);
        const double d2 = dx * dx + dy * dy;
        if (d2 > cap2) alive = false;
        {
          const double fx = fmax(fabs(ox), fabs(ox + S)), fy = fmax(fabs(oy), fabs(oy + S));
          if (fx * fx + fy * fy <= prev2) alive = false;  // every site of this node was handled by an earlier pass
        }
        if (alive && d2 > 0.0 && cannot_cut(d2, wi - wm, R2)) alive = false;  // (a)
        if (alive) {
          // (a') the same with the disk (m, rc) around the polygon: for every site j of the node
          //   c_j - m.D_j = |D_j - m|^2/2 - |m|^2/2 + (w_i - w_j)/2 >= dist(m, node)^2/2 - |m|^2/2 + (w_i - wm)/2
          // and |D_j| <= far corner distance, so no bisector reaches the disk if that bound >= rc * far
          const double qx = fmax(fmax(ox - mx, mx - (ox + S)), 0.0), qy = fmax(fmax(oy - my, my - (oy + S)), 0.0);
          const double emin = 0.5 * ((qx * qx + qy * qy) - (mx * mx + my * my) + (wi - wm));
          const double fx = fmax(fabs(ox), fabs(ox + S)), fy = fmax(fabs(oy), fabs(oy + S));
          if (emin > 0.0 && emin * emin >= rc2 * (fx * fx + fy * fy) * (1.0 + 1e-9)) alive = false;
        }
        if (alive && use_m) {  // (the planes are only built when the weights vary, see PlaneGate)
          const size_t node = level_offset(l) + code;
          const double Gx = p.nodeG[2 * node], Gy = p.nodeG[2 * node + 1], al = dkey_inv(p.nodeA[node]);
          const double hs = 0.5 * S, Zx = ox + hs, Zy = oy + hs;
          bool can = false;
          for (int k = 0; k < n; ++k) {  // (b) supporting plane of the node vs every vertex
            const double ux = P.X(k), uy = P.Y(k), Px = ux - Zx, Py = uy - Zy;
            const double pp = Px * Px + Py * Py, slop = (fabs(2.0 * Px + Gx) + fabs(2.0 * Py + Gy)) * hs;
            const double r2 = ux * ux + uy * uy;
            const double lb = pp + al - slop;
            can = can || !(lb >= (r2 - wi) + 1e-10 * (pp + fabs(al) + slop + r2 + fabs(wi)));
          }
          alive = can;
        } else if (alive) {
          // (b') weights (nearly) constant: the planes are not built; every site of the node lies in its
          // square, so pow_j(p) >= dist(p, square)^2 - wm, tested vertex by vertex (a big boundary cell
          // defeats the disk of test (a), not this one)
          bool can = false;
          for (int k = 0; k < n; ++k) {
            const double ux = P.X(k), uy = P.Y(k);
            const double qx = fmax(fmax(ox - ux, ux - (ox + S)), 0.0), qy = fmax(fmax(oy - uy, uy - (oy + S)), 0.0);
            const double lhs = qx * qx + qy * qy - wm, rhs = ux * ux + uy * uy - wi;
            can = can || !(lhs >= rhs + 1e-10 * (qx * qx + qy * qy + ux * ux + uy * uy + fabs(wm) + fabs(wi)));
          }
          alive = can;
        }
        if (alive) {
          if (l < p.L) {
            // children, nearest first (pushed in reverse)
            const double cx = ox + 0.5 * S, cy = oy + 0.5 * S;
            const unsigned q0 = (cx <= 0.0 ? 1u : 0u) | (cy <= 0.0 ? 2u : 0u);
            const bool xfirst = fabs(cx) < fabs(cy);
            const unsigned q1 = q0 ^ (xfirst ? 1u : 2u), q2 = q0 ^ (xfirst ? 2u : 1u), q3 = q0 ^ 3u;
            const unsigned base = ((unsigned)(l + 1) << 26) | (code << 2);
            if (sp + 4 > CELL_STACK) { status = FLAG_STACK_OVERFLOW; n = 0; phase = 2; }
            else { stk[sp++] = base | q3; stk[sp++] = base | q2; stk[sp++] = base | q1; stk[sp++] = base | q0; }
          } else {
            j = p.bin_start[code]; jend = p.bin_start[code + 1];
          }
        }
      }
    }
  }

  // clip by the held site (requires holding()).  REFRESH = false: the caller keeps R2 current itself (block kernel:
  // sign_mask hands it the radius at the next tested candidate; until then the old, larger value is a valid bound)
  template <bool REFRESH = true> MA_DEV void clip(const Params &p, Poly &P, int maxv) {
    MA_COUNT(2);
    auto lo = [&](int tag, double &nx, double &ny, double &cl) { lineof(p, tag, nx, ny, cl); };
    int n2;
    if constexpr (Poly::PACK) n2 = clip_packed(P, n, maxv, cin, cDx, cDy, cc, jc, lo);
    else n2 = clip_rebuild(P, n, maxv, cin, cDx, cDy, cc, jc, lo);
    if (n2 < 0) { status = FLAG_CELL_OVERFLOW; n = 0; phase = 2; }
    else {
      n = n2;
      cut_in_pass = true;
      if (REFRESH) {
      R2 = 0.0;
      for (int k = 0; k < n; ++k) R2 = fmax(R2, P.X(k) * P.X(k) + P.Y(k) * P.Y(k));
      if (use_m) {  // bounding-box centre, exact radius about it
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int k = 0; k < n; ++k) {
          const double vx = P.X(k), vy = P.Y(k);
          x0 = fmin(x0, vx); x1 = fmax(x1, vx); y0 = fmin(y0, vy); y1 = fmax(y1, vy);
        }
        mx = 0.5 * (x0 + x1); my = 0.5 * (y0 + y1);
        rc2 = 0.0;
        for (int k = 0; k < n; ++k) {
          const double ax = P.X(k) - mx, ay = P.Y(k) - my;
          rc2 = fmax(rc2, ax * ax + ay * ay);
        }
      } else {
        rc2 = R2;
      }
      }
    }
    jc = -1;
  }
};

// One cell per lane.  Returns n (0 = empty, -1 = capacity overflow).
template <class Poly> MA_DEV int cell_build(const Params &p, int i, Poly &P, int maxv, int *flags_out) {
  CellSearch<Poly> S;
  unsigned stk[CELL_STACK];
  S.init(p, i, P);
  for (;;) {
    for (;;) {
      // keep searching while more lanes are searching than are holding a cutting site: both the
      // search steps and the clips then run with at least half of the unfinished lanes
      const int n_search = MA_WARP_COUNT(S.searching()), n_hold = MA_WARP_COUNT(S.holding());
      if (n_search == 0 || n_hold * p.clip_a >= n_search * p.clip_b) break;
      if (S.searching()) S.search_step(p, P, stk);
      MA_WARP_SYNC();
    }
    // warp-uniform exit: nobody holds a site and (see the loop above) nobody is searching
    if (!MA_WARP_ANY(S.holding())) break;
    if (S.holding()) S.clip(p, P, maxv);
    MA_WARP_SYNC();
  }
  if (S.status) { *flags_out |= S.status; return -1; }
  return S.n;
}

// writes the neighbour list / bounding box of a built cell
template <class Poly> MA_DEV void cell_emit(const Params &p, int i, const Poly &P, int n) {
  const double xi = p.xs[i], yi = p.ys[i];
  int *nb = p.nbr + (size_t)i * p.kmax;
  int cnt = 0;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int k = 0; k < n; ++k) {
    int t = P.T(k);
    if (t >= 0) {
      if (cnt < p.kmax) nb[cnt] = t;
      ++cnt;
    }
    x0 = fmin(x0, P.X(k)); x1 = fmax(x1, P.X(k));
    y0 = fmin(y0, P.Y(k)); y1 = fmax(y1, P.Y(k));
  }
  if (cnt > p.kmax) {
    *p.flags |= FLAG_KMAX_OVERFLOW;  // benign race: every writer ORs the same bit in
    cnt = p.kmax;
  }
  for (int k = cnt; k < p.kmax; ++k) nb[k] = -1;
  p.nbr_cnt[i] = (n <= 0) ? -1 : cnt;
  double *bb = p.cell_bb + 4 * (size_t)i;
  bb[0] = x0 + xi; bb[1] = y0 + yi; bb[2] = x1 + xi; bb[3] = y1 + yi;
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
// Per-cell half-plane table (shared memory on the GPU): slot s is neighbour J[s] with bisector
// { u . (Dx,Dy) = C } and Hessian scale S = 1 / (2 |y_i - y_j|) (kantorovich.hpp:118-119).
struct CellTable {
  double *Dx, *Dy, *C, *S;
  int *J;
  int k;
};

MA_DEV void cell_table_fill(const Params &p, int i, int s, const CellTable &T) {
  int j = p.nbr[(size_t)i * p.kmax + s];
  double Dx = p.xs[j] - p.xs[i], Dy = p.ys[j] - p.ys[i];
  double d2 = Dx * Dx + Dy * Dy;
  T.J[s] = j;
  T.Dx[s] = Dx;
  T.Dy[s] = Dy;
  T.C[s] = 0.5 * (d2 + (p.ws[i] - p.ws[j]));
  T.S[s] = 0.5 / sqrt(d2);
}

struct LaneAcc {
  double mass, cost;
  double m[5];  // ∫ρ ux, ∫ρ uy, ∫ρ ux², ∫ρ uy², ∫ρ ux uy (local coordinates)
  unsigned long long touched;
  unsigned long long cnt[CNT_N];
  int npieces, nverts;  // pieces modes
};

struct Tri {  // one candidate face: global vertex coordinates + id
  double gx[3], gy[3];
  int f;
};

// Clip triangle t (local coordinates already in P, tags -1,-2,-3 for edges a->b, b->c, c->a) by the
// k half-planes of the cell.  Returns the vertex count of the piece (0 = empty, -1 overflow).
template <int NT>
MA_DEV int piece_clip(const Params &p, const PolyRef<NT> &P, int maxv, const CellTable &T, const Tri &t, double xi,
                      double yi, double wi, LaneAcc &acc) {
  int n = 3;
  const double ax = t.gx[0] - xi, ay = t.gy[0] - yi, bx = t.gx[1] - xi, by = t.gy[1] - yi, cx = t.gx[2] - xi,
               cy = t.gy[2] - yi;
  P.X(0) = ax; P.Y(0) = ay; P.T(0) = -1;
  P.X(1) = bx; P.Y(1) = by; P.T(1) = -2;
  P.X(2) = cx; P.Y(2) = cy; P.T(2) = -3;
  auto lineof = [&](int tag, double &nx, double &ny, double &cl) {
    if (tag >= 0) { nx = T.Dx[tag]; ny = T.Dy[tag]; cl = T.C[tag]; }
    else {
      double px = tag == -1 ? ax : (tag == -2 ? bx : cx), py = tag == -1 ? ay : (tag == -2 ? by : cy);
      double qx = tag == -1 ? bx : (tag == -2 ? cx : ax), qy = tag == -1 ? by : (tag == -2 ? cy : ay);
      nx = py - qy; ny = qx - px; cl = nx * px + ny * py;
    }
  };
  // exact-ish line of an edge tag from the ORIGINAL inputs (double-double)
  auto ddlineof = [&](int tag) -> ddline {
    if (tag >= 0) { int j = T.J[tag]; return dd_bisector(xi, yi, wi, p.xs[j], p.ys[j], p.ws[j]); }
    int a = -1 - tag, b = (a == 2) ? 0 : a + 1;
    return dd_mesh_edge(xi, yi, t.gx[a], t.gy[a], t.gx[b], t.gy[b]);
  };
  for (int s = 0; s < T.k && n > 0; ++s) {
    const double Dx = T.Dx[s], Dy = T.Dy[s], c = T.C[s];
    unsigned long long in = 0ull;
    for (int k = 0; k < n; ++k) {
      double tx = P.X(k) * Dx, ty = P.Y(k) * Dy;
      double val = c - (tx + ty);
      bool inside = val > 0.0;
      if (fabs(val) <= p.filter_tol * (fabs(c) + fabs(tx) + fabs(ty))) {
        // filtered predicate failed: decide from the original data (Side1 / Side2 / Side3,
        // predicates.hpp:73-135); ties are outside (strict "== SMALLER", App. B T3)
        acc.cnt[CNT_FALLBACK]++;
        int tb = P.T(k), ta = P.T(k == 0 ? n - 1 : k - 1);
        ddline Lt = ddlineof(s);
        int sg;
        if (ta < 0 && tb < 0) {
          int v = -1 - tb;  // edge tb starts at triangle vertex (-1-tb); edge ta ends there
          sg = dd_side_point(xi, yi, t.gx[v], t.gy[v], Lt);
        } else {
          sg = dd_side(ddlineof(ta), ddlineof(tb), Lt);
        }
        inside = sg > 0;
      }
      if (inside) in |= 1ull << k;
    }
    unsigned long long full = (1ull << n) - 1ull;
    if (in == full) continue;
    if (in == 0ull) { n = 0; break; }
    n = clip_rebuild(P, n, maxv, in, Dx, Dy, c, s, lineof);
    if (n < 0) return -1;
  }
  return n;
}

// kantorovich.hpp:105-135 on one piece; rho(u) = a ux + b uy + r0 in local coordinates.
template <int NT>
MA_DEV void piece_kantorovich(const PolyRef<NT> &P, int n, double a, double b, double r0, const CellTable &T,
                              double *hacc, int hstride, LaneAcc &acc) {
  for (int k = 0; k < n; ++k) {  // Hessian: midpoint rule on every Laguerre edge (:110-122, quadrature.hpp:79-85)
    int tg = P.T(k);
    if (tg < 0) continue;
    int kk = (k + 1 == n) ? 0 : k + 1;
    double x0 = P.X(k), y0 = P.Y(k), x1 = P.X(kk), y1 = P.Y(kk);
    double ex = x1 - x0, ey = y1 - y0;
    double r = sqrt(ex * ex + ey * ey) * (a * (0.5 * (x0 + x1)) + b * (0.5 * (y0 + y1)) + r0);
    hacc[tg * hstride] += r * T.S[tg];
    acc.touched |= 1ull << tg;
    acc.cnt[CNT_LEDGES]++;
  }
  const double x0 = P.X(0), y0 = P.Y(0);
  double mass = 0.0, cost = 0.0;
  for (int q = 1; q + 1 < n; ++q) {  // fan from vertex 0 (quadrature.hpp:128-129, 140-141)
    double x1 = P.X(q), y1 = P.Y(q), x2 = P.X(q + 1), y2 = P.Y(q + 1);
    double A = 0.5 * ((x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0));
    const double third = 1.0 / 3.0;
    mass += A * (a * ((x0 + x1 + x2) * third) + b * ((y0 + y1 + y2) * third) + r0);  // centroid rule (:69-77)
    double s = 0.0;
    albrecht_collatz_points(x0, y0, x1, y1, x2, y2, [&](double x, double y, double w) {
      s += w * ((a * x + b * y + r0) * (x * x + y * y));  // fv(p) * |p - y_v|^2 (kantorovich.hpp:126-131)
    });
    cost += A * s;
  }
  acc.mass += mass;
  acc.cost += cost;
}

// lloyd.hpp:57-67 / :100-121 on one piece, in local coordinates (shifted to global by the caller).
template <int NT, int ORDER>
MA_DEV void piece_moments(const PolyRef<NT> &P, int n, double a, double b, double r0, LaneAcc &acc) {
  const double x0 = P.X(0), y0 = P.Y(0);
  for (int q = 1; q + 1 < n; ++q) {
    double x1 = P.X(q), y1 = P.Y(q), x2 = P.X(q + 1), y2 = P.Y(q + 1);
    double A = 0.5 * ((x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0));
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    albrecht_collatz_points(x0, y0, x1, y1, x2, y2, [&](double x, double y, double w) {
      double fp = w * (a * x + b * y + r0);
      s0 += fp; s1 += fp * x; s2 += fp * y;
      if (ORDER == 2) { s3 += fp * x * x; s4 += fp * y * y; s5 += fp * x * y; }
    });
    acc.mass += A * s0; acc.m[0] += A * s1; acc.m[1] += A * s2;
    if (ORDER == 2) { acc.m[2] += A * s3; acc.m[3] += A * s4; acc.m[4] += A * s5; }
  }
}

template <int NT> MA_DEV void piece_stats(const PolyRef<NT> &P, int n, int k, LaneAcc &acc) {
  acc.cnt[CNT_PIECES]++;
  acc.cnt[CNT_VERTS] += n;
  acc.cnt[CNT_SUMKNP] += (unsigned long long)k * n;
  for (int q = 0; q < n; ++q) {
    int ta = P.T(q == 0 ? n - 1 : q - 1), tb = P.T(q);
    if (!(ta < 0 && tb < 0)) acc.cnt[CNT_NEWV]++;
  }
}

// process one candidate face
template <int NT, int MODE>
MA_DEV void lane_face(const Params &p, int i, const PolyRef<NT> &P, int maxv, const CellTable &T, const Tri &t,
                      double xi, double yi, double wi, double *hacc, int hstride, LaneAcc &acc, int lane_piece_base,
                      int lane_vert_base) {
  if (p.stats) acc.cnt[CNT_CAND]++;
  int n = piece_clip<NT>(p, P, maxv, T, t, xi, yi, wi, acc);
  if (n < 0) { *p.flags |= FLAG_PIECE_OVERFLOW; return; }
  if (n < 3) return;
  if (p.stats) piece_stats<NT>(P, n, T.k, acc);
  if (MODE == MODE_PIECES_COUNT) { acc.npieces++; acc.nverts += n; return; }
  if (MODE == MODE_PIECES_FILL) {
    int pi = lane_piece_base + acc.npieces, vi = lane_vert_base + acc.nverts;
    p.pc_cell[pi] = i;
    p.pc_face[pi] = t.f;
    p.pc_ptr[pi] = vi;
    for (int k = 0; k < n; ++k) {
      p.pc_xy[2 * (size_t)(vi + k)] = P.X(k) + xi;
      p.pc_xy[2 * (size_t)(vi + k) + 1] = P.Y(k) + yi;
      int tg = P.T(k);
      p.pc_tag[vi + k] = tg >= 0 ? T.J[tg] : tg;  // -1, -2, -3: edge (0,1), (1,2), (2,0) of the face
    }
    acc.npieces++; acc.nverts += n;
    return;
  }
  const double a = p.abc[3 * (size_t)t.f], b = p.abc[3 * (size_t)t.f + 1], c0 = p.abc[3 * (size_t)t.f + 2];
  const double r0 = c0 + a * xi + b * yi;
  if (MODE == MODE_KANTOROVICH) piece_kantorovich<NT>(P, n, a, b, r0, T, hacc, hstride, acc);
  if (MODE == MODE_MOMENTS1) piece_moments<NT, 1>(P, n, a, b, r0, acc);
  if (MODE == MODE_MOMENTS2) piece_moments<NT, 2>(P, n, a, b, r0, acc);
}

// All candidate faces of cell i handled by `lane` out of `nlanes` (round-robin).  For the pieces
// modes each lane's output range is [lane_piece_base, ...) computed by the caller.
template <int NT, int MODE>
MA_DEV void lane_pieces(const Params &p, int i, int lane, int nlanes, const PolyRef<NT> &P, int maxv,
                        const CellTable &T, double *hacc, int hstride, LaneAcc &acc, int lane_piece_base = 0,
                        int lane_vert_base = 0) {
  const double xi = p.xs[i], yi = p.ys[i], wi = p.ws[i];
  const double *cb = p.cell_bb + 4 * (size_t)i;
  Tri t;
  if (p.mesh_kind == MESH_GRID) {
    // squares overlapped by the cell's bounding box
    const double pad = 1e-12;
    int i0 = (int)floor((cb[0] - p.gx0) / p.gdx - pad), i1 = (int)floor((cb[2] - p.gx0) / p.gdx + pad);
    int j0 = (int)floor((cb[1] - p.gy0) / p.gdy - pad), j1 = (int)floor((cb[3] - p.gy0) / p.gdy + pad);
    i0 = max(i0, 0); j0 = max(j0, 0);
    i1 = min(i1, p.gn - 2); j1 = min(j1, p.gm - 2);
    if (i1 < i0 || j1 < j0) return;
    const int nrows = j1 - j0 + 1;
    const int ncand = (i1 - i0 + 1) * nrows * 2;
    for (int c = lane; c < ncand; c += nlanes) {
      int sq = c >> 1, which = c & 1;
      int ci = sq / nrows, cj = sq - ci * nrows;
      int si = i0 + ci, sj = j0 + cj;
      double X0 = p.gx0 + si * p.gdx, X1 = p.gx0 + (si + 1) * p.gdx;
      double Y0 = p.gy0 + sj * p.gdy, Y1 = p.gy0 + (sj + 1) * p.gdy;
      t.gx[0] = X0; t.gy[0] = Y0;
      if (which == 0) { t.gx[1] = X1; t.gy[1] = Y0; t.gx[2] = X1; t.gy[2] = Y1; }
      else            { t.gx[1] = X1; t.gy[1] = Y1; t.gx[2] = X0; t.gy[2] = Y1; }
      t.f = 2 * (si * (p.gm - 1) + sj) + which;
      lane_face<NT, MODE>(p, i, P, maxv, T, t, xi, yi, wi, hacc, hstride, acc, lane_piece_base, lane_vert_base);
    }
  } else {
    // general mesh: faces binned on a tg x tg grid over the mesh box; a face is handled in the first
    // bin of (its bin range ∩ the query window)
    const int g = p.tg;
    int i0 = min(max((int)floor((cb[0] - p.bb[0]) * p.tinvx - 1e-9), 0), g - 1);
    int i1 = min(max((int)floor((cb[2] - p.bb[0]) * p.tinvx + 1e-9), 0), g - 1);
    int j0 = min(max((int)floor((cb[1] - p.bb[1]) * p.tinvy - 1e-9), 0), g - 1);
    int j1 = min(max((int)floor((cb[3] - p.bb[1]) * p.tinvy + 1e-9), 0), g - 1);
    int seen = 0;
    for (int bj = j0; bj <= j1; ++bj)
      for (int bi = i0; bi <= i1; ++bi) {
        const int q0 = p.tbin_ptr[bj * g + bi], q1 = p.tbin_ptr[bj * g + bi + 1];
        for (int q = q0; q < q1; ++q, ++seen) {
          if (seen % nlanes != lane) continue;
          int f = p.tbin_face[q];
          double fx0 = 1e300, fy0 = 1e300;
          for (int k = 0; k < 3; ++k) {
            int v = p.tri[3 * (size_t)f + k];
            t.gx[k] = p.vx[v]; t.gy[k] = p.vy[v];
            fx0 = fmin(fx0, t.gx[k]); fy0 = fmin(fy0, t.gy[k]);
          }
          int fi0 = max(min(max((int)floor((fx0 - p.bb[0]) * p.tinvx), 0), g - 1), i0);
          int fj0 = max(min(max((int)floor((fy0 - p.bb[1]) * p.tinvy), 0), g - 1), j0);
          if (fi0 != bi || fj0 != bj) continue;
          t.f = f;
          lane_face<NT, MODE>(p, i, P, maxv, T, t, xi, yi, wi, hacc, hstride, acc, lane_piece_base, lane_vert_base);
        }
      }
  }
}

MA_DEV void lane_acc_zero(LaneAcc &a) {
  a.mass = a.cost = 0.0;
  for (int k = 0; k < 5; ++k) a.m[k] = 0.0;
  a.touched = 0ull;
  for (int k = 0; k < CNT_N; ++k) a.cnt[k] = 0ull;
  a.npieces = a.nverts = 0;
}

}  // namespace ma

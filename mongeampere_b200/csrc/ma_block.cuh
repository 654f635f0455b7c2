// ma_block.cuh — K2, fast path: the power cell of a Dirac from the sites of the (2R+1)^2 block of leaf bins
// around its own bin, with a per-vertex security certificate.
//
// Replaces CGAL Regular_triangulation_2's neighbour circulator (kantorovich.hpp:65-72, vti.hpp:265) for the
// cells whose neighbours all live in that block — 85 % of the cells of a uniform point set at ~1 Dirac per
// bin with R = 2, 99.6 % with R = 3 (measured).  The remaining cells go to CellSearch (ma_cell.cuh: ring
// walk + quadtree) through a compacted list.
//
// Why a second kernel for the same job: CellSearch is a resumable state machine that handles one site per
// step with a dozen data-dependent branches around it (16.7 k thread-instructions per cell, 9 of 32 lanes
// active, profiles/r01n_summary.md).  Here the candidate list is known up front — the sites are kept a
// second time in ROW-MAJOR bin order, so the bins [bx-R, bx+R] of one bin row are one contiguous run and a
// block is 2R+1 runs — and every lane of a warp walks its own list in lock step: one candidate per
// iteration, the same straight-line code for all lanes, warp votes to skip the sign loop / the clip when
// no lane needs them.
//
// Certificate.  After the block every site inside the block's rectangle B has been tested.  A site j outside
// B takes vertex p from i iff |p - y_j|^2 - w_j < |p - y_i|^2 - w_i; with w_j <= wmax this needs
// dist(p, outside of B)^2 < |p - y_i|^2 - w_i + wmax.  If that fails for every vertex the polygon is the
// cell (the per-vertex form certifies more cells than "security radius >= 2 R", which measures the worst
// vertex against the nearest side of B).  Sides of B on the rim of the bin grid are unbounded: there are no
// sites beyond it.
#pragma once
#include "ma_cell.cuh"

namespace ma {

// the block kernels run while max w - min w <= MA_LEAN_RANGE bin areas of the block grid
#ifndef MA_LEAN_RANGE
#define MA_LEAN_RANGE 4.0
#endif

#ifdef __CUDA_ARCH__
#define MA_WARP_MAX_INT(v) ((int)__reduce_max_sync(0xffffffffu, (int)(v)))
#else
#define MA_WARP_MAX_INT(v) (v)
#endif

// The candidate runs of one cell: the bins of the block of radius R around bin (cbx, cby) that are NOT in the block
// of radius R0 (R0 = -1: the whole block), row by row, nearest rows first.  Rows inside the inner block contribute
// the two strips left and right of it, the others one full run.
template <int R0, int R> struct BlockRuns {
  static constexpr int NR = (R0 < 0) ? (2 * R + 1) : (2 * (2 * R0 + 1) + 2 * (R - R0));
  int cum[NR + 1], off[NR];
  MA_DEV void build(const Params &p, int cbx, int cby, bool active) {
    const int G = p.bG;
    int q = 0;
    cum[0] = 0;
    auto run = [&](int row, int xa, int xb) {  // bins [xa, xb] of row `row` (may be empty / outside)
      int s = 0, len = 0;
      xa = max(xa, 0); xb = min(xb, G - 1);
      if (active && row >= 0 && row < G && xb >= xa) {
        s = p.rm_start[(size_t)row * G + xa];
        len = p.rm_start[(size_t)row * G + xb + 1] - s;
      }
      off[q] = s - cum[q];
      cum[q + 1] = cum[q] + len;
      ++q;
    };
#pragma unroll
    for (int r = 0; r < 2 * R + 1; ++r) {
      const int dy = (r == 0) ? 0 : ((r & 1) ? -((r + 1) >> 1) : (r >> 1));  // own row first, then -1, +1, -2, +2 ...
      const int ady = dy < 0 ? -dy : dy;
      if (ady <= R0) { run(cby + dy, cbx - R, cbx - R0 - 1); run(cby + dy, cbx + R0 + 1, cbx + R); }
      else run(cby + dy, cbx - R, cbx + R);
    }
  }
  MA_DEV int total() const { return cum[NR]; }
  MA_DEV int position(int t) const {  // index of candidate t in the row-major arrays
    int o = off[0];
#pragma unroll
    for (int r = 1; r < NR; ++r) o = (t >= cum[r]) ? off[r] : o;
    return t + o;
  }
};

// Does every vertex of the polygon keep its distance from the sites outside the block of radius R (see the header)?
template <int R, class Poly>
MA_DEV bool block_certified(const Params &p, const CellSearch<Poly> &S, const Poly &P, int cbx, int cby) {
  const int G = p.bG;
  const double INF = 1.0 / 0.0;
  const double bl = (cbx - R > 0) ? p.px0 + (double)(cbx - R) * p.bph - S.xi : -INF;
  const double br = (cbx + R < G - 1) ? p.px0 + (double)(cbx + R + 1) * p.bph - S.xi : INF;
  const double bb = (cby - R > 0) ? p.py0 + (double)(cby - R) * p.bph - S.yi : -INF;
  const double bt = (cby + R < G - 1) ? p.py0 + (double)(cby + R + 1) * p.bph - S.yi : INF;
  const double dwg = p.wstat[3] - S.wi;  // >= 0 (wstat[3] = the largest weight of this evaluation)
  const double slack = 1e-9 * p.bph;    // a site sits in its bin up to the rounding of the bin index
  bool ok = true;
  for (int k = 0; k < S.n; ++k) {
    const double X = P.X(k), Y = P.Y(k);
    const double rp = X * X + Y * Y + dwg;
    const double d = fmin(fmin(X - bl, br - X), fmin(Y - bb, bt - Y)) - slack;
    ok = ok && (rp <= 0.0 || (d > 0.0 && d * d >= rp * (1.0 + 1e-9)));
  }
  return ok;
}

// Reloads the polygon a previous pass left for cell S.i (k_cells_block stores it for the cells it cannot certify);
// false if there is none (the cell outgrew the 16-vertex class there: it goes straight to CellSearch).
template <class Poly> MA_DEV bool block_reload(const Params &p, CellSearch<Poly> &S, Poly &P) {
  const int n = p.poly_n[S.i];
  if (n < 3 || n > 16) return false;
  double r2 = 0.0;
  for (int k = 0; k < n; ++k) {
    const size_t o = (size_t)k * p.N + S.i;
    const double X = p.poly_x[o], Y = p.poly_y[o];
    P.SX(k) = X; P.SY(k) = Y; P.ST(k) = p.poly_t[o];
    r2 = fmax(r2, X * X + Y * Y);
  }
  P.ord = 0xfedcba9876543210ull;  // vertex k in slot k
  P.used = (1u << n) - 1u;
  S.n = n;
  S.R2 = r2;
  return true;
}

// Builds the cell of S.i from the block of radius R.  R0 < 0: from scratch (S.init() already done); R0 >= 0: the
// polygon in P is what the block of radius R0 left (a previous pass), only the bins beyond it are looked at.
// `active` = false lanes only keep the warp's votes company.  On return S.phase == 0: polygon of S.n vertices in P,
// certified says whether it is final; S.phase == 2: S.n == 0 and S.status tells an empty cell (0) from a capacity overflow.
// One lock-step walk over the candidates of the bins of block R that are not in block R0 (see BlockRuns).
template <int R0, int R, class Poly>
MA_DEV void block_walk(const Params &p, CellSearch<Poly> &S, Poly &P, int maxv, bool active, int cbx, int cby) {
  BlockRuns<R0, R> runs;
  runs.build(p, cbx, cby, active);
  const int T = runs.total();
  // All lanes step through their lists together, candidate t of every lane in iteration t.  (Per-lane cursors with
  // vote-scheduled clips — the CellSearch scheme — were measured on this kernel: 7 % fewer instructions, 15 of 32 lanes
  // instead of 14, but 18 % SLOWER: the candidate loads then issue lane by lane behind data-dependent branches and the
  // kernel is bound by their latency at 20 warps per SM, profiles/r02g.  Queueing the candidates that pass the disk test
  // and clipping them in batches, all lanes busy, was measured too: 1.9x SLOWER (2.73 against 1.43 ms of K2 at c3,
  // profiles/r02k) — a polygon that is clipped late stays large, so far more candidates pass the test and are clipped
  // at all; clipping the moment a candidate is met is what keeps the work per cell at ~8 clips.)
  const int Tmax = MA_WARP_MAX_INT(T);
  // (the next candidate's coordinates are fetched one iteration ahead: the loads then fly while this one is tested / clipped)
  int npos = (active && 0 < T) ? runs.position(0) : 0;
  double nxr = p.xr[npos], nyr = p.yr[npos], nwr = p.wr[npos];
  for (int t = 0; t < Tmax; ++t) {
    const bool act = active && S.phase == 0 && t < T;
    const int pos = npos;
    const double Dx = nxr - S.xi, Dy = nyr - S.yi, wj = nwr;
    npos = (active && t + 1 < T) ? runs.position(t + 1) : 0;
    nxr = p.xr[npos]; nyr = p.yr[npos]; nwr = p.wr[npos];
    const double dd2 = Dx * Dx + Dy * Dy;
    const double dw = S.wi - wj;
    const double c = 0.5 * (dd2 + dw);
    MA_COUNT(0);
    if (act && dd2 == 0.0) {  // itself, or a coincident site: the heavier (then the earlier) one keeps the cell
      const int jj = p.rm2s[pos];
      if (jj != S.i && (wj > S.wi || (wj == S.wi && jj < S.i))) { S.n = 0; S.phase = 2; }
    }
    // the bisector misses the disk of radius sqrt(R2) around y_i that contains the polygon => it cannot cut
    const bool test = act && S.phase == 0 && dd2 > 0.0 && !(c >= 0.0 && c * c >= S.R2 * dd2 * (1.0 + 1e-9));
    bool cut = false;
    if (MA_WARP_ANY(test)) {
      if (test) {
        MA_COUNT(1);
        const int n = S.n;
        const int jj = p.rm2s[pos];
        double r2_seen;
        const unsigned long long in = S.sign_mask(p, P, jj, Dx, Dy, c, dd2, dw, r2_seen);
        S.R2 = r2_seen;  // radius of the polygon as it is now (a clip below leaves it as an upper bound)
        if (in == 0ull) { S.n = 0; S.phase = 2; }
        else if (in != lowmask64(n)) { S.jc = jj; S.cDx = Dx; S.cDy = Dy; S.cc = c; S.cin = in; cut = true; }
      }
      MA_WARP_SYNC();
      if (MA_WARP_ANY(cut)) {
        if (cut) S.template clip<false>(p, P, maxv);  // overflow: status = FLAG_CELL_OVERFLOW, n = 0, phase = 2
        MA_WARP_SYNC();
      }
    }
  }
}

// Builds the cell of S.i from the block of radius R.  R0 < 0: from scratch (S.init() already done); R0 >= 0: the
// polygon in P is what the block of radius R0 left (a previous pass), only the bins beyond it are looked at.
// `active` = false lanes only keep the warp's votes company.  On return S.phase == 0: polygon of S.n vertices in P,
// certified says whether it is final; S.phase == 2: S.n == 0 and S.status tells an empty cell (0) from a capacity overflow.
#ifndef MA_K2B_INNER_FIRST
#define MA_K2B_INNER_FIRST 0  // 1: the 3 x 3 block first, then the ring around it (nearest candidates first)
#endif
template <int R0, int R, class Poly>
MA_DEV void block_search(const Params &p, CellSearch<Poly> &S, Poly &P, int maxv, bool active, bool &certified) {
  const int G = p.bG;
  // the Dirac's bin of the block grid (same expression as k_blk_count, so a site is in the bin it was filed under)
  const int cbx = min(max((int)((S.xi - p.px0) * p.binv), 0), G - 1);
  const int cby = min(max((int)((S.yi - p.py0) * p.binv), 0), G - 1);
  if constexpr (MA_K2B_INNER_FIRST && R0 < 0 && R == 2) {
    block_walk<-1, 1>(p, S, P, maxv, active, cbx, cby);
    block_walk<1, 2>(p, S, P, maxv, active, cbx, cby);
  } else {
    block_walk<R0, R>(p, S, P, maxv, active, cbx, cby);
  }
  // ---- certificate ----
  if (S.phase == 0) certified = block_certified<R>(p, S, P, cbx, cby);
  else certified = S.status == 0;  // hidden Dirac: one site covers a superset of the cell, nothing to certify
}

}  // namespace ma

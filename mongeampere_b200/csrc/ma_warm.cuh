// ma_warm.cuh — K2 from the adjacency of an EARLIER evaluation of the same point set, with a combinatorial
// certificate.  The regime of the damped-Newton loop (optimal_transport.hpp:116-187): the Diracs stay, the weights
// move a little from one evaluation to the next, and they are graded (cells far from their Diracs), which is where
// the distance-based certificates of ma_block.cuh give up and CellSearch's quadtree walk costs ~20 ms per million
// cells — nearly all of it spent PROVING that nobody else cuts, not finding the neighbours.
//
// Step 1 (seed_build): cell i = box ∩ half-planes of the neighbours it had in the seed evaluation: a superset C'_i of
//         the true cell C_i (fewer constraints), ~6 clips.  Its cyclic sequence of edge tags (neighbour index, or
//         -1..-4 for a side of the box) is written to ring[16 i ..].
// Step 2 (ring_match): the vertex of C'_i between the edges tagged (a, b) — incoming a, outgoing b, CCW — is the radical
//         centre of (i, a, b); in a power diagram the same point is the vertex of cell a between (b, i) and of cell b
//         between (i, a).  The check is purely combinatorial: does ring[a] contain b immediately followed by i?  (For a
//         vertex on a side s of the box, (a, s) at i pairs with (s, i) at a.)
// Claim: if EVERY vertex of EVERY cell passes and the cell masses add up to the mass of the box, then C'_i = C_i for
//         all i.  Proof: matched edges coincide with opposite orientation (they join the same two radical centres), and
//         around a vertex the three convex sectors are those of the 3-site diagram, so gluing the polygons along matched
//         edges gives a surface that maps to the box by a local homeomorphism, boundary to boundary: a covering of a
//         simply connected set, i.e. m disjoint sheets, each a tiling of the box; the total mass gives m = 1.  The true
//         cells tile the box too and C_i ⊆ C'_i, hence equality.  Every sign decision on the way is filtered / exact
//         (CellSearch::sign_mask), so the polygons are also bit-identical to what the cold path builds.
// Step 3: the cells at a failing vertex (all three of them) are rebuilt by CellSearch (exact, state 1) and the match is
//         repeated; whatever still fails after the last round makes the host redo the evaluation without seeds.
// A Delaunay flip changes four cells, so the rebuilt fraction is a few times the fraction of flipped edges.
#pragma once
#include "ma_cell.cuh"

namespace ma {

constexpr int RING_STRIDE = 16;  // the warm path is for the 16-vertex class (packed-order polygons)
enum { WARM_SEEDED = 0, WARM_EXACT = 1, WARM_QUEUED = 2 };

// seed row: the neighbours (site indices, `cnt` of them) cell S.i had in the seed evaluation.  All lanes in lock step.
template <class Poly>
MA_DEV void seed_build(const Params &p, CellSearch<Poly> &S, Poly &P, int maxv, bool active, const int *seed_row, int cnt) {
  const int cmax = MA_WARP_MAX_INT(active ? cnt : 0);
  for (int s = 0; s < cmax; ++s) {
    const bool act = active && S.phase == 0 && s < cnt;
    const int j = act ? seed_row[s] : S.i;
    const double Dx = p.xs[j] - S.xi, Dy = p.ys[j] - S.yi, dw = S.wi - p.ws[j];
    const double dd2 = Dx * Dx + Dy * Dy, c = 0.5 * (dd2 + dw);
    bool cut = false;
    if (act && dd2 > 0.0 && !(c >= 0.0 && c * c >= S.R2 * dd2 * (1.0 + 1e-9))) {
      const int n = S.n;
      double r2;
      const unsigned long long in = S.sign_mask(p, P, j, Dx, Dy, c, dd2, dw, r2);
      S.R2 = r2;
      if (in == 0ull) { S.n = 0; S.phase = 2; }  // the superset is empty, so is the cell
      else if (in != lowmask64(n)) { S.jc = j; S.cDx = Dx; S.cDy = Dy; S.cc = c; S.cin = in; cut = true; }
    }
    MA_WARP_SYNC();
    if (MA_WARP_ANY(cut)) {
      if (cut) S.template clip<false>(p, P, maxv);  // overflow: status = FLAG_CELL_OVERFLOW, n = 0, phase = 2
      MA_WARP_SYNC();
    }
  }
}

template <class Poly> MA_DEV void ring_store(int *ring, int *ring_n, int i, const Poly &P, int n) {
  ring_n[i] = n;
  int *r = ring + (size_t)RING_STRIDE * i;
  for (int k = 0; k < n && k < RING_STRIDE; ++k) r[k] = P.T(k);
}

// does the ring of cell a contain the tag u immediately followed by v?
MA_DEV bool ring_has(const int *ring, const int *ring_n, int a, int u, int v) {
  const int n = ring_n[a];
  if (n < 3 || n > RING_STRIDE) return false;
  const int *r = ring + (size_t)RING_STRIDE * a;
  bool f = false;
  int prev = r[n - 1];
  for (int k = 0; k < n; ++k) {
    const int cur = r[k];
    f = f || (prev == u && cur == v);
    prev = cur;
  }
  return f;
}

// All vertices of cell i against the rings of its neighbours; push(cell) is called for the three cells of every
// failing vertex.  An empty cell (n == 0) has nothing to check: its emptiness was proven on a superset.
template <class Push> MA_DEV bool ring_match(const int *ring, const int *ring_n, int i, Push push) {
  const int n = ring_n[i];
  if (n == 0) return true;
  if (n < 3 || n > RING_STRIDE) { push(i); return false; }
  const int *r = ring + (size_t)RING_STRIDE * i;
  bool ok = true;
  int a = r[n - 1];
  for (int k = 0; k < n; ++k) {
    const int b = r[k];  // vertex k: incoming edge tagged a, outgoing edge tagged b
    bool good = true;
    if (a >= 0) good = ring_has(ring, ring_n, a, b, i);
    else if (b >= 0) good = ring_has(ring, ring_n, b, i, a);
    if (!good) {
      ok = false;
      push(i);
      if (a >= 0) push(a);
      if (b >= 0) push(b);
    }
    a = b;
  }
  return ok;
}

}  // namespace ma

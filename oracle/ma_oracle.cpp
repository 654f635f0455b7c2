// ma_oracle.cpp — CPU restatement (ORACLE) of MongeAmpere++'s Kantorovich hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product path
// (mongeampere_b200/, include/) never calls it and has no CPU fallback.
//
// *** PARITY UNPINNED ***: the reference (mrgt/MongeAmpere @84acaeaf) needs CGAL 4.11, Eigen,
// Boost and CImg, none of which exist in this image, and it ships no golden vectors or
// asserting tests.  This file therefore restates the reference's
// arithmetic line by line and is pinned only by (i) closed-form known answers
// (tests/golden/), (ii) the invariants the reference's drivers print (area / mass
// conservation, finite-difference gradient & Hessian), (iii) an independent SciPy/Qhull
// lifted-hull adjacency cross-check, (iv) __float128 arbitration of near-ties and (v) independent golden vectors
// computed in exact rational arithmetic without this file and without the engine (tests/golden/make_independent.py,
// tests/test_independent_golden.py).  Known limit: the reference's global-coordinate radical axis
// (predicates.hpp:46-52), restated here as it stands, loses digits on short Laguerre edges and on cells far from
// their Dirac (up to 3e-8 row-relative in H on graded weights); tests arbitrate those cases with exact arithmetic.
//
// What follows the reference (paths relative to /root/reference):
//   include/MA/kantorovich.hpp:59-141                        -> Oracle::kantorovich / piece_callback
//   include/MA/voronoi_triangulation_intersection.hpp:129-198 -> inside(), clip_by_bisector()
//   include/MA/voronoi_triangulation_intersection.hpp:219-313 -> overlay_bfs()
//   include/MA/predicates.hpp:21-30,35-71,73-135             -> line_line(), radical_axis(), side1/2/3
//   include/MA/quadrature.hpp:24-42,69-85,122-143            -> albrecht_collatz(), centroid rule, midpoint
//   include/MA/functions.hpp:25-80                           -> mao_linear_functions()
//   include/MA/lloyd.hpp:30-144                              -> Oracle::moments
//   include/MA/optimal_transport.hpp:41-87                   -> mao_solve_laplacian (CG stand-in, see note)
//   include/MA/common_rt.hpp:31-53 (CGAL Regular_triangulation_2) -> power_neighbors(): CGAL is an
//       un-vendored dependency (README.md:27-33, CGAL 4.11).  The regular triangulation is
//       unique away from degeneracies, so its neighbour sets are recomputed here by direct
//       half-plane clipping of each power cell against a bounding box (brute force for small N,
//       quadtree-pruned search for large N); tests cross-check against Qhull's lifted hull.
//
// Build: g++ -O3 -march=native -std=c++17 -fopenmp -shared -fPIC ma_oracle.cpp -o _build/libma_oracle.so -lquadmath
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __float128 quad;

namespace {

// ---------------------------------------------------------------------------------------------
// small templated geometry, instantiated for double (fast path) and __float128 (arbitration)
// ---------------------------------------------------------------------------------------------
template <class T> struct Line { T a, b, c; };       // a x + b y + c = 0
template <class T> struct Pt { T x, y; };

template <class T> static inline T tabs(T v) { return v < 0 ? -v : v; }

// CGAL::Line_2(p, q) through two points (used for EDGE_T edges,
// voronoi_triangulation_intersection.hpp:101-105).
template <class T> static inline Line<T> line_through(T px, T py, T qx, T qy) {
  Line<T> l;
  l.a = py - qy;
  l.b = qx - px;
  l.c = -px * l.a - py * l.b;
  return l;
}

// CGAL::radical_axis(p, q) of two weighted points (predicates.hpp:46-52).
template <class T> static inline Line<T> radical_axis(T px, T py, T pw, T qx, T qy, T qw) {
  Line<T> l;
  l.a = T(2) * (px - qx);
  l.b = T(2) * (py - qy);
  l.c = -(px * px) - (py * py) + (qx * qx) + (qy * qy) + pw - qw;
  return l;
}

// CGAL::intersection(Line, Line) (predicates.hpp:21-30).  cond returns |den| / (|a1 b2| + |a2 b1|)
// so that callers can detect ill-conditioned constructions.
template <class T> static inline Pt<T> line_line(const Line<T> &L, const Line<T> &M, T *cond = nullptr) {
  T den = L.a * M.b - M.a * L.b;
  T nx = L.b * M.c - M.b * L.c;
  T ny = M.a * L.c - L.a * M.c;
  if (cond) {
    T mag = tabs(L.a * M.b) + tabs(M.a * L.b);
    *cond = mag > 0 ? tabs(den) / mag : T(0);
  }
  Pt<T> p;
  p.x = nx / den;
  p.y = ny / den;
  return p;
}

// CGAL::weighted_circumcenter(p, q, r) (predicates.hpp:54-61): the point with equal power
// to the three weighted points = intersection of two radical axes.
template <class T>
static inline Pt<T> weighted_circumcenter(T px, T py, T pw, T qx, T qy, T qw, T rx, T ry, T rw, T *cond) {
  return line_line(radical_axis(px, py, pw, qx, qy, qw), radical_axis(px, py, pw, rx, ry, rw), cond);
}

// pow_v(E) - pow_w(E)  (< 0  <=>  E strictly closer to v: predicates.hpp:78-87, "== SMALLER").
template <class T>
static inline T power_diff(T ex, T ey, T vx, T vy, T vw, T wx, T wy, T ww, T *mag) {
  T dv = (ex - vx) * (ex - vx) + (ey - vy) * (ey - vy) - vw;
  T dw = (ex - wx) * (ex - wx) + (ey - wy) * (ey - wy) - ww;
  if (mag) {
    *mag = (ex - vx) * (ex - vx) + (ey - vy) * (ey - vy) + tabs(vw) + (ex - wx) * (ex - wx) +
           (ey - wy) * (ey - wy) + tabs(ww);
  }
  return dv - dw;
}

struct Counters {
  int64_t pieces = 0;        // P: non-empty (cell, face) pieces handed to the callback
  int64_t piece_vertices = 0;  // sum n_p
  int64_t new_vertices = 0;    // sum n_new,p (vertices that are not original mesh vertices)
  int64_t laguerre_edges = 0;  // sum e_p
  int64_t sum_k = 0;           // sum over cells of k_i (neighbours used for clipping)
  int64_t sum_k_np = 0;        // sum over pieces of k_i * n_p
  int64_t fallbacks = 0;       // predicates that needed __float128 arbitration
  int64_t ties = 0;            // predicates that were ties even in __float128
};

struct Mesh {
  int nV = 0, nF = 0;
  std::vector<double> vx, vy;
  std::vector<int> tri;       // 3*nF, CCW
  std::vector<double> abc;    // 3*nF, rho_f(x,y) = a x + b y + c   (functions.hpp:73-77)
  std::vector<int> fnbr;      // 3*nF, face across the edge opposite to local vertex k, -1 if none
  double bb[4] = {0, 0, 0, 0};  // xmin, ymin, xmax, ymax
  // uniform bins of faces for the per-cell (OpenMP) enumeration
  int gbx = 0, gby = 0;
  std::vector<int> bin_ptr, bin_face;
};

struct Edge {  // Tri_intersector::Pgon_edge (voronoi_triangulation_intersection.hpp:54-96)
  int type;    // 0 = EDGE_T (mesh vertices a -> b), 1 = EDGE_DT (bisector of cell a with neighbour b)
  int a, b;
};

struct Oracle {
  Mesh mesh;
  int N = 0;
  std::vector<double> X, Y, W;
  // "regular triangulation": CSR of neighbours of every cell, CCW around the cell
  std::vector<int> nb_ptr, nb_idx;
  std::vector<char> cell_empty;  // power cell ∩ mesh box is empty ("hidden" vertex)
  // bounded sample for the timing legs of bench.py (mao_set_cell_range): only the cells [cell_lo, cell_hi)
  // of the SAME problem are evaluated (per-cell mode); cell_hi < 0 means all
  int cell_lo = 0, cell_hi = -1;
  int lo() const { return cell_hi < 0 ? 0 : std::max(0, std::min(cell_lo, N)); }
  int hi() const { return cell_hi < 0 ? N : std::max(lo(), std::min(cell_hi, N)); }
  // outputs
  double fval = 0;
  std::vector<double> g;
  std::vector<int> h_ptr, h_col;
  std::vector<double> h_val;
  std::vector<double> mom;  // N*6 moments
  Counters cnt;
  // recorded pieces (optional)
  bool record_pieces = false;
  std::vector<int> pc_cell, pc_face, pc_ptr, pc_tag;
  std::vector<double> pc_xy;

  // ------------------------------------------------------------------------------------------
  Line<double> edge_line(const Edge &e) const {  // edge_to_line, vti.hpp:98-110
    if (e.type == 0) return line_through(mesh.vx[e.a], mesh.vy[e.a], mesh.vx[e.b], mesh.vy[e.b]);
    return radical_axis(X[e.a], Y[e.a], W[e.a], X[e.b], Y[e.b], W[e.b]);
  }
  Line<quad> edge_line_q(const Edge &e) const {
    if (e.type == 0)
      return line_through<quad>(mesh.vx[e.a], mesh.vy[e.a], mesh.vx[e.b], mesh.vy[e.b]);
    return radical_axis<quad>(X[e.a], Y[e.a], W[e.a], X[e.b], Y[e.b], W[e.b]);
  }
  Pt<double> vertex_point(const Edge &a, const Edge &b) const {  // vertex_to_point, vti.hpp:112-123
    return line_line(edge_line(a), edge_line(b));
  }

  // decide sign of pow_v(E) - pow_w(E) < 0 with a double filter and a __float128 fallback; this
  // plays the role of CGAL's Filtered_predicate (predicates.hpp:159-167).  construct_q rebuilds
  // E from the original inputs in quad precision.
  template <class ConstructQ>
  bool decide(double ex, double ey, double cond, int v, int w, ConstructQ construct_q, Counters &c) const {
    double mag;
    double val = power_diff(ex, ey, X[v], Y[v], W[v], X[w], Y[w], W[w], &mag);
    double thr = (cond >= 1.0 ? 1e-13 : 1e-10 / std::max(cond, 1e-300)) * mag;
    if (std::fabs(val) > thr && std::isfinite(val)) return val < 0;
    c.fallbacks++;
    Pt<quad> E = construct_q();
    quad qmag;
    quad qv = power_diff<quad>(E.x, E.y, X[v], Y[v], W[v], X[w], Y[w], W[w], &qmag);
    if (!(tabs(qv) > quad(1e-27) * qmag)) {  // tie (or NaN): "== SMALLER" is false, App. B T3
      c.ties++;
      return false;
    }
    return qv < 0;
  }

  // Tri_intersector::inside (vti.hpp:129-166): is the polygon vertex a∩b strictly closer (in power)
  // to v than to w?
  bool inside(int v, int w, const Edge &a, const Edge &b, Counters &c) const {
    if (a.type == 0 && b.type == 0) {  // Side1 on the common mesh vertex (vti.hpp:134-139)
      int p = (a.a == b.a || a.a == b.b) ? a.a : a.b;
      double ex = mesh.vx[p], ey = mesh.vy[p];
      return decide(ex, ey, 1.0, v, w, [&]() { return Pt<quad>{quad(ex), quad(ey)}; }, c);
    }
    if (a.type == 1 && b.type == 1) {  // Side2 (vti.hpp:140-147, predicates.hpp:101-115)
      int u2 = b.b, u3 = a.b;
      double cond;
      Pt<double> E = weighted_circumcenter(X[v], Y[v], W[v], X[u2], Y[u2], W[u2], X[u3], Y[u3], W[u3], &cond);
      return decide(E.x, E.y, cond, v, w,
                    [&]() {
                      quad qc;
                      return weighted_circumcenter<quad>(X[v], Y[v], W[v], X[u2], Y[u2], W[u2], X[u3],
                                                         Y[u3], W[u3], &qc);
                    },
                    c);
    }
    // Side3 (vti.hpp:149-165, predicates.hpp:117-135): bisector(v,u) ∩ line(Sa,Sb)
    int p, q, u;
    if (a.type == 0) { p = a.a; q = a.b; u = b.b; }
    else             { p = b.a; q = b.b; u = a.b; }
    double cond;
    Line<double> L = radical_axis(X[v], Y[v], W[v], X[u], Y[u], W[u]);
    Line<double> M = line_through(mesh.vx[p], mesh.vy[p], mesh.vx[q], mesh.vy[q]);
    Pt<double> E = line_line(M, L, &cond);
    return decide(E.x, E.y, cond, v, w,
                  [&]() {
                    Line<quad> Lq = radical_axis<quad>(X[v], Y[v], W[v], X[u], Y[u], W[u]);
                    Line<quad> Mq = line_through<quad>(mesh.vx[p], mesh.vy[p], mesh.vx[q], mesh.vy[q]);
                    return line_line(Mq, Lq);
                  },
                  c);
  }

  // Tri_intersector::operator() (vti.hpp:168-198): one Sutherland–Hodgman pass on an edge list.
  void clip_by_bisector(const std::vector<Edge> &P, int v, int w, std::vector<Edge> &R, Counters &c) const {
    R.clear();
    size_t n = P.size();
    if (n == 0) return;
    Edge L{1, v, w};
    bool prev_inside = inside(v, w, P[n - 1], P[0], c);
    for (size_t i = 0; i < n; ++i) {
      size_t ii = (i + 1) % n;
      bool cur_inside = inside(v, w, P[i], P[ii], c);
      if (prev_inside) {
        R.push_back(P[i]);
        if (!cur_inside) R.push_back(L);
      } else if (cur_inside) {
        R.push_back(P[i]);
      }
      prev_inside = cur_inside;
    }
  }

  // (cell v) ∩ (face f): start from the triangle as 3 EDGE_T edges and clip by every neighbour
  // of v (vti.hpp:258-274).
  void clip_face(int v, int f, std::vector<Edge> &R, std::vector<Edge> &tmp, Counters &c) const {
    const int *t = &mesh.tri[3 * f];
    R.clear();
    R.push_back(Edge{0, t[0], t[1]});
    R.push_back(Edge{0, t[1], t[2]});
    R.push_back(Edge{0, t[2], t[0]});
    for (int k = nb_ptr[v]; k < nb_ptr[v + 1]; ++k) {
      clip_by_bisector(R, v, nb_idx[k], tmp, c);
      R.swap(tmp);
      if (R.empty()) break;  // (the reference keeps looping on an empty list: same result)
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Quadratures (quadrature.hpp)
// ---------------------------------------------------------------------------------------------
static inline double tri_area(double ax, double ay, double bx, double by, double cx, double cy) {
  return ((bx - ax) * (cy - ay) - (cx - ax) * (by - ay)) / 2;  // CGAL::area, signed
}

// integrate_albrecht_collatz (quadrature.hpp:24-42) for a K-component integrand.
template <int K, class F>
static inline void albrecht_collatz(double ax, double ay, double bx, double by, double cx, double cy, const F &f,
                                    double *acc) {
  const double _1_2 = 1.0 / 2.0, _1_6 = 1.0 / 6.0, _2_3 = 2.0 / 3.0;
  const double _1_30 = 1.0 / 30.0, _9_30 = 9.0 / 30.0;
  double ux = bx - ax, uy = by - ay, vx = cx - ax, vy = cy - ay;
  double r[6][K];
  f(ax + _1_2 * ux + _1_2 * vx, ay + _1_2 * uy + _1_2 * vy, r[0]);
  f(ax + _1_2 * ux, ay + _1_2 * uy, r[1]);
  f(ax + _1_2 * vx, ay + _1_2 * vy, r[2]);
  f(ax + _1_6 * ux + _2_3 * vx, ay + _1_6 * uy + _2_3 * vy, r[3]);
  f(ax + _1_6 * vx + _2_3 * ux, ay + _1_6 * vy + _2_3 * uy, r[4]);
  f(ax + _1_6 * ux + _1_6 * vx, ay + _1_6 * uy + _1_6 * vy, r[5]);
  double A = tri_area(ax, ay, bx, by, cx, cy);
  for (int k = 0; k < K; ++k)
    acc[k] += A * (_1_30 * r[0][k] + _1_30 * r[1][k] + _1_30 * r[2][k] + _9_30 * r[3][k] + _9_30 * r[4][k] +
                   _9_30 * r[5][k]);
}

// ---------------------------------------------------------------------------------------------
// mesh helpers
// ---------------------------------------------------------------------------------------------
static void build_face_adjacency(Mesh &m) {  // Triangulation_incremental_builder_2.h:128-151
  m.fnbr.assign(3 * (size_t)m.nF, -1);
  std::vector<std::pair<uint64_t, int>> he;
  he.reserve(3 * (size_t)m.nF);
  for (int f = 0; f < m.nF; ++f)
    for (int k = 0; k < 3; ++k) {
      int a = m.tri[3 * f + (k + 1) % 3], b = m.tri[3 * f + (k + 2) % 3];
      uint64_t key = ((uint64_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
      he.emplace_back(key, 3 * f + k);
    }
  std::sort(he.begin(), he.end());
  for (size_t i = 0; i + 1 < he.size(); ++i)
    if (he[i].first == he[i + 1].first) {
      m.fnbr[he[i].second] = he[i + 1].second / 3;
      m.fnbr[he[i + 1].second] = he[i].second / 3;
      ++i;
    }
}

static void build_face_bins(Mesh &m) {
  int g = (int)std::ceil(std::sqrt(std::max(1.0, m.nF / 2.0)));
  g = std::min(g, 4096);
  m.gbx = m.gby = g;
  double sx = g / std::max(m.bb[2] - m.bb[0], 1e-300), sy = g / std::max(m.bb[3] - m.bb[1], 1e-300);
  auto range = [&](int f, int &i0, int &i1, int &j0, int &j1) {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int k = 0; k < 3; ++k) {
      int v = m.tri[3 * f + k];
      x0 = std::min(x0, m.vx[v]); x1 = std::max(x1, m.vx[v]);
      y0 = std::min(y0, m.vy[v]); y1 = std::max(y1, m.vy[v]);
    }
    i0 = std::clamp((int)std::floor((x0 - m.bb[0]) * sx), 0, g - 1);
    i1 = std::clamp((int)std::floor((x1 - m.bb[0]) * sx), 0, g - 1);
    j0 = std::clamp((int)std::floor((y0 - m.bb[1]) * sy), 0, g - 1);
    j1 = std::clamp((int)std::floor((y1 - m.bb[1]) * sy), 0, g - 1);
  };
  m.bin_ptr.assign((size_t)g * g + 1, 0);
  for (int f = 0; f < m.nF; ++f) {
    int i0, i1, j0, j1;
    range(f, i0, i1, j0, j1);
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) m.bin_ptr[(size_t)j * g + i + 1]++;
  }
  for (size_t b = 0; b < (size_t)g * g; ++b) m.bin_ptr[b + 1] += m.bin_ptr[b];
  m.bin_face.resize(m.bin_ptr.back());
  std::vector<int> fill(m.bin_ptr.begin(), m.bin_ptr.end() - 1);
  for (int f = 0; f < m.nF; ++f) {
    int i0, i1, j0, j1;
    range(f, i0, i1, j0, j1);
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) m.bin_face[fill[(size_t)j * g + i]++] = f;
  }
}

// ---------------------------------------------------------------------------------------------
// Power-cell neighbour search: stands in for CGAL Regular_triangulation_2 (kantorovich.hpp:65-72).
// Every cell is the intersection of a bounding box with the half-planes h_ij < 0; a neighbour is a
// j whose bisector supports an edge of that polygon.  Plain fp64 clipping with constructed points;
// decisions that matter for the results are re-taken by the filtered predicates above.
// ---------------------------------------------------------------------------------------------
struct CellPoly {
  std::vector<double> x, y;
  std::vector<int> tag;  // edge k goes from vertex k to vertex k+1; tag>=0: neighbour id, <0: box side
  void init_box(const double bb[4]) {
    x = {bb[0], bb[2], bb[2], bb[0]};
    y = {bb[1], bb[1], bb[3], bb[3]};
    tag = {-1, -2, -3, -4};
  }
  double r2(double cx, double cy) const {
    double r = 0;
    for (size_t k = 0; k < x.size(); ++k) r = std::max(r, (x[k] - cx) * (x[k] - cx) + (y[k] - cy) * (y[k] - cy));
    return r;
  }
};

static Line<double> cell_edge_line(const Oracle &o, int i, int tag, const double bb[4]) {
  if (tag >= 0) return radical_axis(o.X[i], o.Y[i], o.W[i], o.X[tag], o.Y[tag], o.W[tag]);
  switch (tag) {
    case -1: return Line<double>{0, -1, bb[1]};   // y >= ymin  -> -(y) + ymin <= 0
    case -2: return Line<double>{1, 0, -bb[2]};   // x <= xmax
    case -3: return Line<double>{0, 1, -bb[3]};   // y <= ymax
    default: return Line<double>{-1, 0, bb[0]};   // x >= xmin
  }
}

// clip cell polygon of i by the half-plane of j; returns true if the polygon changed
static bool clip_cell(const Oracle &o, int i, int j, CellPoly &P, CellPoly &R, const double bb[4]) {
  size_t n = P.x.size();
  Line<double> L = radical_axis(o.X[i], o.Y[i], o.W[i], o.X[j], o.Y[j], o.W[j]);
  // L(x) = pow-diff up to sign: h_ij(x) = pow_i(x) - pow_j(x) = -(L.a x + L.b y + L.c) ... check:
  // pow_i - pow_j = -2x.(yi - yj) + |yi|^2 - |yj|^2 - wi + wj = -(L.a x + L.b y + L.c)
  std::vector<char> in(n);
  bool all_in = true, any_in = false;
  for (size_t k = 0; k < n; ++k) {
    double s = L.a * P.x[k] + L.b * P.y[k] + L.c;  // > 0  <=> strictly closer to i
    in[k] = s > 0;
    all_in &= in[k];
    any_in |= in[k];
  }
  if (all_in) return false;
  R.x.clear(); R.y.clear(); R.tag.clear();
  if (!any_in) { P.x.clear(); P.y.clear(); P.tag.clear(); return true; }
  for (size_t k = 0; k < n; ++k) {
    size_t kk = (k + 1) % n;
    bool a = in[k], b = in[kk];
    if (a) {
      R.x.push_back(P.x[k]); R.y.push_back(P.y[k]); R.tag.push_back(P.tag[k]);
      if (!b) {
        Pt<double> q = line_line(L, cell_edge_line(o, i, P.tag[k], bb));
        R.x.push_back(q.x); R.y.push_back(q.y); R.tag.push_back(j);
      }
    } else if (b) {
      Pt<double> q = line_line(L, cell_edge_line(o, i, P.tag[k], bb));
      R.x.push_back(q.x); R.y.push_back(q.y); R.tag.push_back(P.tag[k]);
    }
  }
  std::swap(P, R);
  return true;
}

// can the half-plane of a point at squared distance >= d2 with weight <= wj cut a polygon contained
// in the disk of squared radius R2 around y_i?  (SURVEY §7.2 "security radius")
static inline bool cannot_cut(double d2, double wi_minus_wj, double R2) {
  // signed distance of the bisector from y_i is t = (d^2 + wi - wj) / (2 d); for wi >= wj use the
  // weaker t >= d/2, otherwise t is increasing in d and decreasing in wj.
  double s = d2 + std::min(wi_minus_wj, 0.0);
  return s > 0 && s * s >= 4.0 * R2 * d2 * (1 + 1e-9);
}

struct PointGrid {
  int L = 0, G = 1;
  double x0 = 0, y0 = 0, inv = 1, h = 1;
  std::vector<int> start, ids;           // bins in Morton order
  std::vector<std::vector<double>> wmax;  // per level
  static uint32_t part1(uint32_t v) {
    v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v;
  }
  static uint32_t morton(uint32_t x, uint32_t y) { return part1(x) | (part1(y) << 1); }
  static uint32_t compact1(uint32_t v) {
    v &= 0x55555555; v = (v | (v >> 1)) & 0x33333333; v = (v | (v >> 2)) & 0x0f0f0f0f;
    v = (v | (v >> 4)) & 0x00ff00ff; v = (v | (v >> 8)) & 0x0000ffff; return v;
  }
  void build(const Oracle &o) {
    int N = o.N;
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = 0; i < N; ++i) {
      xmin = std::min(xmin, o.X[i]); xmax = std::max(xmax, o.X[i]);
      ymin = std::min(ymin, o.Y[i]); ymax = std::max(ymax, o.Y[i]);
    }
    double ext = std::max(std::max(xmax - xmin, ymax - ymin), 1e-300) * (1 + 1e-9);
    L = 0;
    while ((1 << L) * (1 << L) * 2 < N && L < 12) ++L;
    G = 1 << L;
    x0 = xmin; y0 = ymin; h = ext / G; inv = G / ext;
    std::vector<uint32_t> code(N);
    start.assign((size_t)G * G + 1, 0);
    for (int i = 0; i < N; ++i) {
      int bx = std::clamp((int)((o.X[i] - x0) * inv), 0, G - 1);
      int by = std::clamp((int)((o.Y[i] - y0) * inv), 0, G - 1);
      code[i] = morton(bx, by);
      start[code[i] + 1]++;
    }
    for (size_t b = 0; b < (size_t)G * G; ++b) start[b + 1] += start[b];
    ids.resize(N);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < N; ++i) ids[fill[code[i]]++] = i;
    wmax.assign(L + 1, {});
    wmax[L].assign((size_t)G * G, -std::numeric_limits<double>::infinity());
    for (size_t b = 0; b < (size_t)G * G; ++b)
      for (int k = start[b]; k < start[b + 1]; ++k) wmax[L][b] = std::max(wmax[L][b], o.W[ids[k]]);
    for (int l = L - 1; l >= 0; --l) {
      size_t n = (size_t)1 << (2 * l);
      wmax[l].resize(n);
      for (size_t c = 0; c < n; ++c)
        wmax[l][c] = std::max(std::max(wmax[l + 1][4 * c], wmax[l + 1][4 * c + 1]),
                              std::max(wmax[l + 1][4 * c + 2], wmax[l + 1][4 * c + 3]));
    }
  }
  double dist2(int l, uint32_t c, double px, double py) const {
    double s = h * (double)(1 << (L - l));
    double bx = x0 + compact1(c) * s, by = y0 + compact1(c >> 1) * s;
    double dx = std::max(std::max(bx - px, px - (bx + s)), 0.0);
    double dy = std::max(std::max(by - py, py - (by + s)), 0.0);
    return dx * dx + dy * dy;
  }
};

static void power_cell(const Oracle &o, const PointGrid *grid, int i, const double bb[4], CellPoly &P, CellPoly &R) {
  P.init_box(bb);
  double R2 = P.r2(o.X[i], o.Y[i]);
  auto try_point = [&](int j) -> bool {  // returns false when the cell became empty
    if (j == i) return true;
    double dx = o.X[j] - o.X[i], dy = o.Y[j] - o.Y[i], d2 = dx * dx + dy * dy;
    if (d2 == 0) {  // coincident sites: the heavier (then the lower index) one keeps the cell
      if (o.W[j] > o.W[i] || (o.W[j] == o.W[i] && j < i)) { P.x.clear(); P.y.clear(); P.tag.clear(); return false; }
      return true;
    }
    if (cannot_cut(d2, o.W[i] - o.W[j], R2)) return true;
    if (clip_cell(o, i, j, P, R, bb)) {
      if (P.x.empty()) return false;
      R2 = P.r2(o.X[i], o.Y[i]);
    }
    return true;
  };
  if (!grid) {
    // brute force, nearest first
    std::vector<std::pair<double, int>> order;
    order.reserve(o.N);
    for (int j = 0; j < o.N; ++j)
      if (j != i) order.emplace_back((o.X[j] - o.X[i]) * (o.X[j] - o.X[i]) + (o.Y[j] - o.Y[i]) * (o.Y[j] - o.Y[i]), j);
    std::sort(order.begin(), order.end());
    for (auto &pr : order)
      if (!try_point(pr.second)) return;
    return;
  }
  // nearest-first DFS over the quadtree of bins with max-weight pruning
  struct Item { double d2; int l; uint32_t c; };
  std::vector<Item> stack;
  stack.push_back({0.0, 0, 0u});
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    double wm = grid->wmax[it.l][it.c];
    if (wm == -std::numeric_limits<double>::infinity()) continue;
    if (it.d2 > 0 && cannot_cut(it.d2, o.W[i] - wm, R2)) continue;
    if (it.l == grid->L) {
      for (int k = grid->start[it.c]; k < grid->start[it.c + 1]; ++k)
        if (!try_point(grid->ids[k])) return;
      continue;
    }
    Item ch[4];
    for (uint32_t q = 0; q < 4; ++q) {
      ch[q].l = it.l + 1; ch[q].c = 4 * it.c + q;
      ch[q].d2 = grid->dist2(ch[q].l, ch[q].c, o.X[i], o.Y[i]);
    }
    std::sort(ch, ch + 4, [](const Item &a, const Item &b) { return a.d2 > b.d2; });
    for (auto &c : ch) stack.push_back(c);
  }
}

static void power_neighbors(Oracle &o, bool brute, int nthreads) {
  int N = o.N;
  // bounding box: the mesh box (every piece lies inside it)
  double bb[4] = {o.mesh.bb[0], o.mesh.bb[1], o.mesh.bb[2], o.mesh.bb[3]};
  PointGrid grid;
  if (!brute) grid.build(o);
  std::vector<std::vector<int>> nb(N);
  const int lo = o.lo(), hi = o.hi();
  o.cell_empty.assign(N, (lo > 0 || hi < N) ? 1 : 0);  // cells outside the sampled range are skipped downstream
#pragma omp parallel num_threads(nthreads)
  {
    CellPoly P, R;
#pragma omp for schedule(dynamic, 256)
    for (int i = lo; i < hi; ++i) {
      power_cell(o, brute ? nullptr : &grid, i, bb, P, R);
      o.cell_empty[i] = P.x.empty();
      for (int t : P.tag)
        if (t >= 0) nb[i].push_back(t);
    }
  }
  o.nb_ptr.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) o.nb_ptr[i + 1] = o.nb_ptr[i] + (int)nb[i].size();
  o.nb_idx.resize(o.nb_ptr[N]);
  for (int i = 0; i < N; ++i) std::copy(nb[i].begin(), nb[i].end(), o.nb_idx.begin() + o.nb_ptr[i]);
}

// ---------------------------------------------------------------------------------------------
// per-piece callback state
// ---------------------------------------------------------------------------------------------
struct Triplet { int r, c; double v; };

struct Accum {
  double fval = 0;
  std::vector<Triplet> htri;
  Counters cnt;
};

// kantorovich.hpp:87-136
static void piece_kantorovich(Oracle &o, const std::vector<Edge> &pg, int f, int v, double *g, Accum &A) {
  size_t n = pg.size();
  std::vector<double> px(n), py(n);
  std::vector<int> adj(n);
  for (size_t i = 0; i < n; ++i) {  // :93-102, p[i] = pgon[i] ∩ pgon[i-1]
    size_t ii = (i == 0) ? n - 1 : i - 1;
    Pt<double> p = o.vertex_point(pg[i], pg[ii]);
    px[i] = p.x; py[i] = p.y;
    adj[i] = pg[i].type == 1 ? pg[i].b : -1;
  }
  const double a = o.mesh.abc[3 * f], b = o.mesh.abc[3 * f + 1], c = o.mesh.abc[3 * f + 2];
  auto fv = [&](double x, double y) { return a * x + b * y + c; };
  double yx = o.X[v], yy = o.Y[v];
  for (size_t i = 0; i < n; ++i) {  // :110-122, p.edge(i) = (p[i], p[i+1])
    if (adj[i] < 0) continue;
    int w = adj[i];
    size_t ii = (i + 1) % n;
    double ex = px[ii] - px[i], ey = py[ii] - py[i];
    double r = std::sqrt(ex * ex + ey * ey) * fv((px[i] + px[ii]) / 2, (py[i] + py[ii]) / 2);  // quadrature.hpp:79-85
    double d = 2 * std::sqrt((yx - o.X[w]) * (yx - o.X[w]) + (yy - o.Y[w]) * (yy - o.Y[w]));
    A.htri.push_back(Triplet{v, w, -r / d});
    A.htri.push_back(Triplet{v, v, +r / d});
    A.cnt.laguerre_edges++;
  }
  double warea = 0, intg = 0;
  if (n > 2)
    for (size_t i = 1; i + 1 < n; ++i) {
      double A2 = tri_area(px[0], py[0], px[i], py[i], px[i + 1], py[i + 1]);
      warea += A2 * fv((px[0] + px[i] + px[i + 1]) / 3, (py[0] + py[i] + py[i + 1]) / 3);  // quadrature.hpp:69-77
      albrecht_collatz<1>(px[0], py[0], px[i], py[i], px[i + 1], py[i + 1],
                          [&](double x, double y, double *r) {
                            r[0] = fv(x, y) * ((x - yx) * (x - yx) + (y - yy) * (y - yy));
                          },
                          &intg);
    }
  A.fval = A.fval + warea * o.W[v] - intg;  // :132
  g[v] = g[v] + warea;                      // :133
}

// lloyd.hpp:48-68 (order 1) and :91-122 (order 2); the polygon wrapper uses p[i] = R[i] ∩ R[i+1]
// (vti.hpp:336-340).
static void piece_moments(Oracle &o, const std::vector<Edge> &pg, int f, int v, int order, double *mom) {
  size_t n = pg.size();
  std::vector<double> px(n), py(n);
  for (size_t i = 0; i < n; ++i) {
    Pt<double> p = o.vertex_point(pg[i], pg[(i + 1) % n]);
    px[i] = p.x; py[i] = p.y;
  }
  const double a = o.mesh.abc[3 * f], b = o.mesh.abc[3 * f + 1], c = o.mesh.abc[3 * f + 2];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (n > 2)
    for (size_t i = 1; i + 1 < n; ++i) {
      if (order == 1)
        albrecht_collatz<3>(px[0], py[0], px[i], py[i], px[i + 1], py[i + 1],
                            [&](double x, double y, double *r) {
                              double fp = a * x + b * y + c;
                              r[0] = fp; r[1] = fp * x; r[2] = fp * y;
                            },
                            acc);
      else
        albrecht_collatz<6>(px[0], py[0], px[i], py[i], px[i + 1], py[i + 1],
                            [&](double x, double y, double *r) {
                              double fp = a * x + b * y + c;
                              r[0] = fp; r[1] = fp * x; r[2] = fp * y;
                              r[3] = fp * x * x; r[4] = fp * y * y; r[5] = fp * x * y;
                            },
                            acc);
    }
  for (int k = 0; k < 6; ++k) mom[6 * (size_t)v + k] += acc[k];
}

static void count_piece(const Oracle &o, const std::vector<Edge> &pg, int v, Counters &c) {
  c.pieces++;
  c.piece_vertices += (int64_t)pg.size();
  size_t n = pg.size();
  for (size_t i = 0; i < n; ++i)
    if (!(pg[i].type == 0 && pg[(i + 1) % n].type == 0)) c.new_vertices++;
  c.sum_k_np += (int64_t)(o.nb_ptr[v + 1] - o.nb_ptr[v]) * (int64_t)n;
}

static void record_piece(Oracle &o, const std::vector<Edge> &pg, int f, int v) {
  if (o.pc_ptr.empty()) o.pc_ptr.push_back(0);
  o.pc_cell.push_back(v);
  o.pc_face.push_back(f);
  size_t n = pg.size();
  for (size_t i = 0; i < n; ++i) {
    size_t ii = (i == 0) ? n - 1 : i - 1;
    Pt<double> p = o.vertex_point(pg[i], pg[ii]);
    o.pc_xy.push_back(p.x);
    o.pc_xy.push_back(p.y);
    o.pc_tag.push_back(pg[i].type == 1 ? pg[i].b : -1);
  }
  o.pc_ptr.push_back((int)o.pc_tag.size());
}

// voronoi_triangulation_intersection_raw (vti.hpp:219-313): global BFS over (cell, face) pairs from
// one seed.  out(R, f, v) is called for every visited pair, including pairs whose clipped polygon
// came out empty (the reference does the same; its callbacks are no-ops on empty polygons).
template <class Out> static void overlay_bfs(Oracle &o, Out out, Counters &c) {
  const Mesh &m = o.mesh;
  if (m.nV == 0 || m.nF == 0 || o.N == 0) return;
  int f0 = 0;  // t.finite_faces_begin()
  // nearest_power_vertex(f->vertex(0)->point()) (vti.hpp:210-216, :238)
  int v0 = -1;
  {
    double ex = m.vx[m.tri[0]], ey = m.vy[m.tri[0]], best = 0;
    for (int j = 0; j < o.N; ++j) {
      double pj = (ex - o.X[j]) * (ex - o.X[j]) + (ey - o.Y[j]) * (ey - o.Y[j]) - o.W[j];
      if (v0 < 0 || pj < best) { best = pj; v0 = j; }
    }
    if (v0 < 0) v0 = 0;
  }
  typedef std::pair<int, int> VF;
  std::priority_queue<VF> Q;
  std::set<VF> visited;
  Q.push(VF(v0, f0));
  visited.insert(VF(v0, f0));
  std::vector<Edge> R, tmp;
  while (!Q.empty()) {
    VF vf = Q.top();
    Q.pop();
    int v = vf.first, f = vf.second;
    o.clip_face(v, f, R, tmp, c);
    for (const Edge &e : R) {  // propagate (vti.hpp:277-303)
      VF p;
      if (e.type == 0) {
        int i = -1, j = -1;
        for (int k = 0; k < 3; ++k) {
          if (m.tri[3 * f + k] == e.a) i = k;
          if (m.tri[3 * f + k] == e.b) j = k;
        }
        int k = 3 - i - j;
        int fn = m.fnbr[3 * f + k];
        if (fn < 0) continue;
        p = VF(v, fn);
      } else {
        p = VF(e.b, f);
      }
      if (visited.find(p) != visited.end()) continue;
      visited.insert(p);
      Q.push(p);
    }
    out(R, f, v);
  }
}

// Per-cell enumeration (OpenMP variant): every face whose bin overlaps the cell's bounding box is
// clipped; faces are de-duplicated by handling a face only in the first overlapped bin.
template <class Out> static void overlay_cells(Oracle &o, int nthreads, Out out, std::vector<Counters> &cs) {
  const Mesh &m = o.mesh;
  int g = m.gbx;
  double sx = g / std::max(m.bb[2] - m.bb[0], 1e-300), sy = g / std::max(m.bb[3] - m.bb[1], 1e-300);
  cs.assign(nthreads, Counters());
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    std::vector<Edge> R, tmp;
    CellPoly P, Q;
    const int lo = o.lo(), hi = o.hi();
#pragma omp for schedule(dynamic, 64)
    for (int v = lo; v < hi; ++v) {
      if (o.cell_empty[v]) continue;
      // cell polygon in the mesh box from its neighbour list -> bounding box of the cell
      P.init_box(m.bb);
      for (int k = o.nb_ptr[v]; k < o.nb_ptr[v + 1] && !P.x.empty(); ++k) clip_cell(o, v, o.nb_idx[k], P, Q, m.bb);
      if (P.x.empty()) continue;
      double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
      for (size_t k = 0; k < P.x.size(); ++k) {
        x0 = std::min(x0, P.x[k]); x1 = std::max(x1, P.x[k]);
        y0 = std::min(y0, P.y[k]); y1 = std::max(y1, P.y[k]);
      }
      double pad = 1e-9 * std::max(m.bb[2] - m.bb[0], m.bb[3] - m.bb[1]);
      int i0 = std::clamp((int)std::floor((x0 - pad - m.bb[0]) * sx), 0, g - 1);
      int i1 = std::clamp((int)std::floor((x1 + pad - m.bb[0]) * sx), 0, g - 1);
      int j0 = std::clamp((int)std::floor((y0 - pad - m.bb[1]) * sy), 0, g - 1);
      int j1 = std::clamp((int)std::floor((y1 + pad - m.bb[1]) * sy), 0, g - 1);
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i)
          for (int q = m.bin_ptr[(size_t)j * g + i]; q < m.bin_ptr[(size_t)j * g + i + 1]; ++q) {
            int f = m.bin_face[q];
            // first overlapped bin of face f inside the query window?
            double fx0 = 1e300, fy0 = 1e300;
            for (int k = 0; k < 3; ++k) {
              fx0 = std::min(fx0, m.vx[m.tri[3 * f + k]]);
              fy0 = std::min(fy0, m.vy[m.tri[3 * f + k]]);
            }
            int fi0 = std::max(std::clamp((int)std::floor((fx0 - m.bb[0]) * sx), 0, g - 1), i0);
            int fj0 = std::max(std::clamp((int)std::floor((fy0 - m.bb[1]) * sy), 0, g - 1), j0);
            if (fi0 != i || fj0 != j) continue;
            o.clip_face(v, f, R, tmp, cs[tid]);
            if (!R.empty()) out(R, f, v, tid);
          }
    }
  }
}

static void add_counters(Counters &a, const Counters &b) {
  a.pieces += b.pieces; a.piece_vertices += b.piece_vertices; a.new_vertices += b.new_vertices;
  a.laguerre_edges += b.laguerre_edges; a.sum_k += b.sum_k; a.sum_k_np += b.sum_k_np;
  a.fallbacks += b.fallbacks; a.ties += b.ties;
}

// Eigen setFromTriplets + makeCompressed (kantorovich.hpp:137-139): duplicates summed; we emit the
// row-major (CSR) form of the same matrix, columns ascending.
static void assemble(Oracle &o, std::vector<Triplet> &t) {
  std::stable_sort(t.begin(), t.end(), [](const Triplet &a, const Triplet &b) {
    return a.r != b.r ? a.r < b.r : a.c < b.c;
  });
  o.h_ptr.assign(o.N + 1, 0);
  o.h_col.clear();
  o.h_val.clear();
  for (size_t k = 0; k < t.size();) {
    size_t e = k;
    double s = 0;
    while (e < t.size() && t[e].r == t[k].r && t[e].c == t[k].c) s += t[e++].v;
    o.h_col.push_back(t[k].c);
    o.h_val.push_back(s);
    o.h_ptr[t[k].r + 1]++;
    k = e;
  }
  for (int i = 0; i < o.N; ++i) o.h_ptr[i + 1] += o.h_ptr[i];
}

}  // namespace

// =============================================================================================
// C interface (ctypes)
// =============================================================================================
extern "C" {

void *mao_create() { return new Oracle(); }
void mao_destroy(void *h) { delete (Oracle *)h; }

// Linear_function(p,fp,q,fq,r,fr) (functions.hpp:25-53,64-71): barycentric extrapolation to
// (0,0), (1,0), (0,1).
void mao_linear_functions(int nV, const double *vx, const double *vy, const double *rho, int nF, const int *tri,
                          double *abc) {
  (void)nV;
  for (int f = 0; f < nF; ++f) {
    int ia = tri[3 * f], ib = tri[3 * f + 1], ic = tri[3 * f + 2];
    double ax = vx[ia], ay = vy[ia], bx = vx[ib], by = vy[ib], cx = vx[ic], cy = vy[ic];
    auto extrapolate = [&](double px, double py) {
      double v0x = bx - ax, v0y = by - ay, v1x = cx - ax, v1y = cy - ay, v2x = px - ax, v2y = py - ay;
      double d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
      double d20 = v2x * v0x + v2y * v0y, d21 = v2x * v1x + v2y * v1y;
      double denom = d00 * d11 - d01 * d01;
      double v = (d11 * d20 - d01 * d21) / denom, w = (d00 * d21 - d01 * d20) / denom, u = 1.0 - v - w;
      return u * rho[ia] + v * rho[ib] + w * rho[ic];
    };
    double c = extrapolate(0, 0);
    abc[3 * f] = extrapolate(1, 0) - c;
    abc[3 * f + 1] = extrapolate(0, 1) - c;
    abc[3 * f + 2] = c;
  }
}

int mao_set_mesh(void *h, int nV, const double *vx, const double *vy, int nF, const int *tri, const double *abc) {
  Oracle &o = *(Oracle *)h;
  Mesh &m = o.mesh;
  m.nV = nV; m.nF = nF;
  m.vx.assign(vx, vx + nV); m.vy.assign(vy, vy + nV);
  m.tri.assign(tri, tri + 3 * (size_t)nF);
  m.abc.assign(abc, abc + 3 * (size_t)nF);
  m.bb[0] = m.bb[1] = 1e300; m.bb[2] = m.bb[3] = -1e300;
  for (int i = 0; i < nV; ++i) {
    m.bb[0] = std::min(m.bb[0], vx[i]); m.bb[2] = std::max(m.bb[2], vx[i]);
    m.bb[1] = std::min(m.bb[1], vy[i]); m.bb[3] = std::max(m.bb[3], vy[i]);
  }
  for (int f = 0; f < nF; ++f)  // CCW check
    if (tri_area(vx[tri[3 * f]], vy[tri[3 * f]], vx[tri[3 * f + 1]], vy[tri[3 * f + 1]], vx[tri[3 * f + 2]],
                 vy[tri[3 * f + 2]]) <= 0)
      return -1;
  build_face_adjacency(m);
  build_face_bins(m);
  return 0;
}

// Timing legs only: restrict the per-cell mode (mode bit1) to the cells [lo, hi) of the same problem; hi < 0 = all.
void mao_set_cell_range(void *h, int lo, int hi) {
  Oracle &o = *(Oracle *)h;
  o.cell_lo = lo; o.cell_hi = hi;
}

int mao_set_points(void *h, int N, const double *x, const double *y) {
  Oracle &o = *(Oracle *)h;
  o.N = N;
  o.X.assign(x, x + N); o.Y.assign(y, y + N);
  o.W.assign(N, 0.0);
  return 0;
}

// mode bit0: 1 = brute-force neighbour search; bit1: 1 = per-cell OpenMP enumeration instead of the
// reference's global BFS; bit2: record pieces.
int mao_kantorovich(void *h, const double *w, int mode, int nthreads) {
  Oracle &o = *(Oracle *)h;
  if (nthreads < 1) nthreads = 1;
  o.W.assign(w, w + o.N);
  o.cnt = Counters();
  o.record_pieces = mode & 4;
  o.pc_cell.clear(); o.pc_face.clear(); o.pc_ptr.clear(); o.pc_tag.clear(); o.pc_xy.clear();
  power_neighbors(o, mode & 1, nthreads);  // RT dt(Xw.begin(), Xw.end())  kantorovich.hpp:71
  o.cnt.sum_k = o.nb_ptr[o.N];
  o.g.assign(o.N, 0.0);  // :83
  std::vector<Triplet> htri;
  if (!(mode & 2)) {
    Accum A;
    overlay_bfs(o, [&](const std::vector<Edge> &R, int f, int v) {
      if (R.empty()) return;
      count_piece(o, R, v, A.cnt);
      if (o.record_pieces) record_piece(o, R, f, v);
      piece_kantorovich(o, R, f, v, o.g.data(), A);
    }, o.cnt);
    o.fval = A.fval;
    add_counters(o.cnt, A.cnt);
    htri.swap(A.htri);
  } else {
    std::vector<Accum> acc(nthreads);
    std::vector<Counters> cs;
    overlay_cells(o, nthreads, [&](const std::vector<Edge> &R, int f, int v, int tid) {
      count_piece(o, R, v, acc[tid].cnt);
      piece_kantorovich(o, R, f, v, o.g.data(), acc[tid]);  // g[v]: v is owned by one thread
    }, cs);
    o.fval = 0;
    for (int t = 0; t < nthreads; ++t) {
      o.fval += acc[t].fval;
      add_counters(o.cnt, acc[t].cnt);
      add_counters(o.cnt, cs[t]);
      htri.insert(htri.end(), acc[t].htri.begin(), acc[t].htri.end());
    }
  }
  assemble(o, htri);  // :137-139
  return 0;
}

// first_moment / second_moment (lloyd.hpp:30-123).  mom is N x 6 row-major:
// (mass, ∫ρx, ∫ρy, ∫ρx², ∫ρy², ∫ρxy); order-1 leaves the last three zero.
int mao_moments(void *h, const double *w, int order, int mode, int nthreads, double *mom) {
  Oracle &o = *(Oracle *)h;
  if (nthreads < 1) nthreads = 1;
  o.W.assign(w, w + o.N);
  o.cnt = Counters();
  power_neighbors(o, mode & 1, nthreads);  // details::make_regular_triangulation, common_rt.hpp:34-53
  std::fill(mom, mom + 6 * (size_t)o.N, 0.0);
  if (!(mode & 2)) {
    overlay_bfs(o, [&](const std::vector<Edge> &R, int f, int v) {
      if (!R.empty()) piece_moments(o, R, f, v, order, mom);
    }, o.cnt);
  } else {
    std::vector<Counters> cs;
    overlay_cells(o, nthreads, [&](const std::vector<Edge> &R, int f, int v, int) { piece_moments(o, R, f, v, order, mom); }, cs);
  }
  return 0;
}

double mao_fval(void *h) { return ((Oracle *)h)->fval; }
void mao_get_g(void *h, double *g) { Oracle &o = *(Oracle *)h; std::copy(o.g.begin(), o.g.end(), g); }
int mao_nnz(void *h) { return (int)((Oracle *)h)->h_col.size(); }
void mao_get_csr(void *h, int *ptr, int *col, double *val) {
  Oracle &o = *(Oracle *)h;
  std::copy(o.h_ptr.begin(), o.h_ptr.end(), ptr);
  std::copy(o.h_col.begin(), o.h_col.end(), col);
  std::copy(o.h_val.begin(), o.h_val.end(), val);
}
int mao_num_neighbors(void *h) { return (int)((Oracle *)h)->nb_idx.size(); }
void mao_get_neighbors(void *h, int *ptr, int *idx) {
  Oracle &o = *(Oracle *)h;
  std::copy(o.nb_ptr.begin(), o.nb_ptr.end(), ptr);
  std::copy(o.nb_idx.begin(), o.nb_idx.end(), idx);
}
void mao_get_counters(void *h, int64_t *c) {
  const Counters &k = ((Oracle *)h)->cnt;
  c[0] = k.pieces; c[1] = k.piece_vertices; c[2] = k.new_vertices; c[3] = k.laguerre_edges;
  c[4] = k.sum_k; c[5] = k.sum_k_np; c[6] = k.fallbacks; c[7] = k.ties;
}
int mao_num_pieces(void *h) { return (int)((Oracle *)h)->pc_cell.size(); }
int mao_num_piece_vertices(void *h) { return (int)((Oracle *)h)->pc_tag.size(); }
void mao_get_pieces(void *h, int *cell, int *face, int *ptr, int *tag, double *xy) {
  Oracle &o = *(Oracle *)h;
  std::copy(o.pc_cell.begin(), o.pc_cell.end(), cell);
  std::copy(o.pc_face.begin(), o.pc_face.end(), face);
  std::copy(o.pc_ptr.begin(), o.pc_ptr.end(), ptr);
  std::copy(o.pc_tag.begin(), o.pc_tag.end(), tag);
  std::copy(o.pc_xy.begin(), o.pc_xy.end(), xy);
}

// solve_laplacian_matrix (optimal_transport.hpp:41-87): ground the LAST index, solve the leading
// (N-1)x(N-1) block, d[N-1] = 0.  Eigen::SimplicialLLT is an un-vendored dependency (Eigen 3.2.1);
// the SPD system has a unique solution, obtained here by Jacobi-preconditioned CG run to a
// 1e-14 relative residual (tests also cross-check with SciPy's direct SuperLU factorisation).
// Returns the iteration count (negative: singular diagonal, :49-58).
int mao_solve_laplacian(int N, const int *ptr, const int *col, const double *val, const double *g, double *d,
                        double rtol, int maxit) {
  int n = N - 1;
  std::vector<double> diag(n, 0.0), r(n), z(n), p(n), q(n), x(n, 0.0);
  bool singular = false;
  for (int i = 0; i < N; ++i) {
    double dd = 0;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k)
      if (col[k] == i) dd = val[k];
    if (dd == 0) singular = true;
    if (i < n) diag[i] = dd;
  }
  if (singular) { std::fill(d, d + N, 0.0); return -1; }
  double gn = 0;
  for (int i = 0; i < n; ++i) { r[i] = g[i]; gn += g[i] * g[i]; }
  gn = std::sqrt(gn);
  d[N - 1] = 0;
  if (gn == 0) { std::fill(d, d + N, 0.0); return 0; }
  double rz = 0;
  for (int i = 0; i < n; ++i) { z[i] = r[i] / diag[i]; p[i] = z[i]; rz += r[i] * z[i]; }
  int it = 0;
  for (; it < maxit; ++it) {
    double pq = 0;
#pragma omp parallel for reduction(+ : pq) if (n > 50000)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = ptr[i]; k < ptr[i + 1]; ++k)
        if (col[k] < n) s += val[k] * p[col[k]];
      q[i] = s;
      pq += p[i] * s;
    }
    double alpha = rz / pq, rn = 0, rz2 = 0;
    for (int i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      z[i] = r[i] / diag[i];
      rn += r[i] * r[i];
      rz2 += r[i] * z[i];
    }
    if (std::sqrt(rn) <= rtol * gn) { ++it; break; }
    double beta = rz2 / rz;
    rz = rz2;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  for (int i = 0; i < n; ++i) d[i] = x[i];
  return it;
}

}  // extern "C"

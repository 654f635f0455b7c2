// MA/voronoi_polygon_intersection.hpp — drop-in for the reference's
// include/MA/voronoi_polygon_intersection.hpp:153-188:
//   Polygon MA::voronoi_polygon_intersection(P, dt, v)
// = (Voronoi / Laguerre cell of vertex v of dt) ∩ P for a simple polygon P given counter-clockwise, convex or not
// (tests/test_voronoi.cpp:41-48 and tests/test_power.cpp:43-50 call it for every vertex with P = the unit square and
// sum the areas; tests/test_voronoi_ad.cpp uses a non-convex cross).  dt may be weighted (regular triangulation) or
// not (Delaunay: the overloads of predicates.hpp:38-44,63-70,89-98 — the weight is simply absent).  The cells come
// from the GPU neighbour search (ma_cells_build / ma_cells_get: cell ∩ bounding box of P); as the reference's
// Pgon_intersector does (voronoi_polygon_intersection.hpp:27-151), P is then clipped by the half-planes of the cell, one
// Sutherland–Hodgman pass per cell edge on the host; for a non-convex P the result may contain zero-width bridges,
// its area is exact.  The result has the caller's Polygon type (:153-158).  All cells of one (P, dt) pair are computed
// by the first call and cached, so the driver's loop over the vertices costs one GPU evaluation.
#ifndef MA_VORONOI_POLYGON_INTERSECTION_HPP
#define MA_VORONOI_POLYGON_INTERSECTION_HPP

#include "b200_bridge.hpp"
#include "lite.hpp"
#include "voronoi_triangulation_intersection.hpp"  // details::weight_of

namespace MA {
namespace details {
struct CellCache {
  uint64_t key = 0;
  std::vector<int> ptr, tag;
  std::vector<double> xy;
  std::map<const void *, size_t> index;  // address of a dt vertex -> its rank
};
inline CellCache &cell_cache() {
  static thread_local CellCache c;
  return c;
}
}  // namespace details

template <class Polygon, class DT, class VH> Polygon voronoi_polygon_intersection(const Polygon &P, const DT &dt, const VH &v) {
  typedef decltype(dt.finite_vertices_begin()) VIt;
  details::CellCache &cc = details::cell_cache();
  // fingerprint of (P, dt): every coordinate and weight
  std::vector<double> x, y, w, px, py;
  std::vector<const void *> addr;
  for (VIt it = dt.finite_vertices_begin(); it != dt.finite_vertices_end(); ++it) {
    x.push_back(it->point().x()); y.push_back(it->point().y()); w.push_back(details::weight_of(it->point(), 0));
    addr.push_back((const void *)&*it);
  }
  for (size_t k = 0; k < P.size(); ++k) { px.push_back(P[k].x()); py.push_back(P[k].y()); }
  uint64_t key = b200::fnv(x.data(), x.size() * 8);
  key = b200::fnv(y.data(), y.size() * 8, key); key = b200::fnv(w.data(), w.size() * 8, key);
  key = b200::fnv(px.data(), px.size() * 8, key); key = b200::fnv(py.data(), py.size() * 8, key);
  if (key != cc.key) {
    // bounding box of P as a 2-triangle mesh (only its box matters for the cells)
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (size_t k = 0; k < px.size(); ++k) { x0 = std::min(x0, px[k]); x1 = std::max(x1, px[k]); y0 = std::min(y0, py[k]); y1 = std::max(y1, py[k]); }
    const double vx[4] = {x0, x1, x1, x0}, vy[4] = {y0, y0, y1, y1}, abc[6] = {0, 0, 1, 0, 0, 1};
    const int tri[6] = {0, 1, 2, 0, 2, 3};
    b200::Engine &E = b200::Engine::instance();
    ma_ctx *c = E.get();
    E.invalidate();
    b200::check(c, ma_set_mesh(c, 4, vx, vy, 2, tri, abc), "ma_set_mesh");
    b200::check(c, ma_set_points(c, (int)x.size(), x.data(), y.data()), "ma_set_points");
    int nv = 0;
    b200::check(c, ma_cells_build(c, w.data(), &nv), "ma_cells_build");
    cc.ptr.assign(x.size() + 1, 0); cc.xy.assign(2 * (size_t)std::max(nv, 1), 0.0); cc.tag.assign(std::max(nv, 1), 0);
    b200::check(c, ma_cells_get(c, cc.ptr.data(), cc.xy.data(), cc.tag.data()), "ma_cells_get");
    cc.index.clear();
    for (size_t k = 0; k < addr.size(); ++k) cc.index[addr[k]] = k;
    cc.key = key;
  }
  std::map<const void *, size_t>::const_iterator it = cc.index.find((const void *)&*v);
  if (it == cc.index.end()) throw std::runtime_error("MA::voronoi_polygon_intersection: v is not a vertex of dt");
  // the cell (convex, counter-clockwise, already inside the bounding box of P) ...
  std::vector<double> cx, cy;
  for (int k = cc.ptr[it->second]; k < cc.ptr[it->second + 1]; ++k) { cx.push_back(cc.xy[2 * (size_t)k]); cy.push_back(cc.xy[2 * (size_t)k + 1]); }
  // ... clips P: one pass per cell edge, inside = left of the directed edge
  std::vector<double> qx(px), qy(py);
  if (cx.size() < 3) { qx.clear(); qy.clear(); }
  for (size_t s = 0; s < cx.size() && !qx.empty(); ++s) {
    const double ax = cx[s], ay = cy[s], bx = cx[(s + 1) % cx.size()], by = cy[(s + 1) % cx.size()];
    std::vector<double> ox, oy;
    const size_t n = qx.size();
    for (size_t k = 0; k < n; ++k) {
      const double ux = qx[k], uy = qy[k], vx = qx[(k + 1) % n], vy = qy[(k + 1) % n];
      const double sc = (bx - ax) * (uy - ay) - (by - ay) * (ux - ax), sd = (bx - ax) * (vy - ay) - (by - ay) * (vx - ax);
      if (sc >= 0) { ox.push_back(ux); oy.push_back(uy); }
      if ((sc >= 0) != (sd >= 0)) { const double t = sc / (sc - sd); ox.push_back(ux + t * (vx - ux)); oy.push_back(uy + t * (vy - uy)); }
    }
    qx.swap(ox); qy.swap(oy);
  }
  typedef typename std::decay<decltype(P[0])>::type PointT;
  Polygon R;
  for (size_t k = 0; k < qx.size(); ++k) R.push_back(PointT(qx[k], qy[k]));
  return R;
}

}  // namespace MA
#endif

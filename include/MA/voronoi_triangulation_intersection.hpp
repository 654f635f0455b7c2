// MA/voronoi_triangulation_intersection.hpp — drop-in for the point-polygon wrapper of the
// reference's include/MA/voronoi_triangulation_intersection.hpp:315-343:
//   MA::voronoi_triangulation_intersection(t, dt, out)   with   out(Polygon, Face_handle, Vertex_handle)
// called once per non-empty piece (Laguerre cell of a vertex of dt) ∩ (face of t).  dt is only read
// through finite_vertices_begin/end and ->point() (.x(), .y(), .weight() when it has one): a CGAL
// Regular/Delaunay triangulation or MA::lite::Weighted_sites.  The pieces are computed on the GPU
// (ma_pieces_build / ma_pieces_get) and replayed on the host in (cell, face) order; the reference's
// order is that of a pointer-keyed priority queue and is not reproducible anyway.
#ifndef MA_VORONOI_TRIANGULATION_INTERSECTION_HPP
#define MA_VORONOI_TRIANGULATION_INTERSECTION_HPP

#include "b200_bridge.hpp"
#include "lite.hpp"

namespace MA {
namespace details {
template <class P> auto weight_of(const P &p, int) -> decltype(p.weight()) { return p.weight(); }
template <class P> double weight_of(const P &, long) { return 0.0; }  // unweighted (Delaunay / Voronoi)
}  // namespace details

template <class T, class DT, class F> void voronoi_triangulation_intersection(const T &t, const DT &dt, F out) {
  typedef decltype(dt.finite_vertices_begin()) VIt;
  typedef decltype(t.finite_faces_begin()) FIt;
  // uniform density: only the geometry of the pieces matters here
  std::vector<VIt> vh;
  std::vector<double> x, y, w;
  for (VIt v = dt.finite_vertices_begin(); v != dt.finite_vertices_end(); ++v) {
    vh.push_back(v);
    x.push_back(v->point().x()); y.push_back(v->point().y());
    w.push_back(details::weight_of(v->point(), 0));
  }
  std::vector<FIt> fh;
  std::vector<double> vx, vy, abc;
  std::vector<int> tri;
  {
    typedef decltype(t.finite_faces_begin()->vertex(0)) VH;
    std::map<VH, int> idx;
    for (FIt f = t.finite_faces_begin(); f != t.finite_faces_end(); ++f) {
      fh.push_back(f);
      abc.push_back(0); abc.push_back(0); abc.push_back(1);
      for (int k = 0; k < 3; ++k) {
        VH v = f->vertex(k);
        typename std::map<VH, int>::iterator it = idx.find(v);
        if (it == idx.end()) {
          it = idx.insert(std::make_pair(v, (int)vx.size())).first;
          vx.push_back(v->point().x()); vy.push_back(v->point().y());
        }
        tri.push_back(it->second);
      }
    }
  }
  b200::Engine &E = b200::Engine::instance();
  ma_ctx *c = E.get();
  E.invalidate();  // this call installs its own mesh / points
  b200::check(c, ma_set_mesh(c, (int)vx.size(), vx.data(), vy.data(), (int)fh.size(), tri.data(), abc.data()), "ma_set_mesh");
  b200::check(c, ma_set_points(c, (int)x.size(), x.data(), y.data()), "ma_set_points");
  int np = 0, nv = 0;
  b200::check(c, ma_pieces_build(c, w.data(), &np, &nv), "ma_pieces_build");
  std::vector<int> cell(np ? np : 1), face(np ? np : 1), ptr(np + 1), tag(nv ? nv : 1);
  std::vector<double> xy(nv ? 2 * (size_t)nv : 2);
  b200::check(c, ma_pieces_get(c, cell.data(), face.data(), ptr.data(), tag.data(), xy.data()), "ma_pieces_get");
  for (int p = 0; p < np; ++p) {
    lite::Polygon poly;
    for (int k = ptr[p]; k < ptr[p + 1]; ++k) poly.push_back(lite::Point(xy[2 * (size_t)k], xy[2 * (size_t)k + 1]));
    out(poly, fh[face[p]], vh[cell[p]]);
  }
}

}  // namespace MA
#endif

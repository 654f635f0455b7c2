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

// The symbolic polygons of the raw traversal (vti.hpp:54-96): a cyclic list of EDGES, each either an edge of the
// source triangle (EDGE_T: its two mesh vertices) or a piece of the bisector between the cell's site and a neighbour
// (EDGE_DT: (site, neighbour)); vertex i of the polygon is R[i] ∩ R[i+1] (vti.hpp:336-340).
template <class Vertex_handle_T, class Vertex_handle_DT> struct Pgon_edge_t {
  enum EdgeType { EDGE_T, EDGE_DT };
  EdgeType type;
  std::pair<Vertex_handle_T, Vertex_handle_T> edge_t;
  std::pair<Vertex_handle_DT, Vertex_handle_DT> edge_dt;
};

namespace details {
// pieces of (t, dt) from the engine; fh / vh receive the face / vertex handles the indices refer to
template <class T, class DT, class FIt, class VIt>
void gpu_pieces(const T &t, const DT &dt, std::vector<FIt> &fh, std::vector<VIt> &vh, std::vector<int> &cell,
                std::vector<int> &face, std::vector<int> &ptr, std::vector<int> &tag, std::vector<double> &xy);
}  // namespace details

template <class T, class DT, class F> void voronoi_triangulation_intersection(const T &t, const DT &dt, F out) {
  typedef decltype(dt.finite_vertices_begin()) VIt;
  typedef decltype(t.finite_faces_begin()) FIt;
  std::vector<VIt> vh;
  std::vector<FIt> fh;
  std::vector<int> cell, face, ptr, tag;
  std::vector<double> xy;
  details::gpu_pieces(t, dt, fh, vh, cell, face, ptr, tag, xy);
  for (size_t p = 0; p < cell.size(); ++p) {
    lite::Polygon poly;
    for (int k = ptr[p]; k < ptr[p + 1]; ++k) poly.push_back(lite::Point(xy[2 * (size_t)k], xy[2 * (size_t)k + 1]));
    out(poly, fh[face[p]], vh[cell[p]]);
  }
}

// voronoi_triangulation_intersection_raw(t, dt, out), vti.hpp:219-313: out(R, f, v) with R the symbolic polygon of
// the piece (cell of v) ∩ (face f).  Same pieces as above; the edge list is rotated so that R[i] ∩ R[i+1] is vertex i
// of the point polygon the wrapper hands out (kantorovich.hpp:95-102 relies on that indexing).
template <class T, class DT, class F> void voronoi_triangulation_intersection_raw(const T &t, const DT &dt, F out) {
  typedef decltype(dt.finite_vertices_begin()) VIt;
  typedef decltype(t.finite_faces_begin()) FIt;
  typedef decltype(t.finite_faces_begin()->vertex(0)) VHT;
  typedef Pgon_edge_t<VHT, VIt> Edge;
  std::vector<VIt> vh;
  std::vector<FIt> fh;
  std::vector<int> cell, face, ptr, tag;
  std::vector<double> xy;
  details::gpu_pieces(t, dt, fh, vh, cell, face, ptr, tag, xy);
  for (size_t p = 0; p < cell.size(); ++p) {
    const int n = ptr[p + 1] - ptr[p];
    std::vector<Edge> R(n);
    for (int i = 0; i < n; ++i) {
      const int k = ptr[p] + (i + n - 1) % n;  // R[i] = the edge that ENDS at vertex i
      Edge e;
      if (tag[k] >= 0) {
        e.type = Edge::EDGE_DT;
        e.edge_dt = std::make_pair(vh[cell[p]], vh[tag[k]]);
      } else {
        const int a = -1 - tag[k];
        e.type = Edge::EDGE_T;
        e.edge_t = std::make_pair(fh[face[p]]->vertex(a), fh[face[p]]->vertex((a + 1) % 3));
      }
      R[i] = e;
    }
    out(R, fh[face[p]], vh[cell[p]]);
  }
}

namespace details {
template <class T, class DT, class FIt, class VIt>
void gpu_pieces(const T &t, const DT &dt, std::vector<FIt> &fh, std::vector<VIt> &vh, std::vector<int> &cell,
                std::vector<int> &face, std::vector<int> &ptr, std::vector<int> &tag, std::vector<double> &xy) {
  // uniform density: only the geometry of the pieces matters here
  std::vector<double> x, y, w;
  for (VIt v = dt.finite_vertices_begin(); v != dt.finite_vertices_end(); ++v) {
    vh.push_back(v);
    x.push_back(v->point().x()); y.push_back(v->point().y());
    w.push_back(details::weight_of(v->point(), 0));
  }
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
  cell.assign(np, 0); face.assign(np, 0); ptr.assign(np + 1, 0); tag.assign(nv ? nv : 1, 0);
  xy.assign(nv ? 2 * (size_t)nv : 2, 0.0);
  std::vector<int> cbuf(np ? np : 1), fbuf(np ? np : 1);
  b200::check(c, ma_pieces_get(c, cbuf.data(), fbuf.data(), ptr.data(), tag.data(), xy.data()), "ma_pieces_get");
  for (int p = 0; p < np; ++p) { cell[p] = cbuf[p]; face[p] = fbuf[p]; }
}
}  // namespace details

}  // namespace MA
#endif

// MA/functions.hpp — drop-in for the reference's include/MA/functions.hpp:
//   MA::Linear_function<K>            (functions.hpp:55-80 there)
//   MA::image_to_pl_function(image, t, fs)  (functions.hpp:82-120 there)
// The per-face linear functions stay host-side objects (they are inputs of the engine, not part of
// the hot path); the triangulation image_to_pl_function builds uses the fixed diagonal
// (i,j)-(i+1,j+1) per pixel square (the reference lets CGAL's Delaunay choose, which is not
// reproducible on a co-circular grid).
#ifndef MA_FUNCTIONS_HPP
#define MA_FUNCTIONS_HPP

#include <map>

#include "lite.hpp"

namespace MA {

// rho(x, y) = a x + b y + c on one face.
template <class K> class Linear_function {
  typedef typename K::Point_2 Point;
  typedef typename K::FT FT;
  FT _a, _b, _c;

 public:
  typedef FT result_type;
  Linear_function() : _a(0), _b(0), _c(0) {}
  // the plane through (p, fp), (q, fq), (r, fr); the reference extrapolates it barycentrically to
  // (0,0), (1,0), (0,1) (functions.hpp:25-53,64-71), which is the solution of this 2x2 system
  Linear_function(const Point &p, FT fp, const Point &q, FT fq, const Point &r, FT fr) {
    const FT e1x = q.x() - p.x(), e1y = q.y() - p.y(), e2x = r.x() - p.x(), e2y = r.y() - p.y();
    const FT f1 = fq - fp, f2 = fr - fp, det = e1x * e2y - e2x * e1y;
    _a = (f1 * e2y - f2 * e1y) / det;
    _b = (e1x * f2 - e2x * f1) / det;
    _c = fp - _a * p.x() - _b * p.y();
  }
  Linear_function(FT a, FT b, FT c) : _a(a), _b(b), _c(c) {}
  FT operator()(const Point &p) const { return _a * p.x() + _b * p.y() + _c; }
};

// image -> triangulation of [-1,1]^2 with one vertex per pixel + PL density
//   vertex (i, j) at (-1 + 2 i/(n-1), -1 + 2 j/(m-1)) carries image(i, m-j-1)/255 + 1e-3
// (functions.hpp:93-102); returns the total mass  sum_f area_f * rho_f(centroid_f)  (:117).
// T must offer make_grid() (MA::lite::Triangulation); with a CGAL triangulation keep the reference's
// own image_to_pl_function — the engine accepts whatever triangulation it produces.
template <class Image, class T, class Function>
double image_to_pl_function(const Image &image, T &t, std::map<typename T::Face_handle, Function> &fs) {
  typedef typename std::decay<decltype(t.finite_faces_begin()->vertex(0)->point())>::type Point;
  const int n = image.width(), m = image.height();
  t.make_grid(n, m, -1.0, -1.0, 1.0, 1.0);
  fs.clear();
  double total = 0;
  for (typename T::Finite_faces_iterator f = t.finite_faces_begin(); f != t.finite_faces_end(); ++f) {
    Point p[3];
    double v[3];
    for (int k = 0; k < 3; ++k) {
      p[k] = f->vertex(k)->point();
      const int id = t.index(f->vertex(k)), i = id / m, j = id % m;
      v[k] = image(i, m - j - 1) / double(255) + 1e-3;
    }
    Function fn(p[0], v[0], p[1], v[1], p[2], v[2]);
    fs[f] = fn;
    const double area = 0.5 * ((p[1].x() - p[0].x()) * (p[2].y() - p[0].y()) - (p[2].x() - p[0].x()) * (p[1].y() - p[0].y()));
    total += area * fn(Point((p[0].x() + p[1].x() + p[2].x()) / 3, (p[0].y() + p[1].y() + p[2].y()) / 3));
  }
  return total;
}

}  // namespace MA
#endif

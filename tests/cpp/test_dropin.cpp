// tests/cpp/test_dropin.cpp — the reference's driver flows (tests/test_opttransport.cpp,
// tests/test_lloyd.cpp, tests/test_voronoi_tri.cpp, tests/test_quantization.cpp) written against THIS
// repository's include/MA with the MA::lite stand-ins for Eigen / CGAL / CImg.  Prints one
// "key value" line per checked quantity; tests/test_cpp_dropin.py compares them with the C-ABI
// results obtained through Python and with the invariants the reference's drivers print.
//
// build: g++ -std=c++14 -O2 -I include tests/cpp/test_dropin.cpp -L mongeampere_b200 -lma_b200 -Wl,-rpath,...
#include <MA/lloyd.hpp>
#include <MA/optimal_transport.hpp>
#include <MA/voronoi_polygon_intersection.hpp>
#include <MA/rasterization.hpp>
#include <MA/voronoi_triangulation_intersection.hpp>

#include <cstdio>
#include <cstdlib>

typedef MA::lite::Kernel K;
typedef MA::lite::Point Point;
typedef MA::lite::Vector VectorXd;
typedef MA::lite::Matrix MatrixXd;
typedef MA::lite::SparseMatrix SparseMatrix;
typedef MA::lite::Triangulation T;

static double rr() { return 2 * double(rand() / (RAND_MAX + 1.0)) - 1; }  // tests/test_opttransport.cpp:19-22

int main(int argc, const char **argv) {
  const size_t N = argc > 1 ? atoi(argv[1]) : 1000;
  const int n = argc > 2 ? atoi(argv[2]) : 32;
  // a synthetic "image": two Gaussian bumps quantised to 8 bits (functions.hpp:102 adds 1e-3)
  MA::lite::Image image(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double x = -1 + 2.0 * i / (n - 1), y = -1 + 2.0 * j / (n - 1);
      double v = 200 * std::exp(-((x - 0.3) * (x - 0.3) + (y + 0.2) * (y + 0.2)) / 0.08) +
                 120 * std::exp(-((x + 0.4) * (x + 0.4) + (y - 0.4) * (y - 0.4)) / 0.02);
      image(i, j) = std::floor(std::min(v, 255.0));
    }
  T t;
  std::map<T::Face_handle, MA::Linear_function<K>> functions;
  double total_mass = MA::image_to_pl_function(image, t, functions);
  printf("total_mass %.17g\n", total_mass);

  MatrixXd X(N, 2);
  VectorXd masses(N), weights = VectorXd::Zero(N);
  for (size_t i = 0; i < N; ++i) {
    X(i, 0) = 0.999 * rr();
    X(i, 1) = 0.999 * rr();
    masses(i) = total_mass / N;
  }
  // ---- kantorovich at w = 0 (kantorovich.hpp:35-42) ----
  VectorXd g;
  SparseMatrix h;
  double f = MA::kantorovich(t, functions, X, weights, g, h);
  printf("f0 %.17g\n", f);
  printf("sum_g0 %.17g\n", g.sum());
  printf("nnz0 %zu\n", h.nonZeros());
  {
    VectorXd ones = VectorXd::Constant(N, 1.0), r = h * ones;  // Laplacian rows sum to zero
    double mx = 0;
    for (size_t i = 0; i < N; ++i) mx = std::max(mx, std::fabs(r(i)));
    printf("max_rowsum0 %.3g\n", mx);
  }
  // ---- solve_laplacian_matrix (optimal_transport.hpp:41-87) ----
  {
    VectorXd gg = g - masses;
    VectorXd d = MA::solve_laplacian_matrix(h, gg);
    VectorXd r = h * d - gg;
    double mx = 0;
    for (size_t i = 0; i + 1 < N; ++i) mx = std::max(mx, std::fabs(r(i)));
    printf("laplace_residual %.3g\n", mx / std::max(gg.norm(), 1e-300));
    printf("laplace_last %.17g\n", d(N - 1));
  }
  // ---- ot_solve (optimal_transport.hpp:89-193) ----
  VectorXd res;
  MA::Statistics stats;
  MA::ot_solve(t, functions, X, masses, res, 1e-9, 100, false, &stats);
  printf("niter %zu\nneval %zu\n", stats.niter, stats.neval);
  {
    VectorXd g2;
    SparseMatrix h2;
    double f2 = MA::kantorovich(t, functions, X, res, g2, h2);
    printf("f_final %.17g\n", f2);
    printf("final_norm %.3g\n", (g2 - masses).norm());
    printf("w_first %.17g\nw_last %.17g\n", res(0), res(N - 1));
  }
  // ---- lloyd / moments (lloyd.hpp:30-144) ----
  {
    VectorXd m;
    MatrixXd c, c1, inertia;
    MA::lloyd(t, functions, X, res, m, c);
    printf("lloyd_mass_sum %.17g\n", m.sum());
    printf("lloyd_c0 %.17g %.17g\n", c(0, 0), c(0, 1));
    MA::second_moment(t, functions, X, res, m, c1, inertia);
    printf("second_moment0 %.17g %.17g %.17g\n", inertia(0, 0), inertia(0, 1), inertia(0, 2));
    printf("first_moment0 %.17g %.17g\n", c1(0, 0), c1(0, 1));
  }
  // ---- voronoi_triangulation_intersection (vti.hpp:315-343; tests/test_voronoi_tri.cpp:41-67) ----
  {
    MA::lite::Weighted_sites dt(X, res);
    double area = 0;
    size_t pieces = 0;
    std::vector<double> cell_area(N, 0.0);
    MA::voronoi_triangulation_intersection(t, dt, [&](const MA::lite::Polygon &p, T::Face_handle, MA::lite::Weighted_sites::Vertex_handle v) {
      area += p.area();
      cell_area[v->info()] += p.area();
      ++pieces;
    });
    printf("area_sum %.17g\npieces %zu\ncell_area0 %.17g\n", area, pieces, cell_area[0]);
    // the raw traversal (vti.hpp:219-313): symbolic polygons, EDGE_T / EDGE_DT tagged edges.  Rebuild the points from the
    // tags the way kantorovich.hpp:95-102 does (vertex i = line(R[i]) ∩ line(R[i+1])) and compare the areas.
    typedef MA::Pgon_edge_t<T::Vertex_handle, MA::lite::Weighted_sites::Vertex_handle> Edge;
    double raw_area = 0;
    size_t raw_pieces = 0, n_edge_t = 0, n_edge_dt = 0;
    MA::voronoi_triangulation_intersection_raw(t, dt, [&](const std::vector<Edge> &R, T::Face_handle, MA::lite::Weighted_sites::Vertex_handle v) {
      const size_t n = R.size();
      std::vector<double> la(n), lb(n), lc(n);  // a x + b y = c
      for (size_t i = 0; i < n; ++i) {
        if (R[i].type == Edge::EDGE_T) {
          const Point p = R[i].edge_t.first->point(), q = R[i].edge_t.second->point();
          la[i] = p.y() - q.y(); lb[i] = q.x() - p.x(); lc[i] = la[i] * p.x() + lb[i] * p.y();
          ++n_edge_t;
        } else {
          const double xv = R[i].edge_dt.first->point().x(), yv = R[i].edge_dt.first->point().y(), wv = R[i].edge_dt.first->point().weight();
          const double xw = R[i].edge_dt.second->point().x(), yw = R[i].edge_dt.second->point().y(), ww = R[i].edge_dt.second->point().weight();
          if (R[i].edge_dt.first != v) { printf("raw: EDGE_DT does not start at the cell's site\n"); }
          la[i] = 2 * (xw - xv); lb[i] = 2 * (yw - yv); lc[i] = xw * xw + yw * yw - xv * xv - yv * yv + wv - ww;
          ++n_edge_dt;
        }
      }
      MA::lite::Polygon poly;
      for (size_t i = 0; i < n; ++i) {
        const size_t j = (i + 1) % n;
        const double det = la[i] * lb[j] - la[j] * lb[i];
        poly.push_back(Point((lc[i] * lb[j] - lc[j] * lb[i]) / det, (la[i] * lc[j] - la[j] * lc[i]) / det));
      }
      raw_area += poly.area();
      ++raw_pieces;
    });
    printf("raw_area_sum %.17g\nraw_pieces %zu\nraw_edge_t %zu\nraw_edge_dt %zu\n", raw_area, raw_pieces, n_edge_t, n_edge_dt);
  }
  // ---- draw_laguerre_diagram (rasterization.hpp:512-547): colour 1 everywhere => the image integrates the density ----
  {
    MA::lite::Vector colors = MA::lite::Vector::Constant(N, 1.0);
    const size_t W = 37, Hh = 29;
    std::vector<double> img(W * Hh, 0.0);
    MA::draw_laguerre_diagram(t, functions, X, res, colors, -1.0, -1.0, 1.0, 1.0, W, Hh,
                              [&](int x, int y, double v) { img[(size_t)y * W + x] += v; });
    double s = 0;
    for (double v : img) s += v;
    printf("raster_sum %.17g\n", s * (2.0 / W) * (2.0 / Hh));  // pixel area in domain units
  }
  // ---- voronoi_polygon_intersection (voronoi_polygon_intersection.hpp:153-188; tests/test_power.cpp:43-50) ----
  {
    MA::lite::Weighted_sites dt(X, res);
    MA::lite::Polygon P;  // a convex pentagon inside the domain, counter-clockwise
    P.push_back(Point(-0.8, -0.7)); P.push_back(Point(0.6, -0.9)); P.push_back(Point(0.9, 0.1));
    P.push_back(Point(0.2, 0.85)); P.push_back(Point(-0.7, 0.5));
    double area = 0;
    for (MA::lite::Weighted_sites::Finite_vertices_iterator v = dt.finite_vertices_begin(); v != dt.finite_vertices_end(); ++v)
      area += MA::voronoi_polygon_intersection(P, dt, v).area();
    printf("pentagon_area %.17g\ncells_area_sum %.17g\n", P.area(), area);
    // a NON-convex polygon (the cross of tests/test_voronoi_ad.cpp, scaled into the domain): the cells ∩ cross tile it
    MA::lite::Polygon C;
    const double a = 0.3, b = 0.9;
    const double cxs[12] = {-a, a, a, b, b, a, a, -a, -a, -b, -b, -a}, cys[12] = {-b, -b, -a, -a, a, a, b, b, a, a, -a, -a};
    for (int k = 0; k < 12; ++k) C.push_back(Point(cxs[k], cys[k]));
    double carea = 0;
    for (MA::lite::Weighted_sites::Finite_vertices_iterator v = dt.finite_vertices_begin(); v != dt.finite_vertices_end(); ++v)
      carea += MA::voronoi_polygon_intersection(C, dt, v).area();
    printf("cross_area %.17g\ncross_cells_area_sum %.17g\n", C.area(), carea);
  }
  // ---- the same image as an EXPLICIT triangulation (vertices + index triples, as from CGAL's Delaunay of the pixel
  //      grid, which may pick either diagonal of a square: SURVEY App. B T1).  (a) the diagonals of make_grid: the bridge
  //      sends it through ma_set_mesh, which recognises the grid, and kantorovich must give the same numbers;
  //      (b) every second square split the other way: another PL function, its cells still carry its total mass. ----
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<Point> pts((size_t)n * n);
    std::vector<int> tr;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) pts[(size_t)i * n + j] = Point(-1 + 2.0 * i / (n - 1), -1 + 2.0 * j / (n - 1));
    for (int i = 0; i + 1 < n; ++i)
      for (int j = 0; j + 1 < n; ++j) {
        const int a = i * n + j, b = (i + 1) * n + j, c = (i + 1) * n + j + 1, d = i * n + j + 1;
        if (variant == 1 && ((i + j) & 1)) { tr.insert(tr.end(), {a, b, d}); tr.insert(tr.end(), {b, c, d}); }
        else { tr.insert(tr.end(), {a, b, c}); tr.insert(tr.end(), {a, c, d}); }
      }
    T t2;
    t2.assign(pts, tr);
    std::map<T::Face_handle, MA::Linear_function<K>> f2;
    double tm2 = 0;
    for (T::Finite_faces_iterator f = t2.finite_faces_begin(); f != t2.finite_faces_end(); ++f) {
      Point p[3];
      double v[3];
      for (int k = 0; k < 3; ++k) {
        p[k] = f->vertex(k)->point();
        const int id = t2.index(f->vertex(k)), i = id / n, j = id % n;
        v[k] = image(i, n - j - 1) / 255.0 + 1e-3;
      }
      f2[f] = MA::Linear_function<K>(p[0], v[0], p[1], v[1], p[2], v[2]);
      const double area = 0.5 * ((p[1].x() - p[0].x()) * (p[2].y() - p[0].y()) - (p[2].x() - p[0].x()) * (p[1].y() - p[0].y()));
      tm2 += area * (v[0] + v[1] + v[2]) / 3.0;
    }
    VectorXd g2;
    SparseMatrix h2;
    const double fe = MA::kantorovich(t2, f2, X, weights, g2, h2);
    printf(variant == 0 ? "f0_explicit %.17g\n" : "f0_alternating %.17g\n", fe);
    printf(variant == 0 ? "sum_g0_explicit %.17g\n" : "sum_g0_alternating %.17g\n", g2.sum());
    printf(variant == 0 ? "tm_explicit %.17g\n" : "tm_alternating %.17g\n", tm2);
    printf(variant == 0 ? "nnz0_explicit %zu\n" : "nnz0_alternating %zu\n", h2.nonZeros());
  }
  return 0;
}

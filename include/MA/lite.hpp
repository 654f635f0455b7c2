// MA/lite.hpp — minimal stand-ins for the third-party types the MongeAmpere++ templates are
// instantiated with (Eigen::MatrixXd / VectorXd / SparseMatrix<double>, a CGAL triangulation with
// Face_handles, cimg_library::CImg<double>).  None of those libraries is needed by the B200 engine;
// these classes expose exactly the members the reference's templates and drivers touch
// (SURVEY.md §8b "duck-typed members actually used"), so that code written against
// include/MA/*.hpp of the reference compiles against this include/MA/ with a typedef swap.  With the
// real Eigen / CGAL installed the same templates accept the real types.
#ifndef MA_LITE_HPP
#define MA_LITE_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace MA {
namespace lite {

// ---- CGAL::Point_2 / kernel -------------------------------------------------------------------
class Point {
  double _x, _y;

 public:
  Point(double x = 0, double y = 0) : _x(x), _y(y) {}
  double x() const { return _x; }
  double y() const { return _y; }
};
struct Kernel {
  typedef Point Point_2;
  typedef double FT;
};
// CGAL::Weighted_point_2: what a Regular_triangulation_2 vertex carries (kantorovich.hpp:68-69)
class Weighted_point : public Point {
  double _w;

 public:
  Weighted_point(const Point &p = Point(), double w = 0) : Point(p), _w(w) {}
  double weight() const { return _w; }
  const Point &point() const { return *this; }
};

// ---- Eigen::VectorXd --------------------------------------------------------------------------
class Vector {
  std::vector<double> v;

 public:
  Vector() {}
  explicit Vector(size_t n) : v(n, 0.0) {}
  static Vector Zero(size_t n) { return Vector(n); }
  static Vector Constant(size_t n, double c) { Vector r(n); std::fill(r.v.begin(), r.v.end(), c); return r; }
  size_t rows() const { return v.size(); }
  size_t cols() const { return 1; }
  size_t size() const { return v.size(); }
  void resize(size_t n) { v.assign(n, 0.0); }
  double &operator()(size_t i) { return v[i]; }
  double operator()(size_t i) const { return v[i]; }
  double &operator[](size_t i) { return v[i]; }
  double operator[](size_t i) const { return v[i]; }
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
  double minCoeff(size_t *arg = 0) const {
    size_t k = std::min_element(v.begin(), v.end()) - v.begin();
    if (arg) *arg = k;
    return v[k];
  }
  double maxCoeff() const { return *std::max_element(v.begin(), v.end()); }
  double sum() const { double s = 0; for (double a : v) s += a; return s; }
  double dot(const Vector &o) const { double s = 0; for (size_t i = 0; i < v.size(); ++i) s += v[i] * o.v[i]; return s; }
  double norm() const { return std::sqrt(dot(*this)); }
  Vector head(size_t n) const { Vector r(n); std::copy(v.begin(), v.begin() + n, r.v.begin()); return r; }
  Vector operator+(const Vector &o) const { Vector r(*this); for (size_t i = 0; i < v.size(); ++i) r.v[i] += o.v[i]; return r; }
  Vector operator-(const Vector &o) const { Vector r(*this); for (size_t i = 0; i < v.size(); ++i) r.v[i] -= o.v[i]; return r; }
  Vector operator-() const { Vector r(*this); for (double &a : r.v) a = -a; return r; }
  Vector operator*(double s) const { Vector r(*this); for (double &a : r.v) a *= s; return r; }
  Vector &operator-=(const Vector &o) { for (size_t i = 0; i < v.size(); ++i) v[i] -= o.v[i]; return *this; }
  Vector &operator+=(const Vector &o) { for (size_t i = 0; i < v.size(); ++i) v[i] += o.v[i]; return *this; }
};
inline Vector operator*(double s, const Vector &a) { return a * s; }

// ---- Eigen::MatrixXd (column-major) -------------------------------------------------------------
class Matrix {
  size_t r, c;
  std::vector<double> v;

 public:
  Matrix() : r(0), c(0) {}
  Matrix(size_t rows, size_t cols) : r(rows), c(cols), v(rows * cols, 0.0) {}
  static Matrix Zero(size_t rows, size_t cols) { return Matrix(rows, cols); }
  size_t rows() const { return r; }
  size_t cols() const { return c; }
  void resize(size_t rows, size_t cols) { r = rows; c = cols; v.assign(rows * cols, 0.0); }
  double &operator()(size_t i, size_t j) { return v[j * r + i]; }
  double operator()(size_t i, size_t j) const { return v[j * r + i]; }
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
};

// ---- Eigen::Triplet / Eigen::SparseMatrix<double> -------------------------------------------------
// Stored compressed by ROWS: row i holds what the reference's triplets (i, *) sum to
// (kantorovich.hpp:120-121).  The reference's Eigen matrix is column-major; H is symmetric to rounding,
// so the two layouts hold the same numbers up to a transposition (SURVEY.md §8b "H layout note").
struct Triplet {
  int r, c;
  double v;
  Triplet(int r_ = 0, int c_ = 0, double v_ = 0) : r(r_), c(c_), v(v_) {}
  int row() const { return r; }
  int col() const { return c; }
  double value() const { return v; }
};
class SparseMatrix {
  size_t nr, nc;
  std::vector<int> ptr, idx;
  std::vector<double> val;

 public:
  SparseMatrix() : nr(0), nc(0), ptr(1, 0) {}
  SparseMatrix(size_t rows, size_t cols) : nr(rows), nc(cols), ptr(rows + 1, 0) {}
  size_t rows() const { return nr; }
  size_t cols() const { return nc; }
  size_t nonZeros() const { return idx.size(); }
  void makeCompressed() {}
  template <class It> void setFromTriplets(It b, It e) {  // duplicates are summed, as Eigen does
    std::vector<Triplet> t;
    for (It it = b; it != e; ++it) t.push_back(Triplet(it->row(), it->col(), it->value()));
    std::stable_sort(t.begin(), t.end(), [](const Triplet &a, const Triplet &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    ptr.assign(nr + 1, 0); idx.clear(); val.clear();
    for (size_t k = 0; k < t.size(); ++k) {
      if (k && t[k].r == t[k - 1].r && t[k].c == t[k - 1].c) { val.back() += t[k].v; continue; }
      idx.push_back(t[k].c); val.push_back(t[k].v); ptr[t[k].r + 1]++;
    }
    for (size_t i = 0; i < nr; ++i) ptr[i + 1] += ptr[i];
  }
  // adopt ready-made CSR arrays (what the engine hands back)
  void setFromCSR(size_t rows, size_t cols, const int *rowptr, const int *col, const double *v) {
    nr = rows; nc = cols;
    ptr.assign(rowptr, rowptr + rows + 1);
    idx.assign(col, col + rowptr[rows]);
    val.assign(v, v + rowptr[rows]);
  }
  const int *outerIndexPtr() const { return ptr.data(); }
  const int *innerIndexPtr() const { return idx.data(); }
  const double *valuePtr() const { return val.data(); }
  double coeff(size_t i, size_t j) const {
    for (int q = ptr[i]; q < ptr[i + 1]; ++q) if ((size_t)idx[q] == j) return val[q];
    return 0.0;
  }
  Vector diagonal() const { Vector d(nr); for (size_t i = 0; i < nr; ++i) d(i) = coeff(i, i); return d; }
  Vector operator*(const Vector &x) const {
    Vector y(nr);
    for (size_t i = 0; i < nr; ++i) { double s = 0; for (int q = ptr[i]; q < ptr[i + 1]; ++q) s += val[q] * x(idx[q]); y(i) = s; }
    return y;
  }
};

// ---- a triangulation with Face_handles (the members vti.hpp:233-238,261-263,282-291 use) -----------
class Triangulation {
 public:
  struct Vertex {
    Point p;
    int id;
    const Point &point() const { return p; }
  };
  typedef const Vertex *Vertex_handle;
  struct Face {
    Vertex_handle v[3];
    int id;
    Vertex_handle vertex(int i) const { return v[i]; }
  };
  typedef const Face *Face_handle;
  typedef const Vertex *Finite_vertices_iterator;
  typedef const Face *Finite_faces_iterator;
  typedef Point Point_type;

  Triangulation() : gn(0), gm(0) { box[0] = box[1] = box[2] = box[3] = 0; }
  Triangulation(const Triangulation &o) { *this = o; }
  Triangulation &operator=(const Triangulation &o) {
    if (this == &o) return *this;
    std::vector<Point> pts; std::vector<int> tr;
    for (const Vertex &v : o.vs) pts.push_back(v.p);
    for (const Face &f : o.fs) for (int k = 0; k < 3; ++k) tr.push_back(f.v[k]->id);
    assign(pts, tr);
    gn = o.gn; gm = o.gm; for (int k = 0; k < 4; ++k) box[k] = o.box[k];
    return *this;
  }
  // vertices + CCW index triples: the input of CGAL::Triangulation_incremental_builder_2
  // (include/CGAL/Triangulation_incremental_builder_2.h:25-81, tests/test_triangulation.cpp:12-34)
  void assign(const std::vector<Point> &pts, const std::vector<int> &triples) {
    vs.resize(pts.size());
    for (size_t i = 0; i < pts.size(); ++i) { vs[i].p = pts[i]; vs[i].id = (int)i; }
    fs.resize(triples.size() / 3);
    for (size_t f = 0; f < fs.size(); ++f) {
      fs[f].id = (int)f;
      for (int k = 0; k < 3; ++k) fs[f].v[k] = &vs[triples[3 * f + k]];
    }
    gn = gm = 0;
  }
  // n x m vertex grid on [x0,x1] x [y0,y1], vertex (i,j) = index i*m + j, square (i,j) split along
  // (i,j)-(i+1,j+1): the triangulation image_to_pl_function builds (functions.hpp:93-105) with a
  // fixed diagonal (CGAL's own choice is not reproducible)
  void make_grid(int n, int m, double x0, double y0, double x1, double y1) {
    std::vector<Point> pts((size_t)n * m);
    std::vector<int> tr;
    const double dx = (x1 - x0) / double(n - 1), dy = (y1 - y0) / double(m - 1);
    for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) pts[(size_t)i * m + j] = Point(x0 + i * dx, y0 + j * dy);
    tr.reserve((size_t)6 * (n - 1) * (m - 1));
    for (int i = 0; i + 1 < n; ++i) for (int j = 0; j + 1 < m; ++j) {
      int a = i * m + j, b = (i + 1) * m + j, c = (i + 1) * m + j + 1, d = i * m + j + 1;
      tr.push_back(a); tr.push_back(b); tr.push_back(c);
      tr.push_back(a); tr.push_back(c); tr.push_back(d);
    }
    assign(pts, tr);
    gn = n; gm = m; box[0] = x0; box[1] = y0; box[2] = x1; box[3] = y1;
  }
  // structured-grid hint for the engine's fast path (0 x 0 = general mesh)
  bool grid_dims(int &n, int &m, double b[4]) const {
    n = gn; m = gm; for (int k = 0; k < 4; ++k) b[k] = box[k];
    return gn >= 2 && gm >= 2;
  }
  size_t number_of_vertices() const { return vs.size(); }
  size_t number_of_faces() const { return fs.size(); }
  Finite_vertices_iterator finite_vertices_begin() const { return vs.data(); }
  Finite_vertices_iterator finite_vertices_end() const { return vs.data() + vs.size(); }
  Finite_faces_iterator finite_faces_begin() const { return fs.data(); }
  Finite_faces_iterator finite_faces_end() const { return fs.data() + fs.size(); }
  Face_handle face(size_t f) const { return &fs[f]; }
  int index(Vertex_handle v) const { return v->id; }

 private:
  std::vector<Vertex> vs;
  std::vector<Face> fs;
  int gn, gm;
  double box[4];
};

// ---- the Diracs as a "triangulation" argument (vti.hpp:315-319 takes a DT/RT) ---------------------
// Only the vertices are read: finite_vertices_begin/end, ->point() (with .weight()), ->info() = index.
class Weighted_sites {
 public:
  struct Vertex {
    Weighted_point p;
    size_t i;
    const Weighted_point &point() const { return p; }
    size_t info() const { return i; }
  };
  typedef const Vertex *Vertex_handle;
  typedef const Vertex *Finite_vertices_iterator;
  template <class MatrixT, class VectorT> Weighted_sites(const MatrixT &X, const VectorT &w) {
    vs.resize(X.rows());
    for (size_t i = 0; i < vs.size(); ++i) { vs[i].p = Weighted_point(Point(X(i, 0), X(i, 1)), w(i)); vs[i].i = i; }
  }
  size_t number_of_vertices() const { return vs.size(); }
  Finite_vertices_iterator finite_vertices_begin() const { return vs.data(); }
  Finite_vertices_iterator finite_vertices_end() const { return vs.data() + vs.size(); }
  Vertex_handle vertex(size_t i) const { return &vs[i]; }

 private:
  std::vector<Vertex> vs;
};

// ---- CGAL::Polygon_2 as handed to the callbacks (vti.hpp:336-340) ---------------------------------
class Polygon {
  std::vector<Point> p;

 public:
  void push_back(const Point &q) { p.push_back(q); }
  size_t size() const { return p.size(); }
  const Point &operator[](size_t i) const { return p[i]; }
  const Point &vertex(size_t i) const { return p[i]; }
  double area() const {
    double a = 0;
    for (size_t i = 0, n = p.size(); i < n; ++i) { const Point &u = p[i], &w = p[(i + 1) % n]; a += u.x() * w.y() - w.x() * u.y(); }
    return 0.5 * a;
  }
};

// ---- cimg_library::CImg<double> (functions.hpp:91-102, tests/test_zeldovich.cpp:83-84) -------------
class Image {
  int w, h;
  std::vector<double> px;

 public:
  Image() : w(0), h(0) {}
  Image(int width, int height) : w(width), h(height), px((size_t)width * height, 0.0) {}
  // from a PGM file (P2 ascii / P5 binary, 8 or 16 bit), the way the reference's drivers build a CImg<double> from a
  // path (tests/test_opttransport.cpp:45-49); values are kept as stored (0..maxval), (i, j) = (column, row from the top)
  explicit Image(const char *path) : w(0), h(0) {
    std::FILE *f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("MA::lite::Image: cannot open ") + path);
    auto token = [&](std::string &t) {
      t.clear();
      int ch;
      for (;;) {  // skip white space and # comments
        ch = std::fgetc(f);
        if (ch == '#') { while (ch != '\n' && ch != EOF) ch = std::fgetc(f); continue; }
        if (ch == EOF || !std::isspace(ch)) break;
      }
      while (ch != EOF && !std::isspace(ch)) { t.push_back((char)ch); ch = std::fgetc(f); }
      return !t.empty();
    };
    std::string magic, t;
    long maxval = 0;
    bool ok = token(magic) && (magic == "P2" || magic == "P5") && token(t);
    if (ok) { w = std::atoi(t.c_str()); ok = token(t); }
    if (ok) { h = std::atoi(t.c_str()); ok = token(t); }
    if (ok) { maxval = std::atol(t.c_str()); ok = w > 0 && h > 0 && maxval > 0 && maxval < 65536; }
    if (ok) {
      px.resize((size_t)w * h);
      if (magic == "P2") {
        for (size_t k = 0; k < px.size() && ok; ++k) { ok = token(t); px[k] = ok ? std::atof(t.c_str()) : 0.0; }
      } else {  // exactly one white-space byte was consumed after maxval by token()
        const size_t bps = maxval < 256 ? 1 : 2;
        std::vector<unsigned char> raw(px.size() * bps);
        ok = std::fread(raw.data(), 1, raw.size(), f) == raw.size();
        for (size_t k = 0; k < px.size() && ok; ++k) px[k] = bps == 1 ? (double)raw[k] : (double)((raw[2 * k] << 8) | raw[2 * k + 1]);
      }
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error(std::string("MA::lite::Image: not a valid PGM file: ") + path);
  }
  int width() const { return w; }
  int height() const { return h; }
  Image &fill(double v) { std::fill(px.begin(), px.end(), v); return *this; }
  double &operator()(int i, int j) { return px[(size_t)j * w + i]; }
  double operator()(int i, int j) const { return px[(size_t)j * w + i]; }
  const double *data() const { return px.data(); }
};

}  // namespace lite
}  // namespace MA
#endif

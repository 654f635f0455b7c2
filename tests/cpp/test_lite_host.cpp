// tests/cpp/test_lite_host.cpp — host-only parts of the drop-in header layer (no GPU, no libma_b200):
// MA::Linear_function, MA::image_to_pl_function and the MA::lite stand-ins behave like the types they
// replace (functions.hpp:55-120 of the reference; Eigen's setFromTriplets semantics).
#include <MA/functions.hpp>

#include <cmath>
#include <cstdio>

#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  typedef MA::lite::Kernel K;
  typedef MA::lite::Point Point;
  typedef MA::lite::Triangulation T;
  // Linear_function through three points: rho = 2x - 3y + 0.5
  MA::Linear_function<K> f(Point(0.2, 0.1), 2 * 0.2 - 3 * 0.1 + 0.5, Point(1.5, -0.4), 2 * 1.5 + 3 * 0.4 + 0.5, Point(-0.3, 0.9),
                           -2 * 0.3 - 3 * 0.9 + 0.5);
  CHECK(std::fabs(f(Point(0, 0)) - 0.5) < 1e-14 && std::fabs(f(Point(1, 0)) - 2.5) < 1e-14 && std::fabs(f(Point(0, 1)) + 2.5) < 1e-14);
  // image filled with 255 -> rho = 1.001 on [-1,1]^2 (tests/test_zeldovich.cpp:83-85): total mass 4.004
  MA::lite::Image img(2, 2);
  img.fill(255);
  T t;
  std::map<T::Face_handle, MA::Linear_function<K>> fs;
  double tm = MA::image_to_pl_function(img, t, fs);
  CHECK(t.number_of_vertices() == 4 && t.number_of_faces() == 2 && fs.size() == 2);
  CHECK(std::fabs(tm - 4.004) < 1e-13);
  // a 5 x 4 ramp image: rho(x, y) linear in x => total = area * mean
  MA::lite::Image ramp(5, 4);
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 4; ++j) ramp(i, j) = 51.0 * i;  // 0 .. 204
  tm = MA::image_to_pl_function(ramp, t, fs);
  CHECK(t.number_of_vertices() == 20 && t.number_of_faces() == 24);
  CHECK(std::fabs(tm - 4.0 * (0.4 + 1e-3)) < 1e-13);
  int n, m; double box[4];
  CHECK(t.grid_dims(n, m, box) && n == 5 && m == 4 && box[0] == -1 && box[3] == 1);
  // every face is counter-clockwise and its function interpolates the vertex values (image(i, m-j-1)/255 + 1e-3)
  for (T::Finite_faces_iterator fh = t.finite_faces_begin(); fh != t.finite_faces_end(); ++fh) {
    const Point &a = fh->vertex(0)->point(), &b = fh->vertex(1)->point(), &c = fh->vertex(2)->point();
    CHECK((b.x() - a.x()) * (c.y() - a.y()) - (c.x() - a.x()) * (b.y() - a.y()) > 0);
    for (int k = 0; k < 3; ++k) {
      int id = t.index(fh->vertex(k)), i = id / m;
      CHECK(std::fabs(fs[fh](fh->vertex(k)->point()) - (51.0 * i / 255 + 1e-3)) < 1e-14);
    }
  }
  // SparseMatrix::setFromTriplets sums duplicates and sorts (Eigen semantics); diagonal / product
  MA::lite::SparseMatrix A(3, 3);
  std::vector<MA::lite::Triplet> tr = {{0, 1, -1.0}, {0, 0, 1.0}, {2, 2, 4.0}, {0, 1, -0.5}, {1, 1, 2.0}};
  A.setFromTriplets(tr.begin(), tr.end());
  CHECK(A.nonZeros() == 4 && A.coeff(0, 1) == -1.5 && A.coeff(1, 0) == 0.0);
  MA::lite::Vector x = MA::lite::Vector::Constant(3, 1.0), y = A * x;
  CHECK(y(0) == -0.5 && y(1) == 2.0 && y(2) == 4.0 && A.diagonal()(2) == 4.0);
  MA::lite::Matrix X(4, 2);
  X(3, 1) = 7;
  CHECK(X.data()[4 + 3] == 7);  // column-major like Eigen::MatrixXd
  // PGM reader (what stands in for CImg<double>(path), tests/test_opttransport.cpp:45-49): ascii and binary
  {
    const char *pa = "/tmp/ma_lite_test_ascii.pgm", *pb = "/tmp/ma_lite_test_bin.pgm";
    FILE *fa = fopen(pa, "w");
    fprintf(fa, "P2\n# a comment\n3 2\n255\n0 10 20\n30 40 255\n");
    fclose(fa);
    FILE *fb = fopen(pb, "wb");
    fprintf(fb, "P5\n3 2\n255\n");
    const unsigned char raw[6] = {0, 10, 20, 30, 40, 255};
    fwrite(raw, 1, 6, fb);
    fclose(fb);
    MA::lite::Image ia(pa), ib(pb);
    CHECK(ia.width() == 3 && ia.height() == 2 && ib.width() == 3 && ib.height() == 2);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) CHECK(ia(i, j) == raw[j * 3 + i] && ib(i, j) == raw[j * 3 + i]);
    bool threw = false;
    try { MA::lite::Image bad("/tmp/ma_lite_no_such_file.pgm"); } catch (const std::runtime_error &) { threw = true; }
    CHECK(threw);
  }
  printf("ok\n");
  return 0;
}

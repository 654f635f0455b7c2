// MA/b200_bridge.hpp — glue between the MongeAmpere++ header API (this directory) and the C-ABI of
// libma_b200.so (include/ma_b200.h).  Not part of the reference's API; everything lives in
// MA::b200.  The reference is stateless (every call receives t, functions, X), the engine is not
// (device buffers): one engine per host thread is kept alive and its device copy of the mesh / the
// Diracs is refreshed when the arguments change (content fingerprints, see mesh_key / points_key).
//
// There is no CPU fallback: a missing GPU / library failure throws std::runtime_error.
#ifndef MA_B200_BRIDGE_HPP
#define MA_B200_BRIDGE_HPP

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../ma_b200.h"

namespace MA {
namespace b200 {

inline void check(ma_ctx *c, int rc, const char *what) {
  if (rc == MA_OK) return;
  std::string msg = std::string(what) + ": " + (c ? ma_last_error(c) : "no context");
  throw std::runtime_error(msg);
}

inline uint64_t fnv(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
  const unsigned char *p = static_cast<const unsigned char *>(data);
  for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

// has T a structured-grid hint?  (MA::lite::Triangulation::grid_dims)
template <class T> class has_grid_dims {
  template <class U> static auto test(int) -> decltype(std::declval<const U &>().grid_dims(std::declval<int &>(), std::declval<int &>(), std::declval<double *>()), std::true_type());
  template <class U> static std::false_type test(...);

 public:
  static const bool value = decltype(test<T>(0))::value;
};
template <class T> typename std::enable_if<has_grid_dims<T>::value, bool>::type grid_hint(const T &t, int &n, int &m, double box[4]) { return t.grid_dims(n, m, box); }
template <class T> typename std::enable_if<!has_grid_dims<T>::value, bool>::type grid_hint(const T &, int &, int &, double *) { return false; }

// The source mesh as plain arrays, read through the members the reference itself uses on T
// (finite_faces_begin/end, Face_handle::vertex(i)->point(), vti.hpp:233-238) and on Functions
// (find(f)->second(Point), kantorovich.hpp:105-107).
struct MeshArrays {
  std::vector<double> vx, vy, abc;
  std::vector<int> tri;
  int gn = 0, gm = 0;
  double box[4] = {0, 0, 0, 0};
  std::vector<double> rho_v;  // grid meshes: vertex densities
};

template <class T, class Functions> void extract_mesh(const T &t, const Functions &fs, MeshArrays &M) {
  typedef decltype(t.finite_faces_begin()) FIt;
  typedef decltype(t.finite_faces_begin()->vertex(0)) VH;
  typedef typename std::decay<decltype(t.finite_faces_begin()->vertex(0)->point())>::type Pt;
  std::map<VH, int> idx;
  M = MeshArrays();
  for (FIt f = t.finite_faces_begin(); f != t.finite_faces_end(); ++f) {
    typename Functions::const_iterator it = fs.find(f);
    if (it == fs.end()) throw std::runtime_error("MA: a face of the density triangulation has no function");
    const double c = it->second(Pt(0, 0)), a = it->second(Pt(1, 0)) - c, b = it->second(Pt(0, 1)) - c;
    M.abc.push_back(a); M.abc.push_back(b); M.abc.push_back(c);
    for (int k = 0; k < 3; ++k) {
      VH v = f->vertex(k);
      typename std::map<VH, int>::iterator w = idx.find(v);
      if (w == idx.end()) {
        w = idx.insert(std::make_pair(v, (int)M.vx.size())).first;
        M.vx.push_back(v->point().x()); M.vy.push_back(v->point().y());
      }
      M.tri.push_back(w->second);
    }
  }
  // structured grid: same faces in the engine's own numbering => the boundary-segment kernel
  if (grid_hint(t, M.gn, M.gm, M.box) && (size_t)2 * (M.gn - 1) * (M.gm - 1) * 3 == M.tri.size()) {
    M.rho_v.assign((size_t)M.gn * M.gm, 0.0);
    // vertex (i, j) of the grid is recognised by its coordinates; the hint is only trusted if every
    // face vertex sits on a grid node and the faces come in the engine's order (as make_grid builds)
    const double dx = (M.box[2] - M.box[0]) / (M.gn - 1), dy = (M.box[3] - M.box[1]) / (M.gm - 1);
    bool ok = true;
    size_t f = 0;
    for (FIt it = t.finite_faces_begin(); ok && it != t.finite_faces_end(); ++it, ++f) {
      const int si = (int)(f / 2) / (M.gm - 1), sj = (int)(f / 2) % (M.gm - 1);
      for (int k = 0; k < 3; ++k) {
        const double x = it->vertex(k)->point().x(), y = it->vertex(k)->point().y();
        const int i = (int)std::floor((x - M.box[0]) / dx + 0.5), j = (int)std::floor((y - M.box[1]) / dy + 0.5);
        // face 0 of square (si,sj) = {(si,sj),(si+1,sj),(si+1,sj+1)}, face 1 = {(si,sj),(si+1,sj+1),(si,sj+1)}
        const int wi = si + ((f & 1) ? (k == 1) : (k >= 1)), wj = sj + ((f & 1) ? (k >= 1) : (k == 2));
        ok = ok && i == wi && j == wj && std::fabs(x - (M.box[0] + i * dx)) <= 1e-9 * dx && std::fabs(y - (M.box[1] + j * dy)) <= 1e-9 * dy;
        if (ok) M.rho_v[(size_t)i * M.gm + j] = M.abc[3 * f] * x + M.abc[3 * f + 1] * y + M.abc[3 * f + 2];
      }
    }
    if (!ok) { M.gn = M.gm = 0; M.rho_v.clear(); }
  } else {
    M.gn = M.gm = 0;
  }
}

class Engine {
 public:
  ma_ctx *ctx = nullptr;
  uint64_t mesh_key = 0, points_key = 0;
  int N = 0;
  double total_mass = 0;
  std::vector<double> xbuf;

  static Engine &instance() {
    static thread_local Engine e;
    return e;
  }
  ~Engine() { if (ctx) ma_destroy(ctx); }

  ma_ctx *get() {
    if (!ctx) {
      const char *dev = std::getenv("MA_B200_DEVICE");
      int rc = ma_create(&ctx, dev ? std::atoi(dev) : 0);
      if (rc != MA_OK) {
        std::string msg = std::string("MA (libma_b200): ") + (ctx ? ma_last_error(ctx) : "ma_create failed");
        if (ctx) ma_destroy(ctx);
        ctx = nullptr;
        throw std::runtime_error(msg);
      }
    }
    return ctx;
  }
  void invalidate() { mesh_key = points_key = 0; }

  template <class T, class Functions> void set_mesh(const T &t, const Functions &fs) {
    // Fingerprint of the CONTENT: every face's three vertex coordinates and its function's three coefficients (the
    // reference API is stateless, so a caller may rebuild a triangulation at the same address or edit densities in place:
    // addresses and samples prove nothing).  One pass over the faces, cheap next to the extraction + upload it saves.
    uint64_t key = fnv(&t, 0);
    size_t nf = fs.size();
    key = fnv(&nf, sizeof nf, key);
    {
      typedef typename std::decay<decltype(t.finite_faces_begin()->vertex(0)->point())>::type Pt;
      for (typename Functions::const_iterator it = fs.begin(); it != fs.end(); ++it) {
        double s[9];
        for (int k = 0; k < 3; ++k) { s[2 * k] = it->first->vertex(k)->point().x(); s[2 * k + 1] = it->first->vertex(k)->point().y(); }
        s[6] = it->second(Pt(0, 0)); s[7] = it->second(Pt(1, 0)); s[8] = it->second(Pt(0, 1));
        key = fnv(s, sizeof s, key);
      }
    }
    if (key == mesh_key && !std::getenv("MA_B200_NOCACHE")) return;
    MeshArrays M;
    extract_mesh(t, fs, M);
    ma_ctx *c = get();
    if (M.gn) check(c, ma_set_grid(c, M.gn, M.gm, M.box[0], M.box[1], M.box[2], M.box[3], M.rho_v.data(), &total_mass), "ma_set_grid");
    else {
      check(c, ma_set_mesh(c, (int)M.vx.size(), M.vx.data(), M.vy.data(), (int)(M.tri.size() / 3), M.tri.data(), M.abc.data()), "ma_set_mesh");
      total_mass = 0;
    }
    mesh_key = key;
  }

  template <class Matrix> void set_points(const Matrix &X) {
    const size_t n = X.rows();
    xbuf.resize(2 * n);
    for (size_t i = 0; i < n; ++i) { xbuf[i] = X(i, 0); xbuf[n + i] = X(i, 1); }
    uint64_t key = fnv(xbuf.data(), xbuf.size() * 8, fnv(&n, sizeof n));
    if (key == points_key && (int)n == N && !std::getenv("MA_B200_NOCACHE")) return;
    ma_ctx *c = get();
    check(c, ma_set_points(c, (int)n, xbuf.data(), xbuf.data() + n), "ma_set_points");
    N = (int)n;
    points_key = key;
  }
};

template <class Vector> std::vector<double> to_std(const Vector &v) {
  std::vector<double> r(v.size());
  for (size_t i = 0; i < r.size(); ++i) r[i] = v(i);
  return r;
}

struct Triplet {  // what SparseMatrix::setFromTriplets reads (Eigen::Triplet's accessors)
  int r, c;
  double v;
  int row() const { return r; }
  int col() const { return c; }
  double value() const { return v; }
};

}  // namespace b200
}  // namespace MA
#endif

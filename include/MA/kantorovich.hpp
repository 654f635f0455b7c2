// MA/kantorovich.hpp — drop-in for the reference's include/MA/kantorovich.hpp:35-141.
//   double MA::kantorovich(densityT, densityF, X, weights, g, h)
// Same name, argument order and meaning: returns f(w) = sum_i [ w_i m_i - ∫_{Lag_i} rho |x - y_i|^2 ],
// g(i) = m_i = ∫_{Lag_i} rho (overwritten, :83), h = the sparse Hessian (overwritten, :137-139).
// The work is done by libma_b200.so (ma_kantorovich + ma_get_hessian_csr, include/ma_b200.h).
#ifndef MA_KANTOROVICH_HPP
#define MA_KANTOROVICH_HPP

#include <cassert>

#include "b200_bridge.hpp"
#include "functions.hpp"

namespace MA {

template <class T, class Functions, class Matrix, class Vector, class SparseMatrix>
double kantorovich(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights, Vector &g,
                   SparseMatrix &h) {
  const size_t N = X.rows();
  assert((size_t)weights.rows() == N);
  assert(weights.cols() == 1);
  assert(X.cols() == 2);
  b200::Engine &E = b200::Engine::instance();
  E.set_mesh(densityT, densityF);
  E.set_points(X);
  ma_ctx *c = E.get();
  std::vector<double> w = b200::to_std(weights), gm(N);
  double fval = 0;
  int nnz = 0;
  b200::check(c, ma_kantorovich(c, w.data(), &fval, gm.data(), &nnz), "ma_kantorovich");
  std::vector<int> ptr(N + 1), col(nnz > 0 ? nnz : 1);
  std::vector<double> val(nnz > 0 ? nnz : 1);
  b200::check(c, ma_get_hessian_csr(c, ptr.data(), col.data(), val.data()), "ma_get_hessian_csr");
  g = Vector::Zero(N);
  for (size_t i = 0; i < N; ++i) g(i) = gm[i];
  // the same (row, col, value) triplets the reference feeds to setFromTriplets (:79-80,137-139),
  // already summed per entry
  std::vector<b200::Triplet> trip((size_t)nnz);
  for (size_t i = 0; i < N; ++i)
    for (int q = ptr[i]; q < ptr[i + 1]; ++q) { trip[q].r = (int)i; trip[q].c = col[q]; trip[q].v = val[q]; }
  h = SparseMatrix(N, N);
  h.setFromTriplets(trip.begin(), trip.end());
  h.makeCompressed();
  return fval;
}

}  // namespace MA
#endif

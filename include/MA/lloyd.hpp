// MA/lloyd.hpp — drop-in for the reference's include/MA/lloyd.hpp:
//   MA::first_moment (:30-69), MA::second_moment (:71-124), MA::lloyd (:126-144).
// masses(i) = ∫_{Lag_i} rho, centroids(i,:) = ∫ rho (x, y) (first_moment: NOT divided by the mass;
// lloyd divides, :139-143), inertia(i,:) = ∫ rho (x², y², x y).
#ifndef MA_LLOYD_HPP
#define MA_LLOYD_HPP

#include "kantorovich.hpp"

namespace MA {
namespace details {
template <class T, class Functions, class Matrix, class Vector>
void moments(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights, int order, Vector &masses,
             Matrix &centroids, Matrix *inertia) {
  const size_t N = X.rows();
  assert((size_t)weights.rows() == N);
  b200::Engine &E = b200::Engine::instance();
  E.set_mesh(densityT, densityF);
  E.set_points(X);
  ma_ctx *c = E.get();
  std::vector<double> w = b200::to_std(weights), m(N), m1(2 * N), m2(order == 2 ? 3 * N : 1);
  b200::check(c, ma_moments(c, w.data(), order, m.data(), m1.data(), order == 2 ? m2.data() : nullptr), "ma_moments");
  masses = Vector::Zero(N);
  centroids = Matrix::Zero(N, 2);
  if (inertia) *inertia = Matrix::Zero(N, 3);
  for (size_t i = 0; i < N; ++i) {
    masses(i) = m[i];
    centroids(i, 0) = m1[i]; centroids(i, 1) = m1[N + i];
    if (inertia) { (*inertia)(i, 0) = m2[i]; (*inertia)(i, 1) = m2[N + i]; (*inertia)(i, 2) = m2[2 * N + i]; }
  }
}
}  // namespace details

template <class T, class Functions, class Matrix, class Vector>
void first_moment(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights, Vector &masses,
                  Matrix &centroids) {
  details::moments(densityT, densityF, X, weights, 1, masses, centroids, (Matrix *)0);
}

template <class T, class Functions, class Matrix, class Vector>
void second_moment(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights, Vector &masses,
                   Matrix &centroids, Matrix &inertia) {
  details::moments(densityT, densityF, X, weights, 2, masses, centroids, &inertia);
}

template <class T, class Functions, class Matrix, class Vector>
void lloyd(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights, Vector &masses,
           Matrix &centroids) {
  first_moment(densityT, densityF, X, weights, masses, centroids);
  const size_t N = X.rows();
  for (size_t i = 0; i < N; ++i) {
    centroids(i, 0) /= masses[i];
    centroids(i, 1) /= masses[i];
  }
}

}  // namespace MA
#endif

// MA/optimal_transport.hpp — drop-in for the reference's include/MA/optimal_transport.hpp:
//   struct MA::Statistics                                   (:35-39)
//   Vector MA::solve_laplacian_matrix(h, g, verbose)        (:41-87)
//   void   MA::ot_solve(t, functions, X, masses, x, eps_g, maxiter, verbose, stats)   (:89-193)
// The damped Newton loop (including its line search and the `niter++ <= maxiter` quirk) runs inside
// libma_b200.so with the weights resident on the GPU (ma_ot_solve); diagnostics keep the
// reference's wording on std::cerr.
#ifndef MA_OPTIMAL_TRANSPORT_HPP
#define MA_OPTIMAL_TRANSPORT_HPP

#include <iostream>

#include "kantorovich.hpp"

namespace MA {

struct Statistics {
  size_t niter;
  size_t neval;
};

// d with h[0:N-1,0:N-1] d[0:N-1] = g[0:N-1], d[N-1] = 0 (the last index is grounded, :62-65,82-84).
// h is read through outerIndexPtr/innerIndexPtr/valuePtr: a compressed Eigen matrix (column-major) or
// MA::lite::SparseMatrix (row-major) — H is symmetric, so both describe the same system.
template <class SparseMatrix, class Vector> Vector solve_laplacian_matrix(const SparseMatrix &h, const Vector &g, bool verbose = false) {
  const size_t N = g.rows();
  assert((size_t)h.rows() == N && (size_t)h.cols() == N);
  ma_ctx *c = b200::Engine::instance().get();
  std::vector<double> gv = b200::to_std(g), d(N, 0.0);
  int iters = 0;
  int rc = ma_solve_laplacian(c, (int)N, h.outerIndexPtr(), h.innerIndexPtr(), h.valuePtr(), gv.data(), d.data(), &iters);
  if (rc == MA_SINGULAR_HESSIAN) std::cerr << "Error: hessian of Kantorovich's functional is not invertible\n";  // :49-58
  else if (rc == MA_LINSOLVE_RESIDUAL) std::cerr << ma_last_error(c) << "\n";                                 // :68-71
  else b200::check(c, rc, "ma_solve_laplacian");
  if (verbose) std::cerr << "solve_laplacian_matrix: " << iters << " PCG iterations\n";
  Vector r = Vector::Zero(N);
  for (size_t i = 0; i < N; ++i) r(i) = d[i];
  return r;
}

template <class T, class Functions, class Matrix, class Vector>
void ot_solve(const T &t, const Functions &functions, const Matrix &X, const Vector &masses, Vector &x,  // result and initial guess
              double eps_g = 1e-7, size_t maxiter = 100, bool verbose = true, struct Statistics *stats = 0) {
  const size_t N = X.rows();
  assert((size_t)masses.rows() == N);
  assert(masses.cols() == 1);
  assert(X.cols() == 2);
  b200::Engine &E = b200::Engine::instance();
  E.set_mesh(t, functions);
  E.set_points(X);
  ma_ctx *c = E.get();
  const bool have_initial = (size_t)x.size() == N;  // :125-128: otherwise start from zero weights
  std::vector<double> nu = b200::to_std(masses), w(N, 0.0);
  if (have_initial) for (size_t i = 0; i < N; ++i) w[i] = x(i);
  ma_statistics st;
  int rc = ma_ot_solve(c, nu.data(), w.data(), have_initial ? 1 : 0, eps_g, maxiter, verbose ? 1 : 0, &st);
  if (rc == MA_EMPTY_CELL) {
    // :139-148 — the reference prints and returns, leaving x at the initial guess
    if (!have_initial) x = Vector::Zero(N);
    if (stats) { stats->niter = 0; stats->neval = st.neval; }
    return;
  }
  if (rc != MA_OK && rc != MA_NOT_CONVERGED && rc != MA_SINGULAR_HESSIAN && rc != MA_LINSOLVE_RESIDUAL)
    b200::check(c, rc, "ma_ot_solve");
  if (!have_initial) x = Vector::Zero(N);
  for (size_t i = 0; i < N; ++i) x(i) = w[i];
  if (stats) { stats->niter = st.niter; stats->neval = st.neval; }  // :188-192
}

}  // namespace MA
#endif

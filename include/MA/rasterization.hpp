// MA/rasterization.hpp — drop-in for the entry point of the reference's include/MA/rasterization.hpp:512-547:
//   MA::draw_laguerre_diagram(densityT, densityF, X, weights, colors, x0, y0, x1, y1, w, h, put_pixel)
// Every piece (Laguerre cell ∩ face of densityT) is drawn into a w x h image over the box [x0,x1] x [y0,y1] with exact
// pixel coverage; the reference calls put_pixel(x, y, coverage * ff * colors[cell]) once per (piece, pixel) with ff the
// mean of the density at the piece's vertices (:531-541), and its callers accumulate.  Here the GPU accumulates
// (ma_draw_laguerre_diagram, k_raster_pieces) and put_pixel is called ONCE per pixel with the sum — the same image for
// any accumulating put_pixel.  Colours are scalars (one channel per call; call once per channel for RGB).
// The rasteriser's internals (DDA walker, per-pixel case analysis, :164-480) are not part of the API and have no twin.
#ifndef MA_RASTERIZATION_HPP
#define MA_RASTERIZATION_HPP

#include "b200_bridge.hpp"
#include "functions.hpp"

namespace MA {

template <class T, class Functions, class Matrix, class Vector, class ColorVector, class PutPixel>
void draw_laguerre_diagram(const T &densityT, const Functions &densityF, const Matrix &X, const Vector &weights,
                           const ColorVector &colors, double x0, double y0, double x1, double y1, size_t w, size_t h,
                           PutPixel put_pixel) {
  const size_t N = X.rows();
  b200::Engine &E = b200::Engine::instance();
  E.set_mesh(densityT, densityF);
  E.set_points(X);
  ma_ctx *c = E.get();
  std::vector<double> wv = b200::to_std(weights), col(N), img(w * h);
  for (size_t i = 0; i < N; ++i) col[i] = colors[i];
  b200::check(c, ma_draw_laguerre_diagram(c, wv.data(), col.data(), x0, y0, x1, y1, (int)w, (int)h, img.data()), "ma_draw_laguerre_diagram");
  for (size_t y = 0; y < h; ++y)
    for (size_t x = 0; x < w; ++x)
      if (img[y * w + x] != 0.0) put_pixel((int)x, (int)y, img[y * w + x]);
}

}  // namespace MA
#endif

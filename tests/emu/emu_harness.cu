// emu_harness.cu — TEST INFRASTRUCTURE: runs the __host__ __device__ per-thread / per-lane code of
// the CUDA kernels (mongeampere_b200/csrc/ma_cell.cuh) serially on the CPU, so that the clipping /
// integration / neighbour-search logic can be checked against the oracle without a GPU
// (`pytest -m "not gpu"`).  It is NOT a product path: nothing in mongeampere_b200/ links to it.
// The host code below re-does K1 (binning, max-weight pyramid) in the simplest possible way.
//
// Build: nvcc -O2 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr -Xcompiler -fPIC -shared
//        tests/emu/emu_harness.cu -o tests/emu/_build/libma_emu.so
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#define MA_EMU_COUNTERS
#include "../../mongeampere_b200/csrc/ma_seg.cuh"

namespace ma { long long ma_emu_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0}; }
extern "C" void emu_work_counters(long long *out, int reset) {
  for (int k = 0; k < 8; ++k) { out[k] = ma::ma_emu_counters[k]; if (reset) ma::ma_emu_counters[k] = 0; }
}

using namespace ma;

// K2 through the block kernels' lane code first (ma_block.cuh), as launch_cells_lean does on the GPU:
// [0] on/off, [1..3] cells finished by radius 2 / radius 3 / CellSearch in the last evaluation
static int emu_lean[4] = {0, 0, 0, 0};
extern "C" void emu_set_lean(int on) { emu_lean[0] = on; }
// K2 warm path (ma_warm.cuh): seeds = adjacency of an earlier evaluation of the same points, internal order, 16 per cell;
// emu_warm: [0] used (certified), [1] cells rebuilt in round 0, [2] in round 1, [3..5] failing cells per match
static const unsigned *emu_diag = nullptr;  // grid meshes with per-square diagonals: one bit per padded square (ma_seg.cuh)
extern "C" void emu_set_diag(const unsigned *bits) { emu_diag = bits; }
static const int *emu_seed_nbr = nullptr, *emu_seed_cnt = nullptr;
static int emu_warm[6] = {0, 0, 0, 0, 0, 0};
extern "C" void emu_set_seeds(const int *nbr, const int *cnt) { emu_seed_nbr = nbr; emu_seed_cnt = cnt; }
extern "C" void emu_get_warm(int *out) { for (int k = 0; k < 6; ++k) out[k] = emu_warm[k]; }
extern "C" void emu_get_lean(int *out) { for (int k = 0; k < 4; ++k) out[k] = emu_lean[k]; }

extern "C" int emu_eval(int mesh_kind,
                        // grid mesh
                        int gn, int gm, double gx0, double gy0, double gdx, double gdy,
                        // general mesh
                        int nV, const double *vx, const double *vy, int nF, const int *tri,
                        // densities
                        const double *abc, const double *rho_v,
                        // Diracs
                        int N, const double *x, const double *y, const double *w,
                        // knobs
                        int kmax, int maxv_piece, int mode, double filter_tol, int bin_target, int nlanes, int use_seg,
                        // outputs (internal order) + permutation
                        int *perm_out, double *mass, double *fcell, int *nbr_cnt, int *nbr, double *hslot,
                        unsigned long long *touched, double *mom, long long *counters, int *flags_out) {
  Params p;
  memset(&p, 0, sizeof p);
  // ---- mesh ----
  p.mesh_kind = mesh_kind;
  p.abc = abc;
  p.rho_v = rho_v;
  std::vector<int> tbin_ptr, tbin_face;
  if (mesh_kind == MESH_GRID) {
    p.gn = gn; p.gm = gm; p.gx0 = gx0; p.gy0 = gy0; p.gdx = gdx; p.gdy = gdy;
    p.nF = 2 * (gn - 1) * (gm - 1);
    p.bb[0] = gx0; p.bb[1] = gy0; p.bb[2] = gx0 + (gn - 1) * gdx; p.bb[3] = gy0 + (gm - 1) * gdy;
  } else {
    p.nF = nF; p.vx = vx; p.vy = vy; p.tri = tri;
    p.bb[0] = p.bb[1] = 1e300; p.bb[2] = p.bb[3] = -1e300;
    for (int i = 0; i < nV; ++i) {
      p.bb[0] = std::min(p.bb[0], vx[i]); p.bb[2] = std::max(p.bb[2], vx[i]);
      p.bb[1] = std::min(p.bb[1], vy[i]); p.bb[3] = std::max(p.bb[3], vy[i]);
    }
    int g = (int)std::ceil(std::sqrt(std::max(1.0, nF / 2.0)));
    p.tg = g;
    p.tinvx = g / (p.bb[2] - p.bb[0]); p.tinvy = g / (p.bb[3] - p.bb[1]);
    auto range = [&](int f, int &i0, int &i1, int &j0, int &j1) {
      double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
      for (int k = 0; k < 3; ++k) {
        int v = tri[3 * f + k];
        x0 = std::min(x0, vx[v]); x1 = std::max(x1, vx[v]);
        y0 = std::min(y0, vy[v]); y1 = std::max(y1, vy[v]);
      }
      i0 = std::min(std::max((int)std::floor((x0 - p.bb[0]) * p.tinvx), 0), g - 1);
      i1 = std::min(std::max((int)std::floor((x1 - p.bb[0]) * p.tinvx), 0), g - 1);
      j0 = std::min(std::max((int)std::floor((y0 - p.bb[1]) * p.tinvy), 0), g - 1);
      j1 = std::min(std::max((int)std::floor((y1 - p.bb[1]) * p.tinvy), 0), g - 1);
    };
    tbin_ptr.assign((size_t)g * g + 1, 0);
    for (int f = 0; f < nF; ++f) {
      int i0, i1, j0, j1; range(f, i0, i1, j0, j1);
      for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) tbin_ptr[(size_t)j * g + i + 1]++;
    }
    for (size_t b = 0; b < (size_t)g * g; ++b) tbin_ptr[b + 1] += tbin_ptr[b];
    tbin_face.resize(tbin_ptr.back());
    std::vector<int> fill(tbin_ptr.begin(), tbin_ptr.end() - 1);
    for (int f = 0; f < nF; ++f) {
      int i0, i1, j0, j1; range(f, i0, i1, j0, j1);
      for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) tbin_face[fill[(size_t)j * g + i]++] = f;
    }
    p.tbin_ptr = tbin_ptr.data(); p.tbin_face = tbin_face.data();
  }
  std::vector<double> rho_pad;
  if (mesh_kind == MESH_GRID && rho_v) {  // ma_set_grid's padded copy
    rho_pad.resize((size_t)(gn + 2) * (gm + 2));
    for (int a = -1; a <= gn; ++a)
      for (int b = -1; b <= gm; ++b)
        rho_pad[(size_t)(a + 1) * (gm + 2) + (b + 1)] = rho_v[(size_t)std::min(std::max(a, 0), gn - 1) * gm + std::min(std::max(b, 0), gm - 1)];
    p.rho_p = rho_pad.data();
    p.diag = emu_diag;
  }
  // ---- K1 on the host ----
  double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
  for (int i = 0; i < N; ++i) {
    bx0 = std::min(bx0, x[i]); bx1 = std::max(bx1, x[i]);
    by0 = std::min(by0, y[i]); by1 = std::max(by1, y[i]);
  }
  double ext = std::max(std::max(bx1 - bx0, by1 - by0), 1e-300) * (1 + 1e-9);
  int L = 0;
  while (((long long)1 << (2 * L)) * bin_target < N && L < 12) ++L;
  const int G = 1 << L;
  const double pinv = G / ext;
  p.L = L; p.px0 = bx0; p.py0 = by0; p.ph = ext / G;
  std::vector<unsigned> code(N);
  std::vector<int> order(N);
  for (int i = 0; i < N; ++i) {
    int bx = std::min(std::max((int)((x[i] - bx0) * pinv), 0), G - 1);
    int by = std::min(std::max((int)((y[i] - by0) * pinv), 0), G - 1);
    code[i] = morton2(bx, by);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
  const size_t nb = (size_t)G * G;
  std::vector<int> bin_start(nb + 1, 0);
  for (int i = 0; i < N; ++i) bin_start[code[i] + 1]++;
  for (size_t b = 0; b < nb; ++b) bin_start[b + 1] += bin_start[b];
  std::vector<double> xs(N), ys(N), ws(N);
  for (int k = 0; k < N; ++k) { xs[k] = x[order[k]]; ys[k] = y[order[k]]; ws[k] = w[order[k]]; perm_out[k] = order[k]; }
  std::vector<double> wmax((4 * nb - 1) / 3);
  const double NEG = -std::numeric_limits<double>::infinity();
  for (size_t b = 0; b < nb; ++b) {
    double m = NEG;
    for (int k = bin_start[b]; k < bin_start[b + 1]; ++k) m = std::max(m, ws[k]);
    wmax[(nb - 1) / 3 + b] = m;
  }
  for (int l = L - 1; l >= 0; --l) {
    size_t nl = (size_t)1 << (2 * l);
    for (size_t c = 0; c < nl; ++c) {
      const double *ch = &wmax[(4 * nl - 1) / 3 + 4 * c];
      wmax[(nl - 1) / 3 + c] = std::max(std::max(ch[0], ch[1]), std::max(ch[2], ch[3]));
    }
  }
  // per-node supporting planes (the host twin of k_node_fit / k_node_alpha)
  const size_t nnodes = (4 * nb - 1) / 3;
  std::vector<double> nodeG(2 * nnodes, 0.0);
  std::vector<unsigned long long> nodeA(nnodes, dkey(std::numeric_limits<double>::infinity()));
  {
    const double cx = bx0 + 0.5 * ext, cy = by0 + 0.5 * ext;
    auto fit = [&](int l, unsigned c, double &Gx, double &Gy) {
      int sh = 2 * (L - l);
      int s = bin_start[(size_t)c << sh], e = bin_start[((size_t)c + 1) << sh];
      double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = s; j < e; ++j) {
        double X = xs[j] - cx, Y = ys[j] - cy, W = ws[j];
        m[0] += X; m[1] += Y; m[2] += X * X; m[3] += Y * Y; m[4] += X * Y; m[5] += W; m[6] += X * W; m[7] += Y * W;
      }
      return node_gradient((double)(e - s), m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], Gx, Gy);
    };
    for (int l = 0; l <= L; ++l)
      for (unsigned c = 0; c < (1u << (2 * l)); ++c) {
        double Gx = 0, Gy = 0;
        for (int a = l; a >= 0; --a) {
          if (fit(a, c >> (2 * (l - a)), Gx, Gy)) break;
          Gx = Gy = 0;
        }
        size_t node = level_offset(l) + c;
        nodeG[2 * node] = Gx; nodeG[2 * node + 1] = Gy;
        int sh = 2 * (L - l);
        double S = (ext / G) * (double)(1u << (L - l));
        double zx = bx0 + (morton_compact1(c) + 0.5) * S, zy = by0 + (morton_compact1(c >> 1) + 0.5) * S;
        double al = std::numeric_limits<double>::infinity();
        for (int j = bin_start[(size_t)c << sh]; j < bin_start[((size_t)c + 1) << sh]; ++j) {
          double tx = xs[j] - zx, ty = ys[j] - zy;
          al = std::min(al, tx * tx + ty * ty - ws[j] + Gx * tx + Gy * ty);
        }
        nodeA[node] = dkey(al);
      }
  }
  int no_abort = 0;
  p.nodeG = nodeG.data(); p.nodeA = nodeA.data(); p.abort_flag = &no_abort; p.abort_on_empty = 0; p.clip_a = p.clip_b = 1; p.refill_at = 8; p.rmax = 3; p.rtree = 1;
  p.N = N; p.xs = xs.data(); p.ys = ys.data(); p.ws = ws.data();
  p.cell_lo = 0; p.cell_hi = N;
  std::vector<int> bin_rm(2 * nb);
  for (int cy = 0; cy < G; ++cy)
    for (int cx = 0; cx < G; ++cx) {
      unsigned code = morton2(cx, cy);
      bin_rm[2 * ((size_t)cy * G + cx)] = bin_start[code];
      bin_rm[2 * ((size_t)cy * G + cx) + 1] = bin_start[code + 1];
    }
  p.bin_rm = bin_rm.data();
  p.bin_start = bin_start.data(); p.wmax = wmax.data();
  double wstat[4] = {0, 0, 1e300, -1e300};
  for (int k = 0; k < N; ++k) { wstat[0] += ws[k]; wstat[1] += ws[k] * ws[k]; wstat[2] = std::min(wstat[2], ws[k]); wstat[3] = std::max(wstat[3], ws[k]); }
  p.wstat = wstat;
  // ---- outputs ----
  std::vector<double> cell_bb((size_t)4 * N);
  std::vector<unsigned long long> cnt(CNT_N, 0);
  int flags4[4] = {0, 0, 0, 0};  // mirrors the device flags[4]: status bits, abort, K2 exact-stage count, spare
  int &flags = flags4[0];
  p.kmax = kmax; p.nbr = nbr; p.nbr_cnt = nbr_cnt; p.cell_bb = cell_bb.data();
  p.mass = mass; p.fcell = fcell; p.hslot = hslot; p.touched = touched; p.mom = mom;
  p.counters = cnt.data(); p.stats = 1; p.flags = flags4; p.filter_tol = filter_tol;
  // the block grid (k_blk_count / k_blk_scatter / k_blk_fill / k_gather_w): ~1 Dirac per row-major bin
  const int bG = std::max(1, (int)std::ceil(std::sqrt((double)N)));
  const size_t nbb = (size_t)bG * bG;
  p.bG = bG; p.bph = ext / bG; p.binv = bG / ext;
  std::vector<int> rm_start(nbb + 1, 0), rm2s(N), blk(N);
  std::vector<double> xr(N), yr(N), wr(N);
  for (int k = 0; k < N; ++k) {
    const int bx = std::min(std::max((int)((xs[k] - p.px0) * p.binv), 0), bG - 1);
    const int by = std::min(std::max((int)((ys[k] - p.py0) * p.binv), 0), bG - 1);
    blk[k] = by * bG + bx;
    rm_start[blk[k] + 1]++;
  }
  for (size_t q = 0; q < nbb; ++q) rm_start[q + 1] += rm_start[q];
  {
    std::vector<int> fill(rm_start.begin(), rm_start.end() - 1);
    for (int k = 0; k < N; ++k) {  // ascending k inside a bin, as k_blk_fill sorts
      const int pos = fill[blk[k]]++;
      xr[pos] = xs[k]; yr[pos] = ys[k]; wr[pos] = ws[k]; rm2s[pos] = k;
    }
  }
  p.xr = xr.data(); p.yr = yr.data(); p.wr = wr.data(); p.rm2s = rm2s.data(); p.rm_start = rm_start.data();
  const bool graded = (wstat[3] - wstat[2]) > MA_LEAN_RANGE * p.bph * p.bph;
  emu_lean[1] = emu_lean[2] = emu_lean[3] = 0;
  // ---- K2 ----
  const int maxv_cell = kmax + 4;
  // warm path first, as launch_cells_warm does: seed every cell, match, rebuild, match, rebuild, match
  std::vector<int> ring, ring_n, cstate, WT, WN;
  std::vector<double> WX, WY;
  bool warm_ok = false;
  for (int k = 0; k < 6; ++k) emu_warm[k] = 0;
  if (emu_seed_nbr && kmax == 16) {
    ring.assign((size_t)N * RING_STRIDE, -9); ring_n.assign(N, -1); cstate.assign(N, WARM_SEEDED);
    WX.assign((size_t)N * 16, 0.0); WY.assign((size_t)N * 16, 0.0); WT.assign((size_t)N * 16, 0); WN.assign(N, -2);
    p.ring = ring.data(); p.ring_n = ring_n.data(); p.cstate = cstate.data();
    std::vector<double> px(16), py(16);
    std::vector<int> pt(16);
    PolyRef<1, true> PP{px.data(), py.data(), pt.data()};
    std::vector<int> list;
    auto keep = [&](int i, int n) {
      WN[i] = n;
      for (int k = 0; k < n; ++k) { WX[(size_t)i * 16 + k] = PP.X(k); WY[(size_t)i * 16 + k] = PP.Y(k); WT[(size_t)i * 16 + k] = PP.T(k); }
      ring_store(p.ring, p.ring_n, i, PP, n);
    };
    for (int i = 0; i < N; ++i) {
      CellSearch<PolyRef<1, true>> S;
      S.init(p, i, PP);
      const int cnt = emu_seed_cnt[i];
      seed_build(p, S, PP, 16, cnt > 0, emu_seed_nbr + (size_t)i * RING_STRIDE, cnt);
      if (cnt <= 0 || S.status != 0) { ring_n[i] = -1; cstate[i] = WARM_QUEUED; list.push_back(i); }
      else { keep(i, S.n); cstate[i] = S.n == 0 ? WARM_EXACT : WARM_SEEDED; }
    }
    for (int round = 0; round < 3; ++round) {
      int nfail = 0;
      for (int i = 0; i < N; ++i)
        if (!ring_match(p.ring, p.ring_n, i, [&](int c) { if (cstate[c] == WARM_SEEDED) { cstate[c] = WARM_QUEUED; list.push_back(c); } })) ++nfail;
      emu_warm[3 + round] = nfail;
      if (round == 2) break;
      emu_warm[1 + round] = (int)list.size();
      for (int i : list) {
        int fl = 0;
        int n = cell_build(p, i, PP, 16, &fl);
        flags |= fl;
        if (n < 0) n = 0;
        keep(i, n);
        cstate[i] = WARM_EXACT;
      }
      list.clear();
    }
    warm_ok = emu_warm[5] == 0 && !(flags & (FLAG_CELL_OVERFLOW | FLAG_KMAX_OVERFLOW));
    emu_warm[0] = warm_ok;
    if (!warm_ok) flags = 0;  // redone cold below
  }
  {
    std::vector<double> px(maxv_cell), py(maxv_cell);
    std::vector<int> pt(maxv_cell);
    for (int i = 0; i < N; ++i) {
      int fl = 0;
      // kmax = 16 is the packed-polygon class of the GPU (capacity 16), the larger ones shift arrays
      PolyRef<1, true> PP{px.data(), py.data(), pt.data()};
      PolyRef<1, false> PA{px.data(), py.data(), pt.data()};
      int n = -2;
      if (warm_ok) {  // the polygon the warm path certified
        n = WN[i];
        for (int k = 0; k < n; ++k) { PP.SX(k) = WX[(size_t)i * 16 + k]; PP.SY(k) = WY[(size_t)i * 16 + k]; PP.ST(k) = WT[(size_t)i * 16 + k]; }
        PP.ord = 0xfedcba9876543210ull;
        PP.used = (1u << n) - 1u;
      }
      if (n == -2 && kmax == 16 && emu_lean[0] && !graded) {
        {  // pass 1: block of radius 2; pass 2 continues from its polygon with the ring around it; pass 3: radius 5 from scratch
          CellSearch<PolyRef<1, true>> S;
          S.init(p, i, PP);
          bool cert = false;
          block_search<-1, 2>(p, S, PP, 16, true, cert);
          if (cert) { n = S.n; emu_lean[1]++; }
          else if (S.phase == 0) {
            block_search<2, 3>(p, S, PP, 16, true, cert);
            if (cert) { n = S.n; emu_lean[2]++; }
          }
          if (n == -2) {
            S.init(p, i, PP);
            block_search<-1, 5>(p, S, PP, 16, true, cert);
            if (cert) { n = S.n; emu_lean[2]++; }
          }
        }
      }
      if (n == -2) {
        emu_lean[3]++;
        n = (kmax == 16) ? cell_build(p, i, PP, 16, &fl) : cell_build(p, i, PA, maxv_cell, &fl);
      }
      flags |= fl;
      if (n < 0) n = 0;
      auto finish = [&](auto &P) {
      cell_emit(p, i, P, n);
      if (use_seg) {  // what k_cells_seg does after K2
        SegAcc acc;
        unsigned long long tch = 0;
        {  // line-major formulation (what k_seg runs)
          double E[80];
          auto tg = [&](int k) { return P.T(k); };
          if (emu_diag) {
            if (mode == MODE_KANTOROVICH) tch = cell_integrate_lines<MODE_KANTOROVICH, 1, true>(p, i, P, n, acc, hslot + (size_t)i * kmax, E, tg);
            else if (mode == MODE_MOMENTS1) cell_integrate_lines<MODE_MOMENTS1, 1, true>(p, i, P, n, acc, nullptr, E, tg);
            else cell_integrate_lines<MODE_MOMENTS2, 1, true>(p, i, P, n, acc, nullptr, E, tg);
          } else {
            if (mode == MODE_KANTOROVICH) tch = cell_integrate_lines<MODE_KANTOROVICH, 1>(p, i, P, n, acc, hslot + (size_t)i * kmax, E, tg);
            else if (mode == MODE_MOMENTS1) cell_integrate_lines<MODE_MOMENTS1, 1>(p, i, P, n, acc, nullptr, E, tg);
            else cell_integrate_lines<MODE_MOMENTS2, 1>(p, i, P, n, acc, nullptr, E, tg);
          }
        }
        mass[i] = acc.mass;
        fcell[i] = acc.mass * ws[i] - acc.cost;
        touched[i] = tch;
        if (mom) {
          const double xi = xs[i], yi = ys[i], m = acc.mass;
          double *o = mom + 6 * (size_t)i;
          o[0] = m; o[1] = acc.m[0] + xi * m; o[2] = acc.m[1] + yi * m;
          o[3] = acc.m[2] + 2 * xi * acc.m[0] + xi * xi * m;
          o[4] = acc.m[3] + 2 * yi * acc.m[1] + yi * yi * m;
          o[5] = acc.m[4] + xi * acc.m[1] + yi * acc.m[0] + xi * yi * m;
        }
      }
      };
      if (kmax == 16) finish(PP); else finish(PA);
    }
  }
  // ---- K3 ----
  if (!use_seg) {
    std::vector<double> tDx(kmax), tDy(kmax), tC(kmax), tS(kmax), hacc((size_t)kmax * nlanes);
    std::vector<int> tJ(kmax);
    std::vector<double> px(maxv_piece), py(maxv_piece);
    std::vector<int> pt(maxv_piece);
    PolyRef<1> P{px.data(), py.data(), pt.data()};
    for (int i = 0; i < N; ++i) {
      CellTable T{tDx.data(), tDy.data(), tC.data(), tS.data(), tJ.data(), 0};
      int k = nbr_cnt[i];
      T.k = k < 0 ? 0 : k;
      std::fill(hacc.begin(), hacc.end(), 0.0);
      double m = 0, cost = 0, mo[5] = {0, 0, 0, 0, 0};
      unsigned long long tch = 0;
      if (k >= 0) {
        for (int s = 0; s < k; ++s) cell_table_fill(p, i, s, T);
        for (int lane = 0; lane < nlanes; ++lane) {
          LaneAcc acc;
          lane_acc_zero(acc);
          if (mode == MODE_KANTOROVICH)
            lane_pieces<1, MODE_KANTOROVICH>(p, i, lane, nlanes, P, maxv_piece, T, hacc.data() + lane, nlanes, acc);
          else if (mode == MODE_MOMENTS1)
            lane_pieces<1, MODE_MOMENTS1>(p, i, lane, nlanes, P, maxv_piece, T, hacc.data() + lane, nlanes, acc);
          else
            lane_pieces<1, MODE_MOMENTS2>(p, i, lane, nlanes, P, maxv_piece, T, hacc.data() + lane, nlanes, acc);
          m += acc.mass; cost += acc.cost; tch |= acc.touched;
          for (int q = 0; q < 5; ++q) mo[q] += acc.m[q];
          for (int q = 0; q < CNT_N; ++q) cnt[q] += acc.cnt[q];
        }
        cnt[CNT_SUMK] += k;
      }
      mass[i] = m;
      fcell[i] = m * ws[i] - cost;
      touched[i] = tch;
      for (int s = 0; s < kmax; ++s) {
        double h = 0;
        if (s < T.k && ((tch >> s) & 1ull))
          for (int lane = 0; lane < nlanes; ++lane) h += hacc[(size_t)s * nlanes + lane];
        hslot[(size_t)i * kmax + s] = h;
      }
      if (mom) {
        const double xi = xs[i], yi = ys[i];
        double *o = mom + 6 * (size_t)i;
        o[0] = m; o[1] = mo[0] + xi * m; o[2] = mo[1] + yi * m;
        o[3] = mo[2] + 2 * xi * mo[0] + xi * xi * m;
        o[4] = mo[3] + 2 * yi * mo[1] + yi * yi * m;
        o[5] = mo[4] + xi * mo[1] + yi * mo[0] + xi * yi * m;
      }
    }
  }
  for (int q = 0; q < CNT_N; ++q) counters[q] = (long long)cnt[q];
  *flags_out = flags;
  return 0;
}

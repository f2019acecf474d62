// gp_host.hpp -- TEST-ONLY scalar model of the GP pass (csrc/gp2d.cu): same data preparation, the same
// tile-blocked symmetric sweep (block Gauss-Jordan with an augmented right-hand side) and the same
// optimiser (csrc/lbfgs4.cuh), written as plain loops so the algebra can be checked against
// oracle/gp2d.py in a container without a GPU.  Never linked into libmallorn_b200.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "lbfgs4.cuh"
#include "objview.cuh"

namespace gphost {
using namespace mlrn;

constexpr int TB = 16;
static const double WAVE[6] = {3670.0, 4825.0, 6222.0, 7545.0, 8691.0, 9710.0};
constexpr double TINY = 1.25e-12;
constexpr double LOG_2PI = 1.8378770664093453;

struct Gp {
  int n = 0, T = 0;
  std::vector<double> t, lam, y, noise, alpha;
  std::vector<double> A;        // full (16T)^2 row-major, symmetric
  double scale = 1.0;

  bool eval(const double* p, double& f, double* g) {
    const int np = T * TB;
    const double amp = 2.0 * std::exp(p[1]), m0 = std::exp(p[2]), m1 = std::exp(p[3]);
    A.assign((size_t)np * np, 0.0);
    for (int i = 0; i < np; ++i)
      for (int j = 0; j < np; ++j) {
        double v;
        if (i >= n || j >= n) v = (i == j) ? 1.0 : 0.0;
        else {
          double dt = t[i] - t[j], dl = lam[i] - lam[j];
          double s = std::sqrt(3.0 * (dt * dt / m0 + dl * dl / m1));
          v = amp * (1.0 + s) * std::exp(-s);
          if (i == j) v += noise[i];
        }
        A[(size_t)i * np + j] = v;
      }
    std::vector<double> b(np, 0.0);
    for (int i = 0; i < n; ++i) b[i] = y[i] - p[0];
    double quad = 0.0, logdet = 0.0;
    bool ok = true;
    std::vector<double> V((size_t)np * TB);
    for (int k = 0; k < T && ok; ++k) {
      double L[TB][TB] = {}, Li[TB][TB] = {};
      for (int r = 0; r < TB; ++r)                       // Cholesky of the pivot block
        for (int c = 0; c <= r; ++c) {
          double s = A[(size_t)(k * TB + r) * np + k * TB + c];
          for (int m = 0; m < c; ++m) s -= L[r][m] * L[c][m];
          if (r == c) { if (!(s > 0.0)) { ok = false; } L[r][r] = std::sqrt(s); logdet += 2.0 * std::log(L[r][r]); }
          else L[r][c] = s / L[c][c];
        }
      if (!ok) break;
      for (int c = 0; c < TB; ++c)                       // Li = L^-1
        for (int r = c; r < TB; ++r) {
          double s = (r == c) ? 1.0 : 0.0;
          for (int m = c; m < r; ++m) s -= L[r][m] * Li[m][c];
          Li[r][c] = s / L[r][r];
        }
      double w[TB];
      for (int r = 0; r < TB; ++r) { double s = 0.0; for (int c = 0; c <= r; ++c) s += Li[r][c] * b[k * TB + c]; w[r] = s; quad += s * s; }
      for (int i = 0; i < np; ++i) {                     // V_i = X_i Li^T
        if (i / TB == k) continue;
        for (int c = 0; c < TB; ++c) { double s = 0.0; for (int m = 0; m <= c; ++m) s += A[(size_t)i * np + k * TB + m] * Li[c][m]; V[(size_t)i * TB + c] = s; }
      }
      for (int i = 0; i < np; ++i) {
        if (i / TB == k) continue;
        for (int j = 0; j < np; ++j) {
          if (j / TB == k) continue;
          double s = 0.0; for (int m = 0; m < TB; ++m) s += V[(size_t)i * TB + m] * V[(size_t)j * TB + m];
          A[(size_t)i * np + j] -= s;
        }
        double s = 0.0; for (int m = 0; m < TB; ++m) s += V[(size_t)i * TB + m] * w[m];
        b[i] -= s;
      }
      for (int i = 0; i < np; ++i) {                     // column k <- V Li, row k by symmetry
        if (i / TB == k) continue;
        for (int c = 0; c < TB; ++c) {
          double s = 0.0; for (int m = c; m < TB; ++m) s += V[(size_t)i * TB + m] * Li[m][c];
          A[(size_t)i * np + k * TB + c] = s; A[(size_t)(k * TB + c) * np + i] = s;
        }
      }
      for (int r = 0; r < TB; ++r)
        for (int c = 0; c < TB; ++c) { double s = 0.0; for (int m = std::max(r, c); m < TB; ++m) s += Li[m][r] * Li[m][c]; A[(size_t)(k * TB + r) * np + k * TB + c] = -s; }
      for (int c = 0; c < TB; ++c) { double s = 0.0; for (int m = c; m < TB; ++m) s += Li[m][c] * w[m]; b[k * TB + c] = s; }
    }
    double ll = -0.5 * quad - 0.5 * logdet - 0.5 * n * LOG_2PI;
    if (!ok || !std::isfinite(ll)) { f = 1e25; for (int i = 0; i < 4; ++i) g[i] = 0.0; return false; }
    alpha.assign(b.begin(), b.begin() + n);
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
    for (int i = 0; i < n; ++i) {
      g0 += alpha[i];
      for (int j = 0; j < n; ++j) {
        double Aij = alpha[i] * alpha[j] + A[(size_t)i * np + j];     // A holds -K^-1
        double dt = t[i] - t[j], dl = lam[i] - lam[j];
        double a0 = dt * dt / m0, a1 = dl * dl / m1, s = std::sqrt(3.0 * (a0 + a1)), ex = std::exp(-s);
        g1 += Aij * amp * (1.0 + s) * ex; g2 += Aij * 1.5 * amp * ex * a0; g3 += Aij * 1.5 * amp * ex * a1;
      }
    }
    f = -ll; g[0] = -g0; g[1] = -0.5 * g1; g[2] = -0.5 * g2; g[3] = -0.5 * g3;
    return true;
  }

  double predict(const double* p, double tq, double lq) const {
    const double amp = 2.0 * std::exp(p[1]), m0 = std::exp(p[2]), m1 = std::exp(p[3]);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      double dt = tq - t[i], dl = lq - lam[i];
      double s = std::sqrt(3.0 * (dt * dt / m0 + dl * dl / m1));
      acc += amp * (1.0 + s) * std::exp(-s) * alpha[i];
    }
    return acc + p[0];
  }
};

// 27-column row + status {iterations, evaluations, termination, N}
inline void gp2d_row(const ObjView& v, double* row, int* status) {
  for (int k = 0; k < 27; ++k) row[k] = qnan();
  if (status) { status[0] = status[1] = status[2] = 0; status[3] = 0; }
  Gp gp;
  double tmin_all = pinf();
  std::vector<double> f, e;
  std::vector<int> rk;
  for (int b = 0; b < N_SEG; ++b)
    for (int i = v.off[b]; i < v.off[b + 1]; ++i) {
      tmin_all = std::fmin(tmin_all, v.t[i]);
      if (b < 6 && !is_nan(v.f[i]) && !is_nan(v.e[i]) && v.e[i] > 0.0) {
        gp.t.push_back(v.t[i]); gp.lam.push_back(WAVE[b]); f.push_back(v.f[i]); e.push_back(v.e[i]);
      }
    }
  gp.n = (int)f.size();
  if (status) status[3] = gp.n;
  if (gp.n < 10) return;
  double tmin = *std::min_element(gp.t.begin(), gp.t.end());
  for (auto& x : gp.t) x -= tmin;
  std::vector<double> nz;
  for (double x : f) if (x != 0.0) nz.push_back(std::fabs(x));
  std::sort(nz.begin(), nz.end());
  double scale = nz.empty() ? qnan() : sorted_median(nz.data(), (int)nz.size());
  if (scale == 0.0) scale = 1.0;
  gp.scale = scale;
  gp.y.resize(gp.n); gp.noise.resize(gp.n);
  double mean = 0.0;
  for (int i = 0; i < gp.n; ++i) { gp.y[i] = f[i] / scale; double ye = e[i] / scale; gp.noise[i] = ye * ye + TINY; mean += gp.y[i]; }
  mean /= gp.n;
  double var = 0.0;
  for (int i = 0; i < gp.n; ++i) { double d = gp.y[i] - mean; var += d * d; }
  var /= gp.n;
  gp.T = (gp.n + TB - 1) / TB;
  double p0[4] = {mean, std::log(var / 2.0), std::log(100.0 * 100.0), std::log(6000.0 * 6000.0)};
  Lbfgs4 o; o.start(p0);
  for (;;) { double fv, g[4]; gp.eval(o.x, fv, g); if (!o.advance(fv, g)) break; }
  if (status) { status[0] = o.nit; status[1] = o.nfev; status[2] = o.why; }
  const double* p = o.x;
  double amp = std::exp(p[0]), ts = std::sqrt(std::exp(p[1])), ws = std::sqrt(std::exp(p[2]));
  row[0] = amp; row[1] = ts; row[2] = ws; row[3] = -o.f; row[4] = ts / (ws / 1000.0);
  double fv, g[4];
  if (!gp.eval(p, fv, g)) { return; }                   // alpha at the optimum (george recomputes on predict)
  // peak time: first flux maximum of the r band in file order, else of all rows (multiband_gp.py:331-338)
  int pk = -1;
  {
    int lo = v.off[BR], hi = v.off[BR + 1];
    if (hi == lo) { lo = 0; hi = v.n; }
    double best = -pinf(); int brank = 1 << 30;
    for (int i = lo; i < hi; ++i) {
      if (is_nan(v.f[i])) continue;
      if (v.f[i] > best || (v.f[i] == best && v.rank[i] < brank)) { best = v.f[i]; brank = v.rank[i]; pk = i; }
    }
  }
  if (pk < 0) return;
  const double pt = v.t[pk] - tmin_all;
  const double ep[4] = {0.0, 20.0, 50.0, 100.0};
  const int bands[3] = {BG, BR, BI};
  double col[4][2];
  for (int ei = 0; ei < 4; ++ei) {
    double fl[3];
    for (int j = 0; j < 3; ++j) { fl[j] = gp.predict(p, pt + ep[ei], WAVE[bands[j]]) * scale; row[5 + 5 * ei + j] = fl[j]; }
    col[ei][0] = (fl[0] > 0.0 && fl[1] > 0.0) ? -2.5 * std::log10(fl[0] / fl[1]) : qnan();
    col[ei][1] = (fl[1] > 0.0 && fl[2] > 0.0) ? -2.5 * std::log10(fl[1] / fl[2]) : qnan();
    row[8 + 5 * ei] = col[ei][0]; row[9 + 5 * ei] = col[ei][1];
  }
  if (!is_nan(col[0][0]) && !is_nan(col[2][0])) row[25] = (col[2][0] - col[0][0]) / 50.0;
  if (!is_nan(col[0][0]) && !is_nan(col[3][0])) row[26] = (col[3][0] - col[0][0]) / 100.0;
}
}  // namespace gphost

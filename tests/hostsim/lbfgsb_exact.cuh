// lbfgsb_exact.cuh -- TEST-ONLY: L-BFGS-B with box constraints restated routine by routine (compact
// representation, formk/cmprlb/subsm as in the published code).  The product uses csrc/lbfgsb.cuh, which runs
// the same algorithm with the 3 x 3 quasi-Newton matrix formed explicitly; this file is the bridge that shows,
// on the host, that (a) the restated control flow reproduces SciPy's iteration and evaluation counts and
// (b) the dense variant follows it.  Never compiled into libmallorn_b200.so.
//
// The reference fits its per-band GPs with scikit-learn, whose optimiser is
// scipy.optimize.minimize(method='L-BFGS-B', jac=True, bounds=...) with SciPy's defaults
// (src/features/gaussian_process.py:98-106 -> sklearn _gpr.py _constrained_optimization): L-BFGS-B 3.0
// (Byrd, Lu, Nocedal & Zhu 1995; Zhu et al. 1997; Morales & Nocedal 2011), m = 10, factr = 1e7,
// pgtol = 1e-5, maxls = 20, maxiter = maxfun = 15000.  SciPy is a third-party dependency of the
// reference; this file restates that published algorithm for the case exercised there -- every
// variable has a lower and an upper bound (nbd = 2, "boxed") -- keeping the control flow of mainlb:
//   projected-gradient test -> generalised Cauchy point (cauchy, breakpoints taken smallest first)
//   -> free set -> subspace minimisation over the free variables with the compact middle matrix
//   (formk: LEL' factorisation; cmprlb; subsm with the projected-Newton safeguard of v3.0)
//   -> More-Thuente line search (dcsrch, lbfgs4.cuh) capped at the distance to the nearest bound
//   -> BFGS pair update (matupd, formt) or skip; memory refresh whenever a factorisation fails.
// The 2col x 2col matrix K of the subspace problem is formed from the stored pairs on every
// iteration instead of being patched incrementally (same numbers, n is tiny).
//
// Reverse communication, as Lbfgs4:
//     LbfgsBExact<3> o; o.start(x0, lo, hi);
//     for (;;) { eval(o.x, f, g); if (!o.advance(f, g)) break; }     // o.x, o.f = result
// Scalar code: one thread owns the state (about 7 KB; shared memory on the GPU).
#pragma once
#include "lbfgs4.cuh"

namespace mlrn {

template <int N>
struct LbfgsBExact {
  static constexpr int M = 10, M2 = 2 * M;
  static constexpr double EPSMCH = 2.220446049250313e-16;
  enum Why { RUNNING = 0, PGTOL = 1, FTOL = 2, MAXITER = 3, ABNORMAL = 4, MAXFUN = 5 };

  double x[N], f;                 // current trial point (evaluate here) / result
  int nit, nfev, why;
  // problem
  double lo[N], hi[N];
  // iterate state
  double g[N], d[N], z[N], r[N], t[N], xp[N];
  double fold, gd, gdold, stp, dtd, theta, sbgnrm;
  int iwhere[N], index[N], nfree;
  // limited-memory pairs (ring: oldest at head) and the middle matrices
  double ws[M][N], wy[M][N], sy[M][M], ss[M][M], wt[M][M];
  double wn[M2][M2];
  double p[M2], c[M2], wbp[M2], v[M2];
  int col, head, itail, iupdat;
  int ifun, phase;
  Dcsrch ls;

  MLRN_FN static double dotn(const double* a, const double* b) { double s = 0.0; for (int i = 0; i < N; ++i) s += a[i] * b[i]; return s; }

  // ---- LINPACK-style helpers on row-major storage A(i,j) = a[(o+i)*ld + (o+j)], upper triangle ----
  // Cholesky A = R'R, R upper, in place; returns 0 or the 1-based index of the failing pivot (dpofa)
  MLRN_NOINLINE static int chol_upper(double* a, int ld, int o, int n) {
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < j; ++k) {
        double tt = a[(o + k) * ld + o + j];
        for (int q = 0; q < k; ++q) tt -= a[(o + q) * ld + o + k] * a[(o + q) * ld + o + j];
        tt /= a[(o + k) * ld + o + k];
        a[(o + k) * ld + o + j] = tt;
        s += tt * tt;
      }
      s = a[(o + j) * ld + o + j] - s;
      if (!(s > 0.0)) return j + 1;
      a[(o + j) * ld + o + j] = sqrt(s);
    }
    return 0;
  }
  // solve T' x = b (T upper n x n at offset o); b strided by bs  (dtrsl job 11)
  MLRN_NOINLINE static int solve_ut(const double* a, int ld, int o, int n, double* b, int bs) {
    for (int j = 0; j < n; ++j) if (a[(o + j) * ld + o + j] == 0.0) return j + 1;
    b[0] /= a[o * ld + o];
    for (int j = 1; j < n; ++j) {
      double s = 0.0;
      for (int q = 0; q < j; ++q) s += a[(o + q) * ld + o + j] * b[q * bs];
      b[j * bs] = (b[j * bs] - s) / a[(o + j) * ld + o + j];
    }
    return 0;
  }
  // solve T x = b (T upper)  (dtrsl job 01)
  MLRN_NOINLINE static int solve_u(const double* a, int ld, int o, int n, double* b, int bs) {
    for (int j = 0; j < n; ++j) if (a[(o + j) * ld + o + j] == 0.0) return j + 1;
    b[(n - 1) * bs] /= a[(o + n - 1) * ld + o + n - 1];
    for (int j = n - 2; j >= 0; --j) {
      const double tmp = -b[(j + 1) * bs];
      for (int q = 0; q <= j; ++q) b[q * bs] += tmp * a[(o + q) * ld + o + j + 1];
      b[j * bs] /= a[(o + j) * ld + o + j];
    }
    return 0;
  }

  MLRN_FN void refresh() { col = 0; head = 0; theta = 1.0; iupdat = 0; }

  MLRN_FN void start(const double* x0, const double* lo_, const double* hi_) {
    for (int i = 0; i < N; ++i) {
      lo[i] = lo_[i]; hi[i] = hi_[i];
      x[i] = fmin(fmax(x0[i], lo[i]), hi[i]);                 // active(): project the start onto the box
      iwhere[i] = (hi[i] - lo[i] <= 0.0) ? 3 : 0;
    }
    nit = 0; nfev = 0; why = RUNNING; phase = 0; f = 0.0; itail = 0; nfree = N;
    refresh();
  }

  // infinity norm of the projected gradient (projgr)
  MLRN_FN void projgr() {
    double m = 0.0;
    for (int i = 0; i < N; ++i) {
      double gi = g[i];
      if (gi < 0.0) gi = fmax(x[i] - hi[i], gi); else gi = fmin(x[i] - lo[i], gi);
      m = fmax(m, fabs(gi));
    }
    sbgnrm = m;
  }

  // out = M v for the 2col x 2col middle matrix of the compact representation (bmv)
  MLRN_NOINLINE int bmv(const double* vin, double* out) {
    if (col == 0) return 0;
    out[col] = vin[col];
    for (int i = 1; i < col; ++i) {
      double s = 0.0;
      for (int k = 0; k < i; ++k) s += sy[i][k] * vin[k] / sy[k][k];
      out[col + i] = vin[col + i] + s;
    }
    if (solve_ut(&wt[0][0], M, 0, col, out + col, 1)) return 1;
    for (int i = 0; i < col; ++i) out[i] = vin[i] / sqrt(sy[i][i]);
    if (solve_u(&wt[0][0], M, 0, col, out + col, 1)) return 1;
    for (int i = 0; i < col; ++i) out[i] = -out[i] / sqrt(sy[i][i]);
    for (int i = 0; i < col; ++i) {
      double s = 0.0;
      for (int k = i + 1; k < col; ++k) s += sy[k][i] * out[col + k] / sy[i][i];
      out[i] += s;
    }
    return 0;
  }

  // heap maintenance of the breakpoints (hpsolb): after the call tb[n-1] is the smallest of tb[0..n)
  MLRN_NOINLINE static void hpsolb(int n, double* tb, int* io, int iheap) {
    if (iheap == 0) {
      for (int k = 2; k <= n; ++k) {
        double dd = tb[k - 1]; int in = io[k - 1]; int i = k;
        while (i > 1) {
          int j = i / 2;
          if (dd < tb[j - 1]) { tb[i - 1] = tb[j - 1]; io[i - 1] = io[j - 1]; i = j; } else break;
        }
        tb[i - 1] = dd; io[i - 1] = in;
      }
    }
    if (n > 1) {
      int i = 1;
      const double out = tb[0]; const int iout = io[0];
      const double dd = tb[n - 1]; const int in = io[n - 1];
      for (;;) {
        int j = i + i;
        if (j <= n - 1) {
          if (tb[j] < tb[j - 1]) j = j + 1;
          if (tb[j - 1] < dd) { tb[i - 1] = tb[j - 1]; io[i - 1] = io[j - 1]; i = j; } else break;
        } else break;
      }
      tb[i - 1] = dd; io[i - 1] = in;
      tb[n - 1] = out; io[n - 1] = iout;
    }
  }

  // generalised Cauchy point: z = xcp, c = W'(xcp - x), iwhere updated (cauchy)
  MLRN_NOINLINE int cauchy() {
    if (sbgnrm <= 0.0) { for (int i = 0; i < N; ++i) z[i] = x[i]; return 0; }
    bool bnded = true;
    int nfr = N, nbreak = 0, ibkmin = 0;          // iorder[nfr..N) = variables without a breakpoint
    int iorder[N]; double tb[N];
    double bkmin = 0.0, f1 = 0.0;
    const int col2 = 2 * col;
    for (int i = 0; i < col2; ++i) p[i] = 0.0;
    for (int i = 0; i < N; ++i) {
      const double neggi = -g[i];
      double tl = 0.0, tu = 0.0;
      if (iwhere[i] != 3 && iwhere[i] != -1) {
        tl = x[i] - lo[i]; tu = hi[i] - x[i];
        const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
        iwhere[i] = 0;
        if (xlower) { if (neggi <= 0.0) iwhere[i] = 1; }
        else if (xupper) { if (neggi >= 0.0) iwhere[i] = 2; }
        else if (fabs(neggi) <= 0.0) iwhere[i] = -3;
      }
      if (iwhere[i] != 0 && iwhere[i] != -1) d[i] = 0.0;
      else {
        d[i] = neggi;
        f1 -= neggi * neggi;
        int pt = head;
        for (int j = 0; j < col; ++j) {
          p[j] += wy[pt][i] * neggi;
          p[col + j] += ws[pt][i] * neggi;
          pt = (pt + 1) % M;
        }
        if (neggi < 0.0) {
          tb[nbreak] = tl / (-neggi); iorder[nbreak] = i; ++nbreak;
          if (nbreak == 1 || tb[nbreak - 1] < bkmin) { bkmin = tb[nbreak - 1]; ibkmin = nbreak - 1; }
        } else if (neggi > 0.0) {
          tb[nbreak] = tu / neggi; iorder[nbreak] = i; ++nbreak;
          if (nbreak == 1 || tb[nbreak - 1] < bkmin) { bkmin = tb[nbreak - 1]; ibkmin = nbreak - 1; }
        } else {
          --nfr;                                   // (unreachable with two-sided bounds: neggi == 0 gives iwhere = -3)
          if (fabs(neggi) > 0.0) bnded = false;
        }
      }
    }
    if (theta != 1.0) for (int j = 0; j < col; ++j) p[col + j] *= theta;
    for (int i = 0; i < N; ++i) z[i] = x[i];
    if (nbreak == 0 && nfr == N) return 0;
    for (int j = 0; j < col2; ++j) c[j] = 0.0;
    double f2 = -theta * f1;
    const double f2_org = f2;
    if (col > 0) {
      if (bmv(p, v)) return 1;
      double s = 0.0; for (int j = 0; j < col2; ++j) s += v[j] * p[j];
      f2 -= s;
    }
    double dtm = -f1 / f2, tsum = 0.0;
    bool all_fixed = false;
    if (nbreak > 0) {
      int nleft = nbreak, iter = 1;
      double tj = 0.0;
      for (;;) {
        const double tj0 = tj;
        int ibp;
        if (iter == 1) { tj = bkmin; ibp = iorder[ibkmin]; }
        else {
          if (iter == 2) {
            if (ibkmin != nbreak - 1) { tb[ibkmin] = tb[nbreak - 1]; iorder[ibkmin] = iorder[nbreak - 1]; }
          }
          hpsolb(nleft, tb, iorder, iter - 2);
          tj = tb[nleft - 1]; ibp = iorder[nleft - 1];
        }
        const double dt = tj - tj0;
        if (dtm < dt) break;                       // the minimiser lies inside this segment
        tsum += dt; --nleft; ++iter;
        const double dibp = d[ibp];
        d[ibp] = 0.0;
        double zibp;
        if (dibp > 0.0) { zibp = hi[ibp] - x[ibp]; z[ibp] = hi[ibp]; iwhere[ibp] = 2; }
        else { zibp = lo[ibp] - x[ibp]; z[ibp] = lo[ibp]; iwhere[ibp] = 1; }
        if (nleft == 0 && nbreak == N) { dtm = dt; all_fixed = true; break; }
        const double dibp2 = dibp * dibp;
        f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
        f2 = f2 - theta * dibp2;
        if (col > 0) {
          for (int j = 0; j < col2; ++j) c[j] += dt * p[j];
          int pt = head;
          for (int j = 0; j < col; ++j) { wbp[j] = wy[pt][ibp]; wbp[col + j] = theta * ws[pt][ibp]; pt = (pt + 1) % M; }
          if (bmv(wbp, v)) return 1;
          double wmc = 0.0, wmp = 0.0, wmw = 0.0;
          for (int j = 0; j < col2; ++j) { wmc += c[j] * v[j]; wmp += p[j] * v[j]; wmw += wbp[j] * v[j]; }
          for (int j = 0; j < col2; ++j) p[j] -= dibp * wbp[j];
          f1 += dibp * wmc;
          f2 += 2.0 * dibp * wmp - dibp2 * wmw;
        }
        f2 = fmax(EPSMCH * f2_org, f2);
        if (nleft > 0) { dtm = -f1 / f2; continue; }
        if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
        else dtm = -f1 / f2;
        break;
      }
    }
    if (!all_fixed) {
      if (dtm <= 0.0) dtm = 0.0;
      tsum += dtm;
      for (int i = 0; i < N; ++i) z[i] += tsum * d[i];
    }
    if (col > 0) for (int j = 0; j < col2; ++j) c[j] += dtm * p[j];
    return 0;
  }

  // LEL' factorisation of the subspace matrix K into the upper triangle of wn (formk)
  MLRN_NOINLINE int formk() {
    bool fr[N];
    for (int i = 0; i < N; ++i) fr[i] = iwhere[i] <= 0;
    for (int iy = 0; iy < col; ++iy) {
      const int pi = (head + iy) % M, is = col + iy;
      for (int jy = 0; jy < col; ++jy) {
        const int pj = (head + jy) % M, js = col + jy;
        double yy_free = 0.0, ss_act = 0.0, sy_free = 0.0, sy_act = 0.0;
        for (int k = 0; k < N; ++k) {
          if (fr[k]) { yy_free += wy[pi][k] * wy[pj][k]; sy_free += ws[pi][k] * wy[pj][k]; }
          else { ss_act += ws[pi][k] * ws[pj][k]; sy_act += ws[pi][k] * wy[pj][k]; }
        }
        if (jy <= iy) { wn[jy][iy] = yy_free / theta; wn[js][is] = ss_act * theta; }
        wn[jy][is] = (jy < iy) ? -sy_act : sy_free;        // -L_a below the diagonal, R_z on and above
      }
      wn[iy][iy] += sy[iy][iy];
    }
    double* w = &wn[0][0];
    if (chol_upper(w, M2, 0, col)) return -1;
    for (int js = col; js < 2 * col; ++js) if (solve_ut(w, M2, 0, col, w + js, M2)) return -1;
    for (int is = col; is < 2 * col; ++is)
      for (int js = is; js < 2 * col; ++js) {
        double s = 0.0;
        for (int q = 0; q < col; ++q) s += wn[q][is] * wn[q][js];
        wn[is][js] += s;
      }
    if (chol_upper(w, M2, col, col)) return -2;
    return 0;
  }

  // r = -Z'(B(xcp - x) + g) over the free variables (cmprlb)
  MLRN_NOINLINE int cmprlb() {
    for (int i = 0; i < nfree; ++i) { const int k = index[i]; r[i] = -theta * (z[k] - x[k]) - g[k]; }
    if (bmv(c, p)) return -8;
    int pt = head;
    for (int j = 0; j < col; ++j) {
      const double a1 = p[j], a2 = theta * p[col + j];
      for (int i = 0; i < nfree; ++i) { const int k = index[i]; r[i] += wy[pt][k] * a1 + ws[pt][k] * a2; }
      pt = (pt + 1) % M;
    }
    return 0;
  }

  // subspace minimisation: z moves from the Cauchy point to the (projected) subspace Newton point (subsm)
  MLRN_NOINLINE int subsm() {
    if (nfree <= 0) return 0;
    const int col2 = 2 * col;
    double* dd = r;                                // the reduced gradient becomes the Newton direction, in place
    int pt = head;
    for (int i = 0; i < col; ++i) {
      double t1 = 0.0, t2 = 0.0;
      for (int j = 0; j < nfree; ++j) { const int k = index[j]; t1 += wy[pt][k] * dd[j]; t2 += ws[pt][k] * dd[j]; }
      v[i] = t1; v[col + i] = theta * t2;
      pt = (pt + 1) % M;
    }
    double* w = &wn[0][0];
    if (solve_ut(w, M2, 0, col2, v, 1)) return 1;
    for (int i = 0; i < col; ++i) v[i] = -v[i];
    if (solve_u(w, M2, 0, col2, v, 1)) return 1;
    pt = head;
    for (int jy = 0; jy < col; ++jy) {
      const int js = col + jy;
      for (int i = 0; i < nfree; ++i) { const int k = index[i]; dd[i] += wy[pt][k] * v[jy] / theta + ws[pt][k] * v[js]; }
      pt = (pt + 1) % M;
    }
    for (int i = 0; i < nfree; ++i) dd[i] *= 1.0 / theta;
    // projected Newton point; if it leaves the box in an ascent direction fall back to the truncated step
    bool clipped = false;
    for (int i = 0; i < N; ++i) xp[i] = z[i];
    for (int i = 0; i < nfree; ++i) {
      const int k = index[i];
      double xk = fmax(lo[k], z[k] + dd[i]);
      z[k] = fmin(hi[k], xk);
      if (z[k] == lo[k] || z[k] == hi[k]) clipped = true;
    }
    if (!clipped) return 0;
    double dd_p = 0.0;
    for (int i = 0; i < N; ++i) dd_p += (z[i] - x[i]) * g[i];
    if (dd_p > 0.0) {
      for (int i = 0; i < N; ++i) z[i] = xp[i];
      double alpha = 1.0, temp1 = alpha; int ibd = -1;
      for (int i = 0; i < nfree; ++i) {
        const int k = index[i]; const double dk = dd[i];
        if (dk < 0.0) { const double t2 = lo[k] - z[k]; if (t2 >= 0.0) temp1 = 0.0; else if (dk * alpha < t2) temp1 = t2 / dk; }
        else if (dk > 0.0) { const double t2 = hi[k] - z[k]; if (t2 <= 0.0) temp1 = 0.0; else if (dk * alpha > t2) temp1 = t2 / dk; }
        if (temp1 < alpha) { alpha = temp1; ibd = i; }
      }
      if (alpha < 1.0 && ibd >= 0) {
        const double dk = dd[ibd]; const int k = index[ibd];
        if (dk > 0.0) { z[k] = hi[k]; dd[ibd] = 0.0; }
        else if (dk < 0.0) { z[k] = lo[k]; dd[ibd] = 0.0; }
      }
      for (int i = 0; i < nfree; ++i) z[index[i]] += alpha * dd[i];
    }
    return 0;
  }

  // One pass of the main loop from the current iterate up to the first trial point of its line search.
  MLRN_NOINLINE bool begin_iteration() {
    for (;;) {
      if (cauchy()) { refresh(); continue; }
      nfree = 0; { int iact = N; for (int i = 0; i < N; ++i) { if (iwhere[i] <= 0) index[nfree++] = i; else index[--iact] = i; } }
      if (nfree != 0 && col != 0) {
        int info = formk();
        if (!info) info = cmprlb();
        if (!info) info = subsm();
        if (info) { refresh(); continue; }
      }
      for (int i = 0; i < N; ++i) d[i] = z[i] - x[i];
      // lnsrlb, first entry
      dtd = dotn(d, d);
      double stpmx = 1e10;
      if (nit == 0) stpmx = 1.0;
      else
        for (int i = 0; i < N; ++i) {
          const double a1 = d[i];
          if (a1 < 0.0) { const double a2 = lo[i] - x[i]; if (a2 >= 0.0) stpmx = 0.0; else if (a1 * stpmx < a2) stpmx = a2 / a1; }
          else if (a1 > 0.0) { const double a2 = hi[i] - x[i]; if (a2 <= 0.0) stpmx = 0.0; else if (a1 * stpmx > a2) stpmx = a2 / a1; }
        }
      stp = 1.0;                                   // boxed problem: never 1/|d|
      for (int i = 0; i < N; ++i) { t[i] = x[i]; r[i] = g[i]; }
      fold = f; ifun = 0;
      gd = dotn(g, d); gdold = gd;
      int task = (gd >= 0.0 || !(gd == gd)) ? (int)Dcsrch::ERROR : ls.start(f, gd, stp, 1e-3, 0.9, 0.1, 0.0, stpmx);
      if (task == Dcsrch::FG) { trial_point(); phase = 1; return true; }
      if (col == 0) { why = ABNORMAL; return false; }
      refresh();
    }
  }

  MLRN_FN void trial_point() {
    if (stp == 1.0) for (int i = 0; i < N; ++i) x[i] = z[i];
    else for (int i = 0; i < N; ++i) x[i] = stp * d[i] + t[i];
  }

  // BFGS pair (s = d, y = r) into the ring and into S'Y / S'S (matupd), then T = theta S'S + L D^-1 L' (formt)
  MLRN_NOINLINE int update_pairs(double rr, double dr) {
    ++iupdat;
    if (iupdat <= M) { col = iupdat; itail = (head + iupdat - 1) % M; }
    else { itail = (itail + 1) % M; head = (head + 1) % M; }
    for (int i = 0; i < N; ++i) { ws[itail][i] = d[i]; wy[itail][i] = r[i]; }
    theta = rr / dr;
    if (iupdat > M)
      for (int j = 0; j < col - 1; ++j) {
        for (int k = 0; k <= j; ++k) ss[k][j] = ss[k + 1][j + 1];
        for (int k = 0; k < col - 1 - j; ++k) sy[j + k][j] = sy[j + 1 + k][j + 1];
      }
    int pt = head;
    for (int j = 0; j < col - 1; ++j) {
      sy[col - 1][j] = dotn(d, wy[pt]);
      ss[j][col - 1] = dotn(ws[pt], d);
      pt = (pt + 1) % M;
    }
    ss[col - 1][col - 1] = (stp == 1.0) ? dtd : stp * stp * dtd;
    sy[col - 1][col - 1] = dr;
    // formt
    for (int j = 0; j < col; ++j) wt[0][j] = theta * ss[0][j];
    for (int i = 1; i < col; ++i)
      for (int j = i; j < col; ++j) {
        const int k1 = (i < j ? i : j);
        double s = 0.0;
        for (int k = 0; k < k1; ++k) s += sy[i][k] * sy[j][k] / sy[k][k];
        wt[i][j] = s + theta * ss[i][j];
      }
    return chol_upper(&wt[0][0], M, 0, col);
  }

  // Feed f(x), grad f(x) of the current o.x.  Returns true if another evaluation at o.x is wanted.
  MLRN_FN bool advance(double fx_, const double* gx_) {
    ++nfev;
    f = fx_;
    for (int i = 0; i < N; ++i) g[i] = gx_[i];
    if (phase == 0) {
      projgr();
      if (sbgnrm <= 1e-5) { why = PGTOL; return false; }
      return begin_iteration();
    }
    ++ifun;
    gd = dotn(g, d);
    const int task = ls.step(f, gd, stp);
    if (task == Dcsrch::FG) {
      if (ifun >= 20) {                              // maxls trial points without success
        for (int i = 0; i < N; ++i) { x[i] = t[i]; g[i] = r[i]; }
        f = fold;
        if (col == 0) { why = ABNORMAL; return false; }
        refresh();
        return begin_iteration();
      }
      trial_point();
      return true;
    }
    // line search finished: NEW_X
    ++nit;
    projgr();
    if (nit >= 15000) { why = MAXITER; return false; }
    if (nfev > 15000) { why = MAXFUN; return false; }
    if (sbgnrm <= 1e-5) { why = PGTOL; return false; }
    {
      const double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
      if ((fold - f) <= EPSMCH * 1e7 * ddum) { why = FTOL; return false; }
    }
    for (int i = 0; i < N; ++i) r[i] = g[i] - r[i];
    const double rr = dotn(r, r);
    double dr, ddum;
    if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
    else { dr = (gd - gdold) * stp; for (int i = 0; i < N; ++i) d[i] *= stp; ddum = -gdold * stp; }
    if (!(dr <= EPSMCH * ddum)) {
      if (update_pairs(rr, dr)) refresh();
    }
    return begin_iteration();
  }
};

}  // namespace mlrn

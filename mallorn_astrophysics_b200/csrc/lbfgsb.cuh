// lbfgsb.cuh -- L-BFGS-B with box constraints for a handful of variables (the hyper-parameter
// search of the per-band 1-D GP, gp1d.cuh), with the quasi-Newton matrix held explicitly.
//
// The reference fits its per-band GPs with scikit-learn, whose optimiser is
// scipy.optimize.minimize(method='L-BFGS-B', jac=True, bounds=...) with SciPy's defaults
// (src/features/gaussian_process.py:98-106 -> sklearn _gpr.py _constrained_optimization): L-BFGS-B 3.0
// (Byrd, Lu, Nocedal & Zhu 1995; Zhu et al. 1997; Morales & Nocedal 2011), m = 10, factr = 1e7,
// pgtol = 1e-5, maxls = 20, maxiter = maxfun = 15000.  SciPy is a third-party dependency of the
// reference; this file restates that published algorithm for the case exercised there -- every
// variable has a lower and an upper bound ("boxed") -- with the same control flow:
//   projected-gradient test -> generalised Cauchy point (breakpoints taken smallest first)
//   -> free set -> subspace minimisation over the free variables, projected-Newton safeguard of v3.0
//   -> More-Thuente line search (dcsrch, lbfgs4.cuh) capped at the distance to the nearest bound
//   -> BFGS pair kept (or skipped when s'y <= eps |g'd|), memory dropped whenever the model stops
//   being positive definite or a line search fails.
//
// What differs from the published code is only the linear algebra.  The limited-memory matrix
// B = theta I - W M W' of L-BFGS-B is, by construction (Byrd, Nocedal & Schnabel 1994, thm 2.3), the result
// of applying the stored BFGS pairs, oldest first, to theta I.  For n = 3 that matrix is 3 x 3: it is
// rebuilt from the (at most 10) pairs at the top of every iteration -- ~400 flops -- and the Cauchy search
// and the subspace step use it directly (f' = g'd + (xc - x)'B d, f'' = d'B d per segment; B_ff dx = -r_f by a
// 3 x 3 Cholesky) instead of the 2m x 2m compact-form factorisations, which cost ~100x more serial work
// per iteration on one lane and 7 KB of state.  Same iterates up to rounding: tests/hostsim compares this
// file with a routine-by-routine restatement (tests/hostsim/lbfgsb_exact.cuh) and with SciPy itself.
//
// Reverse communication, as Lbfgs4:
//     LbfgsB<3> o; o.start(x0, lo, hi);
//     for (;;) { eval(o.x, f, g); if (!o.advance(f, g)) break; }     // o.x, o.f = result
// Scalar code: one thread owns the state (< 1 KB).
#pragma once
#include "lbfgs4.cuh"

namespace mlrn {

template <int N>
struct LbfgsB {
  static constexpr int M = 10;
  static constexpr double EPSMCH = 2.220446049250313e-16;
  enum Why { RUNNING = 0, PGTOL = 1, FTOL = 2, MAXITER = 3, ABNORMAL = 4, MAXFUN = 5 };

  double x[N], f;                 // current trial point (evaluate here) / result
  int nit, nfev, why;
  double lo[N], hi[N];
  double g[N], d[N], z[N], r[N], t[N];
  double fold, gd, gdold, stp, dtd, theta, sbgnrm;
  double B[N][N];
  double ws[M][N], wy[M][N], ys[M];       // ring of pairs, oldest at head; ys = s'y as computed by the update
  int iwhere[N], index[N], nfree;
  int col, head, iupdat;
  int ifun, phase;
  Dcsrch ls;

  MLRN_FN static double dotn(const double* a, const double* b) { double s = 0.0; for (int i = 0; i < N; ++i) s += a[i] * b[i]; return s; }

  MLRN_FN void refresh() { col = 0; head = 0; theta = 1.0; iupdat = 0; }

  MLRN_FN void start(const double* x0, const double* lo_, const double* hi_) {
    for (int i = 0; i < N; ++i) {
      lo[i] = lo_[i]; hi[i] = hi_[i];
      x[i] = fmin(fmax(x0[i], lo[i]), hi[i]);                 // active(): project the start onto the box
      iwhere[i] = (hi[i] - lo[i] <= 0.0) ? 3 : 0;
    }
    nit = 0; nfev = 0; why = RUNNING; phase = 0; f = 0.0; nfree = N;
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

  // B = BFGS updates of theta I with the stored pairs, oldest first.  Returns false if the model lost
  // positive definiteness (the published code meets that as a failed factorisation of T or K).
  MLRN_NOINLINE bool build_model() {
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) B[i][j] = (i == j) ? theta : 0.0;
    MLRN_ROLL
    for (int k = 0; k < col; ++k) {
      const int pt = (head + k) % M;
      double bs[N];
      for (int i = 0; i < N; ++i) bs[i] = dotn(B[i], ws[pt]);
      const double sbs = dotn(bs, ws[pt]);
      if (!(sbs > 0.0) || !(ys[pt] > 0.0)) return false;
      const double ia = 1.0 / sbs, ib = 1.0 / ys[pt];
      for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) {
          const double v = B[i][j] - (bs[i] * bs[j]) * ia + (wy[pt][i] * wy[pt][j]) * ib;
          B[i][j] = v; B[j][i] = v;
        }
    }
    return true;
  }

  // generalised Cauchy point along the projected steepest-descent path: z = xcp, iwhere updated (cauchy)
  MLRN_NOINLINE void cauchy() {
    for (int i = 0; i < N; ++i) z[i] = x[i];
    if (sbgnrm <= 0.0) return;
    int nbreak = 0, iorder[N]; double tb[N];
    double f1 = 0.0;
    for (int i = 0; i < N; ++i) {
      const double neggi = -g[i];
      double tl = 0.0, tu = 0.0;
      if (iwhere[i] != 3) {
        tl = x[i] - lo[i]; tu = hi[i] - x[i];
        iwhere[i] = 0;
        if (tl <= 0.0) { if (neggi <= 0.0) iwhere[i] = 1; }
        else if (tu <= 0.0) { if (neggi >= 0.0) iwhere[i] = 2; }
        else if (fabs(neggi) <= 0.0) iwhere[i] = -3;
      }
      if (iwhere[i] != 0) d[i] = 0.0;
      else {
        d[i] = neggi;
        f1 -= neggi * neggi;
        tb[nbreak] = (neggi < 0.0) ? tl / (-neggi) : tu / neggi;        // neggi != 0 here
        iorder[nbreak] = i; ++nbreak;
      }
    }
    if (nbreak == 0) return;
    const double f2_org = -theta * f1;
    double bd[N];
    for (int i = 0; i < N; ++i) bd[i] = dotn(B[i], d);
    double f2 = dotn(d, bd);
    double dtm = -f1 / f2, tsum = 0.0, tj = 0.0;
    int nleft = nbreak;
    bool all_fixed = false;
    while (nleft > 0) {
      // smallest remaining breakpoint (ties: the one met first, as the heap of the published code would
      // return for so few entries up to the order of exact ties)
      int q = 0;
      for (int k = 1; k < nleft; ++k) if (tb[k] < tb[q]) q = k;
      const double tj0 = tj;
      tj = tb[q];
      const int ibp = iorder[q];
      const double dt = tj - tj0;
      if (dtm < dt) break;                         // the minimiser lies inside this segment
      tb[q] = tb[nleft - 1]; iorder[q] = iorder[nleft - 1]; --nleft;
      tsum += dt;
      const double dibp = d[ibp];
      d[ibp] = 0.0;
      if (dibp > 0.0) { z[ibp] = hi[ibp]; iwhere[ibp] = 2; }
      else { z[ibp] = lo[ibp]; iwhere[ibp] = 1; }
      if (nleft == 0 && nbreak == N) { dtm = dt; all_fixed = true; break; }
      // slope and curvature of the model on the next segment, from the point reached so far
      double xc[N];
      for (int i = 0; i < N; ++i) xc[i] = (d[i] != 0.0) ? tsum * (-g[i]) : z[i] - x[i];
      for (int i = 0; i < N; ++i) bd[i] = dotn(B[i], d);
      f1 = dotn(g, d) + dotn(xc, bd);
      f2 = fmax(EPSMCH * f2_org, dotn(d, bd));
      if (nleft > 0) dtm = -f1 / f2;
      else { f1 = 0.0; f2 = 0.0; dtm = 0.0; }      // every remaining direction is bounded and used up
    }
    if (!all_fixed) {
      if (dtm <= 0.0) dtm = 0.0;
      tsum += dtm;
      for (int i = 0; i < N; ++i) z[i] += tsum * d[i];
    }
  }

  // subspace minimisation over the free variables: z moves from the Cauchy point to the (projected) Newton point
  // of the model (cmprlb + subsm).  Returns false if B_ff is not positive definite.
  MLRN_NOINLINE bool subspace() {
    double dz[N], rf[N], L[N][N], dd[N], xp[N];
    for (int i = 0; i < N; ++i) dz[i] = z[i] - x[i];
    for (int a = 0; a < nfree; ++a) { const int k = index[a]; rf[a] = -(dotn(B[k], dz) + g[k]); }
    // Cholesky of B_ff, forward and back substitution
    for (int a = 0; a < nfree; ++a)
      for (int b = 0; b <= a; ++b) {
        double s = B[index[a]][index[b]];
        for (int q = 0; q < b; ++q) s -= L[a][q] * L[b][q];
        if (a == b) { if (!(s > 0.0)) return false; L[a][a] = sqrt(s); }
        else L[a][b] = s / L[b][b];
      }
    for (int a = 0; a < nfree; ++a) { double s = rf[a]; for (int q = 0; q < a; ++q) s -= L[a][q] * dd[q]; dd[a] = s / L[a][a]; }
    for (int a = nfree - 1; a >= 0; --a) { double s = dd[a]; for (int q = a + 1; q < nfree; ++q) s -= L[q][a] * dd[q]; dd[a] = s / L[a][a]; }
    // projected Newton point; if it leaves the box in an ascent direction fall back to the truncated step
    bool clipped = false;
    for (int i = 0; i < N; ++i) xp[i] = z[i];
    for (int a = 0; a < nfree; ++a) {
      const int k = index[a];
      const double xk = fmax(lo[k], z[k] + dd[a]);
      z[k] = fmin(hi[k], xk);
      if (z[k] == lo[k] || z[k] == hi[k]) clipped = true;
    }
    if (!clipped) return true;
    double dd_p = 0.0;
    for (int i = 0; i < N; ++i) dd_p += (z[i] - x[i]) * g[i];
    if (dd_p > 0.0) {
      for (int i = 0; i < N; ++i) z[i] = xp[i];
      double alpha = 1.0, temp1 = alpha; int ibd = -1;
      for (int a = 0; a < nfree; ++a) {
        const int k = index[a]; const double dk = dd[a];
        if (dk < 0.0) { const double t2 = lo[k] - z[k]; if (t2 >= 0.0) temp1 = 0.0; else if (dk * alpha < t2) temp1 = t2 / dk; }
        else if (dk > 0.0) { const double t2 = hi[k] - z[k]; if (t2 <= 0.0) temp1 = 0.0; else if (dk * alpha > t2) temp1 = t2 / dk; }
        if (temp1 < alpha) { alpha = temp1; ibd = a; }
      }
      if (alpha < 1.0 && ibd >= 0) {
        const double dk = dd[ibd]; const int k = index[ibd];
        if (dk > 0.0) { z[k] = hi[k]; dd[ibd] = 0.0; }
        else if (dk < 0.0) { z[k] = lo[k]; dd[ibd] = 0.0; }
      }
      for (int a = 0; a < nfree; ++a) z[index[a]] += alpha * dd[a];
    }
    return true;
  }

  // One pass of the main loop from the current iterate up to the first trial point of its line search.
  MLRN_NOINLINE bool begin_iteration() {
    for (;;) {
      if (!build_model()) { refresh(); continue; }
      cauchy();
      nfree = 0; { int iact = N; for (int i = 0; i < N; ++i) { if (iwhere[i] <= 0) index[nfree++] = i; else index[--iact] = i; } }
      if (nfree != 0 && col != 0) {
        if (!subspace()) { refresh(); continue; }
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

  // Feed f(x), grad f(x) of the current o.x.  Returns true if another evaluation at o.x is wanted.
  MLRN_NOINLINE bool advance(double fx_, const double* gx_) {
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
    if (!(dr <= EPSMCH * ddum)) {                    // keep the pair (matupd)
      ++iupdat;
      int slot;
      if (iupdat <= M) { col = iupdat; slot = (head + iupdat - 1) % M; }
      else { slot = head; head = (head + 1) % M; }
      for (int i = 0; i < N; ++i) { ws[slot][i] = d[i]; wy[slot][i] = r[i]; }
      ys[slot] = dr;
      theta = rr / dr;
    }
    return begin_iteration();
  }
};

}  // namespace mlrn

// lbfgs4.cuh -- the hyper-parameter optimiser of the 2-D GP: L-BFGS-B without bounds, n = 4.
//
// The reference calls scipy.optimize.minimize(method='L-BFGS-B', options={'maxiter': 100})
// (src/features/multiband_gp.py:158-164), i.e. L-BFGS-B 3.0 (Zhu, Byrd, Lu, Nocedal; Morales &
// Nocedal 2011) with m = 10, factr = 1e7, pgtol = 1e-5, maxls = 20, and the MINPACK-2 line search
// dcsrch/dcstep (More & Thuente 1994; ftol 1e-3, gtol 0.9, xtol 0.1).  SciPy is a third-party
// dependency of the reference; this file restates that published algorithm for the case the
// reference exercises -- no bounds -- where the generalised Cauchy point plus subspace minimisation
// of L-BFGS-B reduce to the quasi-Newton step -B^-1 g, evaluated here by the two-loop recursion with
// H0 = (s'y / y'y) I.  Control flow (first step 1/|d|, skipped updates, memory refresh after a failed
// search, the two convergence tests and their order, the iteration cap applied before them) follows
// mainlb/lnsrlb and SciPy's driver loop so that iterates track the reference's to rounding
// (tests: iterations and evaluations agree with SciPy on every synthetic object tried).
//
// Reverse communication: the caller owns the objective.  Usage:
//     Lbfgs4 o; o.start(x0);
//     for (;;) { eval(o.x, f, g); if (!o.advance(f, g)) break; }     // o.x, o.f = result
// Scalar code; on the GPU every thread of the CTA runs it redundantly on broadcast (f, g).
#pragma once
#include "team.cuh"

namespace mlrn {

struct Dcsrch {
  double ftol, gtol, xtol, stpmin, stpmax;
  bool brackt; int stage;
  double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;

  enum Task { FG = 0, CONV = 1, WARN = 2, ERROR = 3 };

  MLRN_FN static void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                             double fp, double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf;
    if (fp > fx) {
      double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
      double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
      if (stp < stx) gamma = -gamma;
      double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
      double stpc = stx + r * (stp - stx);
      double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
      stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
      brackt = true;
    } else if (sgnd < 0.0) {
      double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
      double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
      if (stp > stx) gamma = -gamma;
      double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
      double stpc = stp + r * (stx - stp);
      double stpq = stp + (dp / (dp - dx)) * (stx - stp);
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
      double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
      double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
      double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
      if (stp > stx) gamma = -gamma;
      double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
      double stpc;
      if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
      else if (stp > stx) stpc = stpmax;
      else stpc = stpmin;
      double stpq = stp + (dp / (dp - dx)) * (stx - stp);
      if (brackt) {
        stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
        if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
        else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
      } else {
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        stpf = fmin(stpmax, stpf); stpf = fmax(stpmin, stpf);
      }
    } else {
      if (brackt) {
        double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
        double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
        if (stp > sty) gamma = -gamma;
        double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
        stpf = stp + r * (sty - stp);
      } else if (stp > stx) stpf = stpmax;
      else stpf = stpmin;
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
      if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
      stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
  }

  MLRN_FN int start(double f, double g, double& stp, double ftol_, double gtol_, double xtol_, double stpmin_, double stpmax_) {
    ftol = ftol_; gtol = gtol_; xtol = xtol_; stpmin = stpmin_; stpmax = stpmax_;
    if (stp < stpmin || stp > stpmax || !(g < 0.0)) return ERROR;
    brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * g;
    width = stpmax - stpmin; width1 = width / 0.5;
    stx = 0.0; fx = f; gx = g; sty = 0.0; fy = f; gy = g;
    stmin = 0.0; stmax = stp + 4.0 * stp;
    return FG;
  }

  MLRN_FN int step(double f, double g, double& stp) {
    const double ftest = finit + stp * gtest;
    if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
    int task = FG;
    if (brackt && (stp <= stmin || stp >= stmax)) task = WARN;
    if (brackt && stmax - stmin <= xtol * stmax) task = WARN;
    if (stp == stpmax && f <= ftest && g <= gtest) task = WARN;
    if (stp == stpmin && (f > ftest || g >= gtest)) task = WARN;
    if (f <= ftest && fabs(g) <= gtol * (-ginit)) task = CONV;
    if (task != FG) return task;
    if (stage == 1 && f <= fx && f > ftest) {
      double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
      double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
      dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
      fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
    } else {
      dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
    }
    if (brackt) {
      if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
      width1 = width; width = fabs(sty - stx);
    }
    if (brackt) { stmin = fmin(stx, sty); stmax = fmax(stx, sty); }
    else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
    stp = fmax(stp, stpmin); stp = fmin(stp, stpmax);
    if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
    return FG;
  }
};

struct Lbfgs4 {
  static constexpr int N = 4, M = 10;
  static constexpr double EPSMCH = 2.220446049250313e-16;
  enum Why { RUNNING = 0, PGTOL = 1, FTOL = 2, MAXITER = 3, ABNORMAL = 4, MAXFUN = 5 };

  double x[N], f;                 // current trial point (evaluate here) / result
  int nit, nfev, why;
  // state
  double g[N], d[N], xold[N], gold[N], fold, gd, gdold, stp, theta;
  double S[M][N], Y[M][N]; int col, head;        // ring buffer: oldest at head
  int phase;                                     // 0 = waiting for f(x0); 1 = inside a line search
  int ifun;
  Dcsrch ls;

  MLRN_FN static double dot(const double* a, const double* b) { double s = 0.0; for (int i = 0; i < N; ++i) s += a[i] * b[i]; return s; }
  MLRN_FN static double amax(const double* a) { double m = 0.0; for (int i = 0; i < N; ++i) m = fmax(m, fabs(a[i])); return m; }

  MLRN_FN void start(const double* x0) {
    for (int i = 0; i < N; ++i) x[i] = x0[i];
    nit = 0; nfev = 0; why = RUNNING; col = 0; head = 0; theta = 1.0; phase = 0; f = 0.0;
  }

  // quasi-Newton direction d = -H g (two-loop recursion over the stored pairs, oldest -> newest)
  MLRN_FN void direction() {
    if (col == 0) { for (int i = 0; i < N; ++i) d[i] = -g[i]; return; }
    double q[N], al[M], rho[M];
    for (int i = 0; i < N; ++i) q[i] = g[i];
    for (int k = col - 1; k >= 0; --k) {
      const int j = (head + k) % M;
      rho[k] = 1.0 / dot(Y[j], S[j]);
      al[k] = rho[k] * dot(S[j], q);
      for (int i = 0; i < N; ++i) q[i] -= al[k] * Y[j][i];
    }
    for (int i = 0; i < N; ++i) q[i] /= theta;
    for (int k = 0; k < col; ++k) {
      const int j = (head + k) % M;
      double b = rho[k] * dot(Y[j], q);
      for (int i = 0; i < N; ++i) q[i] += (al[k] - b) * S[j][i];
    }
    for (int i = 0; i < N; ++i) d[i] = -q[i];
  }

  // Begin the line search of a new iteration from (x = xold, f, g).  Returns false when finished.
  MLRN_FN bool begin_iteration() {
    for (;;) {
      direction();
      const double dnorm = sqrt(dot(d, d));
      const double stpmx = 1e10;
      stp = (nit == 0) ? fmin(1.0 / dnorm, stpmx) : 1.0;
      for (int i = 0; i < N; ++i) { xold[i] = x[i]; gold[i] = g[i]; }
      fold = f;
      gd = dot(g, d); gdold = gd;
      ifun = 0;
      int task = (gd >= 0.0 || !(gd == gd)) ? Dcsrch::ERROR : ls.start(f, gd, stp, 1e-3, 0.9, 0.1, 0.0, stpmx);
      if (task == Dcsrch::FG) { trial_point(); phase = 1; return true; }
      // ascent / degenerate direction: refresh the memory and retry, or give up when it is already empty
      if (col == 0) { why = ABNORMAL; return false; }
      col = 0; head = 0; theta = 1.0;
    }
  }

  MLRN_FN void trial_point() {
    if (stp == 1.0) for (int i = 0; i < N; ++i) x[i] = xold[i] + d[i];
    else for (int i = 0; i < N; ++i) x[i] = stp * d[i] + xold[i];
  }

  // Feed f(x), grad f(x) of the current o.x.  Returns true if another evaluation at o.x is wanted.
  MLRN_FN bool advance(double fx_, const double* gx_) {
    ++nfev;
    if (phase == 0) {
      f = fx_;
      for (int i = 0; i < N; ++i) g[i] = gx_[i];
      if (amax(g) <= 1e-5) { why = PGTOL; return false; }
      return begin_iteration();
    }
    // inside a line search
    ++ifun;
    f = fx_;
    for (int i = 0; i < N; ++i) g[i] = gx_[i];
    gd = dot(g, d);
    int task = ls.step(f, gd, stp);
    if (task == Dcsrch::FG) {
      if (ifun >= 20) {                                  // maxls evaluations without success
        for (int i = 0; i < N; ++i) { x[i] = xold[i]; g[i] = gold[i]; }
        f = fold;
        if (col == 0) { why = ABNORMAL; return false; }
        col = 0; head = 0; theta = 1.0;
        return begin_iteration();
      }
      trial_point();
      return true;
    }
    // line search finished: NEW_X
    ++nit;
    if (nit >= 100) { why = MAXITER; return false; }      // SciPy's driver checks maxiter before the tests
    if (nfev > 15000) { why = MAXFUN; return false; }
    if (amax(g) <= 1e-5) { why = PGTOL; return false; }
    {
      const double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
      if ((fold - f) <= EPSMCH * 1e7 * ddum) { why = FTOL; return false; }
    }
    double r[N], s[N], dr, ddum;
    for (int i = 0; i < N; ++i) r[i] = g[i] - gold[i];
    if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; for (int i = 0; i < N; ++i) s[i] = d[i]; }
    else { dr = (gd - gdold) * stp; ddum = -gdold * stp; for (int i = 0; i < N; ++i) s[i] = d[i] * stp; }
    if (!(dr <= EPSMCH * ddum)) {
      int slot;
      if (col < M) { slot = (head + col) % M; ++col; }
      else { slot = head; head = (head + 1) % M; }
      for (int i = 0; i < N; ++i) { S[slot][i] = s[i]; Y[slot][i] = r[i]; }
      theta = dot(r, r) / dr;
    }
    return begin_iteration();
  }
};

}  // namespace mlrn

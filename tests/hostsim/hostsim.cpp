// hostsim.cpp -- TEST-ONLY single-member instantiation of the device feature templates.
//
// Compiled by g++ (tests/hostsim/build.py) so the arithmetic in csrc/*.cuh can be exercised
// against the CPU oracle in a container without a GPU.  It is never linked into
// libmallorn_b200.so and never imported by the mallorn_astrophysics_b200 package: the product
// has no CPU compute path.
#include <vector>
#include <cstring>
#include "stat.cuh"
#include "colors.cuh"
#include "shape.cuh"
#include "tde.cuh"
#include "physics.cuh"
#include "research.cuh"
#include "gp_host.hpp"
#include "fits_thread.cuh"
#include "gp1d.cuh"
#include "lbfgsb_exact.cuh"

using namespace mlrn;

namespace {
struct Scratch {
  std::vector<double> t, f, e, sa, sb, sc, sd, row;
  std::vector<unsigned short> rank;
  explicit Scratch(int cap) : t(cap), f(cap), e(cap), sa(cap), sb(cap), sc(cap), sd(cap), row(512), rank(cap) {}
};
}  // namespace

template <int KIND>
static void hs_decline_one(const double* t, const double* f, int n, double k, double* out, int* status) {
  DeclineJob<KIND> job; int code;
  if (job.init(t, f, n, k, code)) { while ((code = job.step()) == TRF_RUNNING) {} }
  *out = job.finish(code);
  if (status) { status[0] = code; status[1] = (code == FIT_SKIPPED || code == FIT_BAD_INPUT) ? 0 : job.trf.nfev; }
}

typedef double (*hs_objective)(const double* x, double* g);
template <class Opt>
static void hs_drive(hs_objective fn, const double* x0, const double* lo, const double* hi, double* x_out, double* f_out, int* counts) {
  Opt o; o.start(x0, lo, hi);
  // SciPy's ScalarFunction does not re-evaluate (or count) a point equal to the previous one; gp1d_band does the same
  double px[3] = {qnan(), qnan(), qnan()}, f = 0.0, g[3]; int evals = 0;
  for (;;) {
    if (!(px[0] == o.x[0] && px[1] == o.x[1] && px[2] == o.x[2])) { f = fn(o.x, g); ++evals; for (int i = 0; i < 3; ++i) px[i] = o.x[i]; }
    if (!o.advance(f, g)) break;
  }
  for (int i = 0; i < 3; ++i) x_out[i] = o.x[i];
  *f_out = o.f; counts[0] = o.nit; counts[1] = evals; counts[2] = o.why;
}

extern "C" {

int hs_stat(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
            const unsigned short* rank, double* out) {
  HostTeam tm;
  Scratch s(2 * MAX_OBS);
  for (long long o = 0; o < n_obj; ++o) {
    ObjView v;
    stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);
    stat_row(tm, v, s.sa.data(), s.sb.data(), s.row.data());
    std::memcpy(out + o * STAT_COLS, s.row.data(), sizeof(double) * STAT_COLS);
  }
  return 0;
}

#define HS_PASS(NAME, COLS, CALL)                                                                        \
  int NAME(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,   \
           const unsigned short* rank, const double* z, double* out) {                                     \
    HostTeam tm;                                                                                           \
    Scratch s(2 * MAX_OBS);                                                                                \
    for (long long o = 0; o < n_obj; ++o) {                                                                \
      ObjView v;                                                                                           \
      stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);   \
      double zz = z ? z[o] : qnan();                                                                       \
      (void)zz;                                                                                            \
      CALL;                                                                                                \
      std::memcpy(out + o * COLS, s.row.data(), sizeof(double) * COLS);                                    \
    }                                                                                                      \
    return 0;                                                                                              \
  }

HS_PASS(hs_color, COLOR_COLS, color_row(tm, v, s.sa.data(), s.sb.data(), s.row.data()))
HS_PASS(hs_shape, SHAPE_COLS, shape_row(tm, v, s.sa.data(), s.sb.data(), s.row.data()))
HS_PASS(hs_tde, TDE_COLS, tde_row(tm, v, s.sa.data(), s.sb.data(), s.sc.data(), s.sd.data(), s.row.data()))
HS_PASS(hs_physics, PHYSICS_COLS, physics_row(tm, v, zz, s.sa.data(), s.sb.data(), s.row.data()))

// research block (research.cuh): extra scratch for the 1-day grid of the r band and the wavelet
int hs_research(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
                const unsigned short* rank, const double* z, double* out) {
  HostTeam tm;
  Scratch s(2 * MAX_OBS);
  std::vector<double> grid(16384), wav(MHPS_WAVELET_CAP);
  for (long long o = 0; o < n_obj; ++o) {
    ObjView v;
    stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);
    research_row(tm, v, z ? z[o] : qnan(), s.sa.data(), s.sb.data(), s.sc.data(), grid.data(), (int)grid.size(), wav.data(), s.row.data());
    std::memcpy(out + o * RESEARCH_COLS, s.row.data(), sizeof(double) * RESEARCH_COLS);
  }
  return 0;
}

// one-thread-per-fit variants (trf_thread.cuh / fits_thread.cuh)
int hs_bazin_t(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
               const unsigned short* rank, double* out, int* status) {
  HostTeam tm;
  for (long long o = 0; o < n_obj; ++o) {
    double* row = out + o * BAZIN_COLS;
    for (int b = 0; b < N_BAND; ++b) {
      const long long lo = seg_off[o * N_SEG + b];
      const int n = (int)(seg_off[o * N_SEG + b + 1] - lo);
      BazinJob job; int code;
      if (job.init(t + lo, f + lo, e + lo, n, code)) { while ((code = job.step()) == TRF_RUNNING) {} }
      job.finish(code, row + 8 * b);
      if (status) { status[o * 12 + 2 * b] = code; status[o * 12 + 2 * b + 1] = (code == FIT_SKIPPED || code == FIT_BAD_INPUT) ? 0 : job.trf.nfev; }
    }
    bazin_tail(tm, row);
  }
  return 0;
}

int hs_powerlaw_t(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
                  const unsigned short* rank, double* out, int* status) {
  const int bands[3] = {BG, BR, BI};
  for (long long o = 0; o < n_obj; ++o)
    for (int j = 0; j < 3; ++j) {
      const long long lo = seg_off[o * N_SEG + bands[j]];
      const int n = (int)(seg_off[o * N_SEG + bands[j] + 1] - lo);
      double* row = out + o * POWERLAW_COLS + 9 * j;
      int* st = status ? status + o * 54 + 18 * j : nullptr;
      for (int q = 0; q < 7; ++q) hs_decline_one<0>(t + lo, f + lo, n, decline_index(q), row + q, st ? st + 2 * q : nullptr);
      hs_decline_one<1>(t + lo, f + lo, n, 0.0, row + 7, st ? st + 14 : nullptr);
      hs_decline_one<2>(t + lo, f + lo, n, 0.0, row + 8, st ? st + 16 : nullptr);
    }
  return 0;
}

int hs_gp2d(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
            const unsigned short* rank, double* out, int* status) {
  HostTeam tm;
  Scratch s(2 * MAX_OBS);
  for (long long o = 0; o < n_obj; ++o) {
    ObjView v;
    stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);
    gphost::gp2d_row(v, out + o * 27, status ? status + o * 4 : nullptr);
  }
  return 0;
}

// The bounded optimisers driven by a caller-supplied objective (python callback), which isolates them from the
// objective's rounding: which = 0 -> csrc/lbfgsb.cuh (the product: explicit 3 x 3 model), 1 -> lbfgsb_exact.cuh
// (TEST-ONLY routine-by-routine restatement of the published compact-form code).
int hs_lbfgsb3(int which, hs_objective fn, const double* x0, const double* lo, const double* hi, double* x_out, double* f_out, int* counts) {
  if (which == 0) hs_drive<LbfgsB<3>>(fn, x0, lo, hi, x_out, f_out, counts);
  else hs_drive<LbfgsBExact<3>>(fn, x0, lo, hi, x_out, f_out, counts);
  return 0;
}

// per-band 1-D GP (gp1d.cuh / lbfgsb.cuh); status: per band {evaluations, iterations, winning start}
int hs_gp1d(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
            const unsigned short* rank, double* out, int* status) {
  HostTeam tm;
  const int bands[4] = {BG, BR, BI, BZ};
  int longest = 5;                                   // scratch for the longest band that will be fitted
  for (long long o = 0; o < n_obj; ++o)
    for (int j = 0; j < 4; ++j) {
      const int n = (int)(seg_off[o * N_SEG + bands[j] + 1] - seg_off[o * N_SEG + bands[j]]);
      if (n > longest && n <= GP1D_MAX_POINTS) longest = n;
    }
  std::vector<double> scratch((size_t)gp1d_scratch_doubles(longest));
  LbfgsB<3> opt; double ctl[8];
  for (long long o = 0; o < n_obj; ++o) {
    Gp1dBandResult res[4];
    for (int j = 0; j < 4; ++j) {
      const long long lo = seg_off[o * N_SEG + bands[j]];
      const int n = (int)(seg_off[o * N_SEG + bands[j] + 1] - lo);
      Gp1dScratch s = gp1d_carve(scratch.data(), n);
      gp1d_band(tm, t + lo, f + lo, e + lo, n, s, &opt, ctl, res[j]);
      if (status) { status[o * 12 + 3 * j] = res[j].nfev; status[o * 12 + 3 * j + 1] = res[j].nit; status[o * 12 + 3 * j + 2] = res[j].winner; }
    }
    gp1d_tail(res, out + o * GP1D_COLS);
  }
  return 0;
}


// the two division helpers of the fit models (trf.cuh, fits.cuh): out[i] = {div_r(a, d, 1 / d), a / d} and, for the
// curve's own division, {div_curve(a, d), a / d}
void hs_div_helpers(long long n, const double* a, const double* d, double* out_r, double* out_curve, double* out_ieee) {
  for (long long i = 0; i < n; ++i) {
    out_r[i] = div_r(a[i], d[i], 1.0 / d[i]);
    out_curve[i] = div_curve(a[i], d[i]);
    out_ieee[i] = a[i] / d[i];
  }
}

}  // extern "C"

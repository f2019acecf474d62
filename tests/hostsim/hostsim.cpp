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
#include "fits.cuh"
#include "gp_host.hpp"

using namespace mlrn;

namespace {
struct Scratch {
  std::vector<double> t, f, e, sa, sb, sc, sd, row;
  std::vector<unsigned short> rank;
  explicit Scratch(int cap) : t(cap), f(cap), e(cap), sa(cap), sb(cap), sc(cap), sd(cap), row(512), rank(cap) {}
};
}  // namespace

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

int hs_bazin(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
             const unsigned short* rank, double* out, int* status) {
  HostTeam tm;
  Scratch s(2 * MAX_OBS);
  std::vector<double> work(trf_work_doubles<5>(MAX_OBS));
  for (long long o = 0; o < n_obj; ++o) {
    ObjView v;
    stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);
    bazin_row(tm, v, s.sa.data(), work.data(), MAX_OBS, s.row.data(), status ? status + o * 12 : nullptr);
    std::memcpy(out + o * BAZIN_COLS, s.row.data(), sizeof(double) * BAZIN_COLS);
  }
  return 0;
}

int hs_powerlaw(long long n_obj, const long long* seg_off, const double* t, const double* f, const double* e,
                const unsigned short* rank, double* out, int* status) {
  HostTeam tm;
  Scratch s(2 * MAX_OBS);
  std::vector<double> work(trf_work_doubles<3>(MAX_OBS));
  for (long long o = 0; o < n_obj; ++o) {
    ObjView v;
    stage_object(tm, seg_off, t, f, e, rank, o, s.t.data(), s.f.data(), s.e.data(), s.rank.data(), v);
    powerlaw_row(tm, v, s.sa.data(), s.sb.data(), work.data(), MAX_OBS, s.row.data(), status ? status + o * 54 : nullptr);
    std::memcpy(out + o * POWERLAW_COLS, s.row.data(), sizeof(double) * POWERLAW_COLS);
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

}  // extern "C"

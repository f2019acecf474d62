// hostsim.cpp -- TEST-ONLY single-member instantiation of the device feature templates.
//
// Compiled by g++ (tests/hostsim/build.py) so the arithmetic in csrc/*.cuh can be exercised
// against the CPU oracle in a container without a GPU.  It is never linked into
// libmallorn_b200.so and never imported by the mallorn_astrophysics_b200 package: the product
// has no CPU compute path.
#include <vector>
#include <cstring>
#include "stat.cuh"

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

}  // extern "C"

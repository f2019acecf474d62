// objview.cuh -- one object's observations as the feature routines see them.
//
// Packed store layout (HBM, struct-of-arrays, written by pack.cu):
//   seg_off[n_obj*7 + 1]  int64   CSR offsets of (object, band) segments; band 0..5 = ugrizy,
//                                 6 = rows whose Filter is none of those
//   t, flux, err [n_rows] float64 observations ordered by (object, band, time, file rank)
//   rank         [n_rows] uint16  position of the row inside its object in FILE order
// File order is semantically significant in the reference: first-argmax
// peaks in colors.py:39-44 and the all-band rise time in lightcurve_shape.py:287-302 are taken
// in groupby order, so every observation keeps its within-object rank.
#pragma once
#include "team.cuh"

namespace mlrn {

constexpr int N_SEG = 7;        // ugrizy + unknown
constexpr int N_BAND = 6;
constexpr int MAX_OBS = 65535;     // per-object cap: within-object ranks are stored as uint16; pack rejects larger objects
constexpr int SHORT_LIMIT = 512;   // objects up to this many observations are staged in shared memory, longer ones in a global scratch

// Which objects a launch of a staging kernel (K1 arrange, K2, K3, research) handles and where their staging area lives.
// The reference takes any length (statistical.py:163-190); survey light curves have ~150 points, so the shared-memory
// launch is sized by the longest SHORT object of the batch and the rare longer ones run the same code in a second launch
// over the list pack.cu makes of them, with the per-warp staging area in global memory (L2-resident).
struct ObjClass {
  const int* list;           // long launch: the objects to handle; nullptr: every object not above n_hi
  int count;                 // entries of list, or n_obj
  int n_hi;                  // short launch: objects with more observations are left to the long launch
  unsigned char* scratch;    // long launch: staging areas [warps of the grid][bytes per warp]; nullptr: shared memory
};
enum Band { BU = 0, BG = 1, BR = 2, BI = 3, BZ = 4, BY = 5, BX = 6 };

struct ObjView {
  int n;                 // observations of the object
  int off[N_SEG + 1];    // local CSR: band b occupies [off[b], off[b+1])
  const double* t;       // staged copies (shared memory on device)
  const double* f;
  const double* e;
  const unsigned short* rank;
  MLRN_FN_FORCE int cnt(int b) const { return off[b + 1] - off[b]; }
};

// Stage one object from the packed store into team scratch (coalesced, lane-strided).
template <class Team>
MLRN_FN void stage_object(const Team& tm, const long long* seg_off, const double* gt, const double* gf,
                          const double* ge, const unsigned short* grank, long long obj,
                          double* st, double* sf, double* se, unsigned short* srank, ObjView& v) {
  const long long* so = seg_off + obj * N_SEG;
  long long base = so[0];
  for (int b = 0; b <= N_SEG; ++b) v.off[b] = (int)(so[b] - base);
  v.n = v.off[N_SEG];
  for (int i = tm.rank(); i < v.n; i += Team::size) {
    st[i] = gt[base + i];
    sf[i] = gf[base + i];
    se[i] = ge[base + i];
    srank[i] = grank[base + i];
  }
  v.t = st; v.f = sf; v.e = se; v.rank = srank;
  tm.sync();
}

// first argmax of flux over [lo, hi) in FILE order: largest value, ties -> smallest file rank
template <class Team>
MLRN_FN int argmax_file_order(const Team& tm, const ObjView& v, int lo, int hi) {
  // two reductions: the max value, then the smallest rank among the elements attaining it
  double m = -pinf(); int any_nan = 0;
  for (int i = lo + tm.rank(); i < hi; i += Team::size) { m = fmax(m, v.f[i]); any_nan |= is_nan(v.f[i]) ? 1 : 0; }
  m = tm.max(m);
  any_nan = tm.imax(any_nan);                        // numpy.argmax: the first NaN is the maximum
  int best = 0x7fffffff;
  for (int i = lo + tm.rank(); i < hi; i += Team::size)
    if (any_nan ? is_nan(v.f[i]) : (v.f[i] == m)) { int key = ((int)v.rank[i] << 11) | (i - lo); if (key < best) best = key; }
  best = tm.imin(best);
  return (best == 0x7fffffff) ? -1 : lo + (best & 2047);
}

}  // namespace mlrn

/* mallorn_b200.h -- C-ABI of libmallorn_b200.so (B200 / sm_100a light-curve featurisation).
 *
 * The reference (Smooth-Cactus0/MALLORN-astrophysics) is pure Python and has no FFI of its own;
 * its boundary for this path is the set of batch extractors in src/features/.  Each entry point
 * below is what the Python shim in mallorn_astrophysics_b200/features/ binds (ctypes) to serve
 * the reference function named beside it.  Plain pointers and sizes only; every function
 * returns 0 on success or a negative MLRN_E* code, with text from mlrn_last_error().  No C++
 * exception crosses the boundary.  There is no CPU implementation behind any of these symbols.
 *
 * Pointers flagged by a *_on_device argument may be host or device memory; a `stream` argument
 * is a cudaStream_t passed as void* (NULL = the legacy default stream).  With device outputs the
 * calls are asynchronous on `stream`; with host outputs they return after the copy completes.
 *
 * Concurrency: a store belongs to one host thread and one stream at a time -- the passes keep their work counters in the
 * store, so two passes on the SAME store must be ordered (same stream, or an event between them; mlrn_fits orders its
 * own internal side stream).  Different stores are independent and may be used from different threads and streams.
 * mlrn_last_error() is per thread.
 *
 * Feature rows are float64, row-major [n_obj, n_cols], columns in the reference's order
 * (mallorn_astrophysics_b200/schema.py).  One row per dense object index 0..n_obj-1; an object
 * without observations yields the reference's "empty" values and is dropped by the shim
 * (statistical.py:163-165).
 */
#ifndef MALLORN_B200_H
#define MALLORN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLRN_OK 0
#define MLRN_EINVAL (-1)      /* bad argument */
#define MLRN_ECUDA (-2)       /* CUDA runtime error (see mlrn_last_error) */
#define MLRN_ETOOLONG (-3)    /* an object has more than MLRN_MAX_OBS observations */
#define MLRN_ENODEVICE (-4)   /* no CUDA device: there is no CPU fallback */

#define MLRN_MAX_OBS 65535    /* per-object observation cap: within-object file ranks are uint16.  Objects up to 512
                               * observations are staged in shared memory, longer ones in a global scratch (same code) */
#define MLRN_N_SEG 7          /* segments per object: u g r i z y + unknown filter */

#define MLRN_STAT_COLS 123      /* extract_statistical_features   statistical.py:135      */
#define MLRN_COLOR_COLS 83      /* extract_color_features         colors.py:347           */
#define MLRN_SHAPE_COLS 65      /* extract_shape_features         lightcurve_shape.py:335 */
#define MLRN_TDE_COLS 25        /* extract_tde_physics_features   tde_physics.py:377      */
#define MLRN_PHYSICS_COLS 32    /* extract_physics_features       physics_based.py:461    */
#define MLRN_BAZIN_COLS 52      /* extract_bazin_features         bazin_fitting.py:254    */
#define MLRN_POWERLAW_COLS 27   /* extract_powerlaw_features      scripts/train_v55_powerlaw.py:196 */
#define MLRN_GP2D_COLS 27       /* extract_multiband_gp_features  multiband_gp.py:347     */
#define MLRN_GP1D_COLS 21       /* extract_gp_features            gaussian_process.py:251 */
#define MLRN_RESEARCH_COLS 40   /* extract_research_features      research_features.py:566 */
#define MLRN_SEGMETA_COLS 22    /* K1 integer outputs per object, see mlrn_store_segments */

typedef struct mlrn_store mlrn_store;   /* packed, device-resident light curves of one batch */

const char* mlrn_last_error(void);
int mlrn_version(void);
int mlrn_device_count(void);            /* number of CUDA devices (0 => every call fails) */
long long mlrn_launch_count(void);      /* kernels this library has launched so far in the process */
/* FP64 roofline denominator measured on `device` at call time ("measure with a DFMA micro-benchmark at
 * bench time"): TFLOP/s (2 x FMA) of a register-resident DFMA loop and of the mma.sync.m8n8k4.f64 loop the GP sweep
 * issues, 16 warps on every SM.  No reference equivalent (measurement only). */
int mlrn_measure_fp64(int device, double* dfma_tflops, double* dmma_tflops);

/* K1 -- pack.  Replaces the per-extractor `lightcurves.groupby('object_id')` dict
 * (statistical.py:155, colors.py:365, lightcurve_shape.py:353, physics_based.py:484,
 * tde_physics.py:395, bazin_fitting.py:272, multiband_gp.py:369) and the per-band
 * `obj_lc[obj_lc['Filter']==band].sort_values('Time (MJD)')` (lightcurve_shape.py:192 ...).
 * Input: a flat long-format table in FILE order -- obj[i] = dense object index of row i
 * (negative = skip the row), band[i] = 0..5 for u g r i z y, >= 6 for any other Filter.
 * Builds CSR (object, band) segments ordered by (time, file rank) in HBM. */
int mlrn_store_create(int device, int64_t n_rows, int32_t n_obj, const int32_t* obj, const uint8_t* band,
                      const double* t, const double* flux, const double* err, int inputs_on_device,
                      void* stream, mlrn_store** out);
int mlrn_store_destroy(mlrn_store* s);
/* rows kept, objects, and the largest per-object observation count */
int mlrn_store_info(const mlrn_store* s, int64_t* n_rows, int32_t* n_obj, int32_t* max_obs);
/* K1 integer outputs, int32 [n_obj, 22] to HOST memory: for each of the 7 segments
 * {count, position of the first flux maximum in file order, same in time order} (-1 if empty),
 * then peak_band (statistical.py:217-222).  Bit-exact contract. */
int mlrn_store_segments(const mlrn_store* s, int32_t* out_host);
/* packed columns to HOST memory (any pointer may be NULL): seg_off int64[n_obj*7+1],
 * t/flux/err float64[n_rows], rank / trank uint16[n_rows] (position inside the object in file order /
 * in (time, file rank) order over all bands) */
int mlrn_store_export(const mlrn_store* s, int64_t* seg_off, double* t, double* flux, double* err, uint16_t* rank,
                      uint16_t* trank);

/* Feature passes.  out: float64 [n_obj, MLRN_*_COLS]. */
int mlrn_stat(const mlrn_store* s, double* out, int out_on_device, void* stream);
int mlrn_color(const mlrn_store* s, double* out, int out_on_device, void* stream);
int mlrn_shape(const mlrn_store* s, double* out, int out_on_device, void* stream);
int mlrn_tde(const mlrn_store* s, double* out, int out_on_device, void* stream);
/* z: redshift per object (metadata['Z'], physics_based.py:481,496); NaN = unknown */
int mlrn_physics(const mlrn_store* s, const double* z, int z_on_device, double* out, int out_on_device, void* stream);
/* status (optional, may be NULL): int32 [n_obj, 6, 2] = {TRF termination status, nfev} per band;
 * status < -99 means curve_fit would have raised (NaN row). */
int mlrn_bazin(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream);
/* status (optional): int32 [n_obj, 27, 2] */
int mlrn_powerlaw(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream);
/* Both fit passes of the caching drivers (cache_bazin_features.py:44-60 and the power-law loop of
 * scripts/train_v55_powerlaw.py:147-202, run one after the other upstream) in ONE call: the Bazin pass is ordered on
 * `stream`, the decline models on an internal side stream forked from and joined to it, so that the straggler stages of
 * one pass (a few long fits, most SMs idle) run under the other's bulk.  Same results as mlrn_bazin + mlrn_powerlaw.
 * Host buffers (out_on_device = 0) are served by the two calls in sequence. */
int mlrn_fits(const mlrn_store* s, double* bazin_out, int32_t* bazin_status, double* powerlaw_out,
              int32_t* powerlaw_status, int out_on_device, void* stream);
/* status (optional): int32 [n_obj, 4] = {L-BFGS-B iterations, function evaluations, termination code, N used} */
int mlrn_gp2d(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream);
/* per-band 1-D GP (sklearn GaussianProcessRegressor in the reference: ConstantKernel * RBF + WhiteKernel, four
 * bounded L-BFGS-B starts per band, gaussian_process.py:29-146) for g, r, i, z plus the cross-band columns.
 * status (optional): int32 [n_obj, 4, 3] = per band {objective evaluations over the four starts, L-BFGS-B
 * iterations over the four starts, index of the winning start (-1: band not fitted)} */
int mlrn_gp1d(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream);

/* K8 -- "research" block of the v115c model (power-law decay quality, nuclear-position proxies, colour at
 * peak, Mexican-hat power spectra of the r band on a 1-day grid, luminosity; research_features.py:44-560), 40 columns,
 * held to the reference's own frames on the B200 (tests/test_gpu_research.py).  z: redshift per object or NULL (luminosity columns NaN, as for Z <= 0 upstream :556);
 * r-band coverage beyond 8192 days leaves the six mhps_* columns NaN. */
int mlrn_research(const mlrn_store* s, const double* z, int z_on_device, double* out, int out_on_device, void* stream);

/* K9 -- feature interactions on a device-resident table.  Replaces create_physics_interactions
 * (src/features/interactions.py:23-196): each new column is one of the seven element-wise forms below of one to three
 * existing columns.  `table` [n_rows, ld] and `out` [n_rows, ld_out] are DEVICE pointers (row-major float64), `ops` is a
 * HOST array of n_ops descriptors; column k of `out` receives op k.  Which interactions exist (the reference only adds
 * those whose inputs are present) is decided by the caller from the column names (features/interactions.py). */
enum { MLRN_IOP_MUL = 0,          /* a * b                       e.g. g_r_post_20d_x_Z           interactions.py:52 */
       MLRN_IOP_DIV_OFFSET = 1,   /* a / (b + k)                 e.g. g_r_post_20d_div_Z (k=0.1) interactions.py:53 */
       MLRN_IOP_DIFF_SCALED = 2,  /* (a - b) / k                 temp_drop_rate (k=50)           interactions.py:73 */
       MLRN_IOP_CLIP_POW = 3,     /* clip(a, 0, k) ** 0.25       temp_peak_4th (k=1e5)           interactions.py:76 */
       MLRN_IOP_SUB = 4,          /* a - b                       gr_color_accel                  interactions.py:138 */
       MLRN_IOP_SQUARE = 5,       /* a ** 2                      r_skew_squared                  interactions.py:170 */
       MLRN_IOP_ASYM_COLOR = 6 }; /* (a / (b + 1)) * c           g_asym_x_color                  interactions.py:116-119 */
typedef struct mlrn_inter_op { int32_t kind, a, b, c; double k; } mlrn_inter_op;   /* b, c = -1 when unused */
int mlrn_interactions(int device, const double* table, int64_t n_rows, int32_t ld, const mlrn_inter_op* ops, int32_t n_ops,
                      double* out, int32_t ld_out, void* stream);

/* K10 -- histogram gradient-boosted trees on a device-resident table.  Replaces the xgboost
 * `tree_method='hist'` training of the downstream drivers (non_successful_tests/scripts/train_v92_focal_adversarial.py:
 * 157-169 parameters, :232-272 5-fold loop; adversarial_validation.py:120-160 the train-vs-test classifier) for the
 * binary:logistic objective; the focal-loss variants of v92a-c are not covered.  All pointers are DEVICE pointers unless
 * named *_host; the caller keeps `bins`, `y` and `weight` alive while the handle exists.  The CPU restatement of the same
 * algorithm (oracle/gbdt.py) grows bit-identical trees. */
typedef struct mlrn_gbdt mlrn_gbdt;
typedef struct mlrn_gbdt_params {
  int32_t max_depth;            /* 1..5                       train_v92...:159 max_depth 5       */
  double learning_rate;         /*                            :160 0.025                         */
  double subsample;             /* rows per tree              :161 0.8                           */
  double min_child_weight;      /*                            :163 3                             */
  double reg_alpha, reg_lambda; /*                            :164-165 0.2, 1.5                  */
  double scale_pos_weight;      /* weight of label-1 rows     :240 n_neg / n_pos                 */
  uint64_t seed;                /* row sub-sampling           :167 random_state 42               */
} mlrn_gbdt_params;
/* one tree as a complete binary tree of depth <= 5: node k has children 2k+1 / 2k+2; feature < 0 marks a leaf */
typedef struct mlrn_gbdt_tree { int32_t feature[63], bin[63], default_left[63]; double value[63]; } mlrn_gbdt_tree;
/* X [n, ld] float64 -> bins [n, F] uint8: bin b <=> cuts[f][b-1] < x <= cuts[f][b] (at most 254 cuts per feature, row f of
 * `cuts` [F, 255]), NaN -> 255 (missing). */
int mlrn_gbdt_bin(int device, const double* X, int64_t n, int32_t F, int32_t ld, const double* cuts, const int32_t* ncuts, uint8_t* bins,
                  void* stream);
int mlrn_gbdt_create(int device, const uint8_t* bins, int64_t n, int32_t F, const double* y, const double* weight /* or NULL */,
                     const mlrn_gbdt_params* params, int32_t max_rounds, void* stream, mlrn_gbdt** out);
int mlrn_gbdt_destroy(mlrn_gbdt* g);
/* grow one tree (colmask_host [F]: 1 = column sampled for this tree, colsample_bytree) and add it to the training margins */
int mlrn_gbdt_round(mlrn_gbdt* g, const uint8_t* colmask_host, void* stream);
/* margin[i] += sum of trees [tree_from, tree_to) for other rows (validation / test), binned with the same cuts */
int mlrn_gbdt_predict(const mlrn_gbdt* g, const uint8_t* bins, int64_t n, int32_t tree_from, int32_t tree_to, double* margin, void* stream);
int mlrn_gbdt_export(const mlrn_gbdt* g, mlrn_gbdt_tree* trees_host, double* train_margin_host);   /* synchronises */
int32_t mlrn_gbdt_rounds(const mlrn_gbdt* g);

#ifdef __cplusplus
}
#endif
#endif /* MALLORN_B200_H */

/*
 * sdb.h — C ABI of the B200-native pointwise statistical-downscaling engine.
 *
 * Drop-in boundary for the hot path of pangeo-data/scikit-downscale
 * (`PointWiseDownscaler.fit/predict` over BcsdTemperature, BcsdPrecipitation,
 * QuantileMapper, PureAnalog, AnalogRegression).  The reference has no native
 * code: what these entry points replace is its per-cell Python loop
 *   skdownscale/pointwise_models/core.py:69-97   (_fit_wrapper, one estimator.fit per cell)
 *   skdownscale/pointwise_models/core.py:100-143 (_predict_wrapper, one estimator.predict per cell)
 * together with the estimator bodies cited per function below.  A maintainer
 * binds them from Python with ctypes.
 *
 * Conventions
 *  - Every data pointer is a DEVICE pointer (cudaMalloc / torch CUDA tensor); the
 *    caller owns all buffers; nothing is allocated inside.  `stream` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *    enqueue work; they do not synchronise.
 *  - Arrays are "time-major, cell-fastest": element (t, c) of a [T, C] array is
 *    at base[t * ld + c] (ld >= C, in elements) — the C-order (time, lat, lon)
 *    layout of the reference's xarray inputs, zero-copy.  Multi-feature inputs
 *    are [T, p, C]: element (t, f, c) at base[(t * p + f) * ld + c].
 *  - dtype: SDB_F32 or SDB_F64 (inputs and fitted state share one dtype).
 *  - Group tables (device, int32): rows[g * max_len + j] = row number (time
 *    index) of the j-th member of group g in time order, -1 beyond len[g].
 *  - cell_valid (device, uint8, may be NULL = all valid): 0 marks a cell the
 *    reference would skip (core.py:35-37); its outputs are NaN.
 *  - nonfinite (device, int32[1], may be NULL): the kernels OR 1 into it when they meet a
 *    NaN/inf inside a valid cell — the condition on which the reference's sklearn
 *    validation raises ValueError (base.py:18-20,29-31).  The caller zeroes it, reads
 *    it back after synchronising and raises.
 *  - Return value: 0 on success, negative on error (SDB_E_*); the message is
 *    available from sdb_last_error() (thread-local).
 *  - Thread safety: entry points keep no global mutable state apart from the
 *    thread-local error string; distinct streams may be driven concurrently.
 */
#ifndef SDB_H_
#define SDB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_F32 0
#define SDB_F64 1

#define SDB_E_INVALID (-1)   /* bad argument (shape, dtype, NULL pointer, unsupported size) */
#define SDB_E_CUDA    (-2)   /* CUDA launch / runtime error */
#define SDB_E_UNSUPPORTED (-3)

/* quantile-mapping predict modes */
#define SDB_MODE_QM     0    /* QuantileMapper.transform        quantile.py:109-147 */
#define SDB_MODE_BCSD_P 1    /* BcsdPrecipitation.predict       bcsd.py:149-185     */
#define SDB_MODE_BCSD_T 2    /* BcsdTemperature.predict         bcsd.py:230-281     */

/* group-mean arithmetic (what pandas does in the reference, see oracle/bcsd.py) */
#define SDB_MEAN_GROUPBY 0   /* df.groupby().mean(): Kahan sum in input dtype   bcsd.py:138,222-223 */
#define SDB_MEAN_FRAME   1   /* DataFrame.mean(): numpy pairwise sum            groupers.py:84-89   */
#define SDB_MEAN_NUMPY   2   /* one-column frame .mean(): pairwise over all n    quantile.py:673-674 */

/* PureAnalog kinds (gard.py:310-336) */
#define SDB_ANALOG_BEST   0
#define SDB_ANALOG_SAMPLE 1
#define SDB_ANALOG_WEIGHT 2
#define SDB_ANALOG_MEAN   3
#define SDB_ANALOG_REGRESSION 4   /* AnalogRegression._predict_one_step  gard.py:191-224 */

/* CunnaneTransformer settings of the quantile map (quantile.py:420-432, reached through
 * QuantileMapper(qt_kwargs=...) / BcsdBase(qm_kwargs={'qt_kwargs': ...})).  `extrapolate`: which tails of the
 * fitted CDF continue as the OLS line through the n_endpoints end points (quantile.py:526-543); the
 * other tails clamp to the end value like np.interp.  None and '1to1' both mean SDB_EXTRAPOLATE_NONE
 * (CunnaneTransformer treats them alike).  A NULL pointer = the defaults {0.4, 0.4, 10, BOTH}. */
#define SDB_EXTRAPOLATE_NONE 0
#define SDB_EXTRAPOLATE_MIN  1
#define SDB_EXTRAPOLATE_MAX  2
#define SDB_EXTRAPOLATE_BOTH 3
typedef struct sdb_cunnane_opts {
    double alpha;         /* plotting positions (i - alpha) / (n + 1 - alpha - beta)   quantile.py:23-43.
                           * NOTE: the reference's CunnaneTransformer never hands its alpha / beta to
                           * plotting_positions (quantile.py:462), i.e. it always maps with 0.4 / 0.4; a
                           * drop-in caller passes 0.4 here whatever the user asked for. */
    double beta;
    int32_t n_endpoints;  /* >= 1 */
    int32_t extrapolate;  /* SDB_EXTRAPOLATE_* */
} sdb_cunnane_opts;

#define SDB_MAX_GROUP_LEN 16384   /* longest group (padded to a power of two) one CTA sorts */
#define SDB_MAX_ANALOGS   256

int sdb_version(void);
const char* sdb_last_error(void);

/* Testing aid: bit 0 forces the generic (any dtype / any group length) kernels even where the
 * float32 tile kernels apply; bit 1 makes the tile kernels use their 4-byte row accesses (the
 * path rows that are not 16-byte aligned take) everywhere; bits 2 / 3 make sdb_bcsd_fit_predict
 * sort / rank with the sorting network instead of the counting rank.  Returns the previous flags.
 * This is the ONE piece of process-wide mutable state in the library (the exception to the
 * thread-safety paragraph above): set it before issuing work, not concurrently with it. */
int sdb_set_debug_flags(int flags);

/* Strided copy of a [height, width_bytes] block (cudaMemcpy2DAsync) on `stream`: how a column block (cell
 * range) of a time-major array travels host <-> device (kind 0 = host to device, 1 = device to host) and, with
 * kind 2 (cudaMemcpyDefault), device <-> device INCLUDING a peer GPU's memory mapped into this process (CUDA
 * IPC): the gather of the predicted field — every rank's out[T, C_rank] block is pushed by the copy engines
 * into columns [c0, c0 + C_rank) of each peer's full [T, n_cells] field over NVLink while the next cell
 * chunk is still being computed (north_star's "final gather of the predicted field"; the reference has no
 * equivalent, its dask workers return blocks to the client, core.py:262,336).  Pitches in bytes. */
int sdb_memcpy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                       int64_t width_bytes, int64_t height, int kind, void* stream);

/* The same copy for device -> device (peer) blocks whose rows are 16-byte aligned multiples of 16 bytes, done by
 * n_ctas CTAs (<= 0: 32) of plain 16-byte loads / stores instead of the copy engines: the NVLink fast path for
 * the strided destination of the gather (a column block of the peer's full field).  SDB_E_UNSUPPORTED when the
 * geometry is not 16-byte aligned (fall back to sdb_memcpy2d_async kind 2). */
int sdb_peer_copy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                    int64_t width_bytes, int64_t height, int n_ctas, void* stream);

/* One read, n_dst (1..8) peer writes: the block is stored into the same [height, width_bytes] window (pitch
 * dst_pitch) behind every pointer of the HOST array dsts — how a rank pushes a finished cell chunk into all its
 * peers' replicas at once.  Same alignment rule as sdb_peer_copy2d; n_ctas <= 0: 296. */
int sdb_peer_bcast2d(void* const* dsts, int n_dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                     int64_t width_bytes, int64_t height, int n_ctas, void* stream);

/* Enable direct (NVLink) access from `device` to memory of `peer_device` for kernels and copies issued on
 * `device` — required before sdb_peer_copy2d / sdb_memcpy2d_async(kind 2) touch IPC-mapped peer memory (without
 * it the kernel faults and the copy is staged through the host).  Idempotent. */
int sdb_enable_peer_access(int device, int peer_device);

/* Peer-visible device buffers for the gather (CUDA IPC).  sdb_peer_alloc: cudaMalloc on the current device +
 * its 64-byte IPC handle (to be sent to the other ranks by any means); sdb_peer_open: map a peer rank's buffer
 * into this process for access FROM THE CURRENT DEVICE (lazy peer access over NVLink); close / free undo them.
 * These are the only entry points that allocate: a framework's caching allocator cannot export IPC handles
 * for interior pointers. */
int sdb_peer_alloc(int64_t bytes, void** ptr, void* handle64);
int sdb_peer_open(const void* handle64, void** ptr);
int sdb_peer_close(void* ptr);
int sdb_peer_free(void* ptr);

/* Largest group length supported by sdb_qm_fit / sdb_qm_predict. */
int sdb_max_group_len(void);

/*
 * Per-group climatology of every cell: climo[g * ld_out + c] = mean of v over the rows
 * of group g, in the arithmetic the reference's pandas call uses (`how`).
 * Replaces  y_groups.mean() / X.groupby().mean()      bcsd.py:138, 222-223
 *           PaddedDOYGrouper.mean()                    groupers.py:84-89
 */
int sdb_group_mean(const void* v, int dtype, int64_t ld, int64_t n_cells,
                   const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                   int how, void* climo, int64_t ld_out,
                   const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/*
 * Fit the empirical CDFs: for every cell and group, sort the group's values.
 *   sorted_state[c * state_ld + state_off[g] + j] = j-th smallest value of group g in cell c
 * Replaces  BcsdBase._qm_fit_by_group → QuantileMapper.fit → CunnaneTransformer.fit
 *           bcsd.py:59-67, quantile.py:81-107, 438-463 (np.sort at :462).
 * state_off: device int64[n_groups].
 */
int sdb_qm_fit(const void* y, int dtype, int64_t ld, int64_t n_cells,
               const int32_t* rows, const int32_t* len, const int64_t* state_off,
               int n_groups, int max_len,
               void* sorted_state, int64_t state_ld,
               const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/*
 * Quantile-map every cell and group of X through the fitted CDFs.
 *   mode QM      out = QM_g(X)
 *   mode BCSD_P  out = QM_g(X) [ / y_climo ]                         (return_anoms)
 *   mode BCSD_T  roll = centred 9-sample mean inside the climate-trend group (float64),
 *                shift = roll - x_climo, out = shift + QM_g(X - shift) [ - y_climo ]
 * QM_g ranks the group's values among themselves (ties take the highest rank),
 * converts ranks to Cunnane plotting positions and interpolates the fitted
 * sorted values, with the reference's 10-endpoint OLS tails when T_pred > T_fit.
 * Replaces  quantile.py:109-147, 465-545; bcsd.py:69-79, 149-185, 230-281.
 *
 * Predict group g uses fitted group state_gid[g] (device int32[n_groups]); fit_len /
 * state_off (device, indexed by FITTED group) give its length and offset inside a cell's
 * state record; x_climo / y_climo are [n_fit_groups, ld_climo] (may be NULL when unused).
 * roll_nbr: NULL ⇒ the 9-sample window runs inside the predict group itself (monthly
 * mode, bcsd.py:48-49); otherwise device int32[T_pred * 9] with the row numbers of the
 * window members of every row (-1 = absent) — the general case where the climate-trend
 * grouping differs from the mapping groups ('daily_nasa-nex', bcsd.py:53,250,275).
 * cunnane: HOST pointer to the CunnaneTransformer settings, NULL = defaults.
 * rank_out: optional device int32 [T_pred, ld_out] receiving the 1-based in-group rank
 * (parity instrumentation), or NULL.  out_dtype may differ from dtype (the reference's
 * estimators return float64, its wrapper casts to X.dtype — core.py:129).
 */
int sdb_qm_predict(int mode, const void* X, int dtype, int64_t ld, int64_t n_cells,
                   const int32_t* rows, const int32_t* len, const int32_t* state_gid,
                   int n_groups, int max_len,
                   const int32_t* fit_len, const int64_t* state_off, int max_fit_len,
                   const void* sorted_state, int64_t state_ld,
                   const void* x_climo, const void* y_climo, int64_t ld_climo,
                   int return_anoms, const int32_t* roll_nbr, const sdb_cunnane_opts* cunnane,
                   void* out, int out_dtype, int64_t ld_out, int32_t* rank_out,
                   const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/*
 * fit + predict in ONE pass for the case fit and predict share their time index (same groups, same
 * lengths — the headline workload): climatologies (sdb_group_mean), then one kernel that sorts the
 * training group of every cell in shared memory, ranks the prediction group against itself and maps
 * rank r to the r-th order statistic (n == m: np.interp lands on the knot, quantile.py:523-530) —
 * the fitted state never travels through HBM.  Results are bit-identical to sdb_qm_fit + sdb_qm_predict.
 * Replaces  BcsdTemperature.fit + .predict   bcsd.py:197-269
 *           BcsdPrecipitation.fit + .predict bcsd.py:115-185
 *           QuantileMapper.fit + .transform  quantile.py:81-147   (mode QM: y_train = the fitted series)
 *   X_train   [T, ld_train]  only read for mode BCSD_T (x_climo); NULL otherwise
 *   y_train   [T, ld_train], X_pred [T, ld_pred], out [T, ld_out]; float32 only, groups of <= 1024 steps
 *             (SDB_E_UNSUPPORTED otherwise: use the split calls)
 *   x_climo / y_climo  OUTPUTS [n_groups, ld_climo] (y_climo: modes BCSD_*; x_climo: BCSD_T), pandas
 *             groupby().mean() arithmetic; NULL when the mode does not use them
 *   sorted_state / state_ld / state_off: optional OUTPUT, the fitted state exactly as sdb_qm_fit writes
 *             it (so the call leaves a fitted model behind); NULL = not kept
 *   stats     optional device uint64[8] instrumentation, accumulated: [0] series, [1] / [2] series whose
 *             training / prediction side took the sorting-network path, [3] / [4] elements that needed
 *             an exact comparison on the training / prediction side
 */
int sdb_bcsd_fit_predict(int mode, const void* X_train, const void* y_train, const void* X_pred, int dtype,
                         int64_t ld_train, int64_t ld_pred, int64_t n_cells,
                         const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                         void* x_climo, void* y_climo, int64_t ld_climo, int return_anoms,
                         void* sorted_state, int64_t state_ld, const int64_t* state_off,
                         void* out, int64_t ld_out,
                         const uint8_t* cell_valid, int32_t* nonfinite, uint64_t* stats, void* stream);

/*
 * Analog downscaling for every cell: exact k-nearest-neighbour search of every query
 * timestep in the cell's training window (float64 squared Euclidean distance,
 * features accumulated in order, lowest index first on ties) followed by the
 * PureAnalog statistic `kind` or the AnalogRegression OLS epilogue.
 *   X_train [T_fit, p, C], y_train [T_fit, C], X_query [T_q, p, C]  (dtype)
 *   out     [T_q, 3, C] (out_dtype): pred, exceedance_prob, prediction_error
 *   knn_idx optional device int32 [T_q, k, C] (parity instrumentation) or NULL
 *   rand_idx device int32 [T_q, C], only for SDB_ANALOG_SAMPLE (host-drawn, gard.py:315)
 *   has_thresh/thresh: PureAnalog threshold masking (gard.py:303-308, 338-343); for
 *           SDB_ANALOG_REGRESSION the exceedance split of gard.py:201-215: exceedance_prob =
 *           P(class 0) of an L2-regularised logistic regression (inverse strength logistic_c,
 *           sklearn's C, default 1.0; solved to machine precision where the reference's lbfgs
 *           stops at tol 1e-4) and the least-squares fit restricted to the analogs above thresh.
 *           A query whose analogs are ALL at or below thresh makes the reference raise
 *           ("needs samples of at least 2 classes"): the kernel ORs 2 into *nonfinite.
 * Replaces  AnalogBase.fit + PureAnalog.predict / AnalogRegression.predict
 *           gard.py:58-87, 152-224, 273-364.
 */
int sdb_analog_predict(int kind, const void* X_train, const void* y_train, const void* X_query,
                       int dtype, int64_t ld, int64_t n_cells,
                       int t_fit, int t_query, int n_features, int k,
                       int has_thresh, double thresh, double logistic_c, const int32_t* rand_idx,
                       void* out, int out_dtype, int64_t ld_out, int32_t* knn_idx,
                       const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/*
 * The spatial index of the analog search, for all cells at once — the counterpart of AnalogBase.fit's KDTree
 * (gard.py:82).  Every series' training rows are cut into up to sdb_analog_grid_boxes() = 512 boxes by quantile
 * planes of the first predictors (1 predictor: 512 slabs, 2: 22 x 22, 3: 8 x 8 x 8).
 *   sdb_analog_grid_fit     X_train [t_fit, p, ld] → bounds [sdb_analog_grid_planes() = 511, ld_grid] (the planes),
 *                           perm_train [t_fit, ld_grid] (rows grouped by box), box_start [513, ld_grid];
 *                           workspace: int32 [t_fit, ld_grid] scratch.  float32, 64 <= t_fit <= 32768.
 *   sdb_analog_grid_assign  X_query → perm_query [t_query, ld_grid]: the query steps grouped by the SAME boxes, which
 *                           hands neighbouring queries to the same warp.
 * sdb_analog_predict_pruned is sdb_analog_predict with these tables: ONE CTA per cell stages the box-ordered
 * training window once into shared memory (float32) and every query walks the boxes shell by shell around its own,
 * skipping boxes whose distance lower bound exceeds its k-th best distance and stopping when the next shell is out
 * of reach.  Identical results (neighbours, order, outputs; exact distance ties → lowest training row).  Covers
 * what sdb_analog_pruned_supported() reports: float32, 1..3 predictors, k <= 16, t_fit <= 18 958 (3 predictors).
 */
int sdb_analog_grid_fit(const void* X_train, int dtype, int64_t ld, int64_t n_cells, int t_fit, int n_features,
                        int32_t* workspace, float* bounds, int32_t* perm_train, int32_t* box_start, int64_t ld_grid,
                        const uint8_t* cell_valid, void* stream);
int sdb_analog_grid_assign(const void* X_query, int dtype, int64_t ld, int64_t n_cells, int t_query, int n_features,
                           const float* bounds, int32_t* perm_query, int64_t ld_grid,
                           const uint8_t* cell_valid, void* stream);
int sdb_analog_grid_boxes(void);
int sdb_analog_grid_planes(void);
int sdb_analog_pruned_supported(int dtype, int t_fit, int n_features, int k);
int sdb_analog_predict_pruned(int kind, const void* X_train, const void* y_train, const void* X_query,
                              int dtype, int64_t ld, int64_t n_cells,
                              int t_fit, int t_query, int n_features, int k,
                              int has_thresh, double thresh, double logistic_c, const int32_t* rand_idx,
                              void* out, int out_dtype, int64_t ld_out, int32_t* knn_idx,
                              const uint8_t* cell_valid, int32_t* nonfinite,
                              const int32_t* perm_train, const int32_t* perm_query, const int32_t* box_start,
                              const float* bounds, int64_t ld_grid, void* stream);

/*
 * order[r * ld_order + c] = index t of the r-th smallest x[t * row_stride + c], t = 0 .. n_steps - 1, for every cell
 * (equal values in index order; float32; n_steps <= sdb_series_argsort_max_steps() = 32768).  One CTA per cell,
 * block-wide counting rank (csrc/qm_long.cu).
 */
int sdb_series_argsort(const void* x, int dtype, int64_t row_stride, int64_t n_cells, int n_steps,
                       int32_t* order, int64_t ld_order, const uint8_t* cell_valid, void* stream);
int sdb_series_argsort_max_steps(void);

/*
 * 1-based rank of every value among its (cell, group) series: the tie-max rank of the quantile map
 * (ordinal = 0; ties share the highest rank, quantile.py:138,488) or the position in (value, time
 * index) order (ordinal = 1; what `np.argsort` yields on tie-free data, quantile.py:239,607).
 *   rank_out[t * ld_rank + c], int32; cells masked by cell_valid get 0.
 */
int sdb_series_rank(const void* X, int dtype, int64_t ld, int64_t n_cells,
                    const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                    int ordinal, int32_t* rank_out, int64_t ld_rank,
                    const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/* CDF-to-CDF regressors */
#define SDB_QMR_REGRESSOR         0   /* QuantileMappingReressor.predict   quantile.py:224-266 */
#define SDB_QMR_EDCDF_DIFFERENCE  1   /* EquidistantCdfMatcher.predict, kind='difference'   quantile.py:594-636 */
#define SDB_QMR_EDCDF_RATIO       2   /* ... kind='ratio' (max_ratio=None) */

/*
 * Synthetic frame points of the two fitted CDFs of every cell (`_calc_extrapolated_cdf`,
 * quantile.py:311-388): frame[c * 4 + {0, 1, 2, 3}] = lower / upper frame value of sorted X, then of
 * sorted y — the OLS line through the n_endpoints end points evaluated at pp = -1e20 / +1e20 on the
 * tails `extrapolate` names, the repeated end value otherwise.  sorted_x / sorted_y: the fitted
 * states of sdb_qm_fit on ONE whole-series group (cell record = the n_fit sorted values).
 */
int sdb_qmr_frame(const void* sorted_x, const void* sorted_y, int dtype, int64_t state_ld,
                  int64_t n_cells, int n_fit, int extrapolate, int n_endpoints,
                  double* frame, const uint8_t* cell_valid, void* stream);

/*
 * QuantileMappingReressor / EquidistantCdfMatcher predict for every cell and time step:
 *   REGRESSOR         out = interp(interp(x, X_cdf.vals, X_cdf.pp), y_cdf.pp, y_cdf.vals)
 *   EDCDF_DIFFERENCE  q = Cunnane position of x among the T_pred values of its cell (rank, ordinal);
 *                     out = y_cdf(q) + (x - X_cdf(q));     EDCDF_RATIO  out = y_cdf(q) * (x / X_cdf(q))
 * one_to_one: extrapolate='1to1' — values outside the fitted X range keep their distance to the
 * range end (quantile.py:268-309).  The result has the dtype of X (out_dtype == dtype).
 * Under 'min' / 'max' / 'both' the reference's own results OUTSIDE the fitted range are ill-conditioned
 * (a difference of two ~1e21 numbers, quantile.py:17-18,375-386): they are computed the same way but
 * are not comparable bit for bit.  Replaces quantile.py:224-266, 594-636.
 */
int sdb_qmr_predict(int kind, const void* X, int dtype, int64_t ld, int64_t n_cells, int t_pred,
                    const void* sorted_x, const void* sorted_y, int64_t state_ld, int n_fit,
                    const double* frame, int extrapolate, int one_to_one,
                    const int32_t* rank, int64_t ld_rank,
                    void* out, int out_dtype, int64_t ld_out,
                    const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/*
 * Detrending quantile map (QuantileMapper(detrend=True), quantile.py:94-98,127-145; per time group
 * through BcsdBase(qm_kwargs={'detrend': True})) — SURVEY.md §8(f) row 2.  The mapper itself is
 * sdb_qm_fit / sdb_qm_predict on float64 residuals; these entry points are the steps around it.
 */
#define SDB_TREND_REMOVE  0   /* out = v - (j * slope + intercept)                          trend.py:54-64   */
#define SDB_TREND_RESTORE 1   /* out = (v + (j * slope + intercept)) - (intercept - intercept_ref)   trend.py:66-77, quantile.py:143-145 */

/* LinearTrendTransformer.fit for every (cell, group): least-squares line of the group's series (time
 * order) on its positions 0..len-1, float64.  slope / intercept: [n_groups, ld_out].   trend.py:40-52 */
int sdb_group_trend(const void* v, int dtype, int64_t ld, int64_t n_cells,
                    const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                    double* slope, double* intercept, int64_t ld_out,
                    const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/* Remove / restore the lines of sdb_group_trend; j = position of the step inside its group.
 * intercept_ref (RESTORE): the intercepts found at FIT time, indexed like slope / intercept. */
int sdb_trend_apply(int mode, const void* v, int dtype, int64_t ld, int64_t n_cells,
                    const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                    const double* slope, const double* intercept, const double* intercept_ref, int64_t ld_coef,
                    double* out, int64_t ld_out, const uint8_t* cell_valid, void* stream);

/* BcsdTemperature.predict up to the mapper (bcsd.py:247-256): shift = centred 9-sample mean of the
 * climate-trend group - x_climo, key = X - shift, both float64 [T, ld_out].  roll_nbr as in sdb_qm_predict. */
int sdb_bcsd_shift(const void* X, int dtype, int64_t ld, int64_t n_cells,
                   const int32_t* rows, const int32_t* len, const int32_t* state_gid, int n_groups, int max_len,
                   const int32_t* roll_nbr, const void* x_climo, int64_t ld_climo,
                   double* shift, double* key, int64_t ld_out,
                   const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/* ... and after it (bcsd.py:263-269, 170-185): BCSD_T out = shift + mapped [- y_climo];
 * BCSD_P out = mapped [/ y_climo]; QM out = mapped.  mapped / shift: float64 [T, ld_in]. */
int sdb_bcsd_combine(int mode, const double* mapped, const double* shift, int64_t ld_in, int64_t n_cells,
                     const int32_t* rows, const int32_t* len, const int32_t* state_gid, int n_groups, int max_len,
                     const void* y_climo, int climo_dtype, int64_t ld_climo, int return_anoms,
                     void* out, int out_dtype, int64_t ld_out, const uint8_t* cell_valid, void* stream);

/*
 * PureRegression (gard.py:367-504) — SURVEY.md §8(f) row 4: one least-squares fit per cell on the rows whose
 * target exceeds thresh (all rows without one), its in-sample RMSE, and — with a threshold — the logistic
 * exceedance model of all rows (exceedance_prob = P(class 1), gard.py:467; solved to its optimum like
 * sdb_analog_predict's).  model: device float64 [n_cells, sdb_pure_regression_model_ld()], private layout.
 * A cell with NO row above thresh makes the reference's LinearRegression raise on an empty selection
 * (gard.py:435): the kernel ORs 4 into *nonfinite.  float32 inputs: the reference fits in float32 through
 * LAPACK, here the accumulation is float64 — parity to the stated tolerance, not to the bit.
 *   X_train [T_fit, p, C], y_train [T_fit, C], X_query [T_q, p, C] (dtype); out [T_q, 3, C] (out_dtype).
 */
int sdb_pure_regression_model_ld(void);
int sdb_pure_regression_fit(const void* X_train, const void* y_train, int dtype, int64_t ld, int64_t n_cells,
                            int t_fit, int n_features, int has_thresh, double thresh, double logistic_c,
                            double* model, const uint8_t* cell_valid, int32_t* nonfinite, void* stream);
int sdb_pure_regression_predict(const void* X_query, int dtype, int64_t ld, int64_t n_cells, int t_query,
                                int n_features, const double* model, void* out, int out_dtype, int64_t ld_out,
                                const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/* ------------------------------------------------------------------ ZScoreRegressor (zscore.py:11-353)
 * fit (zscore.py:32-66, 124-239): the record is laid out as [year, day of year] by the caller's table
 *   day_rows[n_years * n_days] (row of the record, -1 = that year has no such day); pos_col[n_days + window] maps the
 *   positions of the bookended year (last ceil(w/2) day columns | all | first w/2) to day columns, col_count[n_days]
 *   is the number of years holding each column; retained window k (0 <= k < n_kept <= n_days) pools positions
 *   k+1 .. k+window of ALL years.  shift / scale: [n_kept, ld_out] in the input dtype (mean_y - mean_X, std_y / std_X,
 *   population std); stats (optional): [4, n_kept, ld_out] = X_mean, X_std, y_mean, y_std (fit_stats_dict_).
 *   workspace: sdb_zscore_workspace_bytes(n_cells, n_days) bytes of device memory. */
int64_t sdb_zscore_workspace_bytes(int64_t n_cells, int n_days);
int sdb_zscore_fit(const void* X, const void* y, int dtype, int64_t ld, int64_t n_cells,
                   const int32_t* day_rows, int n_years, int n_days,
                   const int32_t* pos_col, const int32_t* col_count, int window, int n_kept,
                   void* workspace, void* shift, void* scale, void* stats, int64_t ld_out,
                   const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

/* predict (zscore.py:68-110, 242-353): centred rolling mean / sample standard deviation over `window` steps (NaN where
 * the window is incomplete), z-score, corrected with shift / scale[t mod min(n_steps, 364)].  SDB_E_INVALID when
 * fewer than min(n_steps, 364) fitted values exist (the reference's positional IndexError, zscore.py:314). */
int sdb_zscore_predict(const void* X, int dtype, int64_t ld, int64_t n_cells, int n_steps, int window,
                       const void* shift, const void* scale, int64_t ld_stats, int n_stats,
                       void* out, int out_dtype, int64_t ld_out,
                       const uint8_t* cell_valid, int32_t* nonfinite, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDB_H_ */

// qm_fused.cu — sdb_bcsd_fit_predict: fit + predict in one pass (qm_fused.cuh) behind the C ABI.
#include "qm_kernels.cuh"
#include "qm_fused.cuh"

using namespace sdb;

extern "C" int sdb_bcsd_fit_predict(int mode, const void* X_train, const void* y_train, const void* X_pred, int dtype,
                                    int64_t ld_train, int64_t ld_pred, int64_t n_cells,
                                    const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                                    void* x_climo, void* y_climo, int64_t ld_climo, int return_anoms,
                                    void* sorted_state, int64_t state_ld, const int64_t* state_off,
                                    void* out, int64_t ld_out,
                                    const uint8_t* cell_valid, int32_t* nonfinite, uint64_t* stats, void* stream) {
    if (!y_train || !X_pred || !rows || !len || !out) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: NULL pointer");
    if (mode != SDB_MODE_QM && mode != SDB_MODE_BCSD_P && mode != SDB_MODE_BCSD_T)
        return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: unknown mode %d", mode);
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld_train < n_cells || ld_pred < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: bad shape");
    if (dtype != SDB_F32) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_bcsd_fit_predict: float32 only (use sdb_qm_fit + sdb_qm_predict)");
    if (max_len > 1024) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_bcsd_fit_predict: groups of up to 1024 steps (got %d)", max_len);
    if (ld_train >= (1LL << 32) || ld_pred >= (1LL << 32) || ld_out >= (1LL << 32))
        return sdb_fail(SDB_E_UNSUPPORTED, "sdb_bcsd_fit_predict: row stride must be below 2^32 elements");
    if (mode == SDB_MODE_BCSD_T && (!X_train || !x_climo)) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: BCSD_T needs X_train and x_climo");
    if (mode != SDB_MODE_QM && return_anoms && !y_climo) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: return_anoms needs y_climo");
    if ((x_climo || y_climo) && ld_climo < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: bad ld_climo");
    if (sorted_state && (!state_off || state_ld <= 0)) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_fit_predict: sorted_state needs state_off / state_ld");
    // climatologies first (bcsd.py:138, 222-223): the map needs them
    if (mode != SDB_MODE_QM && y_climo) {
        const int rc = sdb_group_mean(y_train, dtype, ld_train, n_cells, rows, len, n_groups, max_len, SDB_MEAN_GROUPBY,
                                      y_climo, ld_climo, cell_valid, nonfinite, stream);
        if (rc) return rc;
    }
    if (mode == SDB_MODE_BCSD_T) {
        const int rc = sdb_group_mean(X_train, dtype, ld_train, n_cells, rows, len, n_groups, max_len, SDB_MEAN_GROUPBY,
                                      x_climo, ld_climo, cell_valid, nonfinite, stream);
        if (rc) return rc;
    }
    FusedParams p;
    p.y = (const float*)y_train; p.ld_y = ld_train; p.X = (const float*)X_pred; p.ld_x = ld_pred; p.C = n_cells;
    p.rows = rows; p.len = len; p.max_len = max_len; p.n_groups = n_groups;
    p.x_climo = (const float*)x_climo; p.y_climo = (const float*)y_climo; p.ld_climo = ld_climo;
    p.mode = mode; p.return_anoms = return_anoms;
    p.out = (float*)out; p.ld_out = ld_out;
    p.state = (float*)sorted_state; p.state_ld = state_ld; p.state_off = state_off;
    p.valid = cell_valid; p.nonfinite = nonfinite; p.stats = (unsigned long long*)stats;
    p.no_vec = (g_debug_flags & 2) != 0;
    p.force_network = (g_debug_flags >> 2) & 3;
    cudaStream_t st = (cudaStream_t)stream;
    return mode == SDB_MODE_BCSD_T ? launch_fused<true>(p, st) : launch_fused<false>(p, st);
}

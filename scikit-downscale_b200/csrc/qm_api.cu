// qm_api.cu — extern "C" entry points of the quantile-mapping path (include/sdb.h).
#include "qm_kernels.cuh"
#include "np_pairwise.cuh"

namespace sdb {

// ---------------------------------------------------------------- group means (climatologies)
template <typename T>
struct RowReader {
    const T* v; int64_t ld; int64_t c; const int32_t* rg;
    __device__ __forceinline__ T operator()(int j) const { return v[(int64_t)rg[j] * ld + c]; }
};

template <typename T>
__global__ void group_mean_kernel(const T* __restrict__ v, int64_t ld, int64_t C,
                                  const int32_t* __restrict__ rows, const int32_t* __restrict__ len,
                                  int max_len, int how, T* __restrict__ climo, int64_t ld_out,
                                  const uint8_t* __restrict__ valid, int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= C) return;
    T* dst = climo + (int64_t)g * ld_out + c;
    if (valid && !valid[c]) { *dst = (T)NAN; return; }
    const int n = len[g];
    RowReader<T> a{v, ld, c, rows + (int64_t)g * max_len};
    if (n <= 0) { *dst = (T)NAN; return; }
    if (how == SDB_MEAN_GROUPBY) {
        // pandas group_mean: Kahan-compensated running sum in the input dtype, time order
        T s = (T)0, comp = (T)0;
        for (int j = 0; j < n; ++j) {
            T x = a(j);
            T y = x - comp;
            T t = s + y;
            comp = (t - s) - y;
            if (comp != comp) comp = (T)0;
            s = t;
        }
        flag_nonfinite(s, nonfinite);
        *dst = s / (T)n;
    } else if (how == SDB_MEAN_NUMPY) {
        // contiguous 1-D ndarray.sum in the input dtype: pairwise over all n (a one-column DataFrame.mean())
        const T s = np_pairwise<T>(a, 0, n);
        flag_nonfinite(s, nonfinite);
        *dst = s / (T)n;
    } else {
        // ndarray.sum: first element copied as the initial value, pairwise over the rest
        T s = a(0);
        if (n > 1) s = s + np_pairwise<T>(a, 1, n - 1);
        flag_nonfinite(s, nonfinite);
        *dst = s / (T)n;
    }
}


static int pick_np(int max_len) {
    if (max_len <= 256) return 256;
    if (max_len <= 1024) return 1024;
    if (max_len <= 4096) return 4096;
    if (max_len <= SDB_MAX_GROUP_LEN) return 16384;
    return -1;
}


}  // namespace sdb

using namespace sdb;

extern "C" int sdb_max_group_len(void) { return SDB_MAX_GROUP_LEN; }

extern "C" int sdb_group_mean(const void* v, int dtype, int64_t ld, int64_t n_cells,
                              const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                              int how, void* climo, int64_t ld_out,
                              const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!v || !rows || !len || !climo) return sdb_fail(SDB_E_INVALID, "sdb_group_mean: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_group_mean: bad shape");
    if (how != SDB_MEAN_GROUPBY && how != SDB_MEAN_FRAME && how != SDB_MEAN_NUMPY) return sdb_fail(SDB_E_INVALID, "sdb_group_mean: bad how");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((n_cells + 127) / 128), (unsigned)n_groups);
    if (dtype == SDB_F32)
        group_mean_kernel<float><<<grid, 128, 0, st>>>((const float*)v, ld, n_cells, rows, len, max_len, how, (float*)climo, ld_out, cell_valid, nonfinite);
    else if (dtype == SDB_F64)
        group_mean_kernel<double><<<grid, 128, 0, st>>>((const double*)v, ld, n_cells, rows, len, max_len, how, (double*)climo, ld_out, cell_valid, nonfinite);
    else return sdb_fail(SDB_E_INVALID, "sdb_group_mean: bad dtype %d", dtype);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_qm_fit(const void* y, int dtype, int64_t ld, int64_t n_cells,
                          const int32_t* rows, const int32_t* len, const int64_t* state_off,
                          int n_groups, int max_len, void* sorted_state, int64_t state_ld,
                          const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!y || !rows || !len || !state_off || !sorted_state) return sdb_fail(SDB_E_INVALID, "sdb_qm_fit: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || state_ld <= 0)
        return sdb_fail(SDB_E_INVALID, "sdb_qm_fit: bad shape");
    if (ld >= (1LL << 32)) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_qm_fit: row stride must be below 2^32 elements");
    if (dtype != SDB_F32 && dtype != SDB_F64) return sdb_fail(SDB_E_INVALID, "sdb_qm_fit: bad dtype %d", dtype);
    FitParams f{y, ld, n_cells, rows, len, state_off, n_groups, max_len, sorted_state, state_ld, cell_valid, nonfinite,
                (g_debug_flags & 2) != 0};
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SDB_F32 && !(g_debug_flags & 1)) {
        if (max_len <= 256) return qm_fit_tile_np256(f, st);
        if (max_len <= 1024) return qm_fit_tile_np1024(f, st);
        if (max_len <= SDB_MAX_GROUP_LEN) return qm_fit_long(f, st);
    }
    switch (pick_np(max_len)) {
        case 256:   return qm_fit_np256(dtype, f, st);
        case 1024:  return qm_fit_np1024(dtype, f, st);
        case 4096:  return qm_fit_np4096(dtype, f, st);
        case 16384: return qm_fit_np16384(dtype, f, st);
    }
    return sdb_fail(SDB_E_UNSUPPORTED, "sdb_qm_fit: group length %d > %d", max_len, SDB_MAX_GROUP_LEN);
}

extern "C" int sdb_qm_predict(int mode, const void* X, int dtype, int64_t ld, int64_t n_cells,
                              const int32_t* rows, const int32_t* len, const int32_t* state_gid,
                              int n_groups, int max_len,
                              const int32_t* fit_len, const int64_t* state_off, int max_fit_len,
                              const void* sorted_state, int64_t state_ld,
                              const void* x_climo, const void* y_climo, int64_t ld_climo,
                              int return_anoms, const int32_t* roll_nbr, const sdb_cunnane_opts* cunnane,
                              void* out, int out_dtype, int64_t ld_out, int32_t* rank_out,
                              const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !rows || !len || !state_gid || !fit_len || !state_off || !sorted_state || !out)
        return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || max_fit_len <= 0 || ld < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: bad shape");
    if (ld >= (1LL << 32) || ld_out >= (1LL << 32)) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_qm_predict: row stride must be below 2^32 elements");
    if (mode != SDB_MODE_QM && mode != SDB_MODE_BCSD_P && mode != SDB_MODE_BCSD_T)
        return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: unknown mode %d", mode);
    if ((dtype != SDB_F32 && dtype != SDB_F64) || (out_dtype != SDB_F32 && out_dtype != SDB_F64))
        return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: bad dtype %d / out_dtype %d", dtype, out_dtype);
    if (mode == SDB_MODE_BCSD_T && !x_climo) return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: BCSD_T needs x_climo");
    if (mode != SDB_MODE_QM && return_anoms && !y_climo) return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: return_anoms needs y_climo");
    if (mode != SDB_MODE_BCSD_T && roll_nbr) return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: roll_nbr only applies to BCSD_T");
    PredictParams p;
    p.X = X; p.ld = ld; p.C = n_cells; p.rows = rows; p.len = len; p.state_gid = state_gid; p.max_len = max_len;
    p.fit_len = fit_len; p.state_off = state_off; p.state = sorted_state; p.state_ld = state_ld;
    p.x_climo = x_climo; p.y_climo = y_climo; p.ld_climo = ld_climo; p.return_anoms = return_anoms;
    p.roll_nbr = roll_nbr; p.out = out; p.ld_out = ld_out; p.rank_out = rank_out; p.valid = cell_valid; p.nonfinite = nonfinite;
    p.mode = mode; p.out_f64 = (out_dtype == SDB_F64); p.n_groups = n_groups;
    p.no_vec = (g_debug_flags & 2) != 0;
    p.rank_only = 0; p.rank_ordinal = 0;
    p.alpha = 0.4; p.beta = 0.4; p.n_endpoints = 10; p.extrap_lo = 1; p.extrap_hi = 1;   // quantile.py:420-432 defaults
    if (cunnane) {
        if (cunnane->n_endpoints < 1) return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: n_endpoints must be >= 1");
        if (cunnane->extrapolate < SDB_EXTRAPOLATE_NONE || cunnane->extrapolate > SDB_EXTRAPOLATE_BOTH)
            return sdb_fail(SDB_E_INVALID, "sdb_qm_predict: unknown extrapolate code %d", cunnane->extrapolate);
        p.alpha = cunnane->alpha; p.beta = cunnane->beta; p.n_endpoints = cunnane->n_endpoints;
        p.extrap_lo = (cunnane->extrapolate & SDB_EXTRAPOLATE_MIN) != 0;
        p.extrap_hi = (cunnane->extrapolate & SDB_EXTRAPOLATE_MAX) != 0;
    }
    const int kind = (mode != SDB_MODE_BCSD_T) ? KIND_RAW : (roll_nbr ? KIND_SHIFT_TAB : KIND_SHIFT);
    cudaStream_t st = (cudaStream_t)stream;
    const int longest = max_len > max_fit_len ? max_len : max_fit_len;
    if (dtype == SDB_F32 && out_dtype == SDB_F32 && kind != KIND_SHIFT_TAB && longest <= 1024 && !(g_debug_flags & 1)) {
        if (longest <= 256) return qm_predict_tile_np256(kind, p, st);
        return qm_predict_tile_np1024(kind, p, st);
    }
    if (dtype == SDB_F32 && kind == KIND_RAW && max_len > 1024 && max_len <= SDB_MAX_GROUP_LEN && !(g_debug_flags & 1))
        return qm_predict_long(p, st);
    switch (pick_np(max_len)) {
        case 256:   return qm_predict_np256(dtype, kind, p, st);
        case 1024:  return qm_predict_np1024(dtype, kind, p, st);
        case 4096:  return qm_predict_np4096(dtype, kind, p, st);
        case 16384: return qm_predict_np16384(dtype, kind, p, st);
    }
    return sdb_fail(SDB_E_UNSUPPORTED, "sdb_qm_predict: group length %d > %d", max_len, SDB_MAX_GROUP_LEN);
}

extern "C" int sdb_series_rank(const void* X, int dtype, int64_t ld, int64_t n_cells,
                               const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                               int ordinal, int32_t* rank_out, int64_t ld_rank,
                               const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !rows || !len || !rank_out) return sdb_fail(SDB_E_INVALID, "sdb_series_rank: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || ld_rank < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_series_rank: bad shape");
    if (ld >= (1LL << 32) || ld_rank >= (1LL << 32)) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_series_rank: row stride must be below 2^32 elements");
    if (dtype != SDB_F32 && dtype != SDB_F64) return sdb_fail(SDB_E_INVALID, "sdb_series_rank: bad dtype %d", dtype);
    PredictParams p;
    memset(&p, 0, sizeof(p));
    p.X = X; p.ld = ld; p.C = n_cells; p.rows = rows; p.len = len; p.max_len = max_len;
    p.rank_out = rank_out; p.ld_out = ld_rank; p.valid = cell_valid; p.nonfinite = nonfinite;
    p.mode = SDB_MODE_QM; p.n_groups = n_groups; p.rank_only = 1; p.rank_ordinal = ordinal != 0;
    p.alpha = 0.4; p.beta = 0.4; p.n_endpoints = 10; p.extrap_lo = 1; p.extrap_hi = 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pick_np(max_len)) {
        case 256:   return qm_predict_np256(dtype, KIND_RAW, p, st);
        case 1024:  return qm_predict_np1024(dtype, KIND_RAW, p, st);
        case 4096:  return qm_predict_np4096(dtype, KIND_RAW, p, st);
        case 16384: return qm_predict_np16384(dtype, KIND_RAW, p, st);
    }
    return sdb_fail(SDB_E_UNSUPPORTED, "sdb_series_rank: group length %d > %d", max_len, SDB_MAX_GROUP_LEN);
}

// trend_kernels.cu — the pieces of the detrending quantile map, for every (cell, group), sm_100a.
//
// QuantileMapper(detrend=True) (quantile.py:94-98, 127-145; also reached per time group through
// BcsdBase(qm_kwargs={'detrend': True}), bcsd.py:65-67) removes a least-squares line from a series
// before ranking it and puts the line back afterwards:
//   LinearTrendTransformer.fit        trend.py:40-52   OLS of the series on arange(n)
//   ... .transform / inverse          trend.py:54-83   X -+ (arange(n) * slope + intercept), float64
// Here the trend, its removal / restoration and the BCSD shift / combine steps around the mapper are
// small HBM-bound kernels (thread = cell, rows coalesced across cells) composed with the generic
// float64 sdb_qm_fit / sdb_qm_predict: a non-default option, kept off the tile fast path.
// SURVEY.md §8(f) row 2.  Compiled with -fmad=false (separate multiply and add, like numpy).
#include "qm_kernels.cuh"

namespace sdb {

// slope / intercept of sklearn's LinearRegression of group g's series (time order) on 0..n-1:
// centred least squares in float64.
template <typename T>
__global__ void group_trend_kernel(const T* __restrict__ v, int64_t ld, int64_t C,
                                   const int32_t* __restrict__ rows, const int32_t* __restrict__ len, int max_len,
                                   double* __restrict__ slope, double* __restrict__ icpt, int64_t ld_out,
                                   const uint8_t* __restrict__ valid, int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= C) return;
    const int64_t at = (int64_t)g * ld_out + c;
    if (valid && !valid[c]) { slope[at] = NAN; icpt[at] = NAN; return; }
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    if (n <= 0) { slope[at] = NAN; icpt[at] = NAN; return; }
    double sy = 0.0;
    for (int j = 0; j < n; ++j) sy += (double)v[(int64_t)rg[j] * ld + c];
    const double ym = sy / (double)n, tm = 0.5 * (double)(n - 1);
    double sty = 0.0, stt = 0.0;
    for (int j = 0; j < n; ++j) {
        const double dt = (double)j - tm;
        sty += dt * ((double)v[(int64_t)rg[j] * ld + c] - ym);
        stt += dt * dt;
    }
    if (nonfinite && !isfinite(sy)) atomicOr(nonfinite, 1);
    const double s = stt > 0.0 ? sty / stt : 0.0;
    slope[at] = s;
    icpt[at] = ym - tm * s;
}

// mode 0: out = v - (j * slope + icpt)                        X - trendline(X)           trend.py:64
// mode 1: out = (v + (j * slope + icpt)) - (icpt - icpt_ref)  inverse_transform, then the baseline reset  quantile.py:143-145
template <typename T>
__global__ void trend_apply_kernel(int mode, const T* __restrict__ v, int64_t ld, int64_t C,
                                   const int32_t* __restrict__ rows, const int32_t* __restrict__ len, int max_len,
                                   const double* __restrict__ slope, const double* __restrict__ icpt,
                                   const double* __restrict__ icpt_ref, int64_t ld_coef,
                                   double* __restrict__ out, int64_t ld_out, const uint8_t* __restrict__ valid) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= C) return;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    const bool ok = !valid || valid[c];
    const int64_t k = (int64_t)g * ld_coef + c;
    const double s = ok ? slope[k] : 0.0, b = ok ? icpt[k] : 0.0;
    const double reset = (mode == 1 && ok) ? b - icpt_ref[k] : 0.0;
    for (int j = 0; j < n; ++j) {
        const int64_t row = rg[j];
        double r = NAN;
        if (ok) {
            const double line = (double)j * s + b;
            const double x = (double)v[row * ld + c];
            r = (mode == 0) ? x - line : (x + line) - reset;
        }
        out[row * ld_out + c] = r;
    }
}

// BcsdTemperature.predict up to the mapper: shift = rolling mean - x_climo, key = X - shift   bcsd.py:247-256
template <typename T, bool ROLLTAB>
__global__ void bcsd_shift_kernel(const T* __restrict__ X, int64_t ld, int64_t C,
                                  const int32_t* __restrict__ rows, const int32_t* __restrict__ len, int max_len,
                                  const int32_t* __restrict__ state_gid, const int32_t* __restrict__ roll_nbr,
                                  const T* __restrict__ x_climo, int64_t ld_climo,
                                  double* __restrict__ shift, double* __restrict__ key, int64_t ld_out,
                                  const uint8_t* __restrict__ valid, int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= C) return;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    const bool ok = !valid || valid[c];
    const double xc = ok ? (double)x_climo[(int64_t)state_gid[g] * ld_climo + c] : 0.0;
    for (int j = 0; j < n; ++j) {
        const int64_t at = (int64_t)rg[j] * ld_out + c;
        if (!ok) { shift[at] = NAN; key[at] = NAN; continue; }
        double x, s;
        shifted_value<T, ROLLTAB>(X, ld, c, rg, n, j, roll_nbr, xc, x, s);
        if (nonfinite && !isfinite(x)) atomicOr(nonfinite, 1);
        shift[at] = s;
        key[at] = x - s;
    }
}

// after the mapper: BCSD_T  out = shift + mapped [- y_climo]   bcsd.py:263-269
//                   BCSD_P  out = mapped [/ y_climo]           bcsd.py:170-185;   QM  out = mapped
template <typename T>
__global__ void bcsd_combine_kernel(int mode, const double* __restrict__ mapped, const double* __restrict__ shift,
                                    int64_t ld_in, int64_t C,
                                    const int32_t* __restrict__ rows, const int32_t* __restrict__ len, int max_len,
                                    const int32_t* __restrict__ state_gid, const T* __restrict__ y_climo, int64_t ld_climo,
                                    int return_anoms, void* __restrict__ out, int out_f64, int64_t ld_out,
                                    const uint8_t* __restrict__ valid) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= C) return;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    const bool ok = !valid || valid[c];
    const bool anoms = return_anoms && mode != SDB_MODE_QM;
    const double yc = (ok && anoms) ? (double)y_climo[(int64_t)state_gid[g] * ld_climo + c] : 0.0;
    for (int j = 0; j < n; ++j) {
        const int64_t row = rg[j];
        double o = NAN;
        if (ok) {
            const double q = mapped[row * ld_in + c];
            if (mode == SDB_MODE_BCSD_T) { o = shift[row * ld_in + c] + q; if (anoms) o = o - yc; }
            else if (mode == SDB_MODE_BCSD_P) o = anoms ? q / yc : q;
            else o = q;
        }
        store_out(out, out_f64, row * ld_out + c, o);
    }
}

}  // namespace sdb

using namespace sdb;

static inline dim3 cell_group_grid(int64_t n_cells, int n_groups) { return dim3((unsigned)((n_cells + 127) / 128), (unsigned)n_groups); }

extern "C" int sdb_group_trend(const void* v, int dtype, int64_t ld, int64_t n_cells,
                               const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                               double* slope, double* intercept, int64_t ld_out,
                               const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!v || !rows || !len || !slope || !intercept) return sdb_fail(SDB_E_INVALID, "sdb_group_trend: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_group_trend: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = cell_group_grid(n_cells, n_groups);
    if (dtype == SDB_F32) group_trend_kernel<float><<<grid, 128, 0, st>>>((const float*)v, ld, n_cells, rows, len, max_len, slope, intercept, ld_out, cell_valid, nonfinite);
    else if (dtype == SDB_F64) group_trend_kernel<double><<<grid, 128, 0, st>>>((const double*)v, ld, n_cells, rows, len, max_len, slope, intercept, ld_out, cell_valid, nonfinite);
    else return sdb_fail(SDB_E_INVALID, "sdb_group_trend: bad dtype %d", dtype);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_trend_apply(int mode, const void* v, int dtype, int64_t ld, int64_t n_cells,
                               const int32_t* rows, const int32_t* len, int n_groups, int max_len,
                               const double* slope, const double* intercept, const double* intercept_ref, int64_t ld_coef,
                               double* out, int64_t ld_out, const uint8_t* cell_valid, void* stream) {
    if (!v || !rows || !len || !slope || !intercept || !out) return sdb_fail(SDB_E_INVALID, "sdb_trend_apply: NULL pointer");
    if (mode != SDB_TREND_REMOVE && mode != SDB_TREND_RESTORE) return sdb_fail(SDB_E_INVALID, "sdb_trend_apply: unknown mode %d", mode);
    if (mode == SDB_TREND_RESTORE && !intercept_ref) return sdb_fail(SDB_E_INVALID, "sdb_trend_apply: restore needs the fitted intercepts");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || ld_out < n_cells || ld_coef < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_trend_apply: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = cell_group_grid(n_cells, n_groups);
    if (dtype == SDB_F32) trend_apply_kernel<float><<<grid, 128, 0, st>>>(mode, (const float*)v, ld, n_cells, rows, len, max_len, slope, intercept, intercept_ref, ld_coef, out, ld_out, cell_valid);
    else if (dtype == SDB_F64) trend_apply_kernel<double><<<grid, 128, 0, st>>>(mode, (const double*)v, ld, n_cells, rows, len, max_len, slope, intercept, intercept_ref, ld_coef, out, ld_out, cell_valid);
    else return sdb_fail(SDB_E_INVALID, "sdb_trend_apply: bad dtype %d", dtype);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_bcsd_shift(const void* X, int dtype, int64_t ld, int64_t n_cells,
                              const int32_t* rows, const int32_t* len, const int32_t* state_gid, int n_groups, int max_len,
                              const int32_t* roll_nbr, const void* x_climo, int64_t ld_climo,
                              double* shift, double* key, int64_t ld_out,
                              const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !rows || !len || !state_gid || !x_climo || !shift || !key) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_shift: NULL pointer");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld < n_cells || ld_out < n_cells || ld_climo < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_bcsd_shift: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = cell_group_grid(n_cells, n_groups);
#define SDB_SHIFT_LAUNCH(T, TAB) bcsd_shift_kernel<T, TAB><<<grid, 128, 0, st>>>((const T*)X, ld, n_cells, rows, len, max_len, state_gid, roll_nbr, (const T*)x_climo, ld_climo, shift, key, ld_out, cell_valid, nonfinite)
    if (dtype == SDB_F32) { if (roll_nbr) SDB_SHIFT_LAUNCH(float, true); else SDB_SHIFT_LAUNCH(float, false); }
    else if (dtype == SDB_F64) { if (roll_nbr) SDB_SHIFT_LAUNCH(double, true); else SDB_SHIFT_LAUNCH(double, false); }
    else return sdb_fail(SDB_E_INVALID, "sdb_bcsd_shift: bad dtype %d", dtype);
#undef SDB_SHIFT_LAUNCH
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_bcsd_combine(int mode, const double* mapped, const double* shift, int64_t ld_in, int64_t n_cells,
                                const int32_t* rows, const int32_t* len, const int32_t* state_gid, int n_groups, int max_len,
                                const void* y_climo, int climo_dtype, int64_t ld_climo, int return_anoms,
                                void* out, int out_dtype, int64_t ld_out, const uint8_t* cell_valid, void* stream) {
    if (!mapped || !rows || !len || !out) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: NULL pointer");
    if (mode != SDB_MODE_QM && mode != SDB_MODE_BCSD_P && mode != SDB_MODE_BCSD_T) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: unknown mode %d", mode);
    if (mode == SDB_MODE_BCSD_T && !shift) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: BCSD_T needs the shift");
    if (mode != SDB_MODE_QM && return_anoms && (!y_climo || !state_gid)) return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: return_anoms needs y_climo");
    if (n_cells <= 0 || n_groups <= 0 || max_len <= 0 || ld_in < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: bad shape");
    if ((out_dtype != SDB_F32 && out_dtype != SDB_F64) || (climo_dtype != SDB_F32 && climo_dtype != SDB_F64))
        return sdb_fail(SDB_E_INVALID, "sdb_bcsd_combine: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = cell_group_grid(n_cells, n_groups);
    if (climo_dtype == SDB_F32) bcsd_combine_kernel<float><<<grid, 128, 0, st>>>(mode, mapped, shift, ld_in, n_cells, rows, len, max_len, state_gid, (const float*)y_climo, ld_climo, return_anoms, out, out_dtype == SDB_F64, ld_out, cell_valid);
    else bcsd_combine_kernel<double><<<grid, 128, 0, st>>>(mode, mapped, shift, ld_in, n_cells, rows, len, max_len, state_gid, (const double*)y_climo, ld_climo, return_anoms, out, out_dtype == SDB_F64, ld_out, cell_valid);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

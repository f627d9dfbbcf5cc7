// zscore_kernels.cu — ZScoreRegressor for every cell at once, sm_100a.
//
// Reference: skdownscale/pointwise_models/zscore.py
//   fit      :32-66   _calc_stats (:161-193) of X and y, shift = mean_y - mean_X, scale = std_y / std_X (:196-239)
//   _reshape :124-158 the record as [year, day of year], bookended with the same year's last ceil(w/2) and first
//                     w//2 day columns; _calc_stats pools ALL years and a centred window of w day columns
//   predict  :68-110  centred rolling mean / sample std of the new series (pandas, float64), z-score, corrected by
//                     the fitted values repeated every min(n, 364) steps (_expand_params :278-318)
//
// Everything is streaming work over [time, cell] arrays with the cell index as the fast axis (thread = cell, rows
// coalesced across cells): HBM-bound, no shared memory, no tensor cores.
//   zscore_daysum_kernel   thread = (cell, day column): sum / sum of squares over the years        reads X, y once
//   zscore_window_kernel   thread = (cell, 32 consecutive windows): sliding sum over the day columns
//   zscore_predict_kernel  thread = (cell, 256 consecutive steps): sliding window sums over time   reads X, writes out
// Sums are float64 of (value - first value of the series / segment): exact differences of the float32 inputs, so a
// constant window has variance exactly 0 and nothing cancels catastrophically.
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#include "../../include/sdb.h"
#include "common.cuh"

namespace sdb {

constexpr int ZS_SEG_WIN = 32;      // windows per thread in zscore_window_kernel
constexpr int ZS_SEG_T = 364;       // steps per thread in zscore_predict_kernel = the reference's "average year" (zscore.py:300)

// ws layout: [4][n_days][C] float64 = {sum_X, sumsq_X, sum_y, sumsq_y} of (v - v[row 0]) per day column
template <typename T>
__global__ void zscore_daysum_kernel(const T* __restrict__ X, const T* __restrict__ y, int64_t ld, int64_t C,
                                     const int32_t* __restrict__ day_rows, int n_years, int n_days,
                                     double* __restrict__ ws, const uint8_t* __restrict__ valid, int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (c >= C || (valid && !valid[c])) return;
    const double rx = (double)X[c], ry = (double)y[c];          // row 0 of the record
    double s1x = 0.0, s2x = 0.0, s1y = 0.0, s2y = 0.0;
    bool bad = false;
    for (int yr = 0; yr < n_years; ++yr) {
        const int row = day_rows[yr * n_days + d];
        if (row < 0) continue;
        const T xv = X[(int64_t)row * ld + c], yv = y[(int64_t)row * ld + c];
        bad |= !isfinite((double)xv) || !isfinite((double)yv);
        const double dx = (double)xv - rx, dy = (double)yv - ry;
        s1x += dx; s2x += dx * dx; s1y += dy; s2y += dy * dy;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    const int64_t plane = (int64_t)n_days * C;
    double* w = ws + (int64_t)d * C + c;
    w[0] = s1x; w[plane] = s2x; w[2 * plane] = s1y; w[3 * plane] = s2y;
}

// pos_col[p], p in [0, n_days + w): the day column behind position p of the bookended year; window k covers positions
// k + 1 .. k + w (oracle/zscore.py::zscore_window_columns).  col_count[d]: years that have day column d.
template <typename T>
__global__ void zscore_window_kernel(const T* __restrict__ X, const T* __restrict__ y, int64_t C,
                                     const double* __restrict__ ws, const int32_t* __restrict__ pos_col,
                                     const int32_t* __restrict__ col_count, int n_days, int window, int n_kept,
                                     T* __restrict__ shift, T* __restrict__ scale, T* __restrict__ stats, int64_t ld_out,
                                     const uint8_t* __restrict__ valid) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = blockIdx.y * ZS_SEG_WIN;
    if (c >= C) return;
    const int k1 = min(n_kept, k0 + ZS_SEG_WIN);
    if (valid && !valid[c]) {
        for (int k = k0; k < k1; ++k) {
            shift[(int64_t)k * ld_out + c] = (T)NAN;
            scale[(int64_t)k * ld_out + c] = (T)NAN;
            if (stats) for (int s = 0; s < 4; ++s) stats[((int64_t)s * n_kept + k) * ld_out + c] = (T)NAN;
        }
        return;
    }
    const int64_t plane = (int64_t)n_days * C;
    const double rx = (double)X[c], ry = (double)y[c];
    double s1x = 0.0, s2x = 0.0, s1y = 0.0, s2y = 0.0;
    int64_t n = 0;
    auto add = [&](int p, double sign) {
        const int col = pos_col[p];
        const double* w = ws + (int64_t)col * C + c;
        s1x += sign * w[0]; s2x += sign * w[plane]; s1y += sign * w[2 * plane]; s2y += sign * w[3 * plane];
        n += (sign > 0.0 ? 1 : -1) * (int64_t)col_count[col];
    };
    for (int p = k0 + 1; p <= k0 + window; ++p) add(p, 1.0);
    for (int k = k0; k < k1; ++k) {
        const double inv = n > 0 ? 1.0 / (double)n : NAN;
        const double mx = s1x * inv, my = s1y * inv;
        const double vx = fmax(s2x * inv - mx * mx, 0.0), vy = fmax(s2y * inv - my * my, 0.0);
        // the reference's statistics are arrays of the input dtype: round, then combine in that dtype (:237-238)
        const T mean_x = (T)(rx + mx), mean_y = (T)(ry + my), std_x = (T)sqrt(vx), std_y = (T)sqrt(vy);
        const int64_t at = (int64_t)k * ld_out + c;
        shift[at] = mean_y - mean_x;
        scale[at] = std_y / std_x;
        if (stats) {
            stats[((int64_t)0 * n_kept + k) * ld_out + c] = mean_x;
            stats[((int64_t)1 * n_kept + k) * ld_out + c] = std_x;
            stats[((int64_t)2 * n_kept + k) * ld_out + c] = mean_y;
            stats[((int64_t)3 * n_kept + k) * ld_out + c] = std_y;
        }
        if (k + 1 < k1) { add(k + 1, -1.0); add(k + 1 + window, 1.0); }
    }
}

template <typename T, typename TO>
__global__ void zscore_predict_kernel(const T* __restrict__ X, int64_t ld, int64_t C, int n_steps, int window,
                                      const T* __restrict__ shift, const T* __restrict__ scale, int64_t ld_stats, int len_avgyr,
                                      TO* __restrict__ out, int64_t ld_out, const uint8_t* __restrict__ valid,
                                      int32_t* __restrict__ nonfinite) {
    // grid: x = time segment (fast), y = 128-cell tile — the ~30 CTAs that need the same [364, 128-cell] block of shift /
    // scale run next to each other and find it in L2 (the two arrays are 2 x 189 MB for 129 600 cells: read once
    // instead of once per year); a segment is one period of the fitted values, so k = t - t0 (no modulo)
    const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int t0 = blockIdx.x * ZS_SEG_T;
    if (c >= C) return;
    const int t1 = min(n_steps, t0 + ZS_SEG_T);
    const int hi = (window - 1) / 2, lo = window - 1 - hi;       // pandas centred window of step t: [t - lo, t + hi]
    const bool ok = !valid || valid[c];
    const double invw = 1.0 / (double)window;
    const double invw1 = window > 1 ? 1.0 / (double)(window - 1) : NAN;
    double ref = 0.0, s1 = 0.0, s2 = 0.0;
    bool have = false, bad = false;
    for (int t = t0; t < t1; ++t) {
        const int64_t at = (int64_t)t * ld_out + c;
        if (!ok || t - lo < 0 || t + hi > n_steps - 1) { out[at] = (TO)NAN; have = false; continue; }
        if (!have) {
            ref = (double)X[(int64_t)t * ld + c];
            s1 = 0.0; s2 = 0.0;
            for (int u = t - lo; u <= t + hi; ++u) {
                const double v = (double)X[(int64_t)u * ld + c];
                bad |= !isfinite(v);
                const double d = v - ref;
                s1 += d; s2 += d * d;
            }
            have = true;
        } else {
            const double vin = (double)X[(int64_t)(t + hi) * ld + c], vout = (double)X[(int64_t)(t - lo - 1) * ld + c];
            bad |= !isfinite(vin);
            const double din = vin - ref, dout = vout - ref;
            s1 += din - dout;
            s2 += din * din - dout * dout;
        }
        const double x = (double)X[(int64_t)t * ld + c];
        const double m = s1 * invw;
        const double var = (s2 - s1 * m) * invw1;                  // sample variance (ddof = 1); NaN for window 1
        const int k = t - t0;                                     // = t mod min(n_steps, 364)
        const double sh = (double)shift[(int64_t)k * ld_stats + c], sc = (double)scale[(int64_t)k * ld_stats + c];
        // zscore.py:105-108: z * (std * scale) + (mean + shift) with z = (x - mean) / std.  std cancels (to an ulp
        // of float64) unless it is 0 or NaN, where the reference's 0 / 0 makes the step NaN: no sqrt, no division
        const double r = ((x - ref) - m) * sc + ((ref + m) + sh);
        out[at] = (TO)(var > 0.0 ? r : NAN);
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
}

}  // namespace sdb

using namespace sdb;

extern "C" int64_t sdb_zscore_workspace_bytes(int64_t n_cells, int n_days) {
    if (n_cells <= 0 || n_days <= 0) return 0;
    return (int64_t)4 * n_days * n_cells * (int64_t)sizeof(double);
}

extern "C" int sdb_zscore_fit(const void* X, const void* y, int dtype, int64_t ld, int64_t n_cells,
                              const int32_t* day_rows, int n_years, int n_days,
                              const int32_t* pos_col, const int32_t* col_count, int window, int n_kept,
                              void* workspace, void* shift, void* scale, void* stats, int64_t ld_out,
                              const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !y || !day_rows || !pos_col || !col_count || !workspace || !shift || !scale)
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_fit: NULL pointer");
    if (n_cells <= 0 || n_years <= 0 || n_days <= 0 || window <= 0 || n_kept <= 0 || ld < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_fit: bad shape");
    if (n_kept + window > n_days + window)
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_fit: %d windows of %d do not fit %d day columns", n_kept, window, n_days);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 ga((unsigned)((n_cells + 127) / 128), (unsigned)n_days);
    const dim3 gb((unsigned)((n_cells + 127) / 128), (unsigned)((n_kept + ZS_SEG_WIN - 1) / ZS_SEG_WIN));
    if (dtype == SDB_F32) {
        zscore_daysum_kernel<float><<<ga, 128, 0, st>>>((const float*)X, (const float*)y, ld, n_cells, day_rows, n_years, n_days,
                                                       (double*)workspace, cell_valid, nonfinite);
        zscore_window_kernel<float><<<gb, 128, 0, st>>>((const float*)X, (const float*)y, n_cells, (const double*)workspace, pos_col,
                                                       col_count, n_days, window, n_kept, (float*)shift, (float*)scale,
                                                       (float*)stats, ld_out, cell_valid);
    } else if (dtype == SDB_F64) {
        zscore_daysum_kernel<double><<<ga, 128, 0, st>>>((const double*)X, (const double*)y, ld, n_cells, day_rows, n_years, n_days,
                                                        (double*)workspace, cell_valid, nonfinite);
        zscore_window_kernel<double><<<gb, 128, 0, st>>>((const double*)X, (const double*)y, n_cells, (const double*)workspace, pos_col,
                                                        col_count, n_days, window, n_kept, (double*)shift, (double*)scale,
                                                        (double*)stats, ld_out, cell_valid);
    } else {
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_fit: bad dtype %d", dtype);
    }
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_zscore_predict(const void* X, int dtype, int64_t ld, int64_t n_cells, int n_steps, int window,
                                  const void* shift, const void* scale, int64_t ld_stats, int n_stats,
                                  void* out, int out_dtype, int64_t ld_out,
                                  const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !shift || !scale || !out) return sdb_fail(SDB_E_INVALID, "sdb_zscore_predict: NULL pointer");
    if (n_cells <= 0 || n_steps <= 0 || window <= 0 || ld < n_cells || ld_out < n_cells || ld_stats < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_predict: bad shape");
    const int len_avgyr = n_steps < 364 ? n_steps : 364;          // zscore.py:300
    if (n_stats < len_avgyr)                                      // shift.iloc[inds] out of bounds in the reference (:314)
        return sdb_fail(SDB_E_INVALID, "sdb_zscore_predict: %d fitted values, %d needed (positional indexers are out-of-bounds)", n_stats, len_avgyr);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)((n_steps + ZS_SEG_T - 1) / ZS_SEG_T), (unsigned)((n_cells + 127) / 128));
    if (grid.y > 65535u) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_zscore_predict: more than 8 388 480 cells in one call");
#define SDB_ZS_LAUNCH(T, TO) zscore_predict_kernel<T, TO><<<grid, 128, 0, st>>>((const T*)X, ld, n_cells, n_steps, window, (const T*)shift, \
        (const T*)scale, ld_stats, len_avgyr, (TO*)out, ld_out, cell_valid, nonfinite)
    if (dtype == SDB_F32 && out_dtype == SDB_F32) SDB_ZS_LAUNCH(float, float);
    else if (dtype == SDB_F32 && out_dtype == SDB_F64) SDB_ZS_LAUNCH(float, double);
    else if (dtype == SDB_F64 && out_dtype == SDB_F64) SDB_ZS_LAUNCH(double, double);
    else if (dtype == SDB_F64 && out_dtype == SDB_F32) SDB_ZS_LAUNCH(double, float);
    else return sdb_fail(SDB_E_INVALID, "sdb_zscore_predict: bad dtype %d / %d", dtype, out_dtype);
#undef SDB_ZS_LAUNCH
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

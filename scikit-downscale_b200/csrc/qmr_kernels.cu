// qmr_kernels.cu — CDF-to-CDF quantile-mapping regressors for every cell, sm_100a.
//
// Replaces, per cell, the predict side of
//   QuantileMappingReressor        skdownscale/pointwise_models/quantile.py:160-395
//   EquidistantCdfMatcher          skdownscale/pointwise_models/quantile.py:556-636
// (fit = np.sort of X and of y per cell — sdb_qm_fit on one whole-series group each — plus the two
// synthetic frame points of `_calc_extrapolated_cdf`, quantile.py:311-388, computed here by
// sdb_qmr_frame).  SURVEY.md §8(f) row 1.
//
// The "framed" CDF of a sorted series S[0..m) is the m + 2 point polyline
//   index 0        (pp_lo, v_lo)      synthetic lower point
//   index i=1..m   ((i - 0.4) / (m + 0.2), S[i-1])
//   index m+1      (pp_hi, v_hi)      synthetic upper point
// with (pp_lo, v_lo) = (-1e20, OLS line of the first n_endpoints points at -1e20) on a tail that
// extrapolates and (pp_1, S[0]) otherwise — likewise above.  Both regressors are two np.interp
// look-ups through such polylines; every branch of numpy's interp (exact-hit, last-knot, clamp /
// left / right fill) is reproduced.  All arithmetic is float64 like the reference's.
//
// Work unit: CTA = 8 consecutive cells (one warp each) x a slab of time steps; a lane owns one
// time step at a time, so the 8 warps of a CTA touch the same 32-byte row sectors and each warp's
// binary searches stay inside ONE cell's sorted record (L1-resident: 10 950 x 4 B = 44 KB).
// HBM-bound in principle (read X, write out: 8 B per cell-timestep); no tensor-core path.
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#include "../../include/sdb.h"
#include "common.cuh"

namespace sdb {

constexpr double QMR_PP_MIN = -1e20, QMR_PP_MAX = 1e20;      // quantile.py:17-18

template <typename T>
struct Framed {
    const T* S; int m;
    double den;                 // (m + 1.0 - 0.4) - 0.4
    double pp_lo, v_lo, pp_hi, v_hi;
    __device__ __forceinline__ double pp(int i) const {
        return i == 0 ? pp_lo : (i == m + 1 ? pp_hi : ((double)i - 0.4) / den);
    }
    __device__ __forceinline__ double val(int i) const {
        return i == 0 ? v_lo : (i == m + 1 ? v_hi : (double)S[i - 1]);
    }
};

template <typename T>
__device__ __forceinline__ Framed<T> make_framed(const T* S, int m, const double* frame2, bool lo, bool hi) {
    Framed<T> f;
    f.S = S; f.m = m;
    f.den = (((double)m + 1.0) - 0.4) - 0.4;
    f.pp_lo = lo ? QMR_PP_MIN : (1.0 - 0.4) / f.den;
    f.pp_hi = hi ? QMR_PP_MAX : ((double)m - 0.4) / f.den;
    f.v_lo = frame2[0];
    f.v_hi = frame2[1];
    return f;
}

// np.interp(x, vals, pp, left, right) through the framed CDF: value → plotting position.
// left_inf / right_inf: the fill of quantile.py:243-244 (-inf / +inf) instead of numpy's end-value clamp.
template <typename T>
__device__ double value_to_pp(const Framed<T>& f, double x, bool left_inf, bool right_inf) {
    const int last = f.m + 1;
    if (x != x) return x;
    if (x < f.val(0)) return left_inf ? -INFINITY : f.pp(0);
    if (x > f.val(last)) return right_inf ? INFINITY : f.pp(last);
    // largest j in [0, last] with val(j) <= x
    int lo = 0, hi = last;                       // invariant: val(lo) <= x, answer in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (f.val(mid) <= x) lo = mid; else hi = mid - 1;
    }
    const int j = lo;
    if (j == last) return f.pp(j);
    const double xj = f.val(j);
    if (xj == x) return f.pp(j);
    const double slope = (f.pp(j + 1) - f.pp(j)) / (f.val(j + 1) - xj);
    return slope * (x - xj) + f.pp(j);
}

// np.interp(q, pp, vals) through the framed CDF (default end-value clamp): plotting position → value.
template <typename T>
__device__ double pp_to_value(const Framed<T>& f, double q) {
    const int last = f.m + 1;
    if (q != q) return q;
    if (q < f.pp(0)) return f.val(0);
    if (q > f.pp(last)) return f.val(last);
    // largest j with pp(j) <= q: the interior positions are an arithmetic sequence
    int j;
    if (q < f.pp(1)) j = 0;
    else if (q >= f.pp(last)) j = last;
    else {
        j = (int)floor(q * f.den + 0.4);
        j = j < 1 ? 1 : (j > f.m ? f.m : j);
        while (j > 1 && f.pp(j) > q) --j;
        while (j < f.m && f.pp(j + 1) <= q) ++j;
    }
    if (j == last) return f.val(j);
    const double xj = f.pp(j);
    if (xj == q) return f.val(j);
    const double slope = (f.val(j + 1) - f.val(j)) / (f.pp(j + 1) - xj);
    return slope * (q - xj) + f.val(j);
}

// ---------------------------------------------------------------- frame points of the fitted CDFs
// frame[c * 4 + {0,1,2,3}] = v_lo(X), v_hi(X), v_lo(y), v_hi(y)          quantile.py:349-386
template <typename T>
__global__ void qmr_frame_kernel(const T* __restrict__ sx, const T* __restrict__ sy, int64_t state_ld, int64_t C,
                                 int m, int ne, int lo, int hi, double* __restrict__ frame,
                                 const uint8_t* __restrict__ valid) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double* out = frame + c * 4;
    if (valid && !valid[c]) { out[0] = out[1] = out[2] = out[3] = NAN; return; }
    const double den = (((double)m + 1.0) - 0.4) - 0.4;
    for (int which = 0; which < 2; ++which) {
        const T* S = (which == 0 ? sx : sy) + c * state_ld;
        double v_lo = (double)S[0], v_hi = (double)S[m - 1];
        for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? lo : hi)) continue;
            const int i0 = side == 0 ? 1 : m - ne + 1;          // 1-based index of the first tail point
            double xm = 0.0, ym = 0.0;
            for (int k = 0; k < ne; ++k) { xm += ((double)(i0 + k) - 0.4) / den; ym += (double)S[i0 - 1 + k]; }
            xm /= (double)ne; ym /= (double)ne;
            double sxy = 0.0, sxx = 0.0;
            for (int k = 0; k < ne; ++k) {
                const double dx = ((double)(i0 + k) - 0.4) / den - xm;
                sxy += dx * ((double)S[i0 - 1 + k] - ym);
                sxx += dx * dx;
            }
            const double slope = sxx > 0.0 ? sxy / sxx : 0.0;
            const double icpt = ym - slope * xm;
            const double at = side == 0 ? QMR_PP_MIN : QMR_PP_MAX;
            (side == 0 ? v_lo : v_hi) = slope * at + icpt;
        }
        out[2 * which] = v_lo;
        out[2 * which + 1] = v_hi;
    }
}

// ---------------------------------------------------------------- predict
struct QmrParams {
    const void* X; int64_t ld; int64_t C; int t_pred;
    const void* sx; const void* sy; int64_t state_ld; int m;
    const double* frame;
    const int32_t* rank; int64_t ld_rank;        // EDCDFm: 1-based ordinal rank of every step among its cell's series
    int kind;                                    // SDB_QMR_*
    int lo, hi, one_to_one;
    void* out; int out_f64; int64_t ld_out;
    const uint8_t* valid; int32_t* nonfinite;
};

constexpr int QMR_CT = 8;            // cells per CTA (one warp each)
constexpr int QMR_ROWS = 1024;       // time steps per CTA

template <typename T>
__global__ void __launch_bounds__(QMR_CT * 32)
qmr_predict_kernel(const QmrParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * QMR_CT + warp;
    if (c >= p.C) return;
    const int t0 = blockIdx.y * QMR_ROWS;
    const int t1 = t0 + QMR_ROWS < p.t_pred ? t0 + QMR_ROWS : p.t_pred;
    const bool ok = !p.valid || p.valid[c];
    const T* X = (const T*)p.X;
    Framed<T> fx, fy;
    if (ok) {
        fx = make_framed<T>((const T*)p.sx + c * p.state_ld, p.m, p.frame + c * 4, p.lo != 0, p.hi != 0);
        fy = make_framed<T>((const T*)p.sy + c * p.state_ld, p.m, p.frame + c * 4 + 2, p.lo != 0, p.hi != 0);
    }
    const double pden = (((double)p.t_pred + 1.0) - 0.4) - 0.4;
    for (int t = t0 + lane; t < t1; t += 32) {
        const int64_t at = (int64_t)t * p.ld_out + c;
        double res;
        if (!ok) {
            res = NAN;
        } else {
            const T xr = X[(int64_t)t * p.ld + c];
            const double x = (double)xr;
            if (p.nonfinite && !isfinite(x)) atomicOr(p.nonfinite, 1);
            if (p.kind == SDB_QMR_REGRESSOR) {
                // percentile of x in the fitted X, then the fitted y at that percentile   quantile.py:246-266
                double q = value_to_pp(fx, x, p.lo != 0, p.hi != 0);
                if (isinf(q)) q = NAN;     // beyond the synthetic frame (|x| ~ 1e20 x slope): not reproduced
                res = pp_to_value(fy, q);
            } else {
                // the step's own plotting position, the fitted X and y there, equidistant shift / ratio   quantile.py:609-627
                const int r = p.rank[(int64_t)t * p.ld_rank + c];
                const double q = ((double)r - 0.4) / pden;
                const double x_train = pp_to_value(fx, q);
                const double y_at = pp_to_value(fy, q);
                res = (p.kind == SDB_QMR_EDCDF_DIFFERENCE) ? y_at + (x - x_train) : y_at * (x / x_train);
            }
            res = (double)(T)res;                               // y_hat = np.full_like(X): stored in X's dtype first
            if (p.one_to_one) {                                  // quantile.py:268-309 (X and y fitted on equal lengths)
                const double x_min = (double)fx.S[0], x_max = (double)fx.S[p.m - 1];
                if (x > x_max) res = (double)fy.S[p.m - 1] + (x - x_max);
                if (x < x_min) res = (double)fy.S[0] + (x - x_min);
            }
        }
        if (p.out_f64) ((double*)p.out)[at] = res; else ((T*)p.out)[at] = (T)res;
    }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_qmr_frame(const void* sorted_x, const void* sorted_y, int dtype, int64_t state_ld,
                             int64_t n_cells, int n_fit, int extrapolate, int n_endpoints,
                             double* frame, const uint8_t* cell_valid, void* stream) {
    if (!sorted_x || !sorted_y || !frame) return sdb_fail(SDB_E_INVALID, "sdb_qmr_frame: NULL pointer");
    if (n_cells <= 0 || n_fit <= 0 || state_ld < n_fit) return sdb_fail(SDB_E_INVALID, "sdb_qmr_frame: bad shape");
    if (extrapolate < SDB_EXTRAPOLATE_NONE || extrapolate > SDB_EXTRAPOLATE_BOTH)
        return sdb_fail(SDB_E_INVALID, "sdb_qmr_frame: unknown extrapolate code %d", extrapolate);
    if (n_endpoints < 2 || n_fit < 2 * n_endpoints + 1)
        return sdb_fail(SDB_E_INVALID, "sdb_qmr_frame: need n_endpoints >= 2 and n_fit >= 2 * n_endpoints + 1");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n_cells + 127) / 128);
    const int lo = (extrapolate & SDB_EXTRAPOLATE_MIN) != 0, hi = (extrapolate & SDB_EXTRAPOLATE_MAX) != 0;
    if (dtype == SDB_F32)
        qmr_frame_kernel<float><<<grid, 128, 0, st>>>((const float*)sorted_x, (const float*)sorted_y, state_ld, n_cells, n_fit, n_endpoints, lo, hi, frame, cell_valid);
    else if (dtype == SDB_F64)
        qmr_frame_kernel<double><<<grid, 128, 0, st>>>((const double*)sorted_x, (const double*)sorted_y, state_ld, n_cells, n_fit, n_endpoints, lo, hi, frame, cell_valid);
    else return sdb_fail(SDB_E_INVALID, "sdb_qmr_frame: bad dtype %d", dtype);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_qmr_predict(int kind, const void* X, int dtype, int64_t ld, int64_t n_cells, int t_pred,
                               const void* sorted_x, const void* sorted_y, int64_t state_ld, int n_fit,
                               const double* frame, int extrapolate, int one_to_one,
                               const int32_t* rank, int64_t ld_rank,
                               void* out, int out_dtype, int64_t ld_out,
                               const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X || !sorted_x || !sorted_y || !frame || !out) return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: NULL pointer");
    if (n_cells <= 0 || t_pred <= 0 || n_fit <= 0 || ld < n_cells || ld_out < n_cells || state_ld < n_fit)
        return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: bad shape");
    if (kind < SDB_QMR_REGRESSOR || kind > SDB_QMR_EDCDF_RATIO) return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: unknown kind %d", kind);
    if (kind != SDB_QMR_REGRESSOR && (!rank || ld_rank < n_cells)) return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: EDCDFm needs the rank array");
    if (extrapolate < SDB_EXTRAPOLATE_NONE || extrapolate > SDB_EXTRAPOLATE_BOTH)
        return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: unknown extrapolate code %d", extrapolate);
    if ((dtype != SDB_F32 && dtype != SDB_F64) || (out_dtype != SDB_F32 && out_dtype != SDB_F64))
        return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: bad dtype");
    if (out_dtype != dtype) return sdb_fail(SDB_E_INVALID, "sdb_qmr_predict: the result has the dtype of X (np.full_like, quantile.py:265)");
    QmrParams p;
    p.X = X; p.ld = ld; p.C = n_cells; p.t_pred = t_pred; p.sx = sorted_x; p.sy = sorted_y; p.state_ld = state_ld; p.m = n_fit;
    p.frame = frame; p.rank = rank; p.ld_rank = ld_rank; p.kind = kind;
    p.lo = (extrapolate & SDB_EXTRAPOLATE_MIN) != 0; p.hi = (extrapolate & SDB_EXTRAPOLATE_MAX) != 0; p.one_to_one = one_to_one;
    p.out = out; p.out_f64 = (out_dtype == SDB_F64); p.ld_out = ld_out; p.valid = cell_valid; p.nonfinite = nonfinite;
    dim3 grid((unsigned)((n_cells + QMR_CT - 1) / QMR_CT), (unsigned)((t_pred + QMR_ROWS - 1) / QMR_ROWS));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SDB_F32) qmr_predict_kernel<float><<<grid, QMR_CT * 32, 0, st>>>(p);
    else                  qmr_predict_kernel<double><<<grid, QMR_CT * 32, 0, st>>>(p);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

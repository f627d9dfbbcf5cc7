// qm_kernels.cu — empirical-CDF quantile mapping for every (cell, time-group), sm_100a.
//
// Replaces the reference's per-cell Python loop
//   skdownscale/pointwise_models/core.py:69-143
// around
//   QuantileMapper / CunnaneTransformer   skdownscale/pointwise_models/quantile.py:81-147, 438-545
//   BcsdTemperature / BcsdPrecipitation    skdownscale/pointwise_models/bcsd.py:115-185, 197-281
//
// Data layout: inputs [T, C] time-major / cell-fastest (the reference's (time, lat, lon)
// C-order, zero copy); fitted state [C, state_ld] cell-major (each cell's sorted groups
// contiguous, private to this library).  Work unit = (cell, group): NT threads hold the
// group's series in registers (E per thread, blocked) and sort it with the network of
// sort.cuh.  No tensor cores: there is no contraction on this path.
//
// Compiled with -fmad=false: the float64 interpolation must round like numpy's
// `slope*(x - xp[j]) + fp[j]` (separate multiply and add).
//
// This header holds the kernels; qm_np<N>.cu instantiate them for one padded group size
// each (parallel compilation), qm_api.cu holds the extern "C" entry points.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>

#include "../../include/sdb.h"
#include "sort.cuh"
#include "common.cuh"

namespace sdb {

// ---------------------------------------------------------------- item construction
template <typename T> struct FitItemOf;
template <> struct FitItemOf<float>  { using type = K32; };
template <> struct FitItemOf<double> { using type = K64; };

__device__ __forceinline__ K32 make_fit_item(float x)  { K32 r; r.k = f32_to_sortable(x + 0.0f); return r; }
__device__ __forceinline__ K64 make_fit_item(double x) {
    uint64_t s = f64_to_sortable(x + 0.0);
    K64 r; r.hi = (uint32_t)(s >> 32); r.lo = (uint32_t)s; return r;
}
__device__ __forceinline__ float  fit_item_value(const K32& v) { return sortable_to_f32(v.k); }
__device__ __forceinline__ double fit_item_value(const K64& v) { return sortable_to_f64(((uint64_t)v.hi << 32) | v.lo); }
template <class I> __device__ __forceinline__ I sentinel_item(uint32_t pos);
template <> __device__ __forceinline__ K32  sentinel_item<K32>(uint32_t)    { K32 r;  r.k = 0xffffffffu; return r; }
template <> __device__ __forceinline__ K64  sentinel_item<K64>(uint32_t)    { K64 r;  r.hi = r.lo = 0xffffffffu; return r; }
template <> __device__ __forceinline__ K32I sentinel_item<K32I>(uint32_t p) { K32I r; r.k = 0xffffffffu; r.i = p; return r; }
template <> __device__ __forceinline__ K64I sentinel_item<K64I>(uint32_t p) { K64I r; r.hi = r.lo = 0xffffffffu; r.i = p; return r; }

__device__ __forceinline__ K32I make_rank_item32(float key, uint32_t pos) {
    K32I r; r.k = f32_to_sortable(key + 0.0f); r.i = pos; return r;
}
__device__ __forceinline__ K64I make_rank_item64(double key, uint32_t pos) {
    uint64_t s = f64_to_sortable(key + 0.0);
    K64I r; r.hi = (uint32_t)(s >> 32); r.lo = (uint32_t)s; r.i = pos; return r;
}

template <typename T>
__device__ __forceinline__ void flag_nonfinite(T x, int32_t* flag) {
    if (flag && !isfinite(x)) atomicOr(flag, 1);
}

template <int NT> struct Cfg {
    static constexpr int THREADS = (NT >= 256) ? NT : 256;   // CTA size
    static constexpr int CPB = (NT > 32) ? 1 : THREADS / NT; // cells per CTA (multi-warp groups: one cell)
    static constexpr int BLOCK = CPB * NT;
};

// ---------------------------------------------------------------- fit: sort every group
template <typename T, int E, int NT>
__global__ void __launch_bounds__(Cfg<NT>::BLOCK)
qm_fit_kernel(const T* __restrict__ y, int64_t ld, int64_t C,
              const int32_t* __restrict__ rows, const int32_t* __restrict__ len,
              const int64_t* __restrict__ off, int max_len,
              T* __restrict__ state, int64_t state_ld, const uint8_t* __restrict__ valid,
              int32_t* __restrict__ nonfinite) {
    extern __shared__ uint32_t smem[];
    using Item = typename FitItemOf<T>::type;
    const int sub = threadIdx.x / NT, tid = threadIdx.x % NT;
    const int64_t c = (int64_t)blockIdx.x * Cfg<NT>::CPB + sub;
    const int g = blockIdx.y;
    if (c >= C || (valid && !valid[c])) return;      // uniform per sorting group (and per CTA when NT > 32)
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    Item v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = tid * E + e;
        if (j < n) {
            const T x = y[(int64_t)rg[j] * ld + c];
            flag_nonfinite(x, nonfinite);
            v[e] = make_fit_item(x);
        } else {
            v[e] = sentinel_item<Item>(j);
        }
    }
    sort_blocked<Item, E, NT>(v, tid, smem);
    T* dst = state + c * state_ld + off[g];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = tid * E + e;
        if (j < n) dst[j] = fit_item_value(v[e]);
    }
}

// ---------------------------------------------------------------- predict
struct PredictParams {
    const void* X; int64_t ld; int64_t C;
    const int32_t* rows; const int32_t* len; const int32_t* state_gid; int max_len;
    const int32_t* fit_len; const int64_t* state_off;
    const void* state; int64_t state_ld;
    const void* x_climo; const void* y_climo; int64_t ld_climo;
    int return_anoms; const int32_t* roll_nbr;
    void* out; int64_t ld_out; int32_t* rank_out;
    const uint8_t* valid; int32_t* nonfinite;
    int mode;        // SDB_MODE_*
    int out_f64;     // output element type: 0 float, 1 double
    int n_groups;
    int no_vec;      // testing: tile kernels use the 4-byte (unaligned-safe) row loads and stores
    // CunnaneTransformer settings (quantile.py:420-432): plotting-position parameters, number of tail
    // points of the OLS extrapolation, and which tails extrapolate (otherwise np.interp clamps)
    double alpha, beta;
    int n_endpoints, extrap_lo, extrap_hi;
    // sdb_series_rank: stop after the ranking pass and write rank_out only; rank_ordinal = 1-based
    // position in the (value, time index) order instead of the tie-max rank
    int rank_only, rank_ordinal;
};

// kernel flavours (compile-time): what the rank keys are
constexpr int KIND_RAW = 0;        // QuantileMapper / BcsdPrecipitation: keys are the inputs themselves
constexpr int KIND_SHIFT = 1;      // BcsdTemperature, 9-sample window inside the mapping group
constexpr int KIND_SHIFT_TAB = 2;  // BcsdTemperature, window members given by a neighbour table

__device__ __forceinline__ void store_out(void* out, int out_f64, int64_t at, double v) {
    if (out_f64) ((double*)out)[at] = v; else ((float*)out)[at] = (float)v;
}

// Cunnane plotting position pp_len(i), i 1-based, exactly as numpy evaluates
// (np.arange(1, n+1) - 0.4) / (n + 1.0 - 0.4 - 0.4)        quantile.py:43
struct Cunnane { double alpha, beta; int ne; bool lo, hi; };
__device__ __forceinline__ Cunnane cunnane_of(const PredictParams& p) {
    return Cunnane{p.alpha, p.beta, p.n_endpoints, p.extrap_lo != 0, p.extrap_hi != 0};
}
__device__ __forceinline__ double pp_denominator(int len, const Cunnane& cu) { return (((double)len + 1.0) - cu.alpha) - cu.beta; }
__device__ __forceinline__ double pp_of(int i, double den, const Cunnane& cu) { return ((double)i - cu.alpha) / den; }

// OLS line through `ne` (pp, value) points starting at 1-based index i0 (quantile.py:532-543;
// sklearn LinearRegression = centred least squares; slope 0 when the abscissae coincide,
// which is the minimum-norm lstsq answer for a single point).  S(i): 0-based sorted value.
template <class Acc>
__device__ void ols_tail(const Acc& S, int i0, int ne, double den, const Cunnane& cu, double& slope, double& icpt) {
    double xm = 0.0, ym = 0.0;
    for (int k = 0; k < ne; ++k) { xm += pp_of(i0 + k, den, cu); ym += S(i0 - 1 + k); }
    xm /= (double)ne; ym /= (double)ne;
    double sxy = 0.0, sxx = 0.0;
    for (int k = 0; k < ne; ++k) {
        double dx = pp_of(i0 + k, den, cu) - xm;
        sxy += dx * (S(i0 - 1 + k) - ym);
        sxx += dx * dx;
    }
    slope = (sxx > 0.0) ? sxy / sxx : 0.0;
    icpt = ym - slope * xm;
}

// inverse CDF: value of the fitted sorted series S (length m, accessor returning double) at the
// quantile of rank r of n (CunnaneTransformer.inverse_transform, quantile.py:523-545 =
// np.interp + OLS tails)
template <class Acc>
__device__ double inverse_cdf_acc(int r, int n, int m, const Acc& S, double dn, double dm, const Cunnane& cu) {
    if (n == m) return S(r - 1);                     // q lands exactly on knot r: np.interp returns fp[r-1]
    const double q = pp_of(r, dn, cu);
    const double p1 = pp_of(1, dm, cu), pm = pp_of(m, dm, cu);
    const int ne = m < cu.ne ? m : cu.ne;
    if (q < p1) {                                    // left=-inf + OLS tail, or np.interp's default clamp
        if (!cu.lo) return S(0);
        double a, b; ols_tail(S, 1, ne, dm, cu, a, b); return a * q + b;
    }
    if (q > pm) {
        if (!cu.hi) return S(m - 1);
        double a, b; ols_tail(S, m - ne + 1, ne, dm, cu, a, b); return a * q + b;
    }
    if (q == pm || m == 1) return S(m - 1);
    int j = (int)floor(q * dm + cu.alpha);
    j = j < 1 ? 1 : (j > m - 1 ? m - 1 : j);
    while (j > 1 && pp_of(j, dm, cu) > q) --j;
    while (j < m - 1 && pp_of(j + 1, dm, cu) <= q) ++j;
    const double xj = pp_of(j, dm, cu);
    if (q == xj) return S(j - 1);
    const double yj = S(j - 1), yj1 = S(j);
    const double slope = (yj1 - yj) / (pp_of(j + 1, dm, cu) - xj);
    return slope * (q - xj) + yj;
}

template <typename T>
__device__ double inverse_cdf(int r, int n, int m, const T* __restrict__ S, double dn, double dm, const Cunnane& cu) {
    auto acc = [&](int i) -> double { return (double)S[i]; };
    return inverse_cdf_acc(r, n, m, acc, dn, dm, cu);
}

// (x_j, shift_j) of member j of the group: shift = centred 9-sample mean of the climate-trend
// group minus the x climatology (bcsd.py:247-256), float64 like pandas' rolling mean.
template <typename T, bool ROLLTAB>
__device__ __forceinline__ void shifted_value(const T* __restrict__ X, int64_t ld, int64_t c,
                                              const int32_t* __restrict__ rg, int n, int j,
                                              const int32_t* __restrict__ roll_nbr, double xc,
                                              double& x, double& shift) {
    double acc = 0.0; int cnt = 0;
    if (ROLLTAB) {
        const int32_t row = rg[j];
        const int32_t* nb = roll_nbr + (int64_t)row * 9;
        x = (double)X[(int64_t)row * ld + c];
#pragma unroll
        for (int d = 0; d < 9; ++d) {
            const int32_t r = nb[d];
            if (r >= 0) { acc += (double)X[(int64_t)r * ld + c]; ++cnt; }
        }
    } else {
        x = 0.0;
#pragma unroll
        for (int d = -4; d <= 4; ++d) {
            const int jj = j + d;
            if (jj >= 0 && jj < n) {
                const double xv = (double)X[(int64_t)rg[jj] * ld + c];
                acc += xv; ++cnt;
                if (d == 0) x = xv;
            }
        }
    }
    shift = acc / (double)cnt - xc;
}

template <typename T, int E, int NT, int KIND>
__global__ void __launch_bounds__(Cfg<NT>::BLOCK)
qm_predict_kernel(const PredictParams p) {
    extern __shared__ uint32_t smem[];
    constexpr bool SHIFT = (KIND != KIND_RAW);
    constexpr bool ROLLTAB = (KIND == KIND_SHIFT_TAB);
    constexpr bool KEY64 = SHIFT || (sizeof(T) == 8);
    using Item = typename std::conditional<KEY64, K64I, K32I>::type;
    constexpr int NP = E * NT;
    const int sub = threadIdx.x / NT, tid = threadIdx.x % NT;
    const int64_t c = (int64_t)blockIdx.x * Cfg<NT>::CPB + sub;
    const int g = blockIdx.y;
    if (c >= p.C) return;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    if (p.valid && !p.valid[c]) {
        for (int j = tid; j < n; j += NT) {
            if (p.out) store_out(p.out, p.out_f64, (int64_t)rg[j] * p.ld_out + c, (double)NAN);
            if (p.rank_out) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = 0;
        }
        return;
    }
    const T* X = (const T*)p.X;
    const int sg = p.rank_only ? 0 : p.state_gid[g];
    const int m = p.rank_only ? n : p.fit_len[sg];
    const T* S = p.rank_only ? nullptr : (const T*)p.state + c * p.state_ld + p.state_off[sg];
    double xc = 0.0, yc = 0.0;
    if (SHIFT) xc = (double)((const T*)p.x_climo)[(int64_t)sg * p.ld_climo + c];
    if (p.mode != SDB_MODE_QM && p.return_anoms) yc = (double)((const T*)p.y_climo)[(int64_t)sg * p.ld_climo + c];

    // ---- pass 1: keys → sort → tie-max ranks → rank_of[position in group]
    Item v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = tid * E + e;
        if (j < n) {
            if constexpr (SHIFT) {
                double x, s;
                shifted_value<T, ROLLTAB>(X, p.ld, c, rg, n, j, p.roll_nbr, xc, x, s);
                flag_nonfinite(x, p.nonfinite);
                v[e] = make_rank_item64(x - s, (uint32_t)j);
            } else {
                const T x = X[(int64_t)rg[j] * p.ld + c];
                flag_nonfinite(x, p.nonfinite);
                if constexpr (KEY64) v[e] = make_rank_item64((double)x, (uint32_t)j);
                else                 v[e] = make_rank_item32((float)x, (uint32_t)j);
            }
        } else {
            v[e] = sentinel_item<Item>((uint32_t)j);
        }
    }
    sort_blocked<Item, E, NT>(v, tid, smem);
    int r[E];
    if (p.rank_ordinal) {
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = tid * E + e + 1;
    } else {
        tie_max_ranks<Item, E, NT>(v, tid, r, smem);
    }
    if (NT > 32) __syncthreads();                       // scratch is about to be reused as rank_of
    uint16_t* rank_of = reinterpret_cast<uint16_t*>(smem) + (size_t)sub * NP;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (v[e].i < (uint32_t)n) rank_of[v[e].i] = (uint16_t)r[e];
    if (NT > 32) __syncthreads(); else __syncwarp();

    if (p.rank_only) {
        for (int j = tid; j < n; j += NT) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = (int)rank_of[j];
        return;
    }
    // ---- pass 2: rank → quantile → inverse CDF of the fitted group → output
    const Cunnane cu = cunnane_of(p);
    const double dn = pp_denominator(n, cu), dm = pp_denominator(m, cu);
#pragma unroll 4
    for (int e = 0; e < E; ++e) {
        const int j = tid * E + e;
        if (j >= n) break;
        const int rk = (int)rank_of[j];
        const double val = inverse_cdf<T>(rk, n, m, S, dn, dm, cu);
        double o;
        if constexpr (SHIFT) {
            double x, s;
            shifted_value<T, ROLLTAB>(X, p.ld, c, rg, n, j, p.roll_nbr, xc, x, s);
            o = s + val;                                   // bcsd.py:263
            if (p.return_anoms) o = o - yc;                // bcsd.py:267
        } else if (p.mode == SDB_MODE_BCSD_P) {
            o = p.return_anoms ? val / yc : val;           // bcsd.py:170-185
        } else {
            o = val;
        }
        const int64_t at = (int64_t)rg[j] * p.ld_out + c;
        store_out(p.out, p.out_f64, at, o);
        if (p.rank_out) p.rank_out[at] = rk;
    }
}

// ---------------------------------------------------------------- launchers (one padded size per TU)
struct FitParams {
    const void* y; int64_t ld; int64_t C;
    const int32_t* rows; const int32_t* len; const int64_t* off; int n_groups; int max_len;
    void* state; int64_t state_ld; const uint8_t* valid; int32_t* nonfinite;
    int no_vec;      // testing: see PredictParams
};

template <typename T, int E, int NT>
static int launch_fit(const FitParams& f, cudaStream_t st) {
    using Item = typename FitItemOf<T>::type;
    const size_t smem = (NT > 32) ? (size_t)E * NT * item_words<Item>::value * 4 : 0;
    auto kern = qm_fit_kernel<T, E, NT>;
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((f.C + Cfg<NT>::CPB - 1) / Cfg<NT>::CPB), (unsigned)f.n_groups);
    kern<<<grid, Cfg<NT>::BLOCK, smem, st>>>((const T*)f.y, f.ld, f.C, f.rows, f.len, f.off, f.max_len, (T*)f.state, f.state_ld, f.valid, f.nonfinite);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int E, int NT, int KIND>
static int launch_predict(const PredictParams& p, cudaStream_t st) {
    constexpr bool KEY64 = (KIND != KIND_RAW) || (sizeof(T) == 8);
    using Item = typename std::conditional<KEY64, K64I, K32I>::type;
    const size_t xchg = (NT > 32) ? (size_t)E * NT * item_words<Item>::value * 4 : 0;
    const size_t ranks = (size_t)Cfg<NT>::CPB * E * NT * 2;
    const size_t smem = xchg > ranks ? xchg : ranks;
    auto kern = qm_predict_kernel<T, E, NT, KIND>;
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.C + Cfg<NT>::CPB - 1) / Cfg<NT>::CPB), (unsigned)p.n_groups);
    kern<<<grid, Cfg<NT>::BLOCK, smem, st>>>(p);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

// fast tile kernels (qm_tile.cuh), float32 only, instantiated in qm_np256.cu / qm_np1024.cu
int qm_fit_tile_np256(const FitParams& f, cudaStream_t st);
int qm_fit_tile_np1024(const FitParams& f, cudaStream_t st);
int qm_predict_tile_np256(int kind, const PredictParams& p, cudaStream_t st);
int qm_predict_tile_np1024(int kind, const PredictParams& p, cudaStream_t st);

// long float32 groups by counting rank (qm_long.cu)
int qm_fit_long(const FitParams& f, cudaStream_t st);
int qm_predict_long(const PredictParams& p, cudaStream_t st);

// per-size entry points, defined in qm_np<N>.cu
#define SDB_DECLARE_SIZE(NP)                                                             \
    int qm_fit_np##NP(int dtype, const FitParams& f, cudaStream_t st);                   \
    int qm_predict_np##NP(int dtype, int kind, const PredictParams& p, cudaStream_t st);
SDB_DECLARE_SIZE(256)
SDB_DECLARE_SIZE(1024)
SDB_DECLARE_SIZE(4096)
SDB_DECLARE_SIZE(16384)

#define SDB_DEFINE_SIZE(NP, E, NT)                                                                    \
    int qm_fit_np##NP(int dtype, const FitParams& f, cudaStream_t st) {                               \
        return dtype == SDB_F32 ? launch_fit<float, E, NT>(f, st) : launch_fit<double, E, NT>(f, st); \
    }                                                                                                 \
    int qm_predict_np##NP(int dtype, int kind, const PredictParams& p, cudaStream_t st) {             \
        if (dtype == SDB_F32) {                                                                       \
            if (kind == KIND_RAW) return launch_predict<float, E, NT, KIND_RAW>(p, st);               \
            if (kind == KIND_SHIFT) return launch_predict<float, E, NT, KIND_SHIFT>(p, st);           \
            return launch_predict<float, E, NT, KIND_SHIFT_TAB>(p, st);                               \
        }                                                                                             \
        if (kind == KIND_RAW) return launch_predict<double, E, NT, KIND_RAW>(p, st);                  \
        if (kind == KIND_SHIFT) return launch_predict<double, E, NT, KIND_SHIFT>(p, st);              \
        return launch_predict<double, E, NT, KIND_SHIFT_TAB>(p, st);                                  \
    }

}  // namespace sdb

// analog_kernels.cu — GARD analog downscaling for every cell, sm_100a.
//
// Replaces, per cell and per query timestep,
//   AnalogBase.fit (KDTree build)                  skdownscale/pointwise_models/gard.py:58-87
//   PureAnalog.predict                              gard.py:273-364
//   AnalogRegression.predict / _predict_one_step    gard.py:152-224 (with and without thresh)
// and the Python loop over cells around them (core.py:69-143).
//
// The k-nearest-neighbour search is an EXACT float64 brute force: squared Euclidean
// distance accumulated feature by feature with separate multiply and add (-fmad=false),
// which reproduces sklearn KDTree's neighbour indices bit for bit on tie-free data
// (lowest train index first on exact ties).  For float32 inputs a float32 distance (FMA,
// one 16-byte shared-memory read per training point) screens the candidates first: its
// relative error (< 1e-6) is covered by the slack of the bound, so every point whose exact
// distance could enter the list is still evaluated exactly — results are unchanged.  One thread owns one query timestep and keeps
// its running top-k; the CTA streams the cell's training window through shared memory
// (float64, broadcast reads).  The work is FP64-pipe bound — there is no GEMM here (K = p
// is 1..8 and the `|a|^2 - 2ab + |b|^2` trick would break index exactness), so no tensor
// cores.
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#include "../../include/sdb.h"
#include "common.cuh"
#include "np_pairwise.cuh"
#include "analog_grid.cuh"

namespace sdb {

constexpr int AN_THREADS = 256;     // queries per CTA
template <int P> struct AnChunk { static constexpr int value = (P <= 3) ? 512 : (P == 4 ? 384 : 192); };    // training points staged per pass
constexpr int AN_QD = 8;            // per-lane queue of accepted candidates (register-list kernels)
template <int P> struct AnFStride { static constexpr int value = (P <= 4) ? 4 : 8; };                   // float32 copy: 16-byte rows
constexpr int AN_KREG = 16;         // top-k list kept in registers up to this k
constexpr int AN_PMAX = 8;          // generic-feature kernel handles up to this many predictors

struct AnalogParams {
    const void* Xtr; const void* ytr; const void* Xq;
    int64_t ld; int64_t C; int t_fit, t_query, p, k, kind;
    int has_thresh; double thresh; double logistic_c; const int32_t* rand_idx;
    void* out; int out_f64; int64_t ld_out; int32_t* knn_idx;
    const uint8_t* valid; int32_t* nonfinite;
    // pruned search (analog_cell_kernel): the quantile grid of every cell's training window (analog_grid.cu)
    const int32_t* perm_t; const int32_t* perm_q; const int32_t* box_start; const float* bounds; int64_t ld_grid;
};

__device__ __forceinline__ void store3(const AnalogParams& a, int q, int64_t c, double pred, double prob, double err) {
    const int64_t base = (int64_t)q * 3 * a.ld_out + c;
    if (a.out_f64) {
        double* o = (double*)a.out;
        o[base] = pred; o[base + a.ld_out] = prob; o[base + 2 * a.ld_out] = err;
    } else {
        float* o = (float*)a.out;
        o[base] = (float)pred; o[base + a.ld_out] = (float)prob; o[base + 2 * a.ld_out] = (float)err;
    }
}

// ---- epilogues.  idx(i)/dist2(i): i-th nearest training row and its squared distance.
template <typename T, typename IdxF, typename D2F>
__device__ void pure_analog_epilogue(const AnalogParams& a, int q, int64_t c, int k, const IdxF& idx, const D2F& dist2) {
    const T* y = (const T*)a.ytr;
    auto yv = [&](int i) -> T { return y[(int64_t)idx(i) * a.ld + c]; };
    const T th = (T)a.thresh;
    int n_exceed = 0;
    if (a.has_thresh) for (int i = 0; i < k; ++i) n_exceed += (yv(i) > th) ? 1 : 0;
    const bool any_masked = a.has_thresh && (n_exceed < k);
    double pred;
    if (a.kind == SDB_ANALOG_BEST) {
        pred = (double)yv(0);                                               // gard.py:310-311
    } else if (a.kind == SDB_ANALOG_SAMPLE) {
        pred = (double)yv(a.rand_idx[(int64_t)q * a.C + c]);               // gard.py:313-317
    } else if (a.kind == SDB_ANALOG_WEIGHT) {                               // gard.py:319-327
        if (any_masked) pred = NAN;
        else {
            auto w = [&](int i) -> double { double d = sqrt(dist2(i)); return 1.0 / (d == 0.0 ? 1e-20 : d); };
            auto wy = [&](int i) -> double { return (double)yv(i) * w(i); };
            const double scl = np_sum<double>(w, k);
            pred = np_sum<double>(wy, k) / scl;
        }
    } else {                                                                // mean_analogs  gard.py:329-333
        if (any_masked) pred = NAN;
        else pred = (double)(np_sum<T>(yv, k) / (T)k);
    }
    double err, prob;
    if (a.has_thresh) {                                                     // gard.py:338-343
        if (pred != pred) pred = 0.0;                                       // nan_to_num
        prob = (double)n_exceed / (double)k;
    } else {
        prob = 1.0;
    }
    if (any_masked) err = NAN;
    else {                                                                  // ndarray.std, ddof=0, in y's dtype
        const T mean = np_sum<T>(yv, k) / (T)k;
        auto sq = [&](int i) -> T { T d = yv(i) - mean; return d * d; };
        err = (double)(T)sqrt(np_sum<T>(sq, k) / (T)k);
    }
    store3(a, q, c, pred, prob, err);
}

// Minimum-norm solution of the symmetric positive semi-definite p x p system A beta = b
// (A given by its lower triangle) through a cyclic Jacobi eigen-decomposition: what
// scipy.linalg.lstsq (gelsd) under sklearn's LinearRegression returns when the centred analog
// cloud is rank deficient (fewer regression samples than predictors + 1).
template <int P>
__device__ void sym_minnorm_solve(const double (&A)[P][P], const double (&b)[P], int p, double (&beta)[P]) {
    double M[P][P], V[P][P];
#pragma unroll
    for (int f = 0; f < P; ++f)
#pragma unroll
        for (int g = 0; g < P; ++g) {
            M[f][g] = (f < p && g < p) ? (g <= f ? A[f][g] : A[g][f]) : 0.0;
            V[f][g] = (f == g) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int f = 0; f < P; ++f) {
            diag += M[f][f] * M[f][f];
#pragma unroll
            for (int g = 0; g < f; ++g) off += M[f][g] * M[f][g];
        }
        if (!(off > 1e-60 * diag) || !(off > 0.0)) break;
#pragma unroll
        for (int f = 1; f < P; ++f) {
#pragma unroll
            for (int g = 0; g < f; ++g) {
                const double apq = M[f][g];
                if (apq == 0.0) continue;
                const double theta = (M[g][g] - M[f][f]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                // rotate rows / columns f and g
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const double mhf = M[h][f], mhg = M[h][g];
                    M[h][f] = cs * mhf - sn * mhg;
                    M[h][g] = sn * mhf + cs * mhg;
                }
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const double mfh = M[f][h], mgh = M[g][h];
                    M[f][h] = cs * mfh - sn * mgh;
                    M[g][h] = sn * mfh + cs * mgh;
                }
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const double vhf = V[h][f], vhg = V[h][g];
                    V[h][f] = cs * vhf - sn * vhg;
                    V[h][g] = sn * vhf + cs * vhg;
                }
            }
        }
    }
    double lmax = 0.0;
#pragma unroll
    for (int f = 0; f < P; ++f) lmax = fmax(lmax, M[f][f]);
#pragma unroll
    for (int f = 0; f < P; ++f) beta[f] = 0.0;
#pragma unroll
    for (int e = 0; e < P; ++e) {
        const double lam = M[e][e];
        if (e < p && lam > 1e-13 * lmax && lam > 0.0) {
            double vb = 0.0;
#pragma unroll
            for (int f = 0; f < P; ++f) vb += V[f][e] * b[f];
            vb /= lam;
#pragma unroll
            for (int f = 0; f < P; ++f) beta[f] += V[f][e] * vb;
        }
    }
}

// L2-regularised logistic regression of the exceedance flags on the k analog predictors, solved to
// machine precision by damped Newton (sklearn's LogisticRegression(C) minimises the same objective
// with lbfgs stopped at tol = 1e-4, so it agrees with this optimum to ~1e-4): returns
// predict_proba(X)[0, 0] = P(class 0) at the query point.  gard.py:205-212
template <int P, typename XF, typename EF>
__device__ void logistic_fit(int k, int p, const XF& xv, const EF& exceeds, double Creg, double (&th)[P + 1]) {
    constexpr int D = P + 1;
#pragma unroll
    for (int f = 0; f < D; ++f) th[f] = 0.0;
    const double lam = 1.0 / Creg;
    auto objective = [&](const double (&v)[D]) -> double {
        double s = 0.0;
        for (int i = 0; i < k; ++i) {
            double z = v[P];
#pragma unroll
            for (int f = 0; f < P; ++f) if (f < p) z += v[f] * xv(i, f);
            s += fmax(z, 0.0) + log1p(exp(-fabs(z))) - (exceeds(i) ? z : 0.0);
        }
        double r = 0.0;
#pragma unroll
        for (int f = 0; f < P; ++f) if (f < p) r += v[f] * v[f];
        return s + 0.5 * lam * r;
    };
    double fcur = objective(th);
    for (int it = 0; it < 100; ++it) {
        double g[D], H[D][D];
#pragma unroll
        for (int f = 0; f < D; ++f) { g[f] = 0.0;
#pragma unroll
            for (int h = 0; h < D; ++h) H[f][h] = 0.0; }
        for (int i = 0; i < k; ++i) {
            double xi[D];
#pragma unroll
            for (int f = 0; f < P; ++f) xi[f] = (f < p) ? xv(i, f) : 0.0;
            xi[P] = 1.0;
            double z = 0.0;
#pragma unroll
            for (int f = 0; f < D; ++f) z += th[f] * xi[f];
            const double mu = 1.0 / (1.0 + exp(-z));
            const double r = mu - (exceeds(i) ? 1.0 : 0.0), w = mu * (1.0 - mu);
#pragma unroll
            for (int f = 0; f < D; ++f) { g[f] += r * xi[f];
#pragma unroll
                for (int h = 0; h <= f; ++h) H[f][h] += w * xi[f] * xi[h]; }
        }
#pragma unroll
        for (int f = 0; f < P; ++f) {
            if (f < p) { g[f] += lam * th[f]; H[f][f] += lam; }
            else H[f][f] = 1.0;                              // unused predictor slots: identity, zero step
        }
        // Cholesky of H (positive definite: ridge on the weights, both classes present for the intercept)
        bool ok = true;
#pragma unroll
        for (int f = 0; f < D; ++f) {
#pragma unroll
            for (int h = 0; h <= f; ++h) {
                double sacc = H[f][h];
#pragma unroll
                for (int u = 0; u < h; ++u) sacc -= H[f][u] * H[h][u];
                if (h == f) { if (!(sacc > 0.0)) { ok = false; sacc = 1.0; } H[f][f] = sqrt(sacc); }
                else H[f][h] = sacc / H[h][h];
            }
        }
        double d[D];
#pragma unroll
        for (int f = 0; f < D; ++f) {
            double sacc = -g[f];
#pragma unroll
            for (int u = 0; u < f; ++u) sacc -= H[f][u] * d[u];
            d[f] = sacc / H[f][f];
        }
#pragma unroll
        for (int f = D - 1; f >= 0; --f) {
            double sacc = d[f];
#pragma unroll
            for (int u = f + 1; u < D; ++u) sacc -= H[u][f] * d[u];
            d[f] = sacc / H[f][f];
        }
        if (!ok) {                                           // numerically flat curvature: plain gradient step
#pragma unroll
            for (int f = 0; f < D; ++f) d[f] = -g[f];
        }
        double gd = 0.0;
#pragma unroll
        for (int f = 0; f < D; ++f) gd += g[f] * d[f];
        double step = 1.0, fnew = fcur;
        double cand[D];
        for (;;) {                                           // Armijo backtracking
#pragma unroll
            for (int f = 0; f < D; ++f) cand[f] = th[f] + step * d[f];
            fnew = objective(cand);
            if (fnew <= fcur + 1e-4 * step * gd || step < 1e-10) break;
            step *= 0.5;
        }
        double dmax = 0.0, tmax = 1.0;
#pragma unroll
        for (int f = 0; f < D; ++f) { dmax = fmax(dmax, fabs(step * d[f])); th[f] = cand[f]; tmax = fmax(tmax, fabs(cand[f])); }
        fcur = fnew;
        if (dmax < 1e-14 * tmax) break;
    }
}

template <int P, typename XF, typename EF>
__device__ double logistic_prob_class0(int k, int p, const XF& xv, const EF& exceeds, const double (&xq)[P], double Creg) {
    double th[P + 1];
    logistic_fit<P>(k, p, xv, exceeds, Creg, th);
    double z = th[P];
#pragma unroll
    for (int f = 0; f < P; ++f) if (f < p) z += th[f] * xq[f];
    return 1.0 - 1.0 / (1.0 + exp(-z));
}

// Ordinary least squares with intercept on the k analogs (sklearn LinearRegression:
// centred least squares, minimum norm when rank deficient), prediction at the query point,
// in-sample RMSE.  With a threshold: exceedance probability from the logistic fit above and the
// regression restricted to the analogs above the threshold.  gard.py:191-224
template <typename T, int P, bool LOGIT, typename IdxF>
__device__ void regression_epilogue(const AnalogParams& a, int q, int64_t c, int k, const IdxF& idx, const double (&xq)[P]) {
    const T* X = (const T*)a.Xtr;
    const T* y = (const T*)a.ytr;
    const int p = (P == AN_PMAX) ? a.p : P;
    auto xv = [&](int i, int f) -> double { return (double)X[((int64_t)idx(i) * a.p + f) * a.ld + c]; };
    auto yv = [&](int i) -> double { return (double)y[(int64_t)idx(i) * a.ld + c]; };
    const T th = (T)a.thresh;
    // LOGIT (compile time) = a threshold was given: the plain regression keeps its register budget
    auto use = [&](int i) -> bool { return !LOGIT || (y[(int64_t)idx(i) * a.ld + c] > th); };   // gard.py:201-204
    int m = k;
    double prob = 1.0;
    if constexpr (LOGIT) {
        m = 0;
        for (int i = 0; i < k; ++i) m += use(i) ? 1 : 0;
        if (m == 0) {
            // no analog above the threshold: sklearn's LogisticRegression raises (one class only) — bit 1
            if (a.nonfinite) atomicOr(a.nonfinite, 2);
            store3(a, q, c, NAN, NAN, NAN);
            return;
        }
        if (m < k) prob = logistic_prob_class0<P>(k, p, xv, use, xq, a.logistic_c);
    }
    double xm[P], ym = 0.0;
#pragma unroll
    for (int f = 0; f < P; ++f) xm[f] = 0.0;
    for (int i = 0; i < k; ++i) {
        if (!use(i)) continue;
        ym += yv(i);
#pragma unroll
        for (int f = 0; f < P; ++f) if (f < p) xm[f] += xv(i, f);
    }
    ym /= (double)m;
#pragma unroll
    for (int f = 0; f < P; ++f) xm[f] /= (double)m;
    // normal equations of the centred problem: A (p x p, symmetric) beta = b
    double A[P][P], b[P];
#pragma unroll
    for (int f = 0; f < P; ++f) { b[f] = 0.0;
#pragma unroll
        for (int g = 0; g < P; ++g) A[f][g] = 0.0; }
    for (int i = 0; i < k; ++i) {
        if (!use(i)) continue;
        double dx[P];
#pragma unroll
        for (int f = 0; f < P; ++f) dx[f] = (f < p) ? xv(i, f) - xm[f] : 0.0;
        const double dy = yv(i) - ym;
#pragma unroll
        for (int f = 0; f < P; ++f) { b[f] += dx[f] * dy;
#pragma unroll
            for (int g = 0; g <= f; ++g) A[f][g] += dx[f] * dx[g]; }
    }
    double beta[P];
    bool solved = false;
    // no more regression samples than predictors: rank-deficient centred cloud → the minimum-norm solution
    // (what scipy's gelsd under sklearn's LinearRegression returns)
    if (m <= p) { sym_minnorm_solve<P>(A, b, p, beta); solved = true; }
    if (!solved) {
        // Cholesky A = L L^T.  A pivot that is zero up to roundoff RELATIVE to its diagonal entry (collinear or
        // duplicated analog rows) means the cloud is rank deficient: minimum-norm solution instead of huge
        // coefficients; an exactly constant predictor (zero diagonal) just drops out.
        bool live[P];
        bool deficient = false;
        double L[P][P];
#pragma unroll
        for (int f = 0; f < P; ++f) {
            live[f] = f < p;
#pragma unroll
            for (int g = 0; g <= f; ++g) {
                double s = A[f][g];
#pragma unroll
                for (int h = 0; h < g; ++h) s -= L[f][h] * L[g][h];
                if (g == f) {
                    if (live[f] && A[f][f] > 0.0 && !(s > 1e-13 * A[f][f])) deficient = true;
                    if (!(s > 1e-300) || !live[f]) { live[f] = false; L[f][f] = 1.0; }
                    else L[f][f] = sqrt(s);
                } else {
                    L[f][g] = live[g] ? s / L[g][g] : 0.0;
                }
            }
        }
        if (deficient) { sym_minnorm_solve<P>(A, b, p, beta); solved = true; }
        if (!solved) {
#pragma unroll
        for (int f = 0; f < P; ++f) {                      // forward substitution
            double s = b[f];
#pragma unroll
            for (int h = 0; h < f; ++h) s -= L[f][h] * beta[h];
            beta[f] = live[f] ? s / L[f][f] : 0.0;
        }
#pragma unroll
        for (int f = P - 1; f >= 0; --f) {                 // back substitution
            double s = beta[f];
#pragma unroll
            for (int h = f + 1; h < P; ++h) s -= L[h][f] * beta[h];
            beta[f] = live[f] ? s / L[f][f] : 0.0;
        }
        }
    }
    double icpt = ym;
#pragma unroll
    for (int f = 0; f < P; ++f) icpt -= xm[f] * beta[f];
    double sse = 0.0;
    for (int i = 0; i < k; ++i) {
        if (!use(i)) continue;
        double yh = icpt;
#pragma unroll
        for (int f = 0; f < P; ++f) if (f < p) yh += xv(i, f) * beta[f];
        const double r = yv(i) - yh;
        sse += r * r;
    }
    double pred = icpt;
#pragma unroll
    for (int f = 0; f < P; ++f) if (f < p) pred += xq[f] * beta[f];
    store3(a, q, c, pred, prob, sqrt(sse / (double)m));
}

// ---- the search kernel
template <typename T, int P, int KREG, bool LOGIT>
__global__ void __launch_bounds__(AN_THREADS)
analog_kernel(const AnalogParams a) {
    constexpr int AN_CHUNK = AnChunk<P>::value;
    constexpr int PF = AnFStride<P>::value;
    constexpr bool FILTER = (sizeof(T) == 4);         // float32 inputs: cheap float32 distance bound first
    __shared__ double chunk[AN_CHUNK * P];
    __shared__ __align__(16) float chunkf[FILTER ? AN_CHUNK * PF : 4];
    __shared__ double qd[(KREG > 0) ? AN_QD * AN_THREADS : 1];
    __shared__ int qi[(KREG > 0) ? AN_QD * AN_THREADS : 1];
    const int n_tiles = (a.t_query + AN_THREADS - 1) / AN_THREADS;
    const int64_t c = blockIdx.x / n_tiles;           // consecutive CTAs share a cell: its window stays in L2
    const int tile = blockIdx.x - (int)(c * n_tiles);
    if (a.valid && !a.valid[c]) {
        const int q = tile * AN_THREADS + threadIdx.x;
        if (q < a.t_query) store3(a, q, c, NAN, NAN, NAN);
        return;
    }
    const int p = (P == AN_PMAX) ? a.p : P;
    const int k = a.k;
    const T* Xtr = (const T*)a.Xtr;
    const T* Xq = (const T*)a.Xq;
    const int q = tile * AN_THREADS + threadIdx.x;
    const bool live = q < a.t_query;
    // the query point: float32 inputs are held as float32 only (the exact float64 value is its
    // widening), float64 inputs as float64
    double xq64[FILTER ? 1 : P];
    float xqf[PF];
#pragma unroll
    for (int f = 0; f < PF; ++f) xqf[f] = 0.0f;
#pragma unroll
    for (int f = 0; f < P; ++f) {
        const T raw = (live && f < p) ? Xq[((int64_t)q * a.p + f) * a.ld + c] : (T)0;
        if (FILTER) xqf[f] = (float)raw; else xq64[FILTER ? 0 : f] = (double)raw;
        if (live && f < p && a.nonfinite && !isfinite((double)raw)) atomicOr(a.nonfinite, 1);
    }
    auto xqv = [&](int f) -> double { return FILTER ? (double)xqf[f] : xq64[FILTER ? 0 : f]; };
    // running top-k, ascending by (distance, index)
    constexpr int KL = (KREG > 0) ? KREG : SDB_MAX_ANALOGS;
    double bd[KL];
    int bi[KL];
#pragma unroll
    for (int i = 0; i < (KREG > 0 ? KREG : 1); ++i) { bd[i] = INFINITY; bi[i] = -1; }
    if (KREG == 0) for (int i = 0; i < k; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }
    double worst = INFINITY;                          // current k-th best distance
    float worstf = live ? INFINITY : -1.0f;           // float32 upper bound of it (relative slack 1e-6 >> float32 error); lanes without a query never pass
    int qn = 0;                                       // candidates parked in this lane's queue

    // insert the parked candidates, oldest first (keeps "lowest index first on ties"); every lane of
    // the warp walks the same number of rounds
    auto drain = [&]() {
        for (int i = 0; __any_sync(0xffffffffu, i < qn); ++i) {
            if (i < qn) {
                const double d = qd[i * AN_THREADS + threadIdx.x];
                const int id = qi[i * AN_THREADS + threadIdx.x];
                if (d < worst) {
                    // branch-free insertion into the register list (strict < keeps the earlier index first on ties)
#pragma unroll
                    for (int j = (KREG > 0 ? KREG : 1) - 1; j > 0; --j) {
                        const bool shift = d < bd[j - 1];
                        const bool here = !shift && (d < bd[j]);
                        const double nd = shift ? bd[j - 1] : (here ? d : bd[j]);
                        const int ni = shift ? bi[j - 1] : (here ? id : bi[j]);
                        bd[j] = nd; bi[j] = ni;
                    }
                    if (d < bd[0]) { bd[0] = d; bi[0] = id; }
                    // k may be smaller than KREG: the k-th entry is the acceptance bound
                    double w = bd[(KREG > 0 ? KREG : 1) - 1];
#pragma unroll
                    for (int j = 0; j < (KREG > 0 ? KREG : 1); ++j) if (j == k - 1) w = bd[j];
                    worst = w;
                }
            }
        }
        if (live) worstf = isfinite(worst) ? __double2float_ru(worst * (1.0 + 1e-6)) : INFINITY;
        qn = 0;
    };

    for (int t0 = 0; t0 < a.t_fit; t0 += AN_CHUNK) {
        const int nt = min(AN_CHUNK, a.t_fit - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * p; i += AN_THREADS) {
            const int t = i / p, f = i - t * p;
            const T raw = Xtr[((int64_t)(t0 + t) * a.p + f) * a.ld + c];
            const double xv = (double)raw;
            if (a.nonfinite && tile == 0 && !isfinite(xv)) atomicOr(a.nonfinite, 1);
            chunk[t * P + f] = xv;
            if (FILTER) chunkf[t * PF + f] = (float)raw;
        }
        if (FILTER && PF > P) {
            for (int i = threadIdx.x; i < nt * (PF - P); i += AN_THREADS) {
                const int t = i / (PF - P), f = P + (i - t * (PF - P));
                chunkf[t * PF + f] = 0.0f;
            }
        }
        if (FILTER && P == AN_PMAX && p < P) {
            for (int i = threadIdx.x; i < nt * (P - p); i += AN_THREADS) {
                const int t = i / (P - p), f = p + (i - t * (P - p));
                chunkf[t * PF + f] = 0.0f;
            }
        }
        __syncthreads();
        if (KREG == 0 && !live) continue;
        // exact float64 distance of training point t (same operation order as the reference's KDTree)
        auto exact_d = [&](int t) -> double {
            double d = 0.0;
#pragma unroll
            for (int f = 0; f < P; ++f) {
                if (f < p) { const double tmp = xqv(f) - chunk[t * P + f]; d += tmp * tmp; }
            }
            return d;
        };
        auto accept = [&](int t, double d) {
            const int id = t0 + t;
            if (KREG > 0) {
                // park the candidate: the (long, branch-free) list insertion runs for the whole warp,
                // so it is batched — one pass inserts up to one candidate for every lane
                qd[qn * AN_THREADS + threadIdx.x] = d;
                qi[qn * AN_THREADS + threadIdx.x] = id;
                ++qn;
            } else {
                // large k: binary max-heap on (distance, index) in local memory — O(log k) per accepted
                // point.  d < root is strict, so a later point at the same distance never evicts an
                // earlier one (lowest index first on ties).
                int pos = 0;
                while (true) {
                    const int l = 2 * pos + 1;
                    if (l >= k) break;
                    const int r = l + 1;
                    int big = l;
                    if (r < k && (bd[l] < bd[r] || (bd[l] == bd[r] && bi[l] < bi[r]))) big = r;
                    if (!(d < bd[big] || (d == bd[big] && id < bi[big]))) break;
                    bd[pos] = bd[big]; bi[pos] = bi[big];
                    pos = big;
                }
                bd[pos] = d; bi[pos] = id;
                worst = bd[0];
                worstf = isfinite(worst) ? __double2float_ru(worst * (1.0 + 1e-6)) : INFINITY;
            }
        };
        if (FILTER) {
            // float32 lower-bound test, four training points per round: the exact float64 distance is
            // only evaluated for the few points that could enter the list (d32 <= worst * (1 + 1e-6)
            // whenever d64 < worst).  The queue (depth 8) is drained when a lane holds more than 4.
            const float4* cf4 = reinterpret_cast<const float4*>(chunkf);
            for (int tb = 0; tb < nt; tb += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = tb + u;
                    if (t < nt) {
                        float d32;
                        if (PF == 4) {
                            const float4 cf = cf4[t];
                            const float d0 = xqf[0] - cf.x;
                            d32 = d0 * d0;
                            if (P > 1) { const float d1 = xqf[1] - cf.y; d32 = fmaf(d1, d1, d32); }
                            if (P > 2) { const float d2 = xqf[2] - cf.z; d32 = fmaf(d2, d2, d32); }
                            if (P > 3) { const float d3 = xqf[3] - cf.w; d32 = fmaf(d3, d3, d32); }
                        } else {
                            const float4 ca = cf4[2 * t], cb = cf4[2 * t + 1];
                            const float d0 = xqf[0] - ca.x, d1 = xqf[1] - ca.y, d2 = xqf[2] - ca.z, d3 = xqf[3] - ca.w;
                            const float d4 = xqf[4] - cb.x, d5 = xqf[5] - cb.y, d6 = xqf[6] - cb.z, d7 = xqf[7] - cb.w;
                            d32 = d0 * d0;
                            d32 = fmaf(d1, d1, d32); d32 = fmaf(d2, d2, d32); d32 = fmaf(d3, d3, d32);
                            d32 = fmaf(d4, d4, d32); d32 = fmaf(d5, d5, d32); d32 = fmaf(d6, d6, d32); d32 = fmaf(d7, d7, d32);
                        }
                        if (d32 <= worstf) {
                            const double d = exact_d(t);
                            if (d < worst) accept(t, d);
                        }
                    }
                }
                if (KREG > 0) {
                    if (__any_sync(0xffffffffu, qn > AN_QD - 4)) drain();
                }
            }
        } else {
            for (int tb = 0; tb < nt; tb += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = tb + u;
                    if (t < nt && live) {
                        const double d = exact_d(t);
                        if (d < worst) accept(t, d);
                    }
                }
                if (KREG > 0) {
                    if (__any_sync(0xffffffffu, qn > AN_QD - 4)) drain();
                }
            }
        }
    }
    if (KREG > 0) drain();
    if (!live) return;
    if (KREG == 0) {
        // heap → ascending (distance, index) order, in place
        for (int end = k - 1; end > 0; --end) {
            const double d = bd[end]; const int id = bi[end];
            bd[end] = bd[0]; bi[end] = bi[0];
            int pos = 0;
            while (true) {
                const int l = 2 * pos + 1;
                if (l >= end) break;
                const int r = l + 1;
                int big = l;
                if (r < end && (bd[l] < bd[r] || (bd[l] == bd[r] && bi[l] < bi[r]))) big = r;
                if (!(d < bd[big] || (d == bd[big] && id < bi[big]))) break;
                bd[pos] = bd[big]; bi[pos] = bi[big];
                pos = big;
            }
            bd[pos] = d; bi[pos] = id;
        }
    }
    if (a.knn_idx) {
        if (KREG > 0) {
#pragma unroll
            for (int i = 0; i < KREG; ++i) if (i < k) a.knn_idx[((int64_t)q * k + i) * a.C + c] = bi[i];
        } else {
            for (int i = 0; i < k; ++i) a.knn_idx[((int64_t)q * k + i) * a.C + c] = bi[i];
        }
    }
    if constexpr (KREG > 0) {
        // park the winners in local arrays the epilogues can index dynamically (the search
        // list itself stays in registers)
        int li[KREG]; double ld2[KREG];
#pragma unroll
        for (int i = 0; i < KREG; ++i) { li[i] = bi[i]; ld2[i] = bd[i]; }
        auto idx = [&](int i) -> int { return li[i]; };
        auto dist2 = [&](int i) -> double { return ld2[i]; };
        double xq[P];
#pragma unroll
        for (int f = 0; f < P; ++f) xq[f] = xqv(f);
        if (LOGIT || a.kind == SDB_ANALOG_REGRESSION) regression_epilogue<T, P, LOGIT>(a, q, c, k, idx, xq);
        else pure_analog_epilogue<T>(a, q, c, k, idx, dist2);
    } else {
        auto idx = [&](int i) -> int { return bi[i]; };
        auto dist2 = [&](int i) -> double { return bd[i]; };
        double xq[P];
#pragma unroll
        for (int f = 0; f < P; ++f) xq[f] = xqv(f);
        if (LOGIT || a.kind == SDB_ANALOG_REGRESSION) regression_epilogue<T, P, LOGIT>(a, q, c, k, idx, xq);
        else pure_analog_epilogue<T>(a, q, c, k, idx, dist2);
    }
}

template <typename T, int P>
static int launch_analog(const AnalogParams& a, cudaStream_t st) {
    const int64_t n_tiles = (a.t_query + AN_THREADS - 1) / AN_THREADS;
    if (n_tiles * a.C > 2147483647LL) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_predict: too many (cell, query tile) pairs for one launch; split the shard");
    dim3 grid((unsigned)(n_tiles * a.C));
    const bool logit = (a.kind == SDB_ANALOG_REGRESSION) && a.has_thresh;
    if (a.k <= AN_KREG) {
        if (logit) analog_kernel<T, P, AN_KREG, true><<<grid, AN_THREADS, 0, st>>>(a);
        else       analog_kernel<T, P, AN_KREG, false><<<grid, AN_THREADS, 0, st>>>(a);
    } else {
        if (logit) analog_kernel<T, P, 0, true><<<grid, AN_THREADS, 0, st>>>(a);
        else       analog_kernel<T, P, 0, false><<<grid, AN_THREADS, 0, st>>>(a);
    }
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
static int dispatch_analog(const AnalogParams& a, cudaStream_t st) {
    switch (a.p) {
        case 1: return launch_analog<T, 1>(a, st);
        case 2: return launch_analog<T, 2>(a, st);
        case 3: return launch_analog<T, 3>(a, st);
        case 4: return launch_analog<T, 4>(a, st);
        default: return launch_analog<T, AN_PMAX>(a, st);
    }
}


// ---------------------------------------------------------------- pruned search: one CTA per cell (round 2)
// The reference prunes its search with a per-cell KDTree (gard.py:82,194,299); the brute-force kernel above
// re-stages the whole training window for every 256 queries and evaluates all T_fit distances per query.  Here
//   * ONE CTA owns a cell: its training window, ordered by the boxes of a quantile grid (analog_grid.cu; up to 512
//     boxes of ~T/512 rows), is staged ONCE into shared memory as float32 (12 bytes per row for 3 predictors:
//     18 250 rows = 219 KB) and serves all of the cell's queries;
//   * one thread = one query at a time.  It visits the boxes shell by shell around its own box (index distance
//     0, 1, 2, ...), skipping boxes whose lower bound sum_f gap_f^2 exceeds its current k-th best distance, and stops
//     when the nearest plane of the next shell is farther than that distance — nothing beyond can enter the list.
//     Bounds and distances are float64 sums of squares taken in the same order (f = 0, 1, 2), and float64
//     additions / squares of non-negative terms are monotone, so bound <= distance holds in floating point:
//     the result is EXACT (same neighbours, same order as the brute force; the lower training row wins a tie);
//   * distances are screened in float32 and re-evaluated in float64 from the same float32 values (the exact
//     widening of the inputs — what the reference's float64 KDTree sees).
// About 1 row in 15 is visited for 3 standard-normal predictors, k = 10, 30 years of days.
constexpr int AC_THREADS = 512;
constexpr int AC_K = 16;             // most neighbours kept in registers (list capacities built: 1, 10, 16)

template <int P> constexpr size_t ac_smem_bytes(int t_fit) {
    return (size_t)t_fit * P * 4 + (size_t)(AG_BOXES + 1) * 4 + (size_t)AG_NBND * 4 + 16;
}

template <int P, bool LOGIT, int KC>
__global__ void __launch_bounds__(AC_THREADS, 1)
analog_cell_kernel(const AnalogParams a) {
    extern __shared__ __align__(16) float ac_smem[];
    const int T = a.t_fit;
    float* pts = ac_smem;                                           // [T][P], box order
    int* bstart = reinterpret_cast<int*>(ac_smem + (size_t)T * P);  // [AG_BOXES + 1]
    float* planes = reinterpret_cast<float*>(bstart + AG_BOXES + 1);
    const int64_t c = blockIdx.x;
    const int tid = threadIdx.x;
    if (a.valid && !a.valid[c]) {
        for (int q = tid; q < a.t_query; q += AC_THREADS) store3(a, q, c, NAN, NAN, NAN);
        return;
    }
    const float* Xtr = (const float*)a.Xtr;
    const float* Xq = (const float*)a.Xq;
    int g[3];
    ag_dims(P, g);
    const int nplanes = (g[0] - 1) + (g[1] - 1) + (g[2] - 1);
    const int nbox = g[0] * g[1] * g[2];
    for (int i = tid; i < nplanes; i += AC_THREADS) planes[i] = a.bounds[(int64_t)i * a.ld_grid + c];
    for (int i = tid; i <= AG_BOXES; i += AC_THREADS) bstart[i] = a.box_start[(int64_t)i * a.ld_grid + c];
    bool bad = false;
    for (int i = tid; i < T; i += AC_THREADS) {
        const int row = a.perm_t[(int64_t)i * a.ld_grid + c];
#pragma unroll
        for (int f = 0; f < P; ++f) {
            const float v = Xtr[((int64_t)row * P + f) * a.ld + c];
            bad |= !isfinite(v);
            pts[(size_t)i * P + f] = v;
        }
    }
    if (bad && a.nonfinite) atomicOr(a.nonfinite, 1);
    __syncthreads();
    const float* pl[3] = {planes, planes + (g[0] - 1), planes + (g[0] - 1) + (g[1] - 1)};
    const int k = a.k;
    auto row_of = [&](int pos) -> int { return a.perm_t[(int64_t)pos * a.ld_grid + c]; };

    for (int qs = tid; qs < a.t_query; qs += AC_THREADS) {
        const int q = a.perm_q[(int64_t)qs * a.ld_grid + c];
        float xqf[P];
        bool qbad = false;
#pragma unroll
        for (int f = 0; f < P; ++f) { xqf[f] = Xq[((int64_t)q * P + f) * a.ld + c]; qbad |= !isfinite(xqf[f]); }
        if (qbad && a.nonfinite) atomicOr(a.nonfinite, 1);
        int qb[3] = {0, 0, 0};
#pragma unroll
        for (int f = 0; f < 3; ++f)
            if (f < P && g[f] > 1) qb[f] = ag_slab(xqf[f], pl[f], g[f]);
        // neighbour list: ascending by (distance, training row); bpos = position in box order, row fetched on demand
        double bd[KC];
        int bpos[KC];
#pragma unroll
        for (int i = 0; i < KC; ++i) { bd[i] = INFINITY; bpos[i] = -1; }
        double worst = INFINITY;
        int worst_pos = -1;
        float worstf = INFINITY;
        auto before = [&](double d, int pos, double d2, int pos2) -> bool {      // (d, row(pos)) < (d2, row(pos2))
            if (d != d2) return d < d2;
            if (pos2 < 0) return true;
            return row_of(pos) < row_of(pos2);
        };
        auto insert = [&](double d, int pos) {
#pragma unroll
            for (int j = KC - 1; j > 0; --j) {
                const bool shift = before(d, pos, bd[j - 1], bpos[j - 1]);
                const bool here = !shift && before(d, pos, bd[j], bpos[j]);
                const double nd = shift ? bd[j - 1] : (here ? d : bd[j]);
                const int np = shift ? bpos[j - 1] : (here ? pos : bpos[j]);
                bd[j] = nd; bpos[j] = np;
            }
            if (before(d, pos, bd[0], bpos[0])) { bd[0] = d; bpos[0] = pos; }
            double w = bd[KC - 1];
            int wp = bpos[KC - 1];
#pragma unroll
            for (int j = 0; j < KC - 1; ++j) if (j == k - 1) { w = bd[j]; wp = bpos[j]; }
            worst = w; worst_pos = wp;
            worstf = isfinite(worst) ? __double2float_ru(worst * (1.0 + 1e-6)) : INFINITY;
        };
        auto gap = [&](int f, int b) -> double {      // distance of q_f to slab b of predictor f: rows there satisfy plane[b-1] <= x < plane[b]
            const double qf = (double)xqf[f < P ? f : 0];
            const double lo = (b > 0) ? (double)pl[f][b - 1] : -INFINITY;
            const double hi = (b < g[f] - 1) ? (double)pl[f][b] : INFINITY;
            return qf < lo ? lo - qf : (qf > hi ? qf - hi : 0.0);
        };
        auto scan_box = [&](int box) {
            const int r0 = bstart[box], r1 = bstart[box + 1];
            for (int t = r0; t < r1; ++t) {
                const float* x = pts + (size_t)t * P;
                const float d0 = xqf[0] - x[0];
                float d32 = d0 * d0;
                if (P > 1) { const float d1 = xqf[1] - x[1]; d32 = fmaf(d1, d1, d32); }
                if (P > 2) { const float d2 = xqf[2] - x[2]; d32 = fmaf(d2, d2, d32); }
                if (d32 <= worstf) {
                    double d = 0.0;
#pragma unroll
                    for (int f = 0; f < P; ++f) { const double tmp = (double)xqf[f] - (double)x[f]; d += tmp * tmp; }
                    if (before(d, t, worst, worst_pos)) insert(d, t);
                }
            }
        };
        const int rmax = max(g[0], max(g[1], g[2]));
        for (int r = 0; r < rmax; ++r) {
            const int r1 = g[1] > 1 ? r : 0, r2 = g[2] > 1 ? r : 0;
            for (int d0 = -r; d0 <= r; ++d0) {
                const int b0 = qb[0] + d0;
                if (b0 < 0 || b0 >= g[0]) continue;
                const double t0 = gap(0, b0);
                const double l0 = t0 * t0;
                if (l0 > worst) continue;
                for (int d1 = -r1; d1 <= r1; ++d1) {
                    const int b1 = qb[1] + d1;
                    if (b1 < 0 || b1 >= g[1]) continue;
                    double l1 = l0;
                    if (g[1] > 1) { const double t1 = gap(1, b1); l1 += t1 * t1; }
                    if (l1 > worst) continue;
                    const bool edge01 = (d0 == -r || d0 == r || (g[1] > 1 && (d1 == -r || d1 == r)));
                    // boxes of this shell: index distance exactly r in at least one predictor
                    const int step2 = (edge01 || r2 == 0) ? 1 : 2 * r2;
                    for (int d2 = -r2; d2 <= r2; d2 += step2) {
                        if (!edge01 && g[2] > 1 && d2 != -r && d2 != r) continue;
                        if (!edge01 && g[2] == 1 && r > 0) continue;
                        const int b2 = qb[2] + d2;
                        if (b2 < 0 || b2 >= g[2]) continue;
                        double l2 = l1;
                        if (g[2] > 1) { const double t2 = gap(2, b2); l2 += t2 * t2; }
                        if (l2 > worst) continue;
                        scan_box((b0 * g[1] + b1) * g[2] + b2);
                    }
                }
            }
            // every unvisited box is at index distance >= r + 1 along some predictor: its gap there is at least the
            // distance to the nearest plane of that shell
            double gmin = INFINITY;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                if (g[f] > 1) {
                    const double qf = (double)xqf[f < P ? f : 0];
                    if (qb[f] - r - 1 >= 0) gmin = fmin(gmin, fmax(qf - (double)pl[f][qb[f] - r - 1], 0.0));
                    if (qb[f] + r + 1 <= g[f] - 1) gmin = fmin(gmin, fmax((double)pl[f][qb[f] + r] - qf, 0.0));
                }
            }
            if (!(gmin * gmin <= worst)) break;
        }
        // rows of the neighbours, then the model's statistic
        int li[KC];
        double ld2[KC];
#pragma unroll
        for (int i = 0; i < KC; ++i) { li[i] = (i < k && bpos[i] >= 0) ? row_of(bpos[i]) : 0; ld2[i] = bd[i]; }
        if (a.knn_idx) {
#pragma unroll
            for (int i = 0; i < KC; ++i) if (i < k) a.knn_idx[((int64_t)q * k + i) * a.C + c] = li[i];
        }
        auto idx = [&](int i) -> int { return li[i]; };
        auto dist2 = [&](int i) -> double { return ld2[i]; };
        double xq[P];
#pragma unroll
        for (int f = 0; f < P; ++f) xq[f] = (double)xqf[f];
        if (LOGIT || a.kind == SDB_ANALOG_REGRESSION) regression_epilogue<float, P, LOGIT>(a, q, c, k, idx, xq);
        else pure_analog_epilogue<float>(a, q, c, k, idx, dist2);
    }
}

template <int P, bool LOGIT, int KC>
static int launch_analog_cell_kc(const AnalogParams& a, cudaStream_t st) {
    const size_t smem = ac_smem_bytes<P>(a.t_fit);
    auto kern = analog_cell_kernel<P, LOGIT, KC>;
    SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)a.C, AC_THREADS, smem, st>>>(a);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int P>
static int launch_analog_cell(const AnalogParams& a, cudaStream_t st) {
    const bool logit = (a.kind == SDB_ANALOG_REGRESSION) && a.has_thresh;
    if (logit) return launch_analog_cell_kc<P, true, AC_K>(a, st);      // the threshold models keep one list capacity
    if (a.k == 1) return launch_analog_cell_kc<P, false, 1>(a, st);
    if (a.k <= 10) return launch_analog_cell_kc<P, false, 10>(a, st);
    return launch_analog_cell_kc<P, false, AC_K>(a, st);
}

// largest training window the per-cell kernel can stage (227 KB of shared memory per CTA)
static int analog_cell_max_steps(int p) {
    const size_t fixed = (size_t)(AG_BOXES + 1) * 4 + (size_t)AG_NBND * 4 + 16;
    return (int)((232448 - fixed) / ((size_t)p * 4));
}

// ---------------------------------------------------------------- PureRegression (gard.py:367-504)
// One thread per cell (rows coalesced across cells): centred normal equations of the rows above the
// threshold accumulated in float64 over the whole training window → Cholesky (Jacobi minimum-norm
// when rank deficient) → in-sample RMSE; with a threshold the logistic exceedance model of ALL rows by
// damped Newton (one pass over the window per evaluation).  model[c * MODEL_LD + ...]:
//   [0..P) beta, [P] intercept, [P+1] rmse, [P+2..2P+2) logistic weights, [2P+2] logistic intercept,
//   [2P+3] status: 0 = two classes, 1 = no threshold / every row exceeds (prob = 1), 2 = no row exceeds
constexpr int PR_MODEL_LD = 2 * AN_PMAX + 4;

template <typename T, int P>
__global__ void pure_regression_fit_kernel(const T* __restrict__ X, const T* __restrict__ y, int64_t ld, int64_t C,
                                           int t_fit, int p_rt, int has_thresh, double thresh, double logistic_c,
                                           double* __restrict__ model, const uint8_t* __restrict__ valid,
                                           int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double* mo = model + c * PR_MODEL_LD;
    if (valid && !valid[c]) { for (int i = 0; i < PR_MODEL_LD; ++i) mo[i] = NAN; return; }
    const int p = (P == AN_PMAX) ? p_rt : P;
    const T th = (T)thresh;
    auto xv = [&](int i, int f) -> double { return (double)X[((int64_t)i * p + f) * ld + c]; };
    auto yraw = [&](int i) -> T { return y[(int64_t)i * ld + c]; };
    auto use = [&](int i) -> bool { return !has_thresh || (yraw(i) > th); };
    int m = 0;
    double xm[P], ym = 0.0;
    bool bad = false;
#pragma unroll
    for (int f = 0; f < P; ++f) xm[f] = 0.0;
    for (int i = 0; i < t_fit; ++i) {
        const double yi = (double)yraw(i);
        bad |= !isfinite(yi);
        const bool u = use(i);
        m += u ? 1 : 0;
        if (u) ym += yi;
#pragma unroll
        for (int f = 0; f < P; ++f) if (f < p) { const double x = xv(i, f); bad |= !isfinite(x); if (u) xm[f] += x; }
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    double status = (has_thresh && m < t_fit) ? 0.0 : 1.0;
    if (m == 0) {                                   // gard.py:435: LinearRegression on an empty selection raises — bit 2
        if (nonfinite) atomicOr(nonfinite, 4);
        for (int i = 0; i < PR_MODEL_LD; ++i) mo[i] = NAN;
        mo[2 * P + 3] = 2.0;
        return;
    }
    ym /= (double)m;
#pragma unroll
    for (int f = 0; f < P; ++f) xm[f] /= (double)m;
    double A[P][P], b[P];
#pragma unroll
    for (int f = 0; f < P; ++f) { b[f] = 0.0;
#pragma unroll
        for (int g = 0; g < P; ++g) A[f][g] = 0.0; }
    for (int i = 0; i < t_fit; ++i) {
        if (!use(i)) continue;
        double dx[P];
#pragma unroll
        for (int f = 0; f < P; ++f) dx[f] = (f < p) ? xv(i, f) - xm[f] : 0.0;
        const double dy = (double)yraw(i) - ym;
#pragma unroll
        for (int f = 0; f < P; ++f) { b[f] += dx[f] * dy;
#pragma unroll
            for (int g = 0; g <= f; ++g) A[f][g] += dx[f] * dx[g]; }
    }
    double beta[P];
    bool solved = false;
    if (m <= p) { sym_minnorm_solve<P>(A, b, p, beta); solved = true; }
    if (!solved) {
        bool live[P];
        bool deficient = false;
        double L[P][P];
#pragma unroll
        for (int f = 0; f < P; ++f) {
            live[f] = f < p;
#pragma unroll
            for (int g = 0; g <= f; ++g) {
                double sacc = A[f][g];
#pragma unroll
                for (int h = 0; h < g; ++h) sacc -= L[f][h] * L[g][h];
                if (g == f) {
                    if (live[f] && !(sacc > 1e-13 * A[f][f])) deficient = true;
                    if (!(sacc > 0.0) || !live[f]) { live[f] = false; L[f][f] = 1.0; } else L[f][f] = sqrt(sacc);
                } else {
                    L[f][g] = live[g] ? sacc / L[g][g] : 0.0;
                }
            }
        }
        if (deficient) {
            sym_minnorm_solve<P>(A, b, p, beta);       // collinear predictors: lstsq's minimum-norm answer
        } else {
#pragma unroll
            for (int f = 0; f < P; ++f) {
                double sacc = b[f];
#pragma unroll
                for (int h = 0; h < f; ++h) sacc -= L[f][h] * beta[h];
                beta[f] = live[f] ? sacc / L[f][f] : 0.0;
            }
#pragma unroll
            for (int f = P - 1; f >= 0; --f) {
                double sacc = beta[f];
#pragma unroll
                for (int h = f + 1; h < P; ++h) sacc -= L[h][f] * beta[h];
                beta[f] = live[f] ? sacc / L[f][f] : 0.0;
            }
        }
    }
    double icpt = ym;
#pragma unroll
    for (int f = 0; f < P; ++f) icpt -= xm[f] * beta[f];
    double sse = 0.0;
    for (int i = 0; i < t_fit; ++i) {
        if (!use(i)) continue;
        double yh = icpt;
#pragma unroll
        for (int f = 0; f < P; ++f) if (f < p) yh += xv(i, f) * beta[f];
        const double r = (double)yraw(i) - yh;
        sse += r * r;
    }
#pragma unroll
    for (int f = 0; f < P; ++f) mo[f] = beta[f];
    mo[P] = icpt;
    mo[P + 1] = sqrt(sse / (double)m);
    mo[2 * P + 3] = status;
    if (status == 0.0) {
        // logistic exceedance model of ALL rows (gard.py:417): the Newton solver of the analog epilogue
        double w[P + 1];
        logistic_fit<P>(t_fit, p, xv, use, logistic_c, w);
#pragma unroll
        for (int f = 0; f <= P; ++f) mo[P + 2 + f] = w[f];
    } else {
#pragma unroll
        for (int f = 0; f <= P; ++f) mo[P + 2 + f] = 0.0;
    }
}

template <typename T, int P>
__global__ void pure_regression_predict_kernel(const T* __restrict__ Xq, int64_t ld, int64_t C, int t_query, int p_rt,
                                               const double* __restrict__ model, void* __restrict__ out, int out_f64,
                                               int64_t ld_out, const uint8_t* __restrict__ valid,
                                               int32_t* __restrict__ nonfinite) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (c >= C || q >= t_query) return;
    const int p = (P == AN_PMAX) ? p_rt : P;
    const int64_t base = (int64_t)q * 3 * ld_out + c;
    double pred = NAN, prob = NAN, err = NAN;
    if (!valid || valid[c]) {
        const double* mo = model + c * PR_MODEL_LD;
        pred = mo[P];
        double z = mo[2 * P + 2];
#pragma unroll
        for (int f = 0; f < P; ++f) {
            if (f < p) {
                const double x = (double)Xq[((int64_t)q * p + f) * ld + c];
                if (nonfinite && !isfinite(x)) atomicOr(nonfinite, 1);
                pred += x * mo[f];
                z += x * mo[P + 2 + f];
            }
        }
        err = mo[P + 1];
        prob = (mo[2 * P + 3] == 0.0) ? 1.0 / (1.0 + exp(-z)) : 1.0;      // predict_proba[:, 1]   gard.py:467
    }
    if (out_f64) { double* o = (double*)out; o[base] = pred; o[base + ld_out] = prob; o[base + 2 * ld_out] = err; }
    else { float* o = (float*)out; o[base] = (float)pred; o[base + ld_out] = (float)prob; o[base + 2 * ld_out] = (float)err; }
}

template <typename T, int P>
static int launch_pure_regression_fit(const T* X, const T* y, int64_t ld, int64_t C, int t_fit, int p, int has_thresh,
                                      double thresh, double logistic_c, double* model, const uint8_t* valid,
                                      int32_t* nonfinite, cudaStream_t st) {
    pure_regression_fit_kernel<T, P><<<(unsigned)((C + 63) / 64), 64, 0, st>>>(X, y, ld, C, t_fit, p, has_thresh, thresh, logistic_c, model, valid, nonfinite);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}
template <typename T, int P>
static int launch_pure_regression_predict(const T* Xq, int64_t ld, int64_t C, int t_query, int p, const double* model,
                                          void* out, int out_f64, int64_t ld_out, const uint8_t* valid,
                                          int32_t* nonfinite, cudaStream_t st) {
    dim3 grid((unsigned)((C + 127) / 128), (unsigned)t_query);
    pure_regression_predict_kernel<T, P><<<grid, 128, 0, st>>>(Xq, ld, C, t_query, p, model, out, out_f64, ld_out, valid, nonfinite);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace sdb

using namespace sdb;

#define SDB_PR_DISPATCH(FN, T, ...)                                             \
    switch (n_features) {                                                       \
        case 1: return FN<T, 1>(__VA_ARGS__);                                   \
        case 2: return FN<T, 2>(__VA_ARGS__);                                   \
        case 3: return FN<T, 3>(__VA_ARGS__);                                   \
        case 4: return FN<T, 4>(__VA_ARGS__);                                   \
        default: return FN<T, AN_PMAX>(__VA_ARGS__);                            \
    }

extern "C" int sdb_pure_regression_model_ld(void) { return PR_MODEL_LD; }

extern "C" int sdb_pure_regression_fit(const void* X_train, const void* y_train, int dtype, int64_t ld, int64_t n_cells,
                                       int t_fit, int n_features, int has_thresh, double thresh, double logistic_c,
                                       double* model, const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X_train || !y_train || !model) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_fit: NULL pointer");
    if (n_cells <= 0 || t_fit <= 0 || ld < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_fit: bad shape");
    if (n_features < 1 || n_features > AN_PMAX) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_pure_regression_fit: 1..%d predictors supported, got %d", AN_PMAX, n_features);
    if (has_thresh && !(logistic_c > 0.0)) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_fit: logistic_c must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SDB_F32) { SDB_PR_DISPATCH(launch_pure_regression_fit, float, (const float*)X_train, (const float*)y_train, ld, n_cells, t_fit, n_features, has_thresh, thresh, logistic_c, model, cell_valid, nonfinite, st) }
    if (dtype == SDB_F64) { SDB_PR_DISPATCH(launch_pure_regression_fit, double, (const double*)X_train, (const double*)y_train, ld, n_cells, t_fit, n_features, has_thresh, thresh, logistic_c, model, cell_valid, nonfinite, st) }
    return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_fit: bad dtype %d", dtype);
}

extern "C" int sdb_pure_regression_predict(const void* X_query, int dtype, int64_t ld, int64_t n_cells, int t_query,
                                           int n_features, const double* model, void* out, int out_dtype, int64_t ld_out,
                                           const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    if (!X_query || !model || !out) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_predict: NULL pointer");
    if (n_cells <= 0 || t_query <= 0 || ld < n_cells || ld_out < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_predict: bad shape");
    if (t_query > 65535) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_pure_regression_predict: at most 65535 query steps per call");
    if (n_features < 1 || n_features > AN_PMAX) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_pure_regression_predict: 1..%d predictors supported, got %d", AN_PMAX, n_features);
    if (out_dtype != SDB_F32 && out_dtype != SDB_F64) return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_predict: bad out dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const int of64 = out_dtype == SDB_F64;
    if (dtype == SDB_F32) { SDB_PR_DISPATCH(launch_pure_regression_predict, float, (const float*)X_query, ld, n_cells, t_query, n_features, model, out, of64, ld_out, cell_valid, nonfinite, st) }
    if (dtype == SDB_F64) { SDB_PR_DISPATCH(launch_pure_regression_predict, double, (const double*)X_query, ld, n_cells, t_query, n_features, model, out, of64, ld_out, cell_valid, nonfinite, st) }
    return sdb_fail(SDB_E_INVALID, "sdb_pure_regression_predict: bad dtype %d", dtype);
}

static int analog_predict_impl(int kind, const void* X_train, const void* y_train, const void* X_query,
                                  int dtype, int64_t ld, int64_t n_cells,
                                  int t_fit, int t_query, int n_features, int k,
                                  int has_thresh, double thresh, double logistic_c, const int32_t* rand_idx,
                                  void* out, int out_dtype, int64_t ld_out, int32_t* knn_idx,
                                  const uint8_t* cell_valid, int32_t* nonfinite,
                                  const int32_t* perm_train, const int32_t* perm_query, const int32_t* box_start, const float* bounds,
                                  int64_t ld_grid, void* stream) {
    if (!X_train || !y_train || !X_query || !out) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: NULL pointer");
    if (n_cells <= 0 || t_fit <= 0 || t_query <= 0 || ld < n_cells || ld_out < n_cells)
        return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: bad shape");
    if (n_features < 1 || n_features > AN_PMAX) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_predict: 1..%d predictors supported, got %d", AN_PMAX, n_features);
    if (k < 1 || k > SDB_MAX_ANALOGS || k > t_fit) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: k=%d out of range (1..min(%d, T_fit))", k, SDB_MAX_ANALOGS);
    if (kind < SDB_ANALOG_BEST || kind > SDB_ANALOG_REGRESSION) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: unknown kind %d", kind);
    if (kind == SDB_ANALOG_SAMPLE && !rand_idx) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: sample_analogs needs rand_idx");
    if (kind == SDB_ANALOG_REGRESSION && has_thresh && !(logistic_c > 0.0)) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: logistic_c must be positive");
    if ((dtype != SDB_F32 && dtype != SDB_F64) || (out_dtype != SDB_F32 && out_dtype != SDB_F64))
        return sdb_fail(SDB_E_INVALID, "sdb_analog_predict: bad dtype");
    AnalogParams a;
    a.Xtr = X_train; a.ytr = y_train; a.Xq = X_query; a.ld = ld; a.C = n_cells;
    a.t_fit = t_fit; a.t_query = t_query; a.p = n_features; a.k = k; a.kind = kind;
    a.has_thresh = has_thresh; a.thresh = thresh; a.logistic_c = logistic_c; a.rand_idx = rand_idx;
    a.out = out; a.out_f64 = (out_dtype == SDB_F64); a.ld_out = ld_out; a.knn_idx = knn_idx;
    a.valid = cell_valid; a.nonfinite = nonfinite;
    a.perm_t = perm_train; a.perm_q = perm_query; a.box_start = box_start; a.bounds = bounds; a.ld_grid = ld_grid;
    cudaStream_t st = (cudaStream_t)stream;
    if (perm_train) {
        switch (n_features) {
            case 1: return launch_analog_cell<1>(a, st);
            case 2: return launch_analog_cell<2>(a, st);
            default: return launch_analog_cell<3>(a, st);
        }
    }
    return dtype == SDB_F32 ? dispatch_analog<float>(a, st) : dispatch_analog<double>(a, st);
}

extern "C" int sdb_analog_predict(int kind, const void* X_train, const void* y_train, const void* X_query,
                                  int dtype, int64_t ld, int64_t n_cells,
                                  int t_fit, int t_query, int n_features, int k,
                                  int has_thresh, double thresh, double logistic_c, const int32_t* rand_idx,
                                  void* out, int out_dtype, int64_t ld_out, int32_t* knn_idx,
                                  const uint8_t* cell_valid, int32_t* nonfinite, void* stream) {
    return analog_predict_impl(kind, X_train, y_train, X_query, dtype, ld, n_cells, t_fit, t_query, n_features, k, has_thresh, thresh,
                               logistic_c, rand_idx, out, out_dtype, ld_out, knn_idx, cell_valid, nonfinite,
                               nullptr, nullptr, nullptr, nullptr, 0, stream);
}

// which (n_features, k, t_fit) the pruned per-cell search covers: 1..3 predictors, k <= 16, the window must fit in
// shared memory (18 958 steps for 3 predictors); 0 = use sdb_analog_predict
extern "C" int sdb_analog_pruned_supported(int dtype, int t_fit, int n_features, int k) {
    return dtype == SDB_F32 && n_features >= 1 && n_features <= 3 && k >= 1 && k <= AC_K && t_fit >= 64 &&
           t_fit <= analog_cell_max_steps(n_features) && t_fit <= sdb_series_argsort_max_steps();
}

extern "C" int sdb_analog_predict_pruned(int kind, const void* X_train, const void* y_train, const void* X_query,
                                         int dtype, int64_t ld, int64_t n_cells,
                                         int t_fit, int t_query, int n_features, int k,
                                         int has_thresh, double thresh, double logistic_c, const int32_t* rand_idx,
                                         void* out, int out_dtype, int64_t ld_out, int32_t* knn_idx,
                                         const uint8_t* cell_valid, int32_t* nonfinite,
                                         const int32_t* perm_train, const int32_t* perm_query, const int32_t* box_start,
                                         const float* bounds, int64_t ld_grid, void* stream) {
    if (!perm_train || !perm_query || !box_start || !bounds) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict_pruned: NULL grid table");
    if (ld_grid < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_analog_predict_pruned: bad ld_grid");
    if (!sdb_analog_pruned_supported(dtype, t_fit, n_features, k))
        return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_predict_pruned: float32, 1..3 predictors, k <= %d, at most %d steps (use sdb_analog_predict)",
                        AC_K, analog_cell_max_steps(n_features < 1 ? 1 : (n_features > 3 ? 3 : n_features)));
    return analog_predict_impl(kind, X_train, y_train, X_query, dtype, ld, n_cells, t_fit, t_query, n_features, k, has_thresh, thresh,
                               logistic_c, rand_idx, out, out_dtype, ld_out, knn_idx, cell_valid, nonfinite,
                               perm_train, perm_query, box_start, bounds, ld_grid, stream);
}

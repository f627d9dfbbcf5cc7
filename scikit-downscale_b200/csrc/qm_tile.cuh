// qm_tile.cuh — the fast path of the quantile-mapping kernels (float32 in/out, groups of up to
// 1024 steps, 9-sample window inside the mapping group): what the headline workload runs.
//
// One CTA = 8 warps owns a tile of 8 consecutive cells x one time group:
//   * the group's rows are staged into shared memory with coalesced loads (each 8-cell row
//     segment is one full 32-byte sector) and written back the same way — the strided
//     (time-major) HBM layout is only ever touched by full-sector row accesses;
//   * each warp then owns one cell: 32 values per lane in registers, the bitonic network of
//     sort.cuh (shuffles only);
//   * predict ranks a group with ONE 32-bit keys-only sort: the float64 rank key is quantised
//     monotonically to (32 - log2 NP) bits and packed with the element's position,
//     w = q << log2(NP) | j.  q_a < q_b implies key_a < key_b, so only elements that share a q
//     bucket need the exact float64 comparison; that fix-up is local (runs of <= 16) and
//     otherwise the warp falls back to the exact 64-bit key+payload sort.  Results are
//     therefore identical to the generic kernel for every input.
#pragma once
#include "qm_kernels.cuh"

namespace sdb {

constexpr int TILE_CT = 8;          // cells per CTA (= warps per CTA)
constexpr int TILE_THREADS = 32 * TILE_CT;
constexpr int TILE_LMAX = 16;       // longest same-bucket run fixed up locally

template <int E> struct TileGeom {
    static constexpr int NP = 32 * E;
    static constexpr int NPS = NP + NP / 32 + 4;       // padded row: conflict-free for every access pattern used
    static constexpr int LOG = (E == 8) ? 8 : 10;
    static constexpr uint32_t QMAX = (1u << (32 - LOG)) - 1u;   // bucket of the padding items
    static_assert(E == 8 || E == 32, "tile kernels are instantiated for NP = 256 and 1024");
};
__device__ __forceinline__ int skew(int j) { return j + (j >> 5); }

template <int E>
constexpr size_t fit_tile_smem() { return (size_t)TILE_CT * TileGeom<E>::NPS * 4; }
template <int E>
constexpr size_t predict_tile_smem() { return (size_t)TILE_CT * (2 * TileGeom<E>::NPS * 4 + TileGeom<E>::NP * 2); }

// cooperative, coalesced load of one group's rows for the CTA's 8 cells into tile[cell][skew(j)]
template <int E>
__device__ __forceinline__ void load_tile(float* tile, const float* __restrict__ src, int64_t ld, int64_t C,
                                          int64_t c0, const int32_t* __restrict__ rg, int n,
                                          const uint8_t* __restrict__ valid) {
    constexpr int NPS = TileGeom<E>::NPS;
    const int cc = threadIdx.x & (TILE_CT - 1);
    const int64_t c = c0 + cc;
    const bool ok = c < C && (!valid || valid[c]);
    float* dst = tile + cc * NPS;
    const float* col = src + c;
#pragma unroll 4
    for (int j = threadIdx.x >> 3; j < n; j += TILE_THREADS / TILE_CT)
        dst[skew(j)] = ok ? __ldcs(col + (int64_t)rg[j] * ld) : 0.0f;
}

// ---------------------------------------------------------------- fit
template <int E>
__global__ void __launch_bounds__(TILE_THREADS)
qm_fit_tile_kernel(const float* __restrict__ y, int64_t ld, int64_t C,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ len,
                   const int64_t* __restrict__ off, int max_len,
                   float* __restrict__ state, int64_t state_ld, const uint8_t* __restrict__ valid,
                   int32_t* __restrict__ nonfinite) {
    using G = TileGeom<E>;
    extern __shared__ float tile_f[];
    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * TILE_CT;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    load_tile<E>(tile_f, y, ld, C, c0, rg, n, valid);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = c0 + warp;
    if (c >= C || (valid && !valid[c])) return;
    float* my = tile_f + warp * G::NPS;
    K32 v[E];
    bool bad = false;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        if (j < n) {
            const float x = my[skew(j)];
            bad |= !isfinite(x);
            v[e] = make_fit_item(x);
        } else {
            v[e] = sentinel_item<K32>(j);
        }
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    sort_blocked<K32, E, 32>(v, lane, nullptr);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        if (j < n) my[skew(j)] = fit_item_value(v[e]);
    }
    __syncwarp();
    float* dst = state + c * state_ld + off[g];
    for (int j = lane; j < n; j += 32) dst[j] = my[skew(j)];      // 128-byte coalesced rows of the cell record
}

// ---------------------------------------------------------------- predict helpers
// window sum / count of the centred 9-sample window of member j (members outside [0, n) absent)
__device__ __forceinline__ double window_key(const float* myX, int n, int j, double xc, double& shift) {
    double acc = 0.0;
    const int lo = j - 4 < 0 ? 0 : j - 4, hi = j + 4 > n - 1 ? n - 1 : j + 4;
    for (int jj = lo; jj <= hi; ++jj) acc += (double)myX[skew(jj)];
    shift = acc / (double)(hi - lo + 1) - xc;
    return (double)myX[skew(j)] - shift;
}

// Visit the E members owned by this lane: f(e, x, shift) with shift = rolling mean - xc in float64.
// The window sum slides (add the entering value, subtract the leaving one): exact for float32
// data of ordinary dynamic range, like pandas' own online add/remove kernel.
template <int E, bool EXACT, class F>
__device__ __forceinline__ void visit_shift(const float* myX, int n, int j0, double xc, F&& f) {
    float xh[E + 9];                                   // members j0-4 .. j0+E+4
#pragma unroll
    for (int i = 0; i < E + 9; ++i) {
        const int jj = j0 - 4 + i;
        xh[i] = (jj >= 0 && jj < n) ? myX[skew(jj)] : 0.0f;
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) sum += (double)xh[i];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = j0 + e;
        if (j < n) {
            const int lo = j - 4 < 0 ? 0 : j - 4, hi = j + 4 > n - 1 ? n - 1 : j + 4;
            const int cnt = hi - lo + 1;
            double roll;
            if (EXACT) roll = sum / (double)cnt;
            else       roll = sum * (1.0 / 9.0) * (9.0 / (double)cnt);      // bounds only
            f(e, (double)xh[e + 4], roll - xc);
        }
        sum += (double)xh[e + 9];
        sum -= (double)xh[e];
    }
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// run ends (1-based count) of equal-bucket runs over the first n sorted positions, blocked layout
template <int E, int LOG>
__device__ __forceinline__ void bucket_run_bounds(const K32 (&v)[E], int lane, int n, uint32_t nxt_first,
                                                  uint32_t prv_last, int (&run_end)[E], int (&run_start)[E]) {
    uint32_t bm_last = 0, bm_first = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int p = lane * E + e;
        const uint32_t q = v[e].k >> LOG;
        const uint32_t qn = ((e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k) >> LOG;
        const uint32_t qp = ((e == 0) ? prv_last : v[e == 0 ? e : e - 1].k) >> LOG;
        const bool last = (p >= n - 1) || (q != qn);
        const bool first = (p == 0) || (p >= n) || (q != qp);
        bm_last |= last ? (1u << e) : 0u;
        bm_first |= first ? (1u << e) : 0u;
    }
    if (E < 32) { bm_last &= (1u << E) - 1u; bm_first &= (1u << E) - 1u; }
    const int base = lane * E;
    // ends: nearest "last" at or after p
    const int minb = base + __ffs(bm_last);
    const uint32_t has_l = __ballot_sync(0xffffffffu, bm_last != 0);
    const uint32_t higher = (lane == 31) ? 0u : (has_l & ~((2u << lane) - 1u));
    const int carry_e = __shfl_sync(0xffffffffu, minb, higher ? (__ffs(higher) - 1) : lane);
    // starts: nearest "first" at or before p
    const int maxf = base + (31 - __clz(bm_first | 0u));           // valid when bm_first != 0
    const uint32_t has_f = __ballot_sync(0xffffffffu, bm_first != 0);
    const uint32_t lower = has_f & ((1u << lane) - 1u);
    const int carry_s = __shfl_sync(0xffffffffu, maxf, lower ? (31 - __clz(lower)) : lane);
    int cur = carry_e;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        if ((bm_last >> e) & 1u) cur = base + e + 1;
        run_end[e] = cur;
    }
    cur = carry_s;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((bm_first >> e) & 1u) cur = base + e;
        run_start[e] = cur;
    }
}

// ---------------------------------------------------------------- predict
template <int E, bool SHIFT>
__global__ void __launch_bounds__(TILE_THREADS, 2)
qm_predict_tile_kernel(const PredictParams p) {
    using G = TileGeom<E>;
    constexpr int NP = G::NP, NPS = G::NPS, LOG = G::LOG;
    constexpr uint32_t QMAX = G::QMAX;
    extern __shared__ uint32_t smem_u[];
    float* tileX = reinterpret_cast<float*>(smem_u);
    float* tileS = tileX + TILE_CT * NPS;
    uint16_t* rank_all = reinterpret_cast<uint16_t*>(tileS + TILE_CT * NPS);

    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * TILE_CT;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    const float* X = (const float*)p.X;
    load_tile<E>(tileX, X, p.ld, p.C, c0, rg, n, p.valid);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = c0 + warp;
    float* myX = tileX + warp * NPS;
    float* myS = tileS + warp * NPS;
    uint32_t* sw = reinterpret_cast<uint32_t*>(myS);
    uint16_t* myR = rank_all + warp * NP;
    const int j0 = lane * E;
    const bool in_range = c < p.C;
    const bool active = in_range && (!p.valid || p.valid[c]);

    if (in_range && !active) {
        for (int j = lane; j < n; j += 32) myX[skew(j)] = NAN;
    } else if (active) {
        const int sg = p.state_gid[g];
        const int m = p.fit_len[sg];
        const float* S = (const float*)p.state + c * p.state_ld + p.state_off[sg];
        double xc = 0.0, yc = 0.0;
        if (SHIFT) xc = (double)((const float*)p.x_climo)[(int64_t)sg * p.ld_climo + c];
        if (p.mode != SDB_MODE_QM && p.return_anoms) yc = (double)((const float*)p.y_climo)[(int64_t)sg * p.ld_climo + c];

        // ---- 1. monotone quantisation of the rank keys, packed with the member position
        K32 v[E];
        {
            float lo32 = INFINITY, hi32 = -INFINITY;
            bool bad = false;
            if (SHIFT) {
                visit_shift<E, false>(myX, n, j0, xc, [&](int, double x, double s) {
                    const double k = x - s;
                    lo32 = fminf(lo32, __double2float_rd(k));
                    hi32 = fmaxf(hi32, __double2float_ru(k));
                    bad |= !isfinite(x);
                });
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if (j0 + e < n) {
                        const float x = myX[skew(j0 + e)];
                        lo32 = fminf(lo32, x); hi32 = fmaxf(hi32, x);
                        bad |= !isfinite(x);
                    }
                }
            }
            if (bad && p.nonfinite) atomicOr(p.nonfinite, 1);
            lo32 = warp_min(lo32); hi32 = warp_max(hi32);
            if (SHIFT) {
                // the bounds came from an approximate rolling mean: widen them by a few float32 ulps
                const double lo = (double)lo32 - fabs((double)lo32) * 1e-6 - 1e-30;
                const double hi = (double)hi32 + fabs((double)hi32) * 1e-6 + 1e-30;
                const double scale = (hi > lo && isfinite(hi - lo)) ? (double)(QMAX - 1) / (hi - lo) : 0.0;
                visit_shift<E, true>(myX, n, j0, xc, [&](int e, double x, double s) {
                    const double t = ((x - s) - lo) * scale;
                    uint32_t q = (t > 0.0) ? __double2uint_rd(t) : 0u;      // also maps NaN to 0
                    q = q > QMAX - 1 ? QMAX - 1 : q;
                    v[e].k = (q << LOG) | (uint32_t)(j0 + e);
                });
            } else {
                const float range = hi32 - lo32;
                const float scale = (range > 0.0f && isfinite(range)) ? (float)(QMAX - 1) / range : 0.0f;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if (j0 + e < n) {
                        const float t = (myX[skew(j0 + e)] - lo32) * scale;
                        uint32_t q = (t > 0.0f) ? __float2uint_rd(t) : 0u;
                        q = q > QMAX - 1 ? QMAX - 1 : q;
                        v[e].k = (q << LOG) | (uint32_t)(j0 + e);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (j0 + e >= n) v[e].k = 0xffffffffu;
        }

        // ---- 2. sort, then 1-based tie-max ranks
        sort_blocked<K32, E, 32>(v, lane, nullptr);
        const uint32_t nxt_first = __shfl_down_sync(0xffffffffu, v[0].k, 1);
        const uint32_t prv_last = __shfl_up_sync(0xffffffffu, v[E - 1].k, 1);
        bool any_eq = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pos = j0 + e;
            const uint32_t kn = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
            any_eq |= (pos + 1 < n) && ((v[e].k >> LOG) == (kn >> LOG));
        }
        constexpr uint32_t IDX = (uint32_t)NP - 1u;
        if (!__any_sync(0xffffffffu, any_eq)) {
            // every bucket holds one member: sorted position = rank - 1
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (j0 + e < n) myR[v[e].k & IDX] = (uint16_t)(j0 + e + 1);
        } else {
            // exact key of member j (what the reference compares): the value itself, or x - shift in float64
            auto exact_key = [&](int j) -> double {
                if (SHIFT) { double s; return window_key(myX, n, j, xc, s); }
                return (double)(myX[skew(j)] + 0.0f);
            };
            int run_end[E], run_start[E];
            bucket_run_bounds<E, LOG>(v, lane, n, nxt_first, prv_last, run_end, run_start);
            // bad pair = neighbours in one bucket whose exact keys differ
            uint32_t bm_bad = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = j0 + e;
                const uint32_t kn = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
                if ((pos + 1 < n) && ((v[e].k >> LOG) == (kn >> LOG))) {
                    if (exact_key((int)(v[e].k & IDX)) != exact_key((int)(kn & IDX))) bm_bad |= 1u << e;
                }
            }
            if (!__any_sync(0xffffffffu, bm_bad != 0)) {
                // buckets with several members hold exact ties only: everyone takes the run end
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (j0 + e < n) myR[v[e].k & IDX] = (uint16_t)run_end[e];
            } else {
                // exclusive prefix count of bad pairs over sorted positions → B[pos] (stored in myR for now)
                const int mine = __popc(bm_bad);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int run = incl - mine;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    sw[skew(j0 + e)] = v[e].k;
                    myR[j0 + e] = (uint16_t)run;
                    run += (bm_bad >> e) & 1u;
                }
                __syncwarp();
                int r[E];
                bool need_fallback = false;
#pragma unroll
                for (int e = 0; e < E; ++e) {      // unrolled: v / run bounds must stay in registers
                    const int pos = j0 + e;
                    if (pos >= n) { r[e] = 0; continue; }
                    const int s = run_start[e], en = run_end[e];
                    const int L = en - s;
                    if (L == 1) { r[e] = pos + 1; continue; }
                    const int nbad = (int)myR[en - 1] - (int)myR[s];
                    if (nbad == 0) { r[e] = en; continue; }
                    if (L > TILE_LMAX) { need_fallback = true; r[e] = en; continue; }
                    const double kp = exact_key((int)(v[e].k & IDX));
                    int cnt = 0;
                    for (int q2 = s; q2 < en; ++q2) cnt += exact_key((int)(sw[skew(q2)] & IDX)) <= kp ? 1 : 0;
                    r[e] = s + cnt;
                }
                __syncwarp();
                if (!__any_sync(0xffffffffu, need_fallback)) {
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (j0 + e < n) myR[v[e].k & IDX] = (uint16_t)r[e];
                } else {
                    // a long bucket with distinct keys (e.g. an outlier squeezing the rest): exact 64-bit sort
                    K64I u[E];
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int j = j0 + e;
                        u[e] = (j < n) ? make_rank_item64(exact_key(j), (uint32_t)j) : sentinel_item<K64I>((uint32_t)j);
                    }
                    sort_blocked<K64I, E, 32>(u, lane, nullptr);
                    int r2[E];
                    tie_max_ranks<K64I, E, 32>(u, lane, r2, nullptr);
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (u[e].i < (uint32_t)n) myR[u[e].i] = (uint16_t)r2[e];
                }
            }
        }
        __syncwarp();

        // ---- 3. fitted sorted values of this (cell, group) → shared memory (coalesced)
        for (int j = lane; j < m; j += 32) myS[skew(j)] = S[j];
        __syncwarp();

        // ---- 4. rank → quantile → inverse CDF → output, in member order
        const double dn = pp_denominator(n), dm = pp_denominator(m);
        auto Sat = [&](int i) -> double { return (double)myS[skew(i)]; };
        float o[E];
        auto finish = [&](int e, double shift) {
            const int j = j0 + e;
            const int rk = (int)myR[j];
            const double val = inverse_cdf_acc(rk, n, m, Sat, dn, dm);
            double res;
            if (SHIFT) {
                res = shift + val;                               // bcsd.py:263
                if (p.return_anoms) res = res - yc;              // bcsd.py:267
            } else if (p.mode == SDB_MODE_BCSD_P) {
                res = p.return_anoms ? val / yc : val;           // bcsd.py:170-185
            } else {
                res = val;
            }
            o[e] = (float)res;
            if (p.rank_out) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = rk;
        };
        if (SHIFT) {
            visit_shift<E, true>(myX, n, j0, xc, [&](int e, double, double s) { finish(e, s); });
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (j0 + e < n) finish(e, 0.0);
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (j0 + e < n) myX[skew(j0 + e)] = o[e];
    }
    __syncthreads();
    // ---- 5. coalesced store of the tile (rows of 8 cells)
    {
        const int cc = threadIdx.x & (TILE_CT - 1);
        const int64_t cs = c0 + cc;
        if (cs < p.C) {
            float* outp = (float*)p.out + cs;
            const float* srcp = tileX + cc * NPS;
#pragma unroll 4
            for (int j = threadIdx.x >> 3; j < n; j += TILE_THREADS / TILE_CT)
                __stcs(outp + (int64_t)rg[j] * p.ld_out, srcp[skew(j)]);
        }
    }
}

// ---------------------------------------------------------------- launchers
template <int E>
static int launch_fit_tile(const FitParams& f, cudaStream_t st) {
    auto kern = qm_fit_tile_kernel<E>;
    const size_t smem = fit_tile_smem<E>();
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((f.C + TILE_CT - 1) / TILE_CT), (unsigned)f.n_groups);
    kern<<<grid, TILE_THREADS, smem, st>>>((const float*)f.y, f.ld, f.C, f.rows, f.len, f.off, f.max_len,
                                           (float*)f.state, f.state_ld, f.valid, f.nonfinite);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int E, bool SHIFT>
static int launch_predict_tile(const PredictParams& p, cudaStream_t st) {
    auto kern = qm_predict_tile_kernel<E, SHIFT>;
    const size_t smem = predict_tile_smem<E>();
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.C + TILE_CT - 1) / TILE_CT), (unsigned)p.n_groups);
    kern<<<grid, TILE_THREADS, smem, st>>>(p);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace sdb

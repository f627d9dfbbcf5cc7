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
constexpr size_t predict_tile_smem() { return (size_t)TILE_CT * 3 * TileGeom<E>::NPS * 4; }

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
// sum / cnt for the window counts 5..9, correctly rounded: q0 = sum*rc, one FMA residual
// correction (Markstein).  cnt == 9 everywhere except the first / last four members of a group.
__device__ __forceinline__ double div_count(double sum, int cnt) {
    if (cnt == 9) {
        const double rc = 1.0 / 9.0;
        const double q0 = sum * rc;
        const double r = fma(-q0, 9.0, sum);
        return fma(r, rc, q0);
    }
    return sum / (double)cnt;
}

// window bounds of member j inside a group of n
__device__ __forceinline__ int win_count(int j, int n) {
    const int lo = j - 4 < 0 ? 0 : j - 4, hi = j + 4 > n - 1 ? n - 1 : j + 4;
    return hi - lo + 1;
}

// exact rank key of member j as the reference computes it: x - (rolling mean - xc) in float64
__device__ __forceinline__ double window_key(const float* myX, int n, int j, double xc) {
    double acc = 0.0;
    const int lo = j - 4 < 0 ? 0 : j - 4, hi = j + 4 > n - 1 ? n - 1 : j + 4;
    for (int jj = lo; jj <= hi; ++jj) acc += (double)myX[skew(jj)];
    return (double)myX[skew(j)] - (div_count(acc, hi - lo + 1) - xc);
}
template <bool SHIFT>
__device__ __forceinline__ double exact_key(const float* myX, int n, int j, double xc) {
    if (SHIFT) return window_key(myX, n, j, xc);
    return (double)(myX[skew(j)] + 0.0f);
}

// Sliding 9-sample window over the members [j0, j0+cnt) owned by one lane, values read from the
// shared-memory row: f(j, x_j, window_sum_j).  The sum slides (add the entering member, subtract
// the leaving one) — exact for float32 data of ordinary dynamic range, like pandas' own online
// add/remove kernel (pandas/_libs/window/aggregations.pyx roll_mean).
template <class F>
__device__ __forceinline__ void slide_window(const float* myX, int n, int j0, int j1, F&& f) {
    if (j0 >= j1) return;
    double sum = 0.0;
    for (int jj = (j0 - 4 < 0 ? 0 : j0 - 4); jj <= (j0 + 4 > n - 1 ? n - 1 : j0 + 4); ++jj) sum += (double)myX[skew(jj)];
#pragma unroll 2
    for (int j = j0; j < j1; ++j) {
        f(j, myX[skew(j)], sum);
        if (j + 5 < n) sum += (double)myX[skew(j + 5)];
        if (j - 4 >= 0) sum -= (double)myX[skew(j - 4)];
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

// run [start, end) of equal-bucket sorted positions around every position, packed
// start | end << 10 (blocked layout, first n positions real).  O(1) per position: boundary
// bitmaps per lane + one ballot each way.
template <int E, int LOG>
__device__ __forceinline__ void bucket_run_bounds(const K32 (&v)[E], int lane, int n, uint32_t nxt_first,
                                                  uint32_t prv_last, uint32_t (&packed)[E]) {
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
    const int base = lane * E;
    const int minb = base + __ffs(bm_last);                          // nearest run end at/after the lane start
    const uint32_t has_l = __ballot_sync(0xffffffffu, bm_last != 0);
    const uint32_t higher = (lane == 31) ? 0u : (has_l & ~((2u << lane) - 1u));
    const int carry_e = __shfl_sync(0xffffffffu, minb, higher ? (__ffs(higher) - 1) : lane);
    const int maxf = base + (31 - __clz(bm_first | 1u));             // nearest run start at/before the lane end
    const uint32_t has_f = __ballot_sync(0xffffffffu, bm_first != 0);
    const uint32_t lower = has_f & ((1u << lane) - 1u);
    const int carry_s = __shfl_sync(0xffffffffu, maxf, lower ? (31 - __clz(lower)) : lane);
    int cur = carry_e;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        if ((bm_last >> e) & 1u) cur = base + e + 1;
        packed[e] = (uint32_t)cur << 10;
    }
    cur = carry_s;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((bm_first >> e) & 1u) cur = base + e;
        packed[e] |= (uint32_t)cur;
    }
}

// Out-of-line exact ranking of one (cell, group): 64-bit key + position sort (the generic
// algorithm) for the rare group whose keys defeat the bucket quantisation (a long bucket holding
// distinct keys).  Writes the 1-based tie-max rank of member j to R[skew(j)].
template <int E, bool SHIFT>
__device__ __noinline__ void rank_exact64(const float* myX, int n, double xc, int lane, uint32_t* R) {
    K64I u[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        u[e] = (j < n) ? make_rank_item64(exact_key<SHIFT>(myX, n, j, xc), (uint32_t)j) : sentinel_item<K64I>((uint32_t)j);
    }
    sort_blocked<K64I, E, 32>(u, lane, nullptr);
    int r2[E];
    tie_max_ranks<K64I, E, 32>(u, lane, r2, nullptr);
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (u[e].i < (uint32_t)n) R[skew((int)u[e].i)] = (uint32_t)r2[e];
}

// Exact ranks when some bucket holds several members.  sw[] holds the sorted packed words
// (bucket << LOG | member), R[] receives start | end << 10 | B << 21 per sorted position and,
// at the end, the rank of member j at R[skew(j)].  Runtime loops over shared memory only.
template <int E, int LOG, bool SHIFT>
__device__ __noinline__ bool rank_with_ties(const float* myX, uint32_t* sw, uint32_t* R, int n, double xc, int lane) {
    constexpr uint32_t IDX = (1u << LOG) - 1u;
    const int p0 = lane * E, p1 = (p0 + E < n) ? p0 + E : n;
    // bad pair (pos, pos+1): same bucket, different exact keys
    uint32_t bm_bad = 0;
    for (int pos = p0; pos < p1; ++pos) {
        const uint32_t pk = R[skew(pos)];
        const int en = (int)((pk >> 10) & 0x7ffu);
        if (pos + 1 < en) {
            const double a = exact_key<SHIFT>(myX, n, (int)(sw[skew(pos)] & IDX), xc);
            const double b = exact_key<SHIFT>(myX, n, (int)(sw[skew(pos + 1)] & IDX), xc);
            if (a != b) bm_bad |= 1u << (pos - p0);
        }
    }
    if (!__any_sync(0xffffffffu, bm_bad != 0)) {
        // multi-member buckets hold exact ties only: every member takes the end of its run
        __syncwarp();
        uint32_t rk[1];
        (void)rk;
        // ranks are scattered by member position; the run table is still needed by other lanes
        // only through R[pos] of THEIR positions, so stage through sw (idx | rank << LOG)
        for (int pos = p0; pos < p1; ++pos) {
            const uint32_t en = (R[skew(pos)] >> 10) & 0x7ffu;
            sw[skew(pos)] = (sw[skew(pos)] & IDX) | (en << LOG);
        }
        __syncwarp();
        for (int pos = p0; pos < p1; ++pos) {
            const uint32_t t = sw[skew(pos)];
            R[skew((int)(t & IDX))] = t >> LOG;
        }
        return true;
    }
    // exclusive prefix count of bad pairs over sorted positions → B, packed into R[pos] bits 21..31
    const int mine = __popc(bm_bad);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int before = incl - mine;
    for (int pos = p0; pos < p1; ++pos) {
        const uint32_t B = (uint32_t)(before + __popc(bm_bad & ((1u << (pos - p0)) - 1u)));
        R[skew(pos)] |= B << 21;
    }
    __syncwarp();
    bool need_fallback = false;
    for (int pos = p0; pos < p1; ++pos) {
        const uint32_t pk = R[skew(pos)];
        const int s = (int)(pk & 0x3ffu), en = (int)((pk >> 10) & 0x7ffu);
        const int L = en - s;
        int r;
        if (L == 1) r = pos + 1;
        else {
            const int nbad = (int)(R[skew(en - 1)] >> 21) - (int)(R[skew(s)] >> 21);
            if (nbad == 0) r = en;
            else if (L > TILE_LMAX) { need_fallback = true; r = en; }
            else {
                const double kp = exact_key<SHIFT>(myX, n, (int)(sw[skew(pos)] & IDX), xc);
                int cnt = 0;
                for (int q2 = s; q2 < en; ++q2)
                    cnt += exact_key<SHIFT>(myX, n, (int)(sw[skew(q2)] & IDX), xc) <= kp ? 1 : 0;
                r = s + cnt;
            }
        }
        // members keep their identity in the low bits, so other lanes may still read sw[pos] & IDX
        sw[skew(pos)] = (sw[skew(pos)] & IDX) | ((uint32_t)r << LOG);
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, need_fallback)) return false;
    for (int pos = p0; pos < p1; ++pos) {
        const uint32_t t = sw[skew(pos)];
        R[skew((int)(t & IDX))] = t >> LOG;
    }
    return true;
}

// ---------------------------------------------------------------- predict
template <int E, bool SHIFT>
__global__ void __launch_bounds__(TILE_THREADS, 2)
qm_predict_tile_kernel(const PredictParams p) {
    using G = TileGeom<E>;
    constexpr int NPS = G::NPS, LOG = G::LOG;
    constexpr uint32_t QMAX = G::QMAX;
    constexpr uint32_t IDX = (1u << LOG) - 1u;
    extern __shared__ uint32_t smem_u[];
    float* tileX = reinterpret_cast<float*>(smem_u);                  // inputs of the group
    float* tileS = tileX + TILE_CT * NPS;                             // (slow path) sorted words → fitted values
    uint32_t* tileR = reinterpret_cast<uint32_t*>(tileS + TILE_CT * NPS);   // (run table →) ranks/values → outputs

    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * TILE_CT;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    load_tile<E>(tileX, (const float*)p.X, p.ld, p.C, c0, rg, n, p.valid);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = c0 + warp;
    float* myX = tileX + warp * NPS;
    float* myS = tileS + warp * NPS;
    uint32_t* sw = reinterpret_cast<uint32_t*>(myS);
    uint32_t* R = tileR + warp * NPS;
    const int j0 = lane * E;
    const int j1 = (j0 + E < n) ? j0 + E : n;
    const bool in_range = c < p.C;
    const bool active = in_range && (!p.valid || p.valid[c]);

    if (in_range && !active) {
        for (int j = lane; j < n; j += 32) R[skew(j)] = __float_as_uint(NAN);
    } else if (active) {
        const int sg = p.state_gid[g];
        const int m = p.fit_len[sg];
        const float* S = (const float*)p.state + c * p.state_ld + p.state_off[sg];
        double xc = 0.0, yc = 0.0;
        if (SHIFT) xc = (double)((const float*)p.x_climo)[(int64_t)sg * p.ld_climo + c];
        if (p.mode != SDB_MODE_QM && p.return_anoms) yc = (double)((const float*)p.y_climo)[(int64_t)sg * p.ld_climo + c];
        const bool ratio = (p.mode == SDB_MODE_BCSD_P) && p.return_anoms;

        // ---- 1. own members (+ 4 / 5 halo) to registers; key bounds from the plain value range
        constexpr int HL = SHIFT ? 4 : 0, HR = SHIFT ? 5 : 0;
        float xh[E + HL + HR];
        {
            const int rb = skew(j0);
#pragma unroll
            for (int i = 0; i < E + HL + HR; ++i) {
                const int e = i - HL;                          // member offset inside / around the lane's block
                const int jj = j0 + e;
                // E == 32: the skew step only changes at the block edges → static offsets from the row base
                const int addr = (E == 32) ? rb + e + (e < 0 ? -1 : (e >= 32 ? 1 : 0)) : skew(jj < 0 ? 0 : jj);
                xh[i] = (jj >= 0 && jj < n) ? myX[addr] : 0.0f;
            }
        }
        float lo32 = INFINITY, hi32 = -INFINITY;
        bool bad = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (j0 + e < n) {
                const float x = xh[e + HL];
                lo32 = fminf(lo32, x); hi32 = fmaxf(hi32, x);
                bad |= !isfinite(x);
            }
        }
        if (bad && p.nonfinite) atomicOr(p.nonfinite, 1);
        lo32 = warp_min(lo32); hi32 = warp_max(hi32);

        // ---- 2. exact rank keys → monotone bucket number, packed with the member position.
        // Bounds are a guess (value range + 1/8 margin, centred on the climatology for the shifted
        // key); out-of-range keys clamp to the end buckets, which keeps the map monotone — the
        // exact fix-up below sorts out whatever shares a bucket.
        K32 v[E];
        float sh[SHIFT ? E : 1];
        if (SHIFT) {
            const double range = (double)hi32 - (double)lo32;
            const double lo = (double)lo32 - 0.125 * range, hi = (double)hi32 + 0.125 * range;
            const double scale = (hi > lo && isfinite(hi - lo)) ? (double)(QMAX - 1) / (hi - lo) : 0.0;
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) sum += (double)xh[i];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int j = j0 + e;
                if (j < n) {
                    const double shift = div_count(sum, win_count(j, n)) - xc;
                    const double t = (((double)xh[e + 4] - shift) - lo) * scale;
                    uint32_t q = (t > 0.0) ? __double2uint_rd(t) : 0u;           // also maps NaN to 0
                    q = q > QMAX - 1 ? QMAX - 1 : q;
                    v[e].k = (q << LOG) | (uint32_t)j;
                    sh[e] = (float)shift;
                } else {
                    v[e].k = 0xffffffffu;
                    sh[e] = 0.0f;
                }
                sum += (double)xh[e + 9];
                sum -= (double)xh[e];
            }
        } else {
            const float range = hi32 - lo32;
            const float scale = (range > 0.0f && isfinite(range)) ? (float)(QMAX - 1) / range : 0.0f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int j = j0 + e;
                if (j < n) {
                    const float t = (xh[e] - lo32) * scale;
                    uint32_t q = (t > 0.0f) ? __float2uint_rd(t) : 0u;
                    q = q > QMAX - 1 ? QMAX - 1 : q;
                    v[e].k = (q << LOG) | (uint32_t)j;
                } else {
                    v[e].k = 0xffffffffu;
                }
            }
        }

        // ---- 3. one 32-bit keys-only sort
        sort_blocked<K32, E, 32>(v, lane, nullptr);
        const uint32_t nxt_first = __shfl_down_sync(0xffffffffu, v[0].k, 1);
        bool any_eq = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const uint32_t kn = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
            any_eq |= (j0 + e + 1 < n) && ((v[e].k >> LOG) == (kn >> LOG));
        }
        const bool same = (n == m);
        bool ranks_in_R = false;        // false: R[member] already holds the mapped value (float bits)
        auto park = [&](float val) -> uint32_t { return __float_as_uint(ratio ? (float)((double)val / yc) : val); };

        if (!__any_sync(0xffffffffu, any_eq)) {
            if (same) {
                // one member per bucket, same length: the member at sorted position pos takes the
                // fitted order statistic S[pos] — read straight from the state record
                const bool vec = ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
                if (vec && E % 4 == 0) {
#pragma unroll
                    for (int e = 0; e < E; e += 4) {
                        if (j0 + e + 3 < n) {
                            const float4 s4 = *reinterpret_cast<const float4*>(S + j0 + e);
                            R[skew((int)(v[e].k & IDX))] = park(s4.x);
                            R[skew((int)(v[e + 1].k & IDX))] = park(s4.y);
                            R[skew((int)(v[e + 2].k & IDX))] = park(s4.z);
                            R[skew((int)(v[e + 3].k & IDX))] = park(s4.w);
                        } else {
#pragma unroll
                            for (int d = 0; d < 4; ++d)
                                if (j0 + e + d < n) R[skew((int)(v[e + d].k & IDX))] = park(S[j0 + e + d]);
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (j0 + e < n) R[skew((int)(v[e].k & IDX))] = park(S[j0 + e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (j0 + e < n) R[skew((int)(v[e].k & IDX))] = (uint32_t)(j0 + e + 1);
                ranks_in_R = true;
            }
        } else {
            const uint32_t prv_last = __shfl_up_sync(0xffffffffu, v[E - 1].k, 1);
            uint32_t packed[E];
            bucket_run_bounds<E, LOG>(v, lane, n, nxt_first, prv_last, packed);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                sw[skew(j0 + e)] = v[e].k;
                R[skew(j0 + e)] = packed[e];
            }
            __syncwarp();
            if (!rank_with_ties<E, LOG, SHIFT>(myX, sw, R, n, xc, lane)) {
                __syncwarp();
                rank_exact64<E, SHIFT>(myX, n, xc, lane, R);
            }
            ranks_in_R = true;
        }
        __syncwarp();

        if (ranks_in_R) {
            // ---- 4. (ties, or T_pred != T_fit) rank → quantile → inverse CDF of the fitted values
            for (int j = lane; j < m; j += 32) myS[skew(j)] = S[j];
            __syncwarp();
            const double dn = pp_denominator(n), dm = pp_denominator(m);
            auto Sat = [&](int i) -> double { return (double)myS[skew(i)]; };
            for (int j = j0; j < j1; ++j) {
                const int rk = (int)R[skew(j)];
                if (p.rank_out) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = rk;
                R[skew(j)] = park((float)inverse_cdf_acc(rk, n, m, Sat, dn, dm));
            }
            __syncwarp();
        } else if (p.rank_out) {
#pragma unroll
            for (int e = 0; e < E; ++e)      // instrumentation only: rank of the member at sorted position pos
                if (j0 + e < n) p.rank_out[(int64_t)rg[v[e].k & IDX] * p.ld_out + c] = j0 + e + 1;
        }

        // ---- 5. restore the shift (and remove the target climatology) in member order
        if (SHIFT) {
            uint32_t* rowR = R + skew(j0);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (j0 + e < n) {
                    double res = (double)sh[e] + (double)__uint_as_float(rowR[e]);    // bcsd.py:263
                    if (p.return_anoms) res = res - yc;                              // bcsd.py:267
                    rowR[e] = __float_as_uint((float)res);
                }
            }
        }
    }
    __syncthreads();
    // ---- 6. coalesced store of the tile (rows of 8 cells)
    {
        const int cc = threadIdx.x & (TILE_CT - 1);
        const int64_t cs = c0 + cc;
        if (cs < p.C) {
            float* outp = (float*)p.out + cs;
            const float* srcp = reinterpret_cast<const float*>(tileR) + cc * NPS;
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

// qm_long.cu — quantile mapping of LONG groups (1025 .. 16 384 steps, float32 data, raw keys): the
// whole-series QuantileMapper (quantile.py:81-147; BASELINE config "QuantileMapper 10 000 cells x 10 950
// days") and BcsdPrecipitation with long groups.  One CTA of 512 threads owns one (cell, group).
//
// Round 1 ran these on the generic kernel: a 16 384-point bitonic network on 64-bit (key, position) items,
// 105 compare-exchange stages, the wide ones through shared memory — 17.2 ms for 10 000 x 10 950 (1.2 % of
// the HBM roofline).  A comparison network costs n log^2 n; this file ranks by COUNTING in O(n), the
// block-wide form of bm_rank.cuh:
//   A  every element sets the presence bit of its (monotone) bucket in a shared-memory table of 9 216
//      64-bit entries x 32 buckets with one atomicOr; a second arrival in a bucket bumps the entry's meta;
//   B  a block-wide prefix over the entries (thread t owns 18 consecutive entries) turns meta into "elements
//      below this entry"; entries with a collision are dirty;
//   C  clean entry: position = prefix + popc(bits below) — exact (monotone map, one element per bit);
//      dirty entry: its members take the positions prefix + arrival number;
//   D  the thread that owns a dirty entry orders its (few) members by exact comparison.
// fit writes value → state[position]; predict maps rank = position + 1 through the fitted CDF.  Exact for
// every input (ties → highest rank); degenerate inputs (thousands of equal values away from the minimum)
// only cost time in step D.
#include "qm_kernels.cuh"
#include "bm_rank.cuh"

namespace sdb {

constexpr int LG_NT = 512;                     // threads per CTA of the quantile-mapping kernels (n <= 16 384)
constexpr int LG_NT_BIG = 1024;                // ... of the argsort kernel (n <= 32 768: the 50-year analog windows)
constexpr int LG_EPT = 18;                     // entries per thread: stride 36 words keeps 16-byte accesses conflict-free
constexpr int LG_E = 32;                       // elements per thread, member j = tid + k * NT
template <int NT> struct LgCfg {
    static constexpr int ENT = NT * LG_EPT;    // 512 threads: 9 216 entries = 294 912 buckets (72 KB)
    static constexpr int NB = ENT * 32;
    static constexpr int NMAX = NT * LG_E;
    static constexpr int W_WORDS = 2 * (ENT + 32);   // + one dummy entry per lane
    static constexpr int NWARP = NT / 32;
};
constexpr int LG_ENT = LgCfg<LG_NT>::ENT;
constexpr int LG_NB = LgCfg<LG_NT>::NB;
constexpr int LG_NMAX = LgCfg<LG_NT>::NMAX;
constexpr int LG_W_WORDS = LgCfg<LG_NT>::W_WORDS;
constexpr int LG_SMALL = 8;                    // members of a dirty entry ordered in registers
constexpr int LG_DECAP = 2048;                 // list of dirty entries (uint16 entry numbers); more: every thread walks its own entries

struct LgShared {
    float red_lo[32], red_hi[32];
    uint32_t warp_tot[32];
    uint32_t n_dirty;
    uint32_t n_lo_cursor;
    uint16_t de[LG_DECAP];
};

template <int NT>
__device__ __forceinline__ void lg_block_minmax(float& lo, float& hi, LgShared* sh) {
    bm_warp_minmax(lo, hi);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh->red_lo[warp] = lo; sh->red_hi[warp] = hi; }
    __syncthreads();
    lo = sh->red_lo[lane % LgCfg<NT>::NWARP]; hi = sh->red_hi[lane % LgCfg<NT>::NWARP];
    bm_warp_minmax(lo, hi);
}

// phases A + B for the block: v[k] = key of member tid + k * 512 (k < LG_E), `n` members.  Returns the number of
// inserted elements (keys above the lower bound); afterwards W holds the prefixes.
template <int NT>
__device__ __forceinline__ int lg_build(const float (&v)[LG_E], int n, float lo, float scale, uint32_t* W, LgShared* sh) {
    using G = LgCfg<NT>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t Wsa = bm_saddr(W);
    {
        uint4* p = reinterpret_cast<uint4*>(W);
        for (int i = tid; i < G::W_WORDS / 4; i += NT) p[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) { sh->n_dirty = 0u; sh->n_lo_cursor = 0u; }
    }
    __syncthreads();
    const uint32_t dummy = (uint32_t)(G::ENT + lane) * 8u;
#pragma unroll
    for (int b = 0; b < LG_E; b += 8) {
        if (b * NT < n) {                         // CTA-uniform: batches past the end of the group are skipped
            uint32_t q[8];
            int on[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                q[u] = bm_bucket_f32<G::NB>(v[b + u], lo, scale);
                on[u] = (v[b + u] > lo) ? n - (tid + (b + u) * NT) : 0;
            }
            bm_insert_batch<8>(Wsa, dummy, q, on);
        }
    }
    __syncthreads();
    uint4* base = reinterpret_cast<uint4*>(W + tid * (2 * LG_EPT));
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < LG_EPT / 2; ++i) {
        const uint4 t = base[i];
        run += __popc(t.x) + t.y + __popc(t.z) + t.w;
    }
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh->warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wt = (lane < G::NWARP) ? sh->warp_tot[lane] : 0u, winc = wt;
#pragma unroll
    for (int o = 1; o < G::NWARP; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, winc, G::NWARP - 1);
    const uint32_t wbase = __shfl_sync(0xffffffffu, winc - wt, warp);
    uint32_t p = wbase + incl - run;
#pragma unroll
    for (int i = 0; i < LG_EPT / 2; ++i) {
        uint4 t = base[i];
        const uint32_t c0 = __popc(t.x) + t.y, c1 = __popc(t.z) + t.w;
        if (t.y | t.w) {                             // rare: note the dirty entries
            if (t.y) { const uint32_t k = atomicAdd(&sh->n_dirty, 1u); if (k < LG_DECAP) sh->de[k] = (uint16_t)(tid * LG_EPT + 2 * i); }
            if (t.w) { const uint32_t k = atomicAdd(&sh->n_dirty, 1u); if (k < LG_DECAP) sh->de[k] = (uint16_t)(tid * LG_EPT + 2 * i + 1); }
        }
        t.y = p | (t.y ? 0x80000000u : 0u);          // dirty: bits 16-30 = arrival cursor (starts at 0)
        p += c0;
        t.w = p | (t.w ? 0x80000000u : 0u);
        p += c1;
        base[i] = t;
    }
    __syncthreads();
    return (int)total;
}

// Step D.  Every thread walks the 18 entries it owns; a dirty entry holds its c members at positions
// [start, start + c) in arrival order.  c <= 8: the owning thread orders them in registers.  Larger entries are
// handled by the whole warp, one after the other: if all members are EQUAL (the usual way to get there: many
// exact ties — values on a coarse grid) nothing has to be ordered; otherwise the slow exact path runs.
template <int NT, class Small, class Big>
__device__ __forceinline__ void lg_for_dirty_entries(const uint32_t* W, const LgShared* sh, Small&& small, Big&& big) {
    const int tid = threadIdx.x;
    const int nd = (int)sh->n_dirty;
    // one thread per LISTED dirty entry (dense: a few hundred entries = one pass); if the list overflowed,
    // every thread walks the 18 entries it owns instead
    const bool listed = nd <= LG_DECAP;
    const int trips = listed ? (nd + NT - 1) / NT : LG_EPT;
#pragma unroll 1
    for (int i = 0; i < trips; ++i) {
        int e = -1;
        if (listed) { const int k = i * NT + tid; if (k < nd) e = (int)sh->de[k]; }
        else e = tid * LG_EPT + i;
        const uint32_t meta = (e >= 0) ? W[2 * e + 1] : 0u;
        const int c = ((int32_t)meta < 0) ? (int)((meta >> 16) & 0x7fffu) : 0;
        const int start = (int)(meta & 0xffffu);
        if (c > 1 && c <= LG_SMALL) small(start, c);
        uint32_t bigm = __ballot_sync(0xffffffffu, c > LG_SMALL);
        while (bigm) {
            const int src = __ffs(bigm) - 1;
            bigm &= bigm - 1u;
            big(__shfl_sync(0xffffffffu, start, src), __shfl_sync(0xffffffffu, c, src));
        }
    }
}

// ---------------------------------------------------------------- fit: np.sort of every (cell, group)
__global__ void __launch_bounds__(LG_NT, 2)
qm_fit_long_kernel(const float* __restrict__ y, int64_t ld, int64_t C,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ len,
                   const int64_t* __restrict__ off, int max_len,
                   float* __restrict__ state, int64_t state_ld, const uint8_t* __restrict__ valid,
                   int32_t* __restrict__ nonfinite) {
    extern __shared__ __align__(16) uint32_t lg_smem[];
    uint32_t* W = lg_smem;
    LgShared* sh = reinterpret_cast<LgShared*>(lg_smem + LG_W_WORDS);
    const int64_t c = blockIdx.x;
    const int g = blockIdx.y;
    if (valid && !valid[c]) return;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    const int tid = threadIdx.x;
    float v[LG_E];
    float lo = INFINITY, hi = -INFINITY;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < LG_E; ++k) {
        const int j = tid + k * LG_NT;
        v[k] = 0.0f;
        if (j < n) {
            const float x = y[(int64_t)rg[j] * ld + c];
            bad |= !isfinite(x);
            v[k] = x + 0.0f;
            lo = fminf(lo, v[k]); hi = fmaxf(hi, v[k]);
        }
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    lg_block_minmax<LG_NT>(lo, hi, sh);
    const float scale = bm_scale_f32<LG_NB>(lo, hi);
    const int total = lg_build<LG_NT>(v, n, lo, scale, W, sh);
    const int n_lo = n - total;
    float* dst = state + c * state_ld + off[g] + n_lo;       // the values equal to the lower bound come first
    const uint32_t Wsa = bm_saddr(W);
#pragma unroll
    for (int b = 0; b < LG_E; b += 8) {
        if (b * LG_NT >= n) continue;                        // CTA-uniform
        uint32_t q[8], dirty[8];
        int on[8], pos[8], cur[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            q[u] = bm_bucket_f32<LG_NB>(v[b + u], lo, scale);
            on[u] = (v[b + u] > lo) ? n - (tid + (b + u) * LG_NT) : 0;
        }
        bm_lookup_batch<8>(W, Wsa, q, on, pos, dirty, cur);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (on[u] > 0) dst[pos[u] + cur[u]] = v[b + u];      // cur = 0 for clean entries
    }
    __syncthreads();                                         // the record (global memory) is visible to the whole CTA
    const int lane = tid & 31;
    lg_for_dirty_entries<LG_NT>(W, sh,
        [&](int start, int cnt) {
            float k[LG_SMALL];
#pragma unroll
            for (int a = 0; a < LG_SMALL; ++a) k[a] = (a < cnt) ? __ldcg(dst + start + a) : INFINITY;
            bm_sort8(k);
#pragma unroll
            for (int a = 0; a < LG_SMALL; ++a)
                if (a < cnt) dst[start + a] = k[a];
        },
        [&](int start, int cnt) {
            if (cnt <= 32) {
                // one member per lane, ranks by an all-pairs pass over shuffles (no memory traffic, no in-place hazard)
                const float mine = (lane < cnt) ? __ldcg(dst + start + lane) : INFINITY;
                int r = 0;
                for (int b = 0; b < cnt; ++b) {
                    const float o = __shfl_sync(0xffffffffu, mine, b);
                    r += (o < mine || (o == mine && b < lane)) ? 1 : 0;
                }
                __syncwarp();
                if (lane < cnt) dst[start + r] = mine;
                return;
            }
            float mn = INFINITY, mx = -INFINITY;
            for (int a = lane; a < cnt; a += 32) { const float k = __ldcg(dst + start + a); mn = fminf(mn, k); mx = fmaxf(mx, k); }
            bm_warp_minmax(mn, mx);
            if (mn == mx) return;                            // all equal: already in order
            if (lane == 0) {                                 // slow exact path: in-place insertion sort
                for (int a = 1; a < cnt; ++a) {
                    const float k = __ldcg(dst + start + a);
                    int b = a - 1;
                    while (b >= 0 && __ldcg(dst + start + b) > k) { dst[start + b + 1] = __ldcg(dst + start + b); --b; }
                    dst[start + b + 1] = k;
                }
            }
            __syncwarp();
        });
    float* rec = state + c * state_ld + off[g];
    for (int i = tid; i < n_lo; i += LG_NT) rec[i] = lo;
}

// ---------------------------------------------------------------- predict: self-rank → fitted CDF → output
__global__ void __launch_bounds__(LG_NT, 2)
qm_predict_long_kernel(const PredictParams p) {
    extern __shared__ __align__(16) uint32_t lg_smem[];
    uint32_t* W = lg_smem;
    LgShared* sh = reinterpret_cast<LgShared*>(lg_smem + LG_W_WORDS);
    uint16_t* P2M = reinterpret_cast<uint16_t*>(lg_smem + LG_W_WORDS + (sizeof(LgShared) + 3) / 4);   // position → member
    const int64_t c = blockIdx.x;
    const int g = blockIdx.y;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    const int tid = threadIdx.x;
    if (p.valid && !p.valid[c]) {
        for (int j = tid; j < n; j += LG_NT) {
            store_out(p.out, p.out_f64, (int64_t)rg[j] * p.ld_out + c, (double)NAN);
            if (p.rank_out) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = 0;
        }
        return;
    }
    const float* X = (const float*)p.X;
    const int sg = p.state_gid[g];
    const int m = p.fit_len[sg];
    const float* S = (const float*)p.state + c * p.state_ld + p.state_off[sg];
    const bool ratio = (p.mode == SDB_MODE_BCSD_P) && p.return_anoms;
    const double yc = ratio ? (double)((const float*)p.y_climo)[(int64_t)sg * p.ld_climo + c] : 1.0;
    const Cunnane cu = cunnane_of(p);
    const double dn = pp_denominator(n, cu), dm = pp_denominator(m, cu);
    float v[LG_E];
    float lo = INFINITY, hi = -INFINITY;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < LG_E; ++k) {
        const int j = tid + k * LG_NT;
        v[k] = 0.0f;
        if (j < n) {
            const float x = X[(int64_t)rg[j] * p.ld + c];
            bad |= !isfinite(x);
            v[k] = x + 0.0f;
            lo = fminf(lo, v[k]); hi = fmaxf(hi, v[k]);
        }
    }
    if (bad && p.nonfinite) atomicOr(p.nonfinite, 1);
    lg_block_minmax<LG_NT>(lo, hi, sh);
    const float scale = bm_scale_f32<LG_NB>(lo, hi);
    const int total = lg_build<LG_NT>(v, n, lo, scale, W, sh);
    const int n_lo = n - total;
    auto emit = [&](int j, int rank) {
        double val = inverse_cdf<float>(rank, n, m, S, dn, dm, cu);
        if (ratio) val = val / yc;
        const int64_t at = (int64_t)rg[j] * p.ld_out + c;
        store_out(p.out, p.out_f64, at, val);
        if (p.rank_out) p.rank_out[at] = rank;
    };
    const uint32_t Wsa = bm_saddr(W);
#pragma unroll
    for (int b = 0; b < LG_E; b += 8) {
        if (b * LG_NT >= n) continue;                        // CTA-uniform
        uint32_t q[8], dirty[8];
        int on[8], pos[8], cur[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            q[u] = bm_bucket_f32<LG_NB>(v[b + u], lo, scale);
            on[u] = (v[b + u] > lo) ? n - (tid + (b + u) * LG_NT) : 0;
        }
        bm_lookup_batch<8>(W, Wsa, q, on, pos, dirty, cur);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = tid + (b + u) * LG_NT;
            if (j < n) {
                if (on[u] <= 0) emit(j, n_lo);                                   // ties at the lower bound: highest rank
                else if (!dirty[u]) emit(j, n_lo + pos[u] + 1);
                else P2M[pos[u] + cur[u]] = (uint16_t)j;
            }
        }
    }
    __syncthreads();
    const int lane = tid & 31;
    auto key_at = [&](int at) -> float { return __ldg(X + (int64_t)rg[P2M[at]] * p.ld + c) + 0.0f; };
    lg_for_dirty_entries<LG_NT>(W, sh,
        [&](int start, int cnt) {
            float k[LG_SMALL];
#pragma unroll
            for (int a = 0; a < LG_SMALL; ++a) k[a] = (a < cnt) ? key_at(start + a) : INFINITY;
#pragma unroll
            for (int a = 0; a < LG_SMALL; ++a) {
                int r = 0;
#pragma unroll
                for (int b = 0; b < LG_SMALL; ++b)
                    if (b != a) r += (b < cnt && k[b] <= k[a]) ? 1 : 0;          // ties → highest rank
                if (a < cnt) emit((int)P2M[start + a], n_lo + start + r + 1);
            }
        },
        [&](int start, int cnt) {
            if (cnt <= 32) {                                 // one member per lane, all-pairs over shuffles
                const float mine = (lane < cnt) ? key_at(start + lane) : INFINITY;
                int r = -1;                                  // the member counts itself below
                for (int b = 0; b < cnt; ++b) r += (__shfl_sync(0xffffffffu, mine, b) <= mine) ? 1 : 0;
                if (lane < cnt) emit((int)P2M[start + lane], n_lo + start + r + 1);      // ties → highest rank
                return;
            }
            float mn = INFINITY, mx = -INFINITY;
            for (int a = lane; a < cnt; a += 32) { const float k = key_at(start + a); mn = fminf(mn, k); mx = fmaxf(mx, k); }
            bm_warp_minmax(mn, mx);
            for (int a = lane; a < cnt; a += 32) {
                int r = cnt - 1;                             // all equal: every member takes the highest rank
                if (mn != mx) {                              // slow exact path
                    const float ka = key_at(start + a);
                    r = 0;
                    for (int b = 0; b < cnt; ++b) r += (b != a && key_at(start + b) <= ka) ? 1 : 0;
                }
                emit((int)P2M[start + a], n_lo + start + r + 1);
            }
        });
}

// ---------------------------------------------------------------- argsort of one series per cell
// order[r * ld_order + c] = index of the r-th smallest value of x[t * row_stride + c], t = 0 .. n-1 (equal values in
// index order).  Used to order a cell's analog training window and its query steps by the first predictor, which
// is what lets the kNN search prune (analog_kernels.cu).  1024 threads, n <= 32 768.
__global__ void __launch_bounds__(LG_NT_BIG, 1)
series_argsort_kernel(const float* __restrict__ x, int64_t row_stride, int64_t C, int n,
                      int32_t* __restrict__ order, int64_t ld_order, const uint8_t* __restrict__ valid) {
    constexpr int NT = LG_NT_BIG;
    using G = LgCfg<NT>;
    extern __shared__ __align__(16) uint32_t lg_smem[];
    uint32_t* W = lg_smem;
    LgShared* sh = reinterpret_cast<LgShared*>(lg_smem + G::W_WORDS);
    uint16_t* P2M = reinterpret_cast<uint16_t*>(lg_smem + G::W_WORDS + (sizeof(LgShared) + 3) / 4);   // position → member
    const int64_t c = blockIdx.x;
    if (valid && !valid[c]) return;
    const int tid = threadIdx.x, lane = tid & 31;
    float v[LG_E];
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int k = 0; k < LG_E; ++k) {
        const int j = tid + k * NT;
        v[k] = 0.0f;
        if (j < n) {
            v[k] = x[(int64_t)j * row_stride + c] + 0.0f;
            lo = fminf(lo, v[k]); hi = fmaxf(hi, v[k]);
        }
    }
    lg_block_minmax<NT>(lo, hi, sh);
    const float scale = bm_scale_f32<G::NB>(lo, hi);
    const int total = lg_build<NT>(v, n, lo, scale, W, sh);
    const int n_lo = n - total;
    auto emit = [&](int pos, int j) { order[(int64_t)pos * ld_order + c] = j; };
    const uint32_t Wsa = bm_saddr(W);
#pragma unroll
    for (int b = 0; b < LG_E; b += 8) {
        if (b * NT >= n) continue;                           // CTA-uniform
        uint32_t q[8], dirty[8];
        int on[8], pos[8], cur[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            q[u] = bm_bucket_f32<G::NB>(v[b + u], lo, scale);
            on[u] = (v[b + u] > lo) ? n - (tid + (b + u) * NT) : 0;
        }
        bm_lookup_batch<8>(W, Wsa, q, on, pos, dirty, cur);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = tid + (b + u) * NT;
            if (j < n) {
                if (on[u] <= 0) P2M[atomicAdd(&sh->n_lo_cursor, 1u)] = (uint16_t)j;      // values at the lower bound (NaN too)
                else if (!dirty[u]) emit(n_lo + pos[u], j);
                else P2M[n_lo + pos[u] + cur[u]] = (uint16_t)j;
            }
        }
    }
    __syncthreads();
    auto key_of = [&](int j) -> float { return __ldg(x + (int64_t)j * row_stride + c) + 0.0f; };
    // members of one range of P2M in (key, index) order; keys are skipped for the lower-bound class (all equal)
    auto order_small = [&](int first, int cnt, bool by_key) {
        float k[LG_SMALL];
        int m[LG_SMALL];
#pragma unroll
        for (int a = 0; a < LG_SMALL; ++a) { m[a] = (a < cnt) ? (int)P2M[first + a] : 0x7fffffff; k[a] = (a < cnt && by_key) ? key_of(m[a]) : 0.0f; }
#pragma unroll
        for (int a = 0; a < LG_SMALL; ++a) {
            int r = 0;
#pragma unroll
            for (int b = 0; b < LG_SMALL; ++b)
                if (b != a) r += (b < cnt && (k[b] < k[a] || (k[b] == k[a] && m[b] < m[a]))) ? 1 : 0;
            if (a < cnt) emit(first + r, m[a]);
        }
    };
    auto order_big = [&](int first, int cnt, bool by_key) {            // whole warp, one member per lane and round
        for (int a = lane; a < cnt; a += 32) {
            const int ma = (int)P2M[first + a];
            const float ka = by_key ? key_of(ma) : 0.0f;
            int r = 0;
            for (int b = 0; b < cnt; ++b) {
                const int mb = (int)P2M[first + b];
                const float kb = by_key ? key_of(mb) : 0.0f;
                r += (kb < ka || (kb == ka && mb < ma)) ? 1 : 0;
            }
            emit(first + r, ma);
        }
    };
    lg_for_dirty_entries<NT>(W, sh, [&](int start, int cnt) { order_small(n_lo + start, cnt, true); },
                             [&](int start, int cnt) { order_big(n_lo + start, cnt, true); });
    if (tid < 32) {                                          // the lower-bound class: equal keys, index order
        if (n_lo <= LG_SMALL) { if (lane == 0 && n_lo > 0) order_small(0, n_lo, false); }
        else order_big(0, n_lo, false);
    }
}

constexpr size_t lg_argsort_smem() {
    return (size_t)LgCfg<LG_NT_BIG>::W_WORDS * 4 + ((sizeof(LgShared) + 3) / 4) * 4 + (size_t)LgCfg<LG_NT_BIG>::NMAX * 2;
}

constexpr size_t lg_fit_smem() { return (size_t)LG_W_WORDS * 4 + sizeof(LgShared); }
constexpr size_t lg_predict_smem() { return (size_t)LG_W_WORDS * 4 + ((sizeof(LgShared) + 3) / 4) * 4 + (size_t)LG_NMAX * 2; }

int qm_fit_long(const FitParams& f, cudaStream_t st) {
    auto kern = qm_fit_long_kernel;
    SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lg_fit_smem()));
    dim3 grid((unsigned)f.C, (unsigned)f.n_groups);
    kern<<<grid, LG_NT, lg_fit_smem(), st>>>((const float*)f.y, f.ld, f.C, f.rows, f.len, f.off, f.max_len,
                                             (float*)f.state, f.state_ld, f.valid, f.nonfinite);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

int qm_predict_long(const PredictParams& p, cudaStream_t st) {
    auto kern = qm_predict_long_kernel;
    SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lg_predict_smem()));
    dim3 grid((unsigned)p.C, (unsigned)p.n_groups);
    kern<<<grid, LG_NT, lg_predict_smem(), st>>>(p);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_series_argsort(const void* x, int dtype, int64_t row_stride, int64_t n_cells, int n_steps,
                                  int32_t* order, int64_t ld_order, const uint8_t* cell_valid, void* stream) {
    if (!x || !order) return sdb_fail(SDB_E_INVALID, "sdb_series_argsort: NULL pointer");
    if (n_cells <= 0 || n_steps <= 0 || row_stride < n_cells || ld_order < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_series_argsort: bad shape");
    if (dtype != SDB_F32) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_series_argsort: float32 only");
    if (n_steps > LgCfg<LG_NT_BIG>::NMAX) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_series_argsort: at most %d steps", LgCfg<LG_NT_BIG>::NMAX);
    auto kern = series_argsort_kernel;
    SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lg_argsort_smem()));
    kern<<<(unsigned)n_cells, LG_NT_BIG, lg_argsort_smem(), (cudaStream_t)stream>>>((const float*)x, row_stride, n_cells, n_steps,
                                                                                  order, ld_order, cell_valid);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_series_argsort_max_steps(void) { return LgCfg<LG_NT_BIG>::NMAX; }

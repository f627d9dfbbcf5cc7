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

#ifndef SDB_TILE_CT
#define SDB_TILE_CT 8
#endif
constexpr int TILE_CT = SDB_TILE_CT;   // cells per CTA (= warps per CTA): 8 cells = one 32-byte sector per row
constexpr int TILE_THREADS = 32 * TILE_CT;
constexpr int TILE_LMAX = 16;       // longest same-bucket run fixed up locally

// Cache policy of the tile's row accesses and CTA pairing (measured with ncu, profiles/README.md).
// A CTA touches ONE 32-byte sector per row while DRAM moves 64-byte atoms, so what matters is whether
// the neighbouring CTA's half meets it in L2:
//   fit loads      default policy (ld.global.nc): with streaming loads (ld.cs, evict-first) the other half
//                  was gone before the neighbour asked for it — 11.0 GB read for 5.7 GB of input; now 5.7 GB
//   predict loads  streaming: keeps the output lines (st.cs) in L2 long enough to be completed — with
//                  default-policy loads the kernel wrote 9.6 GB for 5.7 GB of output, with ld.cs 7.4 GB
//   stores         streaming (st.cs): default-policy / st.cg / st.wt stores wrote 11-12 GB
//   SDB_TILE_CLUSTER n: the tile kernels launch as clusters of n CTAs along x — the CTAs of a cluster own
//                  ADJACENT 8-cell row segments (the two halves of a 64-byte atom) and start together
#ifndef SDB_FIT_LD_POLICY
#define SDB_FIT_LD_POLICY 1      // 0 = ld.cs, 1 = default
#endif
#ifndef SDB_PRED_LD_POLICY
#define SDB_PRED_LD_POLICY 0
#endif
#ifndef SDB_ST_POLICY
#define SDB_ST_POLICY 0          // 0 = st.cs, 1 = default, 2 = st.cg, 3 = st.wt
#endif
#ifndef SDB_TILE_CLUSTER
#define SDB_TILE_CLUSTER 2
#endif
template <int POLICY, class P>
__device__ __forceinline__ auto ld_row(const P* p) { if constexpr (POLICY == 0) return __ldcs(p); else return __ldg(p); }
#define SDB_LD_ROW(p) ld_row<LDP>(p)
#if SDB_ST_POLICY == 0
#define SDB_ST_ROW(p, v) __stcs(p, v)
#elif SDB_ST_POLICY == 2
#define SDB_ST_ROW(p, v) __stcg(p, v)
#elif SDB_ST_POLICY == 3
#define SDB_ST_ROW(p, v) __stwt(p, v)
#else
#define SDB_ST_ROW(p, v) (*(p) = (v))
#endif
#ifndef SDB_FIT_BATCH
#define SDB_FIT_BATCH 8
#endif
#ifndef SDB_PRED_BATCH
#define SDB_PRED_BATCH 4
#endif
#ifndef SDB_STORE_BATCH
#define SDB_STORE_BATCH 1
#endif
template <int E> struct TileGeom {
    static constexpr int NP = 32 * E;
    // padded row: 8 words in front (members -4..-1 of the rolling window read as zeros), one extra word
    // per 32 members (skew), room for members n..n+4 behind; NPS mod 32 is 4 (E=32) / 28 (E=8) so the
    // eight rows of a tile start in different banks — conflict-free for every access pattern used
    static constexpr int NPS = (TILE_CT == 8) ? ((E == 32) ? 1092 : 284) : ((E == 32) ? 1096 : 296);
    static constexpr int LOG = (E == 8) ? 8 : 10;
    static constexpr uint32_t QMAX = (1u << (32 - LOG)) - 1u;   // largest bucket number
    // members take buckets 0..QTOP; the padding item at position j takes bucket QTOP + 1 + j, so padding
    // sorts last in position order and NO two padding items share a bucket (the same-bucket scan after
    // the sort then needs no "is this a member" guard)
    static constexpr uint32_t QTOP = QMAX - (uint32_t)NP;
    static constexpr uint32_t PAD_BASE = (QTOP + 1u) << LOG;    // key of padding position j: PAD_BASE + j * (2^LOG + 1)
    static constexpr uint32_t PAD_STEP = (1u << LOG) + 1u;
    static_assert(E == 8 || E == 32, "tile kernels are instantiated for NP = 256 and 1024");
};
__device__ __forceinline__ int skew(int j) { return j + (j >> 5) + 8; }      // valid for j >= -8

template <int E>
constexpr size_t fit_tile_smem() { return (size_t)TILE_CT * TileGeom<E>::NPS * 4; }
template <int E>
constexpr size_t predict_tile_smem() { return (size_t)TILE_CT * 2 * TileGeom<E>::NPS * 4 + (size_t)TileGeom<E>::NP * 4; }

// Cooperative, coalesced load of one group's rows for the CTA's 8 cells into tile[cell][skew(j)],
// transposing on the way.  Fast path: one 16-byte load per (row, 4 cells) — two threads cover the
// 32-byte row segment — and four conflict-free shared stores; it needs 16-byte aligned rows
// (ld % 4 == 0, aligned base) and a full tile.  Otherwise one 4-byte load per (row, cell).
// The row numbers are also left in shared memory (rowtab) for the store pass.
template <int E, int BATCH, int LDP>
__device__ __forceinline__ void load_tile(float* tile, const float* __restrict__ src, int64_t ld, int64_t C,
                                          int64_t c0, const int32_t* __restrict__ rg, int n,
                                          const uint8_t* __restrict__ valid, bool allow_vec,
                                          int32_t* rowtab = nullptr, int tid = -1) {
    const int tix = tid < 0 ? (int)threadIdx.x : tid;      // thread index inside the 256-thread group that loads this tile
    constexpr int NPS = TileGeom<E>::NPS;
    static_assert(TILE_CT == 8, "the vector path assumes 8-cell tiles");
    // zeros around the group: members -4..-1 and n..NP+4 of every row (the rolling window and the
    // padding lanes read them unconditionally)
    const bool vec = allow_vec && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src + c0) & 15) == 0) &&
                     (c0 + TILE_CT <= C);
    if (vec && BATCH != 0) {
        // the vector path below writes every position 0..NP-1 itself (zeros from n on): only the 4 + 5
        // halo slots outside [0, NP) are left
        if (tix < TILE_CT * 9) {
            const int r = tix / 9, k9 = tix - r * 9;
            tile[r * NPS + skew(k9 < 4 ? k9 - 4 : TileGeom<E>::NP + k9 - 4)] = 0.0f;
        }
    } else {
        const int tail = TileGeom<E>::NP + 5 - n;
        const int per_row = 4 + tail;
        for (int i = tix; i < TILE_CT * per_row; i += TILE_THREADS) {
            const int r = i / per_row, k9 = i - r * per_row;
            tile[r * NPS + skew(k9 < 4 ? k9 - 4 : n + k9 - 4)] = 0.0f;
        }
    }
    if (vec) {
        // BATCH row segments per thread are loaded before their shared stores (BATCH = 0: leave the
        // scheduling of the unrolled loop to the compiler — fewer live registers)
        constexpr int IT = TileGeom<E>::NP / (TILE_THREADS / 2);
        const int quad = tix & 1;
        const int64_t cq = c0 + 4 * quad;
        bool ok0 = true, ok1 = true, ok2 = true, ok3 = true;
        if (valid) { ok0 = valid[cq]; ok1 = valid[cq + 1]; ok2 = valid[cq + 2]; ok3 = valid[cq + 3]; }
        float* d0 = tile + (4 * quad) * NPS;
        const float* col = src + cq;
        const int j0 = tix >> 1;
        if constexpr (BATCH == 0) {
#pragma unroll 8
            for (int j = j0; j < n; j += TILE_THREADS / 2) {
                const int32_t row = __ldg(rg + j);
                if (rowtab && quad == 0) rowtab[j] = row;
                const float4 x = SDB_LD_ROW(reinterpret_cast<const float4*>(col + (uint64_t)(uint32_t)row * (uint32_t)ld));
                const int at = skew(j);
                d0[at] = ok0 ? x.x : 0.0f;
                d0[NPS + at] = ok1 ? x.y : 0.0f;
                d0[2 * NPS + at] = ok2 ? x.z : 0.0f;
                d0[3 * NPS + at] = ok3 ? x.w : 0.0f;
            }
        } else {
            constexpr int B = BATCH < IT ? BATCH : IT;
#pragma unroll
            for (int ib = 0; ib < IT; ib += B) {
                int32_t row[B];
                float4 x[B];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int j = j0 + (ib + u) * (TILE_THREADS / 2);
                    row[u] = (j < n) ? __ldg(rg + j) : 0;
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int j = j0 + (ib + u) * (TILE_THREADS / 2);
                    x[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (j < n) x[u] = SDB_LD_ROW(reinterpret_cast<const float4*>(col + (uint64_t)(uint32_t)row[u] * (uint32_t)ld));
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    // every position is stored (zeros from n on), no branch: the rows behind the group are
                    // the zero padding the rolling window and the padding lanes read
                    const int j = j0 + (ib + u) * (TILE_THREADS / 2);
                    if (rowtab && quad == 0) rowtab[j] = row[u];
                    const int at = skew(j);
                    d0[at] = ok0 ? x[u].x : 0.0f;
                    d0[NPS + at] = ok1 ? x[u].y : 0.0f;
                    d0[2 * NPS + at] = ok2 ? x[u].z : 0.0f;
                    d0[3 * NPS + at] = ok3 ? x[u].w : 0.0f;
                }
            }
        }
    } else {
        const int cc = tix & (TILE_CT - 1);
        const int64_t c = c0 + cc;
        const bool ok = c < C && (!valid || valid[c]);
        float* dst = tile + cc * NPS;
        const float* col = src + (ok ? c : 0);
        constexpr int RS = TILE_THREADS / TILE_CT;       // rows per pass
        for (int jb = tix / TILE_CT; jb < n; jb += 8 * RS) {
            int32_t row[8];
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) row[u] = (jb + u * RS < n) ? __ldg(rg + jb + u * RS) : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                x[u] = (ok && jb + u * RS < n) ? SDB_LD_ROW(col + (uint64_t)(uint32_t)row[u] * (uint32_t)ld) : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = jb + u * RS;
                if (j < n) {
                    if (rowtab && cc == 0) rowtab[j] = row[u];
                    dst[skew(j)] = x[u];
                }
            }
        }
    }
}

// ---------------------------------------------------------------- fit
template <int E>
__global__ void __launch_bounds__(TILE_THREADS, 24 / TILE_CT)
qm_fit_tile_kernel(const float* __restrict__ y, int64_t ld, int64_t C,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ len,
                   const int64_t* __restrict__ off, int max_len,
                   float* __restrict__ state, int64_t state_ld, const uint8_t* __restrict__ valid,
                   int32_t* __restrict__ nonfinite, int no_vec) {
    using G = TileGeom<E>;
    extern __shared__ float tile_f[];
    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * TILE_CT;
    const int n = len[g];
    const int32_t* rg = rows + (int64_t)g * max_len;
    load_tile<E, SDB_FIT_BATCH, SDB_FIT_LD_POLICY>(tile_f, y, ld, C, c0, rg, n, valid, !no_vec);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = c0 + warp;
    if (c >= C || (valid && !valid[c])) return;
    float* my = tile_f + warp * G::NPS + skew(lane * E);      // the lane's E members are contiguous: my[e]
    const int nj = n - lane * E;                               // members of this lane that exist: e < nj
    K32 v[E];
    bool bad = false;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float x = my[e];
        bad |= (e < nj) && !isfinite(x);
        v[e].k = (e < nj) ? f32_to_sortable(x + 0.0f) : 0xffffffffu;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
    sort_blocked<K32, E, 32>(v, lane, nullptr);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e) my[e] = sortable_to_f32(v[e].k);       // positions >= n hold padding, never stored
    __syncwarp();
    float* dst = state + c * state_ld + off[g];
    const float* row = tile_f + warp * G::NPS;
    for (int j = lane; j < n; j += 32) dst[j] = row[skew(j)];      // 128-byte coalesced rows of the cell record
}

// ---------------------------------------------------------------- predict helpers
// sum / cnt for the window counts 1..9, correctly rounded: q0 = sum * rc with rc = RN(1/cnt), then one
// exact-residual FMA correction (Markstein) — three FP64 instructions, no division subroutine.
static __constant__ double RC_TAB[10] = {0.0, 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0, 1.0 / 6.0, 1.0 / 7.0,
                                         1.0 / 8.0, 1.0 / 9.0};
static __constant__ double CNT_TAB[10] = {0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0};
__device__ __forceinline__ double div_count(double sum, int cnt) {
    const double rc = RC_TAB[cnt];
    const double q0 = sum * rc;
    const double r = fma(-q0, CNT_TAB[cnt], sum);
    return fma(r, rc, q0);
}

// members of the centred 9-window of member j that exist inside a group of n
__device__ __forceinline__ int win_count(int j, int n) {
    const int a = 4 - j, b = j + 5 - n;
    return 9 - (a > 0 ? a : 0) - (b > 0 ? b : 0);
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

// Out-of-line exact ranking of one (cell, group): 64-bit key + position sort (the generic
// algorithm) for the rare group whose keys defeat the bucket quantisation (a bucket holding three
// or more distinct keys).  Writes the 1-based tie-max rank of member j over the input row:
// Xu[skew(j)] = rank (the inputs are not needed afterwards).
template <int E, bool SHIFT>
__device__ __noinline__ void rank_exact64(float* myX, int n, double xc, int lane) {
    K64I u[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        u[e] = (j < n) ? make_rank_item64(exact_key<SHIFT>(myX, n, j, xc), (uint32_t)j) : sentinel_item<K64I>((uint32_t)j);
    }
    __syncwarp();                                   // every lane has read what it needs from the row
    sort_blocked<K64I, E, 32>(u, lane, nullptr);
    int r2[E];
    tie_max_ranks<K64I, E, 32>(u, lane, r2, nullptr);
    uint32_t* Xu = reinterpret_cast<uint32_t*>(myX);
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (u[e].i < (uint32_t)n) Xu[skew((int)u[e].i)] = (uint32_t)r2[e];
}

// three-way exact comparison of the rank keys of members a and b: -1, 0, +1
static __device__ __noinline__ int cmp_members_shift(const float* myX, int n, int a, int b, double xc) {
    const double ka = window_key(myX, n, a, xc), kb = window_key(myX, n, b, xc);
    return ka < kb ? -1 : (ka > kb ? 1 : 0);
}
template <bool SHIFT>
__device__ __forceinline__ int cmp_members(const float* myX, int n, int a, int b, double xc) {
    if (SHIFT) return cmp_members_shift(myX, n, a, b, xc);
    const float ka = myX[skew(a)], kb = myX[skew(b)];          // the inputs themselves are the keys (-0 == +0)
    return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

// mapped value of 1-based rank rk: the fitted order statistic itself when the lengths agree,
// else the Cunnane interpolation / OLS tails, read from the state record in global memory
static __device__ __noinline__ float mapped_value_general(const PredictParams& p, const float* __restrict__ S, int rk, int n, int m) {
    auto Sat = [&](int i) -> double { return (double)__ldg(S + i); };
    const Cunnane cu = cunnane_of(p);
    return (float)inverse_cdf_acc(rk, n, m, Sat, pp_denominator(n, cu), pp_denominator(m, cu), cu);
}

// The same map for the interior of the fitted CDF with the per-(cell, group) constants hoisted: the
// plotting positions (i - alpha) / den are formed with one reciprocal per group and an exact-residual FMA
// correction (RN(1/den) → q0 = a * rc → r = fma(-q0, den, a) → fma(r, rc, q0): the correctly rounded
// quotient), which leaves ONE true division per value (the slope) where the generic routine spends eight.
// Tails (a handful of ranks when T_pred > T_fit) and degenerate groups go to the generic routine.
struct GroupMap {
    double dn, dm, rdn, rdm, p1, pm, alpha;
    double a_lo, b_lo, a_hi, b_hi;      // value = a * q + b beyond the fitted positions (valid when `tails`)
    int n, m;
    bool tails;
};
__device__ __forceinline__ double fdiv(double a, double den, double rc) {
    const double q0 = a * rc;
    return fma(fma(-q0, den, a), rc, q0);
}
__device__ __forceinline__ GroupMap make_group_map(const PredictParams& p, int n, int m) {
    const Cunnane cu = cunnane_of(p);
    GroupMap g;
    g.n = n; g.m = m; g.alpha = cu.alpha;
    g.dn = pp_denominator(n, cu); g.dm = pp_denominator(m, cu);
    g.rdn = 1.0 / g.dn; g.rdm = 1.0 / g.dm;
    g.p1 = pp_of(1, g.dm, cu); g.pm = pp_of(m, g.dm, cu);
    g.tails = false;
    g.a_lo = g.b_lo = g.a_hi = g.b_hi = 0.0;
    return g;
}
// The two tail lines of a (cell, group), once per warp: lane 0 fits the lower one, lane 31 the upper
// one in the same instruction stream (the per-value generic routine refits the 10-point line for every
// rank that falls outside — four ranks per group when T_pred = 3 T_fit, each on a different lane: they
// made up a third of the instructions of the interpolation path).  A tail that does not extrapolate
// clamps to the end value (a = 0).
__device__ __forceinline__ void fit_group_tails(const PredictParams& p, GroupMap& g, const float* __restrict__ S, int lane) {
    const Cunnane cu = cunnane_of(p);
    double a = 0.0, b = 0.0;
    if (lane == 0 || lane == 31) {
        const bool low = (lane == 0);
        auto Sat = [&](int i) -> double { return (double)__ldg(S + i); };
        const int ne = g.m < cu.ne ? g.m : cu.ne;
        if (low ? cu.lo : cu.hi) ols_tail(Sat, low ? 1 : g.m - ne + 1, ne, g.dm, cu, a, b);
        else b = Sat(low ? 0 : g.m - 1);
    }
    g.a_lo = __shfl_sync(0xffffffffu, a, 0);  g.b_lo = __shfl_sync(0xffffffffu, b, 0);
    g.a_hi = __shfl_sync(0xffffffffu, a, 31); g.b_hi = __shfl_sync(0xffffffffu, b, 31);
    g.tails = true;
}
// position of quantile q inside the fitted CDF (pure arithmetic, no loads): j = 0 → take the generic routine
struct InterpLoc { int j; double q, xj, xj1; };
__device__ __forceinline__ InterpLoc interp_locate(const GroupMap& g, int rk) {
    InterpLoc L;
    L.j = 0;
    L.q = fdiv((double)rk - g.alpha, g.dn, g.rdn);
    L.xj = L.xj1 = 0.0;
    if (g.m < 2) return L;
    if (g.tails && L.q < g.p1) { L.j = -1; return L; }    // value = a_lo * q + b_lo
    if (g.tails && L.q > g.pm) { L.j = -2; return L; }
    if (!(L.q >= g.p1) || !(L.q < g.pm)) return L;
    int j = (int)floor(L.q * g.dm + g.alpha);
    j = j < 1 ? 1 : (j > g.m - 1 ? g.m - 1 : j);
    double xj = fdiv((double)j - g.alpha, g.dm, g.rdm), xj1 = fdiv((double)(j + 1) - g.alpha, g.dm, g.rdm);
    if (xj > L.q) {                                      // the guess is off by at most one
        if (j == 1) return L;
        --j; xj1 = xj; xj = fdiv((double)j - g.alpha, g.dm, g.rdm);
    } else if (xj1 <= L.q) {
        if (j + 1 > g.m - 1) return L;
        ++j; xj = xj1; xj1 = fdiv((double)(j + 1) - g.alpha, g.dm, g.rdm);
    }
    if (xj > L.q || xj1 <= L.q) return L;               // never expected: the generic routine stays exact
    L.j = j; L.xj = xj; L.xj1 = xj1;
    return L;
}
__device__ __forceinline__ float interp_value(const InterpLoc& L, float y0, float y1) {
    if (L.q == L.xj) return y0;
    const double slope = ((double)y1 - (double)y0) / (L.xj1 - L.xj);
    return (float)(slope * (L.q - L.xj) + (double)y0);
}
// one value, self-contained and out of line (the rare exact-sort path)
static __device__ __noinline__ float mapped_value_interp(const PredictParams& p, const float* __restrict__ S, int rk, int n, int m) {
    const GroupMap g = make_group_map(p, n, m);
    const InterpLoc L = interp_locate(g, rk);
    if (L.j == 0) return mapped_value_general(p, S, rk, n, m);
    return interp_value(L, __ldg(S + L.j - 1), __ldg(S + L.j));
}

// ---------------------------------------------------------------- predict
// One warp maps one (cell, group): the group's inputs are in the shared row myX, the outputs
// (float bits) are left in the shared row R.  See the header comment of this file and DESIGN.md §4.
template <int E, bool SHIFT>
__device__ __forceinline__ void map_cell_group(const PredictParams& p, float* myX, uint32_t* R, int lane,
                                               int64_t c, int n, int m, const int32_t* __restrict__ rg,
                                               const float* __restrict__ S, bool quad_ok, double xc, double yc) {
    using G = TileGeom<E>;
    constexpr int LOG = G::LOG;
    constexpr uint32_t QTOP = G::QTOP;
    constexpr uint32_t IDX = (1u << LOG) - 1u;
    const int j0 = lane * E;
    const int j1 = (j0 + E < n) ? j0 + E : n;
    const int nj = n - j0;                     // members of this lane that exist: e < nj
    const int rb = skew(j0);
    const bool ratio = (p.mode == SDB_MODE_BCSD_P) && p.return_anoms;
    const bool same = (n == m);

    // final value of a member from its mapped value (bcsd.py:263,267 / 170-185).  For the shifted
    // model the float32 term (shift - y_climo) was parked in R; adding the float32 mapped value to it
    // with one float32 add is the correctly rounded sum of the two.
    auto finish = [&](int member, float val) {
        const int at = skew(member);
        if (SHIFT) R[at] = __float_as_uint(__fadd_rn(__uint_as_float(R[at]), val));
        else       R[at] = __float_as_uint(ratio ? (float)((double)val / yc) : val);
    };
    const double park_off = (SHIFT && p.return_anoms) ? yc : 0.0;

    // ---- 1. own members (+ 4 / 5 halo) to registers; key bounds from the plain value range
    constexpr int HL = SHIFT ? 4 : 0, HR = SHIFT ? 5 : 0;
    K32 v[E];
    {
        float xh[E + HL + HR];
#pragma unroll
        for (int i = 0; i < E + HL + HR; ++i) {
            const int e = i - HL;                          // member offset inside / around the lane's block
            const int jj = j0 + e;
            // E == 32: the skew step only changes at the block edges → static offsets from the row base
            const int addr = (E == 32) ? rb + e + (e < 0 ? -1 : (e >= 32 ? 1 : 0)) : skew(jj);
            xh[i] = myX[addr];                              // zeros outside [0, n) by construction of the tile
        }
        // The bucket bounds are only a guess (whatever falls outside clamps to the end buckets and is
        // compared exactly), so they come from the lanes whose block is full: no per-member guard.  The
        // one partially filled lane joins only when no lane is full (groups shorter than E).
        float lo32 = xh[HL], hi32 = xh[HL], nanacc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float x = xh[e + HL];
            if (e > 0) { lo32 = fminf(lo32, x); hi32 = fmaxf(hi32, x); }
            nanacc = fmaf(x, 0.0f, nanacc);                 // NaN / inf anywhere → NaN (padding members are zeros)
        }
        if (nj < E) {
            lo32 = INFINITY; hi32 = -INFINITY;
            if (n < E) {                                    // warp-uniform: a group inside lane 0
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (e < nj) { lo32 = fminf(lo32, xh[e + HL]); hi32 = fmaxf(hi32, xh[e + HL]); }
            }
        }
        if (nanacc != nanacc && p.nonfinite) atomicOr(p.nonfinite, 1);
        lo32 = warp_min(lo32); hi32 = warp_max(hi32);

        // ---- 2. exact rank keys → monotone bucket number, packed with the member position.
        // The bounds are a guess (value range + 1/8 margin); keys outside clamp to the end
        // buckets, which keeps the map monotone — whatever shares a bucket is compared exactly.
        // Every position is computed unconditionally (the padding members are zeros and their slots of
        // R exist); only the final key is a select between the member's packed word and the padding key.
        const uint32_t pad0 = G::PAD_BASE + (uint32_t)j0 * G::PAD_STEP;
        if (SHIFT) {
            const double range = (double)hi32 - (double)lo32;
            const double lo = (double)lo32 - 0.125 * range, hi = (double)hi32 + 0.125 * range;
            const double scale = (hi > lo && isfinite(hi - lo)) ? (double)QTOP / (hi - lo) : 0.0;
            const double nls = -lo * scale;                 // t = key * scale - lo * scale: one monotone FMA
            // members of the window that exist: 9 - a - b, a = 4 - j (front, lane 0 only), b = j + 5 - n (back)
            const int back0 = j0 + 5 - n;                   // b of e = 0
            const bool front = (lane == 0);
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) sum += (double)xh[i];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int b = back0 + e;
                b = b < 0 ? 0 : (b > 8 ? 8 : b);            // clamp: padding positions still index the tables
                int cnt = 9 - b;
                if (e < 4) { cnt -= front ? (4 - e) : 0; cnt = cnt < 1 ? 1 : cnt; }
                const double shift = div_count(sum, cnt) - xc;
                const double t = fma((double)xh[e + 4] - shift, scale, nls);
                uint32_t q = __double2uint_rd(t);           // saturating: negative and NaN map to 0
                q = q > QTOP ? QTOP : q;
                const uint32_t w = (q << LOG) + (uint32_t)(j0 + e);
                v[e].k = (e < nj) ? w : pad0 + (uint32_t)e * G::PAD_STEP;
                R[rb + e] = __float_as_uint((float)(shift - park_off));
                sum += (double)xh[e + 9];
                sum -= (double)xh[e];
            }
        } else {
            // Bucket 0 is reserved for the values AT the lower bound, everything above it starts at bucket 1
            // (still monotone).  Zero-inflated precipitation: the run of exact zeros then never shares a
            // bucket with a tiny positive value — that mix is not sorted by value inside the bucket and
            // would send the whole group to the exact 64-bit sort (a 20x straggler that holds its CTA).
            const float range = hi32 - lo32;
            const float scale = (range > 0.0f && isfinite(range)) ? (float)(QTOP - 1u) / range : 0.0f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float t = (xh[e] - lo32) * scale;
                uint32_t q = __float2uint_rd(t) + ((xh[e] > lo32) ? 1u : 0u);   // saturating: negative and NaN map to 0
                q = q > QTOP ? QTOP : q;
                const uint32_t w = (q << LOG) + (uint32_t)(j0 + e);
                v[e].k = (e < nj) ? w : pad0 + (uint32_t)e * G::PAD_STEP;
            }
        }
    }

    // ---- 3. one 32-bit keys-only sort
    sort_blocked<K32, E, 32>(v, lane, nullptr);
    const uint32_t nxt_first = __shfl_down_sync(0xffffffffu, v[0].k, 1);
    uint32_t bm_eq = 0;                       // bit e: sorted positions (pos, pos+1) share a bucket
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const uint32_t kn = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
        bm_eq |= (((v[e].k ^ kn) >> LOG) == 0u) ? (1u << e) : 0u;       // padding never shares a bucket
    }
    if (lane == 31) bm_eq &= 0x7fffffffu >> (32 - E);      // the last position has no successor (nxt_first is the lane's own)
    __syncwarp();                             // shifts parked in R are visible to every lane

    // ---- 4. (member, rank) of every sorted position → mapped value → output
    // mode 0: one member per bucket, rank = position + 1.
    // mode 2: members sharing a bucket are compared exactly; runs of exact ties take the run end
    //         (tie-max rank), an isolated inverted pair is swapped.
    // mode 3: any other structure (3+ distinct keys in a bucket) → exact 64-bit sort of the group.
    int mode = 0;
    uint32_t bm_gt = 0, bm_tie = 0;           // bit e: pair (pos, pos+1) is inverted / exactly tied
    uint32_t prev_gt = 0;                     // the pair (j0 - 1, j0) owned by the previous lane is inverted
    int tie_carry = 0;                        // rank of a tie run that continues past this lane's last member
    constexpr uint32_t EMASK = (E == 32) ? 0xffffffffu : ((1u << (E & 31)) - 1u);
    if (__any_sync(0xffffffffu, bm_eq != 0)) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if ((bm_eq >> e) & 1u) {
                const uint32_t kn = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
                const int cmp = cmp_members<SHIFT>(myX, n, (int)(v[e].k & IDX), (int)(kn & IDX), xc);
                bm_gt |= (cmp > 0) ? (1u << e) : 0u;
                bm_tie |= (cmp == 0) ? (1u << e) : 0u;
            }
        }
        // a non-tied same-bucket pair must be isolated: no other same-bucket pair touching it
        const uint32_t nontie = bm_eq & ~bm_tie;
        const uint32_t nxt_eq0 = __shfl_down_sync(0xffffffffu, bm_eq & 1u, 1);
        const uint32_t prv_eqL = __shfl_up_sync(0xffffffffu, (bm_eq >> (E - 1)) & 1u, 1);
        uint32_t touching = (bm_eq << 1) | (bm_eq >> 1);
        if (lane > 0 && prv_eqL) touching |= 1u;
        if (lane < 31 && nxt_eq0) touching |= 1u << (E - 1);
        mode = __any_sync(0xffffffffu, (nontie & touching) != 0) ? 3 : 2;
        prev_gt = __shfl_up_sync(0xffffffffu, (bm_gt >> (E - 1)) & 1u, 1);
        if (lane == 0) prev_gt = 0;
        // tie runs: a position's rank is 1 + the first position at/after it whose tie bit is clear
        const uint32_t open = ~bm_tie & EMASK;
        const int first_end = j0 + __ffs(open);
        const uint32_t has = __ballot_sync(0xffffffffu, open != 0);
        const uint32_t higher = (lane == 31) ? 0u : (has & ~((2u << lane) - 1u));
        tie_carry = __shfl_sync(0xffffffffu, first_end, higher ? (__ffs(higher) - 1) : lane);
    }

    if ((mode == 0 || mode == 2) && same && !p.rank_out && (E % 4 == 0) && quad_ok) {
        // the common case: same length and (almost) one member per bucket — the member at sorted
        // position pos takes the fitted order statistic S[pos]; the lane's 32 values are 8 vector loads
        float sv[E + 1];
        // quads past the group's last one re-read that last quad (positions >= n only feed padding slots);
        // the last quad may reach up to 3 values into the 16-byte alignment gap behind the group, which
        // belongs to the cell's record
        const int last4 = (n - 1) & ~3;
#pragma unroll
        for (int q4 = 0; q4 < E / 4; ++q4) {
            const int pos = j0 + 4 * q4;
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(S + (pos < last4 ? pos : last4)));
            sv[4 * q4] = t4.x; sv[4 * q4 + 1] = t4.y; sv[4 * q4 + 2] = t4.z; sv[4 * q4 + 3] = t4.w;
        }
        if (mode == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) finish((int)(v[e].k & IDX), sv[e]);    // padding positions land in slots >= n of R
        } else {
            // same scatter with the same-bucket structure patched in registers: an isolated inverted pair
            // exchanges its members; every position of an exact-tie run (any length — the zeros of
            // precipitation are one run of hundreds) takes the order statistic at the END of its run
            // (tie-max rank): a backward scan over the lane's positions, seeded with the value at the run
            // end in a later lane when the lane's last position is still inside a run
            float run_val = ((bm_tie >> (E - 1)) & 1u) ? __ldg(S + tie_carry - 1) : 0.0f;
            const uint32_t prv_last_k = __shfl_up_sync(0xffffffffu, v[E - 1].k, 1);
#pragma unroll
            for (int e = E - 1; e >= 0; --e) {
                uint32_t w = v[e].k;
                if ((bm_gt >> e) & 1u) w = (e == E - 1) ? nxt_first : v[e == E - 1 ? e : e + 1].k;
                else if ((e == 0) ? prev_gt : ((bm_gt >> (e == 0 ? 0 : e - 1)) & 1u)) w = (e == 0) ? prv_last_k : v[e == 0 ? e : e - 1].k;
                run_val = ((bm_tie >> e) & 1u) ? run_val : sv[e];
                finish((int)(w & IDX), run_val);
            }
        }
    } else if (mode == 3) {
        rank_exact64<E, SHIFT>(myX, n, xc, lane);            // ranks by member → input row
        __syncwarp();
        const uint32_t* Xu = reinterpret_cast<const uint32_t*>(myX);
        for (int j = j0; j < j1; ++j) {
            const int rk = (int)Xu[skew(j)];
            if (p.rank_out) p.rank_out[(int64_t)rg[j] * p.ld_out + c] = rk;
            finish(j, same ? __ldg(S + rk - 1) : mapped_value_interp(p, S, rk, n, m));
        }
    } else {
        // every other case (ties / swapped pairs / T_pred != T_fit / rank instrumentation): the exact
        // comparisons are done, the inputs are no longer needed — stage the sorted words over them
        // and walk the lane's positions in a compact runtime loop (neighbours come from the row).
        uint32_t* Xu = reinterpret_cast<uint32_t*>(myX);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) Xu[rb + e] = v[e].k;
        __syncwarp();
        // pass A: rank of every position (the row keeps the original member in the low bits, which is
        // all a neighbour ever reads; the rank goes into the bits above)
        for (int e = 0; e < E; ++e) {
            const int pos = j0 + e;
            if (pos >= n) break;
            int rk = pos + 1;
            if (mode == 2) {
                const uint32_t open = (~bm_tie & EMASK) >> e;                // exact ties: end of the run
                rk = open ? (pos + __ffs(open)) : tie_carry;
            }
            Xu[rb + e] = (Xu[rb + e] & IDX) | ((uint32_t)rk << LOG);
        }
        __syncwarp();
        // pass B: four positions at a time — the fitted values are fetched together, then finished.
        // T_pred != T_fit: the four CDF positions are located first (arithmetic only), then their eight
        // fitted values are requested together, then interpolated — the loads of a batch overlap
        GroupMap gmap;
        if (!same) {
            gmap = make_group_map(p, n, m);
            if (n > m) fit_group_tails(p, gmap, S, lane);       // ranks outside the fitted positions exist only then
        }
        for (int e0 = 0; e0 < E; e0 += 4) {
            if (j0 + e0 >= n) break;
            uint32_t member[4];
            int rk[4];
            float val[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int e = e0 + d, pos = j0 + e;
                const uint32_t w = Xu[rb + e];
                member[d] = w & IDX;
                rk[d] = (int)(w >> LOG);
                if (mode == 2 && pos < n) {
                    const uint32_t gt_here = (bm_gt >> e) & 1u;
                    const uint32_t gt_prev = (e == 0) ? prev_gt : (bm_gt >> (e - 1)) & 1u;
                    if (gt_here) member[d] = Xu[skew(pos + 1)] & IDX;        // inverted pair: take the next member
                    else if (gt_prev) member[d] = Xu[skew(pos - 1)] & IDX;   // ... and the next position takes this one
                }
                val[d] = (same && pos < n) ? __ldg(S + rk[d] - 1) : 0.0f;
            }
            if (!same) {
                InterpLoc loc[4];
                float y0[4], y1[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    loc[d] = interp_locate(gmap, (j0 + e0 + d < n) ? rk[d] : 1);
                    if (j0 + e0 + d >= n) loc[d].j = 0;
                }
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    y0[d] = loc[d].j > 0 ? __ldg(S + loc[d].j - 1) : 0.0f;
                    y1[d] = loc[d].j > 0 ? __ldg(S + loc[d].j) : 0.0f;
                }
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if (j0 + e0 + d < n)
                        val[d] = loc[d].j > 0 ? interp_value(loc[d], y0[d], y1[d])
                               : loc[d].j == -1 ? (float)(gmap.a_lo * loc[d].q + gmap.b_lo)
                               : loc[d].j == -2 ? (float)(gmap.a_hi * loc[d].q + gmap.b_hi)
                               : mapped_value_general(p, S, rk[d], n, m);
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int pos = j0 + e0 + d;
                if (pos < n) {
                    if (p.rank_out) p.rank_out[(int64_t)rg[member[d]] * p.ld_out + c] = rk[d];
                    finish((int)member[d], val[d]);
                }
            }
        }
    }
}

// coalesced store of one output tile (rows of 8 cells): 16-byte stores when the rows are aligned
template <int E>
__device__ __forceinline__ void store_tile(const uint32_t* tileR, const PredictParams& p, int64_t c0,
                                           const int32_t* rowtab, int n) {
    constexpr int NPS = TileGeom<E>::NPS;
    float* out = (float*)p.out;
    const bool vec = !p.no_vec && ((p.ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(out + c0) & 15) == 0) &&
                     (c0 + TILE_CT <= p.C);
    if (vec) {
        const int quad = threadIdx.x & 1;
        const float* s0 = reinterpret_cast<const float*>(tileR) + (4 * quad) * NPS;
        float* outp = out + c0 + 4 * quad;
#if SDB_STORE_BATCH
        constexpr int IT = TileGeom<E>::NP / (TILE_THREADS / 2);
        const int j0 = threadIdx.x >> 1;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int j = j0 + it * (TILE_THREADS / 2);
            if (j < n) {
                const int at = skew(j);
                float4 v4;
                v4.x = s0[at]; v4.y = s0[NPS + at]; v4.z = s0[2 * NPS + at]; v4.w = s0[3 * NPS + at];
                SDB_ST_ROW(reinterpret_cast<float4*>(outp + (uint64_t)(uint32_t)rowtab[j] * (uint32_t)p.ld_out), v4);
            }
        }
#else
#pragma unroll 8
        for (int j = threadIdx.x >> 1; j < n; j += TILE_THREADS / 2) {
            const int at = skew(j);
            float4 v4;
            v4.x = s0[at]; v4.y = s0[NPS + at]; v4.z = s0[2 * NPS + at]; v4.w = s0[3 * NPS + at];
            SDB_ST_ROW(reinterpret_cast<float4*>(outp + (uint64_t)(uint32_t)rowtab[j] * (uint32_t)p.ld_out), v4);
        }
#endif
    } else {
        const int cc = threadIdx.x & (TILE_CT - 1);
        const int64_t cs = c0 + cc;
        if (cs < p.C) {
            float* outp = out + cs;
            const float* srcp = reinterpret_cast<const float*>(tileR) + cc * NPS;
#pragma unroll 8
            for (int j = threadIdx.x / TILE_CT; j < n; j += TILE_THREADS / TILE_CT)
                SDB_ST_ROW(outp + (uint64_t)(uint32_t)rowtab[j] * (uint32_t)p.ld_out, srcp[skew(j)]);
        }
    }
}

// grid = (cell tiles, groups): one (tile, group) per CTA
#ifndef SDB_PRED_CTAS
#define SDB_PRED_CTAS (24 / TILE_CT)
#endif
template <int E, bool SHIFT>
__global__ void __launch_bounds__(TILE_THREADS, SDB_PRED_CTAS)
qm_predict_tile_kernel(const PredictParams p) {   // SDB_PRED_CTAS: experiment builds
    constexpr int NPS = TileGeom<E>::NPS;
    extern __shared__ uint32_t smem_u[];
    float* tileX = reinterpret_cast<float*>(smem_u);                        // inputs of the group
    uint32_t* tileR = smem_u + TILE_CT * NPS;                               // shift → outputs (float bits)
    int32_t* rowtab = reinterpret_cast<int32_t*>(smem_u + 2 * TILE_CT * NPS);   // row numbers of the group
    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * TILE_CT;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = c0 + warp;
    const bool in_range = c < p.C;
    const bool active = in_range && (!p.valid || p.valid[c]);
    // per-(cell, group) scalars: fetched before the tile load so their latency hides behind it
    const int sg = p.state_gid[g];
    const int m = p.fit_len[sg];
    const int64_t soff = p.state_off[sg];
    const float* S = (const float*)p.state + (active ? c : 0) * p.state_ld + soff;
    // 16-byte loads of the fitted values: aligned, and the group's last quad stays inside the cell's record
    const bool quad_ok = ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && (soff + (((m - 1) & ~3) + 4) <= p.state_ld);
    float xc_f = 0.0f, yc_f = 0.0f;
    if (active) {
        if (SHIFT) xc_f = ((const float*)p.x_climo)[(int64_t)sg * p.ld_climo + c];
        if (p.mode != SDB_MODE_QM && p.return_anoms) yc_f = ((const float*)p.y_climo)[(int64_t)sg * p.ld_climo + c];
        // the fitted sorted values are wanted right after the sort: pull the record into L2 now
        if (lane * 32 < m) asm volatile("prefetch.global.L2 [%0];" :: "l"(S + lane * 32));
    }
    load_tile<E, SDB_PRED_BATCH, SDB_PRED_LD_POLICY>(tileX, (const float*)p.X, p.ld, p.C, c0, rg, n, p.valid, !p.no_vec, rowtab);
    __syncthreads();
    uint32_t* R = tileR + warp * NPS;
    if (in_range && !active) {
        for (int j = lane; j < n; j += 32) R[skew(j)] = __float_as_uint(NAN);
    } else if (active) {
        map_cell_group<E, SHIFT>(p, tileX + warp * NPS, R, lane, c, n, m, rg, S, quad_ok, (double)xc_f, (double)yc_f);
    }
    __syncthreads();
    store_tile<E>(tileR, p, c0, rowtab, n);
}

// ---------------------------------------------------------------- launchers
template <int E>
static int launch_fit_tile(const FitParams& f, cudaStream_t st) {
    auto kern = qm_fit_tile_kernel<E>;
    const size_t smem = fit_tile_smem<E>();
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((f.C + TILE_CT - 1) / TILE_CT), (unsigned)f.n_groups);
#if SDB_TILE_CLUSTER > 1
    grid.x = (grid.x + SDB_TILE_CLUSTER - 1) / SDB_TILE_CLUSTER * SDB_TILE_CLUSTER;    // surplus CTAs exit (c >= C)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TILE_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SDB_TILE_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SDB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, (const float*)f.y, f.ld, f.C, f.rows, f.len, f.off, f.max_len,
                                   (float*)f.state, f.state_ld, f.valid, f.nonfinite, f.no_vec));
#else
    kern<<<grid, TILE_THREADS, smem, st>>>((const float*)f.y, f.ld, f.C, f.rows, f.len, f.off, f.max_len,
                                           (float*)f.state, f.state_ld, f.valid, f.nonfinite, f.no_vec);
#endif
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int E, bool SHIFT>
static int launch_predict_tile(const PredictParams& p, cudaStream_t st) {
    auto kern = qm_predict_tile_kernel<E, SHIFT>;
#ifndef SDB_PRED_EXTRA_SMEM
#define SDB_PRED_EXTRA_SMEM 0            // experiment builds: inflate the footprint to probe occupancy sensitivity
#endif
    const size_t smem = predict_tile_smem<E>() + SDB_PRED_EXTRA_SMEM;
    if (smem > 48 * 1024) SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.C + TILE_CT - 1) / TILE_CT), (unsigned)p.n_groups);
#if SDB_TILE_CLUSTER > 1
    grid.x = (grid.x + SDB_TILE_CLUSTER - 1) / SDB_TILE_CLUSTER * SDB_TILE_CLUSTER;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TILE_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SDB_TILE_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SDB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
#else
    kern<<<grid, TILE_THREADS, smem, st>>>(p);
#endif
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace sdb

// qm_fused.cuh — fit + predict of one (cell tile, time group) in ONE kernel, for the case the headline
// workload is: the prediction index IS the training index (same groups, same lengths), float32, groups of
// up to 1024 steps.  BcsdTemperature.fit + .predict (bcsd.py:197-269), BcsdPrecipitation (bcsd.py:115-185),
// QuantileMapper.fit(X).transform(X') (quantile.py:81-147) without the fitted state ever leaving the SM:
//
//   CTA = 16 warps = 8 consecutive cells x 2 roles (a 32-byte row segment, clusters of 2 CTAs = one
//   64-byte DRAM atom).  Warps 0-7 ("y-warps") each sort one cell's training values into a shared-memory
//   row S; warps 8-15 ("x-warps") each build the rank keys of the same cell's prediction values, rank
//   them, and map member j to S[rank_j - 1] (n == m: the quantile lands exactly on a knot,
//   np.interp returns the order statistic itself — quantile.py:138-139, 523-530).  The two warps of a
//   cell meet at ONE named barrier (S complete).  Both use the counting rank of bm_rank.cuh instead of
//   a sorting network; a series that defeats its quantisation takes the network path of qm_tile.cuh in
//   place (same results).  The sorted values can optionally be written out as the fitted state, so the
//   call leaves a fitted model behind like fit() would.
//
//   Arithmetic of the map is the tile kernel's (qm_tile.cuh) to the bit: float64 rank keys
//   x - (rolling9 - x_climo), shift rounded once to float32, one float32 add of the order statistic.
#pragma once
#include "qm_tile.cuh"
#include "bm_rank.cuh"

namespace sdb {

constexpr int FU_CT = 8;                       // cells per CTA
constexpr int FU_THREADS = 64 * FU_CT;         // 512
constexpr int FU_EPL_Y = 26;                   // y table:  832 entries, 26 624 buckets
constexpr int FU_EPL_X = 34;                   // x table: 1088 entries, 34 816 buckets (its keys are only bracketed: 1/8 margin)
constexpr int FU_QCAP_X = 256;                 // x queue: members of dirty entries
static_assert(TILE_CT == FU_CT, "the fused kernel shares the 8-cell tile loaders");

// shared-memory layout (32-bit words)
struct FuLayout {
    static constexpr int NPS = TileGeom<32>::NPS;
    static constexpr int OFF_TILE_Y = 0;
    static constexpr int OFF_TILE_X = FU_CT * NPS;
    static constexpr int OFF_SCR = 2 * FU_CT * NPS;
    static constexpr int Y_WORDS = BmSortScratch<FU_EPL_Y>::WORDS32;
    // x scratch: W | DE[DEMAX] | DQ[DEMAX] | cnt[4] | Qk uint32[QCAP] | Qs float[QCAP] | Qj uint16[QCAP]
    static constexpr int X_OFF_DE = BmT<FU_EPL_X>::WORDS;
    static constexpr int X_OFF_DQ = X_OFF_DE + BM_DEMAX;
    static constexpr int X_OFF_CNT = X_OFF_DQ + BM_DEMAX;
    static constexpr int X_OFF_QK = X_OFF_CNT + 4;
    static constexpr int X_OFF_QS = X_OFF_QK + FU_QCAP_X;
    static constexpr int X_OFF_QJ = X_OFF_QS + FU_QCAP_X;
    static constexpr int X_WORDS = X_OFF_QJ + FU_QCAP_X / 2;
    static constexpr int Y_WORDS_AL = ((Y_WORDS + 3) / 4) * 4;
    static constexpr int CELL_WORDS = ((Y_WORDS_AL + X_WORDS + 3) / 4) * 4;
    static constexpr int TOTAL_WORDS = OFF_SCR + FU_CT * CELL_WORDS;
    static_assert(OFF_SCR % 4 == 0, "tables need 16-byte alignment");
    static_assert((size_t)TOTAL_WORDS * 4 <= 232448, "one CTA per SM: at most 227 KB of shared memory");
};
constexpr size_t fused_smem_bytes() { return (size_t)FuLayout::TOTAL_WORDS * 4; }

struct FusedParams {
    const float* y; int64_t ld_y;              // training target [T, C]
    const float* X; int64_t ld_x;              // prediction input [T, C]
    int64_t C;
    const int32_t* rows; const int32_t* len; int max_len; int n_groups;
    const float* x_climo; const float* y_climo; int64_t ld_climo;
    int mode; int return_anoms;
    float* out; int64_t ld_out;
    float* state; int64_t state_ld; const int64_t* state_off;     // optional: fitted sorted values
    const uint8_t* valid; int32_t* nonfinite;
    unsigned long long* stats;                 // optional [8]: series, y network path, x network path, y queued, x queued
    int no_vec;
    int force_network;                         // testing: bit 0 y-warps, bit 1 x-warps take the network path
};

__device__ __forceinline__ void fu_cell_barrier(int cell) {
    asm volatile("bar.sync %0, 64;" :: "r"(1 + cell) : "memory");
}

// (key, shift) of member j exactly as the reference computes them (bcsd.py:247-256), from the shared row
__device__ __forceinline__ void fu_window(const float* myX, int n, int j, double xc, double& key, double& shift) {
    double acc = 0.0;
    const int lo = j - 4 < 0 ? 0 : j - 4, hi = j + 4 > n - 1 ? n - 1 : j + 4;
    for (int jj = lo; jj <= hi; ++jj) acc += (double)myX[skew(jj)];
    shift = div_count(acc, hi - lo + 1) - xc;
    key = (double)myX[skew(j)] - shift;
}

// ---------------------------------------------------------------- y-warp: np.sort of the training group → S
__device__ __forceinline__ void fu_sort_cell(const FusedParams& p, float* row, uint32_t* scr, int lane, int64_t c, int g, int n) {
    constexpr int E = 32;
    const int rb = skew(lane * E);
    const int nj = n - lane * E;
    float yv[E];
    bool bad = false;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float x = row[rb + e];
        bad |= (e < nj) && !isfinite(x);
        yv[e] = x + 0.0f;                                   // -0 → +0: np.sort treats them as equal
    }
    if (bad && p.nonfinite) atomicOr(p.nonfinite, 1);
    __syncwarp();                                           // every lane holds its values: the row becomes S
    int queued = 0;
    bool ok = !(p.force_network & 1);
    if (ok) ok = bm_sort_values<E, FU_EPL_Y>(yv, n, lane, row, scr, queued);
    if (!ok) {
        K32 v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) v[e].k = (e < nj) ? f32_to_sortable(yv[e]) : 0xffffffffu;
        sort_blocked<K32, E, 32>(v, lane, nullptr);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) row[lane * E + e] = sortable_to_f32(v[e].k);     // positions >= n: never read
    }
    if (p.stats && lane == 0) {
        atomicAdd(p.stats + 0, 1ull);
        if (!ok) atomicAdd(p.stats + 1, 1ull);
        atomicAdd(p.stats + 3, (unsigned long long)queued);
    }
    __syncwarp();
    if (p.state) {
        float* dst = p.state + c * p.state_ld + p.state_off[g];
        for (int j = lane; j < n; j += 32) dst[j] = row[j];
    }
}

// ---------------------------------------------------------------- x-warp, network path (rare, out of line)
// Exact 64-bit key + position sort of the group (rank_exact64 of qm_tile.cuh): for the series whose keys defeat the
// bucket quantisation.  `park` (the cell's table, unused on this path) receives the parked shifts first, because
// the sort overwrites the inputs with the ranks.  Compact runtime loops: this code must stay small.
template <bool SHIFT>
__device__ __noinline__ void fu_map_cell_network(const FusedParams& p, float* xr, const float* S, float* park, int lane,
                                                 int n, double xc, double yc) {
    constexpr int E = 32;
    const bool ratio = (p.mode == SDB_MODE_BCSD_P) && p.return_anoms;
    const double park_off = (SHIFT && p.return_anoms) ? yc : 0.0;
    if (SHIFT) {
#pragma unroll 1
        for (int j = lane; j < n; j += 32) {
            double key, shift;
            fu_window(xr, n, j, xc, key, shift);
            park[j] = (float)(shift - park_off);
        }
    }
    __syncwarp();
    rank_exact64<E, SHIFT>(xr, n, xc, lane);
    __syncwarp();
    uint32_t* Xu = reinterpret_cast<uint32_t*>(xr);
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
        int rk = (int)Xu[skew(j)];
        rk = rk < 1 ? 1 : (rk > n ? n : rk);
        const float val = S[rk - 1];
        float o;
        if (SHIFT) o = __fadd_rn(park[j], val);
        else o = ratio ? (float)((double)val / yc) : val;
        xr[skew(j)] = o;                                  // each lane overwrites only the slots it has just read
    }
}

// ---------------------------------------------------------------- x-warp: rank keys → rank → S[rank - 1] → output
// The prediction values of the cell are in the shared row xr (skewed, zero halo); the outputs replace them.
//
// 32-bit rank keys.  RAW (QuantileMapper / BcsdPrecipitation): the float32 value itself (order-preserving bit
// pattern) — exact.  SHIFT (BcsdTemperature): the float64 key x - (rolling9 - x_climo) is mapped monotonically to
// t = (key - lo) * scale * 65536 and truncated: qf = bucket << 16 | 16 fraction bits.  qf_a < qf_b ⇒ key_a < key_b,
// so members of a dirty entry with DIFFERENT qf are ordered by qf; two members with the SAME qf (2^-16 of a
// bucket apart, or exactly tied) cannot be told apart: the series then takes the exact 64-bit network path.
template <bool SHIFT>
__device__ __forceinline__ void fu_map_cell(const FusedParams& p, float* xr, const float* S, uint32_t* W, int lane, int cell,
                                            int n, double xc, double yc) {
    constexpr int E = 32;
    using T = BmT<FU_EPL_X>;
    using L = FuLayout;
    constexpr int NBATCH = 8;
    uint32_t* DE = W + L::X_OFF_DE;
    uint32_t* DQ = W + L::X_OFF_DQ;
    uint32_t* cnt = W + L::X_OFF_CNT;
    uint32_t* Qk = W + L::X_OFF_QK;
    float* Qs = reinterpret_cast<float*>(W + L::X_OFF_QS);
    uint16_t* Qj = reinterpret_cast<uint16_t*>(W + L::X_OFF_QJ);
    const uint32_t Wsa = bm_saddr(W);
    const int j0 = lane * E;
    const int nj = n - j0;
    const int rb = skew(j0);
    const bool ratio = (p.mode == SDB_MODE_BCSD_P) && p.return_anoms;
    const double park_off = (SHIFT && p.return_anoms) ? yc : 0.0;

    // ---- 1. own members (+ halo) → registers; rank keys and the parked float32 shift
    uint32_t qf[SHIFT ? E : 1];   // SHIFT: bucket << 16 | fraction
    float sh[E];                  // SHIFT: shift - y_climo (float32), later the output; RAW: the key, later the output
    float lo32, hi32, scale32 = 0.0f;
    {
        constexpr int HL = SHIFT ? 4 : 0, HR = SHIFT ? 5 : 0;
        float xh[E + HL + HR];
#pragma unroll
        for (int i = 0; i < E + HL + HR; ++i) {
            const int e = i - HL;
            xh[i] = xr[rb + e + (e < 0 ? -1 : (e >= 32 ? 1 : 0))];      // zeros outside [0, n) by construction of the tile
        }
        lo32 = INFINITY; hi32 = -INFINITY;
        float nanacc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float x = xh[e + HL];
            lo32 = fminf(lo32, (e < nj) ? x : INFINITY);
            hi32 = fmaxf(hi32, (e < nj) ? x : -INFINITY);
            nanacc = fmaf(x, 0.0f, nanacc);
        }
        if (nanacc != nanacc && p.nonfinite) atomicOr(p.nonfinite, 1);
        bm_warp_minmax(lo32, hi32);
        if (SHIFT) {
            // the bounds are a guess (value range + 1/8 margin): keys outside clamp to the end buckets, which
            // keeps the map monotone — whatever shares an entry with a second element is compared by its full qf
            const double range = (double)hi32 - (double)lo32;
            const double lo = (double)lo32 - 0.125 * range, hi = (double)hi32 + 0.125 * range;
            const double scale = (hi > lo && isfinite(hi - lo)) ? (double)(T::NB - 1) * 65536.0 / (hi - lo) : 0.0;
            const double nls = -lo * scale;
            constexpr uint32_t QFMAX = ((uint32_t)T::NB << 16) - 1u;
            const int back0 = j0 + 5 - n;
            const bool front = (lane == 0);
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) sum += (double)xh[i];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int b = back0 + e;
                b = b < 0 ? 0 : (b > 8 ? 8 : b);
                int cn = 9 - b;
                if (e < 4) { cn -= front ? (4 - e) : 0; cn = cn < 1 ? 1 : cn; }
                const double shift = div_count(sum, cn) - xc;
                const double t = fma((double)xh[e + 4] - shift, scale, nls);
                const uint32_t q = __double2uint_rd(t);     // saturating: negative and NaN map to 0
                qf[e] = q > QFMAX ? QFMAX : q;
                sh[e] = (float)(shift - park_off);
                sum += (double)xh[e + 9];
                sum -= (double)xh[e];
            }
        } else {
            scale32 = bm_scale_f32<T::NB>(lo32, hi32);
#pragma unroll
            for (int e = 0; e < E; ++e) sh[e] = xh[e] + 0.0f;   // -0 == +0 for the rank
        }
    }
    auto q_of = [&](int e) -> uint32_t {
        if constexpr (SHIFT) return qf[e] >> 16; else return bm_bucket_f32<T::NB>(sh[e], lo32, scale32);
    };
    // RAW: values AT the lower bound never enter the table (the zeros of precipitation are one run of hundreds)
    auto on_of = [&](int e) -> int {
        if constexpr (SHIFT) return nj - e; else return (sh[e] > lo32) ? nj - e : 0;
    };

    // ---- 2. table: insert, prefix
    bool ok = !(p.force_network & 2);
    int total = 0, n_dirty = 0, n_queued = 0;
    if (ok) {
        bm_clear<FU_EPL_X>(W, lane);
        if (lane < 2) cnt[lane] = 0u;
        __syncwarp();
#pragma unroll
        for (int b = 0; b < E; b += NBATCH) {
            uint32_t q[NBATCH];
            int on[NBATCH];
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) { q[u] = q_of(b + u); on[u] = on_of(b + u); }
            bm_insert_batch<NBATCH>(Wsa, bm_dummy_off<FU_EPL_X>(lane), q, on);
        }
        __syncwarp();
        bool bad;
        total = bm_prefix<FU_EPL_X, true>(W, lane, DE, DQ, cnt, bad);
        __syncwarp();
        n_dirty = (int)cnt[0];
        n_queued = (int)cnt[1];
        ok = !bad && n_dirty <= BM_DEMAX && n_queued <= FU_QCAP_X;
    }
    fu_cell_barrier(cell);                                   // S (the sorted training values) is complete

    auto final_value = [&](float parked, float val) -> float {
        if (SHIFT) return __fadd_rn(parked, val);            // (shift - y_climo) + mapped value: one float32 add
        return ratio ? (float)((double)val / yc) : val;
    };

    uint32_t dmask = 0;                                      // bit e: member e went through the queue
    if (ok) {
        // ---- 3. look-up: position → order statistic → output (registers); members of dirty entries queue up
        const int n_lo = n - total;
#pragma unroll
        for (int b = 0; b < E; b += NBATCH) {
            uint32_t q[NBATCH], dirty[NBATCH];
            int on[NBATCH], pos[NBATCH], cur[NBATCH];
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) { q[u] = q_of(b + u); on[u] = on_of(b + u); }
            bm_lookup_batch<NBATCH>(W, Wsa, q, on, pos, dirty, cur);
            float val[NBATCH];
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) {
                int at = n_lo + pos[u];
                if (!SHIFT) at = (on[u] > 0) ? at : n_lo - 1;    // ties at the lower bound: highest rank of the run
                at = at < 0 ? 0 : (at > n - 1 ? n - 1 : at);     // padding members (e >= nj) read a valid slot
                val[u] = S[at];
            }
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) {
                const int e = b + u;
                const bool queued = dirty[u] && on[u] > 0;
                if (queued) {
                    Qk[cur[u]] = SHIFT ? qf[SHIFT ? e : 0] : f32_to_sortable(sh[e]);
                    Qs[cur[u]] = sh[e];
                    Qj[cur[u]] = (uint16_t)(j0 + e);
                }
                dmask |= queued ? (1u << e) : 0u;
                sh[e] = final_value(sh[e], val[u]);
            }
        }
        __syncwarp();
        // ---- 4. one lane per dirty entry: rank of every member among the entry's members (ties → highest);
        // entries with more than BM_EMAX members are ranked by the whole warp, one member per lane
        bool tie = false;
        for (int k0 = 0; k0 < n_dirty; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t de = (k < n_dirty) ? DE[k] : 0u;
            const int c = (int)(de >> 16), first = n_lo + (int)(de & 0xffffu);
            const int qs = (k < n_dirty) ? (int)DQ[k] : 0;
            if (c > 0 && c <= BM_EMAX) {
                uint32_t key[BM_EMAX];
#pragma unroll
                for (int i = 0; i < BM_EMAX; ++i) key[i] = (i < c) ? Qk[qs + i] : 0xffffffffu;
#pragma unroll
                for (int i = 0; i < BM_EMAX; ++i) {
                    int r = 0;
#pragma unroll
                    for (int j = 0; j < BM_EMAX; ++j)
                        if (j != i) { r += (key[j] <= key[i]) ? 1 : 0; if (SHIFT && j > i) tie |= (key[j] == key[i]) && (j < c); }
                    if (i < c) Qs[qs + i] = final_value(SHIFT ? Qs[qs + i] : 0.0f, S[first + r]);
                }
            }
            uint32_t bigm = __ballot_sync(0xffffffffu, c > BM_EMAX);
            while (bigm) {
                const int src = __ffs(bigm) - 1;
                bigm &= bigm - 1u;
                const int cb = __shfl_sync(0xffffffffu, c, src), fb = __shfl_sync(0xffffffffu, first, src);
                const int qb = __shfl_sync(0xffffffffu, qs, src);
                const uint32_t mine = (lane < cb) ? Qk[qb + lane] : 0xffffffffu;
                int r = -1;                                            // the member itself is counted by "<="
                for (int j = 0; j < cb; ++j) {
                    const uint32_t o = Qk[qb + j];                     // broadcast read
                    r += (o <= mine) ? 1 : 0;
                    if (SHIFT) tie |= (o == mine) && (j != lane) && (lane < cb);
                }
                if (lane < cb) Qs[qb + lane] = final_value(SHIFT ? Qs[qb + lane] : 0.0f, S[fb + r]);
            }
        }
        if (SHIFT) ok = !__any_sync(0xffffffffu, tie);
        __syncwarp();
    }
    if (p.stats && lane == 0) {
        if (!ok) atomicAdd(p.stats + 2, 1ull);
        atomicAdd(p.stats + 4, (unsigned long long)n_queued);
    }
    if (ok) {
        // the row is no longer read: outputs replace the inputs
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (e < nj && !((dmask >> e) & 1u)) xr[rb + e] = sh[e];
        for (int s = lane; s < n_queued; s += 32) xr[skew((int)Qj[s])] = Qs[s];
    } else {
        fu_map_cell_network<SHIFT>(p, xr, S, reinterpret_cast<float*>(W), lane, n, xc, yc);
    }
}

// coalesced store of the output tile by all 512 threads: 16-byte stores when the rows are aligned
__device__ __forceinline__ void fu_store_tile(const float* tile, const FusedParams& p, int64_t c0, const int32_t* __restrict__ rg, int n) {
    constexpr int NPS = FuLayout::NPS;
    float* out = p.out;
    const bool vec = !p.no_vec && ((p.ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(out + c0) & 15) == 0) && (c0 + FU_CT <= p.C);
    if (vec) {
        const int quad = threadIdx.x & 1;
        const float* s0 = tile + (4 * quad) * NPS;
        float* outp = out + c0 + 4 * quad;
#pragma unroll
        for (int it = 0; it < 1024 / (FU_THREADS / 2); ++it) {
            const int j = (threadIdx.x >> 1) + it * (FU_THREADS / 2);
            if (j < n) {
                const int at = skew(j);
                float4 v4;
                v4.x = s0[at]; v4.y = s0[NPS + at]; v4.z = s0[2 * NPS + at]; v4.w = s0[3 * NPS + at];
                SDB_ST_ROW(reinterpret_cast<float4*>(outp + (uint64_t)(uint32_t)__ldg(rg + j) * (uint32_t)p.ld_out), v4);
            }
        }
    } else {
        const int cc = threadIdx.x & (FU_CT - 1);
        const int64_t cs = c0 + cc;
        if (cs < p.C) {
            float* outp = out + cs;
            const float* srcp = tile + cc * NPS;
            for (int j = threadIdx.x / FU_CT; j < n; j += FU_THREADS / FU_CT)
                SDB_ST_ROW(outp + (uint64_t)(uint32_t)__ldg(rg + j) * (uint32_t)p.ld_out, srcp[skew(j)]);
        }
    }
}

template <bool SHIFT>
__global__ void __launch_bounds__(FU_THREADS, 1)
qm_fused_kernel(const FusedParams p) {
    using L = FuLayout;
    extern __shared__ __align__(16) uint32_t smem_u[];
    float* tileY = reinterpret_cast<float*>(smem_u) + L::OFF_TILE_Y;
    float* tileX = reinterpret_cast<float*>(smem_u) + L::OFF_TILE_X;
    const int g = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * FU_CT;
    const int n = p.len[g];
    const int32_t* rg = p.rows + (int64_t)g * p.max_len;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp >> 3, cell = warp & 7;             // every scheduler gets two y-warps and two x-warps
    const int64_t c = c0 + cell;
    const bool in_range = c < p.C;
    const bool active = in_range && (!p.valid || p.valid[c]);
    float xc_f = 0.0f, yc_f = 0.0f;
    if (role == 1 && active) {
        if (SHIFT) xc_f = p.x_climo[(int64_t)g * p.ld_climo + c];
        if (p.mode != SDB_MODE_QM && p.return_anoms) yc_f = p.y_climo[(int64_t)g * p.ld_climo + c];
    }
    if (role == 0) load_tile<32, SDB_FIT_BATCH, SDB_FIT_LD_POLICY>(tileY, p.y, p.ld_y, p.C, c0, rg, n, p.valid, !p.no_vec, nullptr, (int)threadIdx.x);
    else           load_tile<32, SDB_PRED_BATCH, SDB_PRED_LD_POLICY>(tileX, p.X, p.ld_x, p.C, c0, rg, n, p.valid, !p.no_vec, nullptr, (int)threadIdx.x - 256);
    __syncthreads();
    uint32_t* scr = smem_u + L::OFF_SCR + cell * L::CELL_WORDS;
    if (active) {
        if (role == 0) {
            fu_sort_cell(p, tileY + cell * L::NPS, scr, lane, c, g, n);
            fu_cell_barrier(cell);
        } else {
            fu_map_cell<SHIFT>(p, tileX + cell * L::NPS, tileY + cell * L::NPS, scr + L::Y_WORDS_AL, lane, cell, n, (double)xc_f, (double)yc_f);
        }
    } else if (in_range && role == 1) {
        float* xr = tileX + cell * L::NPS;
        for (int j = lane; j < n; j += 32) xr[skew(j)] = NAN;
    }
    __syncthreads();
    fu_store_tile(tileX, p, c0, rg, n);
}

template <bool SHIFT>
static int launch_fused(const FusedParams& p, cudaStream_t st) {
    auto kern = qm_fused_kernel<SHIFT>;
    const size_t smem = fused_smem_bytes();
    SDB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.C + FU_CT - 1) / FU_CT), (unsigned)p.n_groups);
    grid.x = (grid.x + 1) / 2 * 2;                           // clusters of 2 along x: surplus CTAs have no cell in range
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(FU_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SDB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace sdb

// bm_rank.cuh — counting rank of one (cell, group) series by ONE warp, without a sorting network.
//
// Replaces the two 1024-point compare-exchange networks of the round-1 fit / predict tile kernels
// (np.sort of the training group, quantile.py:462; self-rank of the prediction group,
// quantile.py:138,488) by a bitmap counting rank in shared memory.  Per series the warp owns a
// table W of ENT = 32 * EPL 64-bit entries; entry = {32 presence bits, meta}:
//
//   A  every element quantises its key MONOTONICALLY to a bucket q (32 buckets per entry) and sets
//      presence bit q with one shared-memory atomicOr; an element that finds its bit set
//      ("loser") adds 1 to the entry's meta word instead.
//   B  one pass over W (lane L owns entries EPL*L .. EPL*L+EPL-1, one warp scan): meta := number of
//      elements in lower entries.  An entry that holds a loser is "dirty": it is appended to a short
//      list and its members (usually 2-4) will arrive in ARRIVAL order inside their final range.
//   C  an element of a clean entry has
//          position = prefix(entry) + popc(presence bits below its own)
//      — exact, because the bucket map is monotone and a clean entry holds one element per set bit;
//      a member of a dirty entry takes position prefix(entry) + (arrival number from an atomic cursor).
//   D  one lane per dirty entry puts the entry's members into exact key order (a fixed 8-element
//      network in registers).
//
// The result is the exact rank (ties → highest rank) / the exact sorted order for EVERY input.
// Series that defeat the quantisation (more dirty entries than the list holds or an entry with more
// than BM_EMAX members: heavy ties, a far outlier squeezing the range) make the routines return
// false; the caller then takes the sorting-network path (sort.cuh).
//
// Every per-element shared-memory operation is issued as straight-line, PREDICATED code (inline PTX
// for the atomics): the first version of this file used if-guarded atomicOr / atomicAdd and spent a
// quarter of its instructions on BSSY / BRA / BSYNC around them (profiles/r02_fused_v1_*).
//
// Measured on B200 (tools/ubench_rank.cu, profiles/r02_ubench_rank.txt): a shared-memory atomic costs
// the same pipe time as a bank-conflicted load (5.3 against 4.8 SM-cycles per warp instruction at
// random addresses), i.e. the table is bound by bank conflicts, not by atomics.
#pragma once
#include <cstdint>

namespace sdb {

constexpr int BM_EMAX = 8;           // members of one dirty entry that ONE lane orders in registers
constexpr int BM_EBIG = 32;          // ... that the warp orders together (one member per lane); more → network path
constexpr int BM_DEMAX = 128;        // dirty entries per series

template <int EPL> struct BmT {
    // lane stride 2 * EPL words must be = 4 (mod 8) x 4 bytes apart for conflict-free 16-byte accesses
    static_assert(EPL % 2 == 0 && ((2 * EPL) % 32) % 8 == 4, "lane stride of the table must keep 16-byte accesses conflict-free");
    static constexpr int ENT = 32 * EPL;
    static constexpr int NB = ENT * 32;          // buckets
    static constexpr int WORDS = 2 * (ENT + 32); // 32-bit words, incl. one dummy entry per lane (disabled elements aim there)
};

__device__ __forceinline__ uint32_t bm_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// byte offset of the entry of bucket q inside the table
__device__ __forceinline__ uint32_t bm_eoff(uint32_t q) { return (q >> 2) & 0xfffffff8u; }

// shared-memory atomics on 32-bit shared-space addresses.  ptxas turns a predicated atom.shared into a branch
// around it (BSSY / BRA / ATOMS / BSYNC), so the presence atomic of phase A is UNCONDITIONAL: a disabled
// element aims at its lane's dummy entry behind the table.  The rare ones (loser's add, cursor of a dirty
// entry) stay conditional: skipping the shared-memory operation is worth the three control instructions.
__device__ __forceinline__ uint32_t bm_atom_or(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void bm_red_add_if(uint32_t addr, uint32_t v, uint32_t cond) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(addr), "r"(v), "r"(cond) : "memory");
}
// cursor of a dirty entry: enabled when meta (bit 31 = dirty) is negative AND cond > 0
__device__ __forceinline__ uint32_t bm_atom_add_if_dirty(uint32_t addr, uint32_t v, uint32_t meta, int cond) {
    uint32_t old;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %3, 0;\n\tsetp.gt.and.s32 p, %4, 0, p;\n\tmov.u32 %0, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
                 : "=r"(old) : "r"(addr), "r"(v), "r"(meta), "r"(cond) : "memory");
    return old;
}

template <int EPL>
__device__ __forceinline__ void bm_clear(uint32_t* W, int lane) {
    uint4* p = reinterpret_cast<uint4*>(W + lane * (2 * EPL));
#pragma unroll
    for (int i = 0; i < EPL / 2; ++i) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// phase A for a batch: all presence atomics are issued before the first result is needed.
// on[u] > 0 enables element u; `dummy` = byte offset of the lane's dummy entry.
template <int NBATCH>
__device__ __forceinline__ void bm_insert_batch(uint32_t Wsa, uint32_t dummy, const uint32_t (&q)[NBATCH], const int (&on)[NBATCH]) {
    uint32_t hit[NBATCH], off[NBATCH];
#pragma unroll
    for (int u = 0; u < NBATCH; ++u) {
        const uint32_t bit = 1u << (q[u] & 31u);
        off[u] = on[u] > 0 ? bm_eoff(q[u]) : dummy;
        hit[u] = bm_atom_or(Wsa + off[u], bit) & bit;
    }
#pragma unroll
    for (int u = 0; u < NBATCH; ++u) bm_red_add_if(Wsa + off[u] + 4u, 1u, on[u] > 0 ? hit[u] : 0u);
}
template <int EPL>
__device__ __forceinline__ uint32_t bm_dummy_off(int lane) { return (uint32_t)(BmT<EPL>::ENT + lane) * 8u; }

// phase B.  meta after the pass: bits 0-15 = elements in lower entries; dirty entries: bit 31 set, bits 16-30 =
// arrival cursor (starts at 0).  Dirty entries are appended to DE as (first position | members << 16);
// cnt[0] (zeroed by the caller before phase A) counts them.  QUEUE: every dirty entry also reserves `members`
// consecutive slots of a compact queue (cnt[1] = slots handed out so far); the first slot is the initial value
// of the entry's cursor and is recorded in DQ.  Returns the number of inserted elements; `bad` is set when an
// entry holds more than BM_EBIG members.
template <int EPL, bool QUEUE>
__device__ __forceinline__ int bm_prefix(uint32_t* W, int lane, uint32_t* DE, uint32_t* DQ, uint32_t* cnt, bool& bad) {
    uint4* base = reinterpret_cast<uint4*>(W + lane * (2 * EPL));
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < EPL / 2; ++i) {
        const uint4 t = base[i];
        run += __popc(t.x) + t.y + __popc(t.z) + t.w;
    }
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = (int)__shfl_sync(0xffffffffu, incl, 31);
    uint32_t p = incl - run;
    bool big = false;
#pragma unroll
    for (int i = 0; i < EPL / 2; ++i) {
        uint4 t = base[i];
        const uint32_t c0 = __popc(t.x) + t.y, c1 = __popc(t.z) + t.w;
        if (t.y | t.w) {                                   // rare: a dirty entry in this pair
            uint32_t q0 = 0, q1 = 0;
            if (t.y) {
                const uint32_t k = atomicAdd(cnt, 1u);
                if (QUEUE) q0 = atomicAdd(cnt + 1, c0);
                if (k < BM_DEMAX) { DE[k] = p | (c0 << 16); if (QUEUE) DQ[k] = q0; }
                big |= c0 > BM_EBIG;
                t.y = 0x80000000u | (q0 << 16);
            }
            if (t.w) {
                const uint32_t k = atomicAdd(cnt, 1u);
                if (QUEUE) q1 = atomicAdd(cnt + 1, c1);
                if (k < BM_DEMAX) { DE[k] = (p + c0) | (c1 << 16); if (QUEUE) DQ[k] = q1; }
                big |= c1 > BM_EBIG;
                t.w = 0x80000000u | (q1 << 16);
            }
        }
        t.y |= p;                                          // clean entries: meta (the loser count) was 0
        p += c0;
        t.w |= p;
        p += c1;
        base[i] = t;
    }
    bad = __any_sync(0xffffffffu, big);
    return total;
}

// phase C for a batch.  Clean entry: pos[u] = 0-based position among the inserted elements, dirty[u] = 0.
// Dirty entry: dirty[u] = 1, pos[u] = prefix of the entry, cur[u] = the value drawn from the entry's cursor
// (arrival number, plus the entry's first queue slot when the table was built with QUEUE).
template <int NBATCH>
__device__ __forceinline__ void bm_lookup_batch(const uint32_t* W, uint32_t Wsa, const uint32_t (&q)[NBATCH], const int (&on)[NBATCH],
                                                int (&pos)[NBATCH], uint32_t (&dirty)[NBATCH], int (&cur)[NBATCH]) {
    uint2 v[NBATCH];
#pragma unroll
    for (int u = 0; u < NBATCH; ++u) v[u] = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(W) + bm_eoff(q[u]));
    uint32_t arr[NBATCH];
#pragma unroll
    for (int u = 0; u < NBATCH; ++u) arr[u] = bm_atom_add_if_dirty(Wsa + bm_eoff(q[u]) + 4u, 0x10000u, v[u].y, on[u]);
#pragma unroll
    for (int u = 0; u < NBATCH; ++u) {
        dirty[u] = v[u].y >> 31;
        const int clean_pos = __popc(v[u].x & ((1u << (q[u] & 31u)) - 1u));
        cur[u] = (int)((arr[u] >> 16) & 0x7fffu);
        pos[u] = (int)(v[u].y & 0xffffu) + (dirty[u] ? 0 : clean_pos);
    }
}

// monotone bucket of a float32 key at or above the lower bound
template <int NB>
__device__ __forceinline__ uint32_t bm_bucket_f32(float x, float lo, float scale) {
    const uint32_t b = __float2uint_rd((x - lo) * scale);          // saturating; NaN → 0
    return b > (uint32_t)(NB - 1) ? (uint32_t)(NB - 1) : b;
}
template <int NB>
__device__ __forceinline__ float bm_scale_f32(float lo, float hi) {
    const float range = hi - lo;
    return (range > 0.0f && range < INFINITY) ? (float)(NB - 1) / range : 0.0f;
}

__device__ __forceinline__ void bm_warp_minmax(float& lo, float& hi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
}

// ascending sort of 8 floats (odd-even merge sort, 19 compare-exchanges)
__device__ __forceinline__ void bm_sort8(float (&v)[8]) {
#define BM_CE(a, b) { const float lo_ = fminf(v[a], v[b]), hi_ = fmaxf(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
    BM_CE(0, 1) BM_CE(2, 3) BM_CE(4, 5) BM_CE(6, 7)
    BM_CE(0, 2) BM_CE(1, 3) BM_CE(4, 6) BM_CE(5, 7)
    BM_CE(1, 2) BM_CE(5, 6)
    BM_CE(0, 4) BM_CE(1, 5) BM_CE(2, 6) BM_CE(3, 7)
    BM_CE(2, 4) BM_CE(3, 5)
    BM_CE(1, 2) BM_CE(3, 4) BM_CE(5, 6)
#undef BM_CE
}

// ---------------------------------------------------------------- np.sort of one series (fit)
// scratch layout (32-bit words): W[WORDS] | DE[BM_DEMAX] | cnt[4]
template <int EPL> struct BmSortScratch {
    static constexpr int OFF_DE = BmT<EPL>::WORDS;
    static constexpr int OFF_CNT = OFF_DE + BM_DEMAX;
    static constexpr int WORDS32 = OFF_CNT + 4;
};

// In: the lane's E values yv[e] (members lane*E + e < n are real, -0 already folded into +0, finite).
// Out: sorted values at S[0..n).  S may alias the row the values were read from (every lane holds its
// values in registers and the warp is synchronised inside before the first write).  Returns false when
// the series must take the network path (S may hold garbage then; yv is untouched).
template <int E, int EPL>
__device__ __forceinline__ bool bm_sort_values(const float (&yv)[E], int n, int lane, float* S, uint32_t* W, int& n_dirty) {
    using T = BmT<EPL>;
    using L = BmSortScratch<EPL>;
    uint32_t* DE = W + L::OFF_DE;
    uint32_t* cnt = W + L::OFF_CNT;
    const uint32_t Wsa = bm_saddr(W);
    const int nj = n - lane * E;
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        lo = fminf(lo, (e < nj) ? yv[e] : INFINITY);
        hi = fmaxf(hi, (e < nj) ? yv[e] : -INFINITY);
    }
    bm_warp_minmax(lo, hi);
    const float scale = bm_scale_f32<T::NB>(lo, hi);
    bm_clear<EPL>(W, lane);
    if (lane == 0) cnt[0] = 0u;
    __syncwarp();
    constexpr int NBATCH = 8;
    // values AT the lower bound never enter the table (zero-inflated precipitation: one run of hundreds)
#pragma unroll
    for (int b = 0; b < E; b += NBATCH) {
        uint32_t q[NBATCH];
        int on[NBATCH];
#pragma unroll
        for (int u = 0; u < NBATCH; ++u) {
            q[u] = bm_bucket_f32<T::NB>(yv[b + u], lo, scale);
            on[u] = (yv[b + u] > lo) ? nj - (b + u) : 0;
        }
        bm_insert_batch<NBATCH>(Wsa, bm_dummy_off<EPL>(lane), q, on);
    }
    __syncwarp();
    bool bad;
    const int total = bm_prefix<EPL, false>(W, lane, DE, nullptr, cnt, bad);
    __syncwarp();
    n_dirty = (int)cnt[0];
    if (bad || n_dirty > BM_DEMAX) return false;
    const int n_lo = n - total;                      // the values equal to the lower bound come first
#pragma unroll
    for (int b = 0; b < E; b += NBATCH) {
        uint32_t q[NBATCH], dirty[NBATCH];
        int on[NBATCH], pos[NBATCH], cur[NBATCH];
#pragma unroll
        for (int u = 0; u < NBATCH; ++u) {
            q[u] = bm_bucket_f32<T::NB>(yv[b + u], lo, scale);
            on[u] = (yv[b + u] > lo) ? nj - (b + u) : 0;
        }
        bm_lookup_batch<NBATCH>(W, Wsa, q, on, pos, dirty, cur);
#pragma unroll
        for (int u = 0; u < NBATCH; ++u)
            if (on[u] > 0) S[n_lo + pos[u] + cur[u]] = yv[b + u];      // cur = 0 for clean entries
    }
    __syncwarp();
    // one lane per dirty entry: its members sit in arrival order inside their final range → order them.
    // Entries with more than BM_EMAX members (the dense end of a skewed distribution) are ordered by the
    // whole warp, one member per lane, one entry after the other.
    for (int k0 = 0; k0 < n_dirty; k0 += 32) {
        const int k = k0 + lane;
        const uint32_t de = (k < n_dirty) ? DE[k] : 0u;
        float* r = S + n_lo + (int)(de & 0xffffu);
        const int c = (int)(de >> 16);
        if (c > 0 && c <= BM_EMAX) {
            float v[BM_EMAX];
#pragma unroll
            for (int i = 0; i < BM_EMAX; ++i) v[i] = (i < c) ? r[i] : INFINITY;
            bm_sort8(v);
#pragma unroll
            for (int i = 0; i < BM_EMAX; ++i)
                if (i < c) r[i] = v[i];
        }
        uint32_t bigm = __ballot_sync(0xffffffffu, c > BM_EMAX);
        while (bigm) {
            const int src = __ffs(bigm) - 1;
            bigm &= bigm - 1u;
            const uint32_t deb = __shfl_sync(0xffffffffu, de, src);
            float* rb = S + n_lo + (int)(deb & 0xffffu);
            const int cb = (int)(deb >> 16);
            const float mine = (lane < cb) ? rb[lane] : INFINITY;
            int rank = 0;
            for (int j = 0; j < cb; ++j) {
                const float o = rb[j];                                 // broadcast read
                rank += (o < mine || (o == mine && j < lane)) ? 1 : 0;
            }
            __syncwarp();
            if (lane < cb) rb[rank] = mine;
            __syncwarp();
        }
    }
    for (int i = lane; i < n_lo; i += 32) S[i] = lo;
    return true;
}

}  // namespace sdb

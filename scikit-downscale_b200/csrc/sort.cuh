// sort.cuh — register-blocked bitonic sorting for one time-series group (sm_100a).
//
// A "sorting group" of NT threads (NT = 32: one warp, shuffles only; NT > 32: several
// warps of one CTA, the wide stages go through shared memory) sorts NP = E * NT items,
// E per thread in registers, BLOCKED layout: thread t owns positions t*E .. t*E+E-1.
// The network is the "mirror" form of bitonic sort: the first stage of every merge
// level pairs i with i ^ (k-1), the rest are plain half-cleaners i ^ j, so every
// compare-exchange is ascending (min to the lower position) and needs no direction flag.
// Stages whose partner lives in the same thread are straight register min/max; the
// per-thread part of every merge level is the same code, so only the local sort and
// ONE merge body are unrolled (small instruction footprint), the level loop is a
// runtime loop.
//
// Used for: np.sort of a group's training values (quantile.py:462) and the self-rank of
// a group's prediction values (quantile.py:138,488) — see qm_kernels.cu.
#pragma once
#include <cstdint>
#include <type_traits>

namespace sdb {

// ---------------------------------------------------------------- item types
struct K32  { uint32_t k; };                 // 32-bit sortable key
struct K64  { uint32_t hi, lo; };            // 64-bit sortable key
struct K32I { uint32_t k, i; };              // key + original position
struct K64I { uint32_t hi, lo, i; };         // 64-bit key + original position

__device__ __forceinline__ bool item_less(const K32& a, const K32& b)   { return a.k < b.k; }
__device__ __forceinline__ bool item_less(const K32I& a, const K32I& b) { return a.k < b.k; }
__device__ __forceinline__ bool item_less(const K64& a, const K64& b) {
    return (a.hi < b.hi) || (a.hi == b.hi && a.lo < b.lo);
}
__device__ __forceinline__ bool item_less(const K64I& a, const K64I& b) {
    return (a.hi < b.hi) || (a.hi == b.hi && a.lo < b.lo);
}
__device__ __forceinline__ bool key_equal(const K32& a, const K32& b)   { return a.k == b.k; }
__device__ __forceinline__ bool key_equal(const K32I& a, const K32I& b) { return a.k == b.k; }
__device__ __forceinline__ bool key_equal(const K64& a, const K64& b)   { return a.hi == b.hi && a.lo == b.lo; }
__device__ __forceinline__ bool key_equal(const K64I& a, const K64I& b) { return a.hi == b.hi && a.lo == b.lo; }

template <class I> struct item_words;
template <> struct item_words<K32>  { static constexpr int value = 1; };
template <> struct item_words<K64>  { static constexpr int value = 2; };
template <> struct item_words<K32I> { static constexpr int value = 2; };
template <> struct item_words<K64I> { static constexpr int value = 3; };

__device__ __forceinline__ uint32_t get_word(const K32& a, int)  { return a.k; }
__device__ __forceinline__ void set_word(K32& a, int, uint32_t v) { a.k = v; }
__device__ __forceinline__ uint32_t get_word(const K64& a, int w) { return w == 0 ? a.hi : a.lo; }
__device__ __forceinline__ void set_word(K64& a, int w, uint32_t v) { if (w == 0) a.hi = v; else a.lo = v; }
__device__ __forceinline__ uint32_t get_word(const K32I& a, int w) { return w == 0 ? a.k : a.i; }
__device__ __forceinline__ void set_word(K32I& a, int w, uint32_t v) { if (w == 0) a.k = v; else a.i = v; }
__device__ __forceinline__ uint32_t get_word(const K64I& a, int w) { return w == 0 ? a.hi : (w == 1 ? a.lo : a.i); }
__device__ __forceinline__ void set_word(K64I& a, int w, uint32_t v) { if (w == 0) a.hi = v; else if (w == 1) a.lo = v; else a.i = v; }

template <class I>
__device__ __forceinline__ I item_shfl_xor(const I& v, int mask) {
    I o;
#pragma unroll
    for (int w = 0; w < item_words<I>::value; ++w)
        set_word(o, w, __shfl_xor_sync(0xffffffffu, get_word(v, w), mask));
    return o;
}

// Pipe balancing (SDB_HYBRID bit mask, experiment builds): VIMNMX runs on the ALU pipe only (one warp
// instruction per 2 cycles per scheduler), which is what bounds the sorting network.  The "hybrid"
// compare-exchange takes min on the ALU pipe and max = a + b - min as two IMADs on the FMA pipe (exact
// modulo 2^32).  The multipliers come from constant memory so that ptxas cannot fold the IMADs back
// into IADD3 / VIMNMX.  bit 0: thread-local sort, bit 1: thread-local merge, bit 2: cross-lane keep.
#ifndef SDB_HYBRID
#define SDB_HYBRID 2
#endif
static __constant__ uint32_t SDB_MULS[4] = {1u, 0xffffffffu, 0xfffffffeu, 0u};   // 1, -1, -2
__device__ __forceinline__ void cmpswap_hybrid(K32& a, K32& b) {
    const uint32_t lo = min(a.k, b.k);
    const uint32_t t = a.k * SDB_MULS[0] + b.k;
    b.k = lo * SDB_MULS[1] + t;
    a.k = lo;
}
// ascending compare-exchange inside one thread
__device__ __forceinline__ void cmpswap(K32& a, K32& b) {
    uint32_t lo = min(a.k, b.k), hi = max(a.k, b.k);
    a.k = lo; b.k = hi;
}
template <class I>
__device__ __forceinline__ void cmpswap(I& a, I& b) {
    bool s = item_less(b, a);
    I t = a;
    if (s) { a = b; b = t; }
}
// keep the smaller (lower==true) or the larger of own item v and partner item o
__device__ __forceinline__ K32 keep(const K32& v, const K32& o, bool lower) {
    K32 r;
#if !(SDB_HYBRID & 4) && !defined(SDB_KEEP_PLAIN)
    // two complementary predicated VIMNMX writing the SAME register: the result replaces v in place, so
    // the runtime stage loops carry no register copies (the "min, then predicated max over it" form
    // needs a temporary and costs one MOV per item per stage)
    r.k = v.k;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p min.u32 %0, %0, %1;\n\t@!p max.u32 %0, %0, %1;\n\t}"
        : "+r"(r.k) : "r"(o.k), "r"((int)lower));
    return r;
#endif
    r.k = min(v.k, o.k);                 // min, then a predicated max over it: two VIMNMX, no select
#if SDB_HYBRID & 4
    const uint32_t d = r.k * SDB_MULS[2] + (v.k * SDB_MULS[0] + o.k);      // max - min
    r.k = d * (lower ? 0u : 1u) + r.k;
#else
    if (!lower) r.k = max(v.k, o.k);
#endif
    return r;
}
template <class I>
__device__ __forceinline__ I keep(const I& v, const I& o, bool lower) {
    bool take = lower ? item_less(o, v) : item_less(v, o);
    return take ? o : v;
}

// ---------------------------------------------------------------- thread-local pieces
// sort the E items of one thread: Batcher's odd-even merge sort (191 compare-exchanges for
// E = 32 against 240 for the bitonic network); the network is spelled out in sort_net.inc so
// every register index is a literal
#include "sort_net.inc"
template <class I, int E>
__device__ __forceinline__ void local_sort(I (&v)[E]) {
    static_assert(E == 8 || E == 16 || E == 32, "sorting networks are generated for 8, 16 and 32 items");
#if SDB_HYBRID & 1
#define SDB_CE(a, b) if constexpr (std::is_same<I, K32>::value) cmpswap_hybrid((K32&)v[a], (K32&)v[b]); else cmpswap(v[a], v[b]);
#else
#define SDB_CE(a, b) cmpswap(v[a], v[b]);
#endif
    if constexpr (E == 8) { SDB_SORT_NET_8 }
    else if constexpr (E == 16) { SDB_SORT_NET_16 }
    else { SDB_SORT_NET_32 }
#undef SDB_CE
}
// the thread-local half-cleaners j = E/2 .. 1 that finish every wider merge level
template <class I, int E>
__device__ __forceinline__ void local_merge(I (&v)[E]) {
#pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int p = e ^ j;
#if SDB_HYBRID & 2
            if constexpr (std::is_same<I, K32>::value) { if (p > e) cmpswap_hybrid((K32&)v[e], (K32&)v[p]); } else
#endif
            if (p > e) cmpswap(v[e], v[p]);
        }
    }
}

// ---------------------------------------------------------------- cross-thread stages
// Exchange with thread (tid ^ tmask); `mirror` pairs element e with E-1-e of the partner.
// xchg: shared scratch of NT*E*words uint32 for this sorting group (NT > 32 only).
template <class I, int E, int NT, bool MIRROR>
__device__ __forceinline__ void cross_stage(I (&v)[E], int tid, int tmask, bool lower, uint32_t* xchg) {
    if (NT <= 32 || tmask < 32) {
#pragma unroll
        for (int e = 0; e < E / 2; ++e) {
            // handle the pair (e, E-1-e) together so MIRROR needs no temporary copy of v
            const int f = E - 1 - e;
            I oe = item_shfl_xor(MIRROR ? v[f] : v[e], tmask);
            I of = item_shfl_xor(MIRROR ? v[e] : v[f], tmask);
            v[e] = keep(v[e], oe, lower);
            v[f] = keep(v[f], of, lower);
        }
    } else {
        constexpr int W = item_words<I>::value;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int w = 0; w < W; ++w) xchg[(w * E + e) * NT + tid] = get_word(v[e], w);
        __syncthreads();
        const int pt = tid ^ tmask;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pe = MIRROR ? (E - 1 - e) : e;
            I o;
#pragma unroll
            for (int w = 0; w < W; ++w) set_word(o, w, xchg[(w * E + pe) * NT + pt]);
            v[e] = keep(v[e], o, lower);
        }
    }
}

// Full sort of NP = E*NT items, ascending, blocked layout.  All NT threads of the group
// (and, when NT > 32, all threads of the CTA) must call it together.
template <class I, int E, int NT>
__device__ __forceinline__ void sort_blocked(I (&v)[E], int tid, uint32_t* xchg) {
    local_sort<I, E>(v);
#ifndef SDB_SORT_UNROLL
#define SDB_SORT_UNROLL 0
#endif
    if constexpr (SDB_SORT_UNROLL != 0 && NT == 32 && item_words<I>::value == 1) {
        // keys-only warp sort: merge levels unrolled (no register shuffling at loop back-edges)
#pragma unroll
        for (int kt = 2; kt <= NT; kt <<= 1) {
            cross_stage<I, E, NT, true>(v, tid, kt - 1, (tid & (kt >> 1)) == 0, xchg);
            if constexpr (SDB_SORT_UNROLL == 1) {
#pragma unroll
                for (int jt = kt >> 2; jt > 0; jt >>= 1)
                    cross_stage<I, E, NT, false>(v, tid, jt, (tid & jt) == 0, xchg);
            } else {
#pragma unroll 1
                for (int jt = kt >> 2; jt > 0; jt >>= 1)
                    cross_stage<I, E, NT, false>(v, tid, jt, (tid & jt) == 0, xchg);
            }
            local_merge<I, E>(v);
        }
        return;
    }
#pragma unroll 1
    for (int kt = 2; kt <= NT; kt <<= 1) {          // kt = k / E : merge level in units of threads
        cross_stage<I, E, NT, true>(v, tid, kt - 1, (tid & (kt >> 1)) == 0, xchg);
#pragma unroll 1
        for (int jt = kt >> 2; jt > 0; jt >>= 1)
            cross_stage<I, E, NT, false>(v, tid, jt, (tid & jt) == 0, xchg);
        local_merge<I, E>(v);
    }
}

// After sort_blocked: 1-based rank of every sorted position with ties taking the HIGHEST
// rank, i.e. r[e] = 1 + (last position whose key equals the key at this position).
// `scratch` (NT > 32 only): 2*NT uint32 + item scratch, see callers.
template <class I, int E, int NT>
__device__ __forceinline__ void tie_max_ranks(const I (&v)[E], int tid, int (&r)[E], uint32_t* scratch) {
    static_assert(E <= 32, "run-boundary bitmap is 32 bits");
    I nxt;
    if (NT <= 32) {
#pragma unroll
        for (int w = 0; w < item_words<I>::value; ++w)
            set_word(nxt, w, __shfl_down_sync(0xffffffffu, get_word(v[0], w), 1));
    } else {
        constexpr int W = item_words<I>::value;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < W; ++w) scratch[w * NT + tid] = get_word(v[0], w);
        __syncthreads();
        const int nt = (tid + 1 < NT) ? tid + 1 : tid;
#pragma unroll
        for (int w = 0; w < W; ++w) set_word(nxt, w, scratch[w * NT + nt]);
    }
    uint32_t bm = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        bool last = (e == E - 1) ? ((tid == NT - 1) || !key_equal(v[e], nxt)) : !key_equal(v[e], v[e + 1]);
        bm |= last ? (1u << e) : 0u;
    }
    const int base = tid * E;
    const int minb = base + __ffs(bm);                 // valid when bm != 0 (1-based → "+1" included)
    int carry;
    if (NT <= 32) {
        uint32_t has = __ballot_sync(0xffffffffu, bm != 0);
        uint32_t higher = (tid == 31) ? 0u : (has & ~((2u << tid) - 1u));
        int src = higher ? (__ffs(higher) - 1) : tid;
        carry = __shfl_sync(0xffffffffu, minb, src);
    } else {
        __syncthreads();
        scratch[tid] = bm ? (uint32_t)minb : 0u;
        __syncthreads();
        carry = 0;
        for (int t = tid + 1; t < NT; ++t) {
            uint32_t m = scratch[t];
            if (m) { carry = (int)m; break; }
        }
    }
    int cur = carry;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        if ((bm >> e) & 1u) cur = base + e + 1;
        r[e] = cur;
    }
}

// ---------------------------------------------------------------- order-preserving key maps
__device__ __forceinline__ uint32_t f32_to_sortable(float x) {
    uint32_t u = __float_as_uint(x);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float sortable_to_f32(uint32_t s) {
    uint32_t u = s ^ ((s & 0x80000000u) ? 0x80000000u : 0xffffffffu);
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t f64_to_sortable(double x) {
    uint64_t u = (uint64_t)__double_as_longlong(x);
    return u ^ ((uint64_t)((int64_t)u >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double sortable_to_f64(uint64_t s) {
    uint64_t u = s ^ ((s & 0x8000000000000000ull) ? 0x8000000000000000ull : 0xffffffffffffffffull);
    return __longlong_as_double((long long)u);
}

}  // namespace sdb

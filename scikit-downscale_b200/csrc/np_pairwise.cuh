// np_pairwise.cuh — numpy's pairwise summation, iterative (no device recursion).
//
// numpy/_core/src/umath/loops_utils.h.src (*_pairwise_sum): n < 8 plain loop; n <= 128 eight
// interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential
// remainder; larger n split at n/2 rounded down to a multiple of 8, left sum + right sum.
// ndarray.sum along an axis copies the first element as the initial value and reduces the
// rest pairwise.  This is the arithmetic behind DataFrame.mean() (groupers.py:84-89) and
// ndarray.mean/std (gard.py:331-346) in the reference.
#pragma once

namespace sdb {

template <typename F, typename Get>
__device__ __forceinline__ F np_pairwise_base(const Get& get, int lo, int n) {
    if (n < 8) {
        F res = (F)0;
        for (int i = 0; i < n; ++i) res += get(lo + i);
        return res;
    }
    F r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = get(lo + k);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] += get(lo + i + k);
    }
    F res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += get(lo + i);
    return res;
}

template <typename F, typename Get>
__device__ F np_pairwise(const Get& get, int lo0, int n0) {
    constexpr int DEPTH = 24;
    int lo[DEPTH], n[DEPTH], state[DEPTH];
    F left[DEPTH];
    int sp = 0;
    lo[0] = lo0; n[0] = n0; state[0] = 0;
    F val = (F)0;
    bool have = false;
    while (true) {
        if (!have) {
            if (n[sp] <= 128) { val = np_pairwise_base<F>(get, lo[sp], n[sp]); have = true; }
            else {
                int n2 = n[sp] / 2; n2 -= n2 % 8;
                state[sp] = 1;
                lo[sp + 1] = lo[sp]; n[sp + 1] = n2; state[sp + 1] = 0; ++sp;
            }
        } else {
            if (sp == 0) return val;
            --sp;
            if (state[sp] == 1) {
                left[sp] = val; state[sp] = 2;
                int n2 = n[sp] / 2; n2 -= n2 % 8;
                lo[sp + 1] = lo[sp] + n2; n[sp + 1] = n[sp] - n2; state[sp + 1] = 0; ++sp;
                have = false;
            } else {
                val = left[sp] + val;
            }
        }
    }
}

// ndarray.sum along an axis: first element is the initial value, pairwise over the rest
template <typename F, typename Get>
__device__ F np_sum(const Get& get, int n) {
    F s = get(0);
    if (n > 1) s = s + np_pairwise<F>(get, 1, n - 1);
    return s;
}

}  // namespace sdb

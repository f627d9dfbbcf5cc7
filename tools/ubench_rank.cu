// ubench_rank.cu — standalone microbenchmark: sorting one ~930-value series per warp in shared memory,
// (a) the register bitonic / odd-even network of sort.cuh (what round 1 shipped) against
// (b) the bitmap counting rank of bm_rank.cuh, plus raw shared-memory atomic / load throughput.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I scikit-downscale_b200/csrc \
//        tools/ubench_rank.cu -o gpurun_out/ubench_rank && gpurun_out/ubench_rank
// Prints SM-cycles per series (at the SM clock reported by the device) for each variant.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include "sort.cuh"
#include "bm_rank.cuh"

using namespace sdb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int E = 32;
constexpr int NPS = 1092;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float gauss(uint32_t seed) {
    const float u1 = (hash32(seed * 2u + 1u) >> 8) * (1.0f / 16777216.0f) + 1e-7f;
    const float u2 = (hash32(seed * 2u + 2u) >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
}

// mode 0: copy only; 1: network sort of y; 2: bitmap sort of y; 3: two network sorts (proxy of round-1 fit + predict);
// 4: bitmap sort of y + bitmap self-rank of x + gather out[j] = S[rank_j - 1] (the QuantileMapper task)
template <int MODE>
__global__ void __launch_bounds__(256, 2) sort_bench(int n, int reps, int zero_frac_pct, uint32_t* check, int* fallbacks) {
    extern __shared__ uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int SCR = (MODE == 2 || MODE == 4) ? BM_SCRATCH_WORDS : 0;
    float* pristine = reinterpret_cast<float*>(smem) + warp * (3 * NPS + SCR);
    float* row = pristine + NPS;          // y / S
    float* xrow = row + NPS;              // x / out
    uint32_t* scratch = reinterpret_cast<uint32_t*>(xrow + NPS);
    const uint32_t gw = blockIdx.x * 8 + warp;
    for (int j = lane; j < 1024; j += 32) {
        float v = 15.0f + 3.0f * gauss(gw * 1024u + j) + 4.0f * (float)j / (float)n;
        if (zero_frac_pct > 0) v = (hash32(gw * 4096u + j + 77u) % 100u < (uint32_t)zero_frac_pct) ? 0.0f : fabsf(v - 14.0f);
        pristine[j] = (j < n) ? v : 0.0f;
    }
    __syncwarp();
    uint32_t acc = 0;
    int fb = 0;
    for (int r = 0; r < reps; ++r) {
        float yv[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {                       // rotate the members so the loop is not hoisted
            int src = lane * E + e + r;
            src = src >= n ? src - n : src;
            yv[e] = (lane * E + e < n) ? pristine[src] : 0.0f;
        }
        __syncwarp();
        if (MODE == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) row[lane * E + e + (lane)] = yv[e];
        } else if (MODE == 1 || MODE == 3) {
            K32 v[E];
            const int nj = n - lane * E;
#pragma unroll
            for (int e = 0; e < E; ++e) v[e].k = (e < nj) ? f32_to_sortable(yv[e] + 0.0f) : 0xffffffffu;
            sort_blocked<K32, E, 32>(v, lane, nullptr);
#pragma unroll
            for (int e = 0; e < E; ++e) row[lane * E + e + lane] = sortable_to_f32(v[e].k);
            if (MODE == 3) {
                __syncwarp();
#pragma unroll
                for (int e = 0; e < E; ++e) v[e].k = (e < nj) ? f32_to_sortable(row[(lane * E + e + 7) % n] + 1.0f) : 0xffffffffu;
                sort_blocked<K32, E, 32>(v, lane, nullptr);
#pragma unroll
                for (int e = 0; e < E; ++e) xrow[lane * E + e + lane] = sortable_to_f32(v[e].k);
            }
        } else {
            const bool ok = bm_sort_values<E>(yv, n, lane, row, scratch);
            fb += ok ? 0 : 1;
            if (MODE == 4) {
                // x = a different rotation of the same multiset, shifted
                float xv[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    int src = lane * E + e + r + 311;
                    src = src >= n ? src - n : src;
                    src = src >= n ? src - n : src;
                    xv[e] = (lane * E + e < n) ? pristine[src] + 1.5f : 0.0f;
                    xrow[lane * E + e + lane] = xv[e];
                }
                __syncwarp();
                auto key_of = [&](int j) -> float { return xrow[j + (j >> 5)]; };
                auto fin = [&](int j, int pos) { xrow[j + (j >> 5)] = row[pos]; };
                const bool ok2 = bm_rank_raw<E>(xv, n, lane, scratch, key_of, fin);
                fb += ok2 ? 0 : 1;
                __syncwarp();
                if (r == reps - 1 && blockIdx.x == 0 && ok && ok2) {
                    // brute-force check: out[j] == S[#{i : x_i <= x_j} - 1]
                    int bad = 0;
                    for (int j = lane; j < n; j += 32) {
                        int s1 = j + r + 311; s1 = s1 >= n ? s1 - n : s1; s1 = s1 >= n ? s1 - n : s1;
                        const float xj = pristine[s1] + 1.5f;
                        int c = 0;
                        for (int i = 0; i < n; ++i) {
                            int s2 = i + r + 311; s2 = s2 >= n ? s2 - n : s2; s2 = s2 >= n ? s2 - n : s2;
                            c += (pristine[s2] + 1.5f <= xj) ? 1 : 0;
                        }
                        bad += (xrow[j + (j >> 5)] != row[c - 1]) ? 1 : 0;
                    }
                    if (bad) atomicAdd(&check[3], (uint32_t)bad);
                }
            }
        }
        __syncwarp();
        acc += __float_as_uint(row[(lane * 37 + r) % n]) + __float_as_uint(xrow[lane]);
    }
    // verification of the last repetition (bitmap modes write S[0..n) unskewed)
    if (MODE == 2 || MODE == 4) {
        int bad = 0;
        for (int i = lane; i + 1 < n; i += 32) bad += (row[i] > row[i + 1]) ? 1 : 0;
        double s0 = 0.0, s1 = 0.0;
        for (int i = lane; i < n; i += 32) { s0 += (double)row[i]; s1 += (double)pristine[i]; }
        for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (bad) atomicAdd(&check[0], (uint32_t)bad);
        if (lane == 0 && fabs(s0 - s1) > 1e-9 * fabs(s1)) atomicAdd(&check[1], 1u);
    }
    if (lane == 0 && fb) atomicAdd(fallbacks, fb);
    if (acc == 0x12345678u) check[2] = acc;
}

// raw shared-memory op throughput: 8 warps per CTA, every lane hits a pseudo-random word of a per-warp table
template <int OP>   // 0: LDS, 1: ATOMS.OR with return, 2: RED.OR (no return), 3: ATOMS.ADD return, 4: STS
__global__ void __launch_bounds__(256) smem_op_bench(int iters, int words, uint32_t* sink) {
    extern __shared__ uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* W = smem + warp * words;
    for (int i = lane; i < words; i += 32) W[i] = 0;
    __syncwarp();
    uint32_t x = hash32(blockIdx.x * 256u + threadIdx.x + 1u), acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint32_t a = (x >> 12) % (uint32_t)words;
            const uint32_t bit = 1u << (x & 15u);
            if (OP == 0) acc += W[a];
            else if (OP == 1) acc += atomicOr(&W[a], bit);
            else if (OP == 2) atomicOr(&W[a], bit);
            else if (OP == 3) acc += atomicAdd(&W[a], 1u);
            else W[a] = bit;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <class F>
static float time_ms(F&& f, int rep = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < rep; ++i) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, clk_khz);
    uint32_t* check; int* fb;
    CK(cudaMalloc(&check, 64)); CK(cudaMalloc(&fb, 4));
    const int reps = 200;
    auto run = [&](auto kern, const char* name, size_t smem, int ctas_per_sm, int n, int zf) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        const int cps = ctas_per_sm < occ ? ctas_per_sm : occ;
        const int grid = sms * cps;
        CK(cudaMemset(check, 0, 64)); CK(cudaMemset(fb, 0, 4));
        const float ms = time_ms([&] { kern<<<grid, 256, smem>>>(n, reps, zf, check, fb); });
        CK(cudaGetLastError());
        uint32_t h[4]; int hfb;
        CK(cudaMemcpy(h, check, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hfb, fb, 4, cudaMemcpyDeviceToHost));
        const double series = (double)grid * 8 * reps;
        const double cyc_per_series_sm = ms * 1e-3 * clk_khz * 1e3 * sms / series;
        printf("%-28s n=%4d zeros=%2d%% occ=%d(max %d) %8.3f ms  %9.1f SM-cycles/series  unsorted=%u sumbad=%u rankbad=%u fallbacks=%d (of %d)\n",
               name, n, zf, cps, occ, ms, cyc_per_series_sm, h[0] / 4, h[1] / 4, h[3] / 4, hfb / 4, (int)series);  // 4 launches (1 warm-up + 3 timed)
    };
    const size_t sm_plain = 8 * (3 * NPS) * 4, sm_bm = 8 * (3 * NPS + BM_SCRATCH_WORDS) * 4;
    for (int n : {847, 930, 1024}) {
        for (int cps : {1, 2}) {
            run(sort_bench<0>, "copy only", sm_plain, cps, n, 0);
            run(sort_bench<1>, "network y", sm_plain, cps, n, 0);
            run(sort_bench<2>, "bitmap y", sm_bm, cps, n, 0);
            run(sort_bench<3>, "network y + network x", sm_plain, cps, n, 0);
            run(sort_bench<4>, "bitmap y + rank x + gather", sm_bm, cps, n, 0);
        }
    }
    run(sort_bench<2>, "bitmap y zero-infl", sm_bm, 2, 930, 50);
    run(sort_bench<4>, "bitmap y+x zero-infl", sm_bm, 2, 930, 50);
    run(sort_bench<4>, "bitmap y+x zero-infl 90%", sm_bm, 2, 930, 90);

    auto runop = [&](auto kern, const char* name, int words, int cps) {
        const size_t smem = 8 * words * 4;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int iters = 2000;
        const int grid = sms * cps;
        const float ms = time_ms([&] { kern<<<grid, 256, smem>>>(iters, words, check); });
        CK(cudaGetLastError());
        const double warp_ops_per_sm = (double)cps * 8 * iters * 16;
        printf("%-20s words=%5d ctas/SM=%d  %8.3f ms  %6.2f SM-cycles per warp-op\n", name, words, cps, ms,
               ms * 1e-3 * clk_khz * 1e3 / warp_ops_per_sm);
    };
    for (int cps : {1, 2, 4}) {
        runop(smem_op_bench<0>, "LDS random", 1152, cps);
        runop(smem_op_bench<4>, "STS random", 1152, cps);
        runop(smem_op_bench<1>, "ATOMS.OR ret", 1152, cps);
        runop(smem_op_bench<2>, "RED.OR noret", 1152, cps);
        runop(smem_op_bench<3>, "ATOMS.ADD ret", 1152, cps);
    }
    return 0;
}

// analog_grid.cuh — geometry of the quantile grid (up to 512 boxes) shared by the build kernels (analog_grid.cu) and the
// search kernel (analog_kernels.cu).
#pragma once
#include <cstdint>

namespace sdb {

constexpr int AG_BOXES = 512;
constexpr int AG_NBND = 511;         // most planes a grid needs (1 predictor: 511; 2: 42; 3+: 21)

// boxes per predictor: 1 predictor 512 slabs, 2 predictors 22 x 22, otherwise 8 x 8 x 8 on the first three
__host__ __device__ inline void ag_dims(int p, int (&g)[3]) {
    if (p == 1) { g[0] = 512; g[1] = 1; g[2] = 1; }
    else if (p == 2) { g[0] = 22; g[1] = 22; g[2] = 1; }
    else { g[0] = 8; g[1] = 8; g[2] = 8; }
}

// slab of x along one predictor: number of planes <= x (planes ascending; slab b = [plane[b-1], plane[b]))
__device__ __forceinline__ int ag_slab(float x, const float* plane, int g) {
    int lo = 0, hi = g - 1;              // answer in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (plane[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int ag_box(const float (&x)[3], const float* bnd, const int (&g)[3]) {
    const int b0 = ag_slab(x[0], bnd, g[0]);
    const int b1 = g[1] > 1 ? ag_slab(x[1], bnd + (g[0] - 1), g[1]) : 0;
    const int b2 = g[2] > 1 ? ag_slab(x[2], bnd + (g[0] - 1) + (g[1] - 1), g[2]) : 0;
    return (b0 * g[1] + b1) * g[2] + b2;
}

}  // namespace sdb

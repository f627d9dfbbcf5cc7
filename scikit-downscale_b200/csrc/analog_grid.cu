// analog_grid.cu — the spatial index of the analog search: the counterpart of the reference's per-cell
// sklearn KDTree (gard.py:82), built for ALL cells at once.
//
// Per grid cell (series) the training window is cut into up to 512 boxes by QUANTILE planes of the first predictors
// (1 predictor: 512 slabs; 2: 22 x 22; 3 or more: 8 x 8 x 8 on the first three), so every box holds about T / 512
// rows whatever the distribution.  The planes are order statistics (sdb_series_argsort per predictor); rows and
// query steps are then grouped by box with one shared-memory counting sort per series.  The search kernel
// (analog_kernels.cu) visits a box only if its distance lower bound can still beat a query's k-th best distance.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/sdb.h"
#include "common.cuh"
#include "analog_grid.cuh"

namespace sdb {

// planes of predictor f: bounds[(boff + j) * ld_b + c] = ((j + 1) * T / g)-th smallest value of X[:, f, c]
__global__ void ag_bounds_kernel(const float* __restrict__ X, int64_t ld, int64_t C, int T, int p, int f, int g, int boff,
                                 const int32_t* __restrict__ order, int64_t ld_o, float* __restrict__ bounds, int64_t ld_b,
                                 const uint8_t* __restrict__ valid) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (c >= C || j >= g - 1) return;
    float v = 0.0f;
    if (!valid || valid[c]) {
        const int pos = (int)(((int64_t)(j + 1) * T) / g);
        const int row = order[(int64_t)pos * ld_o + c];
        v = X[((int64_t)row * p + f) * ld + c];
    }
    bounds[(int64_t)(boff + j) * ld_b + c] = v;
}

// group the rows of X (training window or query steps) of every series by box: perm[r * ld_p + c] = r-th row in box
// order, start[b * ld_s + c] = first position of box b (start[64] = T); start may be NULL (queries)
__global__ void __launch_bounds__(256) ag_assign_kernel(const float* __restrict__ X, int64_t ld, int64_t C, int T, int p,
                                                        const float* __restrict__ bounds, int64_t ld_b,
                                                        int32_t* __restrict__ perm, int64_t ld_p,
                                                        int32_t* __restrict__ start, int64_t ld_s,
                                                        const uint8_t* __restrict__ valid) {
    __shared__ float bnd[AG_NBND];
    __shared__ int hist[AG_BOXES], base[AG_BOXES + 1];
    const int64_t c = blockIdx.x;
    if (valid && !valid[c]) return;
    int g[3];
    ag_dims(p, g);
    const int nb = (g[0] - 1) + (g[1] - 1) + (g[2] - 1);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bnd[i] = bounds[(int64_t)i * ld_b + c];
    for (int i = threadIdx.x; i < AG_BOXES; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    auto box_of = [&](int t) -> int {
        float x[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int f = 0; f < 3; ++f)
            if (f < p) x[f] = X[((int64_t)t * p + f) * ld + c];
        return ag_box(x, bnd, g);
    };
    for (int t = threadIdx.x; t < T; t += blockDim.x) atomicAdd(&hist[box_of(t)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {
        constexpr int PER = AG_BOXES / 32;
        const int lane = threadIdx.x;
        int sum = 0;
        for (int i = 0; i < PER; ++i) sum += hist[lane * PER + i];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
        for (int i = 0; i < PER; ++i) { base[lane * PER + i] = run; run += hist[lane * PER + i]; }
        if (lane == 31) base[AG_BOXES] = incl;
    }
    __syncthreads();
    if (start) for (int i = threadIdx.x; i <= AG_BOXES; i += blockDim.x) start[(int64_t)i * ld_s + c] = base[i];
    for (int i = threadIdx.x; i < AG_BOXES; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int b = box_of(t);
        perm[(int64_t)(base[b] + atomicAdd(&hist[b], 1)) * ld_p + c] = t;
    }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_analog_grid_fit(const void* X_train, int dtype, int64_t ld, int64_t n_cells, int t_fit, int n_features,
                                   int32_t* workspace, float* bounds, int32_t* perm_train, int32_t* box_start, int64_t ld_grid,
                                   const uint8_t* cell_valid, void* stream) {
    if (!X_train || !workspace || !bounds || !perm_train || !box_start) return sdb_fail(SDB_E_INVALID, "sdb_analog_grid_fit: NULL pointer");
    if (dtype != SDB_F32) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_grid_fit: float32 only");
    if (n_cells <= 0 || t_fit < 64 || n_features < 1 || ld < n_cells || ld_grid < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_analog_grid_fit: bad shape");
    if (t_fit > sdb_series_argsort_max_steps()) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_grid_fit: at most %d steps", sdb_series_argsort_max_steps());
    int g[3];
    ag_dims(n_features, g);
    cudaStream_t st = (cudaStream_t)stream;
    int boff = 0;
    for (int f = 0; f < 3 && f < n_features; ++f) {
        if (g[f] < 2) continue;
        const int rc = sdb_series_argsort((const float*)X_train + (int64_t)f * ld, SDB_F32, (int64_t)n_features * ld, n_cells, t_fit,
                                          workspace, ld_grid, cell_valid, stream);
        if (rc) return rc;
        dim3 grid((unsigned)((n_cells + 127) / 128), (unsigned)(g[f] - 1));
        ag_bounds_kernel<<<grid, 128, 0, st>>>((const float*)X_train, ld, n_cells, t_fit, n_features, f, g[f], boff, workspace, ld_grid,
                                               bounds, ld_grid, cell_valid);
        SDB_CUDA_OK(cudaGetLastError());
        boff += g[f] - 1;
    }
    ag_assign_kernel<<<(unsigned)n_cells, 256, 0, st>>>((const float*)X_train, ld, n_cells, t_fit, n_features, bounds, ld_grid,
                                                       perm_train, ld_grid, box_start, ld_grid, cell_valid);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_analog_grid_assign(const void* X_query, int dtype, int64_t ld, int64_t n_cells, int t_query, int n_features,
                                      const float* bounds, int32_t* perm_query, int64_t ld_grid,
                                      const uint8_t* cell_valid, void* stream) {
    if (!X_query || !bounds || !perm_query) return sdb_fail(SDB_E_INVALID, "sdb_analog_grid_assign: NULL pointer");
    if (dtype != SDB_F32) return sdb_fail(SDB_E_UNSUPPORTED, "sdb_analog_grid_assign: float32 only");
    if (n_cells <= 0 || t_query <= 0 || n_features < 1 || ld < n_cells || ld_grid < n_cells) return sdb_fail(SDB_E_INVALID, "sdb_analog_grid_assign: bad shape");
    ag_assign_kernel<<<(unsigned)n_cells, 256, 0, (cudaStream_t)stream>>>((const float*)X_query, ld, n_cells, t_query, n_features, bounds, ld_grid,
                                                                         perm_query, ld_grid, nullptr, 0, cell_valid);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_analog_grid_boxes(void) { return AG_BOXES; }
extern "C" int sdb_analog_grid_planes(void) { return AG_NBND; }

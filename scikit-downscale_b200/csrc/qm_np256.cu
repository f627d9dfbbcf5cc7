// qm_np256.cu — quantile-mapping kernels for groups padded to 256 items (8 per thread x 32 threads).
#include "qm_kernels.cuh"
#include "qm_tile.cuh"
namespace sdb {
SDB_DEFINE_SIZE(256, 8, 32)
int qm_fit_tile_np256(const FitParams& f, cudaStream_t st) { return launch_fit_tile<8>(f, st); }
int qm_predict_tile_np256(int kind, const PredictParams& p, cudaStream_t st) {
    return kind == KIND_RAW ? launch_predict_tile<8, false>(p, st) : launch_predict_tile<8, true>(p, st);
}
}  // namespace sdb

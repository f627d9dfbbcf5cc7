// qm_np1024.cu — quantile-mapping kernels for groups padded to 1024 items (32 per thread x 32 threads).
#include "qm_kernels.cuh"
#include "qm_tile.cuh"
namespace sdb {
SDB_DEFINE_SIZE(1024, 32, 32)
int qm_fit_tile_np1024(const FitParams& f, cudaStream_t st) { return launch_fit_tile<32>(f, st); }
int qm_predict_tile_np1024(int kind, const PredictParams& p, cudaStream_t st) {
    return kind == KIND_RAW ? launch_predict_tile<32, false>(p, st) : launch_predict_tile<32, true>(p, st);
}
}  // namespace sdb

// qm_np4096.cu — quantile-mapping kernels for groups padded to 4096 items (32 per thread x 128 threads).
#include "qm_kernels.cuh"
namespace sdb {
SDB_DEFINE_SIZE(4096, 32, 128)
}  // namespace sdb

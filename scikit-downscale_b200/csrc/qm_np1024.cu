// qm_np1024.cu — quantile-mapping kernels for groups padded to 1024 items (32 per thread x 32 threads).
#include "qm_kernels.cuh"
namespace sdb {
SDB_DEFINE_SIZE(1024, 32, 32)
}  // namespace sdb

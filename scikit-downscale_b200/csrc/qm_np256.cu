// qm_np256.cu — quantile-mapping kernels for groups padded to 256 items (8 per thread x 32 threads).
#include "qm_kernels.cuh"
namespace sdb {
SDB_DEFINE_SIZE(256, 8, 32)
}  // namespace sdb

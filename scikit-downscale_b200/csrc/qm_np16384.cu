// qm_np16384.cu — quantile-mapping kernels for groups padded to 16384 items (32 per thread x 512 threads).
#include "qm_kernels.cuh"
namespace sdb {
SDB_DEFINE_SIZE(16384, 32, 512)
}  // namespace sdb

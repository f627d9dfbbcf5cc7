// sdb_api.cu — library-level entry points of the C ABI (include/sdb.h).
#include "../../include/sdb.h"
#include "common.cuh"

namespace sdb {
char* sdb_error_buffer() {
    static thread_local char buf[kErrLen] = {0};
    return buf;
}
int g_debug_flags = 0;      // sdb_set_debug_flags: process-wide testing aid (see include/sdb.h)
}  // namespace sdb

extern "C" int sdb_set_debug_flags(int flags) { int old = sdb::g_debug_flags; sdb::g_debug_flags = flags; return old; }
extern "C" int sdb_version(void) { return 200; }   /* 0.2.0 */
extern "C" const char* sdb_last_error(void) { return sdb::sdb_error_buffer(); }

extern "C" int sdb_memcpy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                                  int64_t width_bytes, int64_t height, int kind, void* stream) {
    if (!dst || !src) return sdb::sdb_fail(SDB_E_INVALID, "sdb_memcpy2d_async: NULL pointer");
    if (width_bytes < 0 || height < 0 || dst_pitch < width_bytes || src_pitch < width_bytes || (kind < 0 || kind > 2))
        return sdb::sdb_fail(SDB_E_INVALID, "sdb_memcpy2d_async: bad geometry");
    if (width_bytes == 0 || height == 0) return 0;
    SDB_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)height,
                                  kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDefault),
                                  (cudaStream_t)stream));
    return 0;
}

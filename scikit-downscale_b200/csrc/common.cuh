// common.cuh — error plumbing shared by the translation units of libsdb.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <type_traits>

namespace sdb {

char* sdb_error_buffer();           // thread-local, defined in sdb_api.cu
constexpr int kErrLen = 512;
extern int g_debug_flags;           // defined in sdb_api.cu (sdb_set_debug_flags)

inline int sdb_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(sdb_error_buffer(), kErrLen, fmt, ap);
    va_end(ap);
    return code;
}

#define SDB_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::sdb::sdb_fail(SDB_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,      \
                                   cudaGetErrorString(e__));                                   \
    } while (0)

}  // namespace sdb

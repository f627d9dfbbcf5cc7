// sdb_api.cu — library-level entry points of the C ABI (include/sdb.h).
#include "../../include/sdb.h"
#include "common.cuh"
#include <cstring>

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

// ---------------------------------------------------------------- peer push of a column block (the in-path gather)
// Rows of `width` bytes (a multiple of 16, 16-byte aligned on both sides) are read from local memory and written
// to `dst`, which may be a PEER GPU's memory mapped into this process (CUDA IPC): plain coalesced 16-byte stores
// that the SM's memory system forwards over NVLink.  A few dozen CTAs saturate the link, so the copy runs beside
// the compute kernels of the next cell chunk.  (cudaMemcpy2DAsync on the same pointers measured 25 GB/s on B200:
// the pitched peer copy does not take the NVLink fast path.)
__global__ void __launch_bounds__(256) peer_copy2d_kernel(uint4* __restrict__ dst, int64_t dst_pitch16, const uint4* __restrict__ src,
                                                          int64_t src_pitch16, int width16, int64_t height) {
    const int64_t per_row = width16;
    const int64_t total = per_row * height;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row, c = i - r * per_row;
        dst[r * dst_pitch16 + c] = __ldcs(src + r * src_pitch16 + c);
    }
}

extern "C" int sdb_peer_copy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                               int64_t width_bytes, int64_t height, int n_ctas, void* stream) {
    if (!dst || !src) return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_copy2d: NULL pointer");
    if (width_bytes < 0 || height < 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_copy2d: bad geometry");
    if (width_bytes == 0 || height == 0) return 0;
    if ((width_bytes | dst_pitch | src_pitch | (int64_t)(uintptr_t)dst | (int64_t)(uintptr_t)src) & 15)
        return sdb::sdb_fail(SDB_E_UNSUPPORTED, "sdb_peer_copy2d: rows must be 16-byte aligned multiples of 16 bytes (use sdb_memcpy2d_async)");
    if (width_bytes / 16 > 0x7fffffff) return sdb::sdb_fail(SDB_E_UNSUPPORTED, "sdb_peer_copy2d: row too long");
    if (n_ctas <= 0) n_ctas = 32;
    peer_copy2d_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>((uint4*)dst, dst_pitch / 16, (const uint4*)src, src_pitch / 16,
                                                               (int)(width_bytes / 16), height);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sdb_enable_peer_access(int device, int peer_device) {
    int cur = -1;
    SDB_CUDA_OK(cudaGetDevice(&cur));
    if (device == peer_device) return 0;
    int can = 0;
    SDB_CUDA_OK(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) return sdb::sdb_fail(SDB_E_UNSUPPORTED, "sdb_enable_peer_access: device %d cannot access device %d", device, peer_device);
    SDB_CUDA_OK(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
    cudaSetDevice(cur);
    if (e != cudaSuccess) return sdb::sdb_fail(SDB_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer_device, cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------- peer-visible buffers (CUDA IPC)
// The replica of the gathered field is allocated HERE with cudaMalloc (an IPC handle names a whole allocation;
// a framework's caching allocator hands out interior pointers) and opened by the peers in the context of the
// device that will access it, with lazy peer access — the documented route for kernels and copies on IPC memory.
extern "C" int sdb_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
    if (!ptr || !handle64 || bytes <= 0) return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    SDB_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return sdb::sdb_fail(SDB_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int sdb_peer_open(const void* handle64, void** ptr) {
    if (!ptr || !handle64) return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SDB_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int sdb_peer_close(void* ptr) { if (ptr) SDB_CUDA_OK(cudaIpcCloseMemHandle(ptr)); return 0; }
extern "C" int sdb_peer_free(void* ptr) { if (ptr) SDB_CUDA_OK(cudaFree(ptr)); return 0; }

// ---------------------------------------------------------------- one read, n peer writes (the gather at N = 8)
// Seven pitched copy-engine transfers per chunk, on seven streams, reached 353 GB/s received per GPU with all eight
// GPUs pushing (profiles/r02_bench_n8_ce.json) — no better than the NCCL all-gather of round 1.  This kernel reads
// a 16-byte element of the local block ONCE and stores it into the same place of every peer's replica: the SMs'
// posted remote stores keep all NVLink ports busy at once.
struct PeerDsts { uint4* p[8]; };
__global__ void __launch_bounds__(256) peer_bcast2d_kernel(PeerDsts d, int n_dst, int64_t dst_pitch16, const uint4* __restrict__ src,
                                                           int64_t src_pitch16, int width16, int64_t height) {
    // a CTA walks whole rows (no 64-bit division per element); four loads in flight per thread
    for (int64_t r = blockIdx.x; r < height; r += gridDim.x) {
        const uint4* s = src + r * src_pitch16;
        const int64_t at = r * dst_pitch16;
        for (int c0 = threadIdx.x; c0 < width16; c0 += 4 * 256) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u * 256 < width16) v[u] = __ldcs(s + c0 + u * 256);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < n_dst) {
                    uint4* dk = d.p[k] + at;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (c0 + u * 256 < width16) dk[c0 + u * 256] = v[u];
                }
            }
        }
    }
}

extern "C" int sdb_peer_bcast2d(void* const* dsts, int n_dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                                int64_t width_bytes, int64_t height, int n_ctas, void* stream) {
    if (!dsts || !src || n_dst < 1 || n_dst > 8) return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_bcast2d: 1..8 destinations");
    if (width_bytes < 0 || height < 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_bcast2d: bad geometry");
    if (width_bytes == 0 || height == 0) return 0;
    int64_t bits = width_bytes | dst_pitch | src_pitch | (int64_t)(uintptr_t)src;
    PeerDsts d;
    for (int k = 0; k < 8; ++k) {
        d.p[k] = (k < n_dst) ? (uint4*)dsts[k] : nullptr;
        if (k < n_dst) { if (!dsts[k]) return sdb::sdb_fail(SDB_E_INVALID, "sdb_peer_bcast2d: NULL destination"); bits |= (int64_t)(uintptr_t)dsts[k]; }
    }
    if (bits & 15) return sdb::sdb_fail(SDB_E_UNSUPPORTED, "sdb_peer_bcast2d: rows must be 16-byte aligned multiples of 16 bytes");
    if (width_bytes / 16 > 0x7fffffff) return sdb::sdb_fail(SDB_E_UNSUPPORTED, "sdb_peer_bcast2d: row too long");
    if (n_ctas <= 0) n_ctas = 296;
    peer_bcast2d_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(d, n_dst, dst_pitch / 16, (const uint4*)src, src_pitch / 16,
                                                                (int)(width_bytes / 16), height);
    SDB_CUDA_OK(cudaGetLastError());
    return 0;
}

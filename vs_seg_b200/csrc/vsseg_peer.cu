// Peer-memory plumbing of the multi-GPU sliding window: the ranks of one node blend their windows
// straight into the accumulator of the destination rank over NVLink (red.global.add.f32 in
// vsseg_conv3d_gate_logits), so the overlap-weighted volume is assembled WITHOUT a separate reduce pass.  This file
// holds what that needs besides the blend itself: device allocations that can be mapped by the peer processes
// (CUDA IPC) and system-scope flags for the per-volume hand-shake (arrive / release counters).
#include "vsseg_common.cuh"

namespace vsseg {

__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one thread: spin until every flags[i] >= target (bounded: a dead peer must not hang the GPU)
__global__ void flag_wait_kernel(const long long* flags, int n, long long target, long long timeout_cycles, int* err) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        while (ld_acquire_sys(flags + i) < target) {
            if (clock64() - t0 > timeout_cycles) {
                if (err) atomicExch(err, 1);
                return;
            }
            __nanosleep(256);
        }
    }
}

// one thread: everything this stream did before (including reds into peer memory of earlier kernels, which are
// performed when those kernels complete) is ordered before the flag becomes visible system wide
__global__ void flag_set_kernel(long long* flag, long long value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_peer_alloc(int64_t bytes, void** ptr_host, void* handle64_host) {
    VSSEG_REQUIRE(bytes > 0 && ptr_host && handle64_host, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        set_error("peer_alloc: %s", cudaGetErrorString(e));
        if (p) cudaFree(p);
        cudaGetLastError();
        return (int)e;
    }
    memcpy(handle64_host, &h, sizeof(h));
    *ptr_host = p;
    return 0;
}

int vsseg_peer_open(const void* handle64_host, void** ptr_host) {
    VSSEG_REQUIRE(handle64_host && ptr_host, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("peer_open: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    *ptr_host = p;
    return 0;
}

int vsseg_peer_close(void* ptr) {
    VSSEG_REQUIRE(ptr, "peer_close: NULL");
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("peer_close: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

int vsseg_peer_free(void* ptr) {
    VSSEG_REQUIRE(ptr, "peer_free: NULL");
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("peer_free: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

int vsseg_flag_wait(const int64_t* flags, int32_t n, int64_t target, int64_t timeout_cycles, int32_t* err, void* stream) {
    VSSEG_REQUIRE(flags && n > 0 && timeout_cycles > 0, "flag_wait: bad arguments");
    flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const long long*)flags, n, (long long)target, (long long)timeout_cycles, err);
    return check_launch("flag_wait");
}

int vsseg_flag_set(int64_t* flag, int64_t value, void* stream) {
    VSSEG_REQUIRE(flag, "flag_set: NULL flag");
    flag_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)flag, (long long)value);
    return check_launch("flag_set");
}

}  // extern "C"

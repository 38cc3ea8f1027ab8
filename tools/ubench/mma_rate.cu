// Micro-benchmark: tcgen05.mma issue/throughput for small N, same vs rotating accumulators,
// SWIZZLE_NONE K-major operands (the layout vsseg_tc.cu uses).  One CTA, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N, int NACC, bool ELECT>
__global__ void bench(int iters, int astep, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar; __shared__ uint32_t tptr;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr;
    bool leader;
    if (ELECT) {
        uint32_t pred = 0;
        if (threadIdx.x < 32)
            asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
        leader = pred != 0;
    } else {
        leader = threadIdx.x == 0;
    }
    if (leader) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t da = make_desc(smem_u32(smem), 130 * 16 * 6, 128);
        const uint64_t db = make_desc(smem_u32(smem) + 100 * 1024, N * 16, 128);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                umma(tb + (uint32_t)((u % NACC) * N), da + (uint64_t)(u * astep), db + (uint64_t)(u * 2 * N), idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N, int NACC, bool ELECT>
void run(long long* d) {
    long long h[2];
    cudaFuncSetAttribute(bench<N, NACC, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int astep : {0, 130}) {
        const int iters = 128;
        bench<N, NACC, ELECT><<<1, 128, 200 * 1024>>>(iters, astep, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("N=%3d nacc=%d elect=%d astep=%3d  issue %.1f cyc/mma  complete %.1f cyc/mma (floor %d)\n", N, NACC, (int)ELECT,
               astep, h[0] / (iters * 8.0), h[1] / (iters * 8.0), 128 * N / 256);
    }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    run<16, 1, false>(d); run<16, 1, true>(d); run<16, 4, false>(d); run<16, 8, true>(d);
    run<32, 1, true>(d); run<32, 4, true>(d); run<32, 8, true>(d);
    run<48, 1, true>(d); run<48, 4, true>(d); run<48, 8, true>(d);
    run<64, 1, true>(d); run<64, 8, true>(d);
    run<96, 1, true>(d); run<96, 4, true>(d);
    run<128, 1, true>(d); run<128, 4, true>(d);
    run<256, 1, true>(d); run<256, 2, true>(d);
    return 0;
}

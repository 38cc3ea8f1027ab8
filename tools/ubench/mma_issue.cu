// Micro-benchmark: cost per tcgen05.mma (M=128, K=16, SWIZZLE_NONE K-major) of the ways the issuing thread can
// obtain its operands.  Pattern = narrow conv layer: L A lines x 3 passes, N columns, overlapping accumulators.
//   mode 0: operands from register arithmetic only (regular pattern, no table)
//   mode 1: table in kernel parameters (constant bank, uniform-register indexed loads) - what vsseg_tc.cu does
//   mode 2: table in shared memory
//   mode 3: table in registers of the 32 lanes, broadcast with shfl
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1,%2};\nmov.b64 db, {%3,%4};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}
struct __align__(16) OpD { uint32_t a_lo, a_hi, b_lo, b_hi; };
struct __align__(8) OpC { uint32_t col, idesc; };
constexpr int NOPS = 30;
struct Args { int mode, iters, n8, commit_every; OpD d[NOPS]; OpC c[NOPS]; };

__global__ void __launch_bounds__(128) bench(const __grid_constant__ Args a, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2; __shared__ uint32_t tptr;
    __shared__ OpD sd[NOPS]; __shared__ OpC sc[NOPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2))); }
    for (int i = threadIdx.x; i < 100 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
    for (int i = threadIdx.x; i < NOPS; i += blockDim.x) { sd[i] = a.d[i]; sc[i] = a.c[i]; }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr;
    const uint32_t da = (smem_u32(smem) & 0x3FFFF) >> 4, db = ((smem_u32(smem) + 90 * 1024) & 0x3FFFF) >> 4;
    if (warp == 1) {
        long long t0 = 0, t1 = 0;
        if (a.mode == 3) {
            // every lane keeps one op in registers; the op is broadcast, one elected lane issues
            const OpD od = a.d[lane < NOPS ? lane : 0];
            const OpC oc = a.c[lane < NOPS ? lane : 0];
            const bool leader = elect_one();
            t0 = clock64();
            for (int i = 0; i < a.iters; ++i) {
#pragma unroll 6
                for (int k = 0; k < NOPS; ++k) {
                    const uint32_t alo = __shfl_sync(0xffffffffu, od.a_lo, k), blo = __shfl_sync(0xffffffffu, od.b_lo, k);
                    const uint32_t col = __shfl_sync(0xffffffffu, oc.col, k), id = __shfl_sync(0xffffffffu, oc.idesc, k);
                    if (leader) umma(tb + col, alo + da, od.a_hi, blo + db, od.b_hi, id);
                }
            }
            if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
        } else if (elect_one()) {
            t0 = clock64();
            if (a.mode == 0) {
                const uint32_t n = a.n8 * 8;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a.n8 << 17) | (8u << 24);
                const uint32_t hi = (128u >> 4) | (1u << 14);
                const uint32_t a0 = da + ((20480u >> 4) << 16), b0 = db + ((n * 16u >> 4) << 16);
                for (int i = 0; i < a.iters; ++i) {
                    uint32_t al = a0, col = tb;
#pragma unroll 2
                    for (int line = 0; line < NOPS / 3; ++line) {
                        umma(col, al, hi, b0, hi, idesc);
                        umma(col, al + 1280u, hi, b0, hi, idesc);
                        umma(col, al, hi, b0 + n * 2u, hi, idesc);
                        al += 128u; col += 16u;
                    }
                }
            } else if (a.mode == 1) {
                for (int i = 0; i < a.iters; ++i) {
#pragma unroll 4
                    for (int k = 0; k < NOPS; ++k) {
                        const OpD o = a.d[k]; const OpC c = a.c[k];
                        umma(tb + c.col, o.a_lo + da, o.a_hi, o.b_lo + db, o.b_hi, c.idesc);
                    }
                    if (a.commit_every && (i % a.commit_every) == a.commit_every - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                }
            } else {
                for (int i = 0; i < a.iters; ++i) {
#pragma unroll 4
                    for (int k = 0; k < NOPS; ++k) {
                        const OpD o = sd[k]; const OpC c = sc[k];
                        umma(tb + c.col, o.a_lo + da, o.a_hi, o.b_lo + db, o.b_hi, c.idesc);
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
        }
        if (lane == 0) {
            mbar_wait(smem_u32(&bar), 0);
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[1] = t2 - t0; }
        }
        if (t1 && blockIdx.x == 0 && (a.mode != 3 || lane == 0)) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n8 : {2, 6, 18})
        for (int ce : {0, 1, 3}) {
            const int mode = 1;
            Args a; a.mode = mode; a.iters = 64; a.n8 = n8; a.commit_every = ce;
            const uint32_t n = n8 * 8, hi = (128u >> 4) | (1u << 14);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)n8 << 17) | (8u << 24);
            for (int k = 0; k < NOPS; ++k) {
                const int line = k / 3, pass = k % 3;
                a.d[k] = {(uint32_t)line * 128u + (pass == 1 ? 1280u : 0u) + ((20480u >> 4) << 16), hi, (pass == 2 ? n * 2u : 0u) + ((n * 16u >> 4) << 16), hi};
                a.c[k] = {(uint32_t)line * 16u, idesc};
            }
            long long h[2] = {0, 0};
            cudaMemset(d, 0, 64);
            bench<<<148, 128, 200 * 1024>>>(a, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double nm = 64.0 * NOPS;
            printf("N=%3d mode=%d commit_every=%d stages : issue %.1f  complete %.1f cyc/mma  (model %d)\n", n, mode, ce, h[0] / nm, h[1] / nm,
                   (int)(n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4));
        }
    return 0;
}

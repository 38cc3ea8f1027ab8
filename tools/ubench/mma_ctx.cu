// Micro-benchmark: which concurrent activity of the conv kernel slows its tcgen05.mma stream (M=128, K=16,
// SWIZZLE_NONE K-major, operands from a kernel-parameter table exactly like vsseg_tc.cu)?
// One CTA per SM; warp 1 issues 30-op stages; optional: warp 0 streams bulk copies into shared memory (tma),
// warps 2..9 run tcgen05.ld (ld) and/or tcgen05.st (st) loops, or spin on an mbarrier (poll).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1,%2};\nmov.b64 db, {%3,%4};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}
struct __align__(16) OpD { uint32_t a_lo, a_hi, b_lo, b_hi; };
struct __align__(8) OpC { uint32_t col, idesc; };
constexpr int NOPS = 30;
struct Args { int tma, ld, st, poll, gst, alu, iters, n8; OpD d[NOPS]; OpC c[NOPS]; };

__global__ void __launch_bounds__(576) bench(const __grid_constant__ Args a, const uint4* gsrc, uint4* gdst, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, never, tbar[2]; __shared__ uint32_t tptr; __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[1])));
        stop = 0;
    }
    for (int i = threadIdx.x; i < 100 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
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
    if (warp == 0) {
        if (a.tma && lane == 0) {
            const uint32_t dst = smem_u32(smem) + 100 * 1024;
            int it = 0;
            while (!stop) {
                const int b = it & 1;
                if (it >= 2) mbar_wait(smem_u32(&tbar[b]), ((it >> 1) - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[b])), "r"(40960) : "memory");
                for (int k = 0; k < 20; ++k)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + b * 40960 + k * 2048),
                                 "l"(gsrc + ((size_t)blockIdx.x * 64 + (it * 20 + k) % 4096) * 128), "r"(2048), "r"(smem_u32(&tbar[b])) : "memory");
                ++it;
            }
            for (int k = it > 2 ? it - 2 : 0; k < it; ++k) mbar_wait(smem_u32(&tbar[k & 1]), (k >> 1) & 1);
            if (blockIdx.x == 0) out[3] = it;
        }
    } else if (warp == 1) {
        if (elect_one()) {
            long long t0 = clock64();
            for (int i = 0; i < a.iters; ++i) {
#pragma unroll 4
                for (int k = 0; k < NOPS; ++k) {
                    const OpD o = a.d[k]; const OpC c = a.c[k];
                    umma(tb + c.col, o.a_lo + da, o.a_hi, o.b_lo + db, o.b_hi, c.idesc);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            long long t1 = clock64();
            mbar_wait(smem_u32(&bar), 0);
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
            stop = 1;
        }
    } else {
        const uint32_t tl = tb + (((uint32_t)(warp & 3) * 32) << 16) + 320;   // columns the MMAs do not touch
        uint32_t acc = 0;
        int it = 0;
        uint4* gp = gdst + ((size_t)blockIdx.x * 8 + (warp - 2)) * 32 * 64 + lane;
        while (!stop) {
            if (a.ld) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tl + (it & 7) * 16) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 16; ++q) acc += r[q];
                if (a.gst) {
                    gp[(it & 15) * 32 * 4] = make_uint4(r[0], r[1], r[2], r[3]);
                    gp[(it & 15) * 32 * 4 + 32] = make_uint4(r[4], r[5], r[6], r[7]);
                    gp[(it & 15) * 32 * 4 + 64] = make_uint4(r[8], r[9], r[10], r[11]);
                    gp[(it & 15) * 32 * 4 + 96] = make_uint4(r[12], r[13], r[14], r[15]);
                }
            }
            if (a.st) {
                const uint32_t z = 0;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl + (it & 7) * 16), "r"(z) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            if (a.poll) { if (try_wait(smem_u32(&never), 0)) break; }
            if (a.alu) {
                float f0 = acc, f1 = it, f2 = lane, f3 = warp, f4 = 1.f, f5 = 2.f, f6 = 3.f, f7 = 4.f;
#pragma unroll 16
                for (int q = 0; q < 64 * a.alu; ++q) {
                    f0 = fmaf(f0, 1.0001f, f1); f1 = fmaf(f1, 1.0001f, f2); f2 = fmaf(f2, 1.0001f, f3); f3 = fmaf(f3, 1.0001f, f4);
                    f4 = fmaf(f4, 1.0001f, f5); f5 = fmaf(f5, 1.0001f, f6); f6 = fmaf(f6, 1.0001f, f7); f7 = fmaf(f7, 1.0001f, f0);
                }
                acc += __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
            }
            if (!a.ld && !a.st && !a.poll && !a.alu) __nanosleep(200);
            ++it;
        }
        if (acc == 0x12345) out[2] = acc;
        if (blockIdx.x == 0 && warp == 2 && lane == 0) out[4] = it;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    uint4* g; cudaMalloc(&g, (size_t)(148 * 64 + 4096) * 2048); cudaMemset(g, 0, (size_t)(148 * 64 + 4096) * 2048);
    uint4* go; cudaMalloc(&go, (size_t)148 * 8 * 32 * 64 * 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfgs[][6] = {{0,0,0,0,0,0},{0,0,0,0,0,1},{1,1,1,0,1,0},{1,1,1,0,1,1},{1,1,1,0,1,4}};
    for (int n8 : {2, 6, 12, 18})
        for (auto& cf : cfgs) {
            Args a; a.tma = cf[0]; a.ld = cf[1]; a.st = cf[2]; a.poll = cf[3]; a.gst = cf[4]; a.alu = cf[5]; a.iters = 256; a.n8 = n8;
            const uint32_t n = n8 * 8, hi = (128u >> 4) | (1u << 14);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)n8 << 17) | (8u << 24);
            for (int k = 0; k < NOPS; ++k) {
                const int line = k / 3, pass = k % 3;
                a.d[k] = {(uint32_t)line * 128u + (pass == 1 ? 1280u : 0u) + ((20480u >> 4) << 16), hi, (pass == 2 ? n * 2u : 0u) + ((n * 16u >> 4) << 16), hi};
                a.c[k] = {(uint32_t)line * 16u, idesc};
            }
            long long h[5] = {0, 0, 0, 0, 0};
            cudaMemset(d, 0, 64);
            bench<<<148, 576, 200 * 1024>>>(a, g, go, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
            const double nm = 256.0 * NOPS;
            printf("N=%3d tma=%d ld=%d st=%d poll=%d gst=%d alu=%d : complete %.1f cyc/mma (model %d) | tma %.1f B/cyc, epi iters/mma %.2f\n", n, a.tma, a.ld, a.st, a.poll, a.gst, a.alu,
                   h[1] / nm, (int)(n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4), h[3] * 40960.0 / h[1], h[4] / nm);
        }
    return 0;
}

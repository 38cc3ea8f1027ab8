// Micro-benchmark + semantics check: A operand of tcgen05.mma staged in TMEM (tcgen05.cp.128x256b from the same
// K-major SWIZZLE_NONE shared-memory layout vsseg_tc.cu uses), "TS" MMA vs the "SS" MMA.
//  1. correctness: D_ss = A*B with both operands from shared memory; D_ts = the same product with A copied to TMEM
//     first; both read back with tcgen05.ld and compared on the host against an integer reference.
//  2. rate: cycles per MMA for N = 16..128, SS vs TS, and the cost of the copies.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A: 128 rows x 16 k, two K halves LBO apart, rows 16 B apart in 8-row groups (SBO = 128); B: N rows x 16 k likewise.
constexpr int A_LBO = 2080, A_OFF = 0, B_OFF = 8192;
template <int N>
__global__ void check_kernel(float* out_ss, float* out_ts, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    __nv_bfloat16* A = (__nv_bfloat16*)(smem + A_OFF);
    __nv_bfloat16* B = (__nv_bfloat16*)(smem + B_OFF);
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int m = i / 16, k = i % 16;
        A[(k / 8) * (A_LBO / 2) + m * 8 + (k % 8)] = __float2bfloat16((float)((m * 3 + k * 5) % 7 - 3));
    }
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        B[(k / 8) * (N * 8) + n * 8 + (k % 8)] = __float2bfloat16((float)((n * 2 + k) % 5 - 2));
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t d_ss = tb, d_ts = tb + 128, a_t = tb + 256;
    if (threadIdx.x == 0) {
        const uint64_t da = make_desc(smem_u32(smem) + A_OFF, A_LBO, 128);
        const uint64_t db = make_desc(smem_u32(smem) + B_OFF, N * 16, 128);
        umma_ss(d_ss, da, db, idesc, 0u);
        tmem_cp_128x256b(a_t, da);
        umma_ts(d_ts, a_t, db, idesc, 0u);
        commit(&bar);
        wait(&bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 4 warps: lanes warp*32..
        const int row = warp * 32 + lane;
        for (int c = 0; c < N; c += 16) {
            uint32_t v[16];
            tmem_ld16(d_ss + ((uint32_t)(warp * 32) << 16) + c, v);
            for (int q = 0; q < 16; ++q) out_ss[row * N + c + q] = __uint_as_float(v[q]);
            tmem_ld16(d_ts + ((uint32_t)(warp * 32) << 16) + c, v);
            for (int q = 0; q < 16; ++q) out_ts[row * N + c + q] = __uint_as_float(v[q]);
        }
    }
    // ---- rates (one issuing thread): SS, TS, and copy + TS in the proportion of the conv kernel (1 copy per 3 MMAs)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint64_t da = make_desc(smem_u32(smem) + A_OFF, A_LBO, 128);
        const uint64_t db = make_desc(smem_u32(smem) + B_OFF, N * 16, 128);
        const int iters = 512;
        uint32_t par = 1;
        for (int mode = 0; mode < 4; ++mode) {
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tb + (uint32_t)((i & 1) * 128);
                if (mode == 0) umma_ss(d, da, db, idesc, 1u);
                else if (mode == 1) umma_ts(d, a_t, db, idesc, 1u);
                else if (mode == 2) { if (i % 3 == 0) tmem_cp_128x256b(a_t + (uint32_t)((i / 3 & 1) * 8), da); umma_ts(d, a_t + (uint32_t)((i / 3 & 1) * 8), db, idesc, 1u); }
                else tmem_cp_128x256b(a_t + (uint32_t)((i & 3) * 8), da);
            }
            commit(&bar);
            long long t1 = clock64();
            wait(&bar, par);
            par ^= 1;
            long long t2 = clock64();
            cyc[mode * 2] = (t1 - t0); cyc[mode * 2 + 1] = (t2 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// unrolled issue loops (8 MMAs per iteration, rotating accumulators): execution rate of SS vs TS MMAs
template <int N, bool TS, int NACC>
__global__ void rate_kernel(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t da = make_desc(smem_u32(smem) + A_OFF, A_LBO, 128);
        const uint64_t db = make_desc(smem_u32(smem) + B_OFF, N * 16, 128);
        const uint32_t a_t = tb + 448;
        for (int v = 0; v < 8; ++v) tmem_cp_128x256b(a_t + v * 8, da);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = tb + (uint32_t)((u % NACC) * N);
                if (TS) umma_ts(d, a_t + (uint32_t)(u * 8), db, idesc, 1u);
                else umma_ss(d, da + (uint64_t)(u * 8), db, idesc, 1u);
            }
        }
        commit(&bar);
        long long t1 = clock64();
        wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N, int NACC>
void run_rate(long long* d) {
    long long h[2], g[2];
    const int iters = 256;
    cudaFuncSetAttribute(rate_kernel<N, false, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(rate_kernel<N, true, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    rate_kernel<N, false, NACC><<<1, 128, 64 * 1024>>>(iters, d);
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    rate_kernel<N, true, NACC><<<1, 128, 64 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(g, d, 16, cudaMemcpyDeviceToHost);
    printf("rate N=%3d nacc=%d : SS issue %.1f complete %.1f | TS issue %.1f complete %.1f cycles/MMA (N/2 = %d, 32+N/4 = %d) %s\n", N, NACC,
           h[0] / (iters * 8.0), h[1] / (iters * 8.0), g[0] / (iters * 8.0), g[1] / (iters * 8.0), N / 2, 32 + N / 4,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N>
int run() {
    float *ss, *ts; long long* cyc;
    cudaMalloc(&ss, 128 * N * 4); cudaMalloc(&ts, 128 * N * 4); cudaMalloc(&cyc, 64);
    cudaMemset(ss, 0xff, 128 * N * 4); cudaMemset(ts, 0xff, 128 * N * 4);
    cudaFuncSetAttribute(check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    check_kernel<N><<<1, 128, 64 * 1024>>>(ss, ts, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: error %s\n", N, cudaGetErrorString(e)); return 1; }
    static float hs[128 * 256], ht[128 * 256];
    long long hc[8];
    cudaMemcpy(hs, ss, 128 * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, ts, 128 * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 16; ++k) ref += (float)((m * 3 + k * 5) % 7 - 3) * (float)((n * 2 + k) % 5 - 2);
            bad_ss += hs[m * N + n] != ref;
            bad_ts += ht[m * N + n] != ref;
        }
    const double it = 512.0;
    printf("N=%3d  mismatches: SS %d  TS %d  | cycles/op issue,complete: SS %.1f %.1f | TS %.1f %.1f | cp+3xTS %.1f %.1f | cp only %.1f %.1f\n",
           N, bad_ss, bad_ts, hc[0] / it, hc[1] / it, hc[2] / it, hc[3] / it, hc[4] / it, hc[5] / it, hc[6] / it, hc[7] / it);
    if (bad_ts) {
        printf("   first TS rows: ");
        for (int n = 0; n < 8; ++n) printf("%g ", ht[n]);
        printf("| expected ");
        for (int n = 0; n < 8; ++n) printf("%g ", hs[n]);
        printf("\n");
    }
    return bad_ts != 0;
}
int main() {
    long long* dd; cudaMalloc(&dd, 16);
    run_rate<16, 4>(dd); run_rate<32, 4>(dd); run_rate<48, 4>(dd); run_rate<64, 4>(dd); run_rate<96, 4>(dd); run_rate<128, 2>(dd);
    run_rate<32, 1>(dd); run_rate<96, 1>(dd);
    int bad = 0;
    bad += run<16>(); bad += run<32>(); bad += run<48>(); bad += run<64>(); bad += run<96>(); bad += run<128>();
    return bad ? 2 : 0;
}

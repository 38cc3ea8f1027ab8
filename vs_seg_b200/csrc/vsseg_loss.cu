// Hardness-weighted / attention-supervised Dice loss of VS_Seg as bandwidth-bound kernels
// (reference params/losses/dice_spvPA.py:90-167 DiceLoss.forward, :238-297 Dice_spvPA.forward).
// Forward = one reduction pass per term (logits term: softmax + one-hot + hardness weight fused,
// evaluated once instead of twice as the reference does at :282 and :109), a one-thread finalise that
// turns the sums into the loss and the per-(batch,class) backward coefficients, and one elementwise
// backward pass per term.  128-bit loads, fp64 atomics for the global sums.
#include "vsseg_common.cuh"

namespace vsseg {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void block_atomic_add(const float (&acc)[NV], double* dst) {
    __shared__ double red[NV][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum((double)acc[i]);
        if (lane == 0) red[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(dst + threadIdx.x, s);
    }
}

// ---- label pyramid: max-pool with kernel = stride = ratio (dice_spvPA.py:268-277) ----------------
__global__ void __launch_bounds__(256) maxpool3d_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                        int Xo, int Yo, int Zo, int rx, int ry, int rz) {
    const int64_t n = (int64_t)B * Xo * Yo * Zo;
    const int Yi = Yo * ry, Zi = Zo * rz, Xi = Xo * rx;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Zo), y = (int)((i / Zo) % Yo), x = (int)((i / ((int64_t)Zo * Yo)) % Xo);
        const int b = (int)(i / ((int64_t)Zo * Yo * Xo));
        float m = -INFINITY;
        for (int dx = 0; dx < rx; ++dx)
            for (int dy = 0; dy < ry; ++dy)
                for (int dz = 0; dz < rz; ++dz)
                    m = fmaxf(m, __ldg(in + (((int64_t)b * Xi + x * rx + dx) * Yi + y * ry + dy) * Zi + z * rz + dz));
        out[i] = m;
    }
}

// ---- forward sums ---------------------------------------------------------------------------------
// single-channel term (attention map vs pooled label): sums[b][0..2] += (sum a*g, sum g, sum a)
__global__ void __launch_bounds__(256) dice_sums1_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                         int64_t n, double* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* p = pred + (int64_t)b * n;
    const float* t = tgt + (int64_t)b * n;
    float acc[3] = {0.f, 0.f, 0.f};
    const int64_t n4 = (((uintptr_t)p | (uintptr_t)t) & 15) == 0 ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p) + i), g = __ldg(reinterpret_cast<const float4*>(t) + i);
        acc[0] += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
        acc[1] += g.x + g.y + g.z + g.w;
        acc[2] += a.x + a.y + a.z + a.w;
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = __ldg(p + i), g = __ldg(t + i);
        acc[0] += a * g; acc[1] += g; acc[2] += a;
    }
    block_atomic_add<3>(acc, sums + (int64_t)b * 3);
}

__device__ __forceinline__ void logits_terms(float x0, float x1, float lab, float lam, float (&acc)[6]) {
    // softmax over 2 classes, one-hot of the label, hardness weight w = lam*|p - t| + (1 - lam) (same for both classes)
    const float m = fmaxf(x0, x1);
    const float e0 = __expf(x0 - m), e1 = __expf(x1 - m);
    const float inv = 1.0f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    const float t1 = lab >= 0.5f ? 1.f : 0.f, t0 = 1.f - t1;   // one_hot(target.long()) for labels in {0,1}
    const float w = lam >= 0.f ? lam * fabsf(p1 - t1) + (1.f - lam) : 1.f;
    acc[0] += w * t0 * p0; acc[1] += w * t0; acc[2] += w * p0;
    acc[3] += w * t1 * p1; acc[4] += w * t1; acc[5] += w * p1;
}

// 2-class logits term: sums[b][c][0..2] += (sum w t_c p_c, sum w t_c, sum w p_c)
__global__ void __launch_bounds__(256) dice_sums_logits_kernel(const float* __restrict__ logits, const float* __restrict__ label,
                                                               int64_t n, float lam, double* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* x0 = logits + (int64_t)b * 2 * n;
    const float* x1 = x0 + n;
    const float* t = label + (int64_t)b * n;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t n4 = ((((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)t) & 15) == 0) ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x0) + i), c = __ldg(reinterpret_cast<const float4*>(x1) + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(t) + i);
        logits_terms(a.x, c.x, g.x, lam, acc);
        logits_terms(a.y, c.y, g.y, lam, acc);
        logits_terms(a.z, c.z, g.z, lam, acc);
        logits_terms(a.w, c.w, g.w, lam, acc);
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        logits_terms(__ldg(x0 + i), __ldg(x1 + i), __ldg(t + i), lam, acc);
    block_atomic_add<6>(acc, sums + (int64_t)b * 6);
}

// ---- finalise: loss and backward coefficients -----------------------------------------------------
// row r: f_r = 1 - (2 I + eps) / (G + P + eps); loss = sum_r scale_r f_r;
// coef[r] = scale_r * (alpha, beta), alpha = -2/D, beta = (2I+eps)/D^2, D = G+P+eps  
__global__ void dice_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ scale, int nrows, float eps,
                                     float* __restrict__ loss, float* __restrict__ coef) {
    if (threadIdx.x || blockIdx.x) return;
    double total = 0;
    for (int r = 0; r < nrows; ++r) {
        const double I = sums[3 * r], G = sums[3 * r + 1], P = sums[3 * r + 2];
        const double D = G + P + (double)eps, num = 2.0 * I + (double)eps;
        total += (double)scale[r] * (1.0 - num / D);
        coef[2 * r] = (float)((double)scale[r] * (-2.0 / D));
        coef[2 * r + 1] = (float)((double)scale[r] * (num / (D * D)));
    }
    *loss = (float)total;
}

// ---- backward -------------------------------------------------------------------------------------
// d loss / d att = go * (alpha g + beta)
__global__ void __launch_bounds__(256) dice_bwd1_kernel(const float* __restrict__ tgt, int64_t n, const float* __restrict__ coef,
                                                        const float* __restrict__ go, float* __restrict__ grad) {
    const int b = blockIdx.y;
    const float g0 = __ldg(go), al = __ldg(coef + 2 * b) * g0, be = __ldg(coef + 2 * b + 1) * g0;
    const float* t = tgt + (int64_t)b * n;
    float* o = grad + (int64_t)b * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = al * __ldg(t + i) + be;
}

// d loss / d logits through softmax, one-hot and the (non-detached) hardness weight
__global__ void __launch_bounds__(256) dice_bwd_logits_kernel(const float* __restrict__ logits, const float* __restrict__ label,
                                                              int64_t n, float lam, const float* __restrict__ coef,
                                                              const float* __restrict__ go, float* __restrict__ grad) {
    const int b = blockIdx.y;
    const float g0 = __ldg(go);
    const float a0 = __ldg(coef + 4 * b) * g0, b0 = __ldg(coef + 4 * b + 1) * g0;
    const float a1 = __ldg(coef + 4 * b + 2) * g0, b1 = __ldg(coef + 4 * b + 3) * g0;
    const float* x0 = logits + (int64_t)b * 2 * n;
    const float* x1 = x0 + n;
    const float* t = label + (int64_t)b * n;
    float* o0 = grad + (int64_t)b * 2 * n;
    float* o1 = o0 + n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v0 = __ldg(x0 + i), v1 = __ldg(x1 + i);
        const float m = fmaxf(v0, v1);
        const float e0 = __expf(v0 - m), e1 = __expf(v1 - m);
        const float inv = 1.0f / (e0 + e1);
        const float p0 = e0 * inv, p1 = e1 * inv;
        const float t1 = __ldg(t + i) >= 0.5f ? 1.f : 0.f, t0 = 1.f - t1;
        float gp0, gp1;
        if (lam >= 0.f) {
            const float d0 = p0 - t0, d1 = p1 - t1;
            const float s0 = d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f), s1 = d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f);
            const float w0 = lam * fabsf(d0) + (1.f - lam), w1 = lam * fabsf(d1) + (1.f - lam);
            gp0 = a0 * (w0 * t0 + lam * s0 * t0 * p0) + b0 * (lam * s0 * t0 + w0 + lam * s0 * p0);
            gp1 = a1 * (w1 * t1 + lam * s1 * t1 * p1) + b1 * (lam * s1 * t1 + w1 + lam * s1 * p1);
        } else {
            gp0 = a0 * t0 + b0;
            gp1 = a1 * t1 + b1;
        }
        const float dot = p0 * gp0 + p1 * gp1;   // softmax Jacobian: dx_c = p_c (g_c - sum_k p_k g_k)
        o0[i] = p0 * (gp0 - dot);
        o1[i] = p1 * (gp1 - dot);
    }
}

static unsigned loss_grid(int64_t n, int per_thread) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (n / per_thread + 255) / 256;
    int64_t cap = (int64_t)sms * 8;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_maxpool3d(const float* in, float* out, int32_t B, int32_t Xo, int32_t Yo, int32_t Zo, int32_t rx, int32_t ry,
                    int32_t rz, void* stream) {
    VSSEG_REQUIRE(in && out && B > 0 && Xo > 0 && Yo > 0 && Zo > 0 && rx > 0 && ry > 0 && rz > 0, "maxpool3d: bad arguments");
    maxpool3d_kernel<<<loss_grid((int64_t)B * Xo * Yo * Zo, 1), 256, 0, (cudaStream_t)stream>>>(in, out, B, Xo, Yo, Zo, rx, ry, rz);
    return check_launch("maxpool3d");
}

int vsseg_dice_sums(const float* pred, const float* target, int32_t B, int32_t C, int64_t n, float hardness_lambda,
                    double* sums, void* stream) {
    VSSEG_REQUIRE(pred && target && sums && B > 0 && n > 0, "dice_sums: bad arguments");
    VSSEG_REQUIRE(C == 1 || C == 2, "dice_sums: 1 channel (attention map) or 2-class logits only (got C=%d)", C);
    dim3 grid(loss_grid(n, 8), (unsigned)B);
    if (C == 1) dice_sums1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, n, sums);
    else dice_sums_logits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, n, hardness_lambda, sums);
    return check_launch("dice_sums");
}

int vsseg_dice_finalize(const double* sums, const float* row_scale, int32_t nrows, float smooth, float* loss, float* coef,
                        void* stream) {
    VSSEG_REQUIRE(sums && row_scale && loss && coef && nrows > 0, "dice_finalize: bad arguments");
    dice_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, row_scale, nrows, smooth, loss, coef);
    return check_launch("dice_finalize");
}

int vsseg_dice_backward(const float* pred, const float* target, int32_t B, int32_t C, int64_t n, float hardness_lambda,
                        const float* coef, const float* grad_out, float* grad, void* stream) {
    VSSEG_REQUIRE(target && coef && grad_out && grad && B > 0 && n > 0, "dice_backward: bad arguments");
    VSSEG_REQUIRE(C == 1 || (C == 2 && pred), "dice_backward: 1 channel or 2-class logits only (got C=%d)", C);
    dim3 grid(loss_grid(n, 4), (unsigned)B);
    if (C == 1) dice_bwd1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(target, n, coef, grad_out, grad);
    else dice_bwd_logits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, n, hardness_lambda, coef, grad_out, grad);
    return check_launch("dice_backward");
}

}  // extern "C"

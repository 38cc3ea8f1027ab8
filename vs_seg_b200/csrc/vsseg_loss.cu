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
// d loss / d att = go * (alpha g + beta); 128-bit loads and stores when the rows are 16-byte aligned
__global__ void __launch_bounds__(256) dice_bwd1_kernel(const float* __restrict__ tgt, int64_t n, const float* __restrict__ coef,
                                                        const float* __restrict__ go, float* __restrict__ grad) {
    const int b = blockIdx.y;
    const float g0 = __ldg(go), al = __ldg(coef + 2 * b) * g0, be = __ldg(coef + 2 * b + 1) * g0;
    const float* t = tgt + (int64_t)b * n;
    float* o = grad + (int64_t)b * n;
    const int64_t n4 = (((uintptr_t)t | (uintptr_t)o) & 15) == 0 ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(t) + i);
        reinterpret_cast<float4*>(o)[i] = make_float4(fmaf(al, g.x, be), fmaf(al, g.y, be), fmaf(al, g.z, be), fmaf(al, g.w, be));
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = fmaf(al, __ldg(t + i), be);
}

// one voxel of d loss / d logits through softmax, one-hot and the (non-detached) hardness weight
__device__ __forceinline__ void logits_grad(float v0, float v1, float lab, float lam, float a0, float b0, float a1, float b1,
                                            float& o0, float& o1) {
    const float m = fmaxf(v0, v1);
    const float e0 = __expf(v0 - m), e1 = __expf(v1 - m);
    const float inv = 1.0f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    const float t1 = lab >= 0.5f ? 1.f : 0.f, t0 = 1.f - t1;
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
    o0 = p0 * (gp0 - dot);
    o1 = p1 * (gp1 - dot);
}

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
    const int64_t n4 = (((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)t | (uintptr_t)o0 | (uintptr_t)o1) & 15) == 0 ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(x0) + i), v = __ldg(reinterpret_cast<const float4*>(x1) + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(t) + i);
        float4 r0, r1;
        logits_grad(u.x, v.x, g.x, lam, a0, b0, a1, b1, r0.x, r1.x);
        logits_grad(u.y, v.y, g.y, lam, a0, b0, a1, b1, r0.y, r1.y);
        logits_grad(u.z, v.z, g.z, lam, a0, b0, a1, b1, r0.z, r1.z);
        logits_grad(u.w, v.w, g.w, lam, a0, b0, a1, b1, r0.w, r1.w);
        reinterpret_cast<float4*>(o0)[i] = r0;
        reinterpret_cast<float4*>(o1)[i] = r1;
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        logits_grad(__ldg(x0 + i), __ldg(x1 + i), __ldg(t + i), lam, a0, b0, a1, b1, o0[i], o1[i]);
}

// ---- general DiceLoss (dice_spvPA.py:90-167, every flag) ------------------------------------------
// One voxel, up to DICE_MAXC channels in registers: activation (none | sigmoid | softmax over the channels), target
// (dense [B,C,n] or integer labels [B,1,n] -> one-hot), optional weight [B,C,n], optional squares.
constexpr int DICE_MAXC = 8;
struct DiceGenArgs {
    const float* pred;     // [B,C,n]
    const float* target;   // [B,C,n], or [B,1,n] labels when onehot
    const float* weight;   // [B,C,n] or NULL
    int C, act, onehot, squared;
    int64_t n;
};

template <int NC>
__device__ __forceinline__ void dice_gen_act(const DiceGenArgs& a, const float (&x)[NC], float (&p)[NC]) {
    if (a.act == 2 && NC > 1) {
        float m = x[0];
        _Pragma("unroll") for (int c = 1; c < NC; ++c) m = fmaxf(m, x[c]);
        float s = 0.f;
        _Pragma("unroll") for (int c = 0; c < NC; ++c) { p[c] = __expf(x[c] - m); s += p[c]; }
        const float inv = 1.0f / s;
        _Pragma("unroll") for (int c = 0; c < NC; ++c) p[c] *= inv;
    } else if (a.act == 1) {
        _Pragma("unroll") for (int c = 0; c < NC; ++c) p[c] = 1.0f / (1.0f + __expf(-x[c]));
    } else {
        _Pragma("unroll") for (int c = 0; c < NC; ++c) p[c] = x[c];
    }
}

// sums[b][c][0..2] += (sum w t p, sum w t' , sum w p'), t' = t^2 / p' = p^2 when squared (dice_spvPA.py:133-149)
template <int NC>
__global__ void __launch_bounds__(256) dice_gen_sums_kernel(const DiceGenArgs a, double* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* pr = a.pred + (int64_t)b * NC * a.n;
    const float* tg = a.target + (int64_t)b * (a.onehot ? 1 : NC) * a.n;
    const float* wt = a.weight ? a.weight + (int64_t)b * NC * a.n : nullptr;
    float acc[3 * NC];
#pragma unroll
    for (int i = 0; i < 3 * NC; ++i) acc[i] = 0.f;
    const bool vec = a.n % 4 == 0 && (((uintptr_t)pr | (uintptr_t)tg | (uintptr_t)wt) & 15) == 0;
    const int V = vec ? 4 : 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n / V; i += (int64_t)gridDim.x * blockDim.x) {
        float x[4][NC], t[4][NC], w[4][NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (vec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c * a.n) + i);
                x[0][c] = q.x; x[1][c] = q.y; x[2][c] = q.z; x[3][c] = q.w;
                if (!a.onehot) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(tg + c * a.n) + i);
                    t[0][c] = g.x; t[1][c] = g.y; t[2][c] = g.z; t[3][c] = g.w;
                }
                if (wt) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(wt + c * a.n) + i);
                    w[0][c] = g.x; w[1][c] = g.y; w[2][c] = g.z; w[3][c] = g.w;
                }
            } else {
                x[0][c] = __ldg(pr + c * a.n + i);
                if (!a.onehot) t[0][c] = __ldg(tg + c * a.n + i);
                if (wt) w[0][c] = __ldg(wt + c * a.n + i);
            }
        }
        float lab[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.onehot) {
            if (vec) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(tg) + i);
                lab[0] = g.x; lab[1] = g.y; lab[2] = g.z; lab[3] = g.w;
            } else {
                lab[0] = __ldg(tg + i);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= V) break;
            float p[NC];
            dice_gen_act<NC>(a, x[k], p);
            const int li = (int)lab[k];   // one_hot(target.long())
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float tc = a.onehot ? (c == li ? 1.f : 0.f) : t[k][c];
                const float wc = wt ? w[k][c] : 1.f;
                acc[3 * c] += wc * tc * p[c];
                acc[3 * c + 1] += wc * (a.squared ? tc * tc : tc);
                acc[3 * c + 2] += wc * (a.squared ? p[c] * p[c] : p[c]);
            }
        }
    }
    block_atomic_add<3 * NC>(acc, sums + (int64_t)b * 3 * DICE_MAXC);
}

// grad_pred[b][c][v] = through the activation of  w (gI t + gP (2p | 1)),  (gI, gG, gP) = gsums[b][c][0..2]
template <int NC>
__global__ void __launch_bounds__(256) dice_gen_bwd_kernel(const DiceGenArgs a, const float* __restrict__ gsums,
                                                           float* __restrict__ grad) {
    const int b = blockIdx.y;
    const float* pr = a.pred + (int64_t)b * NC * a.n;
    const float* tg = a.target + (int64_t)b * (a.onehot ? 1 : NC) * a.n;
    const float* wt = a.weight ? a.weight + (int64_t)b * NC * a.n : nullptr;
    float* go = grad + (int64_t)b * NC * a.n;
    float gI[NC], gP[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        gI[c] = __ldg(gsums + ((int64_t)b * DICE_MAXC + c) * 3);
        gP[c] = __ldg(gsums + ((int64_t)b * DICE_MAXC + c) * 3 + 2);
    }
    const bool vec = a.n % 4 == 0 && (((uintptr_t)pr | (uintptr_t)tg | (uintptr_t)wt | (uintptr_t)go) & 15) == 0;
    const int V = vec ? 4 : 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n / V; i += (int64_t)gridDim.x * blockDim.x) {
        float x[4][NC], t[4][NC], w[4][NC], g[4][NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (vec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c * a.n) + i);
                x[0][c] = q.x; x[1][c] = q.y; x[2][c] = q.z; x[3][c] = q.w;
                if (!a.onehot) {
                    const float4 h = __ldg(reinterpret_cast<const float4*>(tg + c * a.n) + i);
                    t[0][c] = h.x; t[1][c] = h.y; t[2][c] = h.z; t[3][c] = h.w;
                }
                if (wt) {
                    const float4 h = __ldg(reinterpret_cast<const float4*>(wt + c * a.n) + i);
                    w[0][c] = h.x; w[1][c] = h.y; w[2][c] = h.z; w[3][c] = h.w;
                }
            } else {
                x[0][c] = __ldg(pr + c * a.n + i);
                if (!a.onehot) t[0][c] = __ldg(tg + c * a.n + i);
                if (wt) w[0][c] = __ldg(wt + c * a.n + i);
            }
        }
        float lab[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.onehot) {
            if (vec) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(tg) + i);
                lab[0] = h.x; lab[1] = h.y; lab[2] = h.z; lab[3] = h.w;
            } else {
                lab[0] = __ldg(tg + i);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= V) break;
            float p[NC], gp[NC];
            dice_gen_act<NC>(a, x[k], p);
            const int li = (int)lab[k];
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float tc = a.onehot ? (c == li ? 1.f : 0.f) : t[k][c];
                const float wc = wt ? w[k][c] : 1.f;
                gp[c] = wc * (gI[c] * tc + gP[c] * (a.squared ? 2.f * p[c] : 1.f));
                dot += p[c] * gp[c];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float d;
                if (a.act == 2 && NC > 1) d = p[c] * (gp[c] - dot);
                else if (a.act == 1) d = gp[c] * p[c] * (1.f - p[c]);
                else d = gp[c];
                g[k][c] = d;
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (vec) reinterpret_cast<float4*>(go + c * a.n)[i] = make_float4(g[0][c], g[1][c], g[2][c], g[3][c]);
            else go[c * a.n + i] = g[0][c];
        }
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

int vsseg_dice_general_sums(const float* pred, const float* target, const float* weight, int32_t B, int32_t C, int64_t n,
                            int32_t act, int32_t target_is_labels, int32_t squared, double* sums, void* stream) {
    VSSEG_REQUIRE(pred && target && sums && B > 0 && n > 0 && C >= 1 && C <= DICE_MAXC && act >= 0 && act <= 2,
                  "dice_general_sums: bad arguments (1 <= C <= %d, act in 0..2)", DICE_MAXC);
    DiceGenArgs a{pred, target, weight, C, act, target_is_labels ? 1 : 0, squared ? 1 : 0, n};
    dim3 grid(loss_grid(n, 8), (unsigned)B);
    switch (C) {
#define VSSEG_DICE_CASE(N) case N: dice_gen_sums_kernel<N><<<grid, 256, 0, (cudaStream_t)stream>>>(a, sums); break;
        VSSEG_DICE_CASE(1) VSSEG_DICE_CASE(2) VSSEG_DICE_CASE(3) VSSEG_DICE_CASE(4)
        VSSEG_DICE_CASE(5) VSSEG_DICE_CASE(6) VSSEG_DICE_CASE(7) VSSEG_DICE_CASE(8)
#undef VSSEG_DICE_CASE
    }
    return check_launch("dice_general_sums");
}

int vsseg_dice_general_backward(const float* pred, const float* target, const float* weight, int32_t B, int32_t C, int64_t n,
                                int32_t act, int32_t target_is_labels, int32_t squared, const float* grad_sums, float* grad,
                                void* stream) {
    VSSEG_REQUIRE(pred && target && grad_sums && grad && B > 0 && n > 0 && C >= 1 && C <= DICE_MAXC && act >= 0 && act <= 2,
                  "dice_general_backward: bad arguments (1 <= C <= %d, act in 0..2)", DICE_MAXC);
    DiceGenArgs a{pred, target, weight, C, act, target_is_labels ? 1 : 0, squared ? 1 : 0, n};
    dim3 grid(loss_grid(n, 4), (unsigned)B);
    switch (C) {
#define VSSEG_DICE_CASE(N) case N: dice_gen_bwd_kernel<N><<<grid, 256, 0, (cudaStream_t)stream>>>(a, grad_sums, grad); break;
        VSSEG_DICE_CASE(1) VSSEG_DICE_CASE(2) VSSEG_DICE_CASE(3) VSSEG_DICE_CASE(4)
        VSSEG_DICE_CASE(5) VSSEG_DICE_CASE(6) VSSEG_DICE_CASE(7) VSSEG_DICE_CASE(8)
#undef VSSEG_DICE_CASE
    }
    return check_launch("dice_general_backward");
}

}  // extern "C"

// Fused multi-tensor Adam over one flat fp32 buffer (reference: torch.optim.Adam built at
// params/VSparams.py:388-391 - lr 1e-4, weight_decay 1e-7 as L2-in-gradient, default betas/eps;
// the reference steps 178 small tensors one by one).  HBM-bound: 16 B read + 12 B written per element.
#include "vsseg_common.cuh"

namespace vsseg {

__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, int64_t n4, float* __restrict__ pt,
                                                        const float* __restrict__ gt, float* __restrict__ mt, float* __restrict__ vt,
                                                        int ntail, float beta1, float beta2, float eps, float wd, float step_size,
                                                        float inv_bc2_sqrt, float grad_scale) {
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg = fmaf(wd, pp, gg * grad_scale);                 // grad (averaged over ranks) + weight_decay * param
        mm = fmaf(beta1, mm, (1.f - beta1) * gg);
        vv = fmaf(beta2, vv, (1.f - beta2) * gg * gg);
        const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
        pp -= step_size * (mm / denom);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        const float4 gg = g[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) {
        const int i = threadIdx.x;
        float pp = pt[i], mm = mt[i], vv = vt[i];
        upd(pp, gt[i], mm, vv);
        pt[i] = pp; mt[i] = mm; vt[i] = vv;
    }
}

}  // namespace vsseg

using namespace vsseg;

extern "C" int vsseg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
    VSSEG_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
    VSSEG_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                  "adam_step: buffers must be 16-byte aligned");
    // bias corrections in double on the host, exactly as torch.optim.Adam computes them
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    const int64_t n4 = n / 4;
    const int ntail = (int)(n - n4 * 4);
    int sms = 148;
    long blocks = (long)((n4 + 255) / 256);
    if (blocks > sms * 8) blocks = sms * 8;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4, param + n4 * 4, grad + n4 * 4, exp_avg + n4 * 4,
        exp_avg_sq + n4 * 4, ntail, beta1, beta2, eps, weight_decay, step_size, inv_bc2_sqrt, grad_scale);
    return check_launch("adam_step");
}

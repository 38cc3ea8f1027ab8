// Fused multi-tensor Adam over one flat fp32 buffer (reference: torch.optim.Adam built at
// params/VSparams.py:388-391 - lr 1e-4, weight_decay 1e-7 as L2-in-gradient, default betas/eps;
// the reference steps 178 small tensors one by one).  HBM-bound: 16 B read + 12 B written per element.
#include "vsseg_common.cuh"

namespace vsseg {

__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, int64_t n4, float* __restrict__ pt,
                                                        const float* __restrict__ gt, float* __restrict__ mt, float* __restrict__ vt,
                                                        int ntail, float beta1, float beta2, float eps, float wd, float step_size,
                                                        float inv_bc2_sqrt, float grad_scale, const long long* __restrict__ step_dev,
                                                        const float* __restrict__ lr_dev) {
    if (step_dev) {
        // graph-capturable form: the step count and the learning rate live in device memory (a captured launch must not
        // bake them in); bias corrections in double like torch.optim.Adam, once per block
        __shared__ float sh[2];
        if (threadIdx.x == 0) {
            const double t = (double)*step_dev;
            const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
            sh[0] = (float)((double)*lr_dev / bc1);
            sh[1] = (float)(1.0 / sqrt(bc2));
        }
        __syncthreads();
        step_size = sh[0];
        inv_bc2_sqrt = sh[1];
    }
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
                               float beta2, float eps, float weight_decay, int64_t step, float grad_scale, const int64_t* step_dev,
                               const float* lr_dev, void* stream) {
    VSSEG_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && (step >= 1 || step_dev), "adam_step: bad arguments");
    VSSEG_REQUIRE(!step_dev == !lr_dev, "adam_step: step_dev and lr_dev go together");
    if (step_dev) step = 1;   // unused by the kernel
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
        exp_avg_sq + n4 * 4, ntail, beta1, beta2, eps, weight_decay, step_size, inv_bc2_sqrt, grad_scale, (const long long*)step_dev,
        lr_dev);
    return check_launch("adam_step");
}

// ---- weight image of the tensor-core conv, built on the device -------------------------------------------------
// Training re-packs every conv weight twice per step (forward conv and the adjoint conv of the data gradient); done
// with torch ops that is ~12 tiny launches per conv (pad, flip, permute, two casts, subtract, stack, copy: ~1000
// launches per step, the step was launch-bound).  One launch per conv here; same image as
// vs_seg_b200.engine.pack_conv_weight_tc, bit for bit.
namespace vsseg {
struct PackArgs {
    const float* w;
    __nv_bfloat16* out;
    int d0, d1, kx, ky, kz, conv_t, phases, flip, cin_pad, n_cta, n_split, nj;
    long long total;   // elements of one plane
};
__global__ void __launch_bounds__(256) pack_weight_tc_kernel(const PackArgs a) {
    const int cout_real = a.conv_t ? a.d1 : a.d0, cin_real = a.conv_t ? a.d0 : a.d1;
    const int nsel = a.n_split * (a.phases ? 2 : 1), nc16 = a.cin_pad / 16;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int e = (int)(t % 8); t /= 8;
        const int n = (int)(t % a.n_cta); t /= a.n_cta;
        const int typ = (int)(t % a.ky); t /= a.ky;
        const int kh = (int)(t % 2); t /= 2;
        const int tz = (int)(t % a.kz); t /= a.kz;
        const int j = (int)(t % a.nj); t /= a.nj;
        const int c16 = (int)(t % nc16); t /= nc16;
        const int sel = (int)t;
        const int px = a.phases ? sel / a.n_split : 0, ns = a.phases ? sel % a.n_split : sel;
        const int co = ns * a.n_cta + n, ci = c16 * 16 + kh * 8 + e;
        int tx = j, ty = a.phases ? typ : a.ky - 1 - typ;
        bool zero = co >= cout_real || ci >= cin_real;
        if (a.phases) {   // sub-pixel phases of a stride-2 transposed conv: px = 0 -> centre tap only; px = 1 -> taps 2, 0
            if (px == 0) { tx = 1; zero = zero || j == 1; }
            else tx = j == 0 ? 2 : 0;
        }
        float v = 0.f;
        if (!zero) {
            int fx = tx, fy = ty, fz = tz;
            if (a.flip) { fx = a.kx - 1 - tx; fy = a.ky - 1 - ty; fz = a.kz - 1 - tz; }
            const long long i0 = a.conv_t ? ci : co, i1 = a.conv_t ? co : ci;
            v = a.w[(((i0 * a.d1 + i1) * a.kx + fx) * a.ky + fy) * a.kz + fz];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        // out index: [sel][c16][j][plane][tz][kh][typ][n][e]
        const long long inner = (((long long)tz * 2 + kh) * a.ky + typ) * a.n_cta * 8 + (long long)n * 8 + e;
        const long long plane_sz = (long long)a.kz * 2 * a.ky * a.n_cta * 8;
        const long long base = ((((long long)sel * nc16 + c16) * a.nj + j) * 2) * plane_sz;
        a.out[base + inner] = h;
        a.out[base + plane_sz + inner] = l;
    }
    (void)nsel;
}
}  // namespace vsseg

extern "C" int vsseg_pack_conv_weight_tc(const float* w, int32_t d0, int32_t d1, int32_t kx, int32_t ky, int32_t kz, int32_t conv_t_layout,
                                         int32_t phases, int32_t flip, int32_t cin_pad, int32_t cout_pad, int32_t n_split,
                                         void* out_bf16, void* stream) {
    VSSEG_REQUIRE(w && out_bf16 && d0 > 0 && d1 > 0 && kx > 0 && ky > 0 && kz > 0, "pack_conv_weight_tc: bad arguments");
    VSSEG_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0 && n_split >= 1 && cout_pad % n_split == 0 && (cout_pad / n_split) % 16 == 0,
                  "pack_conv_weight_tc: Cin % 16, Cout % (16 * n_split) must be 0 after padding");
    VSSEG_REQUIRE(!phases || kx == 3, "pack_conv_weight_tc: phase packing needs kx = 3");
    vsseg::PackArgs a;
    a.w = w; a.out = (__nv_bfloat16*)out_bf16;
    a.d0 = d0; a.d1 = d1; a.kx = kx; a.ky = ky; a.kz = kz; a.conv_t = conv_t_layout; a.phases = phases; a.flip = flip;
    a.cin_pad = cin_pad; a.n_split = n_split; a.n_cta = cout_pad / n_split; a.nj = phases ? 2 : kx;
    const long long nsel = (long long)n_split * (phases ? 2 : 1);
    a.total = nsel * (cin_pad / 16) * a.nj * kz * 2 * ky * a.n_cta * 8;
    long blocks = (long)((a.total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    vsseg::pack_weight_tc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return vsseg::check_launch("pack_conv_weight_tc");
}

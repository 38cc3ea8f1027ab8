// Training-mode kernels of the VS_Seg hot path (reference params/networks/blocks/convolutions.py:148-156:
// Conv -> BatchNorm3d(batch statistics) -> Dropout -> PReLU, and its backward), sm_100a.
// The convolutions themselves (forward and data-gradient) run on the tcgen05 kernel of vsseg_tc.cu: the
// data-gradient of a conv is the transposed conv with the same weights and vice versa.  This file holds
// the bandwidth-bound pieces around them and the weight-gradient reduction:
//   bn_stats / bn_finalize            batch mean / biased variance, running-stat update (momentum, unbiased var)
//   bn_act_fwd                        y = PReLU(dropout(c*scale + shift)) [+ residual]
//   bn_act_bwd_reduce / _apply        d(beta), d(gamma), d(PReLU slope), then dc
//   act_bwd                           ReLU backward of the attention conv1
//   conv3d_wgrad / cin1_wgrad         dW[tap][ci][co] = sum_v x[v_in] * dc[v_out], d(bias)
//   smallcout_bwd                     backward of the 1-2 output-channel convs (attention conv2 + sigmoid, logits)
//   gate_bwd                          backward of x*(1+att)
//   act8_add                          gradient accumulation where a tensor has two consumers
#include "vsseg_common.cuh"

namespace vsseg {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dropout keep-mask: a counter-based hash of (seed, logical NCDHW element index); the same function is
// evaluated in forward and backward, so no mask is stored
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// the 8 keep-scales of one (batch, channel group, voxel) triple from TWO 64-bit hashes (16 bits per element: p is honoured
// to 2^-16); one hash per element made the BatchNorm/activation kernels ALU-bound (31-36 % of the DRAM peak under ncu)
__device__ __forceinline__ void keep_scale8(uint64_t seed, uint64_t group, float p, float (&ks)[8]) {
    if (p <= 0.f) {
#pragma unroll
        for (int k = 0; k < 8; ++k) ks[k] = 1.f;
        return;
    }
    const uint64_t h0 = mix64(seed + (2 * group) * 0x9E3779B97F4A7C15ull), h1 = mix64(seed + (2 * group + 1) * 0x9E3779B97F4A7C15ull);
    const uint32_t thr = (uint32_t)(p * 65536.0f);
    const float inv = 1.0f / (1.0f - p);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t u = (uint32_t)(((k < 4 ? h0 : h1) >> (16 * (k & 3))) & 0xFFFFu);
        ks[k] = u >= thr ? inv : 0.f;
    }
}
__device__ __forceinline__ float keep_scale(uint64_t seed, uint64_t elem, float p) {
    if (p <= 0.f) return 1.f;
    uint64_t z = seed + elem * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    return u >= p ? 1.0f / (1.0f - p) : 0.f;
}

struct G8 {  // iteration over (b, channel group, voxel) triples of an act8 tensor
    int64_t nvox, total;
    int CG;
    __device__ G8(const vsseg_act8& t) : nvox((int64_t)t.X * t.Y * t.Z), CG(t.C / 8) { total = nvox * CG * t.B; }
    __device__ void split(int64_t i, int& b, int& cg, int64_t& v) const {
        v = i % nvox;
        cg = (int)((i / nvox) % CG);
        b = (int)(i / (nvox * CG));
    }
};
__device__ __forceinline__ const __nv_bfloat16* g8_ptr(const vsseg_act8& t, int b, int cg, int64_t v, int64_t nvox) {
    return (const __nv_bfloat16*)t.hi + (int64_t)b * t.batch_stride + ((int64_t)cg * nvox + v) * 8;
}
__device__ __forceinline__ void g8_load(const vsseg_act8& t, int b, int cg, int64_t v, int64_t nvox, float (&f)[8]) {
    const __nv_bfloat16* p = g8_ptr(t, b, cg, v, nvox);
    unpack8(ldg128(p), ldg128(p + t.lo_offset), f);
}
__device__ __forceinline__ void g8_store(const vsseg_act8& t, int b, int cg, int64_t v, int64_t nvox, const float (&f)[8]) {
    __nv_bfloat16* p = (__nv_bfloat16*)t.hi + (int64_t)b * t.batch_stride + ((int64_t)cg * nvox + v) * 8;
    uint4 h, l;
    pack8(f, h, l);
    *reinterpret_cast<uint4*>(p) = h;
    *reinterpret_cast<uint4*>(p + t.lo_offset) = l;
}

// ---- batch statistics ----------------------------------------------------------------------------------
// grid.y = channel group; every block reduces a slice of (b, voxel) and adds 8 sums + 8 sums of squares
__global__ void __launch_bounds__(256) bn_stats_kernel(vsseg_act8 x, double* __restrict__ sums) {
    const int cg = blockIdx.y, b = blockIdx.z;
    const int64_t nvox = (int64_t)x.X * x.Y * x.Z;
    float s[8], q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) s[c] = q[c] = 0.f;
    double ds[8], dq[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) ds[c] = dq[c] = 0.0;
    int n = 0;
    // four independent 32-byte groups in flight per thread (one load pair per iteration left the reduction at a third
    // of the DRAM peak under ncu)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvox; i += 4 * stride) {
        uint4 h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t v = i + u * stride;
            if (v < nvox) {
                const __nv_bfloat16* p = g8_ptr(x, b, cg, v, nvox);
                h[u] = ldg128(p);
                l[u] = ldg128(p + x.lo_offset);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u * stride >= nvox) break;
            float f[8];
            unpack8(h[u], l[u], f);
#pragma unroll
            for (int c = 0; c < 8; ++c) { s[c] += f[c]; q[c] += f[c] * f[c]; }
        }
        if (++n == 16) {  // bounded fp32 partials (64 values), merged in fp64
#pragma unroll
            for (int c = 0; c < 8; ++c) { ds[c] += s[c]; dq[c] += q[c]; s[c] = q[c] = 0.f; }
            n = 0;
        }
    }
    __shared__ double red[16][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double a = warp_sum_d(ds[c] + s[c]), b = warp_sum_d(dq[c] + q[c]);
        if (lane == 0) { red[c][warp] = a; red[8 + c][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        const int c = cg * 8 + (threadIdx.x & 7);
        atomicAdd(sums + (threadIdx.x < 8 ? 0 : x.C) + c, t);
    }
}

// stats layout: [scale | shift | mean | rstd][C]
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* __restrict__ stats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;   // biased, as used for normalisation
    if (var < 0) var = 0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double sc = (double)gamma[c] * rstd;
    stats[c] = (float)sc;
    stats[C + c] = (float)((double)beta[c] - mean * sc);
    stats[2 * C + c] = (float)mean;
    stats[3 * C + c] = (float)rstd;
    if (running_mean) {
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// ---- forward: y = PReLU(dropout(c*scale + shift)) [+ residual] ------------------------------------------
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(vsseg_act8 c, vsseg_act8 y, const float* __restrict__ stats, const float* __restrict__ slope_p,
                                                         float drop_p, uint64_t seed, const uint64_t* __restrict__ seed_base, vsseg_act8 res,
                                                         int has_res) {
    const float slope = __ldg(slope_p);   // the PReLU parameter is read on the device: no host copy, graph-capturable
    if (seed_base) seed += __ldg(reinterpret_cast<const unsigned long long*>(seed_base));
    // grid.y = (batch, channel group): no index arithmetic per element, the 16 per-channel constants live in registers
    const int CG = c.C / 8, b = blockIdx.y / CG, cg = blockIdx.y % CG;
    const int64_t nvox = (int64_t)c.X * c.Y * c.Z;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = __ldg(stats + cg * 8 + k); sh[k] = __ldg(stats + c.C + cg * 8 + k); }
    const uint64_t gbase = (uint64_t)blockIdx.y * (uint64_t)nvox;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
        float f[8], r[8], ks[8];
        g8_load(c, b, cg, v, nvox, f);
        if (has_res) g8_load(res, b, cg, v, nvox, r);
        keep_scale8(seed, gbase + (uint64_t)v, drop_p, ks);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float u = fmaf(f[k], sc[k], sh[k]) * ks[k];
            u = u >= 0.f ? u : u * slope;
            f[k] = has_res ? u + r[k] : u;
        }
        g8_store(y, b, cg, v, nvox, f);
    }
}

// ---- backward reductions: sums[0][C] = sum du, sums[1][C] = sum du*xhat, sums[2C] = d(slope) -------------
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(vsseg_act8 c, vsseg_act8 dy, const float* __restrict__ stats,
                                                                const float* __restrict__ slope_p, float drop_p, uint64_t seed,
                                                                const uint64_t* __restrict__ seed_base, double* __restrict__ sums) {
    const float slope = __ldg(slope_p);
    if (seed_base) seed += __ldg(reinterpret_cast<const unsigned long long*>(seed_base));
    const int cg = blockIdx.y, b = blockIdx.z, CG = c.C / 8;
    const int64_t nvox = (int64_t)c.X * c.Y * c.Z;
    float s[8], q[8], da = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
    double ds[8], dq[8], dda = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) ds[k] = dq[k] = 0.0;
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ch = cg * 8 + k;
        sc[k] = __ldg(stats + ch); sh[k] = __ldg(stats + c.C + ch); mu[k] = __ldg(stats + 2 * c.C + ch); rs[k] = __ldg(stats + 3 * c.C + ch);
    }
    const uint64_t gbase = ((uint64_t)b * CG + cg) * (uint64_t)nvox;
    int n = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v0 < nvox; v0 += 2 * stride) {
        uint4 ch[2], cl[2], dh[2], dl[2];   // two independent groups (8 x 16 B) in flight per thread
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t v = v0 + u * stride;
            if (v < nvox) {
                const __nv_bfloat16* pc = g8_ptr(c, b, cg, v, nvox);
                const __nv_bfloat16* pd = g8_ptr(dy, b, cg, v, nvox);
                ch[u] = ldg128(pc); cl[u] = ldg128(pc + c.lo_offset);
                dh[u] = ldg128(pd); dl[u] = ldg128(pd + dy.lo_offset);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t v = v0 + u * stride;
            if (v >= nvox) break;
            float f[8], g[8], ks[8];
            unpack8(ch[u], cl[u], f);
            unpack8(dh[u], dl[u], g);
            keep_scale8(seed, gbase + (uint64_t)v, drop_p, ks);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float w = fmaf(f[k], sc[k], sh[k]) * ks[k];   // PReLU input
                const float dv = w >= 0.f ? g[k] : g[k] * slope;
                if (w < 0.f) da += g[k] * w;
                const float du = dv * ks[k];
                const float xhat = (f[k] - mu[k]) * rs[k];
                s[k] += du;
                q[k] += du * xhat;
            }
        }
        if (++n == 32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { ds[k] += s[k]; dq[k] += q[k]; s[k] = q[k] = 0.f; }
            dda += da; da = 0.f;
            n = 0;
        }
    }
    __shared__ double red[17][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double a = warp_sum_d(ds[k] + s[k]), b2 = warp_sum_d(dq[k] + q[k]);
        if (lane == 0) { red[k][warp] = a; red[8 + k][warp] = b2; }
    }
    {
        const double a = warp_sum_d(dda + da);
        if (lane == 0) red[16][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x < 17) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        if (threadIdx.x == 16) atomicAdd(sums + 2 * c.C, t);
        else atomicAdd(sums + (threadIdx.x < 8 ? 0 : c.C) + cg * 8 + (threadIdx.x & 7), t);
    }
}

// dc = scale * (du - mean(du) - xhat * mean(du*xhat))
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(vsseg_act8 c, vsseg_act8 dy, const float* __restrict__ stats,
                                                               const double* __restrict__ sums, double inv_count,
                                                               const float* __restrict__ slope_p, float drop_p, uint64_t seed,
                                                               const uint64_t* __restrict__ seed_base, vsseg_act8 dc) {
    const float slope = __ldg(slope_p);
    if (seed_base) seed += __ldg(reinterpret_cast<const unsigned long long*>(seed_base));
    const int CG = c.C / 8, b = blockIdx.y / CG, cg = blockIdx.y % CG;
    const int64_t nvox = (int64_t)c.X * c.Y * c.Z;
    float sc[8], sh[8], mu[8], rs[8], mb[8], mg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ch = cg * 8 + k;
        sc[k] = __ldg(stats + ch); sh[k] = __ldg(stats + c.C + ch); mu[k] = __ldg(stats + 2 * c.C + ch); rs[k] = __ldg(stats + 3 * c.C + ch);
        mb[k] = (float)(sums[ch] * inv_count); mg[k] = (float)(sums[c.C + ch] * inv_count);
    }
    const uint64_t gbase = (uint64_t)blockIdx.y * (uint64_t)nvox;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
        float f[8], g[8], ks[8];
        g8_load(c, b, cg, v, nvox, f);
        g8_load(dy, b, cg, v, nvox, g);
        keep_scale8(seed, gbase + (uint64_t)v, drop_p, ks);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = fmaf(f[k], sc[k], sh[k]) * ks[k];
            const float du = (w >= 0.f ? g[k] : g[k] * slope) * ks[k];
            const float xhat = (f[k] - mu[k]) * rs[k];
            f[k] = sc[k] * (du - mb[k] - xhat * mg[k]);
        }
        g8_store(dc, b, cg, v, nvox, f);
    }
}

// dc = dy * (y > 0 ? 1 : slope)      (attention conv1: ReLU, slope 0)
__global__ void __launch_bounds__(256) act_bwd_kernel(vsseg_act8 y, vsseg_act8 dy, float slope, vsseg_act8 dc) {
    const G8 it(y);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < it.total; i += (int64_t)gridDim.x * blockDim.x) {
        int b, cg;
        int64_t v;
        it.split(i, b, cg, v);
        float f[8], g[8];
        g8_load(y, b, cg, v, it.nvox, f);
        g8_load(dy, b, cg, v, it.nvox, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = f[k] > 0.f ? g[k] : g[k] * slope;
        g8_store(dc, b, cg, v, it.nvox, g);
    }
}

__global__ void __launch_bounds__(256) act8_add_kernel(vsseg_act8 a, vsseg_act8 b2, vsseg_act8 out) {
    const G8 it(a);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < it.total; i += (int64_t)gridDim.x * blockDim.x) {
        int b, cg;
        int64_t v;
        it.split(i, b, cg, v);
        float f[8], g[8];
        g8_load(a, b, cg, v, it.nvox, f);
        g8_load(b2, b, cg, v, it.nvox, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += g[k];
        g8_store(out, b, cg, v, it.nvox, f);
    }
}

// ---- weight gradient --------------------------------------------------------------------------------------
// dW[tap][ci][co] += sum over the M grid (output voxels; input voxels of a transposed conv) of x[v_in][ci] * dc[v_out][co]
// block = (tap, ci group of 8, co group of 16, voxel slice); thread accumulates 8x16 products over its voxels,
// the block reduces through shared memory and issues one atomicAdd per weight.
struct WgradArgs {
    vsseg_act8 x, dc;
    vsseg_conv_geom g;
    float* dw;
    float* dbias;
    int cout_pad;
    int nslice;
};

__global__ void __launch_bounds__(128) conv_wgrad_kernel(const WgradArgs a) {
    const int taps = a.g.kx * a.g.ky * a.g.kz;
    int bid = blockIdx.x;
    const int slice = bid % a.nslice; bid /= a.nslice;
    const int cog = bid % (a.cout_pad / 16); bid /= (a.cout_pad / 16);
    const int cig = bid % (a.x.C / 8); bid /= (a.x.C / 8);
    const int tap = bid;
    const int tx = tap / (a.g.ky * a.g.kz), ty = (tap / a.g.kz) % a.g.ky, tz = tap % a.g.kz;
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    const vsseg_act8& m = a.g.transposed ? a.x : a.dc;   // the M grid
    const int64_t nm = (int64_t)m.X * m.Y * m.Z, total = nm * m.B;
    const int64_t nx = (int64_t)a.x.X * a.x.Y * a.x.Z, nd = (int64_t)a.dc.X * a.dc.Y * a.dc.Z;
    const bool has_co2 = cog * 16 + 8 < a.dc.C;
    float acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    float bsum[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bsum[j] = 0.f;
    const bool do_bias = a.dbias && tap == 0 && cig == 0;
    for (int64_t i = (int64_t)slice * blockDim.x + threadIdx.x; i < total; i += (int64_t)a.nslice * blockDim.x) {
        const int b = (int)(i / nm);
        const int64_t v = i % nm;
        const int z = (int)(v % m.Z), y = (int)((v / m.Z) % m.Y), x = (int)(v / ((int64_t)m.Z * m.Y));
        int64_t vx, vd;
        if (!a.g.transposed) {
            const int xi = x * a.g.sx - px + tx, yi = y * a.g.sy - py + ty, zi = z * a.g.sz - pz + tz;
            if (do_bias) {
                float d0[8];
                g8_load(a.dc, b, cog * 2, v, nd, d0);
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[j] += d0[j];
                if (has_co2) {
                    g8_load(a.dc, b, cog * 2 + 1, v, nd, d0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) bsum[8 + j] += d0[j];
                }
            }
            if (xi < 0 || xi >= a.x.X || yi < 0 || yi >= a.x.Y || zi < 0 || zi >= a.x.Z) continue;
            vx = ((int64_t)xi * a.x.Y + yi) * a.x.Z + zi;
            vd = v;
        } else {
            const int xo = x * a.g.sx - px + tx, yo = y * a.g.sy - py + ty, zo = z * a.g.sz - pz + tz;
            if (xo < 0 || xo >= a.dc.X || yo < 0 || yo >= a.dc.Y || zo < 0 || zo >= a.dc.Z) continue;
            vx = v;
            vd = ((int64_t)xo * a.dc.Y + yo) * a.dc.Z + zo;
        }
        float xf[8], d0[8], d1[8];
        g8_load(a.x, b, cig, vx, nx, xf);
        g8_load(a.dc, b, cog * 2, vd, nd, d0);
        if (has_co2) g8_load(a.dc, b, cog * 2 + 1, vd, nd, d1);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) d1[j] = 0.f;
        }
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i2][j] = fmaf(xf[i2], d0[j], acc[i2][j]);
                acc[i2][8 + j] = fmaf(xf[i2], d1[j], acc[i2][8 + j]);
            }
    }
    // transposed conv bias gradient: sum of dc over ALL output voxels, taken by the (tap 0, cig 0) blocks
    if (a.g.transposed && do_bias) {
        const int64_t tot_d = nd * a.dc.B;
        for (int64_t i = (int64_t)slice * blockDim.x + threadIdx.x; i < tot_d; i += (int64_t)a.nslice * blockDim.x) {
            float d0[8];
            g8_load(a.dc, (int)(i / nd), cog * 2, i % nd, nd, d0);
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[j] += d0[j];
            if (has_co2) {
                g8_load(a.dc, (int)(i / nd), cog * 2 + 1, i % nd, nd, d0);
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[8 + j] += d0[j];
            }
        }
    }
    __shared__ float red[4][8 * 16 + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float s = warp_sum_f(acc[i][j]);
            if (lane == 0) red[warp][i * 16 + j] = s;
        }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float s = warp_sum_f(bsum[j]);
        if (lane == 0) red[warp][128 + j] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 128 + 16; e += blockDim.x) {
        const float s = red[0][e] + red[1][e] + red[2][e] + red[3][e];
        if (e < 128) {
            const int ci = cig * 8 + e / 16, co = cog * 16 + e % 16;
            atomicAdd(a.dw + ((int64_t)tap * a.x.C + ci) * a.cout_pad + co, s);
        } else if (do_bias) {
            atomicAdd(a.dbias + cog * 16 + (e - 128), s);
        }
    }
}

// first conv (Cin = 1, fp32 strided source): dW[tap][co] += sum_v src[v + tap] * dc[v][co]; dbias[co] += sum_v dc[v][co]
struct Cin1WgradArgs {
    vsseg_f32view src;
    vsseg_act8 dc;
    vsseg_conv_geom g;
    float* dw;
    float* dbias;
};
template <int TAPS>
__global__ void __launch_bounds__(128) cin1_wgrad_kernel(const Cin1WgradArgs a) {
    constexpr int taps = TAPS;
    const int cg = blockIdx.y;
    const int64_t nvox = (int64_t)a.dc.X * a.dc.Y * a.dc.Z, total = nvox * a.dc.B;
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    float acc[TAPS][8], bs[8];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[k] = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / nvox);
        const int64_t v = i % nvox;
        const int z = (int)(v % a.dc.Z), y = (int)((v / a.dc.Z) % a.dc.Y), x = (int)(v / ((int64_t)a.dc.Z * a.dc.Y));
        float d[8];
        g8_load(a.dc, b, cg, v, nvox, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) bs[k] += d[k];
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            const int tx = t / (a.g.ky * a.g.kz), ty = (t / a.g.kz) % a.g.ky, tz = t % a.g.kz;
            const int xi = x - px + tx, yi = y - py + ty, zi = z - pz + tz;
            if (xi < 0 || xi >= a.dc.X || yi < 0 || yi >= a.dc.Y || zi < 0 || zi >= a.dc.Z) continue;
            const float s = __ldg(a.src.ptr + b * a.src.sb + xi * a.src.sx + yi * a.src.sy + zi * a.src.sz);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(s, d[k], acc[t][k]);
        }
    }
    const int lane = threadIdx.x & 31;
    const int C = a.dc.C;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = warp_sum_f(acc[t][k]);
            if (lane == 0) atomicAdd(a.dw + t * C + cg * 8 + k, s);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float s = warp_sum_f(bs[k]);
        if (lane == 0 && a.dbias) atomicAdd(a.dbias + cg * 8 + k, s);
    }
}

// ---- small-Cout conv backward (Cout = 1 attention conv2 with sigmoid, Cout = 2 logits, stride 1) -----------
// dz[co][v] = dy[co][v] * (sigmoid ? y(1-y) : 1).  Outputs: dx (act8, all Cin), dW[tap][ci][co], dbias[co].
struct SmallBwdArgs {
    vsseg_act8 x, dx;
    vsseg_f32view dy, y;   // y: forward output (sigmoid) or unused
    vsseg_conv_geom g;
    const float* w;        // [taps][Cin][COUT]
    float* dw;
    float* dbias;
    int sigmoid;
};
template <int COUT>
__global__ void __launch_bounds__(128) smallcout_dgrad_kernel(const SmallBwdArgs a) {
    extern __shared__ float ws[];
    const int taps = a.g.kx * a.g.ky * a.g.kz, C = a.x.C;
    for (int i = threadIdx.x; i < taps * C * COUT; i += blockDim.x) ws[i] = a.w[i];
    __syncthreads();
    const int X = a.x.X, Y = a.x.Y, Z = a.x.Z;
    const int64_t nvox = (int64_t)X * Y * Z, total = nvox * a.x.B * (C / 8);
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i % nvox;
        const int cg = (int)((i / nvox) % (C / 8)), b = (int)(i / (nvox * (C / 8)));
        const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)(v / ((int64_t)Z * Y));
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        int tap = 0;
        for (int tx = 0; tx < a.g.kx; ++tx)
            for (int ty = 0; ty < a.g.ky; ++ty)
                for (int tz = 0; tz < a.g.kz; ++tz, ++tap) {
                    // dx[v] += W[tap] * dz[v - (tap - pad)]
                    const int xo = x + px - tx, yo = y + py - ty, zo = z + pz - tz;
                    if (xo < 0 || xo >= X || yo < 0 || yo >= Y || zo < 0 || zo >= Z) continue;
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        float d = __ldg(a.dy.ptr + b * a.dy.sb + co * a.dy.sc + xo * a.dy.sx + yo * a.dy.sy + zo * a.dy.sz);
                        if (a.sigmoid) {
                            const float s = __ldg(a.y.ptr + b * a.y.sb + co * a.y.sc + xo * a.y.sx + yo * a.y.sy + zo * a.y.sz);
                            d *= s * (1.f - s);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ws[(tap * C + cg * 8 + k) * COUT + co], d, acc[k]);
                    }
                }
        g8_store(a.dx, b, cg, v, nvox, acc);
    }
}
template <int COUT>
__global__ void __launch_bounds__(128) smallcout_wgrad_kernel(const SmallBwdArgs a) {
    // block = (tap, channel group); reduces over all voxels
    const int tap = blockIdx.y, cg = blockIdx.z, C = a.x.C;
    const int tx = tap / (a.g.ky * a.g.kz), ty = (tap / a.g.kz) % a.g.ky, tz = tap % a.g.kz;
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    const int X = a.x.X, Y = a.x.Y, Z = a.x.Z;
    const int64_t nvox = (int64_t)X * Y * Z, total = nvox * a.x.B;
    float acc[8][COUT], bs[COUT];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[k][co] = 0.f;
#pragma unroll
    for (int co = 0; co < COUT; ++co) bs[co] = 0.f;
    const bool do_bias = tap == 0 && cg == 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / nvox);
        const int64_t v = i % nvox;
        const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)(v / ((int64_t)Z * Y));
        float d[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            d[co] = __ldg(a.dy.ptr + b * a.dy.sb + co * a.dy.sc + x * a.dy.sx + y * a.dy.sy + z * a.dy.sz);
            if (a.sigmoid) {
                const float s = __ldg(a.y.ptr + b * a.y.sb + co * a.y.sc + x * a.y.sx + y * a.y.sy + z * a.y.sz);
                d[co] *= s * (1.f - s);
            }
            bs[co] += d[co];
        }
        const int xi = x - px + tx, yi = y - py + ty, zi = z - pz + tz;
        if (xi < 0 || xi >= X || yi < 0 || yi >= Y || zi < 0 || zi >= Z) continue;
        float f[8];
        g8_load(a.x, b, cg, ((int64_t)xi * Y + yi) * Z + zi, nvox, f);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[k][co] = fmaf(f[k], d[co], acc[k][co]);
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float s = warp_sum_f(acc[k][co]);
            if (lane == 0) atomicAdd(a.dw + ((int64_t)tap * C + cg * 8 + k) * COUT + co, s);
        }
    if (do_bias) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float s = warp_sum_f(bs[co]);
            if (lane == 0) atomicAdd(a.dbias + co, s);
        }
    }
}

// Line-structured version for Z >= 32: a block owns one channel group, one z tap and a slab of (b, x, y) lines; its
// threads walk z, so all index arithmetic is per line (the kernel above spends ~100 instructions of 64-bit div/mod per
// voxel and re-reads dy for every (tap, channel group): 2.3 ms for the logits conv of a 2 x 128^3 batch).  A thread keeps
// the KX*KY (dx, dy) taps of its 8 channels in registers and loads dy (and the sigmoid output) once per voxel.
template <int COUT, int KXY>
__global__ void __launch_bounds__(128) smallcout_wgrad_lines_kernel(const SmallBwdArgs a, int lines_per_block) {
    constexpr int KX = KXY == 9 ? 3 : 1, KY = KX;
    const int cg = blockIdx.z, tz = blockIdx.y, C = a.x.C;
    const int X = a.x.X, Y = a.x.Y, Z = a.x.Z;
    const int pz = (a.g.kz - 1) / 2;
    const int64_t nvox = (int64_t)X * Y * Z;
    const int nlines = a.x.B * X * Y;
    float acc[KXY][8][COUT], bs[COUT];
#pragma unroll
    for (int t = 0; t < KXY; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[t][k][co] = 0.f;
#pragma unroll
    for (int co = 0; co < COUT; ++co) bs[co] = 0.f;
    const bool do_bias = tz == 0 && cg == 0;
    const int l0 = blockIdx.x * lines_per_block, l1 = min(l0 + lines_per_block, nlines);
    for (int ln = l0; ln < l1; ++ln) {
        const int y = ln % Y, x = (ln / Y) % X, b = ln / (Y * X);     // per line, block-uniform
        const float* dyp = a.dy.ptr + b * a.dy.sb + x * a.dy.sx + y * a.dy.sy;
        const float* yp = a.y.ptr + b * a.y.sb + x * a.y.sx + y * a.y.sy;
        for (int z = threadIdx.x; z < Z; z += 128) {
            float d[COUT];
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                d[co] = __ldg(dyp + co * a.dy.sc + z * a.dy.sz);
                if (a.sigmoid) {
                    const float sg = __ldg(yp + co * a.y.sc + z * a.y.sz);
                    d[co] *= sg * (1.f - sg);
                }
                bs[co] += d[co];
            }
            const int zi = z - pz + tz;
            if (zi < 0 || zi >= Z) continue;
#pragma unroll
            for (int tx = 0; tx < KX; ++tx)
#pragma unroll
                for (int ty = 0; ty < KY; ++ty) {
                    const int xi = x - (KX - 1) / 2 + tx, yi = y - (KY - 1) / 2 + ty;
                    if (xi < 0 || xi >= X || yi < 0 || yi >= Y) continue;     // block-uniform
                    float f[8];
                    g8_load(a.x, b, cg, ((int64_t)xi * Y + yi) * Z + zi, nvox, f);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int co = 0; co < COUT; ++co) acc[tx * KY + ty][k][co] = fmaf(f[k], d[co], acc[tx * KY + ty][k][co]);
                }
        }
    }
    // block reduction: warp shuffle, then one atomic per (tap, channel, cout) and block
    __shared__ float red[4][KXY * 8 * COUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < KXY; ++t)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const float sum = warp_sum_f(acc[t][k][co]);
                if (lane == 0) red[warp][(t * 8 + k) * COUT + co] = sum;
            }
    __syncthreads();
    for (int i = threadIdx.x; i < KXY * 8 * COUT; i += 128) {
        const float sum = red[0][i] + red[1][i] + red[2][i] + red[3][i];
        const int co = i % COUT, k = (i / COUT) % 8, t = i / (8 * COUT);
        const int tap = t * a.g.kz + tz;      // tap = (tx * KY + ty) * KZ + tz
        atomicAdd(a.dw + ((int64_t)tap * C + cg * 8 + k) * COUT + co, sum);
    }
    if (do_bias) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float sum = warp_sum_f(bs[co]);
            if (lane == 0) atomicAdd(a.dbias + co, sum);
        }
    }
}

// ---- attention gate backward: g = x*(1+att)  =>  dx = dg*(1+att) [+ dx], datt = sum_c dg_c * x_c ------------
__global__ void __launch_bounds__(256) gate_bwd_kernel(vsseg_act8 x, vsseg_f32view att, vsseg_act8 dg, vsseg_act8 dx,
                                                       vsseg_f32view datt, int accumulate) {
    const int64_t nvox = (int64_t)x.X * x.Y * x.Z, total = nvox * x.B;
    const int CG = x.C / 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / nvox);
        const int64_t v = i % nvox;
        const int z = (int)(v % x.Z), y = (int)((v / x.Z) % x.Y), xx = (int)(v / ((int64_t)x.Z * x.Y));
        const float g1 = 1.0f + __ldg(att.ptr + b * att.sb + xx * att.sx + y * att.sy + z * att.sz);
        float da = 0.f;
        for (int cg = 0; cg < CG; ++cg) {
            float f[8], g[8], o[8];
            g8_load(x, b, cg, v, nvox, f);
            g8_load(dg, b, cg, v, nvox, g);
            if (accumulate) g8_load(dx, b, cg, v, nvox, o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                da = fmaf(g[k], f[k], da);
                o[k] = accumulate ? o[k] + g[k] * g1 : g[k] * g1;
            }
            g8_store(dx, b, cg, v, nvox, o);
        }
        datt.ptr[b * datt.sb + xx * datt.sx + y * datt.sy + z * datt.sz] = da;
    }
}

static unsigned ew_grid(int64_t total, int block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (total + block - 1) / block, cap = (int64_t)sms * 16;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}
// x extent of a (x, rows) grid of 256-thread blocks: about 16 blocks per SM in total, at least one block per row
static unsigned ew_grid_2d(int64_t nvox, int rows) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (nvox + 255) / 256, cap = ((int64_t)sms * 16 + rows - 1) / rows;
    if (cap < 1) cap = 1;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}
static bool a8ok(const vsseg_act8* t) {
    return t && t->hi && t->C > 0 && t->C % 8 == 0 && t->B > 0 && t->X > 0 && t->Y > 0 && t->Z > 0;
}
static bool same_shape(const vsseg_act8* a, const vsseg_act8* b) {
    return a->B == b->B && a->C == b->C && a->X == b->X && a->Y == b->Y && a->Z == b->Z;
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_bn_stats(const vsseg_act8* x, double* sums, void* stream) {
    VSSEG_REQUIRE(a8ok(x) && sums, "bn_stats: bad arguments");
    const int64_t nvox = (int64_t)x->X * x->Y * x->Z;
    dim3 grid(ew_grid(nvox, 256 * 8), (unsigned)(x->C / 8), (unsigned)x->B);
    bn_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*x, sums);
    return check_launch("bn_stats");
}

int vsseg_bn_finalize(const double* sums, int32_t C, int64_t count, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, float* stats, void* stream) {
    VSSEG_REQUIRE(sums && gamma && beta && stats && C > 0 && count > 0, "bn_finalize: bad arguments");
    VSSEG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running_mean/var must come together");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, C, (double)count, gamma, beta, eps, momentum,
                                                                          running_mean, running_var, stats);
    return check_launch("bn_finalize");
}

int vsseg_bn_act_fwd(const vsseg_act8* c, const vsseg_act8* y, const float* stats, const float* slope, float drop_p,
                     uint64_t seed, const uint64_t* seed_base, const vsseg_act8* residual, void* stream) {
    VSSEG_REQUIRE(a8ok(c) && a8ok(y) && same_shape(c, y) && stats && slope, "bn_act_fwd: bad arguments");
    VSSEG_REQUIRE(!residual || (a8ok(residual) && same_shape(residual, c)), "bn_act_fwd: residual shape mismatch");
    VSSEG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "bn_act_fwd: dropout probability must be in [0,1)");
    // grid.y = (batch, channel group); x covers the voxels with a few per thread so every SM holds several blocks
    const int64_t nvox = (int64_t)c->X * c->Y * c->Z;
    const int rows = c->B * (c->C / 8);
    VSSEG_REQUIRE(rows <= 65535, "bn_act_fwd: batch x channel groups exceeds the grid limit");
    dim3 grid(ew_grid_2d(nvox, rows), (unsigned)rows);
    bn_act_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*c, *y, stats, slope, drop_p, seed, seed_base,
                                                                            residual ? *residual : *c, residual ? 1 : 0);
    return check_launch("bn_act_fwd");
}

int vsseg_bn_act_bwd_reduce(const vsseg_act8* c, const vsseg_act8* dy, const float* stats, const float* slope, float drop_p,
                            uint64_t seed, const uint64_t* seed_base, double* sums, void* stream) {
    VSSEG_REQUIRE(a8ok(c) && a8ok(dy) && same_shape(c, dy) && stats && sums && slope, "bn_act_bwd_reduce: bad arguments");
    const int64_t nvox = (int64_t)c->X * c->Y * c->Z;
    dim3 grid(ew_grid(nvox, 256 * 8), (unsigned)(c->C / 8), (unsigned)c->B);
    bn_act_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*c, *dy, stats, slope, drop_p, seed, seed_base, sums);
    return check_launch("bn_act_bwd_reduce");
}

int vsseg_bn_act_bwd_apply(const vsseg_act8* c, const vsseg_act8* dy, const float* stats, const double* sums,
                           const float* slope, float drop_p, uint64_t seed, const uint64_t* seed_base, const vsseg_act8* dc,
                           void* stream) {
    VSSEG_REQUIRE(a8ok(c) && a8ok(dy) && a8ok(dc) && same_shape(c, dy) && same_shape(c, dc) && stats && sums && slope,
                  "bn_act_bwd_apply: bad arguments");
    const int64_t count = (int64_t)c->B * c->X * c->Y * c->Z;
    const int64_t nvox = (int64_t)c->X * c->Y * c->Z;
    const int rows = c->B * (c->C / 8);
    VSSEG_REQUIRE(rows <= 65535, "bn_act_bwd_apply: batch x channel groups exceeds the grid limit");
    dim3 grid(ew_grid_2d(nvox, rows), (unsigned)rows);
    bn_act_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *c, *dy, stats, sums, 1.0 / (double)count, slope, drop_p, seed, seed_base, *dc);
    return check_launch("bn_act_bwd_apply");
}

int vsseg_act_bwd(const vsseg_act8* y, const vsseg_act8* dy, float slope, const vsseg_act8* dc, void* stream) {
    VSSEG_REQUIRE(a8ok(y) && a8ok(dy) && a8ok(dc) && same_shape(y, dy) && same_shape(y, dc), "act_bwd: bad arguments");
    const int64_t total = (int64_t)y->B * (y->C / 8) * y->X * y->Y * y->Z;
    act_bwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(*y, *dy, slope, *dc);
    return check_launch("act_bwd");
}

int vsseg_act8_add(const vsseg_act8* a, const vsseg_act8* b, const vsseg_act8* out, void* stream) {
    VSSEG_REQUIRE(a8ok(a) && a8ok(b) && a8ok(out) && same_shape(a, b) && same_shape(a, out), "act8_add: bad arguments");
    const int64_t total = (int64_t)a->B * (a->C / 8) * a->X * a->Y * a->Z;
    act8_add_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(*a, *b, *out);
    return check_launch("act8_add");
}

int vsseg_conv3d_wgrad(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g, float* dw, int32_t cout_pad,
                       float* dbias, void* stream) {
    VSSEG_REQUIRE(a8ok(x) && a8ok(dc) && g && dw && x->B == dc->B, "conv3d_wgrad: bad arguments");
    VSSEG_REQUIRE(cout_pad % 16 == 0 && cout_pad >= dc->C, "conv3d_wgrad: cout_pad must be a multiple of 16 >= Cout");
    VSSEG_REQUIRE((g->kx == 1 || g->kx == 3) && (g->ky == 1 || g->ky == 3) && (g->kz == 1 || g->kz == 3),
                  "conv3d_wgrad: kernel size must be 1 or 3 per axis");
    WgradArgs a{*x, *dc, *g, dw, dbias, cout_pad, 1};
    const int taps = g->kx * g->ky * g->kz;
    const int64_t nblk = (int64_t)taps * (x->C / 8) * (cout_pad / 16);
    const vsseg_act8* m = g->transposed ? x : dc;
    const int64_t total = (int64_t)m->B * m->X * m->Y * m->Z;
    int64_t ns = (148 * 8 + nblk - 1) / nblk;                 // fill the chip
    const int64_t max_ns = (total + 128 * 16 - 1) / (128 * 16);  // at least 16 voxels per thread
    if (ns > max_ns) ns = max_ns;
    if (ns < 1) ns = 1;
    a.nslice = (int)ns;
    conv_wgrad_kernel<<<(unsigned)(nblk * ns), 128, 0, (cudaStream_t)stream>>>(a);
    return check_launch("conv3d_wgrad");
}

int vsseg_conv3d_cin1_wgrad(const vsseg_f32view* src, const vsseg_act8* dc, const vsseg_conv_geom* g, float* dw, float* dbias,
                            void* stream) {
    VSSEG_REQUIRE(f32_direct(src) && a8ok(dc) && g && dw, "cin1_wgrad: bad arguments");
    VSSEG_REQUIRE(!g->transposed && g->sx == 1 && g->sy == 1 && g->sz == 1, "cin1_wgrad: stride-1 conv only");
    Cin1WgradArgs a{*src, *dc, *g, dw, dbias};
    const int64_t total = (int64_t)dc->B * dc->X * dc->Y * dc->Z;
    dim3 grid(ew_grid(total, 128 * 16), (unsigned)(dc->C / 8));
    const int taps = g->kx * g->ky * g->kz;
    if (taps == 1) cin1_wgrad_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else if (taps == 9) cin1_wgrad_kernel<9><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else if (taps == 27) cin1_wgrad_kernel<27><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else { set_error("cin1_wgrad: unsupported tap count %d", taps); return VSSEG_EINVAL; }
    return check_launch("cin1_wgrad");
}

int vsseg_conv3d_smallcout_bwd(const vsseg_act8* x, const vsseg_f32view* dy, const vsseg_f32view* y, const vsseg_conv_geom* g,
                               const float* w, int32_t sigmoid, const vsseg_act8* dx, float* dw, float* dbias, void* stream) {
    VSSEG_REQUIRE(a8ok(x) && f32_direct(dy) && g && w, "smallcout_bwd: bad arguments");
    VSSEG_REQUIRE(dy->C == 1 || dy->C == 2, "smallcout_bwd: Cout must be 1 or 2");
    VSSEG_REQUIRE(!sigmoid || f32_direct(y), "smallcout_bwd: sigmoid backward needs the forward output");
    VSSEG_REQUIRE(!g->transposed && g->sx == 1 && g->sy == 1 && g->sz == 1, "smallcout_bwd: stride-1 conv only");
    SmallBwdArgs a{*x, dx ? *dx : *x, *dy, y ? *y : *dy, *g, w, dw, dbias, sigmoid};
    const int taps = g->kx * g->ky * g->kz;
    const int64_t nvox = (int64_t)x->B * x->X * x->Y * x->Z;
    cudaStream_t s = (cudaStream_t)stream;
    if (dx) {
        VSSEG_REQUIRE(a8ok(dx) && same_shape(dx, x), "smallcout_bwd: dx shape mismatch");
        const size_t smem = (size_t)taps * x->C * dy->C * sizeof(float);
        VSSEG_REQUIRE(smem <= 48 * 1024, "smallcout_bwd: weights exceed 48 KB of shared memory");
        const unsigned grid = ew_grid(nvox * (x->C / 8), 128);
        if (dy->C == 1) smallcout_dgrad_kernel<1><<<grid, 128, smem, s>>>(a);
        else smallcout_dgrad_kernel<2><<<grid, 128, smem, s>>>(a);
        if (int e = check_launch("smallcout_dgrad")) return e;
    }
    if (dw) {
        VSSEG_REQUIRE(dbias, "smallcout_bwd: dw without dbias");
        const int kxy = g->kx * g->ky;
        if (x->Z >= 32 && g->kx == g->ky && (kxy == 9 || kxy == 1)) {
            const int nlines = x->B * x->X * x->Y;
            int lpb = (nlines + 592 - 1) / 592;     // ~4 blocks per SM and (z tap, channel group)
            if (lpb < 1) lpb = 1;
            dim3 grid((unsigned)((nlines + lpb - 1) / lpb), (unsigned)g->kz, (unsigned)(x->C / 8));
            if (dy->C == 1 && kxy == 9) smallcout_wgrad_lines_kernel<1, 9><<<grid, 128, 0, s>>>(a, lpb);
            else if (dy->C == 1) smallcout_wgrad_lines_kernel<1, 1><<<grid, 128, 0, s>>>(a, lpb);
            else if (kxy == 9) smallcout_wgrad_lines_kernel<2, 9><<<grid, 128, 0, s>>>(a, lpb);
            else smallcout_wgrad_lines_kernel<2, 1><<<grid, 128, 0, s>>>(a, lpb);
            return check_launch("smallcout_wgrad_lines");
        }
        dim3 grid(ew_grid(nvox, 128 * 32), (unsigned)taps, (unsigned)(x->C / 8));
        if (dy->C == 1) smallcout_wgrad_kernel<1><<<grid, 128, 0, s>>>(a);
        else smallcout_wgrad_kernel<2><<<grid, 128, 0, s>>>(a);
        if (int e = check_launch("smallcout_wgrad")) return e;
    }
    return 0;
}

int vsseg_att_gate_bwd(const vsseg_act8* x, const vsseg_f32view* att, const vsseg_act8* dg, const vsseg_act8* dx,
                       const vsseg_f32view* datt, int32_t accumulate_dx, void* stream) {
    VSSEG_REQUIRE(a8ok(x) && a8ok(dg) && a8ok(dx) && same_shape(x, dg) && same_shape(x, dx) && f32_direct(att) && f32_direct(datt),
                  "att_gate_bwd: bad arguments");
    const int64_t total = (int64_t)x->B * x->X * x->Y * x->Z;
    gate_bwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(*x, *att, *dg, *dx, *datt, accumulate_dx);
    return check_launch("att_gate_bwd");
}

}  // extern "C"

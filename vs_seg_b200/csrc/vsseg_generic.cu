// Generic (CUDA-core fp32) kernels of the vsseg_b200 hot path: bandwidth-bound layers, odd
// shapes (Cin=1, Cout in {1,2}, strided and transposed convs), layout conversion, attention
// gate and the sliding-window finalise.  The tensor-core (tcgen05) path for the FLOP-heavy
// stride-1 convolutions lives in vsseg_tc.cu.
#include <stdarg.h>

#include <stdlib.h>

#include "vsseg_common.cuh"

namespace vsseg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// -------------------------------------------------------------------------------------------
// pack / unpack
// -------------------------------------------------------------------------------------------
__global__ void pack_act8_kernel(vsseg_f32view src, vsseg_act8 dst) {
    const int64_t nvox = (int64_t)dst.X * dst.Y * dst.Z;
    const int CG = dst.C / 8;
    const int64_t total = nvox * CG * dst.B;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i % nvox;
        int cg = (int)((i / nvox) % CG);
        int b = (int)(i / (nvox * CG));
        int z = (int)(v % dst.Z);
        int y = (int)((v / dst.Z) % dst.Y);
        int x = (int)(v / ((int64_t)dst.Z * dst.Y));
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            f[c] = src.ptr[b * src.sb + (cg * 8 + c) * src.sc + x * src.sx + y * src.sy + z * src.sz];
        uint4 h, l;
        pack8(f, h, l);
        __nv_bfloat16* p = (__nv_bfloat16*)dst.hi + act8_off(dst.batch_stride, dst.X, dst.Y, dst.Z, b, cg, x, y, z);
        *reinterpret_cast<uint4*>(p) = h;
        *reinterpret_cast<uint4*>(p + dst.lo_offset) = l;
    }
}

__global__ void unpack_act8_kernel(vsseg_act8 src, vsseg_f32view dst) {
    const int64_t nvox = (int64_t)src.X * src.Y * src.Z;
    const int CG = src.C / 8;
    const int64_t total = nvox * CG * src.B;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i % nvox;
        int cg = (int)((i / nvox) % CG);
        int b = (int)(i / (nvox * CG));
        int z = (int)(v % src.Z);
        int y = (int)((v / src.Z) % src.Y);
        int x = (int)(v / ((int64_t)src.Z * src.Y));
        const __nv_bfloat16* p = (const __nv_bfloat16*)src.hi + act8_off(src.batch_stride, src.X, src.Y, src.Z, b, cg, x, y, z);
        float f[8];
        unpack8(ldg128(p), ldg128(p + src.lo_offset), f);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            dst.ptr[b * dst.sb + (cg * 8 + c) * dst.sc + x * dst.sx + y * dst.sy + z * dst.sz] = f[c];
    }
}

// -------------------------------------------------------------------------------------------
// generic conv: act8 -> act8, fp32 FMA, VOX_T voxels x CO_T output channels per thread
// -------------------------------------------------------------------------------------------
struct ConvArgs {
    vsseg_act8 in, out;
    vsseg_conv_geom g;
    const float* w;
    int cout_pad;
    vsseg_epilogue ep;
    int res_mode;  // 0 none, 1 act8 addend, 2 cin1 affine
    vsseg_act8 res;
    vsseg_f32view rsrc;
    const float* res_w;
    const float* res_b;
};

template <int CO_T, int VOX_T>
__global__ void __launch_bounds__(128) conv_act8_kernel(const ConvArgs a) {
    const int Xo = a.out.X, Yo = a.out.Y, Zo = a.out.Z;
    const int Xi = a.in.X, Yi = a.in.Y, Zi = a.in.Z;
    const int64_t nvox = (int64_t)a.out.B * Xo * Yo * Zo;
    const int co0 = blockIdx.y * CO_T;
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;

    int vb[VOX_T], vx[VOX_T], vy[VOX_T], vz[VOX_T];
    bool vok[VOX_T];
#pragma unroll
    for (int j = 0; j < VOX_T; ++j) {
        int64_t v = (int64_t)blockIdx.x * (128 * VOX_T) + j * 128 + threadIdx.x;
        vok[j] = v < nvox;
        if (!vok[j]) v = 0;
        vz[j] = (int)(v % Zo);
        vy[j] = (int)((v / Zo) % Yo);
        vx[j] = (int)((v / ((int64_t)Zo * Yo)) % Xo);
        vb[j] = (int)(v / ((int64_t)Zo * Yo * Xo));
    }
    float acc[VOX_T][CO_T];
#pragma unroll
    for (int j = 0; j < VOX_T; ++j)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

    const __nv_bfloat16* in_hi = (const __nv_bfloat16*)a.in.hi;
    const int CGi = a.in.C / 8;
    const int64_t cg_stride = (int64_t)Xi * Yi * Zi * 8;

    int tap = 0;
    for (int tx = 0; tx < a.g.kx; ++tx)
        for (int ty = 0; ty < a.g.ky; ++ty)
            for (int tz = 0; tz < a.g.kz; ++tz, ++tap) {
                int64_t off[VOX_T];
                bool ok[VOX_T];
                bool any = false;
#pragma unroll
                for (int j = 0; j < VOX_T; ++j) {
                    int xi, yi, zi;
                    bool o = vok[j];
                    if (!a.g.transposed) {
                        xi = vx[j] * a.g.sx - px + tx;
                        yi = vy[j] * a.g.sy - py + ty;
                        zi = vz[j] * a.g.sz - pz + tz;
                    } else {
                        int t0 = vx[j] + px - tx, t1 = vy[j] + py - ty, t2 = vz[j] + pz - tz;
                        o = o && t0 >= 0 && t1 >= 0 && t2 >= 0 && (t0 % a.g.sx == 0) && (t1 % a.g.sy == 0) &&
                            (t2 % a.g.sz == 0);
                        xi = t0 / a.g.sx;
                        yi = t1 / a.g.sy;
                        zi = t2 / a.g.sz;
                    }
                    o = o && xi >= 0 && xi < Xi && yi >= 0 && yi < Yi && zi >= 0 && zi < Zi;
                    ok[j] = o;
                    any |= o;
                    off[j] = o ? act8_off(a.in.batch_stride, Xi, Yi, Zi, vb[j], 0, xi, yi, zi) : 0;
                }
                if (!__syncthreads_or(any)) continue;
                const float* wt = a.w + ((int64_t)tap * a.in.C) * a.cout_pad + co0;
                for (int cg = 0; cg < CGi; ++cg) {
                    float xin[VOX_T][8];
#pragma unroll
                    for (int j = 0; j < VOX_T; ++j) {
                        if (ok[j]) {
                            const __nv_bfloat16* p = in_hi + off[j] + cg * cg_stride;
                            unpack8(ldg128(p), ldg128(p + a.in.lo_offset), xin[j]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 8; ++c) xin[j][c] = 0.f;
                        }
                    }
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) {
                        float wv[CO_T];
                        const float4* wp = reinterpret_cast<const float4*>(wt + (int64_t)(cg * 8 + ci) * a.cout_pad);
#pragma unroll
                        for (int q = 0; q < CO_T / 4; ++q) {
                            float4 t = __ldg(wp + q);
                            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
                        }
#pragma unroll
                        for (int j = 0; j < VOX_T; ++j)
#pragma unroll
                            for (int c = 0; c < CO_T; ++c) acc[j][c] = fmaf(xin[j][ci], wv[c], acc[j][c]);
                    }
                }
            }

    // epilogue
    __nv_bfloat16* out_hi = (__nv_bfloat16*)a.out.hi;
#pragma unroll
    for (int j = 0; j < VOX_T; ++j) {
        if (!vok[j]) continue;
        float rsrc = 0.f;
        if (a.res_mode == 2)
            rsrc = f32_base(a.rsrc)[vb[j] * a.rsrc.sb + vx[j] * a.rsrc.sx + vy[j] * a.rsrc.sy + vz[j] * a.rsrc.sz];
#pragma unroll
        for (int g8 = 0; g8 < CO_T / 8; ++g8) {
            const int c0 = co0 + g8 * 8;
            if (c0 >= a.out.C) break;
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v = acc[j][g8 * 8 + c] * __ldg(a.ep.scale + c0 + c) + __ldg(a.ep.shift + c0 + c);
                o[c] = apply_act(v, a.ep.act, a.ep.slope);
            }
            if (a.res_mode == 1) {
                const __nv_bfloat16* rp = (const __nv_bfloat16*)a.res.hi +
                                          act8_off(a.res.batch_stride, Xo, Yo, Zo, vb[j], c0 / 8, vx[j], vy[j], vz[j]);
                float r[8];
                unpack8(ldg128(rp), ldg128(rp + a.res.lo_offset), r);
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] += r[c];
            } else if (a.res_mode == 2) {
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] += __ldg(a.res_w + c0 + c) * rsrc + __ldg(a.res_b + c0 + c);
            }
            uint4 h, l;
            pack8(o, h, l);
            __nv_bfloat16* p = out_hi + act8_off(a.out.batch_stride, Xo, Yo, Zo, vb[j], c0 / 8, vx[j], vy[j], vz[j]);
            *reinterpret_cast<uint4*>(p) = h;
            *reinterpret_cast<uint4*>(p + a.out.lo_offset) = l;
        }
    }
}

// -------------------------------------------------------------------------------------------
// first conv: Cin = 1 fp32 strided source -> act8 (COUT = 16)
// -------------------------------------------------------------------------------------------
struct Cin1Args {
    vsseg_f32view in;
    vsseg_act8 out;
    vsseg_conv_geom g;
    const float* w;  // [taps][Cout]
    vsseg_epilogue ep;
    WinTab win;      // `in` is a window set: batch item b is read at win.off[b]
};

template <int COUT>
__global__ void __launch_bounds__(128) conv_cin1_kernel(const Cin1Args a) {
    // block = one (b, x, y) line segment of 128 z; all per-voxel index arithmetic is 32-bit and block-uniform
    // (the first version spent ~1000 instructions per voxel, most of them 64-bit div/mod: instruction-bound)
    __shared__ float4 ws4[27 * COUT / 4];
    float* ws = reinterpret_cast<float*>(ws4);
    const int taps = a.g.kx * a.g.ky * a.g.kz;
    for (int i = threadIdx.x; i < taps * COUT; i += blockDim.x) ws[i] = a.w[i];
    __syncthreads();
    const int X = a.out.X, Y = a.out.Y, Z = a.out.Z;
    const int nzt = (Z + 127) / 128;
    int t = blockIdx.x;
    const int zt = t % nzt; t /= nzt;
    const int y = t % Y; t /= Y;
    const int x = t % X;
    const int b = t / X;
    const int z = zt * 128 + threadIdx.x;
    if (z >= Z) return;
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    const float* src = f32_base(a.in) + win_off(a.in, a.win, b);
    int tap = 0;
    for (int tx = 0; tx < a.g.kx; ++tx)
        for (int ty = 0; ty < a.g.ky; ++ty) {
            const int xi = x - px + tx, yi = y - py + ty;
            const bool line_in = xi >= 0 && xi < X && yi >= 0 && yi < Y;      // block-uniform
            const float* lp = src + (int64_t)xi * a.in.sx + (int64_t)yi * a.in.sy;
            for (int tz = 0; tz < a.g.kz; ++tz, ++tap) {
                const int zi = z - pz + tz;
                const float sv = (line_in && zi >= 0 && zi < Z) ? __ldg(lp + (int64_t)zi * a.in.sz) : 0.f;
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    const float4 w = ws4[tap * (COUT / 4) + q];
                    acc[4 * q] = fmaf(sv, w.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(sv, w.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(sv, w.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(sv, w.w, acc[4 * q + 3]);
                }
            }
        }
    __nv_bfloat16* out_hi = (__nv_bfloat16*)a.out.hi;
#pragma unroll
    for (int g8 = 0; g8 < COUT / 8; ++g8) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int cc = g8 * 8 + c;
            o[c] = apply_act(acc[cc] * __ldg(a.ep.scale + cc) + __ldg(a.ep.shift + cc), a.ep.act, a.ep.slope);
        }
        uint4 h, l;
        pack8(o, h, l);
        __nv_bfloat16* p = out_hi + act8_off(a.out.batch_stride, X, Y, Z, b, g8, x, y, z);
        *reinterpret_cast<uint4*>(p) = h;
        *reinterpret_cast<uint4*>(p + a.out.lo_offset) = l;
    }
}

// Specialisation for the network's actual first conv: k = (3,3,1), PReLU.  A thread owns YB = 4 consecutive
// y outputs at one (x, z): 18 coalesced input loads (3 x planes x 6 y lines) feed 4 x 16 accumulators, every
// 128-bit shared-memory weight read is used for 16 FMAs.  ~290 instructions per voxel (the generic kernel
// above: ~1100, instruction-bound at 90 us for a 128^3 patch).
template <int CIN1_YB>
__global__ void __launch_bounds__(128) conv_cin1_k331_kernel(const Cin1Args a) {
    __shared__ float4 ws4[9 * 4];      // [tap][16]
    __shared__ float4 ep4[2 * 4];      // scale[16], shift[16]
    {
        float* ws = reinterpret_cast<float*>(ws4);
        float* es = reinterpret_cast<float*>(ep4);
        for (int i = threadIdx.x; i < 144; i += 128) ws[i] = a.w[i];
        if (threadIdx.x < 16) es[threadIdx.x] = a.ep.scale[threadIdx.x];
        else if (threadIdx.x < 32) es[threadIdx.x] = a.ep.shift[threadIdx.x - 16];
    }
    __syncthreads();
    const int X = a.out.X, Y = a.out.Y, Z = a.out.Z;
    const int nzt = (Z + 127) / 128, nyt = (Y + CIN1_YB - 1) / CIN1_YB;
    int t = blockIdx.x;
    const int zt = t % nzt; t /= nzt;
    const int y0 = (t % nyt) * CIN1_YB; t /= nyt;
    const int x = t % X;
    const int b = t / X;
    const int z = zt * 128 + threadIdx.x;
    if (z >= Z) return;
    float2 acc2[CIN1_YB][8];   // channel pairs: packed fp32 FMAs (fma.rn.f32x2), the kernel is instruction-bound
#pragma unroll
    for (int v = 0; v < CIN1_YB; ++v)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc2[v][c] = make_float2(0.f, 0.f);
    const float* src = f32_base(a.in) + win_off(a.in, a.win, b) + (int64_t)z * a.in.sz;
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
        const int xi = x - 1 + tx;
        const bool xin = xi >= 0 && xi < X;     // block-uniform
        float sv[CIN1_YB + 2];
#pragma unroll
        for (int j = 0; j < CIN1_YB + 2; ++j) {
            const int yi = y0 - 1 + j;
            sv[j] = (xin && yi >= 0 && yi < Y) ? __ldg(src + (int64_t)xi * a.in.sx + (int64_t)yi * a.in.sy) : 0.f;
        }
#pragma unroll
        for (int ty = 0; ty < 3; ++ty)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = ws4[(tx * 3 + ty) * 4 + q];
#pragma unroll
                for (int v = 0; v < CIN1_YB; ++v) {
                    const float2 s2 = make_float2(sv[v + ty], sv[v + ty]);
                    acc2[v][2 * q] = __ffma2_rn(s2, make_float2(w.x, w.y), acc2[v][2 * q]);
                    acc2[v][2 * q + 1] = __ffma2_rn(s2, make_float2(w.z, w.w), acc2[v][2 * q + 1]);
                }
            }
    }
    const float sm1 = a.ep.slope - 1.0f;
    __nv_bfloat16* out_hi = (__nv_bfloat16*)a.out.hi;
    const int64_t cgs = (int64_t)X * Y * Z * 8;
#pragma unroll
    for (int v = 0; v < CIN1_YB; ++v) {
        const int y = y0 + v;
        if (y >= Y) break;
        __nv_bfloat16* p = out_hi + (int64_t)b * a.out.batch_stride + (((int64_t)x * Y + y) * Z + z) * 8;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
            float o[8], sc[8], sh[8];
            *reinterpret_cast<float4*>(sc) = ep4[g8 * 2];
            *reinterpret_cast<float4*>(sc + 4) = ep4[g8 * 2 + 1];
            *reinterpret_cast<float4*>(sh) = ep4[4 + g8 * 2];
            *reinterpret_cast<float4*>(sh + 4) = ep4[4 + g8 * 2 + 1];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float av = (c & 1) ? acc2[v][g8 * 4 + c / 2].y : acc2[v][g8 * 4 + c / 2].x;
                const float f = fmaf(av, sc[c], sh[c]);
                o[c] = fmaf(fminf(f, 0.f), sm1, f);   // PReLU / ReLU / identity
            }
            uint4 h, l;
            pack8(o, h, l);
            *reinterpret_cast<uint4*>(p + g8 * cgs) = h;
            *reinterpret_cast<uint4*>(p + g8 * cgs + a.out.lo_offset) = l;
        }
    }
}

// -------------------------------------------------------------------------------------------
// small-Cout conv: act8 -> planar fp32 (Cout 1 or 2), optional sliding-window blend
// -------------------------------------------------------------------------------------------
struct SmallCoutArgs {
    vsseg_act8 in;
    vsseg_f32view out;
    vsseg_conv_geom g;
    const float* w;  // [taps][Cin][COUT]
    const float* bias;
    int act;
    float slope;
    const float* sw_weight;
};

template <int COUT>
__global__ void __launch_bounds__(128) conv_smallcout_kernel(const SmallCoutArgs a) {
    extern __shared__ float ws[];
    const int taps = a.g.kx * a.g.ky * a.g.kz;
    const int nw = taps * a.in.C * COUT;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) ws[i] = a.w[i];
    __syncthreads();
    const int X = a.in.X, Y = a.in.Y, Z = a.in.Z;
    const int64_t nvox = (int64_t)a.in.B * X * Y * Z;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)((v / ((int64_t)Z * Y)) % X);
    const int b = (int)(v / ((int64_t)Z * Y * X));
    const int px = (a.g.kx - 1) / 2, py = (a.g.ky - 1) / 2, pz = (a.g.kz - 1) / 2;
    const int CG = a.in.C / 8;
    const int64_t cg_stride = (int64_t)X * Y * Z * 8;
    const __nv_bfloat16* in_hi = (const __nv_bfloat16*)a.in.hi;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    int tap = 0;
    for (int tx = 0; tx < a.g.kx; ++tx)
        for (int ty = 0; ty < a.g.ky; ++ty)
            for (int tz = 0; tz < a.g.kz; ++tz, ++tap) {
                int xi = x - px + tx, yi = y - py + ty, zi = z - pz + tz;
                if (xi < 0 || xi >= X || yi < 0 || yi >= Y || zi < 0 || zi >= Z) continue;
                const __nv_bfloat16* p = in_hi + act8_off(a.in.batch_stride, X, Y, Z, b, 0, xi, yi, zi);
                const float* wt = ws + tap * a.in.C * COUT;
                for (int cg = 0; cg < CG; ++cg) {
                    float f[8];
                    unpack8(ldg128(p + cg * cg_stride), ldg128(p + cg * cg_stride + a.in.lo_offset), f);
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
                        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(f[ci], wt[(cg * 8 + ci) * COUT + c], acc[c]);
                }
            }
    const float sw = a.sw_weight ? __ldg(a.sw_weight + ((int64_t)x * Y + y) * Z + z) : 0.f;
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        float r = apply_act(acc[c] + __ldg(a.bias + c), a.act, a.slope);
        float* o = f32_base(a.out) + b * a.out.sb + c * a.out.sc + x * a.out.sx + y * a.out.sy + z * a.out.sz;
        if (a.sw_weight)
            *o += sw * r;
        else
            *o = r;
    }
}

// -------------------------------------------------------------------------------------------
// attention gate: out = x * (1 + att)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) att_gate_kernel(vsseg_act8 x, vsseg_f32view att, vsseg_act8 out) {
    const int64_t nvox = (int64_t)x.X * x.Y * x.Z;
    const int CG = x.C / 8;
    const int64_t total = nvox * CG * x.B;
    const __nv_bfloat16* xh = (const __nv_bfloat16*)x.hi;
    __nv_bfloat16* oh = (__nv_bfloat16*)out.hi;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i % nvox;
        int cg = (int)((i / nvox) % CG);
        int b = (int)(i / (nvox * CG));
        int z = (int)(v % x.Z);
        int y = (int)((v / x.Z) % x.Y);
        int xx = (int)(v / ((int64_t)x.Z * x.Y));
        float g = 1.0f + __ldg(f32_base(att) + b * att.sb + xx * att.sx + y * att.sy + z * att.sz);
        const __nv_bfloat16* p = xh + (int64_t)b * x.batch_stride + ((int64_t)cg * nvox + v) * 8;
        float f[8];
        unpack8(ldg128(p), ldg128(p + x.lo_offset), f);
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] *= g;
        uint4 h, l;
        pack8(f, h, l);
        __nv_bfloat16* q = oh + (int64_t)b * out.batch_stride + ((int64_t)cg * nvox + v) * 8;
        *reinterpret_cast<uint4*>(q) = h;
        *reinterpret_cast<uint4*>(q + out.lo_offset) = l;
    }
}

// -------------------------------------------------------------------------------------------
// sliding-window finalise
// -------------------------------------------------------------------------------------------
// 4 voxels per thread: 128-bit loads of every accumulator channel and of the weight-sum map, 128-bit store of the
// probabilities, 32-bit store of the mask (C <= 8 channels stay in registers)
template <typename LabelT, bool VEC>
__global__ void __launch_bounds__(256) sw_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ cnt,
                                                          float* __restrict__ out, int C, int64_t n,
                                                          uint8_t* __restrict__ mask, const LabelT* __restrict__ label,
                                                          double* sums) {
    constexpr int V = VEC ? 4 : 1;
    float s_i = 0.f, s_l = 0.f, s_p = 0.f;
    const int64_t nv = n / V;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        float inv[V], best[V];
        int arg[V];
        if (cnt) {
            if constexpr (VEC) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cnt) + i);
                inv[0] = c4.x; inv[1] = c4.y; inv[2] = c4.z; inv[3] = c4.w;
            } else {
                inv[0] = cnt[i];
            }
        }
        for (int c = 0; c < C; ++c) {
            float p[V];
            if constexpr (VEC) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(acc + c * n) + i);
                p[0] = a4.x; p[1] = a4.y; p[2] = a4.z; p[3] = a4.w;
            } else {
                p[0] = acc[c * n + i];
            }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                if (cnt) p[k] = p[k] / inv[k];   // true division, as MONAI's out / count
                if (c == 0 || p[k] > best[k]) {  // first maximum wins, as torch.argmax
                    best[k] = p[k];
                    arg[k] = c;
                }
            }
            if (out) {
                if constexpr (VEC) reinterpret_cast<float4*>(out + c * n)[i] = make_float4(p[0], p[1], p[2], p[3]);
                else out[c * n + i] = p[0];
            }
        }
        if (mask) {
            if constexpr (VEC) reinterpret_cast<uchar4*>(mask)[i] = make_uchar4((uint8_t)arg[0], (uint8_t)arg[1], (uint8_t)arg[2], (uint8_t)arg[3]);
            else mask[i] = (uint8_t)arg[0];
        }
        if (label) {
            float lb[V];
            if constexpr (VEC) {
                if constexpr (sizeof(LabelT) == 4) {
                    const float4 l4 = __ldg(reinterpret_cast<const float4*>(label) + i);
                    lb[0] = l4.x; lb[1] = l4.y; lb[2] = l4.z; lb[3] = l4.w;
                } else {
                    const uchar4 l4 = __ldg(reinterpret_cast<const uchar4*>(label) + i);
                    lb[0] = l4.x; lb[1] = l4.y; lb[2] = l4.z; lb[3] = l4.w;
                }
            } else {
                lb[0] = (float)label[i];
            }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float pr = arg[k] == 1 ? 1.f : 0.f;
                s_i += pr * lb[k];
                s_l += lb[k];
                s_p += pr;
            }
        }
    }
    if (label && sums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_i += __shfl_xor_sync(0xffffffffu, s_i, o);
            s_l += __shfl_xor_sync(0xffffffffu, s_l, o);
            s_p += __shfl_xor_sync(0xffffffffu, s_p, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(sums + 0, (double)s_i);
            atomicAdd(sums + 1, (double)s_l);
            atomicAdd(sums + 2, (double)s_p);
        }
    }
}

static int grid_for(int64_t total, int block, int sm_mult = 8) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (total + block - 1) / block;
    int64_t cap = (int64_t)sms * sm_mult;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

static bool act8_ok(const vsseg_act8* t) {
    return t && t->hi && t->C > 0 && t->C % 8 == 0 && t->B > 0 && t->X > 0 && t->Y > 0 && t->Z > 0 &&
           ((uintptr_t)t->hi % 16 == 0) && (t->lo_offset % 8 == 0) && (t->batch_stride % 8 == 0);
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_abi_version(void) { return VSSEG_ABI_VERSION; }

const char* vsseg_last_error(void) { return g_err; }

int vsseg_device_sm_count(int device, int* sm_count_host) {
    VSSEG_REQUIRE(sm_count_host, "sm_count_host is NULL");
    cudaError_t e = cudaDeviceGetAttribute(sm_count_host, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int vsseg_pack_act8(const vsseg_f32view* src, const vsseg_act8* dst, void* stream) {
    VSSEG_REQUIRE(f32_direct(src) && act8_ok(dst), "pack_act8: bad tensor descriptor");
    VSSEG_REQUIRE(src->C == dst->C && src->B == dst->B && src->X == dst->X && src->Y == dst->Y && src->Z == dst->Z,
                  "pack_act8: shape mismatch");
    int64_t total = (int64_t)dst->B * (dst->C / 8) * dst->X * dst->Y * dst->Z;
    pack_act8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
    return check_launch("pack_act8");
}

int vsseg_unpack_act8(const vsseg_act8* src, const vsseg_f32view* dst, void* stream) {
    VSSEG_REQUIRE(f32_direct(dst) && act8_ok(src), "unpack_act8: bad tensor descriptor");
    VSSEG_REQUIRE(src->C == dst->C && src->B == dst->B && src->X == dst->X && src->Y == dst->Y && src->Z == dst->Z,
                  "unpack_act8: shape mismatch");
    int64_t total = (int64_t)src->B * (src->C / 8) * src->X * src->Y * src->Z;
    unpack_act8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
    return check_launch("unpack_act8");
}

static int check_geom(const vsseg_conv_geom* g, const char* who) {
    VSSEG_REQUIRE(g, "%s: geometry is NULL", who);
    VSSEG_REQUIRE((g->kx == 1 || g->kx == 3) && (g->ky == 1 || g->ky == 3) && (g->kz == 1 || g->kz == 3),
                  "%s: kernel size must be 1 or 3 per axis", who);
    VSSEG_REQUIRE(g->sx >= 1 && g->sx <= 2 && g->sy >= 1 && g->sy <= 2 && g->sz >= 1 && g->sz <= 2,
                  "%s: stride must be 1 or 2 per axis", who);
    return 0;
}

int vsseg_conv3d_act8(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, const float* w,
                      int32_t cout_pad, const vsseg_epilogue* ep, const vsseg_act8* res_act8,
                      const vsseg_f32view* res_src, const float* res_w, const float* res_b, void* stream) {
    VSSEG_REQUIRE(act8_ok(in) && act8_ok(out), "conv3d_act8: bad tensor descriptor");
    if (int e = check_geom(g, "conv3d_act8")) return e;
    VSSEG_REQUIRE(w && ep && ep->scale && ep->shift, "conv3d_act8: NULL weights/epilogue");
    VSSEG_REQUIRE(cout_pad % 16 == 0 && cout_pad >= out->C, "conv3d_act8: cout_pad must be a multiple of 16 >= Cout");
    VSSEG_REQUIRE(in->B == out->B, "conv3d_act8: batch mismatch");
    if (g->transposed) {
        VSSEG_REQUIRE(out->X == in->X * g->sx && out->Y == in->Y * g->sy && out->Z == in->Z * g->sz,
                      "conv3d_act8: transposed output must be input*stride");
    } else {
        VSSEG_REQUIRE(out->X == (in->X + g->sx - 1) / g->sx && out->Y == (in->Y + g->sy - 1) / g->sy &&
                          out->Z == (in->Z + g->sz - 1) / g->sz,
                      "conv3d_act8: output shape does not match ceil(input/stride)");
    }
    VSSEG_REQUIRE(!(res_act8 && res_src), "conv3d_act8: at most one residual source");
    ConvArgs a{};
    a.in = *in;
    a.out = *out;
    a.g = *g;
    a.w = w;
    a.cout_pad = cout_pad;
    a.ep = *ep;
    a.res_mode = 0;
    if (res_act8) {
        VSSEG_REQUIRE(act8_ok(res_act8) && res_act8->C == out->C && res_act8->X == out->X && res_act8->Y == out->Y &&
                          res_act8->Z == out->Z && res_act8->B == out->B,
                      "conv3d_act8: residual shape mismatch");
        a.res_mode = 1;
        a.res = *res_act8;
    } else if (res_src) {
        VSSEG_REQUIRE(f32_ok(res_src) && res_w && res_b, "conv3d_act8: NULL cin1 residual");
        a.res_mode = 2;
        a.rsrc = *res_src;
        a.res_w = res_w;
        a.res_b = res_b;
    }
    const int64_t nvox = (int64_t)out->B * out->X * out->Y * out->Z;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid;
    grid.y = (out->C + 15) / 16;
    if (nvox >= 148 * 128 * 4 * 2) {
        grid.x = (unsigned)((nvox + 511) / 512);
        conv_act8_kernel<16, 4><<<grid, 128, 0, s>>>(a);
    } else {
        grid.x = (unsigned)((nvox + 127) / 128);
        conv_act8_kernel<16, 1><<<grid, 128, 0, s>>>(a);
    }
    return check_launch("conv3d_act8");
}

int vsseg_conv3d_cin1(const vsseg_f32view* in, const vsseg_act8* out, const vsseg_conv_geom* g, const float* w,
                      const vsseg_epilogue* ep, void* stream) {
    VSSEG_REQUIRE(f32_set_ok(in) && act8_ok(out), "conv3d_cin1: bad tensor descriptor");
    if (int e = check_geom(g, "conv3d_cin1")) return e;
    VSSEG_REQUIRE(g->sx == 1 && g->sy == 1 && g->sz == 1 && !g->transposed, "conv3d_cin1: stride-1 conv only");
    VSSEG_REQUIRE(out->C == 16, "conv3d_cin1: Cout must be 16 (got %d)", out->C);
    const bool set = in->n_windows > 1;   // every window of a sliding-window group in one launch
    VSSEG_REQUIRE(in->X == out->X && in->Y == out->Y && in->Z == out->Z && (set ? in->n_windows : in->B) == out->B,
                  "conv3d_cin1: shape mismatch");
    VSSEG_REQUIRE(w && ep && ep->scale && ep->shift, "conv3d_cin1: NULL weights/epilogue");
    Cin1Args a{*in, *out, *g, w, *ep, {}};
    VSSEG_REQUIRE(win_tab(in, out->B, &a.win), "conv3d_cin1: inconsistent window set (%d records for batch %d)",
                  in->n_windows, out->B);
    const int64_t nvox = (int64_t)out->B * out->X * out->Y * out->Z;
    const unsigned nblk = (unsigned)((int64_t)out->B * out->X * out->Y * ((out->Z + 127) / 128));
    if (g->kx == 3 && g->ky == 3 && g->kz == 1 && ep->act != 1) {
        static const int yb = getenv("VSSEG_CIN1_YB") ? atoi(getenv("VSSEG_CIN1_YB")) : 4;   // y outputs per thread (tuning knob)
        const int YB = yb == 2 ? 2 : 4;
        const unsigned nb4 = (unsigned)((int64_t)out->B * out->X * ((out->Y + YB - 1) / YB) * ((out->Z + 127) / 128));
        if (YB == 2) conv_cin1_k331_kernel<2><<<nb4, 128, 0, (cudaStream_t)stream>>>(a);
        else conv_cin1_k331_kernel<4><<<nb4, 128, 0, (cudaStream_t)stream>>>(a);
    } else {
        conv_cin1_kernel<16><<<nblk, 128, 0, (cudaStream_t)stream>>>(a);
    }
    return check_launch("conv3d_cin1");
}

int vsseg_conv3d_smallcout(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g, const float* w,
                           const float* bias, int32_t act, float slope, const float* sw_weight, void* stream) {
    VSSEG_REQUIRE(act8_ok(in) && f32_ok(out), "conv3d_smallcout: bad tensor descriptor");
    if (int e = check_geom(g, "conv3d_smallcout")) return e;
    VSSEG_REQUIRE(g->sx == 1 && g->sy == 1 && g->sz == 1 && !g->transposed, "conv3d_smallcout: stride-1 conv only");
    VSSEG_REQUIRE(out->C == 1 || out->C == 2, "conv3d_smallcout: Cout must be 1 or 2");
    VSSEG_REQUIRE(in->X == out->X && in->Y == out->Y && in->Z == out->Z && in->B == out->B,
                  "conv3d_smallcout: shape mismatch");
    VSSEG_REQUIRE(w && bias, "conv3d_smallcout: NULL weights");
    SmallCoutArgs a{*in, *out, *g, w, bias, act, slope, sw_weight};
    const int64_t nvox = (int64_t)in->B * in->X * in->Y * in->Z;
    const int taps = g->kx * g->ky * g->kz;
    const size_t smem = (size_t)taps * in->C * out->C * sizeof(float);
    VSSEG_REQUIRE(smem <= 48 * 1024, "conv3d_smallcout: weights (%zu B) exceed 48 KB of shared memory", smem);
    const unsigned grid = (unsigned)((nvox + 127) / 128);
    if (out->C == 1)
        conv_smallcout_kernel<1><<<grid, 128, smem, (cudaStream_t)stream>>>(a);
    else
        conv_smallcout_kernel<2><<<grid, 128, smem, (cudaStream_t)stream>>>(a);
    return check_launch("conv3d_smallcout");
}

int vsseg_att_gate(const vsseg_act8* x, const vsseg_f32view* att, const vsseg_act8* out, void* stream) {
    VSSEG_REQUIRE(act8_ok(x) && act8_ok(out) && f32_ok(att), "att_gate: bad tensor descriptor");
    VSSEG_REQUIRE(x->C == out->C && x->B == out->B && x->X == out->X && x->Y == out->Y && x->Z == out->Z &&
                      att->B == x->B && att->X == x->X && att->Y == x->Y && att->Z == x->Z,
                  "att_gate: shape mismatch");
    int64_t total = (int64_t)x->B * (x->C / 8) * x->X * x->Y * x->Z;
    att_gate_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(*x, *att, *out);
    return check_launch("att_gate");
}

int vsseg_sw_finalize(const float* acc, const float* cnt, float* out, int32_t C, int64_t n, uint8_t* mask,
                      const void* label, int32_t label_u8, double* sums, void* stream) {
    VSSEG_REQUIRE(acc && C >= 1 && n > 0, "sw_finalize: bad arguments");
    VSSEG_REQUIRE(!label || sums, "sw_finalize: label given without sums");
    auto al = [](const void* p, size_t a) { return ((uintptr_t)p % a) == 0; };
    const bool vec = n % 4 == 0 && al(acc, 16) && al(cnt, 16) && al(out, 16) && al(mask, 4) && al(label, label_u8 ? 4 : 16);
    const int grid = grid_for(vec ? n / 4 : n, 256, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (label_u8) {
        const uint8_t* lb = (const uint8_t*)label;
        if (vec) sw_finalize_kernel<uint8_t, true><<<grid, 256, 0, st>>>(acc, cnt, out, C, n, mask, lb, sums);
        else sw_finalize_kernel<uint8_t, false><<<grid, 256, 0, st>>>(acc, cnt, out, C, n, mask, lb, sums);
    } else {
        const float* lb = (const float*)label;
        if (vec) sw_finalize_kernel<float, true><<<grid, 256, 0, st>>>(acc, cnt, out, C, n, mask, lb, sums);
        else sw_finalize_kernel<float, false><<<grid, 256, 0, st>>>(acc, cnt, out, C, n, mask, lb, sums);
    }
    return check_launch("sw_finalize");
}

}  // extern "C"

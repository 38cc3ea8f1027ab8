// Top of the decoder as ONE bandwidth-bound launch: AttentionBlock2 gate (reference attentionblock.py:44-47) fused
// into the last ResidualUnit (conv_only (3,3,1) conv + 1x1x1 shortcut folded into its centre tap,
// unet2d5_spvPA.py:186-190) and the sliding-window blend (MONAI sliding_window_inference step 6, call site
// VSparams.py:568-574).  The gated 2c-channel tensor x*(1+att) is consumed by nothing else, so it is never
// written: that removes the largest pure round trip of the network (8 B x 32 channels per voxel).
//
// Formulation.  conv(g*x)[xo,yo] = sum_{tx,ty} W[tx][ty] . (g*x)[xo+tx-1, yo+ty-1].  Every INPUT voxel is read
// exactly once, projected onto its 9 x Cout contributions P[tx][ty] = g * sum_c W[tx][ty][c] x[c] (fp32 FMAs,
// weights are kernel parameters = constant-bank operands), and the stencil is closed in two cheap steps:
//   y: the three lines of a plane exchange P[.][0] / P[.][2] through shared memory (one barrier per plane,
//      double buffered)  ->  Q[tx](yo) = P[tx][0](yo-1) + P[tx][1](yo) + P[tx][2](yo+1)
//   x: the CTA marches along x and keeps the two open output rows in registers:
//      row(xi-1) = r_prev + Q[2] (complete -> emitted), r_prev' = r_cur + Q[1], r_cur' = Q[0].
// A CTA is (up to 66 y lines) x (8 z) threads, one voxel per thread and plane: a warp reads 4 x 128 B per
// channel group and plane.  Algorithmic bytes per voxel: 4*Cin (x) + 4 (att) + 8*Cout (+4 weight map) blend.
#include <stdlib.h>
#include <string.h>

#include "vsseg_common.cuh"

namespace vsseg {

constexpr int GL_TZ = 8;
#ifndef VSSEG_GL_MAXL
#define VSSEG_GL_MAXL 66
#endif
#ifndef VSSEG_GL_CTAS
#define VSSEG_GL_CTAS 2
#endif
constexpr int GL_MAXL = VSSEG_GL_MAXL;   // y lines per CTA (incl. the two halo lines of an interior tile)
constexpr int GL_CTAS = VSSEG_GL_CTAS;   // CTAs per SM the register budget is cut for
constexpr int GL_MAXW = 16;
constexpr int GL_CIN = 32;

struct GateLogitsArgs {
    vsseg_act8 x;
    vsseg_f32view att, out;
    long long out_off[GL_MAXW];   // n_outs > 1: address (or byte offset from the cell when out.indirect) of entry b's view
    const float* sw_weight;
    int has_att, n_outs, atomic;
    int TY, ny, L;                // y lines emitted per tile, y tiles, lines staged per tile (TY + 2 halo lines when ny > 1)
    int XT, nxs, nz;              // x rows per segment, x segments, z tiles
    float bias[2];
    alignas(16) float w[9 * GL_CIN * 2];      // [tap = tx*3+ty][cin][COUT]
};

template <int COUT>
__global__ void __launch_bounds__(GL_MAXL* GL_TZ, GL_CTAS) gate_logits_kernel(const __grid_constant__ GateLogitsArgs a) {
    extern __shared__ float ex[];   // [2 buffers][ty = 0 | 2][tx][COUT][threads]
    constexpr int NCG = GL_CIN / 8;
    const int nthr = blockDim.x, tid = threadIdx.x;
    const int zl = tid % GL_TZ, line = tid / GL_TZ;
    int t = blockIdx.x;
    const int tz = t % a.nz; t /= a.nz;
    const int tyi = t % a.ny; t /= a.ny;
    const int xs = t % a.nxs; t /= a.nxs;
    const int b = t;
    const int X = a.x.X, Y = a.x.Y, Z = a.x.Z;
    const int z = tz * GL_TZ + zl;
    const int y0 = tyi * a.TY;
    const int y = y0 - (a.ny > 1 ? 1 : 0) + line;   // the input line this thread projects (and emits when inside the tile)
    const bool in_y = y >= 0 && y < Y;
    const bool emit_y = in_y && y >= y0 && y < y0 + a.TY;
    const int x_lo = xs * a.XT, x_hi = min(x_lo + a.XT, X);
    const int64_t cgs = (int64_t)X * Y * Z * 8;
    const __nv_bfloat16* xp = (const __nv_bfloat16*)a.x.hi + (int64_t)b * a.x.batch_stride + ((int64_t)y * Z + z) * 8;
    const int64_t xstep = (int64_t)Y * Z * 8, lo = a.x.lo_offset;
    const float* attp = a.has_att ? f32_base(a.att) + b * a.att.sb + y * a.att.sy + z * a.att.sz : nullptr;
    float* outp;
    if (a.n_outs > 1) {
        const long long base = a.out.indirect ? __ldg(reinterpret_cast<const long long*>(a.out.indirect)) : 0ll;
        outp = reinterpret_cast<float*>(base + a.out_off[b]);
    } else {
        outp = f32_base(a.out) + b * a.out.sb;
    }
    outp += y * a.out.sy + z * a.out.sz;
    const float* swp = a.sw_weight ? a.sw_weight + (int64_t)y * Z + z : nullptr;
    const int64_t sw_xstep = (int64_t)Y * Z;

    float r_prev[COUT], r_cur[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) r_prev[o] = r_cur[o] = 0.f;

    auto emit = [&](int xo, const float (&v)[COUT]) {
        if (!emit_y || xo < x_lo || xo >= x_hi) return;
        const float sw = swp ? __ldg(swp + xo * sw_xstep) : 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            float* q = outp + o * a.out.sc + xo * a.out.sx;
            const float r = v[o] + a.bias[o];
            if (!swp) *q = r;
            else if (a.atomic) atomicAdd(q, sw * r);   // overlapping windows of one launch (red.global.add.f32)
            else *q += sw * r;
        }
    };

    const int xi0 = max(x_lo - 1, 0), xi1 = min(x_hi, X - 1);
    int it = 0;
    for (int xi = xi0; xi <= xi1; ++xi, ++it) {
        float P[9][COUT];
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int o = 0; o < COUT; ++o) P[k][o] = 0.f;
        if (in_y) {
            const __nv_bfloat16* p = xp + xi * xstep;
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) {
                float f[8];
                unpack8(ldg128(p + cg * cgs), ldg128(p + cg * cgs + lo), f);
                if constexpr (COUT == 2) {
                    // packed fp32 FMAs (fma.rn.f32x2, sm_100): both output channels of a tap in one instruction -
                    // the 576 FMAs per voxel are what bounds this kernel, not its 150 B of traffic
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float2 ff = make_float2(f[c], f[c]);
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            const float2 w2 = *reinterpret_cast<const float2*>(&a.w[((k * NCG + cg) * 8 + c) * 2]);
                            const float2 r = __ffma2_rn(ff, w2, make_float2(P[k][0], P[k][1]));
                            P[k][0] = r.x;
                            P[k][1] = r.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int k = 0; k < 9; ++k)
#pragma unroll
                            for (int o = 0; o < COUT; ++o)
                                P[k][o] = fmaf(f[c], a.w[((k * NCG + cg) * 8 + c) * COUT + o], P[k][o]);
                }
            }
            if (attp) {
                const float g = 1.0f + __ldg(attp + xi * a.att.sx);
#pragma unroll
                for (int k = 0; k < 9; ++k)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) P[k][o] *= g;
            }
        }
        float* buf = ex + (it & 1) * (6 * COUT * nthr);
#pragma unroll
        for (int tx = 0; tx < 3; ++tx)
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                buf[((0 * 3 + tx) * COUT + o) * nthr + tid] = P[tx * 3 + 0][o];
                buf[((1 * 3 + tx) * COUT + o) * nthr + tid] = P[tx * 3 + 2][o];
            }
        __syncthreads();
        float Q[3][COUT];
#pragma unroll
        for (int tx = 0; tx < 3; ++tx)
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float q = P[tx * 3 + 1][o];
                if (line > 0) q += buf[((0 * 3 + tx) * COUT + o) * nthr + tid - GL_TZ];          // ty = 0 tap of line y-1
                if (line < a.L - 1) q += buf[((1 * 3 + tx) * COUT + o) * nthr + tid + GL_TZ];    // ty = 2 tap of line y+1
                Q[tx][o] = q;
            }
        float done[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            done[o] = r_prev[o] + Q[2][o];
            r_prev[o] = r_cur[o] + Q[1][o];
            r_cur[o] = Q[0][o];
        }
        emit(xi - 1, done);
    }
    if (x_hi == X) emit(X - 1, r_prev);   // the last row of the volume has no plane behind it
}

}  // namespace vsseg

using namespace vsseg;

extern "C" int vsseg_conv3d_gate_logits(const vsseg_act8* x, const vsseg_f32view* att, const float* w_host,
                                        const float* bias_host, int32_t cout, const vsseg_f32view* outs, int32_t n_outs,
                                        const float* sw_weight, int32_t atomic_blend, void* stream) {
    VSSEG_REQUIRE(x && x->hi && x->C == GL_CIN && x->B >= 1 && x->X >= 1 && x->Y >= 1 && x->Z >= GL_TZ && x->Z % GL_TZ == 0,
                  "conv3d_gate_logits: x must be act8 with %d channels and Z %% %d == 0", GL_CIN, GL_TZ);
    VSSEG_REQUIRE(w_host && bias_host && (cout == 1 || cout == 2), "conv3d_gate_logits: Cout must be 1 or 2");
    VSSEG_REQUIRE(outs && (n_outs == 1 || (n_outs == x->B && n_outs <= GL_MAXW)),
                  "conv3d_gate_logits: n_outs must be 1 or B (<= %d)", GL_MAXW);
    VSSEG_REQUIRE(!att || (f32_ok(att) && att->B == x->B && att->X == x->X && att->Y == x->Y && att->Z == x->Z),
                  "conv3d_gate_logits: attention map extents differ from x");
    static GateLogitsArgs a;   // ~2.7 KB: kept off the stack; calls are serialised by the host thread
    memset(&a, 0, sizeof(a));
    a.x = *x;
    if (att) { a.att = *att; a.has_att = 1; }
    a.out = outs[0];
    a.n_outs = n_outs;
    for (int i = 0; i < n_outs; ++i) {
        const vsseg_f32view& o = outs[i];
        VSSEG_REQUIRE(f32_ok(&o) && o.C == cout && o.X == x->X && o.Y == x->Y && o.Z == x->Z && o.B == (n_outs == 1 ? x->B : 1) &&
                          o.sc == a.out.sc && o.sx == a.out.sx && o.sy == a.out.sy && o.sz == a.out.sz && o.indirect == a.out.indirect,
                      "conv3d_gate_logits: output view %d does not match (extents, strides, base cell)", i);
        a.out_off[i] = (long long)(intptr_t)o.ptr;
    }
    a.sw_weight = sw_weight;
    a.atomic = atomic_blend ? 1 : 0;
    const int Y = x->Y, X = x->X;
    a.ny = Y <= GL_MAXL ? 1 : (Y + GL_MAXL - 3) / (GL_MAXL - 2);
    a.TY = (Y + a.ny - 1) / a.ny;
    a.L = a.ny == 1 ? Y : a.TY + 2;
    a.nz = x->Z / GL_TZ;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaGetLastError();
    // x segments: every segment re-projects one plane on either side, a CTA costs ~(XT + 2) planes; pick the
    // segmentation with the fewest plane-steps over the waves of 2 CTAs per SM
    static const int xt_env = getenv("VSSEG_GL_XT") ? atoi(getenv("VSSEG_GL_XT")) : 0;
    const long base = (long)x->B * a.nz * a.ny, slots = (long)GL_CTAS * sms;
    long best_cost = -1;
    for (int nxs = 1; nxs <= (X + 3) / 4; ++nxs) {
        const int XT = (X + nxs - 1) / nxs;
        if ((X + XT - 1) / XT != nxs) continue;
        if (xt_env > 0 && XT != xt_env && !(xt_env >= X && nxs == 1)) continue;
        const long waves = (base * nxs + slots - 1) / slots;
        const long cost = waves * (XT + (nxs > 1 ? 2 : 0)) + waves;   // + a fixed cost per wave
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; a.XT = XT; a.nxs = nxs; }
    }
    VSSEG_REQUIRE(best_cost >= 0, "conv3d_gate_logits: no x segmentation (VSSEG_GL_XT=%d)", xt_env);
    a.bias[0] = bias_host[0];
    a.bias[1] = cout > 1 ? bias_host[1] : 0.f;
    memcpy(a.w, w_host, sizeof(float) * 9 * GL_CIN * cout);
    const int threads = a.L * GL_TZ;
    const size_t smem = (size_t)2 * 6 * cout * threads * sizeof(float);
    auto kern = cout == 1 ? gate_logits_kernel<1> : gate_logits_kernel<2>;
    static bool attr[2] = {false, false};
    if (!attr[cout - 1]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 6 * 2 * GL_MAXL * GL_TZ * (int)sizeof(float));
        if (e != cudaSuccess) {
            set_error("conv3d_gate_logits: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[cout - 1] = true;
    }
    kern<<<(unsigned)(base * a.nxs), threads, smem, (cudaStream_t)stream>>>(a);
    return check_launch("conv3d_gate_logits");
}

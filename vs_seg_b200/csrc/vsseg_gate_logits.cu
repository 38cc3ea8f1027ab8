// Top of the decoder as ONE bandwidth-bound launch: AttentionBlock2 gate (reference attentionblock.py:44-47) fused
// into the last ResidualUnit (conv_only (3,3,1) conv + 1x1x1 shortcut folded into its centre tap,
// unet2d5_spvPA.py:186-190) and the sliding-window blend (MONAI sliding_window_inference step 6, call site
// VSparams.py:568-574).  The gated 2c-channel tensor x*(1+att) is consumed by nothing else, so it is never
// written: that removes the largest pure round trip of the network (8 B x 32 channels per voxel).
//
// Formulation.  conv(g*x)[xo,yo] = sum_{tx,ty} W[tx][ty] . (g*x)[xo+tx-1, yo+ty-1].  Every INPUT voxel is read
// exactly once and projected onto its 9 x Cout contributions P[tx][ty] = sum_c W[tx][ty][c] x[c]; the stencil is
// then closed in two cheap steps:
//   y: the lines of a plane exchange P through shared memory
//        Q[tx](yo) = g(yo-1) P[tx][0](yo-1) + g(yo) P[tx][1](yo) + g(yo+1) P[tx][2](yo+1)
//   x: the CTA marches along x and keeps the two open output rows in registers:
//        row(xi-1) = r_prev + Q[2] (complete -> emitted), r_prev' = r_cur + Q[1], r_cur' = Q[0].
// Round-2 history (profiles/r02_ncu_gate_logits_details.txt): with the projection as 576 fp32 FMAs per voxel the
// kernel sat at 0.37 of its HBM roofline; moving the projection to the warp-level tensor path changed nothing,
// which showed the real bound: one plane step is a serial chain (global load -> project -> exchange -> blend
// read-modify-write) with nothing in flight behind it.  This version is a software pipeline:
//   * every thread copies the 128 B of its voxel (4 channel groups x hi/lo plane) with cp.async into a ring of
//     GL_NST plane stages, GL_NST - 1 planes ahead of the one being consumed;
//   * the projection runs on mma.sync m16n8k16 (bf16 operands, fp32 accumulators) as "bf16x3" like every other conv
//     of the network: x_hi*w_hi + x_lo*w_hi share one K = 16 step (8 hi channels | the same 8 lo channels against
//     the same w_hi twice), x_hi*w_lo pairs two channel groups: 6 K steps for 32 channels.  An M tile is 16 voxels
//     = two y lines x 8 z; one ldmatrix.x4 yields the A fragment of a (tile, channel group) from the staged voxels;
//   * the blend operands (weight map, old accumulator value) of the row about to be emitted are fetched at the top
//     of the step.
// A CTA is (up to GL_MAXL y lines) x (8 z) threads; a thread is one voxel in the copy and in the stencil-closing
// half of a plane step.  Algorithmic bytes per voxel: 4*Cin (x) + 4 (att) + 8*Cout (+4 weight map) blend.
#include <stdlib.h>
#include <string.h>

#include "vsseg_common.cuh"

namespace vsseg {

constexpr int GL_TZ = 8;
#ifndef VSSEG_GL_MAXL
#define VSSEG_GL_MAXL 34
#endif
#ifndef VSSEG_GL_NST
#define VSSEG_GL_NST 2
#endif
#ifndef VSSEG_GL_BREG
#define VSSEG_GL_BREG 1
#endif
#ifndef VSSEG_GL_CTAS
#define VSSEG_GL_CTAS 2
#endif
constexpr int GL_CTAS = VSSEG_GL_CTAS;   // CTAs per SM (registers and the plane ring are cut for it)
constexpr int GL_MAXL = VSSEG_GL_MAXL;   // y lines per CTA (incl. the two halo lines of an interior tile)
constexpr int GL_NST = VSSEG_GL_NST;     // plane stages of the cp.async ring
constexpr int GL_MAXT = (GL_MAXL * GL_TZ + 31) / 32 * 32;   // threads: whole warps (mma.sync), lines beyond L are idle rows
constexpr int GL_MAXW = 16;
constexpr int GL_CIN = 32;
constexpr int GL_NT = 3;                 // n tiles of 8 columns (column = tap * Cout + o; 18 of 24 used for Cout = 2)

struct GateLogitsArgs {
    vsseg_act8 x;
    vsseg_f32view att, out;
    long long out_off[GL_MAXW];   // n_outs > 1: address (or byte offset from the cell when out.indirect) of entry b's view
    const float* sw_weight;
    int has_att, n_outs, atomic;
    int TY, ny, L;                // y lines emitted per tile, y tiles, lines staged per tile (TY + 2 halo lines when ny > 1)
    int XT, nxs, nz;              // x rows per segment, x segments, z tiles
    float bias[2];
    // B fragments in lane order (lane = 4 * g + t holds column g, k pair 2t / 2t+1 of the tile):
    //   wh[cg][nt]       : bf16x2 of w_hi[col][cg*8 + 2t, +1]   (used for BOTH k halves: x_hi and x_lo of the group)
    //   wl[pair][nt][h]  : bf16x2 of w_lo[col][(2*pair + h)*8 + 2t, +1]
    uint32_t wh[GL_CIN / 8][GL_NT][32];
    uint32_t wl[GL_CIN / 16][GL_NT][2][32];
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&d)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(saddr)
                 : "memory");
}

template <int COUT>
#ifdef VSSEG_GL_MAXREG
__global__ void __maxnreg__(VSSEG_GL_MAXREG) gate_logits_kernel(const __grid_constant__ GateLogitsArgs a) {
#else
__global__ void __launch_bounds__(GL_MAXT, GL_CTAS) gate_logits_kernel(const __grid_constant__ GateLogitsArgs a) {
#endif
    // shared memory: [GL_NST stages][hi | lo][channel group][thread] x 16 B (a consumed stage doubles as the exchange
    // buffer of its plane), the gate factors [thread] and the B fragments
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NCG = GL_CIN / 8, NCOL = 9 * COUT, NT = (NCOL + 7) / 8;
    const int nthr = blockDim.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, fg = lane >> 2, ft = lane & 3;
    const int zl = tid % GL_TZ, line = tid / GL_TZ;
    const uint32_t stage_bytes = (uint32_t)nthr * 128u;
    float* gate_s = reinterpret_cast<float*>(smem_raw + (size_t)GL_NST * stage_bytes);
    int t = blockIdx.x;
    const int tz = t % a.nz; t /= a.nz;
    const int tyi = t % a.ny; t /= a.ny;
    const int xs = t % a.nxs; t /= a.nxs;
    const int b = t;
    const int X = a.x.X, Y = a.x.Y, Z = a.x.Z;
    const int y0 = tyi * a.TY;
    const int ybase = y0 - (a.ny > 1 ? 1 : 0);   // input line of CTA line 0
    const int x_lo = xs * a.XT, x_hi = min(x_lo + a.XT, X);
    const int64_t xstep = (int64_t)Y * Z * 8, lo = a.x.lo_offset;
    const int cgbytes = X * Y * Z * 16;          // one channel group (bytes; < 2^29, checked by the host wrapper)

    // ---- this thread's voxel (line, zl): copy source, gate, blend destination.  Lines outside the tile or the volume
    // copy the nearest valid line and are cancelled by a zero gate factor, so every copy is unconditional
    const int z = tz * GL_TZ + zl;
    const int y = ybase + line, yc = min(max(y, 0), Y - 1);
    const bool in_y = line < a.L && y == yc;
    const bool emit_y = in_y && y >= y0 && y < y0 + a.TY;
    const char* src = reinterpret_cast<const char*>((const __nv_bfloat16*)a.x.hi + (int64_t)b * a.x.batch_stride + ((int64_t)yc * Z + z) * 8);
    const float* attp = a.has_att ? f32_base(a.att) + b * a.att.sb + yc * a.att.sy + z * a.att.sz : nullptr;
    float* outp;
    if (a.n_outs > 1) {
        const long long base = a.out.indirect ? __ldg(reinterpret_cast<const long long*>(a.out.indirect)) : 0ll;
        outp = reinterpret_cast<float*>(base + a.out_off[b]);
    } else {
        outp = f32_base(a.out) + b * a.out.sb;
    }
    outp += yc * a.out.sy + z * a.out.sz;
    const float* swp = a.sw_weight ? a.sw_weight + (int64_t)yc * Z + z : nullptr;
    const int64_t sw_xstep = (int64_t)Y * Z;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t my_slot = smem_base + (uint32_t)tid * 16u;   // + stage * stage_bytes + (plane * NCG + cg) * nthr * 16

    auto copy_plane = [&](int xi, int stage) {
        const char* ph = src + 2 * (xi * xstep);
        const char* pl = ph + 2 * lo;
        const uint32_t d = my_slot + (uint32_t)stage * stage_bytes;
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) {
            cp_async16(d + (uint32_t)(cg * nthr) * 16u, ph + cg * cgbytes);
            cp_async16(d + (uint32_t)((NCG + cg) * nthr) * 16u, pl + cg * cgbytes);
        }
    };

    // B fragments: kernel parameters -> shared memory (lane-indexed, conflict-free reads)
    constexpr int NWH = NCG * GL_NT * 32, NWL = (NCG / 2) * GL_NT * 2 * 32;
#if VSSEG_GL_BREG
    // ... and on into registers (24 for Cout = 2): the plane loop is bound by shared-memory instructions, not by
    // registers.  The staging area is the last plane stage, which is first filled after the loop's first barrier.
    // (Read straight from the parameters, ptxas re-materialises them in the loop as divergent constant-bank loads.)
    uint32_t* wsm = reinterpret_cast<uint32_t*>(smem_raw + (size_t)(GL_NST - 1) * stage_bytes);
#else
    uint32_t* wsm = reinterpret_cast<uint32_t*>(gate_s + nthr);
#endif
    for (int i = tid; i < NWH + NWL; i += nthr) wsm[i] = i < NWH ? (&a.wh[0][0][0])[i] : (&a.wl[0][0][0][0])[i - NWH];
    const uint32_t* whs = wsm + lane;          // [cg][nt] at (cg * GL_NT + nt) * 32
    const uint32_t* wls = wsm + NWH + lane;    // [pair][nt][h] at ((pair * GL_NT + nt) * 2 + h) * 32
#if VSSEG_GL_BREG
    __syncthreads();
    uint32_t bh[NCG][NT], bl[NCG / 2][NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) bh[cg][nt] = whs[(cg * GL_NT + nt) * 32];
#pragma unroll
        for (int s = 0; s < NCG / 2; ++s) {
            bl[s][nt][0] = wls[((s * GL_NT + nt) * 2 + 0) * 32];
            bl[s][nt][1] = wls[((s * GL_NT + nt) * 2 + 1) * 32];
        }
    }
#endif
    // ldmatrix row address of this lane: matrix j = lane / 8 -> (plane j / 2, line half j % 2), row lane % 8 = z
    const uint32_t ldm_off = (uint32_t)(((lane >> 4) * NCG) * nthr + warp * 32 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16u;

    float r_prev[COUT], r_cur[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) r_prev[o] = r_cur[o] = 0.f;

    const int xi0 = max(x_lo - 1, 0), xi1 = min(x_hi, X - 1);
#pragma unroll
    for (int s = 0; s < GL_NST - 1; ++s) {
        if (xi0 + s <= xi1) copy_plane(xi0 + s, s);
        cp_async_commit();
    }
    int it = 0;
    for (int xi = xi0; xi <= xi1; ++xi, ++it) {
        // operands of this step's blend (row xi - 1) and this voxel's gate: in flight while the projection runs
        const int xo = xi - 1;
        const bool will_emit = emit_y && xo >= x_lo && xo < x_hi;
        float sw = 0.f, old[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) old[o] = 0.f;
        if (will_emit && swp) {
            sw = __ldg(swp + xo * sw_xstep);
            if (!a.atomic) {
#pragma unroll
                for (int o = 0; o < COUT; ++o) old[o] = outp[o * a.out.sc + xo * a.out.sx];
            }
        }
        const float gate = !in_y ? 0.f : attp ? 1.0f + __ldg(attp + xi * a.att.sx) : 1.0f;
        cp_async_wait<GL_NST - 2>();
        __syncthreads();   // plane xi has landed for every thread; the previous step's exchange reads are done,
                           // so the stage that held them can be refilled
        if (xi + GL_NST - 1 <= xi1) copy_plane(xi + GL_NST - 1, (it + GL_NST - 1) % GL_NST);
        cp_async_commit();

        // The projections P of a voxel overwrite the staged input of the SAME voxel (an M tile is read completely by
        // its own warp before that warp stores): column c of voxel v lives in the 16-byte cell (slot c / 4, thread v)
        unsigned char* cells = smem_raw + (size_t)(it % GL_NST) * stage_bytes;
        const uint32_t stage = smem_base + (uint32_t)(it % GL_NST) * stage_bytes + ldm_off;
        auto cell = [&](int col, int v) { return reinterpret_cast<float*>(cells + ((size_t)(col >> 2) * nthr + v) * 16 + (col & 3) * 4); };
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float c[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
            uint32_t A[NCG][4];   // per channel group: hi line 0, hi line 1, lo line 0, lo line 1 of the tile
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) ldmatrix_x4(A[cg], stage + (uint32_t)(cg * nthr + mt * 16) * 16u);
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg)   // x_hi * w_hi + x_lo * w_hi: k 0..7 = hi channels, k 8..15 = lo channels
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
#if VSSEG_GL_BREG
                    const uint32_t bw = bh[cg][nt];
#else
                    const uint32_t bw = whs[(cg * GL_NT + nt) * 32];
#endif
                    mma_bf16_16816(c[nt], A[cg][0], A[cg][1], A[cg][2], A[cg][3], bw, bw);
                }
#pragma unroll
            for (int s = 0; s < NCG / 2; ++s)  // x_hi * w_lo: two channel groups per K step
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#if VSSEG_GL_BREG
                    mma_bf16_16816(c[nt], A[2 * s][0], A[2 * s][1], A[2 * s + 1][0], A[2 * s + 1][1], bl[s][nt][0], bl[s][nt][1]);
#else
                    mma_bf16_16816(c[nt], A[2 * s][0], A[2 * s][1], A[2 * s + 1][0], A[2 * s + 1][1],
                                   wls[((s * GL_NT + nt) * 2 + 0) * 32], wls[((s * GL_NT + nt) * 2 + 1) * 32]);
#endif
            // accumulator fragment: c[nt][0..1] = row fg (first line of the tile), columns nt*8 + 2*ft, +1; c[nt][2..3] =
            // row fg + 8 (second line).  Row fg is the voxel of thread v0, row fg + 8 the voxel 8 threads on
            const int v0 = warp * 32 + mt * 16 + fg;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = nt * 8 + 2 * ft;
                if (col < NCOL) {
                    *reinterpret_cast<float2*>(cell(col, v0)) = make_float2(c[nt][0], c[nt][1]);
                    *reinterpret_cast<float2*>(cell(col, v0 + 8)) = make_float2(c[nt][2], c[nt][3]);
                }
            }
        }
        gate_s[tid] = gate;
        __syncthreads();
        if (line < a.L) {
            const float gm = line > 0 ? gate_s[tid - GL_TZ] : 0.f, gp = line < a.L - 1 ? gate_s[tid + GL_TZ] : 0.f;
            const int tm = line > 0 ? tid - GL_TZ : tid, tp = line < a.L - 1 ? tid + GL_TZ : tid;
            float Q[3][COUT];
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                if constexpr (COUT == 2) {
                    const float2 p1 = *reinterpret_cast<const float2*>(cell((tx * 3 + 1) * 2, tid));
                    const float2 p0 = *reinterpret_cast<const float2*>(cell((tx * 3 + 0) * 2, tm));   // ty = 0 tap of line y-1
                    const float2 p2 = *reinterpret_cast<const float2*>(cell((tx * 3 + 2) * 2, tp));   // ty = 2 tap of line y+1
                    Q[tx][0] = fmaf(gp, p2.x, fmaf(gm, p0.x, gate * p1.x));
                    Q[tx][1] = fmaf(gp, p2.y, fmaf(gm, p0.y, gate * p1.y));
                } else {
                    Q[tx][0] = fmaf(gp, *cell(tx * 3 + 2, tp), fmaf(gm, *cell(tx * 3 + 0, tm), gate * *cell(tx * 3 + 1, tid)));
                }
            }
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                const float done = r_prev[o] + Q[2][o] + a.bias[o];
                r_prev[o] = r_cur[o] + Q[1][o];
                r_cur[o] = Q[0][o];
                if (will_emit) {
                    float* q = outp + o * a.out.sc + xo * a.out.sx;
                    if (!swp) *q = done;
                    else if (a.atomic) atomicAdd(q, sw * done);   // overlapping windows of one launch (red.global.add.f32)
                    else *q = old[o] + sw * done;
                }
            }
        }
    }
    if (x_hi == X && emit_y) {   // the last row of the volume has no plane behind it
        const int xo = X - 1;
        const float sw = swp ? __ldg(swp + xo * sw_xstep) : 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            float* q = outp + o * a.out.sc + xo * a.out.sx;
            const float r = r_prev[o] + a.bias[o];
            if (!swp) *q = r;
            else if (a.atomic) atomicAdd(q, sw * r);
            else *q += sw * r;
        }
    }
}

}  // namespace vsseg

using namespace vsseg;

static size_t gl_smem_bytes(int threads, int cout) {
    (void)cout;
    return (size_t)GL_NST * threads * 128 + (size_t)threads * sizeof(float) +
           (VSSEG_GL_BREG ? 0 : sizeof(((GateLogitsArgs*)0)->wh) + sizeof(((GateLogitsArgs*)0)->wl));
}

static uint32_t bf16_bits_rn(float f) {   // round to nearest even (finite inputs: these are conv weights)
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}

extern "C" int vsseg_conv3d_gate_logits(const vsseg_act8* x, const vsseg_f32view* att, const float* w_host,
                                        const float* bias_host, int32_t cout, const vsseg_f32view* outs, int32_t n_outs,
                                        const float* sw_weight, int32_t atomic_blend, void* stream) {
    VSSEG_REQUIRE(x && x->hi && x->C == GL_CIN && x->B >= 1 && x->X >= 1 && x->Y >= 1 && x->Z >= GL_TZ && x->Z % GL_TZ == 0,
                  "conv3d_gate_logits: x must be act8 with %d channels and Z %% %d == 0", GL_CIN, GL_TZ);
    VSSEG_REQUIRE(w_host && bias_host && (cout == 1 || cout == 2), "conv3d_gate_logits: Cout must be 1 or 2");
    VSSEG_REQUIRE((int64_t)x->X * x->Y * x->Z * 16 * (GL_CIN / 8) < (1ll << 31), "conv3d_gate_logits: window too large for 32-bit plane offsets");
    VSSEG_REQUIRE(outs && (n_outs == 1 || (n_outs == x->B && n_outs <= GL_MAXW)),
                  "conv3d_gate_logits: n_outs must be 1 or B (<= %d)", GL_MAXW);
    VSSEG_REQUIRE(!att || (f32_ok(att) && att->B == x->B && att->X == x->X && att->Y == x->Y && att->Z == x->Z),
                  "conv3d_gate_logits: attention map extents differ from x");
    static GateLogitsArgs a;   // ~3.5 KB: kept off the stack; calls are serialised by the host thread
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
    // segmentation with the fewest plane-steps over the waves of GL_CTAS CTAs per SM
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
    // split-bf16 B fragments (same rounding as the act8 planes: hi = rn(w), lo = rn(w - hi))
    for (int nt = 0; nt < GL_NT; ++nt)
        for (int l = 0; l < 32; ++l) {
            const int col = nt * 8 + (l >> 2), k0 = 2 * (l & 3);
            for (int cg = 0; cg < GL_CIN / 8; ++cg) {
                uint32_t hi2 = 0, lo2 = 0;
                if (col < 9 * cout)
                    for (int e = 0; e < 2; ++e) {
                        const float w = w_host[((size_t)(col / cout) * GL_CIN + cg * 8 + k0 + e) * cout + col % cout];
                        const uint32_t h = bf16_bits_rn(w);
                        float hf;
                        const uint32_t hb = h << 16;
                        memcpy(&hf, &hb, 4);
                        hi2 |= h << (16 * e);
                        lo2 |= bf16_bits_rn(w - hf) << (16 * e);
                    }
                a.wh[cg][nt][l] = hi2;
                a.wl[cg / 2][nt][cg % 2][l] = lo2;
            }
        }
    const int threads = (a.L * GL_TZ + 31) / 32 * 32;
    const size_t smem = gl_smem_bytes(threads, cout);
    auto kern = cout == 1 ? gate_logits_kernel<1> : gate_logits_kernel<2>;
    static bool attr[2] = {false, false};
    if (!attr[cout - 1]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gl_smem_bytes(GL_MAXT, 2));
        if (e != cudaSuccess) {
            set_error("conv3d_gate_logits: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[cout - 1] = true;
    }
    kern<<<(unsigned)(base * a.nxs), threads, smem, (cudaStream_t)stream>>>(a);
    return check_launch("conv3d_gate_logits");
}

// Tensor-core (tcgen05 + TMEM + TMA) implicit-GEMM 3-D convolution of UNet2d5_spvPA, sm_100a only.
// Covers every Conv3d / ConvTranspose3d of the network whose channels are multiples of 16/8
// (reference params/networks/blocks/convolutions.py:125-156): stride-1 "same" convs, the strided
// downsample convs, the transposed (sub-pixel phase decomposed) upsample convs and the 1x1x1
// ResidualUnit shortcut fused as a second accumulator (convolutions.py:241-255).
//
// GEMM view.  An M tile is 128 positions of the "M grid" (the output grid of a conv, the INPUT grid
// of a transposed conv): LY consecutive y lines x LZ consecutive z (LY*LZ = 128, LZ = the largest power of two <= 128
// that divides Z).
// N = a slice of Cout, K = Cin x taps walked as stages (16-channel chunk c, x tap j).  Operands
// are the two bf16 planes of the act8 layout (value = hi + lo); every product is evaluated as
// hi*hi + lo*hi + hi*lo ("bf16x3") into one fp32 TMEM accumulator, which keeps the network inside
// the 1e-3 parity bar that single-pass bf16/tf32 misses.
//
// A CTA owns NACC accumulators (y line groups x output phases) of one x row.  Per stage the TMA
// producer stages a few boxes of the input (5-D tiled map; OOB zero fill = "same" padding and the
// output_padding row of the transposed conv; traversal strides = conv stride) plus the packed
// weights of the stage; the single MMA thread then walks a host-built op table
// (A offset, B offset, accumulator) - the shifted views of the staged boxes ARE the im2col matrix,
// nothing is gathered by threads.  Epilogue warps: tcgen05.ld -> BN scale/shift -> PReLU/ReLU ->
// (+ residual / + shortcut accumulator) -> split-bf16 -> global (act8).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "vsseg_ptx.cuh"

namespace vsseg {

// ---- host-built tables ------------------------------------------------------------------------
struct TcOp {        // one tcgen05.mma: D[128 x n8*8 columns at col] += A[128 x 16] * B[16 x n8*8]
    uint16_t a16;    // A start offset inside the stage (hi or lo plane included), in 16 B units
    uint16_t b16;    // B start offset inside the stage's B region (hi or lo plane included), 16 B units
    uint16_t col;    // first TMEM column written
    uint16_t n8;     // N / 8 of this instruction (several adjacent accumulators share one A read)
};
// The MMA-issuing warp shares its scheduler with four ALU-busy epilogue warps, so every instruction it
// needs per MMA costs issue slots it only gets a fraction of (tools/ubench/mma_ctx.cu: 45 -> 110-160
// cycles per N=48 MMA next to busy warps; measured ~75-100 in this kernel with 15 instructions per MMA).
// The device therefore reads one ready-made 16-byte record per merged product and derives the three
// bf16x3 passes from it with two adds: ~4 uniform instructions per MMA.
struct __align__(16) TcEl { uint32_t a_lo, b_lo, col, idesc; };   // one merged product = 3 MMAs (hi*hi, lo*hi, hi*lo)
struct TcBox {       // one TMA box of a stage (issued for the hi and the lo plane)
    uint16_t dst16;  // offset inside the hi-plane A region, 16 B units
    int8_t dy, dz;   // input-coordinate shift of the box origin
};
struct TcAcc {       // where an accumulator lands in the output
    int16_t y_add;   // output y of M-tile line 0, relative to the tile's output y base
    int8_t z_add;    // output z phase
    int8_t yl;       // M-grid line offset of the accumulator's line group (for the partial-group mask)
    int32_t off8;    // (y_add * Zout + z_add) * 8: element offset of the accumulator inside an act8 channel group
};

constexpr int TC_MAX_OPS = 144;
constexpr int TC_MAX_OPS2 = 48;
constexpr int TC_MAX_BOX = 9;
constexpr int TC_MAX_ACC = 32;

struct TcArgs {
    vsseg_act8 out;
    vsseg_epilogue ep;
    int res_mode;            // 0 none, 1 act8 addend, 2 cin1 affine
    vsseg_act8 res;
    vsseg_f32view rsrc;
    WinTab rwin;             // rsrc is a window set: batch item b reads its window at rwin.off[b]
    const float* res_w;
    const float* res_b;
    const float* bias2;      // shortcut bias [Cout] (segment 2)
    const uint8_t* w;        // packed main weights  [sel][chunk][j][b_bytes]
    const uint8_t* w2;       // packed shortcut weights [nsel][chunk2][b2_bytes]
    int nchunk, nj, nchunk2; // stages: nchunk*nj main + nchunk2 shortcut
    int nop, nop2, nbox, nacc;
    int nstage;              // ring depth
    uint32_t a_plane, b_plane, b_off, stage_bytes;  // hi->lo distance of A / B inside a stage, B region start
    uint32_t box_tx, b_bytes, b2_bytes, b2_plane;   // bytes of one box (one plane), of the weight slices
    uint32_t lbo_a, lbo_b, lbo_b2, idesc, tmem_cols;
    // tile decomposition of the M grid
    int ntz, nty, ntx, nsel, npx, nsplit, ntiles, nbuf;
    int XT, Xm;              // x rows per tile (the CTA marches along them), x extent of the M grid
    int LZ, LY, YL, Ym;      // M-tile shape, y lines of the M grid per CTA tile, y extent of the M grid
    int n_cta;               // output channels per CTA (multiple of 16), n_real: channels to store
    int cout;                // real Cout (multiple of 8)
    // input coordinates of box origin: x = mx*sx + j + xoff ; y = my0*sy + box.dy ; z = mz0*sz + box.dz
    int sx, sy, sz, xoff, Xin;
    int cg_plane, cg_batch;        // merged-cg index strides of the main map (lo plane, batch)
    int cg_plane2, cg_batch2;      // same for the shortcut source
    // output coordinates: ox = mx*ux + px ; oy = (my0 + ly)*uy + acc.y_add ; oz = (mz0 + zz)*uz + acc.z_add
    int ux, uy, uz;
    int dy2, dz2;            // box origin shift of the shortcut source
    // line mode (M tile = one z line of 128): A staged by per-line bulk copies instead of tensor-map boxes
    int line_mode, map_wide, BY, pitch, hy, hz, Yin, Zin;
    int out_mode;            // 0: act8 output, 1: planar fp32 output (Cout <= 16) with optional sliding-window blend
    vsseg_f32view outf;
    const float* sw_weight;
    vsseg_act8 in, in2;
    unsigned long long* dbg;  // VSSEG_TC_DEBUG: per-CTA cycle counters [gridDim.x][8]
    int dbgf;                 // VSSEG_TC_DBGF experiments (timing only, results invalid): 1 no epilogue work, 2 no copies
    TcOp ops[TC_MAX_OPS];      // host-side description (plan dump)
    TcOp ops2[TC_MAX_OPS2];
    TcBox boxes[TC_MAX_BOX];
    TcAcc accs[TC_MAX_ACC];
    TcEl el[TC_MAX_OPS / 3];   // what the MMA thread reads
    TcEl el2[TC_MAX_OPS2 / 3];
    uint32_t desc_hi;          // upper descriptor word (SBO = 128 B, version 1)
    int partial;               // the last y line group is partial (rows beyond Ym are masked)
    vsseg_act8 gate;           // out mode 2: AttentionBlock2 gate applied in place to this tensor, x *= 1 + att
    int two_pass;              // Cout <= 2 planar output: weights packed [hi | lo] along N, two MMAs per product
    // TS mode (k = (.,.,1) stride-1 convs on 128-long z lines): the A operand is copied once per staged plane from shared
    // memory to tensor memory (tcgen05.cp) and every MMA reads it from there - the SS-mode MMA re-reads its 4 KB A view
    // from shared memory on every instruction (the 32 + N/4 cycle floor), which is what bounds the narrow layers
    int ts_mode;
    uint32_t a_tmem_col;       // first TMEM column of the A buffers
    uint32_t a_cols;           // columns of one A buffer: BY views x 2 planes x 8
    int na;                    // A buffers
    // The shortcut reads the SAME tensor as the conv (decoder ResidualUnits): no shortcut stages - the stage of the
    // centre x tap of a row already holds the shortcut's A views (centre y/z tap); it also carries the shortcut weights
    // of its channel chunk at b2_off and the row's issuer adds the shortcut MMAs to it
    int sc_self;
    uint32_t b2_off;
    // programmatic dependent launch (VSSEG_TC_PDL): the kernel lets its successor's CTAs become resident as SMs free up
    // and itself waits for its predecessor only after its own set-up (barriers, tensor memory, epilogue constants)
    int pdl;
};

constexpr int TC_HDR = 1024 + 5 * 1024;  // barriers + epilogue constants
#ifndef VSSEG_GATE_BATCH
#define VSSEG_GATE_BATCH 4   // fused attention gate (out mode 2): channel groups loaded ahead of the first store (measured: 1 -> 4: dec3 0.119 -> 0.080, dec2 0.222 -> 0.198 ms per group of 8 windows; 8 spills and is slower)
#endif
#ifndef VSSEG_TC_EPI_WARPS
#define VSSEG_TC_EPI_WARPS 16
#endif
#ifndef VSSEG_TC_MMA_WARPS
#define VSSEG_TC_MMA_WARPS 3
#endif
constexpr int TC_EPI_WARPS = VSSEG_TC_EPI_WARPS;  // four warps per TMEM lane quadrant, each draining every fourth (accumulator, 16-column) unit
// MMA issue is the bottleneck of the narrow layers: one thread sustains one MMA per ~60 cycles (5 uniform
// instructions each, slower next to ALU-busy epilogue warps; tools/ubench/mma_issue.cu, mma_ctx.cu) while an
// N=48 MMA executes in 44.  Three issuing warps therefore own the output ROWS round-robin (row g -> issuer
// g % 3): a plane of the x march feeds three consecutive rows, i.e. one per issuer, and every TMEM
// accumulator is only ever written by one thread (MMAs of different threads into the SAME columns are
// not ordered - splitting one stage's products over warps gave sporadic lost updates).  Completion is
// tracked by one commit per issuer on every barrier.
constexpr int TC_MMA_WARPS = VSSEG_TC_MMA_WARPS;
constexpr int TC_EPI_WARP0 = 1 + TC_MMA_WARPS;   // first epilogue warp (a multiple of 4 plus 1: quadrant = warp % 4)
constexpr int TC_CP_WARP = TC_EPI_WARP0 + TC_EPI_WARPS;   // TS mode: the warp that copies staged A views to tensor memory
constexpr int TC_THREADS = 32 * (TC_CP_WARP + 1);  // warp 0: producer, warps 1-3: MMA issuers (warp 1 owns TMEM), 16 epilogue warps, copier
constexpr int TC_MAX_SLOT = 8;    // accumulator row slots in TMEM

__device__ uint4 g_zero_line[136];  // source of padding lines / halo rows for the bulk-copy producer (zero-initialised)

struct TcTile {
    int sel, tz, ty, mx, b, px, ns, my0, mz0;
    int xt;          // x rows of this tile (the last x segment may be shorter)
    int q_lo, q_hi;  // first / last staged x plane that lies inside the volume
};
__device__ __forceinline__ TcTile decode_tile(const TcArgs& a, int t) {
    TcTile T;
    T.sel = t % a.nsel; t /= a.nsel;
    T.tz = t % a.ntz; t /= a.ntz;
    T.ty = t % a.nty; t /= a.nty;
    T.mx = (t % a.ntx) * a.XT; t /= a.ntx;   // first x row of the tile
    T.b = t;
    T.px = T.sel / a.nsplit; T.ns = T.sel % a.nsplit;
    T.my0 = T.ty * a.YL; T.mz0 = T.tz * a.LZ;
    T.xt = min(a.XT, a.Xm - T.mx);
    // plane q holds input x = mx*sx + q + xoff and feeds row r = q - dx through x tap dx.
    // transposed conv, even x phase: only the centre x tap (plane 0)
    const int qlim = (a.npx == 2 && T.px == 0) ? 1 : T.xt + a.nj - 1;
    const int x0 = T.mx * a.sx + a.xoff;
    T.q_lo = max(0, -x0);
    T.q_hi = min(qlim - 1, a.Xin - 1 - x0);
    return T;
}

struct EpiCtx {
    const TcArgs& a;
    const float* ep_c;
    uint32_t tacc, row_cols;
    int64_t row_off, cgs, lo_out, lo_res;
    __nv_bfloat16* out_b;
    const __nv_bfloat16* res_b;
    int b, ox, my0, mz0, ly, zz, co0, nreal;
    bool sc, sig;
    float slope;
    const float* rsrc_p;   // base pointers of the fp32 views (relocatable views resolved once per thread); rsrc_p: of batch item b
    float* outf_p;
};
template <bool SC, int RM>
struct EpiUnit {
    int ai, c;
    uint32_t v[16], v2[SC ? 16 : 1];
    uint4 rh[RM == 1 ? 2 : 1], rl[RM == 1 ? 2 : 1];
    float rsrc;
    int oy, oz;
    bool live, valid;
};

// issue the TMEM loads of one unit and fetch its residual operands (they arrive while the loads are in flight)
template <int OM, bool SC, int RM>
__device__ __forceinline__ void epi_load(const EpiCtx& X, EpiUnit<SC, RM>& U) {
    const TcArgs& a = X.a;
    const int c0 = U.c << 4;
    U.live = c0 < X.nreal;   // padding columns of the Cout slice (warp-uniform)
    if (!U.live) return;
    const uint32_t taddr = X.tacc + (uint32_t)(U.ai * a.n_cta + c0);
    tmem_ld16(taddr, U.v);
    if constexpr (SC) tmem_ld16(taddr + X.row_cols, U.v2);
    // rows of a partial (zero-padded) line group load their TMEM lane like everyone else
    // (tcgen05.ld is warp-collective) but neither read the residual nor store
    U.valid = !a.partial || X.my0 + X.ly + a.accs[U.ai].yl < a.Ym;
    U.rsrc = 0.f;
    U.oy = U.oz = 0;
    if constexpr (RM == 2 || OM >= 1) {
        U.oy = (X.my0 + X.ly) * a.uy + a.accs[U.ai].y_add;
        U.oz = (X.mz0 + X.zz) * a.uz + a.accs[U.ai].z_add;
    }
    if constexpr (RM == 1) {
      if (U.valid) {
        const __nv_bfloat16* rp = X.res_b + X.row_off + a.accs[U.ai].off8 + (int64_t)(U.c * 2) * X.cgs;
        U.rh[0] = ldg128(rp);
        U.rl[0] = ldg128(rp + X.lo_res);
        if (X.nreal - c0 > 8) {
            U.rh[1] = ldg128(rp + X.cgs);
            U.rl[1] = ldg128(rp + X.cgs + X.lo_res);
        }
      }
    } else if constexpr (RM == 2) {
        if (U.valid) U.rsrc = __ldg(X.rsrc_p + X.ox * a.rsrc.sx + U.oy * a.rsrc.sy + U.oz * a.rsrc.sz);
    }
}

// BN scale/shift -> PReLU/ReLU/sigmoid -> (+ shortcut accumulator | + residual) -> split-bf16 -> global
template <int OM, bool SC, int RM>
__device__ __forceinline__ void epi_finish(const EpiCtx& X, EpiUnit<SC, RM>& U) {
    const TcArgs& a = X.a;
    if (!U.live || !U.valid || (a.dbgf & 1)) return;
    const int c0 = U.c << 4;
    const float* ep_c = X.ep_c;
    if constexpr (OM >= 1) {
        // planar fp32 output: attention map (sigmoid) or logits, optionally blended into the
        // sliding-window accumulator (MONAI sliding_window_inference step 6)
        const int Yo = a.out.Y, Zo = a.out.Z;
        const float sw = a.sw_weight ? __ldg(a.sw_weight + ((int64_t)X.ox * Yo + U.oy) * Zo + U.oz) : 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {   // Cout <= 2 on this path (static indices keep v[] in registers)
            if (q >= a.cout) break;
            float raw = __uint_as_float(U.v[q]);
            if (a.two_pass) raw += __uint_as_float(a.cout == 1 ? U.v[q + 1] : U.v[q + 2]);   // + hi*lo partial sums
            float f = raw * ep_c[q] + ep_c[256 + q];
            f = apply_act(f, a.ep.act, X.slope);
            float* o = X.outf_p + X.b * a.outf.sb + q * a.outf.sc + X.ox * a.outf.sx + U.oy * a.outf.sy + U.oz * a.outf.sz;
            if (a.sw_weight) *o += sw * f;
            else *o = f;
            if constexpr (OM == 2) {
                // AttentionBlock2 (reference attentionblock.py:39-47) fused into the conv that produces the map:
                // every channel group of this voxel is scaled by 1 + att in place (split-bf16 -> fp32 -> split-bf16)
                if (q == 0) {
                    const float gsc = 1.0f + f;
                    const int Xg = a.gate.X, Yg = a.gate.Y, Zg = a.gate.Z;
                    const int64_t cgs_g = (int64_t)Xg * Yg * Zg * 8;
                    __nv_bfloat16* gp = (__nv_bfloat16*)a.gate.hi + (int64_t)X.b * a.gate.batch_stride +
                                        (((int64_t)X.ox * Yg + U.oy) * Zg + U.oz) * 8;
                    const int ncg = a.gate.C >> 3;
                    int cg = 0;
#if VSSEG_GATE_BATCH > 1
                    // the loads of VSSEG_GATE_BATCH channel groups are issued before the first store (the compiler cannot
                    // move a load above a store through the same pointer): 2 x VSSEG_GATE_BATCH 128-bit loads in flight
                    for (; cg + VSSEG_GATE_BATCH <= ncg; cg += VSSEG_GATE_BATCH, gp += VSSEG_GATE_BATCH * cgs_g) {
                        uint4 hh[VSSEG_GATE_BATCH], ll[VSSEG_GATE_BATCH];
#pragma unroll
                        for (int j = 0; j < VSSEG_GATE_BATCH; ++j) {
                            hh[j] = ldg128_plain(gp + j * cgs_g);
                            ll[j] = ldg128_plain(gp + j * cgs_g + a.gate.lo_offset);
                        }
#pragma unroll
                        for (int j = 0; j < VSSEG_GATE_BATCH; ++j) {
                            float xv[8];
                            unpack8(hh[j], ll[j], xv);
#pragma unroll
                            for (int k = 0; k < 8; ++k) xv[k] *= gsc;
                            uint4 h, l;
                            pack8(xv, h, l);
                            *reinterpret_cast<uint4*>(gp + j * cgs_g) = h;
                            *reinterpret_cast<uint4*>(gp + j * cgs_g + a.gate.lo_offset) = l;
                        }
                    }
#endif
#pragma unroll 2
                    for (; cg < ncg; ++cg, gp += cgs_g) {
                        float xv[8];
                        unpack8(ldg128_plain(gp), ldg128_plain(gp + a.gate.lo_offset), xv);
#pragma unroll
                        for (int k = 0; k < 8; ++k) xv[k] *= gsc;
                        uint4 h, l;
                        pack8(xv, h, l);
                        *reinterpret_cast<uint4*>(gp) = h;
                        *reinterpret_cast<uint4*>(gp + a.gate.lo_offset) = l;
                    }
                }
            }
        }
        return;
    } else {
    const int ngrp = min(2, (X.nreal - c0 + 7) >> 3);
    __nv_bfloat16* op = X.out_b + X.row_off + a.accs[U.ai].off8 + (int64_t)(U.c * 2) * X.cgs;
    const float* e = ep_c + X.co0 + c0;
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
        if (g8 >= ngrp) break;
        float scl[8], sft[8], o[8];
        *reinterpret_cast<float4*>(scl) = *reinterpret_cast<const float4*>(e + g8 * 8);
        *reinterpret_cast<float4*>(scl + 4) = *reinterpret_cast<const float4*>(e + g8 * 8 + 4);
        *reinterpret_cast<float4*>(sft) = *reinterpret_cast<const float4*>(e + g8 * 8 + 256);
        *reinterpret_cast<float4*>(sft + 4) = *reinterpret_cast<const float4*>(e + g8 * 8 + 260);
        if (X.sig) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float f = fmaf(__uint_as_float(U.v[g8 * 8 + q]), scl[q], sft[q]);
                o[q] = 1.0f / (1.0f + expf(-f));
            }
        } else {
            // PReLU / ReLU / identity: f + (slope - 1) * min(f, 0)
            const float sm1 = X.slope - 1.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float f = fmaf(__uint_as_float(U.v[g8 * 8 + q]), scl[q], sft[q]);
                o[q] = fmaf(fminf(f, 0.f), sm1, f);
            }
        }
        if constexpr (SC || RM == 2) {
            float post[8];
            *reinterpret_cast<float4*>(post) = *reinterpret_cast<const float4*>(e + g8 * 8 + 512);
            *reinterpret_cast<float4*>(post + 4) = *reinterpret_cast<const float4*>(e + g8 * 8 + 516);
            if constexpr (SC) {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] += __uint_as_float(U.v2[g8 * 8 + q]) + post[q];
            } else {
                float rw[8];
                *reinterpret_cast<float4*>(rw) = *reinterpret_cast<const float4*>(e + g8 * 8 + 768);
                *reinterpret_cast<float4*>(rw + 4) = *reinterpret_cast<const float4*>(e + g8 * 8 + 772);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] += fmaf(rw[q], U.rsrc, post[q]);
            }
        }
        if constexpr (RM == 1) {
            float rr[8];
            unpack8(U.rh[g8], U.rl[g8], rr);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] += rr[q];
        }
        uint4 h, l;
        pack8(o, h, l);
        __nv_bfloat16* p = op + g8 * X.cgs;
        *reinterpret_cast<uint4*>(p) = h;
        *reinterpret_cast<uint4*>(p + X.lo_out) = l;
    }
    }
}

// one MMA with the A operand from shared memory (descriptor low word) or, TS, from tensor memory (column address)
template <bool TS>
__device__ __forceinline__ void mma_a(uint32_t t, uint32_t a, uint32_t dh, uint32_t bl, uint32_t idesc) {
    if constexpr (TS) umma_bf16_ts(t, a, bl, dh, idesc);
    else umma_bf16_w(t, a, dh, bl, dh, idesc);
}

// Persistent kernel: CTA i walks tiles i, i+gridDim.x, ...  A tile is XT consecutive x rows of one
// (y line group, z tile, Cout slice).  The CTA marches along x: every input plane is staged ONCE per
// 16-channel chunk and feeds the (up to) three output rows that use it through the three x taps, so
// the L2->shared traffic per output row is (XT+2)/XT planes instead of 3.  Output rows live in R
// TMEM "row slots" used round-robin: a row is complete after the plane behind it, its slot is drained
// by the epilogue warps while the MMA thread works on the following planes, then re-zeroed and handed
// back.  The shared-memory ring runs across row and tile boundaries.  XT = 1 is the plain
// one-row-per-tile schedule (strided / transposed convs, layers whose accumulators fill TMEM).
template <int OM, bool SC, int RM, bool TS>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                             const __grid_constant__ CUtensorMap tmap2,
                                                             const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [nstage]
    uint64_t* empty = full + 8;                          // [nstage]
    uint64_t* acc_full = full + 16;                      // [nslot]
    uint64_t* acc_empty = full + 24;                     // [nslot]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 32);
    uint64_t* a_full = full + 34;                        // [na]  TS mode: A views of a stage are in tensor memory
    uint64_t* a_empty = full + 38;                       // [na]  ... and have been consumed by every issuer
    // per-channel epilogue constants of this CTA's Cout slice(s), staged once: [scale | shift | bias2 | res_w | res_b][256]
    float* ep_c = reinterpret_cast<float*>(smem + 1024);
    const uint32_t ring = smem_u32(smem) + TC_HDR;

    if (a.pdl) asm volatile("griddepcontrol.launch_dependents;");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_cols = (uint32_t)(a.nacc * a.n_cta);
    const uint32_t slot_cols = row_cols * (a.nchunk2 ? 2 : 1);
    const int R = a.nbuf;
    const int njm1 = a.nj - 1, jc = a.nj >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstage; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, TC_MMA_WARPS + (TS ? 1 : 0));   // TS: the copier's reads of the stage count too
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(a_full + i, 1);
            mbar_init(a_empty + i, TC_MMA_WARPS);
        }
        for (int i = 0; i < TC_MAX_SLOT; ++i) {
            mbar_init(acc_full + i, TC_MMA_WARPS);
            mbar_init(acc_empty + i, TC_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0 && !a.line_mode) {
        prefetch_tmap(&tmap);
        if (a.nchunk2) prefetch_tmap(&tmap2);
    }
    if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
    {
        const int ncp = (a.out_mode >= 1 ? 16 : (a.cout + 15) / 16 * 16);   // all n-slices (<= 256 channels)
        for (int i = threadIdx.x; i < ncp; i += TC_THREADS) {
            ep_c[i] = a.ep.scale[i];
            ep_c[256 + i] = a.ep.shift[i];
            ep_c[512 + i] = a.nchunk2 ? a.bias2[i] : (a.res_mode == 2 ? a.res_b[i] : 0.f);   // added after the activation
            ep_c[768 + i] = a.res_mode == 2 ? a.res_w[i] : 0.f;
        }
    }
    if (a.line_mode && a.hz > 0) {
        // z-halo rows at the volume border are never written when the z tile spans the whole volume:
        // zero the ring once (with several z tiles the producer writes them from g_zero_line instead)
        uint4* q = reinterpret_cast<uint4*>(smem + TC_HDR);
        const int n16 = (int)(a.nstage * a.stage_bytes / 16);
        for (int i = threadIdx.x; i < n16; i += TC_THREADS) q[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above touched only constants of the plan (weights and epilogue tables are never written by a launch);
    // from here on the kernel reads what its predecessor wrote
    if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0 && a.line_mode) {
        // ===== bulk-copy producer: every lane issues whole z lines (2 KB contiguous in act8) =====
        int it = 0;
        long long t_wait = 0, t_beg = clock64();
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const TcTile T = decode_tile(a, tile);
            for (int q = T.q_lo; q <= T.q_hi; ++q) {
                const int rs = q - jc;
                const int n2 = (a.nchunk2 && !a.sc_self && rs >= 0 && rs < T.xt) ? a.nchunk2 : 0;
                for (int cs = 0; cs < a.nchunk + n2; ++cs) {
                    const bool seg2 = cs >= a.nchunk;
                    const int c = seg2 ? cs - a.nchunk : cs;
                    const int x = seg2 ? T.mx + rs : T.mx * a.sx + q + a.xoff;
                    const int st = it % a.nstage;
                    const long long t0 = a.dbg ? clock64() : 0;
                    mbar_wait_relaxed(empty + st, ((it / a.nstage) & 1) ^ 1);
                    if (a.dbg) t_wait += clock64() - t0;
                    const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
                    const vsseg_act8& src = seg2 ? a.in2 : a.in;
                    // staged lines i in [i0, i1): input y = ybase + i (zero line when outside); rows z in [zlo, zhi)
                    const int sy = seg2 ? 1 : a.sy;
                    const int ybase = T.my0 * sy - a.hy;
                    const int i0 = seg2 ? a.hy : 0, i1 = seg2 ? a.hy + a.YL : a.BY;
                    const int zlo = max(T.mz0 - a.hz, 0), zhi = min(T.mz0 + a.LZ + a.hz, a.Zin);
                    const uint32_t row_bytes = (uint32_t)(zhi - zlo) * 16;
                    const int ncopy = (i1 - i0) * 4;
                    // with several z tiles the border halo rows must be rewritten as zeros
                    const bool zl = a.ntz > 1 && a.hz > 0 && T.mz0 - a.hz < 0, zh = a.ntz > 1 && a.hz > 0 && T.mz0 + a.LZ + a.hz > a.Zin;
                    if (lane == 0) {
                        // XT > 1: the plane feeds all x taps, so the stage carries the weights of every tap
                        const uint32_t bb = seg2 ? a.b2_bytes : (a.XT > 1 ? a.nj * a.b_bytes : a.b_bytes);
                        const bool sc_here = a.sc_self && rs >= 0 && rs < T.xt;   // centre x tap of row rs: + shortcut weights
                        mbar_expect_tx(full + st, (uint32_t)ncopy * (row_bytes + (zl ? 16u : 0u) + (zh ? 16u : 0u)) + bb +
                                                      (sc_here ? a.b2_bytes : 0u));
                        const uint8_t* wsrc = seg2 ? a.w2 + ((size_t)T.ns * a.nchunk2 + c) * a.b2_bytes
                                                   : a.w + ((size_t)(T.sel * a.nchunk + c) * a.nj + (a.XT > 1 ? 0 : q)) * a.b_bytes;
                        bulk_load(base + a.b_off, wsrc, bb, full + st);
                        if (sc_here) bulk_load(base + a.b2_off, a.w2 + ((size_t)T.ns * a.nchunk2 + c) * a.b2_bytes, a.b2_bytes, full + st);
                    }
                    const __nv_bfloat16* g0 = (const __nv_bfloat16*)src.hi + (int64_t)T.b * src.batch_stride;
                    for (int k = lane; k < ncopy; k += 32) {
                        const int i = i0 + (k >> 2), plane = (k >> 1) & 1, cg = k & 1;
                        const int y = ybase + i;
                        const void* gp = (y >= 0 && y < a.Yin)
                                             ? (const void*)(g0 + (int64_t)plane * src.lo_offset +
                                                             ((((int64_t)(2 * c + cg) * a.Xin + x) * a.Yin + y) * a.Zin + zlo) * 8)
                                             : (const void*)g_zero_line;
                        const uint32_t line = base + (uint32_t)plane * a.a_plane + (uint32_t)cg * a.lbo_a + (uint32_t)(i * a.pitch) * 16;
                        bulk_load(line + (uint32_t)(zlo - (T.mz0 - a.hz)) * 16, gp, row_bytes, full + st);
                        if (zl) bulk_load(line, g_zero_line, 16, full + st);
                        if (zh) bulk_load(line + (uint32_t)(a.pitch - 1) * 16, g_zero_line, 16, full + st);
                    }
                    __syncwarp();
                    ++it;
                }
            }
        }
        if (a.dbg && lane == 0) {
            a.dbg[blockIdx.x * 8 + 3] = (unsigned long long)t_wait;
            a.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_beg);
        }
    } else if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            int it = 0;
            long long t_wait = 0, t_beg = clock64();
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const TcTile T = decode_tile(a, tile);
                for (int q = T.q_lo; q <= T.q_hi; ++q) {
                    const int rs = q - jc;
                    const int n2 = (a.nchunk2 && !a.sc_self && rs >= 0 && rs < T.xt) ? a.nchunk2 : 0;
                    for (int cs = 0; cs < a.nchunk + n2; ++cs) {
                        const bool seg2 = cs >= a.nchunk;
                        const int c = seg2 ? cs - a.nchunk : cs;
                        const int x = seg2 ? T.mx + rs : T.mx * a.sx + q + a.xoff;
                        const int st = it % a.nstage;
                        const long long t0 = a.dbg ? clock64() : 0;
                        mbar_wait_relaxed(empty + st, ((it / a.nstage) & 1) ^ 1);
                        if (a.dbg) t_wait += clock64() - t0;
                        const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
                        const CUtensorMap* map = seg2 ? &tmap2 : &tmap;
                        const int cgp = seg2 ? a.cg_plane2 : a.cg_plane;
                        const int cgi = T.b * (seg2 ? a.cg_batch2 : a.cg_batch) + c * 2;
                        if (a.dbgf & 2) {
                            mbar_arrive(full + st);
                        } else if (!seg2) {
                            const uint32_t bb = a.XT > 1 ? a.nj * a.b_bytes : a.b_bytes;
                            const bool sc_here = a.sc_self && rs >= 0 && rs < T.xt;   // centre x tap of row rs: + shortcut weights
                            mbar_expect_tx(full + st, 2 * a.nbox * a.box_tx + bb + (sc_here ? a.b2_bytes : 0u));
                            if (sc_here) bulk_load(base + a.b2_off, a.w2 + ((size_t)T.ns * a.nchunk2 + c) * a.b2_bytes, a.b2_bytes, full + st);
                            for (int i = 0; i < a.nbox; ++i) {
                                const uint32_t dst = base + (uint32_t)a.boxes[i].dst16 * 16;
                                const int zc = T.mz0 * a.sz + a.boxes[i].dz, yc = T.my0 * a.sy + a.boxes[i].dy;
                                tma_box(a.map_wide, dst, map, full + st, zc, yc, x, cgi);
                                tma_box(a.map_wide, dst + a.a_plane, map, full + st, zc, yc, x, cgi + cgp);
                            }
                            bulk_load(base + a.b_off, a.w + ((size_t)(T.sel * a.nchunk + c) * a.nj + (a.XT > 1 ? 0 : q)) * a.b_bytes, bb,
                                      full + st);
                        } else {
                            // shortcut source: same box shape, unshifted in z, 1x1x1 weights
                            mbar_expect_tx(full + st, 2 * a.box_tx + a.b2_bytes);
                            const int zc = T.mz0 + a.dz2, yc = T.my0 + a.dy2;
                            tma_box(a.map_wide, base, map, full + st, zc, yc, x, cgi);
                            tma_box(a.map_wide, base + a.a_plane, map, full + st, zc, yc, x, cgi + cgp);
                            bulk_load(base + a.b_off, a.w2 + ((size_t)T.ns * a.nchunk2 + c) * a.b2_bytes, a.b2_bytes, full + st);
                        }
                        ++it;
                    }
                }
            }
            if (a.dbg) {
                a.dbg[blockIdx.x * 8 + 3] = (unsigned long long)t_wait;
                a.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_beg);
            }
        }
    } else if (warp < TC_EPI_WARP0) {
        const int mw = warp - 1;   // this issuer owns the rows g with g % TC_MMA_WARPS == mw
        if (elect_one()) {
            // ===== MMA issuer (one elected lane; every MMA accumulates into TMEM zeroed by the epilogue warps) =====
            int it = 0, rowbase = 0;
            long long t_full = 0, t_acc = 0, n_mma = 0, t_beg = clock64();
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const TcTile T = decode_tile(a, tile);
                int opened = 0;
                for (int q = T.q_lo; q <= T.q_hi; ++q) {
                    // rows 0..min(q, xt-1) receive MMAs from this plane on: their slots must be drained and re-zeroed
                    const int need = min(q, T.xt - 1);
                    if (opened <= need) {
                        const long long t0 = a.dbg ? clock64() : 0;
                        for (; opened <= need; ++opened) {
                            const int g = rowbase + opened;
                            mbar_wait(acc_empty + g % R, (g / R) & 1);
                        }
                        tc_fence_after();
                        if (a.dbg) t_acc += clock64() - t0;
                    }
                    const int rs = q - jc;
                    const int n2 = (a.nchunk2 && !a.sc_self && rs >= 0 && rs < T.xt) ? a.nchunk2 : 0;
                    for (int cs = 0; cs < a.nchunk + n2; ++cs) {
                        const bool seg2 = cs >= a.nchunk;
                        const int st = it % a.nstage;
                        const long long t0 = a.dbg ? clock64() : 0;
                        mbar_wait(full + st, (it / a.nstage) & 1);
                        const int ab = TS ? it % a.na : 0;
                        if constexpr (TS) mbar_wait(a_full + ab, (it / a.na) & 1);   // the stage's A views are in tensor memory
                        tc_fence_after();
                        if (a.dbg) { t_full += clock64() - t0; n_mma += seg2 ? a.nop2 : a.nop * (a.XT == 1 ? 1 : min(q, T.xt - 1) - max(q - njm1, 0) + 1); }   // all issuers' MMAs
                        const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
                        // A: stage start in 16 B units (descriptor low word) | TS: first column of the stage's A buffer
                        const uint32_t da = TS ? tmem_base + a.a_tmem_col + (uint32_t)ab * a.a_cols : (base & 0x3FFFF) >> 4;
                        const uint32_t db = ((base + a.b_off) & 0x3FFFF) >> 4;
                        const uint32_t dh = a.desc_hi;
                        const bool tp = a.two_pass != 0;   // (A_hi, [W_hi | W_lo]) and (A_lo, [W_hi | 0]) instead of three passes
                        if (!seg2) {
                            const uint32_t ap = TS ? (uint32_t)a.BY * 8u : a.a_plane >> 4, bp = a.b_plane >> 4;
                            const int nel = a.nop / 3;
                            if (a.XT == 1) {
                                const uint32_t tr = tmem_base + (uint32_t)(rowbase % R) * slot_cols;
#pragma unroll 2
                                for (int i = rowbase % TC_MMA_WARPS == mw ? 0 : nel; i < nel; ++i) {
                                    const TcEl e = a.el[i];
                                    const uint32_t al = e.a_lo + da, bl = e.b_lo + db, t = tr + e.col;
                                    mma_a<TS>(t, al, dh, bl, e.idesc);
                                    if (tp) {
                                        mma_a<TS>(t, al + ap, dh, bl + bp, e.idesc);
                                    } else {
                                        mma_a<TS>(t, al + ap, dh, bl, e.idesc);
                                        mma_a<TS>(t, al, dh, bl + bp, e.idesc);
                                    }
                                }
                            } else {
                                // plane q of the haloed tile feeds output row r = q - dx through x tap dx
                                for (int dx = 0; dx < a.nj; ++dx) {
                                    const int r = q - dx;
                                    if (r < 0 || r >= T.xt || (rowbase + r) % TC_MMA_WARPS != mw) continue;
                                    const uint32_t tr = tmem_base + (uint32_t)((rowbase + r) % R) * slot_cols;
                                    const uint32_t dbx = db + (uint32_t)dx * (a.b_bytes >> 4);
#pragma unroll 2
                                    for (int i = 0; i < nel; ++i) {
                                        const TcEl e = a.el[i];
                                        const uint32_t al = e.a_lo + da, bl = e.b_lo + dbx, t = tr + e.col;
                                        mma_a<TS>(t, al, dh, bl, e.idesc);
                                        if (tp) {
                                            mma_a<TS>(t, al + ap, dh, bl + bp, e.idesc);
                                        } else {
                                            mma_a<TS>(t, al + ap, dh, bl, e.idesc);
                                            mma_a<TS>(t, al, dh, bl + bp, e.idesc);
                                        }
                                    }
                                }
                            }
                        } else {
                            const uint32_t ap = TS ? (uint32_t)a.BY * 8u : a.a_plane >> 4, bp = a.b2_plane >> 4;
                            const uint32_t tr = tmem_base + (uint32_t)((rowbase + rs) % R) * slot_cols + row_cols;
                            for (int i = (rowbase + rs) % TC_MMA_WARPS == mw ? 0 : a.nop2 / 3; i < a.nop2 / 3; ++i) {
                                const TcEl e = a.el2[i];
                                const uint32_t al = e.a_lo + da, bl = e.b_lo + db, t = tr + e.col;
                                mma_a<TS>(t, al, dh, bl, e.idesc);
                                mma_a<TS>(t, al + ap, dh, bl, e.idesc);
                                mma_a<TS>(t, al, dh, bl + bp, e.idesc);
                            }
                        }
                        if (a.sc_self && rs >= 0 && rs < T.xt && (rowbase + rs) % TC_MMA_WARPS == mw) {
                            // this stage is the centre x tap of row rs: its centre (y, z) views are the shortcut's A operand
                            const uint32_t ap = a.a_plane >> 4, bp = a.b2_plane >> 4;
                            const uint32_t tr = tmem_base + (uint32_t)((rowbase + rs) % R) * slot_cols + row_cols;
                            const uint32_t db2 = ((base + a.b2_off) & 0x3FFFF) >> 4;
                            for (int i = 0; i < a.nop2 / 3; ++i) {
                                const TcEl e = a.el2[i];
                                const uint32_t al = e.a_lo + da, bl = e.b_lo + db2, t = tr + e.col;
                                mma_a<TS>(t, al, dh, bl, e.idesc);
                                mma_a<TS>(t, al + ap, dh, bl, e.idesc);
                                mma_a<TS>(t, al, dh, bl + bp, e.idesc);
                            }
                        }
                        umma_commit(empty + st);
                        if constexpr (TS) umma_commit(a_empty + ab);
                        ++it;
                    }
                    // rows whose last plane this was are complete
                    const int r0 = max(q - njm1, 0), r1 = q == T.q_hi ? T.xt - 1 : q - njm1;
                    for (int r = r0; r <= r1; ++r) umma_commit(acc_full + (rowbase + r) % R);
                }
                rowbase += T.xt;
            }
            if (a.dbg && mw == 0) {
                a.dbg[blockIdx.x * 8 + 0] = (unsigned long long)t_full;
                a.dbg[blockIdx.x * 8 + 1] = (unsigned long long)t_acc;
                a.dbg[blockIdx.x * 8 + 2] = (unsigned long long)(clock64() - t_beg);
                a.dbg[blockIdx.x * 8 + 7] = (unsigned long long)n_mma;
            }
        }
    } else if (warp == TC_CP_WARP) {
        // ===== TS mode: copy the staged A views (BY lines x hi/lo planes, 128 x 16 bf16 each) from shared memory to the
        // stage's A buffer in tensor memory.  One copy per view and staged plane replaces one 4 KB shared-memory read per
        // MMA (an x-march plane feeds up to three rows x three passes per view).  Same walk as the producer / issuers.
        if constexpr (TS) {
          if (elect_one()) {
            int it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const TcTile T = decode_tile(a, tile);
                for (int q = T.q_lo; q <= T.q_hi; ++q) {
                    const int rs = q - jc;
                    const int n2 = (a.nchunk2 && !a.sc_self && rs >= 0 && rs < T.xt) ? a.nchunk2 : 0;
                    for (int cs = 0; cs < a.nchunk + n2; ++cs) {
                        const bool seg2 = cs >= a.nchunk;
                        const int st = it % a.nstage, ab = it % a.na;
                        mbar_wait(full + st, (it / a.nstage) & 1);
                        mbar_wait(a_empty + ab, ((it / a.na) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
                        const uint32_t tdst = tmem_base + a.a_tmem_col + (uint32_t)ab * a.a_cols;
                        const uint32_t lbo = ((a.lbo_a >> 4) & 0x3FFFu) << 16;
                        const int v0 = seg2 ? a.hy : 0, v1 = seg2 ? a.hy + a.YL : a.BY;   // the shortcut reads the tile's own lines
                        for (int v = v0; v < v1; ++v) {
                            const uint32_t sa = base + (uint32_t)v * 2048u;
                            tmem_cp_128x256b(tdst + (uint32_t)v * 8u, (((sa) & 0x3FFFF) >> 4) | lbo, a.desc_hi);
                            tmem_cp_128x256b(tdst + (uint32_t)(a.BY + v) * 8u, (((sa + a.a_plane) & 0x3FFFF) >> 4) | lbo, a.desc_hi);
                        }
                        umma_commit(a_full + ab);
                        umma_commit(empty + st);
                        ++it;
                    }
                }
            }
          }
        }
    } else {
        // ===== epilogue: TC_EPI_WARPS warps, TMEM lanes (warp % 4) * 32 .. +31, thread = one M-tile row;
        // the warps sharing a lane quadrant take alternate (accumulator, 16-column) units of every row slot =====
        const int lane_base = (warp & 3) * 32;
        const int esub = (warp - TC_EPI_WARP0) >> 2, nsub = TC_EPI_WARPS / 4;
        const uint32_t tlane = tmem_base + ((uint32_t)lane_base << 16);
        const int n16 = a.n_cta >> 4;
        constexpr bool sc = SC;
        // zero this warp's lanes of the units it will drain in every slot, then release the MMA issuer
        for (int sl = 0; sl < R; ++sl) {
            for (int u = esub; u < a.nacc * n16; u += nsub) {
                tmem_st16_zero(tlane + (uint32_t)sl * slot_cols + (uint32_t)(u * 16));
                if (sc) tmem_st16_zero(tlane + (uint32_t)sl * slot_cols + row_cols + (uint32_t)(u * 16));
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + sl);
        }
        const int rr_ = lane_base + lane;
        const int ly = rr_ / a.LZ, zz = rr_ % a.LZ;
        const int Xo = a.out.X, Yo = a.out.Y, Zo = a.out.Z;
        const int64_t cgs = (int64_t)Xo * Yo * Zo * 8;   // elements per 8-channel group
        const int thr_off = (ly * a.uy * Zo + zz * a.uz) * 8;   // this thread's row of the M tile inside the output tile
        const bool sig = a.ep.act == 1;
        const float slope = a.ep.slope;
        const int64_t lo_out = a.out.lo_offset, lo_res = a.res.lo_offset;
        int rowbase = 0;
        const float* const rsrc_p = RM == 2 ? f32_base(a.rsrc) : nullptr;
        float* const outf_p = OM >= 1 ? f32_base(a.outf) : nullptr;
        long long t_wait = 0, t_beg = clock64();
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const TcTile T = decode_tile(a, tile);
            const int b = T.b, my0 = T.my0, mz0 = T.mz0;
            const int co0 = T.ns * a.n_cta;
            const int nreal = min(a.n_cta, a.cout - co0);
            // element offset (without batch and x) of this thread's voxel in accumulator 0, channel group co0/8
            const int64_t tile_off = ((int64_t)(my0 * a.uy) * Zo + mz0 * a.uz) * 8 + thr_off + (int64_t)(co0 >> 3) * cgs;
            __nv_bfloat16* const out_b = (__nv_bfloat16*)a.out.hi + (int64_t)b * a.out.batch_stride;
            const __nv_bfloat16* const res_b = (const __nv_bfloat16*)a.res.hi + (int64_t)b * a.res.batch_stride;
            const float* const rsrc_b = RM == 2 ? rsrc_p + win_off(a.rsrc, a.rwin, b) : nullptr;   // window b of a window set
            for (int r = 0; r < T.xt; ++r) {
                const int g = rowbase + r, slot = g % R;
                const uint32_t tacc = tlane + (uint32_t)slot * slot_cols;
                const int ox = (T.mx + r) * a.ux + T.px;
                const int64_t row_off = tile_off + (int64_t)ox * Yo * Zo * 8;
                const long long t0 = a.dbg ? clock64() : 0;
                mbar_wait_relaxed(acc_full + slot, (g / R) & 1);
                tc_fence_after();
                if (a.dbg) t_wait += clock64() - t0;
                // units = (accumulator ai, 16-column chunk c), dealt to the quadrant's warps like the zeroing below
                // (two units in flight per thread were tried: the extra registers spill and it is slower)
                EpiCtx X{a, ep_c, tacc, row_cols, row_off, cgs, lo_out, lo_res, out_b, res_b, b, ox, my0, mz0, ly, zz, co0, nreal, sc, sig, slope, rsrc_b, outf_p};
                int ai = 0, c = esub;
                while (c >= n16) { c -= n16; ++ai; }
                while (ai < a.nacc) {
                    EpiUnit<SC, RM> U;
                    U.ai = ai; U.c = c;
                    c += nsub;
                    while (c >= n16) { c -= n16; ++ai; }
                    epi_load<OM, SC, RM>(X, U);
                    tmem_ld_wait();
                    epi_finish<OM, SC, RM>(X, U);
                }
                // re-zero the drained units of the slot and hand it back to the MMA issuer
                for (int u = esub; u < a.nacc * n16; u += nsub) {
                    tmem_st16_zero(tacc + (uint32_t)(u * 16));
                    if (sc) tmem_st16_zero(tacc + row_cols + (uint32_t)(u * 16));
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + slot);
            }
            rowbase += T.xt;
        }
        if (a.dbg && warp == TC_EPI_WARP0 && lane == 0) {
            a.dbg[blockIdx.x * 8 + 5] = (unsigned long long)t_wait;
            a.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - t_beg);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---- host side: geometry -> tables ------------------------------------------------------------
struct TcPlan {
    TcArgs a;
    cuuint32_t box[5], estr[5];
    size_t smem;
    unsigned grid;
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200; also the answer on the GPU-less build box
        cudaGetLastError();
    }
    return sms;
}

struct TcGeom {   // what gen_ops needs
    bool tr, strided, line;
    int KY, KZ, LY, LZ, BZ, npz, n_cta, sy;
    uint32_t box_bytes, a_plane, b_plane;
};

// Elementary products (A view, weight tap, accumulator) of one stage, then merged: products that
// read the SAME A view and write ADJACENT accumulators with ADJACENT weight rows become one
// tcgen05.mma with a wider N.  In SS mode an MMA costs max(N/2, 32 + N/4) cycles (the 4 KB A read
// from shared memory is the floor for N < 128, measured with tools/ubench/mma_rate.cu), so sharing the
// A read between the (up to) three y taps that use it is worth up to 2.7x for the narrow layers.
static bool gen_ops(const TcGeom& G, int YT, TcOp* ops, int* nop_out, TcAcc* accs, int* nacc_out) {
    struct El { uint32_t a16; int tz, brow, col, n; };
    El el[TC_MAX_ACC * 27];
    int ne = 0, nacc = 0;
    const int n = G.n_cta;
    if (!G.tr) {
        for (int y = 0; y < YT; ++y) {
            accs[nacc].y_add = (int16_t)(y * G.LY);
            accs[nacc].z_add = 0;
            accs[nacc].yl = (int8_t)(y * G.LY);
            for (int dy = 0; dy < G.KY; ++dy)
                for (int dz = 0; dz < G.KZ; ++dz) {
                    uint32_t aoff;
                    if (G.line) aoff = (uint32_t)(((y * G.sy + dy) * G.BZ + dz) * 16);
                    else if (G.strided) aoff = (uint32_t)(dy * G.KZ + dz) * G.box_bytes + (uint32_t)y * 2048;
                    else aoff = (uint32_t)dz * G.box_bytes + (uint32_t)((y * G.LY + dy) * G.LZ * 16);
                    el[ne++] = {aoff / 16, dz, (G.KY - 1 - dy) * n, nacc * n, n};  // weight rows in reversed y order
                }
            ++nacc;
        }
    } else {
        // phase p along an axis: p=0 -> (shift 0, tap 1); p=1 -> (shift 0, tap 2), (shift 1, tap 0).
        // accumulator index = pz*(2*YT) + 2*y + py, so the phases fed by one A view are adjacent.
        for (int pz = 0; pz < G.npz; ++pz)
            for (int y = 0; y < YT; ++y)
                for (int py = 0; py < 2; ++py) {
                    const int ai = pz * 2 * YT + 2 * y + py;
                    accs[ai].y_add = (int16_t)(y * G.LY * 2 + py);
                    accs[ai].z_add = (int8_t)pz;
                    accs[ai].yl = (int8_t)(y * G.LY);
                    for (int ddy = 0; ddy <= py; ++ddy)
                        for (int ddz = 0; ddz <= pz; ++ddz) {
                            const int ky = py == 0 ? 1 : (ddy == 0 ? 2 : 0);
                            const int kz = G.npz == 1 ? 0 : (pz == 0 ? 1 : (ddz == 0 ? 2 : 0));
                            const uint32_t aoff = G.line ? (uint32_t)((y + ddy) * G.BZ * 16)
                                                         : (uint32_t)ddz * G.box_bytes + (uint32_t)((y * G.LY + ddy) * G.LZ * 16);
                            el[ne++] = {aoff / 16, kz, ky * n, ai * n, n};
                        }
                    ++nacc;
                }
    }
    // merge: sort by (A view, weight z tap, column), then join runs
    for (int i = 1; i < ne; ++i) {
        El k = el[i];
        int j = i - 1;
        auto less = [](const El& p, const El& q) {
            if (p.a16 != q.a16) return p.a16 < q.a16;
            if (p.tz != q.tz) return p.tz < q.tz;
            return p.col < q.col;
        };
        while (j >= 0 && less(k, el[j])) { el[j + 1] = el[j]; --j; }
        el[j + 1] = k;
    }
    int nm = 0;
    for (int i = 0; i < ne; ++i) {
        if (nm && el[nm - 1].a16 == el[i].a16 && el[nm - 1].tz == el[i].tz && el[nm - 1].col + el[nm - 1].n == el[i].col &&
            el[nm - 1].brow + el[nm - 1].n == el[i].brow && el[nm - 1].n + el[i].n <= 256)
            el[nm - 1].n += el[i].n;
        else
            el[nm++] = el[i];
    }
    if (nm * 3 > TC_MAX_OPS) return false;
    int nop = 0;
    for (int i = 0; i < nm; ++i) {
        const uint32_t b16 = (uint32_t)(el[i].tz * 2 * G.KY * n + el[i].brow);
        for (int pass = 0; pass < 3; ++pass) {
            TcOp& op = ops[nop++];
            op.a16 = (uint16_t)(el[i].a16 + (pass == 1 ? G.a_plane / 16 : 0));
            op.b16 = (uint16_t)(b16 + (pass == 2 ? G.b_plane / 16 : 0));
            op.col = (uint16_t)el[i].col;
            op.n8 = (uint16_t)(el[i].n / 8);
        }
    }
    *nop_out = nop;
    *nacc_out = nacc;
    return true;
}

// Measured tile choices (tools/autotune_tiles.py on a B200: every launch of a group of 8 windows of 128^3 re-timed with
// every feasible (XT, YT)) for the layers where the cost model below misses the fastest tile by more than the ~3 %
// run-to-run noise of the sweep (profiles/r02_autotune_tiles.tsv).  Keyed by the layer's geometry; anything else,
// and batches of fewer than 4 windows, goes through the cost model.
struct TileHint { int cin, cout_pad, Xm, Ym, Zm, k, s, tr, sc, xt, yt; };   // k, s: kx ky kz / sx sy sz as decimal digits
static const TileHint kTileHints[] = {
    {32, 32, 64, 64, 128, 331, 111, 0, 1, 1, 4},    // enc1.unit1 (+ fused shortcut): 0.399 -> 0.384 ms
    {64, 48, 16, 16, 64, 333, 222, 1, 0, 1, 1},     // up2 (transposed): 0.139 -> 0.113 ms
    {80, 64, 8, 8, 32, 333, 222, 1, 0, 1, 1},       // up3 (transposed): 0.034 -> 0.031 ms
    {64, 16, 16, 16, 64, 333, 111, 0, 0, 2, 2},     // dec3.att.conv2 (+ fused gate): 0.112 -> 0.080 ms
};

// Fills the plan for (in -> out, geometry, n_split); returns false when the shape is not covered
// (the caller then uses the generic CUDA-core kernel).
static bool make_plan_uncached(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, int n_split,
                               const vsseg_act8* src2, TcPlan* P) {
    if (!in || !out || !g || !in->hi || !out->hi) return false;
    if (in->C % 16 || out->C % 8 || in->B != out->B || n_split < 1) return false;
    const int KX = g->kx, KY = g->ky, KZ = g->kz;
    if ((KX != 1 && KX != 3) || (KY != 1 && KY != 3) || (KZ != 1 && KZ != 3)) return false;
    const bool tr = g->transposed != 0;
    const bool strided = !tr && (g->sx != 1 || g->sy != 1 || g->sz != 1);
    if (g->sx < 1 || g->sx > 2 || g->sy < 1 || g->sy > 2 || g->sz < 1 || g->sz > 2) return false;
    // M grid
    int Xm, Ym, Zm;
    if (tr) {
        // sub-pixel phases need k=3 on every stride-2 axis and k=1 on stride-1 axes; x and y strided
        if (g->sx != 2 || g->sy != 2 || KX != 3 || KY != 3) return false;
        if ((g->sz == 2) != (KZ == 3)) return false;
        if (out->X != in->X * 2 || out->Y != in->Y * 2 || out->Z != in->Z * g->sz) return false;
        Xm = in->X; Ym = in->Y; Zm = in->Z;
    } else {
        if (out->X != (in->X + g->sx - 1) / g->sx || out->Y != (in->Y + g->sy - 1) / g->sy ||
            out->Z != (in->Z + g->sz - 1) / g->sz)
            return false;
        if (in->X % g->sx || in->Y % g->sy || in->Z % g->sz) return false;
        Xm = out->X; Ym = out->Y; Zm = out->Z;
    }
    // z extent of the M tile: the largest power of two (<= 128) that divides the z extent of the M grid, so any
    // crop tiles (Z = 40 -> 8, 96 -> 32, 160 -> 32; the coarse levels of shallow crops run with LZ = 4, 2, 1 and
    // LY = 32..128 lines per tile, rows beyond the y extent masked)
    static const int lz_min = getenv("VSSEG_TC_LZ_MIN") ? atoi(getenv("VSSEG_TC_LZ_MIN")) : 1;
    int LZ = 128;
    while (LZ > 1 && Zm % LZ) LZ >>= 1;
    if (LZ < lz_min) return false;
    const int LY = 128 / LZ;
    const int64_t cgs = (int64_t)in->X * in->Y * in->Z * 8;
    if (in->lo_offset % cgs || in->batch_stride % cgs) return false;
    if (src2) {
        if (tr || strided) return false;
        if (src2->C % 16 || src2->X != in->X || src2->Y != in->Y || src2->Z != in->Z || src2->B != in->B) return false;
        if (src2->lo_offset % cgs || src2->batch_stride % cgs) return false;
    }
    const int cout_pad = round_up(out->C, 16);
    if (cout_pad % n_split || (cout_pad / n_split) % 16) return false;
    const int n_cta = cout_pad / n_split;
    if (n_cta > 256) return false;

    TcArgs& a = P->a;
    memset(&a, 0, sizeof(a));
    const int hz = KZ == 3 ? 1 : 0, hy = KY == 3 ? 1 : 0;
    const int nphase = tr ? 2 * g->sz : 1;           // (py, pz) phases per CTA; px is a grid dimension
    const int acc_mult = nphase * (src2 ? 2 : 1);
    // the shortcut of a decoder ResidualUnit reads the conv's own input: its MMAs ride on the centre-tap stages
    static const bool sc_self_env = !(getenv("VSSEG_TC_SC_SELF") && atoi(getenv("VSSEG_TC_SC_SELF")) == 0);
    const bool sc_self = src2 && sc_self_env && src2->hi == in->hi && src2->C == in->C && src2->lo_offset == in->lo_offset &&
                         src2->batch_stride == in->batch_stride;
    const uint32_t b2_bytes = (uint32_t)(2 * n_cta * 32);
    // flavour
    // line mode: M tile = one z line; the stage holds BY whole input lines (pitch = 128 + 2*hz rows),
    // every tap is an address offset (line, dz) into it.  Needs unit z stride.
    const bool line = LY == 1 && g->sz == 1 && KZ == 3;   // without a z halo one wide TMA box moves whole lines
    const int sy_in = tr ? 1 : g->sy;
    // choose YT = number of y line groups per CTA
    const int ygroups = (Ym + LY - 1) / LY;   // a partial last group is zero-filled by TMA and masked in the epilogue
    int best = 0, best_xt = 1;
    double best_cost = 1e30;
    size_t best_stage = 0;
    int best_nstage = 0;
    const int taps_stage = KY * KZ;
    const uint32_t b_bytes = (uint32_t)(2 * taps_stage * n_cta * 32);
    // XT > 1: the CTA marches along x over XT rows and stages every input plane once for its (up to) three
    // rows; stride-1 convs with x taps only.  A stage then carries the weights of all three x taps.
    const bool xt_ok = !tr && !strided && KX == 3;
    static const int xt_max = getenv("VSSEG_TC_XT_MAX") ? atoi(getenv("VSSEG_TC_XT_MAX")) : 64;   // tuning knobs
    static const int xt_min = getenv("VSSEG_TC_XT_MIN") ? atoi(getenv("VSSEG_TC_XT_MIN")) : 1;
    static const int yt_max = getenv("VSSEG_TC_YT_MAX") ? atoi(getenv("VSSEG_TC_YT_MAX")) : 16;
    static const double fill_bpc = getenv("VSSEG_TC_FILL_BPC") ? atof(getenv("VSSEG_TC_FILL_BPC")) : 36.0;   // L2 -> shared bytes/cycle/SM
    static const double epi_unit = getenv("VSSEG_TC_EPI_UNIT") ? atof(getenv("VSSEG_TC_EPI_UNIT")) : 400.0;  // cycles per (accumulator, 16 columns) unit per warp
    int best_slots = 1;
    bool best_ts = false;
    static const int nstage_max = getenv("VSSEG_TC_NSTAGE_MAX") ? atoi(getenv("VSSEG_TC_NSTAGE_MAX")) : 6;   // <= 8 (barrier arrays)
    // TS mode: stride-1 convs without z taps on 128-long z lines (box mode, every A view is one whole z line)
    // Measured (tools/ubench/mma_ts.cu, profiles/r02_*): a tcgen05.mma M=128 K=16 has an execution floor of ~45 cycles
    // for N <= 90 in BOTH modes (SS: max(45, 32 + N/4); TS: max(45, N/2)), so moving A to tensor memory only pays for
    // N >= 96, while the two A buffers cost TMEM row slots - on this network TS is 5-40 % slower on every layer.
    // Kept as an experiment (default off): VSSEG_TC_TS = 0 off, 1 cost model, 2 forced where possible.
    static const int ts_env = getenv("VSSEG_TC_TS") ? atoi(getenv("VSSEG_TC_TS")) : 0;
    const bool ts_ok = ts_env != 0 && !tr && !strided && !line && KZ == 1 && LY == 1 && !sc_self;
    const int sms_ = sm_count();
    // VSSEG_TC_FORCE="XT,YT[,nstage]" (tools/autotune_tiles.py): only that tile is considered; read on every call
    int f_xt = 0, f_yt = 0, f_nst = 0;
    static const bool use_hints = !(getenv("VSSEG_TC_HINTS") && atoi(getenv("VSSEG_TC_HINTS")) == 0);
    bool hinted = false;
    if (const char* f = getenv("VSSEG_TC_FORCE")) {
        sscanf(f, "%d,%d,%d", &f_xt, &f_yt, &f_nst);
    } else if (use_hints && in->B >= 4) {
        const int kk = KX * 100 + KY * 10 + KZ, ss = g->sx * 100 + g->sy * 10 + g->sz;
        for (const TileHint& h : kTileHints)
            if (h.cin == in->C && h.cout_pad == cout_pad && h.Xm == Xm && h.Ym == Ym && h.Zm == Zm && h.k == kk && h.s == ss &&
                h.tr == (tr ? 1 : 0) && h.sc == (src2 ? 1 : 0) && n_split == 1) {
                f_xt = h.xt; f_yt = h.yt; hinted = true;
            }
    }
    // two attempts at most: with the hinted tile, and - when that tile does not fit this variant of the layer -
    // with the cost model alone
    for (int attempt = 0; attempt < 2 && !best; ++attempt) {
    if (attempt == 1) {
        if (!hinted) break;
        f_xt = f_yt = 0;
    }
    for (int ts = 0; ts <= (ts_ok ? 1 : 0); ++ts)
    for (int XT = 1; XT <= (xt_ok ? (xt_max < Xm ? xt_max : Xm) : 1); ++XT) {
        if (f_xt && XT != f_xt) continue;
        if (XT > 1 && XT < xt_min && xt_min <= Xm) continue;
        const int nseg = (Xm + XT - 1) / XT;
        if (XT > 1 && (Xm + nseg - 1) / nseg != XT) continue;   // keep the segments balanced
        const size_t bstage = round_up((int)((XT > 1 ? KX : 1) * b_bytes + (sc_self ? b2_bytes : 0)), 128);
        const long total_tiles_1 = (long)in->B * nseg * (Zm / LZ) * (tr ? 2 : 1) * n_split;
        for (int YT = 1; YT <= ygroups && YT <= yt_max; ++YT) {
            if (ygroups % YT) continue;
            if (f_yt && YT != f_yt) continue;
            const int cols1 = YT * acc_mult * n_cta;   // TMEM columns of one row slot
            if (cols1 > 512) break;
            if (YT * acc_mult > TC_MAX_ACC) break;
            // row slots: XT = 1 double-buffers whole tiles when two fit; the x march needs the three
            // rows a plane feeds plus (ideally) one being drained.  TS: two A buffers of BY views x 2 planes x 8
            // columns come out of the same 512 columns
            const int by_ts = (YT - 1) * sy_in + KY;
            const int acc_cols = ts ? 512 - 2 * 16 * by_ts : 512;
            if (acc_cols < cols1) break;
            int slots = acc_cols / cols1;
            if (XT == 1) slots = slots >= 2 ? 2 : 1;
            else if (slots < 3) continue;
            else if (slots > 6) slots = 6;
            int nbox, BY, BZ;
            if (line) { nbox = 1; BY = tr ? YT + 1 : (YT - 1) * sy_in + KY; BZ = LZ + 2 * hz; }
            else if (tr) { nbox = g->sz; BY = YT * LY + 1; BZ = LZ; }
            else if (strided) { nbox = KY * KZ; BY = YT * LY; BZ = LZ; }
            else { nbox = KZ; BY = YT * LY + 2 * hy; BZ = LZ; }
            if (!line && (BY * (strided ? g->sy : 1) > 256 || BZ * (strided ? g->sz : 1) > 256)) break;
            // every TMA box must land 128-byte aligned: boxes of tiny-LZ tiles are spaced at a padded stride
            const size_t box_bytes = (size_t)round_up(2 * BY * BZ * 16, 128);
            const size_t a_plane = nbox * box_bytes;
            double mma_cyc = 0;
            {
                TcGeom G{tr, strided, line, KY, KZ, LY, LZ, BZ, (int)g->sz, n_cta, sy_in, (uint32_t)box_bytes,
                         (uint32_t)a_plane, b_bytes / 2};
                static TcOp scratch[TC_MAX_OPS];
                static TcAcc scratch_acc[TC_MAX_ACC];
                int n1 = 0, n2 = 0;
                if (!gen_ops(G, YT, scratch, &n1, scratch_acc, &n2) || (src2 && YT * 3 > TC_MAX_OPS2)) break;
                double b_rd = 0;
                for (int i = 0; i < n1; ++i) {   // SS-mode MMA cost, measured (tools/ubench/mma_rate.cu)
                    const double N = scratch[i].n8 * 8.0;
                    if (ts) { mma_cyc += N / 2 > 16 ? N / 2 : 16; b_rd += N * 32; }   // TS: N/2 cycles, only B comes from shared memory
                    else mma_cyc += N / 2 > 32 + N / 4 ? N / 2 : 32 + N / 4;
                }
                if (ts && b_rd / 128 > mma_cyc) mma_cyc = b_rd / 128;
            }
            // TS: the copies read every staged view once more from shared memory (128 B/cycle)
            const double cp_cyc = ts ? (double)(2 * a_plane) / 128.0 : 0.0;
            const size_t stage = 2 * a_plane + bstage;
            if (stage / 16 >= 16000) break;
            const long budget = 227L * 1024 - TC_HDR;
            int nst = (int)(budget / (long)stage);
            if (nst < 2) break;
            // cost model (persistent CTAs, one per SM): waves x tile time.  Main loop: per staged plane-chunk
            // max(MMA issue, shared-memory fill); epilogue: (accumulator, 16-column) units spread over the four
            // warps of a TMEM lane quadrant; the two overlap when a row slot is free while the next fills
            const long tiles = total_tiles_1 * (ygroups / YT);
            const int nch = in->C / 16;
            // a stage can only be refilled after its MMAs have drained: with a 2-deep ring the copy latency
            // (~1500 cycles issue-to-arrival) is exposed on every stage, with 3+ it hides behind the other stages
            static const double fill_lat = getenv("VSSEG_TC_FILL_LAT") ? atof(getenv("VSSEG_TC_FILL_LAT")) : 0.0;   // measured: a latency term makes the choice worse overall
            const double fill_cyc = (double)stage / fill_bpc + (nst == 2 ? fill_lat : 0.0) + cp_cyc;
            double main_cyc;
            if (XT > 1) {
                const double per_plane = 3.0 * XT / (XT + 2) * mma_cyc;   // x taps served per staged plane, on average
                main_cyc = (double)(XT + 2) * nch * (per_plane > fill_cyc ? per_plane : fill_cyc);
            } else {
                const double nst_main = tr ? 1.5 * nch : (double)nch * KX;
                main_cyc = nst_main * (mma_cyc > fill_cyc ? mma_cyc : fill_cyc);
            }
            if (src2) {
                const double f2 = sc_self ? 0.0 : (double)(2 * a_plane + 2 * n_cta * 32) / fill_bpc, m2 = YT * 3.0 * (32 + n_cta / 4.0);
                main_cyc += (double)XT * (src2->C / 16) * (f2 > m2 ? f2 : m2);
            }
            const int units = YT * nphase * (n_cta / 16);
            const double epi_cyc = XT * (150.0 + ((units + 3) / 4) * epi_unit);
            const bool overlap = XT == 1 ? slots == 2 : slots >= 4;
            const double tile_cyc = overlap ? (main_cyc > epi_cyc ? main_cyc : epi_cyc) + 300.0
                                            : main_cyc + epi_cyc + (XT == 1 ? 1500.0 : 0.0);
            const double cost = (double)((tiles + sms_ - 1) / sms_) * tile_cyc + 4000.0;
            if (cost < best_cost || (ts_env == 2 && ts && !best_ts)) {
                best_cost = cost; best = YT; best_xt = XT; best_stage = stage; best_slots = slots;
                best_nstage = nst > nstage_max ? nstage_max : nst;
                if (f_nst >= 2 && f_nst < best_nstage) best_nstage = f_nst;
                best_ts = ts != 0;
            }
        }
    }
    }
    if (!best) return false;
    const int YT = best, XT = best_xt;
    int nbox, BY, BZ;
    if (line) { nbox = 1; BY = tr ? YT + 1 : (YT - 1) * sy_in + KY; BZ = LZ + 2 * hz; }
    else if (tr) { nbox = g->sz; BY = YT * LY + 1; BZ = LZ; }
    else if (strided) { nbox = KY * KZ; BY = YT * LY; BZ = LZ; }
    else { nbox = KZ; BY = YT * LY + 2 * hy; BZ = LZ; }
    const uint32_t box_tx = (uint32_t)(2 * BY * BZ * 16);           // bytes one box moves
    const uint32_t box_bytes = (uint32_t)round_up((int)box_tx, 128);   // distance between boxes of a stage
    a.out = *out;
    a.nchunk = in->C / 16;
    a.nj = tr ? 2 : KX;   // transposed: px=0 CTAs use j=0 only (see nj_eff below)
    a.nchunk2 = src2 ? src2->C / 16 : 0;
    a.nbox = nbox;
    a.nstage = best_nstage;
    a.a_plane = (uint32_t)(nbox * box_bytes);
    a.b_off = 2 * a.a_plane;
    a.b_bytes = b_bytes;
    a.b_plane = b_bytes / 2;
    a.b2_bytes = b2_bytes;
    a.b2_plane = a.b2_bytes / 2;
    a.sc_self = sc_self ? 1 : 0;
    a.b2_off = a.b_off + (uint32_t)((XT > 1 ? KX : 1) * b_bytes);
    a.stage_bytes = (uint32_t)best_stage;
    a.box_tx = box_tx;
    a.lbo_a = (uint32_t)(BY * BZ * 16);
    a.lbo_b = (uint32_t)(KY * n_cta * 16);   // B region: [tz][khalf][ty'][n][8]
    a.lbo_b2 = (uint32_t)(n_cta * 16);
    const TcGeom G{tr, strided, line, KY, KZ, LY, LZ, BZ, (int)g->sz, n_cta, sy_in, box_bytes, a.a_plane, a.b_plane};
    a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);   // N field (bits 17..22) comes from the op
    a.ntz = Zm / LZ; a.nty = ygroups / YT; a.ntx = (Xm + XT - 1) / XT;
    a.XT = XT; a.Xm = Xm;
    a.npx = tr ? 2 : 1; a.nsplit = n_split; a.nsel = a.npx * n_split;
    a.LZ = LZ; a.LY = LY; a.YL = YT * LY; a.Ym = Ym;
    a.n_cta = n_cta; a.cout = out->C;
    a.sx = tr ? 1 : g->sx; a.sy = tr ? 1 : g->sy; a.sz = tr ? 1 : g->sz;
    a.xoff = tr ? 0 : -(KX / 2);
    a.Xin = in->X;
    a.cg_plane = (int)(in->lo_offset / cgs);
    a.cg_batch = (int)(in->batch_stride / cgs);
    if (src2) {
        a.cg_plane2 = (int)(src2->lo_offset / cgs);
        a.cg_batch2 = (int)(src2->batch_stride / cgs);
    }
    a.ux = tr ? 2 : 1; a.uy = tr ? 2 : 1; a.uz = tr ? g->sz : 1;
    a.line_mode = line ? 1 : 0;
    a.map_wide = (!line && !(strided && g->sz == 2)) ? 1 : 0;
    a.BY = BY; a.pitch = BZ; a.hy = tr ? 0 : hy; a.hz = hz; a.Yin = in->Y; a.Zin = in->Z;
    a.in = *in;
    if (src2) a.in2 = *src2;
    a.ts_mode = best_ts ? 1 : 0;
    a.na = 2;
    a.a_cols = (uint32_t)(16 * BY);
    a.a_tmem_col = 512u - 2u * a.a_cols;
    // boxes
    for (int i = 0; i < nbox; ++i) {
        TcBox& bx = a.boxes[i];
        bx.dst16 = (uint16_t)((size_t)i * box_bytes / 16);
        if (tr) { bx.dy = 0; bx.dz = (int8_t)i; }
        else if (strided) { bx.dy = (int8_t)(i / KZ - hy); bx.dz = (int8_t)(i % KZ - hz); }
        else { bx.dy = (int8_t)-hy; bx.dz = (int8_t)(i - hz); }
    }
    // accumulators + ops
    int nacc = 0, nop = 0;
    if (!gen_ops(G, YT, a.ops, &nop, a.accs, &nacc)) return false;
    a.nacc = nacc;
    a.nop = nop;
    if (src2) {
        int n2 = 0;
        for (int y = 0; y < YT; ++y) {
            uint32_t aoff;
            // the shortcut source is staged with the main box shape: the centre sits at (+hy, +hz)
            if (line) aoff = (uint32_t)(((y + hy) * BZ + hz) * 16);
            else aoff = (uint32_t)((y * LY + hy) * LZ * 16);   // staged unshifted in z (dz2 = 0)
            if (sc_self && !line) aoff += (uint32_t)hz * box_bytes;   // main stage: the unshifted box is the centre z tap
            for (int pass = 0; pass < 3; ++pass) {
                TcOp& op = a.ops2[n2++];
                op.a16 = (uint16_t)(aoff / 16 + (pass == 1 ? a.a_plane / 16 : 0));
                op.b16 = (uint16_t)(pass == 2 ? a.b2_plane / 16 : 0);
                op.col = (uint16_t)(y * n_cta);   // relative to the shortcut accumulators of the stage's x row
                op.n8 = (uint16_t)(n_cta / 8);
            }
        }
        a.nop2 = n2;
        a.dy2 = -hy;
        a.dz2 = 0;
    }
    {
        // descriptor words (sm_100 K-major SWIZZLE_NONE: start >> 4 | LBO >> 4 << 16 ; SBO >> 4 | version 1 << 14).
        // gen_ops emits the three passes of a merged product back to back; pass 0 carries the plane-0 offsets
        a.desc_hi = (128u >> 4) | (1u << 14);
        // TS mode: the A word is the view's first column inside the stage's A buffer (a view = one 2 KB z line)
        for (int i = 0; i < a.nop / 3; ++i) {
            const TcOp& o = a.ops[3 * i];
            const uint32_t aw = a.ts_mode ? (uint32_t)(o.a16 / 128) * 8u : o.a16 + (((a.lbo_a >> 4) & 0x3FFFu) << 16);
            a.el[i] = {aw, o.b16 + (((a.lbo_b >> 4) & 0x3FFFu) << 16), o.col, a.idesc | ((uint32_t)o.n8 << 17)};
        }
        for (int i = 0; i < a.nop2 / 3; ++i) {
            const TcOp& o = a.ops2[3 * i];
            const uint32_t aw = a.ts_mode ? (uint32_t)(o.a16 / 128) * 8u : o.a16 + (((a.lbo_a >> 4) & 0x3FFFu) << 16);
            a.el2[i] = {aw, o.b16 + (((a.lbo_b2 >> 4) & 0x3FFFu) << 16), o.col, a.idesc | ((uint32_t)o.n8 << 17)};
        }
        a.partial = (Ym % LY) != 0 ? 1 : 0;
        for (int i = 0; i < a.nacc; ++i) a.accs[i].off8 = (a.accs[i].y_add * out->Z + a.accs[i].z_add) * 8;
    }
    a.ntiles = (int)((long)in->B * a.ntx * a.nty * a.ntz * a.nsel);
    const int sms = sm_count();
    const int cols1 = nacc * (src2 ? 2 : 1) * n_cta;
    a.nbuf = best_slots;   // TMEM row slots
    const int cols = cols1 * a.nbuf;
    a.tmem_cols = a.ts_mode ? 512 : cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    P->box[0] = 8; P->box[1] = (cuuint32_t)(BZ * (strided ? g->sz : 1)); P->box[2] = (cuuint32_t)(BY * (strided ? g->sy : 1));
    P->box[3] = 1; P->box[4] = 2;
    P->estr[0] = 1; P->estr[1] = (cuuint32_t)(strided ? g->sz : 1); P->estr[2] = (cuuint32_t)(strided ? g->sy : 1);
    P->estr[3] = 1; P->estr[4] = 1;
    P->smem = TC_HDR + (size_t)a.nstage * a.stage_bytes;
    P->grid = (unsigned)(a.ntiles < sms ? a.ntiles : sms);   // persistent: one CTA per SM
    // transposed conv: the x phase is the fastest tile index and the odd phase stages twice the planes of the even
    // one; with an even grid every CTA would keep one phase for all its tiles (2x imbalance) - an odd grid alternates
    if (a.npx == 2 && a.ntiles > (int)P->grid && P->grid % 2 == 0) P->grid -= 1;
    return true;
}

// The tile search walks a few hundred candidates: plans are cached per shape (everything but the
// base pointers).  Calls are serialised by the host thread that drives the stream.
struct TcKey {
    int64_t v[26];
    bool operator==(const TcKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
static void key_act(const vsseg_act8* t, int64_t* v) {
    v[0] = t->lo_offset; v[1] = t->batch_stride; v[2] = t->B; v[3] = t->C; v[4] = t->X; v[5] = t->Y; v[6] = t->Z;
}
static bool make_plan(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, int n_split,
                      const vsseg_act8* src2, TcPlan* P) {
    if (!in || !out || !g || !in->hi || !out->hi) return false;
    TcKey k;
    memset(&k, 0, sizeof(k));
    key_act(in, k.v); key_act(out, k.v + 7);
    if (src2) key_act(src2, k.v + 14);
    k.v[21] = src2 ? (src2->hi == in->hi ? 2 : 1) : 0;   // a shortcut that reads the conv's own input is planned differently
    k.v[22] = g->kx | (g->ky << 4) | (g->kz << 8) | (g->sx << 12) | (g->sy << 16) | (g->sz << 20) | ((g->transposed ? 1 : 0) << 24);
    k.v[23] = n_split;
    if (const char* f = getenv("VSSEG_TC_FORCE")) {   // forced tiles are separate cache entries
        int fx = 0, fy = 0, fn = 0;
        sscanf(f, "%d,%d,%d", &fx, &fy, &fn);
        k.v[24] = fx; k.v[25] = fy * 16 + fn;
    }
    struct Entry { TcKey k; bool ok; TcPlan p; };
    static std::vector<Entry>* cache = new std::vector<Entry>();
    for (const Entry& e : *cache)
        if (e.k == k) {
            if (!e.ok) return false;
            *P = e.p;
            P->a.in = *in; P->a.out = *out;
            if (src2) P->a.in2 = *src2;
            return true;
        }
    Entry e;
    e.k = k;
    e.ok = make_plan_uncached(in, out, g, n_split, src2, &e.p);
    if (cache->size() < 4096) cache->push_back(e);
    if (e.ok) *P = e.p;
    return e.ok;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(CUtensorMap* tmap, const vsseg_act8* t, int cg_plane, int cg_batch, const TcPlan& P) {
    // the driver entry point is resolved through the runtime so the library has no link-time
    // dependency on libcuda.so (it must load on a GPU-less build box)
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("conv3d_tc: cuTensorMapEncodeTiled is not available from the driver");
            return e != cudaSuccess ? (int)e : VSSEG_EINVAL;
        }
        encode = (EncodeFn)fn;
    }
    const int64_t cgs = (int64_t)t->X * t->Y * t->Z * 8;
    const cuuint64_t gdim[5] = {8, (cuuint64_t)t->Z, (cuuint64_t)t->Y, (cuuint64_t)t->X,
                                (cuuint64_t)(cg_plane + (t->B - 1) * cg_batch + t->C / 8)};
    const cuuint64_t gstr[4] = {16, (cuuint64_t)t->Z * 16, (cuuint64_t)t->Y * t->Z * 16, (cuuint64_t)cgs * 2};
    CUresult cr;
    if (P.a.map_wide) {
        // (z, 8 channels) merged into one dimension of 8-byte elements: 2 per voxel, a box row = a z line
        const cuuint64_t gdim4[4] = {(cuuint64_t)t->Z * 2, gdim[2], gdim[3], gdim[4]};
        const cuuint32_t box4[4] = {P.box[1] * 2, P.box[2], 1, 2};
        const cuuint32_t estr4[4] = {1, P.estr[2], 1, 1};
        cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, t->hi, gdim4, gstr + 1, box4, estr4,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t->hi, gdim, gstr, P.box, P.estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) {
        set_error("conv3d_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return (int)cr;
    }
    return 0;
}

// epilogue flavours are compile-time (out mode, fused shortcut, residual mode): the dead paths cost registers
typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const TcArgs);
template <bool TS>
static TcKernel pick_kernel_ts(const TcArgs& a) {
    const bool sc = a.nchunk2 != 0;
    if (a.out_mode == 2) return conv_tc_kernel<2, false, 0, TS>;
    if (a.out_mode == 1) return conv_tc_kernel<1, false, 0, TS>;
    if (sc) return a.res_mode == 0 ? conv_tc_kernel<0, true, 0, TS> : a.res_mode == 1 ? conv_tc_kernel<0, true, 1, TS> : conv_tc_kernel<0, true, 2, TS>;
    return a.res_mode == 0 ? conv_tc_kernel<0, false, 0, TS> : a.res_mode == 1 ? conv_tc_kernel<0, false, 1, TS> : conv_tc_kernel<0, false, 2, TS>;
}
static TcKernel pick_kernel(const TcArgs& a) { return a.ts_mode ? pick_kernel_ts<true>(a) : pick_kernel_ts<false>(a); }
static int set_smem_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        const TcKernel all[] = {conv_tc_kernel<2, false, 0, false>, conv_tc_kernel<1, false, 0, false>, conv_tc_kernel<0, true, 0, false>,
                                conv_tc_kernel<0, true, 1, false>, conv_tc_kernel<0, true, 2, false>, conv_tc_kernel<0, false, 0, false>,
                                conv_tc_kernel<0, false, 1, false>, conv_tc_kernel<0, false, 2, false>,
                                conv_tc_kernel<2, false, 0, true>, conv_tc_kernel<1, false, 0, true>, conv_tc_kernel<0, true, 0, true>,
                                conv_tc_kernel<0, true, 1, true>, conv_tc_kernel<0, true, 2, true>, conv_tc_kernel<0, false, 0, true>,
                                conv_tc_kernel<0, false, 1, true>, conv_tc_kernel<0, false, 2, true>};
        for (TcKernel k : all) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_error("conv3d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return (int)e;
            }
        }
        attr_set = true;
    }
    return 0;
}

// VSSEG_TC_DEBUG=1: every launch is followed by a synchronous dump of the per-role cycle counters
static void launch_tc(const TcPlan& P, const CUtensorMap& tmap, const CUtensorMap& tmap2, TcArgs& a, cudaStream_t stream,
                      const char* what) {
    static const bool debug = getenv("VSSEG_TC_DEBUG") && atoi(getenv("VSSEG_TC_DEBUG"));
    static unsigned long long* dbuf = nullptr;
    a.dbg = nullptr;
    a.dbgf = debug && getenv("VSSEG_TC_DBGF") ? atoi(getenv("VSSEG_TC_DBGF")) : 0;
    if (debug) {
        if (!dbuf) cudaMalloc(&dbuf, 256 * 8 * sizeof(unsigned long long));
        cudaMemsetAsync(dbuf, 0, 256 * 8 * sizeof(unsigned long long), stream);
        a.dbg = dbuf;
    }
    static const bool pdl = getenv("VSSEG_TC_PDL") && atoi(getenv("VSSEG_TC_PDL")) != 0;
    a.pdl = pdl && !debug ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = P.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, pick_kernel(a), tmap, tmap2, a);
    if (debug) {
        static unsigned long long h[256 * 8];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
        double s[8] = {0};
        for (unsigned i = 0; i < P.grid; ++i)
            for (int k = 0; k < 8; ++k) s[k] += (double)h[i * 8 + k] / P.grid;
        fprintf(stderr, "[tc-debug] %s cin=%d cout=%d X=%d XT=%d YL=%d slots=%d nstage=%d grid=%u | mma: total %.0f wait_full %.0f wait_acc %.0f n_mma %.0f "
                        "(%.1f cyc/mma busy) | prod: total %.0f wait_empty %.0f | epi: total %.0f wait_full %.0f\n",
                what, a.nchunk * 16, a.cout, a.out.X, a.XT, a.YL, a.nbuf, a.nstage, P.grid, s[2], s[0], s[1], s[7],
                s[7] > 0 ? (s[2] - s[0] - s[1]) / s[7] : 0.0, s[4], s[3], s[6], s[5]);
    }
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_conv3d_tc_supported(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, int32_t n_split,
                              const vsseg_act8* shortcut_src) {
    TcPlan P;
    return make_plan(in, out, g, n_split, shortcut_src, &P) ? 1 : 0;
}

int vsseg_conv3d_tc_suggest_split(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                                  const vsseg_act8* shortcut_src) {
    static TcPlan P;
    if (!make_plan(in, out, g, 1, shortcut_src, &P)) {
        // a full-width N may not fit TMEM next to a shortcut accumulator: try the finest split
        const int n16 = out ? (out->C + 15) / 16 : 0;
        for (int s = 2; s <= n16; ++s)
            if (n16 % s == 0 && make_plan(in, out, g, s, shortcut_src, &P)) return s;
        return 0;
    }
    const long tiles = (long)P.grid;
    const int n16 = (out->C + 15) / 16;
    int best = 1;
    for (int s = 1; s <= n16; ++s) {
        if (n16 % s) continue;
        if (!make_plan(in, out, g, s, shortcut_src, &P)) continue;
        best = s;
        if ((long)P.grid >= 120 || tiles * s >= 120) break;
    }
    return best;
}

static int tc_f32out(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g, const void* w_packed,
                     const vsseg_epilogue* ep, const float* sw_weight, void* stream, int two_pass, const vsseg_act8* gate = nullptr) {
    VSSEG_REQUIRE(in && f32_ok(out) && out->C >= 1 && out->C <= 2, "conv3d_tc_f32out: Cout must be 1 or 2");
    // the plan only needs the output extents: describe the planar output as a 16-channel act8 tensor
    vsseg_act8 o16{};
    o16.hi = (void*)16; o16.B = out->B; o16.C = 16; o16.X = out->X; o16.Y = out->Y; o16.Z = out->Z;
    static TcPlan P;
    VSSEG_REQUIRE(g && !g->transposed && g->sx == 1 && g->sy == 1 && g->sz == 1 && make_plan(in, &o16, g, 1, nullptr, &P),
                  "conv3d_tc_f32out: unsupported shape (stride-1 convs covered by vsseg_conv3d_tc_supported)");
    VSSEG_REQUIRE(w_packed && ep && ep->scale && ep->shift, "conv3d_tc_f32out: NULL weights/epilogue");
    TcArgs& a = P.a;
    a.ep = *ep;
    a.w = (const uint8_t*)w_packed;
    a.out_mode = 1;
    if (gate) {
        VSSEG_REQUIRE(gate->hi && gate->C % 8 == 0 && gate->B == out->B && gate->X == out->X && gate->Y == out->Y && gate->Z == out->Z &&
                          out->C == 1 && !sw_weight, "conv3d_tc_attgate: the gated tensor must have the map's extents (Cout = 1)");
        a.out_mode = 2;
        a.gate = *gate;
    }
    a.outf = *out;
    a.cout = out->C;
    a.two_pass = two_pass;
    a.sw_weight = sw_weight;
    CUtensorMap tmap;
    if (a.line_mode) memset(&tmap, 0, sizeof(tmap));
    else if (int e = encode_map(&tmap, in, a.cg_plane, a.cg_batch, P)) return e;
    if (int e = set_smem_attr()) return e;
    launch_tc(P, tmap, tmap, a, (cudaStream_t)stream, "f32out");
    return check_launch("conv3d_tc_f32out");
}

int vsseg_conv3d_tc_f32out(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g, const void* w_packed,
                           const vsseg_epilogue* ep, const float* sw_weight, void* stream) {
    return tc_f32out(in, out, g, w_packed, ep, sw_weight, stream, 0);
}

int vsseg_conv3d_tc_f32out_2p(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g, const void* w_packed,
                              const vsseg_epilogue* ep, const float* sw_weight, void* stream) {
    return tc_f32out(in, out, g, w_packed, ep, sw_weight, stream, 1);
}

int vsseg_conv3d_tc_attgate(const vsseg_act8* in, const vsseg_f32view* att, const vsseg_conv_geom* g, const void* w_packed,
                            const vsseg_epilogue* ep, const vsseg_act8* gated, void* stream) {
    VSSEG_REQUIRE(gated, "conv3d_tc_attgate: NULL gated tensor");
    return tc_f32out(in, att, g, w_packed, ep, nullptr, stream, 1, gated);
}

int vsseg_conv3d_tc_describe(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, int32_t n_split,
                             const vsseg_act8* shortcut_src, char* buf, int32_t buflen) {
    static TcPlan P;
    VSSEG_REQUIRE(buf && buflen > 0, "conv3d_tc_describe: no buffer");
    if (!make_plan(in, out, g, n_split, shortcut_src, &P)) {
        snprintf(buf, buflen, "unsupported");
        return 0;
    }
    const TcArgs& a = P.a;
    int n = snprintf(buf, buflen,
                     "grid=%u smem=%zu nstage=%d stage_bytes=%u line_mode=%d LZ=%d LY=%d YL=%d BY=%d pitch=%d nbox=%d "
                     "box=[%u,%u,%u,%u,%u] estr=[%u,%u] box_tx=%u a_plane=%u b_off=%u b_bytes=%u lbo_a=%u lbo_b=%u nacc=%d "
                     "n_cta=%d tmem_cols=%u nchunk=%d nj=%d nchunk2=%d nop=%d nop2=%d ntx=%d nty=%d ntz=%d nsel=%d XT=%d slots=%d ts=%d sc_self=%d ops:",
                     P.grid, P.smem, a.nstage, a.stage_bytes, a.line_mode, a.LZ, a.LY, a.YL, a.BY, a.pitch, a.nbox, P.box[0],
                     P.box[1], P.box[2], P.box[3], P.box[4], P.estr[1], P.estr[2], a.box_tx, a.a_plane, a.b_off, a.b_bytes,
                     a.lbo_a, a.lbo_b, a.nacc, a.n_cta, a.tmem_cols, a.nchunk, a.nj, a.nchunk2, a.nop, a.nop2, a.ntx, a.nty,
                     a.ntz, a.nsel, a.XT, a.nbuf, a.ts_mode, a.sc_self);
    for (int i = 0; i < a.nop && n < buflen - 40; ++i)
        n += snprintf(buf + n, buflen - n, " (a%u b%u c%u n%u)", a.ops[i].a16, a.ops[i].b16, a.ops[i].col, a.ops[i].n8 * 8);
    if (a.nop2 && n < buflen - 40) n += snprintf(buf + n, buflen - n, " ops2:");   // the fused shortcut's products
    for (int i = 0; i < a.nop2 && n < buflen - 40; ++i)
        n += snprintf(buf + n, buflen - n, " (a%u b%u c%u n%u)", a.ops2[i].a16, a.ops2[i].b16, a.ops2[i].col, a.ops2[i].n8 * 8);
    return 0;
}

int vsseg_conv3d_tc(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, const void* w_packed,
                    int32_t n_split, const vsseg_epilogue* ep, const vsseg_act8* res_act8, const vsseg_f32view* res_src,
                    const float* res_w, const float* res_b, const vsseg_act8* shortcut_src, const void* shortcut_w,
                    const float* shortcut_bias, void* stream) {
    static TcPlan P;  // 2.5 KB of tables: kept off the stack; calls are serialised by the host thread
    VSSEG_REQUIRE(make_plan(in, out, g, n_split, shortcut_src, &P),
                  "conv3d_tc: unsupported shape (see vsseg_conv3d_tc_supported)");
    VSSEG_REQUIRE(w_packed && ep && ep->scale && ep->shift, "conv3d_tc: NULL weights/epilogue");
    VSSEG_REQUIRE(!(res_act8 && res_src), "conv3d_tc: at most one residual source");
    VSSEG_REQUIRE(!shortcut_src || (shortcut_w && shortcut_bias), "conv3d_tc: shortcut needs weights and bias");
    TcArgs& a = P.a;
    a.ep = *ep;
    a.w = (const uint8_t*)w_packed;
    a.w2 = (const uint8_t*)shortcut_w;
    a.bias2 = shortcut_bias;
    if (res_act8) {
        VSSEG_REQUIRE(res_act8->hi && res_act8->C == out->C && res_act8->X == out->X && res_act8->Y == out->Y &&
                          res_act8->Z == out->Z && res_act8->B == out->B, "conv3d_tc: residual shape mismatch");
        a.res_mode = 1;
        a.res = *res_act8;
    } else if (res_src) {
        VSSEG_REQUIRE(f32_set_ok(res_src) && res_w && res_b, "conv3d_tc: NULL cin1 residual");
        a.res_mode = 2;
        a.rsrc = *res_src; a.res_w = res_w; a.res_b = res_b;
        VSSEG_REQUIRE(win_tab(res_src, out->B, &a.rwin), "conv3d_tc: inconsistent window set (%d records for batch %d)",
                      res_src->n_windows, out->B);
    }
    CUtensorMap tmap, tmap2;
    if (a.line_mode) {  // staged with plain bulk copies: no tensor map needed
        memset(&tmap, 0, sizeof(tmap));
        tmap2 = tmap;
    } else {
        if (int e = encode_map(&tmap, in, a.cg_plane, a.cg_batch, P)) return e;
        if (shortcut_src) {
            if (int e = encode_map(&tmap2, shortcut_src, a.cg_plane2, a.cg_batch2, P)) return e;  // never strided
        } else {
            tmap2 = tmap;
        }
    }
    if (int e = set_smem_attr()) return e;
    launch_tc(P, tmap, tmap2, a, (cudaStream_t)stream, "conv");
    return check_launch("conv3d_tc");
}

}  // extern "C"

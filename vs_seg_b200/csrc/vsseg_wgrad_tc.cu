// Tensor-core weight gradient of the stride-1 convolutions on 128-voxel z lines (levels 1-3 of the U-Net, where
// ~95 % of the weight-gradient time is): dW[tap][ci][co] = sum_v x[v + tap][ci] * dc[v][co]   (autograd of
// reference convolutions.py:137-146).
//
// GEMM view per tap: D[co (M = 128, rows >= Cout unused), ci (N = Cin)] += sum over K = voxels of dc^T[co, v] * x[v+tap, ci].
// Both operands are read "MN-major" straight from the act8 line layout [z][8 channels]: the 8 channels of a voxel are
// the contiguous (MN) direction of a core matrix, consecutive z rows (16 B apart) its K direction - so a staged z line
// of dc is an A operand and a staged (haloed) z line of x, started dz rows later, is the B operand of tap dz.
// A CTA owns one (dx, dy) tap pair (all KZ z-taps as separate TMEM accumulators) and a slice of the output lines
// (split-K); it ends with fp32 atomics into dW.  bf16x3 like the forward pass (hi*hi + lo*hi + hi*lo).
#include "vsseg_ptx.cuh"

namespace vsseg {

struct WgArgs {
    vsseg_act8 x, dc;
    float* dw;               // [taps][Cin][cout_pad]
    int cout_pad, KX, KY, KZ;
    int nslice;              // CTAs per (dx, dy) pair
    // line pairing on the coarse (loop) grid XL x YL: x line = l * ax + (d - p) * bx, dc line = l * ad + (d - p) * bd.
    // stride-1: (1,1,1,0); strided conv (x fine, dc coarse): (2,1,1,0); transposed conv (x coarse, dc fine): (1,0,2,1)
    int ax, bx, ad, bd, XL, YL;
    int LZ;                  // z extent of a staged line: 128, or the whole Z (32 / 64) on the coarse levels
    int nstage;
    uint32_t dc_plane, x_plane, x_off, stage_bytes;   // bytes: dc plane size, x plane size, x region start
    uint32_t idesc;
    uint32_t tmem_cols;
};

__device__ uint4 g_zero_line_wg[8];   // zero source for re-zeroing halo rows

constexpr int WG_SLACK = 32 * 1024;    // M = 128 reads 16 channel-group slots of dc: the slots beyond Cout may reach past the last stage
constexpr int WG_THREADS = 192;   // warp 0: bulk-copy producer, warp 1: MMA issuer, warps 2-5: final TMEM -> atomics
constexpr int WG_HDR = 1024;

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_tc_kernel(const __grid_constant__ WgArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    uint64_t* acc_full = full + 16;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 18);
    const uint32_t ring = smem_u32(smem) + WG_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x % a.nslice, pair = blockIdx.x / a.nslice;
    const int dx = pair / a.KY, dy = pair % a.KY;
    const int px = (a.KX - 1) / 2, py = (a.KY - 1) / 2, hz = (a.KZ - 1) / 2;
    const int X = a.XL, Y = a.YL, Z = a.dc.Z;      // loop grid (the coarser of the two tensors)
    const int Xx = a.x.X, Yx = a.x.Y, Xd = a.dc.X, Yd = a.dc.Y;
    const int LZ = a.LZ;
    const int nzt = Z / LZ;
    const int nlines = a.dc.B * X * Y * nzt;
    const int pitch = LZ + 2 * hz;
    const int Cin = a.x.C, Cout = a.dc.C;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstage; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
    {   // zero the ring: halo rows at the z border of the volume stay zero (never written by the line copies)
        uint4* q = reinterpret_cast<uint4*>(smem + WG_HDR);
        const int n16 = (int)(a.nstage * a.stage_bytes / 16);
        for (int i = threadIdx.x; i < n16; i += WG_THREADS) q[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== producer: one stage = the dc line and the (dx, dy)-shifted x line of one output line =====
        int it = 0;
        for (int ln = slice; ln < nlines; ln += a.nslice) {
            int t = ln;
            const int zt = t % nzt; t /= nzt;
            const int y = t % Y; t /= Y;
            const int xx = t % X; t /= X;
            const int b = t;
            const int xi = xx * a.ax + (dx - px) * a.bx, yi = y * a.ax + (dy - py) * a.bx;
            const int xd = xx * a.ad + (dx - px) * a.bd, yd = y * a.ad + (dy - py) * a.bd;
            if (xi < 0 || xi >= Xx || yi < 0 || yi >= Yx || xd < 0 || xd >= Xd || yd < 0 || yd >= Yd) continue;   // padding line
            const int st = it % a.nstage;
            mbar_wait(empty + st, ((it / a.nstage) & 1) ^ 1);
            const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
            const int z0 = zt * LZ;
            const int zlo = max(z0 - hz, 0), zhi = min(z0 + LZ + hz, Z);
            const uint32_t xrow = (uint32_t)(zhi - zlo) * 16;
            const int ndc = (Cout / 8) * 2, nx = (Cin / 8) * 2;
            // several z tiles: halo rows of an interior tile carry data, of a border tile must be re-zeroed
            const bool zl = nzt > 1 && hz > 0 && z0 - hz < 0, zh = nzt > 1 && hz > 0 && z0 + LZ + hz > Z;
            if (lane == 0)
                mbar_expect_tx(full + st, (uint32_t)ndc * (uint32_t)(LZ * 16) + (uint32_t)nx * (xrow + (zl ? 16u : 0u) + (zh ? 16u : 0u)));
            const int64_t nvox_d = (int64_t)Xd * Yd * Z, nvox_x = (int64_t)Xx * Yx * Z;
            for (int q = lane; q < ndc + nx; q += 32) {
                if (q < ndc) {
                    const int plane = q & 1, cg = q >> 1;
                    const __nv_bfloat16* gp = (const __nv_bfloat16*)a.dc.hi + (int64_t)plane * a.dc.lo_offset +
                                              (int64_t)b * a.dc.batch_stride +
                                              ((int64_t)cg * nvox_d + ((int64_t)xd * Yd + yd) * Z + z0) * 8;
                    bulk_load(base + (uint32_t)plane * a.dc_plane + (uint32_t)(cg * LZ * 16), gp, (uint32_t)(LZ * 16), full + st);
                } else {
                    const int r = q - ndc, plane = r & 1, cg = r >> 1;
                    const __nv_bfloat16* gp = (const __nv_bfloat16*)a.x.hi + (int64_t)plane * a.x.lo_offset +
                                              (int64_t)b * a.x.batch_stride +
                                              ((int64_t)cg * nvox_x + ((int64_t)xi * Yx + yi) * Z + zlo) * 8;
                    const uint32_t line = base + a.x_off + (uint32_t)plane * a.x_plane + (uint32_t)(cg * pitch) * 16;
                    bulk_load(line + (uint32_t)(zlo - (z0 - hz)) * 16, gp, xrow, full + st);
                    if (zl) bulk_load(line, g_zero_line_wg, 16, full + st);
                    if (zh) bulk_load(line + (uint32_t)(pitch - 1) * 16, g_zero_line_wg, 16, full + st);
                }
            }
            __syncwarp();
            ++it;
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            int it = 0;
            uint32_t first = 1;
            for (int ln = slice; ln < nlines; ln += a.nslice) {
                int t = ln / nzt;
                const int y = t % Y; t /= Y;
                const int xx = t % X;
                const int xi = xx * a.ax + (dx - px) * a.bx, yi = y * a.ax + (dy - py) * a.bx;
                const int xd = xx * a.ad + (dx - px) * a.bd, yd = y * a.ad + (dy - py) * a.bd;
                if (xi < 0 || xi >= Xx || yi < 0 || yi >= Yx || xd < 0 || xd >= Xd || yd < 0 || yd >= Yd) continue;
                const int st = it % a.nstage;
                mbar_wait(full + st, (it / a.nstage) & 1);
                tc_fence_after();
                const uint32_t base = ring + (uint32_t)st * a.stage_bytes;
                // MN-major, no swizzle: SBO = stride between 8-channel groups, LBO = stride between 8-row K blocks
                const uint64_t da_h = make_desc(base, 128, (uint32_t)(LZ * 16));
                const uint64_t da_l = make_desc(base + a.dc_plane, 128, (uint32_t)(LZ * 16));
                const uint64_t db_h = make_desc(base + a.x_off, 128, (uint32_t)pitch * 16);
                const uint64_t db_l = make_desc(base + a.x_off + a.x_plane, 128, (uint32_t)pitch * 16);
                for (int dz = 0; dz < a.KZ; ++dz) {
                    const uint32_t d = tmem_base + (uint32_t)(dz * Cin);
#pragma unroll 4
                    for (int ks = 0; ks < (LZ >> 4); ++ks) {
                        const uint64_t ka = (uint64_t)(ks * 16), kb = (uint64_t)(ks * 16 + dz);   // 16 B units: 16 rows per K step
                        umma_bf16(d, da_h + ka, db_h + kb, a.idesc, (first && ks == 0) ? 0u : 1u);
                        umma_bf16(d, da_l + ka, db_h + kb, a.idesc, 1u);
                        umma_bf16(d, da_h + ka, db_l + kb, a.idesc, 1u);
                    }
                }
                first = 0;
                umma_commit(empty + st);
                ++it;
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== epilogue: TMEM lane = output channel co, column = ci of tap dz -> atomicAdd into dW =====
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int lane_base = (warp & 3) * 32;
        const int co = lane_base + lane;
        // did any line run?  (a slice can be empty for tiny layers)
        bool any = false;
        for (int ln = slice; ln < nlines && !any; ln += a.nslice) {
            int t = ln / nzt;
            const int y = t % Y; t /= Y;
            const int xx = t % X;
            const int xi = xx * a.ax + (dx - px) * a.bx, yi = y * a.ax + (dy - py) * a.bx;
            const int xd = xx * a.ad + (dx - px) * a.bd, yd = y * a.ad + (dy - py) * a.bd;
            any = !(xi < 0 || xi >= Xx || yi < 0 || yi >= Yx || xd < 0 || xd >= Xd || yd < 0 || yd >= Yd);
        }
        if (any) {
            for (int dz = 0; dz < a.KZ; ++dz) {
                const int tap = (dx * a.KY + dy) * a.KZ + dz;
                for (int c0 = 0; c0 < Cin; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(dz * Cin + c0), v);
                    tmem_ld_wait();
                    if (co < Cout) {
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            atomicAdd(a.dw + ((int64_t)tap * Cin + c0 + q) * a.cout_pad + co, __uint_as_float(v[q]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_conv3d_wgrad_tc_supported(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g) {
    if (!x || !dc || !g || g->sz != 1 || g->sx != g->sy || g->sx < 1 || g->sx > 2) return 0;
    if ((g->kx != 1 && g->kx != 3) || (g->ky != 1 && g->ky != 3) || (g->kz != 1 && g->kz != 3)) return 0;
    if (g->sx == 2 && (g->kx != 3 || g->ky != 3)) return 0;
    if (g->transposed && g->sx != 2) return 0;
    // x = the conv's input, dc = gradient of its output: equal extents (stride 1), x twice dc (strided conv),
    // dc twice x (transposed conv: output_padding makes the output exactly 2x); z is never strided here
    const int fx = g->transposed ? 1 : g->sx, fd = g->transposed ? g->sx : 1;
    if (x->X * fd != dc->X * fx || x->Y * fd != dc->Y * fx || x->Z != dc->Z || x->B != dc->B) return 0;
    if ((x->Z % 128 && x->Z != 64 && x->Z != 32) || x->C % 16 || dc->C % 8 || dc->C > 128 || x->C > 256) return 0;
    if (g->kz * x->C > 512) return 0;
    return 1;
}

int vsseg_conv3d_wgrad_tc(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g, float* dw, int32_t cout_pad,
                          void* stream) {
    VSSEG_REQUIRE(vsseg_conv3d_wgrad_tc_supported(x, dc, g), "conv3d_wgrad_tc: unsupported shape");
    VSSEG_REQUIRE(dw && cout_pad >= dc->C, "conv3d_wgrad_tc: bad arguments");
    WgArgs a{};
    a.x = *x; a.dc = *dc; a.dw = dw; a.cout_pad = cout_pad;
    a.KX = g->kx; a.KY = g->ky; a.KZ = g->kz;
    if (g->transposed) { a.ax = 1; a.bx = 0; a.ad = 2; a.bd = 1; a.XL = x->X; a.YL = x->Y; }
    else if (g->sx == 2) { a.ax = 2; a.bx = 1; a.ad = 1; a.bd = 0; a.XL = dc->X; a.YL = dc->Y; }
    else { a.ax = 1; a.bx = 1; a.ad = 1; a.bd = 0; a.XL = dc->X; a.YL = dc->Y; }
    a.LZ = x->Z >= 128 ? 128 : x->Z;
    const int hz = (g->kz - 1) / 2, pitch = a.LZ + 2 * hz;
    // M = 128 always: the A rows beyond Cout read whatever follows the dc plane in shared memory; a row of A only
    // feeds its own row of D, and those TMEM lanes are never read
    a.dc_plane = (uint32_t)((dc->C / 8) * a.LZ * 16);
    a.x_plane = (uint32_t)((x->C / 8) * pitch * 16);
    a.x_plane = (a.x_plane + 127) / 128 * 128;
    a.x_off = 2 * a.dc_plane;
    a.stage_bytes = a.x_off + 2 * a.x_plane;
    const long budget = 227L * 1024 - WG_HDR - WG_SLACK;
    int nst = (int)(budget / (long)a.stage_bytes);
    VSSEG_REQUIRE(nst >= 2, "conv3d_wgrad_tc: stage does not fit twice in shared memory");
    a.nstage = nst > 6 ? 6 : nst;
    // instruction descriptor: f32 accumulate, bf16 x bf16, A and B MN-major (bits 15, 16), N = Cin, M = 128
    a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(x->C >> 3) << 17) | (8u << 24);
    const int cols = g->kz * x->C;
    a.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    const int pairs = g->kx * g->ky;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long nlines = (long)dc->B * a.XL * a.YL * (dc->Z / a.LZ);
    long ns = sms / pairs;
    if (ns < 1) ns = 1;
    if (ns > nlines) ns = nlines;
    a.nslice = (int)ns;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("conv3d_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set = true;
    }
    conv_wgrad_tc_kernel<<<(unsigned)(pairs * a.nslice), WG_THREADS, WG_HDR + WG_SLACK + (size_t)a.nstage * a.stage_bytes,
                           (cudaStream_t)stream>>>(a);
    return check_launch("conv3d_wgrad_tc");
}

}  // extern "C"

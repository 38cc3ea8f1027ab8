// Tensor-core (tcgen05 + TMEM + TMA) implicit-GEMM 3-D convolution for the FLOP-heavy stride-1
// layers of UNet2d5_spvPA (reference params/networks/blocks/convolutions.py:137-156), sm_100a only.
//
// GEMM view: M = 128 consecutive z voxels of one (x,y) line, N = Cout, K = Cin x taps.
// Operands are the two bf16 planes of the act8 layout (value = hi + lo), and every product is
// evaluated as hi*hi + lo*hi + hi*lo ("bf16x3") into one fp32 TMEM accumulator, which keeps the
// network inside the 1e-3 parity bar that single-pass bf16/tf32 misses.
//
// Data flow per CTA (a tile of XT x YT output lines, all Cout channels):
//   for each 16-channel slice of Cin:
//     for each halo x-plane px in [x0-1, x0+XT]:
//       TMA (5-D tiled, OOB zero fill = "same" padding) stages the haloed slab
//         [cg 2][y0-1 .. y0+YT][z0-hz .. z0+127+hz][8 ch]   for the hi and the lo plane  -> smem ring
//       cp.async.bulk stages the packed weights of every tap with this dx (3 slots, one per dx)
//       one thread issues tcgen05.mma for every (output line, dy, dz) that reads this plane:
//         the A descriptor is just a start-address offset into the slab (no-swizzle K-major
//         layout, rows 16 B apart), so each staged element is reused by up to 9*XT... taps from
//         shared memory instead of being re-fetched from L2 per tap.
//   epilogue warps: tcgen05.ld -> BN scale/shift -> PReLU/ReLU -> (+ residual) -> split-bf16 -> global.
#include <cuda.h>

#include "vsseg_common.cuh"

namespace vsseg {

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 format: version 1 at bit 46).
// rows (M or N) are 16 B apart inside an 8-row core matrix, SBO = stride between 8-row groups,
// LBO = stride between the two 8-element K halves of one K=16 instruction.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

struct TcArgs {
    vsseg_act8 out;
    vsseg_epilogue ep;
    int res_mode;
    vsseg_act8 res;
    vsseg_f32view rsrc;
    const float* res_w;
    const float* res_b;
    const __nv_bfloat16* w;  // packed [Cin/16][3 dx][2 plane][3 dy][KZ][2 khalf][Cout][8]
    int Cin, Cout, X, Y, Z, B, KZ;
    int XT, YT, SA;          // tile lines and A-ring depth
    int cg_plane, cg_batch;  // merged-cg index strides of the TMA map (lo plane, batch)
    uint32_t a_block, a_box_bytes, a_stage, b_slot, tmem_cols;
};

constexpr int TC_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);  // [SA]
    uint64_t* a_empty = a_full + 8;                        // [SA]
    uint64_t* b_full = a_full + 16;                        // [3]
    uint64_t* b_empty = a_full + 20;                       // [3]
    uint64_t* acc_full = a_full + 24;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full + 26);
    uint8_t* a_ring = smem + 1024;
    uint8_t* b_ring = a_ring + (size_t)a.SA * a.a_stage;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int XT = a.XT, YT = a.YT, KZ = a.KZ, hz = KZ == 3 ? 1 : 0, ZH = 128 + 2 * hz;
    // tile coordinates
    int t = blockIdx.x;
    const int ntz = a.Z / 128, nty = a.Y / YT, ntx = a.X / XT;
    const int tz = t % ntz; t /= ntz;
    const int ty = t % nty; t /= nty;
    const int tx = t % ntx; t /= ntx;
    const int b = t;
    const int x0 = tx * XT, y0 = ty * YT, z0 = tz * 128;
    const int nchunk = a.Cin / 16;
    const uint32_t slab = (uint32_t)ZH * 16;             // one (cg, line) slab
    const uint32_t lbo_a = (uint32_t)(YT + 2) * slab;     // cg -> cg+1
    const uint32_t b_tap = (uint32_t)a.Cout * 32;         // one tap of one plane: [2 khalf][Cout][8] bf16
    const uint32_t b_plane = (uint32_t)(3 * KZ) * b_tap;  // one plane of a slot

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.SA; ++i) {
            mbar_init(a_full + i, 1);
            mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int it = 0;
            for (int c = 0; c < nchunk; ++c) {
                for (int s = 0; s < XT + 2; ++s) {
                    if (s < 3) {  // weights of taps with dx = s for this channel slice
                        mbar_wait(b_empty + s, (c & 1) ^ 1);
                        mbar_expect_tx(b_full + s, a.b_slot);
                        bulk_load(b_ring + (size_t)s * a.b_slot, (const uint8_t*)a.w + ((size_t)c * 3 + s) * a.b_slot,
                                  a.b_slot, b_full + s);
                    }
                    const int px = x0 - 1 + s;
                    if (px < 0 || px >= a.X) continue;  // whole plane is padding: nothing to stage
                    const int st = it % a.SA;
                    mbar_wait(a_empty + st, ((it / a.SA) & 1) ^ 1);
                    mbar_expect_tx(a_full + st, 2 * a.a_box_bytes);
                    uint8_t* dst = a_ring + (size_t)st * a.a_stage;
                    const int cgi = b * a.cg_batch + c * 2;
                    tma_load_5d(dst, &tmap, a_full + st, 0, z0 - hz, y0 - 1, px, cgi);
                    tma_load_5d(dst + a.a_block, &tmap, a_full + st, 0, z0 - hz, y0 - 1, px, cgi + a.cg_plane);
                    ++it;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.Cout >> 3) << 17) | (8u << 24);
            uint32_t started = 0;  // bit per output line: accumulator already written
            int it = 0;
            for (int c = 0; c < nchunk; ++c) {
                for (int s = 0; s < XT + 2; ++s) {
                    const int px = x0 - 1 + s;
                    const bool plane_ok = px >= 0 && px < a.X;
                    uint32_t a_hi = 0, a_lo = 0;
                    int st = 0;
                    if (plane_ok) {
                        st = it % a.SA;
                        mbar_wait(a_full + st, (it / a.SA) & 1);
                        a_hi = smem_u32(a_ring + (size_t)st * a.a_stage);
                        a_lo = a_hi + a.a_block;
                    }
                    for (int dx = 0; dx < 3; ++dx) {
                        const int oxl = s - dx;  // local output x row fed by this plane through tap dx
                        if (oxl < 0 || oxl >= XT) continue;
                        if (oxl == 0) mbar_wait(b_full + dx, c & 1);  // first use of this slot in the slice
                        tc_fence_after();
                        if (plane_ok) {
                            const uint32_t bs = smem_u32(b_ring + (size_t)dx * a.b_slot);
                            for (int oy = 0; oy < YT; ++oy) {
                                const int line = oxl * YT + oy;
                                const uint32_t d_tmem = tmem_base + (uint32_t)(line * a.Cout);
                                for (int dy = 0; dy < 3; ++dy) {
                                    const int gy = y0 + oy + dy - 1;
                                    if (gy < 0 || gy >= a.Y) continue;  // padding row
                                    for (int dz = 0; dz < KZ; ++dz) {
                                        const uint32_t aoff = (uint32_t)(oy + dy) * slab + (uint32_t)dz * 16;
                                        const uint32_t boff = (uint32_t)(dy * KZ + dz) * b_tap;
                                        const uint64_t dah = make_desc(a_hi + aoff, lbo_a, 128);
                                        const uint64_t dal = make_desc(a_lo + aoff, lbo_a, 128);
                                        const uint64_t dbh = make_desc(bs + boff, (uint32_t)a.Cout * 16, 128);
                                        const uint64_t dbl = make_desc(bs + b_plane + boff, (uint32_t)a.Cout * 16, 128);
                                        const uint32_t acc0 = (started >> line) & 1u;
                                        umma_bf16(d_tmem, dah, dbh, idesc, acc0);
                                        umma_bf16(d_tmem, dal, dbh, idesc, 1u);
                                        umma_bf16(d_tmem, dah, dbl, idesc, 1u);
                                        started |= 1u << line;
                                    }
                                }
                            }
                        }
                        if (oxl == XT - 1) umma_commit(b_empty + dx);  // last use of this weight slot
                    }
                    if (plane_ok) {
                        umma_commit(a_empty + st);
                        ++it;
                    }
                }
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== epilogue: 4 warps, TMEM lanes (warp % 4) * 32 .. +31, thread = one z row =====
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int lane_base = (warp & 3) * 32;
        const int z = z0 + lane_base + lane;
        __nv_bfloat16* out_hi = (__nv_bfloat16*)a.out.hi;
        for (int line = 0; line < XT * YT; ++line) {
            const int ox = x0 + line / YT, oy = y0 + line % YT;
            float rsrc = 0.f;
            if (a.res_mode == 2) rsrc = a.rsrc.ptr[b * a.rsrc.sb + ox * a.rsrc.sx + oy * a.rsrc.sy + z * a.rsrc.sz];
            for (int c0 = 0; c0 < a.Cout; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(line * a.Cout + c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                    const int cc = c0 + g8 * 8;
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = __uint_as_float(r[g8 * 8 + j]) * __ldg(a.ep.scale + cc + j) + __ldg(a.ep.shift + cc + j);
                        o[j] = apply_act(v, a.ep.act, a.ep.slope);
                    }
                    if (a.res_mode == 1) {
                        const __nv_bfloat16* rp = (const __nv_bfloat16*)a.res.hi +
                                                  act8_off(a.res.batch_stride, a.X, a.Y, a.Z, b, cc / 8, ox, oy, z);
                        float rr[8];
                        unpack8(ldg128(rp), ldg128(rp + a.res.lo_offset), rr);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] += rr[j];
                    } else if (a.res_mode == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] += __ldg(a.res_w + cc + j) * rsrc + __ldg(a.res_b + cc + j);
                    }
                    uint4 h, l;
                    pack8(o, h, l);
                    __nv_bfloat16* p = out_hi + act8_off(a.out.batch_stride, a.X, a.Y, a.Z, b, cc / 8, ox, oy, z);
                    *reinterpret_cast<uint4*>(p) = h;
                    *reinterpret_cast<uint4*>(p + a.out.lo_offset) = l;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// Tile = as many output lines per CTA as TMEM (512 columns) and shared memory (>= 2 A stages next
// to the 3 weight slots) allow, preferring 2x4, then 2x2, 1x2, 1x1.
static bool plan_tile(const vsseg_act8* in, int cout, int kz, TcArgs* a) {
    const int cand[4][2] = {{2, 4}, {2, 2}, {1, 2}, {1, 1}};
    const int hz = kz == 3 ? 1 : 0, ZH = 128 + 2 * hz;
    a->b_slot = (uint32_t)(2 * 3 * kz * cout * 32);
    for (auto& c : cand) {
        if (in->X % c[0] || in->Y % c[1]) continue;
        if (c[0] * c[1] * cout > 512) continue;
        const uint32_t box = (uint32_t)(2 * (c[1] + 2) * ZH * 16);
        const uint32_t block = (box + 127) / 128 * 128;
        const long budget = 227L * 1024 - 1024 - 3L * a->b_slot;
        const int sa = (int)(budget / (2L * block));
        if (sa < 2) continue;
        a->XT = c[0];
        a->YT = c[1];
        a->a_box_bytes = box;
        a->a_block = block;
        a->a_stage = 2 * block;
        a->SA = sa > 8 ? 8 : sa;
        return true;
    }
    return false;
}

static bool tc_shape_ok(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g) {
    if (!in || !out || !g) return false;
    if (g->transposed || g->sx != 1 || g->sy != 1 || g->sz != 1) return false;
    if (g->kx != 3 || g->ky != 3 || (g->kz != 1 && g->kz != 3)) return false;
    if (in->C % 16 || out->C % 16 || out->C < 16 || out->C > 96) return false;
    if (in->Z % 128 || in->X != out->X || in->Y != out->Y || in->Z != out->Z || in->B != out->B) return false;
    const int64_t cgs = (int64_t)in->X * in->Y * in->Z * 8;
    if (in->lo_offset % cgs || in->batch_stride % cgs) return false;
    TcArgs tmp{};
    return plan_tile(in, out->C, g->kz, &tmp);
}

}  // namespace vsseg

using namespace vsseg;

extern "C" {

int vsseg_conv3d_tc_supported(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g) {
    return tc_shape_ok(in, out, g) ? 1 : 0;
}

int vsseg_conv3d_tc(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g, const void* w_packed,
                    const vsseg_epilogue* ep, const vsseg_act8* res_act8, const vsseg_f32view* res_src,
                    const float* res_w, const float* res_b, void* stream) {
    VSSEG_REQUIRE(tc_shape_ok(in, out, g), "conv3d_tc: unsupported shape (need stride-1 3x3x{1,3}, Cin,Cout %% 16 == 0, "
                                           "Cout <= 96, Z %% 128 == 0)");
    VSSEG_REQUIRE(w_packed && ep && ep->scale && ep->shift, "conv3d_tc: NULL weights/epilogue");
    VSSEG_REQUIRE(!(res_act8 && res_src), "conv3d_tc: at most one residual source");
    TcArgs a{};
    a.out = *out;
    a.ep = *ep;
    a.w = (const __nv_bfloat16*)w_packed;
    a.Cin = in->C; a.Cout = out->C; a.X = in->X; a.Y = in->Y; a.Z = in->Z; a.B = in->B; a.KZ = g->kz;
    if (res_act8) {
        VSSEG_REQUIRE(res_act8->hi && res_act8->C == out->C && res_act8->X == out->X && res_act8->Y == out->Y &&
                          res_act8->Z == out->Z && res_act8->B == out->B, "conv3d_tc: residual shape mismatch");
        a.res_mode = 1;
        a.res = *res_act8;
    } else if (res_src) {
        VSSEG_REQUIRE(res_src->ptr && res_w && res_b, "conv3d_tc: NULL cin1 residual");
        a.res_mode = 2;
        a.rsrc = *res_src; a.res_w = res_w; a.res_b = res_b;
    }
    VSSEG_REQUIRE(plan_tile(in, a.Cout, a.KZ, &a), "conv3d_tc: tile does not fit in shared memory");
    const int hz = a.KZ == 3 ? 1 : 0, ZH = 128 + 2 * hz;
    const int cols = a.XT * a.YT * a.Cout;
    a.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    const int64_t cgs = (int64_t)in->X * in->Y * in->Z * 8;
    a.cg_plane = (int)(in->lo_offset / cgs);
    a.cg_batch = (int)(in->batch_stride / cgs);

    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {8, (cuuint64_t)in->Z, (cuuint64_t)in->Y, (cuuint64_t)in->X,
                                (cuuint64_t)(a.cg_plane + (in->B - 1) * a.cg_batch + in->C / 8)};
    const cuuint64_t gstr[4] = {16, (cuuint64_t)in->Z * 16, (cuuint64_t)in->Y * in->Z * 16, (cuuint64_t)cgs * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)ZH, (cuuint32_t)(a.YT + 2), 1, 2};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    // the driver entry point is resolved through the runtime so the library has no link-time
    // dependency on libcuda.so (it must load on a GPU-less build box)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, in->hi, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("conv3d_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return (int)cr;
    }
    const size_t smem = 1024 + (size_t)a.SA * a.a_stage + 3 * (size_t)a.b_slot;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("conv3d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set = true;
    }
    const unsigned grid = (unsigned)((in->X / a.XT) * (in->Y / a.YT) * (in->Z / 128) * in->B);
    conv_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmap, a);
    return check_launch("conv3d_tc");
}

}  // extern "C"

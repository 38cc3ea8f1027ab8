// Shared device helpers for the vsseg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vsseg_b200.h"

namespace vsseg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VSSEG_REQUIRE(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            ::vsseg::set_error(__VA_ARGS__);     \
            return VSSEG_EINVAL;                 \
        }                                        \
    } while (0)

// ---- split-bf16 <-> fp32 ------------------------------------------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) {
    return __uint_as_float(bits16 << 16);
}

// 8 channels (one 16-byte group) of the hi and lo planes -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float (&v)[8]) {
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
    const uint32_t ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
        v[2 * i + 1] = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
    }
}

__device__ __forceinline__ void split1(float v, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    float r = v - __bfloat162float(h);
    __nv_bfloat16 l = __float2bfloat16_rn(r);
    hi = (uint32_t)__bfloat16_as_ushort(h);
    lo = (uint32_t)__bfloat16_as_ushort(l);
}

__device__ __forceinline__ void pack8(const float (&v)[8], uint4& h, uint4& l) {
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // two values per cvt: hi pair, residuals against the rounded values, lo pair
        const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        hh[i] = *reinterpret_cast<const uint32_t*>(&hp);
        const float r0 = v[2 * i] - __uint_as_float(hh[i] << 16);
        const float r1 = v[2 * i + 1] - __uint_as_float(hh[i] & 0xffff0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(r0, r1);
        ll[i] = *reinterpret_cast<const uint32_t*>(&lp);
    }
    h = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

__device__ __forceinline__ uint4 ldg128(const void* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// plain (coherent) 128-bit load for data this kernel also writes
__device__ __forceinline__ uint4 ldg128_plain(const void* p) {
    return *reinterpret_cast<const uint4*>(p);
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return 1.0f / (1.0f + expf(-v));
    return v >= 0.f ? v : v * slope;
}

// base pointer of an fp32 view: relocatable views (vsseg_f32view.indirect) keep a byte offset in `ptr` and the base
// address in a device cell, read here at run time (one captured graph serves every volume)
__device__ __forceinline__ float* f32_base(const vsseg_f32view& v) {
    if (v.indirect)
        return reinterpret_cast<float*>(__ldg(reinterpret_cast<const long long*>(v.indirect)) + reinterpret_cast<long long>(v.ptr));
    return v.ptr;
}
inline bool f32_set_ok(const vsseg_f32view* v) { return v && (v->ptr || v->indirect); }            // plain view or window set
inline bool f32_ok(const vsseg_f32view* v) { return f32_set_ok(v) && v->n_windows <= 1; }         // plain view
inline bool f32_direct(const vsseg_f32view* v) { return f32_ok(v) && v->ptr && !v->indirect; }

// Window set (vsseg_f32view.n_windows > 1, include/vsseg_b200.h): batch item b of the view is the window described by
// record b.  The kernels keep record 0 and this table of element offsets of the other windows from it.
struct WinTab {
    int32_t n;                            // 0: plain view, batch item b at b * sb
    int64_t off[VSSEG_MAX_WINDOWS];
};
// fills `t` from the records behind `v`; false when they do not form a window set of `batch` items
inline bool win_tab(const vsseg_f32view* v, int batch, WinTab* t) {
    t->n = 0;
    for (int i = 0; i < VSSEG_MAX_WINDOWS; ++i) t->off[i] = 0;
    if (!v || v->n_windows <= 1) return true;
    if (v->n_windows != batch || batch > VSSEG_MAX_WINDOWS) return false;
    for (int i = 0; i < batch; ++i) {
        const vsseg_f32view& r = v[i];
        if (r.B != 1 || r.C != v->C || r.X != v->X || r.Y != v->Y || r.Z != v->Z || r.sc != v->sc || r.sx != v->sx ||
            r.sy != v->sy || r.sz != v->sz || r.indirect != v->indirect || !f32_set_ok(&r))
            return false;
        const long long d = (long long)(reinterpret_cast<intptr_t>(r.ptr) - reinterpret_cast<intptr_t>(v->ptr));
        if (d % (long long)sizeof(float)) return false;
        t->off[i] = d / (long long)sizeof(float);
    }
    t->n = batch;
    return true;
}
__device__ __forceinline__ int64_t win_off(const vsseg_f32view& v, const WinTab& t, int b) {
    return t.n ? t.off[b] : (int64_t)b * v.sb;
}

// element offset of (b, cg, x, y, z) group start in an act8 plane
__device__ __forceinline__ int64_t act8_off(int64_t bstride, int X, int Y, int Z, int b, int cg, int x, int y, int z) {
    return (int64_t)b * bstride + ((((int64_t)cg * X + x) * Y + y) * Z + z) * 8;
}

}  // namespace vsseg

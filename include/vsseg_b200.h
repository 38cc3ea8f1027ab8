/*
 * vsseg_b200.h — C ABI of the B200-native VS_Seg hot path (libvsseg_b200.so).
 *
 * The reference (KCL-BMEIS/VS_Seg) is pure Python: it has no FFI/plugin boundary of its own;
 * the arithmetic it runs lives behind torch/cuDNN/MONAI call sites.  Each entry point below
 * replaces one of those call sites (cited as file:line under /root/reference) and is what a
 * ctypes binding on the reference side would call.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions are asynchronous on `stream`, never synchronise the device, and return 0 on
 *     success or a non-zero code (cudaError_t value, or VSSEG_EINVAL for bad arguments);
 *     vsseg_last_error() gives the message for the calling thread;
 *   - activations inside the network use the "act8" layout described at vsseg_act8.
 */
#ifndef VSSEG_B200_H
#define VSSEG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSSEG_EINVAL 100001
#define VSSEG_ABI_VERSION 1

/*
 * act8: channel-blocked split-bf16 activation tensor.
 *   logical tensor [B, C, X, Y, Z] (reference layout NCDHW, Z contiguous), C % 8 == 0
 *   hi plane: bf16 [B][C/8][X][Y][Z][8]          (value rounded to bf16)
 *   lo plane: bf16, same shape, at hi + lo_offset  (bf16 of the rounding residual)
 *   value = float(hi) + float(lo)  (16 mantissa bits; what "bf16x3" tensor-core passes consume)
 * A view into a wider buffer (the concat of a skip connection) is a pointer to its first
 * channel group plus the parent's batch_stride: torch.cat (MONAI SkipConnection,
 * reference unet2d5_spvPA.py:89) costs nothing.
 */
typedef struct {
    void*   hi;            /* bf16*: element (b=0, first channel group of the view) of the hi plane */
    int64_t lo_offset;     /* elements from a hi element to its lo element */
    int64_t batch_stride;  /* elements between consecutive batch items */
    int32_t B, C, X, Y, Z; /* C = channels of this view (multiple of 8) */
} vsseg_act8;

/* fp32 tensor addressed by strides (elements): used for the 1-channel network input, read in
 * place from the full volume (the sliding-window gather is free), and for planar outputs. */
typedef struct {
    float*  ptr;           /* element (b=0, c=0, x=0, y=0, z=0) of the region */
    int64_t sb, sc, sx, sy, sz;
    int32_t B, C, X, Y, Z;
} vsseg_f32view;

/* Per-output-channel epilogue y = act(acc * scale[c] + shift[c]):
 * eval-mode BatchNorm3d + conv bias folded (reference convolutions.py:148-156);
 * act 0: leaky with `slope` (PReLU; slope 0 = ReLU; slope 1 = identity), 1: sigmoid. */
typedef struct {
    const float* scale;    /* [CoutPad] */
    const float* shift;    /* [CoutPad] */
    float   slope;
    int32_t act;
} vsseg_epilogue;

typedef struct {
    int32_t kx, ky, kz;    /* kernel size; "same" padding (k-1)/2 as reference convolutions.py:85 */
    int32_t sx, sy, sz;    /* stride */
    int32_t transposed;    /* 1: ConvTranspose3d with output_padding = stride-1 (convolutions.py:114-135) */
} vsseg_conv_geom;

/* library / device ------------------------------------------------------------------------- */
int         vsseg_abi_version(void);
const char* vsseg_last_error(void);
int         vsseg_device_sm_count(int device, int* sm_count_host);

/* layout conversion at the module boundary (NCDHW fp32 <-> act8) --------------------------- */
int vsseg_pack_act8(const vsseg_f32view* src, const vsseg_act8* dst, void* stream);
int vsseg_unpack_act8(const vsseg_act8* src, const vsseg_f32view* dst, void* stream);

/*
 * Fused Convolution block, eval mode: out = act(BN(conv(in))) [+ residual].
 * Replaces torch Conv3d/ConvTranspose3d + BatchNorm3d + Dropout(eval) + PReLU/ReLU
 * (reference convolutions.py:125-156) and, with a residual, the ResidualUnit sum
 * (convolutions.py:252-255).
 *   w: fp32 [taps][Cin][CoutPad] (tap index = (tx*ky + ty)*kz + tz), CoutPad = round_up(Cout,16)
 *   residual (optional, added AFTER the activation):
 *     res_act8 != NULL : addend tensor with Cout channels (a precomputed 1x1x1 shortcut)
 *     res_src  != NULL : 1-channel shortcut conv computed in place: + res_w[c]*src[v] + res_b[c]
 */
int vsseg_conv3d_act8(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                      const float* w, int32_t cout_pad, const vsseg_epilogue* ep,
                      const vsseg_act8* res_act8,
                      const vsseg_f32view* res_src, const float* res_w, const float* res_b,
                      void* stream);

/*
 * Tensor-core path of the same fused block for the FLOP-heavy layers: tcgen05.mma (bf16x3 on the
 * hi/lo planes, fp32 TMEM accumulators), TMA-staged haloed tiles, same epilogue/residual contract
 * as vsseg_conv3d_act8.  Supported: stride 1, kernel 3x3x{1,3}, Cin and Cout multiples of 16,
 * Cout <= 96, Z a multiple of 128 (vsseg_conv3d_tc_supported returns 1).
 *   w_packed: bf16 [Cin/16][dx 3][plane hi,lo][dy 3][dz KZ][khalf 2][Cout][8]
 *             (element = weight[cout][cin = 16*c + 8*khalf + j][dx][dy][dz], split like act8)
 */
int vsseg_conv3d_tc_supported(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g);
int vsseg_conv3d_tc(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                    const void* w_packed, const vsseg_epilogue* ep,
                    const vsseg_act8* res_act8,
                    const vsseg_f32view* res_src, const float* res_w, const float* res_b,
                    void* stream);

/* First encoder conv: 1-channel fp32 input (read in place from the volume) -> act8.
 * Replaces model.0.conv.unit0 (Conv3d(1,16,(3,3,1)) + BN + PReLU).  w: fp32 [taps][Cout]. */
int vsseg_conv3d_cin1(const vsseg_f32view* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                      const float* w, const vsseg_epilogue* ep, void* stream);

/* Conv with 1 or 2 output channels -> planar fp32 (attention conv2 + Sigmoid,
 * reference attentionblock.py:21-30; the top ResidualUnit's conv_only unit + shortcut, whose
 * 1x1x1 shortcut is folded into the centre tap, unet2d5_spvPA.py:186-190).
 *   w: fp32 [taps][Cin][Cout]; bias: [Cout]; act as in vsseg_epilogue.act (slope 1 = none).
 *   If sw_weight != NULL the result is blended into `out` instead of stored:
 *     out[b,c,v] += sw_weight[v] * y[b,c,v]      (MONAI sliding_window_inference step 6;
 *   reference call site VSparams.py:568-574); sw_weight is the [X,Y,Z] importance map. */
int vsseg_conv3d_smallcout(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g,
                           const float* w, const float* bias, int32_t act, float slope,
                           const float* sw_weight, void* stream);

/* Attention gate, AttentionBlock2: out = x * (1 + att) (reference attentionblock.py:44-47).
 * att: planar fp32 [B,1,X,Y,Z]; x/out act8 with the same shape (may alias). */
int vsseg_att_gate(const vsseg_act8* x, const vsseg_f32view* att, const vsseg_act8* out, void* stream);

/* Sliding-window finalise: prob = acc / cnt (MONAI step 7), optional argmax mask (uint8) and
 * hard-Dice partial sums vs label (VSparams.compute_dice_score, VSparams.py:393-408):
 * sums[0] += |pred&label|, sums[1] += |label|, sums[2] += |pred| (fp64 accumulators).
 * acc/out: [C,n] planar; cnt: [n]; label: [n] float 0/1 or NULL; mask: [n] or NULL. */
int vsseg_sw_finalize(const float* acc, const float* cnt, float* out, int32_t C, int64_t n,
                      uint8_t* mask, const float* label, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSSEG_B200_H */

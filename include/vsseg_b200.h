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
#define VSSEG_ABI_VERSION 2

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
 * place from the full volume (the sliding-window gather is free), and for planar outputs.
 * `indirect` (optional) makes the view relocatable: it is the DEVICE address of an 8-byte cell that holds
 * the base address of the volume, and `ptr` is then the BYTE OFFSET of the region from that base.  The
 * kernel reads the cell at run time, so a captured CUDA graph of a whole sliding-window schedule
 * (MONAI sliding_window_inference, reference call site VSparams.py:568-574) is re-targeted to another
 * volume by one 8-byte store.  Honoured by the forward kernels (vsseg_conv3d_cin1, vsseg_conv3d_tc*,
 * vsseg_conv3d_act8, vsseg_conv3d_smallcout, vsseg_conv3d_gate_logits, vsseg_att_gate); the layout
 * conversion and training entry points reject it.
 * `n_windows` > 1 makes the record the first of a WINDOW SET: n_windows contiguous vsseg_f32view records, each one
 * window (B == 1) of the same volume (identical strides, extents and `indirect`), that together stand for a batch
 * of n_windows items - batch item b is record b.  The windows of a sliding-window group (MONAI
 * sliding_window_inference slices its windows out of one padded volume, call site VSparams.py:568-574) are read in
 * place by ONE launch this way.  Honoured where stated (vsseg_conv3d_cin1 `in`, vsseg_conv3d_tc `res_src`);
 * everything else requires n_windows <= 1. */
#define VSSEG_MAX_WINDOWS 16
typedef struct {
    float*  ptr;           /* element (b=0, c=0, x=0, y=0, z=0) of the region (byte offset if indirect) */
    int64_t sb, sc, sx, sy, sz;
    int32_t B, C, X, Y, Z;
    int32_t n_windows;     /* 0 or 1: a plain view; n > 1: first record of a window set (see above) */
    const int64_t* indirect; /* NULL, or device cell holding the base address */
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
 * Tensor-core path of the same fused block: tcgen05.mma (bf16x3 on the hi/lo planes, fp32 TMEM
 * accumulators), TMA-staged input boxes, same epilogue/residual contract as vsseg_conv3d_act8.
 * Covers Conv3d (stride 1 or 2 per axis, kernel 1 or 3 per axis) and ConvTranspose3d (stride (2,2,1|2),
 * kernel (3,3,1|3), output_padding = stride-1; reference convolutions.py:114-135) with Cin % 16 == 0,
 * Cout % 8 == 0; the M tile (128 positions of the output grid; input grid when transposed) is LY y lines x LZ z
 * with LZ the largest power of two (<= 128) dividing the z extent.  vsseg_conv3d_tc_supported returns 1 when the shape is
 * covered; vsseg_conv3d_tc_suggest_split returns the n_split (CTAs per M tile along Cout) that fills
 * the chip for small layers, or 0 when the shape is not covered.
 *
 *   w_packed: bf16 [sel][Cin/16][j][plane hi,lo][tz][khalf 2][ty'][n_cta][8], n_cta = round_up(Cout,16)/n_split
 *     Conv3d:          sel = n-slice, j = kernel x index, ty' = ky-1-ty (reversed kernel y index, so that the
 *                      taps sharing one staged input line are adjacent N rows of a single MMA), tz = kernel z,
 *                      element = W[cout = sel*n_cta + n][cin = 16*c + 8*khalf + e][j][ty][tz]
 *     ConvTranspose3d: sel = px*n_split + n-slice (px = output x parity), j in {0,1} = input x shift, ty' = ty,
 *                      element = W[cin][cout][kx][ty][tz] with kx = 1 (px=0, j=0; j=1 unused, zeros),
 *                      kx = 2 (px=1, j=0), kx = 0 (px=1, j=1)
 *     each value split hi/lo like act8; channels >= Cout are zero.
 *   shortcut_src (optional): input of a fused 1x1x1 ResidualUnit shortcut conv
 *     (convolutions.py:241-255) accumulated in a second TMEM accumulator and added after the
 *     activation: out = act(BN(conv(in))) + shortcut_w * shortcut_src + shortcut_bias.
 *     shortcut_w: bf16 [n-slice][Csrc/16][plane][khalf][n_cta][8]; stride-1 convs only.
 *   res_src may be a window set (vsseg_f32view.n_windows == out->B): batch item b adds the affine map of window b.
 */
int vsseg_conv3d_tc_supported(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                              int32_t n_split, const vsseg_act8* shortcut_src);
int vsseg_conv3d_tc_suggest_split(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                                  const vsseg_act8* shortcut_src);
/* Tensor-core conv with 1 or 2 output channels written as planar fp32 (attention conv2 + Sigmoid,
 * reference attentionblock.py:21-30; the top ResidualUnit's conv_only unit with its 1x1x1 shortcut
 * folded into the centre tap, unet2d5_spvPA.py:186-190).  Same contract as vsseg_conv3d_smallcout:
 * with sw_weight the result is blended into `out` (out += sw_weight[v] * y, MONAI
 * sliding_window_inference step 6, reference call site VSparams.py:568-574).
 * w_packed: as vsseg_conv3d_tc with n_split = 1 and Cout zero-padded to 16; ep->scale/shift: [16]. */
int vsseg_conv3d_tc_f32out(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g,
                           const void* w_packed, const vsseg_epilogue* ep, const float* sw_weight, void* stream);
/* Same operation with the "two-pass" weight image: with C = Cout <= 2 real output channels the 16 weight
 * columns of plane 0 hold [bf16 hi(W) (C columns) | bf16 lo(W) (C columns) | 0] and plane 1 holds
 * [bf16 hi(W) | 0].  The kernel then needs two MMAs per product instead of three
 * (A_hi x plane 0, A_lo x plane 1) and the epilogue adds column C+q to column q: same bf16x3 value. */
int vsseg_conv3d_tc_f32out_2p(const vsseg_act8* in, const vsseg_f32view* out, const vsseg_conv_geom* g,
                              const void* w_packed, const vsseg_epilogue* ep, const float* sw_weight, void* stream);

/* Attention conv2 + Sigmoid (two-pass weight image, Cout = 1) with AttentionBlock2 fused into its epilogue
 * (reference attentionblock.py:21-47): the map is stored to `att` and every channel of `gated` (the tensor
 * the block gates, same extents) is scaled in place by 1 + att.  One launch instead of conv2 + vsseg_att_gate. */
int vsseg_conv3d_tc_attgate(const vsseg_act8* in, const vsseg_f32view* att, const vsseg_conv_geom* g,
                            const void* w_packed, const vsseg_epilogue* ep, const vsseg_act8* gated, void* stream);

/* Human-readable description of the launch plan (tile, stages, op table) for tests and DESIGN.md. */
int vsseg_conv3d_tc_describe(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                             int32_t n_split, const vsseg_act8* shortcut_src, char* buf, int32_t buflen);
int vsseg_conv3d_tc(const vsseg_act8* in, const vsseg_act8* out, const vsseg_conv_geom* g,
                    const void* w_packed, int32_t n_split, const vsseg_epilogue* ep,
                    const vsseg_act8* res_act8,
                    const vsseg_f32view* res_src, const float* res_w, const float* res_b,
                    const vsseg_act8* shortcut_src, const void* shortcut_w, const float* shortcut_bias,
                    void* stream);

/* First encoder conv: 1-channel fp32 input (read in place from the volume) -> act8.
 * Replaces model.0.conv.unit0 (Conv3d(1,16,(3,3,1)) + BN + PReLU).  w: fp32 [taps][Cout].
 * `in` may be a window set (vsseg_f32view.n_windows == out->B): every window of a sliding-window group in one launch. */
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

/* Top of the decoder in one bandwidth-bound launch: AttentionBlock2 gate x*(1+att) (reference
 * attentionblock.py:44-47) applied on the fly to the input of the last ResidualUnit (conv_only unit + 1x1x1
 * shortcut folded into the centre tap, unet2d5_spvPA.py:186-190; kernel (3,3,1), Cout 1 or 2), so the gated
 * 2c-channel tensor is never written: out = conv(x * (1 + att)) + bias, stored or, with sw_weight, blended
 * into the sliding-window accumulator (out += sw_weight * y, MONAI step 6, VSparams.py:568-574).
 *   x: act8 [B,Cin,X,Y,Z], Cin % 8 == 0, Cin <= 32;  att: [B,1,X,Y,Z] or NULL (no gate)
 *   w_host: HOST fp32 [9 taps (tx*3+ty)][Cin][Cout]; bias_host: HOST [Cout] (passed to the kernel by value)
 *   outs: n_outs views [*,Cout,X,Y,Z]; n_outs == 1: one view covering all B entries, n_outs == B: entry b -> outs[b]
 *   (the windows of a sliding-window group land at different offsets of the accumulator).
 *   atomic_blend: blend with red.global.add.f32.  REQUIRED when the destination regions of one launch overlap
 *   (several overlapping windows in one launch); the sum order is then unspecified.  With 0 the blend is a plain
 *   read-modify-write: one window per launch, launches ordered on the stream, i.e. MONAI's accumulation order. */
int vsseg_conv3d_gate_logits(const vsseg_act8* x, const vsseg_f32view* att, const float* w_host,
                             const float* bias_host, int32_t cout, const vsseg_f32view* outs, int32_t n_outs,
                             const float* sw_weight, int32_t atomic_blend, void* stream);

/* Attention gate, AttentionBlock2: out = x * (1 + att) (reference attentionblock.py:44-47).
 * att: planar fp32 [B,1,X,Y,Z]; x/out act8 with the same shape (may alias). */
int vsseg_att_gate(const vsseg_act8* x, const vsseg_f32view* att, const vsseg_act8* out, void* stream);

/* Sliding-window finalise: prob = acc / cnt (MONAI step 7; cnt == NULL: prob = acc, no division), optional
 * argmax mask (uint8) and hard-Dice partial sums vs label (VSparams.compute_dice_score, VSparams.py:393-408:
 * argmax -> one_hot -> 1 - DiceLoss(include_background=False)):
 * sums[0] += |pred&label|, sums[1] += |label|, sums[2] += |pred| (fp64 accumulators), pred = (argmax == 1).
 * acc/out: [C,n] planar (out may be NULL); cnt: [n]; label: [n] float 0/1 (label_u8 = 0) or uint8 (label_u8 = 1),
 * or NULL; mask: [n] or NULL.  128-bit accesses when n % 4 == 0 and the pointers are 16-byte aligned. */
int vsseg_sw_finalize(const float* acc, const float* cnt, float* out, int32_t C, int64_t n,
                      uint8_t* mask, const void* label, int32_t label_u8, double* sums, void* stream);

/* ---- multi-GPU sliding window over peer memory (one node, one process per GPU) -----------------
 * The reference has no multi-GPU path; windows are independent (VSparams.py:556-574: eval mode, sw_batch_size 1), so
 * rank r runs a contiguous block of the window list and blends it with atomic adds (vsseg_conv3d_gate_logits,
 * atomic_blend = 1) into the accumulator of the destination rank, mapped into every process with CUDA IPC.
 *   vsseg_peer_alloc   cudaMalloc + zero fill + IPC handle (64 bytes, HOST buffer) of an allocation peers may map
 *   vsseg_peer_open    maps a peer's allocation into this process (peer access enabled lazily); *ptr_host = device pointer
 *   vsseg_peer_close / vsseg_peer_free   unmap / release
 *   vsseg_flag_set     one-thread kernel: system-scope fence, then *flag = value (flag may live in peer memory): "all my
 *                      blends of volume k are done" / "buffer released"
 *   vsseg_flag_wait    one-thread kernel: spins until every flags[0..n) >= target (acquire, system scope); gives up after
 *                      timeout_cycles SM clocks and sets *err (DEVICE int, may be NULL) instead of hanging the GPU */
int vsseg_peer_alloc(int64_t bytes, void** ptr_host, void* handle64_host);
int vsseg_peer_open(const void* handle64_host, void** ptr_host);
int vsseg_peer_close(void* ptr);
int vsseg_peer_free(void* ptr);
int vsseg_flag_set(int64_t* flag, int64_t value, void* stream);
int vsseg_flag_wait(const int64_t* flags, int32_t n, int64_t target, int64_t timeout_cycles, int32_t* err, void* stream);

/* ---- Dice_spvPA loss (reference params/losses/dice_spvPA.py:90-167, :238-297) -----------------------
 * All tensors planar fp32, contiguous.  The loss is assembled from "terms": the 2-class logits term
 * (softmax + one-hot + hardness weight w = lambda*|p - t| + 1 - lambda, gradient flowing through w) and
 * one single-channel term per attention level against the max-pooled label.
 *   vsseg_maxpool3d      label pyramid G_{l+1} = MaxPool3d(kernel = stride = ratio)(G_l)   (:268-277)
 *   vsseg_dice_sums      C == 1: sums[b][0..2]    += (sum a*g, sum g, sum a)               (:136-149)
 *                        C == 2: sums[b][c][0..2] += (sum w t_c p_c, sum w t_c, sum w p_c), p = softmax(logits),
 *                                t = one_hot(label); hardness_lambda < 0 disables the weight (w = 1)
 *   vsseg_dice_finalize  loss = sum_r row_scale[r] * (1 - (2I+smooth)/(G+P+smooth))         (:156-159, :295-297)
 *                        coef[r] = row_scale[r] * (-2/D, (2I+smooth)/D^2), D = G+P+smooth   (backward factors)
 *   vsseg_dice_backward  C == 1: grad = go * (alpha*g + beta); C == 2: d loss / d logits through the softmax
 *                        Jacobian and the hardness weight; grad_out is the DEVICE scalar d(total)/d(loss).
 * sums must be zeroed by the caller; fp64 accumulators. */
int vsseg_maxpool3d(const float* in, float* out, int32_t B, int32_t Xo, int32_t Yo, int32_t Zo,
                    int32_t rx, int32_t ry, int32_t rz, void* stream);
int vsseg_dice_sums(const float* pred, const float* target, int32_t B, int32_t C, int64_t n,
                    float hardness_lambda, double* sums, void* stream);
int vsseg_dice_finalize(const double* sums, const float* row_scale, int32_t nrows, float smooth,
                        float* loss, float* coef, void* stream);
int vsseg_dice_backward(const float* pred, const float* target, int32_t B, int32_t C, int64_t n,
                        float hardness_lambda, const float* coef, const float* grad_out, float* grad,
                        void* stream);

/* General DiceLoss.forward (reference dice_spvPA.py:90-167) for up to 8 channels, every constructor flag:
 *   act 0 none | 1 sigmoid | 2 softmax over the channels (:105-113); target_is_labels: target is [B,1,n] integer labels
 *   turned into a one-hot (:114-118), else dense [B,C,n]; weight: optional hardness weight [B,C,n] (:136-149);
 *   squared: squared_pred (:141-143).
 *   vsseg_dice_general_sums      sums[b][c][0..2] += (sum w t p, sum w t', sum w p')  (fp64, stride 8 channels per batch
 *                                entry: sums is [B][8][3], zeroed by the caller); include_background / jaccard / reduction
 *                                act on these B x C numbers and stay with the caller
 *   vsseg_dice_general_backward  grad[b][c][v] = d/d pred of sum_c (gI_c I_c + gP_c P_c) through the activation;
 *                                grad_sums float [B][8][3] = (gI, gG, gP); the weight is treated as a constant */
int vsseg_dice_general_sums(const float* pred, const float* target, const float* weight, int32_t B, int32_t C, int64_t n,
                            int32_t act, int32_t target_is_labels, int32_t squared, double* sums, void* stream);
int vsseg_dice_general_backward(const float* pred, const float* target, const float* weight, int32_t B, int32_t C,
                                int64_t n, int32_t act, int32_t target_is_labels, int32_t squared,
                                const float* grad_sums, float* grad, void* stream);

/* ---- training mode (reference convolutions.py:148-156 Conv -> BatchNorm3d -> Dropout -> PReLU, autograd) -----
 * Convolutions (forward and data gradient) use vsseg_conv3d_tc: the data gradient of a Conv3d is the
 * ConvTranspose3d with the same weight tensor and vice versa.  The entry points below are the pieces around it.
 *   stats layout: float [4][C] = scale (gamma*rstd) | shift (beta - mean*scale) | batch mean | rstd
 *   vsseg_bn_stats           sums[0][C] += sum x, sums[1][C] += sum x^2 over (B,X,Y,Z)      (fp64, zeroed by the caller)
 *   vsseg_bn_finalize        batch mean / BIASED variance -> stats; running_mean/var <- (1-m)*old + m*(mean / UNBIASED var)
 *   vsseg_bn_act_fwd         y = PReLU_slope(dropout_p(c*scale + shift)) [+ residual]; dropout mask = hash(seed, element)
 *   vsseg_bn_act_bwd_reduce  sums[0][C] += sum du, sums[1][C] += sum du*xhat, sums[2C] += d(slope)   (du = grad after
 *                            PReLU' and the dropout mask)
 *   vsseg_bn_act_bwd_apply   dc = scale * (du - mean(du) - xhat*mean(du*xhat))
 *   vsseg_act_bwd            dc = dy * (y > 0 ? 1 : slope)                                   (attention conv1 ReLU)
 *   vsseg_act8_add           out = a + b                                                     (fan-out gradient sums)
 *   vsseg_conv3d_wgrad       dw[tap][Cin][cout_pad] += sum x[v_in]*dc[v_out], dbias[c] += sum dc  (fp32 atomics; both
 *                            zeroed by the caller; geometry as the forward conv, transposed included)
 *   vsseg_conv3d_cin1_wgrad  same for the 1-channel fp32 source of the first conv / its 1x1x1 shortcut; dw[tap][Cout]
 *   vsseg_conv3d_smallcout_bwd  backward of vsseg_conv3d_smallcout (Cout 1|2): dz = dy*(sigmoid ? y(1-y) : 1);
 *                            dx (act8, optional), dw[tap][Cin][Cout] and dbias (optional, accumulated)
 *   vsseg_att_gate_bwd       g = x*(1+att): dx = dg*(1+att) (+ dx if accumulate_dx), datt = sum_c dg_c*x_c */
int vsseg_bn_stats(const vsseg_act8* x, double* sums, void* stream);
int vsseg_bn_finalize(const double* sums, int32_t C, int64_t count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, float* stats,
                      void* stream);
/* slope: DEVICE pointer to the PReLU parameter (act.weight, one element): read by the kernel, so a training step never
 * copies parameters to the host and can be captured in a CUDA graph.
 * seed_base: optional DEVICE scalar added to `seed` (the per-layer salt) on the device: a captured graph draws fresh
 * dropout masks on every replay once the host stores a new base value before launching it. */
int vsseg_bn_act_fwd(const vsseg_act8* c, const vsseg_act8* y, const float* stats, const float* slope, float drop_p,
                     uint64_t seed, const uint64_t* seed_base, const vsseg_act8* residual, void* stream);
int vsseg_bn_act_bwd_reduce(const vsseg_act8* c, const vsseg_act8* dy, const float* stats, const float* slope,
                            float drop_p, uint64_t seed, const uint64_t* seed_base, double* sums, void* stream);
int vsseg_bn_act_bwd_apply(const vsseg_act8* c, const vsseg_act8* dy, const float* stats, const double* sums,
                           const float* slope, float drop_p, uint64_t seed, const uint64_t* seed_base,
                           const vsseg_act8* dc, void* stream);
int vsseg_act_bwd(const vsseg_act8* y, const vsseg_act8* dy, float slope, const vsseg_act8* dc, void* stream);
int vsseg_act8_add(const vsseg_act8* a, const vsseg_act8* b, const vsseg_act8* out, void* stream);
int vsseg_conv3d_wgrad(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g, float* dw,
                       int32_t cout_pad, float* dbias, void* stream);
/* Tensor-core weight gradient (tcgen05, MN-major operands read straight from the act8 z lines, bf16x3, split-K with
 * fp32 atomics) for stride-1 convs whose z extent is a multiple of 128, Cin % 16 == 0, Cout <= 128; dw as
 * vsseg_conv3d_wgrad (zeroed by the caller).  The bias gradient is vsseg_bn_stats(dc) (its plain sums). */
int vsseg_conv3d_wgrad_tc_supported(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g);
int vsseg_conv3d_wgrad_tc(const vsseg_act8* x, const vsseg_act8* dc, const vsseg_conv_geom* g, float* dw,
                          int32_t cout_pad, void* stream);
int vsseg_conv3d_cin1_wgrad(const vsseg_f32view* src, const vsseg_act8* dc, const vsseg_conv_geom* g,
                            float* dw, float* dbias, void* stream);
int vsseg_conv3d_smallcout_bwd(const vsseg_act8* x, const vsseg_f32view* dy, const vsseg_f32view* y,
                               const vsseg_conv_geom* g, const float* w, int32_t sigmoid,
                               const vsseg_act8* dx, float* dw, float* dbias, void* stream);
int vsseg_att_gate_bwd(const vsseg_act8* x, const vsseg_f32view* att, const vsseg_act8* dg,
                       const vsseg_act8* dx, const vsseg_f32view* datt, int32_t accumulate_dx, void* stream);

/* ---- optimizer (reference params/VSparams.py:388-391, :457-463: torch.optim.Adam(lr, weight_decay) stepped
 * once per batch over 178 tensors).  One launch over flat fp32 buffers of n elements (16-byte aligned):
 *   g = grad * grad_scale + weight_decay * param;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   param -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)        (torch.optim.Adam, amsgrad off)
 * grad_scale = 1 / world_size folds the data-parallel gradient average into the step.
 * step_dev / lr_dev (both or neither; DEVICE scalars): when given, the step count and learning rate are read on the
 * device instead of `step` / `lr`, so a captured CUDA graph of the training step stays valid as they change. */
int vsseg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    const int64_t* step_dev, const float* lr_dev, void* stream);

/* Builds the weight image vsseg_conv3d_tc streams (layout above) from a torch-layout fp32 weight on the device:
 * w is [d0][d1][kx][ky][kz]; conv_t_layout = 0: Conv3d [Cout][Cin] / 1: ConvTranspose3d [Cin][Cout]; phases = 1: the
 * sub-pixel phase image of a stride-2 transposed conv (2 * n_split selections, x taps {1,-} and {2,0}, y unflipped);
 * flip = 1: all three tap axes reversed (with conv_t_layout = 1 this is the adjoint of a stride-1 Conv3d, i.e. its
 * data gradient, reference convolutions.py:125-156 under autograd).  Channels beyond the real ones are zero.
 * out_bf16: 2 * n_sel * (cin_pad/16) * nj * kz * 2 * ky * (cout_pad/n_split) * 8 bf16 values. */
int vsseg_pack_conv_weight_tc(const float* w, int32_t d0, int32_t d1, int32_t kx, int32_t ky, int32_t kz,
                              int32_t conv_t_layout, int32_t phases, int32_t flip, int32_t cin_pad, int32_t cout_pad,
                              int32_t n_split, void* out_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSSEG_B200_H */

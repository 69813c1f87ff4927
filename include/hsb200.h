/*
 * hsb200.h -- C ABI of libhsb200.so: the B200 (sm_100a) implementation of HyperSeg's
 * decoder hot path (per-patch dynamic convolutions + the signal->weights heads).
 *
 * The reference (YuvalNirkin/hyperseg) has no FFI of its own: its hot path is Python
 * calling ATen.  Each entry point below therefore names the reference Python function
 * whose arithmetic it replaces (paths relative to the reference repo root).  The host
 * side (hyperseg_b200/nn/*.py) keeps the reference's nn.Module surface and calls these
 * through ctypes; see INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer owned by the caller (inputs, outputs,
 *     workspace); the library allocates nothing and keeps no state between calls.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises.
 *   - return value: 0 on success, a negative hsb_status otherwise; the text of the last
 *     failure on the calling thread is available from hsb_last_error().
 *   - there is no CPU path: a NULL/host pointer or a missing device is an error.
 *   - feature maps are NCHW, contiguous.  `dtype` applies to x, w and y alike
 *     (accumulation is always fp32).
 *   - per-patch weight tensors come in one of two layouts (hsb_wlayout):
 *       HSB_W_NCHW         the reference layout (B, hp, fh, fw), contiguous
 *       HSB_W_PATCH_MAJOR  (B, fh, fw, row) with `w_row_stride` elements between
 *                          consecutive patches (>= hp); this is what
 *                          hsb_signal2weights_fwd emits and what torch calls
 *                          channels_last for a (B, hp, fh, fw) tensor.
 *     Patch (b, i, j) is patch index p = (b*fh + i)*fw + j.
 */
#ifndef HSB200_H_
#define HSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSB200_VERSION 100 /* major*100 + minor */

typedef enum hsb_status {
    HSB_OK = 0,
    HSB_ERR_INVALID_ARG = -1,  /* bad shape / null pointer / misaligned buffer          */
    HSB_ERR_UNSUPPORTED = -2,  /* legal in the reference but not implemented by kernels */
    HSB_ERR_CUDA = -3,         /* a CUDA runtime call or launch failed                  */
    HSB_ERR_NO_DEVICE = -4     /* no sm_100 device visible                              */
} hsb_status;

typedef enum hsb_dtype { HSB_F32 = 0, HSB_BF16 = 1 } hsb_dtype;
typedef enum hsb_act { HSB_ACT_NONE = 0, HSB_ACT_RELU = 1, HSB_ACT_RELU6 = 2,
                       HSB_ACT_SILU = 3 /* encoder epilogues only */ } hsb_act;
typedef enum hsb_wlayout { HSB_W_NCHW = 0, HSB_W_PATCH_MAJOR = 1 } hsb_wlayout;
typedef enum hsb_padmode { HSB_PAD_ZEROS = 0, HSB_PAD_REFLECT = 1, HSB_PAD_REPLICATE = 2,
                           HSB_PAD_CIRCULAR = 3 } hsb_padmode;

/* Library / build introspection. */
int hsb_version(void);
const char* hsb_last_error(void);
/* Name of the CUDA kernel the last successful entry point of the calling thread launched ("" before the first call).
 * Entry points that choose between kernels (tensor-core vs CUDA-core paths) record their choice here, so a caller --
 * and the parity tests -- can assert which one ran. */
const char* hsb_last_kernel(void);
/* Writes the number of SMs and the compute capability (major*10+minor) of the current
 * device; HSB_ERR_NO_DEVICE when there is none. */
int hsb_device_info(int* sm_count, int* compute_capability);

/*
 * Patch-wise 1x1 convolution with a fused per-channel affine + activation epilogue.
 *   y[b,o,i*ph+u,j*pw+v] = act( post_scale[o] * sum_c Wm[o,c] * x[b, g*Cin/G + c, ...] + post_shift[o] )
 *   Wm[o,c] = w_patch[o*(Cin/G) + c],  g = o / (Cout/G),  ph = H/fh, pw = W/fw.
 * Replaces HyperPatchNoPadding.forward (hyperseg/models/hyperseg_v1_0.py:486-498,
 * hyperseg_v1_0_unify.py:483-494) and, through post_scale/post_shift/act, the eval-mode
 * BatchNorm2d + ReLU that make_hyper_patch_conv2d_block appends (hyperseg_v1_0.py:753-756).
 * post_scale/post_shift may both be NULL (identity epilogue).  H%fh==0 and W%fw==0.
 */
int hsb_patch_conv1x1_fwd(const void* x, const void* w, void* y,
                          const float* post_scale, const float* post_shift, int act,
                          int B, int Cin, int Cout, int H, int W, int fh, int fw, int groups,
                          int dtype, int w_layout, int64_t w_row_stride, void* stream);

/*
 * Fused patch-wise inverted-residual MetaBlock (v1_0 / unify semantics: tile-local halo).
 *   tile = reflect_pad(x,1)[b, :, i*ph : i*ph+ph+2, j*pw : j*pw+pw+2]
 *   h = relu6(bn1(W1 . tile));  d = relu6(bn2(dw3x3_valid(h; W2)));  o = bn3(W3 . d)
 *   y[b,:,i*ph+u,j*pw+v] = o[:,u,v] (+ x when `residual`)
 *   per-patch weight vector = [W1 (hid x Cin) | W2 (hid x 3 x 3) | W3 (Cout x hid)].
 * Replaces HyperPatchInvertedResidual.conv/forward (hyperseg/models/hyperseg_v1_0.py:328-376,
 * hyperseg_v1_0_unify.py:342-389).  bn*_scale/shift are the eval-mode BatchNorm folded to
 * y = scale*x + shift (scale = gamma/sqrt(var+eps), shift = beta - mean*scale), fp32,
 * lengths hid, hid, Cout.
 */
int hsb_patch_ir_fwd(const void* x, const void* w, void* y,
                     const float* bn1_scale, const float* bn1_shift,
                     const float* bn2_scale, const float* bn2_shift,
                     const float* bn3_scale, const float* bn3_shift,
                     int B, int Cin, int hid, int Cout, int H, int W, int fh, int fw,
                     int residual, int dtype, int w_layout, int64_t w_row_stride, void* stream);

/*
 * Restage-free tensor-core form of the same MetaBlock (bf16; tcgen05 / TMEM / TMA; csrc/patch_ir2.cu).  The per-patch
 * weights come as "arranged" rows: the numbers of [W1 | W2 | W3] with the three BatchNorm scales folded in, ordered as
 * the kernel's UMMA operands (csrc/ir_arranged.cuh):
 *   B1 [ceil(Cin/8)][hid][8] = bn1_scale[n]*W1[n][k] | W2T [9][hid] = bn2_scale[c]*W2[c][tap] (padded to 16 B) |
 *   B2 [ceil(hid/8)][Cout][8] = bn3_scale[n]*W3[n][k],   zero where k runs past Cin / hid.
 *   hsb_ir_arranged_row_elems      bf16 elements of one arranged row (-1 on bad dimensions)
 *   hsb_patch_ir_arranged_supported  1 when (Cin, hid, Cout, patch size) has an instantiation
 *   hsb_ir_arrange_weights         reference-order weights (fp32 / bf16, either layout) -> arranged rows, row stride
 *                                  hsb_ir_arranged_row_elems; the weight head can also emit arranged rows directly
 *                                  (hsb_head_pack_arranged / hsb_signal2weights_packed_fwd)
 *   hsb_patch_ir_arranged_fwd      y = bn3(W3 . relu6(bn2(dw3x3(relu6(bn1(W1 . tile))))));  x, y bf16 NCHW, 16-byte
 *                                  aligned, W % 8 == 0; square 16x16 or 8x8 patches; no residual.
 * Replaces HyperPatchInvertedResidual.conv (hyperseg/models/hyperseg_v1_0.py:328-370, hyperseg_v1_0_unify.py:342-389).
 */
int64_t hsb_ir_arranged_row_elems(int Cin, int hid, int Cout);
int hsb_patch_ir_arranged_supported(int Cin, int hid, int Cout, int patch_size);
int hsb_ir_arrange_weights(const void* w, void* w_arranged,
                           const float* bn1_scale, const float* bn2_scale, const float* bn3_scale,
                           int B, int Cin, int hid, int Cout, int fh, int fw,
                           int dtype, int w_layout, int64_t w_row_stride, void* stream);
int hsb_patch_ir_arranged_fwd(const void* x, const void* w_arranged, void* y,
                              const float* bn1_shift, const float* bn2_shift, const float* bn3_shift,
                              int B, int Cin, int hid, int Cout, int H, int W, int fh, int fw,
                              int64_t w_row_stride, void* stream);

/*
 * Weight head: grouped 1x1 convolution from the signal map to per-patch weights.
 *   Wout[b, o, i, j] = sum_{k < sig_ch/G} Ws[o,k] * s[b, sig_index + (o / (out_ch/G))*(sig_ch/G) + k, i, j],  o < hp
 * Replaces apply_signal2weights + nn.Conv2d(groups) (hyperseg/models/hyperseg_v1_0.py:315-326,
 * :473-484, :531-541), WeightLayer.forward (hyperseg_v1_0_unify.py:287-309) and one head of
 * Conv2dMulti (hyperseg_v0_1.py:336-362).
 *   s   (B, *, fh, fw) signal with element strides s_stride_b / s_stride_c / s_stride_p
 *       (NCHW: C*fh*fw, fh*fw, 1;  NHWC: fh*fw*C, 1, C)
 *   ws  (out_ch, sig_ch/G) contiguous, same dtype (the nn.Conv2d weight, out_ch a multiple of G)
 *   w_out in `out_layout`; rows of HSB_W_PATCH_MAJOR are `out_row_stride` elements apart and
 *       only the first hp entries of each row are written.
 */
int hsb_signal2weights_fwd(const void* s, const void* ws, void* w_out,
                           int B, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                           int fh, int fw,
                           int64_t s_stride_b, int64_t s_stride_c, int64_t s_stride_p,
                           int dtype, int out_layout, int64_t out_row_stride, void* stream);

/*
 * Tensor-core variant of the weight head for bf16 (tcgen05).  The static head weights are first packed, once,
 * into the UMMA operand layout.  The packed buffer is opaque (16-byte aligned, caller-allocated): when the head's signal
 * slice fits one CTA's shared memory it holds an item table followed by 128-column tiles of the reference-order row whose K
 * is the union of the signal groups they read (the kernel keeps the slice resident); otherwise one 256-column tile per
 * (group, tile) with K = sig_ch / groups padded to 16.  hsb_head_pack and hsb_signal2weights_packed_fwd agree on the choice
 * (a function of sig_ch, out_ch, groups only).
 *   hsb_head_packed_elems  number of bf16 elements the packed buffer needs (-1 on bad dimensions)
 *   hsb_head_pack          ws (out_ch, sig_ch/G) of `dtype` -> packed; row_scale (fp32, out_ch entries, may be NULL)
 *                          multiplies each output channel's row, which lets an inference engine fold a BatchNorm
 *                          scale that follows the dynamic convolution into the head
 *   hsb_signal2weights_packed_fwd   same result as hsb_signal2weights_fwd with HSB_BF16 / HSB_W_PATCH_MAJOR.
 * Requirements: s is bf16 with position stride 1 (NCHW), fh*fw % 8 == 0, 16-byte aligned base and strides.
 * Replaces the same reference lines as hsb_signal2weights_fwd.
 */
int64_t hsb_head_packed_elems(int sig_ch, int out_ch, int groups);
int hsb_head_pack(const void* ws, void* packed, const float* row_scale, int sig_ch, int out_ch, int groups,
                  int dtype, void* stream);
int hsb_signal2weights_packed_fwd(const void* s, const void* packed, void* w_out,
                                  int B, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                  int fh, int fw, int64_t s_stride_b, int64_t s_stride_c,
                                  int64_t out_row_stride, void* stream);

/*
 * The weight head emitting "arranged" rows for hsb_patch_ir_arranged_fwd directly (no reference-order tensor in between):
 * the static head weights of the block -- rows [hp_offset, hp_offset + hp) of the head, hp = Cin*hid + 9*hid + hid*Cout --
 * are packed once in the block's operand order with its three BatchNorm scales folded in.  Because that order interleaves the
 * head's groups, every 128-column tile carries its own range of signal channels (a small table, built by the pack call).
 *   hsb_head_arranged_plan           sizes: bf16 elements of the packed buffer, number of tiles (= int4 table entries),
 *                                    widest padded signal range of a tile
 *   hsb_head_pack_arranged           ws (out_ch, sig_ch/G) -> packed + table (both caller-allocated device buffers)
 *   hsb_signal2weights_arranged_fwd  signal -> (B, fh, fw, out_row_stride) arranged rows, row_elems =
 *                                    hsb_ir_arranged_row_elems(Cin, hid, Cout); same signal requirements as the packed head
 * Replaces apply_signal2weights of HyperPatchInvertedResidual (hyperseg/models/hyperseg_v1_0.py:315-326) and the slice of
 * WeightLayer.forward that feeds an inverted-residual level (hyperseg_v1_0_unify.py:246-249, :301-309).
 */
int hsb_head_arranged_plan(int sig_index, int sig_ch, int out_ch, int groups, int hp_offset, int Cin, int hid, int Cout,
                           int64_t* packed_elems, int* n_items, int* kpad_max);
int hsb_head_pack_arranged(const void* ws, void* packed, void* table,
                           const float* bn1_scale, const float* bn2_scale, const float* bn3_scale,
                           int sig_index, int sig_ch, int out_ch, int groups, int hp_offset,
                           int Cin, int hid, int Cout, int dtype, void* stream);
int hsb_signal2weights_arranged_fwd(const void* s, const void* packed, const void* table, void* w_arranged,
                                    int B, int sig_index, int sig_ch, int n_items, int kpad_max, int row_elems,
                                    int fh, int fw, int64_t s_stride_b, int64_t s_stride_c, int64_t out_row_stride,
                                    void* stream);

/*
 * General patch-wise convolution (any kernel size / groups / dilation, stride 1):
 *   tile = pad(x, (pad_h,pad_w), pad_mode)[b, :, i*ph : i*ph+ph+2*pad_h, j*pw : j*pw+pw+2*pad_w]
 *   y patch = valid_conv(tile, Wm),  Wm[o,c,ky,kx] = w_patch[((o*(Cin/G)+c)*kh+ky)*kw+kx]
 * Requires 2*pad == dilation*(k-1) per axis (output patch == input patch, as every
 * reference call site has).  Replaces MetaPatch.forward + MetaConv2d.forward
 * (hyperseg/models/layers/meta_patch.py:35-57, meta_conv.py:163-186) and HyperPatch.forward
 * (hyperseg/models/hyperseg_v1_0.py:543-557).  Same fused epilogue as the 1x1 entry point.
 */
int hsb_patch_conv_fwd(const void* x, const void* w, void* y,
                       const float* post_scale, const float* post_shift, int act,
                       int B, int Cin, int Cout, int H, int W, int fh, int fw,
                       int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                       int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream);

/*
 * Per-batch-element dynamic convolution, stride 1, explicit padding:
 *   y[n] = conv2d(pad(x[n]), Wm[n]),  w is (N, Cout*Cin/G*kh*kw) contiguous.
 * Replaces MetaConv2d.forward used on its own (hyperseg/models/layers/meta_conv.py:163-186).
 */
int hsb_meta_conv2d_fwd(const void* x, const void* w, void* y,
                        int N, int Cin, int Cout, int H, int W,
                        int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                        int pad_mode, int dtype, void* stream);

/*
 * Training path (BASELINE config 4 = HyperSeg-L / hyperseg_v0_1 training step): gradients of
 * hsb_patch_conv_fwd (and of the 1x1 kernel, with kh = kw = 1, pad 0) and of the weight head.  In the reference these
 * come from autograd through F.pad / F.unfold / F.conv2d(groups) / F.fold (hyperseg/models/layers/meta_patch.py:35-57,
 * meta_conv.py:163-186) and through the grouped nn.Conv2d of the heads (hyperseg_v0_1.py:336-362, hyperseg_v1_0.py:315-326).
 *   hsb_patch_conv_bwd_weight   dw (same layout / row stride as w) from x and dy
 *   hsb_patch_conv_bwd_input    dx from w and dy (gathers across patch borders and padding aliases; no atomics)
 *   hsb_signal2weights_bwd_signal   ds (B, sig_total, fh, fw) with the given strides; channels outside the head's slice get 0
 *   hsb_signal2weights_bwd_weight   dws (out_ch, sig_ch/G) contiguous; rows >= hp get 0
 * `dwout` is the gradient of the head's output in `g_layout` (HSB_W_PATCH_MAJOR rows are g_row_stride apart).
 */
int hsb_patch_conv_bwd_weight(const void* x, const void* dy, void* dw,
                              int B, int Cin, int Cout, int H, int W, int fh, int fw,
                              int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                              int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream);
int hsb_patch_conv_bwd_input(const void* w, const void* dy, void* dx,
                             int B, int Cin, int Cout, int H, int W, int fh, int fw,
                             int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                             int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream);
int hsb_signal2weights_bwd_signal(const void* ws, const void* dwout, void* ds,
                                  int B, int sig_total, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                  int fh, int fw, int64_t ds_stride_b, int64_t ds_stride_c, int64_t ds_stride_p,
                                  int dtype, int g_layout, int64_t g_row_stride, void* stream);
int hsb_signal2weights_bwd_weight(const void* s, const void* dwout, void* dws,
                                  int B, int sig_total, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                  int fh, int fw, int64_t s_stride_b, int64_t s_stride_c, int64_t s_stride_p,
                                  int dtype, int g_layout, int64_t g_row_stride, void* stream);

/*
 * Decoder glue, one pass: out = cat(coords, feature, bilinear_upsample(prev, (H, W)))  along channels.
 *   coords  (1, Cc, H, W) contiguous, broadcast over the batch (may be NULL with Cc == 0)
 *   feature (B, Cf, H, W) with arbitrary element strides (NCHW or NHWC)
 *   prev    (B, Cp, h, w) contiguous NCHW, upsampled as F.interpolate(mode='bilinear', align_corners=False)
 *   out     (B, Cc+Cf+Cp, H, W) contiguous NCHW
 * Replaces F.interpolate + torch.cat + torch.cat at hyperseg/models/hyperseg_v1_0.py:235-240
 * (hyperseg_v1_0_unify.py:234-240, hyperseg_v0_1.py:187-194).
 */
int hsb_decoder_input_fwd(const void* coords, const void* feature, const void* prev, void* out,
                          int B, int Cc, int Cf, int Cp, int H, int W, int h, int w,
                          int64_t f_stride_b, int64_t f_stride_c, int64_t f_stride_y, int64_t f_stride_x,
                          int dtype, void* stream);

/*
 * Segmentation tail when only labels are wanted: labels[b,y,x] = argmax_c bilinear_upsample(logits)[b,c,y,x]
 * (uint8).  Replaces the final F.interpolate (hyperseg/models/hyperseg_v1_0.py:250-251) followed by
 * output.argmax(1) (hyperseg/test.py:171); the full-resolution logits are never written.
 */
int hsb_upsample_argmax_fwd(const void* logits, void* labels, int B, int C, int h, int w, int H, int W,
                            int dtype, void* stream);

/*
 * Layout change (B, hp, fh, fw) -> (B, fh, fw, row_stride) for weights handed over in the
 * reference layout.  Replaces weight.permute(0,2,3,1).reshape(...) at
 * hyperseg/models/hyperseg_v1_0.py:345-347, :492-493, :549.
 */
int hsb_weights_to_patch_major(const void* w_nchw, void* w_pm, int B, int hp, int fh, int fw,
                               int64_t row_stride, int dtype, void* stream);

/*
 * Encoder epilogues (engine path, outside the decoder hot path): channels-last (N, HW, C) activations of the stock
 * EfficientNet encoder after its eval-mode BatchNorms have been folded into the convolutions
 * (hyperseg/models/backbones/efficientnet.py:82-123 MBConvBlock.forward, :275 stem, :289 head).
 *   hsb_bias_act_nhwc_fwd      y = act(x + bias[c]) (+ residual); y may alias x.  Replaces BatchNorm + swish (:97-98,
 *                              :101-102, :275, :289), BatchNorm after the projection and the skip add (:113, :122).
 *                              With pool_partial != NULL also writes per-chunk sums of the rounded output,
 *                              (N, chunks, C) float32 with chunks = hsb_bias_act_nhwc_chunks(C, HW, dtype); their sum
 *                              over chunks / HW is F.adaptive_avg_pool2d(y, 1) (:106), in a fixed summation order.
 *   hsb_channel_gate_nhwc_fwd  y = x * sigmoid(gate[n, c])  (:110); gate is (N, C) of the same dtype.
 * C * sizeof(element) must be a multiple of 16, tensors 16-byte aligned.
 */
int hsb_bias_act_nhwc_chunks(int C, int64_t HW, int dtype);
int hsb_bias_act_nhwc_fwd(const void* x, const float* bias, const void* residual, void* y, float* pool_partial,
                          int N, int64_t HW, int C, int act, int dtype, void* stream);
int hsb_channel_gate_nhwc_fwd(const void* x, const void* gate, void* y, int N, int64_t HW, int C, int dtype,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HSB200_H_ */

/*
 * neosr_b200 — C ABI of the B200-native neosr training-step kernels.
 *
 * The reference (muslll/neosr) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md §0 fact 1, §8b).  Each entry point below therefore cites the
 * reference ATen/cuDNN call site (file:line under /root/reference) that it
 * replaces.  Conventions for every function:
 *
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers to
 *     fp32 (unless stated) buffers owned by the caller, borrowed for the call;
 *   - activations are NHWC ("tokens": [batch*h*w, channels]), weights are either
 *     the reference's own layout (OIHW / [out,in]) or an opaque packed buffer
 *     produced by nsr_pack_weight();
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), never
 *     synchronised; entry points are re-entrant (no global mutable state);
 *   - return 0 on success, a negative NSR_E_* code otherwise; nsr_last_error()
 *     returns a thread-local message.  Nothing throws or exits.
 */
#ifndef NEOSR_B200_H
#define NEOSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSR_OK 0
#define NSR_E_INVALID (-1)   /* bad argument / unsupported shape */
#define NSR_E_CUDA (-2)      /* CUDA runtime error (message in nsr_last_error) */
#define NSR_E_WORKSPACE (-3) /* workspace too small */

/* activation codes for the fused GEMM/conv epilogues */
#define NSR_ACT_NONE 0
#define NSR_ACT_RELU 1
#define NSR_ACT_LRELU 2 /* slope = act_slope */
#define NSR_ACT_GELU 3  /* exact erf GELU (torch.nn.GELU default) */
#define NSR_ACT_PRELU 4 /* per-output-channel slope vector */
#define NSR_ACT_MULAUX 5 /* actgrad only: aux already holds act'(pre) (saved by pre_mode = 1) */

/* engine selection for the contraction kernels */
#define NSR_ENGINE_AUTO 0
#define NSR_ENGINE_SIMT 1    /* exact-fp32 CUDA-core implicit GEMM */
#define NSR_ENGINE_TCGEN05 2 /* tcgen05 (UMMA) 3xBF16-split, fp32 accumulate in TMEM */
#define NSR_ENGINE_MMA_SYNC 3 /* nsr_window_attn_wsti_*: the warp-level mma.sync kernels instead of tcgen05 (A/B runs) */
#define NSR_ENGINE_BF16 4     /* nsr_conv_fprop / nsr_conv_wgrad(_partial): routed like NSR_ENGINE_AUTO, but the tcgen05 kernels
                                 issue ONE bf16 pass (hi x hi, fp32 accumulate) and fetch no lo halves - the mixed-precision mode
                                 behind `use_amp` + `bfloat16` (neosr/models/image.py:117-127); everything stored stays fp32 */

const char* nsr_last_error(void);
int nsr_version(void);
/* 1 when the current device is sm_100 and the tcgen05 kernels may be used. */
int nsr_device_supports_tcgen05(void);

/* ------------------------------------------------------------------ contraction ---- */
/*
 * One descriptor drives Linear and Conv2d (stride 1) as an implicit GEMM over NHWC:
 *   y[p, co] = epilogue( sum_{r,s,ci} x[p shifted by (r-pad, s-pad), ci] * w[co, r, s, ci] )
 * Replaces: nn.Linear (swinir_arch.py:27-29,139,143 — qkv/proj/fc1/fc2), nn.Conv2d
 * (swinir_arch.py:889,630,964,978-982,782; compact_arch.py:48-72; esrgan_arch.py:97-106;
 * vgg_arch.py:134 features) and their autograd dgrad (same call with the dgrad-packed
 * weight).  Epilogue, in order:
 *   v = acc + bias[co]                       (bias may be NULL)
 *   if (y_pre) y_pre[p,co] = pre_mode ? act'(v) : v   (saved for the backward pass)
 *   v = act(v)                               (NSR_ACT_*)
 *   if (mul_actgrad) v *= act'(aux[p,co])    (chain rule through the producer's
 *                                             activation; aux = its saved output /
 *                                             pre-activation, code in actgrad)
 *   if (row_scale) v *= row_scale[p / (h*w)] (per-sample DropPath factor, arch_util.py:118-131)
 *   if (residual) v += residual[p,co]
 *   y[p,co] = v
 */
typedef struct NsrConv {
  int32_t batch, h, w;   /* input == output spatial size (stride 1, "same" padding) */
  int32_t cin, cout;
  int32_t kh, kw, pad;
  int32_t x_ld;          /* floats between consecutive pixels of x   (>= cin)  */
  int32_t y_ld;          /* floats between consecutive pixels of y/y_pre/residual/aux (>= cout) */
  int32_t act;           /* NSR_ACT_* applied to the output */
  float act_slope;
  int32_t actgrad;       /* NSR_ACT_* whose derivative multiplies the output (0 = none) */
  float actgrad_slope;
  int32_t engine;        /* NSR_ENGINE_* */
  const float* x;
  const void* w_packed;  /* from nsr_pack_weight (fprop or dgrad flavour) */
  const float* bias;
  const float* prelu;    /* [cout] slopes for NSR_ACT_PRELU (act or actgrad) */
  const float* aux;
  const float* row_scale;
  const float* residual;
  float* y_pre;
  float* y;              /* may be NULL when y_sti is given */
  /* Split-tile-image (STI) operands, see nsr_sti_bytes(): x_sti replaces x as the A operand of a
   * 1x1 contraction (bulk-copied, no conversion in the kernel); y_sti receives the final output
   * value pre-split to bf16 hi/lo in the layout the next contraction bulk-copies. */
  const void* x_sti;
  void* y_sti;
  int32_t res_ld;        /* leading dims of `residual` / `aux` when they differ from y_ld (0 = y_ld): lets an */
  int32_t aux_ld;        /* epilogue read a channel slab of a wider buffer (ESRGAN dense blocks)              */
  int32_t pre_mode;      /* 0: y_pre = pre-activation; 1: y_pre = act'(pre-activation), so the backward
                            epilogue is a plain multiply (actgrad = NSR_ACT_MULAUX) and the erf/exp terms
                            are shared with the forward activation */
  int32_t sti_win;       /* 0: y_sti rows in token (image) order.  ws | shift << 16: WINDOW-ORDERED image - token (b, y, x)
                            is stored at row ((b*H/ws + wy)*W/ws + wx)*ws*ws + iy*ws + ix, (wy,iy) = divmod((y - shift) mod H,
                            ws), (wx,ix) likewise: torch.roll(-shift) + window_partition (swinir_arch.py:41-57,356-366) folded
                            into the producing contraction's store, so nsr_window_attn_wsti_* bulk-copy whole windows */
  int32_t aux_mode;      /* 0: aux is fp32.  2: aux is the 16-bit activation-gradient code below (actgrad = NSR_ACT_MULAUX).
                            pre_mode = 2 likewise makes y_pre a uint16 buffer [pixels, y_ld] of codes
                              code = rint((act'(pre) + 0.25) * 40000)      act' = code / 40000 - 0.25
                            (|error| <= 1.25e-5 for GELU', whose range is [-0.13, 1.13]): the fc1 -> fc2-dgrad hand-over of a
                            Swin MLP (swinir_arch.py:27-37) at half the bytes.  tcgen05 engine, split-tile-image output only:
                            {act = GELU, y_pre, y_sti} and {actgrad = MULAUX, aux, y_sti}, no y / residual / row_scale */
  void* workspace;       /* optional scratch of nsr_conv_fprop_workspace() bytes: lets <= 4-channel convolutions */
  size_t workspace_bytes;/* (conv_first / conv_last, VGG conv1_1) run as im2col + one tensor-core contraction */
} NsrConv;

/* Scratch nsr_conv_fprop can use for this descriptor (0 when it needs none). */
size_t nsr_conv_fprop_workspace(const NsrConv* desc);

int nsr_conv_fprop(const NsrConv* d, void* stream);

/*
 * Split tile image (STI) of a token tensor [rows, c] (fp32 semantics): the tensor is stored as
 * bf16 hi + bf16 lo (x ~= hi + lo, the 3xBF16 operand split) in 128-row x 64-channel blocks,
 * each block = 16 KiB hi image + 16 KiB lo image, a row = 128 bytes, 16-byte chunk `ch` of row
 * `r` stored at chunk position ch ^ (r & 7) (the SWIZZLE_128B pattern tcgen05 reads).  Block
 * (mt, kb) lives at ((mt * ceil(c/64)) + kb) * 32 KiB.  Same bytes as fp32; a contraction
 * bulk-copies blocks straight into shared memory (fprop/dgrad: K-major A tile; wgrad: MN-major
 * panels).  Rows >= `rows` and channels >= `c` must read as zero.
 */
size_t nsr_sti_bytes(long long rows, int c);
/* fp32 [rows, c] (row stride ld) -> STI (utility / tests; producers normally emit STI directly) */
int nsr_sti_from_f32(const float* x, int ld, long long rows, int c, void* sti, void* stream);
int nsr_sti_to_f32(const void* sti, long long rows, int c, float* y, int ld, void* stream);

/* Packed-weight buffers. `flavour` 0 = fprop  (w[co][r][s][ci]),
 *                         1 = dgrad  (w'[ci][kh-1-r][kw-1-s][co], i.e. the transposed,
 *                                     180deg-rotated filter so that dgrad is an fprop).
 * Source is the reference layout OIHW ([cout, cin, kh, kw]; Linear = kh=kw=1).
 * The buffer holds an fp32 copy (SIMT engine) followed by the bf16 hi/lo tile images the
 * tcgen05 engine streams with bulk copies. */
size_t nsr_packed_weight_bytes(int cout, int cin, int kh, int kw, int flavour);
int nsr_pack_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int flavour,
                    void* packed, void* stream);
/* Both flavours in ONE launch (packed_dgrad may be NULL): what the engine calls per parameter after every optimizer step. */
int nsr_pack_weight_pair(const float* w_oihw, int cout, int cin, int kh, int kw, void* packed_fprop,
                         void* packed_dgrad, void* stream);

/* Every weight of a network in ONE launch (what the engines call after an optimizer step).  Entry = one Conv2d / Linear
 * weight, optionally re-indexed through int maps (padded output row / input column -> source index, -1 = zero: the
 * head-padded qkv / proj copies behind nsr_window_attn_wsti_*), packed into the fprop and / or dgrad flavour; bias_out
 * (optional) receives the row-mapped bias.  `cout, cin` are the PACKED (padded) sizes, `src_cin` the source tensor's input
 * channel count.  block_base: first CUDA block of the entry, a running sum of nsr_pack_entry_blocks(cout, cin, kh, kw). */
typedef struct NsrPackEntry {
  const float* w;
  const float* bias;
  void* packed_fprop;
  void* packed_dgrad;
  float* bias_out;
  const int32_t* row_map;
  const int32_t* col_map;
  int32_t cout, cin, kh, kw;
  int32_t src_cin, reserved;
  int64_t block_base;
} NsrPackEntry;
int64_t nsr_pack_entry_blocks(int cout, int cin, int kh, int kw);
int nsr_pack_weights_multi(const NsrPackEntry* table_dev, int n_entries, int64_t total_blocks, void* stream);

/*
 * Weight gradient of the same contraction (autograd of nn.Linear / nn.Conv2d weights):
 *   dw[co, ci, r, s] = sum_p dy[p, co] * x[p shifted by (r-pad, s-pad), ci]      (OIHW out)
 *   dbias[co]        = sum_p dy[p, co]                                            (optional)
 * Deterministic split-K: partials go to `workspace`, a second pass reduces them in a
 * fixed order.  nsr_conv_wgrad_workspace() gives the size needed.
 */
typedef struct NsrWgrad {
  int32_t batch, h, w;
  int32_t cin, cout;
  int32_t kh, kw, pad;
  int32_t x_ld, dy_ld;
  int32_t engine;
  const float* x;
  const float* dy;
  float* dw;     /* [cout, cin, kh, kw] fp32, overwritten */
  float* dbias;  /* [cout] or NULL, overwritten */
  void* workspace;
  size_t workspace_bytes;
  const void* x_sti;   /* optional STI copies of x / dy (1x1 only): operands arrive by bulk copy */
  const void* dy_sti;
} NsrWgrad;

size_t nsr_conv_wgrad_workspace(const NsrWgrad* d);
int nsr_conv_wgrad(const NsrWgrad* d, void* stream);

/* Deferred reduction of 1x1 weight gradients (split-tile-image operands, tcgen05 engine): nsr_conv_wgrad_partial leaves
 * the split-K partials [splitk][cout][cin] of the descriptor's (possibly head-padded, bias-column-extended) problem in
 * d->workspace (>= nsr_conv_wgrad_partial_workspace bytes, owned by the caller until the finalize) and reports splitk;
 * d->dw / d->dbias are ignored.  nsr_wgrad_finalize_multi then reduces ANY number of such work spaces in one launch, in a
 * fixed order, straight into the parameter gradients:
 *   dw[co][ci] = sum_k partial[k][row_map[co]][col_map[ci]]     dbias[co] = sum_k partial[k][row_map[co]][bias_col]
 * (NULL map = identity; p_rows x p_cols = the padded problem; block_base = running sum of nsr_reduce_entry_blocks). */
size_t nsr_conv_wgrad_partial_workspace(const NsrWgrad* d);
int nsr_conv_wgrad_partial(const NsrWgrad* d, int* splitk, void* stream);
typedef struct NsrReduceEntry {
  const float* partial;
  float* dw;
  float* dbias;            /* NULL: no bias gradient */
  const int32_t* row_map;  /* [cout] -> row of the padded problem */
  const int32_t* col_map;  /* [cin]  -> column of the padded problem */
  int32_t splitk, p_rows, p_cols;
  int32_t cout, cin, bias_col;
  int64_t block_base;
} NsrReduceEntry;
int64_t nsr_reduce_entry_blocks(int cout, int cin);
int nsr_wgrad_finalize_multi(const NsrReduceEntry* table_dev, int n_entries, int64_t total_blocks, void* stream);

/* ------------------------------------------------------------------ layout / elementwise */
/* y_nhwc[b,h,w,c] = x_nchw[b,c,h,w] * scale[c] + shift[c]
 * Replaces `(x - self.mean) * self.img_range` (swinir_arch.py:1041-1042) and the VGG input
 * normalisation (vgg_arch.py:189-190) fused with the NCHW->NHWC transpose. */
int nsr_nchw_to_nhwc_affine(const float* x, float* y, int batch, int c, int h, int w,
                            const float* scale, const float* shift, void* stream);
/* y_nchw[b,c,h,w] = x_nhwc[b,h,w,c] * scale[c] + shift[c]   (swinir_arch.py:1077 and the
 * backward of the above: pass shift = NULL). */
int nsr_nhwc_to_nchw_affine(const float* x, float* y, int batch, int c, int h, int w,
                            const float* scale, const float* shift, void* stream);

/* nn.PixelShuffle(r) on NHWC (swinir_arch.py:783,786,810; compact_arch.py:73):
 *   y[b, h*r+i, w*r+j, c] = x[b, h, w, c*r*r + i*r + j]          (bit-exact index map)
 * inverse != 0 runs the map backwards (autograd of pixel_shuffle == pixel_unshuffle). */
int nsr_pixel_shuffle_nhwc(const float* x, float* y, int batch, int h, int w, int c_out, int r,
                           int inverse, void* stream);

/* nn.MaxPool2d(2,2) on NHWC (vgg_arch.py:143) and its backward fused with the ReLU mask of
 * the pooled tensor and an optional extra gradient term:
 *   dx[i] = (x[i] > 0 ? routed(dy) : 0) + (dextra ? dextra[i] : 0) */
int nsr_maxpool2_nhwc(const float* x, float* y, int batch, int h, int w, int c, void* stream);
int nsr_maxpool2_relu_bwd_nhwc(const float* x, const float* dy, const float* dextra, float* dx,
                               int batch, int h, int w, int c, void* stream);

/* nn.PReLU(C) backward on NHWC (compact_arch.py:52-70): dx = dy * (pre > 0 ? 1 : slope[c]),
 * dslope[c] = sum_rows dy * min(pre, 0); deterministic two-pass reduction via workspace. */
size_t nsr_prelu_bwd_workspace(int c);
int nsr_prelu_bwd(const float* dy, const float* pre, const float* slope, float* dx, float* dslope,
                  long long rows, int c, void* workspace, size_t workspace_bytes, void* stream);
/* y_nchw[b,c,Y,X] = x_nhwc[b,Y,X,c] + base_nchw[b,c,Y/s,X/s]: the `out += F.interpolate(x, scale, "nearest")`
 * skip of compact_arch.py:80-84 fused with the NHWC->NCHW output transpose (C <= 4). */
int nsr_nhwc_to_nchw_add_nearest(const float* x, const float* base, float* y, int batch, int c, int h, int w,
                                 int scale, void* stream);

/* 2-D strided glue for channel slabs [rows, cols] with independent leading dims:
 *   axpby2d:       y = a * alpha + b * beta               (b may be NULL)
 *   actgrad_mul2d: dx = dy * act'(aux)                    (LeakyReLU/ReLU backward on a slab slice) */
int nsr_axpby2d(const float* a, int lda, float alpha, const float* b, int ldb, float beta, float* y, int ldy,
                long long rows, int cols, void* stream);
int nsr_actgrad_mul2d(const float* dy, int ld_dy, const float* aux, int ld_aux, float* dx, int ld_dx, long long rows,
                      int cols, int act, float slope, void* stream);
/* F.interpolate(x, scale_factor=2, mode="nearest") on NHWC (esrgan_arch.py:207-211) and its backward
 * (sum of each 2x2 block). */
int nsr_nearest_up2_nhwc(const float* x, float* y, int batch, int h, int w, int c, void* stream);
int nsr_nearest_up2_bwd_nhwc(const float* dy, float* dx, int batch, int h, int w, int c, void* stream);

/* ---- U-Net discriminator pieces (neosr/archs/unet_arch.py:40-67) ---- */
/* F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) on NHWC and its backward. */
int nsr_bilinear_up2_nhwc(const float* x, float* y, int batch, int h, int w, int c, void* stream);
int nsr_bilinear_up2_bwd_nhwc(const float* dy, float* dx, int batch, int h, int w, int c, void* stream);
/* nn.Conv2d(cin, cout, 4, 2, 1) == 3x3 stride-1 "same" conv over pixel_unshuffle(x, 2) with the weight
 * w3[cout, 4*cin, 3, 3] this remap builds from w4[cout, cin, 4, 4] (unused taps are zero); inverse != 0
 * gathers a gradient in the 3x3 layout back into the 4x4 layout. */
int nsr_conv4x4s2_remap(const float* src, float* dst, int cout, int cin, int inverse, void* stream);
/* torch.nn.utils.spectral_norm on W[rows = cout, cols = cin*kh*kw]: `power_iterations` updates of the
 * u / v buffers in place (1 in training mode, 0 in eval), sigma = u^T W v, w_out = W / sigma. */
size_t nsr_spectral_norm_workspace(int rows, int cols);
int nsr_spectral_norm_fwd(const float* w_orig, float* u, float* v, float* w_out, float* sigma, int rows, int cols,
                          int power_iterations, float eps, void* workspace, size_t workspace_bytes, void* stream);
/* dL/dW_orig = (G - <G, W_sn> u v^T) / sigma with G = dL/dW_sn (u, v constants, as in torch);
 * accumulate != 0 adds into dw_orig. */
int nsr_spectral_norm_bwd(const float* g_wsn, const float* w_sn, const float* u, const float* v, const float* sigma,
                          float* dw_orig, int rows, int cols, int accumulate, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- RealPLKSR pieces (neosr/archs/realplksr_arch.py) ---- */
/* nn.Mish (DCCM, :14-23) and its derivative times dy. */
int nsr_mish_fwd(const float* x, float* y, size_t n, void* stream);
int nsr_mish_bwd(const float* dy, const float* x, float* dx, size_t n, void* stream);
/* EA gate (:44-53): y = x * sigmoid(s); backward gives dx (gate path only) and ds. */
int nsr_mul_sigmoid_fwd(const float* x, const float* s, float* y, size_t n, void* stream);
int nsr_mul_sigmoid_bwd(const float* dy, const float* x, const float* s, float* dx, float* ds, size_t n, void* stream);
/* y[rows, c*r] += repeat_interleave(x[rows, c], r) (:158-160), NHWC. */
int nsr_add_repeat_interleave(float* y, const float* x, size_t rows, int c, int r, void* stream);
/* nn.GroupNorm(groups, c) on NHWC [batch, hw, c] (+ optional residual add, PLKBlock :85-99); mean / rstd
 * [batch*groups] are saved for the backward, which also yields dgamma / dbeta (overwritten). */
size_t nsr_groupnorm_workspace(int batch, int c, int groups);
int nsr_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y, float* mean,
                      float* rstd, int batch, int hw, int c, int groups, float eps, void* workspace, size_t workspace_bytes,
                      void* stream);
int nsr_groupnorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx,
                      float* dgamma, float* dbeta, int batch, int hw, int c, int groups, void* workspace,
                      size_t workspace_bytes, void* stream);

/* y = a * alpha + b * beta (b may be NULL). Gradient accumulation glue. */
int nsr_axpby(const float* a, float alpha, const float* b, float beta, float* y, size_t n, void* stream);
/* dx = dy * act'(aux) + (dextra ? dextra : 0) */
int nsr_actgrad_mul(const float* dy, const float* aux, const float* dextra, float* dx, size_t n,
                    int act, float slope, void* stream);

/* RealPLKSR's partial large-kernel conv (neosr/archs/realplksr_arch.py:26-41): dense k x k (odd, <= 17), 16 -> 16
 * channels, stride 1, "same" zero padding, on a channel slab of an NHWC tensor (x_ld / y_ld / dy_ld = floats between
 * consecutive pixels).  Exact fp32.  w_ntc: [n = 16][tap = k*k][c = 16] — the fp32 region at the start of an
 * nsr_pack_weight() buffer (flavour 0 = fprop, flavour 1 = dgrad: rotated + transposed, so dgrad is the same call on
 * dy).  wgrad overwrites dw [16,16,k,k] (reference layout) and dbias [16] (may be NULL); deterministic. */
int nsr_conv_lk16_fprop(const float* x, int x_ld, const float* w_ntc, const float* bias, float* y, int y_ld, int batch,
                        int h, int w, int k, void* stream);
size_t nsr_conv_lk16_wgrad_workspace(int batch, int h, int w, int k);
int nsr_conv_lk16_wgrad(const float* x, int x_ld, const float* dy, int dy_ld, float* dw, float* dbias, int batch, int h,
                        int w, int k, void* workspace, size_t workspace_bytes, void* stream);

/* Split tile images written by nsr_layernorm_fwd, nsr_window_attn_fwd and nsr_conv_fprop carry 1.0 in their first
 * padding channel (channel C when C % 64 != 0; the packed weights are zero there, so no contraction over C sees it).
 * A weight gradient taken over cin + 4 input channels of such an image (same image geometry) therefore returns the bias
 * gradient as column cin: t = [cout, cin + 4]; this call scatters t into dw [cout, cin] and dbias [cout], replacing the
 * separate column-sum pass over dy. */
int nsr_wgrad_split(const float* t, float* dw, float* dbias, int cout, int cin, int cinp, void* stream);

/* ------------------------------------------------------------------ LayerNorm ---------- */
/* nn.LayerNorm(c, eps) over the last dim of [rows, c] (swinir_arch.py:284,297,708,960). */
int nsr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y,
                      float* mean, float* rstd, int rows, int c, float eps, void* y_sti, void* stream);
/* y (fp32) and y_sti (split tile image, see nsr_sti_bytes) are both optional outputs. */
/* dx = LN'(dy) + (dres ? dres : 0); dgamma/dbeta reduced deterministically via workspace
 * (>= nsr_layernorm_bwd_workspace(c) bytes). */
size_t nsr_layernorm_bwd_workspace(int c);
int nsr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                      const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                      int rows, int c, void* workspace, size_t workspace_bytes, void* dx_sti, void* stream);

/* Same, with two more options for the Swin blocks' split-tile-image path: the residual-branch gradient may arrive as a
 * split tile image (dres_sti; hi + lo bf16 ~ 2^-17 relative; alternative to dres), and dx may be NULL when dx_sti is given -
 * inside a residual group the gradient of the token stream then lives in ONE format instead of fp32 + tile image. */
/* dgamma = dbeta = NULL defers the parameter gradients: `workspace` (caller-owned until then) keeps the per-block partial sums
 * [nsr_layernorm_bwd_blocks(rows)][2][c] (dgamma row, dbeta row), which nsr_wgrad_finalize_multi reduces as two entries
 * {splitk = blocks, p_rows = 2, p_cols = c, cout = 1, cin = c, row_map = {0} | {1}} - one launch for every LayerNorm of a
 * network instead of one 10 us reduction per layer. */
int nsr_layernorm_bwd_blocks(int rows);
int nsr_layernorm_bwd2(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* dres, const void* dres_sti, float* dx, float* dgamma, float* dbeta, int rows, int c,
                       void* workspace, size_t workspace_bytes, void* dx_sti, void* stream);

/* ------------------------------------------------------------------ window attention --- */
/*
 * Fused (shifted-)window multi-head self-attention core on token-major qkv:
 *   cyclic shift + window_partition (swinir_arch.py:353-364), q*scale, q@k^T, +relative-
 *   position bias table[index] (185-195), +shift mask (197-202, calculate_mask 313-341 as a
 *   function of token coordinates), softmax, @v, window_reverse + un-shift (376-386).
 * qkv:  [batch*h*w, 3*c]  (output of the qkv Linear in natural token order)
 * out:  [batch*h*w, c]    (input of the proj Linear, natural token order)
 * bias_table: [(2*ws-1)^2, heads].  use_mask != 0 adds the {0,-100} mask for `shift`.
 * Supported: ws*ws <= 64, head_dim <= 32 (SwinIR S/M/L at window 8).
 */
int nsr_window_attn_fwd(const float* qkv, const float* bias_table, float* out, int batch, int h,
                        int w, int c, int heads, int ws, int shift, int use_mask, float scale,
                        void* out_sti, void* stream);
/* out (fp32) may be NULL when out_sti is given; likewise dqkv / dqkv_sti below. */
size_t nsr_window_attn_bwd_workspace(int heads, int ws);
int nsr_window_attn_bwd(const float* qkv, const float* bias_table, const float* dout, float* dqkv,
                        float* dbias_table, int batch, int h, int w, int c, int heads, int ws,
                        int shift, int use_mask, float scale, void* workspace,
                        size_t workspace_bytes, void* dqkv_sti, void* stream);

/*
 * The same attention core on WINDOW-ORDERED operands (the default path of the SwinIR engine on sm_100):
 *   qkv_wsti : split tile image [batch*h*w rows in window order, 3*G channels], G = nsr_window_attn_wsti_channels(heads)
 *              = heads*32 rounded up to 64: q | k | v groups, every head padded from c/heads to 32 channels (padding = 0).
 *              Produced by nsr_conv_fprop with NsrConv.sti_win = ws | shift << 16 from head-padded qkv weights
 *              (nsr_gather2d): roll + window_partition + the q/k/v reshape (swinir_arch.py:353-364,172-183) cost nothing,
 *              and a window's 64 tokens x a head pair of q, k or v is one contiguous 8 KiB run per bf16 half, fetched
 *              with cp.async.bulk (no fp32 gather, no in-kernel bf16 split).
 *   dout_wsti: the proj Linear's input gradient in the same row order, [rows, G] (heads padded to 32), produced by the
 *              proj dgrad contraction with sti_win set and head-padded weights.
 *   out / out_sti, dqkv / dqkv_sti, dbias_table: as nsr_window_attn_fwd / _bwd (natural token order, c and 3*c channels).
 * Needs ws == 8 and an even head dim <= 32; workspace as nsr_window_attn_bwd_workspace.
 * engine: NSR_ENGINE_AUTO / NSR_ENGINE_TCGEN05 = Q K^T and P V as tcgen05.mma with TMEM accumulators (window_attn_tc.cu: one
 * work item = a window pair x a head pair = three 32 KiB operand blocks, softmax by one thread per query row);
 * NSR_ENGINE_MMA_SYNC = the warp-level mma.sync kernels on the same operands.
 * out_padded (tcgen05 kernel): out_sti is [tokens, G] with every head padded to 32 channels like the operands, so a thread
 * stores whole 16-byte chunks; channel c/heads of head 0 (a padding slot when c/heads < 32) carries 1.0 - the bias-gradient
 * column of the proj contraction, which then runs on head-padded weights (K = G).
 */
int nsr_window_attn_wsti_channels(int heads);
int nsr_window_attn_wsti_fwd(const void* qkv_wsti, const float* bias_table, float* out, void* out_sti, int out_padded, int batch,
                             int h, int w, int c, int heads, int ws, int shift, int use_mask, float scale, int engine,
                             void* stream);
/* dqkv_padded (tcgen05 kernel): dqkv_sti is [tokens, 3G] in token order with heads padded to 32 channels (q | k | v groups of
 * G), stored 16 bytes at a time; the qkv dgrad / wgrad contractions then run on the head-padded weights.  The workspace
 * must hold (max(nsr_window_attn_bwd_workspace)) bytes as for nsr_window_attn_bwd. */
int nsr_window_attn_wsti_bwd(const void* qkv_wsti, const float* bias_table, const void* dout_wsti, float* dqkv, void* dqkv_sti,
                             int dqkv_padded, float* dbias_table, int batch, int h, int w, int c, int heads, int ws, int shift,
                             int use_mask, float scale, int engine, void* workspace, size_t workspace_bytes, void* stream);
/* Deferred bias-table gradients: with dbias_table = NULL nsr_window_attn_wsti_bwd leaves the per-CTA partial sums of dS in
 * `workspace` ([gx + 1][heads][64][64] floats, caller-owned until the reduction; gx = nsr_window_attn_wsti_bwd_gx(...) with the
 * same arguments), and nsr_window_attn_dbias_multi reduces the work spaces of ANY number of layers with two launches
 * (fixed order: deterministic) - instead of two ~9 us launches per layer. */
typedef struct NsrAttnBiasEntry {
  float* partial;      /* the layer's work space */
  float* dbias_table;  /* [(2 ws - 1)^2, heads], overwritten */
  int32_t gx, heads, ws, reserved;
} NsrAttnBiasEntry;
int nsr_window_attn_wsti_bwd_gx(int batch, int h, int w, int c, int heads, int ws, int dqkv_padded, int engine);
int nsr_window_attn_dbias_multi(const NsrAttnBiasEntry* table_dev, int n_entries, int max_heads, int max_ws, void* stream);
/* dst[r][c] = src[row_map[r]][col_map[c]]; NULL map = identity, negative entry = 0 (head-padded weight / bias copies). */
int nsr_gather2d(const float* src, int src_ld, const int* row_map, const int* col_map, float* dst, int rows, int cols,
                 void* stream);

/* ------------------------------------------------------------------ losses ------------- */
/* All loss kernels ADD weight*loss into *loss_accum (device scalar) and write d(loss)/d(pred)
 * (already multiplied by weight and upstream 1.0) into dpred; reductions are two-pass and
 * deterministic.  workspace >= nsr_loss_workspace() bytes. */
size_t nsr_loss_workspace(void);
/* L1Loss (basic_loss.py:45-53): weight * mean|pred - target|. */
int nsr_l1_loss(const float* pred, const float* target, float* dpred, size_t n, float weight,
                float* loss_accum, float* loss_value, void* workspace, void* stream);
/* chc_loss(huber, lambda=0, clip [cmin,cmax]) on pre-scaled inputs (basic_loss.py:180-219 as
 * used by vgg_perceptual_loss.py:145,232-236): weight * mean(clamp(sqrt((s*(a-b))^2+1e-12))). */
int nsr_charbonnier_loss(const float* a, const float* b, float* da, size_t n, float in_scale,
                         float clip_min, float clip_max, float weight, float* loss_accum,
                         float* loss_value, void* workspace, void* stream);
/* BCEWithLogits vs a constant label (gan_loss.py:62-82). */
int nsr_bce_logits_loss(const float* logits, float* dlogits, size_t n, float label, float weight,
                        float* loss_accum, float* loss_value, void* workspace, void* stream);

/* ------------------------------------------------------------------ optimizer ---------- */
/* Multi-tensor table: one entry per parameter tensor (device pointers).  The table itself
 * lives in device memory; `chunk_base` is the running count of NSR_OPT_CHUNK-element chunks
 * before this tensor (host-computed prefix sum) so a CTA can binary-search its tensor. */
#define NSR_OPT_CHUNK 4096
typedef struct NsrParamEntry {
  float* p;
  float* g;
  float* exp_avg;
  float* exp_avg_sq;
  float* exp_avg_diff;
  float* z;
  float* neg_pre_grad;
  float* ema;          /* may be NULL */
  int64_t n;
  int64_t chunk_base;
} NsrParamEntry;

/* sum of squares of all gradients -> *sumsq (device float, overwritten); deterministic
 * (fixed chunk->CTA assignment, fixed-order final reduce). */
size_t nsr_grad_sumsq_workspace(void);
int nsr_grad_sumsq(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks,
                   float* sumsq, void* workspace, void* stream);

/* Scalars of one adan_sf step, derived on the host in double precision exactly where the
 * reference derives them (adan_sf.py:176-211, 289-330) and rounded to fp32 once. */
typedef struct NsrAdanSF {
  float beta1, one_minus_beta1;
  float beta2, one_minus_beta2;
  float beta3, one_minus_beta3;
  float bias_correction3_sqrt, eps;
  float decay;            /* 1 - lr * weight_decay */
  float ckp1;             /* schedule-free interpolation weight (lerp params -> z) */
  float step_size;        /* lr * (bias_correction1 * (1 - ckp1))        [sf]  | lr / bc1        */
  float step_size_diff;   /* lr * (beta2 / bias_correction2 * (1 - ckp1)) [sf] | lr * beta2 / bc2 */
  float lr;               /* z -= lr * g */
  int32_t schedule_free;
  int32_t first_step;     /* neg_pre_grad := -g before the update (adan_sf.py:226-227) */
  float max_norm;         /* clip_grad_norm_ threshold (image.py:540-544); <= 0 disables */
  float ema_lerp;         /* 1 - decay of AveragedModel EMA; <= 0 disables */
  int32_t ema_first;      /* first EMA update copies (n_averaged == 0) */
} NsrAdanSF;
/* Fused clip_grad_norm_ + adan_sf._multi_tensor_adan + EMA lerp (image.py:540-544,642,661-662;
 * adan_sf.py:264-330).  `sumsq` is the device scalar from nsr_grad_sumsq (may be NULL when
 * max_norm <= 0).  Gradient buffers are left holding the clipped gradient, as in the reference. */
int nsr_adan_sf_step(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks,
                     const NsrAdanSF* hp, const float* sumsq, void* stream);
/* Same, with the scalars read from DEVICE memory (`hp_dev`): the launch can then sit inside a
 * captured CUDA graph while the host refreshes the per-step scalars with one small async copy. */
int nsr_adan_sf_step_dev(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks,
                         const NsrAdanSF* hp_dev, const float* sumsq, void* stream);

typedef struct NsrAdamW {
  float beta1, one_minus_beta1, beta2, one_minus_beta2;
  float eps, decay;       /* decay = 1 - lr * weight_decay */
  float step_size;        /* lr / bias_correction1 */
  float bias_correction2_sqrt;
  float max_norm, ema_lerp;
  int32_t ema_first;
} NsrAdamW;
/* torch.optim.AdamW (base.py:154-155) with the same clip + EMA fusion; uses exp_avg/exp_avg_sq. */
int nsr_adamw_step(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks,
                   const NsrAdamW* hp, const float* sumsq, void* stream);
/* Scalars from DEVICE memory, for launches inside a captured CUDA graph (see nsr_adan_sf_step_dev). */
int nsr_adamw_step_dev(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks,
                       const NsrAdamW* hp_dev, const float* sumsq, void* stream);

/* F-SAM (Friendly Sharpness-Aware Minimization, neosr/optimizers/fsam.py:36-80) on the multi-tensor table; entry fields:
 * p, g, exp_avg = fsam's `momentum`, z = fsam's `old_p`.  first_step: g -= sigma * momentum (skipped when `first`),
 * momentum = lmbda * momentum + (1 - lmbda) * g_before (or = g when `first`), norm = || (|p| if adaptive) * g ||_2 (left in
 * *sumsq as its square), old_p = p, p += (p^2 if adaptive) * g * rho / (norm + 1e-12).  restore: p = old_p (second_step;
 * the base optimizer then steps with the gradients taken at the perturbed point).  workspace >= nsr_grad_sumsq_workspace(). */
int nsr_fsam_first_step(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks, float rho, float sigma,
                        float lmbda, int adaptive, int first, float* sumsq, void* workspace, void* stream);
int nsr_fsam_restore(const NsrParamEntry* table_dev, int n_tensors, int64_t total_chunks, void* stream);

/* ------------------------------------------------------------------ OTF degradations --- */
/* The on-the-fly degradation pipeline of the `otf` model (neosr/models/otf.py:92-283).  Images are
 * NCHW fp32 planes in [0,1] (the layout feed_data receives); per-sample parameters are small DEVICE
 * arrays of `batch` floats, so no stage synchronises the host. */

/* filter2D (neosr/utils/diffjpeg.py:558-584): reflect-pad k/2, per-sample k x k correlation shared by
 * the channels of a sample.  kernel: [kernel_batch, k, k], kernel_batch == 1 (one kernel for all) or
 * == batch; k odd, <= 21 (any other k is the reference's "Wrong kernel size" ValueError). */
int nsr_filter2d(const float* img, const float* kernel, float* out, int batch, int channels, int h, int w,
                 int k, int kernel_batch, void* stream);
/* F.interpolate(..., mode = area | bilinear | bicubic), align_corners=False, no antialias
 * (otf.py:126,179-186,222-226,243-247).  mode: 0 area (adaptive average), 1 bilinear, 2 bicubic
 * (A = -0.75).  coord_scale_* is the source-coordinate scale torch uses: 1/scale_factor when the call
 * passed scale_factor (otf.py:126), in/out when it passed size. */
int nsr_resize(const float* in, float* out, int planes, int h, int w, int oh, int ow, int mode,
               float coord_scale_h, float coord_scale_w, void* stream);
/* random_add_gaussian_noise_pt(clip=True, rounds=False) (neosr/data/degradations.py:569-605,665-676):
 * out = clamp(img + N(0,1)*sigma[b]/255 mixed with ONE batch-shared [h,w] gray field for samples with
 * gray[b] == 1, 0, 1).  z ([batch,3,h,w]) / z_gray ([h,w]): optional caller-provided standard-normal
 * fields; NULL => drawn in-kernel (Philox4x32-10 keyed by seed, Box-Muller). */
int nsr_gaussian_noise(const float* img, float* out, const float* sigma, const float* gray, int any_gray,
                       const float* z, const float* z_gray, int batch, int h, int w, uint64_t seed, void* stream);
/* random_add_poisson_noise_pt(clip=True, rounds=False) (degradations.py:738-786,851-862).  The
 * per-sample `vals = 2^ceil(log2(#distinct 8-bit levels))` (torch.unique in a Python loop, 766-768,
 * 777-779) comes from a 256-bin presence bitmap built on the device.  counts_color ([batch,3,h,w]) /
 * counts_gray ([batch,1,h,w]): optional caller-provided Poisson draws for lambda = quantised image *
 * vals; NULL => drawn in-kernel (Philox; product method below lambda 10, PTRS rejection above). */
size_t nsr_poisson_noise_workspace(int batch);
int nsr_poisson_noise(const float* img, float* out, const float* scale, const float* gray, int any_gray,
                      const float* counts_color, const float* counts_gray, int batch, int h, int w,
                      uint64_t seed, void* workspace, size_t workspace_bytes, void* stream);
/* DiffJPEG(differentiable=False)(clamp(x,0,1), quality) (neosr/utils/diffjpeg.py:254-291,461-508,531-555; the
 * clamp is the torch.clamp that precedes every call, otf.py:154,232,239): zero-pad to a
 * multiple of 16, *255, RGB->YCbCr, 2x2 chroma mean, 8x8 DCT, quantise by the (transposed) tables *
 * quality_to_factor(quality[b]) with round-half-even, dequantise, IDCT, chroma nearest x2, YCbCr->RGB,
 * clamp, /255, crop — one CTA per 16x16 MCU.  quality: [batch] device floats (the reference converts
 * them to factors in a per-sample Python loop with tensor compares, 542-543). */
int nsr_jpeg(const float* img, float* out, const float* quality, int batch, int h, int w, void* stream);
/* paired_random_crop tensor branch (neosr/data/transforms.py:38-131) for one tensor, optionally fused
 * with the 8-bit quantisation of otf.py:251 (quantise != 0: clamp(round(x*255),0,255)/255). */
int nsr_crop(const float* in, float* out, int planes, int h, int w, int top, int left, int ph, int pw,
             int quantise, void* stream);
/* Training-pair pool (otf.py:37-90) without the full-pool randperm gather: for i < b,
 * out[i] = pool[slots[i]] (when dequeue != 0) and pool[slots[i]] = in[i].  slots: [b] int32 on device. */
int nsr_pool_swap(float* pool, const float* in, float* out, const int32_t* slots, int b, size_t sample_elems,
                  int dequeue, void* stream);

/* apply_augment pieces (neosr/data/augmentations.py:14-310), NCHW fp32.
 * nsr_resize_aa: clamp(F.interpolate(src[perm], size/scale, mode = bilinear | bicubic, antialias=True), 0, 1) written
 * into the window (top, left, oh, ow) of dst [batch, channels, dst_h, dst_w] (whole image when the window is the
 * image): the up/down resizes of apply_augment (257-266, 300-308) and resizemix's paste (104-122).  perm: optional
 * [batch] int32 device permutation of the source batch.  coord_scale_*: in/out, or 1/scale_factor.
 * nsr_batch_mix: mode 0  dst = lam*a + lam2*other[perm[b]], lam2 = fp32(1 - lam)   (mixup, 29-31);
 *                mode 1  dst = a, except dst[:, :, y0:y1, x0:x1] = other[perm ? perm[b] : b] (cutmix 58-59, cutblur 164). */
int nsr_resize_aa(const float* src, float* dst, const int32_t* perm, int batch, int channels, int h, int w, int oh,
                  int ow, int dst_h, int dst_w, int top, int left, int bicubic, float coord_scale_h,
                  float coord_scale_w, void* stream);
int nsr_batch_mix(const float* a, const float* other, float* dst, const int32_t* perm, int batch, int channels, int h,
                  int w, int mode, float lam, float lam2, int y0, int y1, int x0, int x1, void* stream);

/* ------------------------------------------------------------------ MS-SSIM / consistency losses --- */
/* mssim_loss (neosr/losses/ssim_loss.py:66-163) on NCHW fp32 [planes = B*C, h, w] images, built from per-scale
 * calls so the caller owns the pyramid buffers:
 *   for s in 0..4:  nsr_ssim_scale_fwd(x_s, y_s) -> sums[2s] = sum(cs), sums[2s+1] = sum(ssim), partials3_s;
 *                   x_{s+1} = nsr_avgpool2(x_s, pad = size % 2)                      (ssim_loss.py:141-143)
 *   nsr_msssim_finalize: loss = weight * (1 - prod_s mean_s^{w_s}), coef[s] = dloss/dsum_s  (device scalars)
 *   for s in 4..0:  nsr_ssim_scale_bwd -> dx_s = coef[s]*(G(P0) + 2x G(P1) + y G(P2)) + avgpool2^T(dx_{s+1}).
 * window: [window_size, window_size] fp32 on the device (the module's `gaussian_window` buffer, one channel);
 * the filter is F.conv2d with zero padding window_size/2 (ssim_loss.py:57-64).  partials3: [3, planes, h, w]. */
int nsr_avgpool2(const float* in, float* out, int planes, int h, int w, int pad_h, int pad_w, void* stream);
size_t nsr_ssim_scale_workspace(int planes, int h, int w);
int nsr_ssim_scale_fwd(const float* x, const float* y, const float* window, int window_size, float c1, float c2,
                       int use_ssim, float* partials3, float* sums2, int planes, int h, int w, void* workspace,
                       size_t workspace_bytes, void* stream);
int nsr_msssim_finalize(const float* sums, const float* counts, int nscales, float weight, float* coef,
                        float* loss_value, float* loss_accum, void* stream);
int nsr_ssim_scale_bwd(const float* partials3, const float* x, const float* y, const float* window, int window_size,
                       const float* coef, const float* dx_coarse, int coarse_h, int coarse_w, int pad_h, int pad_w,
                       float* dx, int planes, int h, int w, void* stream);
/* consistency_loss (neosr/losses/consistency_loss.py:146-192), criterion "chc".  x_blur / y_blur are
 * nsr_filter2d(nsr_clamp(x, 1/255, 1), GaussianBlur(21, 3) kernel) (reflect padding, torchvision) or the clamped
 * images when blur is off.  fwd writes the CIE-L* planes [B,h,w], the loss and — in the workspace — the column
 * cosine terms and the device flag of the `cosim < 1e-3` branch (186-190); bwd writes g_blur = dloss/dx_blur and
 * d_direct = the Oklab-chroma path gradient; the caller pushes g_blur through the blur's adjoint
 * (nsr_corr2d_zero_ext with ext = 10, flip = 1, then nsr_reflect_fold, which also adds d_direct and applies the
 * clamp mask).  The same workspace must be passed to fwd and bwd. */
int nsr_clamp(const float* in, float* out, size_t n, float lo, float hi, void* stream);
int nsr_corr2d_zero_ext(const float* img, const float* kernel, float* out, int planes, int h, int w, int k, int ext,
                        int flip, void* stream);
size_t nsr_consistency_workspace(int batch, int h, int w);
int nsr_consistency_fwd(const float* x, const float* y, const float* x_blur, const float* y_blur, float saturation,
                        float brightness, int use_cosim, float weight, float* luma_x, float* luma_y,
                        float* loss_value, float* loss_accum, int batch, int h, int w, void* workspace,
                        size_t workspace_bytes, void* stream);
int nsr_consistency_bwd(const float* x, const float* y, const float* x_blur, const float* luma_x, const float* luma_y,
                        float saturation, float weight, float* g_blur, float* d_direct, int batch, int h, int w,
                        const void* workspace, void* stream);
int nsr_reflect_fold(const float* dpad, const float* d_direct, const float* x, float* dx, int planes, int h, int w,
                     int r, void* stream);

/* ------------------------------------------------------------------ HAT pieces (neosr/archs/hat_arch.py) --- */
/* Generic (cross-)window attention over qkv [tokens = batch*h*w, 3*c] (q | k | v, heads split c as in the reference's
 * reshape):  queries = the ws x ws window, keys/values = the ows x ows window centred on it (ows == ws: HAB
 * self-attention with cyclic shift + {0,-100} mask as index math, hat_arch.py:168-215,299-350; ows > ws: OCAB,
 * keys/values are the zero-padded neighbourhood nn.Unfold(ows, stride ws, pad (ows-ws)/2) would materialise,
 * 445-515 — padded keys score `bias` and carry value 0, exactly as in the reference).  bias_table:
 * [(ws+ows-1)^2, heads]; the relative-position index is computed arithmetically (1015-1068; OCA's negative entries
 * wrap to the end of the table as Python indexing does).  out: [tokens, c]; lse: log-sum-exp per (window, head,
 * query) for the backward pass (nsr_xwin_attn_stat_floats floats).  head dim <= 32, (ws+ows-1) <= 39. */
size_t nsr_xwin_attn_stat_floats(int batch, int h, int w, int heads, int ws);
/* engine (per call; the library keeps no mutable state): NSR_ENGINE_AUTO = the mma.sync tensor-core kernels (bf16 hi/lo
 * split, 3 passes, fp32 accumulate) when the shape allows (even head dim <= 32), NSR_ENGINE_SIMT = the exact-fp32
 * CUDA-core kernels. */
int nsr_xwin_attn_fwd(const float* qkv, const float* bias_table, float* out, float* lse, int batch, int h, int w, int c,
                      int heads, int ws, int ows, int shift, int use_mask, float scale, int engine, void* stream);
size_t nsr_xwin_attn_bwd_workspace(int batch, int h, int w, int c, int heads, int ws, int ows);
/* dqkv [tokens, 3c] and dbias_table are OVERWRITTEN; deterministic (fixed-order partial sums). */
int nsr_xwin_attn_bwd(const float* qkv, const float* bias_table, const float* out, const float* dout, const float* lse,
                      float* dqkv, float* dbias_table, int batch, int h, int w, int c, int heads, int ws, int ows,
                      int shift, int use_mask, float scale, int engine, void* workspace, size_t workspace_bytes, void* stream);
/* ChannelAttention (hat_arch.py:15-37) on NHWC [batch, hw, c]:
 *   pooled = nsr_channel_mean(x, NULL, scale = 1/hw)            AdaptiveAvgPool2d(1)
 *   gate   = sigmoid(W2 relu(W1 pooled + b1) + b2)              the two 1x1 convs (w1: [cs, c], w2: [c, cs])
 *   y (+)= alpha * x * gate                                     `x * y`, fused with HAB's `conv_x * conv_scale` (349)
 * backward: dgate = nsr_channel_mean(g, x, scale = alpha) (sum of g*x), nsr_channel_gate_bwd (parameter gradients
 * overwritten, summed over the batch in order), dx = alpha * g * gate + dpooled / hw. */
int nsr_channel_mean(const float* x, const float* mul, float* pooled, int batch, int hw, int c, float scale, void* stream);
int nsr_channel_gate_fwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                         float* hidden, float* gate, int batch, int c, int cs, void* stream);
int nsr_channel_scale_add(const float* x, const float* gate, float* y, int batch, int hw, int c, float alpha,
                          int accumulate, void* stream);
int nsr_channel_gate_bwd(const float* dgate, const float* gate, const float* hidden, const float* pooled, const float* w1,
                         const float* w2, float* dpooled, float* dw1, float* db1, float* dw2, float* db2, int batch,
                         int c, int cs, void* stream);
int nsr_channel_scale_bwd(const float* g, const float* gate, const float* dpooled, float* dx, int batch, int hw, int c,
                          float alpha, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEOSR_B200_H */

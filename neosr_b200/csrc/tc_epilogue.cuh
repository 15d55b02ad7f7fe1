// Shared epilogue of the tcgen05 contraction kernels: one 32-column chunk of this warp's 32
// accumulator rows, TMEM -> registers -> per-warp smem transpose -> row-contiguous global traffic.
#pragma once
#include "tc_common.cuh"

namespace nsr {
namespace tc {

constexpr int EPI_LD = 36;  // padded row stride (floats) of the per-warp transpose tile [32][EPI_LD]

// byte offset (hi image) of the 8-byte half-chunk holding channels col..col+3 (col % 4 == 0) of row p
__device__ __forceinline__ size_t sti_offset(long long p, int col, int kbs) {
  const long long mt = p >> 7;
  const int r = (int)(p & 127);
  const int kb = col >> 6, cc = col & 63;
  return ((size_t)(mt * kbs + kb) << 15) + (size_t)(r * 128 + (((cc >> 3) ^ (r & 7)) << 4) + ((cc >> 2) & 1) * 8);
}

// Window-ordered split tile images (d.sti_win = ws | shift << 16, see NsrConv): token row p = (b, y, x) of the image is
// stored at row  p' = ((b * H/ws + wy) * W/ws + wx) * ws^2 + iy * ws + ix  with (ys, xs) = ((y - shift) mod H, (x - shift)
// mod W) the coordinates after torch.roll(-shift), (wy, iy) = divmod(ys, ws), (wx, ix) = divmod(xs, ws): the rows of one
// attention window are contiguous, in the order window_partition (swinir_arch.py:41-57) flattens them.
__device__ __forceinline__ int sti_win_row(const NsrConv& d, long long p, int hw) {
  const int ws = d.sti_win & 0xFFFF, shift = d.sti_win >> 16;
  const int b = (int)(p / hw), rem = (int)(p - (long long)b * hw);
  const int y = rem / d.w, x = rem - y * d.w;
  int ys = y - shift, xs = x - shift;
  if (ys < 0) ys += d.h;
  if (xs < 0) xs += d.w;
  const int wy = ys / ws, iy = ys - wy * ws, wx = xs / ws, ix = xs - wx * ws;
  return ((b * (d.h / ws) + wy) * (d.w / ws) + wx) * (ws * ws) + iy * ws + ix;
}

// The per-row work of one chunk, specialised at compile time on (activation, activation-gradient) so the
// 8x-unrolled loop carries no per-element switch.  Lane = (row-in-group er, 4 columns starting at ec); group i is row
// p0 + 4 i + er.  p0 is a multiple of 32 (tiles start at multiples of 128 rows, warps at multiples of 32), so the 32
// rows share one 128-row block of a split tile image and every address is a base + compile-time multiple of a stride.
// MODE >= 0 additionally fixes which outputs exist (bit 0: y_pre, 1: y_pre holds the activation gradient, 2: residual,
// 3: y, 4: split tile image; no row_scale), removing the warp-uniform tests from the unrolled loop; -1 reads them from d.
// U16: the activation gradient travels as a 16-bit fixed-point code instead of fp32 (NsrConv.pre_mode / aux_mode = 2):
//   code = rint((g + 0.25) * 40000),  g = code / 40000 - 0.25     (|error| <= 1.25e-5 on gelu' in [-0.13, 1.13])
// halving the traffic of the fc1 -> fc2-dgrad hand-over, the largest fp32 stream of a Swin block.
constexpr float AGC_SCALE = 40000.f, AGC_BIAS = 0.25f;
__device__ __forceinline__ uint32_t agc_pack2(float a, float b) {
  const uint32_t qa = (uint32_t)__float2int_rn(fminf(fmaxf((a + AGC_BIAS) * AGC_SCALE, 0.f), 65535.f));
  const uint32_t qb = (uint32_t)__float2int_rn(fminf(fmaxf((b + AGC_BIAS) * AGC_SCALE, 0.f), 65535.f));
  return qa | (qb << 16);
}
__device__ __forceinline__ float4 agc_unpack4(uint2 q) {
  constexpr float inv = 1.f / AGC_SCALE;
  return make_float4(fmaf((float)(q.x & 0xFFFFu), inv, -AGC_BIAS), fmaf((float)(q.x >> 16), inv, -AGC_BIAS),
                     fmaf((float)(q.y & 0xFFFFu), inv, -AGC_BIAS), fmaf((float)(q.y >> 16), inv, -AGC_BIAS));
}

template <int ACT, int AG, int MODE = -1, bool WIN = false, bool U16 = false>
__device__ __forceinline__ void epi_rows(const NsrConv& d, const float* stg, long long p0, int n, bool ncol, bool nsti,
                                         long long M, int hw, int er, int ec, int kbs_out, const int (&wrow)[8]) {
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(d.act_slope, d.act_slope, d.act_slope, d.act_slope);
  float4 g4 = make_float4(d.actgrad_slope, d.actgrad_slope, d.actgrad_slope, d.actgrad_slope);
  if (ncol && d.bias) b4 = __ldg(reinterpret_cast<const float4*>(d.bias + n));
  if (ncol && (ACT == NSR_ACT_PRELU)) s4 = __ldg(reinterpret_cast<const float4*>(d.prelu + n));
  if (ncol && (AG == NSR_ACT_PRELU)) g4 = __ldg(reinterpret_cast<const float4*>(d.prelu + n));
  const long long pb = p0 + er, left = M - pb;
  const int nval = left <= 0 ? 0 : (left > 28 ? 8 : (int)((left + 3) >> 2));  // groups i < nval have a row < M
  const bool has_pre = MODE >= 0 ? (MODE & 1) != 0 : d.y_pre != nullptr;
  const bool pre_grad = MODE >= 0 ? (MODE & 2) != 0 : d.pre_mode != 0;
  const bool has_res = MODE >= 0 ? (MODE & 4) != 0 : d.residual != nullptr;
  const bool has_y = MODE >= 0 ? (MODE & 8) != 0 : d.y != nullptr;
  const bool has_rs = MODE >= 0 ? false : d.row_scale != nullptr;
  if (MODE >= 0) nsti = nsti && (MODE & 16) != 0;
  const long long yo = pb * d.y_ld + n;
  float* const yp = has_y ? d.y + yo : nullptr;
  float* const prep = has_pre ? d.y_pre + yo : nullptr;
  uint16_t* const prep16 = has_pre ? reinterpret_cast<uint16_t*>(d.y_pre) + yo : nullptr;
  const int res_ld = d.res_ld ? d.res_ld : d.y_ld, aux_ld = d.aux_ld ? d.aux_ld : d.y_ld;
  const float* const resp = has_res ? d.residual + pb * res_ld + n : nullptr;
  const float* const auxp = (AG != NSR_ACT_NONE) ? d.aux + pb * aux_ld + n : nullptr;
  const uint16_t* const auxp16 = (AG != NSR_ACT_NONE) ? reinterpret_cast<const uint16_t*>(d.aux) + pb * aux_ld + n : nullptr;
  const int ystep = 4 * d.y_ld, rstep = 4 * res_ld, astep = 4 * aux_ld;
  // split tile image: 16-byte chunk index is swizzled with (row & 7) = er + 4 (i & 1)
  uint8_t* sp = nullptr;
  int sw0 = 0, sw1 = 0;
  constexpr bool win = WIN;  // window-ordered image: every row has its own block / swizzle phase (wrow[i])
  if (nsti) {
    const int cc = n & 63;
    if (win) {
      sp = reinterpret_cast<uint8_t*>(d.y_sti) + ((size_t)(n >> 6) << 15) + ((cc >> 2) & 1) * 8;
      sw0 = cc >> 3;
    } else {
      sp = reinterpret_cast<uint8_t*>(d.y_sti) + ((size_t)((pb >> 7) * kbs_out + (n >> 6)) << 15) + (size_t)(pb & 127) * 128 +
           ((cc >> 2) & 1) * 8;
      sw0 = ((cc >> 3) ^ er) << 4;
      sw1 = ((cc >> 3) ^ (er + 4)) << 4;
    }
  }
  const float one0 = n == d.cout ? 1.f : 0.f;  // first padding channel of an STI output = 1 (bias-gradient column)
#pragma unroll
  for (int half = 0; half < 2; ++half) {  // two batches of 4 row-groups: 8 independent 16-byte loads in flight
    float4 aux4[4], res4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = half * 4 + j;
      const bool ok = ncol && i < nval;
      if (AG != NSR_ACT_NONE) {
        if (U16) aux4[j] = ok ? agc_unpack4(*reinterpret_cast<const uint2*>(auxp16 + i * astep)) : make_float4(0.f, 0.f, 0.f, 0.f);
        else aux4[j] = ok ? *reinterpret_cast<const float4*>(auxp + i * astep) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      res4[j] = (ok && has_res) ? *reinterpret_cast<const float4*>(resp + i * rstep) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = half * 4 + j;
      if (i >= nval) continue;
      float ov[4] = {one0, 0.f, 0.f, 0.f};
      if (ncol) {
        const float4 a4 = *reinterpret_cast<const float4*>(stg + (i * 4 + er) * EPI_LD + ec);
        const float pre[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
        const float sl[4] = {s4.x, s4.y, s4.z, s4.w};
        float gr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) act_value_grad<ACT>(pre[e], sl[e], ov[e], gr[e]);
        if (has_pre) {
          if (U16) *reinterpret_cast<uint2*>(prep16 + i * ystep) = make_uint2(agc_pack2(gr[0], gr[1]), agc_pack2(gr[2], gr[3]));
          else if (pre_grad) *reinterpret_cast<float4*>(prep + i * ystep) = make_float4(gr[0], gr[1], gr[2], gr[3]);
          else *reinterpret_cast<float4*>(prep + i * ystep) = make_float4(pre[0], pre[1], pre[2], pre[3]);
        }
        if (AG != NSR_ACT_NONE) {
          ov[0] *= act_grad_ct<AG>(aux4[j].x, g4.x); ov[1] *= act_grad_ct<AG>(aux4[j].y, g4.y);
          ov[2] *= act_grad_ct<AG>(aux4[j].z, g4.z); ov[3] *= act_grad_ct<AG>(aux4[j].w, g4.w);
        }
        if (has_rs) {
          const float rs = d.row_scale[(pb + 4 * i) / hw];
#pragma unroll
          for (int e = 0; e < 4; ++e) ov[e] *= rs;
        }
        if (has_res) { ov[0] += res4[j].x; ov[1] += res4[j].y; ov[2] += res4[j].z; ov[3] += res4[j].w; }
        if (has_y) *reinterpret_cast<float4*>(yp + i * ystep) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      }
      if (nsti) {  // channels in [cout, kbs_out*64) are written as zeros (K padding of the next contraction)
        uint2 hi, lo;
        split2(ov[0], ov[1], hi.x, lo.x);
        split2(ov[2], ov[3], hi.y, lo.y);
        uint8_t* dst;
        if (win) {
          const int pr = wrow[i];
          dst = sp + ((size_t)((pr >> 7) * kbs_out) << 15) + (pr & 127) * 128 + ((sw0 ^ (pr & 7)) << 4);
        } else {
          dst = sp + i * 512 + ((i & 1) ? sw1 : sw0);
        }
        *reinterpret_cast<uint2*>(dst) = hi;
        *reinterpret_cast<uint2*>(dst + 16384) = lo;
      }
    }
  }
}

// d: contraction descriptor (epilogue fields), stg: this warp's [32][EPI_LD] fp32 tile,
// taddr: TMEM address of (lane quarter, first column of the chunk), p0: first row of this warp,
// nc0: first output channel of the chunk, kbs_out: 64-channel blocks of the STI output (0 if none)
// wrow_lane: window-order row of tile row p0 + lane (sti_win_row; WIN = the kernel instantiation for d.sti_win != 0)
template <bool WIN = false>
__device__ __forceinline__ void epi_chunk(const NsrConv& d, float* stg, uint32_t taddr, long long p0, int nc0,
                                          long long M_rows, int hw, int lane, int kbs_out, int wrow_lane = 0) {
  const int er = lane >> 3, ec = (lane & 7) * 4;
  int wrow[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (WIN) {  // before any per-lane exit
#pragma unroll
    for (int i = 0; i < 8; ++i) wrow[i] = __shfl_sync(0xffffffffu, wrow_lane, er + 4 * i);
  }
  float v[32];
  tmem_ld_32x32(taddr, v);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    *reinterpret_cast<float4*>(stg + lane * EPI_LD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  __syncwarp();
  const int n = nc0 + ec;
  const bool ncol = n < d.cout;
  const bool nsti = d.y_sti != nullptr && n < kbs_out * 64;
  if (!ncol && !nsti) return;
  // the output combinations of the SwinIR / VGG hot path get loops without warp-uniform tests
  const int mode = d.row_scale ? -1
                               : (d.y_pre ? 1 : 0) | (d.y_pre && d.pre_mode ? 2 : 0) | (d.residual ? 4 : 0) | (d.y ? 8 : 0) |
                                     (d.y_sti ? 16 : 0);
  if (d.pre_mode == 2 || d.aux_mode == 2) {  // 16-bit activation-gradient codes: the two MLP hand-over epilogues only (host-checked)
    if (d.act == NSR_ACT_GELU) epi_rows<NSR_ACT_GELU, NSR_ACT_NONE, 19, WIN, true>(d, stg, p0, n, ncol, nsti, M_rows, hw, er, ec, kbs_out, wrow);
    else epi_rows<NSR_ACT_NONE, NSR_ACT_MULAUX, 16, WIN, true>(d, stg, p0, n, ncol, nsti, M_rows, hw, er, ec, kbs_out, wrow);
    return;
  }
#define NSR_EPI_HOT(A, G, M)                                                         \
  if (d.act == (A) && d.actgrad == (G) && mode == (M)) {                             \
    epi_rows<A, G, M, WIN>(d, stg, p0, n, ncol, nsti, M_rows, hw, er, ec, kbs_out, wrow); \
    return;                                                                          \
  }
  NSR_EPI_HOT(NSR_ACT_GELU, NSR_ACT_NONE, 19)    // fc1: gelu'(pre) + STI of gelu(pre)
  NSR_EPI_HOT(NSR_ACT_NONE, NSR_ACT_MULAUX, 16)  // fc2 dgrad: dy * gelu' -> STI
  NSR_EPI_HOT(NSR_ACT_NONE, NSR_ACT_NONE, 12)    // proj / fc2 / RSTB conv: + residual -> fp32
  NSR_EPI_HOT(NSR_ACT_NONE, NSR_ACT_NONE, 8)     // qkv, plain dgrads -> fp32
  NSR_EPI_HOT(NSR_ACT_NONE, NSR_ACT_NONE, 16)    // qkv / proj dgrad -> (window-ordered) STI for the attention kernels
  NSR_EPI_HOT(NSR_ACT_RELU, NSR_ACT_NONE, 8)     // VGG conv + ReLU
  NSR_EPI_HOT(NSR_ACT_RELU, NSR_ACT_NONE, 9)     // VGG tapped layers: pre-activation kept
  NSR_EPI_HOT(NSR_ACT_NONE, NSR_ACT_RELU, 8)     // VGG dgrad chain
#undef NSR_EPI_HOT
#define NSR_EPI_CASE(A, G) \
  case (A) * 8 + (G): epi_rows<A, G, -1, WIN>(d, stg, p0, n, ncol, nsti, M_rows, hw, er, ec, kbs_out, wrow); break;
  switch (d.act * 8 + d.actgrad) {  // warp-uniform
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_NONE)
    NSR_EPI_CASE(NSR_ACT_GELU, NSR_ACT_NONE)
    NSR_EPI_CASE(NSR_ACT_RELU, NSR_ACT_NONE)
    NSR_EPI_CASE(NSR_ACT_LRELU, NSR_ACT_NONE)
    NSR_EPI_CASE(NSR_ACT_PRELU, NSR_ACT_NONE)
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_MULAUX)
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_GELU)
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_RELU)
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_LRELU)
    NSR_EPI_CASE(NSR_ACT_NONE, NSR_ACT_PRELU)
    default: break;  // act and actgrad together are rejected on the host for this engine
  }
#undef NSR_EPI_CASE
}

}  // namespace tc
}  // namespace nsr

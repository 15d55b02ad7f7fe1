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

// The per-row work of one chunk, specialised at compile time on (activation, activation-gradient) so the
// 8x-unrolled loop carries no per-element switch.  Lane = (row-in-group er, 4 columns starting at ec).
template <int ACT, int AG>
__device__ __forceinline__ void epi_rows(const NsrConv& d, const float* stg, long long p0, int n, bool ncol, bool nsti,
                                         long long M, int hw, int er, int ec, int kbs_out) {
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(d.act_slope, d.act_slope, d.act_slope, d.act_slope);
  float4 g4 = make_float4(d.actgrad_slope, d.actgrad_slope, d.actgrad_slope, d.actgrad_slope);
  if (ncol && d.bias) b4 = __ldg(reinterpret_cast<const float4*>(d.bias + n));
  if (ncol && (ACT == NSR_ACT_PRELU)) s4 = __ldg(reinterpret_cast<const float4*>(d.prelu + n));
  if (ncol && (AG == NSR_ACT_PRELU)) g4 = __ldg(reinterpret_cast<const float4*>(d.prelu + n));
#pragma unroll
  for (int half = 0; half < 2; ++half) {  // two batches of 4 row-groups: 8 independent 16-byte loads in flight
  float4 aux4[8], res4[8];
#pragma unroll
  for (int i = half * 4; i < half * 4 + 4; ++i) {
    const long long p = p0 + i * 4 + er;
    const bool ok = ncol && p < M;
    const long long ores = p * (d.res_ld ? d.res_ld : d.y_ld) + n, oaux = p * (d.aux_ld ? d.aux_ld : d.y_ld) + n;
    if (AG != NSR_ACT_NONE) aux4[i] = ok ? *reinterpret_cast<const float4*>(d.aux + oaux) : make_float4(0.f, 0.f, 0.f, 0.f);
    res4[i] = (ok && d.residual) ? *reinterpret_cast<const float4*>(d.residual + ores) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = half * 4; i < half * 4 + 4; ++i) {
    const long long p = p0 + i * 4 + er;
    if (p >= M) continue;
    float ov[4] = {n == d.cout ? 1.f : 0.f, 0.f, 0.f, 0.f};  // first padding channel of an STI output = 1 (bias-gradient column)
    if (ncol) {
      const long long o = p * d.y_ld + n;
      const float4 a4 = *reinterpret_cast<const float4*>(stg + (i * 4 + er) * EPI_LD + ec);
      const float pre[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
      const float sl[4] = {s4.x, s4.y, s4.z, s4.w};
      float gr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) act_value_grad<ACT>(pre[e], sl[e], ov[e], gr[e]);
      if (d.y_pre) {
        if (d.pre_mode) *reinterpret_cast<float4*>(d.y_pre + o) = make_float4(gr[0], gr[1], gr[2], gr[3]);
        else *reinterpret_cast<float4*>(d.y_pre + o) = make_float4(pre[0], pre[1], pre[2], pre[3]);
      }
      if (AG != NSR_ACT_NONE) {
        ov[0] *= act_grad_ct<AG>(aux4[i].x, g4.x); ov[1] *= act_grad_ct<AG>(aux4[i].y, g4.y);
        ov[2] *= act_grad_ct<AG>(aux4[i].z, g4.z); ov[3] *= act_grad_ct<AG>(aux4[i].w, g4.w);
      }
      if (d.row_scale) {
        const float rs = d.row_scale[p / hw];
#pragma unroll
        for (int e = 0; e < 4; ++e) ov[e] *= rs;
      }
      if (d.residual) { ov[0] += res4[i].x; ov[1] += res4[i].y; ov[2] += res4[i].z; ov[3] += res4[i].w; }
      if (d.y) *reinterpret_cast<float4*>(d.y + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
    if (nsti) {  // channels in [cout, kbs_out*64) are written as zeros (K padding of the next contraction)
      uint2 hi, lo;
      split2(ov[0], ov[1], hi.x, lo.x);
      split2(ov[2], ov[3], hi.y, lo.y);
      uint8_t* dst = reinterpret_cast<uint8_t*>(d.y_sti) + sti_offset(p, n, kbs_out);
      *reinterpret_cast<uint2*>(dst) = hi;
      *reinterpret_cast<uint2*>(dst + 16384) = lo;
    }
  }
  }
}

// d: contraction descriptor (epilogue fields), stg: this warp's [32][EPI_LD] fp32 tile,
// taddr: TMEM address of (lane quarter, first column of the chunk), p0: first row of this warp,
// nc0: first output channel of the chunk, kbs_out: 64-channel blocks of the STI output (0 if none)
__device__ __forceinline__ void epi_chunk(const NsrConv& d, float* stg, uint32_t taddr, long long p0, int nc0,
                                          long long M, int hw, int lane, int kbs_out) {
  const int er = lane >> 3, ec = (lane & 7) * 4;
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
#define NSR_EPI_CASE(A, G) \
  case (A) * 8 + (G): epi_rows<A, G>(d, stg, p0, n, ncol, nsti, M, hw, er, ec, kbs_out); break;
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

// Image-side convolutions with <= 4 channels on one side (conv_first 3->C, conv_last C->3, VGG
// conv1_1 and their dgrad/wgrad).  K or N of the implicit GEMM is 3, so tensor-core tiles would
// be >97% padding: these are HBM-bound streaming kernels instead.  A warp owns a pixel; lanes own
// channel pairs of the wide side (coalesced 256-byte rows), the <= 4-channel side is broadcast.
#include "common.cuh"

namespace nsr {

constexpr int CS_WARPS = 8;
constexpr int CS_WG_WARPS = 4;  // wgrad kernels: 4 warps keep the block-reduce tile under 48 KiB of static smem
constexpr int CS_MAXTAPS = 9;

struct SmallGeom {
  int B, H, W, kh, kw, pad, wide, narrow;  // wide: channels of the wide side (multiple of 64 or <= 256), narrow <= 4
  int x_ld, y_ld;
  long long M;
};

__device__ __forceinline__ void pix_decode(long long p, int H, int W, int& oh, int& ow) {
  const int hw = H * W;
  const int rem = (int)(p % hw);
  oh = rem / W;
  ow = rem - oh * W;
}

// ---------------------------------------------------------------- wide -> narrow (cout <= 4) fprop
// y[p, co] = bias[co] + sum_{tap, ci} x[p @ tap, ci] * w[co][tap][ci];  w is the packed fp32 view.
template <int NARROW>
__global__ void __launch_bounds__(CS_WARPS * 32) conv_wide2narrow(const float* __restrict__ x,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 SmallGeom g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int taps = g.kh * g.kw;
  const int groups = g.wide / 64;
  const long long nwarps = (long long)gridDim.x * CS_WARPS;
  for (int grp = 0; grp < groups; ++grp) {
    const int c = grp * 64 + lane * 2;
    float wr[NARROW][CS_MAXTAPS][2];
#pragma unroll
    for (int co = 0; co < NARROW; ++co)
#pragma unroll
      for (int t = 0; t < CS_MAXTAPS; ++t) {
        const bool ok = t < taps;
        wr[co][t][0] = ok ? w[((size_t)co * taps + t) * g.wide + c] : 0.f;
        wr[co][t][1] = ok ? w[((size_t)co * taps + t) * g.wide + c + 1] : 0.f;
      }
    for (long long p = (long long)blockIdx.x * CS_WARPS + warp; p < g.M; p += nwarps) {
      int oh, ow;
      pix_decode(p, g.H, g.W, oh, ow);
      float acc[NARROW];
#pragma unroll
      for (int co = 0; co < NARROW; ++co) acc[co] = 0.f;
#pragma unroll
      for (int t = 0; t < CS_MAXTAPS; ++t) {
        if (t >= taps) break;
        const int r = t / g.kw, s = t - r * g.kw;
        const int ih = oh + r - g.pad, iw = ow + s - g.pad;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;  // warp-uniform
        const float2 v = *reinterpret_cast<const float2*>(x + (p + (long long)(r - g.pad) * g.W + (s - g.pad)) * g.x_ld + c);
#pragma unroll
        for (int co = 0; co < NARROW; ++co) acc[co] = fmaf(v.x, wr[co][t][0], fmaf(v.y, wr[co][t][1], acc[co]));
      }
#pragma unroll
      for (int co = 0; co < NARROW; ++co) acc[co] = warp_sum(acc[co]);
      if (lane < NARROW) {
        float v = 0.f;
#pragma unroll
        for (int co = 0; co < NARROW; ++co) v = lane == co ? acc[co] : v;
        if (grp == 0) v += bias ? bias[lane] : 0.f;
        else v += y[p * g.y_ld + lane];
        y[p * g.y_ld + lane] = v;
      }
    }
  }
}

// ---------------------------------------------------------------- wide -> narrow wgrad
// dw[co][tap][ci] = sum_p dy[p, co] * x[p @ tap, ci]  ==  sum_q x[q, ci] * dy[q - tapoffset, co]
// The warp walks INPUT pixels q (one coalesced 256-byte row each); the 9 x NARROW dy neighbours are
// broadcast loads.  54-72 accumulators per lane, block-reduced, then a fixed-order final reduce.
template <int NARROW>
__global__ void __launch_bounds__(CS_WG_WARPS * 32) wgrad_wide2narrow(const float* __restrict__ x,
                                                                  const float* __restrict__ dy,
                                                                  float* __restrict__ partial, SmallGeom g, int grp) {
  __shared__ float red[CS_WG_WARPS][NARROW * CS_MAXTAPS * 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int taps = g.kh * g.kw;
  const int c = grp * 64 + lane * 2;
  float acc[NARROW][CS_MAXTAPS][2];
#pragma unroll
  for (int co = 0; co < NARROW; ++co)
#pragma unroll
    for (int t = 0; t < CS_MAXTAPS; ++t) acc[co][t][0] = acc[co][t][1] = 0.f;
  const long long nwarps = (long long)gridDim.x * CS_WG_WARPS;
  for (long long q = (long long)blockIdx.x * CS_WG_WARPS + warp; q < g.M; q += nwarps) {
    int qh, qw;
    pix_decode(q, g.H, g.W, qh, qw);
    const float2 v = *reinterpret_cast<const float2*>(x + q * g.x_ld + c);
#pragma unroll
    for (int t = 0; t < CS_MAXTAPS; ++t) {
      if (t >= taps) break;
      const int r = t / g.kw, s = t - r * g.kw;
      const int oh = qh - (r - g.pad), ow = qw - (s - g.pad);  // output pixel that reads q through tap t
      if (oh < 0 || oh >= g.H || ow < 0 || ow >= g.W) continue;
      const float* d = dy + (q - (long long)(r - g.pad) * g.W - (s - g.pad)) * g.y_ld;
#pragma unroll
      for (int co = 0; co < NARROW; ++co) {
        const float dv = __ldg(d + co);
        acc[co][t][0] = fmaf(dv, v.x, acc[co][t][0]);
        acc[co][t][1] = fmaf(dv, v.y, acc[co][t][1]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < NARROW; ++co)
#pragma unroll
    for (int t = 0; t < CS_MAXTAPS; ++t) {
      red[warp][(co * CS_MAXTAPS + t) * 64 + lane * 2] = acc[co][t][0];
      red[warp][(co * CS_MAXTAPS + t) * 64 + lane * 2 + 1] = acc[co][t][1];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NARROW * CS_MAXTAPS * 64; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < CS_WG_WARPS; ++wv) s += red[wv][i];
    partial[(size_t)blockIdx.x * (NARROW * CS_MAXTAPS * 64) + i] = s;
  }
}
// partial[block][co][tap(9)][64] -> dw[co][ci][tap] (OIHW) for channel group grp
__global__ void wgrad_w2n_final(const float* __restrict__ partial, float* __restrict__ dw, int blocks, int narrow,
                                int taps, int wide, int grp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= narrow * taps * 64) return;
  const int cl = i % 64, t = (i / 64) % taps, co = i / (64 * taps);
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[(size_t)b * (narrow * CS_MAXTAPS * 64) + (co * CS_MAXTAPS + t) * 64 + cl];
  dw[((size_t)co * wide + grp * 64 + cl) * taps + t] = s;
}

// ---------------------------------------------------------------- narrow -> wide (cin <= 4) fprop
// y[p, co] = act(bias[co] + sum_{tap, ci} x[p @ tap, ci] * w[co][tap][ci]); lanes own cout pairs.
template <int NARROW>
__global__ void __launch_bounds__(CS_WARPS * 32) conv_narrow2wide(const float* __restrict__ x,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 SmallGeom g, int act, float slope,
                                                                 const float* __restrict__ prelu, float* __restrict__ y_pre) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int taps = g.kh * g.kw;
  const int groups = (g.wide + 63) / 64;
  const long long nwarps = (long long)gridDim.x * CS_WARPS;
  for (int grp = 0; grp < groups; ++grp) {
    const int c = grp * 64 + lane * 2;
    const bool c0 = c < g.wide, c1 = c + 1 < g.wide;
    float wr[CS_MAXTAPS][NARROW][2];
#pragma unroll
    for (int t = 0; t < CS_MAXTAPS; ++t)
#pragma unroll
      for (int ci = 0; ci < NARROW; ++ci) {
        wr[t][ci][0] = (t < taps && c0) ? w[((size_t)c * taps + t) * NARROW + ci] : 0.f;
        wr[t][ci][1] = (t < taps && c1) ? w[((size_t)(c + 1) * taps + t) * NARROW + ci] : 0.f;
      }
    const float b0 = (bias && c0) ? bias[c] : 0.f, b1 = (bias && c1) ? bias[c + 1] : 0.f;
    const float s0 = (act == NSR_ACT_PRELU && c0) ? prelu[c] : slope, s1 = (act == NSR_ACT_PRELU && c1) ? prelu[c + 1] : slope;
    for (long long p = (long long)blockIdx.x * CS_WARPS + warp; p < g.M; p += nwarps) {
      int oh, ow;
      pix_decode(p, g.H, g.W, oh, ow);
      float a0 = b0, a1 = b1;
#pragma unroll
      for (int t = 0; t < CS_MAXTAPS; ++t) {
        if (t >= taps) break;
        const int r = t / g.kw, s = t - r * g.kw;
        const int ih = oh + r - g.pad, iw = ow + s - g.pad;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
        const float* xp = x + (p + (long long)(r - g.pad) * g.W + (s - g.pad)) * g.x_ld;
#pragma unroll
        for (int ci = 0; ci < NARROW; ++ci) {
          const float v = __ldg(xp + ci);
          a0 = fmaf(v, wr[t][ci][0], a0);
          a1 = fmaf(v, wr[t][ci][1], a1);
        }
      }
      if (y_pre) {
        if (c0) y_pre[p * g.y_ld + c] = a0;
        if (c1) y_pre[p * g.y_ld + c + 1] = a1;
      }
      if (act) { a0 = apply_act(a0, act, s0); a1 = apply_act(a1, act, s1); }
      float* yp = y + p * g.y_ld + c;
      if (c1 && (g.y_ld % 2 == 0)) *reinterpret_cast<float2*>(yp) = make_float2(a0, a1);
      else { if (c0) yp[0] = a0; if (c1) yp[1] = a1; }
    }
  }
}

// ---------------------------------------------------------------- narrow -> wide wgrad (conv_first)
// dw[co][ci][tap] = sum_p dy[p, co] * x[p @ tap, ci];  lanes own cout pairs of one 64-channel group.
template <int NARROW>
__global__ void __launch_bounds__(CS_WG_WARPS * 32) wgrad_narrow2wide(const float* __restrict__ x,
                                                                  const float* __restrict__ dy,
                                                                  float* __restrict__ partial, SmallGeom g, int grp) {
  __shared__ float red[CS_WG_WARPS][CS_MAXTAPS * NARROW * 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int taps = g.kh * g.kw;
  const int c = grp * 64 + lane * 2;
  const bool c0 = c < g.wide, c1 = c + 1 < g.wide;
  float acc[CS_MAXTAPS][NARROW][2];
#pragma unroll
  for (int t = 0; t < CS_MAXTAPS; ++t)
#pragma unroll
    for (int ci = 0; ci < NARROW; ++ci) acc[t][ci][0] = acc[t][ci][1] = 0.f;
  const long long nwarps = (long long)gridDim.x * CS_WG_WARPS;
  for (long long p = (long long)blockIdx.x * CS_WG_WARPS + warp; p < g.M; p += nwarps) {
    int oh, ow;
    pix_decode(p, g.H, g.W, oh, ow);
    const float d0 = c0 ? dy[p * g.y_ld + c] : 0.f, d1 = c1 ? dy[p * g.y_ld + c + 1] : 0.f;
#pragma unroll
    for (int t = 0; t < CS_MAXTAPS; ++t) {
      if (t >= taps) break;
      const int r = t / g.kw, s = t - r * g.kw;
      const int ih = oh + r - g.pad, iw = ow + s - g.pad;
      if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
      const float* xp = x + (p + (long long)(r - g.pad) * g.W + (s - g.pad)) * g.x_ld;
#pragma unroll
      for (int ci = 0; ci < NARROW; ++ci) {
        const float v = __ldg(xp + ci);
        acc[t][ci][0] = fmaf(d0, v, acc[t][ci][0]);
        acc[t][ci][1] = fmaf(d1, v, acc[t][ci][1]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < CS_MAXTAPS; ++t)
#pragma unroll
    for (int ci = 0; ci < NARROW; ++ci) {
      red[warp][(t * NARROW + ci) * 64 + lane * 2] = acc[t][ci][0];
      red[warp][(t * NARROW + ci) * 64 + lane * 2 + 1] = acc[t][ci][1];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < CS_MAXTAPS * NARROW * 64; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < CS_WG_WARPS; ++wv) s += red[wv][i];
    partial[(size_t)blockIdx.x * (CS_MAXTAPS * NARROW * 64) + i] = s;
  }
}
__global__ void wgrad_n2w_final(const float* __restrict__ partial, float* __restrict__ dw, int blocks, int narrow,
                                int taps, int wide, int grp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * narrow * 64) return;
  const int cl = i % 64, ci = (i / 64) % narrow, t = i / (64 * narrow);
  const int co = grp * 64 + cl;
  if (co >= wide) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[(size_t)b * (CS_MAXTAPS * narrow * 64) + (t * narrow + ci) * 64 + cl];
  dw[((size_t)co * narrow + ci) * taps + t] = s;
}

// ---------------------------------------------------------------- host
#define CS_NARROW_SWITCH(n, LAUNCH)                      \
  switch (n) {                                           \
    case 1: { constexpr int NW = 1; LAUNCH; } break;     \
    case 2: { constexpr int NW = 2; LAUNCH; } break;     \
    case 3: { constexpr int NW = 3; LAUNCH; } break;     \
    default: { constexpr int NW = 4; LAUNCH; } break;    \
  }
static int cs_blocks(long long M) {
  long long b = (M + CS_WARPS - 1) / CS_WARPS;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}
constexpr int CS_WGRAD_BLOCKS = kNumSMs * 4;

bool conv_small_fprop_supported(const NsrConv& d) {
  if (d.x == nullptr || d.y == nullptr || d.y_sti || d.x_sti) return false;
  if (d.kh * d.kw > CS_MAXTAPS) return false;
  if (d.actgrad || d.residual || d.row_scale || d.act == NSR_ACT_GELU || d.pre_mode) return false;
  if (d.cout <= 4 && d.cout >= 1 && d.cin % 64 == 0 && d.x_ld % 2 == 0 && d.act == NSR_ACT_NONE && !d.y_pre) return true;
  if (d.cin <= 4 && d.cin >= 1 && d.cout >= 16) return true;
  return false;
}

int conv_small_fprop(const NsrConv& d, cudaStream_t st) {
  SmallGeom g;
  g.B = d.batch; g.H = d.h; g.W = d.w; g.kh = d.kh; g.kw = d.kw; g.pad = d.pad;
  g.x_ld = d.x_ld; g.y_ld = d.y_ld;
  g.M = (long long)d.batch * d.h * d.w;
  const float* w = reinterpret_cast<const float*>(d.w_packed);  // fp32 view W[cout][tap][cin]
  if (d.cout <= 4 && d.cin % 64 == 0) {
    g.wide = d.cin; g.narrow = d.cout;
    CS_NARROW_SWITCH(d.cout, (conv_wide2narrow<NW><<<cs_blocks(g.M), CS_WARPS * 32, 0, st>>>(d.x, w, d.bias, d.y, g)));
  } else {
    g.wide = d.cout; g.narrow = d.cin;
    CS_NARROW_SWITCH(d.cin, (conv_narrow2wide<NW><<<cs_blocks(g.M), CS_WARPS * 32, 0, st>>>(d.x, w, d.bias, d.y, g, d.act, d.act_slope, d.prelu, d.y_pre)));
  }
  NSR_CHECK_LAUNCH("conv_small_fprop");
  return NSR_OK;
}

int conv_bias_grad(const NsrWgrad& d, float* bias_partial, int bias_blocks, cudaStream_t st);

bool conv_small_wgrad_supported(const NsrWgrad& d) {
  if (d.x == nullptr || d.dy == nullptr) return false;
  if (d.kh * d.kw > CS_MAXTAPS) return false;
  if (d.cout <= 4 && d.cout >= 1 && d.cin % 64 == 0 && d.x_ld % 2 == 0) return true;
  if (d.cin <= 4 && d.cin >= 1 && d.cout >= 16) return true;
  return false;
}
size_t conv_small_wgrad_workspace(const NsrWgrad& d) {
  const int narrow = d.cout <= 4 ? d.cout : d.cin;
  return (size_t)CS_WGRAD_BLOCKS * narrow * CS_MAXTAPS * 64 * sizeof(float) + (size_t)kNumSMs * 4 * d.cout * sizeof(float);
}
int conv_small_wgrad(const NsrWgrad& d, cudaStream_t st) {
  const size_t need = conv_small_wgrad_workspace(d);
  if (!d.workspace || d.workspace_bytes < need) {
    set_error("nsr_conv_wgrad(small): workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  SmallGeom g;
  g.B = d.batch; g.H = d.h; g.W = d.w; g.kh = d.kh; g.kw = d.kw; g.pad = d.pad;
  g.x_ld = d.x_ld; g.y_ld = d.dy_ld;
  g.M = (long long)d.batch * d.h * d.w;
  float* partial = reinterpret_cast<float*>(d.workspace);
  const int taps = d.kh * d.kw;
  int blocks = (int)((g.M + CS_WG_WARPS - 1) / CS_WG_WARPS);
  if (blocks > CS_WGRAD_BLOCKS) blocks = CS_WGRAD_BLOCKS;
  if (d.cout <= 4 && d.cin % 64 == 0) {
    g.wide = d.cin; g.narrow = d.cout;
    for (int grp = 0; grp < d.cin / 64; ++grp) {
      CS_NARROW_SWITCH(d.cout, (wgrad_wide2narrow<NW><<<blocks, CS_WG_WARPS * 32, 0, st>>>(d.x, d.dy, partial, g, grp)));
      NSR_CHECK_LAUNCH("wgrad_wide2narrow");
      wgrad_w2n_final<<<ceil_div(d.cout * taps * 64, 128), 128, 0, st>>>(partial, d.dw, blocks, d.cout, taps, d.cin, grp);
      NSR_CHECK_LAUNCH("wgrad_w2n_final");
    }
  } else {
    g.wide = d.cout; g.narrow = d.cin;
    for (int grp = 0; grp < (d.cout + 63) / 64; ++grp) {
      CS_NARROW_SWITCH(d.cin, (wgrad_narrow2wide<NW><<<blocks, CS_WG_WARPS * 32, 0, st>>>(d.x, d.dy, partial, g, grp)));
      NSR_CHECK_LAUNCH("wgrad_narrow2wide");
      wgrad_n2w_final<<<ceil_div(taps * d.cin * 64, 128), 128, 0, st>>>(partial, d.dw, blocks, d.cin, taps, d.cout, grp);
      NSR_CHECK_LAUNCH("wgrad_n2w_final");
    }
  }
  if (d.dbias) {
    float* bias_partial = partial + (size_t)CS_WGRAD_BLOCKS * (d.cout <= 4 ? d.cout : d.cin) * CS_MAXTAPS * 64;
    int bias_blocks = bias_grad_blocks(g.M);
    return conv_bias_grad(d, bias_partial, bias_blocks, st);
  }
  return NSR_OK;
}

}  // namespace nsr

// Direct 3x3 ("same") convolutions from <= 4 input channels at image resolution: VGG conv1_1 (3 -> 64), conv_first
// (3 -> C) and the dgrad of conv_last (3 -> C).  K of the implicit GEMM is 27, so these are streaming kernels bound by the
// wide OUTPUT's HBM traffic (4 B x pixels x cout), not tensor-core work: a CTA stages an (8+2)-row input patch in shared
// memory once, a warp owns a pixel (one broadcast 16-byte patch read per tap), every lane keeps the 9 x cin x 2 weights of
// its two output channels in registers for the whole tile, and the 32 lanes write the pixel's 64 channels as one 256-byte
// row.  0.34-0.36 ms per call at 2 M pixels against 0.56-0.58 ms for the im2col + tcgen05 route of conv_narrow_gemm.cu
// (which writes and re-reads a 32-wide column matrix).  The opposite direction (64 -> 3: conv_last, VGG conv1_1 dgrad)
// stays on that route: a direct kernel with a shared-memory patch of the wide input measured 0.48 vs 0.44 ms.
#include "common.cuh"

namespace nsr {

constexpr int CD_THREADS = 256;
constexpr int CD_TH = 8;     // tile rows
constexpr int CD_NW_TW = 64;  // tile width, narrow -> wide

struct DirectGeom {
  int B, H, W, wide, x_ld, y_ld, tiles_w, tiles_h;
};

// ------------------------------------------------------------------ narrow (cin <= 4) -> wide
// y[p, co] = act(bias[co] + sum_{tap, ci} x[p @ tap, ci] * w[co][tap][ci]);  grid.y = 64-channel group of cout
template <int NARROW>
__global__ void __launch_bounds__(CD_THREADS, 2) conv_n2w_direct(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 float* __restrict__ y_pre, DirectGeom g, int act, float slope,
                                                                 const float* __restrict__ prelu) {
  __shared__ float4 patch[(CD_TH + 2) * (CD_NW_TW + 2)];
  const int q = threadIdx.x & 31, pl = threadIdx.x >> 5;  // a warp = one pixel, lane = output-channel pair
  int tile = blockIdx.x;
  const int tx = tile % g.tiles_w;
  tile /= g.tiles_w;
  const int ty = tile % g.tiles_h, b = tile / g.tiles_h;
  const int h0 = ty * CD_TH, w0 = tx * CD_NW_TW;
  const int c = blockIdx.y * 64 + q * 2;
  const bool cok = c < g.wide;  // wide % 2 == 0
  float wr[9][NARROW][2];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < NARROW; ++ci)
#pragma unroll
      for (int e = 0; e < 2; ++e) wr[t][ci][e] = cok ? __ldg(w + ((size_t)(c + e) * 9 + t) * NARROW + ci) : 0.f;
  float2 b2 = make_float2(0.f, 0.f), s2 = make_float2(slope, slope);
  if (cok && bias) b2 = *reinterpret_cast<const float2*>(bias + c);
  if (cok && act == NSR_ACT_PRELU) s2 = *reinterpret_cast<const float2*>(prelu + c);
  const float* xb = x + (size_t)b * g.H * g.W * g.x_ld;
  for (int i = threadIdx.x; i < (CD_TH + 2) * (CD_NW_TW + 2); i += CD_THREADS) {
    const int ry = i / (CD_NW_TW + 2), rx = i - ry * (CD_NW_TW + 2);
    const int ih = h0 + ry - 1, iw = w0 + rx - 1;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
      const float* p = xb + ((size_t)ih * g.W + iw) * g.x_ld;
#pragma unroll
      for (int ci = 0; ci < NARROW; ++ci) v[ci] = __ldg(p + ci);
    }
    patch[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  if (!cok) return;
#pragma unroll 4
  for (int j = 0; j < CD_TH * CD_NW_TW / 8; ++j) {
    const int pi = pl + 8 * j;
    const int ly = pi / CD_NW_TW, lx = pi % CD_NW_TW;
    const int oh = h0 + ly, ow = w0 + lx;
    if (oh >= g.H || ow >= g.W) continue;
    float a0 = b2.x, a1 = b2.y;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 v4 = patch[(ly + t / 3) * (CD_NW_TW + 2) + lx + t % 3];  // one address per warp: broadcast
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int ci = 0; ci < NARROW; ++ci) {
        a0 = fmaf(v[ci], wr[t][ci][0], a0);
        a1 = fmaf(v[ci], wr[t][ci][1], a1);
      }
    }
    const size_t o = ((size_t)(b * g.H + oh) * g.W + ow) * g.y_ld + c;
    if (y_pre) *reinterpret_cast<float2*>(y_pre + o) = make_float2(a0, a1);
    if (act) { a0 = apply_act(a0, act, s2.x); a1 = apply_act(a1, act, s2.y); }
    *reinterpret_cast<float2*>(y + o) = make_float2(a0, a1);
  }
}

// ------------------------------------------------------------------ host
#define CD_NARROW_SWITCH(n, LAUNCH)                    \
  switch (n) {                                         \
    case 1: { constexpr int NW = 1; LAUNCH; } break;   \
    case 2: { constexpr int NW = 2; LAUNCH; } break;   \
    case 3: { constexpr int NW = 3; LAUNCH; } break;   \
    default: { constexpr int NW = 4; LAUNCH; } break;  \
  }

static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool conv_direct_fprop_supported(const NsrConv& d) {
  if (d.x == nullptr || d.y == nullptr || d.y_sti || d.x_sti) return false;
  if (d.kh != 3 || d.kw != 3 || d.pad != 1) return false;
  if (d.actgrad || d.residual || d.row_scale || d.act == NSR_ACT_GELU || d.pre_mode) return false;
  if (d.cin >= 1 && d.cin <= 4 && d.cout >= 16 && d.cout % 2 == 0 && d.y_ld % 2 == 0 && aligned16p(d.y) && aligned16p(d.y_pre) &&
      aligned16p(d.bias) && aligned16p(d.prelu))
    return true;
  return false;
}

int conv_direct_fprop(const NsrConv& d, cudaStream_t st) {
  DirectGeom g;
  g.B = d.batch; g.H = d.h; g.W = d.w; g.x_ld = d.x_ld; g.y_ld = d.y_ld;
  g.tiles_h = ceil_div(d.h, CD_TH);
  const float* w = reinterpret_cast<const float*>(d.w_packed);  // fp32 view W[cout][tap][cin]
  g.wide = d.cout;
  g.tiles_w = ceil_div(d.w, CD_NW_TW);
  dim3 grid((unsigned)(g.tiles_w * g.tiles_h * d.batch), (unsigned)ceil_div(d.cout, 64));
  CD_NARROW_SWITCH(d.cin, (conv_n2w_direct<NW><<<grid, CD_THREADS, 0, st>>>(d.x, w, d.bias, d.y, d.y_pre, g, d.act, d.act_slope, d.prelu)));
  NSR_CHECK_LAUNCH("conv_direct_fprop");
  return NSR_OK;
}

}  // namespace nsr

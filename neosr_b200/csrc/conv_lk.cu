// Large-kernel partial convolution of RealPLKSR (neosr/archs/realplksr_arch.py:26-41): a dense k x k (k = 17 or 13)
// conv over a 16-channel slab of a 64-channel NHWC tensor.  As an implicit GEMM it is N = 16 wide with K = 16*289:
// far too narrow for a 128 x N tcgen05 tile (the generic engine ran it at 12 TFLOP/s fprop, 2 TFLOP/s wgrad, 66 % of
// the C5 step), so it gets exact-fp32 CUDA-core kernels built around shared-memory reuse of the halo tile:
//   fprop / dgrad: one 16x16 output tile per CTA, the (16+k-1)^2 input halo staged once in shared memory
//                  (channel-chunk-major: conflict-free 16-byte reads), weights streamed one filter row at a time;
//                  each thread owns 2 pixels x 16 output channels (32 independent FMA chains, one broadcast weight
//                  read per 8 FMAs).  dgrad = the same kernel on the 180-degree-rotated, transposed filter.
//   wgrad:         one 16x16 pixel tile per CTA; thread (s, co-block, ci-block) accumulates a 4x4 block of
//                  dW[co][ci][r][s] over the tile's 256 pixels for each filter row r; per-tile partials are reduced in
//                  a fixed order (deterministic).
#include "common.cuh"

namespace nsr {
constexpr int LK_C = 16;       // channels in and out
constexpr int LK_T = 16;       // tile edge (pixels)
constexpr int LK_MAXK = 17;
constexpr int LK_HALO = LK_T + LK_MAXK - 1;  // 32

// ---------------------------------------------------------------- fprop / dgrad
// w: [n = 16][tap = k*k][c = 16] fp32 (the fp32 region of nsr_pack_weight, either flavour)
__global__ void __launch_bounds__(128) lk16_fprop_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int y_ld,
                                                         int H, int W, int k, int tiles_x, int tiles_y) {
  extern __shared__ float4 smem4[];
  const int R = k / 2, halo = LK_T + k - 1, npix = halo * halo;
  float4* xs = smem4;                       // [4 chunks][npix]
  float4* ws = smem4 + 4 * npix;            // [k taps of one row][16 n][4 chunks]
  const int tile = blockIdx.x, b = tile / (tiles_x * tiles_y), trem = tile - b * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * LK_T, tx0 = (trem % tiles_x) * LK_T;
  const float* xb = x + (size_t)b * H * W * x_ld;
  for (int i = threadIdx.x; i < 4 * npix; i += 128) {
    const int pix = i >> 2, ch = i & 3, hy = pix / halo, hx = pix - hy * halo;
    const int gy = ty0 + hy - R, gx = tx0 + hx - R;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = *reinterpret_cast<const float4*>(xb + ((size_t)gy * W + gx) * x_ld + ch * 4);
    xs[ch * npix + pix] = v;
  }
  // thread -> pixels (py, px) and (py + 8, px)
  const int px = threadIdx.x & 15, py = threadIdx.x >> 4;
  float acc0[LK_C], acc1[LK_C];
#pragma unroll
  for (int n = 0; n < LK_C; ++n) { acc0[n] = 0.f; acc1[n] = 0.f; }
  for (int r = 0; r < k; ++r) {
    __syncthreads();  // previous row's weights consumed (and, for r == 0, the halo tile written)
    for (int i = threadIdx.x; i < k * 64; i += 128) {
      const int s = i >> 6, n = (i >> 2) & 15, ch = i & 3;
      ws[i] = *reinterpret_cast<const float4*>(w + ((size_t)n * k * k + r * k + s) * LK_C + ch * 4);
    }
    __syncthreads();
    for (int s = 0; s < k; ++s) {
      const int p0 = (py + r) * halo + px + s, p1 = p0 + 8 * halo;
      float4 a0[4], a1[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) { a0[ch] = xs[ch * npix + p0]; a1[ch] = xs[ch * npix + p1]; }
      const float4* wt = ws + s * 64;
#pragma unroll
      for (int n = 0; n < LK_C; ++n) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 wv = wt[n * 4 + ch];
          acc0[n] = fmaf(a0[ch].x, wv.x, acc0[n]); acc0[n] = fmaf(a0[ch].y, wv.y, acc0[n]);
          acc0[n] = fmaf(a0[ch].z, wv.z, acc0[n]); acc0[n] = fmaf(a0[ch].w, wv.w, acc0[n]);
          acc1[n] = fmaf(a1[ch].x, wv.x, acc1[n]); acc1[n] = fmaf(a1[ch].y, wv.y, acc1[n]);
          acc1[n] = fmaf(a1[ch].z, wv.z, acc1[n]); acc1[n] = fmaf(a1[ch].w, wv.w, acc1[n]);
        }
      }
    }
  }
  float bv[LK_C];
#pragma unroll
  for (int n = 0; n < LK_C; ++n) bv[n] = bias ? bias[n] : 0.f;
  const int gx = tx0 + px;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int gy = ty0 + py + 8 * half;
    if (gy < H && gx < W) {
      float* o = y + ((size_t)b * H * W + (size_t)gy * W + gx) * y_ld;
      const float* a = half ? acc1 : acc0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(o + 4 * q) = make_float4(a[4 * q] + bv[4 * q], a[4 * q + 1] + bv[4 * q + 1],
                                                            a[4 * q + 2] + bv[4 * q + 2], a[4 * q + 3] + bv[4 * q + 3]);
    }
  }
}

// ---------------------------------------------------------------- wgrad
constexpr int LKW_THREADS = 288;  // 17 filter columns x 16 (co-block, ci-block) roles, 9 warps (the last half warp idles)
__global__ void __launch_bounds__(LKW_THREADS) lk16_wgrad_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ dy,
                                                                 int dy_ld, float* __restrict__ partial,
                                                                 float* __restrict__ partial_bias, int H, int W, int k,
                                                                 int tiles_x, int tiles_y) {
  extern __shared__ float4 smem4[];
  const int R = k / 2, halo = LK_T + k - 1, npix = halo * halo;
  float4* xs = smem4;                 // [npix][4 chunks]  (pixel-major: a warp's 8 distinct chunks are contiguous)
  float4* ds = smem4 + 4 * npix;      // [256 pixels][4 chunks]
  const int tile = blockIdx.x, b = tile / (tiles_x * tiles_y), trem = tile - b * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * LK_T, tx0 = (trem % tiles_x) * LK_T;
  const float* xb = x + (size_t)b * H * W * x_ld;
  const float* db = dy + (size_t)b * H * W * dy_ld;
  for (int i = threadIdx.x; i < 4 * npix; i += LKW_THREADS) {
    const int pix = i >> 2, ch = i & 3, hy = pix / halo, hx = pix - hy * halo;
    const int gy = ty0 + hy - R, gx = tx0 + hx - R;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = *reinterpret_cast<const float4*>(xb + ((size_t)gy * W + gx) * x_ld + ch * 4);
    xs[i] = v;
  }
  for (int i = threadIdx.x; i < 4 * LK_T * LK_T; i += LKW_THREADS) {
    const int pix = i >> 2, ch = i & 3, gy = ty0 + pix / LK_T, gx = tx0 + pix % LK_T;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy < H && gx < W) v = *reinterpret_cast<const float4*>(db + ((size_t)gy * W + gx) * dy_ld + ch * 4);
    ds[i] = v;
  }
  __syncthreads();
  if (partial_bias && threadIdx.x < LK_C) {  // dbias partial: column sums of the dy tile
    const float* dsf = reinterpret_cast<const float*>(ds);
    float s = 0.f;
    for (int p = 0; p < LK_T * LK_T; ++p) s += dsf[p * LK_C + threadIdx.x];
    partial_bias[(size_t)tile * LK_C + threadIdx.x] = s;
  }
  const int s = threadIdx.x >> 4, role = threadIdx.x & 15, cob = role >> 2, cib = role & 3;
  if (s >= k) return;
  const size_t taps = (size_t)k * k;
  float* pt = partial + (size_t)tile * LK_C * LK_C * taps;
  for (int r = 0; r < k; ++r) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int py = 0; py < LK_T; ++py) {
      const float4* xr = xs + ((py + r) * halo + s) * 4 + cib;
      const float4* dr = ds + (py * LK_T) * 4 + cob;
#pragma unroll
      for (int px = 0; px < LK_T; ++px) {
        const float4 xv = xr[px * 4], dv = dr[px * 4];
        acc[0][0] = fmaf(dv.x, xv.x, acc[0][0]); acc[0][1] = fmaf(dv.x, xv.y, acc[0][1]);
        acc[0][2] = fmaf(dv.x, xv.z, acc[0][2]); acc[0][3] = fmaf(dv.x, xv.w, acc[0][3]);
        acc[1][0] = fmaf(dv.y, xv.x, acc[1][0]); acc[1][1] = fmaf(dv.y, xv.y, acc[1][1]);
        acc[1][2] = fmaf(dv.y, xv.z, acc[1][2]); acc[1][3] = fmaf(dv.y, xv.w, acc[1][3]);
        acc[2][0] = fmaf(dv.z, xv.x, acc[2][0]); acc[2][1] = fmaf(dv.z, xv.y, acc[2][1]);
        acc[2][2] = fmaf(dv.z, xv.z, acc[2][2]); acc[2][3] = fmaf(dv.z, xv.w, acc[2][3]);
        acc[3][0] = fmaf(dv.w, xv.x, acc[3][0]); acc[3][1] = fmaf(dv.w, xv.y, acc[3][1]);
        acc[3][2] = fmaf(dv.w, xv.z, acc[3][2]); acc[3][3] = fmaf(dv.w, xv.w, acc[3][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) pt[((size_t)(cob * 4 + i) * LK_C + cib * 4 + j) * taps + r * k + s] = acc[i][j];
  }
}
// dw[e] = sum over tiles (fixed order); dbias likewise
__global__ void lk16_wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias,
                                         float* __restrict__ dw, float* __restrict__ dbias, int ntiles, int nelem) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nelem) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int t = 0;
    for (; t + 4 <= ntiles; t += 4) {
      s0 += partial[(size_t)t * nelem + e];
      s1 += partial[(size_t)(t + 1) * nelem + e];
      s2 += partial[(size_t)(t + 2) * nelem + e];
      s3 += partial[(size_t)(t + 3) * nelem + e];
    }
    for (; t < ntiles; ++t) s0 += partial[(size_t)t * nelem + e];
    dw[e] = (s0 + s1) + (s2 + s3);
  } else if (dbias && e < nelem + LK_C) {
    const int c = e - nelem;
    float s = 0.f;
    for (int t = 0; t < ntiles; ++t) s += partial_bias[(size_t)t * LK_C + c];
    dbias[c] = s;
  }
}
static inline size_t lk_fprop_smem(int k) { return ((size_t)4 * (LK_T + k - 1) * (LK_T + k - 1) + (size_t)k * 64) * sizeof(float4); }
static inline size_t lk_wgrad_smem(int k) { return ((size_t)4 * (LK_T + k - 1) * (LK_T + k - 1) + 4 * LK_T * LK_T) * sizeof(float4); }
}  // namespace nsr
using namespace nsr;

static int lk_check(const char* who, int batch, int h, int w, int k, int ld_a, int ld_b, const void* a, const void* b) {
  NSR_CHECK_ARG(batch > 0 && h > 0 && w > 0, "%s: bad shape", who);
  NSR_CHECK_ARG(k % 2 == 1 && k >= 3 && k <= LK_MAXK, "%s: kernel size %d (odd, 3..%d)", who, k, LK_MAXK);
  NSR_CHECK_ARG(ld_a >= LK_C && ld_b >= LK_C && ld_a % 4 == 0 && ld_b % 4 == 0, "%s: leading dims must be >= 16 and multiples of 4", who);
  NSR_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "%s: 16-byte aligned buffers required", who);
  return NSR_OK;
}

extern "C" int nsr_conv_lk16_fprop(const float* x, int x_ld, const float* w_ntc, const float* bias, float* y, int y_ld, int batch,
                                   int h, int w, int k, void* stream) {
  NSR_CHECK_ARG(x && w_ntc && y && x != y, "nsr_conv_lk16_fprop: null or aliased buffers");
  int rc = lk_check("nsr_conv_lk16_fprop", batch, h, w, k, x_ld, y_ld, x, y);
  if (rc) return rc;
  const size_t smem = lk_fprop_smem(k);
  cudaError_t e = cudaFuncSetAttribute(lk16_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_fprop_smem(LK_MAXK));
  if (e != cudaSuccess) { set_error("nsr_conv_lk16_fprop: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NSR_E_CUDA; }
  const int tx = ceil_div(w, LK_T), ty = ceil_div(h, LK_T);
  lk16_fprop_kernel<<<batch * tx * ty, 128, smem, (cudaStream_t)stream>>>(x, x_ld, w_ntc, bias, y, y_ld, h, w, k, tx, ty);
  NSR_CHECK_LAUNCH("nsr_conv_lk16_fprop");
  return NSR_OK;
}

extern "C" size_t nsr_conv_lk16_wgrad_workspace(int batch, int h, int w, int k) {
  const size_t tiles = (size_t)batch * ceil_div(w, LK_T) * ceil_div(h, LK_T);
  return tiles * ((size_t)LK_C * LK_C * k * k + LK_C) * sizeof(float);
}
extern "C" int nsr_conv_lk16_wgrad(const float* x, int x_ld, const float* dy, int dy_ld, float* dw, float* dbias, int batch, int h,
                                   int w, int k, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(x && dy && dw, "nsr_conv_lk16_wgrad: null pointer");
  int rc = lk_check("nsr_conv_lk16_wgrad", batch, h, w, k, x_ld, dy_ld, x, dy);
  if (rc) return rc;
  if (!workspace || workspace_bytes < nsr_conv_lk16_wgrad_workspace(batch, h, w, k)) {
    set_error("nsr_conv_lk16_wgrad: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaError_t e = cudaFuncSetAttribute(lk16_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_wgrad_smem(LK_MAXK));
  if (e != cudaSuccess) { set_error("nsr_conv_lk16_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NSR_E_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  const int tx = ceil_div(w, LK_T), ty = ceil_div(h, LK_T), tiles = batch * tx * ty, nelem = LK_C * LK_C * k * k;
  float* partial = (float*)workspace;
  float* pbias = partial + (size_t)tiles * nelem;
  lk16_wgrad_kernel<<<tiles, LKW_THREADS, lk_wgrad_smem(k), st>>>(x, x_ld, dy, dy_ld, partial, dbias ? pbias : nullptr, h, w, k, tx, ty);
  NSR_CHECK_LAUNCH("nsr_conv_lk16_wgrad");
  lk16_wgrad_reduce_kernel<<<ceil_div(nelem + LK_C, 256), 256, 0, st>>>(partial, pbias, dw, dbias, tiles, nelem);
  NSR_CHECK_LAUNCH("nsr_conv_lk16_wgrad(reduce)");
  return NSR_OK;
}

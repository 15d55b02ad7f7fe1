// Pieces of the U-Net discriminator (neosr/archs/unet_arch.py) that are not contractions:
//   * bilinear x2 upsample (align_corners=False) forward / backward on NHWC,
//   * the weight remap that turns a 4x4 stride-2 pad-1 convolution into a 3x3 stride-1 convolution
//     over the pixel-unshuffled (space-to-depth, r=2) input, so the tcgen05 implicit-GEMM kernels
//     (stride 1, "same" padding) serve it unchanged,
//   * spectral normalisation (torch.nn.utils.spectral_norm: one power iteration per training-mode
//     forward, W / sigma) and its backward.
#include "common.cuh"

namespace nsr {

static inline int uo_blocks(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// ---- bilinear x2, align_corners=False: out[2k] = .25 in[k-1] + .75 in[k], out[2k+1] = .75 in[k] + .25 in[k+1],
//      source indices clamped to [0, n-1] (ATen upsample_bilinear2d).
__device__ __forceinline__ void bil_src(int o, int n, int& i0, int& i1, float& w1) {
  float src = (o + 0.5f) * 0.5f - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  w1 = src - (float)i0;
}
__global__ void bilinear_up2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const size_t total = (size_t)B * 2 * H * 2 * W * C;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(o % C);
    size_t q = o / C;
    const int ow = (int)(q % (2 * W));
    q /= 2 * W;
    const int oh = (int)(q % (2 * H));
    const size_t b = q / (2 * H);
    int h0, h1, w0, w1;
    float lh, lw;
    bil_src(oh, H, h0, h1, lh);
    bil_src(ow, W, w0, w1, lw);
    const float* p = x + b * H * W * C + c;
    const float v00 = p[((size_t)h0 * W + w0) * C], v01 = p[((size_t)h0 * W + w1) * C];
    const float v10 = p[((size_t)h1 * W + w0) * C], v11 = p[((size_t)h1 * W + w1) * C];
    y[o] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
  }
}
// backward as a gather: input row k receives from output rows 2k-1 .. 2k+2 (clamped), same for columns
__device__ __forceinline__ float bil_wt(int o, int n, int k) {  // weight of input k in output o (1-D)
  if (o < 0 || o >= 2 * n) return 0.f;
  int i0, i1;
  float w1;
  bil_src(o, n, i0, i1, w1);
  float w = 0.f;
  if (i0 == k) w += 1.f - w1;
  if (i1 == k) w += w1;
  return w;
}
__global__ void bilinear_up2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const size_t total = (size_t)B * H * W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t q = i / C;
    const int w = (int)(q % W);
    q /= W;
    const int h = (int)(q % H);
    const size_t b = q / H;
    float s = 0.f;
#pragma unroll
    for (int a = -1; a <= 2; ++a) {
      const int oh = 2 * h + a;
      const float wh = bil_wt(oh, H, h);
      if (wh == 0.f) continue;
#pragma unroll
      for (int e = -1; e <= 2; ++e) {
        const int ow = 2 * w + e;
        const float ww = bil_wt(ow, W, w);
        if (ww == 0.f) continue;
        s = fmaf(wh * ww, dy[((b * 2 * H + oh) * 2 * W + ow) * C + c], s);
      }
    }
    dx[i] = s;
  }
}

// ---- 4x4 stride-2 pad-1 conv == 3x3 stride-1 pad-1 conv on pixel_unshuffle(x, 2):
//   w3[co][(c, i, j)][a][b] = w4[co][c][r(a,i)][s(b,j)],  r(0,1)=0, r(1,0)=1, r(1,1)=2, r(2,0)=3, else zero tap.
// inverse != 0 gathers a 3x3-layout gradient back into the 4x4 layout.
__global__ void conv4x4s2_remap_kernel(const float* __restrict__ src, float* __restrict__ dst, int cout, int cin, int inverse) {
  const size_t total3 = (size_t)cout * cin * 4 * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total3; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % 3), a = (int)((i / 3) % 3);
    const size_t ch = i / 9;                       // co * 4cin + (c*4 + ii*2 + jj)
    const int sub = (int)(ch % 4), jj = sub & 1, ii = sub >> 1;
    const size_t cc = ch / 4;                      // co * cin + c
    const int r = (a == 0 && ii == 1) ? 0 : (a == 1 ? 1 + ii : ((a == 2 && ii == 0) ? 3 : -1));
    const int s = (b == 0 && jj == 1) ? 0 : (b == 1 ? 1 + jj : ((b == 2 && jj == 0) ? 3 : -1));
    if (inverse) {
      if (r >= 0 && s >= 0) dst[cc * 16 + r * 4 + s] = src[i];
    } else {
      dst[i] = (r >= 0 && s >= 0) ? src[cc * 16 + r * 4 + s] : 0.f;
    }
  }
}

// ---- spectral norm.  W is [rows = cout, cols = cin*kh*kw] row-major.
__global__ void sn_wt_u(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ t, int rows, int cols) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;  // t[k] = sum_r W[r][k] u[r]
  if (k >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s = fmaf(w[(size_t)r * cols + k], u[r], s);
  t[k] = s;
}
__global__ void sn_w_v(const float* __restrict__ w, const float* __restrict__ v, float* __restrict__ s_out, int rows, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;  // s[r] = sum_k W[r][k] v[k]
  if (warp >= rows) return;
  float s = 0.f;
  for (int k = lane; k < cols; k += 32) s = fmaf(w[(size_t)warp * cols + k], v[k], s);
  s = warp_sum(s);
  if (lane == 0) s_out[warp] = s;
}
// single block: out = in / max(||in||, eps); *norm_out = ||in||  (fixed-order reduction)
__global__ void sn_normalize(const float* __restrict__ in, float* __restrict__ out, int n, float eps, float* norm_out) {
  __shared__ float sm[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s = fmaf(in[i], in[i], s);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  const float nrm = sqrtf(sm[0]);
  const float inv = 1.f / fmaxf(nrm, eps);
  for (int i = threadIdx.x; i < n; i += 256) out[i] = in[i] * inv;
  if (threadIdx.x == 0 && norm_out) *norm_out = nrm;
}
__global__ void sn_dot(const float* __restrict__ a, const float* __restrict__ b, int n, float* out) {
  __shared__ float sm[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s = fmaf(a[i], b[i], s);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}
__global__ void sn_scale(const float* __restrict__ w, const float* __restrict__ sigma, float* __restrict__ out, size_t n) {
  const float inv = 1.f / *sigma;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = w[i] * inv;
}
// <G, Wsn> partial sums, then dW = (G - <G, Wsn> u v^T) / sigma
__global__ void sn_inner_partial(const float* __restrict__ g, const float* __restrict__ wsn, size_t n, float* __restrict__ partial) {
  __shared__ float sm[256];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s = fmaf(g[i], wsn[i], s);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void sn_bwd_apply(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                             const float* __restrict__ sigma, const float* __restrict__ partial, int nparts,
                             float* __restrict__ dw, int rows, int cols, int accumulate) {
  __shared__ float inner_s;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < nparts; ++i) s += partial[i];
    inner_s = s;
  }
  __syncthreads();
  const float inner = inner_s, inv = 1.f / *sigma;
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), k = (int)(i - (size_t)r * cols);
    const float val = (g[i] - inner * u[r] * v[k]) * inv;
    dw[i] = accumulate ? dw[i] + val : val;
  }
}

}  // namespace nsr
using namespace nsr;

extern "C" int nsr_bilinear_up2_nhwc(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  NSR_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0, "nsr_bilinear_up2_nhwc: bad arguments");
  bilinear_up2_kernel<<<uo_blocks((size_t)B * H * W * C * 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, B, H, W, C);
  NSR_CHECK_LAUNCH("bilinear_up2");
  return NSR_OK;
}
extern "C" int nsr_bilinear_up2_bwd_nhwc(const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
  NSR_CHECK_ARG(dy && dx && B > 0 && H > 0 && W > 0 && C > 0, "nsr_bilinear_up2_bwd_nhwc: bad arguments");
  bilinear_up2_bwd_kernel<<<uo_blocks((size_t)B * H * W * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, dx, B, H, W, C);
  NSR_CHECK_LAUNCH("bilinear_up2_bwd");
  return NSR_OK;
}
extern "C" int nsr_conv4x4s2_remap(const float* src, float* dst, int cout, int cin, int inverse, void* stream) {
  NSR_CHECK_ARG(src && dst && cout > 0 && cin > 0, "nsr_conv4x4s2_remap: bad arguments");
  conv4x4s2_remap_kernel<<<uo_blocks((size_t)cout * cin * 36), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, cout, cin, inverse);
  NSR_CHECK_LAUNCH("conv4x4s2_remap");
  return NSR_OK;
}
extern "C" size_t nsr_spectral_norm_workspace(int rows, int cols) { return (size_t)(rows + cols + 64 + kNumSMs * 2) * sizeof(float); }
extern "C" int nsr_spectral_norm_fwd(const float* w_orig, float* u, float* v, float* w_out, float* sigma, int rows, int cols,
                                     int power_iterations, float eps, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(w_orig && u && v && w_out && sigma && rows > 0 && cols > 0 && power_iterations >= 0,
                "nsr_spectral_norm_fwd: bad arguments");
  if (!workspace || workspace_bytes < nsr_spectral_norm_workspace(rows, cols)) {
    set_error("nsr_spectral_norm_fwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* t = reinterpret_cast<float*>(workspace);  // [cols]
  float* s = t + cols;                             // [rows]
  for (int it = 0; it < power_iterations; ++it) {  // v = normalize(W^T u); u = normalize(W v)
    sn_wt_u<<<ceil_div(cols, 128), 128, 0, st>>>(w_orig, u, t, rows, cols);
    sn_normalize<<<1, 256, 0, st>>>(t, v, cols, eps, nullptr);
    sn_w_v<<<ceil_div(rows * 32, 128), 128, 0, st>>>(w_orig, v, s, rows, cols);
    sn_normalize<<<1, 256, 0, st>>>(s, u, rows, eps, nullptr);
  }
  sn_w_v<<<ceil_div(rows * 32, 128), 128, 0, st>>>(w_orig, v, s, rows, cols);  // sigma = u . (W v)
  sn_dot<<<1, 256, 0, st>>>(u, s, rows, sigma);
  sn_scale<<<uo_blocks((size_t)rows * cols), 256, 0, st>>>(w_orig, sigma, w_out, (size_t)rows * cols);
  NSR_CHECK_LAUNCH("spectral_norm_fwd");
  return NSR_OK;
}
extern "C" int nsr_spectral_norm_bwd(const float* g_wsn, const float* w_sn, const float* u, const float* v, const float* sigma,
                                     float* dw_orig, int rows, int cols, int accumulate, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(g_wsn && w_sn && u && v && sigma && dw_orig && rows > 0 && cols > 0, "nsr_spectral_norm_bwd: bad arguments");
  if (!workspace || workspace_bytes < nsr_spectral_norm_workspace(rows, cols)) {
    set_error("nsr_spectral_norm_bwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  const size_t n = (size_t)rows * cols;
  int parts = uo_blocks(n);
  if (parts > kNumSMs * 2) parts = kNumSMs * 2;
  sn_inner_partial<<<parts, 256, 0, st>>>(g_wsn, w_sn, n, partial);
  sn_bwd_apply<<<uo_blocks(n), 256, 0, st>>>(g_wsn, u, v, sigma, partial, parts, dw_orig, rows, cols, accumulate);
  NSR_CHECK_LAUNCH("spectral_norm_bwd");
  return NSR_OK;
}

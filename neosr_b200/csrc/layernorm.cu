// LayerNorm over the channel dim of [rows, C] tokens: warp-per-row, shuffle reductions,
// deterministic two-pass dgamma/dbeta.  HBM-bound (fwd: read x, write y; bwd: read x,dy(,dres), write dx).
#include "common.cuh"

namespace nsr {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_BLOCKS = kNumSMs * 4;

__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta,
                                                                     float* __restrict__ y, float* __restrict__ mean,
                                                                     float* __restrict__ rstd, int rows, int C,
                                                                     float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invC = 1.f / (float)C;
  for (long long r = (long long)blockIdx.x * LN_WARPS + warp; r < rows; r += (long long)gridDim.x * LN_WARPS) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mu = warp_sum(s) * invC;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float dlt = xr[c] - mu;
      v = fmaf(dlt, dlt, v);
    }
    const float rs = rsqrtf(warp_sum(v) * invC + eps);
    float* yr = y + r * C;
    for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mu) * rs * gamma[c] + beta[c];
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
  }
}

__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
    float* __restrict__ dx, float* __restrict__ partial, int rows, int C) {
  extern __shared__ float sm[];  // [LN_WARPS][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* dg = sm + (size_t)warp * 2 * C;
  float* db = dg + C;
  for (int c = lane; c < C; c += 32) { dg[c] = 0.f; db[c] = 0.f; }
  const float invC = 1.f / (float)C;
  for (long long r = (long long)blockIdx.x * LN_WARPS + warp; r < rows; r += (long long)gridDim.x * LN_WARPS) {
    const float* xr = x + r * C;
    const float* gr = dy + r * C;
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs;
      const float g = gr[c];
      const float gg = g * gamma[c];
      s1 += gg;
      s2 = fmaf(gg, xh, s2);
      dg[c] = fmaf(g, xh, dg[c]);
      db[c] += g;
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    float* dxr = dx + r * C;
    const float* rr = dres ? dres + r * C : nullptr;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs;
      float v = rs * (gr[c] * gamma[c] - s1 - xh * s2);
      if (rr) v += rr[c];
      dxr[c] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += sm[(size_t)w * 2 * C + c];
    partial[(size_t)blockIdx.x * 2 * C + c] = s;
  }
}

__global__ void layernorm_bwd_final(const float* __restrict__ partial, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int blocks, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * C) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[(size_t)b * 2 * C + c];
  if (c < C) { if (dgamma) dgamma[c] = s; }
  else if (dbeta) dbeta[c - C] = s;
}

static int ln_blocks(int rows) {
  int b = ceil_div(rows, LN_WARPS);
  return b > LN_MAX_BLOCKS ? LN_MAX_BLOCKS : (b < 1 ? 1 : b);
}
}  // namespace nsr
using namespace nsr;

extern "C" int nsr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                 float* rstd, int rows, int c, float eps, void* stream) {
  NSR_CHECK_ARG(x && gamma && beta && y && rows > 0 && c > 0, "nsr_layernorm_fwd: bad arguments");
  layernorm_fwd_kernel<<<ln_blocks(rows), LN_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, y, mean, rstd, rows, c, eps);
  NSR_CHECK_LAUNCH("layernorm_fwd");
  return NSR_OK;
}
extern "C" size_t nsr_layernorm_bwd_workspace(int c) { return (size_t)LN_MAX_BLOCKS * 2 * c * sizeof(float); }
extern "C" int nsr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta, int rows,
                                 int c, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(dy && x && gamma && mean && rstd && dx && rows > 0 && c > 0, "nsr_layernorm_bwd: bad arguments");
  NSR_CHECK_ARG(c <= 704, "nsr_layernorm_bwd: C > 704 not supported");
  if (!workspace || workspace_bytes < nsr_layernorm_bwd_workspace(c)) {
    set_error("nsr_layernorm_bwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = ln_blocks(rows);
  const size_t smem = (size_t)LN_WARPS * 2 * c * sizeof(float);
  float* partial = reinterpret_cast<float*>(workspace);
  layernorm_bwd_kernel<<<blocks, LN_WARPS * 32, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx, partial, rows, c);
  NSR_CHECK_LAUNCH("layernorm_bwd");
  layernorm_bwd_final<<<ceil_div(2 * c, 128), 128, 0, st>>>(partial, dgamma, dbeta, blocks, c);
  NSR_CHECK_LAUNCH("layernorm_bwd_final");
  return NSR_OK;
}

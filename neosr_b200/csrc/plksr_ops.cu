// Non-contraction pieces of RealPLKSR (neosr/archs/realplksr_arch.py): Mish (DCCM, :14-23), the
// element-wise attention gate x * sigmoid(f(x)) (EA, :44-53), GroupNorm over NHWC (PLKBlock.norm, :85-88)
// fused with the block's skip add (:99), and `feats(x) + repeat_interleave(x, s^2)` (:158-160).
#include "common.cuh"

namespace nsr {

static inline int po_blocks(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// torch.nn.functional.mish: x * tanh(softplus(x)), softplus threshold 20
__device__ __forceinline__ float softplus20(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float mish_f(float x) { return x * tanhf(softplus20(x)); }
__device__ __forceinline__ float mish_grad_f(float x) {
  const float t = tanhf(softplus20(x));
  const float sg = 1.f / (1.f + expf(-x));  // d softplus / dx
  return t + x * (1.f - t * t) * sg;
}
__global__ void mish_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = mish_f(x[i]);
}
__global__ void mish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = dy[i] * mish_grad_f(x[i]);
}
// y = x * sigmoid(s);  backward: dx = dy * sigmoid(s), ds = dy * x * sigmoid(s) * (1 - sigmoid(s))
__global__ void mul_sigmoid_fwd_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = x[i] / (1.f + expf(-s[i]));
}
__global__ void mul_sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ s,
                                       float* __restrict__ dx, float* __restrict__ ds, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float sg = 1.f / (1.f + expf(-s[i]));
    const float g = dy[i];
    dx[i] = g * sg;
    ds[i] = g * x[i] * sg * (1.f - sg);
  }
}
// y[p, c*r + k] += x[p, c]   (torch.repeat_interleave(x, r, dim=1) in NHWC)
__global__ void add_repeat_interleave_kernel(float* __restrict__ y, const float* __restrict__ x, size_t rows, int c, int r) {
  const size_t n = rows * (size_t)c * r;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / ((size_t)c * r);
    const int j = (int)(i - p * (size_t)c * r);
    y[i] += x[p * c + j / r];
  }
}

// ---- GroupNorm, NHWC [B, HW, C], G groups of Cg = C / G channels -------------------------------
// stage 1: per (b, g, chunk) partial sums of (v1, v2); stage 2 reduces the chunks in fixed order.
constexpr int GN_THREADS = 256;
constexpr int GN_CHUNKS = 32;  // pixel chunks per (b, g): B*G*32 blocks
template <bool BWD>
__global__ void __launch_bounds__(GN_THREADS) gn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ gamma, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ partial,
                                                             float* __restrict__ chan_partial, int HW, int C, int G) {
  // grid: (GN_CHUNKS, B*G).  FWD: v1 = sum x, v2 = sum x^2.  BWD: v1 = sum dy*gamma, v2 = sum dy*gamma*xhat,
  // and per-channel partials of dgamma = sum dy*xhat, dbeta = sum dy into chan_partial[b][chunk][2][C].
  const int bg = blockIdx.y, b = bg / G, g = bg - b * G, Cg = C / G;
  const int chunk = blockIdx.x;
  const int per = (HW + GN_CHUNKS - 1) / GN_CHUNKS;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  const int cl = threadIdx.x % Cg, pl = threadIdx.x / Cg, pstep = GN_THREADS / Cg;
  const int c = g * Cg + cl;
  float mu = 0.f, rs = 0.f, gm = 0.f;
  if (BWD) { mu = mean[bg]; rs = rstd[bg]; gm = gamma[c]; }
  float v1 = 0.f, v2 = 0.f, dg = 0.f, db = 0.f;
  if (pl < pstep)
    for (int p = p0 + pl; p < p1; p += pstep) {
      const size_t o = ((size_t)b * HW + p) * C + c;
      const float xv = x[o];
      if (BWD) {
        const float d = dy[o], xh = (xv - mu) * rs;
        v1 = fmaf(d, gm, v1);
        v2 = fmaf(d * gm, xh, v2);
        dg = fmaf(d, xh, dg);
        db += d;
      } else {
        v1 += xv;
        v2 = fmaf(xv, xv, v2);
      }
    }
  __shared__ float s1[GN_THREADS], s2[GN_THREADS], s3[GN_THREADS], s4[GN_THREADS];
  s1[threadIdx.x] = v1; s2[threadIdx.x] = v2; s3[threadIdx.x] = dg; s4[threadIdx.x] = db;
  __syncthreads();
  if (BWD && threadIdx.x < Cg) {  // per-channel sums over this chunk's pixels (fixed order)
    float a = 0.f, bb = 0.f;
    for (int q = 0; q < pstep; ++q) { a += s3[q * Cg + threadIdx.x]; bb += s4[q * Cg + threadIdx.x]; }
    float* cp = chan_partial + ((size_t)(b * GN_CHUNKS + chunk) * 2) * C;
    cp[g * Cg + threadIdx.x] = a;
    cp[C + g * Cg + threadIdx.x] = bb;
  }
  for (int o = GN_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[((size_t)bg * GN_CHUNKS + chunk) * 2] = s1[0];
    partial[((size_t)bg * GN_CHUNKS + chunk) * 2 + 1] = s2[0];
  }
}
__global__ void gn_stats_final(const float* __restrict__ partial, float* __restrict__ mean, float* __restrict__ rstd, int BG,
                               float count, float eps) {
  const int bg = blockIdx.x * blockDim.x + threadIdx.x;
  if (bg >= BG) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < GN_CHUNKS; ++k) { a += partial[((size_t)bg * GN_CHUNKS + k) * 2]; b += partial[((size_t)bg * GN_CHUNKS + k) * 2 + 1]; }
  const double m = a / count;
  double var = b / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[bg] = (float)m;
  rstd[bg] = (float)(1.0 / sqrt(var + (double)eps));
}
// y = (x - mean) * rstd * gamma + beta (+ residual)
__global__ void gn_apply_fwd(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ residual,
                             float* __restrict__ y, int B, int HW, int C, int G) {
  const size_t n = (size_t)B * HW * C;
  const int Cg = C / G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((size_t)HW * C));
    const int bg = b * G + c / Cg;
    float v = (x[i] - mean[bg]) * rstd[bg] * gamma[c] + beta[c];
    if (residual) v += residual[i];
    y[i] = v;
  }
}
// dx = rstd * (dy*gamma - m1 - xhat*m2), m1/m2 = group means of dy*gamma and dy*gamma*xhat
__global__ void gn_apply_bwd(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                             const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ partial,
                             float* __restrict__ dx, int B, int HW, int C, int G) {
  extern __shared__ float m12[];  // [B*G][2]
  const float inv = 1.f / ((float)HW * (float)(C / G));
  for (int bg = threadIdx.x; bg < B * G; bg += blockDim.x) {
    float a = 0.f, b2 = 0.f;
    for (int k = 0; k < GN_CHUNKS; ++k) { a += partial[((size_t)bg * GN_CHUNKS + k) * 2]; b2 += partial[((size_t)bg * GN_CHUNKS + k) * 2 + 1]; }
    m12[bg * 2] = a * inv;
    m12[bg * 2 + 1] = b2 * inv;
  }
  __syncthreads();
  const size_t n = (size_t)B * HW * C;
  const int Cg = C / G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((size_t)HW * C));
    const int bg = b * G + c / Cg;
    const float rs = rstd[bg], xh = (x[i] - mean[bg]) * rs;
    dx[i] = rs * (dy[i] * gamma[c] - m12[bg * 2] - xh * m12[bg * 2 + 1]);
  }
}
__global__ void gn_chan_final(const float* __restrict__ chan_partial, float* __restrict__ dgamma, float* __restrict__ dbeta,
                              int n_parts, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < n_parts; ++k) { a += chan_partial[(size_t)k * 2 * C + c]; b += chan_partial[(size_t)k * 2 * C + C + c]; }
  dgamma[c] = a;
  dbeta[c] = b;
}

}  // namespace nsr
using namespace nsr;

#define ST reinterpret_cast<cudaStream_t>(stream)
extern "C" int nsr_mish_fwd(const float* x, float* y, size_t n, void* stream) {
  NSR_CHECK_ARG(x && y && n > 0, "nsr_mish_fwd: bad arguments");
  mish_fwd_kernel<<<po_blocks(n), 256, 0, ST>>>(x, y, n);
  NSR_CHECK_LAUNCH("mish_fwd");
  return NSR_OK;
}
extern "C" int nsr_mish_bwd(const float* dy, const float* x, float* dx, size_t n, void* stream) {
  NSR_CHECK_ARG(dy && x && dx && n > 0, "nsr_mish_bwd: bad arguments");
  mish_bwd_kernel<<<po_blocks(n), 256, 0, ST>>>(dy, x, dx, n);
  NSR_CHECK_LAUNCH("mish_bwd");
  return NSR_OK;
}
extern "C" int nsr_mul_sigmoid_fwd(const float* x, const float* s, float* y, size_t n, void* stream) {
  NSR_CHECK_ARG(x && s && y && n > 0, "nsr_mul_sigmoid_fwd: bad arguments");
  mul_sigmoid_fwd_kernel<<<po_blocks(n), 256, 0, ST>>>(x, s, y, n);
  NSR_CHECK_LAUNCH("mul_sigmoid_fwd");
  return NSR_OK;
}
extern "C" int nsr_mul_sigmoid_bwd(const float* dy, const float* x, const float* s, float* dx, float* ds, size_t n, void* stream) {
  NSR_CHECK_ARG(dy && x && s && dx && ds && n > 0, "nsr_mul_sigmoid_bwd: bad arguments");
  mul_sigmoid_bwd_kernel<<<po_blocks(n), 256, 0, ST>>>(dy, x, s, dx, ds, n);
  NSR_CHECK_LAUNCH("mul_sigmoid_bwd");
  return NSR_OK;
}
extern "C" int nsr_add_repeat_interleave(float* y, const float* x, size_t rows, int c, int r, void* stream) {
  NSR_CHECK_ARG(y && x && rows > 0 && c > 0 && r > 0, "nsr_add_repeat_interleave: bad arguments");
  add_repeat_interleave_kernel<<<po_blocks(rows * c * r), 256, 0, ST>>>(y, x, rows, c, r);
  NSR_CHECK_LAUNCH("add_repeat_interleave");
  return NSR_OK;
}
extern "C" size_t nsr_groupnorm_workspace(int batch, int c, int groups) {
  return ((size_t)batch * groups * GN_CHUNKS * 2 + (size_t)batch * GN_CHUNKS * 2 * c) * sizeof(float);
}
static int gn_check(int B, int HW, int C, int G) {
  NSR_CHECK_ARG(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0 && GN_THREADS % (C / G) == 0 && C / G <= GN_THREADS,
                "nsr_groupnorm: need C % G == 0 and C/G dividing 256");
  NSR_CHECK_ARG(B * G * 2 * sizeof(float) <= 40000, "nsr_groupnorm: batch * groups too large");
  return NSR_OK;
}
extern "C" int nsr_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                                 float* mean, float* rstd, int B, int HW, int C, int G, float eps, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = gn_check(B, HW, C, G);
  if (rc) return rc;
  NSR_CHECK_ARG(x && gamma && beta && y && mean && rstd, "nsr_groupnorm_fwd: null pointer");
  if (!workspace || workspace_bytes < nsr_groupnorm_workspace(B, C, G)) {
    set_error("nsr_groupnorm_fwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  gn_partial_kernel<false><<<dim3(GN_CHUNKS, B * G), GN_THREADS, 0, ST>>>(x, nullptr, nullptr, nullptr, nullptr, partial, nullptr,
                                                                         HW, C, G);
  gn_stats_final<<<ceil_div(B * G, 128), 128, 0, ST>>>(partial, mean, rstd, B * G, (float)HW * (float)(C / G), eps);
  gn_apply_fwd<<<po_blocks((size_t)B * HW * C), 256, 0, ST>>>(x, gamma, beta, mean, rstd, residual, y, B, HW, C, G);
  NSR_CHECK_LAUNCH("groupnorm_fwd");
  return NSR_OK;
}
extern "C" int nsr_groupnorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                 float* dx, float* dgamma, float* dbeta, int B, int HW, int C, int G, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = gn_check(B, HW, C, G);
  if (rc) return rc;
  NSR_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "nsr_groupnorm_bwd: null pointer");
  if (!workspace || workspace_bytes < nsr_groupnorm_workspace(B, C, G)) {
    set_error("nsr_groupnorm_bwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  float* chan = partial + (size_t)B * G * GN_CHUNKS * 2;
  gn_partial_kernel<true><<<dim3(GN_CHUNKS, B * G), GN_THREADS, 0, ST>>>(x, dy, gamma, mean, rstd, partial, chan, HW, C, G);
  gn_apply_bwd<<<po_blocks((size_t)B * HW * C), 256, (size_t)B * G * 2 * sizeof(float), ST>>>(dy, x, gamma, mean, rstd, partial, dx,
                                                                                            B, HW, C, G);
  gn_chan_final<<<ceil_div(C, 128), 128, 0, ST>>>(chan, dgamma, dbeta, B * GN_CHUNKS, C);
  NSR_CHECK_LAUNCH("groupnorm_bwd");
  return NSR_OK;
}

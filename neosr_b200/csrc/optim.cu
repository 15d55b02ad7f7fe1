// Fused multi-tensor optimizer step: clip_grad_norm_ + adan_sf (or AdamW) + EMA in ONE pass
// over a device-resident table of tensors.  HBM-bound: adan_sf touches p,g,m,v,d,z,npg,(ema)
// once each (read+write) = 64 B/param vs the reference's ~57 foreach tensor passes.
#include "common.cuh"

namespace nsr {
constexpr int OPT_THREADS = 256;
constexpr int SUMSQ_BLOCKS = kNumSMs * 4;

__device__ __forceinline__ int find_tensor(const NsrParamEntry* __restrict__ tab, int n, long long chunk) {
  int lo = 0, hi = n - 1;  // last entry with chunk_base <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk_base <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(OPT_THREADS) grad_sumsq_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                                 long long total_chunks, float* __restrict__ partial) {
  __shared__ float red[OPT_THREADS / 32];
  float s = 0.f;
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) {
      const float g = e.g[i];
      s = fmaf(g, g, s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < OPT_THREADS / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_final(const float* __restrict__ partial, int blocks, float* out) {
  __shared__ float sm[OPT_THREADS];
  float s = 0.f;
  for (int i = threadIdx.x; i < blocks; i += OPT_THREADS) s += partial[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  return w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.f - w);
}
__device__ __forceinline__ float clip_coef(const float* sumsq, float max_norm) {
  if (max_norm <= 0.f || sumsq == nullptr) return 1.f;
  const float c = max_norm / (sqrtf(*sumsq) + 1e-6f);  // torch.nn.utils.clip_grad_norm_
  return c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(OPT_THREADS) adan_sf_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                              long long total_chunks, NsrAdanSF hp_val,
                                                              const NsrAdanSF* __restrict__ hp_dev,
                                                              const float* __restrict__ sumsq) {
  const NsrAdanSF hp = hp_dev ? *hp_dev : hp_val;
  const float clip = clip_coef(sumsq, hp.max_norm);
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) {
      float g = e.g[i] * clip;
      float p = e.p[i], m = e.exp_avg[i], v = e.exp_avg_sq[i], d = e.exp_avg_diff[i];
      float z = hp.first_step ? p : e.z[i];
      float npg = hp.first_step ? -g : e.neg_pre_grad[i];
      npg = npg + g;
      m = m * hp.beta1 + g * hp.one_minus_beta1;
      d = d * hp.beta2 + npg * hp.one_minus_beta2;
      npg = npg * hp.beta2 + g;
      v = v * hp.beta3 + hp.one_minus_beta3 * npg * npg;
      const float denom = sqrtf(v) / hp.bias_correction3_sqrt + hp.eps;
      p = p * hp.decay;
      if (hp.schedule_free) {
        p = torch_lerp(p, z, hp.ckp1);
        p = p + (-hp.step_size) * (m / denom);
        p = p + (-hp.step_size_diff) * (d / denom);
        z = z - hp.lr * g;
      } else {
        p = p + (-hp.step_size) * (m / denom);
        p = p + (-hp.step_size_diff) * (d / denom);
      }
      e.g[i] = g;
      e.p[i] = p;
      e.exp_avg[i] = m;
      e.exp_avg_sq[i] = v;
      e.exp_avg_diff[i] = d;
      e.z[i] = z;
      e.neg_pre_grad[i] = -g;
      if (e.ema && hp.ema_lerp > 0.f) e.ema[i] = hp.ema_first ? p : torch_lerp(e.ema[i], p, hp.ema_lerp);
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                            long long total_chunks, NsrAdamW hp_val,
                                                            const NsrAdamW* __restrict__ hp_dev,
                                                            const float* __restrict__ sumsq) {
  const NsrAdamW hp = hp_dev ? *hp_dev : hp_val;
  const float clip = clip_coef(sumsq, hp.max_norm);
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) {
      const float g = e.g[i] * clip;
      float p = e.p[i] * hp.decay;
      float m = e.exp_avg[i], v = e.exp_avg_sq[i];
      m = torch_lerp(m, g, hp.one_minus_beta1);
      v = v * hp.beta2 + hp.one_minus_beta2 * g * g;
      const float denom = sqrtf(v) / hp.bias_correction2_sqrt + hp.eps;
      p = p + (-hp.step_size) * (m / denom);
      e.g[i] = g;
      e.p[i] = p;
      e.exp_avg[i] = m;
      e.exp_avg_sq[i] = v;
      if (e.ema && hp.ema_lerp > 0.f) e.ema[i] = hp.ema_first ? p : torch_lerp(e.ema[i], p, hp.ema_lerp);
    }
  }
}

// ---- F-SAM (neosr/optimizers/fsam.py:36-80) on the same multi-tensor table: p, g, exp_avg = momentum, z = old_p ----------
// first_step part 1: g <- g - sigma * momentum (not on the first call), momentum <- lmbda * momentum + (1 - lmbda) * g_orig
// (first call: momentum <- g), and the partial sums of ((|p| or 1) * g)^2 for the perturbation norm
__global__ void __launch_bounds__(OPT_THREADS) fsam_grad_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                                long long total_chunks, float sigma, float lmbda, int first,
                                                                int adaptive, float* __restrict__ partial) {
  __shared__ float red[OPT_THREADS / 32];
  float s = 0.f;
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) {
      const float g0 = e.g[i];
      float g = g0;
      if (first) {
        e.exp_avg[i] = g0;
      } else {
        const float m = e.exp_avg[i];
        g = g0 - m * sigma;
        e.exp_avg[i] = m * lmbda + g0 * (1.f - lmbda);
        e.g[i] = g;
      }
      const float w = adaptive ? fabsf(e.p[i]) * g : g;
      s = fmaf(w, w, s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < OPT_THREADS / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
// first_step part 2: old_p <- p;  p <- p + (p^2 or 1) * g * rho / (norm + 1e-12)      (climb to "w + e(w)")
__global__ void __launch_bounds__(OPT_THREADS) fsam_climb_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                                 long long total_chunks, float rho, int adaptive,
                                                                 const float* __restrict__ sumsq) {
  const float scale = rho / (sqrtf(*sumsq) + 1e-12f);
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) {
      const float p = e.p[i];
      e.z[i] = p;
      e.p[i] = p + (adaptive ? p * p : 1.f) * e.g[i] * scale;
    }
  }
}
// second_step: p <- old_p
__global__ void __launch_bounds__(OPT_THREADS) fsam_restore_kernel(const NsrParamEntry* __restrict__ tab, int n_tensors,
                                                                   long long total_chunks) {
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const int ti = find_tensor(tab, n_tensors, ch);
    const NsrParamEntry e = tab[ti];
    const long long off = (ch - e.chunk_base) * NSR_OPT_CHUNK;
    long long end = off + NSR_OPT_CHUNK;
    if (end > e.n) end = e.n;
    for (long long i = off + threadIdx.x; i < end; i += OPT_THREADS) e.p[i] = e.z[i];
  }
}

static int opt_blocks(long long chunks) {
  long long cap = (long long)kNumSMs * 8;
  return (int)(chunks < cap ? (chunks > 0 ? chunks : 1) : cap);
}
}  // namespace nsr
using namespace nsr;

extern "C" size_t nsr_grad_sumsq_workspace(void) { return (size_t)SUMSQ_BLOCKS * sizeof(float); }
extern "C" int nsr_grad_sumsq(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, float* sumsq,
                              void* workspace, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && sumsq && workspace, "nsr_grad_sumsq: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = total_chunks < SUMSQ_BLOCKS ? (int)total_chunks : SUMSQ_BLOCKS;
  grad_sumsq_kernel<<<blocks, OPT_THREADS, 0, st>>>(tab, n_tensors, total_chunks, reinterpret_cast<float*>(workspace));
  NSR_CHECK_LAUNCH("grad_sumsq");
  sumsq_final<<<1, OPT_THREADS, 0, st>>>(reinterpret_cast<const float*>(workspace), blocks, sumsq);
  NSR_CHECK_LAUNCH("sumsq_final");
  return NSR_OK;
}
extern "C" int nsr_adan_sf_step(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, const NsrAdanSF* hp,
                                const float* sumsq, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && hp, "nsr_adan_sf_step: bad arguments");
  adan_sf_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tab, n_tensors, total_chunks, *hp, nullptr, sumsq);
  NSR_CHECK_LAUNCH("adan_sf_step");
  return NSR_OK;
}
extern "C" int nsr_adan_sf_step_dev(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, const NsrAdanSF* hp_dev,
                                    const float* sumsq, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && hp_dev, "nsr_adan_sf_step_dev: bad arguments");
  adan_sf_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tab, n_tensors, total_chunks, NsrAdanSF{}, hp_dev, sumsq);
  NSR_CHECK_LAUNCH("adan_sf_step_dev");
  return NSR_OK;
}
extern "C" int nsr_adamw_step(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, const NsrAdamW* hp,
                              const float* sumsq, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && hp, "nsr_adamw_step: bad arguments");
  adamw_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tab, n_tensors, total_chunks, *hp, nullptr, sumsq);
  NSR_CHECK_LAUNCH("adamw_step");
  return NSR_OK;
}
extern "C" int nsr_adamw_step_dev(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, const NsrAdamW* hp_dev,
                                  const float* sumsq, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && hp_dev, "nsr_adamw_step_dev: bad arguments");
  adamw_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tab, n_tensors, total_chunks, NsrAdamW{}, hp_dev, sumsq);
  NSR_CHECK_LAUNCH("adamw_step_dev");
  return NSR_OK;
}

extern "C" int nsr_fsam_first_step(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, float rho, float sigma,
                                   float lmbda, int adaptive, int first, float* sumsq, void* workspace, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0 && sumsq && workspace, "nsr_fsam_first_step: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = total_chunks < SUMSQ_BLOCKS ? (int)total_chunks : SUMSQ_BLOCKS;
  fsam_grad_kernel<<<blocks, OPT_THREADS, 0, st>>>(tab, n_tensors, total_chunks, sigma, lmbda, first, adaptive,
                                                   reinterpret_cast<float*>(workspace));
  NSR_CHECK_LAUNCH("fsam_grad");
  sumsq_final<<<1, OPT_THREADS, 0, st>>>(reinterpret_cast<const float*>(workspace), blocks, sumsq);
  NSR_CHECK_LAUNCH("sumsq_final");
  fsam_climb_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, st>>>(tab, n_tensors, total_chunks, rho, adaptive, sumsq);
  NSR_CHECK_LAUNCH("fsam_climb");
  return NSR_OK;
}
extern "C" int nsr_fsam_restore(const NsrParamEntry* tab, int n_tensors, int64_t total_chunks, void* stream) {
  NSR_CHECK_ARG(tab && n_tensors > 0 && total_chunks > 0, "nsr_fsam_restore: bad arguments");
  fsam_restore_kernel<<<opt_blocks(total_chunks), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tab, n_tensors,
                                                                                                          total_chunks);
  NSR_CHECK_LAUNCH("fsam_restore");
  return NSR_OK;
}

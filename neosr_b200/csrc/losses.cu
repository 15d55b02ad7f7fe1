// Fused loss value + gradient kernels (one read of each operand, one write of the gradient),
// deterministic two-pass reductions.  HBM-bound: 12 B/element (read pred, target; write grad).
#include "common.cuh"

namespace nsr {
constexpr int LOSS_BLOCKS = kNumSMs * 4;
constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[LOSS_THREADS / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < LOSS_THREADS / 32; ++i) s += red[i];
  }
  return s;  // valid on thread 0
}

enum { K_L1 = 0, K_CHARB = 1, K_BCE = 2 };

// one element: returns the loss term, writes the gradient term to g
template <int KIND>
__device__ __forceinline__ float loss_elem(float a, float b, float p0, float p1, float p2, float gscale, float& g) {
  if (KIND == K_L1) {
    const float d = a - b;
    g = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
    return fabsf(d);
  } else if (KIND == K_CHARB) {  // p0 = in_scale, p1 = clip_min, p2 = clip_max
    const float d = (a - b) * p0;
    const float v = sqrtf(d * d + 1e-12f);
    g = (v >= p1 && v <= p2) ? gscale * p0 * d / v : 0.f;
    return fminf(fmaxf(v, p1), p2);
  } else {  // BCE with logits, p0 = label
    g = (1.f / (1.f + expf(-a)) - p0) * gscale;
    return fmaxf(a, 0.f) - a * p0 + log1pf(expf(-fabsf(a)));
  }
}

// VEC: 16-byte accesses (all pointers 16-byte aligned); the n % 4 tail goes through the scalar path of the last block
template <int KIND, bool VEC>
__global__ void __launch_bounds__(LOSS_THREADS) loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ da, size_t n, float p0, float p1,
                                                            float p2, float gscale, float* __restrict__ partial) {
  float s = 0.f;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (VEC) {
    const size_t n4 = n / 4;
    for (size_t i = tid; i < n4; i += nth) {
      const float4 av = reinterpret_cast<const float4*>(a)[i];
      const float4 bv = (KIND == K_BCE) ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(b)[i];
      float4 g;
      s += loss_elem<KIND>(av.x, bv.x, p0, p1, p2, gscale, g.x);
      s += loss_elem<KIND>(av.y, bv.y, p0, p1, p2, gscale, g.y);
      s += loss_elem<KIND>(av.z, bv.z, p0, p1, p2, gscale, g.z);
      s += loss_elem<KIND>(av.w, bv.w, p0, p1, p2, gscale, g.w);
      if (da) reinterpret_cast<float4*>(da)[i] = g;
    }
    done = n4 * 4;
  }
  for (size_t i = done + tid; i < n; i += nth) {
    float g;
    s += loss_elem<KIND>(a[i], (KIND == K_BCE) ? 0.f : b[i], p0, p1, p2, gscale, g);
    if (da) da[i] = g;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void loss_final(const float* __restrict__ partial, int blocks, float scale, float* loss_accum,
                           float* loss_value) {
  __shared__ float sm[LOSS_THREADS];
  float s = 0.f;
  for (int i = threadIdx.x; i < blocks; i += LOSS_THREADS) s += partial[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = LOSS_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float v = sm[0] * scale;
    if (loss_value) *loss_value = v;
    if (loss_accum) *loss_accum += v;
  }
}

template <int KIND>
static int run_loss(const float* a, const float* b, float* da, size_t n, float p0, float p1, float p2, float weight,
                    float* loss_accum, float* loss_value, void* workspace, void* stream, const char* name) {
  NSR_CHECK_ARG(a && (b || KIND == K_BCE) && n > 0 && workspace, "%s: bad arguments", name);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(da)) & 15) == 0 && n >= 4;
  const size_t items = vec ? (n + 3) / 4 : n;
  int blocks = (int)((items + LOSS_THREADS - 1) / LOSS_THREADS);
  if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
  float* partial = reinterpret_cast<float*>(workspace);
  const float inv_n = (float)(1.0 / (double)n);
  if (vec) loss_kernel<KIND, true><<<blocks, LOSS_THREADS, 0, st>>>(a, b, da, n, p0, p1, p2, weight * inv_n, partial);
  else loss_kernel<KIND, false><<<blocks, LOSS_THREADS, 0, st>>>(a, b, da, n, p0, p1, p2, weight * inv_n, partial);
  NSR_CHECK_LAUNCH(name);
  loss_final<<<1, LOSS_THREADS, 0, st>>>(partial, blocks, weight * inv_n, loss_accum, loss_value);
  NSR_CHECK_LAUNCH(name);
  return NSR_OK;
}
}  // namespace nsr
using namespace nsr;

extern "C" size_t nsr_loss_workspace(void) { return (size_t)LOSS_BLOCKS * sizeof(float); }
extern "C" int nsr_l1_loss(const float* pred, const float* target, float* dpred, size_t n, float weight,
                           float* loss_accum, float* loss_value, void* workspace, void* stream) {
  return run_loss<K_L1>(pred, target, dpred, n, 0.f, 0.f, 0.f, weight, loss_accum, loss_value, workspace, stream, "nsr_l1_loss");
}
extern "C" int nsr_charbonnier_loss(const float* a, const float* b, float* da, size_t n, float in_scale, float clip_min,
                                    float clip_max, float weight, float* loss_accum, float* loss_value, void* workspace,
                                    void* stream) {
  return run_loss<K_CHARB>(a, b, da, n, in_scale, clip_min, clip_max, weight, loss_accum, loss_value, workspace, stream,
                           "nsr_charbonnier_loss");
}
extern "C" int nsr_bce_logits_loss(const float* logits, float* dlogits, size_t n, float label, float weight,
                                   float* loss_accum, float* loss_value, void* workspace, void* stream) {
  return run_loss<K_BCE>(logits, nullptr, dlogits, n, label, 0.f, 0.f, weight, loss_accum, loss_value, workspace, stream,
                         "nsr_bce_logits_loss");
}

// Shared helpers for the neosr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/neosr_b200.h"

namespace nsr {

void set_error(const char* fmt, ...);

#define NSR_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      nsr::set_error(__VA_ARGS__);          \
      return NSR_E_INVALID;                 \
    }                                       \
  } while (0)

#define NSR_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      nsr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return NSR_E_CUDA;                                                          \
    }                                                                             \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
__host__ __device__ static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// NSR_ENGINE_BF16 = NSR_ENGINE_AUTO routing with single-pass tensor-core products
static inline int mma_passes(int engine) { return engine == NSR_ENGINE_BF16 ? 1 : 3; }
static inline int route_engine(int engine) { return engine == NSR_ENGINE_BF16 ? NSR_ENGINE_AUTO : engine; }
constexpr int kNumSMs = 148;  // B200
// row blocks of the bias-gradient column sums (colsum_partial): 256 rows each, at most four waves.  (1024 rows per block
// left a 32768-pixel problem - HAT at B = 8 - with 32 CTAs: 31 us per launch, 11 % of that step.)
static inline int bias_grad_blocks(long long rows) {
  long long b = (rows + 255) / 256;
  return (int)(b > kNumSMs * 4 ? kNumSMs * 4 : (b < 1 ? 1 : b));
}

// ---- activation math shared by every epilogue -------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// Fast exact-form GELU for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, below fp32 resolution of the O(1) activations it feeds), sharing one
// exp(-x^2/2) between the cdf and the pdf.  ~15 instructions instead of ~45 for erff + expf.
// raw approx instructions: no range fix-up code around them (the arguments are bounded: the exponent is <= 0, the
// reciprocal's argument is >= 1); ex2.approx.ftz rel. error 2^-22, rcp.approx.ftz 1 ulp
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_terms(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float e = ex2_approx(z * z * -1.4426950408889634f);  // = exp(-x^2 / 2)
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * e;            // erf(|x| / sqrt 2)
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float c, p;
  gelu_terms(x, c, p);
  return x * c;
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
  float c, p;
  gelu_terms(x, c, p);
  return fmaf(x, p, c);
}
__device__ __forceinline__ float apply_act_fast(float v, int act, float slope) {
  switch (act) {
    case NSR_ACT_RELU: return v > 0.f ? v : 0.f;
    case NSR_ACT_LRELU:
    case NSR_ACT_PRELU: return v > 0.f ? v : v * slope;
    case NSR_ACT_GELU: return gelu_fast(v);
    default: return v;
  }
}
__device__ __forceinline__ float act_grad_fast(float aux, int act, float slope) {
  switch (act) {
    case NSR_ACT_RELU: return aux > 0.f ? 1.f : 0.f;
    case NSR_ACT_LRELU:
    case NSR_ACT_PRELU: return aux > 0.f ? 1.f : slope;
    case NSR_ACT_GELU: return gelu_fast_grad(aux);
    default: return 1.f;
  }
}
// compile-time-selected versions (tcgen05 epilogue): value and derivative in one go
template <int ACT>
__device__ __forceinline__ void act_value_grad(float v, float slope, float& val, float& grad) {
  if (ACT == NSR_ACT_GELU) {
    float c, p;
    gelu_terms(v, c, p);
    val = v * c;
    grad = fmaf(v, p, c);
  } else if (ACT == NSR_ACT_RELU) {
    val = v > 0.f ? v : 0.f;
    grad = v > 0.f ? 1.f : 0.f;
  } else if (ACT == NSR_ACT_LRELU || ACT == NSR_ACT_PRELU) {
    val = v > 0.f ? v : v * slope;
    grad = v > 0.f ? 1.f : slope;
  } else {
    val = v;
    grad = 1.f;
  }
}
template <int AG>
__device__ __forceinline__ float act_grad_ct(float aux, float slope) {
  if (AG == NSR_ACT_MULAUX) return aux;
  if (AG == NSR_ACT_GELU) return gelu_fast_grad(aux);
  if (AG == NSR_ACT_RELU) return aux > 0.f ? 1.f : 0.f;
  if (AG == NSR_ACT_LRELU || AG == NSR_ACT_PRELU) return aux > 0.f ? 1.f : slope;
  return 1.f;
}
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case NSR_ACT_RELU: return v > 0.f ? v : 0.f;
    case NSR_ACT_LRELU:
    case NSR_ACT_PRELU: return v > 0.f ? v : v * slope;
    case NSR_ACT_GELU: return gelu_erf(v);
    default: return v;
  }
}
// derivative of act evaluated from `aux`: for (L)ReLU aux may be the activation OUTPUT
// (sign is preserved for slope > 0); for GELU / PReLU aux is the PRE-activation.
__device__ __forceinline__ float act_grad(float aux, int act, float slope) {
  switch (act) {
    case NSR_ACT_MULAUX: return aux;
    case NSR_ACT_RELU: return aux > 0.f ? 1.f : 0.f;
    case NSR_ACT_LRELU:
    case NSR_ACT_PRELU: return aux > 0.f ? 1.f : slope;
    case NSR_ACT_GELU: return gelu_erf_grad(aux);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// packed-weight buffer geometry (see nsr_pack_weight in include/neosr_b200.h)
struct PackedGeom {
  int n, taps, c;        // GEMM view W[n][tap][c]
  int n_pad64;           // n rounded up to 64
  int cblks;             // ceil(c / 64)
  size_t f32_bytes;      // fp32 region (1024-aligned)
  size_t bf16_bytes;     // hi/lo tile images
};
__host__ __device__ static inline PackedGeom packed_geom(int cout, int cin, int kh, int kw, int flavour) {
  PackedGeom g;
  g.n = flavour == 0 ? cout : cin;
  g.c = flavour == 0 ? cin : cout;
  g.taps = kh * kw;
  g.n_pad64 = (g.n + 63) / 64 * 64;
  g.cblks = (g.c + 63) / 64;
  g.f32_bytes = align_up((size_t)g.n * g.taps * g.c * sizeof(float), 1024);
  g.bf16_bytes = (size_t)g.taps * g.cblks * 2 * g.n_pad64 * 128;
  return g;
}

}  // namespace nsr

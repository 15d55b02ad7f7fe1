// C-ABI entry points for the contraction family (Linear / Conv2d fprop, dgrad, wgrad) and
// the weight packer.  Engine dispatch lives here.
#include <stdlib.h>

#include <cstdlib>

#include "common.cuh"

namespace nsr {
int conv_fprop_simt(const NsrConv& d, cudaStream_t st);
int conv_wgrad_simt(const NsrWgrad& d, cudaStream_t st);
size_t conv_wgrad_workspace_simt(const NsrWgrad& d);
// tcgen05 engine (igemm_tc.cu)
bool conv_fprop_tc_supported(const NsrConv& d);
int conv_fprop_tc(const NsrConv& d, cudaStream_t st);
bool conv_wgrad_tc_supported(const NsrWgrad& d);
size_t conv_wgrad_workspace_tc(const NsrWgrad& d);
int conv_wgrad_tc(const NsrWgrad& d, cudaStream_t st);
int conv_wgrad_tc_partial(const NsrWgrad& d, int* splitk, cudaStream_t st);
size_t conv_wgrad_tc_partial_bytes(const NsrWgrad& d);
// <= 4-channel image-side convolutions (conv_small.cu)
// direct 3x3 kernels for <= 4-channel sides at image resolution (conv_direct.cu)
bool conv_direct_fprop_supported(const NsrConv& d);
int conv_direct_fprop(const NsrConv& d, cudaStream_t st);
bool conv_small_fprop_supported(const NsrConv& d);
int conv_small_fprop(const NsrConv& d, cudaStream_t st);
// <= 4-channel convs as im2col + tensor-core GEMM (conv_narrow_gemm.cu)
bool conv_narrow_gemm_supported(const NsrConv& d);
size_t conv_narrow_gemm_workspace(const NsrConv& d);
int conv_narrow_gemm_fprop(const NsrConv& d, cudaStream_t st);
bool conv_narrow_gemm_wgrad_supported(const NsrWgrad& d);
size_t conv_narrow_gemm_wgrad_workspace(const NsrWgrad& d);
int conv_narrow_gemm_wgrad(const NsrWgrad& d, cudaStream_t st);
bool conv_wgrad_tma_supported(const NsrWgrad& d);
size_t conv_wgrad_workspace_tma(const NsrWgrad& d);
int conv_wgrad_tma(const NsrWgrad& d, cudaStream_t st);
bool conv_small_wgrad_supported(const NsrWgrad& d);
size_t conv_small_wgrad_workspace(const NsrWgrad& d);
int conv_small_wgrad(const NsrWgrad& d, cudaStream_t st);

static int forced_engine() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("NSR_ENGINE");
    cached = NSR_ENGINE_AUTO;
    if (e && (!strcmp(e, "simt") || !strcmp(e, "1"))) cached = NSR_ENGINE_SIMT;
    if (e && (!strcmp(e, "tcgen05") || !strcmp(e, "2"))) cached = NSR_ENGINE_TCGEN05;
  }
  return cached;
}

// ---- weight packing ---------------------------------------------------------------------
// fp32 region: W[n][tap][c]; flavour 0: n=co, c=ci, tap=(r,s); flavour 1: n=ci, c=co, tap flipped.
__device__ __forceinline__ float packed_src(const float* __restrict__ w, int n, int tap, int c, int cin, int taps, int flavour) {
  const int co = flavour == 0 ? n : c, ci = flavour == 0 ? c : n;
  const int src_tap = flavour == 0 ? tap : taps - 1 - tap;  // 180-degree rotation
  return w[((size_t)co * cin + ci) * taps + src_tap];
}
__device__ __forceinline__ void pack_f32_range(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps,
                                               int flavour, size_t first, size_t step) {
  const int N = flavour == 0 ? cout : cin, C = flavour == 0 ? cin : cout;
  const size_t total = (size_t)N * taps * C;
  for (size_t i = first; i < total; i += step) {
    const int c = (int)(i % C);
    const size_t r = i / C;
    out[i] = packed_src(w, (int)(r / taps), (int)(r % taps), c, cin, taps, flavour);
  }
}
__global__ void pack_weight_f32(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int kh,
                                int kw, int flavour) {
  pack_f32_range(w, out, cout, cin, kh * kw, flavour, blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// bf16 hi/lo tile images (see common.cuh PackedGeom and igemm_tc.cu): for k-block kb=(tap,cblk),
// half h, 64-row block j: 8 KiB tile, row r = 128 B, 16-byte chunk ch stored at ch ^ (r & 7).
// SRC_RAW: read the parameter tensor itself (OIHW) instead of the fp32 region (used by the one-launch packer)
template <bool SRC_RAW>
__device__ __forceinline__ void pack_bf16_range(const float* __restrict__ src, uint8_t* __restrict__ img, const PackedGeom& g,
                                                int cin, int flavour, size_t first, size_t step) {
  const int nblk = g.n_pad64 / 64;
  const size_t total = (size_t)g.taps * g.cblks * g.n_pad64 * 8;  // one thread per (kb, n, chunk)
  for (size_t i = first; i < total; i += step) {
    const int ch = (int)(i & 7);
    size_t r = i >> 3;
    const int n = (int)(r % g.n_pad64);
    const int kb = (int)(r / g.n_pad64);
    const int tap = kb / g.cblks, cblk = kb - tap * g.cblks;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cblk * 64 + ch * 8 + e;
      float v = 0.f;
      if (n < g.n && c < g.c) v = SRC_RAW ? packed_src(src, n, tap, c, cin, g.taps, flavour) : src[((size_t)n * g.taps + tap) * g.c + c];
      hi[e] = __float2bfloat16_rn(v);
      lo[e] = __float2bfloat16_rn(v - __bfloat162float(hi[e]));
    }
    const int j = n >> 6, row = n & 63;
    const size_t tile_hi = ((size_t)(kb * 2 + 0) * nblk + j) * 8192;
    const size_t tile_lo = ((size_t)(kb * 2 + 1) * nblk + j) * 8192;
    const size_t off = (size_t)row * 128 + (size_t)((ch ^ (row & 7)) * 16);
    *reinterpret_cast<uint4*>(img + tile_hi + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + tile_lo + off) = *reinterpret_cast<const uint4*>(lo);
  }
}
__global__ void pack_weight_bf16(const float* __restrict__ wf32, uint8_t* __restrict__ img, PackedGeom g) {
  pack_bf16_range<false>(wf32, img, g, 0, 0, blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}
// one launch per parameter: fp32 region + bf16 tile images of the fprop packing and (p1 != nullptr) the dgrad packing
__global__ void pack_weight_all(const float* __restrict__ w, uint8_t* __restrict__ p0, uint8_t* __restrict__ p1, PackedGeom g0,
                                PackedGeom g1, int cout, int cin) {
  const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  pack_f32_range(w, reinterpret_cast<float*>(p0), cout, cin, g0.taps, 0, first, step);
  pack_bf16_range<true>(w, p0 + g0.f32_bytes, g0, cin, 0, first, step);
  if (p1) {
    pack_f32_range(w, reinterpret_cast<float*>(p1), cout, cin, g1.taps, 1, first, step);
    pack_bf16_range<true>(w, p1 + g1.f32_bytes, g1, cin, 1, first, step);
  }
}

// ---- every weight of a network in ONE launch ------------------------------------------------------------------------
// Re-packing after an optimizer step used to be one launch per parameter (+ two gathers for every head-padded copy):
// ~400 launches of ~5 us each per SwinIR-medium step (profiles/r02_r_ncu_kernels_c3.md).  Here a device table lists
// (source, maps, destinations, geometry) per entry and the CTAs are dealt out over the entries.
__device__ __forceinline__ float pack_entry_src(const NsrPackEntry& e, int n, int tap, int c, int taps, int flavour) {
  int co = flavour == 0 ? n : c, ci = flavour == 0 ? c : n;
  const int src_tap = flavour == 0 ? tap : taps - 1 - tap;
  if (e.row_map) co = e.row_map[co];
  if (e.col_map) ci = e.col_map[ci];
  if (co < 0 || ci < 0) return 0.f;
  return e.w[((size_t)co * e.src_cin + ci) * taps + src_tap];
}
__device__ __forceinline__ void pack_entry_flavour(const NsrPackEntry& e, uint8_t* __restrict__ dst, const PackedGeom& g, int flavour,
                                                   size_t first, size_t step) {
  float* f32 = reinterpret_cast<float*>(dst);
  const size_t total_f = (size_t)g.n * g.taps * g.c;
  for (size_t i = first; i < total_f; i += step) {
    const int c = (int)(i % g.c);
    const size_t r = i / g.c;
    f32[i] = pack_entry_src(e, (int)(r / g.taps), (int)(r % g.taps), c, g.taps, flavour);
  }
  uint8_t* img = dst + g.f32_bytes;
  const int nblk = g.n_pad64 / 64;
  const size_t total = (size_t)g.taps * g.cblks * g.n_pad64 * 8;
  for (size_t i = first; i < total; i += step) {
    const int ch = (int)(i & 7);
    const size_t r = i >> 3;
    const int n = (int)(r % g.n_pad64);
    const int kb = (int)(r / g.n_pad64);
    const int tap = kb / g.cblks, cblk = kb - tap * g.cblks;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cblk * 64 + ch * 8 + k;
      const float v = (n < g.n && c < g.c) ? pack_entry_src(e, n, tap, c, g.taps, flavour) : 0.f;
      hi[k] = __float2bfloat16_rn(v);
      lo[k] = __float2bfloat16_rn(v - __bfloat162float(hi[k]));
    }
    const int j = n >> 6, row = n & 63;
    const size_t off = (size_t)row * 128 + (size_t)((ch ^ (row & 7)) * 16);
    *reinterpret_cast<uint4*>(img + ((size_t)(kb * 2 + 0) * nblk + j) * 8192 + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + ((size_t)(kb * 2 + 1) * nblk + j) * 8192 + off) = *reinterpret_cast<const uint4*>(lo);
  }
}
__global__ void __launch_bounds__(256) pack_weights_multi(const NsrPackEntry* __restrict__ tab, int n_entries) {
  int lo_i = 0, hi_i = n_entries - 1;  // last entry with block_base <= blockIdx.x
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (tab[mid].block_base <= (long long)blockIdx.x) lo_i = mid; else hi_i = mid - 1;
  }
  const NsrPackEntry e = tab[lo_i];
  const long long nblocks = (lo_i + 1 < n_entries ? tab[lo_i + 1].block_base : (long long)gridDim.x) - e.block_base;
  const size_t first = (size_t)(blockIdx.x - e.block_base) * 256 + threadIdx.x, step = (size_t)nblocks * 256;
  const PackedGeom g0 = packed_geom(e.cout, e.cin, e.kh, e.kw, 0), g1 = packed_geom(e.cout, e.cin, e.kh, e.kw, 1);
  if (e.packed_fprop) pack_entry_flavour(e, reinterpret_cast<uint8_t*>(e.packed_fprop), g0, 0, first, step);
  if (e.packed_dgrad) pack_entry_flavour(e, reinterpret_cast<uint8_t*>(e.packed_dgrad), g1, 1, first, step);
  if (e.bias_out)
    for (size_t i = first; i < (size_t)e.cout; i += step) {
      const int src = e.row_map ? e.row_map[i] : (int)i;
      e.bias_out[i] = src < 0 ? 0.f : e.bias[src];
    }
}

int launch_pack_weight_bf16(const float* wf32, uint8_t* img, const PackedGeom& g, cudaStream_t st) {
  const size_t t2 = (size_t)g.taps * g.cblks * g.n_pad64 * 8;
  int blocks = ceil_div(t2, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_bf16<<<blocks, 256, 0, st>>>(wf32, img, g);
  NSR_CHECK_LAUNCH("pack_weight_bf16");
  return NSR_OK;
}

// ---- split tile image <-> fp32 (utility; producers emit STI directly on the hot path) -------
__global__ void sti_from_f32_kernel(const float* __restrict__ x, int ld, long long rows, int c, uint8_t* __restrict__ sti) {
  const int kbs = (c + 63) / 64;
  const long long mts = (rows + 127) / 128;
  const long long total = mts * 128 * kbs * 8;  // one thread per (row, 16-byte chunk)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % (kbs * 8));
    const long long p = i / (kbs * 8);
    const int c0 = ch * 8;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
      if (p < rows && c0 + e < c) v = x[p * ld + c0 + e];
      hi[e] = __float2bfloat16_rn(v);
      lo[e] = __float2bfloat16_rn(v - __bfloat162float(hi[e]));
    }
    const long long mt = p >> 7;
    const int r = (int)(p & 127), kb = ch >> 3, cc = ch & 7;
    uint8_t* dst = sti + ((size_t)(mt * kbs + kb) << 15) + r * 128 + ((cc ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + 16384) = *reinterpret_cast<const uint4*>(lo);
  }
}
__global__ void sti_to_f32_kernel(const uint8_t* __restrict__ sti, long long rows, int c, float* __restrict__ y, int ld) {
  const int kbs = (c + 63) / 64;
  const long long total = rows * kbs * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % (kbs * 8));
    const long long p = i / (kbs * 8);
    const long long mt = p >> 7;
    const int r = (int)(p & 127), kb = ch >> 3, cc = ch & 7;
    const uint8_t* src = sti + ((size_t)(mt * kbs + kb) << 15) + r * 128 + ((cc ^ (r & 7)) << 4);
    __align__(16) __nv_bfloat16 hi[8], lo[8];
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(src + 16384);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (ch * 8 + e < c) y[p * ld + ch * 8 + e] = __bfloat162float(hi[e]) + __bfloat162float(lo[e]);
  }
}

}  // namespace nsr

using namespace nsr;

extern "C" size_t nsr_sti_bytes(long long rows, int c) {
  if (rows <= 0 || c <= 0) return 0;
  return (size_t)((rows + 127) / 128) * (size_t)((c + 63) / 64) * 32768;
}
extern "C" int nsr_sti_from_f32(const float* x, int ld, long long rows, int c, void* sti, void* stream) {
  NSR_CHECK_ARG(x && sti && rows > 0 && c > 0 && ld >= c, "nsr_sti_from_f32: bad arguments");
  const long long total = ((rows + 127) / 128) * 128 * ((c + 63) / 64) * 8;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  sti_from_f32_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ld, rows, c, reinterpret_cast<uint8_t*>(sti));
  NSR_CHECK_LAUNCH("sti_from_f32");
  return NSR_OK;
}
extern "C" int nsr_sti_to_f32(const void* sti, long long rows, int c, float* y, int ld, void* stream) {
  NSR_CHECK_ARG(y && sti && rows > 0 && c > 0 && ld >= c, "nsr_sti_to_f32: bad arguments");
  const long long total = rows * ((c + 63) / 64) * 8;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  sti_to_f32_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint8_t*>(sti), rows, c, y, ld);
  NSR_CHECK_LAUNCH("sti_to_f32");
  return NSR_OK;
}

extern "C" size_t nsr_packed_weight_bytes(int cout, int cin, int kh, int kw, int flavour) {
  PackedGeom g = packed_geom(cout, cin, kh, kw, flavour);
  return g.f32_bytes + g.bf16_bytes;
}

extern "C" int nsr_pack_weight(const float* w, int cout, int cin, int kh, int kw, int flavour, void* packed,
                               void* stream) {
  NSR_CHECK_ARG(w && packed && cout > 0 && cin > 0 && kh > 0 && kw > 0 && (flavour == 0 || flavour == 1),
                "nsr_pack_weight: bad arguments");
  NSR_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 1023) == 0 || (reinterpret_cast<uintptr_t>(packed) & 255) == 0,
                "nsr_pack_weight: packed buffer must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PackedGeom g = packed_geom(cout, cin, kh, kw, flavour);
  const size_t total = (size_t)g.n * g.taps * g.c;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_f32<<<blocks, 256, 0, st>>>(w, reinterpret_cast<float*>(packed), cout, cin, kh, kw, flavour);
  NSR_CHECK_LAUNCH("pack_weight_f32");
  const size_t t2 = (size_t)g.taps * g.cblks * g.n_pad64 * 8;
  blocks = ceil_div(t2, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_bf16<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(packed),
                                          reinterpret_cast<uint8_t*>(packed) + g.f32_bytes, g);
  NSR_CHECK_LAUNCH("pack_weight_bf16");
  return NSR_OK;
}

extern "C" int nsr_pack_weight_pair(const float* w, int cout, int cin, int kh, int kw, void* packed_fprop, void* packed_dgrad,
                                    void* stream) {
  NSR_CHECK_ARG(w && packed_fprop && cout > 0 && cin > 0 && kh > 0 && kw > 0, "nsr_pack_weight_pair: bad arguments");
  NSR_CHECK_ARG((reinterpret_cast<uintptr_t>(packed_fprop) & 255) == 0 && (reinterpret_cast<uintptr_t>(packed_dgrad) & 255) == 0,
                "nsr_pack_weight_pair: packed buffers must be 256-byte aligned");
  const PackedGeom g0 = packed_geom(cout, cin, kh, kw, 0), g1 = packed_geom(cout, cin, kh, kw, 1);
  size_t work = (size_t)g0.taps * g0.cblks * g0.n_pad64 * 8;
  const size_t w1 = (size_t)g1.taps * g1.cblks * g1.n_pad64 * 8, wf = (size_t)cout * cin * kh * kw;
  if (packed_dgrad && w1 > work) work = w1;
  if (wf > work) work = wf;
  int blocks = ceil_div(work, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_all<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<uint8_t*>(packed_fprop), reinterpret_cast<uint8_t*>(packed_dgrad), g0, g1, cout, cin);
  NSR_CHECK_LAUNCH("pack_weight_all");
  return NSR_OK;
}

// ---- deferred, batched split-K reduction of the 1x1 weight gradients ------------------------------------------------
// Every STI wgrad used to be followed by its own reduce launch, a bias-column split and (head-padded operands) two
// un-padding gathers: ~360 launches of 3-13 us per SwinIR-medium step.  The partials now stay in per-layer buffers
// until the end of the backward pass, where ONE launch reduces all of them (fixed order: deterministic) straight
// into dW / dbias, un-padding rows and columns and taking the bias-gradient column on the way.
__global__ void __launch_bounds__(256) wgrad_finalize_multi(const NsrReduceEntry* __restrict__ tab, int n_entries) {
  int lo_i = 0, hi_i = n_entries - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (tab[mid].block_base <= (long long)blockIdx.x) lo_i = mid; else hi_i = mid - 1;
  }
  const NsrReduceEntry e = tab[lo_i];
  const long long nblocks = (lo_i + 1 < n_entries ? tab[lo_i + 1].block_base : (long long)gridDim.x) - e.block_base;
  const size_t first = (size_t)(blockIdx.x - e.block_base) * 256 + threadIdx.x, step = (size_t)nblocks * 256;
  const size_t per_split = (size_t)e.p_rows * e.p_cols;
  const int cols1 = e.cin + (e.dbias ? 1 : 0);  // column cin of the work space = the bias gradient
  const size_t n = (size_t)e.cout * cols1;
  for (size_t i = first; i < n; i += step) {
    const int co = (int)(i / cols1), c = (int)(i - (size_t)co * cols1);
    const int pr = e.row_map ? e.row_map[co] : co;
    const int pc = c == e.cin ? e.bias_col : (e.col_map ? e.col_map[c] : c);
    const float* src = e.partial + (size_t)pr * e.p_cols + pc;
    float s = 0.f;
    for (int k = 0; k < e.splitk; ++k) s += src[(size_t)k * per_split];
    if (c == e.cin) e.dbias[co] = s; else e.dw[(size_t)co * e.cin + c] = s;
  }
}

extern "C" size_t nsr_conv_wgrad_partial_workspace(const NsrWgrad* d) {
  if (!d || !conv_wgrad_tc_supported(*d)) return 0;
  return conv_wgrad_tc_partial_bytes(*d);
}
extern "C" int nsr_conv_wgrad_partial(const NsrWgrad* d, int* splitk, void* stream) {
  NSR_CHECK_ARG(d && splitk, "nsr_conv_wgrad_partial: null descriptor / splitk");
  NSR_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "nsr_conv_wgrad_partial: bad geometry");
  NSR_CHECK_ARG(d->x_sti && d->dy_sti && d->kh == 1 && d->kw == 1 && conv_wgrad_tc_supported(*d),
                "nsr_conv_wgrad_partial: 1x1 contractions on split-tile-image operands (tcgen05 engine) only");
  return conv_wgrad_tc_partial(*d, splitk, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int64_t nsr_reduce_entry_blocks(int cout, int cin) {
  const int64_t b = ((int64_t)cout * (cin + 1) + 1023) / 1024;  // ~4 outputs per thread
  return b < 1 ? 1 : b;
}
extern "C" int nsr_wgrad_finalize_multi(const NsrReduceEntry* table_dev, int n_entries, int64_t total_blocks, void* stream) {
  NSR_CHECK_ARG(table_dev && n_entries > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "nsr_wgrad_finalize_multi: bad arguments");
  wgrad_finalize_multi<<<(unsigned)total_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table_dev, n_entries);
  NSR_CHECK_LAUNCH("wgrad_finalize_multi");
  return NSR_OK;
}

extern "C" int64_t nsr_pack_entry_blocks(int cout, int cin, int kh, int kw) {
  const PackedGeom g0 = packed_geom(cout, cin, kh, kw, 0), g1 = packed_geom(cout, cin, kh, kw, 1);
  size_t work = (size_t)g0.taps * g0.cblks * g0.n_pad64 * 8;
  const size_t w1 = (size_t)g1.taps * g1.cblks * g1.n_pad64 * 8, wf = (size_t)cout * cin * kh * kw;
  if (w1 > work) work = w1;
  if (wf > work) work = wf;
  int64_t blocks = (int64_t)((work + 2047) / 2048);  // ~8 items per thread
  return blocks < 1 ? 1 : blocks;
}
extern "C" int nsr_pack_weights_multi(const NsrPackEntry* table_dev, int n_entries, int64_t total_blocks, void* stream) {
  NSR_CHECK_ARG(table_dev && n_entries > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "nsr_pack_weights_multi: bad arguments");
  pack_weights_multi<<<(unsigned)total_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table_dev, n_entries);
  NSR_CHECK_LAUNCH("pack_weights_multi");
  return NSR_OK;
}

// NSR_NARROW_GEMM=0 keeps the SIMT kernels of conv_small.cu for the <= 4-channel convolutions (A/B runs)
// NSR_CONV_DIRECT=0 routes the <= 4-channel 3x3 convolutions back through im2col + tcgen05 (A/B runs)
static bool conv_direct_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NSR_CONV_DIRECT");
    on = !(e && e[0] == '0');
  }
  return on != 0;
}
static bool narrow_gemm_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NSR_NARROW_GEMM");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled != 0;
}

static int check_conv(const NsrConv* d) {
  NSR_CHECK_ARG(d, "nsr_conv_fprop: null descriptor");
  NSR_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "nsr_conv_fprop: bad geometry");
  NSR_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->kh == 2 * d->pad + 1 && d->kw == 2 * d->pad + 1,
                "nsr_conv_fprop: only stride-1 'same' convolutions (k = 2*pad+1) are supported");
  NSR_CHECK_ARG(d->x_ld >= d->cin && d->y_ld >= d->cout, "nsr_conv_fprop: leading dims too small");
  NSR_CHECK_ARG((d->x || d->x_sti) && d->w_packed && (d->y || d->y_sti), "nsr_conv_fprop: null x / w / y");
  NSR_CHECK_ARG(!(d->actgrad) || d->aux, "nsr_conv_fprop: actgrad needs aux");
  NSR_CHECK_ARG(!(d->act == NSR_ACT_PRELU || d->actgrad == NSR_ACT_PRELU) || d->prelu, "nsr_conv_fprop: prelu slopes missing");
  NSR_CHECK_ARG(d->act >= 0 && d->act <= NSR_ACT_PRELU && d->actgrad >= 0 && d->actgrad <= NSR_ACT_MULAUX,
                "nsr_conv_fprop: bad activation code");
  return NSR_OK;
}

extern "C" int nsr_conv_fprop(const NsrConv* d, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int eng = route_engine(d->engine);  // NSR_ENGINE_BF16 routes like AUTO; the tcgen05 launchers read d->engine for the pass count
  if (eng == NSR_ENGINE_AUTO) eng = forced_engine();
  if (eng == NSR_ENGINE_TCGEN05) {
    NSR_CHECK_ARG(conv_fprop_tc_supported(*d), "nsr_conv_fprop: shape not supported by the tcgen05 engine");
    return conv_fprop_tc(*d, st);
  }
  if (eng == NSR_ENGINE_AUTO && conv_fprop_tc_supported(*d)) return conv_fprop_tc(*d, st);
  NSR_CHECK_ARG(d->x && d->y && !d->y_sti, "nsr_conv_fprop: split-tile-image operands need the tcgen05 engine");
  NSR_CHECK_ARG(d->pre_mode != 2 && d->aux_mode == 0, "nsr_conv_fprop: 16-bit activation-gradient codes need the tcgen05 engine");
  if (eng == NSR_ENGINE_AUTO && conv_direct_enabled() && conv_direct_fprop_supported(*d)) return conv_direct_fprop(*d, st);
  if (eng == NSR_ENGINE_AUTO && narrow_gemm_enabled() && conv_narrow_gemm_supported(*d)) return conv_narrow_gemm_fprop(*d, st);
  if (eng == NSR_ENGINE_AUTO && conv_small_fprop_supported(*d)) return conv_small_fprop(*d, st);
  return conv_fprop_simt(*d, st);
}

extern "C" size_t nsr_conv_fprop_workspace(const NsrConv* d) {
  if (!d || !narrow_gemm_enabled() || !nsr_device_supports_tcgen05()) return 0;
  int eng = route_engine(d->engine) == NSR_ENGINE_AUTO ? forced_engine() : d->engine;
  if (eng != NSR_ENGINE_AUTO) return 0;
  return conv_narrow_gemm_workspace(*d);
}

static int check_wgrad(const NsrWgrad* d) {
  NSR_CHECK_ARG(d, "nsr_conv_wgrad: null descriptor");
  NSR_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "nsr_conv_wgrad: bad geometry");
  NSR_CHECK_ARG(d->kh == 2 * d->pad + 1 && d->kw == 2 * d->pad + 1, "nsr_conv_wgrad: only stride-1 'same' convolutions");
  NSR_CHECK_ARG(d->x_ld >= d->cin && d->dy_ld >= d->cout, "nsr_conv_wgrad: leading dims too small");
  NSR_CHECK_ARG(((d->x && d->dy) || (d->x_sti && d->dy_sti)) && d->dw, "nsr_conv_wgrad: null x / dy / dw");
  return NSR_OK;
}

static bool wgrad_use_tc(const NsrWgrad* d) {
  int eng = route_engine(d->engine);
  if (eng == NSR_ENGINE_AUTO) eng = forced_engine();
  if (eng == NSR_ENGINE_SIMT) return false;
  return conv_wgrad_tc_supported(*d);
}

// TMA-staged 3x3 wgrad (igemm_wgrad_tma.cu); NSR_WGRAD_TMA=0 keeps the producer-warp kernel (A/B runs)
static bool wgrad_use_tma(const NsrWgrad* d) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NSR_WGRAD_TMA");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return false;
  const bool sti = d->x_sti != nullptr && d->dy_sti != nullptr && d->kh == 1 && d->kw == 1;
  return !sti && conv_wgrad_tma_supported(*d);
}

extern "C" size_t nsr_conv_wgrad_workspace(const NsrWgrad* d) {
  if (!d) return 0;
  size_t a = conv_wgrad_workspace_simt(*d);
  if (wgrad_use_tma(d)) {
    const size_t t = conv_wgrad_workspace_tma(*d);
    a = a > t ? a : t;
  }
  size_t b = conv_wgrad_tc_supported(*d) ? conv_wgrad_workspace_tc(*d) : 0;
  size_t c = conv_small_wgrad_supported(*d) ? conv_small_wgrad_workspace(*d) : 0;
  a = a > b ? a : b;
  a = a > c ? a : c;
  if (narrow_gemm_enabled() && conv_narrow_gemm_wgrad_supported(*d)) {
    const size_t n = conv_narrow_gemm_wgrad_workspace(*d);
    a = a > n ? a : n;
  }
  return a;
}

extern "C" int nsr_conv_wgrad(const NsrWgrad* d, void* stream) {
  int rc = check_wgrad(d);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int eng = route_engine(d->engine) == NSR_ENGINE_AUTO ? forced_engine() : d->engine;
  if (eng == NSR_ENGINE_TCGEN05)
    NSR_CHECK_ARG(conv_wgrad_tc_supported(*d), "nsr_conv_wgrad: shape not supported by the tcgen05 engine");
  if (wgrad_use_tc(d)) return wgrad_use_tma(d) ? conv_wgrad_tma(*d, st) : conv_wgrad_tc(*d, st);
  NSR_CHECK_ARG(d->x && d->dy, "nsr_conv_wgrad: split-tile-image operands need the tcgen05 engine");
  if (eng == NSR_ENGINE_AUTO && narrow_gemm_enabled() && conv_narrow_gemm_wgrad_supported(*d)) return conv_narrow_gemm_wgrad(*d, st);
  if (eng == NSR_ENGINE_AUTO && conv_small_wgrad_supported(*d)) return conv_small_wgrad(*d, st);
  return conv_wgrad_simt(*d, st);
}

// Exact-fp32 CUDA-core implicit GEMM (engine NSR_ENGINE_SIMT).
//
// This is the correctness engine: plain fp32 FMA, fixed reduction order, every shape.
// The tcgen05 engine (igemm_tc.cu) takes the hot shapes; this one covers the rest
// (cin=3 / cout=3 image-side convs, odd channel counts) and is the on-device cross-check.
#include "common.cuh"

namespace nsr {

// ------------------------------------------------------------------------------ fprop
constexpr int FBM = 128, FBN = 64, FBK = 16, FTHREADS = 256;
constexpr int FAS = FBM + 4, FBS = FBN + 4;  // padded smem row strides (floats)

struct RowInfo {
  long long base;  // pixel offset of (b, 0, 0) in pixels
  int oh, ow;
  bool valid;
};

__device__ __forceinline__ float4 ld4_guard(const float* p, int nvalid, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid <= 0 || p == nullptr) return v;
  if (vec && nvalid >= 4) return *reinterpret_cast<const float4*>(p);
  v.x = p[0];
  if (nvalid > 1) v.y = p[1];
  if (nvalid > 2) v.z = p[2];
  if (nvalid > 3) v.w = p[3];
  return v;
}

__global__ void __launch_bounds__(FTHREADS) igemm_fprop_simt(NsrConv d, const float* __restrict__ wq) {
  __shared__ __align__(16) float As[2][FBK][FAS];
  __shared__ __align__(16) float Bs[2][FBK][FBS];
  const int t = threadIdx.x;
  const long long M = (long long)d.batch * d.h * d.w;
  const long long m0 = (long long)blockIdx.x * FBM;
  const int n0 = blockIdx.y * FBN;
  const int taps = d.kh * d.kw;
  const int cchunks = (d.cin + FBK - 1) / FBK;
  const int nk = taps * cchunks;
  const int hw = d.h * d.w;

  // loader mapping
  const int kq = t & 3, lrow = t >> 2;
  RowInfo ri[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    long long p = m0 + lrow + i * 64;
    ri[i].valid = p < M;
    long long b = ri[i].valid ? p / hw : 0;
    int rem = ri[i].valid ? (int)(p - b * hw) : 0;
    ri[i].oh = rem / d.w;
    ri[i].ow = rem - ri[i].oh * d.w;
    ri[i].base = b * hw;
  }
  const bool xvec = (d.x_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.x) & 15) == 0);
  const bool wvec = (d.cin % 4 == 0);
  const int wn = n0 + lrow;  // weight row this thread loads

  float4 ra[2], rb;
  auto gload = [&](int kk) {
    const int tap = kk / cchunks;
    const int c0 = (kk - tap * cchunks) * FBK + kq * 4;
    const int r = tap / d.kw, s = tap - r * d.kw;
    const int nv = d.cin - c0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ih = ri[i].oh + r - d.pad, iw = ri[i].ow + s - d.pad;
      const bool ok = ri[i].valid && ih >= 0 && ih < d.h && iw >= 0 && iw < d.w;
      const float* p = ok ? d.x + (ri[i].base + (long long)ih * d.w + iw) * d.x_ld + c0 : nullptr;
      ra[i] = ld4_guard(p, nv, xvec);
    }
    const float* pw = (wn < d.cout) ? wq + ((long long)wn * taps + tap) * d.cin + c0 : nullptr;
    rb = ld4_guard(pw, nv, wvec);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][kq * 4 + 0][lrow + i * 64] = ra[i].x;
      As[buf][kq * 4 + 1][lrow + i * 64] = ra[i].y;
      As[buf][kq * 4 + 2][lrow + i * 64] = ra[i].z;
      As[buf][kq * 4 + 3][lrow + i * 64] = ra[i].w;
    }
    Bs[buf][kq * 4 + 0][lrow] = rb.x;
    Bs[buf][kq * 4 + 1][lrow] = rb.y;
    Bs[buf][kq * 4 + 2][lrow] = rb.z;
    Bs[buf][kq * 4 + 3][lrow] = rb.w;
  };

  const int tm = t >> 4, tn = t & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kk = 0; kk < nk; ++kk) {
    const int buf = kk & 1;
    if (kk + 1 < nk) gload(kk + 1);
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kk + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue (order documented in include/neosr_b200.h)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long p = m0 + tm * 8 + i;
    if (p >= M) continue;
    const float rs = d.row_scale ? d.row_scale[p / hw] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n >= d.cout) continue;
      const long long o = p * d.y_ld + n;
      const long long ores = p * (d.res_ld ? d.res_ld : d.y_ld) + n, oaux = p * (d.aux_ld ? d.aux_ld : d.y_ld) + n;
      float v = acc[i][j] + (d.bias ? d.bias[n] : 0.f);
      if (d.y_pre) d.y_pre[o] = d.pre_mode ? act_grad(v, d.act, d.act == NSR_ACT_PRELU ? d.prelu[n] : d.act_slope) : v;
      if (d.act) v = apply_act(v, d.act, d.act == NSR_ACT_PRELU ? d.prelu[n] : d.act_slope);
      if (d.actgrad) v *= act_grad(d.aux[oaux], d.actgrad, d.actgrad == NSR_ACT_PRELU ? d.prelu[n] : d.actgrad_slope);
      if (d.row_scale) v *= rs;
      if (d.residual) v += d.residual[ores];
      d.y[o] = v;
    }
  }
}

int conv_fprop_simt(const NsrConv& d, cudaStream_t st) {
  const long long M = (long long)d.batch * d.h * d.w;
  dim3 grid((unsigned)ceil_div(M, FBM), (unsigned)ceil_div(d.cout, FBN));
  igemm_fprop_simt<<<grid, FTHREADS, 0, st>>>(d, reinterpret_cast<const float*>(d.w_packed));
  NSR_CHECK_LAUNCH("igemm_fprop_simt");
  return NSR_OK;
}

// ------------------------------------------------------------------------------ wgrad
constexpr int WBM = 64, WBN = 64, WBK = 16, WTHREADS = 256;

struct WgradPlan {
  int nt_co, nt_ci, taps, tiles, splitk;
  long long rows_per_split;
  int bias_blocks;
  size_t dw_partial_floats, bias_partial_floats;
};

static WgradPlan wgrad_plan(const NsrWgrad& d) {
  WgradPlan p;
  const long long M = (long long)d.batch * d.h * d.w;
  p.nt_co = ceil_div(d.cout, WBM);
  p.nt_ci = ceil_div(d.cin, WBN);
  p.taps = d.kh * d.kw;
  p.tiles = p.nt_co * p.nt_ci * p.taps;
  int want = ceil_div(kNumSMs * 4, p.tiles);
  int maxs = (int)((M + 511) / 512);
  p.splitk = want < 1 ? 1 : want;
  if (p.splitk > maxs) p.splitk = maxs;
  if (p.splitk < 1) p.splitk = 1;
  long long rps = (M + p.splitk - 1) / p.splitk;
  p.rows_per_split = (rps + WBK - 1) / WBK * WBK;
  p.splitk = (int)((M + p.rows_per_split - 1) / p.rows_per_split);
  p.dw_partial_floats = (size_t)p.splitk * d.cout * p.taps * d.cin;
  p.bias_blocks = bias_grad_blocks(M);
  p.bias_partial_floats = (size_t)p.bias_blocks * d.cout;
  return p;
}

__global__ void __launch_bounds__(WTHREADS) igemm_wgrad_simt(NsrWgrad d, WgradPlan pl, float* __restrict__ partial) {
  __shared__ __align__(16) float As[WBK][WBM + 4];  // dy chunk  [pixel][cout]
  __shared__ __align__(16) float Bs[WBK][WBN + 4];  // x  chunk  [pixel][cin]
  const int t = threadIdx.x;
  int tile = blockIdx.x;
  const int tap = tile % pl.taps;
  tile /= pl.taps;
  const int tci = tile % pl.nt_ci;
  const int tco = tile / pl.nt_ci;
  const int r = tap / d.kw, s = tap - r * d.kw;
  const int co0 = tco * WBM, ci0 = tci * WBN;
  const long long M = (long long)d.batch * d.h * d.w;
  const int hw = d.h * d.w;
  const long long p_begin = (long long)blockIdx.y * pl.rows_per_split;
  long long p_end = p_begin + pl.rows_per_split;
  if (p_end > M) p_end = M;

  const int px = t >> 4, q = t & 15;
  const bool yvec = (d.dy_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.dy) & 15) == 0);
  const bool xvec = (d.x_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.x) & 15) == 0);
  const int tm = t >> 4, tn = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pc = p_begin; pc < p_end; pc += WBK) {
    const long long p = pc + px;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (p < p_end) {
      va = ld4_guard(d.dy + p * d.dy_ld + co0 + q * 4, d.cout - (co0 + q * 4), yvec);
      const long long b = p / hw;
      const int rem = (int)(p - b * hw);
      const int oh = rem / d.w, ow = rem - oh * d.w;
      const int ih = oh + r - d.pad, iw = ow + s - d.pad;
      if (ih >= 0 && ih < d.h && iw >= 0 && iw < d.w)
        vb = ld4_guard(d.x + (b * hw + (long long)ih * d.w + iw) * d.x_ld + ci0 + q * 4, d.cin - (ci0 + q * 4), xvec);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[px][q * 4]) = va;
    *reinterpret_cast<float4*>(&Bs[px][q * 4]) = vb;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* out = partial + (size_t)blockIdx.y * d.cout * pl.taps * d.cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tm * 4 + i;
    if (co >= d.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tn * 4 + j;
      if (ci < d.cin) out[((size_t)co * pl.taps + tap) * d.cin + ci] = acc[i][j];
    }
  }
}

// partial[split][co][tap][ci] -> dw[co][ci][tap] (OIHW), fixed summation order.
__global__ void wgrad_reduce(const float* __restrict__ partial, float* __restrict__ dw, int splitk, int cout,
                             int taps, int cin) {
  const size_t n = (size_t)cout * taps * cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splitk; ++k) s += partial[(size_t)k * n + i];
    const int ci = (int)(i % cin);
    const size_t r = i / cin;
    const int tap = (int)(r % taps);
    const int co = (int)(r / taps);
    dw[((size_t)co * cin + ci) * taps + tap] = s;
  }
}

// column sums of dy[M, cout] (row stride ld): phase 1 -> partial[blockIdx.y][cout].
// blockDim = (64 column quads, 4 row lanes); each thread keeps 8 independent 16-byte loads in
// flight (rows r, r+4, ...), so the pass streams at HBM rate instead of one load per round trip.
__global__ void __launch_bounds__(256) colsum_partial(const float* __restrict__ dy, float* __restrict__ partial,
                                                      long long M, int cout, int ld, int vec) {
  __shared__ float4 red[4][64];
  const int cq = blockIdx.x * 64 + threadIdx.x;       // column quad
  const int c = cq * 4;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rows_per;
  long long r1 = r0 + rows_per;
  if (r1 > M) r1 = M;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cout) {
    long long r = r0 + threadIdx.y;
    if (vec && c + 3 < cout) {
      for (; r + 28 < r1; r += 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(dy + (r + 4 * u) * ld + c));
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      for (; r < r1; r += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dy + r * ld + c));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    } else {
      for (; r < r1; r += 4) {
        const float* p = dy + r * ld + c;
        acc.x += p[0];
        if (c + 1 < cout) acc.y += p[1];
        if (c + 2 < cout) acc.z += p[2];
        if (c + 3 < cout) acc.w += p[3];
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cout) {
    float4 s = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float4 o = red[k][threadIdx.x];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    float* out = partial + (size_t)blockIdx.y * cout + c;
    out[0] = s.x;
    if (c + 1 < cout) out[1] = s.y;
    if (c + 2 < cout) out[2] = s.z;
    if (c + 3 < cout) out[3] = s.w;
  }
}
// same reduction when dy only exists as a split tile image: thread = (16-byte chunk, row lane)
__global__ void __launch_bounds__(256) colsum_sti_partial(const uint8_t* __restrict__ sti, float* __restrict__ partial,
                                                          long long M, int cout) {
  __shared__ float red[32][65];
  const int kbs = (cout + 63) / 64;
  const int kb = blockIdx.x;
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const long long rows_per = ((M + gridDim.y - 1) / gridDim.y + 127) / 128 * 128;
  const long long r0 = (long long)blockIdx.y * rows_per;
  long long r1 = r0 + rows_per;
  if (r1 > M) r1 = M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long p = r0 + rl; p < r1; p += 32) {
    const int r = (int)(p & 127);
    const uint8_t* src = sti + ((size_t)((p >> 7) * kbs + kb) << 15) + r * 128 + ((cl ^ (r & 7)) << 4);
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(src + 16384));
    const uint32_t hv[4] = {h.x, h.y, h.z, h.w}, lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] += __uint_as_float(hv[e] << 16) + __uint_as_float(lv[e] << 16);
      acc[2 * e + 1] += __uint_as_float(hv[e] & 0xFFFF0000u) + __uint_as_float(lv[e] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cl * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float sum = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) sum += red[k][threadIdx.x];
    const int c = kb * 64 + threadIdx.x;
    if (c < cout) partial[(size_t)blockIdx.y * cout + c] = sum;
  }
}

// one CTA per 32 columns: 32 row groups of coalesced 128-byte loads, then a fixed-order reduction over the groups
// (a thread per column walking all <= 592 partial rows took 32 us per call at 2 M pixels)
__global__ void __launch_bounds__(1024) colsum_final(const float* __restrict__ partial, float* __restrict__ out, int blocks, int cout) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cout)
    for (int b = ty; b < blocks; b += 32) s += partial[(size_t)b * cout + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cout) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][tx];
    out[c] = t;
  }
}

int launch_wgrad_reduce(const float* partial, float* dw, int splitk, int cout, int taps, int cin, cudaStream_t st) {
  const size_t n = (size_t)cout * taps * cin;
  int blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  wgrad_reduce<<<blocks, 256, 0, st>>>(partial, dw, splitk, cout, taps, cin);
  NSR_CHECK_LAUNCH("wgrad_reduce");
  return NSR_OK;
}

size_t conv_wgrad_workspace_simt(const NsrWgrad& d) {
  WgradPlan p = wgrad_plan(d);
  return (p.dw_partial_floats + p.bias_partial_floats) * sizeof(float);
}

int conv_bias_grad(const NsrWgrad& d, float* bias_partial, int bias_blocks, cudaStream_t st) {
  const long long M = (long long)d.batch * d.h * d.w;
  if (d.dy == nullptr) {  // only the split tile image of dy exists
    if (bias_blocks > (int)((M + 127) / 128)) bias_blocks = (int)((M + 127) / 128);
    dim3 grid((unsigned)((d.cout + 63) / 64), (unsigned)bias_blocks);
    colsum_sti_partial<<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(d.dy_sti), bias_partial, M, d.cout);
    NSR_CHECK_LAUNCH("colsum_sti_partial");
  } else {
    const int vec = (d.dy_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.dy) & 15) == 0);
    dim3 grid((unsigned)ceil_div(d.cout, 256), (unsigned)bias_blocks), block(64, 4);
    colsum_partial<<<grid, block, 0, st>>>(d.dy, bias_partial, M, d.cout, d.dy_ld, vec);
    NSR_CHECK_LAUNCH("colsum_partial");
  }
  colsum_final<<<ceil_div(d.cout, 32), 1024, 0, st>>>(bias_partial, d.dbias, bias_blocks, d.cout);
  NSR_CHECK_LAUNCH("colsum_final");
  return NSR_OK;
}

int conv_wgrad_simt(const NsrWgrad& d, cudaStream_t st) {
  WgradPlan p = wgrad_plan(d);
  const size_t need = (p.dw_partial_floats + p.bias_partial_floats) * sizeof(float);
  if (d.workspace_bytes < need || d.workspace == nullptr) {
    set_error("nsr_conv_wgrad: workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(d.workspace);
  dim3 grid((unsigned)p.tiles, (unsigned)p.splitk);
  igemm_wgrad_simt<<<grid, WTHREADS, 0, st>>>(d, p, partial);
  NSR_CHECK_LAUNCH("igemm_wgrad_simt");
  int rc = launch_wgrad_reduce(partial, d.dw, p.splitk, d.cout, p.taps, d.cin, st);
  if (rc) return rc;
  if (d.dbias) return conv_bias_grad(d, partial + p.dw_partial_floats, p.bias_blocks, st);
  return NSR_OK;
}

}  // namespace nsr

// On-the-fly degradation pipeline (Real-ESRGAN second-order model) as HBM-bound sm_100a kernels.
// Replaces the ATen call chains of neosr/models/otf.py:92-283 (reference) and its helpers:
//   filter2D            neosr/utils/diffjpeg.py:558-584
//   F.interpolate       neosr/models/otf.py:126,179-186,222-226,243-247  (area | bilinear | bicubic)
//   gaussian / poisson  neosr/data/degradations.py:569-605,665-676,738-786,851-862
//   DiffJPEG            neosr/utils/diffjpeg.py:254-291,461-508,531-555
//   quantise + crop     neosr/models/otf.py:251, neosr/data/transforms.py:38-131
//   training-pair pool  neosr/models/otf.py:37-90
// All images are NCHW fp32 planes (the layout the reference's data pipeline hands to feed_data).
// Every stage reads its input once and writes its output once; per-sample parameters (sigma,
// gray flag, JPEG quality, blur kernel) live in small device arrays, so no stage ever syncs the
// host (the reference syncs 2·B+ times per batch: torch.unique loops and quality_to_factor).
#include "common.cuh"

namespace nsr {

// ------------------------------------------------------------------ Philox4x32-10 -----------
struct Philox {
  uint32_t c[4], k[2];
  uint32_t out[4];
  int have;
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t subsequence, uint64_t offset) {
    k[0] = (uint32_t)seed;
    k[1] = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)offset;
    c[1] = (uint32_t)(offset >> 32);
    c[2] = (uint32_t)subsequence;
    c[3] = (uint32_t)(subsequence >> 32);
    have = 0;
  }
  __device__ __forceinline__ void round4(uint32_t* ctr, const uint32_t* key) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ key[0], n1 = lo1, n2 = hi0 ^ ctr[3] ^ key[1], n3 = lo0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
  }
  __device__ __forceinline__ void refill() {
    uint32_t ctr[4] = {c[0], c[1], c[2], c[3]};
    uint32_t key[2] = {k[0], k[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round4(ctr, key);
      key[0] += 0x9E3779B9u;
      key[1] += 0xBB67AE85u;
    }
    out[0] = ctr[0]; out[1] = ctr[1]; out[2] = ctr[2]; out[3] = ctr[3];
    if (++c[0] == 0) ++c[1];
    have = 4;
  }
  __device__ __forceinline__ uint32_t next() {
    if (have == 0) refill();
    return out[4 - have--];
  }
  // uniform in (0, 1]: never 0, so log() is finite
  __device__ __forceinline__ float uniform() { return ((float)(next() >> 8) + 1.0f) * (1.0f / 16777216.0f); }
  __device__ __forceinline__ void normal2(float& a, float& b) {
    const float u1 = uniform(), u2 = uniform();
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c2;
    sincospif(2.0f * u2, &s, &c2);
    a = r * c2;
    b = r * s;
  }
};

// Poisson(lam) sample.  lam < 10: product-of-uniforms; else Hörmann's PTRS transformed rejection
// (the algorithm numpy and torch's CPU path use).  lam <= 256 here (8-bit levels × vals <= 256).
__device__ float poisson_sample(Philox& g, float lam) {
  if (!(lam > 0.f)) return 0.f;
  if (lam < 10.f) {
    const float enlam = expf(-lam);
    float prod = 1.f;
    int x = 0;
    for (;;) {
      prod *= g.uniform();
      if (prod > enlam) ++x; else return (float)x;
      if (x > 1000) return (float)x;
    }
  }
  const float slam = sqrtf(lam), loglam = logf(lam);
  const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.f);
  for (int it = 0; it < 1000; ++it) {
    const float U = g.uniform() - 0.5f, V = g.uniform();
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.f * a / us + b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= vr) return k;
    if (k < 0.f || (us < 0.013f && V > us)) continue;
    if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -lam + k * loglam - lgammaf(k + 1.f)) return k;
  }
  return floorf(lam);
}

// ------------------------------------------------------------------ filter2D ----------------
constexpr int F2D_TILE = 32, F2D_MAXK = 21;
__device__ __forceinline__ int reflect_idx(int i, int n) {  // F.pad(mode="reflect"): no edge repeat
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__global__ void __launch_bounds__(256) filter2d_kernel(const float* __restrict__ img, const float* __restrict__ kern,
                                                       float* __restrict__ out, int C, int H, int W, int k,
                                                       int kern_batch_stride) {
  __shared__ float tile[(F2D_TILE + F2D_MAXK - 1) * (F2D_TILE + F2D_MAXK - 1 + 1)];
  __shared__ float kw[F2D_MAXK * F2D_MAXK];
  const int plane = blockIdx.z;  // b*C + c
  const int b = plane / C;
  const int x0 = blockIdx.x * F2D_TILE, y0 = blockIdx.y * F2D_TILE;
  const int r = k / 2, span = F2D_TILE + k - 1, pitch = span + 1;
  const float* src = img + (size_t)plane * H * W;
  for (int i = threadIdx.x; i < k * k; i += 256) kw[i] = kern[(size_t)b * kern_batch_stride + i];
  for (int i = threadIdx.x; i < span * span; i += 256) {
    const int ty = i / span, tx = i - ty * span;
    const int gy = reflect_idx(y0 + ty - r, H), gx = reflect_idx(x0 + tx - r, W);
    tile[ty * pitch + tx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? src[(size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j < k; ++j) {
      const float wv = kw[i * k + j];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(tile[(ty0 + 8 * q + i) * pitch + tx + j], wv, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int y = y0 + ty0 + 8 * q, x = x0 + tx;
    if (y < H && x < W) out[(size_t)plane * H * W + (size_t)y * W + x] = acc[q];
  }
}

// ------------------------------------------------------------------ resize ------------------
// torch semantics, align_corners=False, no antialias.  rh / rw = the coordinate scale torch uses
// (1/scale_factor when a scale_factor was given, in/out otherwise) — computed by the host.
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  float x = t + 1.f;
  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}
template <int MODE>  // 0 area (adaptive average), 1 bilinear, 2 bicubic
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                     int H, int W, int OH, int OW, float rh, float rw) {
  const size_t total = (size_t)planes * OH * OW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const size_t p = idx / ((size_t)OW * OH);
    const float* src = in + p * H * W;
    float v;
    if (MODE == 0) {
      const int ys = (int)(((long long)oy * H) / OH), ye = (int)((((long long)oy + 1) * H + OH - 1) / OH);
      const int xs = (int)(((long long)ox * W) / OW), xe = (int)((((long long)ox + 1) * W + OW - 1) / OW);
      float s = 0.f;
      for (int y = ys; y < ye; ++y)
        for (int x = xs; x < xe; ++x) s += src[(size_t)y * W + x];
      v = s / (float)((ye - ys) * (xe - xs));
    } else if (MODE == 1) {
      const float sy = fmaxf(rh * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * ((float)ox + 0.5f) - 0.5f, 0.f);
      const int y1 = min((int)sy, H - 1), x1 = min((int)sx, W - 1);
      const int yp = y1 < H - 1 ? 1 : 0, xp = x1 < W - 1 ? 1 : 0;
      const float ly = sy - (float)y1, lx = sx - (float)x1;
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float* r0 = src + (size_t)y1 * W;
      const float* r1 = src + (size_t)(y1 + yp) * W;
      v = hy * (hx * r0[x1] + lx * r0[x1 + xp]) + ly * (hx * r1[x1] + lx * r1[x1 + xp]);
    } else {
      const float sy = rh * ((float)oy + 0.5f) - 0.5f, sx = rw * ((float)ox + 0.5f) - 0.5f;
      const float fy = floorf(sy), fx = floorf(sx);
      const int iy = (int)fy, ix = (int)fx;
      float cy[4], cx[4];
      cubic_coeffs(sy - fy, cy);
      cubic_coeffs(sx - fx, cx);
      v = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* row = src + (size_t)min(max(iy - 1 + i, 0), H - 1) * W;
        float rv = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) rv += row[min(max(ix - 1 + j, 0), W - 1)] * cx[j];
        v += rv * cy[i];
      }
    }
    out[idx] = v;
  }
}

// ------------------------------------------------------------------ noise -------------------
// out = clamp(img + noise, 0, 1); noise as degradations.py:569-605 (gaussian).  z / zg: optional
// caller-provided standard-normal fields ([B,3,H,W] and the batch-shared gray field [H,W]);
// NULL => drawn in-kernel from Philox(seed).
__global__ void __launch_bounds__(256) gaussian_noise_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                             const float* __restrict__ sigma,
                                                             const float* __restrict__ gray, int any_gray,
                                                             const float* __restrict__ z, const float* __restrict__ zg,
                                                             int B, int H, int W, uint64_t seed) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / hw);
    const size_t pix = idx - (size_t)b * hw;
    const float sg = sigma[b], g = gray[b];
    float n[4], ng = 0.f;
    if (z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) n[c] = z[((size_t)b * 3 + c) * hw + pix];
      if (any_gray) ng = zg[pix];
    } else {
      Philox ph(seed, idx, 0);
      ph.normal2(n[0], n[1]);
      ph.normal2(n[2], n[3]);
      if (any_gray) {  // ONE [H,W] field shared by the whole batch (degradations.py:593-598)
        Philox pg(seed ^ 0x9E3779B97F4A7C15ull, pix, 1);
        float t;
        pg.normal2(ng, t);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // every product and sum rounded separately, as the reference's chain of ATen ops does (no FMA contraction)
      float nz = __fmul_rn(n[c], sg) / 255.0f;
      if (any_gray) nz = __fadd_rn(__fmul_rn(nz, 1.f - g), __fmul_rn(__fmul_rn(ng, sg) / 255.0f, g));
      const size_t o = ((size_t)b * 3 + c) * hw + pix;
      out[o] = fminf(fmaxf(__fadd_rn(img[o], nz), 0.f), 1.f);
    }
  }
}

__device__ __forceinline__ float quant8(float v) { return fminf(fmaxf(rintf(v * 255.0f), 0.f), 255.f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) {  // torchvision rgb_to_grayscale, no FMA
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
// 256-bin presence bitmaps per sample: words [0,8) colour image (all 3 channels), [8,16) gray image
// (degradations.py:762-768,775-779 count distinct 8-bit levels with torch.unique in a Python loop).
__global__ void __launch_bounds__(256) level_bitmap_kernel(const float* __restrict__ img, uint32_t* __restrict__ bitmap,
                                                           int H, int W) {
  __shared__ uint32_t bm[16];
  if (threadIdx.x < 16) bm[threadIdx.x] = 0u;
  __syncthreads();
  const int b = blockIdx.y;
  const size_t hw = (size_t)H * W;
  const float* base = img + (size_t)b * 3 * hw;
  for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < hw; pix += (size_t)gridDim.x * blockDim.x) {
    const float r = base[pix], g = base[hw + pix], bl = base[2 * hw + pix];
    const int lr = (int)quant8(r), lg = (int)quant8(g), lb = (int)quant8(bl), ly = (int)quant8(gray_of(r, g, bl));
    atomicOr(&bm[lr >> 5], 1u << (lr & 31));
    atomicOr(&bm[lg >> 5], 1u << (lg & 31));
    atomicOr(&bm[lb >> 5], 1u << (lb & 31));
    atomicOr(&bm[8 + (ly >> 5)], 1u << (ly & 31));
  }
  __syncthreads();
  if (threadIdx.x < 16 && bm[threadIdx.x]) atomicOr(&bitmap[b * 16 + threadIdx.x], bm[threadIdx.x]);
}
__device__ __forceinline__ float vals_from_bitmap(const uint32_t* bm) {  // 2^ceil(log2(#levels))
  int n = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) n += __popc(bm[i]);
  int v = 1;
  while (v < n) v <<= 1;
  return (float)v;
}
// Poisson (shot) noise, degradations.py:738-786.  counts_c / counts_g: optional caller-provided
// Poisson draws for lambda = quantised image × vals ([B,3,H,W] / [B,1,H,W]); NULL => Philox PTRS.
__global__ void __launch_bounds__(256) poisson_noise_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ gray, int any_gray,
                                                            const uint32_t* __restrict__ bitmap,
                                                            const float* __restrict__ counts_c,
                                                            const float* __restrict__ counts_g, int B, int H, int W,
                                                            uint64_t seed) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / hw);
    const size_t pix = idx - (size_t)b * hw;
    const float vc = vals_from_bitmap(bitmap + b * 16), vg = vals_from_bitmap(bitmap + b * 16 + 8);
    const float sc = scale[b], g = gray[b];
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = img[((size_t)b * 3 + c) * hw + pix];
    Philox ph(seed, idx, 2);
    float ng = 0.f;
    if (any_gray) {
      const float q = quant8(gray_of(v[0], v[1], v[2])) / 255.0f;
      const float cnt = counts_g ? counts_g[idx] : poisson_sample(ph, __fmul_rn(q, vg));
      ng = __fsub_rn(cnt / vg, q);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)b * 3 + c) * hw + pix;
      const float q = quant8(v[c]) / 255.0f;
      const float cnt = counts_c ? counts_c[o] : poisson_sample(ph, __fmul_rn(q, vc));
      float nz = __fsub_rn(cnt / vc, q);  // rounded op by op, as the reference's ATen chain (no FMA contraction)
      if (any_gray) nz = __fadd_rn(__fmul_rn(nz, 1.f - g), __fmul_rn(ng, g));
      out[o] = fminf(fmaxf(__fadd_rn(v[c], __fmul_rn(nz, sc)), 0.f), 1.f);
    }
  }
}

// ------------------------------------------------------------------ JPEG --------------------
// One CTA per 16x16 MCU: RGB*255 -> YCbCr -> 2x2 chroma mean -> six 8x8 blocks: -128, DCT-II,
// divide by (table*factor), round-half-even, multiply back, IDCT, +128 -> chroma nearest x2 ->
// RGB -> clamp -> /255.  12 B/pixel in, 12 B/pixel out; everything else stays in shared memory.
__constant__ float c_ytab[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                                 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
__constant__ float c_ctab4[16] = {17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99};
__global__ void __launch_bounds__(256) jpeg_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                   const float* __restrict__ quality, int H, int W) {
  __shared__ float blk[6][8][9];   // [Y00,Y01,Y10,Y11,Cb,Cr][row][col]
  __shared__ float tmp[6][8][9];
  __shared__ float cb_full[16][17], cr_full[16][17];
  __shared__ float cosm[8][8];     // cos((2x+1) u pi / 16), [x][u]
  const int b = blockIdx.z, t = threadIdx.x;
  const int lx = t & 15, ly = t >> 4;
  const int gx = blockIdx.x * 16 + lx, gy = blockIdx.y * 16 + ly;
  const size_t hw = (size_t)H * W;
  const bool inside = gx < W && gy < H;
  if (t < 64) cosm[t >> 3][t & 7] = cospif((float)((2 * (t >> 3) + 1) * (t & 7)) / 16.0f);
  float r = 0.f, g = 0.f, bl = 0.f;  // zero padding to a multiple of 16 (diffjpeg.py:545-551)
  if (inside) {
    const float* p = img + (size_t)b * 3 * hw + (size_t)gy * W + gx;
    // torch.clamp(out, 0, 1) precedes every jpeger call (otf.py:154,232,239): fused into the load
    r = fminf(fmaxf(p[0], 0.f), 1.f) * 255.0f;
    g = fminf(fmaxf(p[hw], 0.f), 1.f) * 255.0f;
    bl = fminf(fmaxf(p[2 * hw], 0.f), 1.f) * 255.0f;
  }
  const float yv = r * 0.299f + g * 0.587f + bl * 0.114f;
  cb_full[ly][lx] = r * -0.168736f + g * -0.331264f + bl * 0.5f + 128.0f;
  cr_full[ly][lx] = r * 0.5f + g * -0.418688f + bl * -0.081312f + 128.0f;
  blk[(ly >> 3) * 2 + (lx >> 3)][ly & 7][lx & 7] = yv - 128.0f;
  __syncthreads();
  if (t < 128) {
    const int ch = t >> 6, yy = (t >> 3) & 7, xx = t & 7;
    float (*src)[17] = ch ? cr_full : cb_full;
    const float m = (src[2 * yy][2 * xx] + src[2 * yy][2 * xx + 1] + src[2 * yy + 1][2 * xx] + src[2 * yy + 1][2 * xx + 1]) * 0.25f;
    blk[4 + ch][yy][xx] = m - 128.0f;
  }
  __syncthreads();
  float q = quality[b];
  const float factor = (q < 50.f ? 5000.0f / q : 200.0f - q * 2.f) / 100.0f;
  // forward DCT, separable: tmp[u][y] = sum_x blk[x][y] cos[x][u];  X[u][v] = sum_y tmp[u][y] cos[y][v]
  for (int e = t; e < 384; e += 256) {
    const int k = e >> 6, u = (e >> 3) & 7, y = e & 7;
    float s = 0.f;
#pragma unroll
    for (int x = 0; x < 8; ++x) s = fmaf(blk[k][x][y], cosm[x][u], s);
    tmp[k][u][y] = s;
  }
  __syncthreads();
  for (int e = t; e < 384; e += 256) {
    const int k = e >> 6, u = (e >> 3) & 7, v = e & 7;
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s = fmaf(tmp[k][u][y], cosm[y][v], s);
    const float au = u == 0 ? 0.70710678118654752440f : 1.f, av = v == 0 ? 0.70710678118654752440f : 1.f;
    const float coef = s * (au * av * 0.25f);
    // tables are the TRANSPOSED standard ones (diffjpeg.py:16-38): table[u][v] = std[v][u]
    const float tab = (k < 4 ? c_ytab[v * 8 + u] : ((u < 4 && v < 4) ? c_ctab4[v * 4 + u] : 99.f)) * factor;
    const float deq = rintf(coef / tab) * tab;
    blk[k][u][v] = deq * (au * av);  // iDCT8x8: image *= alpha
  }
  __syncthreads();
  // inverse: out[u][v] = 0.25 sum_{x,y} blk[x][y] cos((2u+1)x pi/16) cos((2v+1)y pi/16) + 128
  for (int e = t; e < 384; e += 256) {
    const int k = e >> 6, u = (e >> 3) & 7, y = e & 7;
    float s = 0.f;
#pragma unroll
    for (int x = 0; x < 8; ++x) s = fmaf(blk[k][x][y], cosm[u][x], s);
    tmp[k][u][y] = s;
  }
  __syncthreads();
  for (int e = t; e < 384; e += 256) {
    const int k = e >> 6, u = (e >> 3) & 7, v = e & 7;
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s = fmaf(tmp[k][u][y], cosm[v][y], s);
    blk[k][u][v] = 0.25f * s + 128.0f;
  }
  __syncthreads();
  if (inside) {
    const float Y = blk[(ly >> 3) * 2 + (lx >> 3)][ly & 7][lx & 7];
    const float Cb = blk[4][ly >> 1][lx >> 1] - 128.0f, Cr = blk[5][ly >> 1][lx >> 1] - 128.0f;
    const float R = Y + Cr * 1.402f;
    const float G = Y + Cb * -0.344136f + Cr * -0.714136f;
    const float Bv = Y + Cb * 1.772f;
    float* o = out + (size_t)b * 3 * hw + (size_t)gy * W + gx;
    o[0] = fminf(255.f, fmaxf(0.f, R)) / 255.0f;
    o[hw] = fminf(255.f, fmaxf(0.f, G)) / 255.0f;
    o[2 * hw] = fminf(255.f, fmaxf(0.f, Bv)) / 255.0f;
  }
}

// ------------------------------------------------------------------ crop / quantise / pool --
__global__ void __launch_bounds__(256) crop_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                   int H, int W, int top, int left, int ph, int pw, int quantise) {
  const size_t total = (size_t)planes * ph * pw;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % pw), y = (int)((idx / pw) % ph);
    const size_t p = idx / ((size_t)pw * ph);
    float v = in[p * H * W + (size_t)(top + y) * W + left + x];
    if (quantise) v = quant8(v) / 255.0f;  // otf.py:251
    out[idx] = v;
  }
}
// out[i] = pool[slot[i]];  pool[slot[i]] = in[i]   (element-wise, so in == out is NOT allowed)
__global__ void __launch_bounds__(256) pool_swap_kernel(float* __restrict__ pool, const float* __restrict__ in,
                                                        float* __restrict__ out, const int32_t* __restrict__ slots,
                                                        int b, size_t sample_elems, int do_dequeue) {
  const size_t total = (size_t)b * sample_elems;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / sample_elems);
    const size_t e = idx - (size_t)i * sample_elems;
    float* slot = pool + (size_t)slots[i] * sample_elems + e;
    const float nv = in[idx];
    if (do_dequeue) out[idx] = *slot;
    *slot = nv;
  }
}

static inline int grid_for(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
}  // namespace nsr
using namespace nsr;

extern "C" int nsr_filter2d(const float* img, const float* kernel, float* out, int batch, int channels, int h, int w,
                            int k, int kernel_batch, void* stream) {
  NSR_CHECK_ARG(img && kernel && out && img != out, "nsr_filter2d: null or aliased buffers");
  NSR_CHECK_ARG(batch > 0 && channels > 0 && h > 0 && w > 0, "nsr_filter2d: bad shape");
  NSR_CHECK_ARG(k % 2 == 1 && k >= 1 && k <= F2D_MAXK, "nsr_filter2d: Wrong kernel size %d (odd, <= %d)", k, F2D_MAXK);
  NSR_CHECK_ARG(k / 2 < h && k / 2 < w, "nsr_filter2d: reflect padding %d needs an image larger than that", k / 2);
  NSR_CHECK_ARG(kernel_batch == 1 || kernel_batch == batch, "nsr_filter2d: kernel batch %d vs image batch %d", kernel_batch, batch);
  NSR_CHECK_ARG((long long)batch * channels <= 65535, "nsr_filter2d: too many planes");
  dim3 grid(ceil_div(w, F2D_TILE), ceil_div(h, F2D_TILE), batch * channels);
  filter2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, kernel, out, channels, h, w, k, kernel_batch == 1 ? 0 : k * k);
  NSR_CHECK_LAUNCH("nsr_filter2d");
  return NSR_OK;
}

extern "C" int nsr_resize(const float* in, float* out, int planes, int h, int w, int oh, int ow, int mode,
                          float coord_scale_h, float coord_scale_w, void* stream) {
  NSR_CHECK_ARG(in && out && planes > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "nsr_resize: bad arguments");
  NSR_CHECK_ARG(mode >= 0 && mode <= 2, "nsr_resize: mode %d (0 area, 1 bilinear, 2 bicubic)", mode);
  const int grid = grid_for((size_t)planes * oh * ow);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) resize_kernel<0><<<grid, 256, 0, st>>>(in, out, planes, h, w, oh, ow, coord_scale_h, coord_scale_w);
  else if (mode == 1) resize_kernel<1><<<grid, 256, 0, st>>>(in, out, planes, h, w, oh, ow, coord_scale_h, coord_scale_w);
  else resize_kernel<2><<<grid, 256, 0, st>>>(in, out, planes, h, w, oh, ow, coord_scale_h, coord_scale_w);
  NSR_CHECK_LAUNCH("nsr_resize");
  return NSR_OK;
}

extern "C" int nsr_gaussian_noise(const float* img, float* out, const float* sigma, const float* gray, int any_gray,
                                  const float* z, const float* z_gray, int batch, int h, int w, uint64_t seed,
                                  void* stream) {
  NSR_CHECK_ARG(img && out && sigma && gray && batch > 0 && h > 0 && w > 0, "nsr_gaussian_noise: bad arguments");
  NSR_CHECK_ARG(!z || !any_gray || z_gray, "nsr_gaussian_noise: z given without z_gray");
  gaussian_noise_kernel<<<grid_for((size_t)batch * h * w), 256, 0, (cudaStream_t)stream>>>(
      img, out, sigma, gray, any_gray, z, z_gray, batch, h, w, seed);
  NSR_CHECK_LAUNCH("nsr_gaussian_noise");
  return NSR_OK;
}

extern "C" size_t nsr_poisson_noise_workspace(int batch) { return (size_t)batch * 16 * sizeof(uint32_t); }
extern "C" int nsr_poisson_noise(const float* img, float* out, const float* scale, const float* gray, int any_gray,
                                 const float* counts_color, const float* counts_gray, int batch, int h, int w,
                                 uint64_t seed, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(img && out && scale && gray && batch > 0 && h > 0 && w > 0 && batch <= 65535, "nsr_poisson_noise: bad arguments");
  if (!workspace || workspace_bytes < nsr_poisson_noise_workspace(batch)) {
    set_error("nsr_poisson_noise: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(workspace);
  cudaMemsetAsync(bitmap, 0, nsr_poisson_noise_workspace(batch), st);
  const size_t hw = (size_t)h * w;
  int bx = (int)((hw + 255) / 256);
  if (bx > 64) bx = 64;
  level_bitmap_kernel<<<dim3(bx, batch), 256, 0, st>>>(img, bitmap, h, w);
  NSR_CHECK_LAUNCH("nsr_poisson_noise(bitmap)");
  poisson_noise_kernel<<<grid_for((size_t)batch * hw), 256, 0, st>>>(img, out, scale, gray, any_gray, bitmap, counts_color,
                                                                    counts_gray, batch, h, w, seed);
  NSR_CHECK_LAUNCH("nsr_poisson_noise");
  return NSR_OK;
}

extern "C" int nsr_jpeg(const float* img, float* out, const float* quality, int batch, int h, int w, void* stream) {
  NSR_CHECK_ARG(img && out && quality && batch > 0 && h > 0 && w > 0 && batch <= 65535, "nsr_jpeg: bad arguments");
  dim3 grid(ceil_div(w, 16), ceil_div(h, 16), batch);
  jpeg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, out, quality, h, w);
  NSR_CHECK_LAUNCH("nsr_jpeg");
  return NSR_OK;
}

extern "C" int nsr_crop(const float* in, float* out, int planes, int h, int w, int top, int left, int ph, int pw,
                        int quantise, void* stream) {
  NSR_CHECK_ARG(in && out && planes > 0 && ph > 0 && pw > 0, "nsr_crop: bad arguments");
  NSR_CHECK_ARG(top >= 0 && left >= 0 && top + ph <= h && left + pw <= w, "nsr_crop: window (%d,%d,%d,%d) outside %dx%d", top, left, ph, pw, h, w);
  crop_kernel<<<grid_for((size_t)planes * ph * pw), 256, 0, (cudaStream_t)stream>>>(in, out, planes, h, w, top, left, ph, pw, quantise);
  NSR_CHECK_LAUNCH("nsr_crop");
  return NSR_OK;
}

extern "C" int nsr_pool_swap(float* pool, const float* in, float* out, const int32_t* slots, int b, size_t sample_elems,
                             int dequeue, void* stream) {
  NSR_CHECK_ARG(pool && in && slots && b > 0 && sample_elems > 0, "nsr_pool_swap: bad arguments");
  NSR_CHECK_ARG(!dequeue || (out && out != in), "nsr_pool_swap: dequeue needs a distinct output buffer");
  pool_swap_kernel<<<grid_for((size_t)b * sample_elems), 256, 0, (cudaStream_t)stream>>>(pool, in, out, slots, b, sample_elems, dequeue);
  NSR_CHECK_LAUNCH("nsr_pool_swap");
  return NSR_OK;
}

// ------------------------------------------------------------------ augmentations ----------
// apply_augment pieces (neosr/data/augmentations.py:14-310).  F.interpolate(..., antialias=True) for bilinear /
// bicubic (PIL-style: triangle / cubic a = -0.5 filters whose support widens with the down-scale factor; ATen
// upsample_{bilinear,bicubic}2d_aa), as one 2-D gather kernel writing into a window of the destination, with the
// clamp to [0,1] that follows every call site fused in.
namespace nsr {
__device__ __forceinline__ float aa_filter(float x, int cubic) {
  x = fabsf(x);
  if (!cubic) return x < 1.f ? 1.f - x : 0.f;
  const float a = -0.5f;
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}
struct AAAxis { int xmin, xsize; float center, invscale, total; };
__device__ __forceinline__ AAAxis aa_axis(int o, int in, float scale, int cubic) {
  AAAxis a;
  const float interp = cubic ? 4.f : 2.f;
  const float support = scale >= 1.f ? (interp * 0.5f) * scale : interp * 0.5f;
  a.center = scale * ((float)o + 0.5f);
  a.invscale = scale >= 1.f ? 1.f / scale : 1.f;
  a.xmin = max(0, (int)(a.center - support + 0.5f));
  a.xsize = min(in, (int)(a.center + support + 0.5f)) - a.xmin;
  a.total = 0.f;
  for (int k = 0; k < a.xsize; ++k) a.total += aa_filter(((float)(k + a.xmin) - a.center + 0.5f) * a.invscale, cubic);
  return a;
}
// dst[b, c, top + oy, left + ox] = clamp(resize_aa(src[perm ? perm[b] : b, c])[oy, ox], 0, 1)
__global__ void __launch_bounds__(256) resize_aa_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        const int32_t* __restrict__ perm, int B, int C, int H, int W, int OH,
                                                        int OW, int DH, int DW, int top, int left, float sh, float sw, int cubic) {
  const size_t total = (size_t)B * C * OH * OW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH);
    const int c = (int)((idx / ((size_t)OW * OH)) % C), b = (int)(idx / ((size_t)OW * OH * C));
    const int sb = perm ? perm[b] : b;
    const float* s = src + ((size_t)sb * C + c) * H * W;
    const AAAxis ay = aa_axis(oy, H, sh, cubic), ax = aa_axis(ox, W, sw, cubic);
    float acc = 0.f;
    for (int i = 0; i < ay.xsize; ++i) {
      const float wy = aa_filter(((float)(i + ay.xmin) - ay.center + 0.5f) * ay.invscale, cubic) / ay.total;
      float row = 0.f;
      for (int j = 0; j < ax.xsize; ++j)
        row += s[(size_t)(ay.xmin + i) * W + ax.xmin + j] * (aa_filter(((float)(j + ax.xmin) - ax.center + 0.5f) * ax.invscale, cubic) / ax.total);
      acc += wy * row;
    }
    dst[(((size_t)b * C + c) * DH + top + oy) * DW + left + ox] = fminf(fmaxf(acc, 0.f), 1.f);
  }
}
// mode 0 mixup: dst = lam*a + lam2*other[perm[b]], lam2 = fp32(1-lam)  (other = the GT batch for both GT and LQ, 29-31)
// mode 1 box copy: dst[box] = other[perm ? perm[b] : b][box]  (cutmix 58-59, cutblur 164); rest of dst = a
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ a, const float* __restrict__ other,
                                                  float* __restrict__ dst, const int32_t* __restrict__ perm, int B, size_t chw,
                                                  int H, int W, int mode, float lam, float lam2, int y0, int y1, int x0, int x1) {
  const size_t total = (size_t)B * chw;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / chw);
    const size_t e = idx - (size_t)b * chw;
    const int sb = perm ? perm[b] : b;
    if (mode == 0) {
      dst[idx] = __fadd_rn(__fmul_rn(lam, a[idx]), __fmul_rn(lam2, other[(size_t)sb * chw + e]));
    } else {
      const int x = (int)(e % W), y = (int)((e / W) % H);
      dst[idx] = (y >= y0 && y < y1 && x >= x0 && x < x1) ? other[(size_t)sb * chw + e] : a[idx];
    }
  }
}
}  // namespace nsr

extern "C" int nsr_resize_aa(const float* src, float* dst, const int32_t* perm, int batch, int channels, int h, int w, int oh,
                             int ow, int dst_h, int dst_w, int top, int left, int bicubic, float coord_scale_h,
                             float coord_scale_w, void* stream) {
  NSR_CHECK_ARG(src && dst && src != dst && batch > 0 && channels > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "nsr_resize_aa: bad arguments");
  NSR_CHECK_ARG(top >= 0 && left >= 0 && top + oh <= dst_h && left + ow <= dst_w, "nsr_resize_aa: window outside the destination");
  resize_aa_kernel<<<grid_for((size_t)batch * channels * oh * ow), 256, 0, (cudaStream_t)stream>>>(
      src, dst, perm, batch, channels, h, w, oh, ow, dst_h, dst_w, top, left, coord_scale_h, coord_scale_w, bicubic);
  NSR_CHECK_LAUNCH("nsr_resize_aa");
  return NSR_OK;
}
extern "C" int nsr_batch_mix(const float* a, const float* other, float* dst, const int32_t* perm, int batch, int channels, int h,
                             int w, int mode, float lam, float lam2, int y0, int y1, int x0, int x1, void* stream) {
  NSR_CHECK_ARG(a && other && dst && batch > 0 && channels > 0 && h > 0 && w > 0 && (mode == 0 || mode == 1), "nsr_batch_mix: bad arguments");
  NSR_CHECK_ARG(dst != other || !perm, "nsr_batch_mix: in-place on the permuted source");
  mix_kernel<<<grid_for((size_t)batch * channels * h * w), 256, 0, (cudaStream_t)stream>>>(a, other, dst, perm, batch,
                                                                                         (size_t)channels * h * w, h, w, mode, lam,
                                                                                         lam2, y0, y1, x0, x1);
  NSR_CHECK_LAUNCH("nsr_batch_mix");
  return NSR_OK;
}

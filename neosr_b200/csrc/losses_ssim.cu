// MS-SSIM loss (neosr/losses/ssim_loss.py:66-163) and consistency loss (neosr/losses/consistency_loss.py:
// 14-192) as fused value + gradient kernels over NCHW fp32 images.  HBM-bound: every pass reads each image
// once into a shared-memory tile and keeps the 11x11 / 21x21 window sums on chip; the only intermediates
// that touch HBM are the three per-pixel partial-derivative maps of a scale (12 B/px) and the luma planes.
// All reductions are two-pass and fixed-order (deterministic); data-dependent scalars (the MS-SSIM product
// rule coefficients, the `cosim < 1e-3` branch of consistency_loss.py:186-190) stay on the device.
#include "common.cuh"

namespace nsr {
constexpr int ST = 16;        // output tile edge
constexpr int SMAXW = 11;     // largest SSIM window supported
constexpr int STHREADS = 256;

__device__ __forceinline__ void block_sum2(float& a, float& b) {  // valid on thread 0
  __shared__ float red[2][STHREADS / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.f; b = 0.f;
#pragma unroll
    for (int i = 0; i < STHREADS / 32; ++i) { a += red[0][i]; b += red[1][i]; }
  }
}

// ---- one SSIM scale, forward: sums of cs and ssim (ssim_loss.py:146-163) + the partial-derivative maps
// P[0] = dT/dmu_x, P[1] = dT/dG(x^2), P[2] = dT/dG(xy), T = cs (use_ssim == 0) or ssim = l*cs.
__global__ void __launch_bounds__(STHREADS) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ win, int wsz, float C1, float C2,
                                                            int use_ssim, float* __restrict__ P, float* __restrict__ partial,
                                                            int planes, int H, int W) {
  __shared__ float tx[(ST + SMAXW - 1) * (ST + SMAXW)], ty[(ST + SMAXW - 1) * (ST + SMAXW)];
  __shared__ float wk[SMAXW * SMAXW];
  const int r = wsz / 2, span = ST + wsz - 1, pitch = span + 1;
  const int plane = blockIdx.z, x0 = blockIdx.x * ST, y0 = blockIdx.y * ST;
  const float* xs = x + (size_t)plane * H * W;
  const float* ys = y + (size_t)plane * H * W;
  for (int i = threadIdx.x; i < wsz * wsz; i += STHREADS) wk[i] = win[i];
  for (int i = threadIdx.x; i < span * span; i += STHREADS) {
    const int yy = i / span, xx = i - yy * span, gy = y0 + yy - r, gx = x0 + xx - r;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;  // conv2d zero padding (ssim_loss.py:57-64)
    tx[yy * pitch + xx] = in ? xs[(size_t)gy * W + gx] : 0.f;
    ty[yy * pitch + xx] = in ? ys[(size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  float mx = 0.f, my = 0.f, gxx = 0.f, gyy = 0.f, gxy = 0.f;
  for (int i = 0; i < wsz; ++i)
    for (int j = 0; j < wsz; ++j) {
      const float w = wk[i * wsz + j], a = tx[(ly + i) * pitch + lx + j], b = ty[(ly + i) * pitch + lx + j];
      mx = fmaf(w, a, mx);
      my = fmaf(w, b, my);
      gxx = fmaf(w, a * a, gxx);
      gyy = fmaf(w, b * b, gyy);
      gxy = fmaf(w, a * b, gxy);
    }
  float s_cs = 0.f, s_ssim = 0.f;
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx < W && gy < H) {
    const float A1 = 2.f * mx * my + C1, B1 = mx * mx + my * my + C1;
    const float A2 = 2.f * (gxy - mx * my) + C2, B2 = (gxx - mx * mx) + (gyy - my * my) + C2;
    const float l = A1 / B1, cs = A2 / B2;
    s_cs = cs;
    s_ssim = l * cs;
    if (P) {
      const float dcs_dmx = (2.f * mx * cs - 2.f * my) / B2, dcs_dgxx = -cs / B2, dcs_dgxy = 2.f / B2;
      float p0 = dcs_dmx, p1 = dcs_dgxx, p2 = dcs_dgxy;
      if (use_ssim) {
        const float dl_dmx = (2.f * my - 2.f * mx * l) / B1;
        p0 = cs * dl_dmx + l * dcs_dmx;
        p1 = l * dcs_dgxx;
        p2 = l * dcs_dgxy;
      }
      const size_t n = (size_t)planes * H * W, o = (size_t)plane * H * W + (size_t)gy * W + gx;
      P[o] = p0;
      P[n + o] = p1;
      P[2 * n + o] = p2;
    }
  }
  block_sum2(s_cs, s_ssim);
  if (threadIdx.x == 0) {
    const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partial[2 * blk] = s_cs;
    partial[2 * blk + 1] = s_ssim;
  }
}

// Fixed-order sum of one scale's partials -> sums[2*scale .. 2*scale+1]
__global__ void ssim_reduce_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ sums) {
  __shared__ double sa[STHREADS], sb[STHREADS];
  double a = 0., b = 0.;
  for (int i = threadIdx.x; i < nblocks; i += STHREADS) { a += partial[2 * i]; b += partial[2 * i + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = STHREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sums[0] = (float)sa[0]; sums[1] = (float)sb[0]; }
}

// loss = weight * (1 - prod_s m_s^{w_s}),  m_s = mean cs (s < last) / mean ssim (last)  (ssim_loss.py:131-144);
// coef[s] = dloss/d(sum_s) = -weight * w_s * prod / m_s / count_s.
__global__ void msssim_final_kernel(const float* __restrict__ sums, const float* __restrict__ counts, int nscales,
                                    float weight, float* __restrict__ coef, float* loss_value, float* loss_accum) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float wts[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
  float m[5], prod = 1.f;
  for (int s = 0; s < nscales; ++s) {
    m[s] = sums[2 * s + (s == nscales - 1 ? 1 : 0)] / counts[s];
    prod *= powf(m[s], wts[s]);
  }
  for (int s = 0; s < nscales; ++s) coef[s] = -weight * wts[s] * prod / m[s] / counts[s];
  const float v = weight * (1.f - prod);
  if (loss_value) *loss_value = v;
  if (loss_accum) *loss_accum += v;
}

// dx = coef * ( G(P0) + 2 x G(P1) + y G(P2) ) + 0.25 * dx_coarse[(i+ph)/2, (j+pw)/2]   (G zero-padded, symmetric
// window => self-adjoint; the last term is the backward of F.avg_pool2d(2, 2, padding), ssim_loss.py:141-143)
__global__ void __launch_bounds__(STHREADS) ssim_bwd_kernel(const float* __restrict__ P, const float* __restrict__ x,
                                                            const float* __restrict__ y, const float* __restrict__ win,
                                                            int wsz, const float* __restrict__ coef,
                                                            const float* __restrict__ dx_coarse, int ch, int cw, int ph,
                                                            int pw, float* __restrict__ dx, int planes, int H, int W) {
  __shared__ float t0[(ST + SMAXW - 1) * (ST + SMAXW)], t1[(ST + SMAXW - 1) * (ST + SMAXW)], t2[(ST + SMAXW - 1) * (ST + SMAXW)];
  __shared__ float wk[SMAXW * SMAXW];
  const int r = wsz / 2, span = ST + wsz - 1, pitch = span + 1;
  const int plane = blockIdx.z, x0 = blockIdx.x * ST, y0 = blockIdx.y * ST;
  const size_t n = (size_t)planes * H * W, pb = (size_t)plane * H * W;
  for (int i = threadIdx.x; i < wsz * wsz; i += STHREADS) wk[i] = win[(wsz * wsz - 1) - i];  // flipped: adjoint of a correlation
  for (int i = threadIdx.x; i < span * span; i += STHREADS) {
    const int yy = i / span, xx = i - yy * span, gy = y0 + yy - r, gx = x0 + xx - r;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const size_t o = pb + (size_t)gy * W + gx;
    t0[yy * pitch + xx] = in ? P[o] : 0.f;
    t1[yy * pitch + xx] = in ? P[n + o] : 0.f;
    t2[yy * pitch + xx] = in ? P[2 * n + o] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int i = 0; i < wsz; ++i)
    for (int j = 0; j < wsz; ++j) {
      const float w = wk[i * wsz + j];
      const int o = (ly + i) * pitch + lx + j;
      a0 = fmaf(w, t0[o], a0);
      a1 = fmaf(w, t1[o], a1);
      a2 = fmaf(w, t2[o], a2);
    }
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx < W && gy < H) {
    const size_t o = pb + (size_t)gy * W + gx;
    float g = coef[0] * (a0 + 2.f * x[o] * a1 + y[o] * a2);
    if (dx_coarse) g += 0.25f * dx_coarse[(size_t)plane * ch * cw + (size_t)((gy + ph) >> 1) * cw + ((gx + pw) >> 1)];
    dx[o] = g;
  }
}

// F.avg_pool2d(x, 2, 2, padding=(ph, pw)), count_include_pad=True (divisor always 4)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                       int H, int W, int OH, int OW, int ph, int pw) {
  const size_t total = (size_t)planes * OH * OW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH);
    const float* src = in + (idx / ((size_t)OW * OH)) * H * W;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        const int yy = 2 * oy - ph + dy, xx = 2 * ox - pw + dxx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += src[(size_t)yy * W + xx];
      }
    out[idx] = s * 0.25f;
  }
}

// ------------------------------------------------------------------ consistency loss --------
__device__ __forceinline__ float lin_rgb(float v) {  // consistency_loss.py:59-69
  return v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
}
__device__ __forceinline__ float lin_rgb_grad(float v) {
  return v <= 0.04045f ? 1.f / 12.92f : 2.4f / 1.055f * powf((v + 0.055f) / 1.055f, 1.4f);
}
__device__ __forceinline__ float signed_cbrt(float v) { return copysignf(powf(fabsf(v), 1.f / 3.f), v) * (v != 0.f); }
__device__ __forceinline__ float cbrt_grad(float v) {  // d/dv sign(v)|v|^(1/3); autograd gives 0 at v == 0
  return v == 0.f ? 0.f : powf(fabsf(v), -2.f / 3.f) / 3.f;
}
// CIE L* / 100 of a (blurred, clamped) RGB pixel, as written in consistency_loss.py:112-144
__device__ __forceinline__ float l_star(float r, float g, float b) {
  const float Y = lin_rgb(r) * 0.2126f + lin_rgb(g) * 0.7152f + lin_rgb(b) * 0.0722f;
  const float L = Y <= (216.f / 24389.f) ? Y * (Y * (24389.f / 27.f)) : signed_cbrt(Y) * 116.f - 16.f;
  return fminf(fmaxf(L / 100.f, 0.f), 1.f);
}
__device__ __forceinline__ void l_star_grad(float r, float g, float b, float up, float* d) {
  const float Y = lin_rgb(r) * 0.2126f + lin_rgb(g) * 0.7152f + lin_rgb(b) * 0.0722f;
  const bool low = Y <= (216.f / 24389.f);
  const float L = low ? Y * (Y * (24389.f / 27.f)) : signed_cbrt(Y) * 116.f - 16.f;
  const float v = L / 100.f;
  float dY = (v >= 0.f && v <= 1.f) ? up / 100.f : 0.f;
  dY *= low ? 2.f * Y * (24389.f / 27.f) : 116.f * cbrt_grad(Y);
  d[0] = dY * 0.2126f * lin_rgb_grad(r);
  d[1] = dY * 0.7152f * lin_rgb_grad(g);
  d[2] = dY * 0.0722f * lin_rgb_grad(b);
}
// Oklab chroma (a, b) + 0.5, clamped to [0,1] (consistency_loss.py:71-110,169-171); sat multiplies the target
__device__ __forceinline__ void oklab_ab(float r, float g, float b, float sat, float* ab) {
  r = lin_rgb(r); g = lin_rgb(g); b = lin_rgb(b);
  const float l = signed_cbrt(0.4122214708f * r + 0.5363325363f * g + 0.0514459929f * b);
  const float m = signed_cbrt(0.2119034982f * r + 0.6806995451f * g + 0.1073969566f * b);
  const float s = signed_cbrt(0.0883024619f * r + 0.2817188376f * g + 0.6299787005f * b);
  ab[0] = fminf(fmaxf((1.9779984951f * l - 2.4285922050f * m + 0.4505937099f * s) * sat + 0.5f, 0.f), 1.f);
  ab[1] = fminf(fmaxf((0.0259040371f * l + 0.7827717662f * m - 0.8086757660f * s) * sat + 0.5f, 0.f), 1.f);
}
__device__ __forceinline__ void oklab_ab_grad(float r0, float g0, float b0, const float* up, float* d) {
  const float r = lin_rgb(r0), g = lin_rgb(g0), b = lin_rgb(b0);
  const float lv = 0.4122214708f * r + 0.5363325363f * g + 0.0514459929f * b;
  const float mv = 0.2119034982f * r + 0.6806995451f * g + 0.1073969566f * b;
  const float sv = 0.0883024619f * r + 0.2817188376f * g + 0.6299787005f * b;
  const float l = signed_cbrt(lv), m = signed_cbrt(mv), s = signed_cbrt(sv);
  const float a = 1.9779984951f * l - 2.4285922050f * m + 0.4505937099f * s + 0.5f;
  const float bb = 0.0259040371f * l + 0.7827717662f * m - 0.8086757660f * s + 0.5f;
  const float ua = (a >= 0.f && a <= 1.f) ? up[0] : 0.f, ub = (bb >= 0.f && bb <= 1.f) ? up[1] : 0.f;
  const float dl = (1.9779984951f * ua + 0.0259040371f * ub) * cbrt_grad(lv);
  const float dm = (-2.4285922050f * ua + 0.7827717662f * ub) * cbrt_grad(mv);
  const float dsv = (0.4505937099f * ua - 0.8086757660f * ub) * cbrt_grad(sv);
  d[0] = (0.4122214708f * dl + 0.2119034982f * dm + 0.0883024619f * dsv) * lin_rgb_grad(r0);
  d[1] = (0.5363325363f * dl + 0.6806995451f * dm + 0.2817188376f * dsv) * lin_rgb_grad(g0);
  d[2] = (0.0514459929f * dl + 0.1073969566f * dm + 0.6299787005f * dsv) * lin_rgb_grad(b0);
}
__device__ __forceinline__ float clampc(float v) { return fminf(fmaxf(v, 1.f / 255.f), 1.f); }
__device__ __forceinline__ float chc_val(float d) { return fminf(fmaxf(sqrtf(d * d + 1e-12f), 0.f), 1.f); }
__device__ __forceinline__ float chc_grad(float d) {
  const float v = sqrtf(d * d + 1e-12f);
  return v <= 1.f ? d / v : 0.f;
}

// Per-pixel forward.  xb / yb: the 21x21 Gaussian blurs of clamp(x, 1/255, 1) / clamp(y, 1/255, 1) (or the
// clamped images themselves when blur is off).  Writes the luma planes and block partials:
// [0] chc(luma)  [1] chc(chroma)  [2] sum of chroma cosine similarities.
__global__ void __launch_bounds__(STHREADS) consistency_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                   const float* __restrict__ xb, const float* __restrict__ yb,
                                                                   float sat, float bright, float* __restrict__ luma_x,
                                                                   float* __restrict__ luma_y, float* __restrict__ partial,
                                                                   int B, int H, int W) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / hw, pix = idx - b * hw, o = b * 3 * hw + pix;
    const float lx = l_star(fminf(fmaxf(xb[o], 0.f), 1.f), fminf(fmaxf(xb[o + hw], 0.f), 1.f), fminf(fmaxf(xb[o + 2 * hw], 0.f), 1.f));
    const float ly = l_star(fminf(fmaxf(yb[o], 0.f), 1.f), fminf(fmaxf(yb[o + hw], 0.f), 1.f), fminf(fmaxf(yb[o + 2 * hw], 0.f), 1.f)) * bright;
    luma_x[idx] = lx;
    luma_y[idx] = ly;
    s0 += chc_val(lx - ly);
    float ca[2], cb[2];
    oklab_ab(clampc(x[o]), clampc(x[o + hw]), clampc(x[o + 2 * hw]), 1.f, ca);
    oklab_ab(clampc(y[o]), clampc(y[o + hw]), clampc(y[o + 2 * hw]), sat, cb);
    s1 += chc_val(ca[0] - cb[0]) + chc_val(ca[1] - cb[1]);
    const float na = fmaxf(sqrtf(ca[0] * ca[0] + ca[1] * ca[1]), 1e-20f), nb = fmaxf(sqrtf(cb[0] * cb[0] + cb[1] * cb[1]), 1e-20f);
    s2 += (ca[0] / na) * (cb[0] / nb) + (ca[1] / na) * (cb[1] / nb);
  }
  __shared__ float red[3][STHREADS / 32];
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f, c = 0.f;
    for (int i = 0; i < STHREADS / 32; ++i) { a += red[0][i]; b2 += red[1][i]; c += red[2][i]; }
    partial[3 * blockIdx.x] = a; partial[3 * blockIdx.x + 1] = b2; partial[3 * blockIdx.x + 2] = c;
  }
}
// nn.CosineSimilarity(dim=1) on the [B,H,W] luma planes runs along H (consistency_loss.py:181): one thread per
// (b, w) column; stores (dot, |a|, |b|) clamped as torch does and the column's cosine.
__global__ void consistency_col_kernel(const float* __restrict__ la, const float* __restrict__ lb, float* __restrict__ col,
                                       int B, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const int b = i / W, w = i - b * W;
  const float* a = la + (size_t)b * H * W + w;
  const float* c = lb + (size_t)b * H * W + w;
  float na = 0.f, nb = 0.f;
  for (int h = 0; h < H; ++h) { const float u = a[(size_t)h * W], v = c[(size_t)h * W]; na = fmaf(u, u, na); nb = fmaf(v, v, nb); }
  na = fmaxf(sqrtf(na), 1e-20f); nb = fmaxf(sqrtf(nb), 1e-20f);
  float dot = 0.f;
  for (int h = 0; h < H; ++h) dot += (a[(size_t)h * W] / na) * (c[(size_t)h * W] / nb);
  col[3 * i] = dot; col[3 * i + 1] = na; col[3 * i + 2] = nb;
}
// scal[0] = chc luma mean, [1] = chc chroma mean, [2] = cosim, [3] = flag (cosim < 1e-3), loss out
__global__ void consistency_final_kernel(const float* __restrict__ partial, int nblocks, const float* __restrict__ col,
                                         int ncols, float n_luma, float n_chroma, int use_cosim, float weight,
                                         float* __restrict__ scal, float* loss_value, float* loss_accum) {
  __shared__ double sm[4][STHREADS];
  double a = 0., b = 0., c = 0., d = 0.;
  for (int i = threadIdx.x; i < nblocks; i += STHREADS) { a += partial[3 * i]; b += partial[3 * i + 1]; c += partial[3 * i + 2]; }
  for (int i = threadIdx.x; i < ncols; i += STHREADS) d += col[3 * i];
  sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b; sm[2][threadIdx.x] = c; sm[3][threadIdx.x] = d;
  __syncthreads();
  for (int o = STHREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float luma = (float)(sm[0][0] / n_luma), chroma = (float)(sm[1][0] / n_chroma);
    const float cos_c = 1.f - (float)(sm[2][0] / n_luma), cos_l = 1.f - (float)(sm[3][0] / ncols);
    const float cosim = 0.5f * cos_c + 0.5f * cos_l;
    const float flag = (use_cosim && cosim < 1e-3f) ? 1.f : 0.f;
    scal[0] = luma; scal[1] = chroma; scal[2] = cosim; scal[3] = flag;
    const float v = weight * (luma + chroma + flag * cosim);
    if (loss_value) *loss_value = v;
    if (loss_accum) *loss_accum += v;
  }
}
// Per-pixel backward: g_blur = dloss/d(blurred, pre-clamp image) [B,3,H,W]; d_direct = chroma path gradient wrt x
// (clamp(x,1/255,1) mask applied).  When blur is off g_blur is folded into d_direct by the caller.
__global__ void __launch_bounds__(STHREADS) consistency_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                   const float* __restrict__ xb, const float* __restrict__ luma_x,
                                                                   const float* __restrict__ luma_y, const float* __restrict__ col,
                                                                   const float* __restrict__ scal, float sat, float weight,
                                                                   float* __restrict__ g_blur, float* __restrict__ d_direct,
                                                                   int B, int H, int W) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  const float flag = scal[3];
  const float inv_nl = 1.f / (float)total, inv_nc = 1.f / (float)(2 * total), inv_cols = 1.f / (float)((size_t)B * W);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / hw, pix = idx - b * hw, o = b * 3 * hw + pix;
    const int w = (int)(pix % W);
    // ---- luma: chc + (flagged) column cosine
    const float lx = luma_x[idx], ly = luma_y[idx];
    float up = weight * chc_grad(lx - ly) * inv_nl;
    if (flag != 0.f) {
      const float* cc = col + 3 * (b * W + w);
      const float cosv = cc[0], na = cc[1], nb = cc[2];
      // d/da_h [ sum_h (a_h/na)(b_h/nb) ] = b_h/(na nb) - cos * a_h / na^2   (na above the eps clamp)
      up += weight * (-0.5f * inv_cols) * (ly / (na * nb) - cosv * lx / (na * na));
    }
    float r = xb[o], g = xb[o + hw], bl = xb[o + 2 * hw], d[3];
    l_star_grad(fminf(fmaxf(r, 0.f), 1.f), fminf(fmaxf(g, 0.f), 1.f), fminf(fmaxf(bl, 0.f), 1.f), up, d);
    g_blur[o] = (r >= 0.f && r <= 1.f) ? d[0] : 0.f;
    g_blur[o + hw] = (g >= 0.f && g <= 1.f) ? d[1] : 0.f;
    g_blur[o + 2 * hw] = (bl >= 0.f && bl <= 1.f) ? d[2] : 0.f;
    // ---- chroma: chc + (flagged) per-pixel cosine
    float ca[2], cb[2], upc[2];
    const float x0 = x[o], x1 = x[o + hw], x2 = x[o + 2 * hw];
    oklab_ab(clampc(x0), clampc(x1), clampc(x2), 1.f, ca);
    oklab_ab(clampc(y[o]), clampc(y[o + hw]), clampc(y[o + 2 * hw]), sat, cb);
    upc[0] = weight * chc_grad(ca[0] - cb[0]) * inv_nc;
    upc[1] = weight * chc_grad(ca[1] - cb[1]) * inv_nc;
    if (flag != 0.f) {
      const float na = fmaxf(sqrtf(ca[0] * ca[0] + ca[1] * ca[1]), 1e-20f), nb = fmaxf(sqrtf(cb[0] * cb[0] + cb[1] * cb[1]), 1e-20f);
      const float cosv = (ca[0] / na) * (cb[0] / nb) + (ca[1] / na) * (cb[1] / nb);
      const float k = weight * (-0.5f * inv_nl);
      upc[0] += k * (cb[0] / (na * nb) - cosv * ca[0] / (na * na));
      upc[1] += k * (cb[1] / (na * nb) - cosv * ca[1] / (na * na));
    }
    oklab_ab_grad(clampc(x0), clampc(x1), clampc(x2), upc, d);
    d_direct[o] = (x0 >= 1.f / 255.f && x0 <= 1.f) ? d[0] : 0.f;
    d_direct[o + hw] = (x1 >= 1.f / 255.f && x1 <= 1.f) ? d[1] : 0.f;
    d_direct[o + 2 * hw] = (x2 >= 1.f / 255.f && x2 <= 1.f) ? d[2] : 0.f;
  }
}
// Adjoint of reflect padding: dx[i,j] = sum over the (<= 3 x 3) padded positions that read pixel (i,j) of
// dpad (the zero-extended correlation of g_blur on the [-r, H+r) x [-r, W+r) domain), + d_direct, then the
// clamp(x, 1/255, 1) mask of the blur input.
__global__ void __launch_bounds__(256) reflect_fold_kernel(const float* __restrict__ dpad, const float* __restrict__ d_direct,
                                                           const float* __restrict__ x, float* __restrict__ dx, int planes,
                                                           int H, int W, int r) {
  const size_t total = (size_t)planes * H * W;
  const int PH = H + 2 * r, PW = W + 2 * r;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W), i = (int)((idx / W) % H);
    const size_t p = idx / ((size_t)W * H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = i + r;
    if (i >= 1 && i <= r) ys[ny++] = r - i;
    if (i <= H - 2 && i >= H - 1 - r) ys[ny++] = r + 2 * (H - 1) - i;
    xs[nx++] = j + r;
    if (j >= 1 && j <= r) xs[nx++] = r - j;
    if (j <= W - 2 && j >= W - 1 - r) xs[nx++] = r + 2 * (W - 1) - j;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) s += dpad[p * PH * PW + (size_t)ys[a] * PW + xs[b]];
    const float xv = x[idx];
    dx[idx] = ((xv >= 1.f / 255.f && xv <= 1.f) ? s : 0.f) + d_direct[idx];
  }
}
// clamp(x, lo, hi) elementwise (the clamp in front of the blur, consistency_loss.py:155-156)
__global__ void clamp_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float lo, float hi) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = fminf(fmaxf(in[i], lo), hi);
}
// Zero-padded correlation evaluated on the domain extended by `ext` on every side (ext = 0: "same").
constexpr int ZT = 32, ZMAXK = 21;
__global__ void __launch_bounds__(256) corr_zero_ext_kernel(const float* __restrict__ img, const float* __restrict__ kern,
                                                            float* __restrict__ out, int H, int W, int k, int ext, int flip) {
  __shared__ float tile[(ZT + ZMAXK - 1) * (ZT + ZMAXK)];
  __shared__ float kw[ZMAXK * ZMAXK];
  const int plane = blockIdx.z, x0 = blockIdx.x * ZT, y0 = blockIdx.y * ZT;
  const int r = k / 2, span = ZT + k - 1, pitch = span + 1, OH = H + 2 * ext, OW = W + 2 * ext;
  const float* src = img + (size_t)plane * H * W;
  for (int i = threadIdx.x; i < k * k; i += 256) kw[i] = kern[flip ? k * k - 1 - i : i];
  for (int i = threadIdx.x; i < span * span; i += 256) {
    const int ty = i / span, tx = i - ty * span, gy = y0 + ty - r - ext, gx = x0 + tx - r - ext;
    tile[ty * pitch + tx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? src[(size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      const float wv = kw[i * k + j];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(tile[(ty0 + 8 * q + i) * pitch + tx + j], wv, acc[q]);
    }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int y = y0 + ty0 + 8 * q, x = x0 + tx;
    if (y < OH && x < OW) out[(size_t)plane * OH * OW + (size_t)y * OW + x] = acc[q];
  }
}

static inline int grid1d(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
}  // namespace nsr
using namespace nsr;

extern "C" int nsr_avgpool2(const float* in, float* out, int planes, int h, int w, int pad_h, int pad_w, void* stream) {
  NSR_CHECK_ARG(in && out && planes > 0 && h > 0 && w > 0 && pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1,
                "nsr_avgpool2: bad arguments");
  const int oh = (h + 2 * pad_h - 2) / 2 + 1, ow = (w + 2 * pad_w - 2) / 2 + 1;
  avgpool2_kernel<<<grid1d((size_t)planes * oh * ow), 256, 0, (cudaStream_t)stream>>>(in, out, planes, h, w, oh, ow, pad_h, pad_w);
  NSR_CHECK_LAUNCH("nsr_avgpool2");
  return NSR_OK;
}

static inline size_t ssim_blocks(int planes, int h, int w) { return (size_t)planes * ceil_div(h, ST) * ceil_div(w, ST); }
extern "C" size_t nsr_ssim_scale_workspace(int planes, int h, int w) { return ssim_blocks(planes, h, w) * 2 * sizeof(float); }

extern "C" int nsr_ssim_scale_fwd(const float* x, const float* y, const float* window, int window_size, float c1, float c2,
                                  int use_ssim, float* partials3, float* sums2, int planes, int h, int w, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(x && y && window && sums2 && planes > 0 && h > 0 && w > 0, "nsr_ssim_scale_fwd: bad arguments");
  NSR_CHECK_ARG(window_size % 2 == 1 && window_size <= SMAXW, "nsr_ssim_scale_fwd: Window size must be odd (and <= %d)", SMAXW);
  NSR_CHECK_ARG(planes <= 65535, "nsr_ssim_scale_fwd: too many planes");
  if (!workspace || workspace_bytes < nsr_ssim_scale_workspace(planes, h, w)) {
    set_error("nsr_ssim_scale_fwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(w, ST), ceil_div(h, ST), planes);
  ssim_fwd_kernel<<<grid, STHREADS, 0, st>>>(x, y, window, window_size, c1, c2, use_ssim, partials3, (float*)workspace, planes, h, w);
  NSR_CHECK_LAUNCH("nsr_ssim_scale_fwd");
  ssim_reduce_kernel<<<1, STHREADS, 0, st>>>((const float*)workspace, (int)ssim_blocks(planes, h, w), sums2);
  NSR_CHECK_LAUNCH("nsr_ssim_scale_fwd(reduce)");
  return NSR_OK;
}

extern "C" int nsr_msssim_finalize(const float* sums, const float* counts, int nscales, float weight, float* coef,
                                   float* loss_value, float* loss_accum, void* stream) {
  NSR_CHECK_ARG(sums && counts && coef && nscales >= 1 && nscales <= 5, "nsr_msssim_finalize: bad arguments");
  msssim_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, counts, nscales, weight, coef, loss_value, loss_accum);
  NSR_CHECK_LAUNCH("nsr_msssim_finalize");
  return NSR_OK;
}

extern "C" int nsr_ssim_scale_bwd(const float* partials3, const float* x, const float* y, const float* window, int window_size,
                                  const float* coef, const float* dx_coarse, int coarse_h, int coarse_w, int pad_h, int pad_w,
                                  float* dx, int planes, int h, int w, void* stream) {
  NSR_CHECK_ARG(partials3 && x && y && window && coef && dx && planes > 0 && planes <= 65535, "nsr_ssim_scale_bwd: bad arguments");
  NSR_CHECK_ARG(window_size % 2 == 1 && window_size <= SMAXW, "nsr_ssim_scale_bwd: bad window");
  dim3 grid(ceil_div(w, ST), ceil_div(h, ST), planes);
  ssim_bwd_kernel<<<grid, STHREADS, 0, (cudaStream_t)stream>>>(partials3, x, y, window, window_size, coef, dx_coarse, coarse_h,
                                                               coarse_w, pad_h, pad_w, dx, planes, h, w);
  NSR_CHECK_LAUNCH("nsr_ssim_scale_bwd");
  return NSR_OK;
}

extern "C" int nsr_clamp(const float* in, float* out, size_t n, float lo, float hi, void* stream) {
  NSR_CHECK_ARG(in && out && n > 0, "nsr_clamp: bad arguments");
  clamp_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, lo, hi);
  NSR_CHECK_LAUNCH("nsr_clamp");
  return NSR_OK;
}

extern "C" int nsr_corr2d_zero_ext(const float* img, const float* kernel, float* out, int planes, int h, int w, int k, int ext,
                                   int flip, void* stream) {
  NSR_CHECK_ARG(img && kernel && out && planes > 0 && planes <= 65535 && h > 0 && w > 0, "nsr_corr2d_zero_ext: bad arguments");
  NSR_CHECK_ARG(k % 2 == 1 && k <= ZMAXK && ext >= 0, "nsr_corr2d_zero_ext: bad kernel size / extension");
  dim3 grid(ceil_div(w + 2 * ext, ZT), ceil_div(h + 2 * ext, ZT), planes);
  corr_zero_ext_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, kernel, out, h, w, k, ext, flip);
  NSR_CHECK_LAUNCH("nsr_corr2d_zero_ext");
  return NSR_OK;
}

extern "C" size_t nsr_consistency_workspace(int batch, int h, int w) {
  return ((size_t)grid1d((size_t)batch * h * w) * 3 + (size_t)batch * w * 3 + 8) * sizeof(float);
}
/* workspace layout: [partials: 3*blocks][col: 3*B*W][scal: 4 (+4 pad)] */
extern "C" int nsr_consistency_fwd(const float* x, const float* y, const float* x_blur, const float* y_blur, float saturation,
                                   float brightness, int use_cosim, float weight, float* luma_x, float* luma_y,
                                   float* loss_value, float* loss_accum, int batch, int h, int w, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(x && y && x_blur && y_blur && luma_x && luma_y && batch > 0 && h > 0 && w > 0, "nsr_consistency_fwd: bad arguments");
  if (!workspace || workspace_bytes < nsr_consistency_workspace(batch, h, w)) {
    set_error("nsr_consistency_fwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)batch * h * w;
  const int blocks = grid1d(n);
  float* partial = (float*)workspace;
  float* col = partial + 3 * (size_t)blocks;
  float* scal = col + 3 * (size_t)batch * w;
  consistency_fwd_kernel<<<blocks, STHREADS, 0, st>>>(x, y, x_blur, y_blur, saturation, brightness, luma_x, luma_y, partial, batch, h, w);
  NSR_CHECK_LAUNCH("nsr_consistency_fwd");
  consistency_col_kernel<<<ceil_div((long long)batch * w, 128), 128, 0, st>>>(luma_x, luma_y, col, batch, h, w);
  NSR_CHECK_LAUNCH("nsr_consistency_fwd(col)");
  consistency_final_kernel<<<1, STHREADS, 0, st>>>(partial, blocks, col, batch * w, (float)n, (float)(2 * n), use_cosim, weight, scal,
                                                   loss_value, loss_accum);
  NSR_CHECK_LAUNCH("nsr_consistency_fwd(final)");
  return NSR_OK;
}

extern "C" int nsr_consistency_bwd(const float* x, const float* y, const float* x_blur, const float* luma_x, const float* luma_y,
                                   float saturation, float weight, float* g_blur, float* d_direct, int batch, int h, int w,
                                   const void* workspace, void* stream) {
  NSR_CHECK_ARG(x && y && x_blur && luma_x && luma_y && g_blur && d_direct && workspace && batch > 0, "nsr_consistency_bwd: bad arguments");
  const size_t n = (size_t)batch * h * w;
  const int blocks = grid1d(n);
  const float* col = (const float*)workspace + 3 * (size_t)blocks;
  const float* scal = col + 3 * (size_t)batch * w;
  consistency_bwd_kernel<<<blocks, STHREADS, 0, (cudaStream_t)stream>>>(x, y, x_blur, luma_x, luma_y, col, scal, saturation, weight,
                                                                        g_blur, d_direct, batch, h, w);
  NSR_CHECK_LAUNCH("nsr_consistency_bwd");
  return NSR_OK;
}

extern "C" int nsr_reflect_fold(const float* dpad, const float* d_direct, const float* x, float* dx, int planes, int h, int w,
                                int r, void* stream) {
  NSR_CHECK_ARG(dpad && d_direct && x && dx && planes > 0 && r >= 0 && r < h && r < w, "nsr_reflect_fold: bad arguments");
  reflect_fold_kernel<<<grid1d((size_t)planes * h * w), 256, 0, (cudaStream_t)stream>>>(dpad, d_direct, x, dx, planes, h, w, r);
  NSR_CHECK_LAUNCH("nsr_reflect_fold");
  return NSR_OK;
}

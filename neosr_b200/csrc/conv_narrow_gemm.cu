// <= 4-channel image-side convolutions (conv_first / conv_last, VGG conv1_1, the discriminator's ends)
// as tensor-core GEMMs.  The work is tiny (K or N = taps * narrow <= 36) but the tensors are the largest
// of the step (C3: 2M pixels x 64 channels), so the bound is HBM; the SIMT kernels in conv_small.cu are
// instruction-bound at ~10 % of it.  Here
//   narrow -> wide : col[M, Kp] = im2col(x) (Kp = taps*cin rounded up to 16), then ONE 1x1 contraction
//                    col x W'^T on the tcgen05 engine with the regular fused epilogue;
//   wide -> narrow : z[M, Np] = x x W''^T (row (tap, co) of W'' = filter tap), then
//                    y[p, co] = bias + sum_tap z[p + tap offset][tap, co]  (zero outside the image);
// and the two weight gradients are 1x1 wgrad contractions against the same column matrices.
// Used when the caller passes workspace (NsrConv.workspace / nsr_conv_wgrad_workspace) and M is large.
#include "common.cuh"

namespace nsr {

bool conv_fprop_tc_supported(const NsrConv& d);
int conv_fprop_tc(const NsrConv& d, cudaStream_t st);
bool conv_wgrad_tc_supported(const NsrWgrad& d);
size_t conv_wgrad_workspace_tc(const NsrWgrad& d);
int conv_wgrad_tc(const NsrWgrad& d, cudaStream_t st);
int conv_bias_grad(const NsrWgrad& d, float* bias_partial, int bias_blocks, cudaStream_t st);
int launch_pack_weight_bf16(const float* wf32, uint8_t* img, const PackedGeom& g, cudaStream_t st);

constexpr long long NG_MIN_PIXELS = 16384;  // below this the SIMT kernels are launch-latency bound anyway

struct NgGeom {
  int H, W, kh, kw, pad, C, ld, Kp, flip;
  long long M;
};

static inline int ng_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}
static inline size_t up1k(size_t b) { return (b + 1023) / 1024 * 1024; }

// col[p, t*C + c] = x[p + offset(t'), c], t' = t or (taps-1-t) when flip (i.e. the negated offset); 0 outside
__global__ void narrow_im2col(const float* __restrict__ x, float* __restrict__ col, NgGeom g) {
  const int chunks = g.Kp / 4, taps = g.kh * g.kw, K = taps * g.C;
  const long long total = g.M * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / chunks;
    const int k0 = (int)(i - p * chunks) * 4;
    const int rem = (int)(p % ((long long)g.H * g.W));
    const int oh = rem / g.W, ow = rem - oh * g.W;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      v[e] = 0.f;
      if (k < K) {
        const int t = k / g.C, c = k - t * g.C;
        const int tt = g.flip ? taps - 1 - t : t;
        const int r = tt / g.kw, s = tt - r * g.kw;
        const int ih = oh + r - g.pad, iw = ow + s - g.pad;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          v[e] = __ldg(x + (p + (long long)(r - g.pad) * g.W + (s - g.pad)) * g.ld + c);
      }
    }
    *reinterpret_cast<float4*>(col + p * g.Kp + k0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// y[p, co] = bias[co] + sum_t z[(p + offset(t)) * Np + t*C + co]
__global__ void narrow_col2im(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ y, int y_ld,
                              NgGeom g) {
  const int taps = g.kh * g.kw;
  const long long total = g.M * g.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / g.C;
    const int co = (int)(i - p * g.C);
    const int rem = (int)(p % ((long long)g.H * g.W));
    const int oh = rem / g.W, ow = rem - oh * g.W;
    float acc = bias ? bias[co] : 0.f;
    for (int t = 0; t < taps; ++t) {
      const int r = t / g.kw, s = t - r * g.kw;
      const int ih = oh + r - g.pad, iw = ow + s - g.pad;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
        acc += __ldg(z + (p + (long long)(r - g.pad) * g.W + (s - g.pad)) * g.Kp + t * g.C + co);
    }
    y[p * y_ld + co] = acc;
  }
}
// fp32 view [n][K] -> [n][Kp] (zero padded): GEMM weights of the narrow->wide form
__global__ void narrow_w_rows(const float* __restrict__ view, float* __restrict__ out, int n, int K, int Kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Kp) return;
  const int r = i / Kp, k = i - r * Kp;
  out[i] = k < K ? view[(size_t)r * K + k] : 0.f;
}
// fp32 view [co][t][ci] -> [Np][ci], row t*cout + co (zero rows beyond): GEMM weights of wide->narrow
__global__ void narrow_w_taps(const float* __restrict__ view, float* __restrict__ out, int cout, int taps, int cin, int Np) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np * cin) return;
  const int row = i / cin, ci = i - row * cin;
  const int t = row / cout, co = row - t * cout;
  out[i] = row < taps * cout ? view[((size_t)co * taps + t) * cin + ci] : 0.f;
}
// gradient scatter back to OIHW: n2w: dw[co][ci][t] = g[co][t*cin + ci];  w2n: dw[co][ci][t] = g[t*cout + co][ci]
__global__ void narrow_dw_scatter(const float* __restrict__ gmat, float* __restrict__ dw, int cout, int cin, int taps, int ldg,
                                  int w2n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * taps) return;
  const int t = i % taps, ci = (i / taps) % cin, co = i / (taps * cin);
  dw[i] = w2n ? gmat[(size_t)(t * cout + co) * ldg + ci] : gmat[(size_t)co * ldg + t * cin + ci];
}

static inline int pad_k(int k) { return (k + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------ fprop
static bool ng_shape(const NsrConv& d, bool& n2w) {
  if (d.kh * d.kw > 9) return false;
  if (d.cin <= 4 && d.cout >= 16 && d.cout % 4 == 0) { n2w = true; return true; }
  if (d.cout <= 4 && d.cin >= 16 && d.cin % 4 == 0) { n2w = false; return true; }
  return false;
}
static NsrConv ng_inner_fprop(const NsrConv& d, bool n2w) {
  NsrConv e = d;
  e.kh = e.kw = 1;
  e.pad = 0;
  const int taps = d.kh * d.kw;
  if (n2w) {
    e.cin = e.x_ld = pad_k(taps * d.cin);
  } else {
    e.cout = e.y_ld = pad_k(taps * d.cout);
    e.bias = nullptr;
  }
  return e;
}
size_t conv_narrow_gemm_workspace(const NsrConv& d) {
  bool n2w;
  if (!ng_shape(d, n2w)) return 0;
  const size_t M = (size_t)d.batch * d.h * d.w;
  if ((long long)M < NG_MIN_PIXELS) return 0;
  const NsrConv e = ng_inner_fprop(d, n2w);
  const PackedGeom pg = packed_geom(e.cout, e.cin, 1, 1, 0);
  const size_t mat = n2w ? M * e.cin : M * e.cout;
  return up1k(mat * 4) + up1k((size_t)e.cout * e.cin * 4) + up1k(pg.bf16_bytes) + 2048;
}
bool conv_narrow_gemm_supported(const NsrConv& d) {
  if (!nsr_device_supports_tcgen05()) return false;
  bool n2w;
  if (!ng_shape(d, n2w)) return false;
  if (!d.x || !d.y || d.x_sti || d.y_sti) return false;
  if ((long long)d.batch * d.h * d.w < NG_MIN_PIXELS) return false;
  const size_t need = conv_narrow_gemm_workspace(d);
  if (!d.workspace || d.workspace_bytes < need) return false;
  if (!n2w && (d.act != NSR_ACT_NONE || d.actgrad || d.residual || d.row_scale || d.y_pre || d.pre_mode)) return false;
  NsrConv e = ng_inner_fprop(d, n2w);
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)d.workspace + 1023) & ~(uintptr_t)1023);
  if (n2w) e.x = reinterpret_cast<const float*>(ws);
  else e.y = reinterpret_cast<float*>(ws);
  e.w_packed = ws;  // alignment probe only
  return conv_fprop_tc_supported(e);
}
int conv_narrow_gemm_fprop(const NsrConv& d, cudaStream_t st) {
  bool n2w = false;
  ng_shape(d, n2w);
  const long long M = (long long)d.batch * d.h * d.w;
  const int taps = d.kh * d.kw;
  NsrConv e = ng_inner_fprop(d, n2w);
  const PackedGeom pg = packed_geom(e.cout, e.cin, 1, 1, 0);
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)d.workspace + 1023) & ~(uintptr_t)1023);
  float* mat = reinterpret_cast<float*>(ws);
  const size_t mat_bytes = up1k((size_t)M * (n2w ? e.cin : e.cout) * 4);
  float* wf = reinterpret_cast<float*>(ws + mat_bytes);
  uint8_t* img = ws + mat_bytes + up1k((size_t)e.cout * e.cin * 4);
  const float* view = reinterpret_cast<const float*>(d.w_packed);  // fp32 view W[n][tap][c] of the caller's filter
  NgGeom g;
  g.H = d.h; g.W = d.w; g.kh = d.kh; g.kw = d.kw; g.pad = d.pad; g.M = M; g.flip = 0;
  if (n2w) {
    narrow_w_rows<<<ceil_div(e.cout * e.cin, 256), 256, 0, st>>>(view, wf, d.cout, taps * d.cin, e.cin);
    g.C = d.cin; g.ld = d.x_ld; g.Kp = e.cin;
    narrow_im2col<<<ng_blocks(M * (e.cin / 4)), 256, 0, st>>>(d.x, mat, g);
    e.x = mat;
  } else {
    narrow_w_taps<<<ceil_div(e.cout * e.cin, 256), 256, 0, st>>>(view, wf, d.cout, taps, d.cin, e.cout);
    e.y = mat;
  }
  NSR_CHECK_LAUNCH("conv_narrow_gemm: prepare");
  int rc = launch_pack_weight_bf16(wf, img, pg, st);
  if (rc != NSR_OK) return rc;
  e.w_packed = wf;  // [fp32 view | bf16 tile images], the layout nsr_pack_weight produces (f32 region = up1k)
  rc = conv_fprop_tc(e, st);
  if (rc != NSR_OK) return rc;
  if (!n2w) {
    g.C = d.cout; g.ld = e.cout; g.Kp = e.cout;
    narrow_col2im<<<ng_blocks(M * d.cout), 256, 0, st>>>(mat, d.bias, d.y, d.y_ld, g);
    NSR_CHECK_LAUNCH("narrow_col2im");
  }
  return NSR_OK;
}

// ------------------------------------------------------------------------------------ wgrad
static bool ngw_shape(const NsrWgrad& d, bool& n2w) {
  if (d.kh * d.kw > 9) return false;
  if (d.cin <= 4 && d.cout >= 16 && d.cout % 4 == 0) { n2w = true; return true; }
  if (d.cout <= 4 && d.cin >= 16 && d.cin % 4 == 0) { n2w = false; return true; }
  return false;
}
static NsrWgrad ngw_inner(const NsrWgrad& d, bool n2w) {
  NsrWgrad e = d;
  e.kh = e.kw = 1;
  e.pad = 0;
  e.x_sti = e.dy_sti = nullptr;
  const int taps = d.kh * d.kw;
  if (n2w) {
    e.cin = e.x_ld = pad_k(taps * d.cin);
  } else {
    e.cout = e.dy_ld = pad_k(taps * d.cout);
    e.dbias = nullptr;
  }
  return e;
}
static size_t ngw_bias_floats(const NsrWgrad& d) { return (size_t)kNumSMs * 4 * d.cout; }
size_t conv_narrow_gemm_wgrad_workspace(const NsrWgrad& d) {
  bool n2w;
  if (!ngw_shape(d, n2w)) return 0;
  const size_t M = (size_t)d.batch * d.h * d.w;
  if ((long long)M < NG_MIN_PIXELS) return 0;
  NsrWgrad e = ngw_inner(d, n2w);
  e.x = e.dy = reinterpret_cast<const float*>(uintptr_t(1024));  // non-null, aligned: shape probe only
  const size_t mat = n2w ? M * e.cin : M * e.cout;
  return up1k(mat * 4) + up1k((size_t)e.cout * e.cin * 4) + up1k(ngw_bias_floats(d) * 4) + up1k(conv_wgrad_workspace_tc(e)) + 2048;
}
bool conv_narrow_gemm_wgrad_supported(const NsrWgrad& d) {
  if (!nsr_device_supports_tcgen05()) return false;
  bool n2w;
  if (!ngw_shape(d, n2w)) return false;
  if (!d.x || !d.dy) return false;
  if ((long long)d.batch * d.h * d.w < NG_MIN_PIXELS) return false;
  NsrWgrad e = ngw_inner(d, n2w);
  const float* probe = reinterpret_cast<const float*>(uintptr_t(1024));  // the column matrix is 1024-aligned
  if (n2w) e.x = probe;
  else e.dy = probe;
  return conv_wgrad_tc_supported(e);
}
int conv_narrow_gemm_wgrad(const NsrWgrad& d, cudaStream_t st) {
  const size_t need = conv_narrow_gemm_wgrad_workspace(d);
  if (!d.workspace || d.workspace_bytes < need) {
    set_error("nsr_conv_wgrad(narrow gemm): workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  bool n2w = false;
  ngw_shape(d, n2w);
  const long long M = (long long)d.batch * d.h * d.w;
  const int taps = d.kh * d.kw;
  NsrWgrad e = ngw_inner(d, n2w);
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)d.workspace + 1023) & ~(uintptr_t)1023);
  float* mat = reinterpret_cast<float*>(ws);
  const size_t mat_bytes = up1k((size_t)M * (n2w ? e.cin : e.cout) * 4);
  float* gmat = reinterpret_cast<float*>(ws + mat_bytes);
  const size_t gmat_bytes = up1k((size_t)e.cout * e.cin * 4);
  float* bias_partial = reinterpret_cast<float*>(ws + mat_bytes + gmat_bytes);
  uint8_t* inner_ws = ws + mat_bytes + gmat_bytes + up1k(ngw_bias_floats(d) * 4);
  NgGeom g;
  g.H = d.h; g.W = d.w; g.kh = d.kh; g.kw = d.kw; g.pad = d.pad; g.M = M;
  if (n2w) {
    g.C = d.cin; g.ld = d.x_ld; g.Kp = e.cin; g.flip = 0;
    narrow_im2col<<<ng_blocks(M * (e.cin / 4)), 256, 0, st>>>(d.x, mat, g);
    e.x = mat;
  } else {
    g.C = d.cout; g.ld = d.dy_ld; g.Kp = e.cout; g.flip = 1;   // colY[q, (t, co)] = dy[q - offset(t), co]
    narrow_im2col<<<ng_blocks(M * (e.cout / 4)), 256, 0, st>>>(d.dy, mat, g);
    e.dy = mat;
  }
  NSR_CHECK_LAUNCH("narrow_im2col");
  e.dw = gmat;
  e.workspace = inner_ws;
  e.workspace_bytes = d.workspace_bytes - (size_t)(inner_ws - reinterpret_cast<uint8_t*>(d.workspace));
  int rc = conv_wgrad_tc(e, st);
  if (rc != NSR_OK) return rc;
  narrow_dw_scatter<<<ceil_div(d.cout * d.cin * taps, 256), 256, 0, st>>>(gmat, d.dw, d.cout, d.cin, taps, e.cin, n2w ? 0 : 1);
  NSR_CHECK_LAUNCH("narrow_dw_scatter");
  if (!n2w && d.dbias) {
    int bias_blocks = bias_grad_blocks(M);
    return conv_bias_grad(d, bias_partial, bias_blocks, st);
  }
  return NSR_OK;
}

}  // namespace nsr

// HAT pieces (neosr/archs/hat_arch.py): generic (cross-)window attention for 16x16 query windows — the HAB
// self-attention (121-215, 299-350: cyclic shift + {0,-100} mask as index math) and the OCAB overlapping
// cross-attention (445-515: keys/values from the 24x24 zero-padded neighbourhood that nn.Unfold would
// materialise) — and the CAB channel-attention gate (15-52).  fp32 CUDA-core kernels, flash-style: the
// [Nq x Nk] score matrix never leaves registers; backward recomputes it from the saved log-sum-exp.
// Everything is deterministic (fixed-order partial sums, no float atomics).
#include "xwin_geom.cuh"

namespace nsr {

constexpr int GA_D = 32;       // max head dim
constexpr int GA_THREADS = 128;
constexpr int GA_CHUNK = 32;   // keys (fwd / bwd-q) or queries (bwd-kv) staged per step
constexpr int GA_MAXTAB = 39 * 39;

__global__ void __launch_bounds__(GA_THREADS) ga_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ table,
                                                            float* __restrict__ out, float* __restrict__ lse, GAGeom g) {
  __shared__ float Ks[GA_CHUNK][GA_D], Vs[GA_CHUNK][GA_D];
  __shared__ int ktok[GA_CHUNK], krid[GA_CHUNK], kjy[GA_CHUNK], kjx[GA_CHUNK];
  __shared__ float tab[GA_MAXTAB];
  const int wi = blockIdx.x / g.heads, head = blockIdx.x - wi * g.heads;
  const int qi = blockIdx.y * GA_THREADS + threadIdx.x;
  const bool active = qi < g.Nq;
  for (int e = threadIdx.x; e < g.ntab; e += GA_THREADS) tab[e] = table[(size_t)e * g.heads + head];
  int qtok = 0, qrid = 0, iy = 0, ix = 0;
  float q[GA_D], acc[GA_D];
#pragma unroll
  for (int d = 0; d < GA_D; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  if (active) {
    ga_query(g, wi, qi, qtok, qrid, iy, ix);
    const float* qp = qkv + (size_t)qtok * 3 * g.C + head * g.D;
#pragma unroll
    for (int d = 0; d < GA_D; ++d)
      if (d < g.D) q[d] = qp[d] * g.scale;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < g.Nk; k0 += GA_CHUNK) {
    __syncthreads();
    if (threadIdx.x < GA_CHUNK) {
      int t = -1, r = 0, jy = 0, jx = 0;
      if (k0 + threadIdx.x < g.Nk) ga_key(g, wi, k0 + threadIdx.x, t, r, jy, jx);
      ktok[threadIdx.x] = t; krid[threadIdx.x] = r; kjy[threadIdx.x] = jy; kjx[threadIdx.x] = jx;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < GA_CHUNK * GA_D; e += GA_THREADS) {
      const int j = e / GA_D, d = e - j * GA_D, t = ktok[j];
      const bool ok = t >= 0 && d < g.D;
      Ks[j][d] = ok ? qkv[(size_t)t * 3 * g.C + g.C + head * g.D + d] : 0.f;
      Vs[j][d] = ok ? qkv[(size_t)t * 3 * g.C + 2 * g.C + head * g.D + d] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int nk = min(GA_CHUNK, g.Nk - k0);
    float s[GA_CHUNK], cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < GA_CHUNK; ++j) {
      float v = -INFINITY;
      if (j < nk) {
        v = 0.f;
#pragma unroll
        for (int d = 0; d < GA_D; ++d) v = fmaf(q[d], Ks[j][d], v);
        v += tab[ga_rel(g, iy, ix, kjy[j], kjx[j])];
        if (g.use_mask && krid[j] != qrid) v += -100.f;
      }
      s[j] = v;
      cmax = fmaxf(cmax, v);
    }
    const float mn = fmaxf(m, cmax), corr = expf(m - mn);
    l *= corr;
#pragma unroll
    for (int d = 0; d < GA_D; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < GA_CHUNK; ++j) {
      if (j < nk) {
        const float p = expf(s[j] - mn);
        l += p;
#pragma unroll
        for (int d = 0; d < GA_D; ++d) acc[d] = fmaf(p, Vs[j][d], acc[d]);
      }
    }
    m = mn;
  }
  if (active) {
    const float inv = 1.f / l;
    float* op = out + (size_t)qtok * g.C + head * g.D;
#pragma unroll
    for (int d = 0; d < GA_D; ++d)
      if (d < g.D) op[d] = acc[d] * inv;
    lse[((size_t)wi * g.heads + head) * g.Nq + qi] = m + logf(l);
  }
}

// thread per query: delta_i = dout_i . out_i, dq_i = scale * sum_j dS_ij k_j
__global__ void __launch_bounds__(GA_THREADS) ga_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ table,
                                                              const float* __restrict__ out, const float* __restrict__ dout,
                                                              const float* __restrict__ lse, float* __restrict__ delta,
                                                              float* __restrict__ dqkv, GAGeom g) {
  __shared__ float Ks[GA_CHUNK][GA_D], Vs[GA_CHUNK][GA_D];
  __shared__ int ktok[GA_CHUNK], krid[GA_CHUNK], kjy[GA_CHUNK], kjx[GA_CHUNK];
  __shared__ float tab[GA_MAXTAB];
  const int wi = blockIdx.x / g.heads, head = blockIdx.x - wi * g.heads;
  const int qi = blockIdx.y * GA_THREADS + threadIdx.x;
  const bool active = qi < g.Nq;
  for (int e = threadIdx.x; e < g.ntab; e += GA_THREADS) tab[e] = table[(size_t)e * g.heads + head];
  int qtok = 0, qrid = 0, iy = 0, ix = 0;
  float q[GA_D], dO[GA_D], dq[GA_D], L = 0.f, dl = 0.f;
#pragma unroll
  for (int d = 0; d < GA_D; ++d) { q[d] = 0.f; dO[d] = 0.f; dq[d] = 0.f; }
  if (active) {
    ga_query(g, wi, qi, qtok, qrid, iy, ix);
    const float* qp = qkv + (size_t)qtok * 3 * g.C + head * g.D;
    const float* op = out + (size_t)qtok * g.C + head * g.D;
    const float* gp = dout + (size_t)qtok * g.C + head * g.D;
#pragma unroll
    for (int d = 0; d < GA_D; ++d)
      if (d < g.D) { q[d] = qp[d] * g.scale; dO[d] = gp[d]; dl = fmaf(gp[d], op[d], dl); }
    L = lse[((size_t)wi * g.heads + head) * g.Nq + qi];
    delta[((size_t)wi * g.heads + head) * g.Nq + qi] = dl;
  }
  for (int k0 = 0; k0 < g.Nk; k0 += GA_CHUNK) {
    __syncthreads();
    if (threadIdx.x < GA_CHUNK) {
      int t = -1, r = 0, jy = 0, jx = 0;
      if (k0 + threadIdx.x < g.Nk) ga_key(g, wi, k0 + threadIdx.x, t, r, jy, jx);
      ktok[threadIdx.x] = t; krid[threadIdx.x] = r; kjy[threadIdx.x] = jy; kjx[threadIdx.x] = jx;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < GA_CHUNK * GA_D; e += GA_THREADS) {
      const int j = e / GA_D, d = e - j * GA_D, t = ktok[j];
      const bool ok = t >= 0 && d < g.D;
      Ks[j][d] = ok ? qkv[(size_t)t * 3 * g.C + g.C + head * g.D + d] : 0.f;
      Vs[j][d] = ok ? qkv[(size_t)t * 3 * g.C + 2 * g.C + head * g.D + d] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int nk = min(GA_CHUNK, g.Nk - k0);
    for (int j = 0; j < nk; ++j) {
      float s = 0.f, dP = 0.f;
#pragma unroll
      for (int d = 0; d < GA_D; ++d) { s = fmaf(q[d], Ks[j][d], s); dP = fmaf(dO[d], Vs[j][d], dP); }
      s += tab[ga_rel(g, iy, ix, kjy[j], kjx[j])];
      if (g.use_mask && krid[j] != qrid) s += -100.f;
      const float dS = expf(s - L) * (dP - dl);
#pragma unroll
      for (int d = 0; d < GA_D; ++d) dq[d] = fmaf(dS, Ks[j][d], dq[d]);
    }
  }
  if (active) {
    float* o = dqkv + (size_t)qtok * 3 * g.C + head * g.D;
#pragma unroll
    for (int d = 0; d < GA_D; ++d)
      if (d < g.D) o[d] = dq[d] * g.scale;
  }
}

// thread per key: dk_j = sum_i dS_ij (q_i*scale), dv_j = sum_i P_ij dout_i, bias-table partials per CTA.
// dkv: HAB -> written straight into dqkv (every token is a key of exactly one window); OCAB -> per-window
// buffer [nwin, Nk, 2, C] folded by ga_fold_kernel.
__global__ void __launch_bounds__(GA_THREADS) ga_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ table,
                                                               const float* __restrict__ dout, const float* __restrict__ lse,
                                                               const float* __restrict__ delta, float* __restrict__ dqkv,
                                                               float* __restrict__ dkv_win, float* __restrict__ dtab_part,
                                                               GAGeom g) {
  __shared__ float Qs[GA_CHUNK][GA_D], Gs[GA_CHUNK][GA_D];
  __shared__ float qL[GA_CHUNK], qdl[GA_CHUNK];
  __shared__ int qrid[GA_CHUNK], qiy[GA_CHUNK], qix[GA_CHUNK], qtk[GA_CHUNK];
  __shared__ float tab[GA_MAXTAB], dtab[GA_MAXTAB];
  const int wi = blockIdx.x / g.heads, head = blockIdx.x - wi * g.heads;
  const int kj = blockIdx.y * GA_THREADS + threadIdx.x;
  const bool active = kj < g.Nk;
  for (int e = threadIdx.x; e < g.ntab; e += GA_THREADS) { tab[e] = table[(size_t)e * g.heads + head]; dtab[e] = 0.f; }
  int ktok = -1, krid = 0, jy = 0, jx = 0;
  float k[GA_D], v[GA_D], dk[GA_D], dv[GA_D];
#pragma unroll
  for (int d = 0; d < GA_D; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  if (active) {
    ga_key(g, wi, kj, ktok, krid, jy, jx);
    if (ktok >= 0) {
      const float* kp = qkv + (size_t)ktok * 3 * g.C + g.C + head * g.D;
#pragma unroll
      for (int d = 0; d < GA_D; ++d)
        if (d < g.D) { k[d] = kp[d]; v[d] = kp[g.C + d]; }
    }
  }
  for (int q0 = 0; q0 < g.Nq; q0 += GA_CHUNK) {
    __syncthreads();
    if (threadIdx.x < GA_CHUNK) {
      int t = 0, r = 0, iy = 0, ix = 0;
      const int qi = q0 + threadIdx.x;
      if (qi < g.Nq) {
        ga_query(g, wi, qi, t, r, iy, ix);
        qL[threadIdx.x] = lse[((size_t)wi * g.heads + head) * g.Nq + qi];
        qdl[threadIdx.x] = delta[((size_t)wi * g.heads + head) * g.Nq + qi];
      }
      qtk[threadIdx.x] = t; qrid[threadIdx.x] = r; qiy[threadIdx.x] = iy; qix[threadIdx.x] = ix;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < GA_CHUNK * GA_D; e += GA_THREADS) {
      const int i = e / GA_D, d = e - i * GA_D;
      const bool ok = q0 + i < g.Nq && d < g.D;
      Qs[i][d] = ok ? qkv[(size_t)qtk[i] * 3 * g.C + head * g.D + d] * g.scale : 0.f;
      Gs[i][d] = ok ? dout[(size_t)qtk[i] * g.C + head * g.D + d] : 0.f;
    }
    __syncthreads();
    const int nq = min(GA_CHUNK, g.Nq - q0);
    for (int i = 0; i < nq; ++i) {
      if (active) {
        float s = 0.f, dP = 0.f;
#pragma unroll
        for (int d = 0; d < GA_D; ++d) { s = fmaf(Qs[i][d], k[d], s); dP = fmaf(Gs[i][d], v[d], dP); }
        const int e = ga_rel(g, qiy[i], qix[i], jy, jx);
        s += tab[e];
        if (g.use_mask && krid != qrid[i]) s += -100.f;
        const float p = expf(s - qL[i]), dS = p * (dP - qdl[i]);
#pragma unroll
        for (int d = 0; d < GA_D; ++d) { dk[d] = fmaf(dS, Qs[i][d], dk[d]); dv[d] = fmaf(p, Gs[i][d], dv[d]); }
        dtab[e] += dS;  // for a fixed query the keys of a CTA map to distinct table entries: no conflict
      }
      __syncthreads();  // orders the table updates of consecutive queries (deterministic, atomic-free)
    }
  }
  if (active) {
    if (dkv_win) {
      float* o = dkv_win + (((size_t)wi * g.Nk + kj) * 2) * g.C + head * g.D;
#pragma unroll
      for (int d = 0; d < GA_D; ++d)
        if (d < g.D) { o[d] = dk[d]; o[g.C + d] = dv[d]; }
    } else if (ktok >= 0) {
      float* o = dqkv + (size_t)ktok * 3 * g.C + g.C + head * g.D;
#pragma unroll
      for (int d = 0; d < GA_D; ++d)
        if (d < g.D) { o[d] = dk[d]; o[g.C + d] = dv[d]; }
    }
  }
  float* part = dtab_part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * g.ntab;
  for (int e = threadIdx.x; e < g.ntab; e += GA_THREADS) part[e] = dtab[e];
}
// OCAB: every token is a key of up to 2x2 overlapping windows; sum their dk/dv in a fixed order.
__global__ void __launch_bounds__(256) ga_fold_kernel(const float* __restrict__ dkv_win, float* __restrict__ dqkv, GAGeom g) {
  const size_t total = (size_t)g.B * g.H * g.W * 2 * g.C;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c2 = (int)(idx % (2 * g.C));
    const size_t tok = idx / (2 * g.C);
    const int x = (int)(tok % g.W), y = (int)((tok / g.W) % g.H), b = (int)(tok / ((size_t)g.W * g.H));
    float s = 0.f;
    const int wy0 = max(0, (y + g.pad - g.ows + g.ws) / g.ws), wy1 = min(g.nwh - 1, (y + g.pad) / g.ws);
    const int wx0 = max(0, (x + g.pad - g.ows + g.ws) / g.ws), wx1 = min(g.nww - 1, (x + g.pad) / g.ws);
    for (int wy = wy0; wy <= wy1; ++wy)
      for (int wx = wx0; wx <= wx1; ++wx) {
        const int jy = y - (wy * g.ws - g.pad), jx = x - (wx * g.ws - g.pad);
        if (jy < 0 || jy >= g.ows || jx < 0 || jx >= g.ows) continue;
        const size_t wi = ((size_t)b * g.nwh + wy) * g.nww + wx;
        s += dkv_win[((wi * g.Nk + jy * g.ows + jx) * 2) * g.C + c2];
      }
    dqkv[tok * 3 * g.C + g.C + c2] = s;
  }
}
// dtable[e, head] = sum over windows / key tiles of the per-CTA partials (fixed order)
__global__ void ga_dtab_reduce_kernel(const float* __restrict__ part, float* __restrict__ dtable, int nwin, int ktiles, GAGeom g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.ntab * g.heads) return;
  const int e = idx / g.heads, head = idx - e * g.heads;
  float s = 0.f;
  for (int kt = 0; kt < ktiles; ++kt)
    for (int wi = 0; wi < nwin; ++wi)
      s += part[((size_t)kt * nwin * g.heads + (size_t)wi * g.heads + head) * g.ntab + e];
  dtable[idx] = s;
}

// tensor-core implementations (xwin_attn_mma.cu)
bool xwin_attn_mma_supported(const GAGeom& g);
int xwin_fwd_mma_launch(const float* qkv, const float* table, float* out, float* lse, const GAGeom& g, cudaStream_t st);
int xwin_bwd_mma_launch(const float* qkv, const float* table, const float* out, const float* dout, const float* lse, float* delta,
                        float* dqkv, float* dkv_win, float* part, const GAGeom& g, cudaStream_t st);
constexpr int GA_KTILE_MMA = 64;  // keys per CTA of the tensor-core dK/dV kernel (more, smaller partial tables)
                           // (default 1: measured on B200 at C4 the tensor-core forward wins 9.8 vs 11.3 ms, the backward does not yet:
                           // 38.4 vs 36.0 ms - its tile loads and the fixed-point table atomics dominate, not the MMAs)

static int ga_make(GAGeom& g, int batch, int h, int w, int c, int heads, int ws, int ows, int shift, int use_mask, float scale,
                   const char* who) {
  NSR_CHECK_ARG(batch > 0 && h > 0 && w > 0 && c > 0 && heads > 0 && c % heads == 0, "%s: bad shape", who);
  NSR_CHECK_ARG(c / heads <= GA_D, "%s: head dim %d > %d", who, c / heads, GA_D);
  NSR_CHECK_ARG(ws > 0 && h % ws == 0 && w % ws == 0, "%s: %dx%d is not a multiple of window %d", who, h, w, ws);
  NSR_CHECK_ARG(ows >= ws && (ows - ws) % 2 == 0, "%s: overlap window %d vs window %d", who, ows, ws);
  NSR_CHECK_ARG((ws + ows - 1) * (ws + ows - 1) <= GA_MAXTAB, "%s: bias table too large", who);
  NSR_CHECK_ARG(ows == ws || (shift == 0 && !use_mask), "%s: overlapping windows take no shift/mask", who);
  NSR_CHECK_ARG(shift >= 0 && shift < ws, "%s: shift_size must in 0-window_size", who);
  g.B = batch; g.H = h; g.W = w; g.C = c; g.heads = heads; g.D = c / heads; g.ws = ws; g.ows = ows;
  g.pad = (ows - ws) / 2; g.shift = shift; g.use_mask = use_mask && shift > 0; g.oca = ows != ws;
  g.nwh = h / ws; g.nww = w / ws; g.Nq = ws * ws; g.Nk = ows * ows; g.L = ws + ows - 1; g.ntab = g.L * g.L;
  g.scale = scale;
  return NSR_OK;
}

// ------------------------------------------------------------------ CAB channel attention ---
// pooled[b,c] = mean over pixels of x[b,:,:,c] * (mul ? mul[b,:,:,c] : 1)   (NHWC; deterministic)
// (32 row lanes per CTA: with 8, a 8 x 4096-pixel x 180-channel call - HAT at B = 8 - ran 48 CTAs of 512 dependent
// iterations each, 40 us for 24 MB)
constexpr int CM_ROWS = 32;
__global__ void __launch_bounds__(CM_ROWS * 32) chan_mean_kernel(const float* __restrict__ x, const float* __restrict__ mul,
                                                                float* __restrict__ pooled, int HW, int C, float scale) {
  __shared__ float red[CM_ROWS][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), row = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int p = row; p < HW; p += CM_ROWS) {
      const size_t o = ((size_t)b * HW + p) * C + c;
      s += mul ? x[o] * mul[o] : x[o];
    }
  red[row][threadIdx.x & 31] = s;
  __syncthreads();
  if (row == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < CM_ROWS; ++r) t += red[r][threadIdx.x & 31];
    pooled[(size_t)b * C + c] = t * scale;
  }
}
// gate[b,:] = sigmoid(W2 relu(W1 pooled[b,:] + b1) + b2); hidden kept for the backward pass
__global__ void chan_gate_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ w1, const float* __restrict__ b1,
                                     const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ hidden,
                                     float* __restrict__ gate, int C, int Cs) {
  extern __shared__ float sh[];  // [Cs]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < Cs; j += blockDim.x) {
    float s = b1[j];
    for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)j * C + c], pooled[(size_t)b * C + c], s);
    s = fmaxf(s, 0.f);
    sh[j] = s;
    hidden[(size_t)b * Cs + j] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int j = 0; j < Cs; ++j) s = fmaf(w2[(size_t)c * Cs + j], sh[j], s);
    gate[(size_t)b * C + c] = 1.f / (1.f + expf(-s));
  }
}
// y[p,c] (+)= alpha * x[p,c] * gate[b,c]
__global__ void __launch_bounds__(256) chan_scale_add_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                             float* __restrict__ y, size_t total, int HW, int C, float alpha,
                                                             int accumulate) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)HW * C);
    const float v = alpha * x[i] * gate[b * C + c];
    y[i] = accumulate ? y[i] + v : v;
  }
}
// single CTA: back through sigmoid / W2 / ReLU / W1 for every sample in order (deterministic parameter grads)
// dgate[b,c] = dL/dgate;  outputs dpooled[b,c] and dW1, db1, dW2, db2 (overwrite).
__global__ void chan_gate_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate, const float* __restrict__ hidden,
                                     const float* __restrict__ pooled, const float* __restrict__ w1, const float* __restrict__ w2,
                                     float* __restrict__ dpooled, float* __restrict__ dw1, float* __restrict__ db1,
                                     float* __restrict__ dw2, float* __restrict__ db2, int B, int C, int Cs) {
  extern __shared__ float sh[];  // dz[C], dh[Cs]
  float* dz = sh;
  float* dh = sh + C;
  for (int i = threadIdx.x; i < C * Cs; i += blockDim.x) { dw1[i] = 0.f; dw2[i] = 0.f; }
  for (int i = threadIdx.x; i < Cs; i += blockDim.x) db1[i] = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) db2[i] = 0.f;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float gt = gate[(size_t)b * C + c];
      dz[c] = dgate[(size_t)b * C + c] * gt * (1.f - gt);
      db2[c] += dz[c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * Cs; i += blockDim.x) dw2[i] += dz[i / Cs] * hidden[(size_t)b * Cs + i % Cs];
    for (int j = threadIdx.x; j < Cs; j += blockDim.x) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(w2[(size_t)c * Cs + j], dz[c], s);
      s = hidden[(size_t)b * Cs + j] > 0.f ? s : 0.f;
      dh[j] = s;
      db1[j] += s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * Cs; i += blockDim.x) dw1[i] += dh[i / C] * pooled[(size_t)b * C + i % C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int j = 0; j < Cs; ++j) s = fmaf(w1[(size_t)j * C + c], dh[j], s);
      dpooled[(size_t)b * C + c] = s;
    }
    __syncthreads();
  }
}
// The same gradients with every sample staged in shared memory: three phases, each parallel over (sample, channel) or over
// the weight elements, sums over samples taken by the owning thread in sample order (same order as the serial kernel above,
// which walked the samples one by one with read-modify-writes of dW in global memory: 56 us for ~1000 weights).
__global__ void __launch_bounds__(256) chan_gate_bwd_staged(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                            const float* __restrict__ hidden, const float* __restrict__ pooled,
                                                            const float* __restrict__ w1, const float* __restrict__ w2,
                                                            float* __restrict__ dpooled, float* __restrict__ dw1,
                                                            float* __restrict__ db1, float* __restrict__ dw2,
                                                            float* __restrict__ db2, int B, int C, int Cs) {
  extern __shared__ float sh[];  // dz[B][C], pl[B][C], hd[B][Cs], dh[B][Cs]
  float* dz = sh;
  float* pl = dz + (size_t)B * C;
  float* hd = pl + (size_t)B * C;
  float* dh = hd + (size_t)B * Cs;
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const float gt = gate[i];
    dz[i] = dgate[i] * gt * (1.f - gt);
    pl[i] = pooled[i];
  }
  for (int i = threadIdx.x; i < B * Cs; i += blockDim.x) hd[i] = hidden[i];
  __syncthreads();
  for (int i = threadIdx.x; i < B * Cs; i += blockDim.x) {
    const int b = i / Cs, j = i - b * Cs;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w2[(size_t)c * Cs + j], dz[b * C + c], s);
    dh[i] = hd[i] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cs; i += blockDim.x) {
    const int c2 = i / Cs, j2 = i - c2 * Cs;  // dW2[c][j]
    const int j1 = i / C, c1 = i - j1 * C;    // dW1[j][c]
    float a2 = 0.f, a1 = 0.f;
    for (int b = 0; b < B; ++b) {
      a2 += dz[b * C + c2] * hd[b * Cs + j2];
      a1 += dh[b * Cs + j1] * pl[b * C + c1];
    }
    dw2[i] = a2;
    dw1[i] = a1;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dz[b * C + c];
    db2[c] = s;
  }
  for (int j = threadIdx.x; j < Cs; j += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dh[b * Cs + j];
    db1[j] = s;
  }
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int b = i / C, c = i - b * C;
    float s = 0.f;
    for (int j = 0; j < Cs; ++j) s = fmaf(w1[(size_t)j * C + c], dh[b * Cs + j], s);
    dpooled[i] = s;
  }
}
// dx[p,c] = alpha * g[p,c] * gate[b,c] + dpooled[b,c] / HW
__global__ void __launch_bounds__(256) chan_scale_bwd_kernel(const float* __restrict__ g, const float* __restrict__ gate,
                                                             const float* __restrict__ dpooled, float* __restrict__ dx,
                                                             size_t total, int HW, int C, float alpha) {
  const float inv = 1.f / (float)HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)HW * C);
    dx[i] = alpha * g[i] * gate[b * C + c] + dpooled[b * C + c] * inv;
  }
}
static inline int grid1(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
}  // namespace nsr
using namespace nsr;

extern "C" size_t nsr_xwin_attn_stat_floats(int batch, int h, int w, int heads, int ws) {
  return (size_t)batch * (h / ws) * (w / ws) * heads * ws * ws;
}
extern "C" int nsr_xwin_attn_fwd(const float* qkv, const float* bias_table, float* out, float* lse, int batch, int h, int w, int c,
                                 int heads, int ws, int ows, int shift, int use_mask, float scale, int engine, void* stream) {
  NSR_CHECK_ARG(qkv && bias_table && out && lse, "nsr_xwin_attn_fwd: null pointer");
  NSR_CHECK_ARG(engine == NSR_ENGINE_AUTO || engine == NSR_ENGINE_SIMT, "nsr_xwin_attn_fwd: engine must be NSR_ENGINE_AUTO or NSR_ENGINE_SIMT");
  GAGeom g;
  int rc = ga_make(g, batch, h, w, c, heads, ws, ows, shift, use_mask, scale, "nsr_xwin_attn_fwd");
  if (rc) return rc;
  if (engine != NSR_ENGINE_SIMT && xwin_attn_mma_supported(g)) return xwin_fwd_mma_launch(qkv, bias_table, out, lse, g, (cudaStream_t)stream);
  dim3 grid(batch * g.nwh * g.nww * heads, ceil_div(g.Nq, GA_THREADS));
  ga_fwd_kernel<<<grid, GA_THREADS, 0, (cudaStream_t)stream>>>(qkv, bias_table, out, lse, g);
  NSR_CHECK_LAUNCH("nsr_xwin_attn_fwd");
  return NSR_OK;
}
extern "C" size_t nsr_xwin_attn_bwd_workspace(int batch, int h, int w, int c, int heads, int ws, int ows) {
  const size_t nwin = (size_t)batch * (h / ws) * (w / ws);
  const size_t ntab = (size_t)(ws + ows - 1) * (ws + ows - 1), ktiles = (size_t)ceil_div(ows * ows, GA_KTILE_MMA);
  size_t fl = nwin * heads * ws * ws;              // delta
  fl += ktiles * nwin * heads * ntab;              // bias-table partials
  if (ows != ws) fl += nwin * ows * ows * 2 * c;   // per-window dk/dv
  return fl * sizeof(float);
}
extern "C" int nsr_xwin_attn_bwd(const float* qkv, const float* bias_table, const float* out, const float* dout, const float* lse,
                                 float* dqkv, float* dbias_table, int batch, int h, int w, int c, int heads, int ws, int ows,
                                 int shift, int use_mask, float scale, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(qkv && bias_table && out && dout && lse && dqkv && dbias_table, "nsr_xwin_attn_bwd: null pointer");
  NSR_CHECK_ARG(engine == NSR_ENGINE_AUTO || engine == NSR_ENGINE_SIMT, "nsr_xwin_attn_bwd: engine must be NSR_ENGINE_AUTO or NSR_ENGINE_SIMT");
  GAGeom g;
  int rc = ga_make(g, batch, h, w, c, heads, ws, ows, shift, use_mask, scale, "nsr_xwin_attn_bwd");
  if (rc) return rc;
  if (!workspace || workspace_bytes < nsr_xwin_attn_bwd_workspace(batch, h, w, c, heads, ws, ows)) {
    set_error("nsr_xwin_attn_bwd: workspace too small");
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = engine != NSR_ENGINE_SIMT && xwin_attn_mma_supported(g);
  const int nwin = batch * g.nwh * g.nww, ktiles = ceil_div(g.Nk, tc ? GA_KTILE_MMA : GA_THREADS);
  float* delta = (float*)workspace;
  float* part = delta + (size_t)nwin * heads * g.Nq;
  float* dkv_win = g.oca ? part + (size_t)ceil_div(g.Nk, GA_KTILE_MMA) * nwin * heads * g.ntab : nullptr;
  if (tc) {
    rc = xwin_bwd_mma_launch(qkv, bias_table, out, dout, lse, delta, dqkv, dkv_win, part, g, st);
    if (rc) return rc;
  } else {
    ga_bwd_q_kernel<<<dim3(nwin * heads, ceil_div(g.Nq, GA_THREADS)), GA_THREADS, 0, st>>>(qkv, bias_table, out, dout, lse, delta, dqkv, g);
    NSR_CHECK_LAUNCH("nsr_xwin_attn_bwd(q)");
    ga_bwd_kv_kernel<<<dim3(nwin * heads, ktiles), GA_THREADS, 0, st>>>(qkv, bias_table, dout, lse, delta, dqkv, dkv_win, part, g);
    NSR_CHECK_LAUNCH("nsr_xwin_attn_bwd(kv)");
  }
  if (g.oca) {
    ga_fold_kernel<<<grid1((size_t)batch * h * w * 2 * c), 256, 0, st>>>(dkv_win, dqkv, g);
    NSR_CHECK_LAUNCH("nsr_xwin_attn_bwd(fold)");
  }
  ga_dtab_reduce_kernel<<<ceil_div(g.ntab * heads, 128), 128, 0, st>>>(part, dbias_table, nwin, ktiles, g);
  NSR_CHECK_LAUNCH("nsr_xwin_attn_bwd(table)");
  return NSR_OK;
}

extern "C" int nsr_channel_mean(const float* x, const float* mul, float* pooled, int batch, int hw, int c, float scale, void* stream) {
  NSR_CHECK_ARG(x && pooled && batch > 0 && hw > 0 && c > 0 && batch <= 65535, "nsr_channel_mean: bad arguments");
  chan_mean_kernel<<<dim3(ceil_div(c, 32), batch), CM_ROWS * 32, 0, (cudaStream_t)stream>>>(x, mul, pooled, hw, c, scale);
  NSR_CHECK_LAUNCH("nsr_channel_mean");
  return NSR_OK;
}
extern "C" int nsr_channel_gate_fwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                                    float* hidden, float* gate, int batch, int c, int cs, void* stream) {
  NSR_CHECK_ARG(pooled && w1 && b1 && w2 && b2 && hidden && gate && batch > 0 && c > 0 && cs > 0 && cs <= 4096,
                "nsr_channel_gate_fwd: bad arguments");
  chan_gate_fwd_kernel<<<batch, 128, cs * sizeof(float), (cudaStream_t)stream>>>(pooled, w1, b1, w2, b2, hidden, gate, c, cs);
  NSR_CHECK_LAUNCH("nsr_channel_gate_fwd");
  return NSR_OK;
}
extern "C" int nsr_channel_scale_add(const float* x, const float* gate, float* y, int batch, int hw, int c, float alpha,
                                     int accumulate, void* stream) {
  NSR_CHECK_ARG(x && gate && y && batch > 0 && hw > 0 && c > 0, "nsr_channel_scale_add: bad arguments");
  const size_t total = (size_t)batch * hw * c;
  chan_scale_add_kernel<<<grid1(total), 256, 0, (cudaStream_t)stream>>>(x, gate, y, total, hw, c, alpha, accumulate);
  NSR_CHECK_LAUNCH("nsr_channel_scale_add");
  return NSR_OK;
}
extern "C" int nsr_channel_gate_bwd(const float* dgate, const float* gate, const float* hidden, const float* pooled, const float* w1,
                                    const float* w2, float* dpooled, float* dw1, float* db1, float* dw2, float* db2, int batch,
                                    int c, int cs, void* stream) {
  NSR_CHECK_ARG(dgate && gate && hidden && pooled && w1 && w2 && dpooled && dw1 && db1 && dw2 && db2 && batch > 0 &&
                    (size_t)(c + cs) * sizeof(float) <= 48 * 1024, "nsr_channel_gate_bwd: bad arguments");
  const size_t staged = (size_t)batch * 2 * (c + cs) * sizeof(float);
  if (staged <= 48 * 1024)
    chan_gate_bwd_staged<<<1, 256, staged, (cudaStream_t)stream>>>(dgate, gate, hidden, pooled, w1, w2, dpooled, dw1, db1, dw2, db2,
                                                                   batch, c, cs);
  else
    chan_gate_bwd_kernel<<<1, 256, (c + cs) * sizeof(float), (cudaStream_t)stream>>>(dgate, gate, hidden, pooled, w1, w2, dpooled, dw1,
                                                                                     db1, dw2, db2, batch, c, cs);
  NSR_CHECK_LAUNCH("nsr_channel_gate_bwd");
  return NSR_OK;
}
extern "C" int nsr_channel_scale_bwd(const float* g, const float* gate, const float* dpooled, float* dx, int batch, int hw, int c,
                                     float alpha, void* stream) {
  NSR_CHECK_ARG(g && gate && dpooled && dx && batch > 0 && hw > 0 && c > 0, "nsr_channel_scale_bwd: bad arguments");
  const size_t total = (size_t)batch * hw * c;
  chan_scale_bwd_kernel<<<grid1(total), 256, 0, (cudaStream_t)stream>>>(g, gate, dpooled, dx, total, hw, c, alpha);
  NSR_CHECK_LAUNCH("nsr_channel_scale_bwd");
  return NSR_OK;
}

// Generic (cross-)window attention on the tensor cores: HAT's HAB self-attention (16x16 windows, 256 x 256 scores per
// head) and OCAB overlapping cross-attention (256 queries x 576 zero-padded keys), hat_arch.py:168-215, 445-515.
// Same scheme as the 8x8-window kernels (window_attn_mma.cu): warp-level mma.sync m16n8k16, every operand split into
// bf16 hi + lo and multiplied in three passes with fp32 accumulation; the score tile lives in registers, flash-style
// (running max / sum per query row), and the backward pass recomputes it from the saved log-sum-exp.
//   forward     : CTA = (window, head, 64-query tile), 4 warps x 16 rows, loop over 64-key chunks
//   backward dQ : same tiling; dS = P o (dP - delta), dQ += dS K
//   backward dKV: CTA = (window, head, 64-key tile), loop over 64-query chunks on the TRANSPOSED problem
//                 (S^T = K Qs^T), dK += dS^T Qs, dV += P^T dO; the bias-table gradient is accumulated with 64-bit
//                 fixed-point shared-memory atomics (integer adds commute: deterministic), per-CTA partials reduced in
//                 a fixed order.
#include "attn_mma.cuh"
#include "xwin_geom.cuh"

namespace nsr {

constexpr int XM_T = 64;        // tile edge (queries / keys per step)
constexpr int XM_THREADS = 128;
constexpr int XM_MAXTAB = 39 * 39;
constexpr float XM_FIX = 1099511627776.0f;  // 2^40: fixed-point scale of the bias-table gradient accumulators

// rows n = 0..63 -> tokens tokv[n] (-1: zero row); D channels at column offset coff of a [tokens, ld] fp32 tensor, times mul;
// written as bf16 hi/lo [n][AM_LD] tiles.  16 lanes per row (D / 2 <= 16 float2 pairs), 8 rows per pass.
__device__ __forceinline__ void xm_load_tile(const float* __restrict__ base, size_t ld, int coff, const int* tokv, int D, float mul,
                                             __nv_bfloat16* Th, __nv_bfloat16* Tl, int t) {
  const int pr = t & 15, rbase = t >> 4;
  if (pr >= (D >> 1)) return;
  float2 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int tk = tokv[rbase + 8 * u];
    v[u] = make_float2(0.f, 0.f);
    if (tk >= 0) v[u] = *reinterpret_cast<const float2*>(base + (size_t)tk * ld + coff + 2 * pr);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    uint32_t hi, lo;
    split_pair(v[u].x * mul, v[u].y * mul, hi, lo);
    const int o1 = (rbase + 8 * u) * AM_LD + 2 * pr;
    *reinterpret_cast<uint32_t*>(&Th[o1]) = hi;
    *reinterpret_cast<uint32_t*>(&Tl[o1]) = lo;
  }
}
__device__ __forceinline__ void xm_zero(__nv_bfloat16* p, int n, int t) {
  for (int i = t; i < n / 2; i += XM_THREADS) reinterpret_cast<uint32_t*>(p)[i] = 0u;
}

struct XmMeta {  // per-token metadata of a 64-token tile
  int tok[XM_T], rid[XM_T], y[XM_T], x[XM_T];
};
__device__ __forceinline__ void xm_query_meta(const GAGeom& g, int wi, int q0, XmMeta& m, int t) {
  if (t < XM_T) {
    int tk = -1, r = -1, iy = 0, ix = 0;
    if (q0 + t < g.Nq) ga_query(g, wi, q0 + t, tk, r, iy, ix);
    m.tok[t] = tk; m.rid[t] = r; m.y[t] = iy; m.x[t] = ix;
  }
}
__device__ __forceinline__ void xm_key_meta(const GAGeom& g, int wi, int k0, XmMeta& m, int t) {
  if (t < XM_T) {
    int tk = -1, r = -2, jy = 0, jx = 0;  // rid -2: beyond Nk (excluded); a zero-padded key of nn.Unfold keeps tok -1, rid 0
    if (k0 + t < g.Nk) ga_key(g, wi, k0 + t, tk, r, jy, jx);
    m.tok[t] = tk; m.rid[t] = r; m.y[t] = jy; m.x[t] = jx;
  }
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(XM_THREADS) xwin_fwd_mma(const float* __restrict__ qkv, const float* __restrict__ table,
                                                           float* __restrict__ out, float* __restrict__ lse, GAGeom gm) {
  __shared__ __align__(16) __nv_bfloat16 Qh[XM_T * AM_LD], Ql[XM_T * AM_LD], Kh[XM_T * AM_LD], Kl[XM_T * AM_LD];
  __shared__ __align__(16) __nv_bfloat16 Vh[XM_T * AM_LD], Vl[XM_T * AM_LD];
  __shared__ float tab[XM_MAXTAB];
  __shared__ XmMeta qm, km;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int wi = blockIdx.x / gm.heads, head = blockIdx.x - wi * gm.heads, q0 = blockIdx.y * XM_T;
  const int row0 = warp * 16;
  for (int e = t; e < gm.ntab; e += XM_THREADS) tab[e] = table[(size_t)e * gm.heads + head];
  xm_zero(Qh, XM_T * AM_LD, t); xm_zero(Ql, XM_T * AM_LD, t); xm_zero(Kh, XM_T * AM_LD, t); xm_zero(Kl, XM_T * AM_LD, t);
  xm_zero(Vh, XM_T * AM_LD, t); xm_zero(Vl, XM_T * AM_LD, t);
  xm_query_meta(gm, wi, q0, qm, t);
  __syncthreads();
  xm_load_tile(qkv, (size_t)3 * gm.C, head * gm.D, qm.tok, gm.D, gm.scale, Qh, Ql, t);
  float o[4][4], mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  int qy[2], qx[2], qr[2];
  for (int k0 = 0; k0 < gm.Nk; k0 += XM_T) {
    __syncthreads();  // previous chunk consumed (first pass: Q tile / metadata written)
    xm_key_meta(gm, wi, k0, km, t);
    __syncthreads();
    xm_load_tile(qkv, (size_t)3 * gm.C, gm.C + head * gm.D, km.tok, gm.D, 1.f, Kh, Kl, t);
    xm_load_tile(qkv, (size_t)3 * gm.C, 2 * gm.C + head * gm.D, km.tok, gm.D, 1.f, Vh, Vl, t);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) { qy[h] = qm.y[row0 + g + 8 * h]; qx[h] = qm.x[row0 + g + 8 * h]; qr[h] = qm.rid[row0 + g + 8 * h]; }
    float s[8][4];
    qk_scores_ldm(Qh, Ql, Kh, Kl, row0, lane, s);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float cm = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + tid * 2 + e;
          const int kr = km.rid[j];
          float v = -INFINITY;
          if (kr != -2) {
            v = s[nt][2 * h + e] + tab[ga_rel(gm, qy[h], qx[h], km.y[j], km.x[j])];
            if (gm.use_mask && kr != qr[h]) v += -100.f;
          }
          s[nt][2 * h + e] = v;
          cm = fmaxf(cm, v);
        }
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
      const float mn = fmaxf(mrun[h], cm), corr = __expf(mrun[h] - mn);
      mrun[h] = mn;
      float ps = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p = __expf(s[nt][2 * h + e] - mn);
          s[nt][2 * h + e] = p;
          ps += p;
        }
      lrun[h] = lrun[h] * corr + ps;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { o[nt][2 * h] *= corr; o[nt][2 * h + 1] *= corr; }
    }
    float oc[4][4];
    acc_times_ldm(s, Vh, Vl, lane, oc);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nt][e] += oc[nt][e];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float l = lrun[h];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int i = row0 + g + 8 * h, tk = qm.tok[i];
    if (tk < 0) continue;
    const float inv = 1.f / l;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + tid * 2;
      if (c < gm.D) *reinterpret_cast<float2*>(out + (size_t)tk * gm.C + head * gm.D + c) = make_float2(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
    }
    if (tid == 0) lse[((size_t)wi * gm.heads + head) * gm.Nq + q0 + i] = mrun[h] + logf(l);
  }
}

// ------------------------------------------------------------------ backward: dQ
// dynamic smem: Qh Ql Oh Ol Kh Kl Vh Vl [64][AM_LD] | tab | metas | delta
constexpr size_t XQ_SMEM = (size_t)(8 * XM_T * AM_LD) * sizeof(__nv_bfloat16) + XM_MAXTAB * sizeof(float) +
                           2 * sizeof(XmMeta) + 2 * XM_T * sizeof(float);
__global__ void __launch_bounds__(XM_THREADS) xwin_bwd_q_mma(const float* __restrict__ qkv, const float* __restrict__ table,
                                                             const float* __restrict__ out, const float* __restrict__ dout,
                                                             const float* __restrict__ lse, float* __restrict__ delta,
                                                             float* __restrict__ dqkv, GAGeom gm) {
  extern __shared__ __align__(16) uint8_t xsm[];
  __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(xsm);
  __nv_bfloat16 *Ql = Qh + XM_T * AM_LD, *Oh = Ql + XM_T * AM_LD, *Ol = Oh + XM_T * AM_LD, *Kh = Ol + XM_T * AM_LD;
  __nv_bfloat16 *Kl = Kh + XM_T * AM_LD, *Vh = Kl + XM_T * AM_LD, *Vl = Vh + XM_T * AM_LD;
  float* tab = reinterpret_cast<float*>(Vl + XM_T * AM_LD);
  XmMeta* qm = reinterpret_cast<XmMeta*>(tab + XM_MAXTAB);
  XmMeta* km = qm + 1;
  float* lse_s = reinterpret_cast<float*>(km + 1);
  float* del_s = lse_s + XM_T;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int wi = blockIdx.x / gm.heads, head = blockIdx.x - wi * gm.heads, q0 = blockIdx.y * XM_T;
  const int row0 = warp * 16;
  for (int e = t; e < gm.ntab; e += XM_THREADS) tab[e] = table[(size_t)e * gm.heads + head];
  xm_zero(Qh, 8 * XM_T * AM_LD, t);
  xm_query_meta(gm, wi, q0, *qm, t);
  __syncthreads();
  if (t < XM_T) {
    const int tk = qm->tok[t];
    float dl = 0.f, L = 0.f;
    if (tk >= 0) {
      const float* op = out + (size_t)tk * gm.C + head * gm.D;
      const float* gp = dout + (size_t)tk * gm.C + head * gm.D;
      for (int d = 0; d < gm.D; ++d) dl = fmaf(gp[d], op[d], dl);
      const size_t si = ((size_t)wi * gm.heads + head) * gm.Nq + q0 + t;
      L = lse[si];
      delta[si] = dl;
    }
    lse_s[t] = L;
    del_s[t] = dl;
  }
  xm_load_tile(qkv, (size_t)3 * gm.C, head * gm.D, qm->tok, gm.D, gm.scale, Qh, Ql, t);
  xm_load_tile(dout, (size_t)gm.C, head * gm.D, qm->tok, gm.D, 1.f, Oh, Ol, t);
  float dq[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
  for (int k0 = 0; k0 < gm.Nk; k0 += XM_T) {
    __syncthreads();
    xm_key_meta(gm, wi, k0, *km, t);
    __syncthreads();
    xm_load_tile(qkv, (size_t)3 * gm.C, gm.C + head * gm.D, km->tok, gm.D, 1.f, Kh, Kl, t);
    xm_load_tile(qkv, (size_t)3 * gm.C, 2 * gm.C + head * gm.D, km->tok, gm.D, 1.f, Vh, Vl, t);
    __syncthreads();
    float s[8][4], dp[8][4];
    qk_scores_ldm(Qh, Ql, Kh, Kl, row0, lane, s);
    qk_scores_ldm(Oh, Ol, Vh, Vl, row0, lane, dp);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = row0 + g + 8 * h;
      const int qy = qm->y[i], qx = qm->x[i], qr = qm->rid[i];
      const float L = lse_s[i], dl = del_s[i];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + tid * 2 + e;
          const int kr = km->rid[j];
          float ds = 0.f;
          if (kr != -2) {
            float v = s[nt][2 * h + e] + tab[ga_rel(gm, qy, qx, km->y[j], km->x[j])];
            if (gm.use_mask && kr != qr) v += -100.f;
            ds = __expf(v - L) * (dp[nt][2 * h + e] - dl);
          }
          s[nt][2 * h + e] = ds;
        }
    }
    float oc[4][4];
    acc_times_ldm(s, Kh, Kl, lane, oc);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nt][e] += oc[nt][e];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int tk = qm->tok[row0 + g + 8 * h];
    if (tk < 0) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + tid * 2;
      if (c < gm.D)
        *reinterpret_cast<float2*>(dqkv + (size_t)tk * 3 * gm.C + head * gm.D + c) = make_float2(dq[nt][2 * h] * gm.scale, dq[nt][2 * h + 1] * gm.scale);
    }
  }
}

// ------------------------------------------------------------------ backward: dK, dV, bias-table gradient
// dynamic smem: Kh Kl Vh Vl Qh Ql Oh Ol [64][AM_LD] | dtab64 | tab | metas | lse, delta
constexpr size_t XK_SMEM = (size_t)(8 * XM_T * AM_LD) * sizeof(__nv_bfloat16) + XM_MAXTAB * (sizeof(float) + 8) +
                           2 * sizeof(XmMeta) + 2 * XM_T * sizeof(float) + 16;
__global__ void __launch_bounds__(XM_THREADS) xwin_bwd_kv_mma(const float* __restrict__ qkv, const float* __restrict__ table,
                                                              const float* __restrict__ dout, const float* __restrict__ lse,
                                                              const float* __restrict__ delta, float* __restrict__ dqkv,
                                                              float* __restrict__ dkv_win, float* __restrict__ dtab_part, GAGeom gm) {
  extern __shared__ __align__(16) uint8_t xsm[];
  unsigned long long* dtab64 = reinterpret_cast<unsigned long long*>(xsm);
  __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(dtab64 + ((XM_MAXTAB + 1) & ~1));  // 16-byte aligned (ldmatrix rows)
  __nv_bfloat16 *Kl = Kh + XM_T * AM_LD, *Vh = Kl + XM_T * AM_LD, *Vl = Vh + XM_T * AM_LD, *Qh = Vl + XM_T * AM_LD;
  __nv_bfloat16 *Ql = Qh + XM_T * AM_LD, *Oh = Ql + XM_T * AM_LD, *Ol = Oh + XM_T * AM_LD;
  float* tab = reinterpret_cast<float*>(Ol + XM_T * AM_LD);
  XmMeta* km = reinterpret_cast<XmMeta*>(tab + XM_MAXTAB);
  XmMeta* qm = km + 1;
  float* lse_s = reinterpret_cast<float*>(qm + 1);
  float* del_s = lse_s + XM_T;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int wi = blockIdx.x / gm.heads, head = blockIdx.x - wi * gm.heads, k0 = blockIdx.y * XM_T;
  const int row0 = warp * 16;
  for (int e = t; e < gm.ntab; e += XM_THREADS) { tab[e] = table[(size_t)e * gm.heads + head]; dtab64[e] = 0ull; }
  xm_zero(Kh, 8 * XM_T * AM_LD, t);
  xm_key_meta(gm, wi, k0, *km, t);
  __syncthreads();
  xm_load_tile(qkv, (size_t)3 * gm.C, gm.C + head * gm.D, km->tok, gm.D, 1.f, Kh, Kl, t);
  xm_load_tile(qkv, (size_t)3 * gm.C, 2 * gm.C + head * gm.D, km->tok, gm.D, 1.f, Vh, Vl, t);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[nt][e] = 0.f; dv[nt][e] = 0.f; }
  for (int q0 = 0; q0 < gm.Nq; q0 += XM_T) {
    __syncthreads();
    xm_query_meta(gm, wi, q0, *qm, t);
    if (t < XM_T) {
      const bool ok = q0 + t < gm.Nq;
      const size_t si = ((size_t)wi * gm.heads + head) * gm.Nq + q0 + t;
      lse_s[t] = ok ? lse[si] : 0.f;
      del_s[t] = ok ? delta[si] : 0.f;
    }
    __syncthreads();
    xm_load_tile(qkv, (size_t)3 * gm.C, head * gm.D, qm->tok, gm.D, gm.scale, Qh, Ql, t);
    xm_load_tile(dout, (size_t)gm.C, head * gm.D, qm->tok, gm.D, 1.f, Oh, Ol, t);
    __syncthreads();
    float st[8][4], dpt[8][4];
    qk_scores_ldm(Kh, Kl, Qh, Ql, row0, lane, st);    // S^T[j][i]  = K_j . Qs_i
    qk_scores_ldm(Vh, Vl, Oh, Ol, row0, lane, dpt);   // dP^T[j][i] = V_j . dO_i
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = row0 + g + 8 * h;
      const int ky = km->y[j], kx = km->x[j], kr = km->rid[j];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = nt * 8 + tid * 2 + e;
          float p = 0.f, ds = 0.f;
          if (kr != -2 && qm->tok[i] >= 0) {
            const int ti = ga_rel(gm, qm->y[i], qm->x[i], ky, kx);
            float v = st[nt][2 * h + e] + tab[ti];
            if (gm.use_mask && kr != qm->rid[i]) v += -100.f;
            p = __expf(v - lse_s[i]);
            ds = p * (dpt[nt][2 * h + e] - del_s[i]);
            atomicAdd(&dtab64[ti], (unsigned long long)__float2ll_rn(ds * XM_FIX));
          }
          st[nt][2 * h + e] = ds;
          dpt[nt][2 * h + e] = p;
        }
    }
    float oc[4][4];
    acc_times_ldm(st, Qh, Ql, lane, oc);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[nt][e] += oc[nt][e];
    acc_times_ldm(dpt, Oh, Ol, lane, oc);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dv[nt][e] += oc[nt][e];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = row0 + g + 8 * h;
    if (km->rid[j] == -2) continue;
    const int tk = km->tok[j];
    float* o = nullptr;
    if (dkv_win) o = dkv_win + (((size_t)wi * gm.Nk + k0 + j) * 2) * gm.C + head * gm.D;
    else if (tk >= 0) o = dqkv + (size_t)tk * 3 * gm.C + gm.C + head * gm.D;
    if (!o) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + tid * 2;
      if (c < gm.D) {
        *reinterpret_cast<float2*>(o + c) = make_float2(dk[nt][2 * h], dk[nt][2 * h + 1]);
        *reinterpret_cast<float2*>(o + gm.C + c) = make_float2(dv[nt][2 * h], dv[nt][2 * h + 1]);
      }
    }
  }
  __syncthreads();
  float* part = dtab_part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * gm.ntab;
  for (int e = t; e < gm.ntab; e += XM_THREADS) part[e] = (float)((double)(long long)dtab64[e] * (1.0 / (double)XM_FIX));
}

int xwin_bwd_mma_launch(const float* qkv, const float* table, const float* out, const float* dout, const float* lse, float* delta,
                        float* dqkv, float* dkv_win, float* part, const GAGeom& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(xwin_bwd_q_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xwin_bwd_kv_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XK_SMEM);
    if (e != cudaSuccess) {
      set_error("xwin_bwd_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int nwh = g.B * g.nwh * g.nww * g.heads;
  xwin_bwd_q_mma<<<dim3(nwh, ceil_div(g.Nq, XM_T)), XM_THREADS, XQ_SMEM, st>>>(qkv, table, out, dout, lse, delta, dqkv, g);
  NSR_CHECK_LAUNCH("xwin_bwd_q_mma");
  xwin_bwd_kv_mma<<<dim3(nwh, ceil_div(g.Nk, XM_T)), XM_THREADS, XK_SMEM, st>>>(qkv, table, dout, lse, delta, dqkv, dkv_win, part, g);
  NSR_CHECK_LAUNCH("xwin_bwd_kv_mma");
  return NSR_OK;
}

bool xwin_attn_mma_supported(const GAGeom& g) { return g.D % 2 == 0 && g.D <= 32 && g.C % 2 == 0 && g.ntab <= XM_MAXTAB; }

int xwin_fwd_mma_launch(const float* qkv, const float* table, float* out, float* lse, const GAGeom& g, cudaStream_t st) {
  dim3 grid(g.B * g.nwh * g.nww * g.heads, ceil_div(g.Nq, XM_T));
  xwin_fwd_mma<<<grid, XM_THREADS, 0, st>>>(qkv, table, out, lse, g);
  NSR_CHECK_LAUNCH("xwin_fwd_mma");
  return NSR_OK;
}

}  // namespace nsr

// Fused (shifted-)window attention core, fp32 CUDA-core version.
//
// One CTA = one (window, head).  The cyclic shift, window partition/reverse and the {0,-100}
// shift mask are pure index math on token coordinates (nothing is materialised); the relative
// position index is computed arithmetically ((iy-jy+ws-1)*(2ws-1) + (ix-jx+ws-1), identical to
// the reference's registered buffer, swinir_arch.py:120-137).  Softmax uses warp shuffles.
// Backward recomputes P, and reduces the bias-table gradient deterministically.
#include <stdlib.h>

#include "common.cuh"

namespace nsr {

constexpr int WA_N = 64;    // max tokens per window
constexpr int WA_D = 32;    // max head dim
constexpr int WA_LD = 33;   // padded row stride for [N][D] tiles
constexpr int WA_PLD = 65;  // padded row stride for [N][N] tiles
constexpr int WA_THREADS = 128;

struct WinGeom {
  int B, H, W, C, heads, ws, shift, use_mask, N, D, nwh, nww;
  float scale;
};

__device__ __forceinline__ void win_token_map(const WinGeom& g, int wi, int n, int& tok, int& rid) {
  const int per = g.nwh * g.nww;
  const int b = wi / per, rem = wi - b * per;
  const int wy = rem / g.nww, wx = rem - wy * g.nww;
  const int iy = n / g.ws, ix = n - iy * g.ws;
  const int hs = wy * g.ws + iy, wsx = wx * g.ws + ix;  // coordinates in the SHIFTED image
  int ho = hs + g.shift, wo = wsx + g.shift;            // torch.roll(x, -shift): shifted[i] = x[(i+shift) % n]
  if (ho >= g.H) ho -= g.H;
  if (wo >= g.W) wo -= g.W;
  tok = (b * g.H + ho) * g.W + wo;
  const int rh = hs < g.H - g.ws ? 0 : (hs < g.H - g.shift ? 1 : 2);  // calculate_mask slices
  const int rw = wsx < g.W - g.ws ? 0 : (wsx < g.W - g.shift ? 1 : 2);
  rid = rh * 3 + rw;
}

__device__ __forceinline__ int rel_index(int ws, int i, int j) {
  const int iy = i / ws, ix = i - iy * ws, jy = j / ws, jx = j - jy * ws;
  return (iy - jy + ws - 1) * (2 * ws - 1) + (ix - jx + ws - 1);
}

// scores for rows [wp*16, wp*16+16) x cols {lane, lane+32}: acc = A[i][:] . Bm[j][:]
__device__ __forceinline__ void rows_dot(const float* __restrict__ A, const float* __restrict__ Bm, int N, int D,
                                         int wp, int lane, float (&acc)[16][2]) {
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r][0] = acc[r][1] = 0.f;
  const bool has1 = lane + 32 < N;
  const bool has0 = lane < N;
  for (int d = 0; d < D; ++d) {
    const float b0 = has0 ? Bm[lane * WA_LD + d] : 0.f;
    const float b1 = has1 ? Bm[(lane + 32) * WA_LD + d] : 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float a = A[(wp * 16 + r) * WA_LD + d];
      acc[r][0] = fmaf(a, b0, acc[r][0]);
      acc[r][1] = fmaf(a, b1, acc[r][1]);
    }
  }
}

// softmax over the 2-per-lane row fragments, in place; masked-out columns (>= N) get 0.
__device__ __forceinline__ void rows_bias_softmax(const WinGeom& g, const float* __restrict__ bias_s,
                                                  const int* __restrict__ rid, int wp, int lane,
                                                  float (&acc)[16][2]) {
  const bool masked = g.use_mask && g.shift > 0;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int i = wp * 16 + r;
    if (i >= g.N) { acc[r][0] = acc[r][1] = 0.f; continue; }  // warp-uniform
    float s[2];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int j = lane + 32 * cc;
      if (j < g.N) {
        float v = acc[r][cc] + bias_s[rel_index(g.ws, i, j)];
        if (masked && rid[i] != rid[j]) v += -100.0f;
        s[cc] = v;
      } else {
        s[cc] = -INFINITY;
      }
    }
    const float m = warp_max(fmaxf(s[0], s[1]));
    const float e0 = s[0] == -INFINITY ? 0.f : expf(s[0] - m);
    const float e1 = s[1] == -INFINITY ? 0.f : expf(s[1] - m);
    const float inv = 1.f / warp_sum(e0 + e1);
    acc[r][0] = e0 * inv;
    acc[r][1] = e1 * inv;
  }
}

__global__ void __launch_bounds__(WA_THREADS) window_attn_fwd_kernel(const float* __restrict__ qkv,
                                                                    const float* __restrict__ table,
                                                                    float* __restrict__ out, WinGeom g) {
  __shared__ float Qs[WA_N * WA_LD], Ks[WA_N * WA_LD], Vs[WA_N * WA_LD], Ps[WA_N * WA_PLD];
  __shared__ float bias_s[(2 * 8 - 1) * (2 * 8 - 1)];
  __shared__ int tok[WA_N], rid[WA_N];
  const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
  const int wi = blockIdx.x, head = blockIdx.y;
  if (t < g.N) win_token_map(g, wi, t, tok[t], rid[t]);
  for (int i = t; i < (2 * g.ws - 1) * (2 * g.ws - 1); i += WA_THREADS) bias_s[i] = table[i * g.heads + head];
  for (int i = t; i < WA_N * WA_LD; i += WA_THREADS) { Qs[i] = 0.f; Ks[i] = 0.f; Vs[i] = 0.f; }
  __syncthreads();
  for (int idx = t; idx < g.N * g.D; idx += WA_THREADS) {
    const int n = idx / g.D, d = idx - n * g.D;
    const float* p = qkv + (size_t)tok[n] * 3 * g.C + head * g.D + d;
    Qs[n * WA_LD + d] = p[0] * g.scale;
    Ks[n * WA_LD + d] = p[g.C];
    Vs[n * WA_LD + d] = p[2 * g.C];
  }
  __syncthreads();
  float acc[16][2];
  rows_dot(Qs, Ks, g.N, g.D, wp, lane, acc);
  rows_bias_softmax(g, bias_s, rid, wp, lane, acc);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    Ps[(wp * 16 + r) * WA_PLD + lane] = acc[r][0];
    Ps[(wp * 16 + r) * WA_PLD + lane + 32] = acc[r][1];
  }
  __syncwarp();
  if (lane < g.D) {
    for (int r = 0; r < 16; ++r) {
      const int i = wp * 16 + r;
      if (i >= g.N) break;
      float o = 0.f;
      for (int j = 0; j < g.N; ++j) o = fmaf(Ps[i * WA_PLD + j], Vs[j * WA_LD + lane], o);
      out[(size_t)tok[i] * g.C + head * g.D + lane] = o;
    }
  }
}

constexpr size_t WA_BWD_SMEM = (size_t)(4 * WA_N * WA_LD + 2 * WA_N * WA_PLD + WA_N * WA_N) * sizeof(float);

__global__ void __launch_bounds__(WA_THREADS) window_attn_bwd_kernel(const float* __restrict__ qkv,
                                                                    const float* __restrict__ table,
                                                                    const float* __restrict__ dout,
                                                                    float* __restrict__ dqkv,
                                                                    float* __restrict__ partial, WinGeom g, int nwin) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + WA_N * WA_LD;
  float* Vs = Ks + WA_N * WA_LD;
  float* Os = Vs + WA_N * WA_LD;   // dO
  float* Ps = Os + WA_N * WA_LD;
  float* Ss = Ps + WA_N * WA_PLD;  // dS
  float* Acc = Ss + WA_N * WA_PLD; // sum over this CTA's windows of dS, [N][N] (stride WA_N)
  __shared__ float bias_s[(2 * 8 - 1) * (2 * 8 - 1)];
  __shared__ int tok[WA_N], rid[WA_N];
  const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
  const int head = blockIdx.y;
  for (int i = t; i < (2 * g.ws - 1) * (2 * g.ws - 1); i += WA_THREADS) bias_s[i] = table[i * g.heads + head];
  for (int i = t; i < WA_N * WA_N; i += WA_THREADS) Acc[i] = 0.f;
  for (int i = t; i < 4 * WA_N * WA_LD; i += WA_THREADS) sm[i] = 0.f;

  for (int wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    __syncthreads();
    if (t < g.N) win_token_map(g, wi, t, tok[t], rid[t]);
    __syncthreads();
    for (int idx = t; idx < g.N * g.D; idx += WA_THREADS) {
      const int n = idx / g.D, d = idx - n * g.D;
      const float* p = qkv + (size_t)tok[n] * 3 * g.C + head * g.D + d;
      Qs[n * WA_LD + d] = p[0] * g.scale;
      Ks[n * WA_LD + d] = p[g.C];
      Vs[n * WA_LD + d] = p[2 * g.C];
      Os[n * WA_LD + d] = dout[(size_t)tok[n] * g.C + head * g.D + d];
    }
    __syncthreads();
    float pr[16][2], dp[16][2];
    rows_dot(Qs, Ks, g.N, g.D, wp, lane, pr);
    rows_bias_softmax(g, bias_s, rid, wp, lane, pr);
    rows_dot(Os, Vs, g.N, g.D, wp, lane, dp);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = wp * 16 + r;
      const float delta = warp_sum(pr[r][0] * dp[r][0] + pr[r][1] * dp[r][1]);
      const float s0 = pr[r][0] * (dp[r][0] - delta), s1 = pr[r][1] * (dp[r][1] - delta);
      Ps[i * WA_PLD + lane] = pr[r][0];
      Ps[i * WA_PLD + lane + 32] = pr[r][1];
      Ss[i * WA_PLD + lane] = s0;
      Ss[i * WA_PLD + lane + 32] = s1;
      Acc[i * WA_N + lane] += s0;
      Acc[i * WA_N + lane + 32] += s1;
    }
    __syncthreads();
    if (lane < g.D) {
      for (int r = 0; r < 16; ++r) {
        const int n = wp * 16 + r;  // row index for dQ, column index for dK/dV
        if (n >= g.N) break;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int m = 0; m < g.N; ++m) {
          dq = fmaf(Ss[n * WA_PLD + m], Ks[m * WA_LD + lane], dq);
          dk = fmaf(Ss[m * WA_PLD + n], Qs[m * WA_LD + lane], dk);
          dv = fmaf(Ps[m * WA_PLD + n], Os[m * WA_LD + lane], dv);
        }
        float* o = dqkv + (size_t)tok[n] * 3 * g.C + head * g.D + lane;
        o[0] = dq * g.scale;
        o[g.C] = dk;
        o[2 * g.C] = dv;
      }
    }
  }
  __syncthreads();
  float* outp = partial + ((size_t)blockIdx.x * g.heads + head) * WA_N * WA_N;
  for (int i = t; i < WA_N * WA_N; i += WA_THREADS) outp[i] = Acc[i];
}

// Bias-table gradient, two deterministic passes:
//  (1) dS_sum[head][i][j] = sum over CTA partials (one thread per element, coalesced over j)
//  (2) dtable[tidx, head] = sum of dS_sum over the (i, j) pairs with rel_index(i, j) == tidx
__global__ void window_attn_dbias_sum(const float* __restrict__ partial, float* __restrict__ dssum, int gx, int heads) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // head * N*N + i*N + j
  if (e >= heads * WA_N * WA_N) return;
  float s = 0.f;
  for (int b = 0; b < gx; ++b) s += partial[(size_t)b * heads * WA_N * WA_N + e];
  dssum[e] = s;
}
__global__ void window_attn_dbias_kernel(const float* __restrict__ dssum, float* __restrict__ dtable, int heads, int ws) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int span = 2 * ws - 1;
  if (id >= span * span * heads) return;
  const int head = id % heads, tidx = id / heads;
  const int dy = tidx / span - (ws - 1), dx = tidx % span - (ws - 1);
  float s = 0.f;
  for (int jy = 0; jy < ws; ++jy) {
    const int iy = jy + dy;
    if (iy < 0 || iy >= ws) continue;
    for (int jx = 0; jx < ws; ++jx) {
      const int ix = jx + dx;
      if (ix < 0 || ix >= ws) continue;
      const int i = iy * ws + ix, j = jy * ws + jx;
      s += dssum[(size_t)head * WA_N * WA_N + i * WA_N + j];
    }
  }
  dtable[tidx * heads + head] = s;
}

// The same two steps for MANY layers in one launch each (blockIdx.y = layer): nsr_window_attn_dbias_multi
__global__ void window_attn_dbias_sum_multi(const NsrAttnBiasEntry* __restrict__ tab) {
  const NsrAttnBiasEntry e = tab[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // head * N*N + i*N + j
  const int n = e.heads * WA_N * WA_N;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < e.gx; ++b) s += e.partial[(size_t)b * n + i];
  e.partial[(size_t)e.gx * n + i] = s;  // the dssum slot behind the partials
}
__global__ void window_attn_dbias_multi(const NsrAttnBiasEntry* __restrict__ tab) {
  const NsrAttnBiasEntry e = tab[blockIdx.y];
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int span = 2 * e.ws - 1;
  if (id >= span * span * e.heads) return;
  const float* dssum = e.partial + (size_t)e.gx * e.heads * WA_N * WA_N;
  const int head = id % e.heads, tidx = id / e.heads;
  const int dy = tidx / span - (e.ws - 1), dx = tidx % span - (e.ws - 1);
  float s = 0.f;
  for (int jy = 0; jy < e.ws; ++jy) {
    const int iy = jy + dy;
    if (iy < 0 || iy >= e.ws) continue;
    for (int jx = 0; jx < e.ws; ++jx) {
      const int ix = jx + dx;
      if (ix < 0 || ix >= e.ws) continue;
      s += dssum[(size_t)head * WA_N * WA_N + (iy * e.ws + ix) * WA_N + jy * e.ws + jx];
    }
  }
  e.dbias_table[tidx * e.heads + head] = s;
}

bool window_attn_mma_supported(int c, int heads, int ws);
int window_attn_fwd_mma_launch(const float* qkv, const float* table, float* out, void* out_sti, int batch, int h, int w,
                               int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st);
int window_attn_bwd_mma_ctas();
int window_attn_bwd_mma_launch(const float* qkv, const float* table, const float* dout, float* dqkv, void* dqkv_sti,
                               float* partial, int gx, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask,
                               float scale, cudaStream_t st);
int window_attn_wsti_fwd_launch(const void* qkv, const float* table, float* out, void* out_sti, int batch, int h, int w,
                                int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st);
int window_attn_wsti_bwd_launch(const void* qkv, const float* table, const void* dout, float* dqkv, void* dqkv_sti,
                                float* partial, int gx, int batch, int h, int w, int c, int heads, int ws, int shift,
                                int use_mask, float scale, cudaStream_t st);
bool window_attn_tc_supported(int c, int heads, int ws);
int window_attn_tc_fwd_launch(const void* qkv, const float* table, float* out, void* out_sti, int out_padded, int batch, int h,
                              int w, int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st);
int window_attn_tc_bwd_gx(int heads);
int window_attn_tc_bwd_launch(const void* qkv, const float* table, const void* dout, float* dqkv, void* dqkv_sti, float* partial,
                              int gx, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask, float scale,
                              cudaStream_t st);
static bool use_mma(int c, int heads, int ws) {
  static int simt_forced = -1;
  if (simt_forced < 0) {
    const char* e = getenv("NSR_ATTN");
    simt_forced = (e && !strcmp(e, "simt")) ? 1 : 0;
  }
  return !simt_forced && window_attn_mma_supported(c, heads, ws);
}

static int bwd_gx(int nwin, int heads, bool mma) {
  int gx = ((mma ? window_attn_bwd_mma_ctas() : 2) * kNumSMs) / heads;
  if (gx < 1) gx = 1;
  return gx > nwin ? nwin : gx;
}
static int make_geom(WinGeom& g, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask,
                     float scale, const char* who) {
  NSR_CHECK_ARG(batch > 0 && h > 0 && w > 0 && c > 0 && heads > 0 && ws > 0, "%s: bad geometry", who);
  NSR_CHECK_ARG(c % heads == 0 && c / heads <= WA_D, "%s: head_dim must be <= %d", who, WA_D);
  NSR_CHECK_ARG(ws * ws <= WA_N && ws <= 8, "%s: window %d not supported (ws*ws <= %d)", who, ws, WA_N);
  NSR_CHECK_ARG(h % ws == 0 && w % ws == 0, "%s: h, w must be multiples of the window size", who);
  NSR_CHECK_ARG(shift >= 0 && shift < ws, "%s: shift must be in [0, ws)", who);
  g = WinGeom{batch, h, w, c, heads, ws, shift, use_mask, ws * ws, c / heads, h / ws, w / ws, scale};
  return NSR_OK;
}
}  // namespace nsr
using namespace nsr;

extern "C" int nsr_window_attn_fwd(const float* qkv, const float* bias_table, float* out, int batch, int h, int w,
                                   int c, int heads, int ws, int shift, int use_mask, float scale, void* out_sti,
                                   void* stream) {
  NSR_CHECK_ARG(qkv && bias_table && (out || out_sti), "nsr_window_attn_fwd: null pointer");
  WinGeom g;
  int rc = make_geom(g, batch, h, w, c, heads, ws, shift, use_mask, scale, "nsr_window_attn_fwd");
  if (rc) return rc;
  if (use_mma(c, heads, ws))
    return window_attn_fwd_mma_launch(qkv, bias_table, out, out_sti, batch, h, w, c, heads, ws, shift, use_mask, scale,
                                      reinterpret_cast<cudaStream_t>(stream));
  NSR_CHECK_ARG(out, "nsr_window_attn_fwd: this shape runs on the fp32 kernel, which needs the fp32 output buffer");
  dim3 grid(batch * g.nwh * g.nww, heads);
  window_attn_fwd_kernel<<<grid, WA_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, bias_table, out, g);
  NSR_CHECK_LAUNCH("window_attn_fwd");
  if (out_sti) return nsr_sti_from_f32(out, c, (long long)batch * h * w, c, out_sti, stream);
  return NSR_OK;
}

extern "C" size_t nsr_window_attn_bwd_workspace(int heads, int ws) {
  (void)ws;
  return (size_t)(window_attn_bwd_mma_ctas() * kNumSMs + 2 * (heads > 0 ? heads : 1)) * WA_N * WA_N * sizeof(float);
}

extern "C" int nsr_window_attn_bwd(const float* qkv, const float* bias_table, const float* dout, float* dqkv,
                                   float* dbias_table, int batch, int h, int w, int c, int heads, int ws, int shift,
                                   int use_mask, float scale, void* workspace, size_t workspace_bytes, void* dqkv_sti,
                                   void* stream) {
  NSR_CHECK_ARG(qkv && bias_table && dout && (dqkv || dqkv_sti) && dbias_table, "nsr_window_attn_bwd: null pointer");
  WinGeom g;
  int rc = make_geom(g, batch, h, w, c, heads, ws, shift, use_mask, scale, "nsr_window_attn_bwd");
  if (rc) return rc;
  const int nwin = batch * g.nwh * g.nww;
  const bool mma = use_mma(c, heads, ws);
  const int gx = bwd_gx(nwin, heads, mma);
  const size_t need = (size_t)(gx + 1) * heads * WA_N * WA_N * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_error("nsr_window_attn_bwd: workspace %zu < %zu", workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)WA_BWD_SMEM);
    if (e != cudaSuccess) {
      set_error("nsr_window_attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr_set = true;
  }
  float* partial = reinterpret_cast<float*>(workspace);
  if (mma) {
    rc = window_attn_bwd_mma_launch(qkv, bias_table, dout, dqkv, dqkv_sti, partial, gx, batch, h, w, c, heads, ws, shift,
                                    use_mask, scale, st);
    if (rc) return rc;
  } else {
    NSR_CHECK_ARG(dqkv, "nsr_window_attn_bwd: this shape runs on the fp32 kernel, which needs the fp32 dqkv buffer");
    dim3 grid(gx, heads);
    window_attn_bwd_kernel<<<grid, WA_THREADS, WA_BWD_SMEM, st>>>(qkv, bias_table, dout, dqkv, partial, g, nwin);
    NSR_CHECK_LAUNCH("window_attn_bwd");
    if (dqkv_sti) {
      rc = nsr_sti_from_f32(dqkv, 3 * c, (long long)batch * h * w, 3 * c, dqkv_sti, stream);
      if (rc) return rc;
    }
  }
  float* dssum = partial + (size_t)gx * heads * WA_N * WA_N;  // the workspace has `heads` spare tiles after the partials
  window_attn_dbias_sum<<<ceil_div(heads * WA_N * WA_N, 256), 256, 0, st>>>(partial, dssum, gx, heads);
  NSR_CHECK_LAUNCH("window_attn_dbias_sum");
  const int n = (2 * ws - 1) * (2 * ws - 1) * heads;
  window_attn_dbias_kernel<<<ceil_div(n, 128), 128, 0, st>>>(dssum, dbias_table, heads, ws);
  NSR_CHECK_LAUNCH("window_attn_dbias");
  return NSR_OK;
}

// ---- window-ordered operands (see window_attn_mma.cu) -------------------------------------------------------------
extern "C" int nsr_window_attn_wsti_channels(int heads) { return (heads * 32 + 63) / 64 * 64; }

static int wsti_check(WinGeom& g, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask, float scale,
                      const char* who) {
  int rc = make_geom(g, batch, h, w, c, heads, ws, shift, use_mask, scale, who);
  if (rc) return rc;
  NSR_CHECK_ARG(window_attn_mma_supported(c, heads, ws) && nsr_device_supports_tcgen05(),
                "%s: needs window 8, an even head dim <= 32 and an sm_100 device (bulk-copied operands)", who);
  return NSR_OK;
}

extern "C" int nsr_window_attn_wsti_fwd(const void* qkv_wsti, const float* bias_table, float* out, void* out_sti,
                                        int out_padded, int batch, int h, int w, int c, int heads, int ws, int shift,
                                        int use_mask, float scale, int engine, void* stream) {
  NSR_CHECK_ARG(qkv_wsti && bias_table && (out || out_sti), "nsr_window_attn_wsti_fwd: null pointer");
  NSR_CHECK_ARG(engine == NSR_ENGINE_AUTO || engine == NSR_ENGINE_TCGEN05 || engine == NSR_ENGINE_MMA_SYNC,
                "nsr_window_attn_wsti_fwd: engine must be NSR_ENGINE_AUTO, _TCGEN05 or _MMA_SYNC");
  NSR_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv_wsti) & 15) == 0, "nsr_window_attn_wsti_fwd: image must be 16-byte aligned");
  WinGeom g;
  int rc = wsti_check(g, batch, h, w, c, heads, ws, shift, use_mask, scale, "nsr_window_attn_wsti_fwd");
  if (rc) return rc;
  if (engine != NSR_ENGINE_MMA_SYNC && window_attn_tc_supported(c, heads, ws))
    return window_attn_tc_fwd_launch(qkv_wsti, bias_table, out, out_sti, out_padded, batch, h, w, c, heads, ws, shift, use_mask,
                                     scale, reinterpret_cast<cudaStream_t>(stream));
  NSR_CHECK_ARG(engine != NSR_ENGINE_TCGEN05, "nsr_window_attn_wsti_fwd: shape not supported by the tcgen05 kernel");
  NSR_CHECK_ARG(!out_padded, "nsr_window_attn_wsti_fwd: the head-padded output image is written by the tcgen05 kernel only");
  return window_attn_wsti_fwd_launch(qkv_wsti, bias_table, out, out_sti, batch, h, w, c, heads, ws, shift, use_mask, scale,
                                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int nsr_window_attn_wsti_bwd(const void* qkv_wsti, const float* bias_table, const void* dout_wsti, float* dqkv,
                                        void* dqkv_sti, int dqkv_padded, float* dbias_table, int batch, int h, int w, int c,
                                        int heads, int ws, int shift, int use_mask, float scale, int engine, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  NSR_CHECK_ARG(qkv_wsti && bias_table && dout_wsti && (dqkv || dqkv_sti), "nsr_window_attn_wsti_bwd: null pointer");
  NSR_CHECK_ARG(engine == NSR_ENGINE_AUTO || engine == NSR_ENGINE_TCGEN05 || engine == NSR_ENGINE_MMA_SYNC,
                "nsr_window_attn_wsti_bwd: engine must be NSR_ENGINE_AUTO, _TCGEN05 or _MMA_SYNC");
  NSR_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv_wsti) | reinterpret_cast<uintptr_t>(dout_wsti)) & 15) == 0,
                "nsr_window_attn_wsti_bwd: images must be 16-byte aligned");
  WinGeom g;
  int rc = wsti_check(g, batch, h, w, c, heads, ws, shift, use_mask, scale, "nsr_window_attn_wsti_bwd");
  if (rc) return rc;
  const int nwin = batch * g.nwh * g.nww;
  // tcgen05 kernel: needs the head-padded dqkv image (16-byte stores); the mma.sync kernel writes the compact one
  const bool tc_ok = window_attn_tc_supported(c, heads, ws) && (dqkv_sti == nullptr || dqkv_padded);
  const bool use_tc = engine != NSR_ENGINE_MMA_SYNC && tc_ok;
  NSR_CHECK_ARG(use_tc || engine != NSR_ENGINE_TCGEN05,
                "nsr_window_attn_wsti_bwd: the tcgen05 kernel needs a supported shape and dqkv_padded = 1");
  NSR_CHECK_ARG(use_tc || !dqkv_padded, "nsr_window_attn_wsti_bwd: the head-padded dqkv image is written by the tcgen05 kernel only");
  const int gx = use_tc ? window_attn_tc_bwd_gx(heads) : bwd_gx(nwin, heads, true);
  const size_t need = (size_t)(gx + 1) * heads * WA_N * WA_N * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_error("nsr_window_attn_wsti_bwd: workspace %zu < %zu", workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  if (use_tc)
    rc = window_attn_tc_bwd_launch(qkv_wsti, bias_table, dout_wsti, dqkv, dqkv_sti, partial, gx, batch, h, w, c, heads, ws, shift,
                                   use_mask, scale, st);
  else
    rc = window_attn_wsti_bwd_launch(qkv_wsti, bias_table, dout_wsti, dqkv, dqkv_sti, partial, gx, batch, h, w, c, heads, ws,
                                     shift, use_mask, scale, st);
  if (rc) return rc;
  // dbias_table == NULL: the caller keeps `workspace` ([gx + 1][heads][64][64]: per-CTA partials of dS, one scratch slot)
  // and reduces it later together with the other layers' (nsr_window_attn_dbias_multi, gx = nsr_window_attn_wsti_bwd_gx)
  if (dbias_table == nullptr) return NSR_OK;
  float* dssum = partial + (size_t)gx * heads * WA_N * WA_N;
  window_attn_dbias_sum<<<ceil_div(heads * WA_N * WA_N, 256), 256, 0, st>>>(partial, dssum, gx, heads);
  NSR_CHECK_LAUNCH("window_attn_dbias_sum");
  const int n = (2 * ws - 1) * (2 * ws - 1) * heads;
  window_attn_dbias_kernel<<<ceil_div(n, 128), 128, 0, st>>>(dssum, dbias_table, heads, ws);
  NSR_CHECK_LAUNCH("window_attn_dbias");
  return NSR_OK;
}

extern "C" int nsr_window_attn_wsti_bwd_gx(int batch, int h, int w, int c, int heads, int ws, int dqkv_padded, int engine) {
  if (ws <= 0 || h % ws || w % ws) return 0;
  const int nwin = batch * (h / ws) * (w / ws);
  const bool use_tc = engine != NSR_ENGINE_MMA_SYNC && window_attn_tc_supported(c, heads, ws) && dqkv_padded;
  return use_tc ? window_attn_tc_bwd_gx(heads) : bwd_gx(nwin, heads, true);
}
extern "C" int nsr_window_attn_dbias_multi(const NsrAttnBiasEntry* table_dev, int n_entries, int max_heads, int max_ws, void* stream) {
  NSR_CHECK_ARG(table_dev && n_entries > 0 && n_entries <= 65535 && max_heads > 0 && max_ws > 0 && max_ws * max_ws <= WA_N,
                "nsr_window_attn_dbias_multi: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  window_attn_dbias_sum_multi<<<dim3(ceil_div(max_heads * WA_N * WA_N, 256), n_entries), 256, 0, st>>>(table_dev);
  NSR_CHECK_LAUNCH("window_attn_dbias_sum_multi");
  const int n = (2 * max_ws - 1) * (2 * max_ws - 1) * max_heads;
  window_attn_dbias_multi<<<dim3(ceil_div(n, 128), n_entries), 128, 0, st>>>(table_dev);
  NSR_CHECK_LAUNCH("window_attn_dbias_multi");
  return NSR_OK;
}

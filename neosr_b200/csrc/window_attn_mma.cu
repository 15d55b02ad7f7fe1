// Fused (shifted-)window attention core on the tensor cores (warp-level mma.sync m16n8k16,
// bf16 operands split hi/lo in three passes, fp32 accumulate) for 8x8 windows, head_dim <= 32.
//
// One CTA iteration = one (window, head): 64 tokens, 4 warps x 16 query rows.  The 64x64x32
// micro-GEMMs are far too small for a tcgen05/TMEM pipeline to pay off (a 128-row UMMA would
// need two windows with different K/V); the register-resident mma.sync form keeps S, P, dP, dS in
// registers between the GEMMs (the accumulator layout of one MMA is the A-operand layout of the
// next), so nothing but Q/K/V/dO tiles ever touches shared memory in the forward pass.
// Shift / window partition / mask / relative-position bias are index math, as in window_attn.cu.
#include "attn_mma.cuh"
#include "tc_common.cuh"

namespace nsr {

constexpr int AM_N = 64, AM_THREADS = 128;

struct AttnGeom {
  int B, H, W, C, heads, ws, shift, use_mask, D, nwh, nww;
  float scale;
};

__device__ __forceinline__ void attn_token_map(const AttnGeom& g, int wi, int n, int& tok, int& rid) {
  const int per = g.nwh * g.nww;
  const int b = wi / per, rem = wi - b * per;
  const int wy = rem / g.nww, wx = rem - wy * g.nww;
  const int iy = n / g.ws, ix = n - iy * g.ws;
  const int hs = wy * g.ws + iy, wsx = wx * g.ws + ix;
  int ho = hs + g.shift, wo = wsx + g.shift;
  if (ho >= g.H) ho -= g.H;
  if (wo >= g.W) wo -= g.W;
  tok = (b * g.H + ho) * g.W + wo;
  const int rh = hs < g.H - g.ws ? 0 : (hs < g.H - g.shift ? 1 : 2);
  const int rw = wsx < g.W - g.ws ? 0 : (wsx < g.W - g.shift ? 1 : 2);
  rid = rh * 3 + rw;
}

// store the pair (v0, v1) for channels (cidx, cidx+1) of token row p into a split tile image
__device__ __forceinline__ void sti_store_pair(uint8_t* sti, int kbs, long long p, int cidx, float v0, float v1) {
  uint32_t hi, lo;
  split_pair(v0, v1, hi, lo);
  const int r = (int)(p & 127), kb = cidx >> 6, cc = cidx & 63;
  uint8_t* dst = sti + ((size_t)((p >> 7) * kbs + kb) << 15) + r * 128 + (((cc >> 3) ^ (r & 7)) << 4) + (cc & 7) * 2;
  *reinterpret_cast<uint32_t*>(dst) = hi;
  *reinterpret_cast<uint32_t*>(dst + 16384) = lo;
}
// same, with the row part of the address (block row base, r & 7) hoisted by the caller
__device__ __forceinline__ void sti_store_pair_row(uint8_t* rowbase, int r7, int cidx, float v0, float v1) {
  uint32_t hi, lo;
  split_pair(v0, v1, hi, lo);
  const int cc = cidx & 63;
  uint8_t* dst = rowbase + ((size_t)(cidx >> 6) << 15) + ((((cc >> 3) ^ r7) << 4) + (cc & 7) * 2);
  *reinterpret_cast<uint32_t*>(dst) = hi;
  *reinterpret_cast<uint32_t*>(dst + 16384) = lo;
}
__device__ __forceinline__ uint8_t* sti_row_base(uint8_t* sti, int kbs, long long p) {
  return sti + ((size_t)((p >> 7) * kbs) << 15) + (size_t)(p & 127) * 128;
}
// zero the channel padding [cfirst, kbs*64) of this window's 64 tokens (done by the last head's CTA)
// ones: channel cfirst carries 1.0 (the bias-gradient column of the consumer's wgrad, see layernorm.cu)
__device__ __forceinline__ void sti_zero_padding(uint8_t* sti, int kbs, const int* tok, int cfirst, int t, bool ones = false) {
  const int pairs = (kbs * 64 - cfirst) / 2;
  for (int i = t; i < AM_N * pairs; i += AM_THREADS) {
    const int n = i / pairs, pr = i - n * pairs;
    sti_store_pair(sti, kbs, tok[n], cfirst + 2 * pr, (ones && pr == 0) ? 1.f : 0.f, 0.f);
  }
}

// bias + mask + softmax on the accumulator layout (rows row0+g and row0+g+8), in place -> P.
// ws == 8: token j = 8*jy + jx, so column j = 8*nt + 2*tid + e has jy = nt, jx = 2*tid + e and the
// relative-position index (iy-jy+7)*15 + (ix-jx+7) = (15*iy + ix + 112 - 2*tid) - 15*nt - e: one
// integer add per element.  `masked` is false for windows whose 64 tokens share one region id.
template <bool STATS = false>
__device__ __forceinline__ void bias_mask_softmax(const float* bias_s, const int* rid, bool masked, int row0, int g,
                                                  int tid, float (&acc)[8][4], float* lse = nullptr) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = row0 + g + 8 * h;
    const int base = 15 * (i >> 3) + (i & 7) + 112 - 2 * tid;
    float m = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float v = acc[nt][2 * h + e] + bias_s[base - 15 * nt - e];
        acc[nt][2 * h + e] = v;
      }
    }
    if (masked) {
      const int ri = rid[i];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int2 rj = *reinterpret_cast<const int2*>(&rid[nt * 8 + tid * 2]);
        if (ri != rj.x) acc[nt][2 * h] += -100.0f;
        if (ri != rj.y) acc[nt][2 * h + 1] += -100.0f;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) m = fmaxf(m, fmaxf(acc[nt][2 * h], acc[nt][2 * h + 1]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float s = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ev = __expf(acc[nt][2 * h + e] - m);
        acc[nt][2 * h + e] = ev;
        s += ev;
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float inv = 1.f / s;
    if (STATS) lse[h] = m + logf(s);  // P = exp(score - lse): what the backward's transposed pass recomputes
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc[nt][2 * h] *= inv;
      acc[nt][2 * h + 1] *= inv;
    }
  }
}


// ------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(AM_THREADS) window_attn_fwd_mma(const float* __restrict__ qkv,
                                                                 const float* __restrict__ table,
                                                                 float* __restrict__ out, uint8_t* __restrict__ out_sti,
                                                                 AttnGeom gm) {
  __shared__ __align__(16) __nv_bfloat16 tiles[6 * AM_N * AM_LD];  // Q, K, V hi/lo, all [token][d]
  __nv_bfloat16* Qh = tiles;                  __nv_bfloat16* Ql = Qh + AM_N * AM_LD;
  __nv_bfloat16* Kh = Ql + AM_N * AM_LD;      __nv_bfloat16* Kl = Kh + AM_N * AM_LD;
  __nv_bfloat16* Vh = Kl + AM_N * AM_LD;      __nv_bfloat16* Vl = Vh + AM_N * AM_LD;
  __shared__ float bias_s[225];
  __shared__ __align__(8) int tok[AM_N], rid[AM_N];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  // heads vary fastest over the grid: the six CTAs that read the same 64 qkv rows (120 of each row's 2160 bytes per
  // head) run together, so the rows come from DRAM once and from L2 five times (was: 2.2x the algorithmic DRAM reads)
  const int wi = blockIdx.x / gm.heads, head = blockIdx.x - wi * gm.heads;
  int differs = 0;
  if (t < AM_N) {
    attn_token_map(gm, wi, t, tok[t], rid[t]);
    int t0, r0;
    attn_token_map(gm, wi, 0, t0, r0);
    differs = rid[t] != r0;
  }
  for (int i = t; i < (2 * gm.ws - 1) * (2 * gm.ws - 1); i += AM_THREADS) bias_s[i] = table[i * gm.heads + head];
  // zero the d-padding (columns D..31) of the six tiles
  {
    const int per_row = (32 - gm.D) / 2;
    for (int i = t; i < 6 * AM_N * per_row; i += AM_THREADS) {
      const int cpair = i % per_row, rowt = i / per_row;  // rowt = tile * 64 + row
      *reinterpret_cast<uint32_t*>(&tiles[rowt * AM_LD + gm.D + 2 * cpair]) = 0;
    }
  }
  const bool masked = __syncthreads_or(differs) && gm.use_mask && gm.shift > 0;
  // 16 lanes per token row (D / 2 <= 16 float2 pairs, coalesced 4 D-byte rows), 8 rows per pass, 8 passes;
  // LB passes are issued together so 3 * LB requests are in flight per thread
  const int pr = t & 15, rbase = t >> 4;
  const bool act = pr < gm.D / 2;
  constexpr int LB = 4;
#pragma unroll
  for (int u0 = 0; u0 < 8; u0 += LB) {
    float2 q[LB], k[LB], v[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int n = rbase + 8 * (u0 + u);
      if (act) {
        const float* p = qkv + (size_t)tok[n] * 3 * gm.C + head * gm.D + 2 * pr;
        q[u] = *reinterpret_cast<const float2*>(p);
        k[u] = *reinterpret_cast<const float2*>(p + gm.C);
        v[u] = *reinterpret_cast<const float2*>(p + 2 * gm.C);
      }
    }
    if (act) {
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        const int o1 = (rbase + 8 * (u0 + u)) * AM_LD + 2 * pr;
        uint32_t hi, lo;
        split_pair(q[u].x * gm.scale, q[u].y * gm.scale, hi, lo);
        *reinterpret_cast<uint32_t*>(&Qh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Ql[o1]) = lo;
        split_pair(k[u].x, k[u].y, hi, lo);
        *reinterpret_cast<uint32_t*>(&Kh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Kl[o1]) = lo;
        split_pair(v[u].x, v[u].y, hi, lo);
        *reinterpret_cast<uint32_t*>(&Vh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Vl[o1]) = lo;
      }
    }
  }
  __syncthreads();
  const int row0 = warp * 16;
  float acc[8][4];
  qk_scores_ldm(Qh, Ql, Kh, Kl, row0, lane, acc);
  bias_mask_softmax(bias_s, rid, masked, row0, g, tid, acc);
  float o[4][4];
  acc_times_ldm(acc, Vh, Vl, lane, o);  // V read [token][d] through ldmatrix.trans
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long tk = tok[row0 + g + 8 * h];
    uint8_t* rb = out_sti ? sti_row_base(out_sti, (gm.C + 63) / 64, tk) : nullptr;
    const int r7 = (int)(tk & 7);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + tid * 2;
      if (c < gm.D) {
        if (out) *reinterpret_cast<float2*>(out + (size_t)tk * gm.C + head * gm.D + c) = make_float2(o[nt][2 * h], o[nt][2 * h + 1]);
        if (out_sti) sti_store_pair_row(rb, r7, head * gm.D + c, o[nt][2 * h], o[nt][2 * h + 1]);
      }
    }
  }
  if (out_sti && head == gm.heads - 1) sti_zero_padding(out_sti, (gm.C + 63) / 64, tok, gm.C, t, true);
}

// ------------------------------------------------------------------------------------ backward
constexpr int BW_TILE = AM_N * AM_LD;      // 2560 bf16 per [token][d] tile
constexpr int BW_PHASE1 = 8 * BW_TILE;     // Q, K, V, dO hi + lo
// One persistent CTA iterates over windows of one head; no transposed operand copies and no P / dS exchange
// through shared memory.  Row pass (warp = 16 queries i): P, dP, delta_i, dS -> dQ and the bias-gradient accumulators;
// it leaves lse_i and delta_i in shared memory.  Column pass (warp = 16 keys j): recomputes the TRANSPOSED tiles
// S^T = K Qs^T and dP^T = V dO^T on the tensor cores (2 x 48 extra MMAs, cheaper than 256 two-byte scatter stores),
// so P^T and dS^T appear in the accumulator layout = the A operand of dV = P^T dO and dK = dS^T Qs.  The B operands
// of dQ / dK / dV are K / Qs / dO read [token][d] through ldmatrix.trans.  Shared memory: 8 tiles (40 KB).
constexpr size_t BW2_SMEM = (size_t)BW_PHASE1 * sizeof(__nv_bfloat16);

// 3 CTAs / SM (168 registers): 4 (128 registers, 330 B of spills) and 5 (96) measured the same and 19% slower.
__global__ void __launch_bounds__(AM_THREADS, 3) window_attn_bwd_mma(const float* __restrict__ qkv,
                                                                  const float* __restrict__ table,
                                                                  const float* __restrict__ dout,
                                                                  float* __restrict__ dqkv,
                                                                  uint8_t* __restrict__ dqkv_sti,
                                                                  float* __restrict__ partial, AttnGeom gm, int nwin) {
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16* Qh = sm;                 __nv_bfloat16* Ql = Qh + BW_TILE;
  __nv_bfloat16* Kh = Ql + BW_TILE;       __nv_bfloat16* Kl = Kh + BW_TILE;
  __nv_bfloat16* Vh = Kl + BW_TILE;       __nv_bfloat16* Vl = Vh + BW_TILE;
  __nv_bfloat16* Oh = Vl + BW_TILE;       __nv_bfloat16* Ol = Oh + BW_TILE;    // dO
  __shared__ float bias_s[225];
  __shared__ __align__(8) int tok[AM_N], rid[AM_N];
  __shared__ __align__(8) float lse_s[AM_N], delta_s[AM_N];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int head = blockIdx.y;
  const int row0 = warp * 16;
  for (int i = t; i < (2 * gm.ws - 1) * (2 * gm.ws - 1); i += AM_THREADS) bias_s[i] = table[i * gm.heads + head];
  for (int i = t; i < BW_PHASE1 / 2; i += AM_THREADS) reinterpret_cast<uint32_t*>(sm)[i] = 0;  // d padding columns stay 0
  float dacc[8][4];  // sum over this CTA's windows of dS, accumulator layout (fixed (i,j) per thread)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dacc[nt][0] = dacc[nt][1] = dacc[nt][2] = dacc[nt][3] = 0.f;
  const int pr = t & 15, rbase = t >> 4;  // loads: 16 lanes per token row, 8 rows per pass (see the forward kernel)
  const bool act = pr < gm.D / 2;
  const int kbs3 = (3 * gm.C + 63) / 64;

  for (int wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    __syncthreads();  // previous window's column pass is done with the tiles, tok, rid and the row statistics
    int differs = 0;
    if (t < AM_N) {
      attn_token_map(gm, wi, t, tok[t], rid[t]);
      int t0, r0;
      attn_token_map(gm, wi, 0, t0, r0);
      differs = rid[t] != r0;
    }
    const bool masked = __syncthreads_or(differs) && gm.use_mask && gm.shift > 0;
    constexpr int LB = 4;  // 16 independent 8-byte loads in flight per thread
#pragma unroll
    for (int u0 = 0; u0 < 8; u0 += LB) {
      float2 qv[LB], kv[LB], vv[LB], dv[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        if (act) {
          const long long tk = tok[rbase + 8 * (u0 + u)];
          const float* p = qkv + (size_t)tk * 3 * gm.C + head * gm.D + 2 * pr;
          qv[u] = *reinterpret_cast<const float2*>(p);
          kv[u] = *reinterpret_cast<const float2*>(p + gm.C);
          vv[u] = *reinterpret_cast<const float2*>(p + 2 * gm.C);
          dv[u] = *reinterpret_cast<const float2*>(dout + (size_t)tk * gm.C + head * gm.D + 2 * pr);
        }
      }
      if (act) {
#pragma unroll
        for (int u = 0; u < LB; ++u) {
          uint32_t hi, lo;
          const int o1 = (rbase + 8 * (u0 + u)) * AM_LD + 2 * pr;
          split_pair(qv[u].x * gm.scale, qv[u].y * gm.scale, hi, lo);
          *reinterpret_cast<uint32_t*>(&Qh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Ql[o1]) = lo;
          split_pair(kv[u].x, kv[u].y, hi, lo);
          *reinterpret_cast<uint32_t*>(&Kh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Kl[o1]) = lo;
          split_pair(vv[u].x, vv[u].y, hi, lo);
          *reinterpret_cast<uint32_t*>(&Vh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Vl[o1]) = lo;
          split_pair(dv[u].x, dv[u].y, hi, lo);
          *reinterpret_cast<uint32_t*>(&Oh[o1]) = hi; *reinterpret_cast<uint32_t*>(&Ol[o1]) = lo;
        }
      }
    }
    __syncthreads();
    uint8_t* rb[2];
    int r7[2];
    long long tkk[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // rows row0+g, row0+g+8: the queries of the row pass and the keys of the column pass
      tkk[h] = tok[row0 + g + 8 * h];
      rb[h] = dqkv_sti ? sti_row_base(dqkv_sti, kbs3, tkk[h]) : nullptr;
      r7[h] = (int)(tkk[h] & 7);
    }
    float pr_[8][4], ds[8][4], o[4][4];
    // ---- row pass: P = softmax(Qs K^T + bias + mask); dP = dO V^T; dS = P o (dP - rowsum(P o dP)); dQ = scale dS K
    {
      float lse[2];
      qk_scores_ldm(Qh, Ql, Kh, Kl, row0, lane, pr_);
      bias_mask_softmax<true>(bias_s, rid, masked, row0, g, tid, pr_, lse);
      qk_scores_ldm(Oh, Ol, Vh, Vl, row0, lane, ds);  // dP[i][j] = sum_d dO[i][d] V[j][d]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float delta = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) delta += pr_[nt][2 * h] * ds[nt][2 * h] + pr_[nt][2 * h + 1] * ds[nt][2 * h + 1];
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
        if (tid == 0) { lse_s[row0 + g + 8 * h] = lse[h]; delta_s[row0 + g + 8 * h] = delta; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float v = pr_[nt][2 * h + e] * (ds[nt][2 * h + e] - delta);
            ds[nt][2 * h + e] = v;
            dacc[nt][2 * h + e] += v;
          }
        }
      }
      acc_times_ldm(ds, Kh, Kl, lane, o);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = nt * 8 + tid * 2;
          if (c < gm.D) {
            const float v0 = o[nt][2 * h] * gm.scale, v1 = o[nt][2 * h + 1] * gm.scale;
            if (dqkv) *reinterpret_cast<float2*>(dqkv + tkk[h] * 3 * gm.C + head * gm.D + c) = make_float2(v0, v1);
            if (dqkv_sti) sti_store_pair_row(rb[h], r7[h], head * gm.D + c, v0, v1);
          }
        }
      }
    }
    __syncthreads();  // lse / delta of all 64 queries are visible
    // ---- column pass (rows = keys j, columns = queries i = 8 nt + 2 tid + e): P^T, dP^T, dS^T -> dK, dV
    qk_scores_ldm(Kh, Kl, Qh, Ql, row0, lane, pr_);   // S^T[j][i] = sum_d K[j][d] Qs[i][d]
    qk_scores_ldm(Vh, Vl, Oh, Ol, row0, lane, ds);    // dP^T[j][i] = sum_d V[j][d] dO[i][d]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = row0 + g + 8 * h;
      // relative-position index of (i, j): (iy - jy + 7) * 15 + (ix - jx + 7) with iy = nt, ix = 2 tid + e
      const int base = 15 * (7 - (j >> 3)) + 7 - (j & 7) + 2 * tid;
      const int rj = rid[j];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 ls = *reinterpret_cast<const float2*>(&lse_s[nt * 8 + tid * 2]);
        const float2 dl = *reinterpret_cast<const float2*>(&delta_s[nt * 8 + tid * 2]);
        float s0 = pr_[nt][2 * h] + bias_s[base + 15 * nt], s1 = pr_[nt][2 * h + 1] + bias_s[base + 15 * nt + 1];
        if (masked) {
          const int2 ri = *reinterpret_cast<const int2*>(&rid[nt * 8 + tid * 2]);
          if (ri.x != rj) s0 += -100.0f;
          if (ri.y != rj) s1 += -100.0f;
        }
        const float p0 = __expf(s0 - ls.x), p1 = __expf(s1 - ls.y);
        pr_[nt][2 * h] = p0;
        pr_[nt][2 * h + 1] = p1;
        ds[nt][2 * h] = p0 * (ds[nt][2 * h] - dl.x);
        ds[nt][2 * h + 1] = p1 * (ds[nt][2 * h + 1] - dl.y);
      }
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      if (which == 0) acc_times_ldm(ds, Qh, Ql, lane, o);   // dK[j][d] = sum_i dS[i][j] Qs[i][d]
      else acc_times_ldm(pr_, Oh, Ol, lane, o);             // dV[j][d] = sum_i P[i][j] dO[i][d]
      const int cbase = (which == 0 ? gm.C : 2 * gm.C) + head * gm.D;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = nt * 8 + tid * 2;
          if (c < gm.D) {
            if (dqkv) *reinterpret_cast<float2*>(dqkv + tkk[h] * 3 * gm.C + cbase + c) = make_float2(o[nt][2 * h], o[nt][2 * h + 1]);
            if (dqkv_sti) sti_store_pair_row(rb[h], r7[h], cbase + c, o[nt][2 * h], o[nt][2 * h + 1]);
          }
        }
      }
    }
    if (dqkv_sti && head == gm.heads - 1) sti_zero_padding(dqkv_sti, kbs3, tok, 3 * gm.C, t);
  }
  // bias-table gradient partial: [blockIdx.x][head][i][j]
  float* outp = partial + ((size_t)blockIdx.x * gm.heads + head) * AM_N * AM_N;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = row0 + g + 8 * h;
      *reinterpret_cast<float2*>(outp + i * AM_N + nt * 8 + tid * 2) = make_float2(dacc[nt][2 * h], dacc[nt][2 * h + 1]);
    }
  }
}

// ====================================================================================================
// Window-ordered operands ("WSTI"): the qkv contraction (and proj's dgrad for dO) writes its output as a split tile
// image whose ROWS are in window order (NsrConv.sti_win: roll + window_partition folded into the store) and whose
// CHANNELS pad every head to 32 (q | k | v groups of G = ceil(heads * 32, 64) channels; the padded weight rows are
// zero, so the padding is exactly 0).  A window's 64 tokens x one head pair (64 channels) of q, k or v is then ONE
// contiguous 8 KiB run of bf16 hi (+ 8 KiB lo 16 KiB further on): the kernels fetch their operands with six (forward)
// or eight (backward) bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) instead of gathering 120-byte
// fp32 row pieces and splitting them to bf16 in registers, and read fragments with ldmatrix from the 128-byte-swizzled
// rows (chunk c of row r at c ^ (r & 7): the image tcgen05 reads is also conflict-free for ldmatrix).
// The softmax scale multiplies the fp32 scores instead of q (same value up to one rounding).
constexpr int WT_BYTES = AM_N * 128;  // one tile: 64 rows x 128 B (a head PAIR; the CTA's head is one 64-byte half)

__device__ __forceinline__ const __nv_bfloat16* sw_ptr(const uint8_t* tile, int row, int chunk) {
  return reinterpret_cast<const __nv_bfloat16*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4));
}
// acc[16 x 64] = A[row0..+15][h-th 32 columns] B[0..63][same columns]^T; A, B: swizzled [64][64] bf16 tiles (hi, lo)
__device__ __forceinline__ void qk_scores_sw(const uint8_t* Ah, const uint8_t* Al, const uint8_t* Bh, const uint8_t* Bl,
                                             int hsel, int row0, int lane, float (&acc)[8][4]) {
  const int lr = lane & 7, lm = lane >> 3;
  const int arow = row0 + (lm & 1) * 8 + lr, ach = hsel * 4 + (lm >> 1);
  const int brow = (lm >> 1) * 8 + lr, bch = hsel * 4 + (lm & 1);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t ah[4], al[4];
    ldsm_x4(ah, sw_ptr(Ah, arow, ach + 2 * kk));
    ldsm_x4(al, sw_ptr(Al, arow, ach + 2 * kk));
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t bh[4], bl[4];
      ldsm_x4(bh, sw_ptr(Bh, brow + 16 * p, bch + 2 * kk));
      ldsm_x4(bl, sw_ptr(Bl, brow + 16 * p, bch + 2 * kk));
      mma3(acc[2 * p], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma3(acc[2 * p + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}
// out[16 x 32] = X[16 x 64] (accumulator layout) * B[k = token][h-th 32 columns] (ldmatrix.trans from the tile as stored)
__device__ __forceinline__ void acc_times_sw(const float (&x)[8][4], const uint8_t* Bh, const uint8_t* Bl, int hsel, int lane,
                                             float (&o)[4][4]) {
  const int lr = lane & 7, lm = lane >> 3;
  const int brow = (lm & 1) * 8 + lr, bch = hsel * 4 + (lm >> 1);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t ah[4], al[4];
    split_pair(x[2 * kk][0], x[2 * kk][1], ah[0], al[0]);
    split_pair(x[2 * kk][2], x[2 * kk][3], ah[1], al[1]);
    split_pair(x[2 * kk + 1][0], x[2 * kk + 1][1], ah[2], al[2]);
    split_pair(x[2 * kk + 1][2], x[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t bh[4], bl[4];
      ldsm_x4_t(bh, sw_ptr(Bh, brow + 16 * kk, bch + 2 * q));
      ldsm_x4_t(bl, sw_ptr(Bl, brow + 16 * kk, bch + 2 * q));
      mma3(o[2 * q], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma3(o[2 * q + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}
__device__ __forceinline__ void scale_acc(float (&a)[8][4], float s) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { a[nt][0] *= s; a[nt][1] *= s; a[nt][2] *= s; a[nt][3] *= s; }
}
// bulk-copy `n_ops` operands (hi + lo tile each) of window wi, head pair hp: operand j lives in 64-channel block
// kb0[j] + hp of the image `src[j]` whose rows have kbs[j] blocks
struct WstiOperand { const uint8_t* base; int kbs, kb0; };
__device__ __forceinline__ void wsti_load(uint8_t* tiles, const WstiOperand* ops, int n_ops, int wi, int hp, uint64_t* bar) {
  tc::mbar_arrive_expect_tx(bar, (uint32_t)n_ops * 2 * WT_BYTES);
  for (int j = 0; j < n_ops; ++j) {
    const uint8_t* src = ops[j].base + (((size_t)(wi >> 1) * ops[j].kbs + ops[j].kb0 + hp) << 15) + (wi & 1) * WT_BYTES;
    tc::bulk_g2s(tiles + (2 * j) * WT_BYTES, src, WT_BYTES, bar);
    tc::bulk_g2s(tiles + (2 * j + 1) * WT_BYTES, src + 16384, WT_BYTES, bar);
  }
}

__global__ void __launch_bounds__(AM_THREADS) window_attn_wsti_fwd_kernel(const uint8_t* __restrict__ qkv,
                                                                          const float* __restrict__ table,
                                                                          float* __restrict__ out, uint8_t* __restrict__ out_sti,
                                                                          AttnGeom gm, int G) {
  extern __shared__ __align__(128) uint8_t wt[];  // Q, K, V: hi, lo tiles
  __shared__ float bias_s[225];
  __shared__ __align__(8) int tok[AM_N], rid[AM_N];
  __shared__ __align__(8) uint64_t bar;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int wi = blockIdx.x / gm.heads, head = blockIdx.x - wi * gm.heads;  // heads fastest: a head pair shares its tiles in L2
  const int hp = head >> 1, hsel = head & 1;
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
    const int kbs = 3 * G / 64;
    const WstiOperand ops[3] = {{qkv, kbs, 0}, {qkv, kbs, G / 64}, {qkv, kbs, 2 * G / 64}};
    wsti_load(wt, ops, 3, wi, hp, &bar);
  }
  int differs = 0;
  if (t < AM_N) {
    attn_token_map(gm, wi, t, tok[t], rid[t]);
    int t0, r0;
    attn_token_map(gm, wi, 0, t0, r0);
    differs = rid[t] != r0;
  }
  for (int i = t; i < (2 * gm.ws - 1) * (2 * gm.ws - 1); i += AM_THREADS) bias_s[i] = table[i * gm.heads + head];
  const bool masked = __syncthreads_or(differs) && gm.use_mask && gm.shift > 0;  // also publishes the barrier init
  tc::mbar_wait(&bar, 0);
  const uint8_t *Qh = wt, *Ql = wt + WT_BYTES, *Kh = wt + 2 * WT_BYTES, *Kl = wt + 3 * WT_BYTES, *Vh = wt + 4 * WT_BYTES,
                *Vl = wt + 5 * WT_BYTES;
  const int row0 = warp * 16;
  float acc[8][4];
  qk_scores_sw(Qh, Ql, Kh, Kl, hsel, row0, lane, acc);
  scale_acc(acc, gm.scale);
  bias_mask_softmax(bias_s, rid, masked, row0, g, tid, acc);
  float o[4][4];
  acc_times_sw(acc, Vh, Vl, hsel, lane, o);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long tk = tok[row0 + g + 8 * h];
    uint8_t* rb = out_sti ? sti_row_base(out_sti, (gm.C + 63) / 64, tk) : nullptr;
    const int r7 = (int)(tk & 7);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + tid * 2;
      if (c < gm.D) {
        if (out) *reinterpret_cast<float2*>(out + (size_t)tk * gm.C + head * gm.D + c) = make_float2(o[nt][2 * h], o[nt][2 * h + 1]);
        if (out_sti) sti_store_pair_row(rb, r7, head * gm.D + c, o[nt][2 * h], o[nt][2 * h + 1]);
      }
    }
  }
  if (out_sti && head == gm.heads - 1) sti_zero_padding(out_sti, (gm.C + 63) / 64, tok, gm.C, t, true);
}

constexpr size_t WT_FWD_SMEM = 6 * WT_BYTES, WT_BWD_SMEM = 8 * WT_BYTES;

// backward: same two passes as window_attn_bwd_mma; operands Q, K, V (qkv image) and dO (proj-dgrad image) arrive by bulk copy
__global__ void __launch_bounds__(AM_THREADS, 3) window_attn_wsti_bwd_kernel(const uint8_t* __restrict__ qkv,
                                                                             const float* __restrict__ table,
                                                                             const uint8_t* __restrict__ dout,
                                                                             float* __restrict__ dqkv,
                                                                             uint8_t* __restrict__ dqkv_sti,
                                                                             float* __restrict__ partial, AttnGeom gm, int G,
                                                                             int nwin) {
  extern __shared__ __align__(128) uint8_t wt[];  // Q, K, V, dO: hi, lo tiles
  const uint8_t *Qh = wt, *Ql = wt + WT_BYTES, *Kh = wt + 2 * WT_BYTES, *Kl = wt + 3 * WT_BYTES, *Vh = wt + 4 * WT_BYTES,
                *Vl = wt + 5 * WT_BYTES, *Oh = wt + 6 * WT_BYTES, *Ol = wt + 7 * WT_BYTES;
  __shared__ float bias_s[225];
  __shared__ __align__(8) int tok[AM_N], rid[AM_N];
  __shared__ __align__(8) float lse_s[AM_N], delta_s[AM_N];
  __shared__ __align__(8) uint64_t bar;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tid = lane & 3;
  const int head = blockIdx.y, hp = head >> 1, hsel = head & 1;
  const int row0 = warp * 16;
  for (int i = t; i < (2 * gm.ws - 1) * (2 * gm.ws - 1); i += AM_THREADS) bias_s[i] = table[i * gm.heads + head];
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  float dacc[8][4];  // sum over this CTA's windows of dS, accumulator layout (fixed (i,j) per thread)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dacc[nt][0] = dacc[nt][1] = dacc[nt][2] = dacc[nt][3] = 0.f;
  const int kbs3 = (3 * gm.C + 63) / 64, kbsq = 3 * G / 64, kbso = G / 64;
  uint32_t phase = 0;

  for (int wi = blockIdx.x; wi < nwin; wi += gridDim.x) {
    __syncthreads();  // previous window's column pass is done with the tiles, tok, rid and the row statistics
    if (t == 0) {
      const WstiOperand ops[4] = {{qkv, kbsq, 0}, {qkv, kbsq, G / 64}, {qkv, kbsq, 2 * G / 64}, {dout, kbso, 0}};
      wsti_load(wt, ops, 4, wi, hp, &bar);
    }
    int differs = 0;
    if (t < AM_N) {
      attn_token_map(gm, wi, t, tok[t], rid[t]);
      int t0, r0;
      attn_token_map(gm, wi, 0, t0, r0);
      differs = rid[t] != r0;
    }
    const bool masked = __syncthreads_or(differs) && gm.use_mask && gm.shift > 0;
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    uint8_t* rb[2];
    int r7[2];
    long long tkk[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // rows row0+g, row0+g+8: the queries of the row pass and the keys of the column pass
      tkk[h] = tok[row0 + g + 8 * h];
      rb[h] = dqkv_sti ? sti_row_base(dqkv_sti, kbs3, tkk[h]) : nullptr;
      r7[h] = (int)(tkk[h] & 7);
    }
    float pr_[8][4], ds[8][4], o[4][4];
    // ---- row pass: P = softmax(scale Q K^T + bias + mask); dP = dO V^T; dS = P o (dP - rowsum(P o dP)); dQ = scale dS K
    {
      float lse[2];
      qk_scores_sw(Qh, Ql, Kh, Kl, hsel, row0, lane, pr_);
      scale_acc(pr_, gm.scale);
      bias_mask_softmax<true>(bias_s, rid, masked, row0, g, tid, pr_, lse);
      qk_scores_sw(Oh, Ol, Vh, Vl, hsel, row0, lane, ds);  // dP[i][j] = sum_d dO[i][d] V[j][d]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float delta = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) delta += pr_[nt][2 * h] * ds[nt][2 * h] + pr_[nt][2 * h + 1] * ds[nt][2 * h + 1];
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
        if (tid == 0) { lse_s[row0 + g + 8 * h] = lse[h]; delta_s[row0 + g + 8 * h] = delta; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float v = pr_[nt][2 * h + e] * (ds[nt][2 * h + e] - delta);
            ds[nt][2 * h + e] = v;
            dacc[nt][2 * h + e] += v;
          }
        }
      }
      acc_times_sw(ds, Kh, Kl, hsel, lane, o);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = nt * 8 + tid * 2;
          if (c < gm.D) {
            const float v0 = o[nt][2 * h] * gm.scale, v1 = o[nt][2 * h + 1] * gm.scale;
            if (dqkv) *reinterpret_cast<float2*>(dqkv + tkk[h] * 3 * gm.C + head * gm.D + c) = make_float2(v0, v1);
            if (dqkv_sti) sti_store_pair_row(rb[h], r7[h], head * gm.D + c, v0, v1);
          }
        }
      }
    }
    __syncthreads();  // lse / delta of all 64 queries are visible
    // ---- column pass (rows = keys j, columns = queries i = 8 nt + 2 tid + e): P^T, dP^T, dS^T -> dK, dV
    qk_scores_sw(Kh, Kl, Qh, Ql, hsel, row0, lane, pr_);   // S^T[j][i] = sum_d K[j][d] Q[i][d]  (x scale below)
    qk_scores_sw(Vh, Vl, Oh, Ol, hsel, row0, lane, ds);    // dP^T[j][i] = sum_d V[j][d] dO[i][d]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = row0 + g + 8 * h;
      const int base = 15 * (7 - (j >> 3)) + 7 - (j & 7) + 2 * tid;
      const int rj = rid[j];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 ls = *reinterpret_cast<const float2*>(&lse_s[nt * 8 + tid * 2]);
        const float2 dl = *reinterpret_cast<const float2*>(&delta_s[nt * 8 + tid * 2]);
        float s0 = pr_[nt][2 * h] * gm.scale + bias_s[base + 15 * nt], s1 = pr_[nt][2 * h + 1] * gm.scale + bias_s[base + 15 * nt + 1];
        if (masked) {
          const int2 ri = *reinterpret_cast<const int2*>(&rid[nt * 8 + tid * 2]);
          if (ri.x != rj) s0 += -100.0f;
          if (ri.y != rj) s1 += -100.0f;
        }
        const float p0 = __expf(s0 - ls.x), p1 = __expf(s1 - ls.y);
        pr_[nt][2 * h] = p0;
        pr_[nt][2 * h + 1] = p1;
        ds[nt][2 * h] = p0 * (ds[nt][2 * h] - dl.x);
        ds[nt][2 * h + 1] = p1 * (ds[nt][2 * h + 1] - dl.y);
      }
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      if (which == 0) acc_times_sw(ds, Qh, Ql, hsel, lane, o);   // dK[j][d] = scale sum_i dS[i][j] Q[i][d]
      else acc_times_sw(pr_, Oh, Ol, hsel, lane, o);             // dV[j][d] = sum_i P[i][j] dO[i][d]
      const float sc = which == 0 ? gm.scale : 1.f;
      const int cbase = (which == 0 ? gm.C : 2 * gm.C) + head * gm.D;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = nt * 8 + tid * 2;
          if (c < gm.D) {
            const float v0 = o[nt][2 * h] * sc, v1 = o[nt][2 * h + 1] * sc;
            if (dqkv) *reinterpret_cast<float2*>(dqkv + tkk[h] * 3 * gm.C + cbase + c) = make_float2(v0, v1);
            if (dqkv_sti) sti_store_pair_row(rb[h], r7[h], cbase + c, v0, v1);
          }
        }
      }
    }
    if (dqkv_sti && head == gm.heads - 1) sti_zero_padding(dqkv_sti, kbs3, tok, 3 * gm.C, t);
  }
  // bias-table gradient partial: [blockIdx.x][head][i][j]
  float* outp = partial + ((size_t)blockIdx.x * gm.heads + head) * AM_N * AM_N;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = row0 + g + 8 * h;
      *reinterpret_cast<float2*>(outp + i * AM_N + nt * 8 + tid * 2) = make_float2(dacc[nt][2 * h], dacc[nt][2 * h + 1]);
    }
  }
}

int window_attn_wsti_fwd_launch(const void* qkv, const float* table, float* out, void* out_sti, int batch, int h, int w,
                                int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st) {
  AttnGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, scale};
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_wsti_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_FWD_SMEM);
    if (e != cudaSuccess) {
      set_error("window_attn_wsti_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int G = (heads * 32 + 63) / 64 * 64;
  dim3 grid(batch * g.nwh * g.nww * heads);
  window_attn_wsti_fwd_kernel<<<grid, AM_THREADS, WT_FWD_SMEM, st>>>(reinterpret_cast<const uint8_t*>(qkv), table, out,
                                                                    reinterpret_cast<uint8_t*>(out_sti), g, G);
  NSR_CHECK_LAUNCH("window_attn_wsti_fwd");
  return NSR_OK;
}

int window_attn_wsti_bwd_launch(const void* qkv, const float* table, const void* dout, float* dqkv, void* dqkv_sti,
                                float* partial, int gx, int batch, int h, int w, int c, int heads, int ws, int shift,
                                int use_mask, float scale, cudaStream_t st) {
  AttnGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, scale};
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_wsti_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_BWD_SMEM);
    if (e != cudaSuccess) {
      set_error("window_attn_wsti_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int G = (heads * 32 + 63) / 64 * 64;
  dim3 grid(gx, heads);
  window_attn_wsti_bwd_kernel<<<grid, AM_THREADS, WT_BWD_SMEM, st>>>(
      reinterpret_cast<const uint8_t*>(qkv), table, reinterpret_cast<const uint8_t*>(dout), dqkv,
      reinterpret_cast<uint8_t*>(dqkv_sti), partial, g, G, batch * g.nwh * g.nww);
  NSR_CHECK_LAUNCH("window_attn_wsti_bwd");
  return NSR_OK;
}

// ------------------------------------------------------------------------------------ host
int window_attn_bwd_mma_ctas() { return 3; }  // persistent CTAs per SM of the backward kernel (grid and workspace follow it)
bool window_attn_mma_supported(int c, int heads, int ws) {
  const int d = c / heads;
  return ws == 8 && d <= 32 && d % 2 == 0 && (c % 2 == 0);
}

int window_attn_fwd_mma_launch(const float* qkv, const float* table, float* out, void* out_sti, int batch, int h, int w,
                               int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st) {
  AttnGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, scale};
  dim3 grid(batch * g.nwh * g.nww * heads);
  window_attn_fwd_mma<<<grid, AM_THREADS, 0, st>>>(qkv, table, out, reinterpret_cast<uint8_t*>(out_sti), g);
  NSR_CHECK_LAUNCH("window_attn_fwd_mma");
  return NSR_OK;
}

int window_attn_bwd_mma_launch(const float* qkv, const float* table, const float* dout, float* dqkv, void* dqkv_sti,
                               float* partial, int gx, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask,
                               float scale, cudaStream_t st) {
  AttnGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, scale};
  dim3 grid(gx, heads);
  window_attn_bwd_mma<<<grid, AM_THREADS, BW2_SMEM, st>>>(qkv, table, dout, dqkv, reinterpret_cast<uint8_t*>(dqkv_sti), partial,
                                                          g, batch * g.nwh * g.nww);
  NSR_CHECK_LAUNCH("window_attn_bwd_mma");
  return NSR_OK;
}

}  // namespace nsr

// Backward of the (shifted-)window attention core on tcgen05: all five products of the attention gradient are UMMA
// instructions with TMEM accumulators, operands are the window-ordered, head-padded split tile images (qkv from the
// forward pass, dO from the proj dgrad contraction) fetched with four 32 KiB bulk copies per work item.
//
// Work item = (window pair, head pair), as in the forward kernel (window_attn_tc.cu).  A CTA keeps ONE head pair for all
// its items (grid = gx x head pairs), so a thread's bias-gradient accumulators belong to one fixed (query, key) set.
//
//   phase 1 (M = 128 rows = 2 windows x 64 queries, N = 128, per head h):
//        S_h  = Q[:, 32h..] K[:, 32h..]^T          dP_h = dO[:, 32h..] V[:, 32h..]^T          -> TMEM columns 0..511
//   softmax warps (two threads per query row, each 32 of the 64 own-window keys, one (max, sum, delta) merge):
//        P = softmax(scale S + bias + mask),  delta = sum_j P_j dP_j,  dS = P o (dP - delta),  dB += dS
//        P, then dS, are written (bf16 hi / lo) as tiles T[window][(head, query)][64 keys]: 128 rows x 128 B per window
//   phase 2 (per window w; M = 128 rows = (head, key) or (head, query), N = 64 channels of the head pair, K = 64):
//        dV_w = T_P[w]^T  dO[rows of w]     A = T as an MN-major operand: K index = query, two 64-key panels = the two heads
//        dQ_w = T_dS[w]   K[rows of w]      A = T K-major
//        dK_w = T_dS[w]^T Q[rows of w]      A = T MN-major           (B = MN-major view of the operand block: the wgrad layout)
//        row (h, x) of an output holds head h's 32 channels in columns 32h..32h+31; the other half is never read
//   epilogue: dQ, dK (x scale), dV -> the head-padded split tile image dqkv [tokens, 3G] in token order (16-byte stores),
//        consumed by the qkv dgrad / wgrad contractions on head-padded weights.
//
// TMEM is time-shared: phase-2 accumulators overlay the phase-1 columns once every thread has read them.  Shared memory:
// four operand blocks (128 KiB) + one 64 KiB tile region used for P and then for dS; the V and dO blocks are released as soon
// as their last product retires, Q and K at the end of the item.
#include "attn_tc.cuh"

namespace nsr {
using namespace tc;

constexpr int AB_THREADS = 576;               // loader + MMA issuer + 16 softmax warps
constexpr int AB_SMEM_T = 4 * AT_BLK;         // P / dS tiles: 2 windows x (hi 16 KiB + lo 16 KiB)
constexpr int AB_SMEM_BAR = 6 * AT_BLK;
constexpr int AB_SMEM_XCH = AB_SMEM_BAR + 256;  // float4 [2 parities][2 heads][128 rows][2 halves] = 16 KiB
constexpr size_t AB_SMEM = 6 * AT_BLK + 256 + 16384 + 1024;

struct AbBars {
  uint64_t in_full[4], in_empty[4];  // Q, K, V, dO blocks
  uint64_t s_full, p_full, dv_done, ds_full, out_full, out_empty;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(AB_THREADS, 1) window_attn_tc_bwd_kernel(const uint8_t* __restrict__ qkv,
                                                                           const float* __restrict__ table,
                                                                           const uint8_t* __restrict__ dout,
                                                                           float* __restrict__ dqkv, uint8_t* __restrict__ dqkv_sti,
                                                                           float* __restrict__ partial, AtGeom gm, int gx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  AbBars* bars = reinterpret_cast<AbBars*>(smem + AB_SMEM_BAR);
  __shared__ float bias_s[2 * 225];
  __shared__ int rid_s[128];
  __shared__ int tok_s[128];  // token of row (window, x) of the current pair, -1 beyond the last window
  __shared__ int mask_s[4];   // per 32-row quarter: its rows lie in more than one shift-mask region
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_wp = (gm.nwin + 1) >> 1, n_hp = (gm.heads + 1) >> 1;
  const int hp = blockIdx.x % n_hp, b = blockIdx.x / n_hp;  // this CTA's head pair and its slot among the gx CTAs of the pair
  const bool head_b = hp * 2 + 1 < gm.heads;                // odd head count: the last pair has one head
  const int kbs = 3 * gm.G / 64, kbs_o = gm.G / 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->in_full[i], 1);
      mbar_init(&bars->in_empty[i], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 512);
    mbar_init(&bars->dv_done, 1);
    mbar_init(&bars->ds_full, 512);
    mbar_init(&bars->out_full, 1);
    mbar_init(&bars->out_empty, 512);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * 225; i += AB_THREADS) {  // bias tables of the two heads, pre-multiplied by log2(e)
    const int head = hp * 2 + i / 225;
    bias_s[i] = head < gm.heads ? table[(i % 225) * gm.heads + head] * 1.4426950408889634f : 0.f;
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ================================ loader =============================================
    if (lane == 0) {
      uint32_t par = 0;
      for (int wp = b; wp < n_wp; wp += gx, par ^= 1) {
        const uint8_t* qrow = qkv + ((size_t)wp * kbs << 15);
        const uint8_t* src[4] = {qrow + ((size_t)hp << 15), qrow + ((size_t)(gm.G / 64 + hp) << 15),
                                 qrow + ((size_t)(2 * gm.G / 64 + hp) << 15), dout + ((size_t)(wp * kbs_o + hp) << 15)};
        const int order[4] = {2, 3, 0, 1};  // V and dO are released first by the previous item
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int sl = order[k];
          mbar_wait<32>(&bars->in_empty[sl], par ^ 1);
          mbar_arrive_expect_tx(&bars->in_full[sl], AT_BLK);
          bulk_g2s(smem + sl * AT_BLK, src[sl], AT_BLK, &bars->in_full[sl]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc_1 = umma_idesc_bf16(128, 0, 0);  // phase 1: A, B K-major, N = 128
      constexpr uint32_t idesc_t = umma_idesc_bf16(64, 1, 1);   // dV, dK: A = T MN-major, B MN-major, N = 64
      constexpr uint32_t idesc_q = umma_idesc_bf16(64, 0, 1);   // dQ: A = T K-major, B MN-major, N = 64
      const uint32_t sQ = smem_u32(smem), sK = sQ + AT_BLK, sV = sQ + 2 * AT_BLK, sO = sQ + 3 * AT_BLK, sT = sQ + AB_SMEM_T;
      // three-pass product: D (+)= (A_hi + A_lo)(B_hi + B_lo) without lo.lo; `ks` K-steps, descriptor advances da / db per step
      auto mma3x = [&](uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, int ks, int da, int db) {
        for (int k = 0; k < ks; ++k) umma_bf16(d, a_hi + da * k, b_hi + db * k, idesc, k != 0);
        for (int k = 0; k < ks; ++k) umma_bf16(d, a_hi + da * k, b_lo + db * k, idesc, 1);
        for (int k = 0; k < ks; ++k) umma_bf16(d, a_lo + da * k, b_hi + db * k, idesc, 1);
      };
      uint32_t par = 0;
      for (int wp = b; wp < n_wp; wp += gx, par ^= 1) {
        for (int i = 0; i < 4; ++i) mbar_wait(&bars->in_full[i], par);
        mbar_wait(&bars->out_empty, par ^ 1);  // the previous item's accumulators have been read out of TMEM
        tc_fence_after();
        // ---- phase 1
        for (int h = 0; h < (head_b ? 2 : 1); ++h) {
          mma3x(tmem_base + h * 128, umma_desc_sw128(sQ, 1, 64) + 4 * h, umma_desc_sw128(sQ + 16384, 1, 64) + 4 * h,
                umma_desc_sw128(sK, 1, 64) + 4 * h, umma_desc_sw128(sK + 16384, 1, 64) + 4 * h, idesc_1, 2, 2, 2);
          mma3x(tmem_base + 256 + h * 128, umma_desc_sw128(sO, 1, 64) + 4 * h, umma_desc_sw128(sO + 16384, 1, 64) + 4 * h,
                umma_desc_sw128(sV, 1, 64) + 4 * h, umma_desc_sw128(sV + 16384, 1, 64) + 4 * h, idesc_1, 2, 2, 2);
        }
        umma_commit(&bars->s_full);
        umma_commit(&bars->in_empty[2]);  // V
        // ---- dV = P^T dO per window (T holds P)
        mbar_wait(&bars->p_full, par);
        tc_fence_after();
        for (int w = 0; w < 2; ++w)
          mma3x(tmem_base + w * 64, umma_desc_sw128(sT + w * AT_BLK, 512, 64), umma_desc_sw128(sT + w * AT_BLK + 16384, 512, 64),
                umma_desc_sw128(sO + w * 8192, 512, 64), umma_desc_sw128(sO + 16384 + w * 8192, 512, 64), idesc_t, 4, 128, 128);
        umma_commit(&bars->dv_done);      // T may be overwritten with dS
        umma_commit(&bars->in_empty[3]);  // dO
        // ---- dQ = dS K, dK = dS^T Q per window (T holds dS)
        mbar_wait(&bars->ds_full, par);
        tc_fence_after();
        for (int w = 0; w < 2; ++w) {
          mma3x(tmem_base + 128 + w * 64, umma_desc_sw128(sT + w * AT_BLK, 1, 64), umma_desc_sw128(sT + w * AT_BLK + 16384, 1, 64),
                umma_desc_sw128(sK + w * 8192, 512, 64), umma_desc_sw128(sK + 16384 + w * 8192, 512, 64), idesc_q, 4, 2, 128);
          mma3x(tmem_base + 256 + w * 64, umma_desc_sw128(sT + w * AT_BLK, 512, 64), umma_desc_sw128(sT + w * AT_BLK + 16384, 512, 64),
                umma_desc_sw128(sQ + w * 8192, 512, 64), umma_desc_sw128(sQ + 16384 + w * 8192, 512, 64), idesc_t, 4, 128, 128);
        }
        umma_commit(&bars->out_full);
        umma_commit(&bars->in_empty[0]);
        umma_commit(&bars->in_empty[1]);
      }
    }
  } else {
    // ================================ softmax / dS warps + epilogue ========================
    const int cw = warp - 2;
    const int h = (cw >> 2) & 1;            // head A / B (phase 1); window 0 / 1 in the epilogue
    const int half = cw >> 3;               // 32 of the row's 64 keys (phase 1); 16 of the 32 output channels (epilogue)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // TMEM lane: (window, query) in phase 1, (head, key | query) in phase 2
    const int win = r >> 6, qi = r & 63;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int base_i = 15 * (qi >> 3) + (qi & 7) + 112 - 60 * half;
    const float scale_l2 = gm.scale * 1.4426950408889634f;
    float4* xch = reinterpret_cast<float4*>(smem + AB_SMEM_XCH);
    const bool head_ok = h == 0 || head_b;                 // does this thread's phase-1 head exist
    const int e_head = hp * 2 + win;                        // epilogue: head of TMEM lane r = (head-in-pair, x)
    const bool e_head_ok = e_head < gm.heads;
    float acc[32];  // bias-table gradient: sum over this CTA's items (and this thread's window) of dS[qi][32 half + j]
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    uint8_t* trow = smem + AB_SMEM_T + win * AT_BLK + (h * 64 + qi) * 128;  // this thread's row of T[window][(head, query)]
    uint32_t par = 0;
    for (int wp = b; wp < n_wp; wp += gx, par ^= 1) {
      // token index / shift-mask region of the pair's 128 rows: one thread per row computes them (integer divisions),
      // everybody reads them from shared memory; a window is "masked" if its 64 tokens do not share one region
      const bool e_valid = wp * 2 + h < gm.nwin;  // epilogue: this thread stores rows of window h of the pair
      named_bar_sync(1, 512);  // every thread is done with the previous pair's tokens / region ids
      if (h == 0 && half == 0) {
        int tok = -1, rid0 = 0;
        if (wp * 2 + win < gm.nwin) at_token_map(gm, wp * 2 + win, qi, tok, rid0);
        rid_s[r] = rid0;
        tok_s[r] = tok;
        const unsigned differs = __ballot_sync(0xffffffffu, rid0 != __shfl_sync(0xffffffffu, rid0, 0));
        if (lane == 0) mask_s[q] = differs != 0;  // per 32-row quarter; quarters of a window are combined below
      }
      named_bar_sync(1, 512);
      const int rid = rid_s[r];
      const int e_tok = tok_s[h * 64 + qi];
      bool masked = false;
      if (gm.use_mask && gm.shift > 0)
        masked = mask_s[win * 2] || mask_s[win * 2 + 1] || rid_s[win * 64] != rid_s[win * 64 + 32];
      // ---- phase 1 results -> P, delta
      mbar_wait(&bars->s_full, par);
      tc_fence_after();
      float s[32];
      float m = 0.f, l = 1.f, du = 0.f;
      if (head_ok) {
        tmem_ld_32x32(lane_addr + h * 128 + win * 64 + half * 32, s);
        const float* bt = bias_s + h * 225 + base_i;
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = fmaf(s[j], scale_l2, bt[-(15 * (j >> 3) + (j & 7))]);
        if (masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (rid_s[win * 64 + half * 32 + j] != rid) s[j] += -100.0f * 1.4426950408889634f;
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] = ex2_approx(s[j] - m);
          l4[j & 3] += s[j];
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float dp[16];
          tmem_ld_32x16(lane_addr + 256 + h * 128 + win * 64 + half * 32 + c * 16, dp);
#pragma unroll
          for (int j = 0; j < 16; ++j) d4[j & 3] = fmaf(s[16 * c + j], dp[j], d4[j & 3]);
        }
        du = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      // merge the two halves of the row: P = e 2^(m-M) / Z, delta = sum over both halves of P dP
      float4* xs = xch + ((par * 2 + h) * 128 + r) * 2;
      xs[half] = make_float4(m, l, du, 0.f);
      named_bar_sync(2 + h, 256);
      const float4 ot = xs[half ^ 1];
      const float M = fmaxf(m, ot.x);
      const float f = ex2_approx(m - M), fo = ex2_approx(ot.x - M);
      const float invz = 1.f / fmaf(l, f, ot.y * fo);
      const float ct = f * invz;
      const float delta = fmaf(ct, du, fo * invz * ot.z);
      if (head_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] *= ct;  // P
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 hi, lo;
          split_pair(s[8 * c], s[8 * c + 1], hi.x, lo.x);
          split_pair(s[8 * c + 2], s[8 * c + 3], hi.y, lo.y);
          split_pair(s[8 * c + 4], s[8 * c + 5], hi.z, lo.z);
          split_pair(s[8 * c + 6], s[8 * c + 7], hi.w, lo.w);
          const int off = ((half * 4 + c) ^ (qi & 7)) << 4;
          *reinterpret_cast<uint4*>(trow + off) = hi;
          *reinterpret_cast<uint4*>(trow + 16384 + off) = lo;
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(&bars->p_full);
      // ---- dS = P o (dP - delta)   (second pass over dP: it is still in TMEM, and P can be overwritten in place)
      if (head_ok) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float dp[16];
          tmem_ld_32x16(lane_addr + 256 + h * 128 + win * 64 + half * 32 + c * 16, dp);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = s[16 * c + j] * (dp[j] - delta);
            s[16 * c + j] = v;
            acc[16 * c + j] += v;
          }
        }
      }
      tc_fence_before();
      mbar_wait(&bars->dv_done, par);  // P^T dO has retired: the tile region is free for dS
      if (head_ok) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 hi, lo;
          split_pair(s[8 * c], s[8 * c + 1], hi.x, lo.x);
          split_pair(s[8 * c + 2], s[8 * c + 3], hi.y, lo.y);
          split_pair(s[8 * c + 4], s[8 * c + 5], hi.z, lo.z);
          split_pair(s[8 * c + 6], s[8 * c + 7], hi.w, lo.w);
          const int off = ((half * 4 + c) ^ (qi & 7)) << 4;
          *reinterpret_cast<uint4*>(trow + off) = hi;
          *reinterpret_cast<uint4*>(trow + 16384 + off) = lo;
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(&bars->ds_full);
      // ---- epilogue: lane r = (head-in-pair `win`, token x = qi) of window `h`; this thread's 16 channels of dQ, dK, dV.
      // The values go through shared memory (the tile region is idle now) so that the global stores are whole 128-byte
      // rows of the image: 4 rows per warp instruction instead of 32 scattered 16-byte pieces (the direct stores were 27 %
      // of the kernel, profiles/r02_m_ncu_attn_tc_bwd.txt).
      mbar_wait(&bars->out_full, par);
      tc_fence_after();
      {
        uint8_t* stage = smem + AB_SMEM_T;  // 2 buffers x [128 token rows (window, x)][hi 128 B | lo 128 B], chunks swizzled by x
        const int c0 = half * 16;
        auto stage_out = [&](int s3, int buf) {
          float o[16];
          tmem_ld_32x16(lane_addr + (s3 == 0 ? 128 : (s3 == 1 ? 256 : 0)) + h * 64 + win * 32 + half * 16, o);
          const float sc = s3 == 2 ? 1.f : gm.scale;
          if (dqkv && e_tok >= 0 && e_head_ok) {  // fp32 output (tests)
            float* dst = dqkv + (size_t)e_tok * 3 * gm.C + s3 * gm.C + e_head * gm.D + c0;
#pragma unroll
            for (int c = 0; c < 16; c += 2)
              if (c0 + c < gm.D) *reinterpret_cast<float2*>(dst + c) = make_float2(o[c] * sc, o[c + 1] * sc);
          }
          uint8_t* srow = stage + buf * AT_BLK + (h * 64 + qi) * 256;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (e_head_ok) {  // a head beyond `heads` (odd head count) is padding: zeros
              split_pair(o[8 * j] * sc, o[8 * j + 1] * sc, hi.x, lo.x);
              split_pair(o[8 * j + 2] * sc, o[8 * j + 3] * sc, hi.y, lo.y);
              split_pair(o[8 * j + 4] * sc, o[8 * j + 5] * sc, hi.z, lo.z);
              split_pair(o[8 * j + 6] * sc, o[8 * j + 7] * sc, hi.w, lo.w);
            }
            const int off = ((win * 4 + half * 2 + j) ^ (qi & 7)) << 4;
            *reinterpret_cast<uint4*>(srow + off) = hi;
            *reinterpret_cast<uint4*>(srow + 128 + off) = lo;
          }
        };
        auto copy_out = [&](int s3, int buf) {
          if (!dqkv_sti) return;
          const int t = threadIdx.x - 64;  // 0..511
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int idx = it * 512 + t, row = idx >> 4, slot = idx & 15, c = slot & 7;
            const int tk = tok_s[row];
            if (tk < 0) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(stage + buf * AT_BLK + row * 256 + (slot >> 3) * 128 + ((c ^ (row & 7)) << 4));
            uint8_t* dst = dqkv_sti + ((size_t)((tk >> 7) * kbs + s3 * kbs_o + hp) << 15) + (slot >> 3) * 16384 +
                           (size_t)(tk & 127) * 128 + ((c ^ (tk & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = v;
          }
        };
        stage_out(0, 0);
        stage_out(1, 1);
        named_bar_sync(1, 512);
        copy_out(0, 0);
        copy_out(1, 1);
        named_bar_sync(1, 512);
        stage_out(2, 0);
        tc_fence_before();
        mbar_arrive(&bars->out_empty);
        named_bar_sync(1, 512);
        copy_out(2, 0);
        named_bar_sync(1, 512);  // the tile region is free again for the next item's P
      }
    }
    // ---- bias-table gradient partial of this CTA: sum the two windows' accumulators, [b][head][i][j]
    named_bar_sync(1, 512);  // all tiles consumed: the tile region is free as scratch
    float* scr = reinterpret_cast<float*>(smem + AB_SMEM_T);  // [head][query][64 keys] fp32 = 32 KiB
    float* mine = scr + ((h * 64 + qi) * 64 + half * 32);
    if (win == 1) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(mine + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    named_bar_sync(1, 512);
    if (win == 0 && head_ok) {
      float* outp = partial + ((size_t)b * gm.heads + hp * 2 + h) * 4096 + qi * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 o4 = *reinterpret_cast<const float4*>(mine + j);
        *reinterpret_cast<float4*>(outp + j) = make_float4(acc[j] + o4.x, acc[j + 1] + o4.y, acc[j + 2] + o4.z, acc[j + 3] + o4.w);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int window_attn_tc_bwd_gx(int heads) {
  const int n_hp = (heads + 1) / 2;
  const int gx = kNumSMs / n_hp;
  return gx < 1 ? 1 : gx;
}

int window_attn_tc_bwd_launch(const void* qkv, const float* table, const void* dout, float* dqkv, void* dqkv_sti, float* partial,
                              int gx, int batch, int h, int w, int c, int heads, int ws, int shift, int use_mask, float scale,
                              cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM);
    if (e != cudaSuccess) {
      set_error("window_attn_tc_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  AtGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, (heads * 32 + 63) / 64 * 64,
           batch * (h / ws) * (w / ws), scale, 1};
  const int n_hp = (heads + 1) / 2;
  window_attn_tc_bwd_kernel<<<gx * n_hp, AB_THREADS, AB_SMEM, st>>>(reinterpret_cast<const uint8_t*>(qkv), table,
                                                                  reinterpret_cast<const uint8_t*>(dout), dqkv,
                                                                  reinterpret_cast<uint8_t*>(dqkv_sti), partial, g, gx);
  NSR_CHECK_LAUNCH("window_attn_tc_bwd");
  return NSR_OK;
}

}  // namespace nsr

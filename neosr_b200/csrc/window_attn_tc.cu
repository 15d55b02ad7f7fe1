// (Shifted-)window attention core on the 5th-generation tensor cores: Q K^T and P V are tcgen05.mma instructions with
// their accumulators in TMEM; operands arrive by bulk copy (cp.async.bulk -> UBLKCP) from the window-ordered,
// head-padded split tile image the qkv contraction wrote (NsrConv.sti_win, see window_attn_mma.cu / neosr_b200.h).
//
// Work item = (window PAIR, head PAIR): the 128 rows of one STI block row are the 2 x 64 tokens of two windows, a
// 64-channel block holds two heads padded to 32 channels - so Q, K, V of an item are three whole 32 KiB blocks
// (bf16 hi image + lo image), each ONE bulk copy, already in the SWIZZLE_128B layout UMMA reads:
//
//   S_h [128 x 128] = Q[:, 32h..32h+31] K[:, 32h..]^T   A, B K-major; 2 K-steps x (hi.hi + hi.lo + lo.hi)     -> TMEM
//        rows 0..63 x columns 0..63 = window 0, rows 64..127 x columns 64..127 = window 1 (the off-diagonal
//        quarters cost tensor time that is otherwise idle and are never read)
//   softmax (one thread per query row: TMEM lane = row, so max / sum are thread-local), relative-position bias from
//        a 225-entry table in shared memory, {0,-100} shift mask from region ids, then P -> bf16 hi / lo, stored as the
//        K-major A tile of the next product: [128 rows][64 own-window keys]
//   O_h = P_h V: D1 = P_h V[keys of window 0], D2 = P_h V[keys of window 1]  (B = MN-major view of the V block, the
//        wgrad layout); rows 0..63 take D1, rows 64..127 take D2 - the block-diagonal product without storing zeros
//
// Warp roles (320 threads, one CTA per SM, persistent over window pairs):
//   warp 0     loader      : bulk copies, one thread
//   warp 1     MMA issuer  : one thread; issues S of item i+1 before P V of item i, so the softmax warps never wait
//                            for scores
//   warps 2-5  head A, 6-9 head B: softmax of item i, then the epilogue (O -> split tile image in token order) of
//                            item i-1, whose P V ran meanwhile
// TMEM: S_A, S_B (2 x 128 columns) + O_A, O_B (2 x (D1 | D2) = 2 x 128) = 512 columns.
#include "attn_tc.cuh"

namespace nsr {
using namespace tc;

constexpr int AT_THREADS = 576;                   // loader + MMA issuer + 16 softmax warps
constexpr int AT_SLOTS = 4;                        // operand slots: Q, K, V (even items), V (odd items)
constexpr int AT_SMEM_P = AT_SLOTS * AT_BLK;       // P tiles: 2 heads x 32 KiB
constexpr int AT_SMEM_BAR = (AT_SLOTS + 2) * AT_BLK;
constexpr int AT_SMEM_XCH = AT_SMEM_BAR + 256;     // float2 [2 parities][2 heads][128 rows][2 halves] = 8 KiB
constexpr size_t AT_FWD_SMEM = (AT_SLOTS + 2) * AT_BLK + 256 + 8192 + 1024;  // + barriers + exchange + alignment slack

struct AtBars {
  uint64_t full[AT_SLOTS], empty[AT_SLOTS];
  uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(AT_THREADS, 1) window_attn_tc_fwd_kernel(const uint8_t* __restrict__ qkv,
                                                                           const float* __restrict__ table,
                                                                           float* __restrict__ out, uint8_t* __restrict__ out_sti,
                                                                           AtGeom gm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  AtBars* bars = reinterpret_cast<AtBars*>(smem + AT_SMEM_BAR);
  __shared__ float bias_s[AT_MAX_HEADS * 225];
  __shared__ int rid_s[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_wp = (gm.nwin + 1) >> 1, n_hp = (gm.heads + 1) >> 1, kbs = 3 * gm.G / 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < AT_SLOTS; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&bars->s_full[h], 1);
      mbar_init(&bars->s_empty[h], 256);
      mbar_init(&bars->p_full[h], 256);
      mbar_init(&bars->p_empty[h], 1);
      mbar_init(&bars->o_full[h], 1);
      mbar_init(&bars->o_empty[h], 256);
    }
    fence_mbar_init();
  }
  // bias table per head, pre-multiplied by log2(e): the softmax runs on exp2
  for (int i = threadIdx.x; i < gm.heads * 225; i += AT_THREADS)
    bias_s[i] = table[(i % 225) * gm.heads + i / 225] * 1.4426950408889634f;
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ================================ loader =============================================
    // Four 32 KiB operand slots: Q and K of the next item (released as soon as its two S products retire, which the MMA
    // thread issues one item ahead) and a double-buffered V (held until both P V products of its item retire).  Every
    // block therefore has a whole item's duration to arrive.  (A single V slot - and, worse, a plain ring in consumption
    // order - exposes the load latency once per item: the kernel then ran at ~5.5 us per item instead of ~1,
    // profiles/r02_g_ncu_attn_tc_fwd.txt, r02_i.)
    if (lane == 0) {
      int item = 0;
      for (int wp = blockIdx.x; wp < n_wp; wp += gridDim.x) {
        const uint8_t* row = qkv + ((size_t)wp * kbs << 15);
        for (int hp = 0; hp < n_hp; ++hp, ++item) {
          const uint32_t par = item & 1, par2 = (item >> 1) & 1;
          mbar_wait<32>(&bars->empty[0], par ^ 1);
          mbar_arrive_expect_tx(&bars->full[0], AT_BLK);
          bulk_g2s(smem, row + ((size_t)hp << 15), AT_BLK, &bars->full[0]);
          mbar_wait<32>(&bars->empty[1], par ^ 1);
          mbar_arrive_expect_tx(&bars->full[1], AT_BLK);
          bulk_g2s(smem + AT_BLK, row + ((size_t)(gm.G / 64 + hp) << 15), AT_BLK, &bars->full[1]);
          const int vs = 2 + (item & 1);
          mbar_wait<32>(&bars->empty[vs], par2 ^ 1);
          mbar_arrive_expect_tx(&bars->full[vs], AT_BLK);
          bulk_g2s(smem + vs * AT_BLK, row + ((size_t)(2 * gm.G / 64 + hp) << 15), AT_BLK, &bars->full[vs]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 0, 0);  // S: A, B K-major, N = 128
      constexpr uint32_t idesc_o = umma_idesc_bf16(64, 0, 1);   // O: A K-major (P), B MN-major (V), N = 64
      const uint32_t s0 = smem_u32(smem);
      int my_items = 0;
      for (int wp = blockIdx.x; wp < n_wp; wp += gridDim.x) my_items += n_hp;
      uint32_t ph_sh[2] = {0, 0}, ph_oh[2] = {0, 0};  // per head: a head beyond `heads` (odd head count) skips its items
      auto head_valid = [&](int item, int h) { return (item % n_hp) * 2 + h < gm.heads; };
      auto issue_s = [&](int item) {
        mbar_wait(&bars->full[0], item & 1);
        mbar_wait(&bars->full[1], item & 1);
        tc_fence_after();
        const uint32_t sq = s0, sk = s0 + AT_BLK;
        for (int h = 0; h < 2; ++h) {
          if (!head_valid(item, h)) continue;
          mbar_wait(&bars->s_empty[h], ph_sh[h] ^ 1);
          ph_sh[h] ^= 1;
          tc_fence_after();
          const uint32_t d = tmem_base + h * 128;
          const uint64_t a_hi = umma_desc_sw128(sq, 1, 64) + 4 * h, a_lo = umma_desc_sw128(sq + 16384, 1, 64) + 4 * h;
          const uint64_t b_hi = umma_desc_sw128(sk, 1, 64) + 4 * h, b_lo = umma_desc_sw128(sk + 16384, 1, 64) + 4 * h;
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(d, a_hi + 2 * k, b_hi + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(d, a_hi + 2 * k, b_lo + 2 * k, idesc_s, 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(d, a_lo + 2 * k, b_hi + 2 * k, idesc_s, 1);
          umma_commit(&bars->s_full[h]);
        }
        umma_commit(&bars->empty[0]);  // Q, K slots free once these MMAs retire
        umma_commit(&bars->empty[1]);
      };
      if (my_items > 0) issue_s(0);
      for (int item = 0; item < my_items; ++item) {
        if (item + 1 < my_items) issue_s(item + 1);
        const int vs = 2 + (item & 1);
        mbar_wait(&bars->full[vs], (item >> 1) & 1);
        tc_fence_after();
        const uint32_t sv = s0 + vs * AT_BLK;
        for (int h = 0; h < 2; ++h) {
          if (!head_valid(item, h)) continue;
          mbar_wait(&bars->p_full[h], ph_oh[h]);
          mbar_wait(&bars->o_empty[h], ph_oh[h] ^ 1);
          ph_oh[h] ^= 1;
          tc_fence_after();
          const uint32_t sp = smem_u32(smem + AT_SMEM_P + h * AT_BLK);
          const uint64_t a_hi = umma_desc_sw128(sp, 1, 64), a_lo = umma_desc_sw128(sp + 16384, 1, 64);
#pragma unroll
          for (int win = 0; win < 2; ++win) {
            const uint32_t d = tmem_base + 256 + h * 128 + win * 64;
            // B: rows = keys of window `win` (K index), 64 channels per 128-byte row; 16 keys per K-step = 2048 B
            const uint64_t b_hi = umma_desc_sw128(sv + win * 8192, 512, 64), b_lo = umma_desc_sw128(sv + 16384 + win * 8192, 512, 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a_hi + 2 * k, b_hi + 128 * k, idesc_o, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a_hi + 2 * k, b_lo + 128 * k, idesc_o, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a_lo + 2 * k, b_hi + 128 * k, idesc_o, 1);
          }
          umma_commit(&bars->o_full[h]);
          umma_commit(&bars->p_empty[h]);
        }
        umma_commit(&bars->empty[vs]);
      }
    }
  } else {
    // ================================ softmax + epilogue ==================================
    // 16 warps: (head A / B) x (column half 0 / 1) x (TMEM lane quarter).  Two threads share a query row, each owns 32 of
    // its 64 scores; they merge (max, sum) once per item through shared memory (the flash-attention rescale), so a warp
    // scheduler has four softmax warps to interleave instead of two.
    const int cw = warp - 2;
    const int h = (cw >> 2) & 1;            // head A / B of the pair
    const int half = cw >> 3;               // which 32 of the row's 64 keys (and which 16 of its 32 output channels)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // row of the 128-row item = TMEM lane
    const int win = r >> 6, i = r & 63;     // window of the pair, token within the window
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int base_i = 15 * (i >> 3) + (i & 7) + 112 - 60 * half;  // keys j = 32 half + jj: 15 (j >> 3) + (j & 7) = 60 half + ...
    const float scale_l2 = gm.scale * 1.4426950408889634f;
    float2* xch = reinterpret_cast<float2*>(smem + AT_SMEM_XCH);   // [parity][head][row][half] (max, sum) of a half row
    uint32_t ph = 0;
    // epilogue state of the previous item
    bool pend = false;
    int pend_tok = 0, pend_head = 0;
    bool pend_valid = false, pend_last = false;
    auto epilogue = [&](uint32_t phase) {
      mbar_wait(&bars->o_full[h], phase);
      tc_fence_after();
      float o[16];
      tmem_ld_32x16(lane_addr + 256 + h * 128 + win * 64 + h * 32 + half * 16, o);
      tc_fence_before();
      mbar_arrive(&bars->o_empty[h]);
      if (!pend_valid) return;
      const int kbs_o = (gm.C + 63) / 64;
      const long long tk = pend_tok;
      const int c0 = half * 16;  // first of this thread's 16 channels within the head
      if (out) {
        float* dst = out + (size_t)tk * gm.C + pend_head * gm.D + c0;
#pragma unroll
        for (int c = 0; c < 16; c += 2)
          if (c0 + c < gm.D) *reinterpret_cast<float2*>(dst + c) = make_float2(o[c], o[c + 1]);
      }
      if (out_sti && gm.pad_out) {
        // head-padded image [tokens, G]: this thread's 16 channels are two whole 16-byte chunks of the token's row in
        // block `head / 2`; channel D of head 0 carries 1.0 (bias-gradient column of proj's wgrad), other padding is 0
        if (pend_head == 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c0 + c == gm.D) o[c] = 1.f;
        }
        uint8_t* rb = out_sti + ((size_t)((tk >> 7) * (gm.G / 64) + (pend_head >> 1)) << 15) + (size_t)(tk & 127) * 128;
        const int r7 = (int)(tk & 7);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 hi, lo;
          split_pair(o[8 * j], o[8 * j + 1], hi.x, lo.x);
          split_pair(o[8 * j + 2], o[8 * j + 3], hi.y, lo.y);
          split_pair(o[8 * j + 4], o[8 * j + 5], hi.z, lo.z);
          split_pair(o[8 * j + 6], o[8 * j + 7], hi.w, lo.w);
          const int off = ((h * 4 + half * 2 + j) ^ r7) << 4;
          *reinterpret_cast<uint4*>(rb + off) = hi;
          *reinterpret_cast<uint4*>(rb + 16384 + off) = lo;
        }
        if (pend_last && (gm.heads & 1)) {  // odd head count: the second half of the last block is padding
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int off = ((4 + half * 2 + j) ^ r7) << 4;
            *reinterpret_cast<uint4*>(rb + off) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(rb + 16384 + off) = make_uint4(0, 0, 0, 0);
          }
        }
      } else if (out_sti) {
        uint8_t* rb = out_sti + ((size_t)((tk >> 7) * kbs_o) << 15) + (size_t)(tk & 127) * 128;
        const int r7 = (int)(tk & 7);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          if (c0 + c < gm.D) {
            const int cidx = pend_head * gm.D + c0 + c, cc = cidx & 63;
            uint32_t hi, lo;
            split_pair(o[c], o[c + 1], hi, lo);
            uint8_t* dst = rb + ((size_t)(cidx >> 6) << 15) + ((((cc >> 3) ^ r7) << 4) + (cc & 7) * 2);
            *reinterpret_cast<uint32_t*>(dst) = hi;
            *reinterpret_cast<uint32_t*>(dst + 16384) = lo;
          }
        }
        if (pend_last && half == 1) {  // channel padding [C, kbs*64): 1.0 in channel C (bias-gradient column), zeros after
          for (int cidx = gm.C; cidx < kbs_o * 64; cidx += 2) {
            const int cc = cidx & 63;
            uint32_t hi, lo;
            split_pair(cidx == gm.C ? 1.f : 0.f, 0.f, hi, lo);
            uint8_t* dst = rb + ((size_t)(cidx >> 6) << 15) + ((((cc >> 3) ^ r7) << 4) + (cc & 7) * 2);
            *reinterpret_cast<uint32_t*>(dst) = hi;
            *reinterpret_cast<uint32_t*>(dst + 16384) = lo;
          }
        }
      }
    };
    for (int wp = blockIdx.x; wp < n_wp; wp += gridDim.x) {
      const int wi = wp * 2 + win;
      const bool row_valid = wi < gm.nwin;
      int tok = 0, rid = 0;
      if (row_valid) at_token_map(gm, wi, i, tok, rid);
      named_bar_sync(1, 512);  // every softmax thread is done with the previous pair's region ids
      if (h == 0 && half == 0) rid_s[r] = rid;
      named_bar_sync(1, 512);
      bool masked = false;  // does any of this thread's keys lie in another shift-mask region than its query?
      if (gm.use_mask && gm.shift > 0) {
        for (int j = 0; j < 32; ++j) masked |= rid_s[win * 64 + half * 32 + j] != rid;
      }
      for (int hp = 0; hp < n_hp; ++hp) {
        const int head = hp * 2 + h;
        if (head < gm.heads) {
          // ---- this thread's 32 scores of the item
          mbar_wait(&bars->s_full[h], ph);
          tc_fence_after();
          float s[32];
          tmem_ld_32x32(lane_addr + h * 128 + win * 64 + half * 32, s);
          tc_fence_before();
          mbar_arrive(&bars->s_empty[h]);
          const float* bt = bias_s + head * 225 + base_i;
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = fmaf(s[j], scale_l2, bt[-(15 * (j >> 3) + (j & 7))]);
          if (masked) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (rid_s[win * 64 + half * 32 + j] != rid) s[j] += -100.0f * 1.4426950408889634f;
          }
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: the adds / maxes are 4 cycles deep
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
          const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s[j] = ex2_approx(s[j] - m);
            l4[j & 3] += s[j];
          }
          const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
          // merge with the other half of the row: P = e * 2^(m - M) / (l 2^(m - M) + l' 2^(m' - M)), M = max(m, m')
          float2* xs = xch + (((ph & 1) * 2 + h) * 128 + r) * 2;
          xs[half] = make_float2(m, l);
          named_bar_sync(2 + h, 256);
          const float2 ot = xs[half ^ 1];
          const float M = fmaxf(m, ot.x);
          const float f = ex2_approx(m - M);
          const float inv = f / fmaf(l, f, ot.y * ex2_approx(ot.x - M));
          // ---- P -> the K-major A tile of P V (previous item's P V must have retired)
          mbar_wait(&bars->p_empty[h], ph ^ 1);
          uint8_t* prow = smem + AT_SMEM_P + h * AT_BLK + r * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 hi, lo;
            split_pair(s[8 * c] * inv, s[8 * c + 1] * inv, hi.x, lo.x);
            split_pair(s[8 * c + 2] * inv, s[8 * c + 3] * inv, hi.y, lo.y);
            split_pair(s[8 * c + 4] * inv, s[8 * c + 5] * inv, hi.z, lo.z);
            split_pair(s[8 * c + 6] * inv, s[8 * c + 7] * inv, hi.w, lo.w);
            const int off = ((half * 4 + c) ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + off) = hi;
            *reinterpret_cast<uint4*>(prow + 16384 + off) = lo;
          }
          fence_proxy_async_smem();
          mbar_arrive(&bars->p_full[h]);
          // ---- epilogue of the previous item (its P V ran while this softmax was computed)
          if (pend) epilogue(ph ^ 1);  // parity of the previous item of this head
          pend = true;
          pend_tok = tok;
          pend_head = head;
          pend_valid = row_valid;
          pend_last = head == gm.heads - 1;
          ph ^= 1;  // per head: toggles only on the items this head takes part in
        }
      }
    }
    if (pend) epilogue(ph ^ 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool window_attn_tc_supported(int c, int heads, int ws) {
  const int d = c / heads;
  return ws == 8 && d <= 32 && d % 2 == 0 && heads <= AT_MAX_HEADS && nsr_device_supports_tcgen05();
}

int window_attn_tc_fwd_launch(const void* qkv, const float* table, float* out, void* out_sti, int out_padded, int batch, int h,
                              int w, int c, int heads, int ws, int shift, int use_mask, float scale, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_FWD_SMEM);
    if (e != cudaSuccess) {
      set_error("window_attn_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  AtGeom g{batch, h, w, c, heads, ws, shift, use_mask, c / heads, h / ws, w / ws, (heads * 32 + 63) / 64 * 64,
           batch * (h / ws) * (w / ws), scale, out_padded};
  const int n_wp = (g.nwin + 1) / 2;
  const int grid = n_wp < kNumSMs ? n_wp : kNumSMs;
  window_attn_tc_fwd_kernel<<<grid, AT_THREADS, AT_FWD_SMEM, st>>>(reinterpret_cast<const uint8_t*>(qkv), table, out,
                                                                  reinterpret_cast<uint8_t*>(out_sti), g);
  NSR_CHECK_LAUNCH("window_attn_tc_fwd");
  return NSR_OK;
}

}  // namespace nsr

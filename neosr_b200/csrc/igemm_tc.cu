// tcgen05 (UMMA) implicit-GEMM engine: 3xBF16-split operands, fp32 accumulation in TMEM.
//
//   y[p, n] = epilogue( sum_{tap, c} x[p @ tap, c] * w[n, tap, c] )        (NHWC activations)
//
// fp32 parity needs more mantissa than one bf16/tf32 pass gives (SURVEY.md §7 hard part 1), so
// every operand is split x ~= hi + lo (two bf16) and each k-block issues three MMA groups
// hi*hi + hi*lo + lo*hi into the same TMEM accumulator (error ~2^-16 relative, fp32 accumulate).
//
// Persistent, warp-specialised CTA (1 per SM, 448 threads):
//   warps 0-7  A producers : gather the im2col tile (zero padding / channel tail by predication)
//                            straight from fp32 HBM, split to bf16 hi/lo in registers, store into
//                            the SWIZZLE_128B K-major smem image the tensor core reads;
//   warp  8    B loader    : weights are pre-split and pre-swizzled by nsr_pack_weight, so one
//                            bulk copy (cp.async.bulk -> UBLKCP) per operand half fills a stage;
//   warp  9    MMA issuer  : one elected thread issues tcgen05.mma (M=128, N=BN, K=16) and
//                            tcgen05.commit to recycle smem stages / publish the accumulator;
//   warps 10-13 epilogue   : tcgen05.ld TMEM -> registers, bias/activation/act-grad/row-scale/
//                            residual, fp32 stores.  TMEM is double-buffered (2 x BN columns) so
//                            the epilogue of tile i overlaps the main loop of tile i+1.
#include <cstdlib>
#include <cstring>
#include "tc_epilogue.cuh"
#include "wgrad_geom.cuh"

namespace nsr {
using namespace tc;

constexpr int TC_BM = 128;          // pixels per tile (UMMA M)
constexpr int TC_BK = 64;           // bf16 elements per k-block (one 128-byte swizzle row)
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_PROD_WARPS + 2 + 4) * 32;  // 448
constexpr int TC_A_BYTES = TC_BM * 128;                   // one half (hi or lo) of an A stage

// STI = the A operand arrives as a split tile image (bulk copy, no conversion warps); the eight
// producer warps then serve as extra epilogue warps (12 instead of 4).
template <int BN, bool STI>
struct TcCfg {
  static constexpr int b_bytes = BN * 128;  // one half of a B stage
  static constexpr int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
  static constexpr int epi_warps = STI ? 12 : 4;
  static constexpr int epi_bytes = epi_warps * 32 * EPI_LD * 4;  // per-warp transpose tiles
  static constexpr int budget = 227 * 1024 - 1024 - 256 - epi_bytes;
  static constexpr int stages = budget / stage_bytes > 4 ? 4 : budget / stage_bytes;
  static constexpr int smem_bytes = stages * stage_bytes + 1024 /* align slack */ + 256 /* barriers */ + epi_bytes;
};

struct TcGeom {
  int n_tiles, m_tiles, num_tiles, cblks, nk, n_pad64;
  int passes;  // 3: hi*hi + hi*lo + lo*hi;  1: hi*hi only (NSR_ENGINE_BF16: the lo halves are not even fetched)
  long long M;
};

template <int BN, bool STI, bool WIN = false>
__global__ void __launch_bounds__(TC_THREADS, 1) igemm_fprop_tc(NsrConv d, TcGeom g, const uint8_t* __restrict__ wimg) {
  using Cfg = TcCfg<BN, STI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::stages * Cfg::stage_bytes);
  uint64_t* full = bars;                      // [stages]  producers + B loader -> MMA
  uint64_t* empty = bars + Cfg::stages;       // [stages]  MMA -> producers / B loader
  uint64_t* tfull = bars + 2 * Cfg::stages;   // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;               // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::stages; ++s) {
      mbar_init(&full[s], STI ? 1 : TC_PROD_WARPS * 32 + 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], Cfg::epi_warps * 32);
    }
    fence_mbar_init();
  }
  if (warp == TC_PROD_WARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int hw = d.h * d.w;
  if (!STI && warp < TC_PROD_WARPS) {
    // ================================ A producers =========================================
    // Software pipelined over the flattened (tile, k-block) sequence: the loads of step i+1 are in
    // flight (8 x LDG.128 per thread = one whole A k-block per CTA) while step i is split to
    // bf16 hi/lo and stored, so HBM/L2 latency is never exposed between k-blocks or tiles.
    const int t = threadIdx.x;        // 0..255
    const int chunk = t & 7;          // 8-float chunk within the 64-channel k-block
    const int row0 = t >> 3;          // rows row0 + 32*i, i = 0..3
    int stage = 0;
    uint32_t phase = 0;
    // load-side cursor (tile, filter row, filter column, channel block): k-block kb = (r * kw + s) * cblks + cblk is
    // walked with counters, and everything that depends only on the tile row - the pixel's base address and which filter
    // rows / columns land inside the image - is computed once per tile, so a k-block costs one offset, four bit tests and
    // eight loads per thread (the gather used to spend more instructions on divisions, bounds and 64-bit addresses than on
    // the fp32 -> hi/lo conversion: profiles/r01_v_ncu_vgg_conv1_2.txt)
    int l_tile = blockIdx.x, l_kb = 0, l_r = 0, l_s = 0, l_cblk = 0;
    const float* rowp[4];
    uint32_t rmask[4], cmask[4];  // bit r / bit s: input row oh + r - pad / column ow + s - pad exists (kh, kw <= 32)
    const bool small_m = g.M < (1LL << 31);
    auto decode_rows = [&](int tile) {
      const long long m0 = (long long)(tile / g.n_tiles) * TC_BM;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long p = m0 + row0 + 32 * i;
        rmask[i] = cmask[i] = 0u;
        rowp[i] = d.x;
        if (p < g.M) {
          const int rem = small_m ? (int)((unsigned)p % (unsigned)hw) : (int)(p % hw);
          const int oh = rem / d.w, ow = rem - oh * d.w;
          for (int r = 0; r < d.kh; ++r) rmask[i] |= (unsigned)(oh + r - d.pad >= 0 && oh + r - d.pad < d.h) << r;
          for (int sx = 0; sx < d.kw; ++sx) cmask[i] |= (unsigned)(ow + sx - d.pad >= 0 && ow + sx - d.pad < d.w) << sx;
          rowp[i] = d.x + p * d.x_ld + chunk * 8;
        }
      }
    };
    auto load = [&](float4 (&f)[4][2]) {  // loads (l_tile, l_kb) and advances the cursor
      if (l_kb == 0) {
        decode_rows(l_tile);
        l_r = l_s = l_cblk = 0;
      }
      const int c = l_cblk * TC_BK + chunk * 8;
      const bool c0 = c < d.cin, c1 = c + 4 < d.cin;
      const long long off = ((long long)(l_r - d.pad) * d.w + (l_s - d.pad)) * d.x_ld + l_cblk * TC_BK;  // same for the four rows
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = ((rmask[i] >> l_r) & (cmask[i] >> l_s) & 1u) != 0;
        const float* src = rowp[i] + off;
        f[i][0] = (ok && c0) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[i][1] = (ok && c1) ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++l_cblk == g.cblks) {
        l_cblk = 0;
        if (++l_s == d.kw) { l_s = 0; ++l_r; }
      }
      if (++l_kb == g.nk) { l_kb = 0; l_tile += gridDim.x; }
    };
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) ++my_tiles;
    const int total = my_tiles * g.nk;
    float4 cur[4][2], nxt[4][2];
    if (total > 0) load(cur);
    for (int it = 0; it < total; ++it) {
      const bool more = it + 1 < total;
      if (more) load(nxt);
      mbar_wait<64>(&empty[stage], phase ^ 1);  // MMA-bound layers: eight waiting producer warps sleep, not spin
      uint8_t* a_hi = smem + stage * Cfg::stage_bytes;
      uint8_t* a_lo = a_hi + TC_A_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row0 + 32 * i;
        uint4 hi, lo;
        split8(cur[i][0], cur[i][1], hi, lo);
        const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[stage]);
      if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
      if (more) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { cur[i][0] = nxt[i][0]; cur[i][1] = nxt[i][1]; }
      }
    }
  } else if (warp == TC_PROD_WARPS) {
    // ================================ B loader ============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int nblk = g.n_pad64 / 64;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int n0 = (tile % g.n_tiles) * BN;
        int rows = g.n_pad64 - n0;
        if (rows > BN) rows = BN;
        const uint32_t bytes = (uint32_t)rows * 128;
        const size_t mt = (size_t)(tile / g.n_tiles);
        for (int kb = 0; kb < g.nk; ++kb) {
          mbar_wait<64>(&empty[stage], phase ^ 1);
          uint8_t* b_hi = smem + stage * Cfg::stage_bytes + 2 * TC_A_BYTES;
          uint8_t* b_lo = b_hi + Cfg::b_bytes;
          const uint8_t* src_hi = wimg + ((size_t)(kb * 2 + 0) * nblk + (n0 >> 6)) * 8192;
          const uint8_t* src_lo = wimg + ((size_t)(kb * 2 + 1) * nblk + (n0 >> 6)) * 8192;
          const uint32_t halves = g.passes == 3 ? 2u : 1u;
          mbar_arrive_expect_tx(&full[stage], halves * bytes + (STI ? halves * TC_A_BYTES : 0));
          if (STI)  // A tile = one 32 KiB block (hi image + lo image) of the split tile image
            bulk_g2s(smem + stage * Cfg::stage_bytes,
                     reinterpret_cast<const uint8_t*>(d.x_sti) + ((mt * g.nk + kb) << 15), halves * TC_A_BYTES, &full[stage]);
          bulk_g2s(b_hi, src_hi, bytes, &full[stage]);
          if (halves == 2) bulk_g2s(b_lo, src_lo, bytes, &full[stage]);
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_PROD_WARPS + 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        const uint32_t bphase = (local >> 1) & 1;
        mbar_wait<64>(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < g.nk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::stage_bytes);
          const uint64_t a_hi = umma_desc_sw128(sa, 1, 64);
          const uint64_t a_lo = umma_desc_sw128(sa + TC_A_BYTES, 1, 64);
          const uint64_t b_hi = umma_desc_sw128(sa + 2 * TC_A_BYTES, 1, 64);
          const uint64_t b_lo = umma_desc_sw128(sa + 2 * TC_A_BYTES + Cfg::b_bytes, 1, 64);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)  // +32 bytes (2 x 16 B units) per K=16 step
            umma_bf16(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
          if (g.passes == 3) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
          }
          umma_commit(&empty[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);      // accumulator complete
      }
    }
  } else if (warp >= TC_PROD_WARPS + 2 || (STI && warp < TC_PROD_WARPS)) {
    // ================================ epilogue ============================================
    // TMEM hands each thread one accumulator ROW (32 consecutive columns).  A per-warp smem
    // transpose turns that into row-contiguous global traffic: every warp instruction reads
    // (aux / residual) or writes (y, y_pre) 4 rows x 128 contiguous bytes.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int NSLOT = STI ? 3 : 1;                                   // warps sharing a lane quarter
    const int slot = STI ? (warp < TC_PROD_WARPS ? (warp >> 2) : 2) : 0;  // they interleave 32-col chunks
    float* stg = reinterpret_cast<float*>(smem + Cfg::stages * Cfg::stage_bytes + 256) + (slot * 4 + q) * (32 * EPI_LD);
    const int kbs_out = d.y_sti ? (d.cout + 63) / 64 : 0;
    const int ncols = kbs_out ? kbs_out * 64 : d.cout;
    int local = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      const long long p0 = (long long)(tile / g.n_tiles) * TC_BM + q * 32;   // first row of this warp
      const int n0 = (tile % g.n_tiles) * BN;
      int wrow_lane = 0;  // window-ordered STI output: where tile row p0 + lane goes (computed while the MMAs run)
      if (WIN && p0 + lane < g.M) wrow_lane = sti_win_row(d, p0 + lane, hw);
      // The rows the epilogue will add (residual) or multiply by (aux) were last touched several kernels ago: pull this
      // warp's pieces of them into L2 while the tile's MMAs run, so the loads inside the row loop see L2 latency, not
      // HBM's (profiles/r02_x_ncu_stifprop_proj.txt: 56 % of the proj kernel's stall samples were the first FADD on them)
      if (STI && (d.residual != nullptr || d.aux != nullptr) && p0 + lane < g.M) {
        const long long pr = p0 + lane;
        for (int c0 = slot * 32; c0 < BN; c0 += 32 * NSLOT) {
          const int n = n0 + c0;
          if (n >= d.cout) break;
          if (d.residual) {
            const char* a = reinterpret_cast<const char*>(d.residual + pr * (d.res_ld ? d.res_ld : d.y_ld) + n);
            prefetch_l2(a);
            prefetch_l2(a + 112);
          }
          if (d.aux) {
            const long long e = pr * (d.aux_ld ? d.aux_ld : d.y_ld) + n;
            const char* a = d.aux_mode == 2 ? reinterpret_cast<const char*>(reinterpret_cast<const uint16_t*>(d.aux) + e)
                                            : reinterpret_cast<const char*>(d.aux + e);
            prefetch_l2(a);
            if (d.aux_mode != 2) prefetch_l2(a + 112);
          }
        }
      }
      mbar_wait<STI ? 32 : 128>(&tfull[buf], bphase);  // idle epilogue warps must not spin against the producers
      tc_fence_after();
#pragma unroll 1
      for (int c0 = slot * 32; c0 < BN; c0 += 32 * NSLOT) {
        if (n0 + c0 >= ncols) break;  // warp-uniform
        epi_chunk<WIN>(d, stg, tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c0, p0, n0 + c0, g.M, hw, lane, kbs_out, wrow_lane);
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_PROD_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pick_bn(int cout) {
  // smallest tile count first, then least padding waste
  int best = 64, best_cost = 1 << 30;
  const int cands[4] = {64, 128, 192, 256};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int tiles = (cout + bn - 1) / bn;
    const int cost = tiles * bn;  // total padded N; ties -> bigger tile (fewer A re-reads)
    if (cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
  }
  return best;
}

bool conv_fprop_tc_supported(const NsrConv& d) {
  static int ok_dev = -1;
  if (ok_dev < 0) ok_dev = nsr_device_supports_tcgen05();
  if (!ok_dev) return false;
  if (d.cin % 4 || d.x_ld % 4 || d.cout % 4 || d.y_ld % 4 || d.res_ld % 4 || d.aux_ld % 4) return false;
  if (d.cin < 16 || d.cout < 16) return false;  // image-side 3-channel convs stay on the SIMT engine
  if (d.kh > 32 || d.kw > 32) return false;     // the gather keeps per-row validity masks of the filter rows / columns
  if (d.x_sti != nullptr) {
    if (d.kh != 1 || d.kw != 1) return false;  // tile images carry no halo: 1x1 contractions only
    if (!aligned16(d.x_sti)) return false;
  } else if (d.x == nullptr) {
    return false;
  }
  if (d.y == nullptr && d.y_sti == nullptr) return false;
  if (d.sti_win) {  // window-ordered STI output: whole windows only
    const int ws = d.sti_win & 0xFFFF, shift = d.sti_win >> 16;
    if (d.y_sti == nullptr || d.x_sti == nullptr || ws <= 0 || d.h % ws || d.w % ws || shift < 0 || shift >= ws) return false;
  }
  if (d.act != NSR_ACT_NONE && d.actgrad != NSR_ACT_NONE) return false;  // epilogue is specialised on one of them
  if (d.pre_mode == 2 || d.aux_mode == 2) {  // 16-bit activation-gradient codes: exactly the two MLP hand-over epilogues
    const bool plain = d.y_sti != nullptr && d.y == nullptr && d.residual == nullptr && d.row_scale == nullptr;
    const bool fwd = d.pre_mode == 2 && d.aux_mode == 0 && d.act == NSR_ACT_GELU && d.actgrad == NSR_ACT_NONE && d.y_pre != nullptr;
    const bool bwd = d.aux_mode == 2 && d.pre_mode == 0 && d.act == NSR_ACT_NONE && d.actgrad == NSR_ACT_MULAUX && d.y_pre == nullptr;
    if (!plain || !(fwd || bwd)) return false;
  }
  if (!aligned16(d.x) || !aligned16(d.y) || !aligned16(d.y_sti) || !aligned16(d.bias) || !aligned16(d.aux) || !aligned16(d.residual) ||
      !aligned16(d.y_pre) || !aligned16(d.prelu) || !aligned16(d.w_packed))
    return false;
  return true;
}

template <int BN, bool STI, bool WIN = false>
static int launch_fprop_tc(const NsrConv& d, cudaStream_t st) {
  using Cfg = TcCfg<BN, STI>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(igemm_fprop_tc<BN, STI, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes);
    if (e != cudaSuccess) {
      set_error("igemm_fprop_tc<%d>: cudaFuncSetAttribute(%d): %s", BN, Cfg::smem_bytes, cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  // which packed flavour this is does not matter: the GEMM view is W[n = d.cout][tap][c = d.cin]
  PackedGeom pg = packed_geom(d.cout, d.cin, d.kh, d.kw, 0);
  TcGeom g;
  g.M = (long long)d.batch * d.h * d.w;
  g.m_tiles = ceil_div(g.M, TC_BM);
  g.n_tiles = ceil_div(d.cout, BN);
  g.num_tiles = g.m_tiles * g.n_tiles;
  g.cblks = pg.cblks;
  g.nk = pg.taps * pg.cblks;
  g.n_pad64 = pg.n_pad64;
  g.passes = mma_passes(d.engine);
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(d.w_packed) + pg.f32_bytes;
  const int grid = g.num_tiles < kNumSMs ? g.num_tiles : kNumSMs;
  igemm_fprop_tc<BN, STI, WIN><<<grid, TC_THREADS, Cfg::smem_bytes, st>>>(d, g, wimg);
  NSR_CHECK_LAUNCH("igemm_fprop_tc");
  return NSR_OK;
}

int conv_fprop_tc(const NsrConv& d, cudaStream_t st) {
  if (d.x_sti != nullptr) {
    int bn = pick_bn(d.cout);
    if (bn == 256) bn = 128;  // the 12-warp epilogue staging leaves no room for BN=256 stages
    if (d.sti_win) {  // window-ordered STI output (qkv fprop / proj dgrad feeding nsr_window_attn_wsti_*)
      switch (bn) {
        case 64: return launch_fprop_tc<64, true, true>(d, st);
        case 128: return launch_fprop_tc<128, true, true>(d, st);
        default: return launch_fprop_tc<192, true, true>(d, st);
      }
    }
    switch (bn) {
      case 64: return launch_fprop_tc<64, true>(d, st);
      case 128: return launch_fprop_tc<128, true>(d, st);
      default: return launch_fprop_tc<192, true>(d, st);
    }
  }
  switch (pick_bn(d.cout)) {
    case 64: return launch_fprop_tc<64, false>(d, st);
    case 128: return launch_fprop_tc<128, false>(d, st);
    case 192: return launch_fprop_tc<192, false>(d, st);
    default: return launch_fprop_tc<256, false>(d, st);
  }
}

// ==========================================================================================
// wgrad:  dw[co, tap, ci] = sum_p dy[p, co] * x[p @ tap, ci]
//
// The reduction runs over pixels, so both UMMA operands are "MN-major": a smem row is one pixel
// (the K index) holding 64 consecutive channels (128 B, SWIZZLE_128B) — byte-for-byte the tile
// the fprop producers build, only the descriptor says MN-major.  P (rows of D, 128 channels =
// two 64-channel panels) and Q (columns of D, BN channels) are dy and x in whichever orientation
// wastes less padding.  Pixels are split across CTAs (deterministic split-K: per-split partials
// go to the workspace, wgrad_reduce sums them in a fixed order).
constexpr int WG_KPIX = 64;  // pixels per k-block
constexpr int WG_MAX_ROWS_PER_SPLIT = 4096;

template <int BN>
struct WgCfg {
  static constexpr int p_bytes = 2 * WG_KPIX * 128;       // one half (hi or lo): 2 panels
  static constexpr int q_bytes = (BN / 64) * WG_KPIX * 128;
  static constexpr int stage_bytes = 2 * p_bytes + 2 * q_bytes;
  static constexpr int stages = (200 * 1024) / stage_bytes > 4 ? 4 : (200 * 1024) / stage_bytes;
  static constexpr int smem_bytes = stages * stage_bytes + 1024 + 256;
};

int launch_wgrad_reduce(const float* partial, float* dw, int splitk, int cout, int taps, int cin, cudaStream_t st);
int conv_bias_grad(const NsrWgrad& d, float* bias_partial, int bias_blocks, cudaStream_t st);

template <int BN, bool STI>
__global__ void __launch_bounds__(TC_THREADS, 1) igemm_wgrad_tc(NsrWgrad d, WgGeom g, float* __restrict__ partial) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::stages * Cfg::stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::stages;
  uint64_t* tfull = bars + 2 * Cfg::stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::stages; ++s) {
      mbar_init(&full[s], STI ? 1 : TC_PROD_WARPS * 32);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == TC_PROD_WARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int hw = d.h * d.w;

  // item -> (m_tile, n_tile, tap, split)
  auto decode = [&](int item, int& mt, int& nt, int& tap, int& split) {
    split = item % g.splitk;
    item /= g.splitk;
    tap = item % g.taps;
    item /= g.taps;
    nt = item % g.n_tiles;
    mt = item / g.n_tiles;
  };

  if (STI && warp == TC_PROD_WARPS) {
    // ================================ bulk loader (operands are split tile images) ==========
    // A k-block = 64 pixels = rows [r0, r0+64) of the 128-row blocks: 8 KiB per 64-channel panel
    // and half, contiguous in the image, so each panel half is one bulk copy.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int kbp = (g.pc + 63) / 64, kbq = (g.qc + 63) / 64;
      const uint8_t* psti = reinterpret_cast<const uint8_t*>(g.p_sti);
      const uint8_t* qsti = reinterpret_cast<const uint8_t*>(g.q_sti);
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x) {
        int mt, nt, tap, split;
        decode(item, mt, nt, tap, split);
        const long long p_begin = (long long)split * g.rows_per_split;
        long long p_end = p_begin + g.rows_per_split;
        if (p_end > g.M) p_end = g.M;
        const int nkb = (int)((p_end - p_begin + WG_KPIX - 1) / WG_KPIX);
        int np = kbp - mt * 2;
        np = np > 2 ? 2 : np;
        int nq = kbq - nt * (BN / 64);
        nq = nq > BN / 64 ? BN / 64 : nq;
        for (int kb = 0; kb < nkb; ++kb) {
          const long long pk = p_begin + (long long)kb * WG_KPIX;
          const size_t pm = (size_t)(pk >> 7);
          const uint32_t roff = (uint32_t)(pk & 127) * 128;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * Cfg::stage_bytes;
          const bool lo = g.passes == 3;
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(np + nq) * (lo ? 2 : 1) * (WG_KPIX * 128));
          for (int j = 0; j < np; ++j) {
            const uint8_t* src = psti + ((pm * kbp + (size_t)(mt * 2 + j)) << 15) + roff;
            bulk_g2s(sb + j * (WG_KPIX * 128), src, WG_KPIX * 128, &full[stage]);
            if (lo) bulk_g2s(sb + Cfg::p_bytes + j * (WG_KPIX * 128), src + 16384, WG_KPIX * 128, &full[stage]);
          }
          for (int j = 0; j < nq; ++j) {
            const uint8_t* src = qsti + ((pm * kbq + (size_t)(nt * (BN / 64) + j)) << 15) + roff;
            bulk_g2s(sb + 2 * Cfg::p_bytes + j * (WG_KPIX * 128), src, WG_KPIX * 128, &full[stage]);
            if (lo) bulk_g2s(sb + 2 * Cfg::p_bytes + Cfg::q_bytes + j * (WG_KPIX * 128), src + 16384, WG_KPIX * 128, &full[stage]);
          }
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (!STI && warp < TC_PROD_WARPS) {
    // ================================ producers ===========================================
    // Each k-block needs (128 + BN) / 32 (row, 8-channel chunk) items per thread; they are
    // processed in groups of G with the loads of the next group (possibly of the next k-block or
    // work item) in flight while the current group is split and stored.
    constexpr int P_ITEMS = WG_KPIX * 16;                // (row, chunk) pairs in the P tile
    constexpr int ITEMS = WG_KPIX * (128 + BN) / 8;      // P then Q
    constexpr int PER_THREAD = ITEMS / (TC_PROD_WARPS * 32);
    static_assert(ITEMS % (TC_PROD_WARPS * 32) == 0, "tile items must divide evenly");
    constexpr int G = (PER_THREAD % 5 == 0) ? 5 : ((PER_THREAD % 4 == 0) ? 4 : 3);
    constexpr int GROUPS = PER_THREAD / G;
    static_assert(PER_THREAD % G == 0, "group size must divide the per-thread item count");
    const int t = threadIdx.x;
    int stage = 0;
    uint32_t phase = 0;

    // load-side cursor over (work item, k-block, group)
    int l_item = blockIdx.x, l_kb = 0, l_grp = 0, l_nkb = 0;
    int l_mt = 0, l_nt = 0, l_tap = 0, l_split = 0, l_dh = 0, l_dw = 0;
    long long l_pbegin = 0, l_pend = 0;
    auto begin_item = [&]() {
      decode(l_item, l_mt, l_nt, l_tap, l_split);
      const int r = l_tap / d.kw, sx = l_tap - r * d.kw;
      l_dh = r - d.pad;
      l_dw = sx - d.pad;
      l_pbegin = (long long)l_split * g.rows_per_split;
      l_pend = l_pbegin + g.rows_per_split;
      if (l_pend > g.M) l_pend = g.M;
      l_nkb = (int)((l_pend - l_pbegin + WG_KPIX - 1) / WG_KPIX);
    };
    auto load = [&](float4 (&f)[G][2]) {
      if (l_kb == 0 && l_grp == 0) begin_item();
      const long long pk = l_pbegin + (long long)l_kb * WG_KPIX;
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int idx = t + (l_grp * G + u) * (TC_PROD_WARPS * 32);
        const bool isP = idx < P_ITEMS;
        const int li = isP ? idx : idx - P_ITEMS;
        const int chunks = isP ? 16 : BN / 8;
        const int row = li / chunks, chunk = li - row * chunks;   // row = pixel within the k-block
        const int c0 = (isP ? l_mt * 128 : l_nt * BN) + chunk * 8;
        const int cmax = isP ? g.pc : g.qc;
        const int ld = isP ? g.p_ld : g.q_ld;
        const float* base = isP ? g.p_ptr : g.q_ptr;
        const bool shifted = (isP == (g.swap != 0));              // the x operand carries the tap shift
        const int sh = shifted ? l_dh : 0, sw = shifted ? l_dw : 0;
        const long long p = pk + row;
        bool ok = p < l_pend;
        long long sp = p;
        if (ok && (sh | sw)) {
          const long long b = p / hw;
          const int rem = (int)(p - b * hw);
          const int oh = rem / d.w, ow = rem - oh * d.w;
          const int ih = oh + sh, iw = ow + sw;
          ok = ih >= 0 && ih < d.h && iw >= 0 && iw < d.w;
          sp = p + (long long)sh * d.w + sw;
        }
        const float* src = base + sp * ld + c0;
        f[u][0] = (ok && c0 < cmax) ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[u][1] = (ok && c0 + 4 < cmax) ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++l_grp == GROUPS) {
        l_grp = 0;
        if (++l_kb == l_nkb) { l_kb = 0; l_item += gridDim.x; }
      }
    };
    // total number of groups this CTA will process
    long long total = 0;
    for (int item = blockIdx.x; item < g.num_items; item += gridDim.x) {
      int mt, nt, tap, split;
      decode(item, mt, nt, tap, split);
      const long long pb = (long long)split * g.rows_per_split;
      long long pe = pb + g.rows_per_split;
      if (pe > g.M) pe = g.M;
      total += (long long)((pe - pb + WG_KPIX - 1) / WG_KPIX) * GROUPS;
    }
    float4 cur[G][2], nxt[G][2];
    if (total > 0) load(cur);
    int s_grp = 0;
    for (long long it = 0; it < total; ++it) {
      const bool more = it + 1 < total;
      if (more) load(nxt);
      if (s_grp == 0) mbar_wait<64>(&empty[stage], phase ^ 1);
      uint8_t* st_base = smem + stage * Cfg::stage_bytes;
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int idx = t + (s_grp * G + u) * (TC_PROD_WARPS * 32);
        const bool isP = idx < P_ITEMS;
        const int li = isP ? idx : idx - P_ITEMS;
        const int chunks = isP ? 16 : BN / 8;
        const int row = li / chunks, chunk = li - row * chunks;
        // panel = 64-channel group; inside a panel: row * 128 B, 16-byte chunk swizzled by row
        const int panel = chunk >> 3, cc = chunk & 7;
        const int off = panel * (WG_KPIX * 128) + row * 128 + ((cc ^ (row & 7)) << 4);
        uint8_t* dst_hi = st_base + (isP ? 0 : 2 * Cfg::p_bytes);
        uint8_t* dst_lo = dst_hi + (isP ? Cfg::p_bytes : Cfg::q_bytes);
        uint4 hi, lo;
        split8(cur[u][0], cur[u][1], hi, lo);
        *reinterpret_cast<uint4*>(dst_hi + off) = hi;
        *reinterpret_cast<uint4*>(dst_lo + off) = lo;
      }
      if (++s_grp == GROUPS) {
        s_grp = 0;
        fence_proxy_async_smem();
        mbar_arrive(&full[stage]);
        if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
      }
      if (more) {
#pragma unroll
        for (int u = 0; u < G; ++u) { cur[u][0] = nxt[u][0]; cur[u][1] = nxt[u][1]; }
      }
    }
  } else if (warp == TC_PROD_WARPS + 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, 1, 1);  // both operands MN-major
      constexpr uint32_t LBO = (WG_KPIX * 128) >> 4;         // next 64-channel panel
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
        int mt, nt, tap, split;
        decode(item, mt, nt, tap, split);
        const long long p_begin = (long long)split * g.rows_per_split;
        long long p_end = p_begin + g.rows_per_split;
        if (p_end > g.M) p_end = g.M;
        const int nkb = (int)((p_end - p_begin + WG_KPIX - 1) / WG_KPIX);
        const int buf = local & 1;
        const uint32_t bphase = (local >> 1) & 1;
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::stage_bytes);
          const uint64_t p_hi = umma_desc_sw128(sa, LBO, 64);
          const uint64_t p_lo = umma_desc_sw128(sa + Cfg::p_bytes, LBO, 64);
          const uint64_t q_hi = umma_desc_sw128(sa + 2 * Cfg::p_bytes, LBO, 64);
          const uint64_t q_lo = umma_desc_sw128(sa + 2 * Cfg::p_bytes + Cfg::q_bytes, LBO, 64);
          // K = 16 pixels per MMA = two 8-row swizzle atoms = 2048 B = 128 x 16 B units
#pragma unroll
          for (int k = 0; k < WG_KPIX / 16; ++k) umma_bf16(tmem_d, p_hi + 128 * k, q_hi + 128 * k, idesc, (kb | k) != 0);
          if (g.passes == 3) {
#pragma unroll
            for (int k = 0; k < WG_KPIX / 16; ++k) umma_bf16(tmem_d, p_hi + 128 * k, q_lo + 128 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < WG_KPIX / 16; ++k) umma_bf16(tmem_d, p_lo + 128 * k, q_hi + 128 * k, idesc, 1);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else if (warp >= TC_PROD_WARPS + 2) {
    // ================================ epilogue: TMEM -> split-K partial ====================
    const int q = warp & 3;
    int local = 0;
    const size_t per_split = (size_t)d.cout * g.taps * d.cin;
    for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
      int mt, nt, tap, split;
      decode(item, mt, nt, tap, split);
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      mbar_wait<128>(&tfull[buf], bphase);  // long K loops: sleep instead of spinning
      tc_fence_after();
      float* out = partial + (size_t)split * per_split;
      const int m = mt * 128 + q * 32 + lane;  // P-channel of this thread's accumulator row
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (nt * BN + c0 >= g.qc) break;
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c0, v);
        if (m < g.pc) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nt * BN + c0 + j;
            if (n < g.qc) {
              const int co = g.swap ? n : m, ci = g.swap ? m : n;
              out[((size_t)co * g.taps + tap) * d.cin + ci] = v[j];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_PROD_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct WgPlan {
  WgGeom g;
  int bn;
  int bias_blocks;
  size_t dw_partial_floats, bias_partial_floats;
};

static int wg_pick_bn(int qc, int& cost) {
  int best = 64;
  cost = 1 << 30;
  const int cands[4] = {64, 128, 192, 256};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int c = ((qc + bn - 1) / bn) * bn;
    if (c < cost || (c == cost && bn > best)) { best = bn; cost = c; }
  }
  return best;
}

static bool wg_is_sti(const NsrWgrad& d) { return d.x_sti != nullptr && d.dy_sti != nullptr && d.kh == 1 && d.kw == 1; }
// process-wide tuning switches, read once (A/B measurements; the defaults are the measured best)
static int wg_sti_knob(const char* name, int dflt) {
  static const char* names[3] = {"NSR_WG_STI2", "NSR_WG_KPIX", "NSR_WG_PAIR"};
  static int cached[3] = {-1, -1, -1};
  for (int i = 0; i < 3; ++i)
    if (name == names[i] || strcmp(name, names[i]) == 0) {
      if (cached[i] < 0) {
        const char* v = getenv(name);
        cached[i] = v && *v ? atoi(v) : dflt;
      }
      return cached[i];
    }
  return dflt;
}

static WgPlan wg_plan(const NsrWgrad& d) {
  WgPlan p;
  WgGeom& g = p.g;
  g.M = (long long)d.batch * d.h * d.w;
  g.taps = d.kh * d.kw;
  // orientation: rows (P, tiles of 128) x cols (Q, tiles of BN); minimise padded area
  int cost_a, cost_b;
  const int bn_a = wg_pick_bn(d.cin, cost_a);   // P = dy (cout), Q = x (cin)
  const int bn_b = wg_pick_bn(d.cout, cost_b);  // P = x (cin),  Q = dy (cout)
  const long long area_a = (long long)((d.cout + 127) / 128 * 128) * cost_a;
  const long long area_b = (long long)((d.cin + 127) / 128 * 128) * cost_b;
  g.swap = area_b < area_a ? 1 : 0;
  p.bn = g.swap ? bn_b : bn_a;
  g.pc = g.swap ? d.cin : d.cout;
  g.qc = g.swap ? d.cout : d.cin;
  g.p_ld = g.swap ? d.x_ld : d.dy_ld;
  g.q_ld = g.swap ? d.dy_ld : d.x_ld;
  g.p_ptr = g.swap ? d.x : d.dy;
  g.q_ptr = g.swap ? d.dy : d.x;
  g.p_sti = g.swap ? d.x_sti : d.dy_sti;
  g.q_sti = g.swap ? d.dy_sti : d.x_sti;
  g.m_tiles = (g.pc + 127) / 128;
  g.n_tiles = (g.qc + p.bn - 1) / p.bn;
  // split-tile-image operands run on igemm_wgrad_sti (short k-blocks, deep ring, optionally two row tiles per item)
  g.mt = 1;
  g.kpix = WG_KPIX;
  if (wg_is_sti(d) && wg_sti_knob("NSR_WG_STI2", 1)) {
    g.kpix = wg_sti_knob("NSR_WG_KPIX", 32) == 64 ? 64 : 32;
    // two row tiles per item pay once P is wide (qkv: 576 rows, 122 -> 110 us); for 2 - 3 row tiles the longer items
    // and the doubled partials cost more than the saved Q traffic (measured, tools/bench_wgrad.py).  NSR_WG_PAIR: 0 never, 2 always
    const int pair = wg_sti_knob("NSR_WG_PAIR", 1);
    g.mt = (pair == 2 && g.m_tiles >= 2) || (pair == 1 && g.m_tiles >= 4) ? 2 : 1;
    if (g.mt == 2 && g.kpix == 64 && p.bn == 256) g.kpix = 32;  // that stage would not fit twice
  }
  g.passes = mma_passes(d.engine);
  g.m_groups = (g.m_tiles + g.mt - 1) / g.mt;
  const int tiles = g.m_groups * g.n_tiles * g.taps;
  // Pixels per split: at most WG_MAX_ROWS_PER_SPLIT (the tensor-core accumulator truncates on every
  // add, so its error grows with the number of sequential accumulations; the fixed-order fp32 reduce
  // over split partials rounds to nearest), and such that tiles * splits fills whole waves of 148 CTAs.
  const int min_splits = (int)((g.M + WG_MAX_ROWS_PER_SPLIT - 1) / WG_MAX_ROWS_PER_SPLIT);
  const int waves = (tiles * min_splits + kNumSMs - 1) / kNumSMs;
  int want = (waves * kNumSMs) / tiles;
  if (want < min_splits) want = min_splits;
  const int maxs = (int)((g.M + 255) / 256);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  long long rps = (g.M + want - 1) / want;
  g.rows_per_split = (rps + WG_KPIX - 1) / WG_KPIX * WG_KPIX;
  g.splitk = (int)((g.M + g.rows_per_split - 1) / g.rows_per_split);
  g.num_items = tiles * g.splitk;
  p.dw_partial_floats = (size_t)g.splitk * d.cout * g.taps * d.cin;
  p.bias_blocks = bias_grad_blocks(g.M);
  p.bias_partial_floats = (size_t)p.bias_blocks * d.cout;
  return p;
}

bool conv_wgrad_tc_supported(const NsrWgrad& d) {
  static int ok_dev = -1;
  if (ok_dev < 0) ok_dev = nsr_device_supports_tcgen05();
  if (!ok_dev) return false;
  if (d.cin % 4 || d.cout % 4 || d.x_ld % 4 || d.dy_ld % 4) return false;
  if (d.cin < 16 || d.cout < 16) return false;
  const bool sti = d.x_sti != nullptr && d.dy_sti != nullptr && d.kh == 1 && d.kw == 1;
  if (!sti && (d.x == nullptr || d.dy == nullptr)) return false;
  if (!aligned16(d.x) || !aligned16(d.dy) || !aligned16(d.x_sti) || !aligned16(d.dy_sti)) return false;
  return true;
}
size_t conv_wgrad_workspace_tc(const NsrWgrad& d) {
  WgPlan p = wg_plan(d);
  return (p.dw_partial_floats + p.bias_partial_floats) * sizeof(float);
}

template <int BN, bool STI>
static int launch_wgrad_tc(const NsrWgrad& d, const WgPlan& p, float* partial, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_tc<BN, STI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes);
    if (e != cudaSuccess) {
      set_error("igemm_wgrad_tc<%d>: cudaFuncSetAttribute: %s", BN, cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int grid = p.g.num_items < kNumSMs ? p.g.num_items : kNumSMs;
  igemm_wgrad_tc<BN, STI><<<grid, TC_THREADS, Cfg::smem_bytes, st>>>(d, p.g, partial);
  NSR_CHECK_LAUNCH("igemm_wgrad_tc");
  return NSR_OK;
}

static int launch_wgrad_sti_any(const NsrWgrad& d, const WgPlan& p, float* partial, cudaStream_t st) {
  if (wg_sti_knob("NSR_WG_STI2", 1)) return launch_wgrad_sti(d, p.g, p.bn, partial, st);
  switch (p.bn) {
    case 64: return launch_wgrad_tc<64, true>(d, p, partial, st);
    case 128: return launch_wgrad_tc<128, true>(d, p, partial, st);
    case 192: return launch_wgrad_tc<192, true>(d, p, partial, st);
    default: return launch_wgrad_tc<256, true>(d, p, partial, st);
  }
}

// split-K partials only (the caller reduces them later: nsr_wgrad_finalize_multi); returns the number of splits
int conv_wgrad_tc_partial(const NsrWgrad& d, int* splitk, cudaStream_t st) {
  WgPlan p = wg_plan(d);
  const size_t need = p.dw_partial_floats * sizeof(float);
  if (d.workspace_bytes < need || d.workspace == nullptr) {
    set_error("nsr_conv_wgrad_partial: workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(d.workspace);
  *splitk = p.g.splitk;
  return launch_wgrad_sti_any(d, p, partial, st);
}
size_t conv_wgrad_tc_partial_bytes(const NsrWgrad& d) { return wg_plan(d).dw_partial_floats * sizeof(float); }

int conv_wgrad_tc(const NsrWgrad& d, cudaStream_t st) {
  WgPlan p = wg_plan(d);
  const size_t need = (p.dw_partial_floats + p.bias_partial_floats) * sizeof(float);
  if (d.workspace_bytes < need || d.workspace == nullptr) {
    set_error("nsr_conv_wgrad(tcgen05): workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  float* partial = reinterpret_cast<float*>(d.workspace);
  int rc;
  const bool sti = d.x_sti != nullptr && d.dy_sti != nullptr && d.kh == 1 && d.kw == 1;
  if (sti) {
    rc = launch_wgrad_sti_any(d, p, partial, st);
  } else {
    switch (p.bn) {
      case 64: rc = launch_wgrad_tc<64, false>(d, p, partial, st); break;
      case 128: rc = launch_wgrad_tc<128, false>(d, p, partial, st); break;
      case 192: rc = launch_wgrad_tc<192, false>(d, p, partial, st); break;
      default: rc = launch_wgrad_tc<256, false>(d, p, partial, st); break;
    }
  }
  if (rc) return rc;
  rc = launch_wgrad_reduce(partial, d.dw, p.g.splitk, d.cout, p.g.taps, d.cin, st);
  if (rc) return rc;
  if (d.dbias) return conv_bias_grad(d, partial + p.dw_partial_floats, p.bias_blocks, st);
  return NSR_OK;
}

}  // namespace nsr

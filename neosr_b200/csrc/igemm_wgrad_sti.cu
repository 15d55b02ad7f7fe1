// Weight gradients of the 1x1 contractions from split tile images (both operands bulk-copied, no conversion warps):
//   partial[split][co][ci] = sum_{p in split} dy[p, co] * x[p, ci]
// Same math as igemm_wgrad_tc<BN, true> (igemm_tc.cu); what differs is the pipeline.  That kernel moved one 80 KB k-block
// (64 pixels) per stage through a 2-stage ring and measured 2 - 3 us per k-block against 1 us of MMA time: with a single
// stage in flight while the other is being consumed, a CTA never has more than ~80 KB outstanding, about half of what the
// HBM latency-bandwidth product asks of each SM.  Here
//   * a k-block is KPIX = 32 pixels (4 swizzle atoms), so the same shared memory holds 4 - 5 stages and all but one of
//     them are in flight;
//   * an item may cover MT = 2 row tiles of P (256 output rows, two accumulators in TMEM): Q is fetched once for both,
//     and the per-stage MMA work doubles, which hides the issue / commit latency of the short k-blocks;
//   * 192 threads: loader warp, MMA warp, four epilogue warps.  Items are long (>= 14 k-blocks) and CTAs get one or two
//     of them, so the accumulator is double-buffered only when that is free (2 * MT * BN <= 512 columns).
#include "tc_common.cuh"
#include "wgrad_geom.cuh"

namespace nsr {
using namespace tc;

constexpr int WS_THREADS = 6 * 32;

template <int BN, int KPIX, int MT>
struct WsCfg {
  static constexpr int panel = KPIX * 128;                 // one 64-channel panel, one half (hi or lo)
  static constexpr int p_bytes = 2 * MT * panel;           // per half
  static constexpr int q_bytes = (BN / 64) * panel;
  static constexpr int stage_bytes = 2 * p_bytes + 2 * q_bytes;
  static constexpr int budget = 227 * 1024 - 1024 - 256;
  static constexpr int stages = budget / stage_bytes > 8 ? 8 : budget / stage_bytes;
  static constexpr int smem_bytes = stages * stage_bytes + 1024 + 256;
  static constexpr int nbuf = 2 * MT * BN <= 512 ? 2 : 1;  // TMEM accumulator buffers
};

template <int BN, int KPIX, int MT>
__global__ void __launch_bounds__(WS_THREADS, 1) igemm_wgrad_sti(NsrWgrad d, WgGeom g, float* __restrict__ partial) {
  using Cfg = WsCfg<BN, KPIX, MT>;
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
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (m_group, n_tile, split); splits vary fastest, so the CTAs that share a Q range run side by side
  auto decode = [&](int item, int& mg, int& nt, int& split, long long& p_begin, int& nkb) {
    split = item % g.splitk;
    item /= g.splitk;
    nt = item % g.n_tiles;
    mg = item / g.n_tiles;
    p_begin = (long long)split * g.rows_per_split;
    long long p_end = p_begin + g.rows_per_split;
    if (p_end > g.M) p_end = g.M;
    nkb = (int)((p_end - p_begin + KPIX - 1) / KPIX);
  };

  if (warp == 0) {
    // ================================ bulk loader ==========================================
    // k-block = rows [r0, r0 + KPIX) of a 128-row block of the image: KPIX * 128 contiguous bytes per panel and half
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int kbp = (g.pc + 63) / 64, kbq = (g.qc + 63) / 64;
      const uint8_t* psti = reinterpret_cast<const uint8_t*>(g.p_sti);
      const uint8_t* qsti = reinterpret_cast<const uint8_t*>(g.q_sti);
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x) {
        int mg, nt, split, nkb;
        long long p_begin;
        decode(item, mg, nt, split, p_begin, nkb);
        int np = kbp - mg * 2 * MT;
        np = np > 2 * MT ? 2 * MT : np;
        int nq = kbq - nt * (BN / 64);
        nq = nq > BN / 64 ? BN / 64 : nq;
        const bool lo = g.passes == 3;
        const uint32_t tx = (uint32_t)(np + nq) * (lo ? 2 : 1) * Cfg::panel;
        for (int kb = 0; kb < nkb; ++kb) {
          const long long pk = p_begin + (long long)kb * KPIX;
          const size_t pm = (size_t)(pk >> 7);
          const uint32_t roff = (uint32_t)(pk & 127) * 128;
          mbar_wait<32>(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * Cfg::stage_bytes;
          mbar_arrive_expect_tx(&full[stage], tx);
          for (int j = 0; j < nq; ++j) {
            const uint8_t* src = qsti + ((pm * kbq + (size_t)(nt * (BN / 64) + j)) << 15) + roff;
            bulk_g2s(sb + 2 * Cfg::p_bytes + j * Cfg::panel, src, Cfg::panel, &full[stage]);
            if (lo) bulk_g2s(sb + 2 * Cfg::p_bytes + Cfg::q_bytes + j * Cfg::panel, src + 16384, Cfg::panel, &full[stage]);
          }
          for (int j = 0; j < np; ++j) {
            const uint8_t* src = psti + ((pm * kbp + (size_t)(mg * 2 * MT + j)) << 15) + roff;
            bulk_g2s(sb + j * Cfg::panel, src, Cfg::panel, &full[stage]);
            if (lo) bulk_g2s(sb + Cfg::p_bytes + j * Cfg::panel, src + 16384, Cfg::panel, &full[stage]);
          }
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, 1, 1);  // both operands MN-major
      constexpr uint32_t LBO = Cfg::panel >> 4;              // next 64-channel panel
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const int kbp = (g.pc + 63) / 64;
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
        int mg, nt, split, nkb;
        long long p_begin;
        decode(item, mg, nt, split, p_begin, nkb);
        const int buf = Cfg::nbuf == 2 ? (local & 1) : 0;
        const uint32_t bphase = Cfg::nbuf == 2 ? ((local >> 1) & 1) : (local & 1);
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        // row tiles of this group that hold at least one real panel (the rest of a short group is skipped)
        int nmt = (kbp - mg * 2 * MT + 1) / 2;
        nmt = nmt > MT ? MT : nmt;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::stage_bytes);
          const uint64_t q_hi = umma_desc_sw128(sa + 2 * Cfg::p_bytes, LBO, 64);
          const uint64_t q_lo = umma_desc_sw128(sa + 2 * Cfg::p_bytes + Cfg::q_bytes, LBO, 64);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m >= nmt) break;
            const uint32_t tmem_d = tmem_base + buf * (MT * BN) + m * BN;
            const uint64_t p_hi = umma_desc_sw128(sa + m * 2 * Cfg::panel, LBO, 64);
            const uint64_t p_lo = umma_desc_sw128(sa + Cfg::p_bytes + m * 2 * Cfg::panel, LBO, 64);
            // K = 16 pixels per MMA = two 8-row swizzle atoms = 2048 B = 128 x 16 B units
#pragma unroll
            for (int k = 0; k < KPIX / 16; ++k) umma_bf16(tmem_d, p_hi + 128 * k, q_hi + 128 * k, idesc, (kb | k) != 0);
            if (g.passes == 3) {
#pragma unroll
              for (int k = 0; k < KPIX / 16; ++k) umma_bf16(tmem_d, p_hi + 128 * k, q_lo + 128 * k, idesc, 1);
#pragma unroll
              for (int k = 0; k < KPIX / 16; ++k) umma_bf16(tmem_d, p_lo + 128 * k, q_hi + 128 * k, idesc, 1);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ================================ epilogue: TMEM -> split-K partial ====================
    const int q = warp & 3;
    int local = 0;
    const size_t per_split = (size_t)d.cout * d.cin;
    const int kbp = (g.pc + 63) / 64;
    for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
      int mg, nt, split, nkb;
      long long p_begin;
      decode(item, mg, nt, split, p_begin, nkb);
      const int buf = Cfg::nbuf == 2 ? (local & 1) : 0;
      const uint32_t bphase = Cfg::nbuf == 2 ? ((local >> 1) & 1) : (local & 1);
      mbar_wait<128>(&tfull[buf], bphase);
      tc_fence_after();
      float* out = partial + (size_t)split * per_split;
      int nmt = (kbp - mg * 2 * MT + 1) / 2;
      nmt = nmt > MT ? MT : nmt;
#pragma unroll 1
      for (int mi = 0; mi < nmt; ++mi) {
        const int m = (mg * MT + mi) * 128 + q * 32 + lane;  // P-channel of this thread's accumulator row
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (nt * BN + c0 >= g.qc) break;
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) + mi * BN + c0, v);
          if (m < g.pc) {
            const int n0 = nt * BN + c0;
            if (g.swap) {  // co = n, ci = m: consecutive lanes write consecutive floats
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.qc) out[(size_t)(n0 + j) * d.cin + m] = v[j];
            } else if (n0 + 32 <= g.qc) {  // co = m, ci = n: 128 contiguous bytes per thread (cin % 4 == 0)
              float4* o4 = reinterpret_cast<float4*>(out + (size_t)m * d.cin + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.qc) out[(size_t)m * d.cin + n0 + j] = v[j];
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int KPIX, int MT>
static int launch_ws(const NsrWgrad& d, const WgGeom& g, float* partial, cudaStream_t st) {
  using Cfg = WsCfg<BN, KPIX, MT>;
  static_assert(Cfg::stages >= 2, "at least two stages");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_sti<BN, KPIX, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes);
    if (e != cudaSuccess) {
      set_error("igemm_wgrad_sti<%d,%d,%d>: cudaFuncSetAttribute: %s", BN, KPIX, MT, cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int grid = g.num_items < kNumSMs ? g.num_items : kNumSMs;
  igemm_wgrad_sti<BN, KPIX, MT><<<grid, WS_THREADS, Cfg::smem_bytes, st>>>(d, g, partial);
  NSR_CHECK_LAUNCH("igemm_wgrad_sti");
  return NSR_OK;
}

template <int BN>
static int launch_ws_bn(const NsrWgrad& d, const WgGeom& g, float* partial, cudaStream_t st) {
  if (g.kpix == 32) return g.mt == 2 ? launch_ws<BN, 32, 2>(d, g, partial, st) : launch_ws<BN, 32, 1>(d, g, partial, st);
  if (g.mt == 1) return launch_ws<BN, 64, 1>(d, g, partial, st);
  if constexpr (WsCfg<BN, 64, 2>::stages >= 2) return launch_ws<BN, 64, 2>(d, g, partial, st);
  set_error("igemm_wgrad_sti<%d>: 64-pixel k-blocks of two row tiles do not fit two stages", BN);
  return NSR_E_INVALID;
}

int launch_wgrad_sti(const NsrWgrad& d, const WgGeom& g, int bn, float* partial, cudaStream_t st) {
  if ((g.kpix != 32 && g.kpix != 64) || (g.mt != 1 && g.mt != 2) || g.taps != 1 || !g.p_sti || !g.q_sti) {
    set_error("igemm_wgrad_sti: unsupported plan (kpix %d, mt %d, taps %d)", g.kpix, g.mt, g.taps);
    return NSR_E_INVALID;
  }
  switch (bn) {
    case 64: return launch_ws_bn<64>(d, g, partial, st);
    case 128: return launch_ws_bn<128>(d, g, partial, st);
    case 192: return launch_ws_bn<192>(d, g, partial, st);
    case 256: return launch_ws_bn<256>(d, g, partial, st);
    default: set_error("igemm_wgrad_sti: BN %d", bn); return NSR_E_INVALID;
  }
}

}  // namespace nsr

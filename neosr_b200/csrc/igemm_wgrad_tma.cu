// 3x3 (any kh x kw<=3, stride 1, "same") weight gradient on tcgen05 with TMA-staged operands.
//
//   dw[co][r][s][ci] = sum_p dy[p, co] * x[p + (r-pad, s-pad), ci]
//
// Operands are first split into bf16 (hi, lo) PLANES in NHWC order (x ~= hi + lo), so that TMA
// (cp.async.bulk.tensor, 5-D tile mode, SWIZZLE_128B) can stage them: a box of {64 channels, bw(+halo)
// pixels, bh rows} lands in shared memory as rows of 128 bytes - exactly the MN-major UMMA operand
// panel (row = pixel = K index).  Image-border padding is TMA's out-of-bounds zero fill; the tap shift
// is a coordinate offset.  Per k-block (bh x bw <= 64 pixels) one CTA loads
//     P : the unshifted operand,   128 channels  (rows of the accumulator),
//     Q : the shifted operand,     BN channels, ONE box with a (kw-1)-pixel halo in w,
// and issues the kw taps of one filter row against kw TMEM accumulators; tap s just starts its Q
// descriptor s rows (s * 128 B) further into the halo tile (the swizzle is a function of absolute
// shared-memory address bits, so a start that is not 1024-byte aligned needs no descriptor change).  So the shifted operand is fetched once per
// filter ROW instead of once per tap, no thread touches operand data, and the only warps are:
// 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM -> split-K partials).
// Products are formed as hi*hi + hi*lo + lo*hi (3 bf16 passes, fp32 accumulate), as in igemm_tc.cu.
#include <cuda.h>

#include <cstdlib>

#include "tc_common.cuh"

namespace nsr {
using namespace tc;

constexpr int WT_THREADS = 192;
constexpr int WT_MAX_ROWS_PER_SPLIT = 4096;  // accumulator truncation bound, see igemm_tc.cu
constexpr int WT_P_PANEL = 64 * 128;         // bytes of one 64-channel P panel (64 pixel rows)
constexpr int WT_Q_ROWS = 72;                // max rows of a Q halo tile: bh * (bw + 2)
constexpr int WT_Q_PANEL = WT_Q_ROWS * 128;

template <int KW>
struct WtCfg {
  static constexpr int bn = 64;                             // Q channels per tile (one 128-byte panel)
  static constexpr int n_mma = KW * bn;                     // N of one MMA = all taps of a filter row
  static constexpr int p_bytes = 2 * WT_P_PANEL;            // one plane (hi or lo) of P: 2 panels
  static constexpr int q_bytes = WT_Q_PANEL;                // one plane of Q
  static constexpr int stage_bytes = 2 * p_bytes + 2 * q_bytes;
  static constexpr int stages = 4;
  static constexpr int smem_bytes = stages * stage_bytes + 1024 + 256;
  static constexpr int nbuf = 2;                            // TMEM accumulator sets (2 * 192 columns)
};

struct WtGeom {
  int swap;             // 0: P = dy, Q = x shifted by +(r-pad, s-pad);  1: P = x, Q = dy shifted by -(...)
  int pc, qc;           // channels of P and Q
  int m_tiles, n_tiles, kh, kw, pad, splitk, num_items;
  int W, H, B, bw, bh, kpix, wblocks, hblocks;
  int nblocks, blocks_per_split;
  int passes;           // 3: hi*hi + hi*lo + lo*hi;  1: hi*hi only (NSR_ENGINE_BF16: the lo planes are not fetched)
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
// MN-major SWIZZLE_128B descriptor whose start may sit on any 128-byte row of the swizzle atom
// (measured on B200: the descriptor's base_offset field must stay 0 - the UMMA unit applies the 128-byte
// swizzle XOR to absolute shared-memory address bits, matching what TMA wrote; setting base_offset to
// (addr >> 7) & 7 or its negation produces wrong operands)
__device__ __forceinline__ uint64_t umma_desc_sw128_row(uint32_t smem_addr, uint32_t lbo_units) {
  return umma_desc_sw128(smem_addr, lbo_units, 64);
}

template <int KW>
__global__ void __launch_bounds__(WT_THREADS, 1) igemm_wgrad_tma(const __grid_constant__ CUtensorMap tm_p,
                                                                 const __grid_constant__ CUtensorMap tm_q, WtGeom g,
                                                                 int cin, int cout, float* __restrict__ partial) {
  using Cfg = WtCfg<KW>;
  constexpr int BN = Cfg::bn;
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

  // item -> (m_tile, n_tile, filter row, split)
  auto decode = [&](int item, int& mt, int& nt, int& r, int& split) {
    split = item % g.splitk;
    item /= g.splitk;
    r = item % g.kh;
    item /= g.kh;
    nt = item % g.n_tiles;
    mt = item / g.n_tiles;
  };
  auto block_range = [&](int split, int& b0, int& b1) {
    b0 = split * g.blocks_per_split;
    b1 = b0 + g.blocks_per_split;
    if (b1 > g.nblocks) b1 = g.nblocks;
  };
  const int qw = g.bw + KW - 1;  // halo tile width in pixels

  if (warp == 0) {
    // ================================ TMA producer ========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t p_box = (uint32_t)g.kpix * 128u, q_box = (uint32_t)(qw * g.bh) * 128u;
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x) {
        int mt, nt, r, split, b0, b1;
        decode(item, mt, nt, r, split);
        block_range(split, b0, b1);
        int np = (g.pc - mt * 128 + 63) / 64;
        np = np > 2 ? 2 : np;
        const int nq = 1;
        const int dh = g.swap ? -(r - g.pad) : (r - g.pad);
        for (int blk = b0; blk < b1; ++blk) {
          const int wb = blk % g.wblocks;
          const int t = blk / g.wblocks;
          const int hb = t % g.hblocks, b = t / g.hblocks;
          const int w0 = wb * g.bw, h0 = hb * g.bh;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * Cfg::stage_bytes;
          const int planes = g.passes == 3 ? 2 : 1;
          mbar_arrive_expect_tx(&full[stage], (uint32_t)planes * ((uint32_t)np * p_box + (uint32_t)nq * q_box));
          for (int plane = 0; plane < planes; ++plane) {
            for (int j = 0; j < np; ++j)
              tma_load_5d(sb + plane * Cfg::p_bytes + j * WT_P_PANEL, &tm_p, &full[stage], mt * 128 + j * 64, w0, h0, b, plane);
            for (int j = 0; j < nq; ++j)
              tma_load_5d(sb + 2 * Cfg::p_bytes + plane * Cfg::q_bytes + j * WT_Q_PANEL, &tm_q, &full[stage],
                          nt * BN + j * 64, w0 - g.pad, h0 + dh, b, plane);
          }
          if (++stage == Cfg::stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(Cfg::n_mma, 1, 1);  // both operands MN-major
      constexpr uint32_t LBO_P = WT_P_PANEL >> 4;
      constexpr uint32_t LBO_Q = 128 >> 4;  // N "panel" s = halo tile shifted by s pixel rows = tap s
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      // a k-block is nk <= 4 MMA K-steps of 16 pixels; their row offsets (in 16-byte descriptor units)
      // inside the P tile and the Q halo tile are the same for every k-block
      const int k16_per_row = g.bw / 16, nk = g.kpix / 16;
      uint32_t poff[4], qoff[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int h = i / k16_per_row, k = i - h * k16_per_row;
        poff[i] = (uint32_t)(h * g.bw + k * 16) * 8u;
        qoff[i] = (uint32_t)(h * qw + k * 16) * 8u;
      }
      for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
        int mt, nt, r, split, b0, b1;
        decode(item, mt, nt, r, split);
        block_range(split, b0, b1);
        const int buf = local & 1;
        const uint32_t bphase = (local >> 1) & 1;
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::n_mma;
        for (int blk = b0; blk < b1; ++blk) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::stage_bytes);
          const uint64_t p_hi = umma_desc_sw128(sa, LBO_P, 64), p_lo = umma_desc_sw128(sa + Cfg::p_bytes, LBO_P, 64);
          const uint64_t q_hi = umma_desc_sw128(sa + 2 * Cfg::p_bytes, LBO_Q, 64);
          const uint64_t q_lo = umma_desc_sw128(sa + 2 * Cfg::p_bytes + Cfg::q_bytes, LBO_Q, 64);
          const uint32_t first = blk != b0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nk) umma_bf16(tmem_d, p_hi + poff[i], q_hi + qoff[i], idesc, first | (uint32_t)i);
          if (g.passes == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nk) umma_bf16(tmem_d, p_hi + poff[i], q_lo + qoff[i], idesc, 1);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nk) umma_bf16(tmem_d, p_lo + poff[i], q_hi + qoff[i], idesc, 1);
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
    const int taps = g.kh * g.kw;
    const size_t per_split = (size_t)cout * taps * cin;
    for (int item = blockIdx.x; item < g.num_items; item += gridDim.x, ++local) {
      int mt, nt, r, split;
      decode(item, mt, nt, r, split);
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      mbar_wait<128>(&tfull[buf], bphase);
      tc_fence_after();
      float* out = partial + (size_t)split * per_split;
      const int m = mt * 128 + q * 32 + lane;  // P-channel of this thread's accumulator row
#pragma unroll 1
      for (int j = 0; j < KW; ++j) {           // accumulator column block j = halo offset j
        const int tap = r * KW + (g.swap ? (KW - 1 - j) : j);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (nt * BN + c0 >= g.qc) break;
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::n_mma + j * BN + c0, v);
          if (m < g.pc) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int n = nt * BN + c0 + e;
              if (n < g.qc) {
                const int co = g.swap ? n : m, ci = g.swap ? m : n;
                out[((size_t)co * taps + tap) * cin + ci] = v[e];
              }
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

// fp32 [rows, ld] (first C columns) -> bf16 hi plane and lo plane [rows, Cp], Cp = C rounded up to 8, pad = 0
__global__ void split_planes_kernel(const float* __restrict__ x, int ld, int C, int Cp, long long rows,
                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const int chunks = Cp / 8;
  const long long total = rows * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / chunks;
    const int c = (int)(i - row * chunks) * 8;
    const float* src = x + row * ld + c;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
    if (c + 4 <= C) f0 = __ldg(reinterpret_cast<const float4*>(src));
    if (c + 8 <= C) f1 = __ldg(reinterpret_cast<const float4*>(src + 4));
    uint4 h, l;
    split8(f0, f1, h, l);
    *reinterpret_cast<uint4*>(hi + row * Cp + c) = h;
    *reinterpret_cast<uint4*>(lo + row * Cp + c) = l;
  }
}

// ------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// planes: [2][B][H][W][Cp] bf16; box {64, box_w, box_h, 1, 1}
static bool make_plane_map(CUtensorMap* tm, void* planes, int B, int H, int W, int C, int Cp, int box_w, int box_h) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2,
                                 (cuuint64_t)B * H * W * Cp * 2};
  const cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WtPlan {
  WtGeom g;
  int bn, cpx, cpy, bias_blocks;
  size_t plane_x_bytes, plane_y_bytes, dw_partial_floats, bias_partial_floats;
};

static bool wt_blocking(int H, int W, int& bw, int& bh) {
  if (W % 64 == 0) bw = 64;
  else if (W <= 64 && W % 16 == 0) bw = W;
  else if (W % 48 == 0) bw = 48;
  else if (W % 32 == 0) bw = 32;
  else if (W % 16 == 0) bw = 16;
  else return false;
  bh = 1;
  while (bh * 2 * bw <= 64 && H % (bh * 2) == 0 && (bh * 2) * (bw + 2) <= WT_Q_ROWS) bh *= 2;
  return true;
}

static WtPlan wt_plan(const NsrWgrad& d) {
  WtPlan p;
  WtGeom& g = p.g;
  g.kh = d.kh; g.kw = d.kw; g.pad = d.pad;
  g.W = d.w; g.H = d.h; g.B = d.batch;
  wt_blocking(d.h, d.w, g.bw, g.bh);
  g.kpix = g.bw * g.bh;
  g.wblocks = d.w / g.bw;
  g.hblocks = d.h / g.bh;
  g.nblocks = d.batch * g.hblocks * g.wblocks;
  g.passes = mma_passes(d.engine);
  // orientation: P tiles are 128 channels, Q tiles 64; every (P tile, Q tile) pair costs the same
  const int units_a = ((d.cout + 127) / 128) * ((d.cin + 63) / 64);   // P = dy, Q = x
  const int units_b = ((d.cin + 127) / 128) * ((d.cout + 63) / 64);   // P = x,  Q = dy
  g.swap = units_b < units_a ? 1 : 0;
  p.bn = 64;
  g.pc = g.swap ? d.cin : d.cout;
  g.qc = g.swap ? d.cout : d.cin;
  g.m_tiles = (g.pc + 127) / 128;
  g.n_tiles = (g.qc + p.bn - 1) / p.bn;
  const int tiles = g.m_tiles * g.n_tiles * g.kh;
  const int max_blocks = WT_MAX_ROWS_PER_SPLIT / g.kpix;
  const int min_splits = (g.nblocks + max_blocks - 1) / max_blocks;
  const int waves = (tiles * min_splits + kNumSMs - 1) / kNumSMs;
  int want = (waves * kNumSMs) / tiles;
  if (want < min_splits) want = min_splits;
  const int maxs = (g.nblocks + 3) / 4;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  g.blocks_per_split = (g.nblocks + want - 1) / want;
  g.splitk = (g.nblocks + g.blocks_per_split - 1) / g.blocks_per_split;
  g.num_items = tiles * g.splitk;
  p.cpx = (d.cin + 7) / 8 * 8;
  p.cpy = (d.cout + 7) / 8 * 8;
  const size_t M = (size_t)d.batch * d.h * d.w;
  p.plane_x_bytes = (M * p.cpx * 2 * 2 + 1023) / 1024 * 1024;
  p.plane_y_bytes = (M * p.cpy * 2 * 2 + 1023) / 1024 * 1024;
  p.dw_partial_floats = (size_t)g.splitk * d.cout * g.kh * g.kw * d.cin;
  p.bias_blocks = bias_grad_blocks(M);
  p.bias_partial_floats = (size_t)p.bias_blocks * d.cout;
  return p;
}

bool conv_wgrad_tma_supported(const NsrWgrad& d) {
  static int ok_dev = -1;
  if (ok_dev < 0) ok_dev = nsr_device_supports_tcgen05() && encode_tiled() != nullptr;
  if (!ok_dev) return false;
  if (d.x == nullptr || d.dy == nullptr) return false;
  if (d.kw != 3 || d.kh != d.kw || d.pad != 1) return false;
  if (d.cin % 4 || d.cout % 4 || d.x_ld % 4 || d.dy_ld % 4 || d.cin < 16 || d.cout < 16) return false;
  if (((uintptr_t)d.x | (uintptr_t)d.dy) & 15) return false;
  int bw, bh;
  if (!wt_blocking(d.h, d.w, bw, bh)) return false;
  if ((size_t)d.batch * d.h * d.w * (size_t)((d.cin > d.cout ? d.cin : d.cout) + 8) * 2 >= ((size_t)1 << 40)) return false;
  return true;
}
size_t conv_wgrad_workspace_tma(const NsrWgrad& d) {
  WtPlan p = wt_plan(d);
  return p.plane_x_bytes + p.plane_y_bytes + (p.dw_partial_floats + p.bias_partial_floats) * sizeof(float) + 1024;
}

int launch_wgrad_reduce(const float* partial, float* dw, int splitk, int cout, int taps, int cin, cudaStream_t st);
int conv_bias_grad(const NsrWgrad& d, float* bias_partial, int bias_blocks, cudaStream_t st);

template <int KW>
static int launch_wgrad_tma(const CUtensorMap& tp, const CUtensorMap& tq, const NsrWgrad& d, const WtPlan& p, float* partial,
                            cudaStream_t st) {
  using Cfg = WtCfg<KW>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_tma<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes);
    if (e != cudaSuccess) {
      set_error("igemm_wgrad_tma<%d>: cudaFuncSetAttribute: %s", KW, cudaGetErrorString(e));
      return NSR_E_CUDA;
    }
    attr = true;
  }
  const int grid = p.g.num_items < kNumSMs ? p.g.num_items : kNumSMs;
  igemm_wgrad_tma<KW><<<grid, WT_THREADS, Cfg::smem_bytes, st>>>(tp, tq, p.g, d.cin, d.cout, partial);
  NSR_CHECK_LAUNCH("igemm_wgrad_tma");
  return NSR_OK;
}

int conv_wgrad_tma(const NsrWgrad& d, cudaStream_t st) {
  WtPlan p = wt_plan(d);
  const size_t need = conv_wgrad_workspace_tma(d);
  if (d.workspace_bytes < need || d.workspace == nullptr) {
    set_error("nsr_conv_wgrad(tma): workspace %zu < %zu", d.workspace_bytes, need);
    return NSR_E_WORKSPACE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)d.workspace + 1023) & ~(uintptr_t)1023);
  uint16_t* px = reinterpret_cast<uint16_t*>(ws);
  uint16_t* py = reinterpret_cast<uint16_t*>(ws + p.plane_x_bytes);
  float* partial = reinterpret_cast<float*>(ws + p.plane_x_bytes + p.plane_y_bytes);
  const long long M = (long long)d.batch * d.h * d.w;
  {
    const long long tx = M * (p.cpx / 8), ty = M * (p.cpy / 8);
    const long long cap = (long long)kNumSMs * 16;
    long long bx = (tx + 255) / 256, by = (ty + 255) / 256;
    split_planes_kernel<<<(int)(bx < cap ? bx : cap), 256, 0, st>>>(d.x, d.x_ld, d.cin, p.cpx, M, px, px + M * p.cpx);
    split_planes_kernel<<<(int)(by < cap ? by : cap), 256, 0, st>>>(d.dy, d.dy_ld, d.cout, p.cpy, M, py, py + M * p.cpy);
    NSR_CHECK_LAUNCH("split_planes");
  }
  const WtGeom& g = p.g;
  CUtensorMap tm_x_plain, tm_x_halo, tm_y_plain, tm_y_halo;
  // P is loaded with the plain box, Q with the halo box
  const bool okm = g.swap
      ? (make_plane_map(&tm_x_plain, px, d.batch, d.h, d.w, d.cin, p.cpx, g.bw, g.bh) &&
         make_plane_map(&tm_y_halo, py, d.batch, d.h, d.w, d.cout, p.cpy, g.bw + g.kw - 1, g.bh))
      : (make_plane_map(&tm_y_plain, py, d.batch, d.h, d.w, d.cout, p.cpy, g.bw, g.bh) &&
         make_plane_map(&tm_x_halo, px, d.batch, d.h, d.w, d.cin, p.cpx, g.bw + g.kw - 1, g.bh));
  if (!okm) {
    set_error("nsr_conv_wgrad(tma): cuTensorMapEncodeTiled failed");
    return NSR_E_CUDA;
  }
  const CUtensorMap& tp = g.swap ? tm_x_plain : tm_y_plain;
  const CUtensorMap& tq = g.swap ? tm_y_halo : tm_x_halo;
  int rc = launch_wgrad_tma<3>(tp, tq, d, p, partial, st);
  if (rc != NSR_OK) return rc;
  rc = launch_wgrad_reduce(partial, d.dw, g.splitk, d.cout, g.kh * g.kw, d.cin, st);
  if (rc != NSR_OK) return rc;
  if (d.dbias) return conv_bias_grad(d, partial + p.dw_partial_floats, p.bias_blocks, st);
  return NSR_OK;
}

}  // namespace nsr

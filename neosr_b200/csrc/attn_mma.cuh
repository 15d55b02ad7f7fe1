// mma.sync (m16n8k16, bf16 operands split hi/lo in three passes, fp32 accumulate) building blocks shared by the
// window-attention kernels (window_attn_mma.cu: 8x8 windows; xwin_attn_mma.cu: HAT's 16x16 / 24x24 windows).
#pragma once
#include "common.cuh"

namespace nsr {

constexpr int AM_LD = 40;   // bf16 row stride of [token][d] tiles (80 B: conflict-free fragment loads)
constexpr int AM_LDT = 72;  // bf16 row stride of [d][token] and [token][token] tiles (144 B)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (a_hi + a_lo) * (b_hi + b_lo) without the lo*lo term
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_bf16(c, ah, bh0, bh1);
  mma_bf16(c, ah, bl0, bl1);
  mma_bf16(c, al, bh0, bh1);
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t*>(&h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows x 16 k) from a row-major [row][k] bf16 tile
__device__ __forceinline__ void load_a(const __nv_bfloat16* base, int ld, int row0, int k0, int g, int tid,
                                       uint32_t (&a)[4]) {
  const __nv_bfloat16* p = base + (row0 + g) * ld + k0 + tid * 2;
  a[0] = lds32(p);
  a[1] = lds32(p + 8 * ld);
  a[2] = lds32(p + 8);
  a[3] = lds32(p + 8 * ld + 8);
}
// B fragment (16 k x 8 n) from an [n][k] bf16 tile (k contiguous)
__device__ __forceinline__ void load_b(const __nv_bfloat16* base, int ld, int n0, int k0, int g, int tid, uint32_t& b0,
                                       uint32_t& b1) {
  const __nv_bfloat16* p = base + (n0 + g) * ld + k0 + tid * 2;
  b0 = lds32(p);
  b1 = lds32(p + 8);
}

// S = Qs K^T for this warp's 16 rows: acc[nt] covers columns 8nt..8nt+7
__device__ __forceinline__ void qk_scores(const __nv_bfloat16* Ah, const __nv_bfloat16* Al, const __nv_bfloat16* Bh,
                                          const __nv_bfloat16* Bl, int row0, int g, int tid, float (&acc)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t ah[4], al[4];
    load_a(Ah, AM_LD, row0, kk * 16, g, tid, ah);
    load_a(Al, AM_LD, row0, kk * 16, g, tid, al);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t bh0, bh1, bl0, bl1;
      load_b(Bh, AM_LD, nt * 8, kk * 16, g, tid, bh0, bh1);
      load_b(Bl, AM_LD, nt * 8, kk * 16, g, tid, bl0, bl1);
      mma3(acc[nt], ah, al, bh0, bh1, bl0, bl1);
    }
  }
}

// out[16 x 32] = X[16 x 64] (accumulator layout, as A operand) * Bt, Bt = [n = d][k = token] tiles
__device__ __forceinline__ void acc_times(const float (&x)[8][4], const __nv_bfloat16* Bh, const __nv_bfloat16* Bl,
                                          int g, int tid, float (&o)[4][4]) {
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
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bh0, bh1, bl0, bl1;
      load_b(Bh, AM_LDT, nt * 8, kk * 16, g, tid, bh0, bh1);
      load_b(Bl, AM_LDT, nt * 8, kk * 16, g, tid, bl0, bl1);
      mma3(o[nt], ah, al, bh0, bh1, bl0, bl1);
    }
  }
}

}  // namespace nsr

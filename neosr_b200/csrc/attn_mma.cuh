// mma.sync (m16n8k16, bf16 operands split hi/lo in three passes, fp32 accumulate) building blocks shared by the
// window-attention kernels (window_attn_mma.cu: 8x8 windows; xwin_attn_mma.cu: HAT's 16x16 / 24x24 windows).
#pragma once
#include "common.cuh"

namespace nsr {

constexpr int AM_LD = 40;   // bf16 row stride of [token][d] tiles (80 B: conflict-free fragment loads)

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
// ldmatrix: four 8x8 bf16 matrices per instruction; lanes 8m..8m+7 supply the 16-byte row addresses of matrix m.
// Plain form: thread T receives (row T/4, columns 2(T%4), 2(T%4)+1) of each matrix - the mma A / B fragment layout
// of a [row][k] tile.  .trans: thread T receives (rows 2(T%4), 2(T%4)+1, column T/4) - the B fragment of a tile
// stored [k][n], which is how the backward passes read K / Q / dO without keeping transposed copies.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}

// acc[16 x 64] = A[row0..+15][0..31] B^T (S = Qs K^T for a warp's 16 rows; acc[nt] covers columns 8nt..8nt+7),
// A and B both [token][d] tiles with row stride AM_LD, fragments through ldmatrix
__device__ __forceinline__ void qk_scores_ldm(const __nv_bfloat16* Ah, const __nv_bfloat16* Al, const __nv_bfloat16* Bh,
                                              const __nv_bfloat16* Bl, int row0, int lane, float (&acc)[8][4]) {
  const int lr = lane & 7, lm = lane >> 3;
  const int a_off = (row0 + (lm & 1) * 8 + lr) * AM_LD + (lm >> 1) * 8;  // matrices: (rows, k), (rows+8, k), (rows, k+8), (rows+8, k+8)
  const int b_off = ((lm >> 1) * 8 + lr) * AM_LD + (lm & 1) * 8;         // matrices: (n, k), (n, k+8), (n+8, k), (n+8, k+8)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t ah[4], al[4];
    ldsm_x4(ah, Ah + a_off + kk * 16);
    ldsm_x4(al, Al + a_off + kk * 16);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t bh[4], bl[4];
      ldsm_x4(bh, Bh + b_off + p * 16 * AM_LD + kk * 16);
      ldsm_x4(bl, Bl + b_off + p * 16 * AM_LD + kk * 16);
      mma3(acc[2 * p], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma3(acc[2 * p + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}

// out[16 x 32] = X[16 x 64] (accumulator layout, as A operand) * B, B = [k = token][n = d] tiles with row stride AM_LD
// (the tiles as loaded - fragments come through ldmatrix.trans)
__device__ __forceinline__ void acc_times_ldm(const float (&x)[8][4], const __nv_bfloat16* Bh, const __nv_bfloat16* Bl,
                                              int lane, float (&o)[4][4]) {
  const int lr = lane & 7, lm = lane >> 3;
  const int off = ((lm & 1) * 8 + lr) * AM_LD + (lm >> 1) * 8;  // matrices: (k, n), (k+8, n), (k, n+8), (k+8, n+8)
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
      ldsm_x4_t(bh, Bh + off + kk * 16 * AM_LD + q * 16);
      ldsm_x4_t(bl, Bl + off + kk * 16 * AM_LD + q * 16);
      mma3(o[2 * q], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma3(o[2 * q + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}

}  // namespace nsr

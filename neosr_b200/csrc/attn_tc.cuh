// Shared pieces of the tcgen05 window-attention kernels (window_attn_tc.cu: forward, window_attn_tc_bwd.cu: backward).
#pragma once
#include "attn_mma.cuh"
#include "tc_common.cuh"

namespace nsr {

constexpr int AT_BLK = 32768;  // one STI block: 128 rows x 64 channels, bf16 hi image 16 KiB + lo image 16 KiB
constexpr int AT_MAX_HEADS = 8;

struct AtGeom {
  int B, H, W, C, heads, ws, shift, use_mask, D, nwh, nww, G, nwin;
  float scale;
  int pad_out;  // out_sti is [tokens, G]: heads padded to 32 channels (whole 16-byte chunks per thread), 1.0 in channel D
};

// token index and shift-mask region id of row n (0..63) of window wi (same map as attn_token_map)
__device__ __forceinline__ void at_token_map(const AtGeom& g, int wi, int n, int& tok, int& rid) {
  const int per = g.nwh * g.nww;
  const int b = wi / per, rem = wi - b * per;
  const int wy = rem / g.nww, wx = rem - wy * g.nww;
  const int iy = n >> 3, ix = n & 7;
  const int hs = wy * 8 + iy, wsx = wx * 8 + ix;
  int ho = hs + g.shift, wo = wsx + g.shift;
  if (ho >= g.H) ho -= g.H;
  if (wo >= g.W) wo -= g.W;
  tok = (b * g.H + ho) * g.W + wo;
  const int rh = hs < g.H - 8 ? 0 : (hs < g.H - g.shift ? 1 : 2);
  const int rw = wsx < g.W - 8 ? 0 : (wsx < g.W - g.shift ? 1 : 2);
  rid = rh * 3 + rw;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace nsr

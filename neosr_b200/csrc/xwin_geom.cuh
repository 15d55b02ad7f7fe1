// Geometry and index math of the generic (cross-)window attention (HAT's HAB / OCAB), shared by the exact-fp32
// kernels (hat_ops.cu) and the tensor-core kernels (xwin_attn_mma.cu).
#pragma once
#include "common.cuh"

namespace nsr {

struct GAGeom {
  int B, H, W, C, heads, D, ws, ows, pad, shift, use_mask, oca, nwh, nww, Nq, Nk, L, ntab;
  float scale;
};

// query n of window wi -> token index + mask region id (shifted frame), HAB only uses shift/mask
__device__ __forceinline__ void ga_query(const GAGeom& g, int wi, int n, int& tok, int& rid, int& iy, int& ix) {
  const int per = g.nwh * g.nww, b = wi / per, rem = wi - b * per, wy = rem / g.nww, wx = rem - wy * g.nww;
  iy = n / g.ws; ix = n - iy * g.ws;
  const int hs = wy * g.ws + iy, wsx = wx * g.ws + ix;
  int ho = hs + g.shift, wo = wsx + g.shift;  // torch.roll(x, -shift)
  if (ho >= g.H) ho -= g.H;
  if (wo >= g.W) wo -= g.W;
  tok = (b * g.H + ho) * g.W + wo;
  const int rh = hs < g.H - g.ws ? 0 : (hs < g.H - g.shift ? 1 : 2), rw = wsx < g.W - g.ws ? 0 : (wsx < g.W - g.shift ? 1 : 2);
  rid = rh * 3 + rw;
}
// key n of window wi -> token index (-1: zero padding of nn.Unfold) + region id
__device__ __forceinline__ void ga_key(const GAGeom& g, int wi, int n, int& tok, int& rid, int& jy, int& jx) {
  const int per = g.nwh * g.nww, b = wi / per, rem = wi - b * per, wy = rem / g.nww, wx = rem - wy * g.nww;
  jy = n / g.ows; jx = n - jy * g.ows;
  const int hs = wy * g.ws - g.pad + jy, wsx = wx * g.ws - g.pad + jx;
  if (hs < 0 || hs >= g.H || wsx < 0 || wsx >= g.W) { tok = -1; rid = 0; return; }
  int ho = hs + g.shift, wo = wsx + g.shift;
  if (ho >= g.H) ho -= g.H;
  if (wo >= g.W) wo -= g.W;
  tok = (b * g.H + ho) * g.W + wo;
  const int rh = hs < g.H - g.ws ? 0 : (hs < g.H - g.shift ? 1 : 2), rw = wsx < g.W - g.ws ? 0 : (wsx < g.W - g.shift ? 1 : 2);
  rid = rh * 3 + rw;
}
// relative_position_index: SA (hat_arch.py:1015-1033) / OCA (1035-1068; negative entries index the table from
// its end, as Python indexing does in the reference)
__device__ __forceinline__ int ga_rel(const GAGeom& g, int iy, int ix, int jy, int jx) {
  if (!g.oca) return (iy - jy + g.ws - 1) * g.L + (ix - jx + g.ws - 1);
  const int off = g.ws - g.ows + 1;
  int e = (jy - iy + off) * g.L + (jx - ix + off);
  if (e < 0) e += g.ntab;
  return e;
}


}  // namespace nsr

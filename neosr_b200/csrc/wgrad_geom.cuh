// Work decomposition of the tcgen05 weight-gradient kernels (igemm_tc.cu: generic, igemm_wgrad_sti.cu: split-tile-image
// operands), shared so that workspace sizing and both launch paths agree.
#pragma once
#include "common.cuh"

namespace nsr {

struct WgGeom {
  int swap;            // 0: P = dy (rows = cout), Q = x (cols = cin); 1: P = x, Q = dy
  int pc, qc;          // channel counts of P and Q
  int p_ld, q_ld;
  int m_tiles, n_tiles, taps, splitk, num_items;
  int mt;              // 128-row tiles of P per work item (1, or 2 on the STI kernel: Q is loaded once for both)
  int m_groups;        // ceil(m_tiles / mt)
  int kpix;            // pixels per k-block
  int passes;          // 3: hi*hi + hi*lo + lo*hi;  1: hi*hi only (NSR_ENGINE_BF16)
  long long M, rows_per_split;
  const float* p_ptr;
  const float* q_ptr;
  const void* p_sti;
  const void* q_sti;
};

// igemm_wgrad_sti.cu: dW partials from split tile images; g.mt / g.kpix select the instantiation
int launch_wgrad_sti(const NsrWgrad& d, const WgGeom& g, int bn, float* partial, cudaStream_t st);

}  // namespace nsr

// Shared definitions for the triplane field kernels (query_triplane + NeRFMLP).
// Semantics restated from the reference (paths relative to /root/reference):
//   TripoSR/tsr/models/nerf_renderer.py:41-91   query_triplane
//   TripoSR/tsr/models/network_utils.py:35-124  NeRFMLP (Linear/SiLU chain)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sculptmate_b200.h"

namespace smb {

constexpr int kCp = SMB_PLANE_CHANNELS;      // 40 channels per plane
constexpr int kFeat = 3 * kCp;               // 120 = NeRFMLP in_channels
constexpr int kHid = SMB_HIDDEN;             // 64 neurons
constexpr int kOut = 4;                      // density + 3 features
constexpr int kMaxHidden = SMB_MAX_HIDDEN_LAYERS;

// aten grid_sampler "unnormalize" (CPU form): align_corners=False maps [-1,1] to
// [-0.5, size-0.5]; align_corners=True maps it to [0, size-1].
__device__ __forceinline__ float unnormalize(float u, int size, int align_corners) {
  if (align_corners) return __fmul_rn(__fadd_rn(u, 1.0f), 0.5f * (float)(size - 1));
  return __fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), 0.5f * (float)size), 0.5f);
}

struct Tap2 {  // 1-D linear tap pair with zero padding folded into the weights
  int i0, i1;  // clamped indices (always valid addresses)
  float w0, w1;
};
__device__ __forceinline__ Tap2 make_tap(float u, int size, int align_corners) {
  float f = unnormalize(u, size, align_corners);
  float fl = floorf(f);
  float w1 = __fsub_rn(f, fl);
  float w0 = __fsub_rn(1.0f, w1);
  int i0 = (int)fl, i1 = i0 + 1;
  Tap2 t;
  t.w0 = (i0 >= 0 && i0 < size) ? w0 : 0.0f;
  t.w1 = (i1 >= 0 && i1 < size) ? w1 : 0.0f;
  t.i0 = min(max(i0, 0), size - 1);
  t.i1 = min(max(i1, 0), size - 1);
  return t;
}

// positions in (-radius, radius) -> (-1, 1): utils.py:222-231 as aten CPU
// evaluates it with python-float scales: ((p - a0) / (a1 - a0)) * (b1 - b0) + b0.
struct PosScale {
  float sub, div, mul, add;
};
__device__ __forceinline__ float scale_pos(float p, const PosScale& s) {
  return __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, s.sub), s.div), s.mul), s.add);
}

inline int smb_check(cudaError_t e) { return e == cudaSuccess ? SMB_OK : SMB_ERR_CUDA; }

}  // namespace smb

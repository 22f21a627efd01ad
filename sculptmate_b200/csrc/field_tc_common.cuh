// Shared declarations of the lattice tensor-core kernels (field_tc.cu, field_tc_ta.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "field_common.cuh"
#include "ptx_sm100.cuh"

namespace smb {

constexpr int kTileM = 128;
constexpr int kTRowsMax = 66;  // Hp + 2 zero borders for Hp = 64
constexpr int kTPitch = 68;    // floats per T row: 64 + 4 pad (adjacent rows land 4 banks apart)
constexpr int kWBytes = kHid * kHid * 2;     // 8192: one hidden layer, fp16
constexpr int kWFinalBytes = 16 * kHid * 2;  // 2048: head padded to N=16
constexpr int kABytes = kTileM * kHid * 2;   // 16384
constexpr int kSlotsPerWG = 2;

struct TcParams {
  const float* planes_q;  // (3,H,W,64) fp32
  const unsigned char* tc_weights;  // blob + off_tc_hidden: hidden images, head image, biases (contiguous)
  const unsigned char* tc_biasblk;  // blob + off_tc_biasblk: (n_hidden-1) bias K-block images
  const float* bias0_half;          // b0/2 (64)
  const float* axis_u;
  int R, x_begin, nx, H, W, align_corners, n_hidden;
  int trows;       // rows of a slot's T table
  int slot_bytes;  // 1024-aligned: A | T | c
  float density_bias;
  float* out_act;
  float* out_raw;
  int wait_ns;  // suspend-time hint of the consumer mbarrier waits (0: plain try_wait polling)
  uint32_t* sign_out;  // optional: marching-cubes sign masks of the slab (word = line * sign_wz + k/32)
  float sign_sub, sign_mul;
  int sign_wz;
  int stagger_clk;  // field_tc_ta: warpgroup g starts its first tile g*stagger_clk clocks late (breaks the SFU convoy)
  int xu_tokens;    // field_tc_ta: > 0 = at most this many warps per SM sub-partition inside an activation stretch
  int dbg;  // 1: developer timeline instrumentation (SMB_TC_TRACE); 0 in production
};

__host__ __device__ inline int tc_weight_bytes(int n_hidden) {
  // the blob keeps [hidden | head | bias_half (n_hidden x 64) | bias_final (4)] contiguous
  return (n_hidden - 1) * kWBytes + kWFinalBytes + n_hidden * kHid * 4 + 16;
}
__host__ __device__ inline int tc_slot_bytes(int trows) {
  return ((kABytes + trows * kTPitch * 4 + kHid * 4 + 1023) / 1024) * 1024;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t atom_inc_acq_rel(uint32_t addr) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}


struct TileGeom {  // what producer and consumer both derive from a tile index
  long long line;  // (i*R + j)
  int i, j, k0, nvalid, hlo, nrow;
};

__device__ __forceinline__ TileGeom tile_geom(long long t, int tiles_per_line, const TcParams& p) {
  TileGeom g;
  const int seg = (int)(t % tiles_per_line);
  g.line = t / tiles_per_line;
  g.i = (int)(g.line / p.R);
  g.j = (int)(g.line - (long long)g.i * p.R);
  g.k0 = seg * kTileM;
  g.nvalid = min(kTileM, p.R - g.k0);
  const float fz_first = unnormalize(p.axis_u[g.k0], p.H, p.align_corners);
  const float fz_last = unnormalize(p.axis_u[g.k0 + g.nvalid - 1], p.H, p.align_corners);
  g.hlo = (int)floorf(fz_first);
  g.nrow = min((int)floorf(fz_last) + 1 - g.hlo + 1, p.trows);
  return g;
}

// field_tc_ta.cu: the activations-in-TMEM variant
int launch_tc_ta(const TcParams& p, int sms, cudaStream_t st);
int read_trace_ta(long long* host, int n);  // SMB_TC_TRACE=2 timeline of the last launch
// mcubes.cu: sign masks of a density slab (the stand-alone pass the fused path replaces) / where they live
int launch_mc_signs(const float* grid, int nx, int ny, int nz, float sub, float sign, void* workspace, size_t workspace_bytes,
                    cudaStream_t st);

}  // namespace smb

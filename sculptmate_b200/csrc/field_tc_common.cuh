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
  // per-launch axis tables (lattice_api.cu): t1[i][h] = x-interpolated row h of the (x,z) plane at lattice x = x_begin + i,
  // t2[j][h] = the same for the (y,z) plane at lattice y = j; (rows, H, 64) fp32.  A tile's table row is t1 + t2.
  const float* t1;
  const float* t2;
  const float* head_w_f32;          // row 0 of the last Linear, fp32 (64): the density head, fused into the last epilogue
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
  // 32-bit arithmetic: the launchers reject lattices with 2^31 tiles or more (64-bit divisions cost ~100 instructions each
  // and this runs once per tile in every producer lane)
  const unsigned tu = (unsigned)t, tpl = (unsigned)tiles_per_line, Ru = (unsigned)p.R;
  const unsigned line = tu / tpl;
  const int seg = (int)(tu - line * tpl);
  g.line = (long long)line;
  g.i = (int)(line / Ru);
  g.j = (int)(line - (unsigned)g.i * Ru);
  g.k0 = seg * kTileM;
  g.nvalid = min(kTileM, p.R - g.k0);
  const float fz_first = unnormalize(p.axis_u[g.k0], p.H, p.align_corners);
  const float fz_last = unnormalize(p.axis_u[g.k0 + g.nvalid - 1], p.H, p.align_corners);
  g.hlo = (int)floorf(fz_first);
  g.nrow = min((int)floorf(fz_last) + 1 - g.hlo + 1, p.trows);
  return g;
}

// ---- operand plumbing shared by the activations-in-TMEM kernels (field_tc_pair.cu, field_tc_ta.cu)
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// thread i writes 8 consecutive 32-bit columns of TMEM lane (base+i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline int ta_table_bytes(int trows) { return ((trows * kTPitch * 4 + kHid * 4 + 127) / 128) * 128; }

// Layer-0 table of tile t (z rows of the (.,z) planes at this line's x / y taps, and the constant vector of the
// (x,y) plane + bias): one producer warp; the caller owns the hand-off barriers.
struct TableGeom {
  TileGeom tg;
  Tap2 txw, tyw, tyh;
};
__device__ __forceinline__ TableGeom table_geom(const TcParams& p, long long t, int tiles_per_line) {
  TableGeom g;
  g.tg = tile_geom(t, tiles_per_line, p);
  const float ux = p.axis_u[p.x_begin + g.tg.i];
  const float uy = p.axis_u[g.tg.j];
  g.txw = make_tap(ux, p.W, p.align_corners);
  g.tyw = make_tap(uy, p.W, p.align_corners);
  g.tyh = make_tap(uy, p.H, p.align_corners);
  return g;
}
__device__ __forceinline__ void build_table(const TcParams& p, const TableGeom& G, float* sT, int lane, int part = 0, int nparts = 1) {
  const long long HW = (long long)p.H * p.W;
  const float* Q0 = p.planes_q;
  const float* Q1 = Q0 + HW * kHid;
  const float* Q2 = Q1 + HW * kHid;
  float* sC = sT + p.trows * kTPitch;
  const TileGeom& tg = G.tg;
  const Tap2 &txw = G.txw, &tyw = G.tyw, &tyh = G.tyh;
  if (part == 0) {
    const float2 q00 = __ldg(reinterpret_cast<const float2*>(Q0 + ((long long)tyh.i0 * p.W + txw.i0) * kHid) + lane);
    const float2 q01 = __ldg(reinterpret_cast<const float2*>(Q0 + ((long long)tyh.i0 * p.W + txw.i1) * kHid) + lane);
    const float2 q10 = __ldg(reinterpret_cast<const float2*>(Q0 + ((long long)tyh.i1 * p.W + txw.i0) * kHid) + lane);
    const float2 q11 = __ldg(reinterpret_cast<const float2*>(Q0 + ((long long)tyh.i1 * p.W + txw.i1) * kHid) + lane);
    const float2 b0 = __ldg(reinterpret_cast<const float2*>(p.bias0_half) + lane);
    float2 c;
    c.x = b0.x + (tyh.w0 * (txw.w0 * q00.x + txw.w1 * q01.x) + tyh.w1 * (txw.w0 * q10.x + txw.w1 * q11.x));
    c.y = b0.y + (tyh.w0 * (txw.w0 * q00.y + txw.w1 * q01.y) + tyh.w1 * (txw.w0 * q10.y + txw.w1 * q11.y));
    reinterpret_cast<float2*>(sC)[lane] = c;
  }
  {
    const int n4 = lane & 15;
#pragma unroll 4
    for (int r = 2 * part + (lane >> 4); r < tg.nrow; r += 2 * nparts) {
      const int h = tg.hlo + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (h >= 0 && h < p.H && p.t1) {
        // both in-plane interpolations were done once per launch for every lattice x / y (axis tables): two loads and an add
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.t1 + ((long long)tg.i * p.H + h) * kHid) + n4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.t2 + ((long long)tg.j * p.H + h) * kHid) + n4);
        v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      } else if (h >= 0 && h < p.H) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(Q1 + ((long long)h * p.W + txw.i0) * kHid) + n4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(Q1 + ((long long)h * p.W + txw.i1) * kHid) + n4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(Q2 + ((long long)h * p.W + tyw.i0) * kHid) + n4);
        const float4 d = __ldg(reinterpret_cast<const float4*>(Q2 + ((long long)h * p.W + tyw.i1) * kHid) + n4);
        v.x = txw.w0 * a.x + txw.w1 * b.x + tyw.w0 * c.x + tyw.w1 * d.x;
        v.y = txw.w0 * a.y + txw.w1 * b.y + tyw.w0 * c.y + tyw.w1 * d.y;
        v.z = txw.w0 * a.z + txw.w1 * b.z + tyw.w0 * c.z + tyw.w1 * d.z;
        v.w = txw.w0 * a.w + txw.w1 * b.w + tyw.w0 * c.w + tyw.w1 * d.w;
      }
      *reinterpret_cast<float4*>(sT + r * kTPitch + 4 * n4) = v;
    }
  }
}

// field_tc_pair.cu: tile pairs sharing one accumulator (default); poly_pairs of every 8 activation pairs on the FMA pipe
int launch_tc_pair(const TcParams& p, int sms, int poly_pairs, cudaStream_t st);
// field_tc_ta.cu: the activations-in-TMEM variant
int launch_tc_ta(const TcParams& p, int sms, cudaStream_t st);
void keep_async_scratch(int dev);  // lattice_api.cu: release threshold of the default memory pool
int read_trace_ta(long long* host, int n);  // developer build: SMB_TC_TRACE=2 timeline of the last launch
// mcubes.cu: sign masks of a density slab (the stand-alone pass the fused path replaces) / where they live
int launch_mc_signs(const float* grid, int nx, int ny, int nz, float sub, float sign, void* workspace, size_t workspace_bytes,
                    cudaStream_t st);

}  // namespace smb

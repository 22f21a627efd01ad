// SF3D on a LATTICE-ordered tetrahedral grid: query_triplane + the MaterialMLP heads `density` and `vertex_offset` at every
// grid vertex (StableFast/sf3d/system.py:141-198, sf3d/models/network.py:148-208) without gathering a single plane texel per
// vertex.
//
// When the tet grid's vertex array is an outer product of three coordinate lists (sf3d/models/isosurface.py
// detects that once at load; a Kuhn / marching-cubes-style grid is), vertex (a, b, c) of the lattice samples the three planes
// at positions that depend on TWO lattice indices each.  Interpolation and the first Linear are linear, so
//     W0 . [f_xy ; f_xz ; f_yz] + b0  =  C[a][b] + T1[a][c] + T2[b][c]
// with three tables of n^2 x 64 entries per head (3 x 161^2 instead of 161^3 bilinear gathers of 120 channels and 120 x 64
// contractions), built per call by tetgrid_tables_kernel in fp32 -- closer to the reference's fp32 than the fp16 operands of
// the points kernel (field_pts_tc.cu), which stays the path for arbitrary positions.
//
// tetgrid_tc_kernel then is the lattice kernel of field_tc_ta.cu with a different front end: a tile = 128 consecutive rows
// (b, c) of one lattice plane a (rows flattened, so 161-long lines do not leave a 33-row tail tile); 4 producer warps sum the
// three table rows of every sample into a shared-memory tile (coalesced 16-byte loads from the L2-resident tables, 16-byte
// chunks XOR-swizzled so that the consumers' row-per-thread LDS.128 are conflict-free); 4 consumer warpgroups apply SiLU,
// write the fp16 activations to TENSOR MEMORY as the A operand of the hidden layer's tcgen05.mma (weights and the bias
// K-block resident in shared memory), and evaluate the head (1 or 3 outputs) as fp32 dot products in the last epilogue.
// Both heads run in one launch (the tile index carries the head).
#include <cuda_runtime.h>
#include <math.h>

#include "field_tc_common.cuh"

namespace smb {

constexpr int kTgWG = 4;         // consumer warpgroups (4 x 96 + 8 = 392 TMEM columns); 4 producer warps, 640 threads
constexpr int kTgMaxHeads = 2;
// A tile = a 4 x 4 x 8 block (slow x mid x fast index) of the lattice = 128 samples that need only 32 + 32 + 16 DISTINCT table
// rows (20 KB) -- 0.6 rows per sample instead of 2: the tables come from L2, and two rows per sample made the kernel
// L2-latency-bound (780 us: the consumers spent half their instructions waiting for the producers).
constexpr int kTgA = 4, kTgB = 4, kTgC = 8;
constexpr int kTgRow = kHid * 4;                                          // 256 B: one table row
constexpr int kTgStageBytes = (kTgA * kTgC + kTgB * kTgC + kTgA * kTgB) * kTgRow;  // T1 | T2 | C = 20 KB
constexpr int kTgStages = 2;                                              // per warpgroup

struct TgHead {
  const float* C;   // [nA][nB][64]  (b0/2 folded in)
  const float* T1;  // [nA][nC][64]  16-byte chunks of a row stored at chunk ^ (c & 7)
  const float* T2;  // [nB][nC][64]  same
  const unsigned char* tc_weights;  // (n_hidden-1) x 8 KB fp16 UMMA images of W_l/2
  const unsigned char* tc_biasblk;  // (n_hidden-1) x 8 KB bias K-blocks
  const float* head_w;              // last Linear, (4,64) fp32 row-major (rows >= n_out unused), then its bias (4)
  float* out;                       // (N, n_out)
  int n_out;                        // 1 or 3
  int exp_act;                      // 1: out = exp(x + out_bias) (trunc_exp forward); 0: out = x
  float out_bias;
  float out_sub;  // subtracted after the activation (the caller's `density - threshold`, sf3d/system.py:155); 0 = off
};
struct TgParams {
  TgHead head[kTgMaxHeads];
  int nheads, n_hidden;
  int nA, nB, nC;  // lattice extents: slow, mid, fast index of the vertex order
  int wait_ns;
};

struct TgTile {
  int head, a0, b0, c0;
};
__device__ __forceinline__ TgTile tg_tile(unsigned t, unsigned ntb, unsigned ntc, unsigned tiles_per_head) {
  TgTile g;
  g.head = (int)(t / tiles_per_head);
  unsigned rem = t - (unsigned)g.head * tiles_per_head;
  const unsigned ta = rem / (ntb * ntc);
  rem -= ta * ntb * ntc;
  const unsigned tb = rem / ntc;
  g.a0 = (int)ta * kTgA;
  g.b0 = (int)tb * kTgB;
  g.c0 = (int)(rem - tb * ntc) * kTgC;
  return g;
}

__global__ void __launch_bounds__(kTgWG * 128 + 128, 1) tetgrid_tc_kernel(TgParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wper = (nh - 1) * kWBytes;  // per head: hidden images, and as much again for the bias K-blocks
  unsigned char* sW = smem;                              // [head][l-1]
  unsigned char* sBB = sW + p.nheads * wper;             // [head][l-1]
  unsigned char* sS = sBB + p.nheads * wper;             // [wg][stage] 20 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + kTgWG * kTgStages * kTgStageBytes);
  // bars[0] = weights; per warpgroup g: [1+5g+s] full[s], [3+5g+s] empty[s], [5+5g] acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 5 * kTgWG);
  float* sHeadW = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);  // [head][4*64 + 4]

  const int tid_cta = threadIdx.x;
  const int wid = tid_cta >> 5;
  const int lane = tid_cta & 31;

  if (tid_cta == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int g = 0; g < kTgWG; ++g) {
      for (int st = 0; st < kTgStages; ++st) {
        mbar_init(smem_u32(&bars[1 + 5 * g + st]), 1);  // full: the producer's expect_tx arrival + the copies' bytes
        mbar_init(smem_u32(&bars[3 + 5 * g + st]), 4);  // empty: one elected lane per consumer warp
      }
      mbar_init(smem_u32(&bars[5 + 5 * g]), 1);  // acc_full: tcgen05.commit
    }
    mbar_fence_init();
  }
  for (int i = tid_cta; i < p.nheads * (4 * kHid + 4); i += blockDim.x) {
    const int h = i / (4 * kHid + 4);
    sHeadW[i] = p.head[h].head_w[i - h * (4 * kHid + 4)];
  }
  if (wid == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_w = smem_u32(&bars[0]);
  if (tid_cta == 0) {
    mbar_expect_tx(bar_w, (uint32_t)(2 * p.nheads * wper));
    for (int h = 0; h < p.nheads; ++h)
      for (int off = 0; off < wper; off += kWBytes) {
        bulk_g2s(smem_u32(sW + h * wper + off), p.head[h].tc_weights + off, (uint32_t)kWBytes, bar_w);
        bulk_g2s(smem_u32(sBB + h * wper + off), p.head[h].tc_biasblk + off, (uint32_t)kWBytes, bar_w);
      }
  }

  const unsigned nta = (unsigned)((p.nA + kTgA - 1) / kTgA), ntb = (unsigned)((p.nB + kTgB - 1) / kTgB);
  const unsigned ntc = (unsigned)((p.nC + kTgC - 1) / kTgC);
  const unsigned tiles_per_head = nta * ntb * ntc;
  const unsigned ntiles = tiles_per_head * (unsigned)p.nheads;

  if (wid >= kTgWG * 4) {
    // =================================================================== producer: one warp per warpgroup
    // Lanes 0..11 each issue one bulk copy per tile: the fast-index runs T1[a0+l][c0..], T2[b0+l][c0..] (contiguous rows,
    // <= 2 KB) and C[a0+l][b0..] (<= 1 KB) straight into the stage; indices behind the lattice's end are clamped (their
    // samples are never stored).  No registers, no per-thread load latency: the copies of two tiles are in flight.
    const int g = wid - kTgWG * 4;
    uint32_t par_empty = 3;  // bit s: parity to wait on; a fresh barrier passes a wait on parity 1
    unsigned k = 0;
    for (unsigned n = 0;; ++n, ++k) {
      const unsigned t = (n * gridDim.x + blockIdx.x) * kTgWG + g;
      if (t >= ntiles) break;
      const int st = (int)(k & 1u);
      const TgTile tg = tg_tile(t, ntb, ntc, tiles_per_head);
      const TgHead& H = p.head[tg.head];
      const uint32_t bar_full = smem_u32(&bars[1 + 5 * g + st]);
      mbar_wait_sleep(smem_u32(&bars[3 + 5 * g + st]), (par_empty >> st) & 1u, 20000u);
      par_empty ^= 1u << st;
      const int ncr = min(kTgC, p.nC - tg.c0), nbr = min(kTgB, p.nB - tg.b0);
      const uint32_t bytes = (uint32_t)((kTgA + kTgB) * ncr + kTgA * nbr) * kTgRow;
      unsigned char* stage = sS + (g * kTgStages + st) * kTgStageBytes;
      if (lane == 0) mbar_expect_tx(bar_full, bytes);
      __syncwarp();
      if (lane < kTgA) {
        const int a = min(tg.a0 + lane, p.nA - 1);
        bulk_g2s(smem_u32(stage + lane * kTgC * kTgRow), H.T1 + ((long long)a * p.nC + tg.c0) * kHid, (uint32_t)(ncr * kTgRow), bar_full);
      } else if (lane < kTgA + kTgB) {
        const int l = lane - kTgA, b = min(tg.b0 + l, p.nB - 1);
        bulk_g2s(smem_u32(stage + (kTgA + l) * kTgC * kTgRow), H.T2 + ((long long)b * p.nC + tg.c0) * kHid, (uint32_t)(ncr * kTgRow), bar_full);
      } else if (lane < 2 * kTgA + kTgB) {
        const int l = lane - kTgA - kTgB, a = min(tg.a0 + l, p.nA - 1);
        bulk_g2s(smem_u32(stage + ((kTgA + kTgB) * kTgC + l * kTgB) * kTgRow), H.C + ((long long)a * p.nB + tg.b0) * kHid,
                 (uint32_t)(nbr * kTgRow), bar_full);
      }
    }
  } else {
    // =================================================================== consumer
    const int wg = wid >> 2;
    const int q = wid & 3;
    const int m = (q << 5) | lane;
    const int la = m >> 5, lb = (m >> 3) & 3, lc = m & 7;  // the sample's place in the tile's 4 x 4 x 8 block
    const int tid_wg = tid_cta & 127;
    const uint32_t d_tmem = tmem_base + (uint32_t)(wg * 96);
    const uint32_t a_tmem = d_tmem + 64;
    const uint32_t ones_tmem = tmem_base + (uint32_t)(kTgWG * 96);
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t idesc_hidden = umma_idesc_f16_f32(128, 64);
    const uint32_t bar_acc = smem_u32(&bars[5 + 5 * wg]);
    uint32_t par_full = 0, par_acc = 0;
    unsigned k = 0;

    {  // constant activation block of the bias MMA: k = 64, 65 -> 1.0 (bias hi, lo rows), k = 66..79 -> 0
      const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(ones_tmem + lane_off, one);  // every warpgroup writes the same constants (idempotent)
    }
    mbar_wait(bar_w, 0);

    // hand the finished activation columns to the tensor core: hidden layer L = 1..nh-1 of head h
    auto issue_layer = [&](int h, int L) {
      tmem_st_wait();
      tc_fence_before();
      named_bar_sync(1 + wg, 128);
      if (tid_wg == 0) {
        tc_fence_after();
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sW + h * wper + (L - 1) * kWBytes));
#pragma unroll
        for (int kc = 0; kc < kHid / 16; ++kc)  // K = 16 per instruction = 8 packed columns of A, +32 B of B
          umma_f16_ts(d_tmem, a_tmem + 8 * kc, b_desc + 2 * kc, idesc_hidden, kc > 0 ? 1u : 0u);
        umma_f16_ts(d_tmem, ones_tmem, umma_desc_k_sw128(smem_u32(sBB + h * wper + (L - 1) * kWBytes)), idesc_hidden, 1u);
        umma_commit(bar_acc);
      }
    };

    for (unsigned n = 0;; ++n, ++k) {
      const unsigned t = (n * gridDim.x + blockIdx.x) * kTgWG + wg;
      if (t >= ntiles) break;
      const int st = (int)(k & 1u);
      const TgTile tg = tg_tile(t, ntb, ntc, tiles_per_head);
      const TgHead& H = p.head[tg.head];
      {
        // ---- layer 0: C[a][b] + T1[a][c] + T2[b][c] from the stage -> SiLU -> activation columns ------------------
        const unsigned char* stage = sS + (wg * kTgStages + st) * kTgStageBytes;
        const float4* t1 = reinterpret_cast<const float4*>(stage + (la * kTgC + lc) * kTgRow);
        const float4* t2 = reinterpret_cast<const float4*>(stage + ((kTgA + lb) * kTgC + lc) * kTgRow);
        const float4* cc = reinterpret_cast<const float4*>(stage + ((kTgA + kTgB) * kTgC + la * kTgB + lb) * kTgRow);
        mbar_wait_sleep(smem_u32(&bars[1 + 5 * wg + st]), (par_full >> st) & 1u, (uint32_t)p.wait_ns);
        par_full ^= 1u << st;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 4 chunks of 16 columns
          uint32_t pk[8];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int ch = 4 * c + g4;
            const float4 x = t1[ch ^ lc], y = t2[ch ^ lc], z = cc[ch];  // rows of the fast index are stored chunk-swizzled
            // packed fp32 pairs (FADD2 / FFMA2): the kernel is short of issue slots, not of FMA-pipe time
            const float2 lo = silu2_from_half_arg(__fadd2_rn(__fadd2_rn(make_float2(x.x, x.y), make_float2(y.x, y.y)), make_float2(z.x, z.y)));
            const float2 hi = silu2_from_half_arg(__fadd2_rn(__fadd2_rn(make_float2(x.z, x.w), make_float2(y.z, y.w)), make_float2(z.z, z.w)));
            pk[2 * g4 + 0] = pack_half2(lo.x, lo.y);
            pk[2 * g4 + 1] = pack_half2(hi.x, hi.y);
          }
          tmem_st8(a_tmem + lane_off + 8 * c, pk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[3 + 5 * wg + st]));
        issue_layer(tg.head, 1);
      }
      for (int l = 1; l < nh; ++l) {
        mbar_wait_sleep(bar_acc, par_acc, (uint32_t)p.wait_ns);
        par_acc ^= 1u;
        tc_fence_after();
        const bool last = l == nh - 1;
        const float* hwf = sHeadW + tg.head * (4 * kHid + 4);
        const float4* hw = reinterpret_cast<const float4*>(hwf);
        const bool wide = H.n_out > 1;  // warp-uniform
        float2 d0 = make_float2(0.f, 0.f), d1 = d0, d2 = d0;  // even / odd partial sums of the head's dot products
        uint32_t r[2][16];
        tmem_ld16(d_tmem + lane_off, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c + 1 < 4) tmem_ld16(d_tmem + lane_off + (c + 1) * 16, r[(c + 1) & 1]);
          const uint32_t* rc = r[c & 1];
          float2 h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = silu2_from_half_arg(make_float2(__uint_as_float(rc[2 * i]), __uint_as_float(rc[2 * i + 1])));
          if (!last) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_half2(h[i].x, h[i].y);
            tmem_st8(a_tmem + lane_off + 8 * c, pk);
          } else {
            // the head (network.py:176-178: last Linear of the head) as fp32 dot products on the activations in registers
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 w = hw[4 * c + i];
              d0 = __ffma2_rn(h[2 * i], make_float2(w.x, w.y), d0);
              d0 = __ffma2_rn(h[2 * i + 1], make_float2(w.z, w.w), d0);
            }
            if (wide) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 w1 = hw[16 + 4 * c + i], w2 = hw[32 + 4 * c + i];
                d1 = __ffma2_rn(h[2 * i], make_float2(w1.x, w1.y), d1);
                d1 = __ffma2_rn(h[2 * i + 1], make_float2(w1.z, w1.w), d1);
                d2 = __ffma2_rn(h[2 * i], make_float2(w2.x, w2.y), d2);
                d2 = __ffma2_rn(h[2 * i + 1], make_float2(w2.z, w2.w), d2);
              }
            }
          }
        }
        if (!last) {
          issue_layer(tg.head, l + 1);
        } else if (tg.a0 + la < p.nA && tg.b0 + lb < p.nB && tg.c0 + lc < p.nC) {
          const long long row = ((long long)(tg.a0 + la) * p.nB + tg.b0 + lb) * p.nC + tg.c0 + lc;
          const float* hb = hwf + 4 * kHid;
          float o0 = d0.x + d0.y + hb[0];
          if (H.exp_act) o0 = __fsub_rn(expf(__fadd_rn(o0, H.out_bias)), H.out_sub);
          if (!wide) {
            H.out[row] = o0;
          } else {
            float* o = H.out + row * H.n_out;
            o[0] = o0;
            o[1] = d1.x + d1.y + hb[1];
            if (H.n_out > 2) o[2] = d2.x + d2.y + hb[2];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------- tables
struct TgTabParams {
  const float* planes_cl;  // (3, H, W, 40) fp32 channels-last
  int H, W, align_corners;
  const float* axis_u[3];  // normalised coordinate in (-1, 1) of the lattice's slow / mid / fast index
  int n[3];                // extents of the three lattice indices
  int sdim[3];             // spatial dimension (0 x, 1 y, 2 z) each lattice index runs along
  int nheads;
  const float* w0_half[kTgMaxHeads];  // (64, 120) fp32 = W_0 / 2
  const float* b0_half[kTgMaxHeads];  // (64) fp32 = b_0 / 2
  float* tab[kTgMaxHeads][3];         // C (slow, mid), T1 (slow, fast), T2 (mid, fast): (n_p, n_q, 64) fp32
};

// blockIdx.y = pair.  A warp works on 8 consecutive table entries at a time, for all heads at once: lanes 0..7 set up the
// bilinear taps of one entry each (the reference's grid_sample with zero padding, system.py:186-195), the warp gathers
// the 8 x 40 channel values (160-byte channels-last runs) into shared memory, then every lane contracts them with two
// rows of W_0/2 per head (shared memory, transposed so that consecutive lanes read consecutive words; the eight
// entries share every weight it loads) and stores its 2 x 8 outputs per head.
constexpr int kTabE = 8;  // entries per warp iteration
template <int kNH>
__global__ void __launch_bounds__(256) tetgrid_tables_kernel(TgTabParams p) {
  __shared__ float sWt[kNH][kCp][kHid];
  __shared__ __align__(16) float sF[8][kCp][kTabE];
  __shared__ int sOff[8][kTabE][4];
  __shared__ float sWgt[8][kTabE][4];
  const int pair = blockIdx.y;
  const int lp = pair == 2 ? 1 : 0, lq = pair == 0 ? 1 : 2;  // the pair's two lattice indices
  const int dp = p.sdim[lp], dq = p.sdim[lq];
  const int plane = dp + dq - 1;  // {x,y} -> 0, {x,z} -> 1, {y,z} -> 2 (system.py:181-184)
  // the plane's first coordinate (grid_sample's x = width) is the lower spatial dimension
  const bool p_is_u = dp < dq;
  for (int i = threadIdx.x; i < kNH * kCp * kHid; i += blockDim.x) {
    const int h = i / (kCp * kHid), r = i - h * (kCp * kHid);
    const int c = r / kHid, n = r - c * kHid;
    sWt[h][c][n] = p.w0_half[h][n * kFeat + plane * kCp + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = p.n[lp], nq = p.n[lq];
  const long long nent = (long long)np * nq;
  const float* P = p.planes_cl + (long long)plane * p.H * p.W * kCp;
  float bias[kNH][2];
#pragma unroll
  for (int h = 0; h < kNH; ++h) {
    bias[h][0] = pair == 0 ? p.b0_half[h][lane] : 0.0f;
    bias[h][1] = pair == 0 ? p.b0_half[h][lane + 32] : 0.0f;
  }
  for (long long e0 = ((long long)blockIdx.x * 8 + warp) * kTabE; e0 < nent; e0 += (long long)gridDim.x * 8 * kTabE) {
    if (lane < kTabE) {
      const long long e = min(e0 + lane, nent - 1);
      const int ip = (int)(e / nq), iq = (int)(e - (long long)ip * nq);
      const float up = p.axis_u[lp][ip], uq = p.axis_u[lq][iq];
      const Tap2 tu = make_tap(p_is_u ? up : uq, p.W, p.align_corners);
      const Tap2 tv = make_tap(p_is_u ? uq : up, p.H, p.align_corners);
      sOff[warp][lane][0] = (tv.i0 * p.W + tu.i0) * kCp;
      sOff[warp][lane][1] = (tv.i0 * p.W + tu.i1) * kCp;
      sOff[warp][lane][2] = (tv.i1 * p.W + tu.i0) * kCp;
      sOff[warp][lane][3] = (tv.i1 * p.W + tu.i1) * kCp;
      sWgt[warp][lane][0] = tv.w0;
      sWgt[warp][lane][1] = tv.w1;
      sWgt[warp][lane][2] = tu.w0;
      sWgt[warp][lane][3] = tu.w1;
    }
    __syncwarp();
    for (int idx = lane; idx < kTabE * kCp; idx += 32) {
      const int en = idx / kCp, c = idx - en * kCp;
      const int* o = sOff[warp][en];
      const float* w = sWgt[warp][en];
      const float v00 = __ldg(P + o[0] + c), v01 = __ldg(P + o[1] + c), v10 = __ldg(P + o[2] + c), v11 = __ldg(P + o[3] + c);
      sF[warp][c][en] = w[0] * (w[2] * v00 + w[3] * v01) + w[1] * (w[2] * v10 + w[3] * v11);
    }
    __syncwarp();
    float acc[kNH][2][kTabE];
#pragma unroll
    for (int h = 0; h < kNH; ++h)
#pragma unroll
      for (int en = 0; en < kTabE; ++en) {
        acc[h][0][en] = bias[h][0];
        acc[h][1][en] = bias[h][1];
      }
#pragma unroll 4
    for (int c = 0; c < kCp; ++c) {
      const float4 fa = *reinterpret_cast<const float4*>(&sF[warp][c][0]), fb = *reinterpret_cast<const float4*>(&sF[warp][c][4]);
      const float f[kTabE] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int h = 0; h < kNH; ++h) {
        const float w0 = sWt[h][c][lane], w1 = sWt[h][c][lane + 32];
#pragma unroll
        for (int en = 0; en < kTabE; ++en) {
          acc[h][0][en] = fmaf(w0, f[en], acc[h][0][en]);
          acc[h][1][en] = fmaf(w1, f[en], acc[h][1][en]);
        }
      }
    }
#pragma unroll
    for (int en = 0; en < kTabE; ++en) {
      const long long e = e0 + en;
      if (e < nent) {
        // rows indexed by the fast lattice index (T1, T2) are stored with their 16-byte chunks at chunk ^ (c & 7): the kernel
        // above bulk-copies 8 consecutive rows into shared memory and reads them one row per thread
        const int iq = (int)(e % nq);
        const int sw = pair == 0 ? 0 : ((iq & 7) << 2);
#pragma unroll
        for (int h = 0; h < kNH; ++h) {
          float* out = p.tab[h][pair];
          out[e * kHid + (lane ^ sw)] = acc[h][0][en];
          out[e * kHid + ((lane + 32) ^ sw)] = acc[h][1][en];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace smb

using namespace smb;

// See include/sculptmate_b200.h.
extern "C" int smb_query_tetgrid_tc(const float* planes_cl, int Hp, int Wp, int align_corners, int nheads,
                                    const void* const* decoder_blobs, const smb_decoder_layout* const* layouts, const int* n_out,
                                    const int* exp_act, const float* out_bias, const float* out_sub, const float* const* axis_u,
                                    const int* extents, const int* spatial_dim, float* const* outs, void* stream) {
  if (!planes_cl || !decoder_blobs || !layouts || !n_out || !exp_act || !out_bias || !axis_u || !extents || !spatial_dim || !outs)
    return SMB_ERR_BAD_ARG;
  if (nheads < 1 || nheads > kTgMaxHeads || Hp < 1 || Wp < 1) return SMB_ERR_BAD_ARG;
  int seen = 0;
  for (int a = 0; a < 3; ++a) {
    if (extents[a] < 1 || spatial_dim[a] < 0 || spatial_dim[a] > 2 || !axis_u[a]) return SMB_ERR_BAD_ARG;
    seen |= 1 << spatial_dim[a];
  }
  if (seen != 7) return SMB_ERR_BAD_ARG;
  const int nh = (int)layouts[0]->n_hidden;
  for (int h = 0; h < nheads; ++h) {
    if (!decoder_blobs[h] || !layouts[h] || !outs[h] || (int)layouts[h]->n_hidden != nh) return SMB_ERR_BAD_ARG;
    if (n_out[h] < 1 || n_out[h] > 3) return SMB_ERR_BAD_ARG;
  }
  if (nh < 2 || nh > kMaxHidden) return SMB_ERR_BAD_ARG;
  const long long nA = extents[0], nB = extents[1], nC = extents[2];
  if (nB * nC >= (1LL << 30) || nC < 2) return SMB_ERR_BAD_ARG;
  const long long tiles = (long long)nheads * ((nA + kTgA - 1) / kTgA) * ((nB + kTgB - 1) / kTgB) * ((nC + kTgC - 1) / kTgC);
  if (tiles >= (1LL << 31) / kTgWG) return SMB_ERR_BAD_ARG;  // 32-bit tile arithmetic in the kernel
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  // scratch for the tables: stream-ordered (the pool keeps the block between calls, see lattice_api.cu)
  keep_async_scratch(dev);
  const size_t per_head = (size_t)(nA * nB + nA * nC + nB * nC) * kHid;
  float* tabs = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&tabs), per_head * nheads * sizeof(float), st) != cudaSuccess) return SMB_ERR_CUDA;

  TgTabParams tp{};
  tp.planes_cl = planes_cl;
  tp.H = Hp;
  tp.W = Wp;
  tp.align_corners = align_corners;
  TgParams p{};
  p.nheads = nheads;
  p.n_hidden = nh;
  p.nA = (int)nA;
  p.nB = (int)nB;
  p.nC = (int)nC;
  p.wait_ns = 200;
  for (int a = 0; a < 3; ++a) {
    tp.axis_u[a] = axis_u[a];
    tp.n[a] = extents[a];
    tp.sdim[a] = spatial_dim[a];
  }
  tp.nheads = nheads;
  for (int h = 0; h < nheads; ++h) {
    const unsigned char* blob = static_cast<const unsigned char*>(decoder_blobs[h]);
    const smb_decoder_layout* L = layouts[h];
    tp.w0_half[h] = reinterpret_cast<const float*>(blob + L->off_w0_half);
    tp.b0_half[h] = reinterpret_cast<const float*>(blob + L->off_bias_half);
    float* base = tabs + per_head * h;
    tp.tab[h][0] = base;
    tp.tab[h][1] = base + (size_t)nA * nB * kHid;
    tp.tab[h][2] = base + (size_t)(nA * nB + nA * nC) * kHid;
    TgHead& H = p.head[h];
    H.C = tp.tab[h][0];
    H.T1 = tp.tab[h][1];
    H.T2 = tp.tab[h][2];
    H.tc_weights = blob + L->off_tc_hidden;
    H.tc_biasblk = blob + L->off_tc_biasblk;
    // plain fp32 copy of the parameters: [W0 (64x120) b0 (64)] [(W_l (64x64) b_l (64)) x (nh-1)] [W_L (4x64) b_L (4)]
    H.head_w = reinterpret_cast<const float*>(blob + L->off_f32) + (kHid * kFeat + kHid) + (size_t)(nh - 1) * (kHid * kHid + kHid);
    H.out = outs[h];
    H.n_out = n_out[h];
    H.exp_act = exp_act[h];
    H.out_bias = out_bias[h];
    H.out_sub = out_sub ? out_sub[h] : 0.0f;
  }
  {
    const long long maxent = nA * nB > nA * nC ? (nA * nB > nB * nC ? nA * nB : nB * nC) : (nA * nC > nB * nC ? nA * nC : nB * nC);
    long long bx = (maxent + 8 * kTabE - 1) / (8 * kTabE);
    if (bx > (long long)sms * 4) bx = (long long)sms * 4;
    if (nheads == 1) tetgrid_tables_kernel<1><<<dim3((unsigned)bx, 3), 256, 0, st>>>(tp);
    else tetgrid_tables_kernel<2><<<dim3((unsigned)bx, 3), 256, 0, st>>>(tp);
  }
  const size_t smem = (size_t)2 * nheads * (nh - 1) * kWBytes + (size_t)kTgWG * kTgStages * kTgStageBytes + 8 * (1 + 5 * kTgWG) + 16 + 16 +
                      (size_t)nheads * (4 * kHid + 4) * 4;
  int rc = SMB_OK;
  if (smem > 227 * 1024) {
    rc = SMB_ERR_BAD_ARG;
  } else {
    static const cudaError_t attr = cudaFuncSetAttribute(tetgrid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr != cudaSuccess) {
      rc = SMB_ERR_CUDA;
    } else {
      long long grid = (tiles + kTgWG - 1) / kTgWG;
      if (grid > sms) grid = sms;
      tetgrid_tc_kernel<<<(unsigned)grid, kTgWG * 128 + 128, smem, st>>>(p);
      rc = smb_check(cudaGetLastError());
    }
  }
  cudaFreeAsync(tabs, st);
  return rc;
}

// K1 experiment, round 2: lattice field kernel with TILE PAIRS SHARING ONE ACCUMULATOR.  NOT the default: measured 3.04 ms
// at 256^3 against 2.89 ms of field_tc_ta.cu; kept (developer build) with its instrumentation because DESIGN.md argues
// from these measurements.
//
// Path (reference): TSR.extract_mesh's density query -- F.grid_sample x3 + NeRFMLP over the R^3 lattice
// (/root/reference/TripoSR/tsr/models/nerf_renderer.py:41-91, tsr/models/network_utils.py:35-124).
//
// Same math as field_tc_ta.cu (tile = 128 z-samples of one lattice line, layer 0 from the projected-plane table,
// hidden layers as tcgen05.mma with the fp16 activations as the A operand IN TENSOR MEMORY, SiLU = h + h*tanh(h)),
// different hand-off protocol.  In field_tc_ta.cu every consumer warp owns one tile and, per layer, blocks on
// tcgen05.wait::st + a 128-thread barrier + the MMA round trip: a third of each step it issues no SFU work, and
// five warps per SM sub-partition cannot cover that (XU pipe 79 %, profiles/r01o).  Tensor memory held five such
// tiles (64 accumulator + 32 activation columns).  Here:
//
//   * a consumer warpgroup works on a PAIR of tiles (X, Y) that SHARE one 64-column accumulator D:
//       per warpgroup  D (64 cols) | A_X (32) | A_Y (32)  = 128 columns, 4 warpgroups = all 512 columns,
//     i.e. EIGHT tiles in flight per SM instead of five.  While the warpgroup runs the epilogue of X (reads D,
//     writes A_X), the tensor core computes Y's layer into the half of D the epilogue has already read:
//     each layer is issued as two N = 32 halves (lo: D columns 0-31, hi: 32-63), and the epilogue releases
//     each half (mbarrier arrival) as soon as its tcgen05.ld has completed.
//   * consumers NEVER block on their partners: all hand-offs are mbarrier ARRIVALS (one elected lane per warp);
//     a dedicated issuer warp per warpgroup (one thread) waits for them and issues the MMAs.  The only waits
//     of a consumer warp are on accumulator halves that were issued at least two 16-column chunks of
//     epilogue work earlier.
//   * 16 consumer warps (4 per SM sub-partition) are therefore always inside an activation stretch; the
//     stand-alone epilogue stream reaches 93 % of the MUFU peak with that many warps
//     (tools/microbench/epilogue.cu, profiles/r02a_micro.log).
//   * kPolyPairs of the 8 activation pairs of every 16-column chunk take their tanh from a packed-half
//     polynomial on the FMA pipe (silu_poly_h2) instead of the SFU: with every warp busy the kernel is
//     SFU-throughput-bound, so moving a quarter of the activations off the SFU pays (it did not in
//     field_tc_ta.cu, where the warps were latency-bound).
//
// Per warpgroup g the mbarriers are (index base 1 + 12 g):
//   t_full[s], t_empty[s]  producer <-> consumers: layer-0 table of tile s in {X, Y}
//   a_ready[s]             consumers -> issuer: activation columns A_s written (tcgen05.wait::st done), 4 arrivals
//   d_free[h]              consumers -> issuer: half h of D has been read into registers, 4 arrivals
//   acc[s][h]              issuer -> consumers: tcgen05.commit after the MMAs of half h of tile s
#ifdef SMB_DEV_VARIANTS  // developer build only: an experiment that did not beat field_tc_ta.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "field_tc_common.cuh"

namespace smb {

constexpr int kPairWG = 4;                                 // consumer warpgroups (tile pairs in flight)
constexpr int kPairThreads = kPairWG * 128 + 128 + 128;    // + issuer warpgroup + producer warpgroup
constexpr int kBarsPerWG = 12;
// barrier indices inside a warpgroup's block of 12
constexpr int kTFull = 0, kTEmpty = 2, kAReady = 4, kDFree = 6, kAcc = 8;

// developer aid: a wait that times out records (code, block, warp) in pinned host memory before it traps
__device__ unsigned long long* g_pair_dbg = nullptr;
__device__ __forceinline__ void pair_wait(uint32_t bar, uint32_t parity, uint32_t ns, uint32_t code) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, ns ? ns : 1000u)) {
    if (++spins > (1u << 20)) {
      if (g_pair_dbg && (threadIdx.x & 31) == 0) {
        g_pair_dbg[blockIdx.x * 32 + (threadIdx.x >> 5)] = ((unsigned long long)code << 32) | (blockIdx.x << 8) | (threadIdx.x >> 5) | 0x80000000ull;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// silu(2h) for two activations entirely in packed fp16 on the FMA pipe (no MUFU):
//   tanh(|h|) ~= q(x), x = min(|h| / 2 - 1, 1) in [-1, 1]  (|h| clamped at 4: 1 - tanh(4) = 6.7e-4);
//   silu(2h) = h + |h| tanh(|h|) = fma(|h|, q(x), h).
// q: degree-7 minimax polynomial of tanh(2(x + 1)) on [-1, 1] in the CENTRED variable (all coefficients
// below 1 in magnitude, so fp16 Horner does not cancel; in the monomial basis of |h| or h^2 the coefficients
// reach 150 and fp16 evaluation is useless).  Max abs error of tanh under fp16 evaluation 1.3e-3
// (tools/fit_tanh_poly.py half).  12 FMA/ALU-pipe instructions per PAIR of activations, result already packed.
struct PolyH2 {
  __half2 c[8];
  __half2 half_, mone, one;
};
__device__ __forceinline__ PolyH2 make_poly_h2() {
  PolyH2 p;
  const float c[8] = {0.9638671875f, 0.140380859375f, -0.25830078125f, 0.345703125f, -0.37353515625f, 0.1431884765625f, 0.16748046875f, -0.1297607421875f};
#pragma unroll
  for (int i = 0; i < 8; ++i) p.c[i] = __float2half2_rn(c[i]);
  p.half_ = __float2half2_rn(0.5f);
  p.mone = __float2half2_rn(-1.0f);
  p.one = __float2half2_rn(1.0f);
  return p;
}
__device__ __forceinline__ uint32_t silu_poly_h2(float h0, float h1, const PolyH2& P) {
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 a = __habs2(h);
  const __half2 x = __hmin2(__hfma2(a, P.half_, P.mone), P.one);
  __half2 q = __hfma2(P.c[7], x, P.c[6]);
#pragma unroll
  for (int i = 5; i >= 0; --i) q = __hfma2(q, x, P.c[i]);
  const __half2 r = __hfma2(a, q, h);
  return *reinterpret_cast<const uint32_t*>(&r);
}

// 16 accumulator values (+ bias) -> 8 packed fp16 activation pairs; kPolyPairs of them on the FMA pipe
template <int kPolyPairs>
__device__ __forceinline__ void activate16(const float (&h)[16], uint32_t (&pk)[8], const PolyH2& P) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // spread the polynomial pairs evenly over the chunk so that FMA-pipe and SFU work interleave
    const bool poly = kPolyPairs > 0 && ((i * kPolyPairs) & 7) < kPolyPairs;
    if (poly) pk[i] = silu_poly_h2(h[2 * i], h[2 * i + 1], P);
    else pk[i] = pack_half2(silu_from_half_arg(h[2 * i]), silu_from_half_arg(h[2 * i + 1]));
  }
}

__device__ unsigned int g_pair_prof[148 * 16 * 16 + 148 * 4 * 4];
// kProf: event log of block 0 (20 roles x 512 events x (id, clock)): tools/k1_timeline.py
constexpr int kEvtMax = 512;
__device__ unsigned int g_pair_evt[20 * kEvtMax * 2];

template <int kPolyPairs, bool kProf = false>
__global__ void __launch_bounds__(kPairThreads, 1) lattice_tc_pair_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wbytes = tc_weight_bytes(nh);
  unsigned char* sW = smem;
  const float* sBias = reinterpret_cast<const float*>(sW + (nh - 1) * kWBytes + kWFinalBytes);
  const float* sBiasF = sBias + nh * kHid;
  unsigned char* tables = smem + ((wbytes + 1023) / 1024) * 1024;
  const int tbytes = ta_table_bytes(p.trows);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tables + 2 * kPairWG * tbytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kBarsPerWG * kPairWG);
  // per (z-segment of a line, sample m): layer-0 interpolation weight w1 and table row r0 -- they depend only on the
  // sample's z index, so they are computed once per launch instead of once per tile
  float2* sGeo = reinterpret_cast<float2*>(tmem_slot + 4);

  const int tid_cta = threadIdx.x;
  const int wid = tid_cta >> 5;
  const int lane = tid_cta & 31;

  if (tid_cta == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int g = 0; g < kPairWG; ++g) {
      uint64_t* b = bars + 1 + kBarsPerWG * g;
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&b[kTFull + s]), 4);   // one lane per producer warp (each builds a quarter of the rows)
        mbar_init(smem_u32(&b[kTEmpty + s]), 4);  // one lane per consumer warp
        mbar_init(smem_u32(&b[kAReady + s]), 4);
        mbar_init(smem_u32(&b[kDFree + s]), 4);   // s = half
        mbar_init(smem_u32(&b[kAcc + 2 * s]), 1);      // tcgen05.commit, lo half of tile s
        mbar_init(smem_u32(&b[kAcc + 2 * s + 1]), 1);  // hi half
      }
    }
    mbar_fence_init();
  }
  const int tiles_per_line = (p.R + kTileM - 1) / kTileM;
  for (int i = tid_cta; i < tiles_per_line * kTileM; i += kPairThreads) {
    const int k0 = (i / kTileM) * kTileM;
    const float fz = unnormalize(p.axis_u[min(i, p.R - 1)], p.H, p.align_corners);
    const float hf = floorf(fz);
    const int hlo = (int)floorf(unnormalize(p.axis_u[k0], p.H, p.align_corners));
    const int r0 = min(max((int)hf - hlo, 0), p.trows - 2);
    sGeo[i] = make_float2(__fsub_rn(fz, hf), __int_as_float(r0));
  }
  if (wid == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_w = smem_u32(&bars[0]);
  if (tid_cta == 0) {
    mbar_expect_tx(bar_w, (uint32_t)wbytes);
    int off = 0;
    while (off < wbytes) {
      const int n = min(8192, wbytes - off);
      bulk_g2s(smem_u32(sW + off), p.tc_weights + off, (uint32_t)n, bar_w);
      off += n;
    }
  }

  // pair P = (n * gridDim.x + blockIdx.x) * kPairWG + g (n = 0, 1, ...) of warpgroup g holds tiles 2P and 2P + 1
  const int ntiles = p.nx * p.R * tiles_per_line;  // launch_pair_n rejects shapes beyond 2^31 tiles
  const int g = (wid < kPairWG * 4) ? (wid >> 2) : (wid & 3);
  const int pair0 = blockIdx.x * kPairWG + g, pair_stride = gridDim.x * kPairWG;
  const int npairs_all = (ntiles + 1) >> 1;
  const int my_pairs = pair0 < npairs_all ? (npairs_all - 1 - pair0) / pair_stride + 1 : 0;
  // the very last pair of the launch has one tile when ntiles is odd
  const bool last_single = (ntiles & 1) && my_pairs > 0 && pair0 + (my_pairs - 1) * pair_stride == npairs_all - 1;
  const uint32_t bb = smem_u32(bars + 1 + kBarsPerWG * g);  // this warpgroup's barriers
  auto bar = [&](int idx) { return bb + 8u * (uint32_t)idx; };

  if (wid >= kPairWG * 4 + 4) {
    // =================================================================== producers
    // All four producer warps build EVERY table together (warp w the rows w, w+4, ... in pairs), visiting the
    // warpgroups in a fixed order: the load of a table build is then the same on the four SM sub-partitions.
    // One warp per warpgroup (the first version) slowed sub-partition g while warpgroup g's table was built,
    // and with it lane quarter g of every tile -- the skew the consumers then wait out at their accumulators.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int pw = wid - (kPairWG * 4 + 4);
    uint32_t par_empty = 0xffu;  // bit 2 gg + s: parity to wait on; a fresh barrier passes a wait on parity 1
    const int max_pairs = pair0 - g < npairs_all ? (npairs_all - 1 - (pair0 - g)) / pair_stride + 1 : 0;  // warpgroup 0 has the most
    for (int n = 0; n < max_pairs; ++n) {
#pragma unroll 1
      for (int gg = 0; gg < kPairWG; ++gg) {
        const int pr = pair0 - g + gg + n * pair_stride;
        if (pr >= npairs_all) break;
        const int nvalid = (2 * pr + 1 < ntiles) ? 2 : 1;
        const uint32_t bbg = smem_u32(bars + 1 + kBarsPerWG * gg);
#pragma unroll 1
        for (int s = 0; s < nvalid; ++s) {
          float* sT = reinterpret_cast<float*>(tables + (2 * gg + s) * tbytes);
          const TableGeom G = table_geom(p, 2 * pr + s, tiles_per_line);
          pair_wait(bbg + 8u * (kTEmpty + s), (par_empty >> (2 * gg + s)) & 1u, 20000u, 0x100u + s);
          par_empty ^= 1u << (2 * gg + s);
          build_table(p, G, sT, lane, pw, 4);
          __syncwarp();
          if (lane == 0) mbar_arrive(bbg + 8u * (kTFull + s));
        }
      }
    }
  } else if (wid >= kPairWG * 4) {
    // =================================================================== MMA issuers (one thread per warpgroup)
    // register budget: the CTA is launched at 80 x 768; issuers keep 24 and producers 40, which frees
    // 128 x 56 + 128 x 40 = 12288 = 512 x 24 registers: exactly what lifts the 16 consumer warps to 104
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(g * 128);
      const uint32_t idesc_half = umma_idesc_f16_f32(128, 32);
      const uint32_t idesc_head = umma_idesc_f16_f32(128, 16);
      const uint32_t wait_ns = (uint32_t)p.wait_ns;
      const uint32_t w_base = smem_u32(sW);
      uint32_t par = 0xCu;  // bits 0,1: a_ready[s]; bits 2,3: d_free[h] (first wait on each passes: nobody has used D yet)
      unsigned int prof_par = 0, prof_lat[2] = {0, 0}, prof_cnt = 0;
      int evn = 0;
      auto ev = [&](unsigned int id) {
        if (kProf && blockIdx.x == 0 && evn < kEvtMax) {
          unsigned int* o = g_pair_evt + ((16 + g) * kEvtMax + evn) * 2;
          o[0] = id;
          o[1] = clock();
          ++evn;
        }
      };
      mbar_wait(bar_w, 0);
#pragma unroll 1
      for (int n = 0; n < my_pairs; ++n) {
        const int nvalid = (last_single && n == my_pairs - 1) ? 1 : 2;
        uint32_t w_addr = w_base;  // hidden images, then the head image, contiguous
#pragma unroll 1
        for (int e = 1; e <= nh; ++e, w_addr += kWBytes) {
#pragma unroll 1
          for (int s = 0; s < nvalid; ++s) {
            const uint32_t a_tmem = d_tmem + 64 + 32 * s;
            pair_wait(bar(kAReady + s), (par >> s) & 1u, wait_ns, 0x200u + s);
            par ^= 1u << s;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              pair_wait(bar(kDFree + h), (par >> (2 + h)) & 1u, wait_ns, 0x300u + 2 * s + h);
              par ^= 4u << h;
              tc_fence_after();
              if (e < nh) {
                // weight rows 32h .. 32h+31 = four 8-row swizzle atoms of 1024 B
                const uint64_t b_desc = umma_desc_k_sw128(w_addr + 4096u * h);
#pragma unroll
                for (int kc = 0; kc < kHid / 16; ++kc)
                  umma_f16_ts(d_tmem + 32 * h, a_tmem + 8 * kc, b_desc + 2 * kc, idesc_half, kc > 0 ? 1u : 0u);
              } else if (h == 0) {
                const uint64_t b_desc = umma_desc_k_sw128(w_addr);
#pragma unroll
                for (int kc = 0; kc < kHid / 16; ++kc) umma_f16_ts(d_tmem, a_tmem + 8 * kc, b_desc + 2 * kc, idesc_head, kc > 0 ? 1u : 0u);
              }
              umma_commit(bar(kAcc + 2 * s + h));
              ev(0x10000u + (e << 8) + (s << 4) + h);  // MMA of (e, s, h) issued + committed
            }
          }
        }
      }
      if (kProf) {
        unsigned int* o = g_pair_prof + 148 * 16 * 16 + (blockIdx.x * 4 + g) * 4;
        o[0] = prof_lat[0];
        o[1] = prof_lat[1];
        o[2] = prof_cnt;
      }
    }
  } else {
    // =================================================================== consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = wid & 3;
    const int m = (q << 5) | lane;
    const uint32_t d_tmem = tmem_base + (uint32_t)(g * 128) + ((uint32_t)(q * 32) << 16);
    const uint32_t wait_ns = (uint32_t)p.wait_ns;
    const PolyH2 P = make_poly_h2();
    uint32_t par = 0u;  // bits 0,1: t_full[s]; bits 2..5: acc[s][h] at bit 2 + 2s + h
    mbar_wait(bar_w, 0);  // biases live beside the weights
    unsigned int pc[13];  // kProf: clock() accumulators per phase (developer build only)
    unsigned int pt = 0;
    if (kProf) {
#pragma unroll
      for (int i = 0; i < 13; ++i) pc[i] = 0;
      pt = clock();
      pc[0] = pt;
    }
    auto lap = [&](int i) {
      if (kProf) {
        const unsigned int t = clock();
        pc[i] += t - pt;
        pt = t;
      }
    };
    int evn = 0;
    auto ev = [&](unsigned int id) {
      if (kProf && blockIdx.x == 0 && lane == 0 && evn < kEvtMax) {
        unsigned int* o = g_pair_evt + (wid * kEvtMax + evn) * 2;
        o[0] = id;
        o[1] = clock();
        ++evn;
      }
    };

#pragma unroll 1
    for (int n = 0; n < my_pairs; ++n) {
      const int t0 = 2 * (pair0 + n * pair_stride);
      const int nvalid = (last_single && n == my_pairs - 1) ? 1 : 2;

      // ---- layer 0 of both tiles from the producer's tables -> activation columns --------------------
#pragma unroll 1
      for (int s = 0; s < nvalid; ++s) {
        const float* sT = reinterpret_cast<const float*>(tables + (2 * g + s) * tbytes);
        const float* sC = sT + p.trows * kTPitch;
        const uint32_t a_tmem = d_tmem + 64 + 32 * s;
        const int seg = (t0 + s) % tiles_per_line;
        const float2 ge = sGeo[seg * kTileM + m];
        const float w1 = ge.x;
        const float w0 = __fsub_rn(1.0f, w1);
        const int r0 = __float_as_int(ge.y);
        lap(12);
        pair_wait(bar(kTFull + s), (par >> s) & 1u, wait_ns, 0x400u + s);
        par ^= 1u << s;
        lap(1);
        const float4* t0p = reinterpret_cast<const float4*>(sT + r0 * kTPitch);
        const float4* t1p = reinterpret_cast<const float4*>(sT + (r0 + 1) * kTPitch);
        const float4* cc = reinterpret_cast<const float4*>(sC);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 4 chunks of 16 columns
          float h[16];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 a = t0p[4 * c + g4], bq = t1p[4 * c + g4], cv = cc[4 * c + g4];
            h[4 * g4 + 0] = cv.x + w0 * a.x + w1 * bq.x;
            h[4 * g4 + 1] = cv.y + w0 * a.y + w1 * bq.y;
            h[4 * g4 + 2] = cv.z + w0 * a.z + w1 * bq.z;
            h[4 * g4 + 3] = cv.w + w0 * a.w + w1 * bq.w;
          }
          uint32_t pk[8];
          activate16<kPolyPairs>(h, pk, P);
          tmem_st8(a_tmem + 8 * c, pk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kTEmpty + s));  // the table has been consumed
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kAReady + s));
        lap(2);
      }

      // ---- hidden layers: epilogue e reads layer e's accumulator and writes layer e+1's A operand ----
      // Per half (two 16-column chunks): wait for the half's commit, load both chunks, release the half of D as
      // soon as the loads have landed, then compute.  The lo half of A_s is stored only after acc[s][hi]: until
      // then A_s is still the operand of this layer's hi-half MMAs.
#pragma unroll 1
      for (int e = 1; e < nh; ++e) {
        const float4* bl = reinterpret_cast<const float4*>(sBias + e * kHid);
#pragma unroll 1
        for (int s = 0; s < nvalid; ++s) {
          const uint32_t a_tmem = d_tmem + 64 + 32 * s;
          uint32_t r[2][16];
          uint32_t pk[2][8];
          ev(0x1000u + (e << 8) + (s << 4));
          pair_wait(bar(kAcc + 2 * s), (par >> (2 + 2 * s)) & 1u, wait_ns, 0x500u + s);
          lap(3);
          ev(0x2000u + (e << 8) + (s << 4));
          tc_fence_after();
          tmem_ld16(d_tmem, r[0]);
          tmem_ld16(d_tmem + 16, r[1]);
          tmem_ld_wait();
          lap(4);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kDFree + 0));  // the other tile's next layer may overwrite D columns 0-31
          ev(0x3000u + (e << 8) + (s << 4));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float h[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = bl[4 * c + i];
              h[4 * i + 0] = __uint_as_float(r[c][4 * i + 0]) + bv.x;
              h[4 * i + 1] = __uint_as_float(r[c][4 * i + 1]) + bv.y;
              h[4 * i + 2] = __uint_as_float(r[c][4 * i + 2]) + bv.z;
              h[4 * i + 3] = __uint_as_float(r[c][4 * i + 3]) + bv.w;
            }
            activate16<kPolyPairs>(h, pk[c], P);
          }
          lap(5);
          ev(0x4000u + (e << 8) + (s << 4));
          pair_wait(bar(kAcc + 2 * s + 1), (par >> (3 + 2 * s)) & 1u, wait_ns, 0x600u + s);
          par ^= 3u << (2 + 2 * s);
          lap(6);
          ev(0x5000u + (e << 8) + (s << 4));
          tc_fence_after();
          tmem_ld16(d_tmem + 32, r[0]);
          tmem_ld16(d_tmem + 48, r[1]);
          tmem_st8(a_tmem, pk[0]);
          tmem_st8(a_tmem + 8, pk[1]);
          tmem_ld_wait();
          lap(7);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kDFree + 1));  // D columns 32-63 are in registers
          ev(0x6000u + (e << 8) + (s << 4));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float h[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = bl[8 + 4 * c + i];
              h[4 * i + 0] = __uint_as_float(r[c][4 * i + 0]) + bv.x;
              h[4 * i + 1] = __uint_as_float(r[c][4 * i + 1]) + bv.y;
              h[4 * i + 2] = __uint_as_float(r[c][4 * i + 2]) + bv.z;
              h[4 * i + 3] = __uint_as_float(r[c][4 * i + 3]) + bv.w;
            }
            activate16<kPolyPairs>(h, pk[c], P);
            tmem_st8(a_tmem + 16 + 8 * c, pk[c]);
          }
          lap(8);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kAReady + s));
          lap(9);
          ev(0x7000u + (e << 8) + (s << 4));
        }
      }

      // ---- head: density, exp, store, marching-cubes sign ballot ----------------------------------------
#pragma unroll 1
      for (int s = 0; s < nvalid; ++s) {
        uint32_t r[4];
        pair_wait(bar(kAcc + 2 * s), (par >> (2 + 2 * s)) & 1u, wait_ns, 0x700u + s);
        lap(10);
        tc_fence_after();
        tmem_ld4(d_tmem, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kDFree + 0));
        struct { long long line; int k0, nvalid; } t;
        t.line = (t0 + s) / tiles_per_line;
        t.k0 = ((t0 + s) - (int)t.line * tiles_per_line) * kTileM;
        t.nvalid = min(kTileM, p.R - t.k0);
        const float d = __uint_as_float(r[0]) + sBiasF[0];
        const float act = expf(__fadd_rn(d, p.density_bias));
        if (m < t.nvalid) {
          const long long o = t.line * p.R + t.k0 + m;
          if (p.out_raw) p.out_raw[o] = d;
          p.out_act[o] = act;
        }
        if (p.sign_out) {
          // marching-cubes case bits of this warp's 32 consecutive z-samples = one sign-mask word
          // (same fp32 expression as mc_signs in mcubes.cu, on the value that was just stored)
          const bool bit = m < t.nvalid && __fmul_rn(__fsub_rn(act, p.sign_sub), p.sign_mul) > 0.0f;
          const uint32_t mask = __ballot_sync(0xffffffffu, bit);
          if (lane == 0 && q * 32 < t.nvalid) p.sign_out[t.line * p.sign_wz + (t.k0 >> 5) + q] = mask;
        }
        // d_free[hi] may only be released once the issuer has consumed the previous owner's release of it, i.e.
        // after acc[s][hi] of this step (committed behind that wait): the head needs only acc[s][lo], so releasing
        // both halves at once lets d_free[hi] run two phases ahead of the issuer (parity aliasing -> deadlock;
        // found with a scheduling simulation of this protocol, tools/sim_pair_protocol.py)
        pair_wait(bar(kAcc + 2 * s + 1), (par >> (3 + 2 * s)) & 1u, wait_ns, 0x800u + s);
        par ^= 3u << (2 + 2 * s);
        if (lane == 0) mbar_arrive(bar(kDFree + 1));
        lap(11);
      }
    }
    if (kProf && lane == 0) {
      pc[0] = clock() - pc[0];
#pragma unroll
      for (int i = 0; i < 13; ++i) g_pair_prof[(blockIdx.x * 16 + wid) * 16 + i] = pc[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

template <int kPolyPairs, bool kProf = false>
static int launch_pair_n(const TcParams& p, int sms, cudaStream_t st) {
  const int wbytes = tc_weight_bytes(p.n_hidden);
  const size_t smem = (size_t)((wbytes + 1023) / 1024) * 1024 + (size_t)2 * kPairWG * ta_table_bytes(p.trows) +
                      8 * (1 + kBarsPerWG * kPairWG) + 32 + (size_t)((p.R + kTileM - 1) / kTileM) * kTileM * 8;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  auto kern = lattice_tc_pair_kernel<kPolyPairs, kProf>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  const long long ntiles = (long long)p.nx * p.R * ((p.R + kTileM - 1) / kTileM);
  if (ntiles > 0x7ffffff0LL) return SMB_ERR_BAD_ARG;
  long long grid = (ntiles + 2 * kPairWG - 1) / (2 * kPairWG);
  if (grid > sms) grid = sms;
  kern<<<(unsigned)grid, kPairThreads, smem, st>>>(p);
  return smb_check(cudaGetLastError());
}

int pair_evt_read(unsigned int* host, int n) {
  if (!host || n <= 0 || n > 20 * kEvtMax * 2) return SMB_ERR_BAD_ARG;
  return smb_check(cudaMemcpyFromSymbol(host, g_pair_evt, sizeof(unsigned int) * n));
}
int pair_prof_read(unsigned int* host, int n) {
  if (!host || n <= 0 || n > 148 * 16 * 16 + 148 * 4 * 4) return SMB_ERR_BAD_ARG;
  return smb_check(cudaMemcpyFromSymbol(host, g_pair_prof, sizeof(unsigned int) * n));
}
static unsigned long long* g_dbg_host = nullptr;
int pair_debug_dump() {
  if (!g_dbg_host) return 0;
  int n = 0;
  for (int i = 0; i < 148 * 32; ++i)
    if (g_dbg_host[i]) {
      printf("pair_dbg block %d warp %d code 0x%x\n", (int)((g_dbg_host[i] >> 8) & 0xff), (int)(g_dbg_host[i] & 0xff), (unsigned)(g_dbg_host[i] >> 32));
      ++n;
    }
  fflush(stdout);
  return n;
}
static void pair_debug_init() {
  if (g_dbg_host || !getenv("SMB_TC_DEBUG")) return;
  cudaHostAlloc(&g_dbg_host, 148 * 32 * 8, cudaHostAllocMapped);
  memset(g_dbg_host, 0, 148 * 32 * 8);
  unsigned long long* d = nullptr;
  cudaHostGetDevicePointer(&d, g_dbg_host, 0);
  cudaMemcpyToSymbol(g_pair_dbg, &d, sizeof(d));
}

// poly_pairs: how many of the 8 activation pairs of a 16-column chunk use the FMA-pipe polynomial (0..3)
int launch_tc_pair(const TcParams& p, int sms, int poly_pairs, cudaStream_t st) {
  pair_debug_init();
  if (p.dbg == 3) return launch_pair_n<0, true>(p, sms, st);  // SMB_TC_TRACE=3: per-phase clock accumulators (developer)
  switch (poly_pairs) {
    case 0: return launch_pair_n<0>(p, sms, st);
    case 1: return launch_pair_n<1>(p, sms, st);
    case 2: return launch_pair_n<2>(p, sms, st);
    case 3: return launch_pair_n<3>(p, sms, st);
  }
  return SMB_ERR_BAD_ARG;
}

}  // namespace smb
#endif  // SMB_DEV_VARIANTS

// Lattice field kernel, "activations in TMEM" variant (see field_tc.cu for the path and the
// reference citations).  Same tile / layer-0 / SiLU scheme, different operand plumbing:
//
//   * the A operand of every tcgen05.mma (the fp16 activations of the tile) lives in TENSOR
//     MEMORY: the epilogue writes it with tcgen05.st and the MMA reads it from there, so the only
//     shared-memory traffic of a layer step is the 8 KB weight image.  In field_tc.cu the A tile
//     is in shared memory: each step stores 16 KB (STS), fences the generic->async proxy, and the
//     tensor core re-reads it -- 192 B/clk of operand fetch for an M128 x N64 x K16 instruction
//     against the 128 B/clk the shared memory delivers, competing with the epilogue's own
//     LDS/STS in the same MIO queue as the MUFU ops that bound the kernel.
//   * without A tiles shared memory is no longer the limit: 4 consumer warpgroups (one tile in
//     flight each, 4 warps per SM sub-partition), TMEM = 4 x (64 accumulator + 32 activation
//     columns); a warpgroup's MMA latency is covered by the other three.
//   * 4 producer warps (one per warpgroup) build the layer-0 tables (T, c) one tile ahead, as in
//     field_tc.cu (measured at 256^3: 1 producer 4.69 ms, 2: 2.90 ms, 4: 2.85 ms).
//   * kBiasMMA (default since round 2): the bias of a hidden layer enters the accumulator through a fifth K=16 MMA -- a
//     constant activation block (1, 1, 0, ...) in TMEM against a weight block whose K rows 0/1 hold
//     b_l/2 as fp16 hi + lo -- instead of 16 LDS.128 + 64 FADD per sample and layer in the epilogue (a quarter of the
//     epilogue's instructions).  The constant block is the same for every tile, so ONE copy serves all warpgroups
//     (5 x 96 + 8 = 488 TMEM columns); round 1 gave each warpgroup its own (128 columns per slot -> only four tiles in
//     flight), which is why it measured no gain then.  256^3: 2.85 -> 2.76 ms on the same GPU.
//   * kPoly (default 5): kPoly of the 32 activation PAIRS of a layer step take their tanh from an FMA-pipe polynomial
//     (silu_poly2 below) -- the software-exponential trick of FlashAttention-4 applied to SiLU.  In round 1 (latency-bound
//     kernel, a polynomial with 1e-3 error) it was slower; with the leaner steps of round 2 the SFU is busy 85 % of the time.
//     All epilogue arithmetic runs on fp32 PAIRS (sm_100's FFMA2: one issue slot for two lanes), which halves the issue
//     cost of the polynomial: scalar, 3 of 64 activations was the optimum (2.60 ms); packed, 5 pairs = 10 of 64 (2.50 ms).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "field_tc_common.cuh"

namespace smb {

#ifdef SMB_DEV_VARIANTS
constexpr bool kDev = true;  // developer build: bias-MMA / polynomial / token / stagger / trace variants are compiled
#else
constexpr bool kDev = false;  // product build: one kernel (5 warpgroups) and its 4-warpgroup fallback, no experiment code
#endif

// TMEM columns per warpgroup: 64 accumulator + 32 activation.  With kBiasMMA the constant activation block (8 columns: it is the
// same for every tile) is shared by all warpgroups and sits behind their slots: 5 x 96 + 8 = 488 of 512 columns.
__host__ __device__ constexpr int ta_cols_per_wg(bool) { return 96; }

// silu(2h) = h + |h| tanh(|h|) on the FMA pipe (no MUFU):  tanh(|h|) ~= q(x),  x = min(|h| * (2/4.5) - 1, 1)  in [-1, 1].
// q: degree-9 minimax polynomial of tanh(2.25 (x + 1)) in the CENTRED variable -- every coefficient below 1 in magnitude, so
// the fp32 Horner evaluation is well conditioned (the round-1 form, a polynomial in h^2 on [0, 16], loses 3 digits to
// cancellation and stalled at 1.0e-3).  Max abs error of tanh over all h, clamp included (1 - tanh(4.5) = 2.5e-4), under fp32
// evaluation: 2.5e-4 -- BELOW tanh.approx.f32's 5e-4, so an activation that takes this path is no less accurate than one
// that takes the SFU.  1 FFMA (x) + 1 FMNMX + 9 FFMA + 1 FFMA = 12 instructions.
__device__ __forceinline__ float silu_poly(float h) {
  const float a = fabsf(h);
  const float x = fminf(fmaf(a, 2.0f / 4.5f, -1.0f), 1.0f);
  float q = 3.815771247e-02f;
  q = fmaf(q, x, 1.059716077e-01f);
  q = fmaf(q, x, -2.874662484e-01f);
  q = fmaf(q, x, -1.656126921e-02f);
  q = fmaf(q, x, 3.707452418e-01f);
  q = fmaf(q, x, -3.581689483e-01f);
  q = fmaf(q, x, 2.788679147e-01f);
  q = fmaf(q, x, -2.090362925e-01f);
  q = fmaf(q, x, 9.955515848e-02f);
  q = fmaf(q, x, 9.778961789e-01f);
  return fmaf(a, q, h);
}
// The same on a PAIR of activations with sm_100's packed fp32 instructions (FFMA2: one issue slot for two lanes, two
// FMA-pipe cycles): 15 instructions for two activations instead of 24.
__device__ __forceinline__ float2 silu_poly2(float2 h) {
  const float2 a = make_float2(fabsf(h.x), fabsf(h.y));
  float2 x = __ffma2_rn(a, make_float2(2.0f / 4.5f, 2.0f / 4.5f), make_float2(-1.0f, -1.0f));
  x = make_float2(fminf(x.x, 1.0f), fminf(x.y, 1.0f));
  float2 q = make_float2(3.815771247e-02f, 3.815771247e-02f);
  q = __ffma2_rn(q, x, make_float2(1.059716077e-01f, 1.059716077e-01f));
  q = __ffma2_rn(q, x, make_float2(-2.874662484e-01f, -2.874662484e-01f));
  q = __ffma2_rn(q, x, make_float2(-1.656126921e-02f, -1.656126921e-02f));
  q = __ffma2_rn(q, x, make_float2(3.707452418e-01f, 3.707452418e-01f));
  q = __ffma2_rn(q, x, make_float2(-3.581689483e-01f, -3.581689483e-01f));
  q = __ffma2_rn(q, x, make_float2(2.788679147e-01f, 2.788679147e-01f));
  q = __ffma2_rn(q, x, make_float2(-2.090362925e-01f, -2.090362925e-01f));
  q = __ffma2_rn(q, x, make_float2(9.955515848e-02f, 9.955515848e-02f));
  q = __ffma2_rn(q, x, make_float2(9.778961789e-01f, 9.778961789e-01f));
  return __ffma2_rn(a, q, h);
}
// activation pair jp (0..31) of a layer step: kPoly of the 32 pairs go to the FMA pipe, spread evenly over the step
// The same polynomial in Estrin form: 12 packed instructions + 3 squarings instead of 10 + clamp, but a dependency chain of 4
// instead of 10 (developer variant, kPoly = 2000 + pairs).
__device__ __forceinline__ float2 silu_poly2_estrin(float2 h) {
  const float2 a = make_float2(fabsf(h.x), fabsf(h.y));
  float2 x = __ffma2_rn(a, make_float2(2.0f / 4.5f, 2.0f / 4.5f), make_float2(-1.0f, -1.0f));
  x = make_float2(fminf(x.x, 1.0f), fminf(x.y, 1.0f));
  auto c2 = [](float v) { return make_float2(v, v); };
  const float2 x2 = __fmul2_rn(x, x);
  const float2 p01 = __ffma2_rn(c2(9.955515848e-02f), x, c2(9.778961789e-01f));
  const float2 p23 = __ffma2_rn(c2(2.788679147e-01f), x, c2(-2.090362925e-01f));
  const float2 p45 = __ffma2_rn(c2(3.707452418e-01f), x, c2(-3.581689483e-01f));
  const float2 p67 = __ffma2_rn(c2(-2.874662484e-01f), x, c2(-1.656126921e-02f));
  const float2 p89 = __ffma2_rn(c2(3.815771247e-02f), x, c2(1.059716077e-01f));
  const float2 x4 = __fmul2_rn(x2, x2);
  const float2 q03 = __ffma2_rn(p23, x2, p01);
  const float2 q47 = __ffma2_rn(p67, x2, p45);
  const float2 x8 = __fmul2_rn(x4, x4);
  const float2 q07 = __ffma2_rn(q47, x4, q03);
  const float2 q = __ffma2_rn(p89, x8, q07);
  return __ffma2_rn(a, q, h);
}
// Which pairs: kPoly < 1000 = that many pairs, spread by the rule ((jp + offset) * pairs) & 31 < pairs with offset = kPoly / 100
// (developer sweeps); kPoly >= 1000 = entry kPoly - 1000 of the candidate placements below (developer sweeps).
__host__ __device__ constexpr unsigned poly_bits(unsigned a, unsigned b = 32, unsigned c = 32, unsigned d = 32, unsigned e = 32, unsigned f = 32,
                                                 unsigned g = 32) {
  return (a < 32 ? 1u << a : 0u) | (b < 32 ? 1u << b : 0u) | (c < 32 ? 1u << c : 0u) | (d < 32 ? 1u << d : 0u) | (e < 32 ? 1u << e : 0u) |
         (f < 32 ? 1u << f : 0u) | (g < 32 ? 1u << g : 0u);
}
constexpr unsigned kPolyCand[] = {
    poly_bits(0, 7, 13, 20, 26),      poly_bits(0, 6, 13, 19, 26),      poly_bits(0, 8, 16, 24, 4),      poly_bits(0, 8, 16, 24, 12, 28),
    poly_bits(0, 7, 13, 20, 26, 31),  poly_bits(0, 1, 13, 20, 26),      poly_bits(0, 7, 14, 21, 28),     poly_bits(0, 5, 11, 16, 22, 27),
    poly_bits(0, 3, 7, 13, 20, 26),   poly_bits(0, 7, 13, 20, 26, 10),  poly_bits(0, 7, 13, 20, 26, 16), poly_bits(0, 7, 13, 20, 26, 23),
    poly_bits(0, 9, 13, 20, 26),      poly_bits(0, 7, 15, 20, 26),      poly_bits(0, 7, 13, 22, 26),     poly_bits(0, 7, 13, 20, 29),
};
template <int kPolyArg>
__host__ __device__ constexpr unsigned poly_mask() {
  constexpr int kPoly = kPolyArg >= 2000 ? kPolyArg - 2000 : kPolyArg;  // 2000 + n: the Estrin form, same placement rule
  if (kPoly >= 1000) return kPolyCand[kPoly >= 1000 ? kPoly - 1000 : 0];
  unsigned m = 0;
  for (int jp = 0; jp < 32; ++jp)
    if ((((jp + kPoly / 100) * (kPoly % 100)) & 31) < (kPoly % 100)) m |= 1u << jp;
  return m;
}
template <int kPoly>
__device__ __forceinline__ float2 silu_mix2(float2 h, int jp) {
  constexpr unsigned kMask = poly_mask<kPoly>();
  if (kPoly >= 2000) return ((kMask >> jp) & 1u) ? silu_poly2_estrin(h) : silu2_from_half_arg(h);
  return ((kMask >> jp) & 1u) ? silu_poly2(h) : silu2_from_half_arg(h);
}

// developer instrumentation (kTrace, SMB_TC_TRACE=2): clock64 stamps of block 0, every consumer warp, first kTraceSteps layer steps
constexpr int kTraceSteps = 160;
#ifdef SMB_DEV_VARIANTS
__device__ long long g_trace_ta[5 * 4 * kTraceSteps * 4];
#else
__device__ long long g_trace_ta[4];  // kTrace is never instantiated in the product build
#endif

template <int kTaWG, int kTaProducers, bool kBiasMMA, int kPoly, bool kTrace = false>
__global__ void __launch_bounds__(kTaWG * 128 + kTaProducers * 32, 1) lattice_tc_ta_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wbytes = tc_weight_bytes(nh);
  unsigned char* sW = smem;
  unsigned char* sWf = sW + (nh - 1) * kWBytes;
  const float* sBias = reinterpret_cast<const float*>(sWf + kWFinalBytes);
  const float* sBiasF = sBias + nh * kHid;
  unsigned char* sBB = smem + ((wbytes + 1023) / 1024) * 1024;  // bias K-blocks (kBiasMMA)
  const int bbbytes = kBiasMMA ? (nh - 1) * kWBytes : 0;
  unsigned char* tables = sBB + bbbytes;
  const int tbytes = ta_table_bytes(p.trows);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tables + kTaWG * tbytes);
  // bars[0] = weights; per warpgroup g: [1+3g] t_full, [2+3g] t_empty, [3+3g] acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 3 * kTaWG);
  int* xu_sem = reinterpret_cast<int*>(tmem_slot + 4);  // [4]: permits per SM sub-partition (p.xu_tokens > 0)
  // [64] fp32, 16-byte aligned (read as float4): row 0 of the last Linear (density), see the fused head below
  float* sHeadW = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(xu_sem + 4) + 15) & ~(uintptr_t)15);
  // per (z-segment of a line, sample m): layer-0 interpolation weight w1 and table row r0.  They depend only on the sample's
  // z index, so they are computed once per launch; round 1 recomputed them (a global load, floor, clamps and two 64-bit
  // divisions of the tile index) in every consumer thread for every tile
  float2* sGeo = reinterpret_cast<float2*>(sHeadW + kHid);

  const int tid_cta = threadIdx.x;
  const int wid = tid_cta >> 5;
  const int lane = tid_cta & 31;

  if (tid_cta == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int g = 0; g < kTaWG; ++g) {
      mbar_init(smem_u32(&bars[1 + 3 * g]), 1);  // t_full: elected producer lane
      mbar_init(smem_u32(&bars[2 + 3 * g]), 4);  // t_empty: one elected lane per consumer warp
      mbar_init(smem_u32(&bars[3 + 3 * g]), 1);  // acc_full: tcgen05.commit
    }
    mbar_fence_init();
    for (int q = 0; q < 4; ++q) xu_sem[q] = kDev ? p.xu_tokens : 0;
  }
  if (tid_cta < kHid) sHeadW[tid_cta] = p.head_w_f32[tid_cta];
  {
    const int tpl = (p.R + kTileM - 1) / kTileM;
    for (int i = tid_cta; i < tpl * kTileM; i += blockDim.x) {
      const int k0 = (i / kTileM) * kTileM;
      const float fz = unnormalize(p.axis_u[min(i, p.R - 1)], p.H, p.align_corners);
      const float hf = floorf(fz);
      const int hlo = (int)floorf(unnormalize(p.axis_u[k0], p.H, p.align_corners));
      sGeo[i] = make_float2(__fsub_rn(fz, hf), __int_as_float(min(max((int)hf - hlo, 0), p.trows - 2)));
    }
  }
  if (wid == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_w = smem_u32(&bars[0]);
  if (tid_cta == 0) {
    mbar_expect_tx(bar_w, (uint32_t)(wbytes + bbbytes));
    int off = 0;
    while (off < wbytes) {
      const int n = min(8192, wbytes - off);
      bulk_g2s(smem_u32(sW + off), p.tc_weights + off, (uint32_t)n, bar_w);
      off += n;
    }
    for (off = 0; off < bbbytes; off += kWBytes) bulk_g2s(smem_u32(sBB + off), p.tc_biasblk + off, (uint32_t)kWBytes, bar_w);
  }

  const int tiles_per_line = (p.R + kTileM - 1) / kTileM;
  const long long ntiles = (long long)p.nx * p.R * tiles_per_line;
  const long long HW = (long long)p.H * p.W;

  // kRegShift (5 warpgroups): the CTA is launched at 80 registers per thread (768 threads); the producer
  // warpgroup gives registers back and the consumers take them (the increase is served from the registers the CTA itself gave back: 128 x (80 - 40) = 640 x (88 - 80))
  constexpr bool kRegShift = kTaWG * 128 + kTaProducers * 32 > 736;  // up to 736 threads every thread can have 88 registers
  if (wid >= kTaWG * 4) {
    if (kRegShift) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // =================================================================== producer
    const int pi = wid - kTaWG * 4;
    uint32_t par_empty = 0xff;  // bit g: parity to wait on; a fresh barrier passes a wait on parity 1
    for (long long n = 0;; ++n) {
      bool any = false;
#pragma unroll 1
      for (int g = pi; g < kTaWG; g += kTaProducers) {
        const long long t = (n * gridDim.x + blockIdx.x) * kTaWG + g;
        if (t >= ntiles) continue;
        any = true;
        float* sT = reinterpret_cast<float*>(tables + g * tbytes);
        const TableGeom G = table_geom(p, t, tiles_per_line);
        mbar_wait_sleep(smem_u32(&bars[2 + 3 * g]), (par_empty >> g) & 1u, 20000u);
        par_empty ^= 1u << g;
        build_table(p, G, sT, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[1 + 3 * g]));  // t_full
      }
      if (!any) break;
    }
  } else {
    if (kRegShift) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // =================================================================== consumer
    const int wg = wid >> 2;
    const int q = wid & 3;
    const int m = (q << 5) | lane;
    const int tid_wg = tid_cta & 127;
    const float* sT = reinterpret_cast<const float*>(tables + wg * tbytes);
    const float* sC = sT + p.trows * kTPitch;
    const uint32_t d_tmem = tmem_base + (uint32_t)(wg * ta_cols_per_wg(kBiasMMA));
    const uint32_t a_tmem = d_tmem + 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t idesc_hidden = umma_idesc_f16_f32(128, 64);
    const uint32_t idesc_head = umma_idesc_f16_f32(128, 16);
    const uint32_t bar_acc = smem_u32(&bars[3 + 3 * wg]);
    uint32_t par_t = 0, par_acc = 0;
    int tr_n = 0;  // kTrace: layer steps recorded so far
    long long tr0 = 0, tr1 = 0, tr2 = 0;
    auto tr_put = [&](long long a, long long b, long long c, long long d) {
      if (kTrace && blockIdx.x == 0 && lane == 0 && tr_n < kTraceSteps) {
        long long* o = g_trace_ta + (((long long)(wg * 4 + q) * kTraceSteps) + tr_n) * 4;
        o[0] = a;  // step begins (about to wait for the accumulator / the table)
        o[1] = b;  // accumulator (table) ready
        o[2] = c;  // activations computed and stored to TMEM (tcgen05.st issued)
        o[3] = d;  // past tcgen05.wait::st + named barrier (the MMA of the next layer has been issued by thread 0)
        ++tr_n;
      }
    };

    // The four warps of an SM sub-partition (one per warpgroup) run identical steps, share its SFU
    // equally and therefore finish together and wait for their MMAs together: a convoy that leaves
    // the SFU idle for a whole MMA round trip per step.  Two ways to break it: a one-off start offset
    // per warpgroup (stagger_clk), or a semaphore that lets only xu_tokens warps of a sub-partition
    // into an activation stretch at a time (the others wait while their MMA would be waiting anyway).
    const bool use_tok = kDev && p.xu_tokens > 0;
    auto xu_acquire = [&]() {
      if (!use_tok) return;
      if (lane == 0) {
        uint32_t ns = 32;
        while (atomicSub(&xu_sem[q], 1) <= 0) {
          atomicAdd(&xu_sem[q], 1);
          __nanosleep(ns);
          if (ns < 256) ns <<= 1;
        }
      }
      __syncwarp();
    };
    auto xu_release = [&]() {
      if (!use_tok) return;
      __syncwarp();
      if (lane == 0) atomicAdd(&xu_sem[q], 1);
    };

    if (kBiasMMA) {  // constant activation block: k = 64, 65 -> 1.0 (bias hi, lo rows), k = 66..79 -> 0
      const uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(tmem_base + (uint32_t)(kTaWG * 96) + lane_off, one);  // every warpgroup writes the same constants (idempotent)
    }
    mbar_wait(bar_w, 0);

    // hand the finished activation columns to the tensor core: layer L = 1..nh
    auto issue_layer = [&](int L) {
      tmem_st_wait();
      tc_fence_before();
      named_bar_sync(1 + wg, 128);
      if (tid_wg == 0) {
        tc_fence_after();
        const bool head = (L == nh);
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(head ? sWf : sW + (L - 1) * kWBytes));
        const uint32_t idesc = head ? idesc_head : idesc_hidden;
#pragma unroll
        for (int kc = 0; kc < kHid / 16; ++kc)  // K = 16 per instruction = 8 packed columns of A, +32 B of B
          umma_f16_ts(d_tmem, a_tmem + 8 * kc, b_desc + 2 * kc, idesc, kc > 0 ? 1u : 0u);
        if (kBiasMMA && !head)
          umma_f16_ts(d_tmem, tmem_base + (uint32_t)(kTaWG * 96), umma_desc_k_sw128(smem_u32(sBB + (L - 1) * kWBytes)), idesc, 1u);
        umma_commit(bar_acc);
      }
    };

    for (long long n = 0;; ++n) {
      const long long t = (n * gridDim.x + blockIdx.x) * kTaWG + wg;
      if (t >= ntiles) break;
      struct { long long line; int k0, nvalid; } tg;  // 32-bit tile arithmetic (launch_tc_ta_n rejects >= 2^31 tiles)
      {
        const unsigned tu = (unsigned)t, tpl = (unsigned)tiles_per_line;
        const unsigned line = tu / tpl;
        tg.line = (long long)line;
        tg.k0 = (int)(tu - line * tpl) * kTileM;
        tg.nvalid = min(kTileM, p.R - tg.k0);
      }
      {
        // ---- layer 0 from the producer's table -> activation columns ------------------
        const float2 ge = sGeo[tg.k0 + m];
        const float w1 = ge.x;
        const float w0 = __fsub_rn(1.0f, w1);
        const int r0 = __float_as_int(ge.y);
        if (kTrace) tr0 = clock64();
        mbar_wait_sleep(smem_u32(&bars[1 + 3 * wg]), par_t, (uint32_t)p.wait_ns);
        par_t ^= 1u;
        if (kTrace) tr1 = clock64();
        if (kDev && n == 0 && p.stagger_clk > 0 && wg > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < (long long)wg * p.stagger_clk) {}
        }
        xu_acquire();
        const float4* t0p = reinterpret_cast<const float4*>(sT + r0 * kTPitch);
        const float4* t1p = reinterpret_cast<const float4*>(sT + (r0 + 1) * kTPitch);
        const float4* cc = reinterpret_cast<const float4*>(sC);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 4 chunks of 16 columns
          uint32_t pk[8];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 a = t0p[4 * c + g4], b = t1p[4 * c + g4], cv = cc[4 * c + g4];
            // c + w0 T[h0] + w1 T[h1] on fp32 pairs (FFMA2): same operations and order as the scalar form
            const float2 h01 = __ffma2_rn(make_float2(w1, w1), make_float2(b.x, b.y), __ffma2_rn(make_float2(w0, w0), make_float2(a.x, a.y), make_float2(cv.x, cv.y)));
            const float2 h23 = __ffma2_rn(make_float2(w1, w1), make_float2(b.z, b.w), __ffma2_rn(make_float2(w0, w0), make_float2(a.z, a.w), make_float2(cv.z, cv.w)));
            const float2 s01 = silu_mix2<kPoly>(h01, 8 * c + 2 * g4), s23 = silu_mix2<kPoly>(h23, 8 * c + 2 * g4 + 1);
            pk[2 * g4 + 0] = pack_half2(s01.x, s01.y);
            pk[2 * g4 + 1] = pack_half2(s23.x, s23.y);
          }
          tmem_st8(a_tmem + lane_off + 8 * c, pk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[2 + 3 * wg]));  // t_empty
        xu_release();
        if (kTrace) tr2 = clock64();
        issue_layer(1);
        if (kTrace) tr_put(tr0, tr1, tr2, clock64());
      }
      // Hidden layers 1 .. nh-1.  The epilogue of the LAST one does not write an A operand: the head of the density decoder
      // (network_utils.py:122, out[...,0]) is a 64-term dot product per sample, evaluated right here in fp32 on the
      // activations while they are in registers (weights broadcast from shared memory).  Round 1 ran it as a tenth
      // tensor-core step (N = 16 MMA: 1749 clocks per tile, almost all of it hand-off latency for 3 % of the FLOPs) on
      // activations rounded to fp16; fusing it removes a tcgen05.st + wait::st + barrier + MMA round trip + tcgen05.ld per
      // tile and is closer to the fp32 reference.  The three `features` outputs are not needed on the lattice
      // (extract_mesh uses density_act only, system.py:176-184).
      for (int l = 1; l < nh; ++l) {
        if (kTrace) tr0 = clock64();
        mbar_wait_sleep(bar_acc, par_acc, (uint32_t)p.wait_ns);
        par_acc ^= 1u;
        tc_fence_after();
        if (kTrace) tr1 = clock64();
        const bool last = l == nh - 1;
        xu_acquire();
        const float4* bl = reinterpret_cast<const float4*>(sBias + l * kHid);
        const float4* hw = reinterpret_cast<const float4*>(sHeadW);
        float2 dacc = make_float2(0.0f, 0.0f);  // even / odd partial sums of the head's dot product
        uint32_t r[2][16];
        float4 bb[2][4];
        tmem_ld16(d_tmem + lane_off, r[0]);
        if (!kBiasMMA) {
#pragma unroll
          for (int i = 0; i < 4; ++i) bb[0][i] = bl[i];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c + 1 < 4) {
            tmem_ld16(d_tmem + lane_off + (c + 1) * 16, r[(c + 1) & 1]);
            if (!kBiasMMA) {
#pragma unroll
              for (int i = 0; i < 4; ++i) bb[(c + 1) & 1][i] = bl[(c + 1) * 4 + i];
            }
          }
          const uint32_t* rc = r[c & 1];
          float h[16];
          if (kBiasMMA) {
#pragma unroll
            for (int i = 0; i < 16; ++i) h[i] = __uint_as_float(rc[i]);
          } else {
            const float4* bc = bb[c & 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              h[4 * i + 0] = __uint_as_float(rc[4 * i + 0]) + bc[i].x;
              h[4 * i + 1] = __uint_as_float(rc[4 * i + 1]) + bc[i].y;
              h[4 * i + 2] = __uint_as_float(rc[4 * i + 2]) + bc[i].z;
              h[4 * i + 3] = __uint_as_float(rc[4 * i + 3]) + bc[i].w;
            }
          }
          float2 h2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h2[i] = silu_mix2<kPoly>(make_float2(h[2 * i], h[2 * i + 1]), 8 * c + i);
          if (!last) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_half2(h2[i].x, h2[i].y);
            tmem_st8(a_tmem + lane_off + 8 * c, pk);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 w = hw[4 * c + i];
              dacc = __ffma2_rn(h2[2 * i], make_float2(w.x, w.y), dacc);
              dacc = __ffma2_rn(h2[2 * i + 1], make_float2(w.z, w.w), dacc);
            }
          }
        }
        xu_release();
        if (kTrace) tr2 = clock64();
        if (!last) {
          issue_layer(l + 1);
          if (kTrace) tr_put(tr0, tr1, tr2, clock64());
        } else {
          const float d = dacc.x + dacc.y + sBiasF[0];
          const float act = expf(__fadd_rn(d, p.density_bias));
          if (m < tg.nvalid) {
            const long long o = tg.line * p.R + tg.k0 + m;
            if (p.out_raw) p.out_raw[o] = d;
            p.out_act[o] = act;
          }
          if (p.sign_out) {
            // marching-cubes case bits of this warp's 32 consecutive z-samples = one sign-mask word
            // (same fp32 expression as mc_signs in mcubes.cu, on the value that was just stored)
            const bool bit = m < tg.nvalid && __fmul_rn(__fsub_rn(act, p.sign_sub), p.sign_mul) > 0.0f;
            const uint32_t mask = __ballot_sync(0xffffffffu, bit);
            if (lane == 0 && q * 32 < tg.nvalid) p.sign_out[tg.line * p.sign_wz + (tg.k0 >> 5) + q] = mask;
          }
          if (kTrace) {
            const long long t = clock64();
            tr_put(tr0, tr1, t, t);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}


template <int kTaWG, int kTaProducers, bool kBiasMMA, int kPoly, bool kTrace = false>
static int launch_tc_ta_n(const TcParams& p, int sms, cudaStream_t st) {
  const int wbytes = tc_weight_bytes(p.n_hidden);
  const size_t smem = (size_t)((wbytes + 1023) / 1024) * 1024 + (kBiasMMA ? (size_t)(p.n_hidden - 1) * kWBytes : 0) +
                      (size_t)kTaWG * ta_table_bytes(p.trows) + 8 * (1 + 3 * kTaWG) + 48 + kHid * 4 +
                      (size_t)((p.R + kTileM - 1) / kTileM) * kTileM * 8;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  auto kern = lattice_tc_ta_kernel<kTaWG, kTaProducers, kBiasMMA, kPoly, kTrace>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  const long long ntiles = (long long)p.nx * p.R * ((p.R + kTileM - 1) / kTileM);
  if (ntiles >= 0x7fffffffLL) return SMB_ERR_CUDA;  // 32-bit tile arithmetic in the kernel (R^3 / 128 tiles: R up to ~6500)
  long long grid = (ntiles + kTaWG - 1) / kTaWG;
  if (grid > sms) grid = sms;
  kern<<<(unsigned)grid, kTaWG * 128 + kTaProducers * 32, smem, st>>>(p);
  return smb_check(cudaGetLastError());
}

// Default: FIVE consumer warpgroups (5 x 96 + 8 = 488 of the 512 TMEM columns; 768 threads, the producer warpgroup hands 40
// of its 80 registers to the consumers with setmaxnreg so that they run at 88), hidden-layer bias through a fifth K=16
// MMA, the density head fused into the last epilogue, all epilogue arithmetic on fp32 pairs (FFMA2), and kDefaultPoly = 5
// of the 32 activation pairs of a layer step evaluated on the FMA pipe (silu_poly2, at least as accurate as tanh.approx)
// instead of the SFU.  Measured on one B200 at 256^3 (profiles/r02n_k1_packed.log), pairs on the FMA pipe: 0: 2.651,
// 1: 2.625, 2: 2.575, 3: 2.533, 4: 2.554, 5: 2.500, 6: 2.542, 8: 2.597, 10: 2.698, 12: 2.829 ms; WHICH pairs matters as much
// as how many (5 pairs shifted by 2 or 4 positions: 2.59 ms).  The product library instantiates exactly this kernel and its
// four-warpgroup form (the fallback when five table buffers do not fit in shared memory).  A developer build
// (-DSMB_DEV_VARIANTS) also compiles the variants whose measurements DESIGN.md 4/K1 argues from: SMB_TC_TA_POLY=<pairs of 32>
// (+ 100 x placement offset), SMB_TC_TA_BIAS=0 (bias in the epilogue), SMB_TC_TA_STAGGER=<clk>, SMB_TC_TA_TOKENS=1|2|3,
// SMB_TC_TA_WG=4, SMB_TC_TRACE=2.
constexpr int kDefaultPoly = 5;
int launch_tc_ta(const TcParams& p, int sms, cudaStream_t st) {
#ifdef SMB_DEV_VARIANTS
  static const int bias = getenv("SMB_TC_TA_BIAS") ? atoi(getenv("SMB_TC_TA_BIAS")) : 1;
  static const int poly = getenv("SMB_TC_TA_POLY") ? atoi(getenv("SMB_TC_TA_POLY")) : kDefaultPoly;
  static const int wgs = getenv("SMB_TC_TA_WG") ? atoi(getenv("SMB_TC_TA_WG")) : 5;
  static const int prod = getenv("SMB_TC_TA_PROD") ? atoi(getenv("SMB_TC_TA_PROD")) : 4;
  if (prod == 2) return launch_tc_ta_n<5, 2, true, kDefaultPoly>(p, sms, st);
  if (prod == 3) return launch_tc_ta_n<5, 3, true, kDefaultPoly>(p, sms, st);
  if (prod == 1) return launch_tc_ta_n<5, 1, true, kDefaultPoly>(p, sms, st);
  if (p.dbg == 2) return wgs == 5 ? launch_tc_ta_n<5, 4, false, 0, true>(p, sms, st) : launch_tc_ta_n<4, 4, false, 0, true>(p, sms, st);
  if (bias && wgs == 5 && poly != kDefaultPoly) {
    switch (poly) {
      case 0: return launch_tc_ta_n<5, 4, true, 0>(p, sms, st);
      case 1: return launch_tc_ta_n<5, 4, true, 1>(p, sms, st);
      case 2: return launch_tc_ta_n<5, 4, true, 2>(p, sms, st);
      case 3: return launch_tc_ta_n<5, 4, true, 3>(p, sms, st);
      case 10: return launch_tc_ta_n<5, 4, true, 10>(p, sms, st);
      case 12: return launch_tc_ta_n<5, 4, true, 12>(p, sms, st);
      case 7: return launch_tc_ta_n<5, 4, true, 7>(p, sms, st);
      case 2005: return launch_tc_ta_n<5, 4, true, 2005>(p, sms, st);
      case 2006: return launch_tc_ta_n<5, 4, true, 2006>(p, sms, st);
      case 2007: return launch_tc_ta_n<5, 4, true, 2007>(p, sms, st);
      case 2008: return launch_tc_ta_n<5, 4, true, 2008>(p, sms, st);
      case 1000: return launch_tc_ta_n<5, 4, true, 1000>(p, sms, st);
      case 1001: return launch_tc_ta_n<5, 4, true, 1001>(p, sms, st);
      case 1002: return launch_tc_ta_n<5, 4, true, 1002>(p, sms, st);
      case 1003: return launch_tc_ta_n<5, 4, true, 1003>(p, sms, st);
      case 1004: return launch_tc_ta_n<5, 4, true, 1004>(p, sms, st);
      case 1005: return launch_tc_ta_n<5, 4, true, 1005>(p, sms, st);
      case 1006: return launch_tc_ta_n<5, 4, true, 1006>(p, sms, st);
      case 1007: return launch_tc_ta_n<5, 4, true, 1007>(p, sms, st);
      case 1008: return launch_tc_ta_n<5, 4, true, 1008>(p, sms, st);
      case 1009: return launch_tc_ta_n<5, 4, true, 1009>(p, sms, st);
      case 1010: return launch_tc_ta_n<5, 4, true, 1010>(p, sms, st);
      case 1011: return launch_tc_ta_n<5, 4, true, 1011>(p, sms, st);
      case 1012: return launch_tc_ta_n<5, 4, true, 1012>(p, sms, st);
      case 1013: return launch_tc_ta_n<5, 4, true, 1013>(p, sms, st);
      case 1014: return launch_tc_ta_n<5, 4, true, 1014>(p, sms, st);
      case 1015: return launch_tc_ta_n<5, 4, true, 1015>(p, sms, st);
      case 4: return launch_tc_ta_n<5, 4, true, 4>(p, sms, st);
      case 6: return launch_tc_ta_n<5, 4, true, 6>(p, sms, st);
      default: return launch_tc_ta_n<5, 4, true, 8>(p, sms, st);
    }
  }
  if (!bias) {
    if (wgs == 5) {
      const int rc = poly ? launch_tc_ta_n<5, 4, false, kDefaultPoly>(p, sms, st) : launch_tc_ta_n<5, 4, false, 0>(p, sms, st);
      if (rc != SMB_ERR_BAD_ARG) return rc;
    }
    return launch_tc_ta_n<4, 4, false, 0>(p, sms, st);
  }
  if (wgs != 5) return launch_tc_ta_n<4, 4, true, kDefaultPoly>(p, sms, st);
#endif
  const int rc = launch_tc_ta_n<5, 4, true, kDefaultPoly>(p, sms, st);
  if (rc != SMB_ERR_BAD_ARG) return rc;  // five table buffers did not fit in shared memory: four warpgroups
  return launch_tc_ta_n<4, 4, true, kDefaultPoly>(p, sms, st);
}

#ifdef SMB_DEV_VARIANTS
int read_trace_ta(long long* host, int n) {
  if (!host || n <= 0 || n > 5 * 4 * kTraceSteps * 4) return SMB_ERR_BAD_ARG;
  return smb_check(cudaMemcpyFromSymbol(host, g_trace_ta, sizeof(long long) * n));
}
#endif

}  // namespace smb

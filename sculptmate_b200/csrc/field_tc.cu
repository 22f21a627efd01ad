// Fused lattice field kernel for sm_100a: triplane interpolation + NeRFMLP chain
// on tcgen05 tensor cores with TMEM accumulators -- the variant with the activation (A) tile in
// SHARED memory.  The default is field_tc_ta.cu (activations in tensor memory, 14 % faster); this
// kernel stays as SMB_TC_VARIANT=smem, as the comparison point of DESIGN.md and as the home of the
// timeline instrumentation (SMB_TC_TRACE=1, tools/trace_lattice.py).  This is the density half of
// TSR.extract_mesh (/root/reference/TripoSR/tsr/system.py:171-184), i.e.
// query_triplane (tsr/models/nerf_renderer.py:41-91) + NeRFMLP.forward
// (tsr/models/network_utils.py:116-124) evaluated on the R^3 lattice of
// MarchingCubeHelper.grid_vertices (tsr/models/isosurface.py:25-39).
//
// Work decomposition
//   tile   = 128 consecutive z-samples of one (x,y) lattice line = the M dimension
//            of one UMMA (M=128, N=64, K=64 per hidden layer).
//   slot   = the on-chip home of one tile in flight: a 16 KB fp16 A tile (K-major,
//            128B swizzle), the layer-0 z-row table T + constant vector c (fp32) and a
//            64-column TMEM accumulator.
//   CTA    = persistent, one per SM, warp-specialised:
//            * kWG consumer warpgroups (128 threads each).  A warpgroup owns TWO slots
//              and ping-pongs between them layer by layer: while the tensor core runs
//              layer l+1 of one tile the warpgroup runs the SiLU epilogue of the other,
//              so neither the MMA round trip nor the hand-off is exposed.  There is no
//              warpgroup-wide barrier: every warp publishes its 32 rows, bumps a
//              shared-memory counter, and whichever warp arrives last issues the
//              tcgen05.mma (one elected thread) and commits it to the slot's mbarrier.
//            * kWG producer warps (one per consumer warpgroup) build T and c for the
//              warpgroup's next tiles from the L2-resident projected planes, so global
//              load latency never sits in a consumer's instruction stream.
//   layer 0 is evaluated from the projected planes Q_p = (W0/2).plane_p (fp32):
//            on a z-line the (x,y) plane contributes a constant vector c and the
//            (x,z),(y,z) planes share the same z taps, so the pre-activation is
//            c + w0*T[h0] + w1*T[h1].  Exactly the reference's interpolate->Linear in
//            exact arithmetic, evaluated in fp32.
//   layers 1..L-1 (+ the 64->4 head): activations are written as fp16 into the A tile,
//            weights stay resident in shared memory for the whole kernel (brought in by
//            the bulk-copy engine), the accumulator comes back with tcgen05.ld, bias +
//            SiLU are applied in registers.
//   SiLU    silu(x) = h + h*tanh(h), h = x/2; the 1/2 is folded into weights and
//            biases on the host, so one MUFU.TANH + one FFMA per activation.  The SFU
//            (16 tanh/clk/SM, 576 per sample) is the practical bound of this kernel.
#ifdef SMB_DEV_VARIANTS  // developer build only (python -m sculptmate_b200.build --dev): superseded by field_tc_ta.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "field_tc_common.cuh"

namespace smb {

// developer instrumentation (kTrace, SMB_TC_TRACE=1): clock64 stamps of block 0 / warpgroup 0
__device__ long long g_trace[4 * 512 * 4];

template <int kWG, bool kTrace>
__global__ void __launch_bounds__(kWG * 192, 1) lattice_tc_kernel(TcParams p) {
  constexpr int kSlots = kWG * kSlotsPerWG;
  constexpr int kConsumerWarps = kWG * 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wbytes = tc_weight_bytes(nh);
  unsigned char* sW = smem;                                   // hidden images
  unsigned char* sWf = sW + (nh - 1) * kWBytes;               // head image
  const float* sBias = reinterpret_cast<const float*>(sWf + kWFinalBytes);  // [nh][64], row l = b_l/2
  const float* sBiasF = sBias + nh * kHid;                    // 4
  unsigned char* slots = smem + ((wbytes + 1023) / 1024) * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + kSlots * p.slot_bytes);
  // bars[0] = weights, then per slot s: [1+4s] t_full, [2+4s] t_empty, [3+4s] acc_full, [4+4s] a_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 4 * kSlots);

  const int tid_cta = threadIdx.x;
  const int wid = tid_cta >> 5;
  const int lane = tid_cta & 31;

  // ---- one-time setup ----------------------------------------------------
  if (tid_cta == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(smem_u32(&bars[1 + 4 * s]), 1);    // t_full: the producer warp (one elected lane)
      mbar_init(smem_u32(&bars[2 + 4 * s]), 4);    // t_empty: one elected lane per consumer warp
      mbar_init(smem_u32(&bars[3 + 4 * s]), 1);    // acc_full: tcgen05.commit
      mbar_init(smem_u32(&bars[4 + 4 * s]), 4);    // a_full: one elected lane per consumer warp
    }
    mbar_fence_init();
  }
  if (wid == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_w = smem_u32(&bars[0]);
  if (tid_cta == 0) {
    mbar_expect_tx(bar_w, (uint32_t)wbytes);
    // weights + biases via the bulk-copy engine (UBLKCP), 8 KB pieces
    int off = 0;
    while (off < wbytes) {
      int n = min(8192, wbytes - off);
      bulk_g2s(smem_u32(sW + off), p.tc_weights + off, (uint32_t)n, bar_w);
      off += n;
    }
  }

  const int tiles_per_line = (p.R + kTileM - 1) / kTileM;
  const long long ntiles = (long long)p.nx * p.R * tiles_per_line;
  const long long HW = (long long)p.H * p.W;

  if (wid >= kConsumerWarps + kWG) {
    // =================================================================== MMA issuer
    // One warp per consumer warpgroup, walking the warpgroup's (tile pair, layer, slot)
    // sequence in the same order: wait until all 128 rows of the A tile are published,
    // issue the next layer's tcgen05.mma from one lane, commit it to the slot's acc_full.
    const int g = wid - kConsumerWarps - kWG;
    const uint32_t idesc_hidden = umma_idesc_f16_f32(128, 64);
    const uint32_t idesc_head = umma_idesc_f16_f32(128, 16);
    uint32_t par_a = 0;
    mbar_wait(bar_w, 0);  // weights resident
    for (long long n = 0;; ++n) {
      const long long t0 = 2 * ((n * gridDim.x + blockIdx.x) * kWG + g);
      if (t0 >= ntiles) break;
      const bool have1 = t0 + 1 < ntiles;
      for (int l = 0; l < nh; ++l) {
#pragma unroll 1
        for (int s = 0; s < kSlotsPerWG; ++s) {
          if (s == 1 && !have1) break;
          const int slot = g * kSlotsPerWG + s;
          mbar_wait_sleep(smem_u32(&bars[4 + 4 * slot]), (par_a >> s) & 1u, (uint32_t)p.wait_ns);
          par_a ^= 1u << s;
          tc_fence_after();
          if (lane == 0) {
            const int L = l + 1;
            const bool head = (L == nh);
            const uint64_t a_desc = umma_desc_k_sw128(smem_u32(slots + slot * p.slot_bytes));
            const uint64_t b_desc = umma_desc_k_sw128(smem_u32(head ? sWf : sW + (L - 1) * kWBytes));
            const uint32_t idesc = head ? idesc_head : idesc_hidden;
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 64);
#pragma unroll
            for (int kc = 0; kc < kHid / 16; ++kc)  // K = 16 per instruction: +32 B along the swizzled row
              umma_f16_ss(d_tmem, a_desc + 2 * kc, b_desc + 2 * kc, idesc, kc > 0 ? 1u : 0u);
            umma_commit(smem_u32(&bars[3 + 4 * slot]));
          }
          __syncwarp();
        }
      }
    }
  } else if (wid >= kConsumerWarps) {
    // =================================================================== producer
    const int g = wid - kConsumerWarps;
    const float* Q0 = p.planes_q;
    const float* Q1 = Q0 + HW * kHid;
    const float* Q2 = Q1 + HW * kHid;
    uint32_t par_empty = 0x3;  // bit s: parity to wait on; a fresh barrier passes a wait on parity 1
    for (long long n = 0;; ++n) {
      const long long t0 = 2 * ((n * gridDim.x + blockIdx.x) * kWG + g);
      if (t0 >= ntiles) break;
#pragma unroll 1
      for (int s = 0; s < kSlotsPerWG; ++s) {
        const long long t = t0 + s;
        if (t >= ntiles) break;
        const int slot = g * kSlotsPerWG + s;
        unsigned char* sA = slots + slot * p.slot_bytes;
        float* sT = reinterpret_cast<float*>(sA + kABytes);
        float* sC = sT + p.trows * kTPitch;
        const TileGeom tg = tile_geom(t, tiles_per_line, p);
        const float ux = p.axis_u[p.x_begin + tg.i];
        const float uy = p.axis_u[tg.j];
        const Tap2 txw = make_tap(ux, p.W, p.align_corners);  // x on the W axis (planes 0,1)
        const Tap2 tyw = make_tap(uy, p.W, p.align_corners);  // y on the W axis (plane 2)
        const Tap2 tyh = make_tap(uy, p.H, p.align_corners);  // y on the H axis (plane 0)

        mbar_wait_sleep(smem_u32(&bars[2 + 4 * slot]), (par_empty >> s) & 1u, 20000u);
        par_empty ^= 1u << s;

        {  // c vector: lane owns n = 2*lane, 2*lane+1
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
        {  // T rows: lane owns float4 column n4 of rows (lane>>4), (lane>>4)+2, ...
          const int n4 = lane & 15;
#pragma unroll 4
          for (int r = lane >> 4; r < tg.nrow; r += 2) {
            const int h = tg.hlo + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h >= 0 && h < p.H) {
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
        __syncwarp();  // all lanes' table stores are ordered before the elected arrive (release)
        if (lane == 0) mbar_arrive(smem_u32(&bars[1 + 4 * slot]));  // t_full
      }
    }
  } else {
    // =================================================================== consumer
    const int wg = wid >> 2;
    const int q = wid & 3;               // TMEM lane quarter this warp may access
    const int m = (q << 5) | lane;       // row of the tile owned by this thread
    uint32_t par_t = 0, par_acc = 0;     // bit s: parity of the next completion to wait for

    mbar_wait(bar_w, 0);  // weights resident before anyone may issue an MMA

    for (long long n = 0;; ++n) {
      const long long t0 = 2 * ((n * gridDim.x + blockIdx.x) * kWG + wg);
      if (t0 >= ntiles) break;
      const bool have1 = t0 + 1 < ntiles;
      const TileGeom g0 = tile_geom(t0, tiles_per_line, p);
      const TileGeom g1 = tile_geom(have1 ? t0 + 1 : t0, tiles_per_line, p);

      for (int l = 0; l <= nh; ++l) {
#pragma unroll 1
        for (int s = 0; s < kSlotsPerWG; ++s) {
          if (s == 1 && !have1) break;
          const int slot = wg * kSlotsPerWG + s;
          unsigned char* sA = slots + slot * p.slot_bytes;
          const float* sT = reinterpret_cast<const float*>(sA + kABytes);
          const float* sC = sT + p.trows * kTPitch;
          const uint32_t a_addr = smem_u32(sA);
          unsigned char* a_rowp = sA + (m >> 3) * 1024 + (m & 7) * 128;  // this thread's 128-byte row
          const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 64);
          const uint32_t tmem_acc = d_tmem + ((uint32_t)(q * 32) << 16);
          const int k0 = s ? g1.k0 : g0.k0;
          long long tr0 = 0, tr1 = 0, tr2 = 0;
          if (kTrace) tr0 = clock64();

          if (l == 0) {
            // ---- layer 0 from the producer's table --------------------------------
            const int hlo = s ? g1.hlo : g0.hlo;
            const int kk = min(k0 + m, p.R - 1);
            const float fz = unnormalize(p.axis_u[kk], p.H, p.align_corners);
            const float hf = floorf(fz);
            const float w1 = __fsub_rn(fz, hf);
            const float w0 = __fsub_rn(1.0f, w1);
            int r0 = (int)hf - hlo;
            r0 = min(max(r0, 0), p.trows - 2);
            mbar_wait_sleep(smem_u32(&bars[1 + 4 * slot]), (par_t >> s) & 1u, (uint32_t)p.wait_ns);
            par_t ^= 1u << s;
            const float4* t0p = reinterpret_cast<const float4*>(sT + r0 * kTPitch);
            const float4* t1p = reinterpret_cast<const float4*>(sT + (r0 + 1) * kTPitch);
            const float4* cc = reinterpret_cast<const float4*>(sC);
            // 16 groups of 4 columns; the three table reads of group g+1 are issued before
            // group g is evaluated (the compiler will not hoist them over the A-tile stores)
            float4 ta[2], tb[2], tc[2];
            ta[0] = t0p[0];
            tb[0] = t1p[0];
            tc[0] = cc[0];
            uint32_t pk[4];
#pragma unroll
            for (int g4 = 0; g4 < 16; ++g4) {
              if (g4 + 1 < 16) {
                ta[(g4 + 1) & 1] = t0p[g4 + 1];
                tb[(g4 + 1) & 1] = t1p[g4 + 1];
                tc[(g4 + 1) & 1] = cc[g4 + 1];
              }
              const float4 a = ta[g4 & 1], b = tb[g4 & 1], c = tc[g4 & 1];
              const float h0 = c.x + w0 * a.x + w1 * b.x;
              const float h1 = c.y + w0 * a.y + w1 * b.y;
              const float h2 = c.z + w0 * a.z + w1 * b.z;
              const float h3 = c.w + w0 * a.w + w1 * b.w;
              pk[2 * (g4 & 1) + 0] = pack_half2(silu_from_half_arg(h0), silu_from_half_arg(h1));
              pk[2 * (g4 & 1) + 1] = pack_half2(silu_from_half_arg(h2), silu_from_half_arg(h3));
              if (g4 & 1) {
                const int c8 = g4 >> 1;
                *reinterpret_cast<uint4*>(a_rowp + ((c8 ^ (m & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[2 + 4 * slot]));  // t_empty: this warp is done reading T and c
          } else {
            mbar_wait_sleep(smem_u32(&bars[3 + 4 * slot]), (par_acc >> s) & 1u, (uint32_t)p.wait_ns);
            par_acc ^= 1u << s;
            tc_fence_after();
            if (kTrace) tr1 = clock64();
            if (l < nh) {
              // ---- hidden layer l: bias + SiLU, next A tile -------------------------
              // 4 chunks of 16 accumulator columns; the TMEM load and the bias row of chunk
              // c+1 are in flight while chunk c goes through bias + SiLU + pack.
              const float4* bl = reinterpret_cast<const float4*>(sBias + l * kHid);
              uint32_t r[2][16];
              float4 bb[2][4];
              tmem_ld16(tmem_acc, r[0]);
#pragma unroll
              for (int i = 0; i < 4; ++i) bb[0][i] = bl[i];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                tmem_ld_wait();
                if (c + 1 < 4) {
                  tmem_ld16(tmem_acc + (c + 1) * 16, r[(c + 1) & 1]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) bb[(c + 1) & 1][i] = bl[(c + 1) * 4 + i];
                }
                const uint32_t* rc = r[c & 1];
                const float4* bc = bb[c & 1];
                float h[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  h[4 * i + 0] = __uint_as_float(rc[4 * i + 0]) + bc[i].x;
                  h[4 * i + 1] = __uint_as_float(rc[4 * i + 1]) + bc[i].y;
                  h[4 * i + 2] = __uint_as_float(rc[4 * i + 2]) + bc[i].z;
                  h[4 * i + 3] = __uint_as_float(rc[4 * i + 3]) + bc[i].w;
                }
                float t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = tanh_approx(h[i]);
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = fmaf(h[i], t[i], h[i]);
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                  const uint32_t q0 = pack_half2(h[8 * hc + 0], h[8 * hc + 1]);
                  const uint32_t q1 = pack_half2(h[8 * hc + 2], h[8 * hc + 3]);
                  const uint32_t q2 = pack_half2(h[8 * hc + 4], h[8 * hc + 5]);
                  const uint32_t q3 = pack_half2(h[8 * hc + 6], h[8 * hc + 7]);
                  const int chunk = 2 * c + hc;
                  *reinterpret_cast<uint4*>(a_rowp + ((chunk ^ (m & 7)) << 4)) = make_uint4(q0, q1, q2, q3);
                }
              }
            } else {
              // ---- head: density logit -> exp ---------------------------------------
              uint32_t r[4];
              tmem_ld4(tmem_acc, r);
              tmem_ld_wait();
              const float d = __uint_as_float(r[0]) + sBiasF[0];
              const int nvalid = s ? g1.nvalid : g0.nvalid;
              if (m < nvalid) {
                const long long o = ((s ? g1.line : g0.line) * p.R) + k0 + m;
                if (p.out_raw) p.out_raw[o] = d;
                p.out_act[o] = expf(__fadd_rn(d, p.density_bias));
              }
            }
          }

          if (kTrace) tr2 = clock64();
          if (l < nh) {
            // publish this thread's row of the next A tile: generic-proxy stores -> async proxy,
            // TMEM reads ordered before the MMA that will overwrite the accumulator
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[4 + 4 * slot]));
          }
          if (kTrace && blockIdx.x == 0 && wg == 0 && lane == 0) {
            const long long ev = (n * (nh + 1) + l) * 2 + s;
            if (ev < 512) {
              long long* o = g_trace + ((long long)q * 512 + ev) * 4;
              o[0] = tr0;
              o[1] = tr1;
              o[2] = tr2;
              o[3] = clock64();
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

template <int kWG, bool kTrace>
static int launch_tc(const TcParams& p, int sms, cudaStream_t st) {
  const int wbytes = tc_weight_bytes(p.n_hidden);
  const int kSlots = kWG * kSlotsPerWG;
  const size_t smem = (size_t)((wbytes + 1023) / 1024) * 1024 + (size_t)kSlots * p.slot_bytes + 8 * (1 + 4 * kSlots) + 16;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(lattice_tc_kernel<kWG, kTrace>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  const long long ntiles = (long long)p.nx * p.R * ((p.R + kTileM - 1) / kTileM);
  long long grid = (ntiles + kSlots - 1) / kSlots;
  if (grid > sms) grid = sms;
  lattice_tc_kernel<kWG, kTrace><<<(unsigned)grid, kWG * 192, smem, st>>>(p);
  return smb_check(cudaGetLastError());
}

}  // namespace smb


namespace smb {
int launch_tc_smem(const TcParams& p, int sms, bool trace, cudaStream_t st) {
  if (trace) return launch_tc<3, true>(p, sms, st);
  int rc = launch_tc<3, false>(p, sms, st);
  if (rc == SMB_ERR_BAD_ARG) rc = launch_tc<2, false>(p, sms, st);
  if (rc == SMB_ERR_BAD_ARG) rc = launch_tc<1, false>(p, sms, st);
  return rc;
}
int read_trace_smem(long long* host, int n) {
  if (!host || n <= 0 || n > 4 * 512 * 4) return SMB_ERR_BAD_ARG;
  return smb_check(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * n));
}
}  // namespace smb
#endif  // SMB_DEV_VARIANTS

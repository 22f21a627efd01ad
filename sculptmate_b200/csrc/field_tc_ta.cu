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
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "field_tc_common.cuh"

namespace smb {

constexpr int kTaColsPerWG = 96;  // 64 accumulator + 32 activation columns

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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline int ta_table_bytes(int trows) { return ((trows * kTPitch * 4 + kHid * 4 + 127) / 128) * 128; }

template <int kTaWG, int kTaProducers>
__global__ void __launch_bounds__(kTaWG * 128 + kTaProducers * 32, 1) lattice_tc_ta_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wbytes = tc_weight_bytes(nh);
  unsigned char* sW = smem;
  unsigned char* sWf = sW + (nh - 1) * kWBytes;
  const float* sBias = reinterpret_cast<const float*>(sWf + kWFinalBytes);
  const float* sBiasF = sBias + nh * kHid;
  unsigned char* tables = smem + ((wbytes + 1023) / 1024) * 1024;
  const int tbytes = ta_table_bytes(p.trows);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tables + kTaWG * tbytes);
  // bars[0] = weights; per warpgroup g: [1+3g] t_full, [2+3g] t_empty, [3+3g] acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 3 * kTaWG);

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

  const int tiles_per_line = (p.R + kTileM - 1) / kTileM;
  const long long ntiles = (long long)p.nx * p.R * tiles_per_line;
  const long long HW = (long long)p.H * p.W;

  if (wid >= kTaWG * 4) {
    // =================================================================== producer
    const int pi = wid - kTaWG * 4;
    const float* Q0 = p.planes_q;
    const float* Q1 = Q0 + HW * kHid;
    const float* Q2 = Q1 + HW * kHid;
    uint32_t par_empty = 0xff;  // bit g: parity to wait on; a fresh barrier passes a wait on parity 1
    for (long long n = 0;; ++n) {
      bool any = false;
#pragma unroll 1
      for (int g = pi; g < kTaWG; g += kTaProducers) {
        const long long t = (n * gridDim.x + blockIdx.x) * kTaWG + g;
        if (t >= ntiles) continue;
        any = true;
        float* sT = reinterpret_cast<float*>(tables + g * tbytes);
        float* sC = sT + p.trows * kTPitch;
        const TileGeom tg = tile_geom(t, tiles_per_line, p);
        const float ux = p.axis_u[p.x_begin + tg.i];
        const float uy = p.axis_u[tg.j];
        const Tap2 txw = make_tap(ux, p.W, p.align_corners);
        const Tap2 tyw = make_tap(uy, p.W, p.align_corners);
        const Tap2 tyh = make_tap(uy, p.H, p.align_corners);
        mbar_wait_sleep(smem_u32(&bars[2 + 3 * g]), (par_empty >> g) & 1u, 20000u);
        par_empty ^= 1u << g;
        {
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
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[1 + 3 * g]));  // t_full
      }
      if (!any) break;
    }
  } else {
    // =================================================================== consumer
    const int wg = wid >> 2;
    const int q = wid & 3;
    const int m = (q << 5) | lane;
    const int tid_wg = tid_cta & 127;
    const float* sT = reinterpret_cast<const float*>(tables + wg * tbytes);
    const float* sC = sT + p.trows * kTPitch;
    const uint32_t d_tmem = tmem_base + (uint32_t)(wg * kTaColsPerWG);
    const uint32_t a_tmem = d_tmem + 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t idesc_hidden = umma_idesc_f16_f32(128, 64);
    const uint32_t idesc_head = umma_idesc_f16_f32(128, 16);
    const uint32_t bar_acc = smem_u32(&bars[3 + 3 * wg]);
    uint32_t par_t = 0, par_acc = 0;

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
        umma_commit(bar_acc);
      }
    };

    for (long long n = 0;; ++n) {
      const long long t = (n * gridDim.x + blockIdx.x) * kTaWG + wg;
      if (t >= ntiles) break;
      const TileGeom tg = tile_geom(t, tiles_per_line, p);
      {
        // ---- layer 0 from the producer's table -> activation columns ------------------
        const int kk = min(tg.k0 + m, p.R - 1);
        const float fz = unnormalize(p.axis_u[kk], p.H, p.align_corners);
        const float hf = floorf(fz);
        const float w1 = __fsub_rn(fz, hf);
        const float w0 = __fsub_rn(1.0f, w1);
        int r0 = (int)hf - tg.hlo;
        r0 = min(max(r0, 0), p.trows - 2);
        mbar_wait_sleep(smem_u32(&bars[1 + 3 * wg]), par_t, (uint32_t)p.wait_ns);
        par_t ^= 1u;
        const float4* t0p = reinterpret_cast<const float4*>(sT + r0 * kTPitch);
        const float4* t1p = reinterpret_cast<const float4*>(sT + (r0 + 1) * kTPitch);
        const float4* cc = reinterpret_cast<const float4*>(sC);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 4 chunks of 16 columns
          uint32_t pk[8];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 a = t0p[4 * c + g4], b = t1p[4 * c + g4], cv = cc[4 * c + g4];
            const float h0 = cv.x + w0 * a.x + w1 * b.x;
            const float h1 = cv.y + w0 * a.y + w1 * b.y;
            const float h2 = cv.z + w0 * a.z + w1 * b.z;
            const float h3 = cv.w + w0 * a.w + w1 * b.w;
            pk[2 * g4 + 0] = pack_half2(silu_from_half_arg(h0), silu_from_half_arg(h1));
            pk[2 * g4 + 1] = pack_half2(silu_from_half_arg(h2), silu_from_half_arg(h3));
          }
          tmem_st8(a_tmem + lane_off + 8 * c, pk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[2 + 3 * wg]));  // t_empty
        issue_layer(1);
      }
      for (int l = 1; l <= nh; ++l) {
        mbar_wait_sleep(bar_acc, par_acc, (uint32_t)p.wait_ns);
        par_acc ^= 1u;
        tc_fence_after();
        if (l < nh) {
          const float4* bl = reinterpret_cast<const float4*>(sBias + l * kHid);
          uint32_t r[2][16];
          float4 bb[2][4];
          tmem_ld16(d_tmem + lane_off, r[0]);
#pragma unroll
          for (int i = 0; i < 4; ++i) bb[0][i] = bl[i];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld_wait();
            if (c + 1 < 4) {
              tmem_ld16(d_tmem + lane_off + (c + 1) * 16, r[(c + 1) & 1]);
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
            float tt[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) tt[i] = tanh_approx(h[i]);
#pragma unroll
            for (int i = 0; i < 16; ++i) h[i] = fmaf(h[i], tt[i], h[i]);
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_half2(h[2 * i], h[2 * i + 1]);
            tmem_st8(a_tmem + lane_off + 8 * c, pk);
          }
          issue_layer(l + 1);
        } else {
          uint32_t r[4];
          tmem_ld4(d_tmem + lane_off, r);
          tmem_ld_wait();
          const float d = __uint_as_float(r[0]) + sBiasF[0];
          if (m < tg.nvalid) {
            const long long o = tg.line * p.R + tg.k0 + m;
            if (p.out_raw) p.out_raw[o] = d;
            p.out_act[o] = expf(__fadd_rn(d, p.density_bias));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc<512>(tmem_base);
}

template <int kTaWG, int kTaProducers>
static int launch_tc_ta_n(const TcParams& p, int sms, cudaStream_t st) {
  const int wbytes = tc_weight_bytes(p.n_hidden);
  const size_t smem = (size_t)((wbytes + 1023) / 1024) * 1024 + (size_t)kTaWG * ta_table_bytes(p.trows) + 8 * (1 + 3 * kTaWG) + 16;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(lattice_tc_ta_kernel<kTaWG, kTaProducers>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  const long long ntiles = (long long)p.nx * p.R * ((p.R + kTileM - 1) / kTileM);
  long long grid = (ntiles + kTaWG - 1) / kTaWG;
  if (grid > sms) grid = sms;
  lattice_tc_ta_kernel<kTaWG, kTaProducers><<<(unsigned)grid, kTaWG * 128 + kTaProducers * 32, smem, st>>>(p);
  return smb_check(cudaGetLastError());
}

int launch_tc_ta(const TcParams& p, int sms, cudaStream_t st) {
  const char* v = getenv("SMB_TC_TA_WG");
  const int wg = v ? atoi(v) : 4;
  v = getenv("SMB_TC_TA_PROD");
  const int pr = v ? atoi(v) : 4;
  if (wg == 3) return launch_tc_ta_n<3, 2>(p, sms, st);
  if (pr == 1) return launch_tc_ta_n<4, 1>(p, sms, st);
  if (pr == 4) return launch_tc_ta_n<4, 4>(p, sms, st);
  return launch_tc_ta_n<4, 2>(p, sms, st);
}

}  // namespace smb

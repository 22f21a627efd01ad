// Fused lattice field kernel for sm_100a: triplane interpolation + NeRFMLP chain
// on tcgen05 tensor cores with TMEM accumulators.  This is the density half of
// TSR.extract_mesh (/root/reference/TripoSR/tsr/system.py:171-184), i.e.
// query_triplane (tsr/models/nerf_renderer.py:41-91) + NeRFMLP.forward
// (tsr/models/network_utils.py:116-124) evaluated on the R^3 lattice of
// MarchingCubeHelper.grid_vertices (tsr/models/isosurface.py:25-39).
//
// Work decomposition
//   tile   = 128 consecutive z-samples of one (x,y) lattice line = the M dimension
//            of one UMMA (M=128, N=64, K=64 per hidden layer).
//   CTA    = kWG independent warpgroups (128 threads each), persistent, one CTA/SM.
//            A warpgroup owns a tile end to end, so there is no cross-warpgroup
//            hand-off; the SM's warp schedulers interleave the kWG pipelines and
//            hide each other's MMA / TMEM latencies.
//   layer 0 is evaluated from the projected planes Q_p = (W0/2).plane_p (fp32):
//            on a z-line the (x,y) plane contributes a constant vector and the
//            (x,z),(y,z) planes share the same z taps, so the pre-activation is
//            c + w0*T[h0] + w1*T[h1] with a per-tile table T built cooperatively
//            in shared memory.  Exactly the reference's interpolate->Linear in exact
//            arithmetic, evaluated in fp32.
//   layers 1..L-1 (+ the 64->4 head): activations are written as fp16 into the
//            K-major 128B-swizzled A tile in shared memory, one thread issues
//            tcgen05.mma (weights resident in shared memory for the whole kernel,
//            brought in by the bulk-copy engine), the accumulator comes back with
//            tcgen05.ld, bias + SiLU are applied in registers.
//   SiLU    silu(x) = h + h*tanh(h), h = x/2; the 1/2 is folded into weights and
//            biases on the host, so one MUFU.TANH + one FFMA per activation.
#include <cuda_runtime.h>
#include <math.h>

#include "field_common.cuh"
#include "ptx_sm100.cuh"

namespace smb {

constexpr int kTileM = 128;
constexpr int kTRows = 66;   // max plane rows one tile's z-range touches, incl. zero borders
constexpr int kTPitch = 68;  // floats per T row: 64 + 4 pad (adjacent rows land 4 banks apart)
constexpr int kWBytes = kHid * kHid * 2;     // 8192: one hidden layer, fp16
constexpr int kWFinalBytes = 16 * kHid * 2;  // 2048: head padded to N=16
constexpr int kABytes = kTileM * kHid * 2;   // 16384
constexpr int kWgBytes = ((kABytes + kTRows * kTPitch * 4 + kHid * 4 + 1023) / 1024) * 1024;

struct TcParams {
  const float* planes_q;  // (3,H,W,64) fp32
  const unsigned char* tc_weights;  // blob + off_tc_hidden: hidden images, head image, biases (contiguous)
  const float* bias0_half;          // b0/2 (64)
  const float* axis_u;
  int R, x_begin, nx, H, W, align_corners, n_hidden;
  float density_bias;
  float* out_act;
  float* out_raw;
  int* error_flag;
};

__host__ __device__ inline int tc_weight_bytes(int n_hidden) {
  // hidden images + head image + biases for layers 1..n_hidden-1 are not separate:
  // the blob keeps [hidden | head | bias_half (n_hidden x 64) | bias_final (4)] contiguous
  return (n_hidden - 1) * kWBytes + kWFinalBytes + n_hidden * kHid * 4 + 16;
}

template <int kWG>
__global__ void __launch_bounds__(kWG * 128, 1) lattice_tc_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nh = p.n_hidden;
  const int wbytes = tc_weight_bytes(nh);
  unsigned char* sW = smem;                                   // hidden images
  unsigned char* sWf = sW + (nh - 1) * kWBytes;               // head image
  const float* sBias = reinterpret_cast<const float*>(sWf + kWFinalBytes);  // [nh][64], row l = b_l/2
  const float* sBiasF = sBias + nh * kHid;                    // 4
  unsigned char* wg_region = smem + ((wbytes + 1023) / 1024) * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wg_region + kWG * kWgBytes);  // [0]=weights, [1+wg]=mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kWG);

  const int tid_cta = threadIdx.x;
  const int wg = tid_cta >> 7;
  const int tid = tid_cta & 127;
  const int warp_in_wg = tid >> 5;

  unsigned char* sA = wg_region + wg * kWgBytes;
  float* sT = reinterpret_cast<float*>(sA + kABytes);
  float* sC = sT + kTRows * kTPitch;

  const uint32_t bar_w = smem_u32(&bars[0]);
  const uint32_t bar_mma = smem_u32(&bars[1 + wg]);

  // ---- one-time setup ----------------------------------------------------
  if (tid_cta == 0) {
    mbar_init(bar_w, 1);
    for (int g = 0; g < kWG; ++g) mbar_init(smem_u32(&bars[1 + g]), 1);
    mbar_fence_init();
  }
  if (tid_cta < 32) tmem_alloc<kWG * 64>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
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

  const uint32_t tmem_acc = tmem_base + (uint32_t)(wg * 64) + ((uint32_t)(warp_in_wg * 32) << 16);
  const uint32_t idesc_hidden = umma_idesc_f16_f32(128, 64);
  const uint32_t idesc_head = umma_idesc_f16_f32(128, 16);
  const uint32_t a_addr = smem_u32(sA);
  uint32_t phase = 0;
  bool weights_ready = false;

  const int tiles_per_line = (p.R + kTileM - 1) / kTileM;
  const long long ntiles = (long long)p.nx * p.R * tiles_per_line;
  const long long HW = (long long)p.H * p.W;
  const float* Q0 = p.planes_q;
  const float* Q1 = Q0 + HW * kHid;
  const float* Q2 = Q1 + HW * kHid;

  for (long long t = (long long)blockIdx.x * kWG + wg; t < ntiles; t += (long long)gridDim.x * kWG) {
    const int seg = (int)(t % tiles_per_line);
    const long long line = t / tiles_per_line;
    const int i = (int)(line / p.R);
    const int j = (int)(line - (long long)i * p.R);
    const int k0 = seg * kTileM;
    const int nvalid = min(kTileM, p.R - k0);

    // ---- per-tile constants: c vector and T table --------------------------
    const float ux = p.axis_u[p.x_begin + i];
    const float uy = p.axis_u[j];
    const Tap2 txw = make_tap(ux, p.W, p.align_corners);  // x on the W axis (planes 0,1)
    const Tap2 tyw = make_tap(uy, p.W, p.align_corners);  // y on the W axis (plane 2)
    const Tap2 tyh = make_tap(uy, p.H, p.align_corners);  // y on the H axis (plane 0)
    const float fz_first = unnormalize(p.axis_u[k0], p.H, p.align_corners);
    const float fz_last = unnormalize(p.axis_u[k0 + nvalid - 1], p.H, p.align_corners);
    const int hlo = (int)floorf(fz_first);
    const int nrow = min((int)floorf(fz_last) + 1 - hlo + 1, kTRows);

    if (tid < kHid) {
      const int n = tid;
      const float q00 = __ldg(Q0 + ((long long)tyh.i0 * p.W + txw.i0) * kHid + n);
      const float q01 = __ldg(Q0 + ((long long)tyh.i0 * p.W + txw.i1) * kHid + n);
      const float q10 = __ldg(Q0 + ((long long)tyh.i1 * p.W + txw.i0) * kHid + n);
      const float q11 = __ldg(Q0 + ((long long)tyh.i1 * p.W + txw.i1) * kHid + n);
      float c = __ldg(p.bias0_half + n);
      c += tyh.w0 * (txw.w0 * q00 + txw.w1 * q01) + tyh.w1 * (txw.w0 * q10 + txw.w1 * q11);
      sC[n] = c;
    }
    {
      const int n4 = tid & 15;
      for (int r = tid >> 4; r < nrow; r += 8) {
        const int h = hlo + r;
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
    named_bar_sync(1 + wg, 128);

    // ---- layer 0: this thread's sample (row m of the tile) ------------------
    const int m = tid;
    const uint32_t a_row = a_addr + (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    {
      const int kk = min(k0 + m, p.R - 1);
      const float fz = unnormalize(p.axis_u[kk], p.H, p.align_corners);
      const float hf = floorf(fz);
      const float w1 = __fsub_rn(fz, hf);
      const float w0 = __fsub_rn(1.0f, w1);
      int r0 = (int)hf - hlo;
      r0 = min(max(r0, 0), kTRows - 2);
      const float4* t0 = reinterpret_cast<const float4*>(sT + r0 * kTPitch);
      const float4* t1 = reinterpret_cast<const float4*>(sT + (r0 + 1) * kTPitch);
      const float4* cc = reinterpret_cast<const float4*>(sC);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const float4 a0 = t0[2 * c8], a1 = t0[2 * c8 + 1];
        const float4 b0 = t1[2 * c8], b1 = t1[2 * c8 + 1];
        const float4 c0 = cc[2 * c8], c1 = cc[2 * c8 + 1];
        float h[8];
        h[0] = c0.x + w0 * a0.x + w1 * b0.x;
        h[1] = c0.y + w0 * a0.y + w1 * b0.y;
        h[2] = c0.z + w0 * a0.z + w1 * b0.z;
        h[3] = c0.w + w0 * a0.w + w1 * b0.w;
        h[4] = c1.x + w0 * a1.x + w1 * b1.x;
        h[5] = c1.y + w0 * a1.y + w1 * b1.y;
        h[6] = c1.z + w0 * a1.z + w1 * b1.z;
        h[7] = c1.w + w0 * a1.w + w1 * b1.w;
        uint32_t q0 = pack_half2(silu_from_half_arg(h[0]), silu_from_half_arg(h[1]));
        uint32_t q1 = pack_half2(silu_from_half_arg(h[2]), silu_from_half_arg(h[3]));
        uint32_t q2 = pack_half2(silu_from_half_arg(h[4]), silu_from_half_arg(h[5]));
        uint32_t q3 = pack_half2(silu_from_half_arg(h[6]), silu_from_half_arg(h[7]));
        const uint32_t dst = a_row + (uint32_t)((c8 ^ (m & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(q0), "r"(q1), "r"(q2), "r"(q3)
                     : "memory");
      }
    }

    if (!weights_ready) {  // first tile only: weights must have landed before the first MMA
      mbar_wait(bar_w, 0);
      weights_ready = true;
    }

    // ---- layers 1 .. nh-1 (hidden, N=64) and the head (N=16) -----------------
    for (int l = 1; l <= nh; ++l) {
      const bool head = (l == nh);
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1 + wg, 128);
      if (tid == 0) {
        tc_fence_after();
        const uint64_t a_desc = umma_desc_k_sw128(a_addr);
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(head ? sWf : sW + (l - 1) * kWBytes));
        const uint32_t idesc = head ? idesc_head : idesc_hidden;
        const uint32_t d_tmem = tmem_base + (uint32_t)(wg * 64);
#pragma unroll
        for (int kc = 0; kc < kHid / 16; ++kc)  // K = 16 per instruction: +32 B along the swizzled row
          umma_f16_ss(d_tmem, a_desc + 2 * kc, b_desc + 2 * kc, idesc, kc > 0 ? 1u : 0u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();

      if (!head) {
        const float4* bl = reinterpret_cast<const float4*>(sBias + l * kHid);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32(tmem_acc + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const float4 b0 = bl[half * 8 + 2 * c8], b1 = bl[half * 8 + 2 * c8 + 1];
            const float h0 = __uint_as_float(r[8 * c8 + 0]) + b0.x;
            const float h1 = __uint_as_float(r[8 * c8 + 1]) + b0.y;
            const float h2 = __uint_as_float(r[8 * c8 + 2]) + b0.z;
            const float h3 = __uint_as_float(r[8 * c8 + 3]) + b0.w;
            const float h4 = __uint_as_float(r[8 * c8 + 4]) + b1.x;
            const float h5 = __uint_as_float(r[8 * c8 + 5]) + b1.y;
            const float h6 = __uint_as_float(r[8 * c8 + 6]) + b1.z;
            const float h7 = __uint_as_float(r[8 * c8 + 7]) + b1.w;
            uint32_t q0 = pack_half2(silu_from_half_arg(h0), silu_from_half_arg(h1));
            uint32_t q1 = pack_half2(silu_from_half_arg(h2), silu_from_half_arg(h3));
            uint32_t q2 = pack_half2(silu_from_half_arg(h4), silu_from_half_arg(h5));
            uint32_t q3 = pack_half2(silu_from_half_arg(h6), silu_from_half_arg(h7));
            const int chunk = half * 4 + c8;
            const uint32_t dst = a_row + (uint32_t)((chunk ^ (m & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(q0), "r"(q1), "r"(q2), "r"(q3)
                         : "memory");
          }
        }
      } else {
        uint32_t r[4];
        tmem_ld4(tmem_acc, r);
        tmem_ld_wait();
        const float d = __uint_as_float(r[0]) + sBiasF[0];
        if (m < nvalid) {
          const long long o = (line * p.R) + k0 + m;
          if (p.out_raw) p.out_raw[o] = d;
          p.out_act[o] = expf(__fadd_rn(d, p.density_bias));
        }
      }
    }
    // the next tile's T/c build and A writes are ordered after this tile's last
    // reads by the named barriers above; TMEM reuse is ordered by wait::ld + the
    // before_thread_sync fence issued ahead of the next MMA.
  }

  tc_fence_before();
  __syncthreads();
  if (tid_cta < 32) tmem_dealloc<kWG * 64>(tmem_base);
}

}  // namespace smb

using namespace smb;

extern "C" int smb_query_lattice_tc(const float* planes_q, const void* decoder_blob,
                                    const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                    const float* axis_u, int R, int x_begin, int nx, float* out_density_act,
                                    float* out_density, void* stream) {
  if (!planes_q || !decoder_blob || !layout || !cfg || !axis_u || !out_density_act) return SMB_ERR_BAD_ARG;
  if (R < 2 || nx < 0 || x_begin < 0 || x_begin + nx > R) return SMB_ERR_BAD_ARG;
  if (nx == 0) return SMB_OK;
  const int nh = (int)layout->n_hidden;
  if (nh < 2 || nh > kMaxHidden) return SMB_ERR_BAD_ARG;
  // rows of the (.,z) planes one 128-sample segment can touch (+2 for the taps, +1 slack)
  {
    double span = 127.0 * cfg->Hp / (double)(R - 1);
    int rows = (int)span + 3;
    if (rows > cfg->Hp + 2) rows = cfg->Hp + 2;  // rows -1 .. Hp (the two zero borders)
    if (rows > kTRows) return SMB_ERR_BAD_ARG;
  }
  constexpr int kWG = 4;
  const int wbytes = tc_weight_bytes(nh);
  const size_t smem = (size_t)((wbytes + 1023) / 1024) * 1024 + (size_t)kWG * kWgBytes + 8 * (1 + kWG) + 16;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  // layout contract: [hidden | head | bias_half | bias_final] contiguous in the blob
  if (layout->off_tc_final != layout->off_tc_hidden + (uint32_t)(nh - 1) * kWBytes ||
      layout->off_bias_half != layout->off_tc_final + kWFinalBytes ||
      layout->off_bias_final != layout->off_bias_half + (uint32_t)nh * kHid * 4)
    return SMB_ERR_BAD_ARG;

  cudaError_t e = cudaFuncSetAttribute(lattice_tc_kernel<kWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  TcParams p{};
  p.planes_q = planes_q;
  p.tc_weights = static_cast<const unsigned char*>(decoder_blob) + layout->off_tc_hidden;
  p.bias0_half = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_bias_half);
  p.axis_u = axis_u;
  p.R = R;
  p.x_begin = x_begin;
  p.nx = nx;
  p.H = cfg->Hp;
  p.W = cfg->Wp;
  p.align_corners = cfg->align_corners;
  p.n_hidden = nh;
  p.density_bias = cfg->density_bias;
  p.out_act = out_density_act;
  p.out_raw = out_density;
  const long long ntiles = (long long)nx * R * ((R + kTileM - 1) / kTileM);
  long long grid = (ntiles + kWG - 1) / kWG;
  if (grid > sms) grid = sms;
  lattice_tc_kernel<kWG><<<(unsigned)grid, kWG * 128, smem, (cudaStream_t)stream>>>(p);
  return smb_check(cudaGetLastError());
}

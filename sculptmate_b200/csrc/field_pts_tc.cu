// Field query at ARBITRARY positions on tcgen05 tensor cores: bilinear gathers from the
// three channels-last planes feed an MLP whose every layer is an UMMA with the activations
// kept on-chip.  One kernel serves both decoders of the path:
//   TripoSR  query_triplane + NeRFMLP          tsr/models/nerf_renderer.py:41-91, network_utils.py:116-124
//            120 -> 64 (x9, SiLU) -> 4          (colour query at mesh vertices, system.py:191-198)
//   SF3D     query_triplane + MaterialMLP      sf3d/system.py:170-198, sf3d/models/network.py:191-208
//            heads density + vertex_offset fused: 120 -> 128 -> 128 (block-diagonal) -> 4
// The MLP is described by a packed blob (smb_mlp_tc_pack_host): per layer an fp16 K-major
// 128B-swizzled UMMA image (K padded to 64 or 128, N padded to 16/64/128) and fp32 biases;
// layers followed by SiLU carry W/2, b/2 (silu(x) = h + h*tanh(h), h = x/2).
//
// Work decomposition: tile = 128 positions = M of one UMMA; a CTA holds 4 independent
// warpgroups, each owning a tile end to end (thread m gathers position m, writes row m of the
// fp16 A tile, reads TMEM lane m).  The gathers (120 x 16-byte loads per position, L2-resident
// planes) dominate; four warpgroups per SM keep enough of them in flight.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "field_common.cuh"
#include "ptx_sm100.cuh"

namespace smb {

constexpr int kPtsTile = 128;
constexpr int kPtsWG = 4;
constexpr int kABlock = kPtsTile * 64 * 2;  // one 64-wide K block of the A tile: 16 KB

struct PtsTcParams {
  const void* planes_cl;  // (3,H,W,40) fp32, or fp16 when planes_fp16
  int planes_fp16;
  int H, W, align_corners;
  PosScale ps;
  const unsigned char* blob;  // device copy of the packed MLP
  smb_mlp_tc_layout lay;
  const float* positions;  // (n,3)
  long long n;
  float out0_bias;  // added to output 0 before exp: renderer density_bias / head out_bias
  int sigmoid_vec;  // 1: o_vec_act = sigmoid(outputs 1..3) (TripoSR colour)
  float *o_raw0, *o_act0, *o_vec, *o_vec_act;
};

__device__ __forceinline__ void gather_to_a(const float* __restrict__ plane, int H, int W, int align, float u, float v,
                                            unsigned char* __restrict__ a_rowp, int m, int plane_idx) {
  // 40 channels of one plane -> 5 chunks of 8 halves at k = 40*plane_idx + 8q
  Tap2 tx = make_tap(u, W, align);
  Tap2 ty = make_tap(v, H, align);
  const float4* p00 = reinterpret_cast<const float4*>(plane + ((long long)ty.i0 * W + tx.i0) * kCp);
  const float4* p01 = reinterpret_cast<const float4*>(plane + ((long long)ty.i0 * W + tx.i1) * kCp);
  const float4* p10 = reinterpret_cast<const float4*>(plane + ((long long)ty.i1 * W + tx.i0) * kCp);
  const float4* p11 = reinterpret_cast<const float4*>(plane + ((long long)ty.i1 * W + tx.i1) * kCp);
  const float w00 = ty.w0 * tx.w0, w01 = ty.w0 * tx.w1, w10 = ty.w1 * tx.w0, w11 = ty.w1 * tx.w1;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const float4 a0 = __ldg(p00 + 2 * q), a1 = __ldg(p00 + 2 * q + 1);
    const float4 b0 = __ldg(p01 + 2 * q), b1 = __ldg(p01 + 2 * q + 1);
    const float4 c0 = __ldg(p10 + 2 * q), c1 = __ldg(p10 + 2 * q + 1);
    const float4 d0 = __ldg(p11 + 2 * q), d1 = __ldg(p11 + 2 * q + 1);
    const uint32_t q0 = pack_half2(a0.x * w00 + b0.x * w01 + c0.x * w10 + d0.x * w11, a0.y * w00 + b0.y * w01 + c0.y * w10 + d0.y * w11);
    const uint32_t q1 = pack_half2(a0.z * w00 + b0.z * w01 + c0.z * w10 + d0.z * w11, a0.w * w00 + b0.w * w01 + c0.w * w10 + d0.w * w11);
    const uint32_t q2 = pack_half2(a1.x * w00 + b1.x * w01 + c1.x * w10 + d1.x * w11, a1.y * w00 + b1.y * w01 + c1.y * w10 + d1.y * w11);
    const uint32_t q3 = pack_half2(a1.z * w00 + b1.z * w01 + c1.z * w10 + d1.z * w11, a1.w * w00 + b1.w * w01 + c1.w * w10 + d1.w * w11);
    const int idx = 5 * plane_idx + q;  // 16-byte chunk index along K
    *reinterpret_cast<uint4*>(a_rowp + (idx >> 3) * kABlock + (((idx & 7) ^ (m & 7)) << 4)) = make_uint4(q0, q1, q2, q3);
  }
}

// same from fp16 channels-last planes: half the L2 traffic and half the load instructions of the
// gather (which bounds this kernel); taps are widened to fp32, blended in fp32, rounded once
__device__ __forceinline__ void gather_to_a_h(const __half* __restrict__ plane, int H, int W, int align, float u, float v,
                                              unsigned char* __restrict__ a_rowp, int m, int plane_idx) {
  Tap2 tx = make_tap(u, W, align);
  Tap2 ty = make_tap(v, H, align);
  const uint4* p00 = reinterpret_cast<const uint4*>(plane + ((long long)ty.i0 * W + tx.i0) * kCp);
  const uint4* p01 = reinterpret_cast<const uint4*>(plane + ((long long)ty.i0 * W + tx.i1) * kCp);
  const uint4* p10 = reinterpret_cast<const uint4*>(plane + ((long long)ty.i1 * W + tx.i0) * kCp);
  const uint4* p11 = reinterpret_cast<const uint4*>(plane + ((long long)ty.i1 * W + tx.i1) * kCp);
  const float w00 = ty.w0 * tx.w0, w01 = ty.w0 * tx.w1, w10 = ty.w1 * tx.w0, w11 = ty.w1 * tx.w1;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const uint4 a = __ldg(p00 + q), b = __ldg(p01 + q), c = __ldg(p10 + q), d = __ldg(p11 + q);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&av[e]));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bv[e]));
      const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&cv[e]));
      const float2 fd = __half22float2(*reinterpret_cast<const __half2*>(&dv[e]));
      o[e] = pack_half2(fa.x * w00 + fb.x * w01 + fc.x * w10 + fd.x * w11, fa.y * w00 + fb.y * w01 + fc.y * w10 + fd.y * w11);
    }
    const int idx = 5 * plane_idx + q;
    *reinterpret_cast<uint4*>(a_rowp + (idx >> 3) * kABlock + (((idx & 7) ^ (m & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(kPtsWG * 128, 1) points_tc_kernel(PtsTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const smb_mlp_tc_layout& L = p.lay;
  unsigned char* sBlob = smem;
  unsigned char* wg_region = smem + ((L.total_bytes + 1023) / 1024) * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wg_region + kPtsWG * 2 * kABlock);  // [0] weights, [1+wg] mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kPtsWG);

  const int tid_cta = threadIdx.x;
  const int wg = tid_cta >> 7;
  const int m = tid_cta & 127;
  const int q = (tid_cta >> 5) & 3;
  unsigned char* sA = wg_region + wg * 2 * kABlock;
  unsigned char* a_rowp = sA + (m >> 3) * 1024 + (m & 7) * 128;
  const uint32_t a_addr = smem_u32(sA);
  const uint32_t bar_w = smem_u32(&bars[0]);
  const uint32_t bar_mma = smem_u32(&bars[1 + wg]);

  if (tid_cta == 0) {
    mbar_init(bar_w, 1);
    for (int g = 0; g < kPtsWG; ++g) mbar_init(smem_u32(&bars[1 + g]), 1);
    mbar_fence_init();
  }
  if (tid_cta < 32) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid_cta == 0) {
    mbar_expect_tx(bar_w, L.total_bytes);
    uint32_t off = 0;
    while (off < L.total_bytes) {
      const uint32_t nb = min(8192u, L.total_bytes - off);
      bulk_g2s(smem_u32(sBlob + off), p.blob + off, nb, bar_w);
      off += nb;
    }
  }
  const uint32_t d_tmem = tmem_base + (uint32_t)(wg * 128);
  const uint32_t tmem_acc = d_tmem + ((uint32_t)(q * 32) << 16);
  uint32_t phase = 0;
  bool weights_ready = false;
  const long long psz = (long long)p.H * p.W * kCp;
  const long long ntiles = (p.n + kPtsTile - 1) / kPtsTile;

  for (long long t = (long long)blockIdx.x * kPtsWG + wg; t < ntiles; t += (long long)gridDim.x * kPtsWG) {
    const long long s = min(t * kPtsTile + m, p.n - 1);  // tail rows recompute the last position (not stored)
    const bool live = t * kPtsTile + m < p.n;
    {
      const float ux = scale_pos(p.positions[3 * s + 0], p.ps);
      const float uy = scale_pos(p.positions[3 * s + 1], p.ps);
      const float uz = scale_pos(p.positions[3 * s + 2], p.ps);
      if (p.planes_fp16) {
        const __half* pl = static_cast<const __half*>(p.planes_cl);
        gather_to_a_h(pl + 0 * psz, p.H, p.W, p.align_corners, ux, uy, a_rowp, m, 0);  // (x,y)
        gather_to_a_h(pl + 1 * psz, p.H, p.W, p.align_corners, ux, uz, a_rowp, m, 1);  // (x,z)
        gather_to_a_h(pl + 2 * psz, p.H, p.W, p.align_corners, uy, uz, a_rowp, m, 2);  // (y,z)
      } else {
        const float* pl = static_cast<const float*>(p.planes_cl);
        gather_to_a(pl + 0 * psz, p.H, p.W, p.align_corners, ux, uy, a_rowp, m, 0);  // (x,y)
        gather_to_a(pl + 1 * psz, p.H, p.W, p.align_corners, ux, uz, a_rowp, m, 1);  // (x,z)
        gather_to_a(pl + 2 * psz, p.H, p.W, p.align_corners, uy, uz, a_rowp, m, 2);  // (y,z)
      }
      *reinterpret_cast<uint4*>(a_rowp + kABlock + ((7 ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);  // k = 120..127
    }
    if (!weights_ready) {
      mbar_wait(bar_w, 0);
      weights_ready = true;
    }
    for (uint32_t l = 0; l < L.n_layers; ++l) {
      const uint32_t kb = L.kblocks[l], N = L.n_out[l];
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1 + wg, 128);
      if (m == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16_f32(128, N);
        for (uint32_t b = 0; b < kb; ++b) {
          const uint64_t a_desc = umma_desc_k_sw128(a_addr + b * kABlock);
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sBlob + L.w_off[l] + b * N * 128));
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) umma_f16_ss(d_tmem, a_desc + 2 * kc, b_desc + 2 * kc, idesc, (b | kc) ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      const float* bias = reinterpret_cast<const float*>(sBlob + L.b_off[l]);
      if (l + 1 < L.n_layers) {
        // hidden layer: N columns -> bias + SiLU -> the next A tile (N/64 K blocks)
        for (uint32_t c = 0; c < N / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(tmem_acc + c * 16, r);
          tmem_ld_wait();
          const float4* bl = reinterpret_cast<const float4*>(bias + c * 16);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = bl[i];
            pk[2 * i + 0] = pack_half2(silu_from_half_arg(__uint_as_float(r[4 * i + 0]) + b4.x),
                                       silu_from_half_arg(__uint_as_float(r[4 * i + 1]) + b4.y));
            pk[2 * i + 1] = pack_half2(silu_from_half_arg(__uint_as_float(r[4 * i + 2]) + b4.z),
                                       silu_from_half_arg(__uint_as_float(r[4 * i + 3]) + b4.w));
          }
          const uint32_t idx = 2 * c;  // 16-byte chunk index along the next layer's K
          unsigned char* blk = a_rowp + (idx >> 3) * kABlock;
          *reinterpret_cast<uint4*>(blk + (((idx & 7) ^ (m & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + ((((idx + 1) & 7) ^ (m & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      } else {
        uint32_t r[4];
        tmem_ld4(tmem_acc, r);
        tmem_ld_wait();
        if (live) {
          const long long o = t * kPtsTile + m;
          const float d = __fadd_rn(__uint_as_float(r[0]) + bias[0], p.out0_bias);
          const float a = __uint_as_float(r[1]) + bias[1], b = __uint_as_float(r[2]) + bias[2], c = __uint_as_float(r[3]) + bias[3];
          if (p.o_raw0) p.o_raw0[o] = d;
          if (p.o_act0) p.o_act0[o] = expf(d);
          if (p.o_vec) {
            p.o_vec[3 * o + 0] = a;
            p.o_vec[3 * o + 1] = b;
            p.o_vec[3 * o + 2] = c;
          }
          if (p.o_vec_act && p.sigmoid_vec) {
            p.o_vec_act[3 * o + 0] = 1.0f / (1.0f + expf(-a));
            p.o_vec_act[3 * o + 1] = 1.0f / (1.0f + expf(-b));
            p.o_vec_act[3 * o + 2] = 1.0f / (1.0f + expf(-c));
          }
        }
      }
    }
    // the next tile's gathers overwrite the A tile: this tile's last MMA has completed (acc waited above)
  }
  tc_fence_before();
  __syncthreads();
  if (tid_cta < 32) tmem_dealloc<512>(tmem_base);
}

}  // namespace smb

using namespace smb;

// ------------------------------------------------------------------ host: MLP packing
extern "C" int smb_mlp_tc_layout_for(int n_layers, const int* k_in, const int* n_out, smb_mlp_tc_layout* out) {
  if (!out || !k_in || !n_out || n_layers < 2 || n_layers > SMB_MLP_TC_MAX_LAYERS) return SMB_ERR_BAD_ARG;
  memset(out, 0, sizeof(*out));
  out->n_layers = (uint32_t)n_layers;
  uint32_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    const int kb = (k_in[l] + 63) / 64;
    int N = n_out[l];
    N = last ? 16 : (N <= 64 ? 64 : 128);
    if (kb < 1 || kb > 2 || n_out[l] < 1 || n_out[l] > (last ? 4 : 128)) return SMB_ERR_BAD_ARG;
    if (l > 0 && (uint32_t)kb * 64 != out->n_out[l - 1]) return SMB_ERR_BAD_ARG;  // activations feed the next K exactly
    out->kblocks[l] = (uint32_t)kb;
    out->n_out[l] = (uint32_t)N;
    out->w_off[l] = off;
    off += (uint32_t)kb * N * 128;
  }
  for (int l = 0; l < n_layers; ++l) {
    out->b_off[l] = off;
    off += out->n_out[l] * 4;
  }
  out->total_bytes = (off + 15u) & ~15u;
  return SMB_OK;
}

extern "C" int smb_mlp_tc_pack_host(const float* const* W, const float* const* B, const int* k_in, const int* n_out,
                                    const smb_mlp_tc_layout* L, void* blob_host) {
  if (!W || !B || !k_in || !n_out || !L || !blob_host) return SMB_ERR_BAD_ARG;
  unsigned char* blob = static_cast<unsigned char*>(blob_host);
  memset(blob, 0, L->total_bytes);
  for (uint32_t l = 0; l < L->n_layers; ++l) {
    if (!W[l] || !B[l]) return SMB_ERR_BAD_ARG;
    const bool last = l == L->n_layers - 1;
    const float scale = last ? 1.0f : 0.5f;  // layers followed by SiLU carry W/2, b/2
    const uint32_t N = L->n_out[l];
    for (int n = 0; n < n_out[l]; ++n)
      for (int k = 0; k < k_in[l]; ++k) {
        const __half h = __float2half_rn(scale * W[l][(size_t)n * k_in[l] + k]);
        memcpy(blob + L->w_off[l] + (size_t)(k / 64) * N * 128 + sw128_offset((uint32_t)n, (uint32_t)(k % 64)), &h, 2);
      }
    float* b = reinterpret_cast<float*>(blob + L->b_off[l]);
    for (int n = 0; n < n_out[l]; ++n) b[n] = scale * B[l][n];
  }
  return SMB_OK;
}

// NCHW fp32 (3,Cp,H,W) -> channels-last fp16 (3,H,W,Cp); tiled through shared memory like planes_to_channels_last (field_f32.cu)
constexpr int kClhTexels = 128;
__global__ void __launch_bounds__(256) planes_to_channels_last_half(const float* __restrict__ src, __half* __restrict__ dst, int H, int W) {
  __shared__ float s[smb::kCp][kClhTexels + 1];
  const int p = blockIdx.y, HW = H * W, hw0 = blockIdx.x * kClhTexels;
  const int nt = min(kClhTexels, HW - hw0);
  for (int t = threadIdx.x; t < smb::kCp * kClhTexels; t += blockDim.x) {
    const int c = t / kClhTexels, x = t - c * kClhTexels;
    if (x < nt) s[c][x] = __ldg(src + ((long long)p * smb::kCp + c) * HW + hw0 + x);
  }
  __syncthreads();
  __half2* o = reinterpret_cast<__half2*>(dst + ((long long)p * HW + hw0) * smb::kCp);  // Cp is even: pairs never straddle a texel
  for (int t = threadIdx.x; t < nt * (smb::kCp / 2); t += blockDim.x) {
    const int x = t / (smb::kCp / 2), c = 2 * (t - x * (smb::kCp / 2));
    o[t] = __floats2half2_rn(s[c][x], s[c + 1][x]);
  }
}

extern "C" int smb_scene_prepare_half(const float* triplane, int Hp, int Wp, void* planes_cl_half, void* stream) {
  if (!triplane || !planes_cl_half || Hp <= 0 || Wp <= 0) return SMB_ERR_BAD_ARG;
  planes_to_channels_last_half<<<dim3((Hp * Wp + kClhTexels - 1) / kClhTexels, 3), 256, 0, (cudaStream_t)stream>>>(
      triplane, static_cast<__half*>(planes_cl_half), Hp, Wp);
  return smb_check(cudaGetLastError());
}

extern "C" int smb_query_points_tc(const void* planes_cl, int planes_fp16, int Hp, int Wp, int align_corners, const void* mlp_blob_dev,
                                   const smb_mlp_tc_layout* layout, float radius, float out0_bias, int sigmoid_vec,
                                   const float* positions, int64_t n, float* out0_raw, float* out0_act, float* out_vec,
                                   float* out_vec_act, void* stream) {
  if (!planes_cl || !mlp_blob_dev || !layout || Hp <= 0 || Wp <= 0 || n < 0) return SMB_ERR_BAD_ARG;
  if (n == 0) return SMB_OK;
  if (!positions) return SMB_ERR_BAD_ARG;
  if (layout->n_layers < 2 || layout->n_layers > SMB_MLP_TC_MAX_LAYERS || layout->kblocks[0] != 2) return SMB_ERR_BAD_ARG;
  PtsTcParams p{};
  p.planes_cl = planes_cl;
  p.planes_fp16 = planes_fp16;
  p.H = Hp;
  p.W = Wp;
  p.align_corners = align_corners;
  {
    const double r = (double)radius;
    p.ps.sub = (float)(-r);
    p.ps.div = (float)(r - (-r));
    p.ps.mul = (float)(1.0 - (-1.0));
    p.ps.add = -1.0f;
  }
  p.blob = static_cast<const unsigned char*>(mlp_blob_dev);
  p.lay = *layout;
  p.positions = positions;
  p.n = n;
  p.out0_bias = out0_bias;
  p.sigmoid_vec = sigmoid_vec;
  p.o_raw0 = out0_raw;
  p.o_act0 = out0_act;
  p.o_vec = out_vec;
  p.o_vec_act = out_vec_act;
  const size_t smem = (size_t)((layout->total_bytes + 1023) / 1024) * 1024 + (size_t)kPtsWG * 2 * kABlock + 8 * (1 + kPtsWG) + 16;
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(points_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (n + kPtsTile - 1) / kPtsTile;
  long long grid = (ntiles + kPtsWG - 1) / kPtsWG;
  if (grid > sms) grid = sms;
  points_tc_kernel<<<(unsigned)grid, kPtsWG * 128, smem, (cudaStream_t)stream>>>(p);
  return smb_check(cudaGetLastError());
}

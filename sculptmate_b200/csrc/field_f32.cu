// fp32 CUDA-core field kernels:
//   * scene preparation (channels-last planes, layer-0 projected planes),
//   * query_points_f32 / query_lattice_f32: one thread per sample, bilinear gather
//     from channels-last planes + the NeRFMLP chain in fp32 with weights staged in
//     shared memory.  This is the reference-precision path (arbitrary positions,
//     colour query, parity anchor for the tensor-core kernel), not the fast path.
//
// Reference semantics (relative to /root/reference):
//   TripoSR/tsr/models/nerf_renderer.py:41-91  (plane pairing, grid_sample, exp/sigmoid)
//   TripoSR/tsr/models/network_utils.py:116-124 (Linear/SiLU chain; out[0]=density, out[1:4]=features)
#include <cuda_runtime.h>
#include <math.h>

#include "field_common.cuh"

namespace smb {

// ---------------------------------------------------------------- prepare
constexpr int kClTexels = 128;
// NCHW (3,Cp,H,W) -> channels-last (3,H,W,Cp).  One CTA = 128 consecutive texels of one plane through shared memory:
// 512-byte coalesced row segments in, one contiguous 20 KB run out (the first version read with a stride of H*W floats
// between lanes: 181 us for the 70 MB SF3D triplane; this one is HBM-bound).
__global__ void __launch_bounds__(256) planes_to_channels_last(const float* __restrict__ src, float* __restrict__ dst, int H, int W) {
  __shared__ float s[kCp][kClTexels + 1];
  const int p = blockIdx.y, HW = H * W, hw0 = blockIdx.x * kClTexels;
  const int nt = min(kClTexels, HW - hw0);
  for (int t = threadIdx.x; t < kCp * kClTexels; t += blockDim.x) {
    const int c = t / kClTexels, x = t - c * kClTexels;
    if (x < nt) s[c][x] = __ldg(src + ((long long)p * kCp + c) * HW + hw0 + x);
  }
  __syncthreads();
  float* o = dst + ((long long)p * HW + hw0) * kCp;
  for (int t = threadIdx.x; t < nt * kCp; t += blockDim.x) {
    const int x = t / kCp, c = t - x * kCp;
    o[t] = s[c][x];
  }
}

// planes_q[p][h][w][n] = sum_c (W0/2)[n][p*Cp + c] * plane[p][c][h][w]   (fp32)
// One CTA = 64 consecutive texels of one plane: the (40 x 64) source tile is staged in shared memory with
// coalesced 256-byte rows, the plane's (64 x 40) weight slice transposed to [c][n]; thread (g, t) accumulates
// the 16 outputs n = 16g .. 16g+15 of texel t (source reads conflict-free, weight reads warp-uniform
// broadcasts) in the reference's summation order c = 0 .. 39 and writes them as four 16-byte stores.
constexpr int kProjTexels = 64;
__global__ void __launch_bounds__(256) project_planes(const float* __restrict__ src, const float* __restrict__ w0_half,
                                                      float* __restrict__ dst, int H, int W) {
  __shared__ float sW[kCp][kHid];
  __shared__ float sS[kCp][kProjTexels];
  const int p = blockIdx.y;
  const int HW = H * W;
  const int hw0 = blockIdx.x * kProjTexels;
  for (int t = threadIdx.x; t < kHid * kCp; t += blockDim.x) {
    const int n = t / kCp, c = t - n * kCp;
    sW[c][n] = w0_half[n * kFeat + p * kCp + c];
  }
  for (int t = threadIdx.x; t < kCp * kProjTexels; t += blockDim.x) {
    const int c = t / kProjTexels, x = t - c * kProjTexels;
    sS[c][x] = hw0 + x < HW ? src[((long long)p * kCp + c) * HW + hw0 + x] : 0.0f;
  }
  __syncthreads();
  const int x = threadIdx.x & (kProjTexels - 1);
  const int g = threadIdx.x >> 6;  // 0..3: outputs 16g .. 16g+15
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
#pragma unroll 8
  for (int c = 0; c < kCp; ++c) {
    const float v = sS[c][x];
    const float4* w4 = reinterpret_cast<const float4*>(&sW[c][16 * g]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 w = w4[j];
      acc[4 * j + 0] = fmaf(w.x, v, acc[4 * j + 0]);
      acc[4 * j + 1] = fmaf(w.y, v, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(w.z, v, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(w.w, v, acc[4 * j + 3]);
    }
  }
  if (hw0 + x < HW) {
    float4* o = reinterpret_cast<float4*>(dst + ((long long)p * HW + hw0 + x) * kHid + 16 * g);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
}

// ------------------------------------------------------------- fp32 query
struct F32Params {
  const float* planes_cl;  // (3,H,W,Cp)
  const float* wts;        // fp32 section of the decoder blob
  int n_hidden;
  int H, W, align_corners;
  float density_bias;
  PosScale ps;
  // arbitrary positions
  const float* positions;  // (n,3) or nullptr for lattice mode
  const float* features_in;  // (n,120): run the decoder on given features (NeRFMLP.forward) instead of gathering
  long long n;
  // lattice mode
  const float* axis_u;
  int R, x_begin;
  float *density, *features, *density_act, *color;
};

__device__ __forceinline__ void gather_plane(const float* __restrict__ plane, int H, int W, int align, float u,
                                             float v, float* __restrict__ f /*[kCp]*/) {
  // u -> W axis, v -> H axis (grid_sample's grid[...,0] is x/width)
  Tap2 tx = make_tap(u, W, align);
  Tap2 ty = make_tap(v, H, align);
  const float4* p00 = reinterpret_cast<const float4*>(plane + ((long long)ty.i0 * W + tx.i0) * kCp);
  const float4* p01 = reinterpret_cast<const float4*>(plane + ((long long)ty.i0 * W + tx.i1) * kCp);
  const float4* p10 = reinterpret_cast<const float4*>(plane + ((long long)ty.i1 * W + tx.i0) * kCp);
  const float4* p11 = reinterpret_cast<const float4*>(plane + ((long long)ty.i1 * W + tx.i1) * kCp);
  const float w00 = ty.w0 * tx.w0, w01 = ty.w0 * tx.w1, w10 = ty.w1 * tx.w0, w11 = ty.w1 * tx.w1;
#pragma unroll
  for (int c4 = 0; c4 < kCp / 4; ++c4) {
    float4 a = __ldg(p00 + c4), b = __ldg(p01 + c4), c = __ldg(p10 + c4), d = __ldg(p11 + c4);
    f[4 * c4 + 0] = a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11;
    f[4 * c4 + 1] = a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11;
    f[4 * c4 + 2] = a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11;
    f[4 * c4 + 3] = a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11;
  }
}

__device__ __forceinline__ float silu_f32(float x) { return x / (1.0f + expf(-x)); }

// One thread per sample.  Weights live in shared memory (<= 163 KB for 9 hidden
// layers); all lanes read the same weight -> broadcast LDS.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) query_f32_kernel(F32Params p) {
  extern __shared__ __align__(16) float sw[];
  // fp32 section layout: for l in 0..n_hidden: W_l (out,in) then b_l (out)
  int total = kHid * kFeat + kHid;
  total += (p.n_hidden - 1) * (kHid * kHid + kHid);
  total += kOut * kHid + kOut;
  for (int t = threadIdx.x; t < total; t += kThreads) sw[t] = p.wts[t];
  __syncthreads();

  const long long RR = (long long)p.R * p.R;
  for (long long s = blockIdx.x * (long long)kThreads + threadIdx.x; s < p.n;
       s += (long long)gridDim.x * kThreads) {
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (p.features_in) {
    } else if (p.positions) {
      ux = scale_pos(p.positions[3 * s + 0], p.ps);
      uy = scale_pos(p.positions[3 * s + 1], p.ps);
      uz = scale_pos(p.positions[3 * s + 2], p.ps);
    } else {
      int i = (int)(s / RR);
      int rem = (int)(s - (long long)i * RR);
      int j = rem / p.R, k = rem - j * p.R;
      ux = p.axis_u[p.x_begin + i];
      uy = p.axis_u[j];
      uz = p.axis_u[k];
    }
    float act[kHid];
    float nxt[kHid];
    {
      float f[kFeat];
      if (p.features_in) {
        const float4* src = reinterpret_cast<const float4*>(p.features_in + s * kFeat);
#pragma unroll
        for (int k4 = 0; k4 < kFeat / 4; ++k4) {
          const float4 q = __ldg(src + k4);
          f[4 * k4 + 0] = q.x;
          f[4 * k4 + 1] = q.y;
          f[4 * k4 + 2] = q.z;
          f[4 * k4 + 3] = q.w;
        }
      } else {
        const long long psz = (long long)p.H * p.W * kCp;
        gather_plane(p.planes_cl + 0 * psz, p.H, p.W, p.align_corners, ux, uy, f + 0 * kCp);
        gather_plane(p.planes_cl + 1 * psz, p.H, p.W, p.align_corners, ux, uz, f + 1 * kCp);
        gather_plane(p.planes_cl + 2 * psz, p.H, p.W, p.align_corners, uy, uz, f + 2 * kCp);
      }
      const float* W = sw;
      const float* b = sw + kHid * kFeat;
      for (int n = 0; n < kHid; ++n) {
        float acc = b[n];
        const float4* wr = reinterpret_cast<const float4*>(W + n * kFeat);
#pragma unroll
        for (int k4 = 0; k4 < kFeat / 4; ++k4) {
          float4 w = wr[k4];
          acc = fmaf(w.x, f[4 * k4 + 0], acc);
          acc = fmaf(w.y, f[4 * k4 + 1], acc);
          acc = fmaf(w.z, f[4 * k4 + 2], acc);
          acc = fmaf(w.w, f[4 * k4 + 3], acc);
        }
        nxt[n] = silu_f32(acc);
      }
#pragma unroll
      for (int n = 0; n < kHid; ++n) act[n] = nxt[n];
    }
    const float* base = sw + kHid * kFeat + kHid;
    for (int l = 1; l < p.n_hidden; ++l) {
      const float* W = base;
      const float* b = base + kHid * kHid;
      for (int n = 0; n < kHid; ++n) {
        float acc = b[n];
        const float4* wr = reinterpret_cast<const float4*>(W + n * kHid);
#pragma unroll
        for (int k4 = 0; k4 < kHid / 4; ++k4) {
          float4 w = wr[k4];
          acc = fmaf(w.x, act[4 * k4 + 0], acc);
          acc = fmaf(w.y, act[4 * k4 + 1], acc);
          acc = fmaf(w.z, act[4 * k4 + 2], acc);
          acc = fmaf(w.w, act[4 * k4 + 3], acc);
        }
        nxt[n] = silu_f32(acc);
      }
#pragma unroll
      for (int n = 0; n < kHid; ++n) act[n] = nxt[n];
      base += kHid * kHid + kHid;
    }
    float o[kOut];
    {
      const float* W = base;
      const float* b = base + kOut * kHid;
#pragma unroll
      for (int n = 0; n < kOut; ++n) {
        float acc = b[n];
        const float4* wr = reinterpret_cast<const float4*>(W + n * kHid);
#pragma unroll
        for (int k4 = 0; k4 < kHid / 4; ++k4) {
          float4 w = wr[k4];
          acc = fmaf(w.x, act[4 * k4 + 0], acc);
          acc = fmaf(w.y, act[4 * k4 + 1], acc);
          acc = fmaf(w.z, act[4 * k4 + 2], acc);
          acc = fmaf(w.w, act[4 * k4 + 3], acc);
        }
        o[n] = acc;
      }
    }
    if (p.density) p.density[s] = o[0];
    if (p.density_act) p.density_act[s] = expf(__fadd_rn(o[0], p.density_bias));
    if (p.features) {
      p.features[3 * s + 0] = o[1];
      p.features[3 * s + 1] = o[2];
      p.features[3 * s + 2] = o[3];
    }
    if (p.color) {
      p.color[3 * s + 0] = 1.0f / (1.0f + expf(-o[1]));
      p.color[3 * s + 1] = 1.0f / (1.0f + expf(-o[2]));
      p.color[3 * s + 2] = 1.0f / (1.0f + expf(-o[3]));
    }
  }
}

static int f32_smem_bytes(int n_hidden) {
  int total = kHid * kFeat + kHid + (n_hidden - 1) * (kHid * kHid + kHid) + kOut * kHid + kOut;
  return total * (int)sizeof(float);
}

static int launch_f32(const F32Params& p, cudaStream_t st) {
  constexpr int kThreads = 128;
  int smem = f32_smem_bytes(p.n_hidden);
  if (smem > 227 * 1024) return SMB_ERR_BAD_ARG;
  cudaError_t e =
      cudaFuncSetAttribute(query_f32_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (p.n + kThreads - 1) / kThreads;
  if (blocks > sms) blocks = sms;  // persistent: weights are staged once per CTA
  if (blocks < 1) return SMB_OK;
  query_f32_kernel<kThreads><<<(unsigned)blocks, kThreads, smem, st>>>(p);
  return smb_check(cudaGetLastError());
}

}  // namespace smb

using namespace smb;

extern "C" int smb_scene_prepare(const float* triplane, int Hp, int Wp, const void* decoder_blob,
                                 const smb_decoder_layout* layout, float* planes_cl, float* planes_q,
                                 void* stream) {
  if (!triplane || Hp <= 0 || Wp <= 0 || (!planes_cl && !planes_q)) return SMB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes_cl) {
    planes_to_channels_last<<<dim3((Hp * Wp + kClTexels - 1) / kClTexels, 3), 256, 0, st>>>(triplane, planes_cl, Hp, Wp);
    if (cudaGetLastError() != cudaSuccess) return SMB_ERR_CUDA;
  }
  if (planes_q) {
    if (!decoder_blob || !layout) return SMB_ERR_BAD_ARG;
    const float* w0h = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_w0_half);
    project_planes<<<dim3((Hp * Wp + kProjTexels - 1) / kProjTexels, 3), 256, 0, st>>>(triplane, w0h, planes_q, Hp, Wp);
    if (cudaGetLastError() != cudaSuccess) return SMB_ERR_CUDA;
  }
  return SMB_OK;
}

static PosScale make_pos_scale(float radius) {
  // scale_tensor(positions, (-r, r), (-1, 1)) with python-float scales (utils.py:228-230)
  PosScale ps;
  double r = (double)radius;
  ps.sub = (float)(-r);
  ps.div = (float)(r - (-r));
  ps.mul = (float)(1.0 - (-1.0));
  ps.add = -1.0f;
  return ps;
}

extern "C" int smb_query_points_f32(const float* planes_cl, const void* decoder_blob,
                                    const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                    const float* positions, int64_t n, float* density, float* features,
                                    float* density_act, float* color, void* stream) {
  if (!planes_cl || !decoder_blob || !layout || !cfg || n < 0) return SMB_ERR_BAD_ARG;
  if (n == 0) return SMB_OK;
  if (!positions) return SMB_ERR_BAD_ARG;
  F32Params p{};
  p.planes_cl = planes_cl;
  p.wts = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_f32);
  p.n_hidden = (int)layout->n_hidden;
  p.H = cfg->Hp;
  p.W = cfg->Wp;
  p.align_corners = cfg->align_corners;
  p.density_bias = cfg->density_bias;
  p.ps = make_pos_scale(cfg->radius);
  p.positions = positions;
  p.n = n;
  p.R = 1;
  p.density = density;
  p.features = features;
  p.density_act = density_act;
  p.color = color;
  return launch_f32(p, (cudaStream_t)stream);
}

extern "C" int smb_query_lattice_f32(const float* planes_cl, const void* decoder_blob,
                                     const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                     const float* axis_u, int R, int x_begin, int nx, float* out_density_act,
                                     float* out_density, void* stream) {
  if (!planes_cl || !decoder_blob || !layout || !cfg || !axis_u || R <= 0 || nx < 0 || x_begin < 0 ||
      x_begin + nx > R)
    return SMB_ERR_BAD_ARG;
  if (nx == 0) return SMB_OK;
  F32Params p{};
  p.planes_cl = planes_cl;
  p.wts = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_f32);
  p.n_hidden = (int)layout->n_hidden;
  p.H = cfg->Hp;
  p.W = cfg->Wp;
  p.align_corners = cfg->align_corners;
  p.density_bias = cfg->density_bias;
  p.ps = make_pos_scale(cfg->radius);
  p.positions = nullptr;
  p.n = (long long)nx * R * R;
  p.axis_u = axis_u;
  p.R = R;
  p.x_begin = x_begin;
  p.density = out_density;
  p.density_act = out_density_act;
  return launch_f32(p, (cudaStream_t)stream);
}

// NeRFMLP.forward (network_utils.py:116-124) on pre-computed features: (n,120) -> density (n), features (n,3)
extern "C" int smb_decoder_forward_f32(const void* decoder_blob, const smb_decoder_layout* layout, const float* features_in,
                                       int64_t n, float* density, float* features, void* stream) {
  if (!decoder_blob || !layout || n < 0 || (!density && !features)) return SMB_ERR_BAD_ARG;
  if (n == 0) return SMB_OK;
  if (!features_in) return SMB_ERR_BAD_ARG;
  F32Params p{};
  p.wts = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_f32);
  p.n_hidden = (int)layout->n_hidden;
  p.features_in = features_in;
  p.n = n;
  p.R = 1;
  p.density = density;
  p.features = features;
  return launch_f32(p, (cudaStream_t)stream);
}

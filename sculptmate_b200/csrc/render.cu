// Volume rendering of the triplane field: the two elementwise stages of
// TriplaneNeRFRenderer._forward (/root/reference/TripoSR/tsr/models/nerf_renderer.py:93-152) around the
// field query -- sample positions along the rays (:106-117) and alpha compositing (:125-150).  The query in
// between is the tensor-core points kernel (field_pts_tc.cu) or the fp32 kernel.  Both stages are
// HBM-streaming: 12 B written per sample, 16 B read per sample.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sculptmate_b200.h"

namespace smb {

// xyz[r][s] = rays_o[r] + z * rays_d[r],  z = t_near[r] * (1 - t_mid[s]) + t_far[r] * t_mid[s]
// with the reference's fp32 operation order (every product and sum rounded separately, no FMA contraction),
// so the positions are bit-identical to nerf_renderer.py:113-117.
__global__ void __launch_bounds__(256) ray_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float* __restrict__ t_near, const float* __restrict__ t_far,
                                                         const float* __restrict__ t_mid, long long n_rays, int S,
                                                         float* __restrict__ pos) {
  const long long total = n_rays * S;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / S;
    const int s = (int)(t - r * S);
    const float tm = __ldg(t_mid + s);
    const float z = __fadd_rn(__fmul_rn(__ldg(t_near + r), __fsub_rn(1.0f, tm)), __fmul_rn(__ldg(t_far + r), tm));
    float* o = pos + 3 * t;
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = __fadd_rn(__ldg(rays_o + 3 * r + a), __fmul_rn(z, __ldg(rays_d + 3 * r + a)));
  }
}

// One warp per ray.  alpha_s = 1 - exp(-delta_s * sigma_s); T_s = prod_{j<s} (1 - alpha_j + 1e-10);
// w_s = alpha_s * T_s; rgb = sum_s w_s c_s + (1 - sum_s w_s)   (white background, :146-148).
// Samples are taken 32 at a time: multiplicative warp scan inside the group, running carry across groups.
__global__ void __launch_bounds__(256) ray_composite_kernel(const float* __restrict__ sigma, const float* __restrict__ color,
                                                            const float* __restrict__ deltas, long long n_rays, int S,
                                                            float* __restrict__ comp_rgb, float* __restrict__ opacity) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_rays; r += nwarps) {
    float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, op = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const bool ok = s < S;
      float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (ok) {
        const long long i = r * S + s;
        a = __fsub_rn(1.0f, expf(__fmul_rn(-__ldg(deltas + s), __ldg(sigma + i))));
        c0 = __ldg(color + 3 * i + 0);
        c1 = __ldg(color + 3 * i + 1);
        c2 = __ldg(color + 3 * i + 2);
      }
      const float f = ok ? __fadd_rn(__fsub_rn(1.0f, a), 1e-10f) : 1.0f;
      float inc = f;  // inclusive product over the lanes
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, inc, sft);
        if (lane >= sft) inc *= y;
      }
      float excl = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) excl = 1.0f;
      const float w = a * (carry * excl);
      cr = fmaf(w, c0, cr);
      cg = fmaf(w, c1, cg);
      cb = fmaf(w, c2, cb);
      op += w;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      cr += __shfl_xor_sync(0xffffffffu, cr, sft);
      cg += __shfl_xor_sync(0xffffffffu, cg, sft);
      cb += __shfl_xor_sync(0xffffffffu, cb, sft);
      op += __shfl_xor_sync(0xffffffffu, op, sft);
    }
    if (lane == 0) {
      const float bg = 1.0f - op;
      comp_rgb[3 * r + 0] = cr + bg;
      comp_rgb[3 * r + 1] = cg + bg;
      comp_rgb[3 * r + 2] = cb + bg;
      if (opacity) opacity[r] = op;
    }
  }
}

static unsigned render_grid(long long threads) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (threads + 255) / 256;
  const long long cap = (long long)sms * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace smb

using namespace smb;

extern "C" int smb_ray_sample_positions(const float* rays_o, const float* rays_d, const float* t_near, const float* t_far,
                                        const float* t_mid, int64_t n_rays, int n_samples, float* positions, void* stream) {
  if (n_rays < 0 || n_samples <= 0) return SMB_ERR_BAD_ARG;
  if (n_rays == 0) return SMB_OK;
  if (!rays_o || !rays_d || !t_near || !t_far || !t_mid || !positions) return SMB_ERR_BAD_ARG;
  ray_sample_kernel<<<render_grid((long long)n_rays * n_samples), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_near, t_far, t_mid, n_rays,
                                                                                                  n_samples, positions);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

extern "C" int smb_ray_composite(const float* density_act, const float* color, const float* deltas, int64_t n_rays, int n_samples,
                                 float* comp_rgb, float* opacity, void* stream) {
  if (n_rays < 0 || n_samples <= 0) return SMB_ERR_BAD_ARG;
  if (n_rays == 0) return SMB_OK;
  if (!density_act || !color || !deltas || !comp_rgb) return SMB_ERR_BAD_ARG;
  ray_composite_kernel<<<render_grid((long long)n_rays * 32), 256, 0, (cudaStream_t)stream>>>(density_act, color, deltas, n_rays, n_samples,
                                                                                               comp_rgb, opacity);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

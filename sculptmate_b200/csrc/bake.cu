// UV-space texture baking: TextureBaker.rasterize / interpolate
// (/root/reference/StableFast/sf3d/texture_baker/baker.py:12-118).  The reference calls a Windows-only DLL without
// source (texture_baker.dll: rasterize_cpu / interpolate_cpu); the functions of the same name that ship beside it
// in Python (texture_baker/common.py:104-230) are what this restates:
//   rasterize_cpu : texel (y, x) is the point (x / W, 1 - y / H); the texel takes the triangle that contains it
//                   (u, v, w >= 0, barycentric_coordinates(), common.py:104-121, fp32) and stores (u, v, w, triangle index), or (0, 0, 0, -1).
//   interpolate_cpu: out = attr[i0] * u + attr[i1] * v + attr[i2] * w in fp32, zero where no triangle.
// The Python version finds the triangle through a BVH and returns the first hit of its traversal; here every
// triangle rasterises its own bounding box (one warp per triangle, lanes stride over the box) and a texel covered by
// several triangles -- only possible on shared edges / vertices of a UV atlas -- goes to the LOWEST triangle index
// (atomicMin), which makes the result deterministic.  Integer/bit work + a few flops per covered texel: HBM-bound
// (16 B written per texel).
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "../../include/sculptmate_b200.h"

namespace smb {

struct Bary {
  float u, v, w;
  bool ok;
};

// common.py:104-121.  The Python code runs on numpy float32 scalars (tb_float2 members taken from the float32 UV
// array; the texel point's Python floats are weak scalars under NumPy >= 2 and adopt float32), i.e. in fp32 with every
// operation rounded separately -- the arithmetic a C++ `float` implementation behind the DLL would perform too.
__device__ __forceinline__ Bary barycentric(float px, float py, float x0, float y0, float x1, float y1, float x2, float y2) {
  const float ax = __fsub_rn(x1, x0), ay = __fsub_rn(y1, y0);  // v0v1
  const float bx = __fsub_rn(x2, x0), by = __fsub_rn(y2, y0);  // v0v2
  const float qx = __fsub_rn(px, x0), qy = __fsub_rn(py, y0);  // pv0
  const float d00 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
  const float d01 = __fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by));
  const float d11 = __fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by));
  const float d20 = __fadd_rn(__fmul_rn(qx, ax), __fmul_rn(qy, ay));
  const float d21 = __fadd_rn(__fmul_rn(qx, bx), __fmul_rn(qy, by));
  const float denom = __fsub_rn(__fmul_rn(d00, d11), __fmul_rn(d01, d01));
  Bary r;
  r.ok = denom != 0.0f;  // a degenerate triangle covers nothing (numpy would produce inf / nan here, never >= 0 for all three)
  r.v = __fdiv_rn(__fsub_rn(__fmul_rn(d11, d20), __fmul_rn(d01, d21)), denom);
  r.w = __fdiv_rn(__fsub_rn(__fmul_rn(d00, d21), __fmul_rn(d01, d20)), denom);
  r.u = __fsub_rn(__fsub_rn(1.0f, r.v), r.w);
  return r;
}

// the texel's point: Python floats x / width and 1.0 - y / height (double), rounded to fp32 when they meet the fp32 UVs
__device__ __forceinline__ void texel_point(int x, int y, int res, float& px, float& py) {
  px = (float)__ddiv_rn((double)x, (double)res);
  py = (float)__dsub_rn(1.0, __ddiv_rn((double)y, (double)res));
}

__global__ void bake_owner_init(int* __restrict__ owner, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) owner[t] = INT_MAX;
}

// one warp per triangle
__global__ void __launch_bounds__(256) bake_owner_kernel(const float* __restrict__ uv, const int* __restrict__ faces, long long nverts,
                                                         long long nfaces, int res, int* __restrict__ owner) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long f = warp; f < nfaces; f += nwarps) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= nverts || i1 >= nverts || i2 >= nverts) continue;
    const float x0 = uv[2 * i0], y0 = uv[2 * i0 + 1], x1 = uv[2 * i1], y1 = uv[2 * i1 + 1], x2 = uv[2 * i2], y2 = uv[2 * i2 + 1];
    const double minx = fminf(x0, fminf(x1, x2)), maxx = fmaxf(x0, fmaxf(x1, x2));
    const double miny = fminf(y0, fminf(y1, y2)), maxy = fmaxf(y0, fmaxf(y1, y2));
    // texels whose point can lie in the box, one texel of slack on every side (the containment test decides)
    int xa = (int)floor(minx * res) - 1, xb = (int)ceil(maxx * res) + 1;
    int ya = (int)floor((1.0 - maxy) * res) - 1, yb = (int)ceil((1.0 - miny) * res) + 1;
    xa = max(xa, 0);
    ya = max(ya, 0);
    xb = min(xb, res - 1);
    yb = min(yb, res - 1);
    if (xa > xb || ya > yb) continue;
    const int bw = xb - xa + 1;
    const long long npx = (long long)bw * (yb - ya + 1);
    for (long long t = lane; t < npx; t += 32) {
      const int y = ya + (int)(t / bw), x = xa + (int)(t % bw);
      float px, py;
      texel_point(x, y, res, px, py);
      const Bary b = barycentric(px, py, x0, y0, x1, y1, x2, y2);
      if (b.ok && b.u >= 0.0f && b.v >= 0.0f && b.w >= 0.0f) atomicMin(&owner[(long long)y * res + x], (int)f);
    }
  }
}

__global__ void __launch_bounds__(256) bake_write_kernel(const float* __restrict__ uv, const int* __restrict__ faces, int res,
                                                         const int* __restrict__ owner, float4* __restrict__ rast) {
  const long long n = (long long)res * res;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int f = owner[t];
    float4 o = make_float4(0.f, 0.f, 0.f, -1.f);
    if (f != INT_MAX) {
      const int i0 = faces[3 * (long long)f], i1 = faces[3 * (long long)f + 1], i2 = faces[3 * (long long)f + 2];
      float px, py;
      texel_point((int)(t % res), (int)(t / res), res, px, py);
      const Bary b = barycentric(px, py, uv[2 * i0], uv[2 * i0 + 1], uv[2 * i1], uv[2 * i1 + 1], uv[2 * i2], uv[2 * i2 + 1]);
      o = make_float4(b.u, b.v, b.w, (float)f);
    }
    rast[t] = o;
  }
}

// common.py:214-230: attr[i0] * u + attr[i1] * v + attr[i2] * w, fp32, products and sums rounded separately
__global__ void __launch_bounds__(256) bake_interpolate_kernel(const float* __restrict__ attr, int C, const int* __restrict__ faces,
                                                               long long nfaces, const float4* __restrict__ rast, long long ntex,
                                                               float* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ntex; t += (long long)gridDim.x * blockDim.x) {
    const float4 r = rast[t];
    const long long f = r.w < 0.f ? -1 : (long long)r.w;
    if (f < 0 || f >= nfaces) {
      for (int c = 0; c < C; ++c) out[t * C + c] = 0.f;
      continue;
    }
    const long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    for (int c = 0; c < C; ++c) {
      const float a = __fmul_rn(__ldg(attr + i0 * C + c), r.x), b = __fmul_rn(__ldg(attr + i1 * C + c), r.y);
      out[t * C + c] = __fadd_rn(__fadd_rn(a, b), __fmul_rn(__ldg(attr + i2 * C + c), r.z));
    }
  }
}

static unsigned bake_grid(long long threads) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (threads + 255) / 256;
  const long long cap = (long long)sms * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace smb

using namespace smb;

extern "C" size_t smb_bake_workspace_bytes(int resolution) {
  return resolution > 0 ? sizeof(int) * (size_t)resolution * resolution : 0;
}

extern "C" int smb_bake_rasterize(const float* uv, const int32_t* faces, int64_t nverts, int64_t nfaces, int resolution, float* rast,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (resolution <= 0 || nverts < 0 || nfaces < 0 || nfaces > INT_MAX - 1 || !rast || !workspace) return SMB_ERR_BAD_ARG;
  if (workspace_bytes < smb_bake_workspace_bytes(resolution)) return SMB_ERR_WORKSPACE;
  if (nfaces > 0 && (!uv || !faces)) return SMB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int* owner = static_cast<int*>(workspace);
  const long long ntex = (long long)resolution * resolution;
  bake_owner_init<<<bake_grid(ntex), 256, 0, st>>>(owner, ntex);
  if (nfaces > 0) bake_owner_kernel<<<bake_grid(nfaces * 32), 256, 0, st>>>(uv, faces, nverts, nfaces, resolution, owner);
  bake_write_kernel<<<bake_grid(ntex), 256, 0, st>>>(uv, faces, resolution, owner, reinterpret_cast<float4*>(rast));
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

extern "C" int smb_bake_interpolate(const float* attr, int channels, const int32_t* faces, int64_t nfaces, const float* rast, int resolution,
                                    float* out, void* stream) {
  if (resolution <= 0 || channels <= 0 || channels > 16 || nfaces < 0 || !rast || !out) return SMB_ERR_BAD_ARG;
  if (nfaces > 0 && (!attr || !faces)) return SMB_ERR_BAD_ARG;
  const long long ntex = (long long)resolution * resolution;
  bake_interpolate_kernel<<<bake_grid(ntex), 256, 0, (cudaStream_t)stream>>>(attr, channels, faces, nfaces, reinterpret_cast<const float4*>(rast), ntex, out);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

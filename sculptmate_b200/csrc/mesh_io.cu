// Mesh hand-off: the arrays TSR.import_obj_blender
// (/root/reference/TripoSR/tsr/system.py:127-168) feeds to Blender, produced on the device so that the sink
// can use bulk foreach_set calls instead of per-element Python loops.
//
//   * loop colours: system.py:137-146 walks every polygon loop in Python and assigns
//     color_layer.data[idx].color = vertex_colors[loops[idx].vertex_index] with alpha 1 appended
//     (:133-135).  from_pydata numbers the loops 3*f + corner, so the layer is the gather
//     loop_colors[3*f + c] = (vertex_colors[faces[f][c]], 1) -- one coalesced streaming pass, HBM-bound:
//     8 B read + 16 B written per loop, the (V,3) colour table stays in L2.
//   * int32 faces: Blender stores loop vertex indices as int32 (MeshLoop.vertex_index); the API's LongTensor
//     (isosurface.py:50) is narrowed here when the sink wants to foreach_set them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sculptmate_b200.h"

namespace smb {

__global__ void __launch_bounds__(256) loop_colors_kernel(const float* __restrict__ colors, const long long* __restrict__ faces,
                                                          long long nverts, long long nloops, float alpha,
                                                          float4* __restrict__ out, int* __restrict__ bad) {
  for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < nloops; l += (long long)gridDim.x * blockDim.x) {
    const long long v = faces[l];
    float4 c = make_float4(0.f, 0.f, 0.f, alpha);
    if (v >= 0 && v < nverts) {
      c.x = __ldg(colors + 3 * v + 0);
      c.y = __ldg(colors + 3 * v + 1);
      c.z = __ldg(colors + 3 * v + 2);
    } else if (bad) {
      *bad = 1;  // an index outside the vertex table: reported by the host wrapper
    }
    out[l] = c;
  }
}

__global__ void __launch_bounds__(256) faces_i32_kernel(const long long* __restrict__ faces, long long n, int* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) out[t] = (int)faces[t];
}

static unsigned grid_for(long long n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (n + 255) / 256;
  const long long cap = (long long)sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace smb

using namespace smb;

extern "C" int smb_mesh_loop_colors(const float* vertex_colors, const int64_t* faces, int64_t nverts, int64_t ntris, float alpha,
                                    float* loop_colors, int* bad_index_flag, void* stream) {
  if (nverts < 0 || ntris < 0) return SMB_ERR_BAD_ARG;
  if (ntris == 0) return SMB_OK;
  if (!vertex_colors || !faces || !loop_colors) return SMB_ERR_BAD_ARG;
  const long long nloops = 3 * (long long)ntris;
  loop_colors_kernel<<<grid_for(nloops), 256, 0, (cudaStream_t)stream>>>(vertex_colors, reinterpret_cast<const long long*>(faces), nverts, nloops,
                                                                          alpha, reinterpret_cast<float4*>(loop_colors), bad_index_flag);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

// int32 -> int64 (the reference's LongTensor) for faces that crossed NVLink / were emitted as int32: 4 values per thread,
// one 16-byte load and two 16-byte stores
__global__ void __launch_bounds__(256) faces_i64_kernel(const int* __restrict__ in, long long n, long long* __restrict__ out) {
  const long long n4 = n >> 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(in) + t);
    reinterpret_cast<longlong2*>(out)[2 * t] = make_longlong2(v.x, v.y);
    reinterpret_cast<longlong2*>(out)[2 * t + 1] = make_longlong2(v.z, v.w);
  }
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < (n & 3)) out[4 * n4 + t] = in[4 * n4 + t];
}

extern "C" int smb_mesh_faces_i64(const int32_t* faces_i32, int64_t ntris, int64_t* faces, void* stream) {
  if (ntris < 0) return SMB_ERR_BAD_ARG;
  if (ntris == 0) return SMB_OK;
  if (!faces || !faces_i32 || ((uintptr_t)faces_i32 & 15) || ((uintptr_t)faces & 15)) return SMB_ERR_BAD_ARG;
  const long long n = 3 * (long long)ntris;
  faces_i64_kernel<<<grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(faces_i32, n, reinterpret_cast<long long*>(faces));
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

extern "C" int smb_mesh_faces_i32(const int64_t* faces, int64_t ntris, int32_t* faces_i32, void* stream) {
  if (ntris < 0) return SMB_ERR_BAD_ARG;
  if (ntris == 0) return SMB_OK;
  if (!faces || !faces_i32) return SMB_ERR_BAD_ARG;
  const long long n = 3 * (long long)ntris;
  faces_i32_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(faces), n, faces_i32);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

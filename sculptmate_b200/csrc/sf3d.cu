// Stable Fast 3D ("Pro") variant of the path: triplane query with align_corners=True,
// the two MaterialMLP heads triplane_to_meshes uses, and marching tetrahedra.
// Reference semantics (paths relative to /root/reference/StableFast):
//   sf3d/system.py:170-198          SF3D.query_triplane (plane pairing, grid_sample align_corners=True)
//   sf3d/models/network.py:148-208  MaterialMLP: per-head Linear/SiLU chains, + out_bias, output activation
//   sf3d/models/isosurface.py:24-229 MarchingTetrahedraHelper (deformation, edge numbering, face order)
//   sf3d/system.py:141-168          triplane_to_meshes
//
// fp32 CUDA-core kernels (the reference forces .float() here, system.py:191-192).  The
// marching-tetrahedra kernels are HBM-bound index work: the topology of the tet grid is
// static, so the reference's per-call sort/unique over edge pairs (isosurface.py:153-168)
// is replaced by a flag + prefix scan over the grid's PRE-SORTED unique edge list -- the
// k-th crossing edge in that list is vertex k, exactly the numbering torch.unique yields.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "field_common.cuh"  // also pulls in include/sculptmate_b200.h

namespace smb {

// ------------------------------------------------------------------ field query
constexpr int kHeadHid = 64;
// heads blob (fp32): for head in (density[out 1], vertex_offset[out 3]):
//   W0 (64,120) b0 (64) W1 (64,64) b1 (64) W2 (out,64) b2 (out), each head zero-padded to a multiple of 4 floats
constexpr int kHeadW0 = kHeadHid * kFeat, kHeadW1 = kHeadHid * kHeadHid;
constexpr int kHeadFloats0 = (kHeadW0 + kHeadHid + kHeadW1 + kHeadHid + 1 * kHeadHid + 1 + 3) / 4 * 4;  // padded: head 1 stays 16-byte aligned
constexpr int kHeadFloats1 = (kHeadW0 + kHeadHid + kHeadW1 + kHeadHid + 3 * kHeadHid + 3 + 3) / 4 * 4;
constexpr int kHeadsFloats = kHeadFloats0 + kHeadFloats1;

struct Sf3dParams {
  const float* planes_cl;  // (3,H,W,40)
  const float* heads;      // kHeadsFloats
  int H, W;
  PosScale ps;
  float out_bias;
  const float* positions;    // (n,3) or nullptr
  const float* features_in;  // (n,120) or nullptr
  long long n;
  float *features_out, *density_raw, *density_act, *vertex_offset;
};

__device__ __forceinline__ void gather_plane_ac(const float* __restrict__ plane, int H, int W, float u, float v,
                                                float* __restrict__ f /*[kCp]*/) {
  // u -> W axis, v -> H axis; align_corners=True (system.py:193), zero padding
  Tap2 tx = make_tap(u, W, 1);
  Tap2 ty = make_tap(v, H, 1);
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

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// one head: f[120] -> out[kOutN]; weights in shared memory (broadcast reads)
template <int kOutN>
__device__ __forceinline__ void run_head(const float* __restrict__ w, const float* __restrict__ f, float* __restrict__ out) {
  const float* W0 = w;
  const float* b0 = W0 + kHeadW0;
  const float* W1 = b0 + kHeadHid;
  const float* b1 = W1 + kHeadW1;
  const float* W2 = b1 + kHeadHid;
  const float* b2 = W2 + kOutN * kHeadHid;
  float h1[kHeadHid];
  for (int n = 0; n < kHeadHid; ++n) {
    float acc = b0[n];
    const float4* wr = reinterpret_cast<const float4*>(W0 + n * kFeat);
#pragma unroll
    for (int k4 = 0; k4 < kFeat / 4; ++k4) {
      const float4 q = wr[k4];
      acc = fmaf(q.x, f[4 * k4 + 0], acc);
      acc = fmaf(q.y, f[4 * k4 + 1], acc);
      acc = fmaf(q.z, f[4 * k4 + 2], acc);
      acc = fmaf(q.w, f[4 * k4 + 3], acc);
    }
    h1[n] = silu_exact(acc);
  }
  float h2[kHeadHid];
  for (int n = 0; n < kHeadHid; ++n) {
    float acc = b1[n];
    const float4* wr = reinterpret_cast<const float4*>(W1 + n * kHeadHid);
#pragma unroll
    for (int k4 = 0; k4 < kHeadHid / 4; ++k4) {
      const float4 q = wr[k4];
      acc = fmaf(q.x, h1[4 * k4 + 0], acc);
      acc = fmaf(q.y, h1[4 * k4 + 1], acc);
      acc = fmaf(q.z, h1[4 * k4 + 2], acc);
      acc = fmaf(q.w, h1[4 * k4 + 3], acc);
    }
    h2[n] = silu_exact(acc);
  }
#pragma unroll
  for (int n = 0; n < kOutN; ++n) {
    float acc = b2[n];
    const float4* wr = reinterpret_cast<const float4*>(W2 + n * kHeadHid);
#pragma unroll
    for (int k4 = 0; k4 < kHeadHid / 4; ++k4) {
      const float4 q = wr[k4];
      acc = fmaf(q.x, h2[4 * k4 + 0], acc);
      acc = fmaf(q.y, h2[4 * k4 + 1], acc);
      acc = fmaf(q.z, h2[4 * k4 + 2], acc);
      acc = fmaf(q.w, h2[4 * k4 + 3], acc);
    }
    out[n] = acc;
  }
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) sf3d_query_kernel(Sf3dParams p) {
  extern __shared__ __align__(16) float sw[];
  const bool want_heads = p.density_raw || p.density_act || p.vertex_offset;
  if (want_heads) {
    for (int t = threadIdx.x; t < kHeadsFloats; t += kThreads) sw[t] = p.heads[t];
  }
  __syncthreads();
  const long long psz = (long long)p.H * p.W * kCp;
  for (long long s = blockIdx.x * (long long)kThreads + threadIdx.x; s < p.n; s += (long long)gridDim.x * kThreads) {
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
      const float ux = scale_pos(p.positions[3 * s + 0], p.ps);
      const float uy = scale_pos(p.positions[3 * s + 1], p.ps);
      const float uz = scale_pos(p.positions[3 * s + 2], p.ps);
      gather_plane_ac(p.planes_cl + 0 * psz, p.H, p.W, ux, uy, f + 0 * kCp);  // (x,y)  system.py:186-189
      gather_plane_ac(p.planes_cl + 1 * psz, p.H, p.W, ux, uz, f + 1 * kCp);  // (x,z)
      gather_plane_ac(p.planes_cl + 2 * psz, p.H, p.W, uy, uz, f + 2 * kCp);  // (y,z)
    }
    if (p.features_out) {
      float4* dst = reinterpret_cast<float4*>(p.features_out + s * kFeat);
#pragma unroll
      for (int k4 = 0; k4 < kFeat / 4; ++k4) dst[k4] = make_float4(f[4 * k4], f[4 * k4 + 1], f[4 * k4 + 2], f[4 * k4 + 3]);
    }
    if (p.density_raw || p.density_act) {
      float d[1];
      run_head<1>(sw, f, d);
      const float v = __fadd_rn(d[0], p.out_bias);  // heads[name](x) + out_bias   network.py:203
      if (p.density_raw) p.density_raw[s] = v;
      if (p.density_act) p.density_act[s] = expf(v);  // trunc_exp forward = exp   network.py:85
    }
    if (p.vertex_offset) {
      float o[3];
      run_head<3>(sw + kHeadFloats0, f, o);
      p.vertex_offset[3 * s + 0] = o[0];
      p.vertex_offset[3 * s + 1] = o[1];
      p.vertex_offset[3 * s + 2] = o[2];
    }
  }
}

// ------------------------------------------------------------- marching tets
constexpr int kScan = 2048;  // items per count CTA (256 threads x 8)

struct MtetWs {
  uint32_t* evid;    // per edge: chunk-local vertex prefix | crossing << 31
  uint32_t* echunk;  // per edge chunk: total -> base
  uint32_t* tinfo;   // per tet: c1 local (12b) | c2 local (12b) << 12 | tetindex << 24
  uint32_t* t1chunk; // per tet chunk: #1-triangle tets -> base
  uint32_t* t2chunk; // per tet chunk: #2-triangle tets -> base
  long long* totals; // [0] nverts [1] n1 [2] n2
  size_t bytes;
};
__host__ __device__ inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ inline MtetWs mtet_carve(void* base, long long ne, long long nt) {
  MtetWs w;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* p = b ? b + off : nullptr;
    off += al256(bytes);
    return p;
  };
  const size_t ech = (size_t)((ne + kScan - 1) / kScan) + 1, tch = (size_t)((nt + kScan - 1) / kScan) + 1;
  w.evid = reinterpret_cast<uint32_t*>(take((size_t)ne * 4));
  w.echunk = reinterpret_cast<uint32_t*>(take(ech * 4));
  w.tinfo = reinterpret_cast<uint32_t*>(take((size_t)nt * 4));
  w.t1chunk = reinterpret_cast<uint32_t*>(take(tch * 4));
  w.t2chunk = reinterpret_cast<uint32_t*>(take(tch * 4));
  w.totals = reinterpret_cast<long long*>(take(4 * 8));
  w.bytes = off;
  return w;
}

__constant__ int c_tri_table[16][6] = {  // isosurface.py:30-55
    {-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},
    {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},
    {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},
    {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};
__constant__ int c_ntri_table[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};  // isosurface.py:57-63

// CTA-wide exclusive scan of one packed counter per thread (two 16-bit lanes are enough:
// a chunk holds 2048 items); returns this thread's exclusive base and the CTA total.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= s) inc += y;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
  total = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t w = s_warp[q];
    if (q < warp) base += w;
    total += w;
  }
  __syncthreads();
  return base + inc - v;
}

// K1: crossing flag per static edge + chunk-local prefix.  The index records are read COALESCED (lane-consecutive: edge
// q * 256 + tid of the chunk) and the flags pass through shared memory to the thread that scans 8 consecutive edges --
// reading 8 consecutive records per thread made every load instruction touch 32 different cache lines (129 us at 29 M
// edges; the stream itself is 36 us of HBM time).
__global__ void __launch_bounds__(256) mtet_edge_count(const float* __restrict__ sdf, const int2* __restrict__ edges, long long ne,
                                                       uint32_t* __restrict__ evid, uint32_t* __restrict__ echunk) {
  __shared__ uint32_t s_warp[8];
  __shared__ __align__(8) unsigned char s_flag[kScan];
  const long long c0 = (long long)blockIdx.x * kScan;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const long long e = c0 + q * 256 + threadIdx.x;
    unsigned char f = 0;
    if (e < ne) {
      const int2 ab = __ldg(edges + e);
      const bool oa = __ldg(sdf + ab.x) > 0.0f, ob = __ldg(sdf + ab.y) > 0.0f;  // occ_n = sdf_n > 0  :146
      f = oa != ob ? 1 : 0;                                                     // mask_edges        :158
    }
    s_flag[q * 256 + threadIdx.x] = f;
  }
  __syncthreads();
  const uint2 fw = *reinterpret_cast<const uint2*>(&s_flag[threadIdx.x * 8]);  // this thread's 8 consecutive flags, one per byte
  const uint32_t sum = __popc(fw.x) + __popc(fw.y);
  uint32_t total;
  uint32_t run = block_excl_scan(sum, s_warp, total);
  const long long e0 = c0 + threadIdx.x * 8;
  uint32_t out[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t f = ((k < 4 ? fw.x : fw.y) >> (8 * (k & 3))) & 1u;
    out[k] = run | (f << 31);
    run += f;
  }
  if (e0 + 8 <= ne) {  // evid is 16-byte aligned and chunks are multiples of 8
    uint4* o = reinterpret_cast<uint4*>(evid + e0);
    o[0] = make_uint4(out[0], out[1], out[2], out[3]);
    o[1] = make_uint4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (e0 + k < ne) evid[e0 + k] = out[k];
  }
  if (threadIdx.x == 0) echunk[blockIdx.x] = total;
}

// K2: tet code, triangle count class, chunk-local prefixes of the two classes (coalesced like K1)
__global__ void __launch_bounds__(256) mtet_tet_count(const float* __restrict__ sdf, const int4* __restrict__ tets, long long nt,
                                                      uint32_t* __restrict__ tinfo, uint32_t* __restrict__ t1chunk,
                                                      uint32_t* __restrict__ t2chunk) {
  __shared__ uint32_t s_warp[8];
  __shared__ __align__(8) unsigned char s_code[kScan];
  const long long c0 = (long long)blockIdx.x * kScan;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const long long t = c0 + q * 256 + threadIdx.x;
    unsigned char c = 0;
    if (t < nt) {
      const int4 v = __ldg(tets + t);
      c = (unsigned char)((__ldg(sdf + v.x) > 0.0f ? 1u : 0u) | (__ldg(sdf + v.y) > 0.0f ? 2u : 0u) |
                          (__ldg(sdf + v.z) > 0.0f ? 4u : 0u) | (__ldg(sdf + v.w) > 0.0f ? 8u : 0u));  // :182-183
    }
    s_code[q * 256 + threadIdx.x] = c;
  }
  __syncthreads();
  const uint2 cw = *reinterpret_cast<const uint2*>(&s_code[threadIdx.x * 8]);
  uint32_t code[8], cls[8], sum = 0;  // cls: 1 -> low half, 2 -> high half (packed 16+16)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    code[k] = ((k < 4 ? cw.x : cw.y) >> (8 * (k & 3))) & 0xffu;
    const int n = c_ntri_table[code[k]];  // code 0 (also the padding behind nt) has no triangle
    cls[k] = n == 1 ? 1u : (n == 2 ? 0x10000u : 0u);
    sum += cls[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan(sum, s_warp, total);
  const long long t0 = c0 + threadIdx.x * 8;
  uint32_t out[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    out[k] = (run & 0xfffu) | (((run >> 16) & 0xfffu) << 12) | (code[k] << 24);
    run += cls[k];
  }
  if (t0 + 8 <= nt) {
    uint4* o = reinterpret_cast<uint4*>(tinfo + t0);
    o[0] = make_uint4(out[0], out[1], out[2], out[3]);
    o[1] = make_uint4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (t0 + k < nt) tinfo[t0 + k] = out[k];
  }
  if (threadIdx.x == 0) {
    t1chunk[blockIdx.x] = total & 0xffffu;
    t2chunk[blockIdx.x] = total >> 16;
  }
}

// K3: one CTA: exclusive scans of the three chunk-total arrays (in place) + grand totals; 8 consecutive entries per thread
// and round (a round of 1024 entries cost five barriers: 37 us for the 38 K chunk totals of the n = 160 grid)
__global__ void __launch_bounds__(1024) mtet_totals(uint32_t* echunk, long long nech, uint32_t* t1chunk, uint32_t* t2chunk,
                                                    long long ntch, long long* totals, smb_mtet_counts* counts) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = 0; a < 3; ++a) {
    uint32_t* arr = a == 0 ? echunk : (a == 1 ? t1chunk : t2chunk);
    const long long n = a == 0 ? nech : ntch;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 8192) {
      const long long i0 = base + tid * 8;
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = i0 + k < n ? arr[i0 + k] : 0u;
        sum += v[k];
      }
      uint32_t inc = sum;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += y;
      }
      if (lane == 31) warp_sums[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        uint32_t ws = warp_sums[lane], winc = ws;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          uint32_t y = __shfl_up_sync(0xffffffffu, winc, s);
          if (lane >= s) winc += y;
        }
        warp_sums[lane] = winc - ws;
      }
      __syncthreads();
      uint32_t run = carry_s + warp_sums[warp] + inc - sum;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (i0 + k < n) arr[i0 + k] = run;
        run += v[k];
      }
      __syncthreads();
      if (tid == 1023) carry_s = run;
      __syncthreads();
    }
    if (tid == 0) totals[a] = carry_s;
    __syncthreads();
  }
  if (tid == 0) {
    counts->nverts = totals[0];
    counts->ntris = totals[1] + 2 * totals[2];
    counts->ntris1 = totals[1];
    counts->reserved = 0;
  }
}

struct MtetAffine {
  int on;
  float sub, div, mul[3], add[3];
};
// K4: vertices on crossing edges (isosurface.py:170-178, fp32 operation order kept)
__global__ void __launch_bounds__(256) mtet_emit_verts(const float* __restrict__ pos, const float* __restrict__ sdf,
                                                       const int2* __restrict__ edges, long long ne, const uint32_t* __restrict__ evid,
                                                       const uint32_t* __restrict__ echunk, float* __restrict__ verts, MtetAffine af) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
    const uint32_t r = __ldg(evid + e);
    if (!(r >> 31)) continue;
    const long long vid = (long long)__ldg(echunk + e / kScan) + (r & 0x7fffffffu);
    const int2 ab = __ldg(edges + e);
    const float s0 = __ldg(sdf + ab.x);
    const float s1n = -__ldg(sdf + ab.y);             // edges_to_interp_sdf[:, -1] *= -1
    const float den = __fadd_rn(s0, s1n);             // .sum(1)
    const float w0 = __fdiv_rn(s1n, den);             // flip(...) / denominator
    const float w1 = __fdiv_rn(s0, den);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p0 = __ldg(pos + 3LL * ab.x + c), p1 = __ldg(pos + 3LL * ab.y + c);
      float v = __fadd_rn(__fmul_rn(p0, w0), __fmul_rn(p1, w1));  // (edges * w).sum(1)
      // scale_tensor(v_pos, points_range, bbox) of the caller (sf3d/system.py:162-164, utils.py:222-231), same fp32 ops
      if (af.on) v = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, af.sub), af.div), af.mul[c]), af.add[c]);
      verts[3 * vid + c] = v;
    }
  }
}

// K5: faces; all 1-triangle tets first, then all 2-triangle tets, each in tet order (:187-201)
__global__ void __launch_bounds__(256) mtet_emit_faces(const int* __restrict__ tet_edges /*(T,6)*/, long long nt,
                                                       const uint32_t* __restrict__ tinfo, const uint32_t* __restrict__ t1chunk,
                                                       const uint32_t* __restrict__ t2chunk, const uint32_t* __restrict__ evid,
                                                       const uint32_t* __restrict__ echunk, const long long* __restrict__ totals,
                                                       long long* __restrict__ faces) {
  const long long n1 = totals[1];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nt; t += (long long)gridDim.x * blockDim.x) {
    const uint32_t info = __ldg(tinfo + t);
    const uint32_t code = info >> 24;
    const int ntri = c_ntri_table[code];
    if (ntri == 0) continue;
    long long slot;
    if (ntri == 1) slot = (long long)__ldg(t1chunk + t / kScan) + (info & 0xfffu);
    else slot = n1 + 2 * ((long long)__ldg(t2chunk + t / kScan) + ((info >> 12) & 0xfffu));
    long long vid[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int e = __ldg(tet_edges + 6 * t + k);
      const uint32_t r = __ldg(evid + e);
      vid[k] = (r >> 31) ? (long long)__ldg(echunk + e / kScan) + (r & 0x7fffffffu) : -1;  // mapping = -1 off the surface
    }
    for (int q = 0; q < 3 * ntri; ++q) {
      const int k = c_tri_table[code][q];
      long long v = vid[0];
#pragma unroll
      for (int kk = 1; kk < 6; ++kk) v = (k == kk) ? vid[kk] : v;
      faces[3 * slot + q] = v;
    }
  }
}

__global__ void mtet_deform_kernel(const float* __restrict__ base, const float* __restrict__ deform, float scale, long long n3,
                                   float* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n3; t += (long long)gridDim.x * blockDim.x)
    out[t] = __fadd_rn(base[t], __fmul_rn(scale, tanhf(deform[t])));  // grid + (range/res) * tanh(offset)  :106-113,210-213
}

}  // namespace smb

using namespace smb;

static int sf3d_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

extern "C" int smb_sf3d_heads_floats(void) { return kHeadsFloats; }

extern "C" int smb_sf3d_query_f32(const float* planes_cl, int Hp, int Wp, const float* heads_blob, float radius,
                                  float density_out_bias, const float* positions, const float* features_in, int64_t n,
                                  float* features_out, float* density_raw, float* density_act, float* vertex_offset,
                                  void* stream) {
  if (n < 0) return SMB_ERR_BAD_ARG;
  if (n == 0) return SMB_OK;
  const bool want_heads = density_raw || density_act || vertex_offset;
  if ((!positions && !features_in) || (positions && features_in)) return SMB_ERR_BAD_ARG;
  if (positions && (!planes_cl || Hp <= 0 || Wp <= 0)) return SMB_ERR_BAD_ARG;
  if (want_heads && !heads_blob) return SMB_ERR_BAD_ARG;
  if (!want_heads && !features_out) return SMB_ERR_BAD_ARG;
  Sf3dParams p{};
  p.planes_cl = planes_cl;
  p.heads = heads_blob;
  p.H = Hp;
  p.W = Wp;
  {
    const double r = (double)radius;  // scale_tensor(positions, (-r, r), (-1, 1))  system.py:181-183
    p.ps.sub = (float)(-r);
    p.ps.div = (float)(r - (-r));
    p.ps.mul = (float)(1.0 - (-1.0));
    p.ps.add = -1.0f;
  }
  p.out_bias = density_out_bias;
  p.positions = positions;
  p.features_in = features_in;
  p.n = n;
  p.features_out = features_out;
  p.density_raw = density_raw;
  p.density_act = density_act;
  p.vertex_offset = vertex_offset;
  constexpr int kThreads = 128;
  const int smem = want_heads ? kHeadsFloats * (int)sizeof(float) : 16;
  cudaError_t e = cudaFuncSetAttribute(sf3d_query_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return SMB_ERR_CUDA;
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)sf3d_sms() * 2;
  if (blocks > cap) blocks = cap;
  sf3d_query_kernel<kThreads><<<(unsigned)blocks, kThreads, smem, (cudaStream_t)stream>>>(p);
  return smb_check(cudaGetLastError());
}

extern "C" size_t smb_mtet_workspace_bytes(int64_t n_edges, int64_t n_tets) {
  if (n_edges <= 0 || n_tets <= 0) return 0;
  return mtet_carve(nullptr, n_edges, n_tets).bytes;
}

extern "C" int smb_mtet_count(const float* sdf, const int32_t* edges, int64_t n_edges, const int32_t* tets, int64_t n_tets,
                              void* workspace, size_t workspace_bytes, smb_mtet_counts* counts_dev, void* stream) {
  if (!sdf || !edges || !tets || !workspace || !counts_dev || n_edges <= 0 || n_tets <= 0) return SMB_ERR_BAD_ARG;
  MtetWs w = mtet_carve(workspace, n_edges, n_tets);
  if (w.bytes > workspace_bytes) return SMB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long ech = (n_edges + kScan - 1) / kScan, tch = (n_tets + kScan - 1) / kScan;
  if (ech > 0x7fffffffLL || tch > 0x7fffffffLL) return SMB_ERR_BAD_ARG;
  mtet_edge_count<<<(unsigned)ech, 256, 0, st>>>(sdf, reinterpret_cast<const int2*>(edges), n_edges, w.evid, w.echunk);
  mtet_tet_count<<<(unsigned)tch, 256, 0, st>>>(sdf, reinterpret_cast<const int4*>(tets), n_tets, w.tinfo, w.t1chunk, w.t2chunk);
  mtet_totals<<<1, 1024, 0, st>>>(w.echunk, ech, w.t1chunk, w.t2chunk, tch, w.totals, counts_dev);
  return smb_check(cudaGetLastError());
}

static int mtet_emit_impl(const float* positions, const float* sdf, const int32_t* edges, int64_t n_edges, const int32_t* tet_edges,
                          int64_t n_tets, const void* workspace, float* verts, int64_t* faces, const MtetAffine& af, void* stream) {
  if (!positions || !sdf || !edges || !tet_edges || !workspace || n_edges <= 0 || n_tets <= 0) return SMB_ERR_BAD_ARG;
  MtetWs w = mtet_carve(const_cast<void*>(workspace), n_edges, n_tets);
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = sf3d_sms() * 16;
  long long be = (n_edges + 255) / 256, bt = (n_tets + 255) / 256;
  if (be > cap) be = cap;
  if (bt > cap) bt = cap;
  if (verts)
    mtet_emit_verts<<<(unsigned)be, 256, 0, st>>>(positions, sdf, reinterpret_cast<const int2*>(edges), n_edges, w.evid, w.echunk, verts, af);
  if (faces)
    mtet_emit_faces<<<(unsigned)bt, 256, 0, st>>>(tet_edges, n_tets, w.tinfo, w.t1chunk, w.t2chunk, w.evid, w.echunk, w.totals,
                                                   reinterpret_cast<long long*>(faces));
  return smb_check(cudaGetLastError());
}

extern "C" int smb_mtet_emit(const float* positions, const float* sdf, const int32_t* edges, int64_t n_edges,
                             const int32_t* tet_edges, int64_t n_tets, const void* workspace, float* verts, int64_t* faces,
                             void* stream) {
  MtetAffine af{};
  return mtet_emit_impl(positions, sdf, edges, n_edges, tet_edges, n_tets, workspace, verts, faces, af, stream);
}

extern "C" int smb_mtet_emit_affine(const float* positions, const float* sdf, const int32_t* edges, int64_t n_edges,
                                    const int32_t* tet_edges, int64_t n_tets, const void* workspace, float* verts, int64_t* faces,
                                    const float* affine8, void* stream) {
  if (!affine8) return SMB_ERR_BAD_ARG;
  MtetAffine af{};
  af.on = 1;
  af.sub = affine8[0];
  af.div = affine8[1];
  for (int c = 0; c < 3; ++c) {
    af.mul[c] = affine8[2 + c];
    af.add[c] = affine8[5 + c];
  }
  return mtet_emit_impl(positions, sdf, edges, n_edges, tet_edges, n_tets, workspace, verts, faces, af, stream);
}

extern "C" int smb_mtet_deform(const float* base, const float* deform, float scale, int64_t n_vertices, float* out, void* stream) {
  if (!base || !deform || !out || n_vertices < 0) return SMB_ERR_BAD_ARG;
  if (n_vertices == 0) return SMB_OK;
  const long long n3 = 3LL * n_vertices;
  long long blocks = (n3 + 255) / 256;
  const long long cap = (long long)sf3d_sms() * 16;
  if (blocks > cap) blocks = cap;
  mtet_deform_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(base, deform, scale, n3, out);
  return smb_check(cudaGetLastError());
}

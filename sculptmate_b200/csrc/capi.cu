// Host side of the C ABI: status strings, decoder packing, lattice coordinates and
// the host-buffer extractor (the whole TSR.extract_mesh path for one scene code,
// /root/reference/TripoSR/tsr/system.py:171-200 minus the Blender import).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "field_common.cuh"
#include "ptx_sm100.cuh"

using namespace smb;

extern "C" const char* smb_status_string(int status) {
  switch (status) {
    case SMB_OK: return "ok";
    case SMB_ERR_CUDA: return "CUDA runtime error";
    case SMB_ERR_BAD_ARG: return "bad argument";
    case SMB_ERR_WORKSPACE: return "workspace too small";
    case SMB_ERR_ARCH: return "device is not sm_100 (B200)";
    case SMB_ERR_LEVEL_RANGE: return "Surface level must be within volume data range.";
    case SMB_ERR_NO_SURFACE: return "No surface found at the given iso value.";
    case SMB_ERR_CAPACITY: return "output buffers too small (sizes returned; call again with larger buffers)";
  }
  return "unknown status";
}

extern "C" int smb_version(void) { return 100; }

extern "C" int smb_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SMB_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return SMB_ERR_CUDA;
  return major == 10 ? SMB_OK : SMB_ERR_ARCH;
}

// ------------------------------------------------------------------ decoder
static uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

extern "C" int smb_decoder_layout_for(int n_hidden, smb_decoder_layout* out) {
  if (!out || n_hidden < 2 || n_hidden > kMaxHidden) return SMB_ERR_BAD_ARG;
  memset(out, 0, sizeof(*out));
  uint32_t off = 0;
  out->n_hidden = (uint32_t)n_hidden;
  // [hidden | head | bias_half | bias_final] must stay contiguous (field_tc.cu stages them with one bulk-copy run)
  out->off_tc_hidden = off;
  off += (uint32_t)(n_hidden - 1) * 8192u;
  out->off_tc_final = off;
  off += 2048u;
  out->off_bias_half = off;
  off += (uint32_t)n_hidden * kHid * 4u;
  out->off_bias_final = off;
  off += 16u;
  off = (off + 1023u) & ~1023u;
  out->off_tc_l0 = off;
  off += 16384u;
  out->off_w0_half = off;
  off += kHid * kFeat * 4u;
  out->off_f32 = align16(off);
  off = out->off_f32;
  off += (kHid * kFeat + kHid) * 4u;
  off += (uint32_t)(n_hidden - 1) * (kHid * kHid + kHid) * 4u;
  off += (kOut * kHid + kOut) * 4u;
  off = (off + 1023u) & ~1023u;
  out->off_tc_biasblk = off;
  off += (uint32_t)(n_hidden - 1) * 8192u;
  out->total_bytes = align16(off);
  return SMB_OK;
}

static void put_half(unsigned char* img, uint32_t byte_off, float v) {
  __half h = __float2half_rn(v);
  memcpy(img + byte_off, &h, 2);
}

extern "C" int smb_decoder_pack_host(const float* const* W, const float* const* B, int n_hidden,
                                     const smb_decoder_layout* L, void* blob_host) {
  if (!W || !B || !L || !blob_host || n_hidden != (int)L->n_hidden) return SMB_ERR_BAD_ARG;
  for (int l = 0; l <= n_hidden; ++l)
    if (!W[l] || !B[l]) return SMB_ERR_BAD_ARG;
  unsigned char* blob = static_cast<unsigned char*>(blob_host);
  memset(blob, 0, L->total_bytes);
  // hidden layers 1..n_hidden-1: B operand [n][k] = W_l[n][k] / 2, K-major SW128 image
  for (int l = 1; l < n_hidden; ++l) {
    unsigned char* img = blob + L->off_tc_hidden + (size_t)(l - 1) * 8192;
    for (int n = 0; n < kHid; ++n)
      for (int k = 0; k < kHid; ++k) put_half(img, sw128_offset(n, k), 0.5f * W[l][n * kHid + k]);
  }
  // bias K-blocks of the hidden layers: b_l/2 = hi + lo in fp16 at K rows 0 and 1
  for (int l = 1; l < n_hidden; ++l) {
    unsigned char* img = blob + L->off_tc_biasblk + (size_t)(l - 1) * 8192;
    for (int n = 0; n < kHid; ++n) {
      const float b = 0.5f * B[l][n];
      const __half hi = __float2half_rn(b);
      put_half(img, sw128_offset(n, 0), __half2float(hi));
      put_half(img, sw128_offset(n, 1), b - __half2float(hi));
    }
  }
  {  // head: rows 0..3 = W_L (not halved: no SiLU after it), rows 4..15 zero
    unsigned char* img = blob + L->off_tc_final;
    for (int n = 0; n < kOut; ++n)
      for (int k = 0; k < kHid; ++k) put_half(img, sw128_offset(n, k), W[n_hidden][n * kHid + k]);
  }
  {  // layer 0 padded to K = 128 as two 64-wide K blocks
    unsigned char* img = blob + L->off_tc_l0;
    for (int n = 0; n < kHid; ++n)
      for (int k = 0; k < kFeat; ++k)
        put_half(img + (size_t)(k / 64) * 8192, sw128_offset(n, k % 64), 0.5f * W[0][n * kFeat + k]);
  }
  float* bh = reinterpret_cast<float*>(blob + L->off_bias_half);
  for (int l = 0; l < n_hidden; ++l)
    for (int n = 0; n < kHid; ++n) bh[l * kHid + n] = 0.5f * B[l][n];
  float* bf = reinterpret_cast<float*>(blob + L->off_bias_final);
  for (int n = 0; n < kOut; ++n) bf[n] = B[n_hidden][n];
  float* w0h = reinterpret_cast<float*>(blob + L->off_w0_half);
  for (int t = 0; t < kHid * kFeat; ++t) w0h[t] = 0.5f * W[0][t];
  float* f = reinterpret_cast<float*>(blob + L->off_f32);
  for (int l = 0; l <= n_hidden; ++l) {
    const int out = (l == n_hidden) ? kOut : kHid;
    const int in = (l == 0) ? kFeat : kHid;
    memcpy(f, W[l], sizeof(float) * out * in);
    f += out * in;
    memcpy(f, B[l], sizeof(float) * out);
    f += out;
  }
  return SMB_OK;
}

// ------------------------------------------------------- lattice coordinates
// Per-axis sample coordinate of the R-lattice mapped to (-1,1): a scalar IEEE fp32
// restatement of what the reference computes with torch ops:
//   torch.linspace(0, 1, R)                              isosurface.py:30-32
//   scale_tensor(., (0,1), (-radius, radius))            system.py:177-181
//   scale_tensor(., (-radius, radius), (-1, 1))          nerf_renderer.py:52-54
// aten's CPU linspace evaluates  start + step*i  (i < R/2)  and  end - step*(R-1-i)  (else) with the
// multiply-add CONTRACTED into one fused operation (one rounding; step itself is rounded to fp32): fmaf
// below.  With separately rounded multiply and add the values differ by one ulp at many indices (this was the
// state of round 1).  scale_tensor is four separate elementwise aten ops, each rounded on its own.
// tests/test_capi_host.py asserts equality with torch for every R in 2..400 and the benchmark sizes.
extern "C" int smb_lattice_axis_host(int R, float radius, float* axis_u_host) {
  if (R < 1 || !axis_u_host) return SMB_ERR_BAD_ARG;
  const double r = (double)radius;
  const float a_sub = 0.0f, a_div = (float)(1.0 - 0.0), a_mul = (float)(r - (-r)), a_add = (float)(-r);
  const float b_sub = (float)(-r), b_div = (float)(r - (-r)), b_mul = (float)(1.0 - (-1.0)), b_add = -1.0f;
  const float step = R > 1 ? (1.0f - 0.0f) / (float)(R - 1) : 0.0f;
  for (int i = 0; i < R; ++i) {
    volatile float t;
    if (R == 1) {
      t = 0.0f;
    } else if (i < R / 2) {
      t = fmaf(step, (float)i, 0.0f);
    } else {
      t = fmaf(-step, (float)(R - 1 - i), 1.0f);
    }
    volatile float v = t - a_sub;
    v = v / a_div;
    v = v * a_mul;
    v = v + a_add;
    v = v - b_sub;
    v = v / b_div;
    v = v * b_mul;
    v = v + b_add;
    axis_u_host[i] = v;
  }
  return SMB_OK;
}

// --------------------------------------------------------------- extractor
struct smb_extractor {
  smb_decoder_layout layout;
  smb_query_cfg cfg;
  void* blob_dev = nullptr;
  float* triplane_dev = nullptr;
  float* planes_q = nullptr;
  float* axis_dev = nullptr;
  float* density = nullptr;
  void* mc_ws = nullptr;
  smb_mc_counts* counts_dev = nullptr;
  float* verts_dev = nullptr;
  int64_t* faces_dev = nullptr;
  float* minmax_dev = nullptr;
  // pinned host staging
  float* triplane_pin = nullptr;
  smb_mc_counts* counts_pin = nullptr;
  float* verts_pin = nullptr;
  int64_t* faces_pin = nullptr;
  float* minmax_pin = nullptr;
  size_t verts_cap = 0, faces_cap = 0, verts_pin_cap = 0, faces_pin_cap = 0;
  size_t mc_ws_bytes = 0;
  int res = 0;
  cudaStream_t stream = nullptr;
  // slab pipeline: the mesh of slab k goes back to the host while slab k+1 is computed
  static const int kSlabs = 8;  // maximum; n_slabs of them are used
  int64_t* slab_counts_dev = nullptr;       // (kSlabs,4) = smb_mc_counts per slab
  smb_mc_counts* slab_counts_pin = nullptr; // kSlabs
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t slab_done[kSlabs] = {};
  // colour query at the mesh vertices (enable_texture): tensor-core points kernel
  smb_mlp_tc_layout pts_layout;
  void* pts_blob_dev = nullptr;
  float* planes_cl = nullptr;
  float* colors_dev = nullptr;
  float* loop_colors_dev = nullptr;
  float* colors_pin = nullptr;
  float* loop_colors_pin = nullptr;
  size_t colors_cap = 0, loop_colors_cap = 0;
  // optional phase timing of smb_extract_mesh_device (smb_extractor_enable_timing): events on the caller's stream
  int timing = 0;
  cudaEvent_t tev[4] = {};  // start, after prepare, after the lattice kernel, after emit
  int faces_i32 = 0;  // host extractor delivers (F,3) int32 faces (smb_extractor_set_faces_i32)
  int n_slabs = 3;  // equal slabs, measured on B200 at 256^3: 2-3 slabs 4.42 ms, 1 (off) 4.71 ms, >= 4 slower and noisier (per-slab launch + host event cost)
  // slab k holds a share ~ slab_ratio^k of the cell layers: only the LAST slab's copy is exposed, and the copy of
  // slab k (PCIe, ~1/3 of its compute time) still hides behind the compute of the smaller slab k+1
  double slab_ratio = 0.45;
};

#define EX_CUDA(call)                        \
  do {                                       \
    if ((call) != cudaSuccess) return SMB_ERR_CUDA; \
  } while (0)

extern "C" int smb_extractor_create(const float* const* W, const float* const* B, int n_hidden, float radius,
                                    float density_bias, int Hp, int Wp, smb_extractor** out) {
  if (!out || Hp <= 0 || Wp <= 0) return SMB_ERR_BAD_ARG;
  int rc = smb_device_check();
  if (rc != SMB_OK) return rc;
  smb_extractor* ex = new smb_extractor();
  rc = smb_decoder_layout_for(n_hidden, &ex->layout);
  if (rc != SMB_OK) {
    delete ex;
    return rc;
  }
  std::vector<unsigned char> blob(ex->layout.total_bytes);
  rc = smb_decoder_pack_host(W, B, n_hidden, &ex->layout, blob.data());
  if (rc != SMB_OK) {
    delete ex;
    return rc;
  }
  {  // NeRFMLP as a tensor-core MLP for arbitrary positions (the colour query, system.py:191-198)
    std::vector<int> k_in(n_hidden + 1, kHid), n_out(n_hidden + 1, kHid);
    k_in[0] = kFeat;
    n_out[n_hidden] = kOut;
    rc = smb_mlp_tc_layout_for(n_hidden + 1, k_in.data(), n_out.data(), &ex->pts_layout);
    std::vector<unsigned char> pblob(rc == SMB_OK ? ex->pts_layout.total_bytes : 0);
    if (rc == SMB_OK) rc = smb_mlp_tc_pack_host(W, B, k_in.data(), n_out.data(), &ex->pts_layout, pblob.data());
    if (rc == SMB_OK && (cudaMalloc(&ex->pts_blob_dev, pblob.size()) != cudaSuccess ||
                         cudaMemcpy(ex->pts_blob_dev, pblob.data(), pblob.size(), cudaMemcpyHostToDevice) != cudaSuccess))
      rc = SMB_ERR_CUDA;
    if (rc != SMB_OK) {
      smb_extractor_destroy(ex);
      return rc;
    }
  }
  ex->cfg.radius = radius;
  ex->cfg.density_bias = density_bias;
  ex->cfg.align_corners = 0;
  ex->cfg.Hp = Hp;
  ex->cfg.Wp = Wp;
  const size_t tp_bytes = (size_t)3 * kCp * Hp * Wp * sizeof(float);
  bool ok = cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&ex->blob_dev, blob.size()) == cudaSuccess &&
            cudaMalloc(&ex->triplane_dev, tp_bytes) == cudaSuccess &&
            cudaMalloc(&ex->planes_q, (size_t)3 * Hp * Wp * kHid * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&ex->planes_cl, tp_bytes) == cudaSuccess &&
            cudaMalloc(&ex->counts_dev, sizeof(smb_mc_counts)) == cudaSuccess &&
            cudaMalloc(&ex->minmax_dev, 2 * sizeof(float)) == cudaSuccess &&
            cudaMallocHost(&ex->triplane_pin, tp_bytes) == cudaSuccess &&
            cudaMallocHost(&ex->counts_pin, sizeof(smb_mc_counts)) == cudaSuccess &&
            cudaMallocHost(&ex->minmax_pin, 2 * sizeof(float)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ex->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&ex->slab_counts_dev, sizeof(int64_t) * 4 * smb_extractor::kSlabs) == cudaSuccess &&
            cudaMallocHost(&ex->slab_counts_pin, sizeof(smb_mc_counts) * smb_extractor::kSlabs) == cudaSuccess &&
            cudaMemcpy(ex->blob_dev, blob.data(), blob.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  for (int k = 0; ok && k < smb_extractor::kSlabs; ++k)
    ok = cudaEventCreateWithFlags(&ex->slab_done[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    smb_extractor_destroy(ex);
    return SMB_ERR_CUDA;
  }
  if (const char* e = getenv("SMB_PIPE_SLABS")) {  // tuning knob: 0/1 disables the slab pipeline
    const int v = atoi(e);
    ex->n_slabs = v < 0 ? 0 : (v > smb_extractor::kSlabs ? smb_extractor::kSlabs : v);
  }
  if (const char* e = getenv("SMB_PIPE_RATIO")) {  // tuning knob: 1 = equal slabs
    const double v = atof(e);
    if (v > 0.05 && v <= 1.0) ex->slab_ratio = v;
  }
  *out = ex;
  return SMB_OK;
}

extern "C" void smb_extractor_destroy(smb_extractor* ex) {
  if (!ex) return;
  cudaFree(ex->blob_dev);
  cudaFree(ex->triplane_dev);
  cudaFree(ex->planes_q);
  cudaFree(ex->axis_dev);
  cudaFree(ex->density);
  cudaFree(ex->mc_ws);
  cudaFree(ex->counts_dev);
  cudaFree(ex->verts_dev);
  cudaFree(ex->faces_dev);
  cudaFree(ex->minmax_dev);
  cudaFreeHost(ex->triplane_pin);
  cudaFreeHost(ex->counts_pin);
  cudaFreeHost(ex->verts_pin);
  cudaFreeHost(ex->faces_pin);
  cudaFreeHost(ex->minmax_pin);
  cudaFree(ex->slab_counts_dev);
  cudaFreeHost(ex->slab_counts_pin);
  cudaFree(ex->pts_blob_dev);
  cudaFree(ex->planes_cl);
  cudaFree(ex->colors_dev);
  cudaFree(ex->loop_colors_dev);
  cudaFreeHost(ex->colors_pin);
  cudaFreeHost(ex->loop_colors_pin);
  for (int k = 0; k < smb_extractor::kSlabs; ++k)
    if (ex->slab_done[k]) cudaEventDestroy(ex->slab_done[k]);
  for (int k = 0; k < 4; ++k)
    if (ex->tev[k]) cudaEventDestroy(ex->tev[k]);
  if (ex->copy_stream) cudaStreamDestroy(ex->copy_stream);
  if (ex->stream) cudaStreamDestroy(ex->stream);
  delete ex;
}

extern "C" int smb_extractor_pinned_input(smb_extractor* ex, float** triplane_pinned) {
  if (!ex || !triplane_pinned) return SMB_ERR_BAD_ARG;
  *triplane_pinned = ex->triplane_pin;
  return SMB_OK;
}

static int ensure_resolution(smb_extractor* ex, int R) {
  // per-resolution workspaces are cached like the reference caches its helper
  // (system.py:118-124 set_marching_cubes_resolution)
  if (ex->res == R) return SMB_OK;
  cudaFree(ex->axis_dev);
  cudaFree(ex->density);
  cudaFree(ex->mc_ws);
  ex->axis_dev = nullptr;
  ex->density = nullptr;
  ex->mc_ws = nullptr;
  ex->res = 0;
  std::vector<float> axis(R);
  int rc = smb_lattice_axis_host(R, ex->cfg.radius, axis.data());
  if (rc != SMB_OK) return rc;
  ex->mc_ws_bytes = smb_mc_workspace_bytes(R, R, R);
  EX_CUDA(cudaMalloc(&ex->axis_dev, sizeof(float) * R));
  EX_CUDA(cudaMalloc(&ex->density, sizeof(float) * (size_t)R * R * R));
  EX_CUDA(cudaMalloc(&ex->mc_ws, ex->mc_ws_bytes));
  EX_CUDA(cudaMemcpy(ex->axis_dev, axis.data(), sizeof(float) * R, cudaMemcpyHostToDevice));
  ex->res = R;
  return SMB_OK;
}

extern "C" int smb_extractor_set_axis(smb_extractor* ex, int R, const float* axis_u_host) {
  if (!ex || !axis_u_host || R < 2) return SMB_ERR_BAD_ARG;
  int rc = ensure_resolution(ex, R);
  if (rc != SMB_OK) return rc;
  EX_CUDA(cudaMemcpy(ex->axis_dev, axis_u_host, sizeof(float) * R, cudaMemcpyHostToDevice));
  return SMB_OK;
}

// Slab pipeline (used once the buffers of a previous mesh exist): the lattice is cut into kSlabs x-slabs
// (one-plane halo recomputed, canonical order makes slabs concatenate bit-exactly); every slab runs
// lattice -> count -> emit in gather mode (offsets summed on the device from the earlier slabs' counts),
// all queued without a host round trip; the host only follows the slab events and streams each slab's
// part of the mesh to pinned memory on a second stream while the next slab is being computed.
// Returns 1 when the mesh was delivered, 0 when the caller must take the single-pass path
// (empty surface, or the mesh outgrew the remembered capacities), < 0 on error.
// dst_v / dst_f: pinned host destinations (capacities in vertices / triangles); null = the handle's own staging buffers.
static int extract_pipelined(smb_extractor* ex, int R, float threshold, int64_t* nverts, int64_t* ntris, int faces_i32,
                             float* dst_v = nullptr, size_t dst_v_cap = 0, void* dst_f = nullptr, size_t dst_f_cap = 0) {
  const int S = ex->n_slabs;
  const size_t fbytes = faces_i32 ? sizeof(int32_t) : sizeof(int64_t);  // device buffers are sized for int64 either way
  if (S < 2 || ex->verts_cap == 0 || ex->faces_cap == 0 || R - 1 < 8 * S) return 0;
  float* host_v = dst_v ? dst_v : ex->verts_pin;
  char* host_f = reinterpret_cast<char*>(dst_f ? dst_f : (void*)ex->faces_pin);
  const size_t host_v_cap = dst_v ? dst_v_cap : ex->verts_pin_cap, host_f_cap = dst_f ? dst_f_cap : ex->faces_pin_cap;
  cudaStream_t st = ex->stream;
  const double r = (double)ex->cfg.radius;
  const int flags = SMB_MC_FLIP | SMB_MC_DIV | SMB_MC_AFFINE | (faces_i32 ? SMB_MC_FACES_I32 : 0);
  const float vdiv = (float)(R - 1.0), vmul = (float)(r - (-r)), vadd = (float)(-r);
  // geometric slab sizes (at least 8 cell layers each), first slab largest
  const int cells = R - 1;
  int bounds[smb_extractor::kSlabs + 1];
  {
    double wsum = 0.0, w = 1.0;
    for (int k = 0; k < S; ++k, w *= ex->slab_ratio) wsum += w;
    double acc = 0.0;
    w = 1.0;
    bounds[0] = 0;
    for (int k = 0; k < S; ++k, w *= ex->slab_ratio) {
      acc += w;
      int b = (int)(cells * (acc / wsum) + 0.5);
      if (b < bounds[k] + 8) b = bounds[k] + 8;
      if (b > cells - 8 * (S - 1 - k)) b = cells - 8 * (S - 1 - k);
      bounds[k + 1] = b;
    }
    bounds[S] = cells;
  }
  int a = 0;
  for (int k = 0; k < S; ++k) {
    const int b = bounds[k + 1];
    const int nx = b - a + 1, last = k == S - 1;
    int rc = smb_query_lattice_tc_signs(ex->planes_q, ex->blob_dev, &ex->layout, &ex->cfg, ex->axis_dev, R, a, nx, ex->density, nullptr,
                                        threshold, 1.0f, ex->mc_ws, ex->mc_ws_bytes, st);
    if (rc != SMB_OK) return rc;
    smb_mc_counts* cnt = reinterpret_cast<smb_mc_counts*>(ex->slab_counts_dev + 4 * k);
    rc = smb_mc_count_presigned(nx, R, R, last, ex->mc_ws, ex->mc_ws_bytes, cnt, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaMemcpyAsync(&ex->slab_counts_pin[k], cnt, sizeof(smb_mc_counts), cudaMemcpyDeviceToHost, st));
    rc = smb_mc_emit_gather(ex->density, nx, R, R, threshold, 1.0f, a, last, flags, vdiv, vmul, vadd, ex->mc_ws, ex->slab_counts_dev, k,
                            ex->verts_dev, (int64_t)ex->verts_cap, ex->faces_dev, (int64_t)ex->faces_cap, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaEventRecord(ex->slab_done[k], st));
    a = b;
  }
  int64_t V = 0, F = 0;
  bool fits = true;
  for (int k = 0; k < S; ++k) {
    EX_CUDA(cudaEventSynchronize(ex->slab_done[k]));
    const int64_t Vk = ex->slab_counts_pin[k].nverts, Fk = ex->slab_counts_pin[k].ntris;
    fits = fits && (size_t)(V + Vk) <= ex->verts_cap && (size_t)(F + Fk) <= ex->faces_cap &&
           (size_t)(V + Vk) <= host_v_cap && (size_t)(F + Fk) <= host_f_cap;
    if (fits) {
      EX_CUDA(cudaStreamWaitEvent(ex->copy_stream, ex->slab_done[k], 0));
      if (Vk) EX_CUDA(cudaMemcpyAsync(host_v + 3 * V, ex->verts_dev + 3 * V, sizeof(float) * 3 * Vk, cudaMemcpyDeviceToHost, ex->copy_stream));
      if (Fk) EX_CUDA(cudaMemcpyAsync(host_f + fbytes * 3 * F, reinterpret_cast<char*>(ex->faces_dev) + fbytes * 3 * F,
                                      fbytes * 3 * Fk, cudaMemcpyDeviceToHost, ex->copy_stream));
    }
    V += Vk;
    F += Fk;
  }
  EX_CUDA(cudaStreamSynchronize(ex->copy_stream));
  *nverts = V;
  *ntris = F;
  if (!fits || V == 0 || F == 0) return 0;
  return 1;
}

extern "C" int smb_extract_mesh_host(smb_extractor* ex, const float* triplane_host, int R, float threshold,
                                     const float** verts_host, const int64_t** faces_host, int64_t* nverts,
                                     int64_t* ntris) {
  if (!ex || !triplane_host || R < 2 || !verts_host || !faces_host || !nverts || !ntris) return SMB_ERR_BAD_ARG;
  int rc = ensure_resolution(ex, R);
  if (rc != SMB_OK) return rc;
  cudaStream_t st = ex->stream;
  const size_t tp_bytes = (size_t)3 * kCp * ex->cfg.Hp * ex->cfg.Wp * sizeof(float);
  if (triplane_host != ex->triplane_pin) memcpy(ex->triplane_pin, triplane_host, tp_bytes);  // else: written in place
  EX_CUDA(cudaMemcpyAsync(ex->triplane_dev, ex->triplane_pin, tp_bytes, cudaMemcpyHostToDevice, st));
  rc = smb_scene_prepare(ex->triplane_dev, ex->cfg.Hp, ex->cfg.Wp, ex->blob_dev, &ex->layout, nullptr, ex->planes_q, st);
  if (rc != SMB_OK) return rc;
  rc = extract_pipelined(ex, R, threshold, nverts, ntris, ex->faces_i32);
  if (rc < 0) return rc;
  if (rc == 1) {
    *verts_host = ex->verts_pin;
    *faces_host = ex->faces_pin;
    return SMB_OK;
  }
  // system.py:184  helper(-(density - threshold))  ->  level = density - threshold, iso 0; the case bits are
  // balloted by the lattice kernel
  rc = smb_query_lattice_tc_signs(ex->planes_q, ex->blob_dev, &ex->layout, &ex->cfg, ex->axis_dev, R, 0, R, ex->density,
                                  nullptr, threshold, 1.0f, ex->mc_ws, ex->mc_ws_bytes, st);
  if (rc != SMB_OK) return rc;
  rc = smb_mc_count_presigned(R, R, R, 1, ex->mc_ws, ex->mc_ws_bytes, ex->counts_dev, st);
  if (rc != SMB_OK) return rc;
  // isosurface.py:52-53 (flip, /(R-1)) and system.py:185-189 (scale to +-radius)
  const double r = (double)ex->cfg.radius;
  const int mc_flags = SMB_MC_FLIP | SMB_MC_DIV | SMB_MC_AFFINE | (ex->faces_i32 ? SMB_MC_FACES_I32 : 0);
  const size_t fbytes = ex->faces_i32 ? sizeof(int32_t) : sizeof(int64_t);
  const float vdiv = (float)(R - 1.0), vmul = (float)(r - (-r)), vadd = (float)(-r);
  // emit right behind count into the buffers kept from the previous mesh (no host round trip);
  // the sizes are read afterwards and emit is repeated only if the mesh outgrew them
  const bool speculative = ex->verts_cap > 0 && ex->faces_cap > 0;
  if (speculative) {
    rc = smb_mc_emit_bounded(ex->density, R, R, R, threshold, 1.0f, 0, 1, mc_flags, vdiv, vmul, vadd, 0, ex->mc_ws, ex->verts_dev,
                             (int64_t)ex->verts_cap, ex->faces_dev, (int64_t)ex->faces_cap, st);
    if (rc != SMB_OK) return rc;
  }
  EX_CUDA(cudaMemcpyAsync(ex->counts_pin, ex->counts_dev, sizeof(smb_mc_counts), cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaStreamSynchronize(st));
  const int64_t V = ex->counts_pin->nverts, F = ex->counts_pin->ntris;
  if (V == 0 || F == 0) {
    rc = smb_grid_minmax(ex->density, (int64_t)R * R * R, threshold, 1.0f, ex->minmax_dev, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaMemcpyAsync(ex->minmax_pin, ex->minmax_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    EX_CUDA(cudaStreamSynchronize(st));
    *nverts = 0;
    *ntris = 0;
    if (ex->minmax_pin[0] > 0.0f || ex->minmax_pin[1] < 0.0f) return SMB_ERR_LEVEL_RANGE;
    return SMB_ERR_NO_SURFACE;
  }
  const bool fits = speculative && (size_t)V <= ex->verts_cap && (size_t)F <= ex->faces_cap;
  if ((size_t)V > ex->verts_cap) {
    cudaFree(ex->verts_dev);
    ex->verts_cap = 0;
    EX_CUDA(cudaMalloc(&ex->verts_dev, sizeof(float) * 3 * (size_t)V * 5 / 4));
    ex->verts_cap = (size_t)V * 5 / 4;
  }
  if ((size_t)F > ex->faces_cap) {
    cudaFree(ex->faces_dev);
    ex->faces_cap = 0;
    EX_CUDA(cudaMalloc(&ex->faces_dev, sizeof(int64_t) * 3 * (size_t)F * 5 / 4));
    ex->faces_cap = (size_t)F * 5 / 4;
  }
  if ((size_t)V > ex->verts_pin_cap) {
    cudaFreeHost(ex->verts_pin);
    ex->verts_pin_cap = 0;
    EX_CUDA(cudaMallocHost(&ex->verts_pin, sizeof(float) * 3 * (size_t)V * 5 / 4));
    ex->verts_pin_cap = (size_t)V * 5 / 4;
  }
  if ((size_t)F > ex->faces_pin_cap) {
    cudaFreeHost(ex->faces_pin);
    ex->faces_pin_cap = 0;
    EX_CUDA(cudaMallocHost(&ex->faces_pin, sizeof(int64_t) * 3 * (size_t)F * 5 / 4));
    ex->faces_pin_cap = (size_t)F * 5 / 4;
  }
  if (!fits) {
    rc = smb_mc_emit_bounded(ex->density, R, R, R, threshold, 1.0f, 0, 1, mc_flags, vdiv, vmul, vadd, 0, ex->mc_ws, ex->verts_dev,
                             (int64_t)ex->verts_cap, ex->faces_dev, (int64_t)ex->faces_cap, st);
    if (rc != SMB_OK) return rc;
  }
  EX_CUDA(cudaMemcpyAsync(ex->verts_pin, ex->verts_dev, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaMemcpyAsync(ex->faces_pin, ex->faces_dev, fbytes * 3 * F, cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaStreamSynchronize(st));
  *verts_host = ex->verts_pin;
  *faces_host = ex->faces_pin;
  *nverts = V;
  *ntris = F;
  return SMB_OK;
}

// enable_texture=True (system.py:190-200): mesh as above, then the colour query at the vertices on the tensor
// cores (positions = the vertices already in (-radius, radius), still on the device) and, optionally, the
// per-loop RGBA gather the Blender sink assigns.
extern "C" int smb_extract_mesh_host_textured(smb_extractor* ex, const float* triplane_host, int R, float threshold,
                                              const float** verts_host, const int64_t** faces_host,
                                              const float** colors_host, const float** loop_colors_host, int64_t* nverts,
                                              int64_t* ntris) {
  if (!colors_host) return SMB_ERR_BAD_ARG;
  int rc = smb_extract_mesh_host(ex, triplane_host, R, threshold, verts_host, faces_host, nverts, ntris);
  if (rc != SMB_OK) return rc;
  const int64_t V = *nverts, F = *ntris;
  cudaStream_t st = ex->stream;
  if ((size_t)V > ex->colors_cap) {
    cudaFree(ex->colors_dev);
    cudaFreeHost(ex->colors_pin);
    ex->colors_dev = ex->colors_pin = nullptr;
    ex->colors_cap = 0;
    const size_t cap = (size_t)V * 5 / 4;
    EX_CUDA(cudaMalloc(&ex->colors_dev, sizeof(float) * 3 * cap));
    EX_CUDA(cudaMallocHost(&ex->colors_pin, sizeof(float) * 3 * cap));
    ex->colors_cap = cap;
  }
  rc = smb_scene_prepare(ex->triplane_dev, ex->cfg.Hp, ex->cfg.Wp, nullptr, nullptr, ex->planes_cl, nullptr, st);
  if (rc != SMB_OK) return rc;
  rc = smb_query_points_tc(ex->planes_cl, 0, ex->cfg.Hp, ex->cfg.Wp, 0, ex->pts_blob_dev, &ex->pts_layout, ex->cfg.radius,
                           ex->cfg.density_bias, 1, ex->verts_dev, V, nullptr, nullptr, nullptr, ex->colors_dev, st);
  if (rc != SMB_OK) return rc;
  EX_CUDA(cudaMemcpyAsync(ex->colors_pin, ex->colors_dev, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
  if (loop_colors_host) {
    if ((size_t)F > ex->loop_colors_cap) {
      cudaFree(ex->loop_colors_dev);
      cudaFreeHost(ex->loop_colors_pin);
      ex->loop_colors_dev = ex->loop_colors_pin = nullptr;
      ex->loop_colors_cap = 0;
      const size_t cap = (size_t)F * 5 / 4;
      EX_CUDA(cudaMalloc(&ex->loop_colors_dev, sizeof(float) * 12 * cap));
      EX_CUDA(cudaMallocHost(&ex->loop_colors_pin, sizeof(float) * 12 * cap));
      ex->loop_colors_cap = cap;
    }
    rc = smb_mesh_loop_colors(ex->colors_dev, ex->faces_dev, V, F, 1.0f, ex->loop_colors_dev, nullptr, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaMemcpyAsync(ex->loop_colors_pin, ex->loop_colors_dev, sizeof(float) * 12 * F, cudaMemcpyDeviceToHost, st));
  }
  EX_CUDA(cudaStreamSynchronize(st));
  *colors_host = ex->colors_pin;
  if (loop_colors_host) *loop_colors_host = ex->loop_colors_pin;
  return SMB_OK;
}

// Device triplane in, mesh in the CALLER's pinned host buffers out: what the Python plugin call TSR.extract_mesh needs
// (tsr/system.py:171-200 ends in .cpu().numpy() of both arrays).  Uses the slab pipeline of smb_extract_mesh_host -- the mesh of
// slab k crosses PCIe while slab k+1 is computed -- but writes straight into buffers the caller owns (e.g. a torch tensor
// from the caching pinned allocator), so nothing aliases across calls and nothing is copied twice.  The handle keeps device
// staging buffers of at least the given capacities.  SMB_ERR_CAPACITY (+ sizes) when the mesh does not fit: retry with
// larger buffers.  Work runs on the handle's own streams, ordered after `stream` (where triplane_dev was produced).
extern "C" int smb_extract_mesh_device_to_host(smb_extractor* ex, const float* triplane_dev, int R, float threshold, int face_flags,
                                               float* verts_host_pinned, int64_t verts_capacity, void* faces_host_pinned,
                                               int64_t faces_capacity, void* stream, int64_t* nverts, int64_t* ntris) {
  if (!ex || !triplane_dev || R < 2 || !nverts || !ntris || verts_capacity < 0 || faces_capacity < 0) return SMB_ERR_BAD_ARG;
  if ((verts_capacity > 0 && !verts_host_pinned) || (faces_capacity > 0 && !faces_host_pinned)) return SMB_ERR_BAD_ARG;
  int rc = ensure_resolution(ex, R);
  if (rc != SMB_OK) return rc;
  const int i32 = (face_flags & SMB_MC_FACES_I32) ? 1 : 0;
  cudaStream_t st = ex->stream;
  *nverts = 0;
  *ntris = 0;
  // order the handle's stream after the producer of triplane_dev
  cudaEvent_t ready = ex->slab_done[smb_extractor::kSlabs - 1];
  EX_CUDA(cudaEventRecord(ready, (cudaStream_t)stream));
  EX_CUDA(cudaStreamWaitEvent(st, ready, 0));
  if ((size_t)verts_capacity > ex->verts_cap) {
    EX_CUDA(cudaStreamSynchronize(st));
    cudaFree(ex->verts_dev);
    ex->verts_cap = 0;
    EX_CUDA(cudaMalloc(&ex->verts_dev, sizeof(float) * 3 * (size_t)verts_capacity));
    ex->verts_cap = (size_t)verts_capacity;
  }
  if ((size_t)faces_capacity > ex->faces_cap) {
    EX_CUDA(cudaStreamSynchronize(st));
    cudaFree(ex->faces_dev);
    ex->faces_cap = 0;
    EX_CUDA(cudaMalloc(&ex->faces_dev, sizeof(int64_t) * 3 * (size_t)faces_capacity));
    ex->faces_cap = (size_t)faces_capacity;
  }
  rc = smb_scene_prepare(triplane_dev, ex->cfg.Hp, ex->cfg.Wp, ex->blob_dev, &ex->layout, nullptr, ex->planes_q, st);
  if (rc != SMB_OK) return rc;
  if (verts_capacity > 0 && faces_capacity > 0) {
    rc = extract_pipelined(ex, R, threshold, nverts, ntris, i32, verts_host_pinned, (size_t)verts_capacity, faces_host_pinned,
                           (size_t)faces_capacity);
    if (rc < 0) return rc;
    if (rc == 1) return SMB_OK;
    if (*nverts > 0 && *ntris > 0 && R - 1 >= 8 * ex->n_slabs && ex->n_slabs >= 2) return SMB_ERR_CAPACITY;  // sizes are set
  }
  // first call / tiny lattice / empty surface: single pass to learn the sizes (or classify the error)
  rc = smb_query_lattice_tc_signs(ex->planes_q, ex->blob_dev, &ex->layout, &ex->cfg, ex->axis_dev, R, 0, R, ex->density, nullptr, threshold,
                                  1.0f, ex->mc_ws, ex->mc_ws_bytes, st);
  if (rc != SMB_OK) return rc;
  rc = smb_mc_count_presigned(R, R, R, 1, ex->mc_ws, ex->mc_ws_bytes, ex->counts_dev, st);
  if (rc != SMB_OK) return rc;
  EX_CUDA(cudaMemcpyAsync(ex->counts_pin, ex->counts_dev, sizeof(smb_mc_counts), cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaStreamSynchronize(st));
  const int64_t V = ex->counts_pin->nverts, F = ex->counts_pin->ntris;
  *nverts = V;
  *ntris = F;
  if (V == 0 || F == 0) {
    rc = smb_grid_minmax(ex->density, (int64_t)R * R * R, threshold, 1.0f, ex->minmax_dev, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaMemcpyAsync(ex->minmax_pin, ex->minmax_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    EX_CUDA(cudaStreamSynchronize(st));
    if (ex->minmax_pin[0] > 0.0f || ex->minmax_pin[1] < 0.0f) return SMB_ERR_LEVEL_RANGE;
    return SMB_ERR_NO_SURFACE;
  }
  if (V > verts_capacity || F > faces_capacity) return SMB_ERR_CAPACITY;
  // fits, but the pipeline was not applicable (tiny lattice): emit + copy in one go
  const double r = (double)ex->cfg.radius;
  rc = smb_mc_emit_bounded(ex->density, R, R, R, threshold, 1.0f, 0, 1, SMB_MC_FLIP | SMB_MC_DIV | SMB_MC_AFFINE | (i32 ? SMB_MC_FACES_I32 : 0),
                           (float)(R - 1.0), (float)(r - (-r)), (float)(-r), 0, ex->mc_ws, ex->verts_dev, (int64_t)ex->verts_cap, ex->faces_dev,
                           (int64_t)ex->faces_cap, st);
  if (rc != SMB_OK) return rc;
  EX_CUDA(cudaMemcpyAsync(verts_host_pinned, ex->verts_dev, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaMemcpyAsync(faces_host_pinned, ex->faces_dev, (i32 ? sizeof(int32_t) : sizeof(int64_t)) * 3 * F, cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaStreamSynchronize(st));
  return SMB_OK;
}

extern "C" int smb_extractor_enable_timing(smb_extractor* ex, int enable) {
  if (!ex) return SMB_ERR_BAD_ARG;
  if (enable)
    for (int k = 0; k < 4; ++k)
      if (!ex->tev[k] && cudaEventCreate(&ex->tev[k]) != cudaSuccess) return SMB_ERR_CUDA;
  ex->timing = enable ? 1 : 0;
  return SMB_OK;
}
extern "C" int smb_extractor_last_timing(smb_extractor* ex, float* prepare_ms, float* lattice_ms, float* mc_ms) {
  if (!ex || !ex->timing || !ex->tev[3]) return SMB_ERR_BAD_ARG;
  float a = 0.f, b = 0.f, c = 0.f;
  if (cudaEventElapsedTime(&a, ex->tev[0], ex->tev[1]) != cudaSuccess || cudaEventElapsedTime(&b, ex->tev[1], ex->tev[2]) != cudaSuccess ||
      cudaEventElapsedTime(&c, ex->tev[2], ex->tev[3]) != cudaSuccess)
    return SMB_ERR_CUDA;
  if (prepare_ms) *prepare_ms = a;
  if (lattice_ms) *lattice_ms = b;
  if (mc_ms) *mc_ms = c;
  return SMB_OK;
}

extern "C" int smb_extractor_set_faces_i32(smb_extractor* ex, int enable) {
  if (!ex) return SMB_ERR_BAD_ARG;
  ex->faces_i32 = enable ? 1 : 0;
  return SMB_OK;
}

// The device-resident whole path (what TSR.extract_mesh_tensors runs): triplane already in HBM, mesh left in
// HBM in the CALLER's buffers.  One call queues prepare -> lattice (+ sign ballot) -> count -> totals -> emit on
// `stream` back to back (no host code between the launches, so the device never waits for the host), then reads
// the sizes.  emit_only != 0: density and word records of the previous call are still valid (same handle, same
// R / threshold) and only the emit is repeated -- what a caller does after SMB_ERR_CAPACITY with larger buffers.
extern "C" int smb_extract_mesh_device(smb_extractor* ex, const float* triplane_dev, int R, float threshold, int face_flags,
                                       float* verts_out, int64_t verts_capacity, void* faces_out, int64_t faces_capacity,
                                       float* density_out, int emit_only, void* stream, int64_t* nverts, int64_t* ntris) {
  if (!ex || !triplane_dev || R < 2 || !nverts || !ntris || verts_capacity < 0 || faces_capacity < 0) return SMB_ERR_BAD_ARG;
  if ((verts_capacity > 0 && !verts_out) || (faces_capacity > 0 && !faces_out)) return SMB_ERR_BAD_ARG;
  int rc = ensure_resolution(ex, R);
  if (rc != SMB_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* dens = density_out ? density_out : ex->density;
  const double r = (double)ex->cfg.radius;
  const int mc_flags = SMB_MC_FLIP | SMB_MC_DIV | SMB_MC_AFFINE | (face_flags & SMB_MC_FACES_I32);
  const float vdiv = (float)(R - 1.0), vmul = (float)(r - (-r)), vadd = (float)(-r);
  const bool timed = ex->timing && !emit_only;
  if (!emit_only) {
    if (timed) EX_CUDA(cudaEventRecord(ex->tev[0], st));
    rc = smb_scene_prepare(triplane_dev, ex->cfg.Hp, ex->cfg.Wp, ex->blob_dev, &ex->layout, nullptr, ex->planes_q, st);
    if (rc != SMB_OK) return rc;
    if (timed) EX_CUDA(cudaEventRecord(ex->tev[1], st));
    rc = smb_query_lattice_tc_signs(ex->planes_q, ex->blob_dev, &ex->layout, &ex->cfg, ex->axis_dev, R, 0, R, dens, nullptr, threshold, 1.0f,
                                    ex->mc_ws, ex->mc_ws_bytes, st);
    if (rc != SMB_OK) return rc;
    if (timed) EX_CUDA(cudaEventRecord(ex->tev[2], st));
    rc = smb_mc_count_presigned(R, R, R, 1, ex->mc_ws, ex->mc_ws_bytes, ex->counts_dev, st);
    if (rc != SMB_OK) return rc;
  }
  if (verts_capacity > 0 || faces_capacity > 0) {
    rc = smb_mc_emit_bounded(dens, R, R, R, threshold, 1.0f, 0, 1, mc_flags, vdiv, vmul, vadd, 0, ex->mc_ws, verts_out, verts_capacity,
                             static_cast<int64_t*>(faces_out), faces_capacity, st);
    if (rc != SMB_OK) return rc;
  }
  if (timed) EX_CUDA(cudaEventRecord(ex->tev[3], st));
  EX_CUDA(cudaMemcpyAsync(ex->counts_pin, ex->counts_dev, sizeof(smb_mc_counts), cudaMemcpyDeviceToHost, st));
  EX_CUDA(cudaStreamSynchronize(st));
  const int64_t V = ex->counts_pin->nverts, F = ex->counts_pin->ntris;
  *nverts = V;
  *ntris = F;
  if (V == 0 || F == 0) {
    rc = smb_grid_minmax(dens, (int64_t)R * R * R, threshold, 1.0f, ex->minmax_dev, st);
    if (rc != SMB_OK) return rc;
    EX_CUDA(cudaMemcpyAsync(ex->minmax_pin, ex->minmax_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    EX_CUDA(cudaStreamSynchronize(st));
    if (ex->minmax_pin[0] > 0.0f || ex->minmax_pin[1] < 0.0f) return SMB_ERR_LEVEL_RANGE;
    return SMB_ERR_NO_SURFACE;
  }
  if ((face_flags & SMB_MC_FACES_I32) && V >= (int64_t)1 << 31) return SMB_ERR_BAD_ARG;
  if (V > verts_capacity || F > faces_capacity) return SMB_ERR_CAPACITY;
  return SMB_OK;
}

// ------------------------------------------------- peer memory (multi-GPU gather)
// The sharded path lets every rank's emit kernel store its slab straight into the destination
// rank's mesh buffers over NVLink.  The destination allocates them here (cudaMalloc: the pointer
// IS the allocation base, which cudaIpcGetMemHandle requires), exports a handle, and the other
// processes map it; cudaIpcOpenMemHandle enables peer access between the two devices.
extern "C" int smb_dev_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return SMB_ERR_BAD_ARG;
  return smb_check(cudaMalloc(out, bytes));
}
extern "C" int smb_dev_free(void* ptr) { return smb_check(cudaFree(ptr)); }
extern "C" int smb_ipc_export(const void* dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64) return SMB_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  return smb_check(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(dev_ptr)));
}
extern "C" int smb_ipc_open(const void* handle64, void** out) {
  if (!handle64 || !out) return SMB_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return smb_check(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
}
extern "C" int smb_ipc_close(void* mapped_ptr) { return smb_check(cudaIpcCloseMemHandle(mapped_ptr)); }

// Lattice query entry points (smb_query_lattice_tc / _signs): parameter checks, the tile geometry the kernels share, and the
// dispatch to the tensor-core lattice kernel (field_tc_ta.cu).  The density half of TSR.extract_mesh
// (/root/reference/TripoSR/tsr/system.py:171-184): query_triplane (tsr/models/nerf_renderer.py:41-91) + NeRFMLP.forward
// (tsr/models/network_utils.py:116-124) on the lattice of MarchingCubeHelper.grid_vertices (tsr/models/isosurface.py:25-39).
//
// Environment switches are read ONCE, at the first launch (tc_env): SMB_TC_WAITNS (suspend-time hint of the consumers'
// mbarrier waits, ns).  A developer build (-DSMB_DEV_VARIANTS, `python -m sculptmate_b200.build --dev`) additionally
// compiles the superseded / experimental kernels and their switches (SMB_TC_VARIANT, SMB_TC_TRACE, SMB_TC_TA_*,
// SMB_TC_POLY); the product library carries exactly one lattice kernel plus its four-warpgroup fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "field_tc_common.cuh"

namespace smb {
struct TcEnv {
  int wait_ns = 2000;
  int trace = 0, stagger = 0, tokens = 0, poly = 0;
  char variant = 0;
  bool axis_tables = true;  // SMB_NO_AXIS_TABLES: the producers interpolate per tile (the round-1 path, kept as the no-scratch fallback)
};
static const TcEnv& tc_env() {
  static const TcEnv env = [] {
    TcEnv e;
    if (const char* v = getenv("SMB_TC_WAITNS")) e.wait_ns = atoi(v);
    if (getenv("SMB_NO_AXIS_TABLES")) e.axis_tables = false;
#ifdef SMB_DEV_VARIANTS
    if (const char* v = getenv("SMB_TC_TRACE")) e.trace = atoi(v);
    if (const char* v = getenv("SMB_TC_TA_STAGGER")) e.stagger = atoi(v);
    if (const char* v = getenv("SMB_TC_TA_TOKENS")) e.tokens = atoi(v);
    if (const char* v = getenv("SMB_TC_POLY")) e.poly = atoi(v);
    if (const char* v = getenv("SMB_TC_VARIANT")) e.variant = v[0];
#endif
    return e;
  }();
  return env;
}
// Axis tables of one launch: for every lattice x of the slab the x-interpolated rows of the projected (x,z) plane, for every
// lattice y those of the (y,z) plane -- the two in-plane bilinear interpolations of layer 0 depend on ONE lattice coordinate
// each, so they are done (nx + R) * H times per launch instead of once per tile (R^2 * nx * tiles_per_line times) by the
// producer warps, whose table build then is two loads and an add per element.  (Zero padding is folded into the tap weights.)
__global__ void __launch_bounds__(256) lattice_axis_tables(TcParams p, float* __restrict__ t1, float* __restrict__ t2) {
  const long long HW = (long long)p.H * p.W;
  const float* Q1 = p.planes_q + HW * kHid;
  const float* Q2 = Q1 + HW * kHid;
  const long long n1 = (long long)p.nx * p.H * (kHid / 4), n2 = (long long)p.R * p.H * (kHid / 4);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n1 + n2; t += (long long)gridDim.x * blockDim.x) {
    const bool second = t >= n1;
    const long long u = second ? t - n1 : t;
    const int n4 = (int)(u % (kHid / 4));
    const long long rh = u / (kHid / 4);
    const int h = (int)(rh % p.H), row = (int)(rh / p.H);
    const Tap2 tap = make_tap(p.axis_u[second ? row : p.x_begin + row], p.W, p.align_corners);
    const float* Q = second ? Q2 : Q1;
    const float4 a = __ldg(reinterpret_cast<const float4*>(Q + ((long long)h * p.W + tap.i0) * kHid) + n4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(Q + ((long long)h * p.W + tap.i1) * kHid) + n4);
    const float4 v = make_float4(tap.w0 * a.x + tap.w1 * b.x, tap.w0 * a.y + tap.w1 * b.y, tap.w0 * a.z + tap.w1 * b.z, tap.w0 * a.w + tap.w1 * b.w);
    reinterpret_cast<float4*>(second ? t2 : t1)[u] = v;
  }
}

// Keep freed stream-ordered scratch in the device's default pool instead of returning it to the driver at every
// synchronisation (release threshold 0 is the default: each call would then pay a fresh allocation).
void keep_async_scratch(int dev) {
  static int pool_dev = -1;
  if (pool_dev != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_dev = dev;
  }
}

#ifdef SMB_DEV_VARIANTS
int launch_tc_smem(const TcParams& p, int sms, bool trace, cudaStream_t st);
int read_trace_smem(long long* host, int n);
int pair_debug_dump();
int pair_prof_read(unsigned int* host, int n);
int pair_evt_read(unsigned int* host, int n);
#endif
}  // namespace smb

using namespace smb;

static int query_lattice_tc_impl(const float* planes_q, const void* decoder_blob,
                                 const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                 const float* axis_u, int R, int x_begin, int nx, float* out_density_act,
                                 float* out_density, bool want_signs, float sub, float sign, void* mc_workspace,
                                 size_t mc_workspace_bytes, void* stream) {
  if (!planes_q || !decoder_blob || !layout || !cfg || !axis_u || !out_density_act) return SMB_ERR_BAD_ARG;
  if (R < 2 || nx < 0 || x_begin < 0 || x_begin + nx > R) return SMB_ERR_BAD_ARG;
  if (nx == 0) return SMB_OK;
  const int nh = (int)layout->n_hidden;
  if (nh < 2 || nh > kMaxHidden) return SMB_ERR_BAD_ARG;
  // rows of the (.,z) planes one 128-sample segment can touch (+2 for the taps, +1 slack)
  int rows;
  {
    double span = 127.0 * cfg->Hp / (double)(R - 1);
    rows = (int)span + 3;
    if (rows > cfg->Hp + 2) rows = cfg->Hp + 2;  // rows -1 .. Hp (the two zero borders)
    if (rows < 2) rows = 2;
    if (rows > kTRowsMax) return SMB_ERR_BAD_ARG;
  }
  // layout contract: [hidden | head | bias_half | bias_final] contiguous in the blob
  if (layout->off_tc_final != layout->off_tc_hidden + (uint32_t)(nh - 1) * kWBytes ||
      layout->off_bias_half != layout->off_tc_final + kWFinalBytes ||
      layout->off_bias_final != layout->off_bias_half + (uint32_t)nh * kHid * 4)
    return SMB_ERR_BAD_ARG;

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  TcParams p{};
  p.planes_q = planes_q;
  p.tc_weights = static_cast<const unsigned char*>(decoder_blob) + layout->off_tc_hidden;
  p.tc_biasblk = static_cast<const unsigned char*>(decoder_blob) + layout->off_tc_biasblk;
  p.bias0_half = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_bias_half);
  // plain fp32 copy of the parameters: [W0 (64x120) b0 (64)] [(W_l (64x64) b_l (64)) x (nh-1)] [W_L (4x64) b_L (4)]
  p.head_w_f32 = reinterpret_cast<const float*>(static_cast<const char*>(decoder_blob) + layout->off_f32) + (kHid * kFeat + kHid) +
                 (size_t)(nh - 1) * (kHid * kHid + kHid);
  p.axis_u = axis_u;
  p.R = R;
  p.x_begin = x_begin;
  p.nx = nx;
  p.H = cfg->Hp;
  p.W = cfg->Wp;
  p.align_corners = cfg->align_corners;
  p.n_hidden = nh;
  p.trows = rows;
  p.slot_bytes = tc_slot_bytes(rows);
  p.density_bias = cfg->density_bias;
  p.out_act = out_density_act;
  p.out_raw = out_density;
  const TcEnv& env = tc_env();
  p.wait_ns = env.wait_ns;
  cudaStream_t st = (cudaStream_t)stream;
  if (want_signs) {
    if (!mc_workspace || smb_mc_workspace_bytes(nx, R, R) > mc_workspace_bytes) return SMB_ERR_WORKSPACE;
    p.sign_out = static_cast<uint32_t*>(mc_workspace);  // the sign masks are the first region of an MC workspace
    p.sign_sub = sub;
    p.sign_mul = sign;
    p.sign_wz = (R + 31) / 32;
  }
#ifdef SMB_DEV_VARIANTS
  // developer build: the experiments DESIGN.md argues from (SMB_TC_VARIANT=smem | pair, SMB_TC_TRACE, ...)
  p.dbg = env.trace;
  p.stagger_clk = env.stagger;
  p.xu_tokens = env.tokens;
  if (env.variant == 'p' && (!p.dbg || p.dbg == 3)) {
    const int rc2 = launch_tc_pair(p, sms, env.poly, st);
    if (rc2 != SMB_ERR_BAD_ARG) return rc2;
  }
  if (env.variant == 's' || p.dbg == 1) {
    p.sign_out = nullptr;  // the shared-memory-A kernel does not ballot: stand-alone sign pass below
    int rc = launch_tc_smem(p, sms, p.dbg == 1, st);
    if (rc == SMB_OK && want_signs) rc = launch_mc_signs(out_density_act, nx, R, R, sub, sign, mc_workspace, mc_workspace_bytes, st);
    return rc;
  }
#endif
  // axis tables in stream-ordered scratch memory (the pool keeps the block between calls: no allocation cost after the first)
  const size_t tab_floats = ((size_t)nx + (size_t)R) * cfg->Hp * kHid;
  float* tabs = nullptr;
  keep_async_scratch(dev);
  if (tc_env().axis_tables && cudaMallocAsync(reinterpret_cast<void**>(&tabs), tab_floats * sizeof(float), st) == cudaSuccess) {
    p.t1 = tabs;
    p.t2 = tabs + (size_t)nx * cfg->Hp * kHid;
    const long long items = (long long)tab_floats / 4;
    const int blocks = (int)((items + 255) / 256 < (long long)sms * 8 ? (items + 255) / 256 : (long long)sms * 8);
    lattice_axis_tables<<<blocks, 256, 0, st>>>(p, tabs, tabs + (size_t)nx * cfg->Hp * kHid);
  } else {
    (void)cudaGetLastError();  // no scratch: the producers interpolate per tile (the round-1 path)
    p.t1 = p.t2 = nullptr;
  }
  const int rc = launch_tc_ta(p, sms, st);
  if (tabs) cudaFreeAsync(tabs, st);
  return rc;
}

extern "C" int smb_query_lattice_tc(const float* planes_q, const void* decoder_blob,
                                    const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                    const float* axis_u, int R, int x_begin, int nx, float* out_density_act,
                                    float* out_density, void* stream) {
  return query_lattice_tc_impl(planes_q, decoder_blob, layout, cfg, axis_u, R, x_begin, nx, out_density_act, out_density, false, 0.f,
                               1.f, nullptr, 0, stream);
}

extern "C" int smb_query_lattice_tc_signs(const float* planes_q, const void* decoder_blob,
                                          const smb_decoder_layout* layout, const smb_query_cfg* cfg,
                                          const float* axis_u, int R, int x_begin, int nx, float* out_density_act,
                                          float* out_density, float sub, float sign, void* mc_workspace,
                                          size_t mc_workspace_bytes, void* stream) {
  if (!mc_workspace) return SMB_ERR_BAD_ARG;
  return query_lattice_tc_impl(planes_q, decoder_blob, layout, cfg, axis_u, R, x_begin, nx, out_density_act, out_density, true, sub,
                               sign, mc_workspace, mc_workspace_bytes, stream);
}

#ifdef SMB_DEV_VARIANTS
// developer instrumentation (not declared in include/sculptmate_b200.h; tools/trace_lattice*.py, tools/k1_*.py)
extern "C" int smb_debug_pair_evt(unsigned int* host, int n) { return smb::pair_evt_read(host, n); }
extern "C" int smb_debug_pair_prof(unsigned int* host, int n) { return smb::pair_prof_read(host, n); }
extern "C" int smb_debug_pair_dump(void) { return smb::pair_debug_dump(); }
extern "C" int smb_debug_read_trace_ta(long long* host, int n) { return smb::read_trace_ta(host, n); }
extern "C" int smb_debug_read_trace(long long* host, int n) { return smb::read_trace_smem(host, n); }
#endif

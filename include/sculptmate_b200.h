/* sculptmate_b200 -- C ABI of the B200-native extract_mesh hot path.
 *
 * Drop-in boundary for ONE path of shravan-d/SculptMate: TripoSR's
 * TSR.extract_mesh = dense triplane query (grid_sample x3 -> NeRFMLP) over an
 * R^3 lattice followed by marching cubes.  The reference is pure Python with no
 * FFI of its own, so each entry point cites the Python interface it replaces
 * (paths relative to the reference root); the binding a maintainer would add is
 * a ctypes stub, shown in INTEGRATION.md and shipped as sculptmate_b200/_capi.py.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*;
 *   - every `*_dev` / unqualified pointer is DEVICE memory on the current device,
 *     every `*_host` pointer is host memory;
 *   - functions return SMB_OK (0) or a negative smb_status; they never throw and
 *     never synchronise the device unless the name ends in `_host`;
 *   - all kernels are sm_100a only; there is no CPU fallback.
 */
#ifndef SCULPTMATE_B200_H
#define SCULPTMATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB_PLANE_CHANNELS 40     /* TripoSR/checkpoints/config.yaml:22-23 (out_channels: 40) */
#define SMB_HIDDEN 64             /* config.yaml:27 n_neurons */
#define SMB_MAX_HIDDEN_LAYERS 16  /* config.yaml:28 uses 9 */

typedef enum smb_status {
  SMB_OK = 0,
  SMB_ERR_CUDA = -1,        /* a CUDA runtime call or launch failed */
  SMB_ERR_BAD_ARG = -2,     /* null pointer / non-positive size / unsupported shape */
  SMB_ERR_WORKSPACE = -3,   /* workspace too small */
  SMB_ERR_ARCH = -4,        /* device is not sm_100 */
  SMB_ERR_LEVEL_RANGE = -5, /* iso level outside the data range (skimage: ValueError) */
  SMB_ERR_NO_SURFACE = -6,  /* no triangle produced (skimage: RuntimeError) */
  SMB_ERR_CAPACITY = -7     /* caller-provided output buffers too small; the needed sizes were returned */
} smb_status;

const char* smb_status_string(int status);
/* 0 when the current device is compute capability 10.x, else SMB_ERR_ARCH / SMB_ERR_CUDA. */
int smb_device_check(void);
int smb_version(void);

/* ------------------------------------------------------------------ decoder
 * NeRFMLP weights (tsr/models/network_utils.py:48-79; state-dict keys
 * layers.{0,2,..,2L}.{weight,bias}) packed once into a device blob that holds
 * every form the kernels need.  n_hidden = config n_hidden_layers (9):
 * Linear(120,64)+SiLU, (n_hidden-1) x [Linear(64,64)+SiLU], Linear(64,4).
 *
 * weights_host[l] is (out,in) row-major fp32 exactly as nn.Linear stores it,
 * biases_host[l] is (out) fp32, l = 0..n_hidden (n_hidden+1 layers).
 */
typedef struct smb_decoder_layout {
  uint32_t total_bytes;
  uint32_t n_hidden;
  uint32_t off_tc_hidden;  /* (n_hidden-1) x 8192 B: W_l/2, l=1.., fp16, K-major 128B-swizzle UMMA image */
  uint32_t off_tc_final;   /* 2048 B: last layer padded to 16 rows, fp16, same image format */
  uint32_t off_tc_l0;      /* 16384 B: W_0/2 padded to K=128 as two [64 x 64] K-blocks (points kernel) */
  uint32_t off_bias_half;  /* n_hidden x 64 fp32: b_l/2, l = 0..n_hidden-1 */
  uint32_t off_bias_final; /* 4 fp32 (padded to 16 B) */
  uint32_t off_w0_half;    /* 64 x 120 fp32: W_0/2 (plane projection for the lattice kernel) */
  uint32_t off_f32;        /* plain fp32 copy: for each layer W (out,in) then b (out) */
  uint32_t off_tc_biasblk; /* (n_hidden-1) x 8192 B: per hidden layer l=1.. a [64 x 64] fp16 K-block in the same image format whose
                              K rows 0/1 hold b_l/2 split as fp16 hi + lo (rest zero): the bias enters the accumulator through one
                              extra K=16 MMA against a constant (1,1,0,..) activation block instead of 64 FADDs per sample */
  uint32_t reserved[6];
} smb_decoder_layout;

int smb_decoder_layout_for(int n_hidden, smb_decoder_layout* out);
/* Fills blob_host (layout->total_bytes bytes, host memory); the caller uploads it. */
int smb_decoder_pack_host(const float* const* weights_host, const float* const* biases_host, int n_hidden,
                          const smb_decoder_layout* layout, void* blob_host);

/* ------------------------------------------------------------- scene planes
 * Per scene code (triplane (3,Cp,Hp,Wp) fp32 NCHW, nerf_renderer.py:61-66):
 *   planes_cl : channels-last copy (3,Hp,Wp,Cp) fp32   -- arbitrary-position query
 *   planes_q  : layer-0 projection (3,Hp,Wp,64) fp32 = (W_0/2) . plane -- lattice query
 * (bilinear interpolation is linear, so projecting before interpolating equals
 * the reference's interpolate-then-Linear in exact arithmetic; it is done in fp32.)
 */
int smb_scene_prepare(const float* triplane, int Hp, int Wp, const void* decoder_blob,
                      const smb_decoder_layout* layout, float* planes_cl, float* planes_q, void* stream);

/* ------------------------------------------------------------- field query
 * Replaces TriplaneNeRFRenderer.query_triplane (nerf_renderer.py:41-91) for
 * arbitrary positions.  positions (n,3) fp32 in (-radius, radius); the
 * (-radius,radius)->(-1,1) rescale (nerf_renderer.py:52-54) happens in-kernel
 * with the reference's fp32 operation order.  Any output pointer may be NULL.
 *   density (n), features (n,3), density_act (n) = exp(density+density_bias),
 *   color (n,3) = sigmoid(features).
 * fp32 CUDA-core kernel: bit-comparable to the reference within fp32 noise.
 */
typedef struct smb_query_cfg {
  float radius;        /* renderer.cfg.radius (0.87) */
  float density_bias;  /* renderer.cfg.density_bias (-1.0) */
  int align_corners;   /* 0 TripoSR (nerf_renderer.py:64), 1 SF3D (sf3d/system.py:193) */
  int Hp, Wp;          /* plane size */
} smb_query_cfg;

int smb_query_points_f32(const float* planes_cl, const void* decoder_blob, const smb_decoder_layout* layout,
                         const smb_query_cfg* cfg, const float* positions, int64_t n, float* density,
                         float* features, float* density_act, float* color, void* stream);

/* NeRFMLP.forward (network_utils.py:116-124) on pre-computed features (n,120): density (n) = out[...,0],
 * features (n,3) = out[...,1:4]; fp32, either output may be NULL. */
int smb_decoder_forward_f32(const void* decoder_blob, const smb_decoder_layout* layout, const float* features_in,
                            int64_t n, float* density, float* features, void* stream);

/* Lattice query = the density half of TSR.extract_mesh (tsr/system.py:171-184):
 * evaluates density_act on x-planes [x_begin, x_begin+nx) of the R^3 lattice
 * (row (i*R+j)*R+k = (x_i,y_j,z_k), isosurface.py:25-39) without materialising
 * positions.  axis_u (R) holds the per-axis sample coordinate already mapped to
 * (-1,1) by the caller with the reference's own torch ops (linspace ->
 * scale_tensor -> scale_tensor), so coordinates are bit-identical.
 * out_density_act: (nx,R,R) fp32.  out_density (optional, may be NULL): raw logit.
 *   smb_query_lattice_tc  : fused tcgen05 kernel (fp16 operands, fp32 accumulate)
 *   smb_query_lattice_f32 : fp32 CUDA-core kernel (same result as query_points_f32)
 */
/* Scalar fp32 restatement of linspace(0,1,R) -> (-radius,radius) -> (-1,1) for hosts
 * without torch (within 2 ulp of aten's vectorised linspace; see capi.cu). */
int smb_lattice_axis_host(int R, float radius, float* axis_u_host);

int smb_query_lattice_tc(const float* planes_q, const void* decoder_blob, const smb_decoder_layout* layout,
                         const smb_query_cfg* cfg, const float* axis_u, int R, int x_begin, int nx,
                         float* out_density_act, float* out_density, void* stream);
int smb_query_lattice_f32(const float* planes_cl, const void* decoder_blob, const smb_decoder_layout* layout,
                          const smb_query_cfg* cfg, const float* axis_u, int R, int x_begin, int nx,
                          float* out_density_act, float* out_density, void* stream);
/* smb_query_lattice_tc that also leaves the marching-cubes sign masks of the slab it evaluates in
 * mc_workspace (a workspace of smb_mc_workspace_bytes(nx, R, R)): bit k%32 of word (i*R + j)*ceil(R/32) + k/32
 * = ((density_act - sub) * sign > 0), the case bit of tsr/system.py:184 + isosurface.py:45-46 with
 * sub = threshold, sign = 1.  The warp that holds 32 consecutive z-samples ballots them while they are still in
 * registers, so the marching-cubes pass (smb_mc_count_presigned) never re-reads the 4 R^3-byte density grid to
 * classify it. */
int smb_query_lattice_tc_signs(const float* planes_q, const void* decoder_blob, const smb_decoder_layout* layout,
                               const smb_query_cfg* cfg, const float* axis_u, int R, int x_begin, int nx,
                               float* out_density_act, float* out_density, float sub, float sign,
                               void* mc_workspace, size_t mc_workspace_bytes, void* stream);

/* ---------------------------------------------------------- marching cubes
 * Replaces MarchingCubeHelper.forward (tsr/models/isosurface.py:41-54), i.e.
 * skimage.measure.marching_cubes(level, 0.0) + the wrapper's post-processing,
 * on a slab of x-planes.  val(p) = (grid[p] - sub) * sign, surface at val = 0,
 * case bit = val > 0.  Output order is canonical and deterministic:
 * vertices by (x-plane; in-plane edges by (j,k,axis); then x-edges by (j,k)),
 * triangles by (cell, table order).  Two calls: count, then emit.
 */
#define SMB_MC_FLIP 1    /* faces[:, [1,0,2]]            isosurface.py:52 */
#define SMB_MC_DIV 2     /* verts / vdiv (IEEE fp32)     isosurface.py:53 */
#define SMB_MC_AFFINE 4  /* verts * vmul + vadd          system.py:185-189 */
#define SMB_MC_COALESCE 16 /* emit assembles each warp's output run in shared memory and writes it with coalesced 128-byte
                           * stores: for destinations in PEER memory (NVLink), where scattered 4-byte stores are slow */
#define SMB_MC_FACES_I32 8 /* `faces` is (F,3) int32 instead of int64: the index width Blender's loop arrays use
                           * (system.py:127-131 hands the array to bpy); same values, half the bytes on PCIe / NVLink */

typedef struct smb_mc_counts {
  int64_t nverts;          /* vertices this slab stores */
  int64_t ntris;           /* triangles this slab stores */
  int64_t nverts_numbered; /* nverts + in-plane crossings of the last plane when not emitted */
  int64_t reserved;
} smb_mc_counts;

size_t smb_mc_workspace_bytes(int nx, int ny, int nz);
/* counts_dev: device smb_mc_counts written by the stream; copy it back to size the outputs. */
int smb_mc_count(const float* grid, int nx, int ny, int nz, float sub, float sign, int emit_last_plane,
                 void* workspace, size_t workspace_bytes, smb_mc_counts* counts_dev, void* stream);
/* smb_mc_count for a workspace whose sign masks were already written by smb_query_lattice_tc_signs with the
 * same (nx, ny, nz, sub, sign): count + scan only. */
int smb_mc_count_presigned(int nx, int ny, int nz, int emit_last_plane, void* workspace, size_t workspace_bytes,
                           smb_mc_counts* counts_dev, void* stream);
/* Must follow smb_mc_count on the same grid/workspace.  vertex_id_offset is added to
 * every face index (running vertex count of the lower slabs); x_origin is the global
 * index of the slab's first plane.  verts (nverts,3) fp32, faces (ntris,3) int64. */
int smb_mc_emit(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                int emit_last_plane, int flags, float vdiv, float vmul, float vadd, int64_t vertex_id_offset,
                const void* workspace, float* verts, int64_t* faces, void* stream);
/* Same, into buffers of a given capacity (in vertices / triangles): elements beyond the capacity are
 * dropped.  Lets a caller that remembers the size of its last mesh launch emit right behind count --
 * no host round trip between them -- and read the counts afterwards; it re-emits only on overflow. */
int smb_mc_emit_bounded(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                        int emit_last_plane, int flags, float vdiv, float vmul, float vadd, int64_t vertex_id_offset,
                        const void* workspace, float* verts, int64_t verts_capacity, int64_t* faces,
                        int64_t faces_capacity, void* stream);
/* Gather mode (multi-GPU, one process per GPU): all_counts_dev is (world,4) int64 = every rank's
 * smb_mc_counts (an all-gather of the counts_dev of smb_mc_count, device to device); the kernel derives this
 * slab's vertex / triangle offsets from the lower ranks' counts and stores its part of the mesh, with global
 * vertex ids, straight into verts_dst / faces_dst -- the DESTINATION rank's buffers, mapped into this process
 * with smb_ipc_open (NVLink peer stores; no staging buffer, no send/recv).  Elements beyond the capacities are
 * dropped (the caller re-runs with larger buffers). */
int smb_mc_emit_gather(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                       int emit_last_plane, int flags, float vdiv, float vmul, float vadd, const void* workspace,
                       const int64_t* all_counts_dev, int rank, float* verts_dst, int64_t verts_capacity,
                       int64_t* faces_dst, int64_t faces_capacity, void* stream);
/* Gather mode without collectives: every rank owns a control block of SMB_PEER_CTRL_WORDS int64 (smb_dev_alloc +
 * IPC, mapped by every process; zero it once).  Per call, with a sequence number seq = 1, 2, ... shared by the ranks:
 *   smb_peer_wait_release(ctrl, seq-1)         first: the destination has consumed the previous call
 *   ... lattice, smb_mc_count* into counts_dev ...
 *   smb_peer_publish_counts(counts_dev, peers, rank, world, seq)    NVLink stores of the counts + flag into EVERY block
 *   smb_mc_emit_gather_flags(..., ctrl, seq, rank, dst buffers)     waits (on the device) for the lower ranks' flags,
 *                                                                   then stores its slab into the destination's buffers
 *   smb_peer_signal_done(dst_ctrl, rank, seq)                        flag into the destination's block
 *   smb_peer_wait_all(ctrl, peers, world, seq, wait_done, totals_dev)  -> totals_dev = {sum V, sum F, error, seq}:
 *       destination (wait_done = 1): every slab stored; also releases the peers for call seq+1;
 *       other ranks (wait_done = 0): every rank's counts have arrived (so that all ranks see the same totals)
 * peers = DEVICE array of `world` pointers to the control blocks as mapped in this process.  No NCCL call, no host
 * round trip; waits are bounded (about 3 s) and set the error word instead of hanging the GPU. */
#define SMB_PEER_CTRL_WORDS 128
int smb_peer_wait_release(void* ctrl_local, int64_t need, void* stream);
int smb_peer_publish_counts(const smb_mc_counts* counts_dev, void* const* peer_ctrl_dev, int rank, int world, int64_t seq, void* stream);
int smb_mc_emit_gather_flags(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                             int emit_last_plane, int flags, float vdiv, float vmul, float vadd, const void* workspace,
                             void* ctrl_local, int64_t seq, int rank, float* verts_dst, int64_t verts_capacity,
                             void* faces_dst, int64_t faces_capacity, void* stream);
int smb_peer_signal_done(void* dst_ctrl, int rank, int64_t seq, void* stream);
int smb_peer_wait_all(void* ctrl_local, void* const* peer_ctrl_dev, int world, int64_t seq, int wait_done, int64_t* totals_dev,
                      void* stream);
/* Peer-memory plumbing for the above: the destination rank allocates its mesh buffers with smb_dev_alloc,
 * exports a 64-byte handle, the other ranks map it (this also enables peer access). */
int smb_dev_alloc(size_t bytes, void** out);
int smb_dev_free(void* ptr);
int smb_ipc_export(const void* dev_ptr, void* handle64);
int smb_ipc_open(const void* handle64, void** out);
int smb_ipc_close(void* mapped_ptr);
/* cube-case index of every cell, (nx-1,ny-1,nz-1) uint8 (parity/debug). */
int smb_mc_cases(const float* grid, int nx, int ny, int nz, float sub, float sign, unsigned char* cases,
                 void* stream);
/* min/max of val over the grid -> minmax_dev[2]; used to tell SMB_ERR_LEVEL_RANGE from SMB_ERR_NO_SURFACE. */
int smb_grid_minmax(const float* grid, int64_t n, float sub, float sign, float* minmax_dev, void* stream);

/* ------------------------------------------------- whole path, host buffers
 * TSR.extract_mesh for one scene code with HOST inputs and outputs
 * (tsr/system.py:171-200 minus the Blender import): uploads the triplane,
 * runs prepare -> lattice query (tensor cores) -> marching cubes, downloads
 * the mesh.  Synchronises the stream.  The mesh is returned in buffers owned by
 * the handle and valid until the next call / smb_extractor_destroy.
 */
typedef struct smb_extractor smb_extractor;
int smb_extractor_create(const float* const* weights_host, const float* const* biases_host, int n_hidden,
                         float radius, float density_bias, int Hp, int Wp, smb_extractor** out);
void smb_extractor_destroy(smb_extractor* ex);
/* Optional: supply the per-axis lattice coordinates (R values in (-1,1)) for resolution R instead of
 * the scalar smb_lattice_axis_host restatement -- a host that has torch passes the values the
 * reference's own ops produce (isosurface.py:30-32 -> system.py:177-181 -> nerf_renderer.py:52-54),
 * making the C path bit-identical to the Python drop-in.  Copied; valid until R changes. */
int smb_extractor_set_axis(smb_extractor* ex, int resolution, const float* axis_u_host);
/* The handle's pinned staging buffer for the scene code (3*40*Hp*Wp floats).  A host that writes the triplane
 * there and passes the same pointer as triplane_host skips the pageable->pinned copy (about 0.1 ms for 2 MB). */
int smb_extractor_pinned_input(smb_extractor* ex, float** triplane_pinned);
int smb_extract_mesh_host(smb_extractor* ex, const float* triplane_host, int resolution, float threshold,
                          const float** verts_host, const int64_t** faces_host, int64_t* nverts,
                          int64_t* ntris);

/* Deliver the faces of smb_extract_mesh_host[_textured] as (F,3) int32 instead of int64 (the returned pointer then
 * addresses int32 data): the index width Blender's loop arrays use, and a third less PCIe traffic per mesh. */
int smb_extractor_set_faces_i32(smb_extractor* ex, int enable);

/* Device-resident variant = what TSR.extract_mesh does between "scene code on the GPU" and "v_pos / t_pos_idx on the
 * GPU" (tsr/system.py:173-189): triplane_dev (3,40,Hp,Wp) fp32 in HBM, mesh written to the CALLER's buffers
 * verts_out (verts_capacity,3) fp32 / faces_out (faces_capacity,3) int64 (int32 with face_flags = SMB_MC_FACES_I32).
 * All kernels are queued on `stream` by this one call, then the sizes are read (one stream synchronisation).
 * Returns SMB_ERR_CAPACITY with *nverts / *ntris set when a buffer is too small: allocate and call again with
 * emit_only = 1 (density and records are kept).  density_out (optional, (R,R,R) fp32) receives density_act. */
int smb_extract_mesh_device(smb_extractor* ex, const float* triplane_dev, int resolution, float threshold, int face_flags,
                            float* verts_out, int64_t verts_capacity, void* faces_out, int64_t faces_capacity,
                            float* density_out, int emit_only, void* stream, int64_t* nverts, int64_t* ntris);

/* Device triplane in, mesh out in the CALLER's pinned host buffers (capacities in vertices / triangles): the plugin call
 * TSR.extract_mesh ends in `.cpu().numpy()` of both arrays (system.py:200).  Slab pipeline as in smb_extract_mesh_host (the mesh
 * of slab k crosses PCIe while slab k+1 is computed) but into buffers the caller owns, so results never alias across calls.
 * SMB_ERR_CAPACITY with *nverts / *ntris set when the mesh does not fit (first call: pass capacities 0 to learn the sizes).
 * Runs on the handle's own streams, ordered after `stream`; returns after the copies have landed. */
int smb_extract_mesh_device_to_host(smb_extractor* ex, const float* triplane_dev, int resolution, float threshold, int face_flags,
                                    float* verts_host_pinned, int64_t verts_capacity, void* faces_host_pinned,
                                    int64_t faces_capacity, void* stream, int64_t* nverts, int64_t* ntris);

/* Optional phase timing of smb_extract_mesh_device: CUDA events on the caller's stream around prepare / the lattice
 * kernel / marching cubes (count + totals + emit); smb_extractor_last_timing returns the last call's durations in ms
 * (valid after that call returned: it synchronises the stream).  This is how bench.py measures the dominant kernel
 * live inside its timed region. */
int smb_extractor_enable_timing(smb_extractor* ex, int enable);
int smb_extractor_last_timing(smb_extractor* ex, float* prepare_ms, float* lattice_ms, float* mc_ms);

/* The same with enable_texture=True (system.py:190-200): also returns the vertex colours
 * colors_host (nverts,3) fp32 = query_triplane(decoder, v_pos, scene_code)["color"] (sigmoid of the three
 * feature outputs; tensor-core points kernel) and, when loop_colors_host != NULL, the (3*ntris,4) RGBA-per-loop
 * array the Blender sink assigns (smb_mesh_loop_colors, alpha = 1).  Buffers owned by the handle. */
int smb_extract_mesh_host_textured(smb_extractor* ex, const float* triplane_host, int resolution, float threshold,
                                   const float** verts_host, const int64_t** faces_host, const float** colors_host,
                                   const float** loop_colors_host, int64_t* nverts, int64_t* ntris);

/* ------------------------------------------- field query on tensor cores, any positions
 * One kernel for both decoders of the path: bilinear gathers from the channels-last planes feed an MLP
 * whose every layer is a tcgen05 UMMA (fp16 operands, fp32 accumulate), activations kept on-chip.
 *   TripoSR: query_triplane + NeRFMLP (nerf_renderer.py:41-91) -- the colour query at mesh vertices
 *            (system.py:191-198): layers 120->64, 8 x 64->64, 64->4.
 *   SF3D:    query_triplane + MaterialMLP heads density / vertex_offset (sf3d/system.py:153-154) fused as
 *            120->128 (both first layers), 128->128 (block-diagonal), 128->4.
 * The MLP is packed on the host from dense row-major (out,in) matrices: layer l has k_in[l] inputs
 * (<= 128, padded to 64 or 128) and n_out[l] outputs (hidden: padded to 64 or 128 and equal to the next
 * layer's padded K; last: <= 4).  Every layer but the last is followed by SiLU.
 * Outputs (each optional): out0_raw (n) = y0 + out0_bias, out0_act (n) = exp(out0_raw),
 * out_vec (n,3) = y1..y3, out_vec_act (n,3) = sigmoid(y1..y3) when sigmoid_vec != 0.
 */
#define SMB_MLP_TC_MAX_LAYERS 12
typedef struct smb_mlp_tc_layout {
  uint32_t total_bytes;
  uint32_t n_layers;
  uint32_t kblocks[SMB_MLP_TC_MAX_LAYERS]; /* 64-wide K blocks of the layer's input (1 or 2) */
  uint32_t n_out[SMB_MLP_TC_MAX_LAYERS];   /* padded N of the UMMA (16, 64 or 128) */
  uint32_t w_off[SMB_MLP_TC_MAX_LAYERS];   /* byte offset of the K-major 128B-swizzled fp16 image */
  uint32_t b_off[SMB_MLP_TC_MAX_LAYERS];   /* byte offset of the fp32 bias row */
} smb_mlp_tc_layout;
int smb_mlp_tc_layout_for(int n_layers, const int* k_in, const int* n_out, smb_mlp_tc_layout* out);
int smb_mlp_tc_pack_host(const float* const* weights_host, const float* const* biases_host, const int* k_in,
                         const int* n_out, const smb_mlp_tc_layout* layout, void* blob_host);
/* planes_cl: channels-last (3,Hp,Wp,40) planes, fp32 (smb_scene_prepare) or -- planes_fp16 != 0 -- fp16
 * (smb_scene_prepare_half): half the gather traffic, which bounds this kernel; taps are blended in fp32. */
int smb_scene_prepare_half(const float* triplane, int Hp, int Wp, void* planes_cl_half, void* stream);
int smb_query_points_tc(const void* planes_cl, int planes_fp16, int Hp, int Wp, int align_corners, const void* mlp_blob_dev,
                        const smb_mlp_tc_layout* layout, float radius, float out0_bias, int sigmoid_vec,
                        const float* positions, int64_t n, float* out0_raw, float* out0_act, float* out_vec,
                        float* out_vec_act, void* stream);

/* ------------------------------------------------------------ SF3D variant
 * Stable Fast 3D ("Pro"): StableFast/sf3d/system.py:141-198, sf3d/models/network.py:148-208,
 * sf3d/models/isosurface.py:24-229.  Planes are (3,40,Hp,Wp) (Hp=Wp=384 in the shipped config);
 * make the channels-last copy with smb_scene_prepare(triplane, Hp, Wp, NULL, NULL, planes_cl, NULL, s).
 *
 * heads_blob (device, fp32, smb_sf3d_heads_floats() floats): for head in (density, vertex_offset):
 *   W0 (64,120) b0 (64) W1 (64,64) b1 (64) W2 (out,64) b2 (out), out = 1 then 3, each head
 *   zero-padded to a multiple of 4 floats (11972 + 12100 floats)
 *   = MaterialMLP state-dict heads.<name>.{0,2,4}.{weight,bias} (network.py:158-178).
 *
 * smb_sf3d_query_f32: SF3D.query_triplane (align_corners=True; positions (n,3) in (-radius,radius))
 * fused with the two heads; pass features_in (n,120) INSTEAD of positions to run the heads on
 * pre-computed features (MaterialMLP.forward).  Outputs, each optional:
 *   features_out (n,120); density_raw (n) = head + out_bias; density_act (n) = exp(density_raw)
 *   (trunc_exp forward, network.py:85); vertex_offset (n,3).
 */
int smb_sf3d_heads_floats(void);
int smb_sf3d_query_f32(const float* planes_cl, int Hp, int Wp, const float* heads_blob, float radius,
                       float density_out_bias, const float* positions, const float* features_in, int64_t n,
                       float* features_out, float* density_raw, float* density_act, float* vertex_offset,
                       void* stream);

/* Marching tetrahedra = MarchingTetrahedraHelper._forward (isosurface.py:144-203) on a STATIC tet grid:
 *   edges     (E,2) int32: the grid's unique edges, each (a<b), sorted lexicographically
 *             (= MarchingTetrahedraHelper.all_edges, isosurface.py:119-133);
 *   tets      (T,4) int32 (npz "indices");  tet_edges (T,6) int32: index into `edges` of each tet's
 *             edges in base_tet_edges order (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) (isosurface.py:67).
 * Vertex k is the k-th crossing edge of `edges` -- the numbering torch.unique(dim=0) gives the
 * reference -- and faces list all 1-triangle tets, then all 2-triangle tets (isosurface.py:187-201).
 * Two calls: count (sizes the outputs), then emit.  smb_mtet_deform: grid + scale*tanh(offset)
 * (normalize_grid_deformation, isosurface.py:106-113). */
typedef struct smb_mtet_counts {
  int64_t nverts;
  int64_t ntris;
  int64_t ntris1; /* triangles from 1-triangle tets (they come first) */
  int64_t reserved;
} smb_mtet_counts;
size_t smb_mtet_workspace_bytes(int64_t n_edges, int64_t n_tets);
int smb_mtet_count(const float* sdf, const int32_t* edges, int64_t n_edges, const int32_t* tets, int64_t n_tets,
                   void* workspace, size_t workspace_bytes, smb_mtet_counts* counts_dev, void* stream);
int smb_mtet_emit(const float* positions, const float* sdf, const int32_t* edges, int64_t n_edges,
                  const int32_t* tet_edges, int64_t n_tets, const void* workspace, float* verts, int64_t* faces,
                  void* stream);
/* smb_mtet_emit with the caller's scale_tensor(v_pos, points_range, bbox) (sf3d/system.py:162-164) applied to every
 * vertex in the same fp32 operation order: v = ((v - affine8[0]) / affine8[1]) * affine8[2+c] + affine8[5+c]
 * (affine8: host array {in_lo, in_hi - in_lo, (bbox_hi - bbox_lo)[3], bbox_lo[3]}). */
int smb_mtet_emit_affine(const float* positions, const float* sdf, const int32_t* edges, int64_t n_edges,
                         const int32_t* tet_edges, int64_t n_tets, const void* workspace, float* verts, int64_t* faces,
                         const float* affine8, void* stream);
int smb_mtet_deform(const float* base, const float* deform, float scale, int64_t n_vertices, float* out, void* stream);

/* SF3D.query_triplane + decoder(include=[...]) (sf3d/system.py:153-154, 170-198; sf3d/models/network.py:148-208) at the
 * vertices of a LATTICE-ordered tet grid: grid_vertices[(a*nB + b)*nC + c] = (coord_0[a], coord_1[b], coord_2[c]) up to a
 * permutation of the spatial axes (the caller detects this once when it loads the grid, isosurface.py:71-81).  Every vertex
 * then samples each plane at a position given by TWO lattice indices, so layer 0 of a head is the sum of three table rows,
 *     W0 . [f_xy; f_xz; f_yz] + b0 = C[a][b] + T1[a][c] + T2[b][c],
 * built per call in fp32 (n^2 bilinear interpolations per table instead of n^3 per plane); the hidden layers run on tcgen05
 * (fp16 operands, fp32 accumulate) and the last Linear as fp32 dot products.  Up to 2 heads per call.
 *   planes_cl       (3,Hp,Wp,40) fp32 channels-last (smb_scene_prepare)
 *   decoder_blobs[h], layouts[h]   the head packed with smb_decoder_pack_host: Linear(120,64), (n_hidden-1) x Linear(64,64),
 *                   last Linear padded to 4 rows (n_out[h] = 1..3 of them are read)
 *   exp_act[h], out_bias[h]        1: out = exp(x + out_bias) (trunc_exp forward, config.yaml density head); 0: out = x
 *   out_sub (may be NULL)          out_sub[h] is subtracted from an exp-activated head AFTER the activation, in fp32: the caller's
 *                   `density - isosurface_threshold` (sf3d/system.py:155) without a separate elementwise pass
 *   axis_u[k]       device array (extents[k]) of lattice index k's coordinate already mapped to (-1,1) with the reference's
 *                   own scale_tensor ops; spatial_dim[k] in {0,1,2} says which of x,y,z it is (a permutation)
 *   outs[h]         (extents[0]*extents[1]*extents[2], n_out[h]) fp32, vertex order of the grid */
int smb_query_tetgrid_tc(const float* planes_cl, int Hp, int Wp, int align_corners, int nheads, const void* const* decoder_blobs,
                         const smb_decoder_layout* const* layouts, const int* n_out, const int* exp_act, const float* out_bias,
                         const float* out_sub, const float* const* axis_u, const int* extents, const int* spatial_dim,
                         float* const* outs, void* stream);

/* ------------------------------------------------------------ volume rendering
 * The two elementwise stages of TriplaneNeRFRenderer._forward (tsr/models/nerf_renderer.py:93-152) around the field
 * query (smb_query_points_tc / smb_query_points_f32 at the positions produced here):
 *   smb_ray_sample_positions: positions (n_rays, n_samples, 3) = rays_o + z * rays_d with
 *     z = t_near * (1 - t_mid[s]) + t_far * t_mid[s] (:110-117), in the reference's fp32 operation order;
 *     t_near / t_far (n_rays) from rays_intersect_bbox (tsr/utils.py:115-149), t_mid (n_samples) the bin centres (:108-109).
 *   smb_ray_composite: alpha = 1 - exp(-deltas[s] * density_act), weights = alpha * cumprod(1 - alpha + 1e-10)
 *     (exclusive), comp_rgb (n_rays,3) = sum_s w c + (1 - sum_s w) (white background, :125-150); opacity optional. */
int smb_ray_sample_positions(const float* rays_o, const float* rays_d, const float* t_near, const float* t_far,
                             const float* t_mid, int64_t n_rays, int n_samples, float* positions, void* stream);
int smb_ray_composite(const float* density_act, const float* color, const float* deltas, int64_t n_rays, int n_samples,
                      float* comp_rgb, float* opacity, void* stream);

/* ------------------------------------------------------------ texture baking (SF3D)
 * TextureBaker.rasterize / interpolate (StableFast/sf3d/texture_baker/baker.py:12-118).  The reference calls
 * rasterize_cpu / interpolate_cpu inside a Windows-only DLL without source; the Python functions of the same name
 * beside it (texture_baker/common.py:104-142,214-230) are the algorithm restated here:
 *   smb_bake_rasterize: rast (res,res,4) fp32; texel (y,x) is the point (x/res, 1 - y/res); (u, v, w, triangle index)
 *     of the triangle of the UV atlas that contains it (barycentrics in fp32 as common.py:104-121), else (0,0,0,-1).
 *     Texels covered by several triangles (shared edges) take the lowest index (the reference: first BVH hit).
 *     uv (nverts,2) fp32, faces (nfaces,3) int32 (baker.py:44), workspace of smb_bake_workspace_bytes(res).
 *   smb_bake_interpolate: out (res,res,channels) = attr[i0]*u + attr[i1]*v + attr[i2]*w (fp32), 0 where no triangle. */
size_t smb_bake_workspace_bytes(int resolution);
int smb_bake_rasterize(const float* uv, const int32_t* faces, int64_t nverts, int64_t nfaces, int resolution, float* rast,
                       void* workspace, size_t workspace_bytes, void* stream);
int smb_bake_interpolate(const float* attr, int channels, const int32_t* faces, int64_t nfaces, const float* rast,
                         int resolution, float* out, void* stream);

/* ------------------------------------------------------------ mesh hand-off
 * What the reference's sink TSR.import_obj_blender (tsr/system.py:127-168) needs from the mesh, produced on the
 * device so the sink can use Blender's bulk foreach_set instead of its per-loop Python assignment (:143-146):
 *   smb_mesh_loop_colors: loop_colors (3*ntris, 4) fp32, row 3*f + c = (vertex_colors[faces[f][c]], alpha) --
 *     from_pydata numbers polygon loops 3*f + c; alpha = 1 is the column system.py:133-135 appends.
 *     bad_index_flag (optional device int, caller zeroes it) is set when a face index is outside [0, nverts).
 *   smb_mesh_faces_i32: faces narrowed to int32 (MeshLoop.vertex_index is a 32-bit int);
 *   smb_mesh_faces_i64: the reverse (faces emitted / gathered as int32 widened to the reference's LongTensor width;
 *     both pointers 16-byte aligned). */
int smb_mesh_loop_colors(const float* vertex_colors, const int64_t* faces, int64_t nverts, int64_t ntris, float alpha,
                         float* loop_colors, int* bad_index_flag, void* stream);
int smb_mesh_faces_i32(const int64_t* faces, int64_t ntris, int32_t* faces_i32, void* stream);
int smb_mesh_faces_i64(const int32_t* faces_i32, int64_t ntris, int64_t* faces, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCULPTMATE_B200_H */

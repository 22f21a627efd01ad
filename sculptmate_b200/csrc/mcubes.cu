// Marching cubes on sm_100a: signs -> count+scan -> emit, deterministic and
// duplicate-free.  Replaces the CPU call skimage.measure.marching_cubes(level, 0.0)
// and the wrapper post-processing in MarchingCubeHelper.forward
// (/root/reference/TripoSR/tsr/models/isosurface.py:41-54).
//
// HBM-bound integer/bit work: no tensor cores.  The unit of work is a "word" =
// 32 consecutive z-samples of an (x,y) row.
//
//   K1 mc_signs  : streams the density slab ONCE (coalesced 128-byte row segments,
//                  8 independent loads in flight per lane) and reduces it 32:1 to sign
//                  bit masks with warp ballots.  Everything downstream works on the
//                  masks (R^3/8 bytes, L2 resident).
//   K2 mc_count  : one thread per word, one CTA per chunk of 1024 words of a plane:
//                  crossing masks by XOR of neighbouring sign masks, owned-vertex and
//                  triangle counts (popc / case table), CTA-wide exclusive scan; writes
//                  a 16-byte record per word + the chunk totals.
//   K3 mc_totals : one CTA turns chunk totals into chunk bases in canonical order.
//   K4 mc_emit   : a warp takes 32 consecutive words, compacts their crossing samples and
//                  their active cells across the warp (popc + warp scan), and then works
//                  one ITEM per lane: a crossing sample writes its <= 3 owned vertices
//                  (each lattice edge is owned by exactly one sample -> no duplicates, no
//                  atomics), an active cell writes its <= 5 triangles; vertex ids of
//                  neighbouring words come from prefix[word] + popc(mask & lanes_below).
//                  The surface is sparse (a few % of the cells), so lane-per-item keeps the
//                  SIMT lanes full where lane-per-sample left > 80 % of them idle.  Density
//                  is re-read only at the two end points of crossing edges.
//
// Canonical order (identical to oracle/mc_oracle.c): vertices by x-plane i, inside a
// plane first the in-plane crossings by (j,k) with the y-edge before the z-edge of a
// sample, then the x-edge crossings (plane i -> i+1) by (j,k); triangles by cell
// (i,j,k) then table order.  Slabs concatenate bit-exactly (see smb_mc_emit).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sculptmate_b200.h"
#define SMB_TABLE_QUAL __device__ const
#include "mc_tables.h"

namespace smb {

constexpr int kChunkWords = 1024;  // words per count CTA (256 threads x 4): 512 CTAs at 256^3 (2048 = 256 CTAs left the 148 SMs 1.7 waves)
constexpr int kWordsPerThread = 4;

struct McDims {
  int nx, ny, nz, wz;   // wz = words per z-row
  int pw;               // words per x-plane = ny*wz
  int cpp;              // chunks per plane
  long long nwords;     // nx*pw
  long long nchunks;    // nx*cpp
};

__host__ __device__ inline McDims make_dims(int nx, int ny, int nz) {
  McDims d;
  d.nx = nx;
  d.ny = ny;
  d.nz = nz;
  d.wz = (nz + 31) / 32;
  d.pw = ny * d.wz;
  d.cpp = (d.pw + kChunkWords - 1) / kChunkWords;
  d.nwords = (long long)nx * d.pw;
  d.nchunks = (long long)nx * d.cpp;
  return d;
}

// per word: chunk-local exclusive prefixes of (in-plane vertices, x-edge vertices,
// triangles), an info word (bit 0: the word owns a vertex or has a triangle), the crossing
// masks of the edges its samples own and the mask of its cells that produce triangles.
struct __align__(16) WordRec {
  uint32_t a, b, t, info;
  uint32_t mx, my, mz, act;
};
// per chunk: totals (K2) -> global exclusive bases (K3), same field meaning.
struct __align__(16) ChunkRec {
  uint32_t a, b, t, pad;
};

struct McWorkspace {
  uint32_t* pos;     // nwords sign masks
  WordRec* rec;      // nwords
  ChunkRec* ctot;    // nchunks totals
  ChunkRec* cbase;   // nchunks bases
  size_t bytes;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline McWorkspace carve(void* base, const McDims& d) {
  McWorkspace w;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* p = b ? b + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const size_t nw = (size_t)d.nwords;
  w.pos = reinterpret_cast<uint32_t*>(take((nw + 1) * 4));
  w.rec = reinterpret_cast<WordRec*>(take(nw * sizeof(WordRec)));
  w.ctot = reinterpret_cast<ChunkRec*>(take((size_t)d.nchunks * sizeof(ChunkRec)));
  w.cbase = reinterpret_cast<ChunkRec*>(take((size_t)d.nchunks * sizeof(ChunkRec)));
  w.bytes = off;
  return w;
}

__device__ __forceinline__ float mc_val(const float* __restrict__ g, long long idx, float sub, float sign) {
  return __fmul_rn(__fsub_rn(__ldg(g + idx), sub), sign);
}

// streaming load: the slab is read exactly once by K1, do not keep it in L1
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------ K1 signs
// One warp reduces 8 words per iteration: 8 independent coalesced 128-byte loads in
// flight per warp, then 8 ballots.  Persistent grid-stride over groups of 8 words.
__global__ void __launch_bounds__(256) mc_signs(const float* __restrict__ grid, McDims d, float sub, float sign,
                                                uint32_t* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long ngroups = (d.nwords + 7) / 8;
  const bool dense = (d.nz & 31) == 0;  // rows are whole words: the slab is one flat run of words
  for (long long g = warp; g < ngroups; g += nwarps) {
    const long long w0 = g * 8;
    float v[8];
    bool ok[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long word = w0 + e;
      ok[e] = word < d.nwords;
      long long idx;
      if (dense) {
        idx = word * 32 + lane;
      } else {
        const long long row = word / d.wz;
        const int k = (int)(word - row * d.wz) * 32 + lane;
        ok[e] = ok[e] && k < d.nz;
        idx = row * d.nz + k;
      }
      v[e] = ok[e] ? ld_stream(grid + idx) : 0.0f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool bit = ok[e] && __fmul_rn(__fsub_rn(v[e], sub), sign) > 0.0f;
      const uint32_t m = __ballot_sync(0xffffffffu, bit);
      if (lane == e && w0 + e < d.nwords) pos[w0 + e] = m;
    }
  }
}

// cube-case dump (parity/debug): one thread per cell
__global__ void mc_cases_kernel(const float* __restrict__ grid, int nx, int ny, int nz, float sub, float sign,
                                unsigned char* __restrict__ cases) {
  const long long ncell = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  const long long sy = nz, sx = (long long)ny * nz;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ncell;
       t += (long long)gridDim.x * blockDim.x) {
    int k = (int)(t % (nz - 1));
    long long r = t / (nz - 1);
    int j = (int)(r % (ny - 1));
    int i = (int)(r / (ny - 1));
    const long long p = (long long)i * sx + (long long)j * sy + k;
    uint32_t cs = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const long long q = p + ((c >> 2) & 1) * sx + ((c >> 1) & 1) * sy + (c & 1);
      cs |= (mc_val(grid, q, sub, sign) > 0.0f ? 1u : 0u) << c;
    }
    cases[t] = (unsigned char)cs;
  }
}

// ------------------------------------------------- masks of one word
// Sign masks around word (i,j,w): rows (di,dj) in {0,1}^2 at "k" alignment (m) and
// "k+1" alignment (s), plus validity.  All loads hit the L2-resident mask array.
struct WordMasks {
  uint32_t m[4], s[4];  // row = di*2 + dj
  uint32_t vin, vz;     // lanes with a sample / lanes with k+1 < nz
  bool hx, hy;
};

__device__ __forceinline__ WordMasks load_masks(const uint32_t* __restrict__ pos, const McDims& d, int i, int j,
                                                int w) {
  WordMasks r;
  r.hx = i + 1 < d.nx;
  r.hy = j + 1 < d.ny;
  const bool hw = w + 1 < d.wz;
  const long long base = ((long long)i * d.ny + j) * d.wz + w;
#pragma unroll
  for (int row = 0; row < 4; ++row) {
    const int di = row >> 1, dj = row & 1;
    uint32_t cur = 0u, nxt = 0u;
    if ((di == 0 || r.hx) && (dj == 0 || r.hy)) {
      const long long wd = base + (long long)di * d.pw + (long long)dj * d.wz;
      cur = __ldg(pos + wd);
      if (hw) nxt = __ldg(pos + wd + 1);
    }
    r.m[row] = cur;
    r.s[row] = (cur >> 1) | ((nxt & 1u) << 31);
  }
  const int nin = min(32, d.nz - w * 32);
  r.vin = nin >= 32 ? 0xffffffffu : ((1u << nin) - 1u);
  const int nz1 = min(32, d.nz - 1 - w * 32);
  r.vz = nz1 >= 32 ? 0xffffffffu : (nz1 <= 0 ? 0u : ((1u << nz1) - 1u));
  return r;
}

// crossing masks of the edges OWNED by the samples of word (i,j,w)
__device__ __forceinline__ void owned_masks(const WordMasks& k, uint32_t& mx, uint32_t& my, uint32_t& mz) {
  mx = k.hx ? ((k.m[0] ^ k.m[2]) & k.vin) : 0u;
  my = k.hy ? ((k.m[0] ^ k.m[1]) & k.vin) : 0u;
  mz = (k.m[0] ^ k.s[0]) & k.vz;
}

// lanes whose cell (i,j,k) has corners of both signs
__device__ __forceinline__ uint32_t active_cells(const WordMasks& k) {
  if (!(k.hx && k.hy)) return 0u;
  const uint32_t a = k.m[0];
  return ((a ^ k.s[0]) | (a ^ k.m[1]) | (a ^ k.s[1]) | (a ^ k.m[2]) | (a ^ k.s[2]) | (a ^ k.m[3]) | (a ^ k.s[3])) &
         k.vz;
}

__device__ __forceinline__ uint32_t cell_case(const WordMasks& k, int lane) {
  return ((k.m[0] >> lane) & 1u) | (((k.s[0] >> lane) & 1u) << 1) | (((k.m[1] >> lane) & 1u) << 2) |
         (((k.s[1] >> lane) & 1u) << 3) | (((k.m[2] >> lane) & 1u) << 4) | (((k.s[2] >> lane) & 1u) << 5) |
         (((k.m[3] >> lane) & 1u) << 6) | (((k.s[3] >> lane) & 1u) << 7);
}

// ------------------------------------------------------- K2 count + scan
// blockIdx.x = i*cpp + c : chunk c of plane i.  Thread t owns kWordsPerThread consecutive words
// c*kChunkWords + kWordsPerThread*t ... of the plane (blocked, so the scan order is the canonical order).
// The three counters are packed into one 64-bit lane (21/21/22 bits: a chunk holds at
// most kChunkWords*64 in-plane vertices, kChunkWords*32 x-edge vertices, kChunkWords*160 triangles).
__global__ void __launch_bounds__(256) mc_count(const uint32_t* __restrict__ pos, McDims d, WordRec* __restrict__ rec,
                                                ChunkRec* __restrict__ ctot) {
  __shared__ unsigned char s_ntri[256];
  __shared__ unsigned long long s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  s_ntri[tid] = SMB_MC_NTRI[tid];
  __syncthreads();
  const int i = (int)(blockIdx.x / d.cpp);
  const int c = (int)(blockIdx.x - (long long)i * d.cpp);
  const int jw0 = c * kChunkWords + tid * kWordsPerThread;

  unsigned long long local[kWordsPerThread + 1];
  uint32_t kmx[kWordsPerThread], kmy[kWordsPerThread], kmz[kWordsPerThread], kact[kWordsPerThread];
  unsigned long long sum = 0;
#pragma unroll
  for (int e = 0; e < kWordsPerThread; ++e) {
    const int jw = jw0 + e;
    unsigned long long cnt = 0;
    kmx[e] = kmy[e] = kmz[e] = kact[e] = 0u;
    if (jw < d.pw) {
      const int j = jw / d.wz, w = jw - j * d.wz;
      const WordMasks k = load_masks(pos, d, i, j, w);
      uint32_t mx, my, mz;
      owned_masks(k, mx, my, mz);
      uint32_t act = active_cells(k);
      kmx[e] = mx;
      kmy[e] = my;
      kmz[e] = mz;
      kact[e] = act;
      uint32_t nt = 0;
      while (act) {
        const int b = __ffs(act) - 1;
        act &= act - 1;
        nt += s_ntri[cell_case(k, b)];
      }
      cnt = (unsigned long long)(__popc(my) + __popc(mz)) | ((unsigned long long)__popc(mx) << 21) |
            ((unsigned long long)nt << 42);
    }
    local[e] = sum;  // exclusive within the thread
    sum += cnt;
  }
  local[kWordsPerThread] = sum;
  // CTA-wide exclusive scan of the per-thread sums
  unsigned long long inc = sum;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= s) inc += y;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned long long wbase = 0, total = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const unsigned long long v = s_warp[q];
    if (q < warp) wbase += v;
    total += v;
  }
  const unsigned long long tbase = wbase + inc - sum;
#pragma unroll
  for (int e = 0; e < kWordsPerThread; ++e) {
    const int jw = jw0 + e;
    if (jw < d.pw) {
      const unsigned long long p = tbase + local[e];
      WordRec r;
      r.a = (uint32_t)(p & 0x1fffffu);
      r.b = (uint32_t)((p >> 21) & 0x1fffffu);
      r.t = (uint32_t)(p >> 42);
      r.info = (local[e + 1] != local[e]) ? 1u : 0u;
      r.mx = kmx[e];
      r.my = kmy[e];
      r.mz = kmz[e];
      r.act = kact[e];
      rec[(long long)i * d.pw + jw] = r;
    }
  }
  if (tid == 0) {
    ChunkRec cr;
    cr.a = (uint32_t)(total & 0x1fffffu);
    cr.b = (uint32_t)((total >> 21) & 0x1fffffu);
    cr.t = (uint32_t)(total >> 42);
    cr.pad = 0;
    ctot[blockIdx.x] = cr;
  }
}

// ------------------------------------------------------------ K3 totals
// One CTA: exclusive scan of the chunk totals in canonical order.  Vertices: for each
// plane i, the in-plane counts of its chunks, then the x-edge counts of its chunks.
// Triangles: chunk order.  Sequential over tiles of 1024 with a running carry.
__global__ void __launch_bounds__(1024) mc_totals(const ChunkRec* __restrict__ ctot, ChunkRec* __restrict__ cbase,
                                                  McDims d, int emit_last_plane, smb_mc_counts* __restrict__ counts) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t totals[2];
  for (int a = 0; a < 2; ++a) {
    const long long n = a ? d.nchunks : 2 * d.nchunks;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
      const long long idx = base + tid;
      uint32_t v = 0;
      long long chunk = 0;
      int grp = 0;
      if (idx < n) {
        if (a) {
          chunk = idx;
          v = ctot[chunk].t;
        } else {
          const long long i = idx / (2 * d.cpp);
          const int r = (int)(idx - i * 2 * d.cpp);
          grp = r / d.cpp;
          chunk = i * d.cpp + (r - grp * d.cpp);
          v = grp ? ctot[chunk].b : ctot[chunk].a;
        }
      }
      uint32_t inc = v;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += y;
      }
      if (lane == 31) warp_sums[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        uint32_t ws = warp_sums[lane];
        uint32_t winc = ws;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          uint32_t y = __shfl_up_sync(0xffffffffu, winc, s);
          if (lane >= s) winc += y;
        }
        warp_sums[lane] = winc - ws;  // exclusive warp base
      }
      __syncthreads();
      const uint32_t excl = carry_s + warp_sums[warp] + inc - v;
      if (idx < n) {
        if (a) cbase[chunk].t = excl;
        else if (grp) cbase[chunk].b = excl;
        else cbase[chunk].a = excl;
      }
      __syncthreads();
      if (tid == 1023) carry_s = excl + v;
      __syncthreads();
    }
    totals[a] = carry_s;
    __syncthreads();
  }
  if (tid == 0) {
    long long stored = totals[0];
    // a slab that is not the last one numbers, but does not store, the in-plane
    // crossings of its last plane (they are the first vertices of the next slab)
    if (!emit_last_plane) stored = cbase[(long long)(d.nx - 1) * d.cpp].a;
    counts->nverts = stored;
    counts->ntris = totals[1];
    counts->nverts_numbered = totals[0];
    counts->reserved = 0;
  }
}

// ---------------------------------------------------------------- K4 emit
struct EmitParams {
  const float* grid;
  McDims d;
  float sub, sign;
  int x_origin, emit_last_plane, flags;
  float vdiv, vmul, vadd;
  long long id_offset;
  const uint32_t* pos;
  const WordRec* rec;
  const ChunkRec* cbase;
  float* verts;
  void* faces;  // (F,3) int64, or int32 with SMB_MC_FACES_I32
  long long vcap, fcap;  // capacity of verts / faces in vertices / triangles (writes beyond are dropped)
  // gather mode: (world,4) int64 = every rank's smb_mc_counts; this slab's output offsets are the
  // sums over the lower ranks, and verts/faces point at the destination rank's buffers (peer memory)
  const long long* all_counts;
  int rank;
  // peer-flag mode: all_counts lives in this rank's control block and is valid for the lower ranks once
  // count_seq[g] == seq (written over NVLink by rank g's smb_peer_publish_counts)
  const long long* count_seq;
  long long seq;
  long long* error_flag;
};

__device__ __forceinline__ float mc_xform(float v, int flags, float vdiv, float vmul, float vadd) {
  if (flags & SMB_MC_DIV) v = __fdiv_rn(v, vdiv);
  if (flags & SMB_MC_AFFINE) v = __fadd_rn(__fmul_rn(v, vmul), vadd);
  return v;
}

constexpr int kEmitWarps = 8;

// lane l holds cnt_l items; returns the exclusive prefix and the warp total
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
  uint32_t inc = v;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, sft);
    if (lane >= sft) inc += y;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// item idx of the batch -> (lane that holds its word, position of its bit in that word's mask)
__device__ __forceinline__ void locate_item(uint32_t idx, const uint32_t* __restrict__ s_off /*[33]*/, int& src) {
  int lo = 0;  // largest l with s_off[l] <= idx
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
    if (s_off[lo + step] <= idx) lo += step;
  src = lo;
}

// kStaged (the host picks it when a batch of 32 words never straddles an x-plane, i.e. ny*wz % 32 == 0): the vertices and
// triangles of a 32-item group occupy CONTIGUOUS output slots, so they are assembled in shared memory and written out
// by the whole warp with coalesced 128-byte stores instead of scattered 4-byte stores per lane -- what NVLink peer
// stores (gather mode) and HBM sector writes both want.
template <bool kStaged>
__global__ void __launch_bounds__(kEmitWarps * 32, 4) mc_emit(EmitParams p) {
  __shared__ float s_vst[kStaged ? kEmitWarps : 1][96 * 3];       // 64 in-plane + 32 x-edge vertices of a group
  __shared__ uint32_t s_fst[kStaged ? kEmitWarps : 1][160 * 3];   // triangles of a group (slab-local vertex ids)
  __shared__ signed char s_tri[256][16];
  __shared__ unsigned char s_ntri[256];
  __shared__ unsigned short s_emask[256];
  __shared__ uint32_t s_off[kEmitWarps][33];
  __shared__ uint32_t s_eid[12][kEmitWarps * 32];  // per-thread column of the 12 edge vertex ids
  for (int t = threadIdx.x; t < 256 * 16; t += blockDim.x) (&s_tri[0][0])[t] = (&SMB_MC_TRI[0][0])[t];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) {
    s_ntri[t] = SMB_MC_NTRI[t];
    s_emask[t] = SMB_MC_EDGEMASK[t];
  }
  __syncthreads();

  const McDims d = p.d;
  long long v_off = 0, f_off = 0;
  if (p.all_counts) {
    if (p.count_seq) {
      // the lower ranks' counts arrive as peer stores: one thread per CTA waits for their sequence flags
      if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int g = 0; g < p.rank; ++g) {
          while (*reinterpret_cast<const volatile long long*>(p.count_seq + g) != p.seq) {
            if (clock64() - t0 > 6000000000LL) {  // ~3 s: a peer died; report instead of hanging the GPU
              if (p.error_flag) *p.error_flag = 1;
              break;
            }
            __nanosleep(200);
          }
        }
        __threadfence_system();
      }
      __syncthreads();
    }
    for (int g = 0; g < p.rank; ++g) {
      v_off += *reinterpret_cast<const volatile long long*>(p.all_counts + 4 * g + 0);
      f_off += *reinterpret_cast<const volatile long long*>(p.all_counts + 4 * g + 1);
    }
  }
  const long long id_off = p.id_offset + v_off;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long gwarp = (long long)blockIdx.x * kEmitWarps + warp;
  const long long nwarps = (long long)gridDim.x * kEmitWarps;
  const long long nbatch = (d.nwords + 31) / 32;
  const long long sy = d.nz, sx = (long long)d.ny * d.nz;
  uint32_t* off = s_off[warp];

  for (long long batch = gwarp; batch < nbatch; batch += nwarps) {
    // ---- phase 0: lane <-> word --------------------------------------------------
    const long long wd = batch * 32 + lane;
    uint32_t info = 0;
    if (wd < d.nwords) info = __ldg(&p.rec[wd].info);
    if (__ballot_sync(0xffffffffu, info & 1u) == 0u) continue;  // warp-uniform: nothing in these 32 words
    uint32_t mx = 0, my = 0, mz = 0, act = 0, v0 = 0, v1 = 0, t0 = 0;
    int wi = 0, wj = 0, ww = 0;
    WordMasks km;
#pragma unroll
    for (int r = 0; r < 4; ++r) km.m[r] = km.s[r] = 0u;
    if (info & 1u) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(&p.rec[wd]));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(&p.rec[wd]) + 1);
      mx = hi.x;
      my = hi.y;
      mz = hi.z;
      act = hi.w;
      wi = (int)(wd / d.pw);
      const int jw = (int)(wd - (long long)wi * d.pw);
      wj = jw / d.wz;
      ww = jw - wj * d.wz;
      const uint4 cb = __ldg(reinterpret_cast<const uint4*>(&p.cbase[(long long)wi * d.cpp + jw / kChunkWords]));
      v0 = cb.x + lo.x;
      v1 = cb.y + lo.y;
      t0 = cb.z + lo.z;
      if (act) km = load_masks(p.pos, d, wi, wj, ww);
    }

    // ---- phase V: one lane per crossing sample -> its owned vertices -------------
    {
      const uint32_t own = mx | my | mz;
      uint32_t total;
      const uint32_t ex = warp_excl_scan(__popc(own), lane, total);
      __syncwarp();
      off[lane] = ex;
      if (lane == 0) off[32] = 0xffffffffu;
      __syncwarp();
      for (uint32_t base = 0; base < total; base += 32) {  // warp-uniform
        const uint32_t idx = base + lane;
        const bool valid = idx < total;
        int src = 0;
        if (valid) locate_item(idx, off, src);
        const uint32_t smx = __shfl_sync(0xffffffffu, mx, src), smy = __shfl_sync(0xffffffffu, my, src);
        const uint32_t smz = __shfl_sync(0xffffffffu, mz, src);
        const uint32_t sv0 = __shfl_sync(0xffffffffu, v0, src), sv1 = __shfl_sync(0xffffffffu, v1, src);
        const int si = __shfl_sync(0xffffffffu, wi, src), sj = __shfl_sync(0xffffffffu, wj, src);
        const int sw = __shfl_sync(0xffffffffu, ww, src);
        const uint32_t sex = __shfl_sync(0xffffffffu, ex, src);
        uint32_t sl_in = 0, sl_x = 0, c_in = 0, c_x = 0;
        int bit = 0;
        bool by = false, bz = false, bx = false;
        if (valid) {
          const uint32_t sown = smx | smy | smz;
          bit = __fns(sown, 0, (int)(idx - sex) + 1);
          const uint32_t below = (1u << bit) - 1u;
          by = (smy >> bit) & 1u, bz = (smz >> bit) & 1u, bx = (smx >> bit) & 1u;
          sl_in = sv0 + __popc(smy & below) + __popc(smz & below);
          sl_x = sv1 + __popc(smx & below);
          c_in = (by ? 1u : 0u) + (bz ? 1u : 0u);
          c_x = bx ? 1u : 0u;
        }
        // kStaged: the slots of a group are consecutive from lane 0's (lane 0 always holds an item)
        const uint32_t b_in = __shfl_sync(0xffffffffu, sl_in, 0), b_x = __shfl_sync(0xffffffffu, sl_x, 0);
        if (valid) {
          const int k = sw * 32 + bit;
          const long long pt = (long long)si * sx + (long long)sj * sy + k;
          const float a = mc_val(p.grid, pt, p.sub, p.sign);
          const float fi = (float)(p.x_origin + si), fj = (float)sj, fk = (float)k;
          const bool store_inplane = (si < d.nx - 1) || p.emit_last_plane;
          float* st = s_vst[kStaged ? warp : 0];
          if (by) {
            const float bb = mc_val(p.grid, pt + sy, p.sub, p.sign);
            const float t = __fdiv_rn(a, __fsub_rn(a, bb));
            float* o = kStaged ? st + 3 * (sl_in - b_in) : p.verts + 3 * (v_off + sl_in);
            if (kStaged || (store_inplane && v_off + sl_in < p.vcap)) {
              o[0] = mc_xform(fi, p.flags, p.vdiv, p.vmul, p.vadd);
              o[1] = mc_xform(__fadd_rn(fj, t), p.flags, p.vdiv, p.vmul, p.vadd);
              o[2] = mc_xform(fk, p.flags, p.vdiv, p.vmul, p.vadd);
            }
          }
          if (bz) {
            const float bb = mc_val(p.grid, pt + 1, p.sub, p.sign);
            const float t = __fdiv_rn(a, __fsub_rn(a, bb));
            const uint32_t sl = sl_in + (by ? 1u : 0u);
            float* o = kStaged ? st + 3 * (sl - b_in) : p.verts + 3 * (v_off + sl);
            if (kStaged || (store_inplane && v_off + sl < p.vcap)) {
              o[0] = mc_xform(fi, p.flags, p.vdiv, p.vmul, p.vadd);
              o[1] = mc_xform(fj, p.flags, p.vdiv, p.vmul, p.vadd);
              o[2] = mc_xform(__fadd_rn(fk, t), p.flags, p.vdiv, p.vmul, p.vadd);
            }
          }
          if (bx) {
            const float bb = mc_val(p.grid, pt + sx, p.sub, p.sign);
            const float t = __fdiv_rn(a, __fsub_rn(a, bb));
            float* o = kStaged ? st + 3 * (64 + sl_x - b_x) : p.verts + 3 * (v_off + sl_x);
            if (kStaged || v_off + sl_x < p.vcap) {
              o[0] = mc_xform(__fadd_rn(fi, t), p.flags, p.vdiv, p.vmul, p.vadd);
              o[1] = mc_xform(fj, p.flags, p.vdiv, p.vmul, p.vadd);
              o[2] = mc_xform(fk, p.flags, p.vdiv, p.vmul, p.vadd);
            }
          }
        }
        if (kStaged) {
          const int nlast = (int)min(32u, total - base) - 1;  // lane of the group's last item
          const uint32_t n_in = __shfl_sync(0xffffffffu, sl_in + c_in, nlast) - b_in;
          const uint32_t n_x = __shfl_sync(0xffffffffu, sl_x + c_x, nlast) - b_x;
          const int pl = __shfl_sync(0xffffffffu, si, 0);  // a group lies in one x-plane
          const bool st_in = (pl < d.nx - 1) || p.emit_last_plane;
          __syncwarp();
          const float* st = s_vst[warp];
          if (st_in) {
            const long long o0 = v_off + b_in;
            for (uint32_t t = lane; t < 3 * n_in; t += 32)
              if (o0 + t / 3 < p.vcap) p.verts[3 * o0 + t] = st[t];
          }
          {
            const long long o0 = v_off + b_x;
            for (uint32_t t = lane; t < 3 * n_x; t += 32)
              if (o0 + t / 3 < p.vcap) p.verts[3 * o0 + t] = st[3 * 64 + t];
          }
          __syncwarp();
        }
      }
    }

    // ---- phase T: one lane per active cell -> its triangles ----------------------
    {
      uint32_t total;
      const uint32_t ex = warp_excl_scan(__popc(act), lane, total);
      if (total == 0u) continue;  // warp-uniform
      __syncwarp();
      off[lane] = ex;
      if (lane == 0) off[32] = 0xffffffffu;
      __syncwarp();
      int carry_src = -1;       // word (lane) whose cells straddle the previous 32-item group ...
      uint32_t carry_sum = 0u;  // ... and the triangles it has emitted so far
      for (uint32_t base = 0; base < total; base += 32) {  // warp-uniform
        const uint32_t idx = base + lane;
        const bool valid = idx < total;
        int src = 0;
        if (valid) locate_item(idx, off, src);
        WordMasks sk;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          sk.m[r] = __shfl_sync(0xffffffffu, km.m[r], src);
          sk.s[r] = __shfl_sync(0xffffffffu, km.s[r], src);
        }
        const uint32_t sact = __shfl_sync(0xffffffffu, act, src);
        const uint32_t st0 = __shfl_sync(0xffffffffu, t0, src);
        const int si = __shfl_sync(0xffffffffu, wi, src), sj = __shfl_sync(0xffffffffu, wj, src);
        const int sw = __shfl_sync(0xffffffffu, ww, src);
        const uint32_t sex = __shfl_sync(0xffffffffu, ex, src);
        int bit = 0;
        uint32_t cs = 0, ntri = 0;
        if (valid) {
          bit = __fns(sact, 0, (int)(idx - sex) + 1);
          cs = cell_case(sk, bit);
          ntri = s_ntri[cs];
        }
        // triangles of the lower cells of the same word: segmented inclusive scan over the
        // lanes (items are sorted by word, then bit) + the carry of a straddling word
        const int key = valid ? src : -2;
        uint32_t inc = ntri;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, inc, sft);
          const int ky = __shfl_up_sync(0xffffffffu, key, sft);
          if (lane >= sft && ky == key) inc += y;
        }
        if (key == carry_src) inc += carry_sum;
        carry_src = __shfl_sync(0xffffffffu, key, 31);
        carry_sum = __shfl_sync(0xffffffffu, inc, 31);
        // slot of this cell's first triangle relative to the slab's first (32-bit: a slab holds < 2^32 triangles);
        // consecutive over the lanes of a group, so lane 0's is the start of the group's run
        const uint32_t rel_slot = st0 + (inc - ntri);
        const uint32_t rel_base = __shfl_sync(0xffffffffu, rel_slot, 0);
        if (ntri != 0u) {
          // ids of the vertices on this cell's crossing edges, grouped by the sample that
          // owns them: owner (di,dj,dk) holds x-edge 2dj+dk (di=0), y-edge 4+2di+dk (dj=0),
          // z-edge 8+2di+dj (dk=0)
          const uint32_t emask = s_emask[cs];
          const int jw = sj * d.wz + sw;
          // The four (di,dj) rows of the cell: the owners dk = 0 and dk = 1 of a row live in the same word
          // (bit, bit+1) unless bit == 31.  The record loads of two rows (one x-plane) are issued together and
          // unconditionally (an active cell has all four rows): two trips to L2 per cell instead of one per
          // owner (eight); all four rows at once would spill at the 64 registers that give 4 CTAs per SM.
#pragma unroll
          for (int di = 0; di < 2; ++di) {
            uint2 rlo[2], rcb[2];  // (a, b) prefixes of the word / bases of its chunk
            uint4 rhi[2];          // (mx, my, mz, act)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
              const int jw2 = jw + dj * d.wz;
              const long long w2 = (long long)(si + di) * d.pw + jw2;
              rlo[dj] = __ldg(reinterpret_cast<const uint2*>(&p.rec[w2]));
              rhi[dj] = __ldg(reinterpret_cast<const uint4*>(&p.rec[w2]) + 1);
              rcb[dj] = __ldg(reinterpret_cast<const uint2*>(&p.cbase[(long long)(si + di) * d.cpp + jw2 / kChunkWords]));
            }
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
#pragma unroll
              for (int dk = 0; dk < 2; ++dk) {
                uint32_t ebits = 0u;  // x-edge needs di == 0, y-edge dj == 0, z-edge dk == 0
                if (di == 0) ebits |= 1u << (2 * dj + dk);
                if (dj == 0) ebits |= 1u << (4 + 2 * di + dk);
                if (dk == 0) ebits |= 1u << (8 + 2 * di + dj);
                if (!(emask & ebits)) continue;
                uint2 lo = rlo[dj], cb = rcb[dj];
                uint4 hi = rhi[dj];
                int bb = bit + dk;
                if (bb == 32) {  // the owner is sample 0 of the next word of the row (rare)
                  const int jw2 = jw + dj * d.wz + 1;
                  const long long w2 = (long long)(si + di) * d.pw + jw2;
                  lo = __ldg(reinterpret_cast<const uint2*>(&p.rec[w2]));
                  hi = __ldg(reinterpret_cast<const uint4*>(&p.rec[w2]) + 1);
                  cb = __ldg(reinterpret_cast<const uint2*>(&p.cbase[(long long)(si + di) * d.cpp + jw2 / kChunkWords]));
                  bb = 0;
                }
                const uint32_t below = (1u << bb) - 1u;
                const uint32_t idyz = cb.x + lo.x + __popc(hi.y & below) + __popc(hi.z & below);
                if (di == 0) s_eid[2 * dj + dk][threadIdx.x] = cb.y + lo.y + __popc(hi.x & below);
                if (dj == 0) s_eid[4 + 2 * di + dk][threadIdx.x] = idyz;
                if (dk == 0) s_eid[8 + 2 * di + dj][threadIdx.x] = idyz + ((hi.y >> bb) & 1u);
              }
            }
          }
          const bool flip = p.flags & SMB_MC_FLIP;
          if (kStaged) {
            // slab-local ids of the triangles into the group's contiguous run in shared memory
            uint32_t* o = s_fst[warp] + 3 * (rel_slot - rel_base);
            for (uint32_t t = 0; t < ntri; ++t) {
              const uint32_t i0 = s_eid[s_tri[cs][3 * t + 0]][threadIdx.x];
              const uint32_t i1 = s_eid[s_tri[cs][3 * t + 1]][threadIdx.x];
              const uint32_t i2 = s_eid[s_tri[cs][3 * t + 2]][threadIdx.x];
              o[3 * t + 0] = flip ? i1 : i0;
              o[3 * t + 1] = flip ? i0 : i1;
              o[3 * t + 2] = i2;
            }
          } else {
            const long long slot0 = f_off + rel_slot;
            if (p.flags & SMB_MC_FACES_I32) {
              // int32 indices (what Blender's loop arrays and a PCIe / NVLink wire want); ids < 2^31 is checked on the host
              int* o = static_cast<int*>(p.faces) + 3 * slot0;
              const int ido = (int)id_off;
              for (uint32_t t = 0; t < ntri && slot0 + t < p.fcap; ++t) {
                const int i0 = (int)s_eid[s_tri[cs][3 * t + 0]][threadIdx.x] + ido;
                const int i1 = (int)s_eid[s_tri[cs][3 * t + 1]][threadIdx.x] + ido;
                const int i2 = (int)s_eid[s_tri[cs][3 * t + 2]][threadIdx.x] + ido;
                o[3 * t + 0] = flip ? i1 : i0;
                o[3 * t + 1] = flip ? i0 : i1;
                o[3 * t + 2] = i2;
              }
            } else {
              long long* o = static_cast<long long*>(p.faces) + 3 * slot0;
              for (uint32_t t = 0; t < ntri && slot0 + t < p.fcap; ++t) {
                const long long i0 = (long long)s_eid[s_tri[cs][3 * t + 0]][threadIdx.x] + id_off;
                const long long i1 = (long long)s_eid[s_tri[cs][3 * t + 1]][threadIdx.x] + id_off;
                const long long i2 = (long long)s_eid[s_tri[cs][3 * t + 2]][threadIdx.x] + id_off;
                o[3 * t + 0] = flip ? i1 : i0;
                o[3 * t + 1] = flip ? i0 : i1;
                o[3 * t + 2] = i2;
              }
            }
          }
        }
        if (kStaged) {
          // the whole warp writes the group's run: 128 contiguous bytes (int32) / 256 (int64) per store instruction
          const int nlast = (int)min(32u, total - base) - 1;
          const uint32_t n3 = 3u * (__shfl_sync(0xffffffffu, rel_slot + ntri, nlast) - rel_base);
          __syncwarp();
          const uint32_t* st = s_fst[warp];
          const long long o0 = f_off + rel_base;
          if (p.flags & SMB_MC_FACES_I32) {
            int* o = static_cast<int*>(p.faces) + 3 * o0;
            const int ido = (int)id_off;
            for (uint32_t t = lane; t < n3; t += 32)
              if (o0 + t / 3 < p.fcap) o[t] = (int)st[t] + ido;
          } else {
            long long* o = static_cast<long long*>(p.faces) + 3 * o0;
            for (uint32_t t = lane; t < n3; t += 32)
              if (o0 + t / 3 < p.fcap) o[t] = (long long)st[t] + id_off;
          }
          __syncwarp();
        }
      }
    }
  }
  if (p.all_counts) __threadfence_system();  // peer-memory stores are performed before the grid retires
}

// ----------------------------------------------------------- min / max
__global__ void __launch_bounds__(256) mc_minmax(const float* __restrict__ grid, long long n, float sub, float sign,
                                                 float* __restrict__ out /* [2], pre-set to +inf,-inf bits */) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float v = mc_val(grid, t, sub, sign);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  lo += 0.0f;  // canonicalise -0.0 -> +0.0 (the int ordering below would rank -0.0 below every negative)
  hi += 0.0f;
  if ((threadIdx.x & 31) == 0) {
    // float atomic min/max through the ordered-int trick
    int* o = reinterpret_cast<int*>(out);
    if (lo >= 0.0f) atomicMin(o + 0, __float_as_int(lo)); else atomicMax(reinterpret_cast<unsigned*>(o + 0), __float_as_uint(lo));
    if (hi >= 0.0f) atomicMax(o + 1, __float_as_int(hi)); else atomicMin(reinterpret_cast<unsigned*>(o + 1), __float_as_uint(hi));
  }
}

__global__ void mc_minmax_init(float* out) {
  out[0] = INFINITY;
  out[1] = -INFINITY;
}

}  // namespace smb

using namespace smb;

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

extern "C" size_t smb_mc_workspace_bytes(int nx, int ny, int nz) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
  McDims d = make_dims(nx, ny, nz);
  return carve(nullptr, d).bytes;
}

static int launch_count(const McDims& d, const McWorkspace& w, int emit_last_plane, smb_mc_counts* counts_dev, cudaStream_t st) {
  mc_count<<<(unsigned)d.nchunks, 256, 0, st>>>(w.pos, d, w.rec, w.ctot);
  mc_totals<<<1, 1024, 0, st>>>(w.ctot, w.cbase, d, emit_last_plane, counts_dev);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

namespace smb {
int launch_mc_signs(const float* grid, int nx, int ny, int nz, float sub, float sign, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
  if (!grid || !workspace || nx <= 0 || ny <= 0 || nz <= 0) return SMB_ERR_BAD_ARG;
  McDims d = make_dims(nx, ny, nz);
  McWorkspace w = carve(workspace, d);
  if (w.bytes > workspace_bytes) return SMB_ERR_WORKSPACE;
  // up to 8 CTAs of 8 warps per SM, each warp 8 words per iteration
  long long sg = (d.nwords + 63) / 64;
  if (sg > (long long)sm_count() * 8) sg = (long long)sm_count() * 8;
  mc_signs<<<(unsigned)sg, 256, 0, st>>>(grid, d, sub, sign, w.pos);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}
}  // namespace smb

extern "C" int smb_mc_count(const float* grid, int nx, int ny, int nz, float sub, float sign, int emit_last_plane,
                            void* workspace, size_t workspace_bytes, smb_mc_counts* counts_dev, void* stream) {
  if (!grid || !workspace || !counts_dev || nx <= 0 || ny <= 0 || nz <= 0) return SMB_ERR_BAD_ARG;
  McDims d = make_dims(nx, ny, nz);
  McWorkspace w = carve(workspace, d);
  if (w.bytes > workspace_bytes) return SMB_ERR_WORKSPACE;
  if (d.nchunks > 0x7fffffffLL) return SMB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = launch_mc_signs(grid, nx, ny, nz, sub, sign, workspace, workspace_bytes, st);
  if (rc != SMB_OK) return rc;
  return launch_count(d, w, emit_last_plane, counts_dev, st);
}

extern "C" int smb_mc_count_presigned(int nx, int ny, int nz, int emit_last_plane, void* workspace, size_t workspace_bytes,
                                      smb_mc_counts* counts_dev, void* stream) {
  if (!workspace || !counts_dev || nx <= 0 || ny <= 0 || nz <= 0) return SMB_ERR_BAD_ARG;
  McDims d = make_dims(nx, ny, nz);
  McWorkspace w = carve(workspace, d);
  if (w.bytes > workspace_bytes) return SMB_ERR_WORKSPACE;
  if (d.nchunks > 0x7fffffffLL) return SMB_ERR_BAD_ARG;
  return launch_count(d, w, emit_last_plane, counts_dev, (cudaStream_t)stream);
}

static int launch_emit(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                       int emit_last_plane, int flags, float vdiv, float vmul, float vadd,
                       int64_t vertex_id_offset, const void* workspace, float* verts, int64_t verts_capacity,
                       int64_t* faces, int64_t faces_capacity, const int64_t* all_counts, int rank, void* stream,
                       const int64_t* count_seq = nullptr, int64_t seq = 0, int64_t* error_flag = nullptr) {
  if (!grid || !workspace || nx <= 0 || ny <= 0 || nz <= 0 || verts_capacity < 0 || faces_capacity < 0) return SMB_ERR_BAD_ARG;
  McDims d = make_dims(nx, ny, nz);
  McWorkspace w = carve(const_cast<void*>(workspace), d);
  EmitParams p;
  p.grid = grid;
  p.d = d;
  p.sub = sub;
  p.sign = sign;
  p.x_origin = x_origin;
  p.emit_last_plane = emit_last_plane;
  p.flags = flags;
  p.vdiv = vdiv;
  p.vmul = vmul;
  p.vadd = vadd;
  p.id_offset = vertex_id_offset;
  p.pos = w.pos;
  p.rec = w.rec;
  p.cbase = w.cbase;
  p.verts = verts;
  p.faces = faces;
  p.vcap = verts ? verts_capacity : 0;
  p.fcap = faces ? faces_capacity : 0;
  p.all_counts = reinterpret_cast<const long long*>(all_counts);
  p.rank = rank;
  p.count_seq = reinterpret_cast<const long long*>(count_seq);
  p.seq = seq;
  p.error_flag = reinterpret_cast<long long*>(error_flag);
  const long long nbatch = (d.nwords + 31) / 32;
  long long blocks = (nbatch + kEmitWarps - 1) / kEmitWarps;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  // staged, coalesced output only on request (SMB_MC_COALESCE: destination is peer memory): on local HBM the extra
  // shared-memory pass costs more than the scattered 4-byte stores it replaces (measured 0.18 vs 0.15 ms at 256^3)
  if ((flags & SMB_MC_COALESCE) && d.pw % 32 == 0) mc_emit<true><<<(unsigned)blocks, kEmitWarps * 32, 0, (cudaStream_t)stream>>>(p);
  else mc_emit<false><<<(unsigned)blocks, kEmitWarps * 32, 0, (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

extern "C" int smb_mc_emit_bounded(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                                   int emit_last_plane, int flags, float vdiv, float vmul, float vadd,
                                   int64_t vertex_id_offset, const void* workspace, float* verts, int64_t verts_capacity,
                                   int64_t* faces, int64_t faces_capacity, void* stream) {
  return launch_emit(grid, nx, ny, nz, sub, sign, x_origin, emit_last_plane, flags, vdiv, vmul, vadd, vertex_id_offset, workspace,
                     verts, verts_capacity, faces, faces_capacity, nullptr, 0, stream);
}

extern "C" int smb_mc_emit_gather(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                                  int emit_last_plane, int flags, float vdiv, float vmul, float vadd, const void* workspace,
                                  const int64_t* all_counts_dev, int rank, float* verts_dst, int64_t verts_capacity,
                                  int64_t* faces_dst, int64_t faces_capacity, void* stream) {
  if (!all_counts_dev || rank < 0 || !verts_dst || !faces_dst) return SMB_ERR_BAD_ARG;
  return launch_emit(grid, nx, ny, nz, sub, sign, x_origin, emit_last_plane, flags, vdiv, vmul, vadd, 0, workspace, verts_dst,
                     verts_capacity, faces_dst, faces_capacity, all_counts_dev, rank, stream);
}

// ------------------------------------------------ peer control block (multi-GPU gather without collectives)
// Layout of a rank's control block (int64 words): counts[16][4] at 0, count_seq[16] at 64, done_seq[16] at 80,
// release_seq at 96, error at 97 (SMB_PEER_CTRL_WORDS = 128).  Every rank maps every block (CUDA IPC).
namespace smb {
constexpr int kPeerCounts = 0, kPeerCountSeq = 64, kPeerDoneSeq = 80, kPeerRelease = 96, kPeerError = 97;

__global__ void peer_publish_counts(const smb_mc_counts* __restrict__ mine, long long* const* __restrict__ peers, int rank, int world,
                                    long long seq) {
  const int g = threadIdx.x;
  if (g >= world) return;
  volatile long long* c = peers[g] + kPeerCounts + 4 * rank;
  c[0] = mine->nverts;
  c[1] = mine->ntris;
  c[2] = mine->nverts_numbered;
  c[3] = 0;
  __threadfence_system();
  *reinterpret_cast<volatile long long*>(peers[g] + kPeerCountSeq + rank) = seq;
}

__global__ void peer_signal_done(long long* dst_ctrl, int rank, long long seq) {
  __threadfence_system();  // everything this stream stored before (the emit kernel's peer stores) is performed first
  *reinterpret_cast<volatile long long*>(dst_ctrl + kPeerDoneSeq + rank) = seq;
}

// wait_done != 0 (destination rank): wait until every rank has stored its slab, then release the peers for the next
// call; wait_done == 0 (other ranks): wait until every rank's counts have arrived.  Either way sum the counts.
__global__ void peer_wait_all(long long* ctrl, long long* const* __restrict__ peers, int world, long long seq, int wait_done,
                              long long* totals) {
  const int g = threadIdx.x;
  const long long t0 = clock64();
  if (g < world) {
    const long long* flag = ctrl + (wait_done ? kPeerDoneSeq : kPeerCountSeq) + g;
    while (*reinterpret_cast<const volatile long long*>(flag) != seq) {
      if (clock64() - t0 > 6000000000LL) {
        ctrl[kPeerError] = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
  if (wait_done && g < world) *reinterpret_cast<volatile long long*>(peers[g] + kPeerRelease) = seq;
  if (g == 0) {
    long long v = 0, f = 0;
    for (int r = 0; r < world; ++r) {
      v += *reinterpret_cast<volatile long long*>(ctrl + kPeerCounts + 4 * r + 0);
      f += *reinterpret_cast<volatile long long*>(ctrl + kPeerCounts + 4 * r + 1);
    }
    totals[0] = v;
    totals[1] = f;
    totals[2] = ctrl[kPeerError];
    totals[3] = seq;
  }
}

// every rank, first kernel of a call: the destination has consumed call seq-1 (nobody overwrites counts / mesh
// buffers a slower rank or the destination's reader still needs)
__global__ void peer_wait_release(long long* ctrl, long long need) {
  const long long t0 = clock64();
  while (*reinterpret_cast<volatile long long*>(ctrl + kPeerRelease) < need) {
    if (clock64() - t0 > 6000000000LL) {
      ctrl[kPeerError] = 1;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}
}  // namespace smb

extern "C" int smb_peer_publish_counts(const smb_mc_counts* counts_dev, void* const* peer_ctrl_dev, int rank, int world, int64_t seq,
                                       void* stream) {
  if (!counts_dev || !peer_ctrl_dev || rank < 0 || world < 1 || world > 16 || rank >= world) return SMB_ERR_BAD_ARG;
  peer_publish_counts<<<1, 32, 0, (cudaStream_t)stream>>>(counts_dev, reinterpret_cast<long long* const*>(peer_ctrl_dev), rank, world, seq);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}
extern "C" int smb_peer_signal_done(void* dst_ctrl, int rank, int64_t seq, void* stream) {
  if (!dst_ctrl || rank < 0 || rank >= 16) return SMB_ERR_BAD_ARG;
  peer_signal_done<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<long long*>(dst_ctrl), rank, seq);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}
extern "C" int smb_peer_wait_all(void* ctrl_local, void* const* peer_ctrl_dev, int world, int64_t seq, int wait_done, int64_t* totals_dev,
                                 void* stream) {
  if (!ctrl_local || !peer_ctrl_dev || !totals_dev || world < 1 || world > 16) return SMB_ERR_BAD_ARG;
  peer_wait_all<<<1, 32, 0, (cudaStream_t)stream>>>(static_cast<long long*>(ctrl_local), reinterpret_cast<long long* const*>(peer_ctrl_dev), world,
                                                   seq, wait_done, reinterpret_cast<long long*>(totals_dev));
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}
extern "C" int smb_peer_wait_release(void* ctrl_local, int64_t need, void* stream) {
  if (!ctrl_local) return SMB_ERR_BAD_ARG;
  peer_wait_release<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<long long*>(ctrl_local), need);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}
extern "C" int smb_mc_emit_gather_flags(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                                        int emit_last_plane, int flags, float vdiv, float vmul, float vadd, const void* workspace,
                                        void* ctrl_local, int64_t seq, int rank, float* verts_dst, int64_t verts_capacity,
                                        void* faces_dst, int64_t faces_capacity, void* stream) {
  if (!ctrl_local || rank < 0 || rank >= 16 || !verts_dst || !faces_dst) return SMB_ERR_BAD_ARG;
  int64_t* c = static_cast<int64_t*>(ctrl_local);
  return launch_emit(grid, nx, ny, nz, sub, sign, x_origin, emit_last_plane, flags, vdiv, vmul, vadd, 0, workspace, verts_dst,
                     verts_capacity, static_cast<int64_t*>(faces_dst), faces_capacity, c + smb::kPeerCounts, rank, stream,
                     c + smb::kPeerCountSeq, seq, c + smb::kPeerError);
}

extern "C" int smb_mc_emit(const float* grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                           int emit_last_plane, int flags, float vdiv, float vmul, float vadd,
                           int64_t vertex_id_offset, const void* workspace, float* verts, int64_t* faces,
                           void* stream) {
  return smb_mc_emit_bounded(grid, nx, ny, nz, sub, sign, x_origin, emit_last_plane, flags, vdiv, vmul, vadd, vertex_id_offset,
                             workspace, verts, INT64_MAX, faces, INT64_MAX, stream);
}

extern "C" int smb_mc_cases(const float* grid, int nx, int ny, int nz, float sub, float sign, unsigned char* cases,
                            void* stream) {
  if (!grid || !cases || nx < 2 || ny < 2 || nz < 2) return SMB_ERR_BAD_ARG;
  mc_cases_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(grid, nx, ny, nz, sub, sign, cases);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

extern "C" int smb_grid_minmax(const float* grid, int64_t n, float sub, float sign, float* minmax_dev, void* stream) {
  if (!grid || !minmax_dev || n <= 0) return SMB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  mc_minmax_init<<<1, 1, 0, st>>>(minmax_dev);
  mc_minmax<<<592, 256, 0, st>>>(grid, (long long)n, sub, sign, minmax_dev);
  return cudaGetLastError() == cudaSuccess ? SMB_OK : SMB_ERR_CUDA;
}

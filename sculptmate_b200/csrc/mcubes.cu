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
//                  a 32-byte record per word, the chunk totals and the weight of every
//                  256-word unit.
//   K3 mc_totals : one CTA turns chunk totals into chunk bases in canonical order and builds
//                  mc_emit's work list (non-empty units, heavy ones cut into slices,
//                  heaviest first).
//   K4 mc_emit   : a CTA draws work items from a ticket counter; per item it stages what the
//                  unit's 256 words and their neighbours will be asked for in shared memory
//                  (one round of independent loads), scans the per-word item counts, and its
//                  eight warps work through the unit's crossing samples and active cells in
//                  groups of 32, ONE ITEM PER LANE: a crossing sample writes its <= 3 owned
//                  vertices (each lattice edge is owned by exactly one sample -> no duplicates,
//                  no atomics on the output), an active cell writes its <= 5 triangles, vertex
//                  ids from the staged prefixes + popc(mask & lanes_below).  The surface is
//                  sparse (a few % of the cells), so lane-per-item keeps the SIMT lanes full
//                  where lane-per-sample left > 80 % of them idle.  Density is re-read only at
//                  the two end points of crossing edges.
//
// Canonical order (identical to oracle/mc_oracle.c): vertices by x-plane i, inside a
// plane first the in-plane crossings by (j,k) with the y-edge before the z-edge of a
// sample, then the x-edge crossings (plane i -> i+1) by (j,k); triangles by cell
// (i,j,k) then table order.  Slabs concatenate bit-exactly (see smb_mc_emit).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/sculptmate_b200.h"
#define SMB_TABLE_QUAL __device__ const
#include "mc_tables.h"

namespace smb {

constexpr int kChunkWords = 1024;  // words per count CTA (256 threads x 4): 512 CTAs at 256^3 (2048 = 256 CTAs left the 148 SMs 1.7 waves)
constexpr int kWordsPerThread = 4;
constexpr int kUnitWords = 256;   // mc_emit's unit: words of one x-plane staged per CTA (= its threads); 4 units per count chunk
constexpr int kMaxSlices = 16;    // a heavy unit is worked on by up to 16 CTAs
constexpr int kSliceWeight = 3072;  // vertices + triangles one work item should hold at most

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
  uint32_t* ticket;  // [0] mc_emit's work counter (zeroed by mc_totals, wraps back to 0 at the end of every emit), [1] number of work items
  uint32_t* uw;      // weight (vertices + triangles) of every unit of 256 words (mc_count)
  uint2* items;      // mc_emit's work list, heaviest first (mc_totals): x = plane, y = (first word / 256) << 8 | slice << 4 | slices - 1
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
  w.ticket = reinterpret_cast<uint32_t*>(take(256));
  w.uw = reinterpret_cast<uint32_t*>(take((size_t)d.nchunks * (kChunkWords / kUnitWords) * 4));
  w.items = reinterpret_cast<uint2*>(take((size_t)d.nchunks * (kChunkWords / kUnitWords) * kMaxSlices * sizeof(uint2)));
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
                                                ChunkRec* __restrict__ ctot, uint32_t* __restrict__ uw) {
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
  if (tid < kChunkWords / kUnitWords) {
    // weight of mc_emit's units (256 words = the words of two warps here): vertices + triangles
    constexpr int kWarpsPerUnit = kUnitWords / (32 * kWordsPerThread);
    unsigned long long v = 0;
#pragma unroll
    for (int q = 0; q < kWarpsPerUnit; ++q) v += s_warp[tid * kWarpsPerUnit + q];
    uw[(long long)blockIdx.x * (kChunkWords / kUnitWords) + tid] =
        (uint32_t)(v & 0x1fffffu) + (uint32_t)((v >> 21) & 0x1fffffu) + (uint32_t)(v >> 42);
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
                                                  McDims d, int emit_last_plane, smb_mc_counts* __restrict__ counts,
                                                  const uint32_t* __restrict__ uw, uint32_t* __restrict__ ticket,
                                                  uint2* __restrict__ items) {
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
  // ---- mc_emit's work list.  A unit = 256 consecutive words of a plane (4 per chunk); its weight (vertices + triangles)
  // comes from mc_count.  Empty units produce no work, heavy ones are cut into up to 16
  // slices (each staged by its own CTA, the groups dealt round-robin), and the items are listed heaviest first in 64
  // linear classes: longest-processing-time-first keeps the tail of mc_emit short when the surface is concentrated in
  // a part of the volume.  The order inside a class depends on the atomics; it only affects scheduling, never the output.
  __shared__ uint32_t s_hist[64], s_cur[64];
  if (tid < 64) s_hist[tid] = 0u;
  __syncthreads();
  constexpr int kUnitsPerChunk = kChunkWords / kUnitWords;
  const long long nunits = d.nchunks * kUnitsPerChunk;
  auto slices_of = [](uint32_t w) -> uint32_t { return w == 0u ? 0u : min((uint32_t)kMaxSlices, (w + kSliceWeight - 1) / kSliceWeight); };
  auto class_of = [](uint32_t w, uint32_t n) -> uint32_t { return 63u - min(63u, (w / n) * 63u / kSliceWeight); };
  constexpr int kPer = 8;  // unit weights per thread and round: independent loads, one latency per 8192 units
  for (long long u0 = 0; u0 < nunits; u0 += 1024 * kPer) {
    uint32_t w[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const long long u = u0 + e * 1024 + tid;
      w[e] = u < nunits ? uw[u] : 0u;
    }
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const uint32_t n = slices_of(w[e]);
      if (n) atomicAdd(&s_hist[class_of(w[e], n)], n);
    }
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0u;
    for (int q = 0; q < 64; ++q) {
      s_cur[q] = run;
      run += s_hist[q];
    }
    ticket[0] = 0u;
    ticket[1] = run;
  }
  __syncthreads();
  for (long long u0 = 0; u0 < nunits; u0 += 1024 * kPer) {
    uint32_t w[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const long long u = u0 + e * 1024 + tid;
      w[e] = u < nunits ? uw[u] : 0u;
    }
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const uint32_t n = slices_of(w[e]);
      if (n) {
        const long long u = u0 + e * 1024 + tid;
        const long long c = u / kUnitsPerChunk;
        const uint32_t plane = (uint32_t)(c / d.cpp);
        const uint32_t unit_in_plane = (uint32_t)(c - (long long)plane * d.cpp) * kUnitsPerChunk + (uint32_t)(u - c * kUnitsPerChunk);
        const uint32_t at = atomicAdd(&s_cur[class_of(w[e], n)], n);
        for (uint32_t k = 0; k < n; ++k) items[at + k] = make_uint2(plane, (unit_in_plane << 8) | (k << 4) | (n - 1u));
      }
    }
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
  // dynamic scheduling: every CTA draws positions of the work list from ticket[0] (atomicInc with wrap = items + gridDim - 1:
  // each CTA ends on exactly one failed draw, so the counter is back at 0 when the grid retires)
  uint32_t* ticket;       // [0] counter, [1] number of work items
  const uint2* items;     // the work list (mc_totals)
};

__device__ __forceinline__ float mc_xform(float v, int flags, float vdiv, float vmul, float vadd) {
  if (flags & SMB_MC_DIV) v = __fdiv_rn(v, vdiv);
  if (flags & SMB_MC_AFFINE) v = __fadd_rn(__fmul_rn(v, vmul), vadd);
  return v;
}

constexpr int kEmitWarps = 8;
constexpr int kHaloMax = 72;     // words staged behind the unit (rows j+1 and the next word of a row: wz + 1 <= 72 up to nz = 2272)
constexpr int kStageWords = kUnitWords + kHaloMax;
constexpr int kXfMax = 1024;  // lattice indices with a tabulated output coordinate

// position of the n-th (0-based) set bit of m; popc(m) > n
__device__ __forceinline__ int nth_set_bit(uint32_t m, uint32_t n) {
  int pos = 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const uint32_t low = m & ((1u << s) - 1u);
    const uint32_t c = __popc(low);
    if (n >= c) {
      n -= c;
      m >>= s;
      pos += s;
    } else {
      m = low;
    }
  }
  return pos;
}

// What a unit needs of the words it touches, staged once per unit: plane 0 = the unit's x-plane i, plane 1 = i + 1.
struct EmitStage {
  uint32_t pos[2][kStageWords];  // sign masks
  uint32_t A[2][kStageWords];    // id of the word's first in-plane vertex (chunk base + chunk-local prefix)
  uint32_t my[2][kStageWords];   // crossing masks of the owned y- / z-edges
  uint32_t mz[2][kStageWords];
  uint32_t B[kStageWords];       // id of the word's first x-edge vertex, mask of the owned x-edges (plane i only)
  uint32_t mx[kStageWords];
  uint32_t T[kUnitWords];        // slot of the word's first triangle, mask of its active cells, its row j
  uint32_t act[kUnitWords];
  uint32_t row[kUnitWords];
  uint32_t offV[kUnitWords];     // exclusive prefix over the unit's words of crossing samples / active cells
  uint32_t offC[kUnitWords];
  uint32_t warp_tot[kEmitWarps];
  uint32_t draw;        // the ticket drawn for the next loop iteration and the work item it maps to
  uint2 item;
  uint32_t gctr;        // groups of the current item handed out so far
};

// largest l in [0, kUnitWords) with off[l] <= idx
__device__ __forceinline__ int locate_word(const uint32_t* __restrict__ off, uint32_t idx) {
  int lo = 0;
#pragma unroll
  for (int step = kUnitWords / 2; step > 0; step >>= 1)
    if (off[lo + step] <= idx) lo += step;
  return lo;
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane) {
  uint32_t inc = v;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, sft);
    if (lane >= sft) inc += y;
  }
  return inc - v;
}

// One CTA works through UNITS of 256 consecutive words of one x-plane, drawn from a ticket counter:
//   stage   everything the unit's items will ask for goes to shared memory in ONE round of independent loads (the records
//           of the unit's words and of the wz + 1 words behind them, in planes i and i + 1), then a CTA-wide scan of the
//           per-word item counts;
//   V / T   the unit's crossing samples / active cells are numbered 0..total and handed out in groups of 32 to the eight
//           warps round-robin, one item per lane: full lanes, no load imbalance inside a unit, and neither phase reads
//           global memory again except for the densities at the two ends of a crossing edge.
// The earlier design (a warp per 32 words, items compacted inside the warp, neighbour records fetched from L2 per cell)
// spent 118 us at 256^3: every group iteration was a chain of dependent L2 round trips, and the batches with 20-40 groups
// ran serially in one warp (a third of the kernel was tail).
//
// kFar: rows are longer than the staged halo (wz + 1 > 72, nz > 2272): the neighbour words that fall behind it are read from
// global memory; the common instantiation has no such path.
//
// kStaged (peer-memory destination): the vertices and triangles of a group occupy CONTIGUOUS output slots, so they are
// assembled in shared memory and written out by the whole warp with coalesced 128-byte stores.
template <bool kStaged, bool kFar>
__global__ void __launch_bounds__(kEmitWarps * 32, 4) mc_emit(EmitParams p) {
  __shared__ EmitStage S;
  // kStaged: per warp 64 in-plane + 32 x-edge vertices and <= 160 triangles (slab-local vertex ids) of a group, in dynamic
  // shared memory (the static part alone is 40 KB)
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float (*s_vst)[96 * 3] = reinterpret_cast<float (*)[96 * 3]>(s_dyn);
  uint32_t (*s_fst)[160 * 3] = reinterpret_cast<uint32_t (*)[160 * 3]>(s_dyn + kEmitWarps * 96 * 3 * sizeof(float));
  __shared__ unsigned short s_tri3[256][5];  // triangle t of a case: its three edges, 4 bits each, winding already applied
  __shared__ unsigned char s_ntri[256];
  __shared__ uint32_t s_eid[12][kEmitWarps * 32];  // per-thread column of the 12 edge vertex ids
  __shared__ float s_xf[kXfMax];                   // output coordinate of lattice index n (y and z axes)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool flip = p.flags & SMB_MC_FLIP;
  {
    const int cs = tid;  // 256 threads = 256 cases
    const int nt = SMB_MC_NTRI[cs];
    s_ntri[cs] = (unsigned char)nt;
    for (int t = 0; t < 5; ++t) {
      uint32_t e0 = 0, e1 = 0, e2 = 0;
      if (t < nt) {
        e0 = (uint32_t)SMB_MC_TRI[cs][3 * t + 0];
        e1 = (uint32_t)SMB_MC_TRI[cs][3 * t + 1];
        e2 = (uint32_t)SMB_MC_TRI[cs][3 * t + 2];
      }
      s_tri3[cs][t] = (unsigned short)(flip ? (e1 | (e0 << 4) | (e2 << 8)) : (e0 | (e1 << 4) | (e2 << 8)));
    }
  }
  const McDims d = p.d;
  const bool use_xf = max(d.ny, d.nz) <= kXfMax;
  if (use_xf)
    for (int n = tid; n < max(d.ny, d.nz); n += kEmitWarps * 32) s_xf[n] = mc_xform((float)n, p.flags, p.vdiv, p.vmul, p.vadd);

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
  const int wz = d.wz;
  const long long sy = d.nz, sx = (long long)d.ny * d.nz;

  // Work items (a unit, or one of the slices of a heavy unit) come from mc_totals' list, heaviest first; the draw for the
  // NEXT item is issued at the head of the current one, so the atomic's and the list load's latencies are covered by work.
  const uint32_t nitems = __ldg(p.ticket + 1);
  const uint32_t wrap = nitems + gridDim.x - 1u;
  uint32_t nu = 0;
  uint2 nc = make_uint2(0u, 0u);
  if (tid == 0) {
    nu = atomicInc(p.ticket, wrap);
    if (nu < nitems) nc = __ldg(p.items + nu);
    S.draw = nu;
    S.item = nc;
  }
  for (;;) {
    __syncthreads();  // S.draw / S.item are set, the previous item's shared data is no longer read, the tables are written
    if (S.draw >= nitems) break;
    const uint2 item = S.item;
    const uint32_t slice = (item.y >> 4) & 15u, nslices = (item.y & 15u) + 1u;
    const int i = (int)item.x;
    const int w0 = (int)(item.y >> 8) * kUnitWords;
    const int nown = min(kUnitWords, d.pw - w0);
    const int nst = min(kStageWords, d.pw - w0);
    const bool hx = i + 1 < d.nx;
    const bool store_inplane = hx || p.emit_last_plane;
    // nothing to store: last plane of a slab that is not the last (its in-plane crossings are numbered, not stored)
    const bool skip = !hx && !store_inplane;
    if (tid == 0) {
      nu = atomicInc(p.ticket, wrap);
      S.gctr = 0u;
    }

    // ---- stage -----------------------------------------------------------------------------------------------------
    if (!skip) {
      for (int t = tid; t < nst; t += kEmitWarps * 32) {
        const int jw = w0 + t;
        const long long wd = (long long)i * d.pw + jw;
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(&p.rec[wd]));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(&p.rec[wd]) + 1);
        const uint4 cb = __ldg(reinterpret_cast<const uint4*>(&p.cbase[(long long)i * d.cpp + jw / kChunkWords]));
        const uint32_t ps = __ldg(p.pos + wd);
        uint32_t a1 = 0, my1 = 0, mz1 = 0, ps1 = 0;
        if (hx) {
          const uint4 hi1 = __ldg(reinterpret_cast<const uint4*>(&p.rec[wd + d.pw]) + 1);
          a1 = __ldg(&p.rec[wd + d.pw].a) + __ldg(&p.cbase[(long long)(i + 1) * d.cpp + jw / kChunkWords].a);
          my1 = hi1.y;
          mz1 = hi1.z;
          ps1 = __ldg(p.pos + wd + d.pw);
        }
        S.pos[0][t] = ps;
        S.pos[1][t] = ps1;
        S.A[0][t] = cb.x + lo.x;
        S.A[1][t] = a1;
        S.my[0][t] = hi.y;
        S.my[1][t] = my1;
        S.mz[0][t] = hi.z;
        S.mz[1][t] = mz1;
        S.B[t] = cb.y + lo.y;
        S.mx[t] = hi.x;
        if (t < kUnitWords) {
          S.T[t] = cb.z + lo.z;
          S.act[t] = hi.w;
          S.row[t] = (uint32_t)(jw / wz);
        }
      }
    }
    if (tid == 0 && nu < nitems) nc = __ldg(p.items + nu);
    __syncthreads();  // also: every thread has read S.draw / S.item
    if (skip) {
      if (tid == 0) {
        S.draw = nu;
        S.item = nc;
      }
      continue;
    }
    // ---- CTA-wide exclusive scan of (crossing samples, active cells) per word, packed 16 | 16 -------------------------
    uint32_t totV, totC;
    {
      uint32_t cnt = 0;
      if (tid < nown) cnt = (uint32_t)__popc(S.mx[tid] | S.my[0][tid] | S.mz[0][tid]) | ((uint32_t)__popc(S.act[tid]) << 16);
      const uint32_t ex = warp_excl_scan(cnt, lane);
      if (lane == 31) S.warp_tot[warp] = ex + cnt;
      __syncthreads();
      uint32_t base = 0, total = 0;
#pragma unroll
      for (int q = 0; q < kEmitWarps; ++q) {
        const uint32_t v = S.warp_tot[q];
        if (q < warp) base += v;
        total += v;
      }
      totV = total & 0xffffu;
      totC = total >> 16;
      const uint32_t e = base + ex;
      S.offV[tid] = tid < nown ? (e & 0xffffu) : totV;  // words behind the unit: never <= a valid item index
      S.offC[tid] = tid < nown ? (e >> 16) : totC;
      __syncthreads();
    }
    const float xi = mc_xform((float)(p.x_origin + i), p.flags, p.vdiv, p.vmul, p.vadd);

    // The item's groups of 32 crossing samples (V) and of 32 active cells (T) -- every nslices-th group of the unit, starting
    // at `slice` -- are handed to the warps by a shared counter, T groups (the longer ones) first.
    const uint32_t ngV = (totV + 31u) / 32u, ngC = (totC + 31u) / 32u;
    const uint32_t nVs = ngV > slice ? (ngV - slice + nslices - 1u) / nslices : 0u;
    const uint32_t nCs = ngC > slice ? (ngC - slice + nslices - 1u) / nslices : 0u;
    for (;;) {
    uint32_t gq = 0;
    if (lane == 0) gq = atomicAdd(&S.gctr, 1u);
    gq = __shfl_sync(0xffffffffu, gq, 0);
    if (gq >= nVs + nCs) break;
    // ---- V: one lane per crossing sample -> its owned vertices -----------------------------------------------------------
    if (gq >= nCs) {
      const uint32_t base = (slice + nslices * (gq - nCs)) * 32u;
      const uint32_t idx = base + lane;
      const bool valid = idx < totV;
      uint32_t sl_in = 0, sl_x = 0, c_in = 0, c_x = 0;
      bool by = false, bz = false, bx = false;
      int j = 0, k = 0;
      long long pt = 0;
      if (valid) {
        const int src = locate_word(S.offV, idx);
        const uint32_t smx = S.mx[src], smy = S.my[0][src], smz = S.mz[0][src];
        const int bit = nth_set_bit(smx | smy | smz, idx - S.offV[src]);
        const uint32_t below = (1u << bit) - 1u;
        by = (smy >> bit) & 1u, bz = (smz >> bit) & 1u, bx = (smx >> bit) & 1u;
        sl_in = S.A[0][src] + __popc(smy & below) + __popc(smz & below);
        sl_x = S.B[src] + __popc(smx & below);
        c_in = (by ? 1u : 0u) + (bz ? 1u : 0u);
        c_x = bx ? 1u : 0u;
        j = (int)S.row[src];
        k = (w0 + src - j * wz) * 32 + bit;
        pt = (long long)i * sx + (long long)j * sy + k;
      }
      // kStaged: the slots of a group are consecutive from lane 0's (lane 0 always holds an item)
      const uint32_t b_in = __shfl_sync(0xffffffffu, sl_in, 0), b_x = __shfl_sync(0xffffffffu, sl_x, 0);
      if (valid) {
        // all densities of the item in one round trip
        const float a = mc_val(p.grid, pt, p.sub, p.sign);
        float vy = 0.f, vz = 0.f, vx = 0.f;
        if (by) vy = mc_val(p.grid, pt + sy, p.sub, p.sign);
        if (bz) vz = mc_val(p.grid, pt + 1, p.sub, p.sign);
        if (bx) vx = mc_val(p.grid, pt + sx, p.sub, p.sign);
        const float fj = (float)j, fk = (float)k;
        const float xj = use_xf ? s_xf[j] : mc_xform(fj, p.flags, p.vdiv, p.vmul, p.vadd);
        const float xk = use_xf ? s_xf[k] : mc_xform(fk, p.flags, p.vdiv, p.vmul, p.vadd);
        float* st = s_vst[kStaged ? warp : 0];
        if (by) {
          const float t = __fdiv_rn(a, __fsub_rn(a, vy));
          float* o = kStaged ? st + 3 * (sl_in - b_in) : p.verts + 3 * (v_off + sl_in);
          if (kStaged || (store_inplane && v_off + sl_in < p.vcap)) {
            o[0] = xi;
            o[1] = mc_xform(__fadd_rn(fj, t), p.flags, p.vdiv, p.vmul, p.vadd);
            o[2] = xk;
          }
        }
        if (bz) {
          const float t = __fdiv_rn(a, __fsub_rn(a, vz));
          const uint32_t sl = sl_in + (by ? 1u : 0u);
          float* o = kStaged ? st + 3 * (sl - b_in) : p.verts + 3 * (v_off + sl);
          if (kStaged || (store_inplane && v_off + sl < p.vcap)) {
            o[0] = xi;
            o[1] = xj;
            o[2] = mc_xform(__fadd_rn(fk, t), p.flags, p.vdiv, p.vmul, p.vadd);
          }
        }
        if (bx) {
          const float t = __fdiv_rn(a, __fsub_rn(a, vx));
          float* o = kStaged ? st + 3 * (64 + sl_x - b_x) : p.verts + 3 * (v_off + sl_x);
          if (kStaged || v_off + sl_x < p.vcap) {
            o[0] = mc_xform(__fadd_rn((float)(p.x_origin + i), t), p.flags, p.vdiv, p.vmul, p.vadd);
            o[1] = xj;
            o[2] = xk;
          }
        }
      }
      if (kStaged) {
        const int nlast = (int)min(32u, totV - base) - 1;  // lane of the group's last item
        const uint32_t n_in = __shfl_sync(0xffffffffu, sl_in + c_in, nlast) - b_in;
        const uint32_t n_x = __shfl_sync(0xffffffffu, sl_x + c_x, nlast) - b_x;
        __syncwarp();
        const float* st = s_vst[warp];
        if (store_inplane) {
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
      continue;
    }

    // ---- T: one lane per active cell -> its triangles -------------------------------------------------------------------
    // sign mask / record fields of word t2 (relative to the unit) of plane i + di: staged, or (nz > 2272) from global memory
    auto pos_at = [&](int di, int t2) -> uint32_t {
      return (!kFar || t2 < kStageWords) ? S.pos[di][t2] : __ldg(p.pos + (long long)(i + di) * d.pw + w0 + t2);
    };
    // the cell's 8 corner signs: row r = (di, dj) contributes the bits (k, k + 1) of its word, taken across the word border
    auto cell_case_at = [&](int src, int bit) -> uint32_t {
      uint32_t cs = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int t2 = src + (r & 1) * wz;
        const uint32_t two = __funnelshift_r(pos_at(r >> 1, t2), pos_at(r >> 1, t2 + 1), bit) & 3u;
        cs |= two << (2 * r);
      }
      return cs;
    };
    struct RowInfo {
      uint32_t A, my, mz, B, mx;
    };
    auto row_info = [&](int di, int t2) -> RowInfo {
      RowInfo r;
      if (!kFar || t2 < kStageWords) {
        r.A = S.A[di][t2];
        r.my = S.my[di][t2];
        r.mz = S.mz[di][t2];
        r.B = di == 0 ? S.B[t2] : 0u;
        r.mx = di == 0 ? S.mx[t2] : 0u;
      } else {
        const int jw2 = w0 + t2;
        const long long w2 = (long long)(i + di) * d.pw + jw2;
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(&p.rec[w2]));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(&p.rec[w2]) + 1);
        const uint2 cb = __ldg(reinterpret_cast<const uint2*>(&p.cbase[(long long)(i + di) * d.cpp + jw2 / kChunkWords]));
        r.A = cb.x + lo.x;
        r.B = cb.y + lo.y;
        r.mx = hi.x;
        r.my = hi.y;
        r.mz = hi.z;
      }
      return r;
    };
    {
      const uint32_t base = (slice + nslices * gq) * 32u;
      const uint32_t idx = base + lane;
      const bool valid = idx < totC;
      int src = 0, bit = 0;
      uint32_t cs = 0, ntri = 0;
      if (valid) {
        src = locate_word(S.offC, idx);
        bit = nth_set_bit(S.act[src], idx - S.offC[src]);
        cs = cell_case_at(src, bit);
        ntri = s_ntri[cs];
      }
      // Triangles are numbered in word order, then cell order: the group's run starts at the first triangle of lane 0's
      // cell = the word's first slot + the triangles of the word's lower cells (one cell per lane, summed over the warp),
      // and continues over the lanes.
      const int src0 = __shfl_sync(0xffffffffu, src, 0), bit0 = __shfl_sync(0xffffffffu, bit, 0);
      uint32_t pre = 0;
      if (bit0 != 0) {  // warp-uniform
        if (lane < bit0 && ((S.act[src0] >> lane) & 1u)) pre = s_ntri[cell_case_at(src0, lane)];
        pre = __reduce_add_sync(0xffffffffu, pre);
      }
      const uint32_t rel_base = S.T[src0] + pre;
      const uint32_t rel_slot = rel_base + warp_excl_scan(ntri, lane);
      if (ntri != 0u) {
        // ids of the vertices on the cell's 12 edges, by the sample that owns them: owner (di,dj,dk) holds x-edge 2dj+dk
        // (di=0), y-edge 4+2di+dk (dj=0), z-edge 8+2di+dj (dk=0).  The owners dk = 0, 1 of a row (di,dj) are bits k, k+1 of
        // one word: the second one's ids follow from the first one's by the crossing bits at k.  All 12 are computed
        // (branch-free); the case's triangles only read the ones that exist.
#pragma unroll
        for (int di = 0; di < 2; ++di) {
#pragma unroll
          for (int dj = 0; dj < 2; ++dj) {
            const int t2 = src + dj * wz;
            const RowInfo r = row_info(di, t2);
            const uint32_t below = (1u << bit) - 1u;
            const uint32_t myb = (r.my >> bit) & 1u, mzb = (r.mz >> bit) & 1u;
            const uint32_t idyz0 = r.A + __popc(r.my & below) + __popc(r.mz & below);
            uint32_t idyz1 = idyz0 + myb + mzb;
            uint32_t idx0 = 0, idx1 = 0;
            if (di == 0) {
              idx0 = r.B + __popc(r.mx & below);
              idx1 = idx0 + ((r.mx >> bit) & 1u);
            }
            if (bit == 31) {  // the owner dk = 1 is sample 0 of the next word of the row: nothing of that word lies below it
              const RowInfo r2 = row_info(di, t2 + 1);
              idyz1 = r2.A;
              idx1 = r2.B;
            }
            if (di == 0) {
              s_eid[2 * dj + 0][tid] = idx0;
              s_eid[2 * dj + 1][tid] = idx1;
            }
            if (dj == 0) {
              s_eid[4 + 2 * di + 0][tid] = idyz0;
              s_eid[4 + 2 * di + 1][tid] = idyz1;
            }
            s_eid[8 + 2 * di + dj][tid] = idyz0 + myb;
          }
        }
        const unsigned short* tri = s_tri3[cs];
        if (kStaged) {
          // slab-local ids of the triangles into the group's contiguous run in shared memory
          uint32_t* o = s_fst[warp] + 3 * (rel_slot - rel_base);
          for (uint32_t t = 0; t < ntri; ++t) {
            const uint32_t e = tri[t];
            o[3 * t + 0] = s_eid[e & 15u][tid];
            o[3 * t + 1] = s_eid[(e >> 4) & 15u][tid];
            o[3 * t + 2] = s_eid[e >> 8][tid];
          }
        } else {
          const long long slot0 = f_off + rel_slot;
          const uint32_t nt = (uint32_t)max(0LL, min((long long)ntri, p.fcap - slot0));
          if (p.flags & SMB_MC_FACES_I32) {
            // int32 indices (what Blender's loop arrays and a PCIe / NVLink wire want); ids < 2^31 is checked on the host
            int* o = static_cast<int*>(p.faces) + 3 * slot0;
            const int ido = (int)id_off;
#pragma unroll
            for (uint32_t t = 0; t < 5u; ++t) {
              if (t >= nt) break;
              const uint32_t e = tri[t];
              o[3 * t + 0] = (int)s_eid[e & 15u][tid] + ido;
              o[3 * t + 1] = (int)s_eid[(e >> 4) & 15u][tid] + ido;
              o[3 * t + 2] = (int)s_eid[e >> 8][tid] + ido;
            }
          } else {
            long long* o = static_cast<long long*>(p.faces) + 3 * slot0;
#pragma unroll
            for (uint32_t t = 0; t < 5u; ++t) {
              if (t >= nt) break;
              const uint32_t e = tri[t];
              o[3 * t + 0] = (long long)s_eid[e & 15u][tid] + id_off;
              o[3 * t + 1] = (long long)s_eid[(e >> 4) & 15u][tid] + id_off;
              o[3 * t + 2] = (long long)s_eid[e >> 8][tid] + id_off;
            }
          }
        }
      }
      if (kStaged) {
        // the whole warp writes the group's run: 128 contiguous bytes (int32) / 256 (int64) per store instruction
        const int nlast = (int)min(32u, totC - base) - 1;
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
    }  // groups
    if (tid == 0) {
      S.draw = nu;
      S.item = nc;
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
  mc_count<<<(unsigned)d.nchunks, 256, 0, st>>>(w.pos, d, w.rec, w.ctot, w.uw);
  mc_totals<<<1, 1024, 0, st>>>(w.ctot, w.cbase, d, emit_last_plane, counts_dev, w.uw, w.ticket, w.items);
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
  const long long nunits = d.nchunks * (kChunkWords / kUnitWords);
  if ((d.pw + kUnitWords - 1) / kUnitWords >= (1 << 24)) return SMB_ERR_BAD_ARG;  // a work item carries its unit-in-plane in 24 bits
  long long blocks = nunits * kMaxSlices;
  const long long cap = (long long)sm_count() * 4;  // the resident set: work is drawn from mc_totals' list, not assigned
  if (blocks > cap) blocks = cap;
  p.ticket = w.ticket;
  p.items = w.items;
  // staged, coalesced output only on request (SMB_MC_COALESCE: destination is peer memory): on local HBM the extra
  // shared-memory pass costs more than the scattered 4-byte stores it replaces
  constexpr size_t kStagedDyn = kEmitWarps * (96 * 3 * sizeof(float) + 160 * 3 * sizeof(uint32_t));
  static const cudaError_t staged_attr = [] {
    cudaError_t e = cudaFuncSetAttribute(mc_emit<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStagedDyn);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mc_emit<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStagedDyn);
    return e;
  }();
  if (staged_attr != cudaSuccess) return SMB_ERR_CUDA;
  const bool staged = flags & SMB_MC_COALESCE, far = d.wz + 1 > kHaloMax;
  const unsigned g = (unsigned)blocks, b = kEmitWarps * 32;
  cudaStream_t st = (cudaStream_t)stream;
  if (staged) {
    if (far) mc_emit<true, true><<<g, b, kStagedDyn, st>>>(p);
    else mc_emit<true, false><<<g, b, kStagedDyn, st>>>(p);
  } else {
    if (far) mc_emit<false, true><<<g, b, 0, st>>>(p);
    else mc_emit<false, false><<<g, b, 0, st>>>(p);
  }
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

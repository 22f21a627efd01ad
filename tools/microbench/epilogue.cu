// Microbenchmark: ceiling of the K1 epilogue in isolation (no MMA, no barriers): per 16-column chunk
// tcgen05.ld.x16 (prefetched one ahead) -> bias FADD (LDS.128 x4) -> tanh (MUFU) + FFMA -> F2FP pack ->
// tcgen05.st.x8, four chunks per "step" and a tcgen05.wait::st per step, with W warps per SM.
// Reports SFU lane-ops per clock per SM (16 = the MUFU peak) so that the achievable XU utilisation of
// the epilogue instruction stream itself is known independently of the tile hand-off protocol.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sculptmate_b200/csrc -o epilogue epilogue.cu
// Template switches: LD (tcgen05.ld), ST (tcgen05.st), BIAS (LDS + FADD), POLY (k of every 16
// activations through a packed-half polynomial on the FMA pipe instead of the SFU).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"
using namespace smb;

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// silu(2h) for a pair, all in packed half: tanh(|h|) ~ hc*q(hc^2), hc = min(|h|, 4)
__device__ __forceinline__ uint32_t silu_poly_h2(float h0, float h1) {
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 a = __habs2(h);
  const __half2 hc = __hmin2(a, __float2half2_rn(4.0f));
  const __half2 u = __hmul2(hc, hc);
  __half2 q = __float2half2_rn(-2.9e-4f);
  q = __hfma2(q, u, __float2half2_rn(3.15e-3f));
  q = __hfma2(q, u, __float2half2_rn(-2.13e-2f));
  q = __hfma2(q, u, __float2half2_rn(9.57e-2f));
  q = __hfma2(q, u, __float2half2_rn(-3.115e-1f));
  q = __hfma2(q, u, __float2half2_rn(9.96e-1f));
  const __half2 t = __hmul2(hc, q);
  const __half2 r = __hfma2(a, t, h);
  return *reinterpret_cast<const uint32_t*>(&r);
}

template <int LD, int ST, int BIAS, int POLY>
__global__ void __launch_bounds__(640, 1) k(float* out, int steps) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) float sbias[64];
  if (threadIdx.x < 64) sbias[threadIdx.x] = 0.001f * threadIdx.x;
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  // warp w: lane quadrant w%4, tile slot w/4 (96 columns each: 64 accumulator + 32 activation)
  const uint32_t d_t = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 96);
  const uint32_t a_t = d_t + 64;
  {  // defined accumulator contents
    uint32_t z[8];
    for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(0.01f * (float)((threadIdx.x + i) & 63) - 0.3f);
    for (int c = 0; c < 8; ++c) st8(d_t + 8 * c, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  float acc = 0.f;
  const float4* bl = reinterpret_cast<const float4*>(sbias);
  for (int s = 0; s < steps; ++s) {
    uint32_t r[2][16];
    float4 bb[2][4];
    if (LD) tmem_ld16(d_t, r[0]);
    else
      for (int i = 0; i < 16; ++i) r[0][i] = __float_as_uint(acc + 0.01f * i);
    if (BIAS)
      for (int i = 0; i < 4; ++i) bb[0][i] = bl[i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (LD) tmem_ld_wait();
      if (c + 1 < 4) {
        if (LD) tmem_ld16(d_t + (c + 1) * 16, r[(c + 1) & 1]);
        else
          for (int i = 0; i < 16; ++i) r[(c + 1) & 1][i] = __float_as_uint(acc + 0.02f * i + c);
        if (BIAS)
          for (int i = 0; i < 4; ++i) bb[(c + 1) & 1][i] = bl[(c + 1) * 4 + i];
      }
      const uint32_t* rc = r[c & 1];
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = __uint_as_float(rc[i]);
      if (BIAS) {
        const float4* bc = bb[c & 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[4 * i + 0] += bc[i].x;
          h[4 * i + 1] += bc[i].y;
          h[4 * i + 2] += bc[i].z;
          h[4 * i + 3] += bc[i].w;
        }
      }
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (((i * POLY) & 7) < POLY) {  // POLY of every 8 pairs on the FMA pipe
          pk[i] = silu_poly_h2(h[2 * i], h[2 * i + 1]);
        } else {
          pk[i] = pack_half2(silu_from_half_arg(h[2 * i]), silu_from_half_arg(h[2 * i + 1]));
        }
      }
      if (ST) st8(a_t + 8 * c, pk);
      else
        acc += __uint_as_float(pk[0] ^ pk[3] ^ pk[5] ^ pk[7]) * 1e-30f;
    }
    if (ST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(base);
}

template <int LD, int ST, int BIAS, int POLY>
void run(int warps) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int steps = 4000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<LD, ST, BIAS, POLY><<<sms, warps * 32>>>(out, 50);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<LD, ST, BIAS, POLY><<<sms, warps * 32>>>(out, steps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double clk = 1.965e9;
  const double acts = (double)steps * warps * 32 * 64;  // activations per SM
  const double mufu = acts * (8 - POLY) / 8.0;
  printf("ld=%d st=%d bias=%d poly=%d/8 warps=%2d  %.3f ms  clk/step/warp %.0f  activations/clk/SM %.2f  MUFU/clk/SM %.2f  err=%s\n", LD, ST, BIAS,
         POLY, warps, ms, ms * 1e-3 * clk / steps, acts / (ms * 1e-3) / clk, mufu / (ms * 1e-3) / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}


// ---- second experiment: the epilogue shape of field_tc_pair.cu (two 16-column chunks per half, loaded together)
// with the fast-path cost of its synchronisation: SYNC mbarrier try_waits per step on an ALREADY COMPLETED phase
// (each warp arrives on its own barrier and then waits for that phase) and the same number of extra arrivals.
template <int MODE, int SYNC>
__global__ void __launch_bounds__(640, 1) k2(float* out, int steps) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) float sbias[64];
  __shared__ __align__(8) uint64_t sbar[20][2];
  if (threadIdx.x < 64) sbias[threadIdx.x] = 0.001f * threadIdx.x;
  if (threadIdx.x < 40) mbar_init(smem_u32(&sbar[threadIdx.x >> 1][threadIdx.x & 1]), 1);
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t d_t = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 96);
  const uint32_t a_t = d_t + 64;
  const uint32_t bar0 = smem_u32(&sbar[warp][0]), bar1 = smem_u32(&sbar[warp][1]);
  {
    uint32_t z[8];
    for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(0.01f * (float)((threadIdx.x + i) & 63) - 0.3f);
    for (int c = 0; c < 8; ++c) st8(d_t + 8 * c, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  float acc = 0.f;
  uint32_t par = 0;
  const float4* bl = reinterpret_cast<const float4*>(sbias);
  auto sync_point = [&](uint32_t bar, uint32_t bit) {
    if (SYNC) {
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      uint32_t spins = 0;
      while (!mbar_try_wait_hint(bar, (par >> bit) & 1u, 2000u)) {
        if (++spins > (1u << 20)) __trap();
      }
      par ^= 1u << bit;
      tc_fence_after();
    }
  };
  auto compute = [&](const uint32_t (&r)[16], int c, uint32_t (&pk)[8]) {
    float h[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 bv = bl[4 * c + i];
      h[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bv.x;
      h[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv.y;
      h[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv.z;
      h[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_half2(silu_from_half_arg(h[2 * i]), silu_from_half_arg(h[2 * i + 1]));
  };
  for (int s = 0; s < steps; ++s) {
    uint32_t r[2][16], pk[2][8];
    if (MODE == 1) {  // field_tc_pair.cu, first version: both chunks of a half loaded together, waited for at once
      sync_point(bar0, 0);
      tmem_ld16(d_t, r[0]);
      tmem_ld16(d_t + 16, r[1]);
      tmem_ld_wait();
      compute(r[0], 0, pk[0]);
      compute(r[1], 1, pk[1]);
      sync_point(bar1, 1);
      tmem_ld16(d_t + 32, r[0]);
      tmem_ld16(d_t + 48, r[1]);
      st8(a_t, pk[0]);
      st8(a_t + 8, pk[1]);
      tmem_ld_wait();
      compute(r[0], 2, pk[0]);
      st8(a_t + 16, pk[0]);
      compute(r[1], 3, pk[1]);
      st8(a_t + 24, pk[1]);
    } else {  // MODE 2: loads one chunk ahead; the hi half is requested after chunk 0
      sync_point(bar0, 0);
      tmem_ld16(d_t, r[0]);
      tmem_ld16(d_t + 16, r[1]);
      tmem_ld_wait();
      compute(r[0], 0, pk[0]);
      sync_point(bar1, 1);
      tmem_ld16(d_t + 32, r[0]);
      st8(a_t, pk[0]);
      compute(r[1], 1, pk[1]);
      st8(a_t + 8, pk[1]);
      tmem_ld_wait();
      tmem_ld16(d_t + 48, r[1]);
      compute(r[0], 2, pk[0]);
      st8(a_t + 16, pk[0]);
      tmem_ld_wait();
      compute(r[1], 3, pk[1]);
      st8(a_t + 24, pk[1]);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (SYNC) {  // the a_ready-style arrival
      tc_fence_before();
      __syncwarp();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(par);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(base);
}

template <int MODE, int SYNC>
void run2(int warps) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int steps = 4000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k2<MODE, SYNC><<<sms, warps * 32>>>(out, 50);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k2<MODE, SYNC><<<sms, warps * 32>>>(out, steps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double clk = 1.965e9;
  const double acts = (double)steps * warps * 32 * 64;
  printf("pair-shape mode=%d sync=%d warps=%2d  %.3f ms  clk/step/warp %.0f  MUFU/clk/SM %.2f  err=%s\n", MODE, SYNC, warps, ms, ms * 1e-3 * clk / steps,
         acts / (ms * 1e-3) / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

// ---- third experiment: does the ORDER of the epilogue's instructions matter?  k3<0>: as the compiler schedules the plain
// source (all 16 tanh of a chunk back to back, then the FFMAs and packs: the warp sits in mio_throttle during the burst and
// does its other arithmetic afterwards); k3<1>: software-pipelined by half a chunk -- the tanh of 8 columns are issued, then
// the FFMA / pack / store work of the PREVIOUS 8 columns, fully unrolled -- so that a warp's FMA-pipe work overlaps its own
// queued SFU work.  Same loads (x16, one chunk ahead), same bias LDS, same stores.
template <int PIPE>
__global__ void __launch_bounds__(640, 1) k3(float* out, int steps) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) float sbias[64];
  if (threadIdx.x < 64) sbias[threadIdx.x] = 0.001f * threadIdx.x;
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t d_t = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 96);
  const uint32_t a_t = d_t + 64;
  {
    uint32_t z[8];
    for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(0.01f * (float)((threadIdx.x + i) & 63) - 0.3f);
    for (int c = 0; c < 8; ++c) st8(d_t + 8 * c, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const float4* bl = reinterpret_cast<const float4*>(sbias);
  float keep = 0.f;
  for (int s = 0; s < steps; ++s) {
    uint32_t r[2][16];
    tmem_ld16(d_t, r[0]);
    float hp[8], tp[8];  // previous half chunk: pre-activation and its tanh
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait();
      if (c + 1 < 4) tmem_ld16(d_t + (c + 1) * 16, r[(c + 1) & 1]);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float h[8], t[8];
        const float4 b0 = bl[4 * c + 2 * hh], b1 = bl[4 * c + 2 * hh + 1];
        const uint32_t* rc = r[c & 1] + 8 * hh;
        h[0] = __uint_as_float(rc[0]) + b0.x; h[1] = __uint_as_float(rc[1]) + b0.y; h[2] = __uint_as_float(rc[2]) + b0.z; h[3] = __uint_as_float(rc[3]) + b0.w;
        h[4] = __uint_as_float(rc[4]) + b1.x; h[5] = __uint_as_float(rc[5]) + b1.y; h[6] = __uint_as_float(rc[6]) + b1.z; h[7] = __uint_as_float(rc[7]) + b1.w;
        if (PIPE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t[i]) : "f"(h[i]));
          if (c > 0 || hh > 0) {  // finish the previous half chunk while this one's tanh are in the SFU queue
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[i] = pack_half2(fmaf(hp[2 * i], tp[2 * i], hp[2 * i]), fmaf(hp[2 * i + 1], tp[2 * i + 1], hp[2 * i + 1]));
            const int q = 2 * c + hh - 1;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_t + 4 * q), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { hp[i] = h[i]; tp[i] = t[i]; }
        } else {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[i] = pack_half2(silu_from_half_arg(h[2 * i]), silu_from_half_arg(h[2 * i + 1]));
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_t + 4 * (2 * c + hh)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      }
    }
    if (PIPE) {
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pk[i] = pack_half2(fmaf(hp[2 * i], tp[2 * i], hp[2 * i]), fmaf(hp[2 * i + 1], tp[2 * i + 1], hp[2 * i + 1]));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_t + 28), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    keep += hp[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(base);
}

template <int PIPE>
void run3(int warps) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int steps = 4000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k3<PIPE><<<sms, warps * 32>>>(out, 50);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k3<PIPE><<<sms, warps * 32>>>(out, steps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double clk = 1.965e9;
  const double acts = (double)steps * warps * 32 * 64;
  printf("order pipe=%d warps=%2d  %.3f ms  clk/step/warp %.0f  MUFU/clk/SM %.2f  err=%s\n", PIPE, warps, ms, ms * 1e-3 * clk / steps, acts / (ms * 1e-3) / clk,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'o') {
    for (int w : {4, 8, 12, 16, 20}) run3<0>(w);
    for (int w : {4, 8, 12, 16, 20}) run3<1>(w);
    return 0;
  }
  if (argc > 1) {
    for (int w : {8, 16}) run<1, 1, 1, 0>(w);
    for (int w : {4, 8, 12, 16}) run2<1, 0>(w);
    for (int w : {4, 8, 12, 16}) run2<1, 1>(w);
    for (int w : {4, 8, 12, 16}) run2<2, 0>(w);
    for (int w : {4, 8, 12, 16}) run2<2, 1>(w);
    return 0;
  }
  const int ws[] = {4, 8, 12, 16, 20};
  for (int w : ws) run<0, 0, 0, 0>(w);
  for (int w : ws) run<1, 0, 0, 0>(w);
  for (int w : ws) run<1, 1, 0, 0>(w);
  for (int w : ws) run<1, 1, 1, 0>(w);
  for (int w : ws) run<1, 1, 1, 1>(w);
  for (int w : ws) run<1, 1, 1, 2>(w);
  for (int w : ws) run<1, 1, 1, 3>(w);
  for (int w : ws) run<1, 1, 1, 4>(w);
  for (int w : ws) run<0, 0, 0, 8>(w);
  return 0;
}

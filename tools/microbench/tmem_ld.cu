// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM vs number of warps and
// shape (.32x32b.x16 / .x32), alone and overlapped with MUFU work.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sculptmate_b200/csrc -o tmem_ld tmem_ld.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace smb;

template <int X, int MUFU>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  float acc = 0.f;
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = 0.1f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
    if (X == 32) {
      uint32_t r[32];
      tmem_ld32(taddr + (it & 1) * 32, r);
      tmem_ld_wait();
      acc += __uint_as_float(r[0] ^ r[31]);
    } else {
      uint32_t r[16];
      tmem_ld16(taddr + (it & 3) * 16, r);
      tmem_ld_wait();
      acc += __uint_as_float(r[0] ^ r[15]);
    }
    if (MUFU) {
#pragma unroll
      for (int rep = 0; rep < (X == 32 ? 4 : 2); ++rep)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float y;
          asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
          x[i] = y * 0.5f + 0.25f;
        }
    }
  }
  for (int i = 0; i < 8; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(base);
}

template <int X, int MUFU>
void run(int warps) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<X, MUFU><<<sms, warps * 32>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<X, MUFU><<<sms, warps * 32>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bytes = (double)iters * warps * 32 * X * 4;  // per SM
  double mufu = MUFU ? (double)iters * warps * 32 * (X == 32 ? 32 : 16) : 0;
  printf("ld.32x32b.x%-2d mufu=%d warps=%2d  %.3f ms  %.1f B/ns/SM (= %.1f B/clk @1965)  mufu %.2f/clk  err=%s\n", X, MUFU, warps, ms,
         bytes / (ms * 1e6), bytes / (ms * 1e-3) / 1.965e9, mufu / (ms * 1e-3) / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 12, 16}) run<32, 0>(w);
  for (int w : {4, 8, 12, 16}) run<16, 0>(w);
  for (int w : {4, 8, 12, 16}) run<32, 1>(w);
  for (int w : {4, 8, 12, 16}) run<16, 1>(w);
  return 0;
}

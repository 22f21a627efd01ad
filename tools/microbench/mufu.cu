// Microbenchmark: SFU (MUFU) throughput per SM for tanh.approx / ex2.approx / rcp.approx,
// alone and mixed with the FADD+FFMA+F2FP work of the SiLU epilogue.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP, int MIX>
__global__ void k(float* out, int iters, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + 0.01f * (threadIdx.x + i);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      if (MIX) {  // h = a + b ; s = h*t + h ; pack pairs
        float h = x[i] + seed;
        y = fmaf(h, y, h);
      }
      x[i] = y * 0.5f + 0.25f;
    }
    if (MIX) {
      __half2 p0 = __floats2half2_rn(x[0], x[1]);
      __half2 p1 = __floats2half2_rn(x[2], x[3]);
      acc += __low2float(p0) + __high2float(p1);
    }
  }
  float s = acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int MIX>
void run(const char* name, int warps_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 2048);
  const int iters = 20000;
  int threads = warps_per_sm * 32;
  int blocks = sms;
  if (threads > 1024) { blocks = sms * (threads / 1024); threads = 1024; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP, MIX><<<blocks, threads>>>(out, 100, 0.3f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<OP, MIX><<<blocks, threads>>>(out, iters, 0.3f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = (double)iters * 8 * warps_per_sm * 32;  // per SM
  printf("%-14s mix=%d warps/SM=%2d  %.3f ms  %.2f MUFU lane-ops/ns/SM  (= %.2f per clk at %.0f MHz nominal)\n", name, MIX, warps_per_sm, ms,
         ops / (ms * 1e6), ops / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1e3);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 12, 16, 32}) run<0, 0>("tanh.approx", w);
  for (int w : {4, 16, 32}) run<1, 0>("ex2.approx", w);
  for (int w : {4, 16, 32}) run<2, 0>("rcp.approx", w);
  for (int w : {4, 8, 12, 16, 32}) run<0, 1>("tanh+silu mix", w);
  return 0;
}

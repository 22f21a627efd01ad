// Microbenchmark: SFU throughput of the HALF-precision transcendental forms on sm_100a: tanh.approx.f16 / .f16x2 / .bf16x2,
// ex2.approx.f16x2 -- results per clock per SM (the f32 forms deliver 15.9).  If a packed form delivered two results per SFU
// slot, K1's epilogue (whose activations are rounded to fp16 anyway) would halve its SFU time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_half mufu_half.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  unsigned x[8];
  for (int i = 0; i < 8; ++i) {
    __half2 h = __floats2half2_rn(seed + 0.01f * (threadIdx.x + i), seed - 0.02f * i);
    x[i] = *reinterpret_cast<unsigned*>(&h);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned y;
      if (OP == 0) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x[i]));
      if (OP == 1) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x[i]));
      if (OP == 2) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x[i]));
      if (OP == 3) {
        unsigned short a = (unsigned short)x[i], b;
        asm volatile("tanh.approx.f16 %0, %1;" : "=h"(b) : "h"(a));
        y = b | (x[i] & 0xffff0000u);
      }
      x[i] = y ^ 0x00010001u;
    }
  }
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

template <int OP>
void run(const char* name, int results_per_op, int warps) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<OP><<<sms, warps * 32>>>(out, 100, 0.3f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<OP><<<sms, warps * 32>>>(out, iters, 0.3f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double res = (double)iters * 8 * warps * 32 * results_per_op;
  printf("%-22s warps/SM=%2d  %.3f ms  %.2f results/clk/SM at 1965 MHz  err=%s\n", name, warps, ms, res / (ms * 1e-3) / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) run<0>("tanh.approx.f16x2", 2, w);
  for (int w : {8, 16, 32}) run<1>("tanh.approx.bf16x2", 2, w);
  for (int w : {8, 16, 32}) run<2>("ex2.approx.f16x2", 2, w);
  for (int w : {8, 16, 32}) run<3>("tanh.approx.f16", 1, w);
  return 0;
}

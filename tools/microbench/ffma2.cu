// Throughput of the packed fp32 instructions of sm_100 (FFMA2 / FADD2: two fp32 lanes per instruction, 64-bit register
// pairs) against scalar FFMA, and of a MUFU.TANH stream with the SiLU arithmetic done either way.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int kMode>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kMode == 0) {  // 2 scalar FFMA per pair
        a[i].x = fmaf(a[i].x, m.x, c.x);
        a[i].y = fmaf(a[i].y, m.y, c.y);
      } else if (kMode == 1) {  // 1 FFMA2 per pair
        a[i] = __ffma2_rn(a[i], m, c);
      } else if (kMode == 2) {  // SiLU on a pair, scalar: 2 MUFU + 2 FFMA
        float tx, ty;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(a[i].x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(a[i].y));
        a[i].x = fmaf(a[i].x, tx, a[i].x);
        a[i].y = fmaf(a[i].y, ty, a[i].y);
      } else {  // SiLU on a pair, packed: 2 MUFU + 1 FFMA2
        float2 t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(a[i].x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(a[i].y));
        a[i] = __ffma2_rn(a[i], t, a[i]);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kMode>
void run(const char* name, int warps_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096, blocks = sms * warps_per_sm / 8;
  float* out;
  cudaMalloc(&out, (size_t)blocks * 256 * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<kMode><<<blocks, 256>>>(out, 16);
  cudaEventRecord(a);
  k<kMode><<<blocks, 256>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double pairs = (double)blocks * 256 * iters * 8;
  printf("%-34s warps/SM=%2d  %.3f ms  %.1f fp32 lane-results/clk/SM at %d MHz\n", name, warps_per_sm, ms,
         2 * pairs / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("FFMA x2 (scalar)", w);
    run<1>("FFMA2 (packed)", w);
    run<2>("tanh + FFMA (scalar SiLU)", w);
    run<3>("tanh + FFMA2 (packed SiLU)", w);
  }
  return 0;
}

// Microbenchmark: issue rate of the packed fp32x2 instructions of sm_100 (FADD2 / FFMA2) against the
// scalar ones, and a mix with shared-memory traffic.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp2 fp2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}

// MODE 0: 16 scalar FADD per iteration (8 complex adds); 1: 8 FADD2; 2: 16 scalar FFMA; 3: 8 FFMA2
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* clk) {
  const int lane = threadIdx.x & 31;
  float2 v[8], w[8];
  for (int t = 0; t < 8; ++t) { v[t] = make_float2(lane + t, lane - t); w[t] = make_float2(1e-3f * t * (lane + 1), -1e-3f * t * (lane + 2)); }
  const float2 m2 = make_float2(0.999f + 1e-6f * lane, 1.001f - 1e-6f * lane);
  const long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (MODE == 0) { v[t].x += w[t].x; v[t].y += w[t].y; }
      if (MODE == 1) v[t] = add2(v[t], w[t]);
      if (MODE == 2) { v[t].x = fmaf(v[t].x, m2.x, w[t].x); v[t].y = fmaf(v[t].y, m2.y, w[t].y); }
      if (MODE == 3) v[t] = fma2(v[t], m2, w[t]);
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int t = 0; t < 8; ++t) s += v[t].x + v[t].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 8);
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, iters, clk);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<148, threads>>>(out, iters, clk);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  long long c;
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  // per SM sub-partition: (threads/128) warps, each issuing `n` instructions per iteration
  const int n = (MODE == 0 || MODE == 2) ? 16 : 8;
  printf("%-10s threads %4d: %.3f ms, %lld clk, %.3f clk per warp-instruction per scheduler, %.2f complex ops/clk/SM\n", name,
         threads, ms, c, (double)c / ((double)iters * n * (threads / 128)), (double)iters * 8 * threads / (double)c);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  for (int th : {128, 512}) {
    if (th == 128) { run<0>("FADD", 128); run<1>("FADD2", 128); run<2>("FFMA", 128); run<3>("FFMA2", 128); }
    else { run<0>("FADD", 512); run<1>("FADD2", 512); run<2>("FFMA", 512); run<3>("FFMA2", 512); }
  }
  return 0;
}

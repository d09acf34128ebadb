// Microbenchmark: do warp shuffles share the L1TEX/shared-memory data pipe with LDS/STS?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  __shared__ float2 buf[8][512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* b = buf[warp];
  float2 v[8];
  for (int t = 0; t < 8; ++t) v[t] = make_float2(lane + t, lane - t);
  for (int t = 0; t < 8; ++t) b[lane + 32 * t] = v[t];
  __syncwarp();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {  // 8 STS.64 + 8 LDS.64 (conflict-free): 32 wavefronts
#pragma unroll
      for (int t = 0; t < 8; ++t) b[lane + 32 * t] = v[t];
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = b[((lane + 1) & 31) + 32 * t];
      __syncwarp();
    }
    if (MODE == 1 || MODE == 2) {  // 16 SHFL.32
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        v[t].x = __shfl_xor_sync(0xffffffffu, v[t].x, 1 + (t & 3));
        v[t].y = __shfl_xor_sync(0xffffffffu, v[t].y, 2 + (t & 3));
      }
    }
    if (MODE == 3) {  // FMA only, for reference
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        v[t].x = fmaf(v[t].x, 1.0001f, v[t].y);
        v[t].y = fmaf(v[t].y, 0.9999f, v[t].x);
      }
    }
    if (MODE == 4) {  // match_any
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        unsigned m = __match_any_sync(0xffffffffu, __float_as_int(v[t].x));
        v[t].x += (float)__popc(m);
      }
    }
  }
  float s = 0;
  for (int t = 0; t < 8; ++t) s += v[t].x + v[t].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 256>>>(d, 10);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 256>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // per SM: 16 warps x iters iterations
  double clk = ms * 1e-3 * 1.965e9 / (16.0 * iters);
  printf("%-28s %8.3f ms  -> %6.1f clk per warp-iteration per SM (at 1965 MHz)\n", name, ms, clk);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 2 * 256 * 4);
  const int iters = 20000;
  run<0>("8 STS.64 + 8 LDS.64", d, iters);
  run<1>("16 SHFL", d, iters);
  run<2>("both", d, iters);
  run<3>("16 FFMA (dependent)", d, iters);
  run<4>("8 MATCH.ANY", d, iters);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

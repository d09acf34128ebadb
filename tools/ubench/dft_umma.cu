// dft_umma.cu -- closes the tensor-core question of DESIGN.md 3.1 with a measurement:
// what would the n_fft = 512 frame transform cost on tcgen05.mma kind::tf32?
//
// Two-stage DFT-as-GEMM, 512 = 32 x 16, complex data as real [re | im] blocks, 3xTF32 split for fp32-level bins:
//   stage 1 (32-point DFTs over n1):  D[(frame, n2), (k1 re|im)] = X[(frame, n2), (n1 re|im)] . F32      M = 8 frames x 16
//             = 128, N = 64, K = 64  -> 8 k-steps of K = 8, x3 passes (hi*hi, lo*hi, hi*lo): 24 x mma M128 N64 K8 / 8 frames
//   stage 2 (16-point DFTs over n2):  D[(frame, k1), (k2 re|im)] = Y'[(frame, k1), (n2 re|im)] . F16     M = 4 frames x 32
//             = 128, N = 32, K = 32  -> 4 k-steps x3 passes: 12 x mma M128 N32 K8 / 4 frames
// Between the stages the 128 x 64 fp32 accumulator goes TMEM -> registers (twiddle, hi/lo split) -> shared memory as the
// next A operand; after stage 2 it goes TMEM -> registers for the split / phase transform.
// The kernel measures, per SM (one CTA per SM, operands resident in shared memory, values irrelevant):
//   (a) the MMA stream alone (one thread issuing, tcgen05.commit + mbarrier wait every `batch` groups);
//   (b) the data movement alone: tcgen05.ld of the two accumulators and the shared-memory stores of the split operand;
// and prints clocks per frame for both.  A fused kernel cannot beat max(a, b) and will not do worse than a + b; the
// register FFT of the shipped kernel costs ~232 clk/frame/SM, the whole shipped kernel ~430 (profiles/README.md).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dft_umma tools/ubench/dft_umma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48), layout_type 0 (no swizzle)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format TF32 (2) [7,10) [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void wait_parity(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}

#define TMEM_LD32(r, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
               "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),   \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),  \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])               \
               : "r"(taddr))

// mode 0: MMA stream; mode 1: data movement; mode 2: both in the same CTA (warp 0 lane 0 issues, all 4 warps move data
// of the previous group -- no data dependence is honoured, this is a pure resource-sharing measurement)
__global__ void __launch_bounds__(128, 1) umma_dft_kernel(unsigned long long* cycles, int groups, int batch, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands: A1 8 k-steps x 4 KB (M128 K8 tf32, no swizzle: [16 row groups][2 k-cores][8 rows][16 B]), B1 8 x 2 KB (N64),
  //           A2 4 x 4 KB, B2 4 x 1 KB (N32); hi and lo copies share the buffers (values do not matter for timing)
  unsigned char* A1 = smem;
  unsigned char* B1 = A1 + 8 * 4096;
  unsigned char* A2 = B1 + 8 * 2048;
  unsigned char* B2 = A2 + 4 * 4096;
  float* stage_out = reinterpret_cast<float*>(B2 + 4 * 1024);  // 2 x 32 KB: hi / lo split of the stage-1 result (A2 of the next group)
  for (int i = tid; i < (8 * 4096 + 8 * 2048 + 4 * 4096 + 4 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 2654435761u) >> 20);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes of the operands -> async proxy (tcgen05.mma reads)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA1 = (uint32_t)__cvta_generic_to_shared(A1), sB1 = (uint32_t)__cvta_generic_to_shared(B1);
  const uint32_t sA2 = (uint32_t)__cvta_generic_to_shared(A2), sB2 = (uint32_t)__cvta_generic_to_shared(B2);
  const uint32_t id1 = idesc_tf32(128, 64), id2 = idesc_tf32(128, 32);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0 || mode == 2) {
    if (tid == 0) {
      uint32_t parity = 0;
      for (int g = 0; g < groups; ++g) {  // one group = 8 frames
        for (int p = 0; p < 3; ++p)
          for (int k = 0; k < 8; ++k)
            mma_tf32(tmem, smem_desc(sA1 + k * 4096, 128, 256), smem_desc(sB1 + k * 2048, 128, 256), id1, (p | k) != 0);
        for (int h = 0; h < 2; ++h)
          for (int p = 0; p < 3; ++p)
            for (int k = 0; k < 4; ++k)
              mma_tf32(tmem + 64 + 32 * h, smem_desc(sA2 + k * 4096, 128, 256), smem_desc(sB2 + k * 1024, 128, 256), id2, (p | k) != 0);
        if ((g + 1) % batch == 0 || g + 1 == groups) {
          commit(&bar);
          wait_parity(&bar, parity);
          parity ^= 1;
        }
      }
    }
  }
  if (mode == 1 || mode == 2) {
    // per group of 8 frames: stage-1 accumulator 128 lanes x 64 columns -> registers -> hi / lo -> shared memory (2 x 32 KB),
    // stage-2 accumulators 128 x 64 -> registers (the epilogue would consume them here)
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < groups; ++g) {
      uint32_t r[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        TMEM_LD32(r, lane_base + 32 * half);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4* dst = reinterpret_cast<float4*>(stage_out + (size_t)tid * 64 + 32 * half);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float a = __uint_as_float(r[4 * q]), b = __uint_as_float(r[4 * q + 1]), c = __uint_as_float(r[4 * q + 2]), d = __uint_as_float(r[4 * q + 3]);
          // hi = tf32-truncated value, lo = remainder (the 3xTF32 split of the next stage's operand)
          const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
          const float ch = __uint_as_float(__float_as_uint(c) & 0xffffe000u), dh = __uint_as_float(__float_as_uint(d) & 0xffffe000u);
          dst[q] = make_float4(ah, bh, ch, dh);
          dst[q + 2048] = make_float4(a - ah, b - bh, c - ch, d - dh);  // + 32 KB
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        TMEM_LD32(r, lane_base + 64 + 32 * half);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 32; ++q) sink ^= r[q];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0) + (sink == 0x12345678u);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main(int argc, char** argv) {
  const int groups = argc > 1 ? atoi(argv[1]) : 4096;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 8 * 4096 + 8 * 2048 + 4 * 4096 + 4 * 1024 + 2 * 32768 + 1024;
  cudaFuncSetAttribute(umma_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned long long* d;
  cudaMalloc(&d, sizeof(unsigned long long) * sms);
  std::vector<unsigned long long> h(sms);
  const char* names[3] = {"mma stream only", "tmem->reg->smem movement only", "both in one CTA (no dependences)"};
  printf("{\"sms\": %d, \"frames_per_group\": 8, \"groups\": %d", sms, groups);
  for (int mode = 0; mode < 3; ++mode) {
    for (int batch : {1, 8}) {
      if (mode == 1 && batch != 1) continue;
      for (int rep = 0; rep < 2; ++rep) {
        umma_dft_kernel<<<sms, 128, smem>>>(d, groups, batch, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf(", \"error\": \"%s (mode %d)\"}\n", cudaGetErrorString(e), mode);
          return 1;
        }
      }
      cudaMemcpy(h.data(), d, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      const double med = (double)h[sms / 2];
      printf(", \"%s, commit every %d group(s)\": {\"clk_per_frame_per_sm\": %.1f, \"clk_per_group\": %.1f}", names[mode], batch,
             med / (8.0 * groups), med / groups);
    }
  }
  printf(", \"note\": \"3xTF32 two-stage 32x16 DFT of one 512-point frame: 24 mma M128N64K8 + 24 mma M128N32K8 per 8 frames\"}\n");
  return 0;
}

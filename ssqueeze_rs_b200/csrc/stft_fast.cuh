// stft_fast.cuh -- register-resident fast path of the fused ssq_stft / stft
// kernel for n_fft = 512 (BASELINE.json configs[1], [3], [4]).
//
// One warp owns one frame; each lane keeps 16 complex points in registers and
// does two radix-8 butterflies per stage (512 = 8*8*8, Stockham ordering):
//
//   stage 1  j in {l, l+32}:  in  z[j+64t]                 (straight from the
//            sample tile, window and derivative window live in registers)
//            out idx 8j+t                                   (exchange 1)
//   stage 2  j in {l, l+32}:  in  idx j+64t, times W_64^{(j&7) t}
//            out idx (j>>3)*64 + (j&7) + 8t                 (exchange 2)
//   stage 3  j in {l, 64-l} (lane 0: {0, 32}): in idx j+64t, times W_512^{j t}
//            out X[j+64t] stays in registers: bins k and 512-k needed by the
//            real/imag split sit in the SAME lane, so the split, the phase
//            transform and the bin computation need no further exchange.
//
// Both exchanges go through a 4 KB per-warp buffer with XOR swizzles chosen so
// that every shared-memory access of the FFT is bank-conflict free and all
// per-access address arithmetic folds into immediates.
// All twiddles (21 complex) and the 32 window taps a lane needs are loaded
// once per CTA lifetime (persistent CTAs).
#pragma once
#include "stft_kernels.cuh"

#define SSQ_FAST_N 512
#define SSQ_FAST_F 32            // frames per tile
#define SSQ_FAST_WARPS 8
#define SSQ_FAST_ACC_STRIDE 289  // bin k lives at k + (k>>3) (max 288); odd stride -> conflict-free read-out
#define SSQ_FAST_MAX_HOP 64

template <bool PK = SSQ_PK_DEFAULT>
__device__ __forceinline__ void fft8_fwd(float2 (&v)[8]) {
  const float S = 0.70710678118654752440f;
  float2 a0 = caddf<PK>(v[0], v[4]), a4 = csubf<PK>(v[0], v[4]);
  float2 a1 = caddf<PK>(v[1], v[5]), a5 = csubf<PK>(v[1], v[5]);
  float2 a2 = caddf<PK>(v[2], v[6]), a6 = csubf<PK>(v[2], v[6]);
  float2 a3 = caddf<PK>(v[3], v[7]), a7 = csubf<PK>(v[3], v[7]);
  // odd branch twiddles W8^1, W8^2, W8^3 (the 1/sqrt2 factors are folded into the last layer)
  float2 p5 = caddf<PK>(a5, cmi(a5));  // a5 * (1 - i)      [* S]
  float2 p6 = cmi(a6);             // a6 * (-i)
  float2 p7 = csubf<PK>(cmi(a7), a7);  // a7 * (-1 - i)     [* S]
  float2 b0 = caddf<PK>(a0, a2), b2 = csubf<PK>(a0, a2);
  float2 b1 = caddf<PK>(a1, a3), b3 = csubf<PK>(a1, a3);
  float2 b4 = caddf<PK>(a4, p6), b6 = csubf<PK>(a4, p6);
  float2 b5 = caddf<PK>(p5, p7), b7 = csubf<PK>(p5, p7);  // both still lack the factor S
  float2 r3 = cmi(b3);                            // -i * b3
  float2 r7 = cmi(b7);                            // -i * b7
  v[0] = caddf<PK>(b0, b1);
  v[4] = csubf<PK>(b0, b1);
  v[2] = caddf<PK>(b2, r3);
  v[6] = csubf<PK>(b2, r3);
  v[1] = fma2<PK>(b5, bc2(S), b4);
  v[5] = fma2<PK>(b5, bc2(-S), b4);
  v[3] = fma2<PK>(r7, bc2(S), b6);
  v[7] = fma2<PK>(r7, bc2(-S), b6);
}

// Same butterfly with the conjugate kernel: v[m] <- sum_t v[t] W_8^{-t m}.
template <bool PK = SSQ_PK_DEFAULT>
__device__ __forceinline__ void fft8_inv(float2 (&v)[8]) {
  const float S = 0.70710678118654752440f;
  float2 a0 = caddf<PK>(v[0], v[4]), a4 = csubf<PK>(v[0], v[4]);
  float2 a1 = caddf<PK>(v[1], v[5]), a5 = csubf<PK>(v[1], v[5]);
  float2 a2 = caddf<PK>(v[2], v[6]), a6 = csubf<PK>(v[2], v[6]);
  float2 a3 = caddf<PK>(v[3], v[7]), a7 = csubf<PK>(v[3], v[7]);
  float2 p5 = caddf<PK>(a5, cpi(a5));  // a5 * (1 + i)      [* S]
  float2 p6 = cpi(a6);             // a6 * (+i)
  float2 p7 = csubf<PK>(cpi(a7), a7);  // a7 * (-1 + i)     [* S]
  float2 b0 = caddf<PK>(a0, a2), b2 = csubf<PK>(a0, a2);
  float2 b1 = caddf<PK>(a1, a3), b3 = csubf<PK>(a1, a3);
  float2 b4 = caddf<PK>(a4, p6), b6 = csubf<PK>(a4, p6);
  float2 b5 = caddf<PK>(p5, p7), b7 = csubf<PK>(p5, p7);
  float2 r3 = cpi(b3);  // +i * b3
  float2 r7 = cpi(b7);  // +i * b7
  v[0] = caddf<PK>(b0, b1);
  v[4] = csubf<PK>(b0, b1);
  v[2] = caddf<PK>(b2, r3);
  v[6] = csubf<PK>(b2, r3);
  v[1] = fma2<PK>(b5, bc2(S), b4);
  v[5] = fma2<PK>(b5, bc2(-S), b4);
  v[3] = fma2<PK>(r7, bc2(S), b6);
  v[7] = fma2<PK>(r7, bc2(-S), b6);
}

// Bin k of a frame's accumulator column lives at k + (k >> 3): one pad slot per 8
// bins, so that lanes owning bins 8 apart (the reassignment phase) and lanes owning
// consecutive bins (the stft store) both hit distinct banks.
__device__ __forceinline__ int acc_phys(int k) { return k + (k >> 3); }

// One bin of the epilogue: zk = Z[k], zn = Z[N-k].  MODE 1 stores Sx; MODE 0 parks
// the item (value to add, destination bin or -1 when gated) in the per-warp staging
// area, indexed by the SOURCE bin k.
template <int MODE>
__device__ __forceinline__ void ssq_epilogue_bin(const StftParams& P, float2* col, float2* sval, int* skey, int k,
                                                 float2 zk, float2 zn) {
  const float c = zk.x + zn.x, d = zk.y - zn.y;  // 2*Sx
  if (MODE == 1) {
    col[acc_phys(k)] = make_float2(0.5f * c, 0.5f * d);
    return;
  }
  const float a = zk.y + zn.y, b = zn.x - zk.x;  // 2*V
  const float den = fmaf(c, c, d * d);
  const float num = fmaf(b, c, -a * d);
  const float q = __fdividef(num, den) * P.cphase;
  const float binf = fabsf((float)k - q);
  const float r = ceilf(binf - 0.5f);
  int kb = (int)fminf(fmaxf(r, 0.f), 256.f);  // fmaxf(NaN, 0) = 0 -> bin 0 like the reference
  if (den < P.gate2) kb = -1;                  // |Sx| < gamma (ssq_stft.rs:23): dropped
  float2 v = make_float2(c * P.tx_scale, d * P.tx_scale);
  if (P.squeezing == SSQ_SQUEEZE_LEBESGUE) v = make_float2(P.leb_val, 0.f);
  sval[acc_phys(k)] = v;
  skey[k] = kb;
}

// Adds the lane's item (bin kb or -1) into the frame column.  All 32 lanes call it
// together.  fp32 shared atomics are CAS loops on sm_100a and cost ~2 cycles per
// lane, so none are used: a byte tag per bin detects the (rare) case of two lanes
// aiming at the same bin in the same step; otherwise every lane does a plain
// read-modify-write.  Lanes own 8 CONSECUTIVE source bins each and walk them in
// ascending order, which is the reference's accumulation order (ssq_stft.rs:277).
__device__ __forceinline__ void ssq_accumulate_step(float2* col, unsigned char* tag, int kb, float2 v, int lane) {
  const bool on = kb >= 0;
  if (on) tag[kb] = (unsigned char)lane;
  __syncwarp();
  const bool mine = !on || tag[on ? kb : 0] == (unsigned char)lane;
  if (__all_sync(0xffffffffu, mine)) {
    if (on) {
      float2* p = col + acc_phys(kb);
      float2 t = *p;
      t.x += v.x;
      t.y += v.y;
      *p = t;
    }
  } else {
    for (int src = 0; src < 32; ++src) {  // collision: serialise the lanes (ascending source order)
      if (lane == src && on) {
        float2* p = col + acc_phys(kb);
        float2 t = *p;
        t.x += v.x;
        t.y += v.y;
        *p = t;
      }
      __syncwarp();
    }
  }
  __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(SSQ_FAST_WARPS * 32, 2) ssq_stft512_kernel(const StftParams P) {
  constexpr int N = SSQ_FAST_N, F = SSQ_FAST_F, AS = SSQ_FAST_ACC_STRIDE;
  extern __shared__ float2 smem[];
  float2* acc = smem;                                              // [F][AS]
  float2* xch = acc + F * AS + (threadIdx.x >> 5) * N;            // per warp [512]
  float* tile = reinterpret_cast<float*>(acc + F * AS + SSQ_FAST_WARPS * N);  // [(F-1)*hop + N]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hop = P.hop;
  const int tile_len = (F - 1) * hop + N;

  // ---- per-lane constants (once per CTA lifetime) -------------------------------
  float wr[16], wi[16];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int n = lane + 32 * h + 64 * t;
      wr[h * 8 + t] = P.win[n];
      wi[h * 8 + t] = P.dwin[n];
    }
  const int j2 = lane ? 64 - lane : 32;  // second stage-3 butterfly
  float2 tw2[7], tw3a[7], tw3b[7];
#pragma unroll
  for (int t = 1; t < 8; ++t) {
    tw2[t - 1] = P.tw[((lane & 7) * t * 8) & (N - 1)];
    tw3a[t - 1] = P.tw[(lane * t) & (N - 1)];
    tw3b[t - 1] = P.tw[(j2 * t) & (N - 1)];
  }
  // exchange-1 write: row j (64 B), chunk q stored at q ^ ((j>>1)&3)
  const int f1 = (lane >> 1) & 3;
  // exchange-1 read: element (j + 64t): row (j>>3)+8t, col j&7; chunk (c>>1)^((j>>4)&3)
  // (row = (j>>3)+8t, so (row>>1)&3 = ((j>>3)>>1)&3 independent of t; j = lane or lane+32)
  const int rd1a = (lane >> 3) * 8 + ((((lane & 7) >> 1) ^ ((lane >> 4) & 3)) << 1) + (lane & 1);
  const int rd1b = ((lane >> 3) + 4) * 8 + ((((lane & 7) >> 1) ^ (((lane >> 4) + 2) & 3)) << 1) + (lane & 1);
  // exchange-2 write: idx = (j>>3)*64 + (j&7) + 8t, stored at idx ^ (8*((j>>3)&1))
  const int g2 = (lane >> 3) & 1;
  const int wr2 = (lane >> 3) * 64 + (lane & 7);
  // exchange-2 read: idx = j + 64t stored at idx ^ (8*(t&1))

  for (int i = threadIdx.x; i < F * AS; i += blockDim.x) acc[i] = make_float2(0.f, 0.f);

  for (int64_t tl = blockIdx.x; tl < P.total_tiles; tl += gridDim.x) {
    const int ch = (int)(tl / P.tiles_per_channel);
    const int64_t f0 = (tl % P.tiles_per_channel) * F;
    const int nf = (int)min((int64_t)F, P.n_frames - f0);
    const float* xc = P.x + (size_t)ch * P.x_stride;
    const int64_t p0 = f0 * hop;
    {
      const int64_t o0 = p0 - P.left;
      const int need = (nf - 1) * hop + N;
      if (o0 >= 0 && o0 + need <= P.n) {
        for (int i = threadIdx.x; i < need; i += blockDim.x) tile[i] = __ldg(xc + o0 + i);
      } else {
        for (int i = threadIdx.x; i < need; i += blockDim.x) tile[i] = stft_sample(xc, P.n, p0 + i, P.left, P.padtype);
      }
    }
    __syncthreads();

    for (int fl = warp; fl < nf; fl += SSQ_FAST_WARPS) {
      const float* fr = tile + fl * hop + lane;
      float2 va[8], vb[8];
      // ---- stage 1 -----------------------------------------------------------------
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float x0 = fr[64 * t], x1 = fr[32 + 64 * t];
        va[t] = make_float2(x0 * wr[t], x0 * wi[t]);
        vb[t] = make_float2(x1 * wr[8 + t], x1 * wi[8 + t]);
      }
      fft8_fwd(va);
      fft8_fwd(vb);
      {
        float4* rowa = reinterpret_cast<float4*>(xch + lane * 8);
        float4* rowb = reinterpret_cast<float4*>(xch + (lane + 32) * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          rowa[q ^ f1] = make_float4(va[2 * q].x, va[2 * q].y, va[2 * q + 1].x, va[2 * q + 1].y);
          rowb[q ^ f1] = make_float4(vb[2 * q].x, vb[2 * q].y, vb[2 * q + 1].x, vb[2 * q + 1].y);
        }
      }
      __syncwarp();
      // ---- stage 2 -----------------------------------------------------------------
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        va[t] = xch[rd1a + 64 * t];
        vb[t] = xch[rd1b + 64 * t];
      }
      __syncwarp();
#pragma unroll
      for (int t = 1; t < 8; ++t) {
        va[t] = cmulf(va[t], tw2[t - 1]);
        vb[t] = cmulf(vb[t], tw2[t - 1]);
      }
      fft8_fwd(va);
      fft8_fwd(vb);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        xch[wr2 + 8 * (t ^ g2)] = va[t];
        xch[wr2 + 256 + 8 * (t ^ g2)] = vb[t];
      }
      __syncwarp();
      // ---- stage 3 -----------------------------------------------------------------
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        va[t] = xch[(lane ^ (8 * (t & 1))) + 64 * t];
        vb[t] = xch[(j2 ^ (8 * (t & 1))) + 64 * t];
      }
      __syncwarp();
#pragma unroll
      for (int t = 1; t < 8; ++t) {
        va[t] = cmulf(va[t], tw3a[t - 1]);
        vb[t] = cmulf(vb[t], tw3b[t - 1]);
      }
      fft8_fwd(va);  // va[m] = X[lane + 64 m]
      fft8_fwd(vb);  // vb[m] = X[j2 + 64 m]
      // ---- split + phase transform (lanes own strided bins here) -----------------------
      float2* col = acc + fl * AS;
      float2* sval = xch;                                        // [289] items by source bin
      int* skey = reinterpret_cast<int*>(xch + 296);             // [257]
      unsigned char* tag = reinterpret_cast<unsigned char*>(xch + 432);  // [257]
      const bool l0 = (lane == 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        // bin lane+64m pairs with 512-(lane+64m) = (64-lane)+64(7-m): vb[7-m]; lane 0: va[(8-m)&7]
        float2 pa = vb[7 - m];
        const float2 alt = va[(8 - m) & 7];
        if (l0) pa = alt;
        ssq_epilogue_bin<MODE>(P, col, sval, skey, lane + 64 * m, va[m], pa);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        // bin j2+64m pairs with 512-(j2+64m) = lane+64(7-m): va[7-m]; lane 0 (j2=32): vb[7-m]
        float2 pb = va[7 - m];
        const float2 alt = vb[7 - m];
        if (l0) pb = alt;
        ssq_epilogue_bin<MODE>(P, col, sval, skey, j2 + 64 * m, vb[m], pb);
      }
      if (l0) ssq_epilogue_bin<MODE>(P, col, sval, skey, 256, va[4], va[4]);
      if (MODE == 0) {
        __syncwarp();
        // ---- reassignment: lane owns source bins 8*lane .. 8*lane+7 (lane 31 also 256) ----
        const int4 k0 = *reinterpret_cast<const int4*>(skey + 8 * lane);
        const int4 k1 = *reinterpret_cast<const int4*>(skey + 8 * lane + 4);
        float2 it[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) it[j] = sval[9 * lane + j];
        const int k8 = (lane == 31) ? skey[256] : -1;
        const float2 it8 = sval[288];
        __syncwarp();
        ssq_accumulate_step(col, tag, k0.x, it[0], lane);
        ssq_accumulate_step(col, tag, k0.y, it[1], lane);
        ssq_accumulate_step(col, tag, k0.z, it[2], lane);
        ssq_accumulate_step(col, tag, k0.w, it[3], lane);
        ssq_accumulate_step(col, tag, k1.x, it[4], lane);
        ssq_accumulate_step(col, tag, k1.y, it[5], lane);
        ssq_accumulate_step(col, tag, k1.z, it[6], lane);
        ssq_accumulate_step(col, tag, k1.w, it[7], lane);
        ssq_accumulate_step(col, tag, k8, it8, lane);
      }
      __syncwarp();  // the staging area is the exchange buffer of the next frame
    }
    __syncthreads();
    // ---- coalesced store of the tile (and re-zero of the accumulator) -------------------
    {
      float2* outc = P.out + (size_t)ch * P.n_freqs * P.n_frames + f0;
      // each warp re-zeroes exactly the slots it has just read (the pad slots 9m+8 are never touched)
      for (int k = warp; k < 257; k += SSQ_FAST_WARPS) {
        float2* a = acc + lane * AS + acc_phys(k);
        if (lane < nf) outc[(size_t)k * P.n_frames + lane] = *a;
        if (MODE == 0) *a = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();
  }
}

// Tries the specialised kernels; *done=false means "use the generic kernel".
static ssq_status stft_fast_launch(ssq_ctx* ctx, StftParams& P, bool* done) {
  *done = false;
  if (P.n_fft != SSQ_FAST_N || P.hop > SSQ_FAST_MAX_HOP || P.modulated) return SSQ_OK;
  P.F = SSQ_FAST_F;
  P.acc_stride = SSQ_FAST_ACC_STRIDE;
  P.tiles_per_channel = (P.n_frames + P.F - 1) / P.F;
  P.total_tiles = P.tiles_per_channel * P.channels;
  const size_t smem = ((size_t)SSQ_FAST_F * SSQ_FAST_ACC_STRIDE + (size_t)SSQ_FAST_WARPS * SSQ_FAST_N) * sizeof(float2) +
                      ((size_t)(SSQ_FAST_F - 1) * P.hop + SSQ_FAST_N) * sizeof(float);
  const int grid = (int)std::min<int64_t>(P.total_tiles, (int64_t)ctx->num_sms * 2);
  if (P.mode == 0) {
    SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(ssq_stft512_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssq_stft512_kernel<0><<<grid, SSQ_FAST_WARPS * 32, smem, ctx->stream>>>(P);
    SSQ_TRY(ssq_check_launch(ctx, "ssq_stft512_kernel<ssq>"));
    ctx->last_kernel = "ssq_stft512_kernel<ssq>";
  } else {
    SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(ssq_stft512_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssq_stft512_kernel<1><<<grid, SSQ_FAST_WARPS * 32, smem, ctx->stream>>>(P);
    SSQ_TRY(ssq_check_launch(ctx, "ssq_stft512_kernel<stft>"));
    ctx->last_kernel = "ssq_stft512_kernel<stft>";
  }
  *done = true;
  return SSQ_OK;
}

// stft_h32.cuh -- the benchmarked kernel: fused ssq_stft / stft for n_fft = 512,
// hop = 32 (BASELINE.json configs[1], [3], [4]).
//
// Same FFT core as the tile kernel of stft_fast.cuh (which stays as the path for
// other hops); what changed comes from the ncu readings in profiles/README.md:
//
//  * the L1TEX data pipe (1 shared-memory or global wavefront per clock per SM;
//    warp shuffles use the same pipe -- tools/ubench/pipes.cu) is the binding
//    resource, so every change below removes wavefronts or keeps the pipe busy;
//  * samples live in a register sliding window.  With hop == warp size the lane
//    that needs sample n = lane + 32 j of frame f needs the SAME value at position
//    j+1 for frame f-1: a frame costs ONE coalesced 128 B global load and 15
//    register moves, issued right after stage 1 of the previous frame (a whole
//    frame ahead of its use).  No sample tile in shared memory, no load phase;
//  * a warp owns 4 CONSECUTIVE frames of the CTA's 32-frame tile;
//  * the window pair (w, dw) comes from a CTA-shared table (frees 32 registers for
//    the sliding window), the stage-2 twiddles from an 8-row table;
//  * reassignment: SoA staging (every staging access conflict-free), one
//    __syncwarp per step (the tag of step j+1 is written in the shadow of step
//    j's add), bin 256 added by its owner lane without the protocol, collisions
//    resolved in ascending source order (the reference's order), tonal frames
//    (all lanes -> one bin) by a shuffle reduction;
//  * the tile leaves as 256 B row segments.  (Per-warp 32 B quads were tried:
//    partially written sectors thrash L2 at 384 channels -- DRAM traffic 92 GB
//    instead of 47 GB, profiles/README.md r1d.)
#pragma once
#include "stft_fast.cuh"

#define H32_WARPS 8
#define H32_AS 289  // column stride (float2): bin k at k + (k>>3) (max 288); odd -> conflict-free transposed read-out

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void smem_rmw_add(float2* p, float re, float im) {
  float2 t = *p;
  t.x += re;
  t.y += im;
  *p = t;
}

// One bin of the epilogue.  zk = Z[k], zn = Z[512-k]; kf = (float)k.
// MODE 1: store Sx into the frame's column.  MODE 0: returns the item (dest bin or
// -1 when gated, value to add).
template <int MODE, int SQZ>
__device__ __forceinline__ void h32_bin(const StftParams& P, float2* col, int k, float kf, float2 zk, float2 zn, int& kb,
                                        float& vre, float& vim) {
  const float c = zk.x + zn.x, d = zk.y - zn.y;  // 2*Sx
  if (MODE == 1) {
    col[acc_phys(k)] = make_float2(0.5f * c, 0.5f * d);
    return;
  }
  const float a = zk.y + zn.y, b = zn.x - zk.x;  // 2*V
  const float den = fmaf(c, c, d * d);
  const float num = fmaf(b, c, -a * d);
  const float q = num * rcp_approx(den);
  const float binf = fabsf(fmaf(-q, P.cphase, kf));
  const float r = ceilf(binf - 0.5f);
  kb = (int)fminf(fmaxf(r, 0.f), 256.f);  // fmaxf(NaN, 0) = 0 -> bin 0 like the reference
  if (den < P.gate2) kb = -1;              // |Sx| < gamma (ssq_stft.rs:23): dropped
  if (SQZ == SSQ_SQUEEZE_LEBESGUE) {
    vre = P.leb_val;
    vim = 0.f;
  } else {
    vre = c * P.tx_scale;
    vim = d * P.tx_scale;
  }
}

// Rare path of one reassignment step: at least two lanes aim at the same bin.
// Lanes whose bin is uncontended add at once; contended lanes add one at a time in
// ascending lane order = ascending source bin (ssq_stft.rs:277-298).
__device__ __noinline__ void h32_collision(float2* col, unsigned char* T, int kb, float vre, float vim, bool mine,
                                           int lane) {
  const bool on = kb >= 0;
  // tonal frames: every active lane aims at the same bin -> one shuffle reduction, one add
  {
    const unsigned act = __ballot_sync(0xffffffffu, on);
    const int first = __ffs(act) - 1;
    const int kb0 = __shfl_sync(0xffffffffu, kb, first);
    if (__all_sync(0xffffffffu, !on || kb == kb0)) {
      float sr = on ? vre : 0.f, si = on ? vim : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == first) smem_rmw_add(col + acc_phys(kb0), sr, si);
      return;
    }
  }
  if (on && !mine) T[kb] = 0xFF;
  __syncwarp();
  const bool contended = on && T[kb] == 0xFF;
  if (on && !contended) smem_rmw_add(col + acc_phys(kb), vre, vim);
  unsigned m = __ballot_sync(0xffffffffu, contended);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    if (lane == src) smem_rmw_add(col + acc_phys(kb), vre, vim);
    __syncwarp();
  }
}

// Stages 1..3 exchange buffers, split, phase transform and reassignment of ONE frame.
// va/vb hold the windowed samples (stage-1 inputs); col is the frame's Tx (or Sx) column.
struct H32Lane {
  int lane, j2, f1, rd1a, rd1b, g2, wr2;
  float lane_f, j2_f;
  bool l0;
  const float2* tw2;
  float2 tw3a[7], tw3b[7];
};

// 512-point forward DFT of one frame: inputs va[t] = z[lane + 64 t], vb[t] = z[lane + 32 + 64 t];
// outputs va[m] = Z[lane + 64 m], vb[m] = Z[L.j2 + 64 m].  Two swizzled exchanges through xch.
// ROT: the last butterfly of vb uses the conjugate kernel (see stft_h32r.cuh).
// TW2POW: stage-2 twiddles by powers of one table entry (saves 12 shared-memory wavefronts, costs 24
// multiplies: a gain where the data pipe binds -- stft/ssq_stft -- and a loss for istft).
template <bool ROT = false, bool TW2POW = ROT, bool PK = SSQ_PK_DEFAULT>
__device__ __forceinline__ void h32_fft512(const H32Lane& L, float2* xch, float2 (&va)[8], float2 (&vb)[8]) {
  const int lane = L.lane, j2 = L.j2;
  fft8_fwd<PK>(va);
  fft8_fwd<PK>(vb);
  {
    float4* rowa = reinterpret_cast<float4*>(xch + lane * 8);
    float4* rowb = reinterpret_cast<float4*>(xch + (lane + 32) * 8);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      rowa[q ^ L.f1] = make_float4(va[2 * q].x, va[2 * q].y, va[2 * q + 1].x, va[2 * q + 1].y);
      rowb[q ^ L.f1] = make_float4(vb[2 * q].x, vb[2 * q].y, vb[2 * q + 1].x, vb[2 * q + 1].y);
    }
  }
  __syncwarp();
  // ---- stage 2 ---------------------------------------------------------------------------
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    va[t] = xch[L.rd1a + 64 * t];
    vb[t] = xch[L.rd1b + 64 * t];
  }
  __syncwarp();
  {
    float2 w[8];
    w[1] = L.tw2[1];  // W_64^{r t}, r = lane & 7
    if (TW2POW) {     // the other powers by products at most 3 deep
      w[2] = cmulf<PK>(w[1], w[1]);
      w[3] = cmulf<PK>(w[2], w[1]);
      w[4] = cmulf<PK>(w[2], w[2]);
      w[5] = cmulf<PK>(w[4], w[1]);
      w[6] = cmulf<PK>(w[4], w[2]);
      w[7] = cmulf<PK>(w[4], w[3]);
    } else {
#pragma unroll
      for (int t = 2; t < 8; ++t) w[t] = L.tw2[t];
    }
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      va[t] = cmulf<PK>(va[t], w[t]);
      vb[t] = cmulf<PK>(vb[t], w[t]);
    }
  }
  fft8_fwd<PK>(va);
  fft8_fwd<PK>(vb);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    xch[L.wr2 + 8 * (t ^ L.g2)] = va[t];
    xch[L.wr2 + 256 + 8 * (t ^ L.g2)] = vb[t];
  }
  __syncwarp();
  // ---- stage 3 ---------------------------------------------------------------------------
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    va[t] = xch[(lane ^ (8 * (t & 1))) + 64 * t];
    vb[t] = xch[(j2 ^ (8 * (t & 1))) + 64 * t];
  }
  __syncwarp();
  if (ROT) {
    // rotated variant (stft_h32r.cuh): the partner butterfly's twiddles are the conjugates of tw3a
    // (W^{t (512 - e_a)}), except for lane 0 whose partner is j = 32: W^{480 t} = conj(W_16^t)
    const float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, H = 0.70710678118654752440f;
    const float2 w16[7] = {{C1, -S1}, {H, -H}, {S1, -C1}, {0.f, -1.f}, {-S1, -C1}, {-H, -H}, {-C1, -S1}};
    // stage-3 twiddles W^{t e_a} as powers of the first (at most 3 products deep): 12 registers of table
    // traded for 6 complex multiplications -- the kernel is bound by latency, not by issue slots
    float2 w3[7];
    w3[0] = L.tw3a[0];
    w3[1] = cmulf<PK>(w3[0], w3[0]);
    w3[2] = cmulf<PK>(w3[1], w3[0]);
    w3[3] = cmulf<PK>(w3[1], w3[1]);
    w3[4] = cmulf<PK>(w3[3], w3[0]);
    w3[5] = cmulf<PK>(w3[3], w3[1]);
    w3[6] = cmulf<PK>(w3[3], w3[2]);
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      const float2 wa = w3[t - 1];
      va[t] = cmulf<PK>(va[t], wa);
      const float2 wb = L.l0 ? w16[t - 1] : wa;
      vb[t] = cmulcf<PK>(vb[t], wb);  // vb * conj(wb)
    }
    fft8_fwd<PK>(va);
    fft8_inv<PK>(vb);
    return;
  }
#pragma unroll
  for (int t = 1; t < 8; ++t) {
    va[t] = cmulf<PK>(va[t], L.tw3a[t - 1]);
    vb[t] = cmulf<PK>(vb[t], L.tw3b[t - 1]);
  }
  fft8_fwd<PK>(va);  // va[m] = Z[lane + 64 m]
  fft8_fwd<PK>(vb);  // vb[m] = Z[j2 + 64 m]
}

template <int MODE, int SQZ>
__device__ __forceinline__ void h32_frame(const StftParams& P, const H32Lane& L, float2* xch, float2* col,
                                          float2 (&va)[8], float2 (&vb)[8]) {
  const int lane = L.lane, j2 = L.j2;
  float* sre = reinterpret_cast<float*>(xch);                          // [264] staging by SOURCE bin
  float* sim = sre + 264;                                               // [264]
  int* skey = reinterpret_cast<int*>(sre + 528);                        // [264]
  unsigned char* tagA = reinterpret_cast<unsigned char*>(sre + 792);    // [264]
  unsigned char* tagB = tagA + 264;                                     // [264]
  h32_fft512<false, true>(L, xch, va, vb);

  // ---- split + phase transform: lane owns bins lane+64m and j2+64m (and lane 0: 256) ------
  const bool l0 = L.l0;
  int kb256 = -1;
  float re256 = 0.f, im256 = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    // bin lane+64m pairs with 512-(lane+64m) = (64-lane)+64(7-m): vb[7-m]; lane 0: va[(8-m)&7]
    float2 pa = vb[7 - m];
    const float2 alt = va[(8 - m) & 7];
    if (l0) pa = alt;
    int kb;
    float vre, vim;
    h32_bin<MODE, SQZ>(P, col, lane + 64 * m, L.lane_f + (float)(64 * m), va[m], pa, kb, vre, vim);
    if (MODE == 0) {
      sre[lane + 64 * m] = vre;
      sim[lane + 64 * m] = vim;
      skey[lane + 64 * m] = kb;
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    // bin j2+64m pairs with 512-(j2+64m) = lane+64(7-m): va[7-m]; lane 0 (j2=32): vb[7-m]
    float2 pb = va[7 - m];
    const float2 alt = vb[7 - m];
    if (l0) pb = alt;
    int kb;
    float vre, vim;
    h32_bin<MODE, SQZ>(P, col, j2 + 64 * m, L.j2_f + (float)(64 * m), vb[m], pb, kb, vre, vim);
    if (MODE == 0) {
      sre[j2 + 64 * m] = vre;
      sim[j2 + 64 * m] = vim;
      skey[j2 + 64 * m] = kb;
    }
  }
  if (l0) h32_bin<MODE, SQZ>(P, col, 256, 256.f, va[4], va[4], kb256, re256, im256);

  if (MODE == 0) {
    __syncwarp();
    // ---- reassignment: lane owns source bins 8*lane .. 8*lane+7, ascending --------------------
    int kk[8];
    float vr[8], vi[8];
    {
      const int4 k0 = *reinterpret_cast<const int4*>(skey + 8 * lane);
      const int4 k1 = *reinterpret_cast<const int4*>(skey + 8 * lane + 4);
      const float4 r0 = *reinterpret_cast<const float4*>(sre + 8 * lane);
      const float4 r1 = *reinterpret_cast<const float4*>(sre + 8 * lane + 4);
      const float4 i0 = *reinterpret_cast<const float4*>(sim + 8 * lane);
      const float4 i1 = *reinterpret_cast<const float4*>(sim + 8 * lane + 4);
      kk[0] = k0.x; kk[1] = k0.y; kk[2] = k0.z; kk[3] = k0.w; kk[4] = k1.x; kk[5] = k1.y; kk[6] = k1.z; kk[7] = k1.w;
      vr[0] = r0.x; vr[1] = r0.y; vr[2] = r0.z; vr[3] = r0.w; vr[4] = r1.x; vr[5] = r1.y; vr[6] = r1.z; vr[7] = r1.w;
      vi[0] = i0.x; vi[1] = i0.y; vi[2] = i0.z; vi[3] = i0.w; vi[4] = i1.x; vi[5] = i1.y; vi[6] = i1.z; vi[7] = i1.w;
    }
    if (kk[0] >= 0) tagA[kk[0]] = (unsigned char)lane;  // tags do not overlap the staging arrays
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unsigned char* T = (j & 1) ? tagB : tagA;
      unsigned char* Tn = (j & 1) ? tagA : tagB;
      const int kb = kk[j];
      const bool on = kb >= 0;
      const bool mine = !on || T[kb] == (unsigned char)lane;
      if (__all_sync(0xffffffffu, mine)) {
        if (on) smem_rmw_add(col + acc_phys(kb), vr[j], vi[j]);
      } else {
        h32_collision(col, T, kb, vr[j], vi[j], mine, lane);
      }
      if (j < 7 && kk[j + 1] >= 0) Tn[kk[j + 1]] = (unsigned char)lane;
      __syncwarp();
    }
    // bin 256 is the last source: only lane 0 is active, no protocol needed
    if (l0 && kb256 >= 0) smem_rmw_add(col + acc_phys(kb256), re256, im256);
  }
  __syncwarp();  // the staging area is the exchange buffer of the next frame
}

template <int MODE, int SQZ>
__global__ void __launch_bounds__(H32_WARPS * 32, 2) ssq_stft512_h32_kernel(const StftParams P) {
  constexpr int N = 512, AS = H32_AS, F = 32;
  extern __shared__ float2 smem[];
  float2* wtab = smem;        // [512] (w, dw*s)
  float2* tw2tab = smem + N;  // [8][9]: W_64^{r t}, r = lane & 7 (row r, column t; row stride 9 -> distinct banks)
  float2* acc = smem + N + 72;  // [32][AS] the tile's Tx columns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* xch = acc + F * AS + warp * N;  // per-warp exchange / staging buffer [512]

  for (int i = threadIdx.x; i < N; i += blockDim.x) wtab[i] = make_float2(P.win[i], P.dwin[i]);
  if (threadIdx.x < 64)
    tw2tab[(threadIdx.x >> 3) * 9 + (threadIdx.x & 7)] = P.tw[((threadIdx.x >> 3) * (threadIdx.x & 7) * 8) & (N - 1)];
  for (int i = threadIdx.x; i < F * AS; i += blockDim.x) acc[i] = make_float2(0.f, 0.f);

  // ---- per-lane constants ----------------------------------------------------------
  H32Lane L;
  L.lane = lane;
  L.j2 = lane ? 64 - lane : 32;  // second stage-3 butterfly
  L.tw2 = tw2tab + (lane & 7) * 9;  // 8 distinct rows per warp: one wavefront per read
#pragma unroll
  for (int t = 1; t < 8; ++t) {
    L.tw3a[t - 1] = P.tw[(lane * t) & (N - 1)];
    L.tw3b[t - 1] = P.tw[(L.j2 * t) & (N - 1)];
  }
  L.f1 = (lane >> 1) & 3;
  L.rd1a = (lane >> 3) * 8 + ((((lane & 7) >> 1) ^ ((lane >> 4) & 3)) << 1) + (lane & 1);
  L.rd1b = ((lane >> 3) + 4) * 8 + ((((lane & 7) >> 1) ^ (((lane >> 4) + 2) & 3)) << 1) + (lane & 1);
  L.g2 = (lane >> 3) & 1;
  L.wr2 = (lane >> 3) * 64 + (lane & 7);
  L.lane_f = (float)lane;
  L.j2_f = (float)L.j2;
  L.l0 = (lane == 0);
  __syncthreads();

  // ---- the warp's frames of a tile: [f0, f0 + nfr), nfr in 0..4 ------------------------------
  const float* xc = nullptr;
  int64_t f0 = 0;
  int nfr = 0;
  bool inner = false;
  float xw[16];
  auto open_tile = [&](int64_t tile) {  // sets xc/f0/nfr/inner and loads the 16-sample window of frame f0
    const int ch = (int)(tile / P.tiles_per_channel);
    f0 = (tile % P.tiles_per_channel) * F + 4 * warp;
    nfr = (int)max((int64_t)0, min((int64_t)4, P.n_frames - f0));
    xc = P.x + (size_t)ch * P.x_stride;
    // padded index of sample n of frame f is 32 f + n; interior frames map it to x[p - left]
    inner = f0 * 32 - P.left >= 0 && (f0 + 3) * 32 + N - 1 - P.left < P.n;
    if (nfr > 0) {
      const int64_t p = f0 * 32 + lane;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        xw[j] = inner ? __ldg(xc + (p + 32 * j - P.left)) : stft_sample(xc, P.n, p + 32 * j, P.left, P.padtype);
    }
  };

  int64_t tile = blockIdx.x;
  if (tile < P.total_tiles) open_tile(tile);
  for (; tile < P.total_tiles; tile += gridDim.x) {
    const int tch = (int)(tile / P.tiles_per_channel);
    const int64_t tf0 = (tile % P.tiles_per_channel) * F;
    const int tnf = (int)min((int64_t)F, P.n_frames - tf0);
    const int64_t next = tile + gridDim.x;
    const int my_n = nfr;
    if (my_n == 0 && next < P.total_tiles) open_tile(next);  // idle warp of a short last tile
    for (int s = 0; s < my_n; ++s) {
      // ---- stage 1 (consumes the sample window) ---------------------------------------------
      float2 va[8], vb[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float2 w0 = wtab[lane + 64 * t], w1 = wtab[lane + 32 + 64 * t];
        va[t] = make_float2(xw[2 * t] * w0.x, xw[2 * t] * w0.y);
        vb[t] = make_float2(xw[2 * t + 1] * w1.x, xw[2 * t + 1] * w1.y);
      }
      // ---- advance the window: one new sample per lane, or the next tile's first window ------
      if (s + 1 < my_n) {
#pragma unroll
        for (int j = 0; j < 15; ++j) xw[j] = xw[j + 1];
        const int64_t p = (f0 + s + 1) * 32 + lane + 480;
        xw[15] = inner ? __ldg(xc + (p - P.left)) : stft_sample(xc, P.n, p, P.left, P.padtype);
      } else if (next < P.total_tiles) {
        open_tile(next);
      }
      h32_frame<MODE, SQZ>(P, L, xch, acc + (4 * warp + s) * AS, va, vb);
    }
    __syncthreads();
    // ---- coalesced store of the tile: warp -> rows warp + 8 i (physical warp + 9 i), lane -> frame
    {
      float2* a = acc + lane * AS + warp;
      float2* g = P.out + ((size_t)tch * 257 + warp) * P.n_frames + tf0 + lane;
      const size_t gstep = (size_t)8 * P.n_frames;
      const bool ok = lane < tnf;
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const float2 v = a[9 * it];
        if (MODE == 0) a[9 * it] = make_float2(0.f, 0.f);
        if (ok) *g = v;
        g += gstep;
      }
      if (warp == 0) {  // row 256
        const float2 v = a[288];
        if (MODE == 0) a[288] = make_float2(0.f, 0.f);
        if (ok) *g = v;
      }
    }
    __syncthreads();
  }
}

// hop == 32 launch; *done=false means "not applicable".
static ssq_status stft_h32_launch(ssq_ctx* ctx, StftParams& P, bool* done) {
  *done = false;
  if (P.n_fft != 512 || P.hop != 32 || P.modulated || getenv("SSQ_NO_H32")) return SSQ_OK;
  P.F = 32;
  P.acc_stride = H32_AS;
  P.tiles_per_channel = (P.n_frames + 31) / 32;
  P.total_tiles = P.tiles_per_channel * P.channels;
  const size_t smem = ((size_t)512 + 72 + (size_t)32 * H32_AS + (size_t)H32_WARPS * 512) * sizeof(float2);
  const int grid = (int)std::min<int64_t>(P.total_tiles, (int64_t)ctx->num_sms * 2);
  const bool leb = P.squeezing == SSQ_SQUEEZE_LEBESGUE;
  void (*k)(const StftParams) = P.mode == 1 ? ssq_stft512_h32_kernel<1, 0>
                                : leb       ? ssq_stft512_h32_kernel<0, 1>
                                            : ssq_stft512_h32_kernel<0, 0>;
  SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, H32_WARPS * 32, smem, ctx->stream>>>(P);
  const char* name = P.mode == 1 ? "ssq_stft512_h32_kernel<stft>" : "ssq_stft512_h32_kernel<ssq>";
  SSQ_TRY(ssq_check_launch(ctx, name));
  ctx->last_kernel = name;
  *done = true;
  return SSQ_OK;
}

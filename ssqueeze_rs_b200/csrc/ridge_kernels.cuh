// ridge_kernels.cuh -- ridge extraction on a time-frequency map that stays in HBM.
//
// The reference crate declares `ridge::extraction` and leaves it empty (rust/src/ridge/{mod,extraction}.rs: 0 bytes);
// the specification is upstream's forward/backward penalised ridge tracking,
// old/ssqueezepy/ridge_extraction.py:11-232, with the semantics of its sequential JIT code:
//   energy = |Tf|^2;  per ridge:  E = -log(energy / max_f energy + eps)
//   forward   P[f, 0] = E[f, 0];  P[f, t] = E[f, t] + min_g (P[g, t-1] + penalty (s_f - s_g)^2)        (:171-177)
//             ridge[t] = first arg-min_f P[f, t]                                                       (:163-165)
//   backward  t = T-2 .. 0:  r = ridge[t+1], val = P[r, t+1] - E[r, t+1]; the LAST f with
//             |val - (P[f, t] + penalty (s_r - s_f)^2)| < eps replaces ridge[t] (none: it stays)       (:204-214)
//   then energy[ridx - bw : ridx + bw, t] = 0 with Python slice semantics                              (:145-147)
// Arithmetic type follows the input as upstream does (:118-121): complex64 -> float / EPS32, complex128 -> double /
// EPS64.  Every step of the dynamic programme is an IEEE add / multiply / min / compare written with explicit
// round-to-nearest intrinsics (no FMA contraction), so given the same E the indices equal the NumPy restatement
// the tests check it against (pinned to upstream's output) bit for bit; `log` is the one libm call.
//
// One CTA per channel walks the time axis (the recursion is sequential in t); E / P move through shared memory in
// tiles of TT time steps so that every global access is a row segment of TT values.
#pragma once
#include "ssq_common.cuh"

template <typename T> struct RidgeOps;
template <> struct RidgeOps<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float neglog(float a) { return -logf(a); }
  static __device__ __forceinline__ float hyp(float a, float b) { return hypotf(a, b); }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct RidgeOps<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double neglog(double a) { return -log(a); }
  static __device__ __forceinline__ double hyp(double a, double b) { return hypot(a, b); }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};

template <typename T>
struct RidgeParams {
  int channels, F;         // rows (frequencies / scales)
  int64_t Tn;              // time steps
  const T* s;              // [F] log(scales) ('cwt') or scales ('stft'), in T
  const T* s_orig;         // [F] scales as given (ridge_f)
  T penalty, eps;
  T* energy;               // [channels, F, Tn]
  T* E;                    // [channels, F, Tn]
  T* P;                    // [channels, F, Tn]
  int* ridge;              // [channels, Tn] work (forward arg-min, then the backward result)
  int TT;                  // tile length along t
  int bw, n_ridges, ridge_no;
  int* out_idx;            // [channels, Tn, n_ridges]
  T* out_f;                // optional, same shape
  T* out_e;                // optional
};

// energy = |Tf|^2 (np.abs(Tf) ** 2: hypot, then the square)
template <typename T, typename C2>
__global__ void ridge_energy_kernel(const C2* __restrict__ Tf, T* __restrict__ energy, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const C2 v = Tf[i];
  const T a = RidgeOps<T>::hyp((T)v.x, (T)v.y);
  energy[i] = RidgeOps<T>::mul(a, a);
}

// E = -log(energy / column max + eps): one thread per time column (rows are Tn apart, columns adjacent)
template <typename T>
__global__ void ridge_neglog_kernel(const RidgeParams<T> P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ch = blockIdx.y;
  if (t >= P.Tn) return;
  const T* en = P.energy + (size_t)ch * P.F * P.Tn + t;
  T* E = P.E + (size_t)ch * P.F * P.Tn + t;
  T mx = en[0];
  for (int f = 1; f < P.F; ++f) {
    const T v = en[(size_t)f * P.Tn];
    mx = v > mx ? v : mx;  // energies are >= 0
  }
  for (int f = 0; f < P.F; ++f)
    E[(size_t)f * P.Tn] = RidgeOps<T>::neglog(RidgeOps<T>::add(RidgeOps<T>::div(en[(size_t)f * P.Tn], mx), P.eps));
}

// forward recursion; dynamic smem: tile[F][TT + 1] (rounded up to 4 elements) | prev[2][F4] | s[F4], F4 = F rounded up to 4
// (pads: prev = +inf, s = 0, so that a padded candidate never wins the minimum).
// float: the inner minimisation -- 3.7 10^9 candidate evaluations per channel at 257 rows x 56 250 frames, all of
// the kernel's time -- works on four candidates per 128-bit shared-memory load with packed fp32x2 instructions
// (sub / mul / mul / add in round-to-nearest, the same IEEE operations in the same order as the scalar code, so the
// values are bit-identical; the minimum is order-independent on NaN-free input): 3.5 instead of 8 instructions per
// candidate.
template <typename T>
__global__ void __launch_bounds__(1024) ridge_forward_kernel(const RidgeParams<T> P) {
  extern __shared__ __align__(16) unsigned char ridge_smem_raw[];
  T* tile = reinterpret_cast<T*>(ridge_smem_raw);
  const int F = P.F, TT = P.TT, TS = TT + 1;
  const int F4 = (F + 3) & ~3;
  T* prev = tile + (((size_t)F * TS + 3) & ~(size_t)3);
  T* sv = prev + 2 * (size_t)F4;
  const int ch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const T* Eg = P.E + (size_t)ch * F * P.Tn;
  T* Pg = P.P + (size_t)ch * F * P.Tn;
  for (int f = tid; f < F4; f += nt) {
    sv[f] = f < F ? P.s[f] : (T)0;
    prev[f] = prev[F4 + f] = RidgeOps<T>::inf();
  }
  int cur = 0;
  for (int64_t t0 = 0; t0 < P.Tn; t0 += TT) {
    const int nt_tile = (int)min((int64_t)TT, P.Tn - t0);
    __syncthreads();
    for (int i = tid; i < F * TT; i += nt) {
      const int f = i / TT, tt = i - f * TT;
      if (tt < nt_tile) tile[f * TS + tt] = Eg[(size_t)f * P.Tn + t0 + tt];
    }
    __syncthreads();
    for (int tt = 0; tt < nt_tile; ++tt) {
      T* pw = prev + (size_t)(cur ^ 1) * F4;
      const T* pr = prev + (size_t)cur * F4;
      if (t0 + tt == 0) {
        for (int f = tid; f < F; f += nt) pw[f] = tile[f * TS];  // P[:, 0] = E[:, 0]
      } else {
        for (int f = tid; f < F; f += nt) {
          const T sf = sv[f];
          T m = RidgeOps<T>::inf();
          if constexpr (sizeof(T) == 4) {
            const float2 sf2 = make_float2(sf, sf), pen2 = make_float2(P.penalty, P.penalty);
            float m0 = m, m1 = m, m2 = m, m3 = m;
#pragma unroll 2
            for (int g = 0; g < F4; g += 4) {
              const float4 p4 = *reinterpret_cast<const float4*>(pr + g);
              const float4 s4 = *reinterpret_cast<const float4*>(sv + g);
              const float2 d01 = sub2<true>(sf2, make_float2(s4.x, s4.y)), d23 = sub2<true>(sf2, make_float2(s4.z, s4.w));
              const float2 c01 = add2<true>(make_float2(p4.x, p4.y), mul2<true>(pen2, mul2<true>(d01, d01)));
              const float2 c23 = add2<true>(make_float2(p4.z, p4.w), mul2<true>(pen2, mul2<true>(d23, d23)));
              m0 = fminf(m0, c01.x);
              m1 = fminf(m1, c01.y);
              m2 = fminf(m2, c23.x);
              m3 = fminf(m3, c23.y);
            }
            m = fminf(fminf(m0, m1), fminf(m2, m3));
          } else {
            for (int g = 0; g < F; ++g) {
              const T d = RidgeOps<T>::sub(sf, sv[g]);
              const T c = RidgeOps<T>::add(pr[g], RidgeOps<T>::mul(P.penalty, RidgeOps<T>::mul(d, d)));
              m = (c < m) ? c : m;  // np.amin (NaN-free inputs)
            }
          }
          const T pc = RidgeOps<T>::add(tile[f * TS + tt], m);
          tile[f * TS + tt] = pc;
          pw[f] = pc;
        }
      }
      cur ^= 1;
      __syncthreads();
    }
    for (int i = tid; i < F * TT; i += nt) {
      const int f = i / TT, tt = i - f * TT;
      if (tt < nt_tile) Pg[(size_t)f * P.Tn + t0 + tt] = tile[f * TS + tt];
    }
  }
}

// ridge[t] = first arg-min over f of P[f, t] (np.argmin): one thread per column
template <typename T>
__global__ void ridge_argmin_kernel(const RidgeParams<T> P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ch = blockIdx.y;
  if (t >= P.Tn) return;
  const T* p = P.P + (size_t)ch * P.F * P.Tn + t;
  T best = p[0];
  int bi = 0;
  for (int f = 1; f < P.F; ++f) {
    const T v = p[(size_t)f * P.Tn];
    if (v < best) {
      best = v;
      bi = f;
    }
  }
  P.ridge[(size_t)ch * P.Tn + t] = bi;
}

// backward pass; dynamic smem: tileP[F][TT + 1] | tileE[F][TT + 1] | s[F] | int rgs[64]
// The recursion is one short step per time column (F candidates); with the whole CTA on it every step paid two CTA
// barriers (56 250 steps: 180 ms per pass on configs[1]).  The CTA loads a tile of TT columns, then ONE warp walks
// them with warp-level reductions only; the other warps wait at the next tile's barrier.
template <typename T>
__global__ void __launch_bounds__(1024) ridge_backward_kernel(const RidgeParams<T> P) {
  extern __shared__ __align__(16) unsigned char ridge_smem_raw[];
  T* tP = reinterpret_cast<T*>(ridge_smem_raw);
  const int F = P.F, TT = P.TT, TS = TT + 1;
  T* tE = tP + (size_t)F * TS;
  T* sv = tE + (size_t)F * TS;
  int* rgs = reinterpret_cast<int*>(sv + F);  // ridge[t0 .. t0 + TT) of the tile (forward arg-min)
  const int ch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const T* Eg = P.E + (size_t)ch * F * P.Tn;
  const T* Pg = P.P + (size_t)ch * F * P.Tn;
  int* rg = P.ridge + (size_t)ch * P.Tn;
  for (int f = tid; f < F; f += nt) sv[f] = P.s[f];
  if (P.Tn < 2) return;
  // tiles cover [t0, t0 + TT); walk them from the last one down; the step at t needs column t of P and column t+1
  // of P and E at row r only -- r and val for the next step are produced while column t+1 is still resident
  const int64_t last_t0 = ((P.Tn - 1) / TT) * TT;
  bool have = false;  // (r, val) describe ridge[t+1]; registers of warp 0, identical in all its lanes
  int r = 0;
  T val = (T)0;
  for (int64_t t0 = last_t0; t0 >= 0; t0 -= TT) {
    const int nt_tile = (int)min((int64_t)TT, P.Tn - t0);
    __syncthreads();
    for (int i = tid; i < F * TT; i += nt) {
      const int f = i / TT, tt = i - f * TT;
      if (tt < nt_tile) {
        tP[f * TS + tt] = Pg[(size_t)f * P.Tn + t0 + tt];
        tE[f * TS + tt] = Eg[(size_t)f * P.Tn + t0 + tt];
      }
    }
    if (tid < nt_tile) rgs[tid] = rg[t0 + tid];
    __syncthreads();
    if (warp != 0) continue;
    for (int tt = nt_tile - 1; tt >= 0; --tt) {
      int rt = rgs[tt];
      if (have) {
        const T sr = sv[r];
        int hit = -1;
        for (int f = lane; f < F; f += 32) {
          const T d = RidgeOps<T>::sub(sr, sv[f]);
          const T c = RidgeOps<T>::add(tP[f * TS + tt], RidgeOps<T>::mul(P.penalty, RidgeOps<T>::mul(d, d)));
          const T diff = RidgeOps<T>::sub(val, c);
          if ((diff < 0 ? -diff : diff) < P.eps) hit = f;  // ascending f: the last one stays
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hit = max(hit, __shfl_xor_sync(0xffffffffu, hit, o));
        if (hit >= 0) {
          rt = hit;
          if (lane == 0) rg[t0 + tt] = hit;
        }
      }
      // ridge[t] is final now: r and val for the step at t - 1
      r = rt;
      val = RidgeOps<T>::sub(tP[r * TS + tt], tE[r * TS + tt]);
      have = true;
    }
  }
}

// outputs of one ridge and removal of its band from the energy map (Python slice semantics, see header)
template <typename T>
__global__ void ridge_finish_kernel(const RidgeParams<T> P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ch = blockIdx.y;
  if (t >= P.Tn) return;
  const int r = P.ridge[(size_t)ch * P.Tn + t];
  T* en = P.energy + (size_t)ch * P.F * P.Tn + t;
  const size_t o = ((size_t)ch * P.Tn + t) * P.n_ridges + P.ridge_no;
  P.out_idx[o] = r;
  if (P.out_f) P.out_f[o] = P.s_orig[r];
  if (P.out_e) P.out_e[o] = en[(size_t)r * P.Tn];
  int lo = r - P.bw, hi = r + P.bw;
  if (lo < 0) lo = max(lo + P.F, 0);
  if (hi < 0) hi = max(hi + P.F, 0);
  lo = min(lo, P.F);
  hi = min(hi, P.F);
  for (int f = lo; f < hi; ++f) en[(size_t)f * P.Tn] = (T)0;
}

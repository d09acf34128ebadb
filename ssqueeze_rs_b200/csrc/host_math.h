// host_math.h -- small float64 host-side helpers of libssqcuda (window fit,
// derivative window, default scales).  These are O(n_fft) / O(n_fft log n_fft)
// set-up steps the reference also does once per call on the host
// (ssq_stft.rs:104-179), not part of the per-sample hot path.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstdint>

namespace ssqhost {

typedef std::complex<double> cd;

// exp(-i pi m^2 / n) with the phase reduced in integers (m^2 mod 2n): exact argument for any m
inline cd chirp(int64_t m, int64_t n) {
  const double PI = 3.14159265358979323846;
  const int64_t r = (int64_t)(((unsigned __int128)m * (unsigned __int128)m) % (unsigned __int128)(2 * n));
  const double ang = -PI * (double)r / (double)n;
  return cd(std::cos(ang), std::sin(ang));
}

// plain DFT / FFT in double: radix-2 for powers of two, Bluestein on top of it otherwise (tiny n: direct sum)
inline void dft(std::vector<cd>& a, bool inverse) {
  const size_t n = a.size();
  if (n <= 1) return;
  const double PI = 3.14159265358979323846;
  if ((n & (n - 1)) != 0 && n > 32) {
    // X[k] = c[k] sum_j (a[j] c[j]) conj(c[k - j]), c[m] = exp(-i pi m^2 / n) (conjugated for the inverse)
    size_t M = 1;
    while (M < 2 * n - 1) M <<= 1;
    std::vector<cd> u(M, cd(0.0, 0.0)), h(M, cd(0.0, 0.0));
    for (size_t j = 0; j < n; ++j) {
      cd c = chirp((int64_t)j, (int64_t)n);
      if (inverse) c = std::conj(c);
      u[j] = a[j] * c;
      h[j] = std::conj(c);
      if (j) h[M - j] = std::conj(c);
    }
    dft(u, false);
    dft(h, false);
    for (size_t i = 0; i < M; ++i) u[i] *= h[i];
    dft(u, true);
    for (size_t k = 0; k < n; ++k) {
      cd c = chirp((int64_t)k, (int64_t)n);
      if (inverse) c = std::conj(c);
      a[k] = u[k] * c / (double)M;
    }
    return;
  }
  if ((n & (n - 1)) == 0) {
    for (size_t i = 1, j = 0; i < n; ++i) {
      size_t bit = n >> 1;
      for (; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
      const double ang = 2.0 * PI / (double)len * (inverse ? 1.0 : -1.0);
      for (size_t i = 0; i < n; i += len) {
        for (size_t k = 0; k < len / 2; ++k) {
          cd w(std::cos(ang * (double)k), std::sin(ang * (double)k));
          cd u = a[i + k], v = a[i + k + len / 2] * w;
          a[i + k] = u + v;
          a[i + k + len / 2] = u - v;
        }
      }
    }
    return;
  }
  std::vector<cd> out(n);
  for (size_t k = 0; k < n; ++k) {
    cd s(0.0, 0.0);
    for (size_t j = 0; j < n; ++j) {
      const double ang = 2.0 * PI * (double)((k * j) % n) / (double)n * (inverse ? 1.0 : -1.0);
      s += a[j] * cd(std::cos(ang), std::sin(ang));
    }
    out[k] = s;
  }
  a.swap(out);
}

// ssq_stft.rs:104-119
inline std::vector<double> fit_window(const double* w, int64_t len, int n_fft) {
  std::vector<double> out((size_t)n_fft, 0.0);
  if (len < n_fft) {
    const int64_t left = (n_fft - len) / 2;
    for (int64_t i = 0; i < len; ++i) out[(size_t)(i + left)] = w[i];
  } else {
    const int64_t start = (len - n_fft) / 2;
    for (int i = 0; i < n_fft; ++i) out[(size_t)i] = w[start + i];
  }
  return out;
}

// ssq_stft.rs:131-179
inline std::vector<double> diff_window(const std::vector<double>& w) {
  const size_t n = w.size();
  const double PI = 3.14159265358979323846;
  std::vector<cd> W(n);
  for (size_t i = 0; i < n; ++i) W[i] = cd(w[i], 0.0);
  dft(W, false);
  for (size_t i = 0; i < n; ++i) {
    double f = (i <= n / 2) ? (double)i : (double)i - (double)n;
    f *= 2.0 * PI / (double)n;
    W[i] = cd(-W[i].imag() * f, W[i].real() * f);
  }
  dft(W, true);
  std::vector<double> out(n);
  const double sc = 1.0 / (double)n;
  for (size_t i = 0; i < n; ++i) out[i] = W[i].real() * sc;
  return out;
}

// utils/array.rs:9-11
inline int64_t next_power_of_2(int64_t n) {
  if (n <= 0) return 1;
  return (int64_t)1 << (int64_t)std::ceil(std::log2((double)n));
}

// cwt.rs:461-489 / cwt_simd.rs:474-545
inline int64_t default_scales(int64_t n, int nv, int simd, double* out) {
  const double log_min = std::log2(2.0);
  const double log_max = std::log2((double)n * 0.5);
  const double v = std::ceil((log_max - log_min) * (double)nv);
  if (!(v > 0.0)) return 0;
  const int64_t ns = (int64_t)v;
  if (!out) return ns;
  const double sf = ns > 1 ? (log_max - log_min) / (double)(ns - 1) : 0.0;
  const double ln2 = 0.693147180559945309417232121458;
  for (int64_t i = 0; i < ns; ++i) {
    const double p = log_min + (double)i * sf;
    out[i] = (simd && ns >= 16) ? std::exp(p * ln2) : std::pow(2.0, p);
  }
  return ns;
}

// Synchrosqueezing admissibility constant of the reference's wavelets (cwt.rs:492-547):
// Css = integral_0^inf psi-hat(w) / w dw  (old/ssqueezepy/utils/cwt_utils.py:28-47).
//   gmw (gamma 3, beta 60, unnormalised): 2 int w^59 exp(-w^3) dw = (2/3) Gamma(20) = (2/3) 19!
//   morlet (mu 6): composite Simpson on [0, 48]; the integrand tends to 6 c exp(-18) at w -> 0.
inline double admissibility_ssq(bool morlet) {
  if (!morlet) return (2.0 / 3.0) * std::tgamma(20.0);
  const double c = std::pow(3.14159265358979323846, -0.25) * std::sqrt(2.0);
  auto f = [c](double w) {
    if (w < 1e-8) return c * std::exp(-18.0) * 6.0;
    return c * (std::exp(-0.5 * (w - 6.0) * (w - 6.0)) - std::exp(-18.0) * std::exp(-0.5 * w * w)) / w;
  };
  const int m = 1 << 18;
  const double h = 48.0 / m;
  double acc = f(0.0) + f(48.0);
  for (int i = 1; i < m; ++i) acc += f(i * h) * ((i & 1) ? 4.0 : 2.0);
  return acc * h / 3.0;
}

}  // namespace ssqhost

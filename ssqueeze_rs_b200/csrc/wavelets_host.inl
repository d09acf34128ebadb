// wavelets_host.inl -- the wavelet generator surface the reference declares in src/ssqueeze/_rs.pyi:91-132 and
// implements in rust/src/wavelets/{morlet,gmw,base}.rs (the #[pyfunction]s are never registered by lib.rs:25-32).
// O(n) host arithmetic in double, as in the reference; SURVEY 8(f) rank 2.
namespace ssqwav {

// gmw.rs:165-190 (Lanczos, g = 7, the standard nine coefficients)
static double gamma_fn(double x) {
  const double PI = 3.14159265358979323846;
  if (x < 0.5) return PI / (std::sin(PI * x) * gamma_fn(1.0 - x));
  static const double p[8] = {676.5203681218851,     -1259.1392167224028,  771.32342877765313,
                              -176.61502916214059,   12.507343278686905,   -0.13857109526572012,
                              9.9843695780195716e-6, 1.5056327351493116e-7};
  x -= 1.0;
  double y = 0.99999999999980993;
  for (int i = 0; i < 8; ++i) y += p[i] / (x + (double)i + 1.0);
  const double t = x + 8.0 - 0.5;
  return std::sqrt(2.0 * PI) * std::pow(t, x + 0.5) * std::exp(-t) * y;
}

static double factorial(int n) {  // gmw.rs:193-198
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= (double)i;
  return r;
}

static double binomial(int n, int k) {  // gmw.rs:201-221
  if (k < 0 || k > n) return 0.0;
  if (k == 0 || k == n) return 1.0;
  if (n <= 20) return factorial(n) / (factorial(k) * factorial(n - k));
  double c = 0.0;
  for (int i = 1; i <= k; ++i) c += std::log((double)(n - k + i)) - std::log((double)i);
  return std::exp(c);
}

// morlet.rs:21-45
static void morlet_psih(const double* w, int64_t n, double mu, double* out) {
  const double PI = 3.14159265358979323846;
  const double cs = std::pow(1.0 + std::exp(-mu * mu) - 2.0 * std::exp(-0.75 * mu * mu), -0.5);
  const double ks = std::exp(-0.5 * mu * mu);
  const double factor = std::sqrt(2.0) * cs * std::pow(PI, 0.25);
  for (int64_t i = 0; i < n; ++i) {
    const double t1 = std::exp(-0.5 * (w[i] - mu) * (w[i] - mu));
    const double t2 = ks * std::exp(-0.5 * w[i] * w[i]);
    out[2 * i] = factor * (t1 - t2);
    out[2 * i + 1] = 0.0;
  }
}

// gmw.rs:30-163
static void gmw_psih(const double* w, int64_t n, double gamma, double beta, bool bandpass, int order, double* out) {
  const double PI = 3.14159265358979323846;
  const double wc = std::pow(beta / gamma, 1.0 / gamma);
  const double r = (2.0 * beta + 1.0) / gamma;
  for (int64_t i = 0; i < 2 * n; ++i) out[i] = 0.0;
  if (order == 0) {
    const double nc = bandpass ? 2.0 / std::exp(beta * std::log(wc) - std::pow(wc, gamma))
                               : std::sqrt(2.0 * PI * gamma * std::pow(2.0, r) / gamma_fn(r));
    for (int64_t i = 0; i < n; ++i) {
      if (w[i] <= 0.0) continue;
      out[2 * i] = bandpass ? nc * std::exp(beta * std::log(w[i]) - std::pow(w[i], gamma))
                            : nc * std::pow(w[i], beta) * std::exp(-std::pow(w[i], gamma));
    }
    return;
  }
  const double c = r - 1.0;
  const int k = order, ci = (int)c;
  const double coeff = bandpass ? 2.0 * std::sqrt(gamma_fn(r) * gamma_fn((double)k + 1.0) / gamma_fn((double)k + r))
                                : std::sqrt(2.0 * PI * gamma * std::pow(2.0, r) * gamma_fn((double)k + 1.0) /
                                            gamma_fn((double)k + r));
  for (int64_t i = 0; i < n; ++i) {
    if (w[i] <= 0.0) continue;
    const double x = 2.0 * std::pow(w[i], gamma);
    double lag = 0.0;  // gmw.rs:54-66
    for (int m = 0; m <= k; ++m)
      lag += binomial(k + ci + 1, ci + m + 1) * binomial(k, m) * ((m & 1) ? -1.0 : 1.0) * std::pow(x, m) / factorial(m);
    if (bandpass)
      out[2 * i] = coeff * lag * std::exp(-beta * std::log(wc) + std::pow(wc, gamma) + beta * std::log(w[i]) - std::pow(w[i], gamma));
    else
      out[2 * i] = coeff * lag * std::pow(w[i], beta) * std::exp(-std::pow(w[i], gamma));
  }
}

// base.rs:18-33
static std::vector<double> xifn(double scale, int64_t n) {
  std::vector<double> xi((size_t)n);
  const double h = scale * (2.0 * 3.14159265358979323846) / (double)n;
  for (int64_t i = 0; i < n; ++i) xi[(size_t)i] = (i <= n / 2 ? (double)i : (double)(i - n)) * h;
  return xi;
}

// morlet.rs:104-145 / gmw.rs:283-327: psih (-1)^i, Nyquist halved for even n, unnormalised inverse DFT / n
static void to_time(double* psih, int64_t n) {
  std::vector<ssqhost::cd> a((size_t)n);
  for (int64_t i = 0; i < n; ++i) a[(size_t)i] = ssqhost::cd(psih[2 * i], psih[2 * i + 1]) * ((i & 1) ? -1.0 : 1.0);
  if (n % 2 == 0 && n > 0) a[(size_t)(n / 2)] /= 2.0;
  ssqhost::dft(a, true);
  for (int64_t i = 0; i < n; ++i) {
    psih[2 * i] = a[(size_t)i].real() / (double)n;
    psih[2 * i + 1] = a[(size_t)i].imag() / (double)n;
  }
}

}  // namespace ssqwav

// kind: 0 = evaluate at the given w[n]; 1 = on xifn(scale, n) (the *_freq functions); 2 = time domain (*_time).
// out: complex128 [n].
extern "C" ssq_status ssq_wavelet_morlet(int kind, const double* w, int64_t n, double scale, double mu, double* out) {
  if (!out || n < 0 || (kind == 0 && !w && n > 0)) return ssq_fail(nullptr, SSQ_EINVAL, "morlet: NULL argument");
  if (kind == 0) {
    ssqwav::morlet_psih(w, n, mu, out);
    return SSQ_OK;
  }
  const std::vector<double> xi = ssqwav::xifn(scale, n);
  ssqwav::morlet_psih(xi.data(), n, mu, out);
  if (kind == 2) ssqwav::to_time(out, n);
  return SSQ_OK;
}

// norm_bandpass: norm.to_lowercase() == "bandpass" (gmw.rs:24, :44); validate: the checks of `gmw` (gmw.rs:238-246),
// which gmw_freq / gmw_time do not make
extern "C" ssq_status ssq_wavelet_gmw(int kind, const double* w, int64_t n, double scale, double gamma, double beta,
                                      int norm_bandpass, int order, double* out) {
  if (!out || n < 0 || (kind == 0 && !w && n > 0)) return ssq_fail(nullptr, SSQ_EINVAL, "gmw: NULL argument");
  if (kind == 0) {
    if (gamma <= 0.0) return ssq_fail(nullptr, SSQ_EINVAL, "gamma must be positive");
    if (beta < 0.0) return ssq_fail(nullptr, SSQ_EINVAL, "beta must be non-negative");
    if (order < 0) return ssq_fail(nullptr, SSQ_EINVAL, "order must be non-negative");
    ssqwav::gmw_psih(w, n, gamma, beta, norm_bandpass != 0, order, out);
    return SSQ_OK;
  }
  const std::vector<double> xi = ssqwav::xifn(scale, n);
  ssqwav::gmw_psih(xi.data(), n, gamma, beta, norm_bandpass != 0, order, out);
  if (kind == 2) ssqwav::to_time(out, n);
  return SSQ_OK;
}

// gmw.rs:331-357; kind: 0 "peak", 1 "energy"
extern "C" ssq_status ssq_wavelet_gmw_center_frequency(double gamma, double beta, int kind, double* out) {
  if (!out) return ssq_fail(nullptr, SSQ_EINVAL, "out is NULL");
  if (kind == 0)
    *out = std::pow(beta / gamma, 1.0 / gamma);
  else if (kind == 1)
    *out = (1.0 / std::pow(2.0, 1.0 / gamma)) * (ssqwav::gamma_fn((2.0 * beta + 2.0) / gamma) / ssqwav::gamma_fn((2.0 * beta + 1.0) / gamma));
  else
    return ssq_fail(nullptr, SSQ_EINVAL, "Unknown center frequency kind");
  return SSQ_OK;
}

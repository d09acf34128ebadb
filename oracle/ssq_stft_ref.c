/* ssq_stft_ref.c -- CPU restatement of the reference's ssq_stft / stft hot path in
 * plain C (float64, OpenMP).  TEST INFRASTRUCTURE / TIMED CPU BASELINE, NOT
 * PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.
 *
 * Why a restatement: the reference (Rust + rustfft 6.2 + rayon 1.10) cannot be
 * built in this image (no rustc/cargo, crates not vendored), so the "Rayon CPU
 * path timed on the box's host cores" is reproduced here with the same
 * algorithm and the same parallel structure:
 *
 *   ssq_stft_as_written   (ssq_stft.rs:122-303)
 *     - reflect/zero pad                                   stft_utils.rs:19-65
 *     - derivative window via FFT                          ssq_stft.rs:131-179
 *     - frames in parallel (rayon par_iter -> omp for), a NEW FFT plan per
 *       frame (`FftPlanner::new()` + `plan_fft_forward` inside the closure,
 *       :198-199; here: the twiddle table is rebuilt per frame), two full
 *       complex FFTs per frame (:226-227)
 *     - serial strided gather into Sx/dSx [n_freqs, n_frames]   :247-252
 *     - SERIAL phase transform                             :11-39, 264
 *     - SERIAL reassignment with the O(n_freqs) linear arg-min per bin :276-301
 *   ssq_stft_as_intended: same numerics; one shared plan, O(1) closed-form bin
 *     (checked against the arg-min neighbours), phase + reassignment parallel
 *     over frames.  Reported beside the as-written figure so the GPU speed-up
 *     is not inflated by the reference's quadratic loop.
 *
 * parity pin: UNPINNED against the Rust binary (see oracle/ssq_oracle.py);
 * pinned against oracle/ssq_oracle.py by tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

typedef struct { double re, im; } cplx;

/* ---- a plain iterative radix-2 FFT with an explicit "plan" (twiddle table) -- */
typedef struct { int n; int log2n; cplx* tw; int* rev; } plan_t;

static int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

static plan_t* plan_create(int n) {
  plan_t* p = (plan_t*)malloc(sizeof(plan_t));
  p->n = n; p->log2n = ilog2(n);
  p->tw = (cplx*)malloc(sizeof(cplx) * (size_t)n);
  p->rev = (int*)malloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double a = -2.0 * PI * (double)i / (double)n;
    p->tw[i].re = cos(a); p->tw[i].im = sin(a);
  }
  if ((1 << p->log2n) == n) {
    for (int i = 0; i < n; ++i) {
      int r = 0;
      for (int b = 0; b < p->log2n; ++b) if (i & (1 << b)) r |= 1 << (p->log2n - 1 - b);
      p->rev[i] = r;
    }
  }
  return p;
}
static void plan_destroy(plan_t* p) { free(p->tw); free(p->rev); free(p); }

/* forward (sign=-1) or inverse (sign=+1) unnormalised DFT, in place */
static void fft_exec(const plan_t* p, cplx* a, int sign, cplx* scratch) {
  const int n = p->n;
  if ((1 << p->log2n) == n) {
    for (int i = 0; i < n; ++i) { int r = p->rev[i]; if (i < r) { cplx t = a[i]; a[i] = a[r]; a[r] = t; } }
    for (int len = 2; len <= n; len <<= 1) {
      const int half = len >> 1, step = n / len;
      for (int i = 0; i < n; i += len) {
        for (int k = 0; k < half; ++k) {
          cplx w = p->tw[k * step];
          if (sign > 0) w.im = -w.im;
          cplx u = a[i + k], v = a[i + k + half];
          cplx t = { v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re };
          a[i + k].re = u.re + t.re; a[i + k].im = u.im + t.im;
          a[i + k + half].re = u.re - t.re; a[i + k + half].im = u.im - t.im;
        }
      }
    }
    return;
  }
  for (int k = 0; k < n; ++k) {  /* any length: direct DFT */
    double sr = 0, si = 0;
    for (int j = 0; j < n; ++j) {
      cplx w = p->tw[(int)(((int64_t)k * j) % n)];
      if (sign > 0) w.im = -w.im;
      sr += a[j].re * w.re - a[j].im * w.im;
      si += a[j].re * w.im + a[j].im * w.re;
    }
    scratch[k].re = sr; scratch[k].im = si;
  }
  memcpy(a, scratch, sizeof(cplx) * (size_t)n);
}

/* stft_utils.rs:19-65 */
static double* pad_signal(const double* x, int64_t n, int n_fft, int zero) {
  const int64_t pad = n_fft - 1, left = pad / 2, right = pad - left;
  double* p = (double*)calloc((size_t)(n + pad), sizeof(double));
  memcpy(p + left, x, sizeof(double) * (size_t)n);
  if (!zero) {
    for (int64_t i = 0; i < left; ++i) { int64_t m = left - i; if (m < n) p[i] = x[m]; }
    for (int64_t i = 0; i < right; ++i) { int64_t m = n - 2 - i; if (m >= 0 && m < n) p[n + left + i] = x[m]; }
  }
  return p;
}

/* ssq_stft.rs:131-179 */
static void diff_window(const double* w, int n, double* dw) {
  plan_t* p = plan_create(n);
  cplx* a = (cplx*)malloc(sizeof(cplx) * (size_t)n);
  cplx* s = (cplx*)malloc(sizeof(cplx) * (size_t)n);
  for (int i = 0; i < n; ++i) { a[i].re = w[i]; a[i].im = 0; }
  fft_exec(p, a, -1, s);
  for (int i = 0; i < n; ++i) {
    double f = (i <= n / 2) ? (double)i : (double)i - (double)n;
    f *= 2.0 * PI / (double)n;
    cplx t = { -a[i].im * f, a[i].re * f };
    a[i] = t;
  }
  fft_exec(p, a, +1, s);
  for (int i = 0; i < n; ++i) dw[i] = a[i].re * (1.0 / (double)n);
  free(a); free(s); plan_destroy(p);
}

/* x[n] -> Tx complex128 [n_freqs, n_frames] (interleaved), ssq_freqs[n_freqs].
 * `window` must already be fitted to n_fft.  mode: 0 as-written, 1 as-intended.
 * squeezing: 0 sum, 1 lebesgue.  gamma < 0: default 10*EPS64.
 * Optional Sx_out (complex128, same shape) for the stft check. */
int ssq_stft_ref(const double* x, int64_t n, const double* window, int n_fft, int hop, double fs,
                 int zero_pad, int squeezing, double gamma, int mode, double* Tx, double* ssq_freqs,
                 double* Sx_out) {
  if (n < 1 || hop < 1 || n_fft < 2) return 1;
  const int n_freqs = n_fft / 2 + 1;
  const int64_t n_frames = (n - 1) / hop + 1;
  double* padded = pad_signal(x, n, n_fft, zero_pad);
  double* dw = (double*)malloc(sizeof(double) * (size_t)n_fft);
  diff_window(window, n_fft, dw);
  cplx* Sx = (cplx*)malloc(sizeof(cplx) * (size_t)n_freqs * (size_t)n_frames);
  cplx* dSx = (cplx*)malloc(sizeof(cplx) * (size_t)n_freqs * (size_t)n_frames);
  cplx* T = (cplx*)Tx;
  memset(T, 0, sizeof(cplx) * (size_t)n_freqs * (size_t)n_frames);
  if (gamma < 0) gamma = 10.0 * 2.2204460492503131e-16;
  for (int i = 0; i < n_freqs; ++i) ssq_freqs[i] = (double)i * 0.5 * fs / ((double)n_freqs - 1.0);
  const double dwf = ssq_freqs[1] - ssq_freqs[0];
  const double sfs_step = (0.5 * fs) / (double)(n_freqs - 1);  /* Array1::linspace(0, fs/2, n_freqs) */

  if (mode == 0) {
    /* HOT LOOP A: frames in parallel, plan per frame, per-frame output vectors */
    cplx** outs = (cplx**)malloc(sizeof(cplx*) * (size_t)n_frames);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t f = 0; f < n_frames; ++f) {
      plan_t* p = plan_create(n_fft);               /* ssq_stft.rs:198-199 */
      cplx* a = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
      cplx* b = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
      cplx* s = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
      const double* fr = padded + f * hop;
      for (int i = 0; i < n_fft; ++i) { a[i].re = fr[i] * window[i]; a[i].im = 0; }
      for (int i = 0; i < n_fft; ++i) { b[i].re = fr[i] * dw[i] * fs; b[i].im = 0; }
      fft_exec(p, a, -1, s);
      fft_exec(p, b, -1, s);
      cplx* o = (cplx*)malloc(sizeof(cplx) * 2 * (size_t)n_freqs);
      memcpy(o, a, sizeof(cplx) * (size_t)n_freqs);
      memcpy(o + n_freqs, b, sizeof(cplx) * (size_t)n_freqs);
      outs[f] = o;
      free(a); free(b); free(s); plan_destroy(p);
    }
    /* serial gather, strided writes (ssq_stft.rs:247-252) */
    for (int64_t f = 0; f < n_frames; ++f) {
      for (int i = 0; i < n_freqs; ++i) {
        Sx[(size_t)i * n_frames + f] = outs[f][i];
        dSx[(size_t)i * n_frames + f] = outs[f][n_freqs + i];
      }
      free(outs[f]);
    }
    free(outs);
    /* HOT LOOP B: serial phase transform, row-major (ssq_stft.rs:21-36) */
    double* w = (double*)malloc(sizeof(double) * (size_t)n_freqs * (size_t)n_frames);
    for (int i = 0; i < n_freqs; ++i) {
      const double sfs = (i == n_freqs - 1) ? 0.5 * fs : sfs_step * (double)i;
      for (int64_t j = 0; j < n_frames; ++j) {
        const cplx S = Sx[(size_t)i * n_frames + j], D = dSx[(size_t)i * n_frames + j];
        if (hypot(S.re, S.im) < gamma) w[(size_t)i * n_frames + j] = INFINITY;
        else {
          const double pd = (D.im * S.re - D.re * S.im) / ((S.re * S.re + S.im * S.im) * 6.283185307179586);
          w[(size_t)i * n_frames + j] = fabs(sfs - pd);
        }
      }
    }
    /* HOT LOOP C: serial reassignment, linear arg-min (ssq_stft.rs:276-301) */
    for (int64_t j = 0; j < n_frames; ++j) {
      for (int i = 0; i < n_freqs; ++i) {
        const double wv = w[(size_t)i * n_frames + j];
        if (!isinf(wv)) {
          int k = 0; double md = INFINITY;
          for (int idx = 0; idx < n_freqs; ++idx) {
            const double d = fabs(wv - ssq_freqs[idx]);
            if (d < md) { md = d; k = idx; }
          }
          cplx wt = Sx[(size_t)i * n_frames + j];
          if (squeezing == 1) { wt.re = 1.0 / (double)n_freqs; wt.im = 0; }
          T[(size_t)k * n_frames + j].re += wt.re * dwf;
          T[(size_t)k * n_frames + j].im += wt.im * dwf;
        }
      }
    }
    free(w);
  } else {
    plan_t* p = plan_create(n_fft);
#pragma omp parallel
    {
      cplx* a = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
      cplx* b = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
      cplx* s = (cplx*)malloc(sizeof(cplx) * (size_t)n_fft);
#pragma omp for schedule(static)
      for (int64_t f = 0; f < n_frames; ++f) {
        const double* fr = padded + f * hop;
        for (int i = 0; i < n_fft; ++i) { a[i].re = fr[i] * window[i]; a[i].im = 0; }
        for (int i = 0; i < n_fft; ++i) { b[i].re = fr[i] * dw[i] * fs; b[i].im = 0; }
        fft_exec(p, a, -1, s);
        fft_exec(p, b, -1, s);
        for (int i = 0; i < n_freqs; ++i) {
          const cplx S = a[i], D = b[i];
          Sx[(size_t)i * n_frames + f] = S;
          dSx[(size_t)i * n_frames + f] = D;
          if (hypot(S.re, S.im) < gamma) continue;
          const double sfs = (i == n_freqs - 1) ? 0.5 * fs : sfs_step * (double)i;
          const double pd = (D.im * S.re - D.re * S.im) / ((S.re * S.re + S.im * S.im) * 6.283185307179586);
          const double wv = fabs(sfs - pd);
          int k;
          if (wv != wv) k = 0;
          else {
            double r = ceil(wv / dwf - 0.5);
            if (r < 0) r = 0; if (r > n_freqs - 1) r = n_freqs - 1;
            k = (int)r;
            /* settle against the neighbours with the reference's strict '<' rule */
            double best = fabs(wv - ssq_freqs[k]);
            if (k > 0 && fabs(wv - ssq_freqs[k - 1]) <= best) { best = fabs(wv - ssq_freqs[k - 1]); k -= 1; }
            else if (k + 1 < n_freqs && fabs(wv - ssq_freqs[k + 1]) < best) k += 1;
          }
          cplx wt = S;
          if (squeezing == 1) { wt.re = 1.0 / (double)n_freqs; wt.im = 0; }
          T[(size_t)k * n_frames + f].re += wt.re * dwf;
          T[(size_t)k * n_frames + f].im += wt.im * dwf;
        }
      }
      free(a); free(b); free(s);
    }
    plan_destroy(p);
  }
  if (Sx_out) memcpy(Sx_out, Sx, sizeof(cplx) * (size_t)n_freqs * (size_t)n_frames);
  free(Sx); free(dSx); free(dw); free(padded);
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants all host cores */
void ssq_ref_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

int ssq_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

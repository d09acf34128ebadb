// cwt_host.inl -- host side of the CWT entry points (included by ssqcuda.cu).

struct FftPlanHost {
  int log2L;
  int npass;
  int r[8];
  int log2T;
};

static FftPlanHost fft_plan(int log2L) {
  FftPlanHost p;
  p.log2L = log2L;
  if (log2L <= 12) {
    p.npass = 1;
    p.r[0] = log2L;
    p.log2T = 0;
    return p;
  }
  // 7-bit register passes (fft128_pass_kernel) first AND last, the remaining bits in one or two generic passes in
  // the middle: the first pass generates x-hat psi-hat on the fly and can be skipped for band-limited rows
  // (cwt_skip_level), the last pass carries the fused ssq_cwt epilogue -- both want the fast kernel.
  p.npass = 0;
  if (log2L >= 18) {
    const int rem = log2L - 14;  // 4 .. 13 bits between the two 7-bit passes
    p.r[p.npass++] = 7;
    if (rem == 7) {
      p.r[p.npass++] = 7;
    } else if (rem <= 7) {
      p.r[p.npass++] = rem;
    } else if (rem >= 11) {
      p.r[p.npass++] = 7;
      p.r[p.npass++] = rem - 7;
    } else {  // 8 .. 10
      p.r[p.npass++] = (rem + 1) / 2;
      p.r[p.npass++] = rem / 2;
    }
    p.r[p.npass++] = 7;
  } else if (log2L == 14) {
    p.r[p.npass++] = 7;
    p.r[p.npass++] = 7;
  } else {  // 13, 15 .. 17: 7-bit passes first, the remainder in one or two generic passes of >= 4 bits
    const int n7 = log2L / 7, rem = log2L - 7 * n7;
    if (rem >= 4) {
      for (int i = 0; i < n7; ++i) p.r[p.npass++] = 7;
      p.r[p.npass++] = rem;
    } else {
      for (int i = 0; i < n7 - 1; ++i) p.r[p.npass++] = 7;
      const int two = 7 + rem;  // 8 .. 10 bits in two generic passes
      p.r[p.npass++] = (two + 1) / 2;
      p.r[p.npass++] = two / 2;
    }
  }
  p.log2T = 5;
  return p;
}

static ssq_status cwt_twiddles(ssq_ctx* ctx, int log2L, const float2** lo, const float2** hi, int* tw_s) {
  const int64_t L = (int64_t)1 << log2L;
  const int s = (log2L + 1) / 2;
  const int64_t nlo = (int64_t)1 << s, nhi = (int64_t)1 << (log2L - s);
  if (ctx->cwt_tw_n != L) {
    std::vector<float> h((size_t)(nlo + nhi) * 2);
    for (int64_t i = 0; i < nlo; ++i) {
      const double a = -2.0 * SSQ_PI * (double)i / (double)L;
      h[(size_t)2 * i] = (float)std::cos(a);
      h[(size_t)2 * i + 1] = (float)std::sin(a);
    }
    for (int64_t i = 0; i < nhi; ++i) {
      const double a = -2.0 * SSQ_PI * (double)(i << s) / (double)L;
      h[(size_t)2 * (nlo + i)] = (float)std::cos(a);
      h[(size_t)2 * (nlo + i) + 1] = (float)std::sin(a);
    }
    SSQ_TRY(devbuf_reserve(ctx, ctx->cwt_tw, h.size() * sizeof(float)));
    SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->cwt_tw.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->cwt_tw_n = L;
  }
  *lo = (const float2*)ctx->cwt_tw.p;
  *hi = *lo + nlo;
  *tw_s = s;
  return SSQ_OK;
}

// Runs all passes of one batched FFT.  `base` carries the functor fields; the
// first pass uses base.load_mode, the last base.store_mode; in between plain.
static ssq_status fft_run(ssq_ctx* ctx, const FftPlanHost& pl, FftPass base, int rows, float2* ws0, float2* ws1,
                          int skip = 0) {
  const int64_t L = (int64_t)1 << pl.log2L;
  int log2Ns = 0;
  for (int i = 0; i < skip; ++i) log2Ns += pl.r[i];
  base.up_shift = log2Ns;
  const float2* cur_in = nullptr;
  for (int i = skip; i < pl.npass; ++i) {
    FftPass P = base;
    P.log2L = pl.log2L;
    P.log2Ns = log2Ns;
    P.r = pl.r[i];
    P.log2T = pl.log2T;
    const bool first = (i == skip), last = (i == pl.npass - 1);
    if (!first) P.load_mode = 0;
    if (!last) P.store_mode = 0;
    P.in = first ? base.in : cur_in;  // load_mode 0 on the first pass reads base.in (two-integral icwt)
    float2* o = last ? base.out : ((i & 1) ? ws1 : ws0);
    P.out = o;
    const int R = 1 << P.r, T = 1 << P.log2T;
    dim3 grid((unsigned)(L / ((int64_t)R * T)), (unsigned)rows);
    if (last && base.store_mode == 2) {
      // fused ssq_cwt epilogue: lane pairs carry the W and dW rows of one scale; 16 columns per CTA of 256 threads by
      // default (half the barrier domain of the 512-thread shape: 13.19 vs 13.55 ms per channel on C3), option cwt_fused_tc
      // (cwt_fused_ok guarantees r == 7, log2Ns >= 6 and an even number of rows starting at an even row)
      if (ctx->opt.cwt_fused_tc == 32) {
        dim3 g2((unsigned)(L / ((int64_t)128 * 16)), (unsigned)(rows / 2));
        fft128_pass_kernel<32, true><<<g2, 256, (size_t)32 * 129 * sizeof(float2), ctx->stream>>>(P);
      } else {
      dim3 g2((unsigned)(L / ((int64_t)128 * 32)), (unsigned)(rows / 2));
      const size_t sm = (size_t)64 * 129 * sizeof(float2);
      SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(fft128_pass_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      fft128_pass_kernel<64, true><<<g2, 512, sm, ctx->stream>>>(P);
      }
      SSQ_TRY(ssq_check_launch(ctx, "fft128_pass_kernel<fused ssq_cwt>"));
    } else if (P.r == 7 && P.log2T == 5 && (log2Ns == 0 || log2Ns >= 5) && !ctx->opt.no_fft128) {
      // columns per CTA (option fft128_tc): 32 by default -- 256-thread CTAs, 256 B runs; the 512-thread shape (64
      // columns) was the faster one in round 1 and is 8 % (ssq_cwt) to 15 % (cwt) slower now that the passes issue
      // half the instructions: the barrier domain matters more than the run length
      // (0 = automatic: 16 columns for rows that end in a plain store -- cwt, the STFT rows path: 9.7 vs 10.2 ms per
      // channel pair of C3 --, 32 for the rows whose last pass is the fused ssq_cwt epilogue: 23.5 vs 24.4 ms)
      const int tc_env = ctx->opt.fft128_tc ? ctx->opt.fft128_tc : (base.store_mode == 2 ? 32 : 16);
      int tc = (tc_env == 64 && L >= (int64_t)128 * 64 && !(log2Ns > 0 && log2Ns < 6)) ? 64 : 32;
      if (tc_env == 16 && (log2Ns == 0 || log2Ns >= 4) && L >= (int64_t)128 * 16) tc = 16;
      dim3 g2((unsigned)(L / ((int64_t)128 * tc)), (unsigned)rows);
      const size_t sm = (size_t)tc * 129 * sizeof(float2);
      if (tc == 64) {
        SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(fft128_pass_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        fft128_pass_kernel<64><<<g2, 512, sm, ctx->stream>>>(P);
      } else if (tc == 16) {
        fft128_pass_kernel<16><<<g2, 128, sm, ctx->stream>>>(P);
      } else {
        fft128_pass_kernel<32><<<g2, 256, sm, ctx->stream>>>(P);
      }
      SSQ_TRY(ssq_check_launch(ctx, "fft128_pass_kernel"));
    } else {
      const size_t smem = (size_t)2 * R * T * sizeof(float2);
      SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(fft_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      fft_pass_kernel<<<grid, 256, smem, ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "fft_pass_kernel"));
    }
    cur_in = o;
    log2Ns += P.r;
  }
  return SSQ_OK;
}

struct CwtCall {
  const float* d_x;
  int64_t channels, n, x_stride;
  int wavelet;
  const double* scales;
  int64_t ns;
  double dt;
  int padtype;
  unsigned flags;
  bool derivative;
};

// Computes x-hat for `channels` rows into ctx->ws_fft0 region; returns pointers.
static ssq_status cwt_forward(ssq_ctx* ctx, const CwtCall& c, int log2L, const FftPlanHost& pl, const float2* lo,
                              const float2* hi, int tw_s, float2* xhat, float2* ws0, float2* ws1) {
  FftPass B;
  memset(&B, 0, sizeof(B));
  B.sign = -1;
  B.tw_lo = lo;
  B.tw_hi = hi;
  B.tw_s = tw_s;
  B.load_mode = 1;
  B.x = c.d_x;
  B.x_stride = c.x_stride;
  B.n = c.n;
  B.padtype = c.padtype == SSQ_PAD_ZERO ? SSQ_PAD_ZERO : SSQ_PAD_REFLECT;
  B.store_mode = 0;
  const int64_t L = (int64_t)1 << log2L;
  const int64_t max_rows = std::max<int64_t>(1, std::min<int64_t>(32768, ((int64_t)1 << 31) / (L * 8)));
  for (int64_t r0 = 0; r0 < c.channels; r0 += max_rows) {
    const int rows = (int)std::min<int64_t>(max_rows, c.channels - r0);
    B.row0 = r0;
    B.out = xhat + (size_t)r0 * L;
    SSQ_TRY(fft_run(ctx, pl, B, rows, ws0, ws1));
  }
  return SSQ_OK;
}

static double cwt_denorm_constant(int wavelet) {
  return wavelet == SSQ_WAVELET_MORLET ? 1.0 : 2.0 * std::exp(SSQ_GMW_LOGPEAK);
}

// psi-hat(scale * xi) is exactly 0 in the kernel for scale * xi above the cut-off of psihat()
// (cwt_kernels.cuh), i.e. for spectrum indices >= B = wmax L / (2 pi scale).  When B <= L/128 the
// first radix-128 pass sees a single non-zero input per butterfly (t = 0): it is a broadcast,
// its output at m is the spectrum at m >> 7, bit for bit.  When B <= L/128^2 the same holds for
// the second pass.  Those passes are not run; the next pass reads the spectrum directly.
static int cwt_skip_level(const ssq_ctx* ctx, const CwtCall& c, const FftPlanHost& pl, int64_t si) {
  if (pl.npass < 2 || pl.log2T != 5 || ctx->opt.no_cwt_prune) return 0;
  const double scale = c.scales[si];
  if (!(scale > 0.0)) return 0;
  const double wmax = (c.wavelet == SSQ_WAVELET_MORLET) ? 14.5 : 4.5;
  const double L = (double)((int64_t)1 << pl.log2L);
  const double B = wmax * L / (2.0 * SSQ_PI * scale) * (1.0 + 1e-4) + 2.0;
  // the first pass (and then the second) is a broadcast when the band ends below L / 2^r0 (L / 2^(r0 + r1)); at
  // least one pass must remain
  // (any radices: with only t = 0 non-zero a Stockham pass of radix 2^r copies its input at j to the outputs
  // Ns (q 2^r + t') + k, j = Ns q + k -- the spectrum at m >> (bits so far))
  int lvl = 0;
  if (B <= L / (double)((int64_t)1 << pl.r[0])) lvl = 1;
  if (lvl == 1 && pl.npass >= 3 && B <= L / (double)((int64_t)1 << (pl.r[0] + pl.r[1]))) lvl = 2;
  return lvl;
}

// Shared driver: inverse transforms of rows [g0, g0+rows) in global (channel, scale, which) order.
static ssq_status cwt_inverse_rows(ssq_ctx* ctx, const CwtCall& c, int log2L, const FftPlanHost& pl,
                                   const float2* lo, const float2* hi, int tw_s, const float2* xhat,
                                   const float* d_scales, int nd, float2* outW, float2* outD, int64_t out_cols,
                                   int64_t n1, float out_scale, int64_t g0, int64_t g1, float2* ws0, float2* ws1,
                                   int64_t max_rows, const SsqCwtParams* fused = nullptr) {
  FftPass B;
  memset(&B, 0, sizeof(B));
  B.sign = +1;
  B.tw_lo = lo;
  B.tw_hi = hi;
  B.tw_s = tw_s;
  B.load_mode = 2;
  B.xhat = xhat;
  B.scales = d_scales;
  B.ns = (int)c.ns;
  B.nd = nd;
  B.wavelet = c.wavelet == SSQ_WAVELET_MORLET ? SSQ_WAVELET_MORLET : SSQ_WAVELET_GMW;
  B.inv_dt = (float)(1.0 / c.dt);
  B.store_mode = 1;
  B.outW = outW;
  B.outD = outD;
  B.out_cols = out_cols;
  B.n1 = n1;
  B.out_scale = out_scale;
  B.l2_norm = (c.flags & SSQ_FLAG_L2_NORM) ? 1 : 0;
  if (fused) {  // last pass = fused ssq_cwt epilogue: row pairs (W, dW) must stay together
    B.store_mode = 2;
    B.E = *fused;
    max_rows = std::max<int64_t>(2, max_rows & ~(int64_t)1);
  }
  if (g1 > (int64_t)0x7fffffff) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "cwt: more than 2^31 rows in one call");
  // rows are ordered (channel, scale, which): cut the range into runs of equal skip level
  int64_t r0 = g0;
  while (r0 < g1) {
    const int lvl = cwt_skip_level(ctx, c, pl, (r0 / nd) % c.ns);
    int64_t r1 = r0 + 1;
    while (r1 < g1 && r1 - r0 < max_rows && cwt_skip_level(ctx, c, pl, (r1 / nd) % c.ns) == lvl) ++r1;
    B.row0 = r0;
    SSQ_TRY(fft_run(ctx, pl, B, (int)(r1 - r0), ws0, ws1, lvl));
    r0 = r1;
  }
  return SSQ_OK;
}

// scales -> device fp32, cached by content (as stft_tables caches the window): a batched caller that repeats
// the same scales pays neither a stream synchronisation nor a blocking copy per call
static ssq_status cwt_upload_scales(ssq_ctx* ctx, const double* scales, int64_t ns) {
  if (ctx->cwt_scales.p && (int64_t)ctx->cwt_scales_host.size() == ns &&
      memcmp(ctx->cwt_scales_host.data(), scales, sizeof(double) * (size_t)ns) == 0)
    return SSQ_OK;
  std::vector<float> hs((size_t)ns);
  for (int64_t i = 0; i < ns; ++i) hs[(size_t)i] = (float)scales[i];
  SSQ_TRY(devbuf_reserve(ctx, ctx->cwt_scales, hs.size() * sizeof(float)));
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // queued kernels may still read the previous table
  SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->cwt_scales.p, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->cwt_scales_host.assign(scales, scales + ns);
  return SSQ_OK;
}

static ssq_status rows_bluestein_tables(ssq_ctx* ctx, int N, int64_t M, const float2** d_chirp, const float2** d_filt);  // stft_rows.inl

static ssq_status cwt_prepare(ssq_ctx* ctx, const CwtCall& c, int* log2L, FftPlanHost* pl, const float2** lo,
                              const float2** hi, int* tw_s, const float** d_scales, float2** xhat, float2** ws0,
                              float2** ws1, int64_t* max_rows) {
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (c.n < 1 || c.channels < 1) return ssq_fail(ctx, SSQ_EINVAL, "cwt: empty input");
  if (c.ns < 1) return ssq_fail(ctx, SSQ_EINVAL, "cwt: no scales");
  if (!(c.dt == c.dt) || c.dt == 0.0) return ssq_fail(ctx, SSQ_EINVAL, "cwt: dt must be non-zero");
  const int64_t L = ssqhost::next_power_of_2(c.n + c.n / 2);
  int l2 = 0;
  while (((int64_t)1 << l2) < L) ++l2;
  if (l2 > 27) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "cwt: padded length 2^%d too large", l2);
  *log2L = l2;
  *pl = fft_plan(l2);
  SSQ_TRY(cwt_twiddles(ctx, l2, lo, hi, tw_s));
  // scales -> device fp32
  SSQ_TRY(cwt_upload_scales(ctx, c.scales, c.ns));
  *d_scales = (const float*)ctx->cwt_scales.p;
  // workspaces: x-hat [channels, L]; two ping-pong buffers of max_rows rows (multi-pass only)
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft0, (size_t)c.channels * L * sizeof(float2)));
  *xhat = (float2*)ctx->ws_fft0.p;
  *max_rows = std::max<int64_t>(1, std::min<int64_t>(32768, ((int64_t)1 << 31) / (L * 8)));
  if (pl->npass > 1) {
    // multi-pass rows: keep the ping-pong workspaces of one batch inside L2 (126 MB) so that the
    // intermediate passes never reach HBM; the same addresses are rewritten by the next batch
    int64_t budget = (int64_t)2048 << 20;  // bytes per workspace (SSQ_CWT_WS_MB to experiment with L2-resident batches)
    if (ctx->opt.cwt_ws_mb > 0) budget = ctx->opt.cwt_ws_mb << 20;
    *max_rows = std::max<int64_t>(1, std::min<int64_t>(*max_rows, budget / (L * 8)));
  }
  *ws0 = *ws1 = nullptr;
  if (pl->npass > 1) {
    const int64_t total_rows = std::max<int64_t>(c.channels, c.channels * c.ns * 2);
    const int64_t mr = std::min<int64_t>(*max_rows, total_rows);
    *max_rows = mr;
    const size_t per = (size_t)mr * L * sizeof(float2);
    SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft1, per * (pl->npass > 2 ? 2 : 1)));
    *ws0 = (float2*)ctx->ws_fft1.p;
    *ws1 = pl->npass > 2 ? (float2*)((char*)ctx->ws_fft1.p + per) : nullptr;
  }
  return SSQ_OK;
}

extern "C" ssq_status ssq_cwt_batch_f32(ssq_ctx* ctx, const float* d_x, int64_t channels, int64_t n,
                                        int64_t x_stride, int wavelet, const double* scales, int64_t ns, double dt,
                                        int padtype, unsigned flags, float* d_Wx, float* d_dWx) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!d_x || !scales || !d_Wx) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  CwtCall c;
  c.d_x = d_x;
  c.channels = channels;
  c.n = n;
  c.x_stride = x_stride > 0 ? x_stride : n;
  c.wavelet = wavelet;
  c.scales = scales;
  c.ns = ns;
  c.dt = dt;
  c.padtype = padtype;
  c.flags = flags;
  c.derivative = d_dWx != nullptr;
  int log2L, tw_s;
  FftPlanHost pl;
  const float2 *lo, *hi;
  const float* d_scales;
  float2 *xhat, *ws0, *ws1;
  int64_t max_rows;
  SSQ_TRY(cwt_prepare(ctx, c, &log2L, &pl, &lo, &hi, &tw_s, &d_scales, &xhat, &ws0, &ws1, &max_rows));
  const int64_t L = (int64_t)1 << log2L;
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  SSQ_TRY(cwt_forward(ctx, c, log2L, pl, lo, hi, tw_s, xhat, ws0, ws1));
  const bool rp = (flags & SSQ_FLAG_RPADDED) != 0;
  const int nd = c.derivative ? 2 : 1;
  const float out_scale = (float)(cwt_denorm_constant(c.wavelet) / (double)L);
  SSQ_TRY(cwt_inverse_rows(ctx, c, log2L, pl, lo, hi, tw_s, xhat, d_scales, nd, (float2*)d_Wx, (float2*)d_dWx,
                           rp ? L : n, rp ? 0 : (L - n) / 2, out_scale, 0, channels * ns * nd, ws0, ws1, max_rows));
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  ctx->last_kernel = "fft_pass_kernel";
  return SSQ_OK;
}

// ssq_cwt.rs:447-469 + :135-158: ssq grid and binning constants, in double on the host
static void ssq_cwt_grid(const double* scales, int64_t ns, int64_t n, double dt, int maprange, int freq_dist,
                         std::vector<double>& f, int* is_log, double* f0, double* inv_step) {
  double fmin, fmax;
  if (maprange == SSQ_MAPRANGE_MAXIMAL) {
    fmin = 1.0 / ((double)n * dt);
    fmax = 0.5 / dt;
  } else {
    fmin = 1.0 / scales[ns - 1];
    fmax = 1.0 / scales[0];
  }
  f.resize((size_t)ns);
  if (freq_dist == SSQ_FREQS_LINEAR) {
    const double step = ns > 1 ? (fmax - fmin) / (double)(ns - 1) : 0.0;
    for (int64_t i = 0; i < ns; ++i) f[(size_t)i] = fmin + (double)i * step;
  } else {
    const double lmin = std::log2(fmin), lmax = std::log2(fmax);
    const double sf = ns > 1 ? (lmax - lmin) / (double)(ns - 1) : 0.0;
    for (int64_t i = 0; i < ns; ++i) f[(size_t)i] = std::pow(2.0, lmin + (double)i * sf);
  }
  *is_log = (ns > 1) ? (f[1] / f[0] > 1.1) : 0;
  if (*is_log) {
    const double lm = std::log2(f[0]);
    const double ls = ns > 1 ? (std::log2(f[(size_t)ns - 1]) - lm) / (double)(ns - 1) : 1.0;
    *f0 = lm;
    *inv_step = 1.0 / ls;
  } else {
    const double st = ns > 1 ? (f[(size_t)ns - 1] - f[0]) / (double)(ns - 1) : 1.0;
    *f0 = f[0];
    *inv_step = 1.0 / st;
  }
}

// The last pass can carry the fused epilogue when it is a 7-bit register pass whose outputs of adjacent columns are
// adjacent (Ns >= 32) and the row fits 32-bit column arithmetic.
static bool cwt_fused_ok(const ssq_ctx* ctx, const FftPlanHost& pl, int64_t n, int64_t ns) {
  if (ctx->opt.no_cwt_fused || ctx->opt.no_fft128 || pl.npass < 2 || pl.log2T != 5) return false;
  // (the fused epilogue addresses a channel's Tx with signed 32-bit element offsets)
  return pl.r[pl.npass - 1] == 7 && pl.log2L - 7 >= 5 && pl.log2L >= 12 && n < ((int64_t)1 << 30) &&
         ns * n < ((int64_t)1 << 31);
}

extern "C" ssq_status ssq_ssq_cwt_batch_diag_f32(ssq_ctx* ctx, const float* d_x, int64_t channels, int64_t n,
                                                 int64_t x_stride, int wavelet, const double* scales, int64_t ns,
                                                 double dt, int freq_dist, int padtype, int squeezing, int maprange,
                                                 double gamma, unsigned flags, float* d_Tx, double* ssq_freqs,
                                                 float* d_w, int32_t* d_kb) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!d_x || !scales || !d_Tx) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  CwtCall c;
  c.d_x = d_x;
  c.channels = channels;
  c.n = n;
  c.x_stride = x_stride > 0 ? x_stride : n;
  c.wavelet = wavelet;
  c.scales = scales;
  c.ns = ns;
  c.dt = dt;
  c.padtype = padtype;
  c.flags = 0;  // ssq_cwt is always L1-normalised, unpadded (ssq_cwt.rs:405-435)
  c.derivative = true;
  int log2L, tw_s;
  FftPlanHost pl;
  const float2 *lo, *hi;
  const float* d_scales;
  float2 *xhat, *ws0, *ws1;
  int64_t max_rows;
  SSQ_TRY(cwt_prepare(ctx, c, &log2L, &pl, &lo, &hi, &tw_s, &d_scales, &xhat, &ws0, &ws1, &max_rows));
  const int64_t L = (int64_t)1 << log2L;
  std::vector<double> f;
  int is_log;
  double f0, inv_step;
  ssq_cwt_grid(scales, ns, n, dt, maprange, freq_dist, f, &is_log, &f0, &inv_step);
  if (ssq_freqs) memcpy(ssq_freqs, f.data(), sizeof(double) * (size_t)ns);
  const double K = cwt_denorm_constant(c.wavelet);
  const double g = (gamma != gamma) ? 10.0 * kEps64 : gamma;  // NaN: not given; negative: never gates
  SsqCwtParams S;
  memset(&S, 0, sizeof(S));
  S.ns = (int)ns;
  S.n = n;
  S.gate = (float)(g / K);
  S.gate2f = (float)std::min(std::max(g < 0.0 ? 0.0 : (g / K) * (g / K), 1e-30), 3.0e38);
  S.is_log = is_log;
  S.f0s = (float)(f0 * inv_step);
  S.inv_step = (float)inv_step;
  S.flipud = (flags & SSQ_FLAG_NO_FLIPUD) ? 0 : 1;
  S.squeezing = squeezing == SSQ_SQUEEZE_LEBESGUE ? SSQ_SQUEEZE_LEBESGUE : SSQ_SQUEEZE_SUM;
  S.K = (float)K;
  S.leb_val = (float)(1.0 / (double)ns);
  const size_t stage = (size_t)ns * n * sizeof(float2);
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  SSQ_TRY(cwt_forward(ctx, c, log2L, pl, lo, hi, tw_s, xhat, ws0, ws1));
  SSQ_CUDA_TRY(ctx, cudaMemsetAsync(d_Tx, 0, (size_t)channels * stage, ctx->stream));
  if (cwt_fused_ok(ctx, pl, n, ns)) {
    // one sweep over all (channel, scale) row pairs: the last pass of every pair reassigns straight into Tx
    // the fused epilogue sees the rows before the 1/L of the inverse transform: fold it into K and the gate
    S.Tx = (float2*)d_Tx;
    S.K = (float)(K / (double)L);
    S.gate = (float)(g / K * (double)L);
    S.gate2f = (float)std::min(std::max(g < 0.0 ? 0.0 : (g / K * (double)L) * (g / K * (double)L), 1e-30), 3.0e38);
    S.aux_kb = d_kb;
    S.aux_w = d_w;
    if (d_kb) SSQ_CUDA_TRY(ctx, cudaMemsetAsync(d_kb, 0xff, (size_t)channels * ns * n * sizeof(int), ctx->stream));
    SSQ_TRY(cwt_inverse_rows(ctx, c, log2L, pl, lo, hi, tw_s, xhat, d_scales, 2, nullptr, nullptr, n, (L - n) / 2,
                             (float)(1.0 / (double)L), 0, channels * ns * 2, ws0, ws1, max_rows, &S));
    SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ev_valid = true;
    ctx->last_kernel = "fft128_pass_kernel<fused ssq_cwt>";
    return SSQ_OK;
  }
  // per-channel W', dW' staging [ns, n] each, then the stand-alone reassignment
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux0, stage));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux1, stage));
  for (int64_t ch = 0; ch < channels; ++ch) {
    // rows of channel ch; outputs land at row (cs - ch*ns) of the staging buffers
    float2* W = (float2*)ctx->ws_aux0.p - (size_t)ch * ns * n;
    float2* D = (float2*)ctx->ws_aux1.p - (size_t)ch * ns * n;
    SSQ_TRY(cwt_inverse_rows(ctx, c, log2L, pl, lo, hi, tw_s, xhat, d_scales, 2, W, D, n, (L - n) / 2,
                             (float)(1.0 / (double)L), ch * ns * 2, (ch + 1) * ns * 2, ws0, ws1, max_rows));
    S.W = (const float2*)ctx->ws_aux0.p;
    S.D = (const float2*)ctx->ws_aux1.p;
    S.Tx = (float2*)d_Tx + (size_t)ch * ns * n;
    S.aux_kb = d_kb ? d_kb + (size_t)ch * ns * n : nullptr;
    S.aux_w = d_w ? d_w + (size_t)ch * ns * n : nullptr;
    ssq_cwt_reassign_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(S);
    SSQ_TRY(ssq_check_launch(ctx, "ssq_cwt_reassign_kernel"));
  }
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  ctx->last_kernel = "fft128_pass_kernel+ssq_cwt_reassign_kernel";
  return SSQ_OK;
}

extern "C" ssq_status ssq_ssq_cwt_batch_f32(ssq_ctx* ctx, const float* d_x, int64_t channels, int64_t n,
                                            int64_t x_stride, int wavelet, const double* scales, int64_t ns,
                                            double dt, int freq_dist, int padtype, int squeezing, int maprange,
                                            double gamma, unsigned flags, float* d_Tx, double* ssq_freqs) {
  return ssq_ssq_cwt_batch_diag_f32(ctx, d_x, channels, n, x_stride, wavelet, scales, ns, dt, freq_dist, padtype,
                                    squeezing, maprange, gamma, flags, d_Tx, ssq_freqs, nullptr, nullptr);
}

extern "C" ssq_status ssq_cwt_f64(ssq_ctx* ctx, const double* x, int64_t n, int wavelet, const double* scales,
                                  int64_t ns, double dt, int padtype, unsigned flags, double* Wx, double* dWx) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!x || !scales || !Wx) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (n < 1 || ns < 1) return ssq_fail(ctx, SSQ_EINVAL, "cwt: empty input");
  const int64_t L = ssqhost::next_power_of_2(n + n / 2);
  const int64_t cols = (flags & SSQ_FLAG_RPADDED) ? L : n;
  const size_t cnt = (size_t)ns * cols;
  SSQ_TRY(upload_f64_as_f32(ctx, x, (size_t)n, ctx->ws_in));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, cnt * sizeof(float2)));
  if (dWx) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_misc, cnt * sizeof(float2)));
  SSQ_TRY(ssq_cwt_batch_f32(ctx, (const float*)ctx->ws_in.p, 1, n, n, wavelet, scales, ns, dt, padtype, flags,
                            (float*)ctx->ws_out.p, dWx ? (float*)ctx->ws_misc.p : nullptr));
  SSQ_TRY(download_f32_as_f64(ctx, ctx->ws_out.p, cnt * 2, Wx));
  if (dWx) SSQ_TRY(download_f32_as_f64(ctx, ctx->ws_misc.p, cnt * 2, dWx));
  return SSQ_OK;
}

extern "C" ssq_status ssq_ssq_cwt_f64(ssq_ctx* ctx, const double* x, int64_t n, int wavelet, const double* scales,
                                      int64_t ns, double dt, int freq_dist, int padtype, int squeezing, int maprange,
                                      double gamma, unsigned flags, double* Tx, double* ssq_freqs, double* w,
                                      int32_t* kb) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!x || !scales || !Tx) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (n < 1 || ns < 1) return ssq_fail(ctx, SSQ_EINVAL, "ssq_cwt: empty input");
  const size_t cnt = (size_t)ns * n;
  SSQ_TRY(upload_f64_as_f32(ctx, x, (size_t)n, ctx->ws_in));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, cnt * sizeof(float2)));
  if (w) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux2, cnt * sizeof(float)));
  if (kb) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_misc, cnt * sizeof(int)));
  SSQ_TRY(ssq_ssq_cwt_batch_diag_f32(ctx, (const float*)ctx->ws_in.p, 1, n, n, wavelet, scales, ns, dt, freq_dist,
                                     padtype, squeezing, maprange, gamma, flags, (float*)ctx->ws_out.p, ssq_freqs,
                                     w ? (float*)ctx->ws_aux2.p : nullptr, kb ? (int*)ctx->ws_misc.p : nullptr));
  SSQ_TRY(download_f32_as_f64(ctx, ctx->ws_out.p, cnt * 2, Tx));
  if (w) SSQ_TRY(download_f32_as_f64(ctx, ctx->ws_aux2.p, cnt, w));
  if (kb) {
    SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(kb, ctx->ws_misc.p, cnt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SSQ_OK;
}

// icwt: cwt.rs:548-718, one-integral branch only.
extern "C" ssq_status ssq_icwt_batch_f32(ssq_ctx* ctx, const float* d_Wx, int64_t channels, int64_t ns, int64_t n_cols,
                                         int wavelet, const double* scales, int one_int, int64_t x_len, double x_mean,
                                         unsigned flags, float* d_x) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!d_Wx || !d_x) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (!scales) return ssq_fail(ctx, SSQ_EINVAL, "Scales must be provided");  // cwt.rs:572-575
  if (channels < 1 || ns < 1 || n_cols < 1) return ssq_fail(ctx, SSQ_EINVAL, "icwt: empty Wx");
  if (x_len <= 0) x_len = n_cols;
  if (x_len > n_cols)
    return ssq_fail(ctx, SSQ_EPANIC, "x_len %lld > Wx.shape[1] %lld: index out of bounds in the reference (cwt.rs:613)",
                    (long long)x_len, (long long)n_cols);
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const double adm = (flags & SSQ_FLAG_ADM_EXACT) ? ssqhost::admissibility_ssq(wavelet == SSQ_WAVELET_MORLET)
                     : wavelet == SSQ_WAVELET_MORLET ? 0.776 : 1.0;                                 // cwt.rs:579-583
  const double dj = (ns > 1 && scales[1] > scales[0]) ? std::log(scales[1] / scales[0]) : 0.1;     // :595-599
  const double final_norm = (2.0 / adm) * dj;
  if (!one_int) {
    // Two-integral branch (cwt.rs:629-712): per scale FFT(Wx[i, :x_len]) * conj(psi-hat_i) -> IFFT, real part / x_len
    // / scale, summed over the scales.  By linearity the sum is taken in the frequency domain and ONE inverse
    // transform follows.  Any x_len (rustfft takes any length): powers of two run the row FFT directly, other
    // lengths through Bluestein's identity on rows of M = 2^k >= 2 x_len - 1 (stft_rows.inl).
    const int64_t N = x_len;
    if (N < 2) return ssq_fail(ctx, SSQ_EINVAL, "icwt: x_len %lld", (long long)N);
    const bool pow2 = (N & (N - 1)) == 0;
    int l2 = 0;
    int64_t M = 1;
    while (M < (pow2 ? N : 2 * N - 1)) {
      M <<= 1;
      ++l2;
    }
    if (l2 > 27) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "icwt: rows of 2^%d points", l2);
    const FftPlanHost pl = fft_plan(l2);
    const float2 *lo, *hi;
    int tw_s;
    SSQ_TRY(cwt_twiddles(ctx, l2, &lo, &hi, &tw_s));
    SSQ_TRY(cwt_upload_scales(ctx, scales, ns));
    const float2 *d_chirp = nullptr, *d_filt = nullptr;
    if (!pow2) SSQ_TRY(rows_bluestein_tables(ctx, (int)N, M, &d_chirp, &d_filt));
    const int64_t max_rows = std::max<int64_t>(1, std::min<int64_t>(ns, ((int64_t)1 << 30) / (M * 8)));
    SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft0, (size_t)ns * M * sizeof(float2)));
    SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft1, (size_t)2 * max_rows * M * sizeof(float2)));
    SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux0, (size_t)3 * M * sizeof(float2) + (pow2 ? 0 : (size_t)max_rows * M * sizeof(float2))));
    float2* What = (float2*)ctx->ws_fft0.p;
    float2* ws0 = (float2*)ctx->ws_fft1.p;
    float2* ws1 = ws0 + (size_t)max_rows * M;
    float2* S = (float2*)ctx->ws_aux0.p;
    float2* S2 = S + M;
    float2* S3 = S2 + M;
    float2* tmp = S3 + M;  // Bluestein: forward spectra of a batch of rows
    FftPass B;
    const double K = cwt_denorm_constant(wavelet);
    const int wav = wavelet == SSQ_WAVELET_MORLET ? SSQ_WAVELET_MORLET : SSQ_WAVELET_GMW;
    SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int64_t ch = 0; ch < channels; ++ch) {
      for (int64_t r0 = 0; r0 < ns; r0 += max_rows) {
        const int rows = (int)std::min<int64_t>(max_rows, ns - r0);
        memset(&B, 0, sizeof(B));
        B.tw_lo = lo;
        B.tw_hi = hi;
        B.tw_s = tw_s;
        B.sign = -1;
        B.load_mode = 6;
        B.in = (const float2*)d_Wx + ((size_t)ch * ns + r0) * n_cols;
        B.fr_ld = n_cols;
        B.fr_nfft = (int)N;
        B.fr_chirp = d_chirp;
        B.out = pow2 ? What + (size_t)r0 * M : tmp;
        SSQ_TRY(fft_run(ctx, pl, B, rows, ws0, ws1));
        if (!pow2) {
          FftPass I;
          memset(&I, 0, sizeof(I));
          I.tw_lo = lo;
          I.tw_hi = hi;
          I.tw_s = tw_s;
          I.sign = +1;
          I.load_mode = 4;
          I.in = tmp;
          I.mul = d_filt;
          I.out = What + (size_t)r0 * M;
          SSQ_TRY(fft_run(ctx, pl, I, rows, ws0, ws1));
        }
      }
      // S[k] = sum_i What_i[k] (c[k]) psi-hat_i / scale_i, k < N
      icwt2_accum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(What, M, d_chirp, (int)ns, (int)N,
                                                                              (const float*)ctx->cwt_scales.p, wav, S);
      SSQ_TRY(ssq_check_launch(ctx, "icwt2_accum_kernel"));
      // inverse DFT of length N: directly, or as conj(DFT(conj S)) through the same Bluestein rows
      memset(&B, 0, sizeof(B));
      B.tw_lo = lo;
      B.tw_hi = hi;
      B.tw_s = tw_s;
      if (pow2) {
        B.sign = +1;
        B.in = S;
        B.out = S2;
        SSQ_TRY(fft_run(ctx, pl, B, 1, ws0, ws1));
      } else {
        B.sign = -1;
        B.load_mode = 6;
        B.in = S;
        B.fr_ld = M;
        B.fr_nfft = (int)N;
        B.fr_chirp = d_chirp;
        B.conj_in = 1;
        B.out = S3;
        SSQ_TRY(fft_run(ctx, pl, B, 1, ws0, ws1));
        FftPass I;
        memset(&I, 0, sizeof(I));
        I.tw_lo = lo;
        I.tw_hi = hi;
        I.tw_s = tw_s;
        I.sign = +1;
        I.load_mode = 4;
        I.in = S3;
        I.mul = d_filt;
        I.out = S2;
        SSQ_TRY(fft_run(ctx, pl, I, 1, ws0, ws1));
      }
      // x = Re(.) K final_norm / N + x_mean (the conjugation of the Bluestein route leaves the real part alone)
      icwt2_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(
          S2, pow2 ? nullptr : d_chirp, N, (float)(K * final_norm / (double)N), (float)x_mean, d_x + (size_t)ch * N);
      SSQ_TRY(ssq_check_launch(ctx, "icwt2_finalize_kernel"));
    }
    ctx->last_kernel = pow2 ? "icwt2_accum_kernel" : "icwt2_accum_kernel<bluestein>";
    SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ev_valid = true;
    return SSQ_OK;
  }
  std::vector<float> hn((size_t)ns);
  for (int64_t i = 0; i < ns; ++i)
    hn[(size_t)i] = (flags & SSQ_FLAG_L2_NORM) ? (float)(1.0 / std::sqrt(scales[i])) : 1.f;       // :606-610
  SSQ_TRY(devbuf_reserve(ctx, ctx->cwt_scales, hn.size() * sizeof(float)));
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->cwt_scales.p, hn.data(), hn.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->cwt_scales_host.clear();  // the buffer no longer holds a scale table
  dim3 g((unsigned)((x_len + 255) / 256), (unsigned)channels);
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  icwt_kernel<<<g, 256, 0, ctx->stream>>>((const float2*)d_Wx, ns, n_cols, x_len, (const float*)ctx->cwt_scales.p,
                                          (float)final_norm, (float)x_mean, d_x);
  SSQ_TRY(ssq_check_launch(ctx, "icwt_kernel"));
  ctx->last_kernel = "icwt_kernel";
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  return SSQ_OK;
}

extern "C" ssq_status ssq_icwt_f64(ssq_ctx* ctx, const double* Wx, int64_t ns, int64_t n_cols, int wavelet,
                                   const double* scales, int one_int, int64_t x_len, double x_mean, unsigned flags,
                                   double* x) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!Wx || !x) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (!scales) return ssq_fail(ctx, SSQ_EINVAL, "Scales must be provided");
  if (ns < 1 || n_cols < 1) return ssq_fail(ctx, SSQ_EINVAL, "icwt: empty Wx");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (x_len <= 0) x_len = n_cols;
  const size_t cnt = (size_t)ns * n_cols;
  SSQ_TRY(upload_f64_as_f32(ctx, Wx, cnt * 2, ctx->ws_in));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, (size_t)std::max<int64_t>(x_len, 1) * sizeof(float)));
  SSQ_TRY(ssq_icwt_batch_f32(ctx, (const float*)ctx->ws_in.p, 1, ns, n_cols, wavelet, scales, one_int, x_len, x_mean,
                             flags, (float*)ctx->ws_out.p));
  return download_f32_as_f64(ctx, ctx->ws_out.p, (size_t)x_len, x);
}

extern "C" ssq_status ssq_cwt_admissibility(int wavelet, double* css) {
  if (!css) return SSQ_EINVAL;
  *css = ssqhost::admissibility_ssq(wavelet == SSQ_WAVELET_MORLET);
  return SSQ_OK;
}

// issq_cwt: old/ssqueezepy/_ssq_cwt.py:313-378, full inversion; same column sum as icwt.
extern "C" ssq_status ssq_issq_cwt_batch_f32(ssq_ctx* ctx, const float* d_Tx, int64_t channels, int64_t ns, int64_t n,
                                             int wavelet, const double* scales, float* d_x) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!d_Tx || !d_x) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (!scales) return ssq_fail(ctx, SSQ_EINVAL, "Scales must be provided");
  if (channels < 1 || ns < 1 || n < 1) return ssq_fail(ctx, SSQ_EINVAL, "issq_cwt: empty Tx");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const double css = ssqhost::admissibility_ssq(wavelet == SSQ_WAVELET_MORLET);
  const double dj = (ns > 1 && scales[1] > scales[0]) ? std::log(scales[1] / scales[0]) : 0.1;
  std::vector<float> ones((size_t)ns, 1.f);
  SSQ_TRY(devbuf_reserve(ctx, ctx->cwt_scales, ones.size() * sizeof(float)));
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->cwt_scales.p, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->cwt_scales_host.clear();
  dim3 g((unsigned)((n + 255) / 256), (unsigned)channels);
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  icwt_kernel<<<g, 256, 0, ctx->stream>>>((const float2*)d_Tx, ns, n, n, (const float*)ctx->cwt_scales.p,
                                          (float)((2.0 / css) * dj), 0.f, d_x);
  SSQ_TRY(ssq_check_launch(ctx, "icwt_kernel"));
  ctx->last_kernel = "icwt_kernel";
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  return SSQ_OK;
}

extern "C" ssq_status ssq_issq_cwt_f64(ssq_ctx* ctx, const double* Tx, int64_t ns, int64_t n, int wavelet,
                                       const double* scales, double* x) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!Tx || !x) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (!scales) return ssq_fail(ctx, SSQ_EINVAL, "Scales must be provided");
  if (ns < 1 || n < 1) return ssq_fail(ctx, SSQ_EINVAL, "issq_cwt: empty Tx");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t cnt = (size_t)ns * n;
  SSQ_TRY(upload_f64_as_f32(ctx, Tx, cnt * 2, ctx->ws_in));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, (size_t)n * sizeof(float)));
  SSQ_TRY(ssq_issq_cwt_batch_f32(ctx, (const float*)ctx->ws_in.p, 1, ns, n, wavelet, scales, (float*)ctx->ws_out.p));
  return download_f32_as_f64(ctx, ctx->ws_out.p, (size_t)n, x);
}

// issq_cwt with curve bands (old/ssqueezepy/_ssq_cwt.py:313-402): cc / cw int32 [n][K] (centre row and half-width per
// column and component; cc == -1: no curve at that column).  x: [K + 1][n], the last row is the residual.
extern "C" ssq_status ssq_issq_cwt_components_f64(ssq_ctx* ctx, const double* Tx, int64_t ns, int64_t n, int wavelet,
                                                  const double* scales, const int* cc, const int* cw, int K,
                                                  double* x) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!Tx || !x || !cc || !cw) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (!scales) return ssq_fail(ctx, SSQ_EINVAL, "Scales must be provided");
  if (ns < 1 || n < 1) return ssq_fail(ctx, SSQ_EINVAL, "issq_cwt: empty Tx");
  if (K < 1 || K > SSQ_MAX_COMPONENTS)
    return ssq_fail(ctx, SSQ_EUNSUPPORTED, "issq_cwt: %d components (1..%d supported)", K, SSQ_MAX_COMPONENTS);
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // row bands per column, clipped as upstream does (:388-394): [lower, upper] inclusive, empty where cc == -1
  std::vector<int> lo((size_t)n * K), hi((size_t)n * K);
  for (int64_t j = 0; j < n; ++j)
    for (int c = 0; c < K; ++c) {
      const int64_t ce = cc[j * K + c], wi = cw[j * K + c];
      int64_t u = std::min<int64_t>(std::max<int64_t>(ce + wi, 0), ns);
      int64_t l = std::min<int64_t>(std::max<int64_t>(ce - wi, 0), ns);
      if (ce == -1) {
        u = 0;
        l = 1;
      }
      lo[(size_t)j * K + c] = (int)l;
      hi[(size_t)j * K + c] = (int)u;  // the slice [l, u + 1) of upstream; rows >= ns do not exist
    }
  const double css = ssqhost::admissibility_ssq(wavelet == SSQ_WAVELET_MORLET);
  const double dj = (ns > 1 && scales[1] > scales[0]) ? std::log(scales[1] / scales[0]) : 0.1;
  const size_t cnt = (size_t)ns * n;
  SSQ_TRY(upload_f64_as_f32(ctx, Tx, cnt * 2, ctx->ws_in));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux0, (size_t)2 * n * K * sizeof(int)));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, (size_t)(K + 1) * n * sizeof(float)));
  int* d_lo = (int*)ctx->ws_aux0.p;
  int* d_hi = d_lo + (size_t)n * K;
  SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(d_lo, lo.data(), lo.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(d_hi, hi.data(), hi.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  issq_cwt_components_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      (const float2*)ctx->ws_in.p, (int)ns, n, K, d_lo, d_hi, (float)((2.0 / css) * dj), (float*)ctx->ws_out.p);
  SSQ_TRY(ssq_check_launch(ctx, "issq_cwt_components_kernel"));
  ctx->last_kernel = "issq_cwt_components_kernel";
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // lo / hi staging vectors go out of scope
  return download_f32_as_f64(ctx, ctx->ws_out.p, (size_t)(K + 1) * n, x);
}

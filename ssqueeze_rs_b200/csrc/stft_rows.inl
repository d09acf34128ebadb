// stft_rows.inl -- STFT family through the batched row passes (included by ssqcuda.cu after cwt_host.inl).
//
// The reference takes any n_fft (rustfft: ssq_stft.rs:92,198-199, stft.rs:43-44).  Frames of up to 1024 points that
// are powers of two live in one warp's shared memory (stft_generic_kernel) or in registers (the 256 / 512 / 1024
// kernels); everything else -- n_fft >= 2048, and lengths that are not powers of two -- goes through the row FFT of
// the CWT path: a batch of frames is a batch of rows, the first pass windows the samples on the fly (load functor 3),
// and the generic kernel in rows mode does split, phase transform, reassignment and the coalesced store from the
// spectra.  Lengths that are not powers of two use Bluestein's identity on rows of M = 2^k >= 2 n_fft - 1:
//   Z[k] = c[k] sum_n (z[n] c[n]) conj(c[k - n]),  c[m] = exp(-i pi m^2 / n_fft)
// i.e. one forward row FFT of z c, a multiplication by the (host, float64) spectrum of the chirp filter, one inverse
// row FFT, and the factor c[k] applied where the spectra are read.  The chirp tables are computed in double with the
// phase m^2 reduced mod 2 n_fft in integers.
struct RowsTables {
  std::vector<double> key;  // fitted window + n_fft + kind
  DevBuf buf;               // [chirp n_fft][filter spectrum M] float2
  int64_t M = 0;
};

// Bluestein tables of length N on rows of M points, cached by N: chirp c[n] = exp(-i pi n^2 / N) and the spectrum of
// the chirp filter conj(c[|m|]) (1/M of the inverse row FFT folded in), both from double.
static ssq_status rows_bluestein_tables(ssq_ctx* ctx, int N, int64_t M, const float2** d_chirp, const float2** d_filt) {
  if (ctx->rows_n != N || !ctx->rows_tab.p) {
    std::vector<ssqhost::cd> h((size_t)M, ssqhost::cd(0.0, 0.0));
    std::vector<float> t((size_t)2 * (N + M));
    for (int j = 0; j < N; ++j) {
      const ssqhost::cd c = ssqhost::chirp(j, N);
      t[(size_t)2 * j] = (float)c.real();
      t[(size_t)2 * j + 1] = (float)c.imag();
      h[(size_t)j] = std::conj(c);
      if (j) h[(size_t)(M - j)] = std::conj(c);
    }
    ssqhost::dft(h, false);
    for (int64_t i = 0; i < M; ++i) {
      t[(size_t)2 * (N + i)] = (float)(h[(size_t)i].real() / (double)M);
      t[(size_t)2 * (N + i) + 1] = (float)(h[(size_t)i].imag() / (double)M);
    }
    SSQ_TRY(devbuf_reserve(ctx, ctx->rows_tab, t.size() * sizeof(float)));
    SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->rows_tab.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->rows_n = N;
  }
  *d_chirp = (const float2*)ctx->rows_tab.p;
  *d_filt = *d_chirp + N;
  return SSQ_OK;
}

// Rows of 2^11 / 2^12 points (frames of 2048 / 4096 samples): a radix-128 register pass + one generic pass instead of
// the single shared-memory pass fft_plan() gives short CWT rows -- with it the rows path beats the warp-per-frame
// shared-memory kernel from 2048 points on (n_fft 4096, 32 ch x 1.8 M, hop 1024: 17.7 ms in the generic kernel).
static FftPlanHost fft_plan_rows(int l2) {
  if (l2 != 11 && l2 != 12) return fft_plan(l2);
  FftPlanHost p;
  p.log2L = l2;
  p.npass = 2;
  p.r[0] = 7;
  p.r[1] = l2 - 7;
  p.log2T = 5;
  return p;
}
// frames the shared-memory kernel keeps: powers of two up to 1024 (the register kernels take 256 / 512 / 1024 before
// it), other lengths up to 32
static inline bool stft_rows_skip(bool pow2, int N) { return pow2 ? N <= 1024 : N <= 32; }

static ssq_status stft_rows_run(ssq_ctx* ctx, StftParams P /* by value: per-batch copies */, bool* done) {
  *done = false;
  const int N = P.n_fft;
  const bool pow2 = P.is_pow2;
  if (stft_rows_skip(pow2, N)) return SSQ_OK;  // the shared-memory kernel (radix-4/2, or a direct sum for tiny n)
  int l2 = 0;
  int64_t M = 1;
  while (M < (pow2 ? (int64_t)N : 2 * (int64_t)N - 1)) {
    M <<= 1;
    ++l2;
  }
  if (l2 > 27) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "n_fft=%d: rows of 2^%d points", N, l2);
  const FftPlanHost pl = fft_plan_rows(l2);
  const float2 *lo, *hi;
  int tw_s;
  // (the CWT twiddle cache is keyed by the row length; a CWT call after this one rebuilds it)
  SSQ_TRY(cwt_twiddles(ctx, l2, &lo, &hi, &tw_s));
  const float2 *d_chirp = nullptr, *d_filt = nullptr;
  if (!pow2) SSQ_TRY(rows_bluestein_tables(ctx, N, M, &d_chirp, &d_filt));
  // batches of rows: spectra + (Bluestein: a second row buffer) + ping-pong workspaces of the passes
  const size_t row_bytes = (size_t)M * sizeof(float2);
  const int64_t rows_max = std::max<int64_t>(2, (int64_t)(((size_t)512 << 20) / row_bytes));
  const int64_t cc_max = std::min<int64_t>(P.channels, rows_max);
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft0, (size_t)rows_max * row_bytes * (pow2 ? 1 : 2)));
  if (pl.npass > 1) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft1, (size_t)rows_max * row_bytes * (pl.npass > 2 ? 2 : 1)));
  float2* Z = (float2*)ctx->ws_fft0.p;
  float2* Z2 = pow2 ? nullptr : Z + (size_t)rows_max * M;
  float2* ws0 = pl.npass > 1 ? (float2*)ctx->ws_fft1.p : nullptr;
  float2* ws1 = pl.npass > 2 ? ws0 + (size_t)rows_max * M : nullptr;
  const int n_freqs = P.n_freqs;
  const int64_t n_frames_total = P.n_frames;
  const float* x_all = P.x;
  float2* out_all = P.out;
  const int channels_all = P.channels;
  for (int64_t c0 = 0; c0 < channels_all; c0 += cc_max) {
    const int64_t cc = std::min<int64_t>(cc_max, channels_all - c0);
    const int64_t fb = std::max<int64_t>(1, rows_max / cc);
    for (int64_t f0 = 0; f0 < n_frames_total; f0 += fb) {
      const int64_t nf = std::min<int64_t>(fb, n_frames_total - f0);
      const int rows = (int)(cc * nf);
      FftPass B;
      memset(&B, 0, sizeof(B));
      B.sign = -1;
      B.tw_lo = lo;
      B.tw_hi = hi;
      B.tw_s = tw_s;
      B.load_mode = 3;
      B.x = x_all + (size_t)c0 * P.x_stride;
      B.x_stride = P.x_stride;
      B.n = P.n;
      B.padtype = P.padtype;
      B.fr_wpair = P.wpair;
      B.fr_chirp = d_chirp;
      B.fr_nfft = N;
      B.fr_hop = P.hop;
      B.fr_left = P.left;
      B.fr_count = (int)nf;
      B.fr_first = P.frame0 + f0;
      B.fr_origin = P.x_origin;
      B.store_mode = 0;
      B.out = Z;
      SSQ_TRY(fft_run(ctx, pl, B, rows, ws0, ws1));
      const float2* spectra = Z;
      if (!pow2) {
        FftPass I;
        memset(&I, 0, sizeof(I));
        I.sign = +1;
        I.tw_lo = lo;
        I.tw_hi = hi;
        I.tw_s = tw_s;
        I.load_mode = 4;
        I.in = Z;
        I.mul = d_filt;
        I.store_mode = 0;
        I.out = Z2;
        SSQ_TRY(fft_run(ctx, pl, I, rows, ws0, ws1));
        spectra = Z2;
      }
      StftParams Q = P;
      Q.channels = (int)cc;
      Q.n_frames = nf;
      Q.frame0 = P.frame0 + f0;
      Q.out_ld = n_frames_total;
      const size_t col0 = (size_t)c0 * n_freqs * n_frames_total + f0;
      Q.out = out_all + col0;
      if (P.aux_Sx) Q.aux_Sx = P.aux_Sx + col0;
      if (P.aux_dSx) Q.aux_dSx = P.aux_dSx + col0;
      if (P.aux_w) Q.aux_w = P.aux_w + col0;
      if (P.aux_kb) Q.aux_kb = P.aux_kb + col0;
      Q.zrows = spectra;
      Q.zmul = d_chirp;
      Q.zld = M;
      // tiles of the rows-mode generic kernel: accumulator tile + per-warp tags only
      const size_t budget = std::min<size_t>((size_t)ctx->max_smem_optin, 200 * 1024);
      const int acc_stride = n_freqs | 1;
      int F = 32, nw = 8;
      auto need = [&](int F_, int nw_) { return (size_t)F_ * acc_stride * sizeof(float2) + (size_t)nw_ * ((n_freqs + 7) & ~7); };
      while (F > 1 && need(F, nw) > budget) F >>= 1;
      if (need(F, nw) > budget) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "n_fft=%d: one Tx column does not fit shared memory", N);
      nw = std::min(nw, std::max(1, F));
      if ((int64_t)F > nf) {
        int f2 = 1;
        while (f2 < nf) f2 <<= 1;
        F = std::max(1, std::min(F, f2));
      }
      Q.F = F;
      Q.acc_stride = acc_stride;
      Q.tiles_per_channel = (nf + F - 1) / F;
      Q.total_tiles = Q.tiles_per_channel * cc;
      const size_t smem = need(F, nw);
      const int per_sm = std::max<int>(1, (int)(((size_t)220 * 1024) / (smem + 1024)));
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(Q.total_tiles, (int64_t)ctx->num_sms * per_sm));
      SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(stft_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      stft_generic_kernel<<<grid, nw * 32, smem, ctx->stream>>>(Q);
      SSQ_TRY(ssq_check_launch(ctx, "stft_generic_kernel<rows>"));
    }
  }
  ctx->last_kernel = pow2 ? "fft_pass rows + stft_generic_kernel<rows>" : "bluestein fft_pass rows + stft_generic_kernel<rows>";
  *done = true;
  return SSQ_OK;
}


// istft through the row passes: x = Re(IFFT(Zfull)) = Re(conj(FFT(conj(Zfull)))), so the forward machinery above is
// reused on conj of the Hermitian extension (load functor 5); istft_ola_kernel in rows mode windows, overlap-adds and
// merges tile seams as it does for the shared-memory path.  P: filled as for istft_ola_kernel (wa, hop, L, xacc ...).
static ssq_status istft_rows_run(ssq_ctx* ctx, IstftParams P, bool* done) {
  *done = false;
  const int N = P.n_fft;
  const bool pow2 = P.is_pow2;
  if (stft_rows_skip(pow2, N)) return SSQ_OK;
  int l2 = 0;
  int64_t M = 1;
  while (M < (pow2 ? (int64_t)N : 2 * (int64_t)N - 1)) {
    M <<= 1;
    ++l2;
  }
  if (l2 > 27) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "n_fft=%d: rows of 2^%d points", N, l2);
  const FftPlanHost pl = fft_plan_rows(l2);
  const float2 *lo, *hi;
  int tw_s;
  SSQ_TRY(cwt_twiddles(ctx, l2, &lo, &hi, &tw_s));
  const float2 *d_chirp = nullptr, *d_filt = nullptr;
  if (!pow2) SSQ_TRY(rows_bluestein_tables(ctx, N, M, &d_chirp, &d_filt));
  const size_t row_bytes = (size_t)M * sizeof(float2);
  const int64_t rows_max = std::max<int64_t>(2, (int64_t)(((size_t)512 << 20) / row_bytes));
  const int64_t cc_max = std::min<int64_t>(P.channels, rows_max);
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft0, (size_t)rows_max * row_bytes * (pow2 ? 1 : 2)));
  if (pl.npass > 1) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_fft1, (size_t)rows_max * row_bytes * (pl.npass > 2 ? 2 : 1)));
  float2* Z = (float2*)ctx->ws_fft0.p;
  float2* Z2 = pow2 ? nullptr : Z + (size_t)rows_max * M;
  float2* ws0 = pl.npass > 1 ? (float2*)ctx->ws_fft1.p : nullptr;
  float2* ws1 = pl.npass > 2 ? ws0 + (size_t)rows_max * M : nullptr;
  const int channels_all = P.channels;
  const int64_t n_use_total = P.n_use;
  const float2* Sx_all = P.Sx;
  float* xacc_all = P.xacc;
  const size_t budget = std::min<size_t>((size_t)ctx->max_smem_optin, 200 * 1024);
  for (int64_t c0 = 0; c0 < channels_all; c0 += cc_max) {
    const int64_t cc = std::min<int64_t>(cc_max, channels_all - c0);
    const int64_t fb = std::max<int64_t>(1, rows_max / cc);
    for (int64_t f0 = 0; f0 < n_use_total; f0 += fb) {
      const int64_t nf = std::min<int64_t>(fb, n_use_total - f0);
      const int rows = (int)(cc * nf);
      FftPass B;
      memset(&B, 0, sizeof(B));
      B.sign = -1;
      B.tw_lo = lo;
      B.tw_hi = hi;
      B.tw_s = tw_s;
      B.load_mode = 5;
      B.in = Sx_all + (size_t)c0 * P.n_freqs * P.n_frames;
      B.fr_ld = P.n_frames;
      B.fr_chirp = d_chirp;
      B.fr_nfft = N;
      B.fr_count = (int)nf;
      B.fr_first = f0;
      B.store_mode = 0;
      B.out = Z;
      SSQ_TRY(fft_run(ctx, pl, B, rows, ws0, ws1));
      const float2* spectra = Z;
      if (!pow2) {
        FftPass I;
        memset(&I, 0, sizeof(I));
        I.sign = +1;
        I.tw_lo = lo;
        I.tw_hi = hi;
        I.tw_s = tw_s;
        I.load_mode = 4;
        I.in = Z;
        I.mul = d_filt;
        I.store_mode = 0;
        I.out = Z2;
        SSQ_TRY(fft_run(ctx, pl, I, rows, ws0, ws1));
        spectra = Z2;
      }
      IstftParams Q = P;
      Q.channels = (int)cc;
      Q.n_use = nf;
      Q.xacc = xacc_all + (size_t)c0 * P.L;
      Q.zrows = spectra;
      Q.zmul = d_chirp;
      Q.zld = M;
      Q.fbase = f0;
      const int acc_stride = P.n_freqs | 1;
      int F = 32;
      while (F > 1 && (size_t)F * acc_stride * sizeof(float2) > budget) F >>= 1;
      if ((size_t)F * acc_stride * sizeof(float2) > budget)
        return ssq_fail(ctx, SSQ_EUNSUPPORTED, "n_fft=%d: one frame does not fit shared memory", N);
      if ((int64_t)F > nf) {
        int f2 = 1;
        while (f2 < nf) f2 <<= 1;
        F = std::max(1, std::min(F, f2));
      }
      Q.F = F;
      Q.acc_stride = acc_stride;
      Q.tiles_per_channel = (nf + F - 1) / F;
      Q.total_tiles = Q.tiles_per_channel * cc;
      const size_t smem = (size_t)F * acc_stride * sizeof(float2);
      const int per_sm = std::max<int>(1, (int)(((size_t)220 * 1024) / (smem + 1024)));
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(Q.total_tiles, (int64_t)ctx->num_sms * per_sm));
      SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(istft_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      istft_ola_kernel<<<grid, 256, smem, ctx->stream>>>(Q);
      SSQ_TRY(ssq_check_launch(ctx, "istft_ola_kernel<rows>"));
    }
  }
  ctx->last_kernel = pow2 ? "fft_pass rows + istft_ola_kernel<rows>" : "bluestein fft_pass rows + istft_ola_kernel<rows>";
  *done = true;
  return SSQ_OK;
}

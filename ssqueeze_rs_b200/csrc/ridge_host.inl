// ridge_host.inl -- host side of ridge extraction (included by ssqcuda.cu); device code in ridge_kernels.cuh.

template <typename T, typename C2>
static ssq_status ridge_run(ssq_ctx* ctx, const C2* d_Tf, int64_t channels, int64_t F, int64_t Tn, const double* scales,
                            double penalty, int n_ridges, int bw, int transform, int32_t* d_idx, T* d_f, T* d_e,
                            T* d_E_all) {
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (channels < 1 || F < 1 || Tn < 1) return ssq_fail(ctx, SSQ_EINVAL, "extract_ridges: empty Tf");
  if (n_ridges < 1 || bw < 0) return ssq_fail(ctx, SSQ_EINVAL, "extract_ridges: n_ridges >= 1 and bw >= 0");
  if (F > 8192) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "extract_ridges: %lld rows (up to 8192)", (long long)F);
  // penalty coordinates: log(scales) for 'cwt', scales otherwise (ridge_extraction.py:123-125), in T.  The logarithm
  // is taken in double and rounded (upstream: NumPy's log in the array's dtype, a few ulp from it at most).
  std::vector<T> tab((size_t)2 * F);
  for (int64_t i = 0; i < F; ++i) {
    const T sc = (T)scales[i];
    tab[(size_t)i] = transform == 0 ? (T)std::log((double)sc) : sc;
    tab[(size_t)F + i] = sc;
  }
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_ridge[4], tab.size() * sizeof(T)));
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaMemcpy(ctx->ws_ridge[4].p, tab.data(), tab.size() * sizeof(T), cudaMemcpyHostToDevice));
  // tile length along t: the backward pass holds two [F][TT + 1] tiles
  int TT = 32;
  auto smem_bw = [&](int tt) { return ((size_t)2 * F * (tt + 1) + F) * sizeof(T) + 64 * sizeof(int); };
  while (TT > 1 && smem_bw(TT) > (size_t)200 * 1024) TT >>= 1;
  if (smem_bw(TT) > (size_t)200 * 1024) return ssq_fail(ctx, SSQ_EUNSUPPORTED, "extract_ridges: %lld rows do not fit", (long long)F);
  const size_t smem_fw = ((((size_t)F * (TT + 1) + 3) & ~(size_t)3) + 3 * (((size_t)F + 3) & ~(size_t)3)) * sizeof(T);
  SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(ridge_forward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fw));
  SSQ_CUDA_TRY(ctx, cudaFuncSetAttribute(ridge_backward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bw(TT)));
  const int threads = (int)std::min<int64_t>(1024, ((F + 31) / 32) * 32);
  // channel batches: three [F, Tn] maps of T per channel
  const size_t per_ch = (size_t)F * Tn * sizeof(T);
  const int64_t cb = std::max<int64_t>(1, std::min<int64_t>(channels, (int64_t)(((size_t)20 << 30) / std::max<size_t>(per_ch, 1))));
  for (int k = 0; k < 3; ++k) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_ridge[k], (size_t)cb * per_ch));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_ridge[3], (size_t)cb * Tn * sizeof(int)));
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  for (int64_t c0 = 0; c0 < channels; c0 += cb) {
    const int64_t cc = std::min(cb, channels - c0);
    RidgeParams<T> P;
    memset(&P, 0, sizeof(P));
    P.channels = (int)cc;
    P.F = (int)F;
    P.Tn = Tn;
    P.s = (const T*)ctx->ws_ridge[4].p;
    P.s_orig = P.s + F;
    P.penalty = (T)penalty;
    P.eps = sizeof(T) == 8 ? (T)2.2204460492503131e-16 : (T)1.1920928955078125e-07;
    P.energy = (T*)ctx->ws_ridge[0].p;
    P.E = (T*)ctx->ws_ridge[1].p;
    P.P = (T*)ctx->ws_ridge[2].p;
    P.ridge = (int*)ctx->ws_ridge[3].p;
    P.TT = TT;
    P.bw = bw;
    P.n_ridges = n_ridges;
    P.out_idx = d_idx + (size_t)c0 * Tn * n_ridges;
    P.out_f = d_f ? d_f + (size_t)c0 * Tn * n_ridges : nullptr;
    P.out_e = d_e ? d_e + (size_t)c0 * Tn * n_ridges : nullptr;
    const size_t cnt = (size_t)cc * F * Tn;
    ridge_energy_kernel<T, C2><<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(d_Tf + (size_t)c0 * F * Tn, P.energy, cnt);
    SSQ_TRY(ssq_check_launch(ctx, "ridge_energy_kernel"));
    dim3 gcol((unsigned)((Tn + 127) / 128), (unsigned)cc);
    for (int i = 0; i < n_ridges; ++i) {
      P.ridge_no = i;
      ridge_neglog_kernel<T><<<gcol, 128, 0, ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "ridge_neglog_kernel"));
      if (d_E_all)
        for (int64_t c = 0; c < cc; ++c)
          SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(d_E_all + (((size_t)(c0 + c) * n_ridges + i) * F) * Tn, P.E + (size_t)c * F * Tn,
                                            per_ch, cudaMemcpyDeviceToDevice, ctx->stream));
      ridge_forward_kernel<T><<<(unsigned)cc, threads, smem_fw, ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "ridge_forward_kernel"));
      ridge_argmin_kernel<T><<<gcol, 128, 0, ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "ridge_argmin_kernel"));
      ridge_backward_kernel<T><<<(unsigned)cc, threads, smem_bw(TT), ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "ridge_backward_kernel"));
      ridge_finish_kernel<T><<<gcol, 128, 0, ctx->stream>>>(P);
      SSQ_TRY(ssq_check_launch(ctx, "ridge_finish_kernel"));
    }
  }
  SSQ_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  ctx->last_kernel = "ridge_forward_kernel";
  return SSQ_OK;
}

extern "C" ssq_status ssq_extract_ridges_batch(ssq_ctx* ctx, const void* d_Tf, int is_f64, int64_t channels,
                                               int64_t n_freq, int64_t n_time, const double* scales, double penalty,
                                               int n_ridges, int bw, int transform, int32_t* d_ridge_idxs,
                                               void* d_ridge_f, void* d_ridge_e, void* d_E_all) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!d_Tf || !scales || !d_ridge_idxs) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (is_f64)
    return ridge_run<double, double2>(ctx, (const double2*)d_Tf, channels, n_freq, n_time, scales, penalty, n_ridges, bw,
                                      transform, d_ridge_idxs, (double*)d_ridge_f, (double*)d_ridge_e, (double*)d_E_all);
  return ridge_run<float, float2>(ctx, (const float2*)d_Tf, channels, n_freq, n_time, scales, penalty, n_ridges, bw,
                                  transform, d_ridge_idxs, (float*)d_ridge_f, (float*)d_ridge_e, (float*)d_E_all);
}

// host buffers: Tf complex128 (is_f64) or complex64 [n_freq, n_time]; ridge_idxs int32 [n_time, n_ridges];
// ridge_f / ridge_e / E_all in the real type of Tf (may be NULL)
extern "C" ssq_status ssq_extract_ridges_host(ssq_ctx* ctx, const void* Tf, int is_f64, int64_t n_freq, int64_t n_time,
                                              const double* scales, double penalty, int n_ridges, int bw, int transform,
                                              int32_t* ridge_idxs, void* ridge_f, void* ridge_e, void* E_all) {
  if (!ctx) return ssq_fail(nullptr, SSQ_EINVAL, "ctx is NULL");
  if (!Tf || !scales || !ridge_idxs) return ssq_fail(ctx, SSQ_EINVAL, "NULL argument");
  if (n_freq < 1 || n_time < 1 || n_ridges < 1) return ssq_fail(ctx, SSQ_EINVAL, "extract_ridges: empty Tf / n_ridges < 1");
  SSQ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t rs = is_f64 ? sizeof(double) : sizeof(float);
  const size_t cnt = (size_t)n_freq * n_time, nout = (size_t)n_time * n_ridges;
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_in, cnt * 2 * rs));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_out, nout * sizeof(int)));
  SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux0, 2 * nout * rs));
  if (E_all) SSQ_TRY(devbuf_reserve(ctx, ctx->ws_aux1, (size_t)n_ridges * cnt * rs));
  SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ws_in.p, Tf, cnt * 2 * rs, cudaMemcpyHostToDevice, ctx->stream));
  char* fe = (char*)ctx->ws_aux0.p;
  SSQ_TRY(ssq_extract_ridges_batch(ctx, ctx->ws_in.p, is_f64, 1, n_freq, n_time, scales, penalty, n_ridges, bw, transform,
                                   (int32_t*)ctx->ws_out.p, ridge_f ? fe : nullptr, ridge_e ? fe + nout * rs : nullptr,
                                   E_all ? ctx->ws_aux1.p : nullptr));
  SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(ridge_idxs, ctx->ws_out.p, nout * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (ridge_f) SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(ridge_f, fe, nout * rs, cudaMemcpyDeviceToHost, ctx->stream));
  if (ridge_e) SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(ridge_e, fe + nout * rs, nout * rs, cudaMemcpyDeviceToHost, ctx->stream));
  if (E_all) SSQ_CUDA_TRY(ctx, cudaMemcpyAsync(E_all, ctx->ws_aux1.p, (size_t)n_ridges * cnt * rs, cudaMemcpyDeviceToHost, ctx->stream));
  SSQ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return SSQ_OK;
}

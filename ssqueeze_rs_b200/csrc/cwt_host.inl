// cwt_host.inl -- host side of the CWT entry points (included by ssqcuda.cu)
extern "C" ssq_status ssq_cwt_batch_f32(ssq_ctx* ctx, const float*, int64_t, int64_t, int64_t, int, const double*,
                                        int64_t, double, int, unsigned, float*, float*) {
  return ssq_fail(ctx, SSQ_EUNSUPPORTED, "cwt: not built yet");
}
extern "C" ssq_status ssq_ssq_cwt_batch_f32(ssq_ctx* ctx, const float*, int64_t, int64_t, int64_t, int,
                                            const double*, int64_t, double, int, int, int, int, double, unsigned,
                                            float*, double*) {
  return ssq_fail(ctx, SSQ_EUNSUPPORTED, "ssq_cwt: not built yet");
}
extern "C" ssq_status ssq_cwt_f64(ssq_ctx* ctx, const double*, int64_t, int, const double*, int64_t, double, int,
                                  unsigned, double*, double*) {
  return ssq_fail(ctx, SSQ_EUNSUPPORTED, "cwt: not built yet");
}
extern "C" ssq_status ssq_ssq_cwt_f64(ssq_ctx* ctx, const double*, int64_t, int, const double*, int64_t, double,
                                      int, int, int, int, double, unsigned, double*, double*) {
  return ssq_fail(ctx, SSQ_EUNSUPPORTED, "ssq_cwt: not built yet");
}

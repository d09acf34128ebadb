// stft_fast.cuh -- register-resident fast path of the fused ssq_stft kernel.
#pragma once
#include "stft_kernels.cuh"

// Tries the specialised kernels; *done=false means "use the generic kernel".
static ssq_status stft_fast_launch(ssq_ctx* ctx, StftParams& P, bool* done) {
  (void)ctx; (void)P;
  *done = false;
  return SSQ_OK;
}

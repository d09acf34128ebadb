// cwt_kernels.cuh -- CWT / ssq_cwt device code (filled in below).
#pragma once
#include "ssq_common.cuh"

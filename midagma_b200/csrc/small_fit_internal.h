// Internal launch hooks shared by the two on-chip fit kernels (small_fit.cu dispatches).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "../../include/dagma_b200.h"

namespace dagma {

// DMMA block-Gauss-Jordan kernel (small_fit_dmma.cu), 32 < d <= 64
int fit_dmma_geometry(int batch, int sms, int* ctas, int* threads, size_t* smem_bytes);
int launch_fit_dmma(cudaStream_t stream, const dagma_small_fit_args& a, int sms);

}  // namespace dagma

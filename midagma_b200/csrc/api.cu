// Library-level entry points of libdagma_b200.so (see include/dagma_b200.h).
#include "common.cuh"
#include "../../include/dagma_b200.h"
#include <vector>

namespace dagma {
thread_local char g_last_error[512] = "";
int logdet_inv_small(cudaStream_t, int, int, double, const double*, int, int, double*, double*,
                     double*, double*, int, double*, int*);
int logdet_inv_large(cudaStream_t, int, int, double, const double*, int, int, double*, double*,
                     double*, double*, int, double*, int*);
}  // namespace dagma

using namespace dagma;

extern "C" int dagma_version(void) { return 100; }

extern "C" const char* dagma_last_error(void) { return g_last_error; }

extern "C" int dagma_device_check(int* sm_count) {
    int dev = 0;
    DAGMA_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DAGMA_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (prop.major != 10)
        return set_error(-2, "libdagma_b200 is built for sm_100a only; device is not compute capability 10.x", prop.name);
    return 0;
}

extern "C" int dagma_logdet_inv_f64(dagma_stream_t stream, int batch, int d, double s,
                                    const double* a_dev, int lda, int square_input,
                                    double* logabsdet_dev, double* h_dev, double* minv_dev,
                                    double* grad_dev, int ldo, double* min_entry_dev, int* info_dev) {
    DAGMA_REQUIRE(batch >= 1 && d >= 1, "batch and d must be positive");
    DAGMA_REQUIRE(a_dev != nullptr, "null input");
    DAGMA_REQUIRE(lda >= d && ((minv_dev == nullptr && grad_dev == nullptr) || ldo >= d), "leading dimension < d");
    if (d <= DAGMA_ONCHIP_INV_MAX_D)
        return logdet_inv_small((cudaStream_t)stream, batch, d, s, a_dev, lda, square_input,
                                logabsdet_dev, h_dev, minv_dev, grad_dev, ldo, min_entry_dev, info_dev);
    return logdet_inv_large((cudaStream_t)stream, batch, d, s, a_dev, lda, square_input,
                            logabsdet_dev, h_dev, minv_dev, grad_dev, ldo, min_entry_dev, info_dev);
}

// Host-buffer wrapper: H2D of cov / W / lambda1, fit, D2H of W and the reports.
extern "C" int dagma_linear_fit_small_host_f64(dagma_stream_t stream_, const dagma_small_fit_args* host) {
    DAGMA_REQUIRE(host != nullptr, "null args");
    cudaStream_t stream = (cudaStream_t)stream_;
    dagma_small_fit_args a = *host;
    const size_t B = (size_t)a.batch, dd = (size_t)a.d * a.d;
    DAGMA_REQUIRE(a.batch >= 1 && a.d >= 1 && a.d <= DAGMA_SMALL_MAX_D, "bad shape");
    const size_t n_mat = B * dd * sizeof(double);
    const size_t n_stats = B * (size_t)(a.n_stages > 0 ? a.n_stages : 1) * 8 * sizeof(double);
    const size_t n_log = B * (size_t)a.ckpt_log_cap * 6 * sizeof(double);
    const size_t n_diag = host->ckpt_diag_dev ? B * (size_t)a.ckpt_log_cap * DAGMA_DIAG_COLS * sizeof(double) : 0;
    size_t total = 2 * n_mat + B * sizeof(double) + 2 * dd + B * sizeof(int32_t) + n_stats +
                   2 * B * sizeof(double) + n_log + n_diag + B * sizeof(int32_t) + 64 + 17 * 256;
    unsigned char* base = nullptr;
    DAGMA_CUDA_OK(cudaMallocAsync((void**)&base, total, stream));
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; };
    double* cov_d = (double*)take(n_mat);
    double* w_d = (double*)take(n_mat);
    double* lam_d = (double*)take(B * sizeof(double));
    uint8_t* exc_d = host->mask_exc_dev ? (uint8_t*)take(dd) : nullptr;
    uint8_t* inc_d = host->mask_inc_dev ? (uint8_t*)take(dd) : nullptr;
    int32_t* status_d = (int32_t*)take(B * sizeof(int32_t));
    double* stats_d = (double*)take(n_stats);
    double* final_d = (double*)take(2 * B * sizeof(double));
    double* log_d = a.ckpt_log_cap > 0 ? (double*)take(n_log) : nullptr;
    double* diag_d = n_diag ? (double*)take(n_diag) : nullptr;
    int32_t* cnt_d = (int32_t*)take(B * sizeof(int32_t));
    uint32_t* ctr_d = (uint32_t*)take(64);
    int rc = 0;
    // outputs of stages / problems the kernel never reaches (out of domain, retry limit, T = 0) must read as zero
    if (cudaMemsetAsync(status_d, 0, (size_t)((base + off) - (unsigned char*)status_d), stream) != cudaSuccess)
        rc = set_error(-3, "memset of the output block failed");
#define H2D(dst, src, n) if (!rc && cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, stream) != cudaSuccess) rc = set_error(-3, "H2D copy failed")
#define D2H(dst, src, n) if (!rc && (dst) && cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, stream) != cudaSuccess) rc = set_error(-3, "D2H copy failed")
    H2D(cov_d, host->cov_dev, n_mat);
    H2D(w_d, host->w_dev, n_mat);
    H2D(lam_d, host->lambda1_dev, B * sizeof(double));
    if (exc_d) H2D(exc_d, host->mask_exc_dev, dd);
    if (inc_d) H2D(inc_d, host->mask_inc_dev, dd);
    a.cov_dev = cov_d; a.w_dev = w_d; a.lambda1_dev = lam_d; a.mask_exc_dev = exc_d; a.mask_inc_dev = inc_d;
    a.status_dev = status_d; a.stage_stats_dev = stats_d; a.final_dev = host->final_dev ? final_d : nullptr;
    a.ckpt_log_dev = log_d; a.ckpt_count_dev = cnt_d; a.work_counter_dev = ctr_d; a.ckpt_diag_dev = diag_d;
    if (!rc) rc = dagma_linear_fit_small_f64(stream, &a);
    D2H(host->w_dev, w_d, n_mat);
    D2H(host->status_dev, status_d, B * sizeof(int32_t));
    D2H(host->stage_stats_dev, stats_d, n_stats);
    D2H(host->final_dev, final_d, 2 * B * sizeof(double));
    if (log_d) D2H(host->ckpt_log_dev, log_d, n_log);
    if (diag_d) D2H(host->ckpt_diag_dev, diag_d, n_diag);
    D2H(host->ckpt_count_dev, cnt_d, B * sizeof(int32_t));
#undef H2D
#undef D2H
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFreeAsync(base, stream);
    if (!rc && e != cudaSuccess) rc = set_error(-100 - (int)e, "fit kernel failed", cudaGetErrorString(e));
    return rc;
}

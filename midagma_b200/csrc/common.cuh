// Shared host-side helpers for libdagma_b200.so (error text, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

namespace dagma {

extern thread_local char g_last_error[512];

inline int set_error(int code, const char* what, const char* detail = "") {
    snprintf(g_last_error, sizeof(g_last_error), "%s%s%s", what, detail[0] ? ": " : "", detail);
    return code;
}

#define DAGMA_CUDA_OK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return ::dagma::set_error(-100 - (int)_e, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define DAGMA_REQUIRE(cond, msg)                                                         \
    do {                                                                                 \
        if (!(cond)) return ::dagma::set_error(-1, msg, #cond);                          \
    } while (0)

}  // namespace dagma

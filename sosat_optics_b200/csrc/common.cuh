// Shared host-side helpers for libsosat_b200.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "sosat_b200.h"

namespace sosat {

// thread-local message returned by sosat_last_error()
char *error_buffer();
int set_error(int code, const char *fmt, ...);

#define SOSAT_CUDA_OK(expr)                                                              \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess)                                                        \
            return ::sosat::set_error(SOSAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,    \
                                      cudaGetErrorString(err__), __FILE__, __LINE__);    \
    } while (0)

#define SOSAT_REQUIRE(cond, ...)                                                         \
    do {                                                                                 \
        if (!(cond)) return ::sosat::set_error(SOSAT_ERR_INVALID_ARG, __VA_ARGS__);      \
    } while (0)

inline int num_sms() {
    static int cache[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    if (n > 0 && dev >= 0 && dev < 64) cache[dev] = n;
    return n > 0 ? n : 148;
}

}  // namespace sosat

// common.cuh -- shared helpers for libl3c_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/l3c_b200.h"

namespace l3c {

void set_error(const char *fmt, ...);

#define L3C_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            ::l3c::set_error(__VA_ARGS__);          \
            return L3C_EINVAL;                      \
        }                                           \
    } while (0)

#define L3C_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::l3c::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                             __FILE__, __LINE__);                                        \
            return L3C_ECUDA;                                                            \
        }                                                                                \
    } while (0)

#define L3C_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            ::l3c::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));  \
            return L3C_ECUDA;                                                            \
        }                                                                                \
        ::l3c::count_launch(name);                                                       \
    } while (0)

void count_launch(const char *kernel_name);   // per-kernel launch counters (l3c_launch_log)

int sm_count();                          // SMs of the current device
int stream_sm_count(cudaStream_t st);    // SMs available to kernels launched into `st` (partition.cu)

// cudaFuncSetAttribute is per DEVICE: "done once" flags must be kept per device, or a process that drives a second
// GPU launches its kernels there without the opt-in for > 48 KB of dynamic shared memory
static inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace l3c

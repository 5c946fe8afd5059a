// Shared helpers for the promonet_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/promonet_b200.h"

namespace pmn {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launch_count;

inline int fail(int status, const std::string& message) {
    g_last_error = message;
    return status;
}

inline int check_cuda(cudaError_t error, const char* what) {
    if (error == cudaSuccess) return PMN_OK;
    return fail(PMN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(error));
}

// Optional per-kernel device timing (pmn_profile_*): when enabled, every launch
// is bracketed by CUDA events on its own stream; totals are read back per kernel.
void profile_before(const char* kernel, cudaStream_t stream);
void profile_after(const char* kernel, cudaStream_t stream);
extern bool g_profile_enabled;

// Declare before a kernel launch; counts it, times it when profiling is on
struct LaunchScope {
    const char* kernel;
    cudaStream_t stream;
    LaunchScope(const char* kernel_, cudaStream_t stream_) : kernel(kernel_), stream(stream_) {
        if (g_profile_enabled) profile_before(kernel, stream);
    }
    ~LaunchScope() {
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        if (g_profile_enabled) profile_after(kernel, stream);
    }
};

// Call after every kernel launch: surfaces launch errors
inline int launched(const char* kernel) {
    return check_cuda(cudaGetLastError(), kernel);
}

#define PMN_TRY(expression)                     \
    do {                                        \
        int pmn_status_ = (expression);         \
        if (pmn_status_ != PMN_OK) return pmn_status_; \
    } while (0)

#define PMN_REQUIRE(condition, message)                       \
    do {                                                      \
        if (!(condition)) return pmn::fail(PMN_ERR_ARGUMENT, message); \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float leaky(float x, float slope) {
    return x > 0.f ? x : x * slope;
}

// ---------------------------------------------------------------------------
// Kernel launchers shared between the C ABI and the generator orchestration
// ---------------------------------------------------------------------------

struct Conv1dArgs {
    const float* x = nullptr;        // (B, C_in, T_in)
    const float* weight = nullptr;   // packed (C_in, K, C_out)
    const float* bias = nullptr;     // (C_out) or null
    const float* bias2 = nullptr;    // (B, C_out) or null
    const float* residual = nullptr; // (B, C_out, T_out) or null
    float* out = nullptr;            // (B, C_out, T_out) or null
    float* accum = nullptr;          // (B, C_out, T_out) or null
    int accum_mode = 0;              // 0 unused, 1 store, 2 add
    float accum_scale = 1.f;
    int batch = 0, c_in = 0, c_out = 0, t_in = 0, t_out = 0;
    int k = 1, dilation = 1, padding = 0;
    float in_slope = 1.f;
    int out_act = 0;                 // 0 none, 1 tanh, 2 relu
};

int launch_conv1d(const Conv1dArgs& args, cudaStream_t stream);

int launch_conv_transpose1d(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope,
    cudaStream_t stream);

int launch_weight_norm_fold(
    const float* v, const float* g, float* w, int dim0, int inner, cudaStream_t stream);

int launch_pack_conv1d_weight(
    const float* w, float* packed, int c_out, int c_in, int k, cudaStream_t stream);

}  // namespace pmn

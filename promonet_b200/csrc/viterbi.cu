// Viterbi decoding (torbi.from_probabilities as called at
// promonet/preprocess/harmonics.py:270-276 and inside penn's 'viterbi' decoder,
// promonet/preprocess/core.py:64-81).
//
//   delta_0[j] = log pi[j] + log o_0[j]
//   delta_t[j] = max_i (delta_{t-1}[i] + log A[i, j]) + log o_t[j],  psi_t[j] = argmax_i (lowest i on ties)
//   path: backtrace from argmax_j delta_{T-1}[j]
//
// The recurrence is sequential in t, so one CTA walks one utterance with delta
// double-buffered in shared memory.  Pitch transition matrices are banded, so a
// prepass stores log A column by column as band[k][j] = log A[lo_j + k, j] for
// the rows lo_j .. hi_j that are not -inf: thread j then reads band[k * S + j]
// (coalesced over j) and delta[lo_j + k] (conflict-free).  A dense matrix is the
// same code with lo = 0.  fp32 add/compare only after the logs, so the indices
// are bit-exact against the CPU recurrence on the same log inputs.
#include <math.h>

#include "spectral.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 1024;

// Per column j: first and one-past-last row with a finite log-probability
__global__ void band_range_kernel(
    const float* __restrict__ transition, bool log_probs, int states,
    int* __restrict__ lo, int* __restrict__ width, int* __restrict__ max_width) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= states) return;
    int first = states, last = 0;
    for (int i = 0; i < states; ++i) {
        const float value = transition[(size_t)i * states + j];
        const bool live = log_probs ? value > -INFINITY : value > 0.f;
        if (live) { first = min(first, i); last = i + 1; }
    }
    if (first >= last) { first = 0; last = 1; }
    lo[j] = first;
    width[j] = last - first;
    atomicMax(max_width, last - first);
}

__global__ void band_fill_kernel(
    const float* __restrict__ transition, bool log_probs, int states,
    const int* __restrict__ lo, const int* __restrict__ width, const int* __restrict__ max_width,
    float* __restrict__ band) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (j >= states || k >= *max_width) return;
    float value = -INFINITY;
    if (k < width[j]) {
        value = transition[(size_t)(lo[j] + k) * states + j];
        if (!log_probs) value = logf(value);
    }
    band[(size_t)k * states + j] = value;
}

__global__ void __launch_bounds__(kThreads) viterbi_kernel(
    const float* __restrict__ observation, const int* __restrict__ batch_frames,
    const float* __restrict__ initial, bool log_probs,
    const float* __restrict__ band, const int* __restrict__ lo, const int* __restrict__ width,
    short* __restrict__ psi, int* __restrict__ indices, int frames, int states) {
    extern __shared__ float delta[];  // [2][states]
    __shared__ float best_value[32];
    __shared__ int best_index[32];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int length = batch_frames ? min(batch_frames[b], frames) : frames;
    const float* obs = observation + (size_t)b * frames * states;
    short* back = psi + (size_t)b * frames * states;
    int* path = indices + (size_t)b * frames;
    if (length <= 0) {
        for (int t = tid; t < frames; t += kThreads) path[t] = 0;
        return;
    }

    for (int j = tid; j < states; j += kThreads) {
        const float o = log_probs ? obs[j] : logf(obs[j]);
        const float p = log_probs ? initial[j] : logf(initial[j]);
        delta[j] = p + o;
    }
    __syncthreads();

    int current = 0;
    for (int t = 1; t < length; ++t) {
        const float* previous = delta + current * states;
        float* next = delta + (current ^ 1) * states;
        const float* row = obs + (size_t)t * states;
        for (int j = tid; j < states; j += kThreads) {
            const int first = lo[j], count = width[j];
            float best = -INFINITY;
            int arg = 0;
            const float* column = band + j;
            const float* source = previous + first;
            for (int k = 0; k < count; ++k) {
                const float value = source[k] + column[(size_t)k * states];
                if (value > best) { best = value; arg = first + k; }
            }
            const float o = log_probs ? row[j] : logf(row[j]);
            next[j] = best + o;
            back[(size_t)t * states + j] = (short)arg;
        }
        current ^= 1;
        __syncthreads();
    }

    // argmax over the final scores, lowest index on ties
    const float* final_scores = delta + current * states;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = tid; j < states; j += kThreads) {
        const float value = final_scores[j];
        if (value > best) { best = value; arg = j; }
    }
    if (arg == 0x7fffffff) arg = tid < states ? tid : 0x7fffffff;  // all -inf: candidates by index
    for (int offset = 16; offset > 0; offset >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, best, offset);
        const int other_arg = __shfl_xor_sync(0xffffffffu, arg, offset);
        if (other > best || (other == best && other_arg < arg)) { best = other; arg = other_arg; }
    }
    if ((tid & 31) == 0) { best_value[tid >> 5] = best; best_index[tid >> 5] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            if (best_value[w] > best || (best_value[w] == best && best_index[w] < arg)) {
                best = best_value[w];
                arg = best_index[w];
            }
        }
        int state = arg == 0x7fffffff ? 0 : arg;
        if (!(best > -INFINITY)) state = 0;  // CPU recurrence keeps index 0 when nothing beats -inf
        for (int t = length - 1; t >= 0; --t) {
            path[t] = state;
            if (t > 0) state = back[(size_t)t * states + state];
        }
        for (int t = length; t < frames; ++t) path[t] = 0;
    }
}

struct Workspace {
    float* band;
    int *lo, *width, *max_width;
    short* psi;
    size_t bytes;
};

Workspace carve(void* base, int batch, int frames, int states) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += align_up(bytes, 256);
        return r;
    };
    w.band = (float*)take((size_t)states * states * sizeof(float));
    w.lo = (int*)take((size_t)states * sizeof(int));
    w.width = (int*)take((size_t)states * sizeof(int));
    w.max_width = (int*)take(sizeof(int));
    w.psi = (short*)take((size_t)batch * frames * states * sizeof(short));
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}

}  // namespace

size_t viterbi_workspace_bytes(int batch, int frames, int states) {
    return carve(nullptr, batch, frames, states).bytes;
}

int launch_viterbi(
    const float* observation, const int* batch_frames, const float* transition,
    const float* initial, bool log_probs, int* indices, int batch, int frames, int states,
    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    PMN_REQUIRE(observation && transition && initial && indices && workspace, "viterbi: null pointer");
    PMN_REQUIRE(batch > 0 && frames > 0 && states > 0, "viterbi: empty input");
    PMN_REQUIRE(states <= 32767, "viterbi: more than 32767 states");
    PMN_REQUIRE((size_t)2 * states * sizeof(float) <= 200 * 1024, "viterbi: too many states for shared memory");
    Workspace w = carve(workspace, batch, frames, states);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "viterbi: workspace too small");
    PMN_TRY(check_cuda(cudaMemsetAsync(w.max_width, 0, sizeof(int), stream), "viterbi memset"));
    {
        LaunchScope scope("band_range_kernel", stream);
        band_range_kernel<<<ceil_div(states, 128), 128, 0, stream>>>(
            transition, log_probs, states, w.lo, w.width, w.max_width);
        PMN_TRY(launched("band_range_kernel"));
    }
    {
        dim3 grid(ceil_div(states, 128), states);
        LaunchScope scope("band_fill_kernel", stream);
        band_fill_kernel<<<grid, 128, 0, stream>>>(
            transition, log_probs, states, w.lo, w.width, w.max_width, w.band);
        PMN_TRY(launched("band_fill_kernel"));
    }
    const size_t smem = (size_t)2 * states * sizeof(float);
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
            "viterbi smem attribute"));
        configured = true;
    }
    LaunchScope scope("viterbi_kernel", stream);
    viterbi_kernel<<<batch, kThreads, smem, stream>>>(
        observation, batch_frames, initial, log_probs, w.band, w.lo, w.width, w.psi, indices,
        frames, states);
    return launched("viterbi_kernel");
}

}  // namespace pmn

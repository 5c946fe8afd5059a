/* Oracle: Viterbi decoding (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * Restates torbi.from_probabilities as witnessed at
 * promonet/preprocess/harmonics.py:270-276 (torbi itself is an un-vendored
 * dependency, setup.py:14-36 / README.md:64-66: PARITY UNPINNED):
 *   delta_0[j] = log pi[j] + log o_0[j]
 *   delta_t[j] = max_i (delta_{t-1}[i] + log A[i, j]) + log o_t[j],  psi_t[j] = argmax_i
 *   path by backtrace from argmax_j delta_{T-1}[j]
 * Ties resolve to the lowest index.  All arithmetic in fp32, the operation
 * order (delta + logA, then + log o) is the one the CUDA kernel uses so that
 * the indices can be compared bit-exactly.
 *
 * Inputs are LOG probabilities (the caller takes logs when log_probs=False).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int viterbi_oracle(
    const float* observation,   /* (batch, frames, states) log probabilities */
    const int32_t* batch_frames, /* (batch) valid lengths or NULL */
    const float* transition,    /* (states, states) log, row i -> column j */
    const float* initial,       /* (states) log */
    int32_t* indices,           /* (batch, frames) out */
    int batch, int frames, int states) {
    /* Column j of the transition, contiguous, and the range [lo, hi) of its rows that are
     * not -inf: a -inf candidate never wins the strict comparison below, so skipping the
     * rows outside the range changes neither the maximum nor the (lowest-index) argmax.
     * Only the speed of the checker depends on this (penn's band: 181 of 1440 rows). */
    float* columns = (float*)malloc(sizeof(float) * (size_t)states * states);
    int* lo = (int*)malloc(sizeof(int) * states);
    int* hi = (int*)malloc(sizeof(int) * states);
    if (!columns || !lo || !hi) { free(columns); free(lo); free(hi); return -1; }
    for (int j = 0; j < states; ++j) {
        lo[j] = states; hi[j] = 0;
        for (int i = 0; i < states; ++i) {
            const float value = transition[(size_t)i * states + j];
            columns[(size_t)j * states + i] = value;
            if (value > -INFINITY) { if (i < lo[j]) lo[j] = i; hi[j] = i + 1; }
        }
    }
    int failed = 0;
    for (int b = 0; b < batch; ++b) {
        float* delta = (float*)malloc(sizeof(float) * 2 * states);
        int32_t* psi = (int32_t*)malloc(sizeof(int32_t) * (size_t)frames * states);
        if (!delta || !psi) { free(delta); free(psi); failed = 1; continue; }
        const float* obs = observation + (size_t)b * frames * states;
        const int length = batch_frames ? batch_frames[b] : frames;
        float* previous = delta;
        float* current = delta + states;
        for (int j = 0; j < states; ++j) previous[j] = initial[j] + obs[j];
        for (int t = 1; t < length; ++t) {
            for (int j = 0; j < states; ++j) {
                float best = -INFINITY;
                int32_t arg = 0;
                const float* column = columns + (size_t)j * states;
                for (int i = lo[j]; i < hi[j]; ++i) {
                    const float value = previous[i] + column[i];
                    if (value > best) { best = value; arg = i; }
                }
                current[j] = best + obs[(size_t)t * states + j];
                psi[(size_t)t * states + j] = arg;
            }
            float* swap = previous; previous = current; current = swap;
        }
        int32_t state = 0;
        float best = -INFINITY;
        for (int j = 0; j < states; ++j)
            if (previous[j] > best) { best = previous[j]; state = j; }
        for (int t = length - 1; t >= 0; --t) {
            indices[(size_t)b * frames + t] = state;
            if (t > 0) state = psi[(size_t)t * states + state];
        }
        for (int t = length; t < frames; ++t) indices[(size_t)b * frames + t] = 0;
        free(delta);
        free(psi);
    }
    free(columns); free(lo); free(hi);
    return failed ? -1 : 0;
}

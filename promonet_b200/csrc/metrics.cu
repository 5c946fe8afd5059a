// In-training validation: the prosody / pronunciation metrics and the contour edits around them.
//
// promonet.evaluate.Metrics.update (promonet/evaluate/metrics.py:38-61) is, per utterance, ~40
// small ATen kernels and four host synchronisations (boolean-mask indexing); train.evaluate calls
// it 7 times per validation item (promonet/train/core.py:606-799).  Here one pass over the frames
// adds every running sum of every metric into 12 doubles on the device; nothing comes back to the
// host until Metrics.__call__.  promonet.edit.from_features (promonet/edit/core.py:17-132) for the
// three contours (pitch / periodicity / loudness) is one kernel too.  Both are HBM-bound and tiny.
#include "common.cuh"

namespace pmn {

namespace {

constexpr int kMaxPpg = 64;
constexpr int kThreads = 128;

// ppgs.sparsify(p, 'percentile', q) over the C channels of one frame, in place:
// cutoff = torch.quantile (linear interpolation at rank q (C - 1)), keep p > cutoff,
// renormalise as softmax(log(p + 1e-8)).  Same arithmetic as features_kernel (features.cu).
template <int C>
__device__ __forceinline__ void sparsify(float (&p)[C], float q) {
    const float rank = q * (float)(C - 1);
    const int lower = (int)floorf(rank);
    const int upper = min(lower + 1, C - 1);
    const float weight = rank - (float)lower;
    float below = 0.f, above = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        int less = 0, less_equal = 0;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            less += p[j] < p[i];
            less_equal += p[j] <= p[i];
        }
        if (less <= lower && lower < less_equal) below = p[i];
        if (less <= upper && upper < less_equal) above = p[i];
    }
    const float diff = above - below;
    const float cutoff = weight < 0.5f ? below + weight * diff : above - diff * (1.f - weight);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        p[i] = (p[i] > cutoff ? p[i] : 0.f) + 1e-8f;
        total += p[i];
    }
    const float inv = 1.f / total;
#pragma unroll
    for (int i = 0; i < C; ++i) p[i] *= inv;
}

// Jensen-Shannon distance of one frame (ppgs.distance, un-vendored, restated): clamp to
// [1e-8, 1 - 1e-8], optionally p <- S^T p with the phoneme-similarity matrix S,
// m = (p + q) / 2, sqrt((KL(p || m) + KL(q || m)) / 2)
template <int C>
__device__ __forceinline__ float js_distance(
    float (&p)[C], float (&q)[C], const float* __restrict__ similarity /* shared, (C, C) */) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
        p[i] = fminf(fmaxf(p[i], 1e-8f), 1.f - 1e-8f);
        q[i] = fminf(fmaxf(q[i], 1e-8f), 1.f - 1e-8f);
    }
    float divergence = 0.f;
    if (similarity) {
        for (int i = 0; i < C; ++i) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const float s = similarity[j * C + i];
                a = fmaf(s, p[j], a);
                b = fmaf(s, q[j], b);
            }
            const float m = 0.5f * (a + b);
            divergence += a * (logf(a) - logf(m)) + b * (logf(b) - logf(m));
        }
    } else {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float m = 0.5f * (p[i] + q[i]);
            const float lm = logf(m);
            divergence += p[i] * (logf(p[i]) - lm) + q[i] * (logf(q[i]) - lm);
        }
    }
    return sqrtf(fmaxf(0.5f * divergence, 0.f));
}

// One thread per (item, frame); block sums in shared memory; one double atomic per slot and block.
template <int C>
__global__ void __launch_bounds__(kThreads) metrics_update_kernel(
    const float* __restrict__ predicted_loudness, int predicted_bands,
    const float* __restrict__ target_loudness, int target_bands,
    const float* __restrict__ predicted_pitch, const float* __restrict__ target_pitch,
    const float* __restrict__ predicted_periodicity, const float* __restrict__ target_periodicity,
    const float* __restrict__ predicted_ppg, const float* __restrict__ target_ppg,
    const float* __restrict__ similarity,
    int frames, float loudness_threshold, float voicing_threshold, float ppg_threshold,
    double* __restrict__ sums) {
    __shared__ float s_similarity[C > 0 ? C * C : 1];
    __shared__ float partial[PMN_METRICS_SLOTS][kThreads / 32];
    const bool with_ppg = C > 0 && predicted_ppg && target_ppg;
    const bool with_similarity = with_ppg && similarity;
    if (with_similarity) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_similarity[i] = similarity[i];
        __syncthreads();
    }
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t item = blockIdx.y;
    float v[PMN_METRICS_SLOTS];
#pragma unroll
    for (int i = 0; i < PMN_METRICS_SLOTS; ++i) v[i] = 0.f;
    if (f < frames) {
        // Loudness (metrics.py:185-204): mean over the rows given, then squared error overall and
        // split by whether both contours are above the threshold
        if (predicted_loudness && target_loudness) {
            // up to 513 rows per side: summed in double so the mean is the correctly rounded one
            double sum_a = 0., sum_b = 0.;
            const float* pl = predicted_loudness + item * predicted_bands * frames + f;
            const float* tl = target_loudness + item * target_bands * frames + f;
            for (int r = 0; r < predicted_bands; ++r) sum_a += (double)__ldg(pl + (size_t)r * frames);
            for (int r = 0; r < target_bands; ++r) sum_b += (double)__ldg(tl + (size_t)r * frames);
            const float a = (float)(sum_a / predicted_bands);
            const float b = (float)(sum_b / target_bands);
            const float d = a - b;
            const bool loud = a > loudness_threshold && b > loudness_threshold;
            v[0] = d * d; v[1] = 1.f;
            v[2] = loud ? d * d : 0.f; v[3] = loud ? 1.f : 0.f;
            v[4] = loud ? 0.f : d * d; v[5] = loud ? 0.f : 1.f;
        }
        // Periodicity RMSE (metrics.py:20,55) and voiced pitch error in log2 (metrics.py:249-261)
        if (predicted_periodicity && target_periodicity) {
            const float a = __ldg(predicted_periodicity + item * frames + f);
            const float b = __ldg(target_periodicity + item * frames + f);
            v[6] = (a - b) * (a - b); v[7] = 1.f;
            if (predicted_pitch && target_pitch && a > voicing_threshold && b > voicing_threshold) {
                v[8] = fabsf(
                    log2f(__ldg(predicted_pitch + item * frames + f)) -
                    log2f(__ldg(target_pitch + item * frames + f)));
                v[9] = 1.f;
            }
        }
        // PPG distance (metrics.py:287-312)
        if constexpr (C > 0) {
            if (with_ppg) {
                float p[C], q[C];
                const float* pp = predicted_ppg + item * C * frames + f;
                const float* tp = target_ppg + item * C * frames + f;
#pragma unroll
                for (int i = 0; i < C; ++i) {
                    p[i] = __ldg(pp + (size_t)i * frames);
                    q[i] = __ldg(tp + (size_t)i * frames);
                }
                sparsify<C>(p, ppg_threshold);
                sparsify<C>(q, ppg_threshold);
                v[10] = js_distance<C>(p, q, with_similarity ? s_similarity : nullptr);
                v[11] = 1.f;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < PMN_METRICS_SLOTS; ++i) {
        float x = v[i];
#pragma unroll
        for (int offset = 16; offset > 0; offset >>= 1) x += __shfl_xor_sync(0xffffffffu, x, offset);
        if (lane == 0) partial[i][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < PMN_METRICS_SLOTS) {
        double total = 0.;
        for (int w = 0; w < kThreads / 32; ++w) total += (double)partial[threadIdx.x][w];
        if (total != 0.) atomicAdd(sums + threadIdx.x, total);
    }
}

// promonet.edit.from_features for one contour (edit/core.py:113-128).  The contour (items, t_in)
// is resampled at `grid` (edit/grid.py:12-43: linear, final frame replicated; grid == NULL keeps
// the frames), in the log2 domain when log2_domain (pitch: 2 ** sample(log2(pitch), grid)); then
// out = clip(scale * value + shift, lo, hi): scale = pitch shift, shift = loudness in dB, and the
// clip (pitch shift only) applies when lo < hi.
__global__ void __launch_bounds__(kThreads) edit_contour_kernel(
    const float* __restrict__ sequence, const float* __restrict__ grid, float* __restrict__ out,
    int t_in, int t_out, int log2_domain, float scale, float shift, float lo, float hi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_out) return;
    const float* row = sequence + (size_t)blockIdx.y * t_in;
    float value;
    if (grid) {
        const float x = grid[t];
        int i = x < 0.f ? 0 : min((int)floorf(x) + 1, t_in);
        i = max(i, 1);
        const int lower = i - 1, upper = min(i, t_in - 1);
        const float wl = (float)i - x, wu = x - (float)(i - 1);
        float a = row[lower], b = row[upper];
        if (log2_domain) { a = log2f(a); b = log2f(b); }
        value = a * wl + b * wu;
        if (log2_domain) value = exp2f(value);
    } else {
        value = row[t];
    }
    value = value * scale + shift;
    if (lo < hi) value = fminf(fmaxf(value, lo), hi);
    out[(size_t)blockIdx.y * t_out + t] = value;
}

}  // namespace

}  // namespace pmn

using namespace pmn;

extern "C" {

int pmn_metrics_update(
    const float* predicted_loudness, int predicted_bands,
    const float* target_loudness, int target_bands,
    const float* predicted_pitch, const float* target_pitch,
    const float* predicted_periodicity, const float* target_periodicity,
    const float* predicted_ppg, const float* target_ppg, int ppg_channels,
    const float* similarity, int items, int frames,
    float loudness_threshold, float voicing_threshold, float ppg_threshold,
    double* sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMN_REQUIRE(sums, "metrics_update: null accumulator");
    PMN_REQUIRE(items > 0 && items <= 65535, "metrics_update: bad item count");
    PMN_REQUIRE((predicted_loudness == nullptr) == (target_loudness == nullptr) &&
                (predicted_pitch == nullptr) == (target_pitch == nullptr) &&
                (predicted_periodicity == nullptr) == (target_periodicity == nullptr) &&
                (predicted_ppg == nullptr) == (target_ppg == nullptr),
                "metrics_update: predicted and target must be given together");
    PMN_REQUIRE(!predicted_loudness || (predicted_bands > 0 && target_bands > 0),
                "metrics_update: loudness needs at least one row");
    PMN_REQUIRE(!predicted_pitch || predicted_periodicity,
                "metrics_update: pitch error needs the periodicity for the voicing decision");
    PMN_REQUIRE(!predicted_ppg || ppg_channels == 40,
                "metrics_update: PPGs have 40 channels (PPG_CHANNELS, config/defaults.py:102)");
    PMN_REQUIRE(ppg_channels <= kMaxPpg, "metrics_update: too many PPG channels");
    if (frames <= 0) return PMN_OK;
    dim3 blocks(ceil_div(frames, kThreads), items);
    LaunchScope scope("metrics_update_kernel", stream);
    if (predicted_ppg)
        metrics_update_kernel<40><<<blocks, kThreads, 0, stream>>>(
            predicted_loudness, predicted_bands, target_loudness, target_bands, predicted_pitch,
            target_pitch, predicted_periodicity, target_periodicity, predicted_ppg, target_ppg,
            similarity, frames, loudness_threshold, voicing_threshold, ppg_threshold, sums);
    else
        metrics_update_kernel<0><<<blocks, kThreads, 0, stream>>>(
            predicted_loudness, predicted_bands, target_loudness, target_bands, predicted_pitch,
            target_pitch, predicted_periodicity, target_periodicity, nullptr, nullptr,
            nullptr, frames, loudness_threshold, voicing_threshold, ppg_threshold, sums);
    return launched("metrics_update_kernel");
}

int pmn_edit_contour(
    const float* sequence, const float* grid, float* out, int items, int t_in, int t_out,
    int log2_domain, float scale, float shift, float lo, float hi, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMN_REQUIRE(sequence && out, "edit_contour: null pointer");
    PMN_REQUIRE(items > 0 && items <= 65535 && t_in > 0, "edit_contour: bad shape");
    PMN_REQUIRE(grid || t_out == t_in, "edit_contour: without a grid the length cannot change");
    if (t_out <= 0) return PMN_OK;
    dim3 blocks(ceil_div(t_out, kThreads), items);
    LaunchScope scope("edit_contour_kernel", stream);
    edit_contour_kernel<<<blocks, kThreads, 0, stream>>>(
        sequence, grid, out, t_in, t_out, log2_domain, scale, shift, lo, hi);
    return launched("edit_contour_kernel");
}

}  // extern "C"

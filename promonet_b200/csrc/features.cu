// Feature assembly, speaker projection, weight preparation and output head.
//
// All HBM-bound, one pass each: these are the rows G1, G2, G7, G8 of the scope
// table (Generator.prepare_features promonet/model/generator.py:137-197,
// prepare_global_features :49-70, output head hifigan.py:56-60, weight-norm
// parametrisation model/core.py:43-45).
#include "features.cuh"

namespace pmn {

namespace {

constexpr int kPpg = 40;          // PPG_CHANNELS, config/defaults.py:102
constexpr int kPitchBins = 256;   // PITCH_BINS :96
constexpr int kPitchEmbed = 64;   // PITCH_EMBEDDING_SIZE :99
constexpr int kBands = 8;         // LOUDNESS_BANDS :90
constexpr float kFmin = 50.f, kFmax = 550.f;  // :27-28
constexpr float kMinDb = -100.f, kRefDb = 20.f;  // :37, :46

// One thread per (b, f).  Channel order of the output (generator.py:149-188):
// ppg[0:40] . pitch embedding[40:104] . loudness[104:112] . periodicity[112] (. period[113])
__global__ void __launch_bounds__(128) features_kernel(
    const float* __restrict__ loudness, int rows,
    const float* __restrict__ pitch, const float* __restrict__ periodicity,
    const float* __restrict__ ppg,
    const float* __restrict__ pitch_distribution,  // (256) sorted bin edges
    const float* __restrict__ pitch_embedding,     // (256, 64)
    float threshold, int with_period,
    float* __restrict__ out, int frames) {
    __shared__ float edges[kPitchBins];
    for (int i = threadIdx.x; i < kPitchBins; i += blockDim.x) edges[i] = pitch_distribution[i];
    __syncthreads();

    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (f >= frames) return;
    const int channels = kPpg + kPitchEmbed + kBands + 1 + (with_period ? 1 : 0);
    float* o = out + (size_t)b * channels * frames + f;

    // --- ppgs.sparsify(ppg, 'percentile', threshold): generator.py:140-147 ---
    // torch.quantile, linear interpolation at rank q*(n-1), then keep p > cutoff
    // and renormalise as softmax(log(p + 1e-8)) = (p + 1e-8) / sum(p + 1e-8).
    float p[kPpg];
    const float* pp = ppg + (size_t)b * kPpg * frames + f;
#pragma unroll
    for (int i = 0; i < kPpg; ++i) p[i] = __ldg(pp + (size_t)i * frames);
    const float rank = threshold * (float)(kPpg - 1);
    const int lower = (int)floorf(rank);
    const int upper = min(lower + 1, kPpg - 1);
    const float weight = rank - (float)lower;
    float below = 0.f, above = 0.f;
#pragma unroll
    for (int i = 0; i < kPpg; ++i) {
        int less = 0, less_equal = 0;
#pragma unroll
        for (int j = 0; j < kPpg; ++j) {
            less += p[j] < p[i];
            less_equal += p[j] <= p[i];
        }
        if (less <= lower && lower < less_equal) below = p[i];
        if (less <= upper && upper < less_equal) above = p[i];
    }
    // at::lerp: weight < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
    const float diff = above - below;
    const float cutoff = weight < 0.5f ? below + weight * diff : above - diff * (1.f - weight);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < kPpg; ++i) {
        p[i] = (p[i] > cutoff ? p[i] : 0.f) + 1e-8f;
        total += p[i];
    }
    const float inv = 1.f / total;
#pragma unroll
    for (int i = 0; i < kPpg; ++i) o[(size_t)i * frames] = p[i] * inv;

    // --- pitch: clip, searchsorted(side=left), clip, embedding: :153-164 ---
    const float raw = __ldg(pitch + (size_t)b * frames + f);
    const float hz = fminf(fmaxf(raw, kFmin), kFmax);
    int lo = 0, hi = kPitchBins;  // first index with edges[idx] >= hz
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] < hz) lo = mid + 1; else hi = mid;
    }
    const int bin = min(lo, kPitchBins - 1);
    const float* e = pitch_embedding + (size_t)bin * kPitchEmbed;
#pragma unroll 8
    for (int i = 0; i < kPitchEmbed; ++i) o[(size_t)(kPpg + i) * frames] = __ldg(e + i);

    // --- loudness: band means + normalise: :172-184, loudness.py:144-146 ---
    const float* lp = loudness + (size_t)b * rows * frames + f;
    const float step = (float)rows / (float)kBands;
    for (int band = 0; band < kBands; ++band) {
        // Python: int(band * step) with step a double; rows/8 is exact in fp32
        // for rows < 2^24 / 8, and band * step is then exact too
        const int start = (int)(band * step);
        const int stop = (int)((band + 1) * step);
        float sum = 0.f;
        for (int r = start; r < stop; ++r) sum += __ldg(lp + (size_t)r * frames);
        const float mean = sum / (float)(stop - start);
        o[(size_t)(kPpg + kPitchEmbed + band) * frames] = (mean - kMinDb) / (kRefDb - kMinDb);
    }

    // --- periodicity (:187-188) and FARGAN period (:191-195) ---
    o[(size_t)(kPpg + kPitchEmbed + kBands) * frames] = __ldg(periodicity + (size_t)b * frames + f);
    if (with_period)
        o[(size_t)(kPpg + kPitchEmbed + kBands + 1) * frames] = 22050.f / hz;
}

// bias2[b, o] = conv_bias[o] + sum_c W[o, c] * g[b, c], g = [speaker_embedding[spk], sbr, lr]
// (generator.py:56-68 then input_speaker_conv hifigan.py:68)
__global__ void speaker_bias_kernel(
    const float* __restrict__ speaker_embedding, const int64_t* __restrict__ speakers,
    const float* __restrict__ sbr, const float* __restrict__ lr,
    const float* __restrict__ weight,  // (C_out, C_g)
    const float* __restrict__ bias, float* __restrict__ out,
    int speaker_channels, int c_out, int num_speakers) {
    extern __shared__ float g[];
    const int b = blockIdx.x;
    const int c_g = speaker_channels + 2;
    int64_t speaker = speakers[b];
    speaker = speaker < 0 ? 0 : (speaker >= num_speakers ? num_speakers - 1 : speaker);
    for (int i = threadIdx.x; i < speaker_channels; i += blockDim.x)
        g[i] = speaker_embedding[(size_t)speaker * speaker_channels + i];
    if (threadIdx.x == 0) {
        g[speaker_channels] = sbr[b];
        g[speaker_channels + 1] = lr[b];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < c_out; o += blockDim.x) {
        float acc = bias ? bias[o] : 0.f;
        const float* w = weight + (size_t)o * c_g;
        for (int c = 0; c < c_g; ++c) acc = fmaf(w[c], g[c], acc);
        out[(size_t)b * c_out + o] = acc;
    }
}

// tanh(Conv1d(C -> 1, k7, pad 3, no bias)(lrelu(x))): hifigan.py:56-60.
// Each thread makes 4 consecutive samples from a 12-sample aligned window.
template <int C>
__global__ void __launch_bounds__(256) head_kernel(
    const float* __restrict__ x, const float* __restrict__ weight,  // (1, C, 7)
    float* __restrict__ out, int t_len, float slope) {
    __shared__ float w[C * 7];
    for (int i = threadIdx.x; i < C * 7; i += blockDim.x) w[i] = weight[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= t_len) return;
    const float* xb = x + (size_t)b * C * t_len;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool interior = t >= 4 && t + 8 <= t_len && (t_len % 4 == 0);
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        const float* row = xb + (size_t)c * t_len;
        float v[12];
        if (interior) {
            const float4 a = *reinterpret_cast<const float4*>(row + t - 4);
            const float4 m = *reinterpret_cast<const float4*>(row + t);
            const float4 z = *reinterpret_cast<const float4*>(row + t + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = m.x; v[5] = m.y; v[6] = m.z; v[7] = m.w;
            v[8] = z.x; v[9] = z.y; v[10] = z.z; v[11] = z.w;
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int u = t - 4 + i;
                v[i] = (u >= 0 && u < t_len) ? row[u] : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = leaky(v[i], slope);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float wj = w[c * 7 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(wj, v[i + j + 1], acc[i]);
        }
    }
    float* ob = out + (size_t)b * t_len;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (t + i < t_len) ob[t + i] = tanhf(acc[i]);
}

// w[d0, :] = g[d0] * v[d0, :] / ||v[d0, :]||  (torch.nn.utils.weight_norm, dim=0)
__global__ void weight_norm_fold_kernel(
    const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int inner) {
    __shared__ float partial[32];
    const float* row = v + (size_t)blockIdx.x * inner;
    float sum = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) sum = fmaf(row[i], row[i], sum);
    for (int offset = 16; offset > 0; offset >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, offset);
    if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float s = threadIdx.x < (blockDim.x >> 5) ? partial[threadIdx.x] : 0.f;
        for (int offset = 16; offset > 0; offset >>= 1) s += __shfl_xor_sync(0xffffffffu, s, offset);
        if (threadIdx.x == 0) partial[0] = s;
    }
    __syncthreads();
    const float scale = g[blockIdx.x] / sqrtf(partial[0]);
    float* dst = w + (size_t)blockIdx.x * inner;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) dst[i] = row[i] * scale;
}

// (C_out, C_in, K) -> (C_in, K, C_out)
__global__ void pack_conv1d_weight_kernel(
    const float* __restrict__ w, float* __restrict__ packed, int c_out, int c_in, int k) {
    const size_t total = (size_t)c_out * c_in * k;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx % c_out);
        const size_t cj = idx / c_out;
        const int j = (int)(cj % k);
        const int c = (int)(cj / k);
        packed[idx] = w[((size_t)o * c_in + c) * k + j];
    }
}

// promonet.edit.grid.sample (edit/grid.py:12-38), linear: out[r, t] = lerp of sequence[r, :] at
// grid[t], with the final frame replicated; optionally followed by the distribution-preserving
// renormalisation softmax(log(p + 1e-8)) over the `channels` rows of each item
// (preprocess/core.py:97-103): (p + 1e-8) / sum_c (p + 1e-8).  One thread per (item, t).
__global__ void __launch_bounds__(128) grid_sample_kernel(
    const float* __restrict__ sequence, const float* __restrict__ grid, float* __restrict__ out,
    int channels, int t_in, int t_out, int nearest, int renormalize) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int item = blockIdx.y;
    if (t >= t_out) return;
    const float x = grid[t];
    const float* src = sequence + (size_t)item * channels * t_in;
    float* dst = out + (size_t)item * channels * t_out + t;
    int lower, upper;
    float wl, wu;
    if (nearest) {
        lower = upper = min(max((int)rintf(x), 0), t_in - 1);   // torch.round: half to even
        wl = 1.f; wu = 0.f;
    } else {
        // i = searchsorted(arange(T), x, right) = number of integers k in [0, T) with k <= x
        int i = x < 0.f ? 0 : min((int)floorf(x) + 1, t_in);
        i = max(i, 1);                      // x < 0 indexes from the end in torch; clamp instead
        lower = i - 1;
        upper = min(i, t_in - 1);           // replicate padding of the final frame
        wl = (float)i - x;                  // xp[i] - x
        wu = x - (float)(i - 1);            // x - xp[i - 1]
    }
    float total = 0.f;
    for (int c = 0; c < channels; ++c) {
        const float* row = src + (size_t)c * t_in;
        const float v = row[lower] * wl + row[upper] * wu;
        if (renormalize) total += v + 1e-8f;
        else dst[(size_t)c * t_out] = v;
    }
    if (renormalize) {
        const float inv = 1.f / total;
        for (int c = 0; c < channels; ++c) {
            const float* row = src + (size_t)c * t_in;
            dst[(size_t)c * t_out] = (row[lower] * wl + row[upper] * wu + 1e-8f) * inv;
        }
    }
}

}  // namespace

int launch_grid_sample(
    const float* sequence, const float* grid, float* out, int items, int channels, int t_in,
    int t_out, bool nearest, bool renormalize, cudaStream_t stream) {
    PMN_REQUIRE(sequence && grid && out, "grid_sample: null pointer");
    PMN_REQUIRE(items > 0 && items <= 65535 && channels > 0 && t_in > 0, "grid_sample: bad shape");
    if (t_out <= 0) return PMN_OK;
    dim3 blocks(ceil_div(t_out, 128), items);
    LaunchScope scope("grid_sample_kernel", stream);
    grid_sample_kernel<<<blocks, 128, 0, stream>>>(
        sequence, grid, out, channels, t_in, t_out, nearest ? 1 : 0, renormalize ? 1 : 0);
    return launched("grid_sample_kernel");
}

int launch_features(
    const float* loudness, int rows, const float* pitch, const float* periodicity,
    const float* ppg, const float* pitch_distribution, const float* pitch_embedding,
    float threshold, bool with_period, float* out, int batch, int frames,
    cudaStream_t stream) {
    PMN_REQUIRE(loudness && pitch && periodicity && ppg && out, "features: null pointer");
    PMN_REQUIRE(rows >= kBands, "features: loudness needs at least 8 rows");
    PMN_REQUIRE(batch > 0 && batch <= 65535, "features: bad batch");
    if (frames <= 0) return PMN_OK;
    dim3 grid(ceil_div(frames, 128), batch);
    LaunchScope scope("features_kernel", stream);
    features_kernel<<<grid, 128, 0, stream>>>(
        loudness, rows, pitch, periodicity, ppg, pitch_distribution, pitch_embedding,
        threshold, with_period ? 1 : 0, out, frames);
    return launched("features_kernel");
}

int launch_speaker_bias(
    const float* speaker_embedding, const int64_t* speakers, const float* sbr,
    const float* lr, const float* weight, const float* bias, float* out,
    int batch, int speaker_channels, int c_out, int num_speakers, cudaStream_t stream) {
    PMN_REQUIRE(speaker_embedding && speakers && sbr && lr && weight && out,
                "speaker_bias: null pointer");
    LaunchScope scope("speaker_bias_kernel", stream);
    // the compiler reads g[] in 16-byte pieces, the last one past c_g floats: size the buffer for it
    const size_t smem = (size_t)(speaker_channels + 2 + 3) / 4 * 4 * sizeof(float);
    speaker_bias_kernel<<<batch, 256, smem, stream>>>(
        speaker_embedding, speakers, sbr, lr, weight, bias, out,
        speaker_channels, c_out, num_speakers);
    return launched("speaker_bias_kernel");
}

int launch_head(
    const float* x, const float* weight, float* out, int batch, int channels,
    int t_len, float slope, cudaStream_t stream) {
    PMN_REQUIRE(channels == 32, "head: only 32 input channels");
    dim3 grid(ceil_div(ceil_div(t_len, 4), 256), batch);
    LaunchScope scope("head_kernel", stream);
    head_kernel<32><<<grid, 256, 0, stream>>>(x, weight, out, t_len, slope);
    return launched("head_kernel");
}

int launch_weight_norm_fold(
    const float* v, const float* g, float* w, int dim0, int inner, cudaStream_t stream) {
    PMN_REQUIRE(v && g && w && dim0 > 0 && inner > 0, "weight_norm_fold: bad argument");
    LaunchScope scope("weight_norm_fold_kernel", stream);
    weight_norm_fold_kernel<<<dim0, 256, 0, stream>>>(v, g, w, inner);
    return launched("weight_norm_fold_kernel");
}

int launch_pack_conv1d_weight(
    const float* w, float* packed, int c_out, int c_in, int k, cudaStream_t stream) {
    PMN_REQUIRE(w && packed && c_out > 0 && c_in > 0 && k > 0, "pack_conv1d_weight: bad argument");
    const size_t total = (size_t)c_out * c_in * k;
    const int blocks = (int)min((size_t)4096, (total + 255) / 256);
    LaunchScope scope("pack_conv1d_weight_kernel", stream);
    pack_conv1d_weight_kernel<<<blocks, 256, 0, stream>>>(w, packed, c_out, c_in, k);
    return launched("pack_conv1d_weight_kernel");
}

}  // namespace pmn

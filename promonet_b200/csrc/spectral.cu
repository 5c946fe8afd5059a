// Framed STFT with fused epilogues: magnitude, log-mel, A-weighted loudness.
//
//   magnitude  promonet/preprocess/spectrogram.py:35-52   sqrt(re^2 + im^2 + 1e-6)
//   log-mel    spectrogram.py:111-135                      log(mel_basis @ magnitude)
//   loudness   promonet/preprocess/loudness.py:17-55       A-weighted dB, top_db 80, floor -100, band means
//
// All three share the reflect padding of (1024 - 256) / 2 samples, the periodic
// hann window and the 1024-point transform, so one kernel computes the frame
// spectrum once in shared memory and applies whichever epilogues are requested.
// A CTA processes kFrames consecutive frames so that the (B, bins, F) outputs are
// written in full 32-byte sectors.  Loudness needs the utterance maximum
// (librosa amplitude_to_db top_db): the kernel stores dB and reduces the maximum
// with an atomic, `loudness_finish_kernel` clamps, weights, floors and band-averages.
#include <math.h>

#include <mutex>
#include <vector>

#include "spectral.cuh"

namespace pmn {

namespace {

constexpr int kFft = 1024;      // NUM_FFT / WINDOW_SIZE, config/defaults.py:43,52
constexpr int kHop = 256;       // HOPSIZE :31
constexpr int kBins = kFft / 2 + 1;
constexpr int kMels = 80;       // NUM_MELS :40
constexpr int kPad = (kFft - kHop) / 2;
constexpr int kFrames = 8;      // frames per CTA
constexpr int kThreads = 256;
constexpr float kMinDb = -100.f;  // MIN_DB :37
constexpr double kSampleRate = 22050.;

using Tables = SpectralTables;

std::mutex g_tables_mutex;
std::vector<Tables> g_tables(64);

double hz_to_mel(double f) {
    return f >= 1000. ? 15. + log(f / 1000.) / (log(6.4) / 27.) : f / (200. / 3.);
}
double mel_to_hz(double m) {
    return m >= 15. ? 1000. * exp((log(6.4) / 27.) * (m - 15.)) : m * (200. / 3.);
}

// Device tables, built once per device in double precision
int tables(const Tables** out) {
    int device = 0;
    PMN_TRY(check_cuda(cudaGetDevice(&device), "cudaGetDevice"));
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    Tables& t = g_tables[device];
    if (!t.window) {
        const double pi = 3.14159265358979323846;
        std::vector<float> window(kFft), mel((size_t)kMels * kBins, 0.f), weights(kBins);
        std::vector<float2> twiddle(kFft / 2);
        std::vector<int> range(2 * kMels);
        for (int n = 0; n < kFft; ++n) window[n] = (float)(0.5 - 0.5 * cos(2. * pi * n / kFft));
        for (int k = 0; k < kFft / 2; ++k)
            twiddle[k] = make_float2((float)cos(2. * pi * k / kFft), (float)-sin(2. * pi * k / kFft));
        // librosa.filters.mel(sr=22050, n_fft=1024, n_mels=80): Slaney scale, Slaney norm
        std::vector<double> edges(kMels + 2);
        const double top = hz_to_mel(kSampleRate / 2.);
        for (int i = 0; i < kMels + 2; ++i) edges[i] = mel_to_hz(top * i / (kMels + 1));
        for (int m = 0; m < kMels; ++m) {
            int first = kBins, last = 0;
            const double norm = 2. / (edges[m + 2] - edges[m]);
            for (int k = 0; k < kBins; ++k) {
                const double f = k * kSampleRate / kFft;
                const double lower = (f - edges[m]) / (edges[m + 1] - edges[m]);
                const double upper = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
                const double w = fmax(0., fmin(lower, upper)) * norm;
                mel[(size_t)m * kBins + k] = (float)w;
                if (w > 0.) { first = k < first ? k : first; last = k + 1; }
            }
            range[2 * m] = first < last ? first : 0;
            range[2 * m + 1] = first < last ? last : 0;
        }
        // librosa.A_weighting(fft_frequencies) - REF_DB, loudness.py:149-160
        const double c[4] = {12194.217 * 12194.217, 20.598997 * 20.598997,
                             107.65265 * 107.65265, 737.86223 * 737.86223};
        for (int k = 0; k < kBins; ++k) {
            const double f = k * kSampleRate / kFft, f2 = f * f;
            double a = -80.;
            if (f2 > 0.) {
                a = 2. + 20. * (log10(c[0]) + 2. * log10(f2) - log10(f2 + c[0]) - log10(f2 + c[1]) -
                                0.5 * log10(f2 + c[2]) - 0.5 * log10(f2 + c[3]));
                a = fmax(-80., a);
            }
            weights[k] = (float)(a - 20.);
        }
        auto upload = [](auto** dst, const auto& src) {
            const size_t bytes = src.size() * sizeof(src[0]);
            PMN_TRY(check_cuda(cudaMalloc(dst, bytes), "cudaMalloc tables"));
            return check_cuda(cudaMemcpy(*dst, src.data(), bytes, cudaMemcpyHostToDevice), "upload tables");
        };
        PMN_TRY(upload(&t.window, window));
        PMN_TRY(upload(&t.twiddle, twiddle));
        PMN_TRY(upload(&t.mel_weights, mel));
        PMN_TRY(upload(&t.mel_range, range));
        PMN_TRY(upload(&t.a_weights, weights));
    }
    *out = &t;
    return PMN_OK;
}

}  // namespace

int spectral_tables(const SpectralTables** out) { return tables(out); }

namespace {

// Order-preserving float <-> int map for atomicMax on floats of either sign
__device__ __forceinline__ int float_key(float v) {
    const int bits = __float_as_int(v);
    return bits >= 0 ? bits : bits ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int key) {
    return __int_as_float(key >= 0 ? key : key ^ 0x7fffffff);
}

__global__ void __launch_bounds__(kThreads) stft_kernel(
    const float* __restrict__ audio, int samples, int frames,
    Tables t, float* __restrict__ magnitude, float* __restrict__ mels, float mel_floor,
    float* __restrict__ db, int* __restrict__ db_max) {
    __shared__ float2 buffer[2][kFft];
    __shared__ float2 twiddle[kFft / 2];
    __shared__ float spectrum[kBins][kFrames + 1];  // squared magnitude per frame
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFrames;
    const float* x = audio + (size_t)b * samples;
    for (int k = tid; k < kFft / 2; k += kThreads) twiddle[k] = t.twiddle[k];

    for (int frame = 0; frame < kFrames; ++frame) {
        const int f = f0 + frame;
        if (f >= frames) break;  // uniform across the block
        __syncthreads();
        // Windowed frame, reflect-padded (torch pad mode='reflect': no edge repeat)
        for (int n = tid; n < kFft; n += kThreads) {
            int i = f * kHop - kPad + n;
            if (i < 0) i = -i;
            if (i >= samples) i = 2 * (samples - 1) - i;
            i = min(max(i, 0), samples - 1);
            buffer[0][n] = make_float2(x[i] * t.window[n], 0.f);
        }
        // Stockham radix-2, 10 stages: stage s combines sub-transforms of length 2^s
        int source = 0;
#pragma unroll 1
        for (int half = 1; half < kFft; half <<= 1) {
            __syncthreads();
            const int stride = kFft / (2 * half);  // twiddle step
            for (int j = tid; j < kFft / 2; j += kThreads) {
                const int k = j & (half - 1);        // index within the sub-transform
                const int group = j / half;          // which pair of sub-transforms
                const float2 a = buffer[source][group * half + k];
                const float2 c = buffer[source][group * half + k + kFft / 2];
                const float2 w = twiddle[k * stride];
                const float2 wc = make_float2(w.x * c.x - w.y * c.y, w.x * c.y + w.y * c.x);
                buffer[source ^ 1][2 * group * half + k] = make_float2(a.x + wc.x, a.y + wc.y);
                buffer[source ^ 1][2 * group * half + k + half] = make_float2(a.x - wc.x, a.y - wc.y);
            }
            source ^= 1;
        }
        __syncthreads();
        for (int k = tid; k < kBins; k += kThreads) {
            const float2 v = buffer[source][k];
            spectrum[k][frame] = v.x * v.x + v.y * v.y;
        }
    }
    __syncthreads();
    const int valid = min(kFrames, frames - f0);

    if (magnitude) {
        for (int idx = tid; idx < kBins * kFrames; idx += kThreads) {
            const int k = idx / kFrames, frame = idx % kFrames;
            if (frame < valid)
                magnitude[((size_t)b * kBins + k) * frames + f0 + frame] = sqrtf(spectrum[k][frame] + 1e-6f);
        }
    }
    if (mels) {
        for (int idx = tid; idx < kMels * kFrames; idx += kThreads) {
            const int m = idx / kFrames, frame = idx % kFrames;
            if (frame >= valid) continue;
            const float* w = t.mel_weights + (size_t)m * kBins;
            float sum = 0.f;
            for (int k = t.mel_range[2 * m]; k < t.mel_range[2 * m + 1]; ++k)
                sum = fmaf(w[k], sqrtf(spectrum[k][frame] + 1e-6f), sum);
            mels[((size_t)b * kMels + m) * frames + f0 + frame] = fmaxf(logf(sum), mel_floor);
        }
    }
    if (db) {
        // librosa.amplitude_to_db(|X|) = 10 log10(max(1e-10, |X|^2))
        float local = -INFINITY;
        for (int idx = tid; idx < kBins * kFrames; idx += kThreads) {
            const int k = idx / kFrames, frame = idx % kFrames;
            if (frame >= valid) continue;
            const float value = 10.f * log10f(fmaxf(1e-10f, spectrum[k][frame]));
            db[((size_t)b * kBins + k) * frames + f0 + frame] = value;
            local = fmaxf(local, value);
        }
        for (int offset = 16; offset > 0; offset >>= 1)
            local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, offset));
        if ((tid & 31) == 0 && local > -INFINITY) atomicMax(db_max + b, float_key(local));
    }
}

__global__ void init_max_kernel(int* db_max, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) db_max[i] = float_key(-INFINITY);
}

// max(dB, max - 80) + A-weights, floor -100, then band means (loudness.py:46-55,84-111)
__global__ void __launch_bounds__(128) loudness_finish_kernel(
    const float* __restrict__ db, const int* __restrict__ db_max, const float* __restrict__ a_weights,
    float* __restrict__ out, int frames, int bands) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (f >= frames) return;
    const float threshold = key_float(db_max[b]) - 80.f;
    const float* src = db + (size_t)b * kBins * frames + f;
    if (bands <= 0) {
        float* dst = out + (size_t)b * kBins * frames + f;
        for (int k = 0; k < kBins; ++k)
            dst[(size_t)k * frames] =
                fmaxf(fmaxf(src[(size_t)k * frames], threshold) + a_weights[k], kMinDb);
        return;
    }
    const float step = (float)kBins / (float)bands;
    for (int band = 0; band < bands; ++band) {
        const int start = bands == 1 ? 0 : (int)(band * step);
        const int stop = bands == 1 ? kBins : (int)((band + 1) * step);
        float sum = 0.f;
        for (int k = start; k < stop; ++k)
            sum += fmaxf(fmaxf(src[(size_t)k * frames], threshold) + a_weights[k], kMinDb);
        out[((size_t)b * bands + band) * frames + f] = sum / (float)(stop - start);
    }
}

// log(mel_basis @ spectrogram) for an existing magnitude spectrogram (spectrogram.py:111-135)
__global__ void __launch_bounds__(128) mel_kernel(
    const float* __restrict__ magnitude, Tables t, float* __restrict__ mels, float mel_floor, int frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y, b = blockIdx.z;
    if (f >= frames) return;
    const float* w = t.mel_weights + (size_t)m * kBins;
    const float* src = magnitude + (size_t)b * kBins * frames + f;
    float sum = 0.f;
    for (int k = t.mel_range[2 * m]; k < t.mel_range[2 * m + 1]; ++k)
        sum = fmaf(w[k], src[(size_t)k * frames], sum);
    mels[((size_t)b * kMels + m) * frames + f] = fmaxf(logf(sum), mel_floor);
}

}  // namespace

int launch_linear_to_mel(
    const float* magnitude, float* mels, float mel_floor, int batch, int frames, cudaStream_t stream) {
    PMN_REQUIRE(magnitude && mels && batch > 0 && batch <= 65535, "linear_to_mel: bad argument");
    if (frames <= 0) return PMN_OK;
    const Tables* t;
    PMN_TRY(tables(&t));
    dim3 grid(ceil_div(frames, 128), kMels, batch);
    LaunchScope scope("mel_kernel", stream);
    mel_kernel<<<grid, 128, 0, stream>>>(magnitude, *t, mels, mel_floor, frames);
    return launched("mel_kernel");
}

int spectral_frames(int samples) { return samples / kHop; }

size_t spectral_workspace_bytes(int batch, int samples) {
    return align_up((size_t)batch * kBins * spectral_frames(samples) * sizeof(float), 256) +
           align_up((size_t)batch * sizeof(int), 256);
}

int launch_spectral_features(
    const float* audio, int batch, int samples, float* magnitude, float* mels, float mel_floor,
    float* loudness, int loudness_bands, void* workspace, size_t workspace_bytes,
    cudaStream_t stream) {
    PMN_REQUIRE(audio && batch > 0 && batch <= 65535, "spectral_features: bad argument");
    PMN_REQUIRE(samples >= kHop, "spectral_features: audio shorter than one hop");
    PMN_REQUIRE(samples > kPad, "spectral_features: audio shorter than the reflect padding");
    PMN_REQUIRE(magnitude || mels || loudness, "spectral_features: no output requested");
    const int frames = spectral_frames(samples);
    float* db = nullptr;
    int* db_max = nullptr;
    if (loudness) {
        PMN_REQUIRE(workspace, "spectral_features: loudness needs a workspace");
        if (spectral_workspace_bytes(batch, samples) > workspace_bytes)
            return fail(PMN_ERR_WORKSPACE, "spectral_features: workspace too small");
        db = static_cast<float*>(workspace);
        db_max = reinterpret_cast<int*>(
            static_cast<char*>(workspace) + align_up((size_t)batch * kBins * frames * sizeof(float), 256));
        LaunchScope scope("init_max_kernel", stream);
        init_max_kernel<<<ceil_div(batch, 128), 128, 0, stream>>>(db_max, batch);
        PMN_TRY(launched("init_max_kernel"));
    }
    const Tables* t;
    PMN_TRY(tables(&t));
    {
        dim3 grid(ceil_div(frames, kFrames), batch);
        LaunchScope scope("stft_kernel", stream);
        stft_kernel<<<grid, kThreads, 0, stream>>>(
            audio, samples, frames, *t, magnitude, mels, mel_floor, db, db_max);
        PMN_TRY(launched("stft_kernel"));
    }
    if (loudness) {
        dim3 grid(ceil_div(frames, 128), batch);
        LaunchScope scope("loudness_finish_kernel", stream);
        loudness_finish_kernel<<<grid, 128, 0, stream>>>(
            db, db_max, t->a_weights, loudness, frames, loudness_bands);
        PMN_TRY(launched("loudness_finish_kernel"));
    }
    return PMN_OK;
}

}  // namespace pmn

// Differentiable STFT magnitude and the mel loss of the training step.
//
//   mel loss        promonet/train/core.py:277-305: L1(log(M @ |STFT(y)|), log(M @ S)) * 45 with
//                   |X| = sqrt(re^2 + im^2 + 1e-6), hann 1024 / hop 256, reflect pad 384
//                   (promonet/preprocess/spectrogram.py:15-60,111-135)
//   CMB spectrogram promonet/model/discriminator.py:175-195: the same framing with no
//                   window (rectangular) and |X| = sqrt(re^2 + im^2), laid out (B, 1, F, 513)
//
// Forward keeps the complex spectrum; backward turns dL/d|X| into dL/dX = g X / |X|,
// runs one more 1024-point transform per frame (the adjoint of the one-sided DFT is
// the real part of a DFT of the conjugated, zero-extended gradient) and scatters the
// windowed result back through the overlap and the reflect padding.
#include "spectral.cuh"
#include "train.cuh"

namespace pmn {

namespace {

constexpr int kFft = 1024;
constexpr int kHop = 256;
constexpr int kBins = kFft / 2 + 1;
constexpr int kMels = 80;
constexpr int kPad = (kFft - kHop) / 2;
constexpr int kThreads = 256;

// In-place-pair Stockham radix-2 over buffer[2][1024]; returns which half holds the result
__device__ __forceinline__ int fft1024(float2 (*buffer)[kFft], const float2* twiddle, int tid) {
    int source = 0;
#pragma unroll 1
    for (int half = 1; half < kFft; half <<= 1) {
        __syncthreads();
        const int stride = kFft / (2 * half);
        for (int j = tid; j < kFft / 2; j += kThreads) {
            const int k = j & (half - 1);
            const int group = j / half;
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
    return source;
}

__device__ __forceinline__ int reflect_index(int i, int samples) {
    if (i < 0) i = -i;
    if (i >= samples) i = 2 * (samples - 1) - i;
    return min(max(i, 0), samples - 1);
}

// Band edges of the complex multi-band discriminator: int(fraction * 513),
// fractions 0, .1, .25, .5, .75, 1 (discriminator.py:150,161-163)
__constant__ int kBandEdges[6] = {0, 51, 128, 256, 384, 513};

// magnitude layouts: 0 -> (B, 513, F); 1 -> (B, F, 513); 2 -> the five bands of layout 1
// stored one after the other, band i as a contiguous (B, F, hi_i - lo_i) tensor
__device__ __forceinline__ size_t magnitude_index(
    int layout, int b, int k, int f, int frames, int batch) {
    if (layout == 0) return ((size_t)b * kBins + k) * frames + f;
    if (layout == 1) return ((size_t)b * frames + f) * kBins + k;
    int band = 0;
    while (k >= kBandEdges[band + 1]) ++band;
    const int lo = kBandEdges[band], width = kBandEdges[band + 1] - lo;
    return (size_t)batch * frames * lo + ((size_t)b * frames + f) * width + (k - lo);
}

// One frame per CTA
__global__ void __launch_bounds__(kThreads) stft_train_kernel(
    const float* __restrict__ audio, int samples, int frames, SpectralTables t, int window_kind,
    float eps, int layout, float2* __restrict__ spectrum, float* __restrict__ magnitude) {
    __shared__ float2 buffer[2][kFft];
    __shared__ float2 twiddle[kFft / 2];
    const int tid = threadIdx.x;
    const int f = blockIdx.x, b = blockIdx.y;
    const float* x = audio + (size_t)b * samples;
    for (int k = tid; k < kFft / 2; k += kThreads) twiddle[k] = t.twiddle[k];
    for (int n = tid; n < kFft; n += kThreads) {
        const int i = reflect_index(f * kHop - kPad + n, samples);
        const float w = window_kind == 0 ? t.window[n] : 1.f;
        buffer[0][n] = make_float2(x[i] * w, 0.f);
    }
    const int source = fft1024(buffer, twiddle, tid);
    for (int k = tid; k < kBins; k += kThreads) {
        const float2 v = buffer[source][k];
        if (spectrum) spectrum[((size_t)b * frames + f) * kBins + k] = v;
        const float mag = sqrtf(v.x * v.x + v.y * v.y + eps);
        if (magnitude) magnitude[magnitude_index(layout, b, k, f, frames, gridDim.y)] = mag;
    }
}

__global__ void __launch_bounds__(kThreads) stft_train_backward_kernel(
    const float* __restrict__ gmagnitude, const float2* __restrict__ spectrum, int samples,
    int frames, SpectralTables t, int window_kind, float eps, int layout,
    float* __restrict__ gaudio) {
    __shared__ float2 buffer[2][kFft];
    __shared__ float2 twiddle[kFft / 2];
    const int tid = threadIdx.x;
    const int f = blockIdx.x, b = blockIdx.y;
    for (int k = tid; k < kFft / 2; k += kThreads) twiddle[k] = t.twiddle[k];
    for (int k = tid; k < kFft; k += kThreads) {
        float2 value = make_float2(0.f, 0.f);
        if (k < kBins) {
            const float2 v = spectrum[((size_t)b * frames + f) * kBins + k];
            const size_t idx = magnitude_index(layout, b, k, f, frames, gridDim.y);
            const float mag = sqrtf(v.x * v.x + v.y * v.y + eps);
            // torch.norm's subgradient at 0 is 0
            const float scale = mag > 0.f ? gmagnitude[idx] / mag : 0.f;
            value = make_float2(scale * v.x, -scale * v.y);  // conj(dL/dX)
        }
        buffer[0][k] = value;
    }
    const int source = fft1024(buffer, twiddle, tid);
    float* gx = gaudio + (size_t)b * samples;
    for (int n = tid; n < kFft; n += kThreads) {
        const float w = window_kind == 0 ? t.window[n] : 1.f;
        const float v = buffer[source][n].x * w;
        const int i = reflect_index(f * kHop - kPad + n, samples);
        if (v != 0.f) atomicAdd(gx + i, v);
    }
}

// loss += loss_weight * mean|log(M @ mag) - target|, gmagnitude = grad_weight * dmean / dmag
// (one frame per CTA)
__global__ void __launch_bounds__(128) mel_loss_kernel(
    const float* __restrict__ magnitude, const float* __restrict__ target, int frames, int batch,
    SpectralTables t, float loss_weight, float grad_weight, float* __restrict__ loss,
    float* __restrict__ gmagnitude) {
    __shared__ float mag[kBins];
    __shared__ float gmel[kMels];
    __shared__ float partial[4];
    const int tid = threadIdx.x;
    const int f = blockIdx.x, b = blockIdx.y;
    const float* src = magnitude + (size_t)b * kBins * frames + f;
    for (int k = tid; k < kBins; k += blockDim.x) mag[k] = src[(size_t)k * frames];
    __syncthreads();
    const float count = (float)batch * kMels * frames;
    const float scale = loss_weight / count;
    const float gscale = grad_weight / count;
    float local = 0.f;
    if (tid < kMels) {
        const float* w = t.mel_weights + (size_t)tid * kBins;
        float sum = 0.f;
        for (int k = t.mel_range[2 * tid]; k < t.mel_range[2 * tid + 1]; ++k)
            sum = fmaf(w[k], mag[k], sum);
        const float d = logf(sum) - target[((size_t)b * kMels + tid) * frames + f];
        local = fabsf(d) * scale;
        const float sign = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        gmel[tid] = sign * gscale / sum;
    }
    for (int offset = 16; offset > 0; offset >>= 1) local += __shfl_xor_sync(0xffffffffu, local, offset);
    if ((tid & 31) == 0) partial[tid >> 5] = local;
    __syncthreads();
    if (tid == 0 && loss) atomicAdd(loss, partial[0] + partial[1] + partial[2] + partial[3]);
    if (gmagnitude) {
        float* dst = gmagnitude + (size_t)b * kBins * frames + f;
        for (int k = tid; k < kBins; k += blockDim.x) {
            float sum = 0.f;
            for (int m = 0; m < kMels; ++m) {
                if (k >= t.mel_range[2 * m] && k < t.mel_range[2 * m + 1])
                    sum = fmaf(t.mel_weights[(size_t)m * kBins + k], gmel[m], sum);
            }
            dst[(size_t)k * frames] = sum;
        }
    }
}

// Windowed one-sided DFT basis as a (2 bins, n_fft) weight: rows [0, bins) = hann[n] cos(2 pi k n / N),
// rows [bins, 2 bins) = -hann[n] sin(2 pi k n / N) (periodic hann, torch.hann_window; loss.py:99)
__global__ void dft_basis_kernel(float* __restrict__ out, int n_fft, int bins) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (n >= n_fft) return;
    const double window = 0.5 - 0.5 * cospi(2. * n / n_fft);
    // k n mod N keeps the argument small: exact phase for every (k, n)
    const long long phase = ((long long)k * n) % n_fft;
    double sine, cosine;
    sincospi(2. * (double)phase / n_fft, &sine, &cosine);
    out[(size_t)k * n_fft + n] = (float)(window * cosine);
    out[(size_t)(bins + k) * n_fft + n] = (float)(-window * sine);
}

__device__ __forceinline__ float root_magnitude(float re, float im, float* magnitude) {
    // loss.py:73-80: sqrt(clamp(|X|, 1e-7))
    *magnitude = sqrtf(re * re + im * im);
    return sqrtf(fmaxf(*magnitude, 1e-7f));
}

// spec (2 B, 2 bins, frames): items [0, B) are the target y, [B, 2 B) the prediction x.
// sums[0] += sum |s_y - s_x|, sums[1] += sum s_y   (loss.py:121: ||y - x||_1 / ||y||_1)
__global__ void __launch_bounds__(256) spectral_convergence_sums_kernel(
    const float* __restrict__ spec, int batch, int bins, int frames, float* __restrict__ sums) {
    __shared__ float scratch[2][32];
    const size_t item = (size_t)2 * bins * frames, half = (size_t)bins * frames;
    const size_t total = (size_t)batch * half;
    float difference = 0.f, reference = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / half, rest = idx - b * half;
        const float* y = spec + b * item + rest;
        const float* x = spec + (b + batch) * item + rest;
        float unused;
        const float sy = root_magnitude(y[0], y[half], &unused);
        const float sx = root_magnitude(x[0], x[half], &unused);
        difference += fabsf(sy - sx);
        reference += sy;
    }
    for (int offset = 16; offset > 0; offset >>= 1) {
        difference += __shfl_xor_sync(0xffffffffu, difference, offset);
        reference += __shfl_xor_sync(0xffffffffu, reference, offset);
    }
    if ((threadIdx.x & 31) == 0) {
        scratch[0][threadIdx.x >> 5] = difference;
        scratch[1][threadIdx.x >> 5] = reference;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float a = threadIdx.x < (blockDim.x >> 5) ? scratch[0][threadIdx.x] : 0.f;
        float c = threadIdx.x < (blockDim.x >> 5) ? scratch[1][threadIdx.x] : 0.f;
        for (int offset = 16; offset > 0; offset >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, offset);
            c += __shfl_xor_sync(0xffffffffu, c, offset);
        }
        if (threadIdx.x == 0) {
            atomicAdd(sums, a);
            atomicAdd(sums + 1, c);
        }
    }
}

// *loss += weight sums[0] / sums[1]; gspec (B, 2 bins, frames) = weight d(sums[0] / sums[1]) / d spec_x
__global__ void __launch_bounds__(256) spectral_convergence_backward_kernel(
    const float* __restrict__ spec, int batch, int bins, int frames, const float* __restrict__ sums,
    float weight, float* __restrict__ loss, float* __restrict__ gspec) {
    const size_t item = (size_t)2 * bins * frames, half = (size_t)bins * frames;
    const size_t total = (size_t)batch * half;
    const float scale = weight / sums[1];
    if (loss && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(loss, sums[0] * scale);
    if (!gspec) return;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / half, rest = idx - b * half;
        const float* y = spec + b * item + rest;
        const float* x = spec + (b + batch) * item + rest;
        float unused, magnitude;
        const float sy = root_magnitude(y[0], y[half], &unused);
        const float sx = root_magnitude(x[0], x[half], &magnitude);
        float gre = 0.f, gim = 0.f;
        if (magnitude > 1e-7f) {
            const float sign = sx > sy ? 1.f : (sx < sy ? -1.f : 0.f);
            const float gmagnitude = scale * sign * 0.5f / sx;
            gre = gmagnitude * x[0] / magnitude;
            gim = gmagnitude * x[half] / magnitude;
        }
        gspec[b * item + rest] = gre;
        gspec[b * item + rest + half] = gim;
    }
}

// gsignal[b, f hop + n] += gframes[b, n, f]: adjoint of reading a signal as overlapping frames
__global__ void frame_overlap_add_kernel(
    const float* __restrict__ gframes, float* __restrict__ gsignal, int n_fft, int frames, int hop,
    int samples) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y, b = blockIdx.z;
    if (f >= frames) return;
    atomicAdd(gsignal + (size_t)b * samples + (size_t)f * hop + n,
              gframes[((size_t)b * n_fft + n) * frames + f]);
}

}  // namespace

int launch_dft_basis(float* out, int n_fft, cudaStream_t stream) {
    PMN_REQUIRE(out && n_fft >= 2 && n_fft % 2 == 0, "dft_basis: bad argument");
    const int bins = n_fft / 2 + 1;
    dim3 grid(ceil_div(n_fft, 128), bins);
    LaunchScope scope("dft_basis_kernel", stream);
    dft_basis_kernel<<<grid, 128, 0, stream>>>(out, n_fft, bins);
    return launched("dft_basis_kernel");
}

int launch_spectral_convergence(
    const float* spec, int batch, int bins, int frames, float weight, float* sums, float* loss,
    float* gspec, cudaStream_t stream) {
    PMN_REQUIRE(spec && sums && batch > 0 && bins > 0 && frames > 0, "spectral_convergence: bad argument");
    const size_t total = (size_t)batch * bins * frames;
    const int blocks = (int)min((size_t)148 * 4, (total + 255) / 256);
    PMN_TRY(check_cuda(cudaMemsetAsync(sums, 0, 2 * sizeof(float), stream), "spectral_convergence memset"));
    {
        LaunchScope scope("spectral_convergence_sums_kernel", stream);
        spectral_convergence_sums_kernel<<<blocks, 256, 0, stream>>>(spec, batch, bins, frames, sums);
        PMN_TRY(launched("spectral_convergence_sums_kernel"));
    }
    LaunchScope scope("spectral_convergence_backward_kernel", stream);
    spectral_convergence_backward_kernel<<<blocks, 256, 0, stream>>>(
        spec, batch, bins, frames, sums, weight, loss, gspec);
    return launched("spectral_convergence_backward_kernel");
}

int launch_frame_overlap_add(
    const float* gframes, float* gsignal, int batch, int n_fft, int frames, int hop, int samples,
    cudaStream_t stream) {
    PMN_REQUIRE(gframes && gsignal && batch > 0 && batch <= 65535 && n_fft > 0 && n_fft <= 65535 &&
                frames > 0 && hop > 0 && (frames - 1) * hop + n_fft <= samples,
                "frame_overlap_add: bad argument");
    dim3 grid(ceil_div(frames, 64), n_fft, batch);
    LaunchScope scope("frame_overlap_add_kernel", stream);
    frame_overlap_add_kernel<<<grid, 64, 0, stream>>>(gframes, gsignal, n_fft, frames, hop, samples);
    return launched("frame_overlap_add_kernel");
}

size_t stft_train_frames(int samples) { return (size_t)(samples / kHop); }

int launch_stft_train(
    const float* audio, int batch, int samples, int window_kind, float eps, int layout,
    float* spectrum, float* magnitude, cudaStream_t stream) {
    PMN_REQUIRE(audio && batch > 0 && batch <= 65535, "stft_train: bad argument");
    PMN_REQUIRE(samples > kPad && samples >= kHop, "stft_train: audio too short");
    PMN_REQUIRE(spectrum || magnitude, "stft_train: no output");
    const int frames = samples / kHop;
    const SpectralTables* t;
    PMN_TRY(spectral_tables(&t));
    dim3 grid(frames, batch);
    LaunchScope scope("stft_train_kernel", stream);
    stft_train_kernel<<<grid, kThreads, 0, stream>>>(
        audio, samples, frames, *t, window_kind, eps, layout,
        reinterpret_cast<float2*>(spectrum), magnitude);
    return launched("stft_train_kernel");
}

int launch_stft_train_backward(
    const float* gmagnitude, const float* spectrum, int batch, int samples, int window_kind,
    float eps, int layout, float* gaudio, int accumulate, cudaStream_t stream) {
    PMN_REQUIRE(gmagnitude && spectrum && gaudio && batch > 0 && batch <= 65535,
                "stft_train_backward: bad argument");
    PMN_REQUIRE(samples > kPad && samples >= kHop, "stft_train_backward: audio too short");
    const int frames = samples / kHop;
    const SpectralTables* t;
    PMN_TRY(spectral_tables(&t));
    if (!accumulate)
        PMN_TRY(check_cuda(
            cudaMemsetAsync(gaudio, 0, (size_t)batch * samples * sizeof(float), stream),
            "stft_train_backward memset"));
    dim3 grid(frames, batch);
    LaunchScope scope("stft_train_backward_kernel", stream);
    stft_train_backward_kernel<<<grid, kThreads, 0, stream>>>(
        gmagnitude, reinterpret_cast<const float2*>(spectrum), samples, frames, *t, window_kind,
        eps, layout, gaudio);
    return launched("stft_train_backward_kernel");
}

int launch_mel_loss(
    const float* magnitude, const float* target_mels, int batch, int frames, float loss_weight,
    float grad_weight, float* loss, float* gmagnitude, cudaStream_t stream) {
    PMN_REQUIRE(magnitude && target_mels && (loss || gmagnitude) && batch > 0 && batch <= 65535 &&
                frames > 0, "mel_loss: bad argument");
    const SpectralTables* t;
    PMN_TRY(spectral_tables(&t));
    dim3 grid(frames, batch);
    LaunchScope scope("mel_loss_kernel", stream);
    mel_loss_kernel<<<grid, 128, 0, stream>>>(
        magnitude, target_mels, frames, batch, *t, loss_weight, grad_weight, loss, gmagnitude);
    return launched("mel_loss_kernel");
}

}  // namespace pmn

// Multi-resolution spectrogram discriminator front end (DiscriminatorR.spectrogram,
// promonet/model/discriminator.py:127-141; flag MULTI_RESOLUTION_DISCRIMINATOR, off by default).
//
// The STFTs (n_fft / hop / win = 1024/120/600, 2048/240/1200, 512/50/240: hops that do not divide
// the window, windows shorter than the transform) run as 1 x 1 convolutions of the reflect-padded
// signal read in place as overlapping frames (pmn_conv_gemm[_tc] with pmn_conv_geometry strides,
// exactly like the spectral-convergence loss, train_spectral.cu); this file supplies the weight of
// that convolution for torch.stft(window=None, win_length < n_fft) — a rectangular window of
// win_length samples centred in the frame — and the magnitude between the STFT and the first
// convolution, forward and backward.  HBM-bound, one pass each.  Nothing on the default training
// path launches these kernels (the flag is off in config/promonet.py).
#include "common.cuh"

namespace pmn {

namespace {

// (2 bins, n_fft) weight: rows [0, bins) = w[n] cos(2 pi k n / N), rows [bins, 2 bins) =
// -w[n] sin(2 pi k n / N), w = 1 on [left, left + win_length), left = (n_fft - win_length) / 2
// (torch.stft pads a short window on both sides to n_fft), else 0
__global__ void dft_basis_rect_kernel(float* __restrict__ out, int n_fft, int bins, int win_length) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (n >= n_fft) return;
    const int left = (n_fft - win_length) / 2;
    const bool inside = n >= left && n < left + win_length;
    // k n mod N keeps the argument small: exact phase for every (k, n)
    const long long phase = ((long long)k * n) % n_fft;
    double sine, cosine;
    sincospi(2. * (double)phase / n_fft, &sine, &cosine);
    out[(size_t)k * n_fft + n] = inside ? (float)cosine : 0.f;
    out[(size_t)(bins + k) * n_fft + n] = inside ? (float)(-sine) : 0.f;
}

// spec (items, 2 bins, frames) -> magnitude (items, bins, frames) = sqrt(re^2 + im^2)
// (torch.norm(view_as_real(X), p=2, dim=-1), discriminator.py:141: no epsilon)
__global__ void __launch_bounds__(256) complex_magnitude_kernel(
    const float* __restrict__ spec, float* __restrict__ magnitude, size_t half, size_t total) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t item = idx / half, rest = idx - item * half;
        const float re = spec[item * 2 * half + rest];
        const float im = spec[item * 2 * half + half + rest];
        magnitude[idx] = sqrtf(re * re + im * im);
    }
}

// gspec (items, 2 bins, frames) = gmagnitude * (re, im) / |X|, zero where |X| = 0 (the
// subgradient torch's norm backward takes)
__global__ void __launch_bounds__(256) complex_magnitude_backward_kernel(
    const float* __restrict__ gmagnitude, const float* __restrict__ spec,
    float* __restrict__ gspec, size_t half, size_t total) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t item = idx / half, rest = idx - item * half;
        const size_t real = item * 2 * half + rest, imaginary = real + half;
        const float re = spec[real], im = spec[imaginary];
        const float norm = sqrtf(re * re + im * im);
        const float scale = norm > 0.f ? gmagnitude[idx] / norm : 0.f;
        gspec[real] = scale * re;
        gspec[imaginary] = scale * im;
    }
}

inline int blocks_for(size_t total) {
    return (int)min((size_t)148 * 8, (total + 255) / 256);
}

}  // namespace

}  // namespace pmn

using namespace pmn;

extern "C" {

int pmn_dft_basis_rect(float* out, int n_fft, int win_length, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMN_REQUIRE(out && n_fft >= 2 && n_fft % 2 == 0 && win_length >= 1 && win_length <= n_fft,
                "dft_basis_rect: bad argument");
    const int bins = n_fft / 2 + 1;
    dim3 grid(ceil_div(n_fft, 128), bins);
    LaunchScope scope("dft_basis_rect_kernel", stream);
    dft_basis_rect_kernel<<<grid, 128, 0, stream>>>(out, n_fft, bins, win_length);
    return launched("dft_basis_rect_kernel");
}

int pmn_complex_magnitude(
    const float* spec, float* magnitude, int items, int bins, int frames, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMN_REQUIRE(spec && magnitude && items > 0 && bins > 0 && frames > 0,
                "complex_magnitude: bad argument");
    const size_t half = (size_t)bins * frames, total = half * items;
    LaunchScope scope("complex_magnitude_kernel", stream);
    complex_magnitude_kernel<<<blocks_for(total), 256, 0, stream>>>(spec, magnitude, half, total);
    return launched("complex_magnitude_kernel");
}

int pmn_complex_magnitude_backward(
    const float* gmagnitude, const float* spec, float* gspec, int items, int bins, int frames,
    void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMN_REQUIRE(gmagnitude && spec && gspec && items > 0 && bins > 0 && frames > 0,
                "complex_magnitude_backward: bad argument");
    const size_t half = (size_t)bins * frames, total = half * items;
    LaunchScope scope("complex_magnitude_backward_kernel", stream);
    complex_magnitude_backward_kernel<<<blocks_for(total), 256, 0, stream>>>(
        gmagnitude, spec, gspec, half, total);
    return launched("complex_magnitude_backward_kernel");
}

}  // extern "C"

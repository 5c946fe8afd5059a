// penn-style pitch / periodicity estimation (promonet/preprocess/core.py:64-85:
// penn.from_audio(..., center='half-hop', decoder='viterbi')), FCNF0++ network.
//
// Pipeline (definition pinned in oracle/penn.py; penn itself is un-vendored):
//   resample to 8 kHz (torchaudio sinc_interp_hann polyphase FIR)
//   -> 1024-sample frames every int(hopsize * 8000) samples, reflect-padded, crop [16:-15]
//   -> 6 x [Conv1d k32 -> ReLU -> (MaxPool 2) -> LayerNorm(C, L)] -> Conv1d(512 -> 1440, k4)
//   -> mask bins outside [fmin, fmax), softmax, entropy periodicity
//   -> Viterbi (viterbi.cu) -> local expected value in a 19-bin window -> Hz
//
// Frames are laid end to end on one time axis per channel ([C][frames * L]) so the
// k32 convolutions run as long 1-D convolutions (conv1d.cu); the 31 samples that
// straddle two frames are computed and dropped by the pooling / LayerNorm kernel.
// The last block writes its (512, 4) activations transposed so that the final
// layer is a dense 2048 -> 1440 product over all frames (conv1d with k = 1).
//
// Block 1 (256 -> 32 channels, k = 32: 60 % of the network's FLOPs) has only 32 output
// columns, a quarter of what one M = 128 MMA needs to amortise its A-operand fetch.  On the
// tensor-core path it runs "folded by 4" in time instead: four consecutive samples become
// four channel groups (x4[t4, (p, c)] = x[4 t4 + p, c], 1024 channels) and four consecutive
// outputs become four column groups (y4[t4, (q, o)] = y[4 t4 + q, o], 128 columns) of a
// 9-tap convolution with the block-Toeplitz weight W4[(q, o), (p, c), J] = W[o, c, 4 J + p - q]
// (zero outside 0..31): 12.5 % more multiply-adds, all of them at full tensor rate.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdlib.h>

#include <map>
#include <new>
#include <numeric>
#include <vector>

#include "conv1d_tc.cuh"
#include "pitch.cuh"
#include "spectral.cuh"
#include "tc_ptx.cuh"
#include "tensor_store.cuh"

namespace pmn {

namespace {

constexpr int kRate = 8000;          // penn.SAMPLE_RATE
constexpr int kWindow = 1024;        // penn.WINDOW_SIZE
constexpr int kCropped = 993;        // frames[:, :, 16:-15]
constexpr int kCropStart = 16;
constexpr int kBins = 1440;          // penn.PITCH_BINS
constexpr int kKernel = 32;
constexpr int kLayers = 6;
constexpr int kChannels[kLayers + 1] = {1, 256, 32, 32, 128, 256, 512};
constexpr bool kPooled[kLayers] = {true, true, true, false, false, false};
constexpr int kLength[kLayers + 1] = {993, 481, 225, 97, 66, 35, 4};  // per-frame length after block i
constexpr float kCentsPerBin = 5.f, kFmin = 31.f, kOctave = 1200.f;
constexpr int kLocalWindow = 19;
constexpr int kFold = 4;                       // block 1 on the tensor cores: folded by 4 in time
constexpr int kFoldTaps = 9;                   // ceil((32 + 3) / 4)
constexpr int kFoldStride = 484;               // block 1's frame stride: a multiple of 4 >= 481
constexpr int kFrameMajorFirst = 3;            // blocks 3, 4, 5 and the head run frame-major (below)
// Block 0 on the tensor-core path takes its own operand buffer (im2col rows) and runs in
// sub-chunks of at most kSubFrames frames.  Measured (profiles/r2_preprocess_history.txt):
// sub-chunks small enough for the convolution output to stay in L2 (95 frames, 48 MB) make
// the LayerNorm pass 20 % faster but cost more than that in per-launch overheads (296
// instead of 14 launches of each kernel), so the sub-chunk is the whole default frame batch
constexpr int kSubFrames = 2048;
// Block 0 shared between frames (hop even): consecutive frames overlap by 1 - hop / 1024 (91 % at
// the default hop of 92 samples) and block 0's convolution, ReLU and MaxPool do not depend on the
// frame, only its LayerNorm does.  The convolution therefore runs ONCE over each reflect-padded
// utterance (10.7 x fewer multiply-adds than frame by frame at the default hop) and the LayerNorm
// kernel reads frame f's 481 pooled rows at offset (hop f + 16) / 2 of that shared output.
// Utterances are processed in groups whose pooled output stays below this many rows (x 256
// channels x 4 bytes = 2 GiB).
constexpr int kPooledRowsCap = 1 << 21;

}  // namespace

}  // namespace pmn

struct pmn_pitch {
    pmn::TensorStore store;
    bool finalized = false;
    int math = PMN_MATH_FP32_SIMT;
    __nv_bfloat16* conv_slabs[pmn::kLayers] = {};  // tensor-core path, layers 1..5
    float* conv_weight[pmn::kLayers] = {};   // packed (C_in, 32, C_out)
    const float* conv_bias[pmn::kLayers] = {};
    const float* norm_weight[pmn::kLayers] = {};
    const float* norm_bias[pmn::kLayers] = {};
    float* head_weight = nullptr;            // packed (2048, 1, 1440)
    __nv_bfloat16* head_slabs = nullptr;     // tensor-core path
    const float* head_bias = nullptr;
    const float* folded_bias = nullptr;      // block 1 folded by 4: bias of column (q, o)
    // block 1 with "fp16 + 2 x fp8" operands (conv1d_tc.cuh): its slabs and the weights' power-of-two
    // scale; -1: bf16 x 3 (PMN_PITCH_F8=0 at model construction; the fp8 form is the default)
    void* block1_f8_slabs = nullptr;
    int block1_shift = -1;
    // resampling tables per input rate: (2 width + orig, new) transposed FIR bank
    struct Resampler { float* table; int orig, fresh, width; };
    std::map<int, Resampler> resamplers;

    ~pmn_pitch() {
        for (auto& item : resamplers) cudaFree(item.second.table);
    }
};

namespace pmn {

namespace {

// y[n], n = q * fresh + p:  sum_k table[k][p] * x[q * orig + k - width]
__global__ void __launch_bounds__(256) resample_kernel(
    const float* __restrict__ audio, const float* __restrict__ table,
    float* __restrict__ out, int samples, int out_samples, int orig, int fresh, int width) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (n >= out_samples) return;
    const int q = n / fresh, p = n % fresh;
    const float* x = audio + (size_t)b * samples;
    const int taps = 2 * width + orig;
    const int start = q * orig - width;
    float acc = 0.f;
    for (int k = 0; k < taps; ++k) {
        const int i = start + k;
        if (i >= 0 && i < samples) acc = fmaf(table[(size_t)k * fresh + p], x[i], acc);
    }
    out[(size_t)b * out_samples + n] = acc;
}

// Cropped frames laid end to end: x[g * 993 + n] = padded[b][hop * f + 16 + n], g = first + local
__global__ void __launch_bounds__(256) frames_kernel(
    const float* __restrict__ audio, float* __restrict__ out, int samples, int frames_per_item,
    int first_frame, int count, int hop, int padding, int stride) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int local = blockIdx.y;
    if (n >= stride || local >= count) return;
    if (n >= kCropped) {  // pad sample that keeps frame starts on even rows
        out[(size_t)local * stride + n] = 0.f;
        return;
    }
    const int g = first_frame + local;
    const int b = g / frames_per_item, f = g % frames_per_item;
    int i = hop * f + kCropStart + n - padding;  // index into the unpadded audio
    float value = 0.f;
    if (i < 0) i = -i;
    if (i >= samples) {
        // right reflection covers `padding` samples; beyond that the oracle zero-pads
        const int beyond = i - (samples - 1);
        i = beyond <= padding ? samples - 1 - beyond : -1;
    }
    if (i >= 0 && i < samples) value = audio[(size_t)b * samples + i];
    out[(size_t)local * stride + n] = value;
}

// Reflect-padded utterances laid end to end: out[u * stride + s] = padded[first_item + u][s],
// the signal frames_kernel cuts its frames from (frame f is padded[hop f : hop f + 1024])
__global__ void __launch_bounds__(256) padded_kernel(
    const float* __restrict__ audio, float* __restrict__ out, int samples, int first_item,
    int padding, int stride) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= stride) return;
    const int b = first_item + blockIdx.y;
    int i = s - padding;
    float value = 0.f;
    if (i < 0) i = -i;
    if (i >= samples) {
        const int beyond = i - (samples - 1);
        i = beyond <= padding ? samples - 1 - beyond : -1;
    }
    if (i >= 0 && i < samples) value = audio[(size_t)b * samples + i];
    out[(size_t)blockIdx.y * stride + s] = value;
}

// MaxPool(2) (optional) + LayerNorm over (C, L) with elementwise affine, per frame.
// in: [C][in_row] with frame f at columns f * l_in .. f * l_in + l_conv (already ReLU'd)
// out: [C][count * l_out] or, transposed, [(c * l_out + t)][count]
__global__ void __launch_bounds__(256) pool_norm_kernel(
    const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
    float* __restrict__ out, int channels, int l_in, int l_out, bool pooled, size_t in_row,
    int count, bool transposed) {
    __shared__ double partial[2][8];
    __shared__ float stats[2];
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = channels * l_out;
    const float* base = in + (size_t)f * l_in;
    auto value = [&](int idx) {
        const int c = idx / l_out, t = idx % l_out;
        const float* row = base + (size_t)c * in_row;
        return pooled ? fmaxf(row[2 * t], row[2 * t + 1]) : row[t];
    };
    float sum = 0.f, squares = 0.f;
    for (int idx = tid; idx < total; idx += blockDim.x) {
        const float v = value(idx);
        sum += v;
        squares = fmaf(v, v, squares);
    }
    double dsum = sum, dsquares = squares;
    for (int offset = 16; offset > 0; offset >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, offset);
        dsquares += __shfl_xor_sync(0xffffffffu, dsquares, offset);
    }
    if ((tid & 31) == 0) { partial[0][tid >> 5] = dsum; partial[1][tid >> 5] = dsquares; }
    __syncthreads();
    if (tid == 0) {
        double s = 0., q = 0.;
        for (int w = 0; w < 8; ++w) { s += partial[0][w]; q += partial[1][w]; }
        const double mean = s / total;
        const double variance = fmax(q / total - mean * mean, 0.);
        stats[0] = (float)mean;
        stats[1] = (float)(1. / sqrt(variance + 1e-5));
    }
    __syncthreads();
    const float mean = stats[0], rstd = stats[1];
    for (int idx = tid; idx < total; idx += blockDim.x) {
        const float v = (value(idx) - mean) * rstd * weight[idx] + bias[idx];
        if (transposed) {
            out[(size_t)idx * count + f] = v;
        } else {
            const int c = idx / l_out, t = idx % l_out;
            out[(size_t)c * count * l_out + (size_t)f * l_out + t] = v;
        }
    }
}

// Same normalisation, written as the bf16 hi/lo planes the next tensor-core conv
// reads (conv1d_tc.cuh): planes[plane][c / 8][kTcPad + f * l_out + t][c % 8].
// Frames f >= count (padding up to a multiple of 16 frames) are written as zeros.
__global__ void __launch_bounds__(256) pool_norm_planes_kernel(
    const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
    __nv_bfloat16* __restrict__ planes, int channels, int l_in, int l_out, bool pooled, size_t in_row,
    int count, int t_pad, bool transposed, int stride_out, bool fold_out, bool fold_in, int frame_base,
    int frames_per_item = 0, int item_stride = 0, int first_frame = 0, bool frame_major = false) {
    __shared__ double partial[2][8];
    __shared__ float stats[2];
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int groups = channels / 8;
    const int g_begin = 0, g_end = groups;
    // fold_out: the operand of a convolution folded by 4 in time: sample t of channel group g
    // is row t / 4 of group (t % 4) * groups + g, and a frame has stride_out / 4 rows
    // frame_major: the operand of a frame-major layer, planes[plane][t groups + g][kTcPad + f][8]
    const int plane_groups = frame_major ? l_out * groups : fold_out ? 4 * groups : groups;
    uint4* hi_plane = reinterpret_cast<uint4*>(planes);
    uint4* lo_plane = hi_plane + (size_t)plane_groups * t_pad;
    // frame f of this launch is frame frame_base + f of the operand being written
    auto plane_row = [&](int g, int t) {
        const size_t frame = (size_t)(frame_base + f);
        if (frame_major) return ((size_t)t * groups + g) * t_pad + kTcPad + frame;
        if (fold_out)
            return (size_t)((t & 3) * groups + g) * t_pad + kTcPad + frame * (stride_out >> 2) + (t >> 2);
        return (size_t)g * t_pad + kTcPad + frame * stride_out + t;
    };
    // rows [l_out, stride_out) of a frame and whole frames >= count are zero padding
    for (int idx = tid; idx < (g_end - g_begin) * stride_out; idx += blockDim.x) {
        const int t = idx % stride_out;
        if (!transposed && (f >= count || t >= l_out)) {
            const size_t row = plane_row(g_begin + idx / stride_out, t);
            hi_plane[row] = make_uint4(0, 0, 0, 0);
            lo_plane[row] = make_uint4(0, 0, 0, 0);
        }
    }
    if (f >= count) return;
    const int total = channels * l_out;
    // frames_per_item > 0: `in` is shared by overlapping frames (block 0): frame first_frame + f
    // is frame (.) % frames_per_item of item (.) / frames_per_item, items item_stride columns
    // apart, frames l_in columns apart
    const float* base = frames_per_item > 0
        ? in + (size_t)((first_frame + f) / frames_per_item) * item_stride +
              (size_t)((first_frame + f) % frames_per_item) * l_in
        : in + (size_t)f * l_in;
    auto value = [&](int c, int t) {
        if (fold_in) {
            // output of a folded convolution, not yet pooled: sample 4 t4 + q of channel c is
            // row t4 of column q * channels + c; MaxPool(2) pairs q = 0 | 1 and q = 2 | 3
            const float* row = base + (size_t)((t & 1) * 2 * channels + c) * in_row + (t >> 1);
            return fmaxf(row[0], row[(size_t)channels * in_row]);
        }
        const float* row = base + (size_t)c * in_row;
        return pooled ? fmaxf(row[2 * t], row[2 * t + 1]) : row[t];
    };
    // statistics: the frame's (channel, sample) pairs flat over the threads, 8 independent loads in
    // flight per thread (a warp per channel row left 48 - 64 dependent load latencies in a row on
    // the short rows of blocks 3 - 5: 66, 35 and 4 samples)
    float sum = 0.f, squares = 0.f;
    {
        int idx = tid;
        for (; idx + 7 * 256 < total; idx += 8 * 256) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int flat = idx + i * 256;
                v[i] = value(flat / l_out, flat % l_out);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sum += v[i];
                squares = fmaf(v[i], v[i], squares);
            }
        }
        for (; idx < total; idx += 256) {
            const float v = value(idx / l_out, idx % l_out);
            sum += v;
            squares = fmaf(v, v, squares);
        }
    }
    double dsum = sum, dsquares = squares;
    for (int offset = 16; offset > 0; offset >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, offset);
        dsquares += __shfl_xor_sync(0xffffffffu, dsquares, offset);
    }
    if ((tid & 31) == 0) { partial[0][tid >> 5] = dsum; partial[1][tid >> 5] = dsquares; }
    __syncthreads();
    if (tid == 0) {
        double s = 0., q = 0.;
        for (int w = 0; w < 8; ++w) { s += partial[0][w]; q += partial[1][w]; }
        const double mean = s / total;
        const double variance = fmax(q / total - mean * mean, 0.);
        stats[0] = (float)mean;
        stats[1] = (float)(1. / sqrt(variance + 1e-5));
    }
    __syncthreads();
    const float mean = stats[0], rstd = stats[1];
    if (transposed) {
        // head operand: 2048 "channels" c * l_out + t, one row per frame
        const int wide = channels * l_out / 8;
        uint4* hi_wide = reinterpret_cast<uint4*>(planes);
        uint4* lo_wide = hi_wide + (size_t)wide * t_pad;
        for (int g = tid; g < wide; g += blockDim.x) {
            unsigned int hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int flat = g * 8 + 2 * e + h;
                    y[h] = (value(flat / l_out, flat % l_out) - mean) * rstd * weight[flat] + bias[flat];
                }
                tc::split_pair(y[0], y[1], hi[e], lo[e]);
            }
            hi_wide[(size_t)g * t_pad + kTcPad + f] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            lo_wide[(size_t)g * t_pad + kTcPad + f] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        return;
    }
    for (int idx = tid; idx < (g_end - g_begin) * l_out; idx += blockDim.x) {
        const int g = g_begin + idx / l_out, t = idx % l_out;
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = value(g * 8 + e, t);   // 8 independent loads
        unsigned int hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float z[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = g * 8 + 2 * e + h;
                z[h] = (y[2 * e + h] - mean) * rstd * weight[c * l_out + t] + bias[c * l_out + t];
            }
            tc::split_pair(z[0], z[1], hi[e], lo[e]);
        }
        const size_t row = plane_row(g, t);
        hi_plane[row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        lo_plane[row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ---- block 0 shared between frames: LayerNorm of frame windows of one pooled convolution ----

// sums[0][u] = sum_c in[c][u], sums[1][u] = sum_c in[c][u]^2 (one pass over the shared output)
__global__ void __launch_bounds__(256) column_sums_kernel(
    const float* __restrict__ in, float* __restrict__ sums, int channels, size_t in_row) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= in_row) return;
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, squares[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < channels; c += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = in[(size_t)(c + i) * in_row + u];
            sum[i] += v;
            squares[i] = fmaf(v, v, squares[i]);
        }
    }
    sums[u] = (sum[0] + sum[1]) + (sum[2] + sum[3]);
    sums[in_row + u] = (squares[0] + squares[1]) + (squares[2] + squares[3]);
}

// column offset of frame `frame` (counted from the first frame of the group) in the shared output
__device__ __forceinline__ size_t shared_frame_offset(int frame, int frames_per_item, int item_stride, int l_in) {
    return (size_t)(frame / frames_per_item) * item_stride + (size_t)(frame % frames_per_item) * l_in;
}

// mean and 1 / std of every frame's (channels x l_out) window, one warp per frame
__global__ void __launch_bounds__(256) frame_stats_kernel(
    const float* __restrict__ sums, size_t in_row, float2* __restrict__ stats, int count,
    int frames_per_item, int item_stride, int l_in, int l_out, int first_frame, int channels) {
    const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= count) return;
    const float* column = sums + shared_frame_offset(first_frame + f, frames_per_item, item_stride, l_in);
    double sum = 0., squares = 0.;
    for (int t = lane; t < l_out; t += 32) {
        sum += column[t];
        squares += column[in_row + t];
    }
    for (int offset = 16; offset > 0; offset >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, offset);
        squares += __shfl_xor_sync(0xffffffffu, squares, offset);
    }
    if (lane == 0) {
        const double total = (double)channels * l_out;
        const double mean = sum / total;
        const double variance = fmax(squares / total - mean * mean, 0.);
        stats[f] = make_float2((float)mean, (float)(1. / sqrt(variance + 1e-5)));
    }
}

// (x - mean_f) rstd_f weight[c, t] + bias[c, t] of kSharedFrames consecutive frames x 8 channels
// per CTA, written as the hi / lo planes of block 1's operand folded by 4 in time.  The frames'
// windows overlap (l_in = hop / 2 columns apart, l_out wide), so the CTA stages their union and
// the 8 channels' affine parameters in shared memory once: 0.46 instead of 3 loads per output.
constexpr int kSharedFrames = 8;
// F8: the operand in the "fp16 + 2 x fp8" format (conv1d_tc.cuh), 16 channels per CTA: per sample
// two fp16 rows (8 channels each), one e4m3 row of z s_x and one of (z - fp16 z) s_xl.  With twice the
// shared memory per CTA it runs 512 threads, two per sample on alternate frames, so that as many
// threads stay resident (at 256 it was 1.6 x slower; 8 channels per thread writing half e4m3 rows 2 x)
template <bool F8>
__global__ void __launch_bounds__(F8 ? 512 : 256, F8 ? 2 : 5) shared_norm_planes_kernel(
    const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
    const float2* __restrict__ stats, __nv_bfloat16* __restrict__ planes, int channels, int l_in,
    int l_out, size_t in_row, int count, int t_pad, int stride_out, int frames_per_item,
    int item_stride, int first_frame) {
    constexpr int kC = F8 ? 16 : 8;               // channels per CTA
    extern __shared__ float shared_norm_smem[];
    const int width = l_out + (kSharedFrames - 1) * l_in;
    float* tile = shared_norm_smem;               // [kC][width]
    __shared__ float2 frame_stats[kSharedFrames];
    const int f0 = blockIdx.x * kSharedFrames, g = blockIdx.y;
    const int groups = channels / 8;
    const int n_frames = min(kSharedFrames, count - f0);
    const int item0 = (first_frame + f0) / frames_per_item;
    // frames of this CTA that lie in the first frame's utterance share the staged window
    int same = 1;
    while (same < n_frames && (first_frame + f0 + same) / frames_per_item == item0) ++same;
    const size_t offset0 = shared_frame_offset(first_frame + f0, frames_per_item, item_stride, l_in);
    const int needed = (same - 1) * l_in + l_out;
    const float* source = in + (size_t)(g * kC) * in_row + offset0;
    if (F8) {
        // two CTAs of 64 registers per SM; every channel's load of a position is issued before the
        // first store waits for one (with 128 registers and one CTA the staging waited on its loads)
        for (int pos = threadIdx.x; pos < needed; pos += blockDim.x) {
            float staged[kC];
            const float* from = source + pos;
#pragma unroll
            for (int c = 0; c < kC; ++c, from += in_row) staged[c] = *from;
#pragma unroll
            for (int c = 0; c < kC; ++c) tile[c * width + pos] = staged[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < kC; ++c)
            for (int pos = threadIdx.x; pos < needed; pos += blockDim.x)
                tile[c * width + pos] = source[(size_t)c * in_row + pos];
    }
    if (threadIdx.x < n_frames) frame_stats[threadIdx.x] = stats[f0 + threadIdx.x];
    __syncthreads();
    uint4* rows16 = reinterpret_cast<uint4*>(planes);
    const int step = stride_out >> 2;             // rows between consecutive frames
    // a thread owns sample t of every (F8: every second) frame: the affine parameters of
    // (kC channels, t) stay in registers
    const int j_first = threadIdx.x / 256, j_step = blockDim.x / 256;
    for (int t = threadIdx.x % 256; t < stride_out; t += 256) {
        // sample t of an 8-channel group g8 is row t / 4 of group (t % 4) * groups + g8 (pool_norm_planes_kernel);
        // row groups of this thread's rows: bf16: hi, lo; F8: fp16 a, fp16 b, e4m3 z, e4m3 low
        size_t group[4];
        if (F8) {
            group[0] = (size_t)(t & 3) * groups + 2 * g;
            group[1] = group[0] + 1;
            group[2] = (size_t)4 * groups + (size_t)(t & 3) * (groups / 2) + g;
            group[3] = group[2] + (size_t)4 * (groups / 2);
        } else {
            group[0] = (size_t)(t & 3) * groups + g;
            group[1] = group[0] + (size_t)4 * groups;
            group[2] = group[3] = 0;
        }
        const size_t within = kTcPad + (size_t)f0 * step + (t >> 2);
        uint4* target[4];                             // this thread's rows of frame f0
#pragma unroll
        for (int r = 0; r < 4; ++r) target[r] = rows16 + group[r] * t_pad + within;
        if (t >= l_out) {
            for (int j = j_first; j < n_frames; j += j_step)
                for (int r = 0; r < (F8 ? 4 : 2); ++r) target[r][j * step] = make_uint4(0, 0, 0, 0);
            continue;
        }
        float gamma[kC], beta[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            gamma[c] = weight[(size_t)(g * kC + c) * l_out + t];
            beta[c] = bias[(size_t)(g * kC + c) * l_out + t];
        }
        for (int j = j_first; j < n_frames; j += j_step) {
            const float2 st = frame_stats[j];
            float x[kC];
            if (j < same) {
#pragma unroll
                for (int c = 0; c < kC; ++c) x[c] = tile[c * width + j * l_in + t];
            } else {
                const float* global = in + (size_t)(g * kC) * in_row + t +
                    shared_frame_offset(first_frame + f0 + j, frames_per_item, item_stride, l_in);
#pragma unroll
                for (int c = 0; c < kC; ++c) x[c] = global[(size_t)c * in_row];
            }
#pragma unroll
            for (int c = 0; c < kC; ++c) x[c] = (x[c] - st.x) * st.y * gamma[c] + beta[c];
            const int row = j * step;
            if (F8) {
                unsigned int half_words[8], z8[4], low8[4];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    unsigned int z, low;
                    tc::split_pair_f8(x[2 * e], x[2 * e + 1], half_words[e], z, low);
                    if (e % 2 == 0) { z8[e / 2] = z; low8[e / 2] = low; }
                    else { z8[e / 2] |= z << 16; low8[e / 2] |= low << 16; }
                }
                target[0][row] = make_uint4(half_words[0], half_words[1], half_words[2], half_words[3]);
                target[1][row] = make_uint4(half_words[4], half_words[5], half_words[6], half_words[7]);
                target[2][row] = make_uint4(z8[0], z8[1], z8[2], z8[3]);
                target[3][row] = make_uint4(low8[0], low8[1], low8[2], low8[3]);
            } else {
                unsigned int hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) tc::split_pair(x[2 * e], x[2 * e + 1], hi[e], lo[e]);
                target[0][row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                target[1][row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

// ---- frame-major layers (blocks 3 - 5 and the head on the tensor-core path) ----
//
// Laid end to end on one time axis, a frame of L rows yields L - 31 valid outputs of a k = 32
// convolution and 31 that straddle two frames: 32 % of block 3's rows, 47 % of block 4's and (in
// its 16 frames x 8 rows tiles) 50 % of block 5's were computed and dropped.  These layers
// therefore keep their activations FRAME-MAJOR, planes[plane][position][c / 8][frame][8]: output
// position t of every frame is then one dense product over K = (tap j, channel c) whose operand,
// position t + j, is the contiguous group range [t C / 8, (t + 32) C / 8) -- a k = 1 launch of
// conv1d_tc_kernel with the batch index as the position (TcConvArgs::item_groups) and frames as
// rows, no row wasted.

// (C_out, C_in, K) -> (C_out, K C_in), column j C_in + c
__global__ void tap_major_weight_kernel(
    const float* __restrict__ w, float* __restrict__ out, int c_out, int c_in, int k) {
    const size_t total = (size_t)c_out * c_in * k;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / ((size_t)c_in * k));
        const int rest = (int)(idx % ((size_t)c_in * k));
        const int j = rest / c_in, c = rest % c_in;
        out[idx] = w[((size_t)o * c_in + c) * k + j];
    }
}

constexpr int kStatSlices = 16;

// in: (positions, channels, frames) = `rows` rows of in_row floats.  partial[(slice count + f) 2 + i]:
// sum (i = 0) and sum of squares (i = 1) of frame f over the rows slice, slice + 16, ...
__global__ void __launch_bounds__(256) frame_major_stats_kernel(
    const float* __restrict__ in, double* __restrict__ partial, int rows, size_t in_row, int count) {
    __shared__ float shared[2][8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * 32 + lane, slice = blockIdx.y;
    const bool live = f < count;
    const int step = kStatSlices * 8;
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, squares[4] = {0.f, 0.f, 0.f, 0.f};
    int r = slice + kStatSlices * warp;
    const float* column = in + (live ? f : 0);
    for (; r + 3 * step < rows; r += 4 * step) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = column[(size_t)(r + i * step) * in_row];
#pragma unroll
        for (int i = 0; i < 4; ++i) { sum[i] += v[i]; squares[i] = fmaf(v[i], v[i], squares[i]); }
    }
    for (; r < rows; r += step) {
        const float v = column[(size_t)r * in_row];
        sum[0] += v;
        squares[0] = fmaf(v, v, squares[0]);
    }
    shared[0][warp][lane] = (sum[0] + sum[1]) + (sum[2] + sum[3]);
    shared[1][warp][lane] = (squares[0] + squares[1]) + (squares[2] + squares[3]);
    __syncthreads();
    if (warp < 2 && live) {
        double total = 0.;
        for (int w = 0; w < 8; ++w) total += shared[warp][w][lane];
        partial[((size_t)slice * count + f) * 2 + warp] = total;
    }
}

// LayerNorm over (channels, positions) of every frame, written as the next layer's frame-major
// operand planes[plane][t groups + g][kTcPad + f][8]; lanes are frames: every load and store of a
// warp is one contiguous run
__global__ void __launch_bounds__(256) frame_major_norm_planes_kernel(
    const float* __restrict__ in, const double* __restrict__ partial, const float* __restrict__ weight,
    const float* __restrict__ bias, __nv_bfloat16* __restrict__ planes, int channels, int positions,
    size_t in_row, int count, int t_pad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * 32 + lane;
    if (f >= count) return;
    double sum = 0., squares = 0.;
    for (int slice = 0; slice < kStatSlices; ++slice) {
        sum += partial[((size_t)slice * count + f) * 2];
        squares += partial[((size_t)slice * count + f) * 2 + 1];
    }
    const double total = (double)channels * positions;
    const double mean_d = sum / total;
    const float mean = (float)mean_d;
    const float rstd = (float)(1. / sqrt(fmax(squares / total - mean_d * mean_d, 0.) + 1e-5));
    const int groups = channels / 8, work = positions * groups;
    uint4* hi_plane = reinterpret_cast<uint4*>(planes);
    uint4* lo_plane = hi_plane + (size_t)work * t_pad;
    for (int idx = blockIdx.y * 8 + warp; idx < work; idx += gridDim.y * 8) {
        const int t = idx / groups, g = idx % groups;
        const float* source = in + ((size_t)t * channels + g * 8) * in_row + f;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = source[(size_t)e * in_row];
        unsigned int hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = g * 8 + 2 * e;
            tc::split_pair(
                (x[2 * e] - mean) * rstd * weight[(size_t)c * positions + t] + bias[(size_t)c * positions + t],
                (x[2 * e + 1] - mean) * rstd * weight[(size_t)(c + 1) * positions + t] +
                    bias[(size_t)(c + 1) * positions + t],
                hi[e], lo[e]);
        }
        const size_t row = (size_t)idx * t_pad + kTcPad + f;
        hi_plane[row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        lo_plane[row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Block 1's weight (32, 256, 32) as the weight (128, 1024, 9) of the convolution folded by 4
// in time: W4[(q, o), (p, c), J] = W[o, c, 4 J + p - q], zero where that tap does not exist
__global__ void fold_weight_kernel(const float* __restrict__ w, float* __restrict__ folded) {
    constexpr int kOut = 32, kIn = 256, kFold = 4, kTaps = (kKernel + kFold - 1) / kFold + 1;
    const int total = kFold * kOut * kFold * kIn * kTaps;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int rest = idx;
        const int J = rest % kTaps; rest /= kTaps;
        const int c = rest % kIn; rest /= kIn;
        const int pp = rest % kFold; rest /= kFold;
        const int o = rest % kOut; rest /= kOut;
        const int q = rest;
        const int j = kFold * J + pp - q;
        folded[idx] = (j >= 0 && j < kKernel) ? w[((size_t)o * kIn + c) * kKernel + j] : 0.f;
    }
}

// Block 0 on the tensor cores: its 32 taps become 32 "channels" of a 1x1 conv.
// planes[plane][tap / 8][kTcPad + r][tap % 8] = x[r + tap] over the frames laid
// end to end (x is the cropped-frame buffer), r < rows; pad rows are zero.
__global__ void __launch_bounds__(128) im2col_planes_kernel(
    const float* __restrict__ x, __nv_bfloat16* __restrict__ planes, int samples, int rows, int t_pad) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= t_pad) return;
    const int g = blockIdx.y;
    const int r = row - kTcPad;
    unsigned int hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    if (r >= 0 && r < rows) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float y[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = r + g * 8 + 2 * e + h;
                y[h] = i < samples ? x[i] : 0.f;
            }
            tc::split_pair(y[0], y[1], hi[e], lo[e]);
        }
    }
    uint4* hi_plane = reinterpret_cast<uint4*>(planes);
    uint4* lo_plane = hi_plane + (size_t)4 * t_pad;
    hi_plane[(size_t)g * t_pad + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    lo_plane[(size_t)g * t_pad + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// logits^T (1440, count) -> masked logits, softmax and entropy periodicity, 32 frames per CTA
__global__ void __launch_bounds__(256) posterior_kernel(
    const float* __restrict__ logits_t, int count, int min_bin, int max_bin,
    float* __restrict__ masked, float* __restrict__ distribution, float* __restrict__ periodicity) {
    __shared__ float tile[32][33];
    __shared__ float reduce[8][32];
    __shared__ float row_max[32], row_sum[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f0 = blockIdx.x * 32;
    const int f = f0 + tx;
    const bool live = f < count;
    auto load = [&](int o) {
        if (!live || o < min_bin || o >= max_bin) return -INFINITY;
        return logits_t[(size_t)o * count + f];
    };
    float best = -INFINITY;
    for (int o = ty; o < kBins; o += 8) best = fmaxf(best, load(o));
    reduce[ty][tx] = best;
    __syncthreads();
    if (ty == 0) {
        for (int w = 1; w < 8; ++w) best = fmaxf(best, reduce[w][tx]);
        row_max[tx] = best;
    }
    __syncthreads();
    const float maximum = row_max[tx];
    float sum = 0.f;
    for (int o = ty; o < kBins; o += 8) sum += expf(load(o) - maximum);
    __syncthreads();
    reduce[ty][tx] = sum;
    __syncthreads();
    if (ty == 0) {
        for (int w = 1; w < 8; ++w) sum += reduce[w][tx];
        row_sum[tx] = sum;
    }
    __syncthreads();
    const float inverse = 1.f / row_sum[tx];
    float entropy = 0.f;
    for (int o0 = 0; o0 < kBins; o0 += 32) {
        // logits pass through the tile first, then probabilities: both are written
        // row-major (frame, bin) with the bins of a frame contiguous
        float p[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int o = o0 + ty + 8 * i;
            l[i] = load(o);
            p[i] = live ? expf(l[i] - maximum) * inverse : 0.f;
            entropy += p[i] * logf(p[i] + 1e-7f);
            tile[ty + 8 * i][tx] = l[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = f0 + ty + 8 * i;
            if (row < count) masked[(size_t)row * kBins + o0 + tx] = tile[tx][ty + 8 * i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = p[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = f0 + ty + 8 * i;
            if (row < count) distribution[(size_t)row * kBins + o0 + tx] = tile[tx][ty + 8 * i];
        }
        __syncthreads();
    }
    reduce[ty][tx] = entropy;
    __syncthreads();
    if (ty == 0 && live) {
        for (int w = 1; w < 8; ++w) entropy += reduce[w][tx];
        periodicity[f] = 1.f + entropy / logf((float)kBins);
    }
}

// Local expected value around the decoded bin -> Hz
__global__ void __launch_bounds__(128) pitch_kernel(
    const float* __restrict__ masked, const int* __restrict__ bins, float* __restrict__ pitch, int count) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= count) return;
    const int centre = bins[f];
    const float* row = masked + (size_t)f * kBins;
    float values[kLocalWindow];
    float best = -INFINITY;
#pragma unroll
    for (int w = 0; w < kLocalWindow; ++w) {
        const int i = centre - kLocalWindow / 2 + w;
        values[w] = (i >= 0 && i < kBins) ? row[i] : -INFINITY;
        best = fmaxf(best, values[w]);
    }
    float sum = 0.f, expected = 0.f;
#pragma unroll
    for (int w = 0; w < kLocalWindow; ++w) {
        const float e = expf(values[w] - best);
        sum += e;
        expected += e * (kCentsPerBin * (float)(centre - kLocalWindow / 2 + w));
    }
    pitch[f] = kFmin * exp2f(expected / sum / kOctave);
}

int find(const pmn_pitch* p, const std::string& name, const Tensor** out) {
    return p->store.find(name, out);
}

int alloc(pmn_pitch* p, size_t count, float** out) { return p->store.alloc(count, out); }

// torchaudio.functional.resample kernel bank (sinc_interp_hann, width 6, rolloff 0.99),
// stored transposed: table[k][phase]
int resampler(pmn_pitch* p, int sample_rate, const pmn_pitch::Resampler** out) {
    auto it = p->resamplers.find(sample_rate);
    if (it == p->resamplers.end()) {
        const int g = std::gcd(sample_rate, kRate);
        const int orig = sample_rate / g, fresh = kRate / g;
        const double base = (double)std::min(orig, fresh) * 0.99;
        const int width = (int)ceil(6. * orig / base);
        const int taps = 2 * width + orig;
        const double pi = 3.14159265358979323846;
        std::vector<float> table((size_t)taps * fresh);
        for (int phase = 0; phase < fresh; ++phase) {
            const double shift = (double)((float)(-phase) / (float)fresh);  // fp32 like torch.arange(...) / new_freq
            for (int k = 0; k < taps; ++k) {
                double t = (shift + (double)(k - width) / orig) * base;
                t = std::max(-6., std::min(6., t));
                const double window = cos(t * pi / 6. / 2.);
                t *= pi;
                const double sinc = t == 0. ? 1. : sin(t) / t;
                table[(size_t)k * fresh + phase] = (float)(sinc * window * window * (base / orig));
            }
        }
        pmn_pitch::Resampler r{nullptr, orig, fresh, width};
        PMN_TRY(check_cuda(cudaMalloc(&r.table, table.size() * sizeof(float)), "cudaMalloc resampler"));
        PMN_TRY(check_cuda(
            cudaMemcpy(r.table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice),
            "upload resampler"));
        it = p->resamplers.emplace(sample_rate, r).first;
    }
    *out = &it->second;
    return PMN_OK;
}

struct Workspace {
    float *resampled, *conv, *act, *logits_t, *masked, *distribution, *pooled, *sums, *padded;
    float2* stats;
    double* partial;
    __nv_bfloat16 *planes, *planes0;
    int* bins;
    void* viterbi;
    size_t viterbi_bytes, bytes;
};

// Rows of one reflect-padded utterance on the shared block-0 time axis (even, so MaxPool pairs
// never straddle two utterances) and utterances per group; 0 when the hop is odd (frame starts
// then fall on both parities of the pooling grid and block 0 runs frame by frame)
int shared_item_rows(int frames, int hop) {
    if (hop % 2) return 0;
    return (hop * (frames - 1) + kCropStart + kCropped + 1) / 2 * 2;
}
int shared_group(int batch, int item_rows) {
    return std::max(1, std::min(batch, 2 * kPooledRowsCap / item_rows));
}

Workspace carve(void* base, int batch, int out_samples, int frames, int frame_batch, int hop) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += align_up(bytes, 256);
        return r;
    };
    const size_t total = (size_t)batch * frames;
    const size_t fb = (size_t)std::min<size_t>(frame_batch, total);
    // block 0 frame by frame (the fp32 path, and the tensor-core path when the hop is odd) never
    // takes more than kSubFrames frames at a time, whatever the frame batch
    const size_t fb0 = std::min<size_t>(fb, kSubFrames);
    w.resampled = (float*)take((size_t)batch * out_samples * 4);
    // largest conv output: block 0 frame by frame, else block 1 (128 columns x 121 rows per frame)
    w.conv = (float*)take(std::max(fb0 * 256 * (kCropped + 1), fb * 128 * (kFoldStride / kFold)) * 4);
    w.act = (float*)take(fb0 * 256 * 481 * 4);                       // largest block output / input (fp32 path)
    // tensor-core operand planes: the widest is block 0's output (256 channels x 481 rows per frame)
    // (the frame-major layers' operands have one row per frame, padded to the tile, per
    // (position, channel group): block 5's 35 x 256 channels are the most)
    w.planes = (__nv_bfloat16*)take(std::max(std::max(
        tc_planes_elements(1, 256, (int)(fb + 16) * 482),
        tc_planes_elements(1, kFold * 256, (int)(fb + 16) * (kFoldStride / kFold))),
        tc_planes_elements(1, kLength[5] * kChannels[5], (int)fb)) * 2);
    // block 0's own operand (im2col rows of a sub-chunk of frames, 32 "channels")
    const int item_rows = shared_item_rows(frames, hop);
    const size_t shared_rows = item_rows ? (size_t)shared_group(batch, item_rows) * item_rows : 0;
    w.planes0 = (__nv_bfloat16*)take(std::max(
        tc_planes_elements(1, kKernel, (int)std::min<size_t>(kSubFrames, fb) * (kCropped + 1)),
        tc_planes_elements(1, kKernel, (int)shared_rows)) * 2);
    // block 0's pooled output over whole utterances (shared by their frames), plus the padded signal
    w.pooled = (float*)take((shared_rows / 2 + 64) * 256 * 4);
    w.padded = (float*)take(shared_rows * 4);
    w.sums = (float*)take(shared_rows * 4);          // 2 x (shared_rows / 2) column sums
    w.stats = (float2*)take(fb * sizeof(float2));    // mean, 1 / std per frame of a chunk
    w.partial = (double*)take(fb * kStatSlices * 2 * sizeof(double));   // frame-major LayerNorm sums
    w.logits_t = (float*)take(fb * kBins * 4);
    w.masked = (float*)take(total * kBins * 4);
    w.distribution = (float*)take(total * kBins * 4);
    w.bins = (int*)take(total * 4);
    w.viterbi_bytes = viterbi_workspace_bytes(batch, frames, kBins);
    w.viterbi = take(w.viterbi_bytes);
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}

}  // namespace

pmn_pitch* pitch_create() { return new (std::nothrow) pmn_pitch(); }
void pitch_destroy(pmn_pitch* p) { delete p; }

int pitch_set_tensor(pmn_pitch* p, const char* name, const float* data, const int64_t* shape,
                     int ndim, cudaStream_t stream) {
    if (p->finalized) return fail(PMN_ERR_STATE, "set_tensor after finalize");
    return p->store.set(name, data, shape, ndim, stream);
}

int pitch_finalize(pmn_pitch* p, int math, cudaStream_t stream) {
    if (p->finalized) return fail(PMN_ERR_STATE, "pitch model already finalized");
    if (math != PMN_MATH_FP32_SIMT && math != PMN_MATH_BF16X3_TC)
        return fail(PMN_ERR_ARGUMENT, "pitch: unsupported math mode");
    p->math = math;
    for (int i = 0; i < kLayers; ++i) {
        const std::string prefix = "layers." + std::to_string(i);
        const Tensor *w, *b, *nw, *nb;
        PMN_TRY(find(p, prefix + ".conv.weight", &w));
        PMN_TRY(find(p, prefix + ".conv.bias", &b));
        PMN_TRY(find(p, prefix + ".norm.weight", &nw));
        PMN_TRY(find(p, prefix + ".norm.bias", &nb));
        if (w->numel() != (size_t)kChannels[i + 1] * kChannels[i] * kKernel ||
            nw->numel() != (size_t)kChannels[i + 1] * kLength[i + 1] || nb->numel() != nw->numel())
            return fail(PMN_ERR_STATE, "unexpected shapes at " + prefix);
        if (math == PMN_MATH_BF16X3_TC) {
            float* slabs;  // bf16 hi + lo = the bytes of the fp32 tensor
            PMN_TRY(alloc(p, w->numel(), &slabs));
            p->conv_slabs[i] = reinterpret_cast<__nv_bfloat16*>(slabs);
            if (i == 1) {
                // folded by 4: (32, 256, 32) -> (128, 1024, 9), 12.5 % zeros
                const size_t folded_numel = (size_t)kFold * 32 * kFold * 256 * kFoldTaps;
                float *folded, *folded_slabs;
                PMN_TRY(alloc(p, folded_numel, &folded));
                PMN_TRY(alloc(p, folded_numel, &folded_slabs));
                {
                    LaunchScope scope("fold_weight_kernel", stream);
                    fold_weight_kernel<<<1024, 256, 0, stream>>>(w->data, folded);
                    PMN_TRY(launched("fold_weight_kernel"));
                }
                p->conv_slabs[i] = reinterpret_cast<__nv_bfloat16*>(folded_slabs);
                PMN_TRY(launch_pack_tc_weight(
                    folded, p->conv_slabs[i], kFold * 32, kFold * 256, kFoldTaps, false, stream));
                {
                    // "fp16 + 2 x fp8": block 1's input is a LayerNorm output (|x| of a few units), its
                    // weights are scaled by the power of two that brings the largest to (128, 256]
                    // The default (PMN_PITCH_F8=0: bf16 x 3): measured on 32 x 10 s the convolution gains
                    // 2.6 ms and the kernel that writes its four-stream operand loses 0.4 (two CTAs of
                    // 64 registers per SM; at 128 registers it lost 3.8), profiles/r2_preprocess_history.txt
                    const char* flag = getenv("PMN_PITCH_F8");
                    if (!(flag && flag[0] == '0')) {
                        int shift;
                        PMN_TRY(tc_f8_weight_shift_of(w->data, w->numel(), stream, &shift));
                        float* f8_slabs;
                        PMN_TRY(alloc(p, folded_numel, &f8_slabs));
                        PMN_TRY(launch_pack_tc_weight_f8(
                            folded, f8_slabs, kFold * 32, kFold * 256, kFoldTaps, shift, stream));
                        p->block1_f8_slabs = f8_slabs;
                        p->block1_shift = shift;
                    }
                }
                // the bias of column (q, o) is the bias of channel o
                float* bias4;
                PMN_TRY(alloc(p, kFold * 32, &bias4));
                for (int q = 0; q < kFold; ++q)
                    PMN_TRY(check_cuda(
                        cudaMemcpyAsync(bias4 + q * 32, b->data, 32 * sizeof(float),
                                        cudaMemcpyDeviceToDevice, stream), "fold bias"));
                p->folded_bias = bias4;
                p->conv_bias[i] = b->data;
                p->norm_weight[i] = nw->data;
                p->norm_bias[i] = nb->data;
                continue;
            }
            if (i >= kFrameMajorFirst) {
                // frame-major layer: (C_out, C_in, 32) -> (C_out, 32 C_in) with K = (tap, channel)
                float* tap_major;
                PMN_TRY(alloc(p, w->numel(), &tap_major));
                {
                    LaunchScope scope("tap_major_weight_kernel", stream);
                    tap_major_weight_kernel<<<1024, 256, 0, stream>>>(
                        w->data, tap_major, kChannels[i + 1], kChannels[i], kKernel);
                    PMN_TRY(launched("tap_major_weight_kernel"));
                }
                PMN_TRY(launch_pack_tc_weight(
                    tap_major, p->conv_slabs[i], kChannels[i + 1], kKernel * kChannels[i], 1, false, stream));
                p->conv_bias[i] = b->data;
                p->norm_weight[i] = nw->data;
                p->norm_bias[i] = nb->data;
                continue;
            }
            // block 0: (256, 1, 32) is read as a (256, 32, 1) 1x1 conv over the 32 taps
            PMN_TRY(launch_pack_tc_weight(
                w->data, p->conv_slabs[i], kChannels[i + 1], i == 0 ? kKernel : kChannels[i],
                i == 0 ? 1 : kKernel, i == kLayers - 1, stream));
        } else {
            PMN_TRY(alloc(p, w->numel(), &p->conv_weight[i]));
            PMN_TRY(launch_pack_conv1d_weight(
                w->data, p->conv_weight[i], kChannels[i + 1], kChannels[i], kKernel, stream));
        }
        p->conv_bias[i] = b->data;
        p->norm_weight[i] = nw->data;
        p->norm_bias[i] = nb->data;
    }
    const Tensor *w, *b;
    PMN_TRY(find(p, "layers.6.weight", &w));
    PMN_TRY(find(p, "layers.6.bias", &b));
    if (w->numel() != (size_t)kBins * 512 * 4) return fail(PMN_ERR_STATE, "unexpected layers.6 shape");
    // (1440, 512, 4) read as (1440, 2048, 1): input index c * 4 + t matches the transposed activations
    if (math == PMN_MATH_BF16X3_TC) {
        float* slabs;
        PMN_TRY(alloc(p, w->numel(), &slabs));
        p->head_slabs = reinterpret_cast<__nv_bfloat16*>(slabs);
        // the head reads block 5's frame-major output: K = (position t, channel c)
        float* tap_major;
        PMN_TRY(alloc(p, w->numel(), &tap_major));
        {
            LaunchScope scope("tap_major_weight_kernel", stream);
            tap_major_weight_kernel<<<1024, 256, 0, stream>>>(w->data, tap_major, kBins, 512, 4);
            PMN_TRY(launched("tap_major_weight_kernel"));
        }
        PMN_TRY(launch_pack_tc_weight(tap_major, p->head_slabs, kBins, 2048, 1, false, stream));
    } else {
        PMN_TRY(alloc(p, w->numel(), &p->head_weight));
        PMN_TRY(launch_pack_conv1d_weight(w->data, p->head_weight, kBins, 2048, 1, stream));
    }
    p->head_bias = b->data;
    PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "finalize sync"));
    p->finalized = true;
    return PMN_OK;
}

int pitch_frames(int samples, int sample_rate, double hopsize_seconds) {
    const int frames = (int)((double)samples / (hopsize_seconds * sample_rate));
    return frames < 1 ? 1 : frames;
}

static int resampled_length(int samples, int sample_rate) {
    const int g = std::gcd(sample_rate, kRate);
    const long long fresh = kRate / g, orig = sample_rate / g;
    return (int)((fresh * samples + orig - 1) / orig);
}

size_t pitch_workspace_bytes(int batch, int samples, int sample_rate, double hopsize_seconds, int frame_batch) {
    return carve(nullptr, batch, resampled_length(samples, sample_rate),
                 pitch_frames(samples, sample_rate, hopsize_seconds), frame_batch,
                 (int)(hopsize_seconds * kRate)).bytes;
}

int pitch_forward(
    pmn_pitch* p, const float* audio, int batch, int samples, int sample_rate,
    double hopsize_seconds, float fmin, float fmax, const float* transition, const float* initial,
    float* pitch, float* periodicity, float* logits_out, int* bins_out, int frame_batch,
    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!p->finalized) return fail(PMN_ERR_STATE, "pitch model not finalized");
    PMN_REQUIRE(audio && pitch && periodicity && transition && initial && workspace, "pitch: null pointer");
    PMN_REQUIRE(batch > 0 && batch <= 65535 && samples > 0 && sample_rate > 0 && frame_batch > 0,
                "pitch: bad shape");
    const int hop = (int)(hopsize_seconds * kRate);
    PMN_REQUIRE(hop > 0 && hop < kWindow, "pitch: bad hopsize");
    const int padding = (kWindow - hop) / 2;
    const int out_samples = resampled_length(samples, sample_rate);
    PMN_REQUIRE(out_samples > padding, "pitch: audio shorter than the reflect padding");
    const int frames = pitch_frames(samples, sample_rate, hopsize_seconds);
    const int total = batch * frames;
    Workspace w = carve(workspace, batch, out_samples, frames, frame_batch, hop);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "pitch: workspace too small");

    // 1. resample to 8 kHz
    const float* audio8k = audio;
    if (sample_rate != kRate) {
        const pmn_pitch::Resampler* r;
        PMN_TRY(resampler(p, sample_rate, &r));
        dim3 grid(ceil_div(out_samples, 256), batch);
        LaunchScope scope("resample_kernel", stream);
        resample_kernel<<<grid, 256, 0, stream>>>(
            audio, r->table, w.resampled, samples, out_samples, r->orig, r->fresh, r->width);
        PMN_TRY(launched("resample_kernel"));
        audio8k = w.resampled;
    }

    // penn.convert.frequency_to_bins: floor / ceil of 1200 log2(f / 31) / 5
    const int min_bin = std::max(0, (int)floor(kOctave * log2((double)fmin / kFmin) / kCentsPerBin));
    const int max_bin = std::min(kBins, (int)ceil(kOctave * log2((double)fmax / kFmin) / kCentsPerBin));

    // 2. network, frame_batch frames at a time; on the tensor-core path with an even hop block 0's
    // convolution runs once per group of utterances before the group's frames (kPooledRowsCap)
    const bool tensor_cores = p->math == PMN_MATH_BF16X3_TC;
    const int item_rows = tensor_cores ? shared_item_rows(frames, hop) : 0;
    const bool shared0 = item_rows > 0;
    const int group = shared0 ? shared_group(batch, item_rows) : batch;
    size_t pooled_row = 0;     // row length of w.pooled
    for (int item0 = 0; item0 < batch; item0 += group) {
    const int items = std::min(group, batch - item0);
    if (shared0) {
        float* padded = w.padded;
        {
            dim3 grid(ceil_div(item_rows, 256), items);
            LaunchScope scope("padded_kernel", stream);
            padded_kernel<<<grid, 256, 0, stream>>>(
                audio8k, padded, out_samples, item0, padding, item_rows);
            PMN_TRY(launched("padded_kernel"));
        }
        const size_t rows_in = (size_t)items * item_rows;
        const size_t conv_rows = rows_in - (kKernel - 1);
        const int t_pad = tc_padded_length((int)conv_rows);
        {
            dim3 grid(ceil_div(t_pad, 128), 4);
            LaunchScope scope("im2col_planes_kernel", stream);
            im2col_planes_kernel<<<grid, 128, 0, stream>>>(
                padded, w.planes0, (int)rows_in, (int)conv_rows, t_pad);
            PMN_TRY(launched("im2col_planes_kernel"));
        }
        pooled_row = conv_rows / 2;
        TcConvArgs a;
        a.x_planes = w.planes0; a.w_slabs = p->conv_slabs[0]; a.bias = p->conv_bias[0];
        a.out = w.pooled; a.batch = 1; a.c_in = kKernel; a.c_out = kChannels[1]; a.k = 1;
        a.t_len = (int)conv_rows; a.valid = true; a.relu = true; a.pool = true;
        a.out_row = (int)pooled_row;
        PMN_TRY(launch_conv1d_tc(a, stream));
        LaunchScope scope("column_sums_kernel", stream);
        column_sums_kernel<<<(unsigned)((pooled_row + 255) / 256), 256, 0, stream>>>(
            w.pooled, w.sums, kChannels[1], pooled_row);
        PMN_TRY(launched("column_sums_kernel"));
    }
    const int group_end = (item0 + items) * frames;
    const int chunk = tensor_cores ? frame_batch : std::min(frame_batch, kSubFrames);
    for (int first = item0 * frames; first < group_end; first += chunk) {
        const int count = std::min(chunk, group_end - first);
        // Tensor-core path: frame strides are even where a MaxPool follows, so pooling
        // pairs never straddle two frames and the conv epilogue can pool adjacent lanes
        auto stride_of = [&](int i) {
            if (tensor_cores && i == 1) return kFoldStride;
            return tensor_cores && i < 3 ? kLength[i] + 1 : kLength[i];
        };
        if (!tensor_cores) {
            const int stride = stride_of(0);
            dim3 grid(ceil_div(stride, 256), count);
            LaunchScope scope("frames_kernel", stream);
            frames_kernel<<<grid, 256, 0, stream>>>(
                audio8k, w.act, out_samples, frames, first, count, hop, padding, stride);
            PMN_TRY(launched("frames_kernel"));
        } else if (shared0) {
            // Block 0: LayerNorm of every frame's window of the shared pooled convolution ->
            // block 1's folded operand
            const int stride_next = stride_of(1);
            const int t_next = count * (stride_next / kFold);
            PMN_TRY(launch_zero_plane_pads(w.planes, 1, kFold * kChannels[1], t_next, stream));
            const int first_frame = first - item0 * frames;
            {
                LaunchScope scope("frame_stats_kernel", stream);
                frame_stats_kernel<<<ceil_div(count, 8), 256, 0, stream>>>(
                    w.sums + kCropStart / 2, pooled_row, w.stats, count, frames, item_rows / 2,
                    hop / 2, kLength[1], first_frame, kChannels[1]);
                PMN_TRY(launched("frame_stats_kernel"));
            }
            const int width = kLength[1] + (kSharedFrames - 1) * (hop / 2);
            const bool f8 = p->block1_shift >= 0;        // block 1 takes "fp16 + 2 x fp8" operands
            const int smem = (f8 ? 16 : 8) * width * (int)sizeof(float);
            PMN_REQUIRE(smem <= 200 * 1024, "pitch: hop too long for the shared block 0");
            static bool configured = false;
            if (!configured) {
                PMN_TRY(check_cuda(
                    cudaFuncSetAttribute(shared_norm_planes_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                    "shared_norm_planes smem attribute"));
                PMN_TRY(check_cuda(
                    cudaFuncSetAttribute(shared_norm_planes_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                    "shared_norm_planes smem attribute"));
                configured = true;
            }
            dim3 grid(ceil_div(count, kSharedFrames), kChannels[1] / (f8 ? 16 : 8));
            LaunchScope scope("shared_norm_planes_kernel", stream);
            auto kernel = f8 ? shared_norm_planes_kernel<true> : shared_norm_planes_kernel<false>;
            kernel<<<grid, f8 ? 512 : 256, smem, stream>>>(
                w.pooled + kCropStart / 2, p->norm_weight[0], p->norm_bias[0], w.stats, w.planes,
                kChannels[1], hop / 2, kLength[1], pooled_row, count, tc_padded_length(t_next),
                stride_next, frames, item_rows / 2, first_frame);
            PMN_TRY(launched("shared_norm_planes_kernel"));
        } else {
            // Block 0, a sub-chunk of frames at a time (kSubFrames): frames -> im2col operand ->
            // 1x1 convolution over the 32 taps (ReLU + MaxPool in its epilogue) -> LayerNorm ->
            // block 1's folded operand, at the sub-chunk's place in the chunk's operand
            const int stride = stride_of(0), stride_next = stride_of(1);
            const int t_next = count * (stride_next / kFold);
            PMN_TRY(launch_zero_plane_pads(w.planes, 1, kFold * kChannels[1], t_next, stream));
            for (int sub = 0; sub < count; sub += kSubFrames) {
                const int n = std::min(kSubFrames, count - sub);
                {
                    dim3 grid(ceil_div(stride, 256), n);
                    LaunchScope scope("frames_kernel", stream);
                    frames_kernel<<<grid, 256, 0, stream>>>(
                        audio8k, w.act, out_samples, frames, first + sub, n, hop, padding, stride);
                    PMN_TRY(launched("frames_kernel"));
                }
                const size_t conv_rows = (size_t)n * stride - (kKernel - 1);
                const int t_pad = tc_padded_length((int)conv_rows);
                {
                    dim3 grid(ceil_div(t_pad, 128), 4);
                    LaunchScope scope("im2col_planes_kernel", stream);
                    im2col_planes_kernel<<<grid, 128, 0, stream>>>(
                        w.act, w.planes0, n * stride, (int)conv_rows, t_pad);
                    PMN_TRY(launched("im2col_planes_kernel"));
                }
                TcConvArgs a;
                a.x_planes = w.planes0; a.w_slabs = p->conv_slabs[0]; a.bias = p->conv_bias[0];
                a.out = w.conv; a.batch = 1; a.c_in = kKernel; a.c_out = kChannels[1]; a.k = 1;
                a.t_len = (int)conv_rows; a.valid = true; a.relu = true; a.pool = true;
                a.out_row = (int)(conv_rows / 2);
                PMN_TRY(launch_conv1d_tc(a, stream));
                LaunchScope scope("pool_norm_planes_kernel", stream);
                pool_norm_planes_kernel<<<n, 256, 0, stream>>>(
                    w.conv, p->norm_weight[0], p->norm_bias[0], w.planes, kChannels[1], stride / 2,
                    kLength[1], false, conv_rows / 2, n, tc_padded_length(t_next), false, stride_next,
                    true, false, sub);
                PMN_TRY(launched("pool_norm_planes_kernel"));
            }
        }
        const int padded = (count + 15) / 16 * 16;  // frame-mode tiles cover 16 frames
        for (int i = tensor_cores ? 1 : 0; i < kLayers; ++i) {
            if (tensor_cores && i >= kFrameMajorFirst) {
                // frame-major layer: one dense product per output position, rows = frames
                const int c_in = kChannels[i], c_out = kChannels[i + 1];
                const int positions_in = kLength[i], positions_out = kLength[i + 1];
                const int t_pad = tc_padded_length(count);
                TcConvArgs a;
                a.x_planes = w.planes; a.w_slabs = p->conv_slabs[i]; a.bias = p->conv_bias[i];
                a.out = w.conv; a.batch = positions_out; a.c_in = kKernel * c_in; a.c_out = c_out;
                a.k = 1; a.t_len = count; a.valid = true; a.relu = true; a.out_row = count;
                a.item_groups = c_in / 8; a.plane_groups = positions_in * (c_in / 8);
                PMN_TRY(launch_conv1d_tc(a, stream));
                {
                    dim3 grid(ceil_div(count, 32), kStatSlices);
                    LaunchScope scope("frame_major_stats_kernel", stream);
                    frame_major_stats_kernel<<<grid, 256, 0, stream>>>(
                        w.conv, w.partial, positions_out * c_out, count, count);
                    PMN_TRY(launched("frame_major_stats_kernel"));
                }
                dim3 grid(ceil_div(count, 32), 16);
                LaunchScope scope("frame_major_norm_planes_kernel", stream);
                frame_major_norm_planes_kernel<<<grid, 256, 0, stream>>>(
                    w.conv, w.partial, p->norm_weight[i], p->norm_bias[i], w.planes, c_out,
                    positions_out, count, count, t_pad);
                PMN_TRY(launched("frame_major_norm_planes_kernel"));
                continue;
            }
            const int l_in = stride_of(i);                       // frame stride of the input rows
            const bool folded = tensor_cores && i == 1;          // block 1: folded by 4 in time
            const size_t conv_rows = folded ? (size_t)count * (l_in / kFold) - (kFoldTaps - 1)
                                            : (size_t)count * l_in - (kKernel - 1);
            const bool pool_in_conv = tensor_cores && kPooled[i] && !folded;
            const size_t row = pool_in_conv ? conv_rows / 2 : conv_rows;  // row length of w.conv
            // frame stride inside w.conv
            const int l_conv = folded ? l_in / kFold : pool_in_conv ? l_in / 2 : l_in;
            if (tensor_cores) {
                TcConvArgs a;
                a.x_planes = w.planes; a.w_slabs = p->conv_slabs[i]; a.bias = p->conv_bias[i];
                a.out = w.conv; a.batch = 1; a.c_out = kChannels[i + 1];
                a.valid = true; a.relu = true; a.pool = pool_in_conv; a.out_row = (int)row;
                if (folded) {
                    a.c_in = kFold * kChannels[i]; a.c_out = kFold * kChannels[i + 1];
                    a.k = kFoldTaps; a.t_len = count * (l_in / kFold); a.bias = p->folded_bias;
                    if (shared0 && p->block1_shift >= 0) {
                        a.f8x2 = true;
                        a.w_slabs = static_cast<const __nv_bfloat16*>(p->block1_f8_slabs);
                        a.f8_unscale = tc_f8_unscale(p->block1_shift);
                    }
                } else {
                    a.c_in = kChannels[i]; a.k = kKernel; a.t_len = count * l_in;
                }
                if (i == kLayers - 1) {
                    // block 5 keeps 4 of 35 rows per frame: 16 frames x 8 rows per MMA tile
                    a.t_len = padded * l_in;
                    a.frames = count; a.frame_length = l_in; a.frame_valid = kLength[i + 1];
                }
                PMN_TRY(launch_conv1d_tc(a, stream));
            } else {
                Conv1dArgs a;
                a.x = w.act; a.weight = p->conv_weight[i]; a.bias = p->conv_bias[i]; a.out = w.conv;
                a.batch = 1; a.c_in = kChannels[i]; a.c_out = kChannels[i + 1];
                a.t_in = count * l_in; a.t_out = (int)row; a.k = kKernel; a.out_act = 2;
                PMN_TRY(launch_conv1d(a, stream));
            }
            if (tensor_cores && i + 1 == kFrameMajorFirst) {
                // the next block is frame-major: planes[plane][t groups + g][kTcPad + f][8]
                LaunchScope scope("pool_norm_planes_kernel", stream);
                pool_norm_planes_kernel<<<count, 256, 0, stream>>>(
                    w.conv, p->norm_weight[i], p->norm_bias[i], w.planes, kChannels[i + 1], l_conv,
                    kLength[i + 1], false, row, count, tc_padded_length(count), false, kLength[i + 1],
                    false, folded, 0, 0, 0, 0, true);
                PMN_TRY(launched("pool_norm_planes_kernel"));
            } else if (tensor_cores && i < kLayers - 1) {
                // the next block runs on the tensor cores: write its operand planes
                const int stride_next = stride_of(i + 1);
                const int frames_out = i + 1 == kLayers - 1 ? padded : count;
                const bool fold_next = i + 1 == 1;           // the next block reads the folded layout
                const int t_next = fold_next ? frames_out * (stride_next / kFold) : frames_out * stride_next;
                PMN_TRY(launch_zero_plane_pads(
                    w.planes, 1, fold_next ? kFold * kChannels[i + 1] : kChannels[i + 1], t_next, stream));
                LaunchScope scope("pool_norm_planes_kernel", stream);
                pool_norm_planes_kernel<<<frames_out, 256, 0, stream>>>(
                    w.conv, p->norm_weight[i], p->norm_bias[i], w.planes, kChannels[i + 1], l_conv,
                    kLength[i + 1], false, row, count, tc_padded_length(t_next), false, stride_next,
                    fold_next, folded, 0);
                PMN_TRY(launched("pool_norm_planes_kernel"));
            } else if (tensor_cores) {
                // last block: (512, 4) per frame becomes one 2048-channel row of the head's operand
                PMN_TRY(launch_zero_plane_pads(w.planes, 1, 2048, count, stream));
                LaunchScope scope("pool_norm_planes_kernel", stream);
                pool_norm_planes_kernel<<<count, 256, 0, stream>>>(
                    w.conv, p->norm_weight[i], p->norm_bias[i], w.planes, kChannels[i + 1], l_conv,
                    kLength[i + 1], false, row, count, tc_padded_length(count), true, kLength[i + 1],
                    false, false, 0);
                PMN_TRY(launched("pool_norm_planes_kernel"));
            } else {
                LaunchScope scope("pool_norm_kernel", stream);
                pool_norm_kernel<<<count, 256, 0, stream>>>(
                    w.conv, p->norm_weight[i], p->norm_bias[i], w.act, kChannels[i + 1], l_in,
                    kLength[i + 1], kPooled[i], row, count, i == kLayers - 1);
                PMN_TRY(launched("pool_norm_kernel"));
            }
        }
        if (tensor_cores) {
            TcConvArgs a;
            a.x_planes = w.planes; a.w_slabs = p->head_slabs; a.bias = p->head_bias; a.out = w.logits_t;
            a.batch = 1; a.c_in = 2048; a.c_out = kBins; a.k = 1; a.valid = true; a.t_len = count;
            PMN_TRY(launch_conv1d_tc(a, stream));
        } else {
            Conv1dArgs a;
            a.x = w.act; a.weight = p->head_weight; a.bias = p->head_bias; a.out = w.logits_t;
            a.batch = 1; a.c_in = 2048; a.c_out = kBins; a.t_in = a.t_out = count; a.k = 1;
            PMN_TRY(launch_conv1d(a, stream));
        }
        {
            LaunchScope scope("posterior_kernel", stream);
            posterior_kernel<<<ceil_div(count, 32), 256, 0, stream>>>(
                w.logits_t, count, min_bin, max_bin, w.masked + (size_t)first * kBins,
                w.distribution + (size_t)first * kBins, periodicity + first);
            PMN_TRY(launched("posterior_kernel"));
        }
    }
    }

    // 3. Viterbi over the posteriors, 4. local expected value
    int* bins = bins_out ? bins_out : w.bins;
    PMN_TRY(launch_viterbi(
        w.distribution, nullptr, transition, initial, false, bins, batch, frames, kBins,
        w.viterbi, w.viterbi_bytes, stream));
    {
        LaunchScope scope("pitch_kernel", stream);
        pitch_kernel<<<ceil_div(total, 128), 128, 0, stream>>>(w.masked, bins, pitch, total);
        PMN_TRY(launched("pitch_kernel"));
    }
    if (logits_out)
        PMN_TRY(check_cuda(
            cudaMemcpyAsync(logits_out, w.masked, (size_t)total * kBins * 4, cudaMemcpyDeviceToDevice, stream),
            "copy logits"));
    return PMN_OK;
}

}  // namespace pmn

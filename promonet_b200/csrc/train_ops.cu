// Element-wise, reduction and optimizer kernels of the training step
// (promonet/train/core.py:183-369, promonet/train/loss.py:11-53): all HBM-bound,
// one pass over their tensors.
#include "train.cuh"

namespace pmn {

namespace {

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch domain

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float s = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0.f;
        for (int offset = 16; offset > 0; offset >>= 1) s += __shfl_xor_sync(0xffffffffu, s, offset);
        if (threadIdx.x == 0) scratch[0] = s;
    }
    __syncthreads();
    const float total = scratch[0];
    __syncthreads();
    return total;
}

__device__ __forceinline__ int reflect(int i, int n) {
    // torch 'reflect' padding: no repeat of the edge sample
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// out[r, u] = x[r, reflect(u - left)], u in [0, left + t_in + right)
// (discriminator.py:78-81 pads on the right to a multiple of the period)
__global__ void reflect_pad_kernel(
    const float* __restrict__ x, float* __restrict__ out, int t_in, int left, int t_out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (u >= t_out) return;
    out[(size_t)r * t_out + u] = x[(size_t)r * t_in + reflect(u - left, t_in)];
}

// gx[r, i] (+)= sum of gout[r, u] over the u that read sample i
__global__ void reflect_pad_backward_kernel(
    const float* __restrict__ gout, float* __restrict__ gx, int t_in, int left, int right,
    int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= t_in) return;
    const int t_out = left + t_in + right;
    const float* row = gout + (size_t)r * t_out;
    float v = row[left + i];
    // left padding position u = left - i (i in 1..left), right u = left + 2 (t_in - 1) - i
    if (i >= 1 && i <= left) v += row[left - i];
    const int u = left + 2 * (t_in - 1) - i;
    if (i < t_in - 1 && u < t_out && u >= left + t_in) v += row[u];
    float* dst = gx + (size_t)r * t_in + i;
    *dst = accumulate ? *dst + v : v;
}

__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float xv = x ? a * x[i] : 0.f;
        y[i] = b == 0.f ? xv : xv + b * y[i];
    }
}

// LSGAN terms (loss.py:29-53): loss += weight * mean((x - target)^2);
// grad = weight * 2 (x - target) / n
__global__ void __launch_bounds__(256) mse_to_target_kernel(
    const float* __restrict__ x, int64_t n, float target, float weight,
    float* __restrict__ loss, float* __restrict__ grad) {
    __shared__ float scratch[32];
    float sum = 0.f;
    const float scale = weight / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float d = x[i] - target;
        sum = fmaf(d, d, sum);
        if (grad) grad[i] = 2.f * d * scale;
    }
    sum = block_sum(sum, scratch);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, sum * scale);
}

// Feature matching / L1 terms (loss.py:11-26): loss += weight * mean(|fake - real|);
// gfake (+)= weight * sign(fake - real) / n
__global__ void __launch_bounds__(256) l1_mean_kernel(
    const float* __restrict__ fake, const float* __restrict__ real, int64_t n, float weight,
    float* __restrict__ loss, float* __restrict__ gfake, int accumulate) {
    __shared__ float scratch[32];
    float sum = 0.f;
    const float scale = weight / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float d = fake[i] - real[i];
        sum += fabsf(d);
        if (gfake) {
            const float gv = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
            gfake[i] = accumulate ? gfake[i] + gv : gv;
        }
    }
    sum = block_sum(sum, scratch);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, sum * scale);
}

// torch.optim.AdamW (train/core.py:63-64, config/defaults.py:390-394), one fused pass:
//   p *= 1 - lr wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adamw_kernel(
    float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
    float* __restrict__ exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
    float weight_decay, float correction1, float correction2_sqrt, float grad_scale,
    const float* __restrict__ step_device) {
    if (step_device) {
        // step count kept on the device so that a captured CUDA graph stays valid as it advances
        const double t = (double)*step_device;
        correction1 = (float)(1. - pow((double)beta1, t));
        correction2_sqrt = (float)sqrt(1. - pow((double)beta2, t));
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float gr = grad[i] * grad_scale;
        float pv = param[i] * (1.f - lr * weight_decay);
        const float m = beta1 * exp_avg[i] + (1.f - beta1) * gr;
        const float v = beta2 * exp_avg_sq[i] + (1.f - beta2) * gr * gr;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / correction2_sqrt + eps;
        pv -= (lr / correction1) * (m / denom);
        param[i] = pv;
    }
}

// Data-parallel AdamW over NVLink peer memory: reduce-scatter + optimizer + all-gather in one
// kernel.  Every rank owns the elements [begin, end) of the flat parameter buffer: it reads that
// slice of every rank's gradient buffer through peer pointers, averages, takes the AdamW step
// with its (shard-local) moments and writes the new parameters into every rank's buffer.
struct PeerBuffers {
    const float* grad[kMaxPeers];
    float* param[kMaxPeers];
};

__global__ void __launch_bounds__(256) adamw_peer_kernel(
    PeerBuffers peers, int world, int rank, float* __restrict__ exp_avg,
    float* __restrict__ exp_avg_sq, int64_t begin, int64_t end, float lr, float beta1, float beta2,
    float eps, float weight_decay, float correction1, float correction2_sqrt,
    const float* __restrict__ step_device) {
    if (step_device) {
        const double t = (double)*step_device;
        correction1 = (float)(1. - pow((double)beta1, t));
        correction2_sqrt = (float)sqrt(1. - pow((double)beta2, t));
    }
    const float inv_world = 1.f / (float)world;
    // begin and end are multiples of 4 (the flat buffers are 16-byte aligned and padded)
    for (int64_t i = begin + 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); i < end;
         i += 4 * (int64_t)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(peers.grad[r] + i);
            g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
        }
        float4 pv = *reinterpret_cast<const float4*>(peers.param[rank] + i);
        float4 m = *reinterpret_cast<const float4*>(exp_avg + i);
        float4 v = *reinterpret_cast<const float4*>(exp_avg_sq + i);
        float* pe = &pv.x; float* me = &m.x; float* ve = &v.x; const float* ge = &g.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = ge[e] * inv_world;
            float value = pe[e] * (1.f - lr * weight_decay);
            me[e] = beta1 * me[e] + (1.f - beta1) * gr;
            ve[e] = beta2 * ve[e] + (1.f - beta2) * gr * gr;
            value -= (lr / correction1) * (me[e] / (sqrtf(ve[e]) / correction2_sqrt + eps));
            pe[e] = value;
        }
        *reinterpret_cast<float4*>(exp_avg + i) = m;
        *reinterpret_cast<float4*>(exp_avg_sq + i) = v;
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(peers.param[r] + i) = pv;
    }
}

// out[r] (+)= sum_c x[r, c]
__global__ void __launch_bounds__(256) row_sum_kernel(
    const float* __restrict__ x, float* __restrict__ out, int cols, int accumulate) {
    __shared__ float scratch[32];
    const float* row = x + (size_t)blockIdx.x * cols;
    float sum = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) sum += row[i];
    sum = block_sum(sum, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = accumulate ? out[blockIdx.x] + sum : sum;
}

// out[c] (+)= sum_{b, i} x[b, c, i]   (bias gradient of a transposed convolution)
__global__ void __launch_bounds__(256) channel_sum_kernel(
    const float* __restrict__ x, float* __restrict__ out, int batch, int channels, int inner,
    int accumulate) {
    __shared__ float scratch[32];
    const int c = blockIdx.x;
    float sum = 0.f;
    for (int b = 0; b < batch; ++b) {
        const float* row = x + ((size_t)b * channels + c) * inner;
        for (int i = threadIdx.x; i < inner; i += blockDim.x) sum += row[i];
    }
    sum = block_sum(sum, scratch);
    if (threadIdx.x == 0) out[c] = accumulate ? out[c] + sum : sum;
}

// dst[r, dst_offset + j] (+)= src[r, src_offset + j], j < cols: concatenation / split
// along the last axis (torch.cat(..., dim=-1), discriminator.py:204)
__global__ void copy_columns_kernel(
    const float* __restrict__ src, int src_width, int src_offset, float* __restrict__ dst,
    int dst_width, int dst_offset, int64_t rows, int cols, int accumulate) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / cols;
        const int j = (int)(idx - r * cols);
        const float v = src[r * src_width + src_offset + j];
        float* d = dst + r * dst_width + dst_offset + j;
        *d = accumulate ? *d + v : v;
    }
}

// gtable[index[b, f], e] += gout[b, channel_offset + e, f]  (backward of the pitch
// embedding lookup, generator.py:158-164)
__global__ void embedding_backward_kernel(
    const float* __restrict__ gout, const int64_t* __restrict__ index, float* __restrict__ gtable,
    int channels, int frames, int rows, int out_channels, int channel_offset) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y, b = blockIdx.z;
    if (f >= frames) return;
    int64_t row = index[(size_t)b * frames + f];
    row = row < 0 ? 0 : (row >= rows ? rows - 1 : row);
    atomicAdd(gtable + (size_t)row * channels + e,
              gout[((size_t)b * out_channels + channel_offset + e) * frames + f]);
}

// bins = clip(searchsorted(edges, clip(pitch, fmin, fmax), side=left), 0, n - 1)
// (generator.py:153-157)
__global__ void pitch_bins_kernel(
    const float* __restrict__ pitch, const float* __restrict__ edges, int64_t* __restrict__ bins,
    int n, int num_edges, float fmin, float fmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float hz = fminf(fmaxf(pitch[i], fmin), fmax);
    int lo = 0, hi = num_edges;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] < hz) lo = mid + 1; else hi = mid;
    }
    bins[i] = min(lo, num_edges - 1);
}

// g[b] = [speaker_embedding[speakers[b]], sbr[b], lr[b]]  (generator.py:56-68)
__global__ void global_features_kernel(
    const float* __restrict__ speaker_embedding, const int64_t* __restrict__ speakers,
    const float* __restrict__ sbr, const float* __restrict__ lr, float* __restrict__ out,
    int speaker_channels, int num_speakers) {
    const int b = blockIdx.x;
    int64_t speaker = speakers[b];
    speaker = speaker < 0 ? 0 : (speaker >= num_speakers ? num_speakers - 1 : speaker);
    float* dst = out + (size_t)b * (speaker_channels + 2);
    for (int i = threadIdx.x; i < speaker_channels; i += blockDim.x)
        dst[i] = speaker_embedding[(size_t)speaker * speaker_channels + i];
    if (threadIdx.x == 0) {
        dst[speaker_channels] = sbr[b];
        dst[speaker_channels + 1] = lr[b];
    }
}

int grid_for(int64_t n) { return (int)min((int64_t)148 * 8, (n + 255) / 256); }

}  // namespace

int launch_reflect_pad(
    const float* x, float* out, int rows, int t_in, int left, int right, cudaStream_t stream) {
    PMN_REQUIRE(x && out && rows > 0 && rows <= 65535 && t_in > 1, "reflect_pad: bad argument");
    PMN_REQUIRE(left >= 0 && right >= 0 && left < t_in && right < t_in,
                "reflect_pad: padding must be smaller than the input");
    const int t_out = left + t_in + right;
    dim3 grid(ceil_div(t_out, 256), rows);
    LaunchScope scope("reflect_pad_kernel", stream);
    reflect_pad_kernel<<<grid, 256, 0, stream>>>(x, out, t_in, left, t_out);
    return launched("reflect_pad_kernel");
}

int launch_reflect_pad_backward(
    const float* gout, float* gx, int rows, int t_in, int left, int right, int accumulate,
    cudaStream_t stream) {
    PMN_REQUIRE(gout && gx && rows > 0 && rows <= 65535 && t_in > 1, "reflect_pad_backward: bad argument");
    PMN_REQUIRE(left >= 0 && right >= 0 && left < t_in && right < t_in,
                "reflect_pad_backward: padding must be smaller than the input");
    dim3 grid(ceil_div(t_in, 256), rows);
    LaunchScope scope("reflect_pad_backward_kernel", stream);
    reflect_pad_backward_kernel<<<grid, 256, 0, stream>>>(gout, gx, t_in, left, right, accumulate);
    return launched("reflect_pad_backward_kernel");
}

int launch_axpby(float a, const float* x, float b, float* y, int64_t n, cudaStream_t stream) {
    PMN_REQUIRE(y && n >= 0, "axpby: bad argument");
    if (n == 0) return PMN_OK;
    LaunchScope scope("axpby_kernel", stream);
    axpby_kernel<<<grid_for(n), 256, 0, stream>>>(a, x, b, y, n);
    return launched("axpby_kernel");
}

int launch_mse_to_target(
    const float* x, int64_t n, float target, float weight, float* loss, float* grad,
    cudaStream_t stream) {
    PMN_REQUIRE(x && n > 0 && (loss || grad), "mse_to_target: bad argument");
    LaunchScope scope("mse_to_target_kernel", stream);
    mse_to_target_kernel<<<grid_for(n), 256, 0, stream>>>(x, n, target, weight, loss, grad);
    return launched("mse_to_target_kernel");
}

int launch_l1_mean(
    const float* fake, const float* real, int64_t n, float weight, float* loss, float* gfake,
    int accumulate, cudaStream_t stream) {
    PMN_REQUIRE(fake && real && n > 0 && (loss || gfake), "l1_mean: bad argument");
    LaunchScope scope("l1_mean_kernel", stream);
    l1_mean_kernel<<<grid_for(n), 256, 0, stream>>>(fake, real, n, weight, loss, gfake, accumulate);
    return launched("l1_mean_kernel");
}

int launch_adamw(
    float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
    float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
    const float* step_device, cudaStream_t stream) {
    PMN_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && (step >= 1 || step_device),
                "adamw: bad argument");
    const float correction1 = (float)(1. - pow((double)beta1, step));
    const float correction2_sqrt = (float)sqrt(1. - pow((double)beta2, step));
    LaunchScope scope("adamw_kernel", stream);
    adamw_kernel<<<grid_for(n), 256, 0, stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
        correction1, correction2_sqrt, grad_scale, step_device);
    return launched("adamw_kernel");
}

int launch_adamw_peer(
    const float* const* grad_peers, float* const* param_peers, int world, int rank, float* exp_avg,
    float* exp_avg_sq, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps,
    float weight_decay, int step, const float* step_device, cudaStream_t stream) {
    PMN_REQUIRE(grad_peers && param_peers && exp_avg && exp_avg_sq && world >= 1 && world <= kMaxPeers &&
                rank >= 0 && rank < world && begin >= 0 && end >= begin && begin % 4 == 0 && end % 4 == 0 &&
                (step >= 1 || step_device), "adamw_peer: bad argument");
    if (end == begin) return PMN_OK;
    PeerBuffers peers;
    for (int r = 0; r < world; ++r) {
        PMN_REQUIRE(grad_peers[r] && param_peers[r], "adamw_peer: null peer pointer");
        peers.grad[r] = grad_peers[r];
        peers.param[r] = param_peers[r];
    }
    const float correction1 = step_device ? 1.f : (float)(1. - pow((double)beta1, step));
    const float correction2_sqrt = step_device ? 1.f : (float)sqrt(1. - pow((double)beta2, step));
    LaunchScope scope("adamw_peer_kernel", stream);
    adamw_peer_kernel<<<grid_for((end - begin) / 4), 256, 0, stream>>>(
        peers, world, rank, exp_avg, exp_avg_sq, begin, end, lr, beta1, beta2, eps, weight_decay,
        correction1, correction2_sqrt, step_device);
    return launched("adamw_peer_kernel");
}

int launch_row_sum(
    const float* x, float* out, int rows, int cols, int accumulate, cudaStream_t stream) {
    PMN_REQUIRE(x && out && rows > 0 && cols > 0, "row_sum: bad argument");
    LaunchScope scope("row_sum_kernel", stream);
    row_sum_kernel<<<rows, 256, 0, stream>>>(x, out, cols, accumulate);
    return launched("row_sum_kernel");
}

int launch_channel_sum(
    const float* x, float* out, int batch, int channels, int inner, int accumulate,
    cudaStream_t stream) {
    PMN_REQUIRE(x && out && batch > 0 && channels > 0 && inner > 0, "channel_sum: bad argument");
    LaunchScope scope("channel_sum_kernel", stream);
    channel_sum_kernel<<<channels, 256, 0, stream>>>(x, out, batch, channels, inner, accumulate);
    return launched("channel_sum_kernel");
}

int launch_copy_columns(
    const float* src, int src_width, int src_offset, float* dst, int dst_width, int dst_offset,
    int64_t rows, int cols, int accumulate, cudaStream_t stream) {
    PMN_REQUIRE(src && dst && rows > 0 && cols > 0 && src_offset >= 0 && dst_offset >= 0 &&
                src_offset + cols <= src_width && dst_offset + cols <= dst_width,
                "copy_columns: bad argument");
    LaunchScope scope("copy_columns_kernel", stream);
    copy_columns_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(
        src, src_width, src_offset, dst, dst_width, dst_offset, rows, cols, accumulate);
    return launched("copy_columns_kernel");
}

int launch_embedding_backward(
    const float* gout, const int64_t* index, float* gtable, int batch, int channels, int frames,
    int rows, int out_channels, int channel_offset, cudaStream_t stream) {
    PMN_REQUIRE(gout && index && gtable && batch > 0 && batch <= 65535 && channels > 0 &&
                channels <= 65535 && frames > 0 && rows > 0, "embedding_backward: bad argument");
    dim3 grid(ceil_div(frames, 128), channels, batch);
    LaunchScope scope("embedding_backward_kernel", stream);
    embedding_backward_kernel<<<grid, 128, 0, stream>>>(
        gout, index, gtable, channels, frames, rows, out_channels, channel_offset);
    return launched("embedding_backward_kernel");
}

int launch_global_features(
    const float* speaker_embedding, const int64_t* speakers, const float* sbr, const float* lr,
    float* out, int batch, int speaker_channels, int num_speakers, cudaStream_t stream) {
    PMN_REQUIRE(speaker_embedding && speakers && sbr && lr && out && batch > 0 &&
                speaker_channels > 0 && num_speakers > 0, "global_features: bad argument");
    LaunchScope scope("global_features_kernel", stream);
    global_features_kernel<<<batch, 128, 0, stream>>>(
        speaker_embedding, speakers, sbr, lr, out, speaker_channels, num_speakers);
    return launched("global_features_kernel");
}

int launch_pitch_bins(
    const float* pitch, const float* edges, int64_t* bins, int n, int num_edges, float fmin,
    float fmax, cudaStream_t stream) {
    PMN_REQUIRE(pitch && edges && bins && n > 0 && num_edges > 0, "pitch_bins: bad argument");
    LaunchScope scope("pitch_bins_kernel", stream);
    pitch_bins_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(pitch, edges, bins, n, num_edges, fmin, fmax);
    return launched("pitch_bins_kernel");
}

}  // namespace pmn

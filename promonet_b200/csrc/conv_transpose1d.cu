// LeakyReLU + ConvTranspose1d (kernel = 2*stride, padding = stride/2), polyphase.
//
// MultiReceptiveFieldFusion's upsampler, promonet/model/hifigan.py:97-106 with
// (k, s) = (16, 8), (16, 8), (4, 2), (4, 2) (config/defaults.py:259-262).  With
// PyTorch semantics y[o, i*s - p + j] += w[c, o, j] * x[c, i], every output
// sample receives exactly two taps:
//   t = s*i + q, q in [0, s):   q <  s/2: y = w[., q + s/2] x[i]   + w[., q + 3s/2] x[i-1]
//                               q >= s/2: y = w[., q - s/2] x[i+1] + w[., q + s/2]   x[i]
// so a thread that owns input position i produces the s contiguous outputs
// s*i .. s*i+s-1 from x[i-1], x[i], x[i+1] with no wasted multiplies.
#include "common.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 256;
constexpr int kLanes = 32;   // threads along time
constexpr int kGroups = 8;   // thread groups along C_out

template <int S, int RO, int P>
__global__ void __launch_bounds__(kThreads, 2) conv_transpose1d_kernel(
    const float* __restrict__ x, const float* __restrict__ weight,
    const float* __restrict__ bias, float* __restrict__ out,
    int c_in, int c_out, int t_in, float in_slope, int chunk) {
    constexpr int K = 2 * S;
    constexpr int H = S / 2;
    constexpr int TI = kLanes * P;     // input positions per block
    constexpr int CO_T = kGroups * RO; // output channels per block
    constexpr int XW = TI + 2;
    extern __shared__ __align__(16) float smem[];
    float* xsm = smem;                 // [chunk][XW]
    float* wsm = smem + chunk * XW;    // [chunk][CO_T][K]

    const int tid = threadIdx.x;
    const int tlane = tid % kLanes;
    const int og = tid / kLanes;
    const int i0 = blockIdx.x * TI;
    const int o0 = blockIdx.y * CO_T;
    const int b = blockIdx.z;
    const float* xb = x + (size_t)b * c_in * t_in;

    float acc[P][RO][S];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int r = 0; r < RO; ++r)
#pragma unroll
            for (int q = 0; q < S; ++q) acc[p][r][q] = 0.f;

    for (int c0 = 0; c0 < c_in; c0 += chunk) {
        const int cc = min(chunk, c_in - c0);
        for (int idx = tid; idx < cc * XW; idx += kThreads) {
            const int c = idx / XW, u = idx % XW;
            const int i = i0 - 1 + u;
            float v = 0.f;
            if (i >= 0 && i < t_in) v = leaky(__ldg(xb + (size_t)(c0 + c) * t_in + i), in_slope);
            xsm[idx] = v;
        }
        for (int idx = tid; idx < cc * CO_T * K; idx += kThreads) {
            const int c = idx / (CO_T * K), rest = idx % (CO_T * K);
            const int o = rest / K, j = rest % K;
            float v = 0.f;
            if (o0 + o < c_out) v = __ldg(weight + ((size_t)(c0 + c) * c_out + o0 + o) * K + j);
            wsm[idx] = v;
        }
        __syncthreads();
        for (int c = 0; c < cc; ++c) {
            float xm[P], xc[P], xp[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float* s = xsm + c * XW + tlane + kLanes * p;
                xm[p] = s[0]; xc[p] = s[1]; xp[p] = s[2];
            }
#pragma unroll
            for (int r = 0; r < RO; ++r) {
                float w[K];
                const float4* wp = reinterpret_cast<const float4*>(
                    wsm + (c * CO_T + og * RO + r) * K);
#pragma unroll
                for (int v = 0; v < K / 4; ++v) {
                    const float4 f = wp[v];
                    w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
#pragma unroll
                    for (int q = 0; q < H; ++q)
                        acc[p][r][q] = fmaf(w[q + H], xc[p], fmaf(w[q + H + S], xm[p], acc[p][r][q]));
#pragma unroll
                    for (int q = H; q < S; ++q)
                        acc[p][r][q] = fmaf(w[q - H], xp[p], fmaf(w[q + H], xc[p], acc[p][r][q]));
                }
            }
        }
        __syncthreads();
    }

    const int t_out = S * t_in;
#pragma unroll
    for (int r = 0; r < RO; ++r) {
        const int o = o0 + og * RO + r;
        if (o >= c_out) continue;
        const float bv = bias ? __ldg(bias + o) : 0.f;
        float* row = out + ((size_t)b * c_out + o) * t_out;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int i = i0 + tlane + kLanes * p;
            if (i >= t_in) continue;
            if constexpr (S % 4 == 0) {
#pragma unroll
                for (int q = 0; q < S; q += 4)
                    *reinterpret_cast<float4*>(row + (size_t)S * i + q) = make_float4(
                        acc[p][r][q] + bv, acc[p][r][q + 1] + bv,
                        acc[p][r][q + 2] + bv, acc[p][r][q + 3] + bv);
            } else {
#pragma unroll
                for (int q = 0; q < S; q += 2)
                    *reinterpret_cast<float2*>(row + (size_t)S * i + q) =
                        make_float2(acc[p][r][q] + bv, acc[p][r][q + 1] + bv);
            }
        }
    }
}

template <int S, int RO, int P>
int launch_variant(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, float in_slope, cudaStream_t stream) {
    constexpr int K = 2 * S;
    constexpr int TI = kLanes * P;
    constexpr int CO_T = kGroups * RO;
    const int per_channel = (TI + 2 + CO_T * K) * (int)sizeof(float);
    int chunk = max(1, min(min(64 * 1024 / per_channel, 32), c_in));
    const size_t smem = (size_t)chunk * per_channel;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(
                conv_transpose1d_kernel<S, RO, P>,
                cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024),
            "conv_transpose1d smem attribute"));
        configured = true;
    }
    dim3 grid(ceil_div(t_in, TI), ceil_div(c_out, CO_T), batch);
    LaunchScope scope("conv_transpose1d_kernel", stream);
    conv_transpose1d_kernel<S, RO, P><<<grid, kThreads, smem, stream>>>(
        x, weight, bias, out, c_in, c_out, t_in, in_slope, chunk);
    return launched("conv_transpose1d_kernel");
}

}  // namespace

int launch_conv_transpose1d(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope,
    cudaStream_t stream) {
    PMN_REQUIRE(x && weight && out, "conv_transpose1d: null pointer");
    PMN_REQUIRE(batch > 0 && batch <= 65535 && c_in > 0 && c_out > 0, "conv_transpose1d: bad shape");
    PMN_REQUIRE(k == 2 * stride, "conv_transpose1d: only kernel = 2*stride, padding = stride/2");
    if (t_in <= 0) return PMN_OK;
    if (stride == 8)
        return launch_variant<8, 4, 2>(x, weight, bias, out, batch, c_in, c_out, t_in, in_slope, stream);
    if (stride == 2)
        return launch_variant<2, 8, 4>(x, weight, bias, out, batch, c_in, c_out, t_in, in_slope, stream);
    return fail(PMN_ERR_ARGUMENT, "conv_transpose1d: stride must be 8 or 2");
}

}  // namespace pmn

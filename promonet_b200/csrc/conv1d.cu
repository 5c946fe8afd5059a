// Fused dilated Conv1d, fp32 SIMT path.
//
// One launch computes, for a (batch item, C_out tile, time tile):
//   y = act_out(bias + bias2 + residual + sum_{c,j} w[o,c,j] * lrelu(x[c, t + j*d - p]))
// which is one c1 or c2 of Block.forward (promonet/model/hifigan.py:198-210)
// with the activation that precedes it, the residual add that follows it and
// (for the last conv of a Block) the MRF mean of ResidualBlock.forward
// (hifigan.py:141-145) folded into the epilogue.  It also serves the input
// feature conv (hifigan.py:65-68, speaker projection arrives as `bias2`) and the
// penn-style F0 CNN layers (valid convs, padding 0).
//
// Tiling: 256 threads hold an 8 (channels) x 8 (time) register tile each.  The
// time steps of a thread are interleaved (t = lane + TL*i) so that the shared
// memory reads of the activation row are conflict-free for every dilation, and
// all lanes of a warp share the same 8 output channels so weight reads are
// broadcasts.  Activations are staged once per C_in chunk with the LeakyReLU
// applied on the way in; zero padding comes from the bounds check in staging.
#include "common.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 256;
constexpr int kRegO = 8;
constexpr int kRegT = 8;
constexpr int kSmemBudget = 96 * 1024;

template <int CO_T, int TT>
__global__ void __launch_bounds__(kThreads, 2) conv1d_kernel(Conv1dArgs a, int chunk, int xw) {
    constexpr int TL = TT / kRegT;  // threads along time
    extern __shared__ __align__(16) float smem[];
    float* xsm = smem;                // [chunk][xw]
    float* wsm = smem + chunk * xw;   // [chunk][K][CO_T]

    const int tid = threadIdx.x;
    const int tlane = tid % TL;
    const int og = tid / TL;
    const int t0 = blockIdx.x * TT;
    const int o0 = blockIdx.y * CO_T;
    const int b = blockIdx.z;
    const int K = a.k;
    const int dil = a.dilation;

    const float* xb = a.x + (size_t)b * a.c_in * a.t_in;

    float acc[kRegO][kRegT];
#pragma unroll
    for (int r = 0; r < kRegO; ++r)
#pragma unroll
        for (int i = 0; i < kRegT; ++i) acc[r][i] = 0.f;

    const int t_first = t0 - a.padding;  // input time of xsm[.][0]

    for (int c0 = 0; c0 < a.c_in; c0 += chunk) {
        const int cc = min(chunk, a.c_in - c0);

        // Stage activations (LeakyReLU applied once here)
        for (int c = 0; c < cc; ++c) {
            const float* xrow = xb + (size_t)(c0 + c) * a.t_in;
            float* dst = xsm + c * xw;
            for (int u = tid; u < xw; u += kThreads) {
                const int t = t_first + u;
                float v = 0.f;
                if (t >= 0 && t < a.t_in) v = leaky(__ldg(xrow + t), a.in_slope);
                dst[u] = v;
            }
        }
        // Stage weights: packed (C_in, K, C_out) -> [c][j][CO_T]
        {
            const int total = cc * K * CO_T;
            const float* wbase = a.weight + (size_t)c0 * K * a.c_out;
            for (int idx = tid; idx < total; idx += kThreads) {
                const int o = idx % CO_T;
                const int cj = idx / CO_T;
                float v = 0.f;
                if (o0 + o < a.c_out) v = __ldg(wbase + (size_t)cj * a.c_out + o0 + o);
                wsm[idx] = v;
            }
        }
        __syncthreads();

        for (int c = 0; c < cc; ++c) {
            const float* xrow = xsm + c * xw + tlane;
            const float* wrow = wsm + c * K * CO_T + og * kRegO;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                const float4 w0 = *reinterpret_cast<const float4*>(wrow + j * CO_T);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + j * CO_T + 4);
                const float w[kRegO] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float xv[kRegT];
                const float* xp = xrow + j * dil;
#pragma unroll
                for (int i = 0; i < kRegT; ++i) xv[i] = xp[TL * i];
#pragma unroll
                for (int r = 0; r < kRegO; ++r)
#pragma unroll
                    for (int i = 0; i < kRegT; ++i) acc[r][i] = fmaf(w[r], xv[i], acc[r][i]);
            }
        }
        __syncthreads();
    }

    // Epilogue
#pragma unroll
    for (int r = 0; r < kRegO; ++r) {
        const int o = o0 + og * kRegO + r;
        if (o >= a.c_out) continue;
        float base = 0.f;
        if (a.bias) base += __ldg(a.bias + o);
        if (a.bias2) base += __ldg(a.bias2 + (size_t)b * a.c_out + o);
        const size_t row = ((size_t)b * a.c_out + o) * a.t_out;
#pragma unroll
        for (int i = 0; i < kRegT; ++i) {
            const int t = t0 + tlane + TL * i;
            if (t >= a.t_out) continue;
            float v = acc[r][i] + base;
            if (a.residual) v += a.residual[row + t];
            if (a.out_act == 1) v = tanhf(v);
            else if (a.out_act == 2) v = fmaxf(v, 0.f);
            if (a.out) a.out[row + t] = v;
            if (a.accum_mode == 1) a.accum[row + t] = v * a.accum_scale;
            else if (a.accum_mode == 2) a.accum[row + t] += v * a.accum_scale;
        }
    }
}

template <int CO_T, int TT>
int launch_variant(const Conv1dArgs& a, cudaStream_t stream) {
    const int halo = (a.k - 1) * a.dilation;
    const int xw = (TT + halo + 3) & ~3;
    const int per_channel = (xw + a.k * CO_T) * (int)sizeof(float);
    int chunk = kSmemBudget / per_channel;
    chunk = max(1, min(min(chunk, 32), a.c_in));
    const size_t smem = (size_t)chunk * per_channel;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(
                conv1d_kernel<CO_T, TT>,
                cudaFuncAttributeMaxDynamicSharedMemorySize,
                kSmemBudget + 16 * 1024),
            "conv1d smem attribute"));
        configured = true;
    }
    if (smem > (size_t)kSmemBudget + 16 * 1024)
        return fail(PMN_ERR_ARGUMENT, "conv1d: kernel size * dilation too large for one tile");
    dim3 grid(ceil_div(a.t_out, TT), ceil_div(a.c_out, CO_T), a.batch);
    LaunchScope scope("conv1d_kernel", stream);
    conv1d_kernel<CO_T, TT><<<grid, kThreads, smem, stream>>>(a, chunk, xw);
    return launched("conv1d_kernel");
}

}  // namespace

int launch_conv1d(const Conv1dArgs& a, cudaStream_t stream) {
    PMN_REQUIRE(a.x && a.weight, "conv1d: null input");
    PMN_REQUIRE(a.out || (a.accum && a.accum_mode), "conv1d: no output");
    PMN_REQUIRE(a.batch > 0 && a.c_in > 0 && a.c_out > 0 && a.k > 0 && a.dilation > 0,
                "conv1d: bad shape");
    PMN_REQUIRE(a.batch <= 65535, "conv1d: batch > 65535");
    if (a.t_out <= 0) return PMN_OK;
    PMN_REQUIRE(a.t_out <= a.t_in + 2 * a.padding - (a.k - 1) * a.dilation,
                "conv1d: t_out exceeds the valid output length");
    if (a.c_out <= 32) return launch_variant<32, 512>(a, stream);
    return launch_variant<64, 256>(a, stream);
}

}  // namespace pmn

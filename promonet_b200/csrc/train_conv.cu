// Convolution family of the training step: forward, data gradient, weight gradient.
//
// The training step (promonet/train/core.py:183-369) differentiates through every
// convolution of the Generator (hifigan.py: Conv1d / ConvTranspose1d) and of the
// discriminators (discriminator.py:57-93 Conv2d (5,1)/(3,1); :146-208 Conv2d (3,9)).
// All of them are instances of one strided, dilated, zero-padded 2-D convolution
// (a 1-D convolution is the case W = 1, kw = 1 with time on the H axis), so three
// implicit-GEMM kernels cover them:
//
//   conv_gemm_kernel<false>   out[b,n,p] = sum_{c,tap} W[n,(c,tap)] a[b,c,in(p,tap)]      forward
//   conv_gemm_kernel<true>    out[b,n,q] = sum_{c,tap} W[n,(c,tap)] a[b,c,out(q,tap)]     data gradient
//                             (also the forward of a ConvTranspose)
//   conv_wgrad_kernel         gw[n,(c,tap)] += sum_{b,p} dy[b,n,p] x[b,c,in(p,tap)]       weight gradient
//
// The element-wise work around a convolution is fused into the operand loads and
// the epilogue: LeakyReLU on the way in (pre-activation blocks, hifigan.py:204-207),
// LeakyReLU / tanh on the way out (discriminator.py:86-90, hifigan.py:59), the
// backward of either as a mask on the gradient operand, bias, residual add and
// accumulation into an existing gradient.
//
// fp32 FMA tiles: the GEMM M dimension is the flattened (batch, position) index so
// that the short feature maps of the deep discriminator layers still fill a tile.
#include "train.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 256;
constexpr int kBK = 16;

__device__ __forceinline__ float operand_act(float v, float companion, int act, float slope) {
    if (act == kActLrelu) return leaky(v, slope);
    if (act == kActLreluMask) return companion > 0.f ? v : v * slope;
    if (act == kActTanhMask) return v * (1.f - companion * companion);
    return v;
}

struct GemmParams {
    ConvGemmArgs a;
    int a_ch, a_h, a_w;   // gathered tensor
    int o_ch, o_h, o_w;   // produced tensor
    int taps, kdim, m_total, o_positions;
    int channel_stride, position_stride;
    size_t batch_stride;
};

template <int BM, int BN, int TM, int TN, bool TRANSPOSED>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(GemmParams p) {
    static_assert((BM / TM) * (BN / TN) == kThreads, "tile / thread mismatch");
    static_assert(TM == 4 || TM == 8, "TM");
    static_assert(TN == 4, "TN");
    constexpr int A_PER_THREAD = kBK * BM / kThreads;
    constexpr int B_PER_THREAD = (kBK * BN + kThreads - 1) / kThreads;
    constexpr int K_STEP = kThreads / BM > 0 ? kThreads / BM : 1;  // k rows covered per pass
    constexpr int BNP = BN + 4;
    __shared__ __align__(16) float As[kBK][BM];
    __shared__ __align__(16) float Bs[kBK][BNP];

    const pmn_conv_geometry& g = p.a.g;
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- the position this thread gathers for the A tile ----
    // BM is 128 or 256: with 256 threads a thread owns position tid % BM and the
    // k rows tid / BM + i * K_STEP
    const int am = tid % BM;
    const int ak0 = tid / BM;
    const int m_load = m0 + am;
    const bool m_ok = m_load < p.m_total;
    int hb = 0, wb = 0;
    const float* a_base = p.a.a;
    const float* c_base = p.a.a_companion;
    {
        const int mm = m_ok ? m_load : 0;
        const int b = mm / p.o_positions;
        const int rem = mm - b * p.o_positions;
        const int oh = rem / p.o_w;
        const int ow = rem - oh * p.o_w;
        if (TRANSPOSED) {
            hb = oh + g.ph;
            wb = ow + g.pw;
        } else {
            hb = oh * g.sh - g.ph;
            wb = ow * g.sw - g.pw;
        }
        const size_t offset = (size_t)b * p.batch_stride;
        a_base += offset;
        if (c_base) c_base += offset;
    }

    auto gather = [&](int k) -> float {
        if (!m_ok || k >= p.kdim) return 0.f;
        const int c = k / p.taps;
        const int tap = k - c * p.taps;
        const int i = tap / g.kw;
        const int j = tap - i * g.kw;
        int hi, wi;
        if (TRANSPOSED) {
            const int th = hb - i * g.dh;
            const int tw = wb - j * g.dw;
            if (th < 0 || tw < 0) return 0.f;
            hi = th / g.sh;
            wi = tw / g.sw;
            if (hi * g.sh != th || wi * g.sw != tw) return 0.f;
        } else {
            hi = hb + i * g.dh;
            wi = wb + j * g.dw;
            if (hi < 0 || wi < 0) return 0.f;
        }
        if (hi >= p.a_h || wi >= p.a_w) return 0.f;
        const size_t idx = (size_t)c * p.channel_stride + (size_t)(hi * p.a_w + wi) * p.position_stride;
        const float v = __ldg(a_base + idx);
        if (p.a.a_act == kActNone) return v;
        const float companion = c_base ? __ldg(c_base + idx) : 0.f;
        return operand_act(v, companion, p.a.a_act, p.a.a_slope);
    };

    // ---- B tile: Wmat (N, kdim) row-major; consecutive threads read consecutive k ----
    auto load_b = [&](int k0, int i) -> float {
        const int e = tid + i * kThreads;
        if (e >= kBK * BN) return 0.f;
        const int kl = e % kBK;
        const int nl = e / kBK;
        const int n = n0 + nl, k = k0 + kl;
        if (n >= p.o_ch || k >= p.kdim) return 0.f;
        return __ldg(p.a.wmat + (size_t)n * p.kdim + k);
    };

    const int tx = tid % (BM / TM);
    const int ty = tid / (BM / TM);

    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int s = 0; s < TN; ++s) acc[r][s] = 0.f;

    float a_reg[A_PER_THREAD], b_reg[B_PER_THREAD];
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) a_reg[i] = gather(ak0 + i * K_STEP);
#pragma unroll
    for (int i = 0; i < B_PER_THREAD; ++i) b_reg[i] = load_b(0, i);

    for (int k0 = 0; k0 < p.kdim; k0 += kBK) {
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) As[ak0 + i * K_STEP][am] = a_reg[i];
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * kThreads;
            if (e < kBK * BN) Bs[e % kBK][e / kBK] = b_reg[i];
        }
        __syncthreads();
        // prefetch the next tile into registers while this one is consumed
        if (k0 + kBK < p.kdim) {
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) a_reg[i] = gather(k0 + kBK + ak0 + i * K_STEP);
#pragma unroll
            for (int i = 0; i < B_PER_THREAD; ++i) b_reg[i] = load_b(k0 + kBK, i);
        }
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float av[TM], bv[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
            if (TM == 8) {
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][BM / 2 + tx * 4]);
                av[TM - 4] = a1.x; av[TM - 3] = a1.y; av[TM - 2] = a1.z; av[TM - 1] = a1.w;
            }
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
            bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int s = 0; s < TN; ++s) acc[r][s] = fmaf(av[r], bv[s], acc[r][s]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int ml = (r < 4) ? tx * 4 + r : BM / 2 + tx * 4 + (r - 4);
        const int m = m0 + ml;
        if (m >= p.m_total) continue;
        const int b = m / p.o_positions;
        const int rem = m - b * p.o_positions;
#pragma unroll
        for (int s = 0; s < TN; ++s) {
            const int n = n0 + ty * 4 + s;
            if (n >= p.o_ch) continue;
            const size_t idx = ((size_t)b * p.o_ch + n) * p.o_positions + rem;
            float v = acc[r][s];
            if (p.a.bias) v += __ldg(p.a.bias + n);
            if (p.a.bias2) v += __ldg(p.a.bias2 + (size_t)b * p.o_ch + n);
            if (p.a.out_act == kOutLrelu) v = leaky(v, p.a.out_slope);
            else if (p.a.out_act == kOutTanh) v = tanhf(v);
            if (p.a.mask_src) v = __ldg(p.a.mask_src + idx) > 0.f ? v : v * p.a.mask_slope;
            if (p.a.residual) v += __ldg(p.a.residual + idx);
            v *= p.a.alpha;
            if (p.a.accumulate) v += p.a.out[idx];
            p.a.out[idx] = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch_gemm_variant(const GemmParams& p, cudaStream_t stream) {
    dim3 grid(ceil_div(p.m_total, BM), ceil_div(p.o_ch, BN));
    LaunchScope scope(p.a.transposed ? "conv_dgrad_kernel" : "conv_fprop_kernel", stream);
    if (p.a.transposed)
        conv_gemm_kernel<BM, BN, TM, TN, true><<<grid, kThreads, 0, stream>>>(p);
    else
        conv_gemm_kernel<BM, BN, TM, TN, false><<<grid, kThreads, 0, stream>>>(p);
    return launched("conv_gemm_kernel");
}

int check_geometry(const pmn_conv_geometry& g, const char* what) {
    if (g.batch <= 0 || g.c_in <= 0 || g.c_out <= 0 || g.h_in <= 0 || g.w_in <= 0 ||
        g.h_out <= 0 || g.w_out <= 0 || g.kh <= 0 || g.kw <= 0 || g.sh <= 0 || g.sw <= 0 ||
        g.dh <= 0 || g.dw <= 0 || g.ph < 0 || g.pw < 0)
        return fail(PMN_ERR_ARGUMENT, std::string(what) + ": bad geometry");
    // the output extent may not exceed what the padded input supports
    if ((g.h_out - 1) * g.sh + (g.kh - 1) * g.dh + 1 > g.h_in + 2 * g.ph ||
        (g.w_out - 1) * g.sw + (g.kw - 1) * g.dw + 1 > g.w_in + 2 * g.pw)
        return fail(PMN_ERR_ARGUMENT, std::string(what) + ": output larger than the padded input allows");
    if ((int64_t)g.batch * g.h_out * g.w_out >= (int64_t)1 << 31 ||
        (int64_t)g.batch * g.h_in * g.w_in >= (int64_t)1 << 31)
        return fail(PMN_ERR_ARGUMENT, std::string(what) + ": too many positions");
    return PMN_OK;
}

// ---------------------------------------------------------------------------
// Weight gradient
// ---------------------------------------------------------------------------

struct WgradParams {
    ConvWgradArgs a;
    int taps, ncols, positions, o_positions, chunk;
};

constexpr int kWT = 64;    // tile edge (output channels x (c_in, tap) columns)
constexpr int kWP = 16;    // positions per step

__global__ void __launch_bounds__(kThreads) conv_wgrad_kernel(WgradParams p) {
    constexpr int WTP = kWT + 4;
    __shared__ __align__(16) float Ds[kWP][WTP];  // dy   [position][channel]
    __shared__ __align__(16) float Xs[kWP][WTP];  // x    [position][column]
    const pmn_conv_geometry& g = p.a.g;
    const int tid = threadIdx.x;
    const int col0 = blockIdx.x * kWT;
    const int n0 = blockIdx.y * kWT;
    const int first = blockIdx.z * p.chunk;
    const int last = min(first + p.chunk, p.positions);

    // loader role: position tid % 16, channels / columns tid / 16 + 16 i
    const int lp = tid % kWP;
    const int lq = tid / kWP;
    // the columns' (channel, tap) decode never changes
    int col_c[4], col_h[4], col_w[4];
    bool col_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = col0 + lq + 16 * i;
        col_ok[i] = col < p.ncols;
        const int cc = col_ok[i] ? col : 0;
        const int c = cc / p.taps;
        const int tap = cc - c * p.taps;
        const int ti = tap / g.kw;
        col_c[i] = c;
        col_h[i] = ti * g.dh - g.ph;
        col_w[i] = (tap - ti * g.kw) * g.dw - g.pw;
    }
    const size_t x_plane = (size_t)g.h_in * g.w_in;

    auto load = [&](int pos0, float* d_reg, float* x_reg) {
        const int m = pos0 + lp;
        const bool ok = m < last;
        const int mm = ok ? m : 0;
        const int b = mm / p.o_positions;
        const int rem = mm - b * p.o_positions;
        const int oh = rem / g.w_out;
        const int ow = rem - oh * g.w_out;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + lq + 16 * i;
            float v = 0.f;
            if (ok && n < g.c_out) {
                const size_t idx = ((size_t)b * g.c_out + n) * p.o_positions + rem;
                v = __ldg(p.a.dy + idx);
                if (p.a.dy_act != kActNone)
                    v = operand_act(v, p.a.dy_companion ? __ldg(p.a.dy_companion + idx) : 0.f,
                                    p.a.dy_act, p.a.dy_slope);
            }
            d_reg[i] = v;
        }
        const int hb = oh * g.sh, wb = ow * g.sw;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = 0.f;
            const int hi = hb + col_h[i], wi = wb + col_w[i];
            if (ok && col_ok[i] && hi >= 0 && hi < g.h_in && wi >= 0 && wi < g.w_in) {
                const size_t idx = ((size_t)b * g.c_in + col_c[i]) * x_plane + (size_t)hi * g.w_in + wi;
                v = __ldg(p.a.x + idx);
                if (p.a.x_act != kActNone)
                    v = operand_act(v, p.a.x_companion ? __ldg(p.a.x_companion + idx) : 0.f,
                                    p.a.x_act, p.a.x_slope);
            }
            x_reg[i] = v;
        }
    };

    const int tx = tid % 16;  // columns tx * 4 ..
    const int ty = tid / 16;  // channels ty * 4 ..
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[r][s] = 0.f;
    float bias_sum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool do_bias = p.a.gbias != nullptr && blockIdx.x == 0 && tx == 0;

    float d_reg[4], x_reg[4];
    if (first < last) load(first, d_reg, x_reg);
    for (int pos0 = first; pos0 < last; pos0 += kWP) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            Ds[lp][lq + 16 * i] = d_reg[i];
            Xs[lp][lq + 16 * i] = x_reg[i];
        }
        __syncthreads();
        if (pos0 + kWP < last) load(pos0 + kWP, d_reg, x_reg);
#pragma unroll
        for (int kk = 0; kk < kWP; ++kk) {
            const float4 d = *reinterpret_cast<const float4*>(&Ds[kk][ty * 4]);
            const float4 x = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
            const float dv[4] = {d.x, d.y, d.z, d.w};
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) acc[r][s] = fmaf(dv[r], xv[s], acc[r][s]);
            if (do_bias) {
#pragma unroll
                for (int r = 0; r < 4; ++r) bias_sum[r] += dv[r];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = n0 + ty * 4 + r;
        if (n >= g.c_out) continue;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int col = col0 + tx * 4 + s;
            if (col < p.ncols && acc[r][s] != 0.f)
                atomicAdd(p.a.gw + (size_t)n * p.ncols + col, acc[r][s]);
        }
        if (do_bias && bias_sum[r] != 0.f) atomicAdd(p.a.gbias + n, bias_sum[r]);
    }
}

// (dim0, dim1, taps) -> (dim1, dim0, taps)
__global__ void transpose_weight_kernel(
    const float* __restrict__ w, float* __restrict__ wt, int dim0, int dim1, int taps) {
    const size_t total = (size_t)dim0 * dim1 * taps;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % taps);
        const size_t rest = idx / taps;
        const int a = (int)(rest % dim0);
        const int b = (int)(rest / dim0);
        wt[idx] = w[((size_t)a * dim1 + b) * taps + tap];
    }
}

// Backward of w = g v / ||v|| over rows (torch.nn.utils.weight_norm dim=0, model/core.py:43-45):
//   gg = <gw, v> / ||v||,   gv = g / ||v|| (gw - v <gw, v> / ||v||^2)
__device__ __forceinline__ void weight_norm_backward_row(
    const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ gw,
    float* __restrict__ gv, float* __restrict__ gg, int inner, int row_index) {
    __shared__ float partial[2][32];
    const size_t row = (size_t)row_index * inner;
    float norm2 = 0.f, dot = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
        const float vi = v[row + i];
        norm2 = fmaf(vi, vi, norm2);
        dot = fmaf(vi, gw[row + i], dot);
    }
    for (int offset = 16; offset > 0; offset >>= 1) {
        norm2 += __shfl_xor_sync(0xffffffffu, norm2, offset);
        dot += __shfl_xor_sync(0xffffffffu, dot, offset);
    }
    if ((threadIdx.x & 31) == 0) {
        partial[0][threadIdx.x >> 5] = norm2;
        partial[1][threadIdx.x >> 5] = dot;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float a = threadIdx.x < (blockDim.x >> 5) ? partial[0][threadIdx.x] : 0.f;
        float b = threadIdx.x < (blockDim.x >> 5) ? partial[1][threadIdx.x] : 0.f;
        for (int offset = 16; offset > 0; offset >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, offset);
            b += __shfl_xor_sync(0xffffffffu, b, offset);
        }
        if (threadIdx.x == 0) {
            partial[0][0] = a;
            partial[1][0] = b;
        }
    }
    __syncthreads();
    norm2 = partial[0][0];
    dot = partial[1][0];
    const float inv = rsqrtf(norm2);
    const float scale = g[row_index] * inv;
    const float project = dot / norm2;
    for (int i = threadIdx.x; i < inner; i += blockDim.x)
        gv[row + i] = scale * (gw[row + i] - v[row + i] * project);
    if (threadIdx.x == 0) gg[row_index] = dot * inv;
}

__global__ void __launch_bounds__(256) weight_norm_backward_kernel(
    const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ gw,
    float* __restrict__ gv, float* __restrict__ gg, int inner) {
    weight_norm_backward_row(v, g, gw, gv, gg, inner, blockIdx.x);
}

// The same for every weight-normed convolution of a module in one launch (block (row, layer); rows
// past a layer's dim0 leave at once): 132 launches of a few microseconds each per training step were
// 1.25 ms of launch latency
__global__ void __launch_bounds__(256) weight_norm_backward_table_kernel(const pmn_weight_norm_desc* table) {
    const pmn_weight_norm_desc d = table[blockIdx.y];
    if ((int)blockIdx.x >= d.dim0) return;
    weight_norm_backward_row(d.v, d.g, d.gw, d.gv, d.gg, d.inner, blockIdx.x);
}

}  // namespace

int launch_conv_gemm(const ConvGemmArgs& args, cudaStream_t stream) {
    const pmn_conv_geometry& g = args.g;
    PMN_TRY(check_geometry(g, "conv_gemm"));
    PMN_REQUIRE(args.a && args.wmat && args.out, "conv_gemm: null pointer");
    PMN_REQUIRE(args.a_act == kActNone || args.a_act == kActLrelu || args.a_companion,
                "conv_gemm: this operand activation needs a companion tensor");
    GemmParams p;
    p.a = args;
    if (args.transposed) {
        p.a_ch = g.c_out; p.a_h = g.h_out; p.a_w = g.w_out;
        p.o_ch = g.c_in; p.o_h = g.h_in; p.o_w = g.w_in;
    } else {
        p.a_ch = g.c_in; p.a_h = g.h_in; p.a_w = g.w_in;
        p.o_ch = g.c_out; p.o_h = g.h_out; p.o_w = g.w_out;
    }
    p.taps = g.kh * g.kw;
    p.kdim = p.a_ch * p.taps;
    p.o_positions = p.o_h * p.o_w;
    p.m_total = g.batch * p.o_positions;
    const bool strided = g.channel_stride || g.position_stride || g.batch_stride;
    PMN_REQUIRE(!strided || (!args.transposed && g.channel_stride > 0 && g.position_stride > 0 &&
                             g.batch_stride > 0), "conv_gemm: bad tensor strides");
    p.channel_stride = strided ? g.channel_stride : p.a_h * p.a_w;
    p.position_stride = strided ? g.position_stride : 1;
    p.batch_stride = strided ? (size_t)g.batch_stride : (size_t)p.a_ch * p.a_h * p.a_w;
    if (p.o_ch <= 16) return launch_gemm_variant<256, 16, 4, 4>(p, stream);
    return launch_gemm_variant<128, 64, 8, 4>(p, stream);
}

int launch_conv_wgrad(const ConvWgradArgs& args, cudaStream_t stream) {
    const pmn_conv_geometry& g = args.g;
    PMN_TRY(check_geometry(g, "conv_wgrad"));
    PMN_REQUIRE(args.dy && args.x && args.gw, "conv_wgrad: null pointer");
    PMN_REQUIRE(args.dy_act == kActNone || args.dy_act == kActLrelu || args.dy_companion,
                "conv_wgrad: this dy activation needs a companion tensor");
    PMN_REQUIRE(args.x_act == kActNone || args.x_act == kActLrelu || args.x_companion,
                "conv_wgrad: this x activation needs a companion tensor");
    WgradParams p;
    p.a = args;
    p.taps = g.kh * g.kw;
    p.ncols = g.c_in * p.taps;
    p.o_positions = g.h_out * g.w_out;
    p.positions = g.batch * p.o_positions;
    const int tiles = ceil_div(p.ncols, kWT) * ceil_div(g.c_out, kWT);
    // enough CTAs for 148 SMs x 3, but at least 4 steps of positions per split
    int splits = max(1, min(ceil_div(444, tiles), ceil_div(p.positions, 4 * kWP)));
    splits = min(splits, 65535);
    p.chunk = ceil_div(ceil_div(p.positions, splits), kWP) * kWP;
    splits = ceil_div(p.positions, p.chunk);
    dim3 grid(ceil_div(p.ncols, kWT), ceil_div(g.c_out, kWT), splits);
    PMN_REQUIRE(grid.y <= 65535, "conv_wgrad: too many output channels");
    LaunchScope scope("conv_wgrad_kernel", stream);
    conv_wgrad_kernel<<<grid, kThreads, 0, stream>>>(p);
    return launched("conv_wgrad_kernel");
}

int launch_transpose_weight(
    const float* w, float* wt, int dim0, int dim1, int taps, cudaStream_t stream) {
    PMN_REQUIRE(w && wt && dim0 > 0 && dim1 > 0 && taps > 0, "transpose_weight: bad argument");
    const size_t total = (size_t)dim0 * dim1 * taps;
    const int blocks = (int)min((size_t)2048, (total + 255) / 256);
    LaunchScope scope("transpose_weight_kernel", stream);
    transpose_weight_kernel<<<blocks, 256, 0, stream>>>(w, wt, dim0, dim1, taps);
    return launched("transpose_weight_kernel");
}

int launch_weight_norm_backward_table(
    const pmn_weight_norm_desc* table, int layers, int max_dim0, cudaStream_t stream) {
    PMN_REQUIRE(table && layers > 0 && layers <= 65535 && max_dim0 > 0, "weight_norm_backward_table: bad argument");
    LaunchScope scope("weight_norm_backward_table_kernel", stream);
    weight_norm_backward_table_kernel<<<dim3(max_dim0, layers), 256, 0, stream>>>(table);
    return launched("weight_norm_backward_table_kernel");
}

int launch_weight_norm_backward(
    const float* v, const float* g, const float* gw, float* gv, float* gg, int dim0, int inner,
    cudaStream_t stream) {
    PMN_REQUIRE(v && g && gw && gv && gg && dim0 > 0 && inner > 0, "weight_norm_backward: bad argument");
    LaunchScope scope("weight_norm_backward_kernel", stream);
    weight_norm_backward_kernel<<<dim0, 256, 0, stream>>>(v, g, gw, gv, gg, inner);
    return launched("weight_norm_backward_kernel");
}

}  // namespace pmn

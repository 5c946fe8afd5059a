// Training-step convolutions on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as conv_gemm_kernel (train_conv.cu): forward / data-gradient
// implicit GEMM of a strided, dilated 2-D convolution with the activations fused
// into the operand load and the epilogue.  Here the products run as
// tcgen05.mma.kind::tf32 (10-bit mantissa operands, fp32 accumulation in TMEM):
// the training step of the reference runs under fp16 autocast
// (promonet/train/core.py:220,262), so tf32 operands are the more precise choice.
//
// GEMM view: M = 128 flattened (batch, position) rows per CTA (TMEM lanes),
// N = BN output channels (TMEM columns), K = taps x input channels, tap-major
// ("per-tap implicit GEMM"): a K step is 32 consecutive channels at one tap, so a
// row's 32 operands share one spatial offset and one bounds check.  Weights are
// pre-packed tap-major (pack_weight_taps_kernel), so their tile is 16-byte loads.
//
// Per K step all 256 threads gather the A tile (128 x 32, activation applied,
// rounded to tf32) and the B tile (BN x 32) into shared memory in the un-swizzled
// K-major core-matrix layout [k / 4][row][4 x fp32] (16-byte rows, exactly what the
// shared-memory descriptor addresses), make the writes visible to the async proxy,
// and one elected thread issues four K = 8 MMAs and commits them to an mbarrier.
// Two stages: the gather of step i + 1 overlaps the MMAs of step i.  The epilogue
// reads the accumulators with tcgen05.ld (warp w owns lanes 32 (w % 4) ..) and
// applies bias / activation / mask / residual / accumulate exactly like the fp32 path.
#include "train.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 256;
constexpr int kBM = 128;
constexpr int kKStep = 32;            // fp32 operands per row per stage
constexpr int kChunks = kKStep / 4;   // 16-byte chunks per row per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t address = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(address), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_load16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no swizzle: [0,14) start >> 4 | [16,30) byte offset between the two 16-byte
// K chunks of one MMA >> 4 | [32,46) byte offset between 8-row groups >> 4 | version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t address, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((address & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// D fp32 (1 << 4), A and B tf32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float operand_act(float v, float companion, int act, float slope) {
    if (act == kActLrelu) return leaky(v, slope);
    if (act == kActLreluMask) return companion > 0.f ? v : v * slope;
    if (act == kActTanhMask) return v * (1.f - companion * companion);
    return v;
}

struct TcParams {
    ConvGemmArgs a;        // a.wmat = packed weights (o_ch, taps, c_pad)
    int a_ch, a_h, a_w;    // gathered tensor
    int o_ch, o_h, o_w;    // produced tensor
    int taps, c_pad, m_total, o_positions;
};

template <int BN, bool TRANSPOSED>
__global__ void __launch_bounds__(kThreads) conv_gemm_tc_kernel(TcParams p) {
    constexpr uint32_t kABytes = kBM * kKStep * 4;
    constexpr uint32_t kBBytes = BN * kKStep * 4;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mma_done[2];
    __shared__ uint64_t acc_done;
    __shared__ uint32_t tmem_slot;
    auto a_stage = [&](int s) { return smem + s * kABytes; };
    auto b_stage = [&](int s) { return smem + 2 * kABytes + s * kBBytes; };

    const pmn_conv_geometry& g = p.a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;

    if (tid == 0) {
        mbar_init(mma_done + 0, 1);
        mbar_init(mma_done + 1, 1);
        mbar_init(&acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // ---- A gather role: row tid % 128, channels (tid / 128) * 16 .. + 16 of each K step ----
    const int arow = tid & (kBM - 1);
    const int ahalf = tid >> 7;
    const int m_load = m0 + arow;
    const bool m_ok = m_load < p.m_total;
    int hb, wb;
    const float* a_base = p.a.a;
    const float* c_base = p.a.a_companion;
    {
        const int mm = m_ok ? m_load : 0;
        const int b = mm / p.o_positions;
        const int rem = mm - b * p.o_positions;
        const int oh = rem / p.o_w, ow = rem - oh * p.o_w;
        if (TRANSPOSED) { hb = oh + g.ph; wb = ow + g.pw; }
        else { hb = oh * g.sh - g.ph; wb = ow * g.sw - g.pw; }
        const size_t offset = (size_t)b * p.a_ch * p.a_h * p.a_w;
        a_base += offset;
        if (c_base) c_base += offset;
    }
    const size_t plane = (size_t)p.a_h * p.a_w;
    const int blocks_per_tap = p.c_pad / kKStep;
    const int k_steps = p.taps * blocks_per_tap;
    constexpr uint32_t idesc = instr_desc_tf32(kBM, BN);

    int tap = 0, cb = 0;        // position of the NEXT step to gather
    bool tap_ok = false;
    size_t tap_offset = 0;
    auto enter_tap = [&]() {
        const int i = tap / g.kw, j = tap - i * g.kw;
        int hi, wi;
        bool ok = m_ok;
        if (TRANSPOSED) {
            const int th = hb - i * g.dh, tw = wb - j * g.dw;
            hi = th / g.sh; wi = tw / g.sw;
            ok = ok && th >= 0 && tw >= 0 && hi * g.sh == th && wi * g.sw == tw;
        } else {
            hi = hb + i * g.dh; wi = wb + j * g.dw;
            ok = ok && hi >= 0 && wi >= 0;
        }
        ok = ok && hi < p.a_h && wi < p.a_w;
        tap_ok = ok;
        tap_offset = ok ? (size_t)hi * p.a_w + wi : 0;
    };
    enter_tap();

    for (int step = 0; step < k_steps; ++step) {
        const int s = step & 1;
        if (step >= 2) mbar_wait(mma_done + s, ((step >> 1) - 1) & 1);
        // ---- gather A: 16 channels of this row ----
        {
            const int c_first = cb * kKStep + ahalf * 16;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = c_first + i;
                float value = 0.f;
                if (tap_ok && c < p.a_ch) {
                    const size_t idx = (size_t)c * plane + tap_offset;
                    value = __ldg(a_base + idx);
                    if (p.a.a_act != kActNone)
                        value = operand_act(value, c_base ? __ldg(c_base + idx) : 0.f,
                                            p.a.a_act, p.a.a_slope);
                }
                v[i] = to_tf32(value);
            }
            float4* dst = reinterpret_cast<float4*>(a_stage(s));
#pragma unroll
            for (int q = 0; q < 4; ++q)
                dst[(ahalf * 4 + q) * kBM + arow] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        // ---- gather B: warp w takes 16-byte chunk w of rows lane, lane + 32, ... ----
        {
            const size_t k_offset = (size_t)tap * p.c_pad + cb * kKStep + warp * 4;
            const size_t row_stride = (size_t)p.taps * p.c_pad;
            float4* dst = reinterpret_cast<float4*>(b_stage(s));
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) {
                const int n = lane + 32 * i;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + n < p.o_ch)
                    w = __ldg(reinterpret_cast<const float4*>(p.a.wmat + (size_t)(n0 + n) * row_stride + k_offset));
                dst[warp * BN + n] = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
            }
        }
        // advance to the next step's (tap, channel block)
        if (++cb == blocks_per_tap) {
            cb = 0;
            if (++tap < p.taps) enter_tap();
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_stage(s)), b_addr = smem_u32(b_stage(s));
#pragma unroll
            for (int kk = 0; kk < kKStep / 8; ++kk) {
                const uint64_t a_desc = smem_desc(a_addr + 2 * kk * kBM * 16, kBM * 16, 128);
                const uint64_t b_desc = smem_desc(b_addr + 2 * kk * BN * 16, BN * 16, 128);
                tc_mma_tf32(tmem_base, a_desc, b_desc, idesc, (step > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(mma_done + s);
            if (step == k_steps - 1) tc_commit(&acc_done);
        }
    }

    // ---- epilogue ----
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    {
        const int quad = warp & 3, half = warp >> 2;
        const int m = m0 + quad * 32 + lane;
        const bool ok = m < p.m_total;
        const int mm = ok ? m : 0;
        const int b = mm / p.o_positions;
        const int rem = mm - b * p.o_positions;
#pragma unroll 1
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
            uint32_t raw[16];
            __syncwarp();
            tc_load16(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, raw);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + c0 + i;
                if (!ok || n >= p.o_ch) continue;
                const size_t idx = ((size_t)b * p.o_ch + n) * p.o_positions + rem;
                float v = __uint_as_float(raw[i]);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// w (d0, d1, taps) -> tap-major GEMM rows with the reduction channels padded to 32:
//   transposed = 0: out[a][tap][b] (rows d0, reduce over d1: forward of a Conv, dgrad of a ConvTranspose)
//   transposed = 1: out[b][tap][a] (rows d1, reduce over d0: data gradient of a Conv)
__global__ void pack_weight_taps_kernel(
    const float* __restrict__ w, float* __restrict__ out, int d0, int d1, int taps, int transposed,
    int c_pad) {
    const int rows = transposed ? d1 : d0;
    const size_t total = (size_t)rows * taps * c_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % c_pad);
        const size_t rest = idx / c_pad;
        const int tap = (int)(rest % taps);
        const int row = (int)(rest / taps);
        const int reduce = transposed ? d0 : d1;
        float v = 0.f;
        if (c < reduce) {
            const int a = transposed ? c : row, b = transposed ? row : c;
            v = w[((size_t)a * d1 + b) * taps + tap];
        }
        out[idx] = v;
    }
}

template <int BN>
int launch_variant(const TcParams& p, cudaStream_t stream) {
    const size_t smem = 2 * (kBM * kKStep * 4) + 2 * (BN * kKStep * 4);
    static bool configured[2] = {false, false};
    const int t = p.a.transposed ? 1 : 0;
    if (!configured[t]) {
        cudaError_t error = t
            ? cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PMN_TRY(check_cuda(error, "conv_gemm_tc smem attribute"));
        configured[t] = true;
    }
    dim3 grid(ceil_div(p.m_total, kBM), ceil_div(p.o_ch, BN));
    PMN_REQUIRE(grid.y <= 65535, "conv_gemm_tc: too many output channels");
    LaunchScope scope(t ? "conv_dgrad_tc_kernel" : "conv_fprop_tc_kernel", stream);
    if (t) conv_gemm_tc_kernel<BN, true><<<grid, kThreads, smem, stream>>>(p);
    else conv_gemm_tc_kernel<BN, false><<<grid, kThreads, smem, stream>>>(p);
    return launched("conv_gemm_tc_kernel");
}

}  // namespace

int conv_tc_channel_pad(int channels) { return (channels + kKStep - 1) / kKStep * kKStep; }

int launch_pack_weight_taps(
    const float* w, float* out, int d0, int d1, int taps, int transposed, cudaStream_t stream) {
    PMN_REQUIRE(w && out && d0 > 0 && d1 > 0 && taps > 0, "pack_weight_taps: bad argument");
    const int c_pad = conv_tc_channel_pad(transposed ? d0 : d1);
    const size_t total = (size_t)(transposed ? d1 : d0) * taps * c_pad;
    const int blocks = (int)min((size_t)2048, (total + 255) / 256);
    LaunchScope scope("pack_weight_taps_kernel", stream);
    pack_weight_taps_kernel<<<blocks, 256, 0, stream>>>(w, out, d0, d1, taps, transposed, c_pad);
    return launched("pack_weight_taps_kernel");
}

int launch_conv_gemm_tc(const ConvGemmArgs& args, cudaStream_t stream) {
    const pmn_conv_geometry& g = args.g;
    PMN_REQUIRE(args.a && args.wmat && args.out, "conv_gemm_tc: null pointer");
    PMN_REQUIRE(g.batch > 0 && g.c_in > 0 && g.c_out > 0 && g.h_in > 0 && g.w_in > 0 &&
                g.h_out > 0 && g.w_out > 0 && g.kh > 0 && g.kw > 0 && g.sh > 0 && g.sw > 0 &&
                g.dh > 0 && g.dw > 0 && g.ph >= 0 && g.pw >= 0, "conv_gemm_tc: bad geometry");
    PMN_REQUIRE((g.h_out - 1) * g.sh + (g.kh - 1) * g.dh + 1 <= g.h_in + 2 * g.ph &&
                (g.w_out - 1) * g.sw + (g.kw - 1) * g.dw + 1 <= g.w_in + 2 * g.pw,
                "conv_gemm_tc: output larger than the padded input allows");
    PMN_REQUIRE(args.a_act == kActNone || args.a_act == kActLrelu || args.a_companion,
                "conv_gemm_tc: this operand activation needs a companion tensor");
    TcParams p;
    p.a = args;
    if (args.transposed) {
        p.a_ch = g.c_out; p.a_h = g.h_out; p.a_w = g.w_out;
        p.o_ch = g.c_in; p.o_h = g.h_in; p.o_w = g.w_in;
    } else {
        p.a_ch = g.c_in; p.a_h = g.h_in; p.a_w = g.w_in;
        p.o_ch = g.c_out; p.o_h = g.h_out; p.o_w = g.w_out;
    }
    p.taps = g.kh * g.kw;
    p.c_pad = conv_tc_channel_pad(p.a_ch);
    p.o_positions = p.o_h * p.o_w;
    PMN_REQUIRE((int64_t)g.batch * p.o_positions < ((int64_t)1 << 31), "conv_gemm_tc: too many positions");
    p.m_total = g.batch * p.o_positions;
    if (p.o_ch > 64) return launch_variant<128>(p, stream);
    if (p.o_ch > 32) return launch_variant<64>(p, stream);
    return launch_variant<32>(p, stream);
}

}  // namespace pmn

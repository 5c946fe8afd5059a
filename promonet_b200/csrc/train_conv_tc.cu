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

constexpr int kProducers = 512;        // gather / epilogue threads
constexpr int kThreads = kProducers + 64;  // + the MMA issuing warp + the weight-copy warp
constexpr int kStages = 3;             // pipeline stages (2 for the 256-column tiles: 96 KB each)
constexpr int kBM = 128;
constexpr int kMaxPhases = 9;          // stride (3, 3) at most
constexpr int kKStep = 32;            // fp32 operands per row per K step (one tap, 32 channels)
constexpr int kPair = 2;              // K steps per pipeline stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

long long* g_debug_counters = nullptr;   // pmn_debug_train_tc_counters

struct TcParams {
    long long* debug;      // 8 cycle counters written by CTA 0 (profiling aid) or null
    ConvGemmArgs a;        // a.wmat = packed weights (o_ch, taps, c_pad)
    int a_ch, a_h, a_w;    // gathered tensor
    int o_ch, o_h, o_w;    // produced tensor
    int taps, c_pad, m_total, o_positions;
    int channel_stride, position_stride;
    size_t batch_stride;
    // Strided data gradient: output rows grouped by phase ((h + ph) mod sh, (w + pw) mod sw) so that
    // a tile only walks the taps that can reach it (1 / (sh sw) of them).  phases = 0: off.
    int phases;
    int phase_tile_start[kMaxPhases + 1];
    int phase_h0[kMaxPhases], phase_w0[kMaxPhases];   // first output row / column of the phase
    int phase_qh[kMaxPhases], phase_qw[kMaxPhases];   // rows / columns of the phase
};

__device__ float g_zero_words[4] = {0.f, 0.f, 0.f, 0.f};  // what rows outside the input read

template <int BN, bool TRANSPOSED, int A_ACT>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_tc_kernel(TcParams p) {
    constexpr int kNStages = BN > 128 ? 2 : kStages;
    constexpr uint32_t kABytes = kBM * kKStep * 4;            // one K step of A
    // BN = 256: two 128-row weight tiles share one gathered A tile (two N = 128 MMAs per K = 8)
    constexpr int kSlabRows = BN > 128 ? 128 : BN;
    constexpr int kSlabs = BN / kSlabRows;
    constexpr uint32_t kSlabBytes = kSlabRows * kKStep * 4;   // one packed slab: one K step of one tile
    constexpr uint32_t kBBytes = kSlabs * kSlabBytes;
    constexpr uint32_t kStageBytes = kPair * (kABytes + kBBytes);
    constexpr int kAPer = kBM * kKStep / kProducers;          // A operands per thread per K step (8)
    constexpr bool kCompanion = A_ACT == kActLreluMask || A_ACT == kActTanhMask;
    constexpr int kMaxTaps = 64;                              // taps with a shared-memory offset table
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[kStages];    // 16 producer warps + the weight copy (with its bytes)
    __shared__ uint64_t empty[kStages];   // tcgen05.commit: the MMAs that read the stage are done
    __shared__ uint64_t acc_done;
    __shared__ uint32_t tmem_slot;
    __shared__ int2 tap_offsets[kMaxTaps];   // (i dh, j dw) per tap

    const pmn_conv_geometry& g = p.a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * BN;
    constexpr int kMmaWarp = kProducers / 32, kCopyWarp = kMmaWarp + 1;
    __shared__ int tap_ids[kMaxTaps];        // the taps this tile walks (all, or those of its phase)
    __shared__ int tap_count;

    // Row mapping: tile-local row -> (batch item, output position).  Phased: rows enumerate
    // (item, q_h, q_w) of one phase, position = (h0 + q_h sh, w0 + q_w sw).
    int phase = -1, m0 = blockIdx.x * kBM, rows_total = p.m_total;
    if (TRANSPOSED && p.phases > 0) {
        phase = 0;
        while (phase + 1 < p.phases && (int)blockIdx.x >= p.phase_tile_start[phase + 1]) ++phase;
        m0 = ((int)blockIdx.x - p.phase_tile_start[phase]) * kBM;
        rows_total = p.a.g.batch * p.phase_qh[phase] * p.phase_qw[phase];
    }
    auto decode_row = [&](int m, int* b, int* oh, int* ow) {
        if (phase < 0) {
            *b = m / p.o_positions;
            const int rem = m - *b * p.o_positions;
            *oh = rem / p.o_w;
            *ow = rem - *oh * p.o_w;
        } else {
            const int per_item = p.phase_qh[phase] * p.phase_qw[phase];
            *b = m / per_item;
            const int rem = m - *b * per_item;
            const int qh = rem / p.phase_qw[phase];
            *oh = p.phase_h0[phase] + qh * p.a.g.sh;
            *ow = p.phase_w0[phase] + (rem - qh * p.phase_qw[phase]) * p.a.g.sw;
        }
    };

    if (tid == 0) {
        for (int i = 0; i < kNStages; ++i) {
            mbar_init(full + i, kProducers / 32 + 1);
            mbar_init(empty + i, 1);
        }
        mbar_init(&acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) {
        // taps in order; a phased tile keeps those with i = (h + ph) mod sh, j = (w + pw) mod sw
        int count = 0;
        const int rh = phase < 0 ? 0 : (p.phase_h0[phase] + g.ph) % g.sh;
        const int rw = phase < 0 ? 0 : (p.phase_w0[phase] + g.pw) % g.sw;
        for (int t = 0; t < min(p.taps, kMaxTaps); ++t) {
            const int i = t / g.kw, j = t - i * g.kw;
            if (phase >= 0 && (i % g.sh != rh || j % g.sw != rw)) continue;
            tap_ids[count] = t;
            tap_offsets[count] = make_int2(i * g.dh, j * g.dw);
            ++count;
        }
        tap_count = p.taps > kMaxTaps ? p.taps : count;
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int blocks_per_tap = p.c_pad / kKStep;
    const int taps = tap_count;
    const int k_steps = taps * blocks_per_tap;
    const int stages_total = (k_steps + kPair - 1) / kPair;
    const bool timing = p.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    const long long begin = timing ? clock64() : 0;
    // stage layout: [A step 0][A step 1][B tile 0: step 0, step 1][B tile 1: step 0, step 1]
    auto stage_a = [&](int s, int sub) { return smem + s * kStageBytes + sub * kABytes; };
    auto stage_b = [&](int s, int tile, int sub) {
        return smem + s * kStageBytes + kPair * kABytes + (tile * kPair + sub) * kSlabBytes;
    };

    if (warp == kMmaWarp) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(kBM, kSlabRows);
            long long t_wait = 0, mark = 0;
            for (int it = 0; it < stages_total; ++it) {
                const int s = it % kNStages;
                if (timing) mark = clock64();
                mbar_wait(full + s, (it / kNStages) & 1);
                if (timing) t_wait += clock64() - mark;
                tc_fence_after();
                const int subs = min(kPair, k_steps - it * kPair);
                for (int sub = 0; sub < subs; ++sub) {
                    const uint32_t a_addr = smem_u32(stage_a(s, sub));
#pragma unroll
                    for (int kk = 0; kk < kKStep / 8; ++kk) {
                        const uint64_t a_desc = smem_desc(a_addr + 2 * kk * kBM * 16, kBM * 16, 128);
#pragma unroll
                        for (int tile = 0; tile < kSlabs; ++tile) {
                            const uint32_t b_addr = smem_u32(stage_b(s, tile, sub));
                            const uint64_t b_desc =
                                smem_desc(b_addr + 2 * kk * kSlabRows * 16, kSlabRows * 16, 128);
                            tc_mma_tf32(tmem_base + tile * kSlabRows, a_desc, b_desc, idesc,
                                        (it > 0 || sub > 0 || kk > 0) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(empty + s);
            }
            tc_commit(&acc_done);
            if (timing) { p.debug[5] = t_wait; p.debug[6] = k_steps; }
        }
    } else if (warp == kCopyWarp) {
        // ===== weight copy: one thread, one bulk copy per stage (slabs are stored in K-step order) =====
        if (lane == 0) {
            constexpr size_t kSlabFloats = kSlabRows * kKStep;
            const size_t all_steps = (size_t)p.taps * blocks_per_tap;   // slabs per 128-row tile
            const float* slabs = p.a.wmat + (size_t)blockIdx.y * kSlabs * all_steps * kSlabFloats;
            for (int it = 0; it < stages_total; ++it) {
                const int s = it % kNStages;
                if (it >= kNStages) mbar_wait(empty + s, ((it / kNStages) - 1) & 1);
                const int subs = min(kPair, k_steps - it * kPair);
                mbar_expect_tx(full + s, kSlabs * subs * kSlabBytes);
                for (int sub = 0; sub < subs; ++sub) {
                    // K step -> (tap of this tile's list, channel block) -> slab of the packed weight
                    const int step = it * kPair + sub;
                    const int k = step / blocks_per_tap, cb = step - k * blocks_per_tap;
                    const size_t slab = (size_t)(p.taps > kMaxTaps ? k : tap_ids[k]) * blocks_per_tap + cb;
#pragma unroll
                    for (int tile = 0; tile < kSlabs; ++tile)
                        bulk_copy(stage_b(s, tile, sub), slabs + ((size_t)tile * all_steps + slab) * kSlabFloats,
                                  kSlabBytes, full + s);
                }
            }
        }
    } else {
        // ===== producers: 512 threads gather the A tile of every K step =====
        // row tid % 128, channels (tid / 128) * 8 .. + 8 of the step's 32 (two 16-byte chunks)
        const int arow = tid & (kBM - 1);
        const int aquarter = tid >> 7;
        const int m_load = m0 + arow;
        const bool m_ok = m_load < rows_total;
        int hb, wb;
        const float* a_base = p.a.a;
        const float* c_base = kCompanion ? p.a.a_companion : nullptr;
        {
            int b, oh, ow;
            decode_row(m_ok ? m_load : 0, &b, &oh, &ow);
            if (TRANSPOSED) { hb = oh + g.ph; wb = ow + g.pw; }
            else { hb = oh * g.sh - g.ph; wb = ow * g.sw - g.pw; }
            const size_t offset = (size_t)b * p.batch_stride;
            a_base += offset;
            if (kCompanion) c_base += offset;
        }
        const int plane = p.channel_stride;
        const bool padded = p.c_pad != p.a_ch;     // only then can a channel index run past the tensor
        const bool unit_stride = g.sh == 1 && g.sw == 1;
        const int c_thread = aquarter * kAPer;

        // (tap, channel block) of the K step being LOADED.  A row that falls outside the input at
        // this tap reads a word of zeros with stride 0, so the loop has no per-element select.
        int tap = 0, cb = 0, loaded = 0;
        const float* tap_a = nullptr;
        const float* tap_c = nullptr;
        int tap_stride = 0;
        auto enter_tap = [&]() {
            int2 o;
            if (p.taps <= kMaxTaps) o = tap_offsets[tap];
            else { const int i = tap / g.kw; o = make_int2(i * g.dh, (tap - i * g.kw) * g.dw); }
            int hi, wi;
            bool ok = m_ok;
            if (TRANSPOSED) {
                hi = hb - o.x; wi = wb - o.y;
                ok = ok && hi >= 0 && wi >= 0;
                if (!unit_stride) {
                    const int th = hi, tw = wi;
                    hi = th / g.sh; wi = tw / g.sw;
                    ok = ok && hi * g.sh == th && wi * g.sw == tw;
                }
            } else {
                hi = hb + o.x; wi = wb + o.y;
                ok = ok && hi >= 0 && wi >= 0;
            }
            ok = ok && hi < p.a_h && wi < p.a_w;
            const int offset = ok ? (hi * p.a_w + wi) * p.position_stride : 0;
            tap_a = ok ? a_base + offset : g_zero_words;
            if (kCompanion) tap_c = ok ? c_base + offset : g_zero_words;
            tap_stride = ok ? plane : 0;
        };
        enter_tap();

        // issue the loads of the next K step (if any) into the given registers, then advance
        auto issue_loads = [&](float (&va)[kAPer], float (&vc)[kAPer]) {
            if (loaded >= k_steps) return;
            const int c_first = cb * kKStep + c_thread;
            if (padded && cb == blocks_per_tap - 1) {
#pragma unroll
                for (int i = 0; i < kAPer; ++i) {
                    const int c = min(c_first + i, p.a_ch - 1);   // the weights of the padding are zero
                    va[i] = __ldg(tap_a + (long long)c * tap_stride);
                    if (kCompanion) vc[i] = __ldg(tap_c + (long long)c * tap_stride);
                }
            } else {
                const float* src = tap_a + (long long)c_first * tap_stride;
#pragma unroll
                for (int i = 0; i < kAPer; ++i) va[i] = __ldg(src + (long long)i * tap_stride);
                if (kCompanion) {
                    const float* src_c = tap_c + (long long)c_first * tap_stride;
#pragma unroll
                    for (int i = 0; i < kAPer; ++i) vc[i] = __ldg(src_c + (long long)i * tap_stride);
                }
            }
            ++loaded;
            if (++cb == blocks_per_tap) {
                cb = 0;
                if (++tap < taps) enter_tap();
            }
        };
        auto finish = [&](const float (&va)[kAPer], const float (&vc)[kAPer], float4 (&out)[kAPer / 4]) {
            float v[kAPer];
#pragma unroll
            for (int i = 0; i < kAPer; ++i) {
                float value = va[i];
                if (A_ACT == kActLrelu) value = fmaxf(value, value * p.a.a_slope);   // slope in [0, 1]
                else if (A_ACT == kActLreluMask) value = vc[i] > 0.f ? value : value * p.a.a_slope;
                else if (A_ACT == kActTanhMask) value = value * (1.f - vc[i] * vc[i]);
                v[i] = to_tf32(value);
            }
#pragma unroll
            for (int q = 0; q < kAPer / 4; ++q)
                out[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        };

        float va0[kAPer], vc0[kAPer], va1[kAPer], vc1[kAPer];
        long long t_wait = 0, t_store = 0, mark = 0;
        issue_loads(va0, vc0);
        issue_loads(va1, vc1);
        for (int it = 0; it < stages_total; ++it) {
            const int s = it % kNStages;
            const int subs = min(kPair, k_steps - it * kPair);
            // finish this stage's operands and immediately reuse the registers for the next stage
            float4 a0[kAPer / 4], a1[kAPer / 4];
            finish(va0, vc0, a0);
            issue_loads(va0, vc0);
            if (subs > 1) {
                finish(va1, vc1, a1);
                issue_loads(va1, vc1);
            }
            if (timing && tid == 0) mark = clock64();
            if (it >= kNStages) mbar_wait(empty + s, ((it / kNStages) - 1) & 1);
            if (timing && tid == 0) { const long long now = clock64(); t_wait += now - mark; mark = now; }
            float4* dst0 = reinterpret_cast<float4*>(stage_a(s, 0));
#pragma unroll
            for (int q = 0; q < kAPer / 4; ++q)
                dst0[(aquarter * (kAPer / 4) + q) * kBM + arow] = a0[q];
            if (subs > 1) {
                float4* dst1 = reinterpret_cast<float4*>(stage_a(s, 1));
#pragma unroll
                for (int q = 0; q < kAPer / 4; ++q)
                    dst1[(aquarter * (kAPer / 4) + q) * kBM + arow] = a1[q];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full + s);
            if (timing && tid == 0) t_store += clock64() - mark;
        }
        if (timing && tid == 0) {
            p.debug[0] = clock64() - begin; p.debug[1] = t_wait; p.debug[3] = t_store;
        }
    }

    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .., columns of group w / 4 ----
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    if (timing && tid == 0) p.debug[7] = clock64() - begin;
    constexpr int kGroups = BN >= 64 ? 4 : BN / 16;
    constexpr int kGroupCols = BN / kGroups;
    if (warp < 4 * kGroups) {
        const int quad = warp & 3, group = warp >> 2;
        const int m = m0 + quad * 32 + lane;
        const bool ok = m < rows_total;
        int b, oh, ow;
        decode_row(ok ? m : 0, &b, &oh, &ow);
        const int rem = oh * p.o_w + ow;
#pragma unroll 1
        for (int c0 = group * kGroupCols; c0 < (group + 1) * kGroupCols; c0 += 16) {
            uint32_t raw[16];
            __syncwarp();
            tc_load16(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, raw);
            // one 64-bit index per chunk, then a step of o_positions per channel: the per-element
            // ((b o_ch + n) o_positions + rem) and the optional-input tests were most of this loop
            if (!ok) continue;
            const int n_first = n0 + c0;
            size_t idx = ((size_t)b * p.o_ch + n_first) * p.o_positions + rem;
            const float* bias = p.a.bias ? p.a.bias + n_first : nullptr;
            const float* bias2 = p.a.bias2 ? p.a.bias2 + (size_t)b * p.o_ch + n_first : nullptr;
            const bool lrelu = p.a.out_act == kOutLrelu;
            const int live = min(16, p.o_ch - n_first);      // channels of this chunk that exist
#pragma unroll
            for (int i = 0; i < 16; ++i, idx += p.o_positions) {
                if (i >= live) break;
                float v = k_steps > 0 ? __uint_as_float(raw[i]) : 0.f;   // no tap reaches this phase
                if (bias) v += __ldg(bias + i);
                if (bias2) v += __ldg(bias2 + i);
                if (lrelu) v = leaky(v, p.a.out_slope);
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
    if (timing && tid == 0) p.debug[4] = clock64() - begin;   // end of the epilogue
    if (warp == kProducers / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

__host__ __device__ inline int tile_columns(int rows) { return rows > 64 ? 128 : (rows > 32 ? 64 : 32); }
__host__ __device__ inline int conv_tc_pad(int channels) { return (channels + kKStep - 1) / kKStep * kKStep; }

// w (d0, d1, taps) -> the shared-memory images the kernel copies in bulk.  The GEMM rows are
// dim 0 (transposed = 0: forward of a Conv, data gradient of a ConvTranspose) or dim 1
// (transposed = 1: data gradient of a Conv), the reduction runs tap-major over the other
// dimension padded to 32.  Layout: [row tile][K step = tap x channel block][k / 4][row in tile][4],
// values rounded to tf32, padding rows / channels zero.
__global__ void pack_weight_taps_kernel(
    const float* __restrict__ w, float* __restrict__ out, int d0, int d1, int taps, int transposed,
    int c_pad, int bn, int row_tiles) {
    const size_t total = (size_t)row_tiles * bn * taps * c_pad;
    const int blocks = c_pad / kKStep;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 3);
        size_t rest = idx >> 2;
        const int local = (int)(rest % bn); rest /= bn;
        const int chunk = (int)(rest % (kKStep / 4)); rest /= (kKStep / 4);
        const int cb = (int)(rest % blocks); rest /= blocks;
        const int tap = (int)(rest % taps);
        const int tile = (int)(rest / taps);
        const int row = tile * bn + local;
        const int c = cb * kKStep + chunk * 4 + e;
        const int rows = transposed ? d1 : d0, reduce = transposed ? d0 : d1;
        float v = 0.f;
        if (row < rows && c < reduce) {
            const int a = transposed ? c : row, b = transposed ? row : c;
            v = to_tf32(w[((size_t)a * d1 + b) * taps + tap]);
        }
        out[idx] = v;
    }
}

template <int BN, bool TRANSPOSED, int A_ACT>
int launch_instance(const TcParams& p, cudaStream_t stream) {
    const size_t smem = (size_t)(BN > 128 ? 2 : kStages) * kPair * (kBM + BN) * kKStep * 4;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, TRANSPOSED, A_ACT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            "conv_gemm_tc smem attribute"));
        configured = true;
    }
    const int m_tiles = p.phases > 0 ? p.phase_tile_start[p.phases] : ceil_div(p.m_total, kBM);
    dim3 grid(m_tiles, ceil_div(p.o_ch, BN));
    PMN_REQUIRE(grid.y <= 65535, "conv_gemm_tc: too many output channels");
    LaunchScope scope(TRANSPOSED ? "conv_dgrad_tc_kernel" : "conv_fprop_tc_kernel", stream);
    conv_gemm_tc_kernel<BN, TRANSPOSED, A_ACT><<<grid, kThreads, smem, stream>>>(p);
    return launched("conv_gemm_tc_kernel");
}

template <int BN, bool TRANSPOSED>
int launch_activation(const TcParams& p, cudaStream_t stream) {
    switch (p.a.a_act) {
        case kActNone: return launch_instance<BN, TRANSPOSED, kActNone>(p, stream);
        case kActLrelu: return launch_instance<BN, TRANSPOSED, kActLrelu>(p, stream);
        case kActLreluMask: return launch_instance<BN, TRANSPOSED, kActLreluMask>(p, stream);
        default: return launch_instance<BN, TRANSPOSED, kActTanhMask>(p, stream);
    }
}

template <int BN>
int launch_variant(const TcParams& p, cudaStream_t stream) {
    return p.a.transposed ? launch_activation<BN, true>(p, stream)
                          : launch_activation<BN, false>(p, stream);
}


// ---------------------------------------------------------------------------
// Weight gradient on the tensor cores
// ---------------------------------------------------------------------------
//   gw[n, (c, tap)] += sum_{b, pos} act(dy)[b, n, pos] * act(x)[b, c, in(pos, tap)]
// GEMM view: M = 128 (c, tap) columns of the weight matrix per CTA (TMEM lanes; consecutive
// lanes are consecutive addresses of gw, so the atomic epilogue is coalesced), N = BN output
// channels, K = positions, 32 consecutive positions of one batch item per K step, the
// position range split over gridDim.z.  Both operands are gathered by the 512 producer
// threads into the [k / 4][row][4] layout (k = position).  The bias gradient rides along as
// one extra row of ones: row `ncols` of the M dimension accumulates sum_pos dy[n, pos].

struct TcWgradParams {
    ConvWgradArgs a;
    int taps, ncols, rows_total, o_positions, steps_per_item, steps_total, steps_per_split;
};

template <int ACT>
__device__ __forceinline__ float apply_act(float v, float companion, float slope) {
    if (ACT == kActLrelu) return fmaxf(v, v * slope);
    if (ACT == kActLreluMask) return companion > 0.f ? v : v * slope;
    if (ACT == kActTanhMask) return v * (1.f - companion * companion);
    return v;
}

template <int BN, int DY_ACT, int X_ACT>
__global__ void __launch_bounds__(kProducers + 32, 1) conv_wgrad_tc_kernel(TcWgradParams p) {
    constexpr uint32_t kABytes = kBM * kKStep * 4;
    constexpr uint32_t kBBytes = BN * kKStep * 4;
    constexpr uint32_t kStageBytes = kPair * (kABytes + kBBytes);
    constexpr int kAPer = 8;                                   // positions per thread per K step
    constexpr int kBThreads = BN * 4;                          // threads with a dy row segment
    constexpr bool kDyCompanion = DY_ACT == kActLreluMask || DY_ACT == kActTanhMask;
    static_assert(X_ACT == kActNone || X_ACT == kActLrelu, "x activation");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[kStages];
    __shared__ uint64_t empty[kStages];
    __shared__ uint64_t acc_done;
    __shared__ uint32_t tmem_slot;

    const pmn_conv_geometry& g = p.a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col0 = blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int t_begin = blockIdx.z * p.steps_per_split;
    const int t_end = min(t_begin + p.steps_per_split, p.steps_total);
    const int k_steps = t_end - t_begin;
    constexpr int kMmaWarp = kProducers / 32;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(full + i, kProducers / 32);
            mbar_init(empty + i, 1);
        }
        mbar_init(&acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int stages_total = (k_steps + kPair - 1) / kPair;
    auto stage_a = [&](int s, int sub) { return smem + s * kStageBytes + sub * kABytes; };
    auto stage_b = [&](int s, int sub) { return smem + s * kStageBytes + kPair * kABytes + sub * kBBytes; };

    if (warp == kMmaWarp) {
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(kBM, BN);
            for (int it = 0; it < stages_total; ++it) {
                const int s = it % kStages;
                mbar_wait(full + s, (it / kStages) & 1);
                tc_fence_after();
                const int subs = min(kPair, k_steps - it * kPair);
                for (int sub = 0; sub < subs; ++sub) {
                    const uint32_t a_addr = smem_u32(stage_a(s, sub));
                    const uint32_t b_addr = smem_u32(stage_b(s, sub));
#pragma unroll
                    for (int kk = 0; kk < kKStep / 8; ++kk) {
                        const uint64_t a_desc = smem_desc(a_addr + 2 * kk * kBM * 16, kBM * 16, 128);
                        const uint64_t b_desc = smem_desc(b_addr + 2 * kk * BN * 16, BN * 16, 128);
                        tc_mma_tf32(tmem_base, a_desc, b_desc, idesc,
                                    (it > 0 || sub > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                tc_commit(empty + s);
            }
            tc_commit(&acc_done);
        }
    } else {
        // ---- A role: weight-matrix column col0 + tid % 128, positions (tid / 128) * 8 .. + 8 ----
        const int arow = tid & (kBM - 1);
        const int aquarter = tid >> 7;
        const int col = col0 + arow;
        const bool is_ones = p.a.gbias != nullptr && col == p.ncols;   // the bias-gradient row
        const bool col_ok = col < p.ncols;
        int off_h = 0, off_w = 0;
        size_t channel_offset = 0;
        {
            const int cc = col_ok ? col : 0;
            const int c = cc / p.taps, tap = cc - c * p.taps;
            const int ti = tap / g.kw, tj = tap - ti * g.kw;
            off_h = ti * g.dh - g.ph;
            off_w = tj * g.dw - g.pw;
            channel_offset = (size_t)c * g.h_in * g.w_in;
        }
        const size_t x_item = (size_t)g.c_in * g.h_in * g.w_in;
        // ---- B role: output channel n0 + tid % BN, positions (tid / BN) * 8 .. + 8 ----
        const bool b_active = tid < kBThreads;
        const int brow = tid % BN;
        const int bsegment = tid / BN;          // 0..3 when active
        const bool n_ok = b_active && n0 + brow < g.c_out;
        const size_t dy_item = (size_t)g.c_out * p.o_positions;
        const size_t dy_row = (size_t)(n_ok ? n0 + brow : 0) * p.o_positions;

        // Linear case: the input index of a (column, position) pair is position * stride + shift
        // (every Conv1d, and the stride-1 (k, 1) convolutions over (H, W): 84 % of the FLOPs)
        const bool one_d = g.w_in == 1 && g.w_out == 1 && g.kw == 1;
        const bool linear = one_d ||
            (g.kw == 1 && g.sw == 1 && g.pw == 0 && g.w_in == g.w_out && g.sh == 1);
        const int lin_stride = one_d ? g.sh : 1;
        const int lin_shift = one_d ? off_h : off_h * g.w_in;
        const int x_plane = g.h_in * g.w_in;
        const bool dy_vector = (p.o_positions & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(p.a.dy) & 15) == 0 &&
            (!kDyCompanion || (reinterpret_cast<uintptr_t>(p.a.dy_companion) & 15) == 0);

        int loaded = 0;
        auto issue_loads = [&](float (&va)[kAPer], float (&vb)[kAPer], float (&vc)[kAPer]) {
            if (loaded >= k_steps) return;
            const int t = t_begin + loaded;
            ++loaded;
            const int b = t / p.steps_per_item;
            const int p0 = (t - b * p.steps_per_item) * kKStep;
            // x operand
            {
                const int first = p0 + aquarter * kAPer;
                const float* base = p.a.x + (size_t)b * x_item + channel_offset;
                if (is_ones) {
#pragma unroll
                    for (int e = 0; e < kAPer; ++e) va[e] = first + e < p.o_positions ? 1.f : 0.f;
                } else if (linear) {
                    const int start = first * lin_stride + lin_shift;
                    const int count = col_ok ? p.o_positions - first : 0;   // valid positions from `first`
                    const int end = start + (kAPer - 1) * lin_stride;
                    if (count >= kAPer && start >= 0 && end < x_plane) {
                        // interior (almost every step): no per-element predicate
                        const float* src = base + start;
                        if (lin_stride == 1) {
#pragma unroll
                            for (int e = 0; e < kAPer; ++e) va[e] = __ldg(src + e);
                        } else {
#pragma unroll
                            for (int e = 0; e < kAPer; ++e) va[e] = __ldg(src + e * lin_stride);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < kAPer; ++e) {
                            const int idx = start + e * lin_stride;
                            float value = 0.f;
                            if (e < count && (unsigned)idx < (unsigned)x_plane) value = __ldg(base + idx);
                            va[e] = value;
                        }
                    }
                } else {
                    int oh = first / g.w_out, ow = first - oh * g.w_out;
#pragma unroll
                    for (int e = 0; e < kAPer; ++e) {
                        const int hi = oh * g.sh + off_h, wi = ow * g.sw + off_w;
                        const bool ok = col_ok && first + e < p.o_positions && (unsigned)hi < (unsigned)g.h_in &&
                                        (unsigned)wi < (unsigned)g.w_in;
                        float value = 0.f;
                        if (ok) value = __ldg(base + hi * g.w_in + wi);
                        va[e] = value;
                        if (++ow == g.w_out) { ow = 0; ++oh; }
                    }
                }
            }
            // dy operand
            if (b_active) {
                const int first = p0 + bsegment * kAPer;
                const size_t base = (size_t)b * dy_item + dy_row + first;
                if (dy_vector && n_ok && first + kAPer <= p.o_positions) {
                    const float4 lo = __ldg(reinterpret_cast<const float4*>(p.a.dy + base));
                    const float4 hi = __ldg(reinterpret_cast<const float4*>(p.a.dy + base) + 1);
                    vb[0] = lo.x; vb[1] = lo.y; vb[2] = lo.z; vb[3] = lo.w;
                    vb[4] = hi.x; vb[5] = hi.y; vb[6] = hi.z; vb[7] = hi.w;
                    if (kDyCompanion) {
                        const float4 cl = __ldg(reinterpret_cast<const float4*>(p.a.dy_companion + base));
                        const float4 ch = __ldg(reinterpret_cast<const float4*>(p.a.dy_companion + base) + 1);
                        vc[0] = cl.x; vc[1] = cl.y; vc[2] = cl.z; vc[3] = cl.w;
                        vc[4] = ch.x; vc[5] = ch.y; vc[6] = ch.z; vc[7] = ch.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < kAPer; ++e) {
                        const bool ok = n_ok && first + e < p.o_positions;
                        vb[e] = ok ? __ldg(p.a.dy + base + e) : 0.f;
                        if (kDyCompanion) vc[e] = ok ? __ldg(p.a.dy_companion + base + e) : 0.f;
                    }
                }
            }
        };
        auto store = [&](int s, int sub, const float (&va)[kAPer], const float (&vb)[kAPer],
                         const float (&vc)[kAPer]) {
            float4* a_dst = reinterpret_cast<float4*>(stage_a(s, sub));
            float4* b_dst = reinterpret_cast<float4*>(stage_b(s, sub));
            float v[kAPer];
#pragma unroll
            for (int e = 0; e < kAPer; ++e) v[e] = to_tf32(apply_act<X_ACT>(va[e], 0.f, p.a.x_slope));
            a_dst[(aquarter * 2 + 0) * kBM + arow] = make_float4(v[0], v[1], v[2], v[3]);
            a_dst[(aquarter * 2 + 1) * kBM + arow] = make_float4(v[4], v[5], v[6], v[7]);
            if (b_active) {
#pragma unroll
                for (int e = 0; e < kAPer; ++e)
                    v[e] = to_tf32(apply_act<DY_ACT>(vb[e], kDyCompanion ? vc[e] : 0.f, p.a.dy_slope));
                b_dst[(bsegment * 2 + 0) * BN + brow] = make_float4(v[0], v[1], v[2], v[3]);
                b_dst[(bsegment * 2 + 1) * BN + brow] = make_float4(v[4], v[5], v[6], v[7]);
            }
        };

        float va0[kAPer], vb0[kAPer], vc0[kAPer], va1[kAPer], vb1[kAPer], vc1[kAPer];
        issue_loads(va0, vb0, vc0);
        issue_loads(va1, vb1, vc1);
        for (int it = 0; it < stages_total; ++it) {
            const int s = it % kStages;
            const int subs = min(kPair, k_steps - it * kPair);
            if (it >= kStages) mbar_wait(empty + s, ((it / kStages) - 1) & 1);
            store(s, 0, va0, vb0, vc0);
            issue_loads(va0, vb0, vc0);
            if (subs > 1) {
                store(s, 1, va1, vb1, vc1);
                issue_loads(va1, vb1, vc1);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full + s);
        }
    }

    // ---- epilogue: lanes = weight columns (consecutive addresses), TMEM columns = channels ----
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    constexpr int kGroups = BN >= 64 ? 4 : BN / 16;
    constexpr int kGroupCols = BN / kGroups;
    if (warp < 4 * kGroups && k_steps > 0) {
        const int quad = warp & 3, group = warp >> 2;
        const int col = col0 + quad * 32 + lane;
#pragma unroll 1
        for (int c0 = group * kGroupCols; c0 < (group + 1) * kGroupCols; c0 += 16) {
            uint32_t raw[16];
            __syncwarp();
            tc_load16(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, raw);
            const int n_first = n0 + c0;
            const int live = min(16, g.c_out - n_first);
            if (col < p.ncols) {
                float* target = p.a.gw + (size_t)n_first * p.ncols + col;
#pragma unroll
                for (int i = 0; i < 16; ++i, target += p.ncols) {
                    const float v = __uint_as_float(raw[i]);
                    if (i < live && v != 0.f) atomicAdd(target, v);
                }
            } else if (col == p.ncols && p.a.gbias) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v = __uint_as_float(raw[i]);
                    if (i < live && v != 0.f) atomicAdd(p.a.gbias + n_first + i, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

template <int BN, int DY_ACT, int X_ACT>
int launch_wgrad_instance(const TcWgradParams& p, dim3 grid, cudaStream_t stream) {
    const size_t smem = (size_t)kStages * kPair * (kBM + BN) * kKStep * 4;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, DY_ACT, X_ACT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            "conv_wgrad_tc smem attribute"));
        configured = true;
    }
    LaunchScope scope("conv_wgrad_tc_kernel", stream);
    conv_wgrad_tc_kernel<BN, DY_ACT, X_ACT><<<grid, kProducers + 32, smem, stream>>>(p);
    return launched("conv_wgrad_tc_kernel");
}

template <int BN>
int launch_wgrad_activation(const TcWgradParams& p, dim3 grid, cudaStream_t stream) {
    const int dy = p.a.dy_act, x = p.a.x_act;
    if (dy == kActNone && x == kActNone) return launch_wgrad_instance<BN, kActNone, kActNone>(p, grid, stream);
    if (dy == kActNone && x == kActLrelu) return launch_wgrad_instance<BN, kActNone, kActLrelu>(p, grid, stream);
    if (dy == kActLreluMask && x == kActNone)
        return launch_wgrad_instance<BN, kActLreluMask, kActNone>(p, grid, stream);
    if (dy == kActLrelu && x == kActNone) return launch_wgrad_instance<BN, kActLrelu, kActNone>(p, grid, stream);
    return fail(PMN_ERR_ARGUMENT, "conv_wgrad_tc: this activation pair is not built (use pmn_conv_wgrad)");
}


// One launch for every convolution of a module: fold the weight norm (block per output
// row) ...
__global__ void __launch_bounds__(256) fold_weights_kernel(const pmn_weight_desc* table) {
    const pmn_weight_desc d = table[blockIdx.y];
    if ((int)blockIdx.x >= d.dim0 || d.g == nullptr) return;
    __shared__ float partial[32];
    const int inner = d.dim1 / max(d.groups, 1) * d.taps;
    const float* row = d.v + (size_t)blockIdx.x * inner;
    float sum = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) sum = fmaf(row[i], row[i], sum);
    for (int offset = 16; offset > 0; offset >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, offset);
    if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? partial[threadIdx.x] : 0.f;
        for (int offset = 16; offset > 0; offset >>= 1) t += __shfl_xor_sync(0xffffffffu, t, offset);
        if (threadIdx.x == 0) partial[0] = t;
    }
    __syncthreads();
    const float scale = d.g[blockIdx.x] / sqrtf(partial[0]);
    float* dst = d.w + (size_t)blockIdx.x * inner;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) dst[i] = row[i] * scale;
}

// Element (a, b, tap) of the dense (dim0, dim1, taps) weight; a grouped convolution
// (discriminator.py:218-224) stores (dim0, dim1 / groups, taps) and is zero off its diagonal blocks
__device__ __forceinline__ float dense_weight(const pmn_weight_desc& d, const float* w, int a, int b, int tap) {
    if (d.groups <= 1) return w[((size_t)a * d.dim1 + b) * d.taps + tap];
    const int per_in = d.dim1 / d.groups, per_out = d.dim0 / d.groups;
    const int group = a / per_out;
    if (b / per_in != group) return 0.f;
    return w[((size_t)a * per_in + (b - group * per_in)) * d.taps + tap];
}

// ... then write both tensor-core packings (and, for the FMA path, the plain transpose and the
// dense form of a grouped weight)
__global__ void __launch_bounds__(256) pack_weights_kernel(const pmn_weight_desc* table) {
    const pmn_weight_desc d = table[blockIdx.y];
    const float* w = d.g ? d.w : d.v;
    const int blocks[2] = {conv_tc_pad(d.dim1) / kKStep, conv_tc_pad(d.dim0) / kKStep};
    const int bn[2] = {tile_columns(d.dim0), tile_columns(d.dim1)};
    const int rows[2] = {d.dim0, d.dim1};
    float* out[2] = {d.packed, d.packed_t};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (!out[t]) continue;
        const int tiles = (rows[t] + bn[t] - 1) / bn[t];
        // 32-bit index arithmetic (a packing has < 2^31 elements: train/layers.py checks): the
        // five 64-bit divisions per element made this pass instruction-bound at 13 x its memory time
        const uint32_t total = (uint32_t)tiles * bn[t] * d.taps * blocks[t] * kKStep;
        for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
            const int e = (int)(idx & 3);
            uint32_t rest = idx >> 2;
            const int local = (int)(rest % (uint32_t)bn[t]); rest /= (uint32_t)bn[t];
            const int chunk = (int)(rest % (kKStep / 4)); rest /= (kKStep / 4);
            const int cb = (int)(rest % (uint32_t)blocks[t]); rest /= (uint32_t)blocks[t];
            const int tap = (int)(rest % (uint32_t)d.taps);
            const int tile = (int)(rest / (uint32_t)d.taps);
            const int row = tile * bn[t] + local;
            const int c = cb * kKStep + chunk * 4 + e;
            const int reduce = t ? d.dim0 : d.dim1;
            float v = 0.f;
            if (row < rows[t] && c < reduce) {
                const int a = t ? c : row, b = t ? row : c;
                v = to_tf32(dense_weight(d, w, a, b, tap));
            }
            out[t][idx] = v;
        }
    }
    if (d.wt || d.dense) {
        const size_t total = (size_t)d.dim0 * d.dim1 * d.taps;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (size_t)gridDim.x * blockDim.x) {
            const int tap = (int)(idx % d.taps);
            const size_t rest = idx / d.taps;
            if (d.wt) {
                const int a = (int)(rest % d.dim0), b = (int)(rest / d.dim0);
                d.wt[idx] = dense_weight(d, w, a, b, tap);
            }
            if (d.dense) {
                const int b = (int)(rest % d.dim1), a = (int)(rest / d.dim1);
                d.dense[idx] = dense_weight(d, w, a, b, tap);
            }
        }
    }
}

// gw (dim0, dim1 / groups, taps) = the diagonal blocks of the dense gradient (dim0, dim1, taps)
__global__ void extract_grouped_kernel(
    const float* __restrict__ dense, float* __restrict__ gw, int dim0, int dim1, int taps, int groups) {
    const int per_in = dim1 / groups, per_out = dim0 / groups;
    const size_t total = (size_t)dim0 * per_in * taps;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % taps);
        const size_t rest = idx / taps;
        const int local = (int)(rest % per_in), a = (int)(rest / per_in);
        const int b = (a / per_out) * per_in + local;
        gw[idx] = dense[((size_t)a * dim1 + b) * taps + tap];
    }
}

}  // namespace

void set_train_tc_debug(long long* counters) { g_debug_counters = counters; }

int conv_tc_channel_pad(int channels) { return (channels + kKStep - 1) / kKStep * kKStep; }

size_t conv_tc_packed_floats(int rows, int reduce, int taps) {
    const int bn = tile_columns(rows);
    return (size_t)ceil_div(rows, bn) * bn * taps * conv_tc_channel_pad(reduce);
}

int launch_pack_weight_taps(
    const float* w, float* out, int d0, int d1, int taps, int transposed, cudaStream_t stream) {
    PMN_REQUIRE(w && out && d0 > 0 && d1 > 0 && taps > 0, "pack_weight_taps: bad argument");
    const int rows = transposed ? d1 : d0, reduce = transposed ? d0 : d1;
    const int c_pad = conv_tc_channel_pad(reduce);
    const int bn = tile_columns(rows);
    const int row_tiles = ceil_div(rows, bn);
    const size_t total = conv_tc_packed_floats(rows, reduce, taps);
    const int blocks = (int)min((size_t)2048, (total + 255) / 256);
    LaunchScope scope("pack_weight_taps_kernel", stream);
    pack_weight_taps_kernel<<<blocks, 256, 0, stream>>>(
        w, out, d0, d1, taps, transposed, c_pad, bn, row_tiles);
    return launched("pack_weight_taps_kernel");
}

int launch_extract_grouped(
    const float* dense, float* gw, int dim0, int dim1, int taps, int groups, cudaStream_t stream) {
    PMN_REQUIRE(dense && gw && dim0 > 0 && dim1 > 0 && taps > 0 && groups > 0 && dim0 % groups == 0 &&
                dim1 % groups == 0, "extract_grouped: bad argument");
    const size_t total = (size_t)dim0 * (dim1 / groups) * taps;
    LaunchScope scope("extract_grouped_kernel", stream);
    extract_grouped_kernel<<<(int)min((size_t)1024, (total + 255) / 256), 256, 0, stream>>>(
        dense, gw, dim0, dim1, taps, groups);
    return launched("extract_grouped_kernel");
}

int launch_prepare_weights(const pmn_weight_desc* table, int layers, int max_dim0, cudaStream_t stream) {
    PMN_REQUIRE(table && layers > 0 && layers <= 65535 && max_dim0 > 0, "prepare_weights: bad argument");
    {
        LaunchScope scope("fold_weights_kernel", stream);
        fold_weights_kernel<<<dim3(max_dim0, layers), 256, 0, stream>>>(table);
        PMN_TRY(launched("fold_weights_kernel"));
    }
    LaunchScope scope("pack_weights_kernel", stream);
    pack_weights_kernel<<<dim3(256, layers), 256, 0, stream>>>(table);
    return launched("pack_weights_kernel");
}

int launch_conv_wgrad_tc(const ConvWgradArgs& args, cudaStream_t stream) {
    const pmn_conv_geometry& g = args.g;
    PMN_REQUIRE(args.dy && args.x && args.gw, "conv_wgrad_tc: null pointer");
    PMN_REQUIRE(g.batch > 0 && g.c_in > 0 && g.c_out > 0 && g.h_in > 0 && g.w_in > 0 &&
                g.h_out > 0 && g.w_out > 0 && g.kh > 0 && g.kw > 0 && g.sh > 0 && g.sw > 0 &&
                g.dh > 0 && g.dw > 0 && g.ph >= 0 && g.pw >= 0, "conv_wgrad_tc: bad geometry");
    PMN_REQUIRE(args.dy_act == kActNone || args.dy_act == kActLrelu || args.dy_companion,
                "conv_wgrad_tc: this dy activation needs a companion tensor");
    TcWgradParams p;
    p.a = args;
    p.taps = g.kh * g.kw;
    p.ncols = g.c_in * p.taps;
    p.rows_total = p.ncols + (args.gbias ? 1 : 0);
    p.o_positions = g.h_out * g.w_out;
    p.steps_per_item = ceil_div(p.o_positions, kKStep);
    p.steps_total = g.batch * p.steps_per_item;
    const int bn = tile_columns(g.c_out);
    const int tiles = ceil_div(p.rows_total, kBM) * ceil_div(g.c_out, bn);
    // Split the position range over gridDim.z so that the CTAs fill whole waves of the SMs: the cost
    // of a launch is (waves) x (K steps of one CTA + its fixed prologue / epilogue, worth about 6
    // steps), at least 8 steps per CTA.  The heaviest launches of a training step (MPD 1024 -> 1024:
    // 328 tiles) ran unsplit in 3 waves with the last 22 % full, and "two waves' worth" of CTAs
    // (296 / tiles splits) often landed just past two waves.
    static int sms = 0;
    if (!sms) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (sms <= 0) sms = 148;
    }
    int splits = 1;
    {
        const int most = max(1, min(64, p.steps_total / 8));
        long long best = -1;
        for (int candidate = 1; candidate <= most; ++candidate) {
            const int per = ceil_div(p.steps_total, candidate);
            const int real = ceil_div(p.steps_total, per);
            const long long cost = (long long)ceil_div(tiles * real, sms) * (per + 6);
            if (best < 0 || cost < best) { best = cost; splits = candidate; }
        }
    }
    p.steps_per_split = ceil_div(p.steps_total, splits);
    splits = ceil_div(p.steps_total, p.steps_per_split);
    dim3 grid(ceil_div(p.rows_total, kBM), ceil_div(g.c_out, bn), splits);
    PMN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv_wgrad_tc: grid too large");
    switch (bn) {
        case 128: return launch_wgrad_activation<128>(p, grid, stream);
        case 64: return launch_wgrad_activation<64>(p, grid, stream);
        default: return launch_wgrad_activation<32>(p, grid, stream);
    }
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
    PMN_REQUIRE(args.out_act != kOutTanh, "conv_gemm_tc: tanh epilogue is not built (use pmn_conv_gemm)");
    PMN_REQUIRE(args.a_act != kActLrelu || (args.a_slope >= 0.f && args.a_slope <= 1.f),
                "conv_gemm_tc: LeakyReLU slope must be in [0, 1]");
    TcParams p;
    p.debug = g_debug_counters;
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
    const bool strided = g.channel_stride || g.position_stride || g.batch_stride;
    PMN_REQUIRE(!strided || (!args.transposed && g.channel_stride > 0 && g.position_stride > 0 &&
                             g.batch_stride > 0), "conv_gemm_tc: bad tensor strides");
    p.channel_stride = strided ? g.channel_stride : p.a_h * p.a_w;
    p.position_stride = strided ? g.position_stride : 1;
    p.batch_stride = strided ? (size_t)g.batch_stride : (size_t)p.a_ch * p.a_h * p.a_w;
    p.phases = 0;
    if (args.transposed && (g.sh > 1 || g.sw > 1) && g.dh == 1 && g.dw == 1 &&
        g.sh * g.sw <= kMaxPhases && p.taps <= 64) {
        int tiles = 0;
        for (int rh = 0; rh < g.sh; ++rh)
            for (int rw = 0; rw < g.sw; ++rw) {
                // rows h with (h + ph) mod sh = rh start at h0 and step by sh
                const int h0 = ((rh - g.ph) % g.sh + g.sh) % g.sh, w0 = ((rw - g.pw) % g.sw + g.sw) % g.sw;
                const int qh = h0 < p.o_h ? (p.o_h - h0 + g.sh - 1) / g.sh : 0;
                const int qw = w0 < p.o_w ? (p.o_w - w0 + g.sw - 1) / g.sw : 0;
                if (qh == 0 || qw == 0) continue;
                const int k = p.phases++;
                p.phase_tile_start[k] = tiles;
                p.phase_h0[k] = h0; p.phase_w0[k] = w0; p.phase_qh[k] = qh; p.phase_qw[k] = qw;
                tiles += ceil_div(g.batch * qh * qw, kBM);
            }
        p.phase_tile_start[p.phases] = tiles;
    }
    const int m_tiles_all = p.phases > 0 ? p.phase_tile_start[p.phases] : ceil_div(p.m_total, kBM);
    // two 128-column weight tiles per CTA halve the gather work per FLOP; worth it while the
    // grid still covers most of the 148 SMs
    if (p.o_ch % 256 == 0 && m_tiles_all * (p.o_ch / 256) >= 100)
        return launch_variant<256>(p, stream);
    switch (tile_columns(p.o_ch)) {
        case 128: return launch_variant<128>(p, stream);
        case 64: return launch_variant<64>(p, stream);
        default: return launch_variant<32>(p, stream);
    }
}

}  // namespace pmn

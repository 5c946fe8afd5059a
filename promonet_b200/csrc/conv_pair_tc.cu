// One residual pair of HiFi-GAN's Block.forward in ONE kernel (tcgen05 + TMEM):
//     y = x + c2(lrelu(c1(lrelu(x))))         promonet/model/hifigan.py:198-210
// c1 = dilated Conv1d(C, C, k, dilation d), c2 = Conv1d(C, C, k), both "same".
//
// conv1d_tc.cu runs the two convolutions as two launches that hand the activation
// over through HBM as bf16 hi/lo operand planes: 24 B per element and pair (planes in
// and out of c1; planes in, fp32 residual in, fp32 and planes out of c2).  Here the
// fp32 residual stream is read once and written once (8 B per element): nothing
// else leaves the SM.
//
//   converter warps   x (fp32, HBM/L2) -> lrelu -> bf16 hi/lo operand image of the
//                     time window in shared memory (what planes_from_f32 wrote to HBM)
//   MMA thread        c1 over that window into TMEM accumulator 1
//   mid epilogue      accumulator 1 + bias -> lrelu -> hi/lo -> c2's operand image in
//                     shared memory (rows outside [0, T) zeroed: c2's zero padding)
//   MMA thread        c2 over the mid image into TMEM accumulator 2
//   final epilogue    accumulator 2 + bias + x (the residual, an L2 hit) -> fp32 out,
//                     or the MRF mean accumulated in place (hifigan.py:141-145)
//
// A tile is S x 128 rows of c1 output; c2 needs (k - 1) / 2 rows of it on either
// side, so a tile yields S x 128 - (k - 1) output rows and consecutive tiles
// recompute that halo (8 % of c1 at k = 11, S = 1).  Both accumulators are double
// buffered in TMEM (4 x S x columns = all 512 columns) and the MMA thread issues
// c1 of tile i + 1 before c2 of tile i, so the tensor pipe works on c1 while the
// mid epilogue of the tile before it runs.  Weights stream from L2 as in
// conv1d_tc.cu (same slabs: pack_tc_weight_kernel), one bulk copy per (tap, K block).
//
// The arithmetic is that of the two-launch path operation for operation (same
// operand rounding, same products, same epilogue order), so the results are
// bit-identical to it -- which is how the tests pin this kernel.
#include "conv1d_tc.cuh"
#include "tc_ptx.cuh"

namespace pmn {

namespace {

using namespace tc;

constexpr int kMaxSpan1 = 50;       // (k - 1) d of c1: (11 - 1) * 5
constexpr int kMaxSpan2 = 10;       // k - 1 of c2
constexpr int kAlignSlack = 6;      // the staged window starts on a multiple of 4 samples (<= 3
                                    // rows early) and is staged in quads of rows (<= 3 rows late)

long long* g_pair_debug = nullptr;  // optional per-CTA cycle counters (pmn_debug_tc_counters)
int g_pair_variant = -1;            // experiment knob: -1 = the default variant of each C
constexpr int kDefaultVariant = 0;   // measured fastest (profiles/r2_pair_breakdown.txt)

__host__ __device__ constexpr uint32_t pair_instr_desc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// C channels, S x 128 rows of c1 per tile, KB input channels per staged K block, NW
// weight-slab stages; CONCAT as in conv1d_tc.cu (B = [W_hi; W_lo], two MMAs per K chunk);
// NC converter warps; MB buffers of the mid image; NF final-epilogue warps (4 or 8)
template <int C, int S, int KB, int NW, bool CONCAT, int NC, int MB, int NF>
struct PairConfig {
    static constexpr int kConverters = NC * 32;
    static constexpr int kThreads = 32 * (2 + NC + 4 + NF);  // producer, MMA, converters, mid, final
    static constexpr int kMidWarp = 2 + NC;                  // first mid-epilogue warp
    static constexpr int kFinalWarp = 2 + NC + 4;            // first final-epilogue warp
    static constexpr int kMidRows = S * 128;
    static constexpr int kXRows = kMidRows + kMaxSpan1 + kAlignSlack;   // rows of a staged window
    static constexpr int kMRows = kMidRows + kMaxSpan2;   // rows of the mid image (the last
                                                          // k - 1 feed discarded outputs only)
    static constexpr int kGroups = KB / 8;
    static constexpr int kBlocks = C / KB;
    static constexpr int kXStages = 2;
    static constexpr int kXSlab = 2 * kGroups * kXRows * 16;   // bytes, both planes
    static constexpr int kMid = 2 * (C / 8) * kMRows * 16;    // bytes of one mid image
    static constexpr int kWSlab = KB * C * 4;                  // hi + lo
    static constexpr int kCols = CONCAT ? 2 * C : C;           // TMEM columns per 128 rows
    static constexpr int kAcc = S * kCols;                     // columns per accumulator stage
    static constexpr int kBarriers = 2 * kXStages + 2 * NW + 2 * 2 + 2 * MB + 2 * 2;
    static constexpr int kSmem =
        128 + kXStages * kXSlab + MB * kMid + NW * kWSlab + kBarriers * 8 + 16 + 2 * C * 4;
    static_assert(4 * kAcc == 512, "two double-buffered accumulators fill TMEM");
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
    static_assert(C % KB == 0 && KB % 16 == 0 && kCols <= 256, "shape");
};

struct TcPairArgs {
    const float* x = nullptr;                 // (B, C, T) fp32: input and residual
    const __nv_bfloat16* w1 = nullptr;        // slabs of c1 (pack_tc_weight_kernel)
    const __nv_bfloat16* w2 = nullptr;
    const float* bias1 = nullptr;
    const float* bias2 = nullptr;
    float* out = nullptr;                     // (B, C, T) fp32 or null
    float* accum = nullptr;                   // (B, C, T) fp32 or null
    int accum_mode = 0;                       // 0 unused, 1 store, 2 add
    float accum_scale = 1.f;
    int batch = 0, t_len = 0, k = 1, dilation = 1;
    float slope = 0.1f;
    // optional (gridDim.x, 10 roles, 4) cycle counters: [0] total, [1..3] barrier waits
    long long* debug = nullptr;
};

template <int C, int S, int KB, int NW, bool CONCAT, int NC, int MB, int NF>
__global__ void __launch_bounds__(PairConfig<C, S, KB, NW, CONCAT, NC, MB, NF>::kThreads, 1)
conv_pair_tc_kernel(TcPairArgs a, int tiles_per_item, int num_tiles) {
    using Cfg = PairConfig<C, S, KB, NW, CONCAT, NC, MB, NF>;
    constexpr int kConverters = Cfg::kConverters;
    constexpr int kPairThreads = Cfg::kThreads;
    extern __shared__ uint8_t smem_raw[];
    // 128-byte alignment by pointer arithmetic on the __shared__ array: through an integer cast the
    // compiler loses the address space and emits generic LD / ST for every access to the buffers
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* x_slabs = smem;
    uint8_t* mid = x_slabs + Cfg::kXStages * Cfg::kXSlab;
    uint8_t* w_slabs = mid + MB * Cfg::kMid;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_slabs + NW * Cfg::kWSlab);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + Cfg::kXStages;
    uint64_t* w_full = x_empty + Cfg::kXStages;
    uint64_t* w_empty = w_full + NW;
    uint64_t* acc1_full = w_empty + NW;
    uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* mid_full = acc1_empty + 2;
    uint64_t* mid_empty = mid_full + MB;
    uint64_t* acc2_full = mid_empty + MB;
    uint64_t* acc2_empty = acc2_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);
    float* bias_smem = reinterpret_cast<float*>(tmem_slot + 4);   // bias1 | bias2

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int span1 = (a.k - 1) * a.dilation;
    const int span2 = a.k - 1;
    const int h1 = span1 / 2, h2 = span2 / 2;
    const int tile_rows = Cfg::kMidRows - span2;     // output rows a tile yields
    const int x_rows = Cfg::kMidRows + span1;        // staged rows c1 reads
    // 16-byte loads of four consecutive samples need aligned rows; otherwise sample by sample
    const bool vector = (a.t_len & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    // window start of a tile: time of staged row 0, and the rows c1 skips to reach its own first
    auto window = [&](int tile, int& first_time, int& skip) {
        const int t_start = (tile % tiles_per_item) * tile_rows - h2 - h1;
        skip = vector ? (t_start & 3) : 0;           // two's complement: also right for t_start < 0
        first_time = t_start - skip;
    };
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::kXStages; ++i) { mbar_init(x_full + i, kConverters / 32); mbar_init(x_empty + i, 1); }
        for (int i = 0; i < NW; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc1_full + i, 1); mbar_init(acc1_empty + i, 4);
            mbar_init(acc2_full + i, 1); mbar_init(acc2_empty + i, NF);
        }
        for (int i = 0; i < MB; ++i) { mbar_init(mid_full + i, 4); mbar_init(mid_empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * C; i += kPairThreads) {
        const float* source = i < C ? a.bias1 : a.bias2;
        bias_smem[i] = source ? source[i < C ? i : i - C] : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== weight producer: slabs in the order the MMA thread consumes them =====
        if (lane == 0) {
            uint32_t wcount = 0;
            long long begin = a.debug ? clock64() : 0, wait_w = 0, mark = 0;
            auto stream_weights = [&](const __nv_bfloat16* slabs) {
                for (int kb = 0; kb < Cfg::kBlocks; ++kb) {
                    for (int tap = 0; tap < a.k; ++tap) {
                        const uint32_t ws = wcount % NW, wphase = (wcount / NW) & 1;
                        ++wcount;
                        if (a.debug) mark = clock64();
                        mbar_wait(w_empty + ws, wphase ^ 1);
                        if (a.debug) wait_w += clock64() - mark;
                        mbar_expect_tx(w_full + ws, Cfg::kWSlab);
                        bulk_copy(w_slabs + ws * Cfg::kWSlab,
                                  reinterpret_cast<const uint8_t*>(slabs) +
                                      (size_t)(tap * Cfg::kBlocks + kb) * Cfg::kWSlab,
                                  Cfg::kWSlab, w_full + ws);
                    }
                }
            };
            for (int i = 0; i <= my_tiles; ++i) {
                if (i < my_tiles) stream_weights(a.w1);
                if (i >= 1) stream_weights(a.w2);
            }
            if (a.debug) {
                long long* d = a.debug + ((size_t)blockIdx.x * 10 + 0) * 4;
                d[0] = clock64() - begin; d[1] = wait_w;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: c1(0), then c1(i + 1) before c2(i) =====
        if (lane == 0) {
            constexpr uint32_t idesc = pair_instr_desc(128, C);
            constexpr uint32_t idesc_wide = pair_instr_desc(128, Cfg::kCols);
            uint32_t xcount = 0, wcount = 0;
            long long begin = a.debug ? clock64() : 0, mark = 0;
            long long wait_x = 0, wait_w = 0, wait_acc1 = 0, wait_mid = 0, wait_acc2 = 0;
            // one convolution over an operand image: rows0 = image rows per 8-channel
            // group, plane = bytes between its hi and lo halves, step = rows per tap
            auto taps = [&](uint32_t image, uint32_t rows0, uint32_t plane, uint32_t step,
                            uint32_t d_base, bool first_block) {
                for (int tap = 0; tap < a.k; ++tap) {
                    const uint32_t ws = wcount % NW, wphase = (wcount / NW) & 1;
                    ++wcount;
                    if (a.debug) mark = clock64();
                    mbar_wait(w_full + ws, wphase);
                    if (a.debug) wait_w += clock64() - mark;
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(w_slabs + ws * Cfg::kWSlab);
                    const bool first = first_block && tap == 0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const uint32_t row = s * 128 + tap * step;
                        const uint32_t d = d_base + s * Cfg::kCols;
#pragma unroll
                        for (int kk = 0; kk < KB / 16; ++kk) {
                            const uint32_t xa = image + (2 * kk * rows0 + row) * 16;
                            const uint64_t a_hi = smem_desc(xa, rows0 * 16, 128);
                            const uint64_t a_lo = smem_desc(xa + plane, rows0 * 16, 128);
                            if constexpr (CONCAT) {
                                const uint32_t wa = w_addr + 2 * kk * (2 * C) * 16;
                                const uint64_t b_both = smem_desc(wa, 2 * C * 16, 128);
                                tc_mma(d, a_hi, b_both, idesc_wide, !(first && kk == 0));
                                tc_mma(d, a_lo, b_both, idesc, 1);
                            } else {
                                constexpr uint32_t w_plane = Cfg::kGroups * C * 16;
                                const uint32_t wa = w_addr + 2 * kk * C * 16;
                                const uint64_t b_hi = smem_desc(wa, C * 16, 128);
                                const uint64_t b_lo = smem_desc(wa + w_plane, C * 16, 128);
                                tc_mma(d, a_hi, b_hi, idesc, !(first && kk == 0));
                                tc_mma(d, a_lo, b_hi, idesc, 1);
                                tc_mma(d, a_hi, b_lo, idesc, 1);
                            }
                        }
                    }
                    tc_commit(w_empty + ws);
                }
            };
            for (int i = 0; i <= my_tiles; ++i) {
                if (i < my_tiles) {
                    const uint32_t as = i & 1, aphase = (i >> 1) & 1;
                    int first_time, skip;
                    window(blockIdx.x + i * gridDim.x, first_time, skip);
                    if (a.debug) mark = clock64();
                    mbar_wait(acc1_empty + as, aphase ^ 1);
                    if (a.debug) wait_acc1 += clock64() - mark;
                    tc_fence_after();
                    for (int kb = 0; kb < Cfg::kBlocks; ++kb) {
                        const uint32_t xs = xcount % Cfg::kXStages, xphase = (xcount / Cfg::kXStages) & 1;
                        ++xcount;
                        if (a.debug) mark = clock64();
                        mbar_wait(x_full + xs, xphase);
                        if (a.debug) wait_x += clock64() - mark;
                        tc_fence_after();
                        taps(smem_u32(x_slabs + xs * Cfg::kXSlab) + skip * 16, Cfg::kXRows,
                             Cfg::kGroups * Cfg::kXRows * 16, a.dilation,
                             tmem_base + as * Cfg::kAcc, kb == 0);
                        tc_commit(x_empty + xs);
                    }
                    tc_commit(acc1_full + as);
                }
                if (i >= 1) {
                    const int j = i - 1;
                    const uint32_t as = j & 1, aphase = (j >> 1) & 1;
                    const uint32_t ms = j % MB, mphase = (j / MB) & 1;
                    if (a.debug) mark = clock64();
                    mbar_wait(mid_full + ms, mphase);
                    if (a.debug) { wait_mid += clock64() - mark; mark = clock64(); }
                    mbar_wait(acc2_empty + as, aphase ^ 1);
                    if (a.debug) wait_acc2 += clock64() - mark;
                    tc_fence_after();
                    for (int kb = 0; kb < Cfg::kBlocks; ++kb)
                        taps(smem_u32(mid + ms * Cfg::kMid) + kb * Cfg::kGroups * Cfg::kMRows * 16,
                             Cfg::kMRows, (C / 8) * Cfg::kMRows * 16, 1,
                             tmem_base + 2 * Cfg::kAcc + as * Cfg::kAcc, kb == 0);
                    tc_commit(mid_empty + ms);
                    tc_commit(acc2_full + as);
                }
            }
            if (a.debug) {
                long long* d = a.debug + ((size_t)blockIdx.x * 10 + 1) * 4;
                d[0] = clock64() - begin; d[1] = wait_x; d[2] = wait_w; d[3] = wait_acc1;
                d += 4 * 4;   // role 5: the c2 side of the MMA thread
                d[0] = wait_mid; d[1] = wait_acc2;
            }
        }
    } else if (warp < Cfg::kMidWarp) {
        // ===== converters: fp32 x -> lrelu -> hi/lo operand image of the window =====
        const int ctid = threadIdx.x - 64;
        uint32_t xcount = 0;
        long long begin = a.debug ? clock64() : 0, wait_x = 0, mark = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = tile / tiles_per_item;
            int first_time, skip;
            window(tile, first_time, skip);
            for (int kb = 0; kb < Cfg::kBlocks; ++kb) {
                const uint32_t xs = xcount % Cfg::kXStages, xphase = (xcount / Cfg::kXStages) & 1;
                ++xcount;
                if (a.debug && ctid == 0) mark = clock64();
                mbar_wait(x_empty + xs, xphase ^ 1);
                if (a.debug && ctid == 0) wait_x += clock64() - mark;
                uint8_t* dst = x_slabs + xs * Cfg::kXSlab;
                const float* src = a.x + ((size_t)b * C + kb * KB) * a.t_len;
                if (vector) {
                    // a task = 4 consecutive rows x 8 channels: eight 16-byte loads, eight stores
                    constexpr int kU = NC >= 8 ? 1 : 2;
                    const int quads = (x_rows + skip + 3) >> 2;
                    const int tasks = quads * Cfg::kGroups;
                    for (int base = ctid; base < tasks; base += kConverters * kU) {
                        float4 v[kU][8];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            const int idx = base + u * kConverters;
                            const int g = idx / quads, q = idx - g * quads;
                            const int t = first_time + 4 * q;
                            const bool live = idx < tasks && t >= 0 && t < a.t_len;
                            // one address, then pointer increments (conv1d_tc.cu: per-element 64-bit
                            // address arithmetic was half of that epilogue's instructions)
                            const float* p = src + (ptrdiff_t)(g * 8) * a.t_len + t;
                            if (live) {
#pragma unroll
                                for (int e = 0; e < 8; ++e, p += a.t_len)
                                    v[u][e] = __ldg(reinterpret_cast<const float4*>(p));
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[u][e] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            const int idx = base + u * kConverters;
                            if (idx < tasks) {
                                const int g = idx / quads, q = idx - g * quads;
                                uint8_t* row = dst + ((size_t)g * Cfg::kXRows + 4 * q) * 16;
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    uint32_t hi[4], lo[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float4 &c0 = v[u][2 * e], &c1 = v[u][2 * e + 1];
                                        const float y0 = r == 0 ? c0.x : r == 1 ? c0.y : r == 2 ? c0.z : c0.w;
                                        const float y1 = r == 0 ? c1.x : r == 1 ? c1.y : r == 2 ? c1.z : c1.w;
                                        split_pair(leaky(y0, a.slope), leaky(y1, a.slope), hi[e], lo[e]);
                                    }
                                    *reinterpret_cast<uint4*>(row + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                    *reinterpret_cast<uint4*>(row + r * 16 + Cfg::kGroups * Cfg::kXRows * 16) =
                                        make_uint4(lo[0], lo[1], lo[2], lo[3]);
                                }
                            }
                        }
                    }
                } else {
                    constexpr int kU = 4;                // (row, group) tasks in flight per thread
                    const int tasks = x_rows * Cfg::kGroups;
                    for (int base = ctid; base < tasks; base += kConverters * kU) {
                        float v[kU][8];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            const int idx = base + u * kConverters;
                            const int g = idx / x_rows, q = idx - g * x_rows;
                            const int t = first_time + q;
                            const bool live = idx < tasks && t >= 0 && t < a.t_len;
                            const float* p = src + (ptrdiff_t)(g * 8) * a.t_len + t;
                            if (live) {
#pragma unroll
                                for (int e = 0; e < 8; ++e, p += a.t_len) v[u][e] = __ldg(p);
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            const int idx = base + u * kConverters;
                            if (idx < tasks) {
                                const int g = idx / x_rows, q = idx - g * x_rows;
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    split_pair(leaky(v[u][2 * e], a.slope), leaky(v[u][2 * e + 1], a.slope),
                                               hi[e], lo[e]);
                                uint8_t* row = dst + ((size_t)g * Cfg::kXRows + q) * 16;
                                *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(row + Cfg::kGroups * Cfg::kXRows * 16) =
                                    make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_full + xs);
            }
        }
        if (a.debug && ctid == 0) {
            long long* d = a.debug + ((size_t)blockIdx.x * 10 + 2) * 4;
            d[0] = clock64() - begin; d[1] = wait_x;
        }
    } else if (warp < Cfg::kFinalWarp) {
        // ===== mid epilogue: accumulator 1 -> c2's operand image =====
        const int quad = warp & 3;
        long long begin = a.debug ? clock64() : 0, wait_acc = 0, wait_mid = 0, mark = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int t_first = (tile % tiles_per_item) * tile_rows - h2;   // time of mid row 0
            const uint32_t as = i & 1, aphase = (i >> 1) & 1;
            const uint32_t ms = i % MB, mphase = (i / MB) & 1;
            uint8_t* image = mid + ms * Cfg::kMid;
            if (a.debug) mark = clock64();
            mbar_wait(acc1_full + as, aphase);
            if (a.debug) { wait_acc += clock64() - mark; mark = clock64(); }
            mbar_wait(mid_empty + ms, mphase ^ 1);
            if (a.debug) wait_mid += clock64() - mark;
            tc_fence_after();
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
                const int r = s * 128 + quad * 32 + lane;
                const int t = t_first + r;
                const bool live = t >= 0 && t < a.t_len;
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += 16) {
                    const uint32_t address =
                        tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::kAcc + s * Cfg::kCols + c0;
                    uint32_t raw[16];
                    tc_load16(address, raw);
                    if constexpr (CONCAT) {
                        uint32_t other[16];
                        tc_load16(address + C, other);
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(other[e]));
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = c0 + g * 8 + 2 * e;
                            float y0 = leaky(__uint_as_float(raw[g * 8 + 2 * e]) + bias_smem[c], a.slope);
                            float y1 = leaky(__uint_as_float(raw[g * 8 + 2 * e + 1]) + bias_smem[c + 1], a.slope);
                            if (!live) { y0 = 0.f; y1 = 0.f; }
                            split_pair(y0, y1, hi[e], lo[e]);
                        }
                        uint8_t* row = image + ((size_t)(c0 / 8 + g) * Cfg::kMRows + r) * 16;
                        *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(row + (C / 8) * Cfg::kMRows * 16) =
                            make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(acc1_empty + as);
                mbar_arrive(mid_full + ms);
            }
        }
        if (a.debug && lane == 0 && warp == Cfg::kMidWarp) {
            long long* d = a.debug + ((size_t)blockIdx.x * 10 + 3) * 4;
            d[0] = clock64() - begin; d[1] = wait_acc; d[2] = wait_mid;
        }
    } else {
        // ===== final epilogue: accumulator 2 + bias + residual -> fp32 =====
        const int quad = warp & 3;
        constexpr int kSets = NF / 4;                      // warp sets sharing a tile's chunks
        const int half = (warp - Cfg::kFinalWarp) >> 2;
        long long begin = a.debug ? clock64() : 0, wait_acc = 0, mark = 0;
        constexpr int kW = 16;
        constexpr int kPerSub = C / kW;
        constexpr int kChunks = S * kPerSub;
        constexpr int kMine = kChunks / kSets;
        constexpr int kDepth = kMine < 2 ? kMine : 2;     // residual chunks in flight
        static_assert(kChunks % kSets == 0 && (NF == 4 || NF == 8), "chunks are split between the warp sets");
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = tile / tiles_per_item;
            const int t0 = (tile % tiles_per_item) * tile_rows;
            const uint32_t as = i & 1, aphase = (i >> 1) & 1;
            auto row_of = [&](int s) {
                const int r = s * 128 + quad * 32 + lane;
                return (r < tile_rows && t0 + r < a.t_len) ? t0 + r : -1;
            };
            auto fetch = [&](const float* source, int chunk, float (&r)[kW]) {
                const int s = chunk / kPerSub, c0 = (chunk % kPerSub) * kW;
                const int t = row_of(s);
                if (source != nullptr && t >= 0) {
                    const float* from = source + ((size_t)b * C + c0) * a.t_len + t;
#pragma unroll
                    for (int e = 0; e < kW; ++e, from += a.t_len) r[e] = *from;
                } else {
#pragma unroll
                    for (int e = 0; e < kW; ++e) r[e] = 0.f;
                }
            };
            float res[kDepth][kW];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) fetch(a.x, half + kSets * d, res[d]);
            if (a.debug) mark = clock64();
            mbar_wait(acc2_full + as, aphase);
            if (a.debug) wait_acc += clock64() - mark;
            tc_fence_after();
#pragma unroll
            for (int mine = 0; mine < kMine; ++mine) {
                const int chunk = half + kSets * mine;
                const int s = chunk / kPerSub, c0 = (chunk % kPerSub) * kW;
                const int t = row_of(s);
                float acc[kW];
                fetch(a.accum_mode == 2 ? a.accum : nullptr, chunk, acc);
                const uint32_t address =
                    tmem_base + ((uint32_t)(quad * 32) << 16) + 2 * Cfg::kAcc + as * Cfg::kAcc +
                    s * Cfg::kCols + c0;
                uint32_t raw[kW];
                tc_load16(address, raw);
                if constexpr (CONCAT) {
                    uint32_t other[kW];
                    tc_load16(address + C, other);
#pragma unroll
                    for (int e = 0; e < kW; ++e)
                        raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(other[e]));
                }
                float (&r)[kW] = res[mine % kDepth];
                if (t >= 0) {
                    const size_t idx = ((size_t)b * C + c0) * a.t_len + t;
                    float y[kW];
#pragma unroll
                    for (int e = 0; e < kW; ++e) y[e] = __uint_as_float(raw[e]) + r[e] + bias_smem[C + c0 + e];
                    if (a.out) {
                        float* to = a.out + idx;
#pragma unroll
                        for (int e = 0; e < kW; ++e, to += a.t_len) *to = y[e];
                    }
                    if (a.accum_mode) {
                        float* to = a.accum + idx;
#pragma unroll
                        for (int e = 0; e < kW; ++e, to += a.t_len) *to = fmaf(y[e], a.accum_scale, acc[e]);
                    }
                }
                if (mine + kDepth < kMine) fetch(a.x, half + kSets * (mine + kDepth), r);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty + as);
        }
        if (a.debug && lane == 0 && warp == Cfg::kFinalWarp) {
            long long* d = a.debug + ((size_t)blockIdx.x * 10 + 4) * 4;
            d[0] = clock64() - begin; d[1] = wait_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "n"(512) : "memory");
    }
}

int pair_sm_count() {
    static int count = 0;
    if (!count) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
    }
    return count;
}

template <int C, int S, int KB, int NW, bool CONCAT, int NC, int MB, int NF>
int launch_pair_variant(const TcPairArgs& args, cudaStream_t stream) {
    using Cfg = PairConfig<C, S, KB, NW, CONCAT, NC, MB, NF>;
    auto kernel = conv_pair_tc_kernel<C, S, KB, NW, CONCAT, NC, MB, NF>;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem),
            "conv_pair_tc smem attribute"));
        configured = true;
    }
    TcPairArgs a = args;
    a.debug = g_pair_debug;
    const int tile_rows = Cfg::kMidRows - (a.k - 1);
    const int tiles_per_item = ceil_div(a.t_len, tile_rows);
    const int num_tiles = tiles_per_item * a.batch;
    const int grid = min(num_tiles, pair_sm_count());
    LaunchScope scope("conv_pair_tc_kernel", stream);
    kernel<<<grid, Cfg::kThreads, Cfg::kSmem, stream>>>(a, tiles_per_item, num_tiles);
    return launched("conv_pair_tc_kernel");
}

}  // namespace

bool tc_pair_supported(int channels, int k, int dilation) {
    return (channels == 128 || channels == 64 || channels == 32) && k % 2 == 1 && k >= 1 &&
           k - 1 <= kMaxSpan2 && (k - 1) * dilation <= kMaxSpan1;
}

int launch_conv_pair_tc(
    const float* x, const __nv_bfloat16* w1, const float* bias1, const __nv_bfloat16* w2,
    const float* bias2, float* out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float slope, cudaStream_t stream) {
    PMN_REQUIRE(x && w1 && w2, "conv_pair_tc: null input");
    PMN_REQUIRE(out || (accum && accum_mode), "conv_pair_tc: no output");
    PMN_REQUIRE(out != x && accum != x, "conv_pair_tc: the output may not alias the input (halo rows)");
    PMN_REQUIRE(batch > 0 && t_len > 0, "conv_pair_tc: empty input");
    PMN_REQUIRE(tc_pair_supported(channels, k, dilation), "conv_pair_tc: unsupported shape");
    TcPairArgs a;
    a.x = x; a.w1 = w1; a.w2 = w2; a.bias1 = bias1; a.bias2 = bias2;
    a.out = out; a.accum = accum; a.accum_mode = accum ? accum_mode : 0; a.accum_scale = accum_scale;
    a.batch = batch; a.t_len = t_len; a.k = k; a.dilation = dilation; a.slope = slope;
    // variants: 0 = 4 converter + 8 final warps, one mid image; 1 = 8 converter + 4 final warps;
    // 2 = variant 1 with two mid images (C <= 64: C = 128 has no room for the second);
    // 3 = 8 converter + 8 final warps, two mid images
    const int variant = g_pair_variant < 0 ? kDefaultVariant : g_pair_variant;
    if (channels == 128) {
        if (variant == 0) return launch_pair_variant<128, 1, 64, 2, false, 4, 1, 8>(a, stream);
        if (variant == 3) return launch_pair_variant<128, 1, 64, 2, false, 8, 1, 8>(a, stream);
        return launch_pair_variant<128, 1, 64, 2, false, 8, 1, 4>(a, stream);
    }
    if (channels == 64) {
        if (variant == 0) return launch_pair_variant<64, 1, 64, 4, true, 4, 1, 8>(a, stream);
        if (variant == 1) return launch_pair_variant<64, 1, 64, 4, true, 8, 1, 4>(a, stream);
        if (variant == 3) return launch_pair_variant<64, 1, 64, 4, true, 8, 2, 8>(a, stream);
        return launch_pair_variant<64, 1, 64, 4, true, 8, 2, 4>(a, stream);
    }
    if (variant == 0) return launch_pair_variant<32, 2, 32, 8, true, 4, 1, 8>(a, stream);
    if (variant == 1) return launch_pair_variant<32, 2, 32, 8, true, 8, 1, 4>(a, stream);
    if (variant == 3) return launch_pair_variant<32, 2, 32, 8, true, 8, 2, 8>(a, stream);
    return launch_pair_variant<32, 2, 32, 8, true, 8, 2, 4>(a, stream);
}

void tc_pair_set_debug(long long* counters, int variant) {
    g_pair_debug = counters;
    g_pair_variant = variant;
}

}  // namespace pmn

// Dilated Conv1d on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-grade.
//
// The dense residual-block convolutions of HiFi-GAN (Block.forward,
// promonet/model/hifigan.py:198-210; 97 % of the generator's FLOPs) as an
// implicit GEMM per tap:
//     D[t, o] = sum_j sum_c A_j[t, c] * W_j[c, o],   A_j[t, c] = a[c, t + (j - h) d]
// with M = 128 time steps (TMEM lanes), N = C_out (TMEM columns), K = C_in.
//
// fp32 parity on bf16 tensor cores: every operand is split a = a_hi + a_lo
// (bf16 each) and three products a_hi w_hi + a_lo w_hi + a_hi w_lo are
// accumulated in fp32 in TMEM; the dropped a_lo w_lo term is 2^-16 relative.
//
// Operands are K-major, un-swizzled ("interleaved") core matrices laid out as
// [k / 8][row][8]: rows are contiguous at 16 B, so the tap shift (j - h) d is a
// plain 16 B-granular offset of the A descriptor's start address and one staged
// time window (tile + halo) serves all K taps.  The global layout (conv1d_tc.cuh)
// is the same, so staging is a handful of 1-D bulk async copies
// (cp.async.bulk, completion on an mbarrier) per slab -- no tensor maps.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0    producer: bulk copies of activation slabs (per 64-channel block)
//             and weight slabs (per tap x block) into shared-memory rings
//   warp 1    MMA issuer: one thread issues tcgen05.mma, commits to mbarriers
//   warps 4-11 epilogue (two warpgroups, registers raised with setmaxnreg): tcgen05.ld
//             accumulators, bias + residual + LeakyReLU, fp32 and/or operand-plane stores,
//             MRF accumulate (hifigan.py:141-145); warps 2-3 idle (they complete the first warpgroup)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps
// the MMAs of tile i + 1.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdlib.h>

#include <vector>

#include "conv1d_tc.cuh"
#include "tc_ptx.cuh"

namespace pmn {

namespace {

using namespace tc;

// LeakyReLU for slopes in [0, 1] (1 = identity) as a multiply and a maximum
__device__ __forceinline__ float leaky01(float x, float slope) { return fmaxf(x, x * slope); }

// Three warpgroups: {producer, MMA issuer, two idle warps} and two of epilogue warps.  The launch
// bound of 384 threads caps every thread at 168 registers, which the fully unrolled epilogue
// overflows; the first warpgroup hands registers back (setmaxnreg.dec) and the epilogue warps take
// them (setmaxnreg.inc): 128 x 56 + 256 x 224 = 384 x 168.
constexpr int kThreads = 384;
constexpr int kEpilogueWarp0 = 4;
constexpr int kMaxHalo = 25;  // (11 - 1) / 2 * 5

// ---------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------

enum : int {
    kConv = 0,       // Conv1d, "same" or valid padding, rows = time steps
    kTranspose = 1,  // ConvTranspose1d as a 3-tap convolution with UP * c_out columns
    kFrames = 2      // valid Conv1d over frames laid end to end: M = 16 frames x 8 rows
};

constexpr int kFrameRows = 8;      // output rows computed per frame in kFrames mode
constexpr int kFramesPerTile = 16;
constexpr int kMaxFrameLength = 35;

// C_IN input channels (the K of the GEMM), N output columns per tile (TMEM columns
// of one MMA, or half of them when CONCAT), S x 128 output rows per tile, KB input
// channels per shared-memory slab, NW weight-slab stages, AS accumulator stages.
// CONCAT: the B operand is [W_hi; W_lo] (2 N rows), so a_hi w_hi and a_hi w_lo
// come out of ONE MMA (columns [0, N) and [N, 2 N)) and a_lo w_hi of a second one
// that accumulates into [0, N): two reads of the 128-row A operand per K chunk
// instead of three, which is what bounds the narrow (N <= 64) layers.
// XS activation-slab stages: 2 everywhere except the k = 1 layers with a long K (frame-major pitch
// blocks 4 and 5), where a 32-channel slab is only 768 cycles of MMAs and two stages do not cover
// the L2 latency of the next slab.
// F8: "fp16 + 2 x fp8" operands (conv1d_tc.cuh): two kind::f16 and two kind::f8f6f4 MMAs per 32 input
// channels instead of six bf16 ones, all into the same accumulator.
template <int C_IN, int N, int S, int KB, int NW, int AS, int MODE, bool CONCAT, int XS = 2, bool F8 = false>
struct TcConfig {
    static constexpr int kTile = S * 128;
    static constexpr int kRowsMax =
        MODE == kFrames ? (kFramesPerTile - 1) * kMaxFrameLength + kFrameRows + 31 : kTile + 2 * kMaxHalo;
    static constexpr int kGroups = KB / 8;                      // 8-channel groups per K block
    static constexpr int kBlocks = C_IN / KB;
    static constexpr int kXSlab = 2 * kGroups * kRowsMax * 16;  // bytes, both planes
    static constexpr int kWSlab = KB * N * 4;                   // bytes, both planes
    static constexpr int kXStages = XS;
    static constexpr int kBarriers = 2 * kXStages + 2 * NW + 2 * AS;
    static constexpr int kSmem = kXStages * kXSlab + NW * kWSlab + kBarriers * 8 + 16 + 128 + 6144;
    static constexpr int kCols = CONCAT ? 2 * N : N;            // TMEM columns per 128 rows
    static constexpr int kStageCols = S * kCols;                // TMEM columns of one accumulator stage
    static constexpr int kColumns = AS * kStageCols;
    static constexpr int kAlloc = kColumns <= 32 ? 32 : kColumns <= 64 ? 64 : kColumns <= 128 ? 128
                                  : kColumns <= 256 ? 256 : 512;
    static_assert(kColumns <= 512, "accumulators exceed TMEM");
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
    static_assert(C_IN % KB == 0 && KB % 16 == 0 && N % 32 == 0 && kCols <= 256, "shape");
    static_assert(MODE != kFrames || S == 1, "frame mode computes one 128-row tile");
    static_assert(!F8 || (MODE == kConv && !CONCAT && KB % 32 == 0), "fp8 corrections: plain convolutions only");
};

template <int C_IN, int N, int S, int KB, int NW, int AS, int MODE, int UP, bool CONCAT, int XS, bool F8>
__global__ void __launch_bounds__(kThreads, 1) conv1d_tc_kernel(
    TcConvArgs a, int t_pad, int tiles_per_item, int n_tiles, int num_tiles) {
    using Cfg = TcConfig<C_IN, N, S, KB, NW, AS, MODE, CONCAT, XS, F8>;
    extern __shared__ uint8_t smem_raw[];
    // 128-byte alignment by pointer arithmetic on the __shared__ array: through an integer cast the
    // compiler loses the address space and emits generic LD / ST for every access to the buffers
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* x_slabs = smem;
    uint8_t* w_slabs = smem + Cfg::kXStages * Cfg::kXSlab;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_slabs + NW * Cfg::kWSlab);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + Cfg::kXStages;
    uint64_t* w_full = x_empty + Cfg::kXStages;
    uint64_t* w_empty = w_full + NW;
    uint64_t* acc_full = w_empty + NW;
    uint64_t* acc_empty = acc_full + AS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + AS);
    float* bias_smem = reinterpret_cast<float*>(tmem_slot + 4);  // c_out (<= 1536) floats

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int span = (a.k - 1) * a.dilation;               // receptive field minus one
    const int left = a.valid ? 0 : span / 2;               // rows staged before the tile
    // rows of the staged window and rows advanced per tile
    const int rows = MODE == kFrames ? (kFramesPerTile - 1) * a.frame_length + kFrameRows + span
                                     : Cfg::kTile + span;
    const int advance = MODE == kFrames ? kFramesPerTile * a.frame_length : Cfg::kTile;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::kXStages; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        for (int i = 0; i < NW; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
        for (int i = 0; i < AS; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 1536; i += kThreads)
        bias_smem[i] = (a.bias && i < a.c_out) ? a.bias[i] : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kAlloc) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kEpilogueWarp0) {
    // ---- first warpgroup: producer, MMA issuer, two idle warps (role bodies keep their indentation) ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t xcount = 0, wcount = 0;
            long long wait_x = 0, wait_w = 0, begin = a.debug ? clock64() : 0, mark = 0;
            const uint32_t x_bytes = 2 * Cfg::kGroups * rows * 16;
            const int plane_groups = a.plane_groups > 0 ? a.plane_groups : C_IN / 8;
            const int item_groups = a.item_groups > 0 ? a.item_groups : 2 * (C_IN / 8);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % n_tiles, rest = tile / n_tiles;
                const int b = rest / tiles_per_item;
                const int t0 = (rest % tiles_per_item) * advance;
                for (int kb = 0; kb < Cfg::kBlocks; ++kb) {
                    const uint32_t xs = xcount % Cfg::kXStages, xphase = (xcount / Cfg::kXStages) & 1;
                    ++xcount;
                    if (a.debug) mark = clock64();
                    mbar_wait(x_empty + xs, xphase ^ 1);
                    if (a.debug) wait_x += clock64() - mark;
                    mbar_expect_tx(x_full + xs, x_bytes);
                    uint8_t* dst = x_slabs + xs * Cfg::kXSlab;
                    if constexpr (F8) {
                        // fp16 section: kGroups row groups of this K block; the two e4m3 sections:
                        // kGroups / 2 row groups (16 channels a row) each
#pragma unroll 1
                        for (int g = 0; g < 2 * Cfg::kGroups; ++g) {
                            const int section = g < Cfg::kGroups ? 0 : 1 + (g - Cfg::kGroups) / (Cfg::kGroups / 2);
                            const int within = g < Cfg::kGroups ? g : (g - Cfg::kGroups) % (Cfg::kGroups / 2);
                            const size_t group =
                                (size_t)b * item_groups +
                                (section == 0 ? kb * Cfg::kGroups
                                              : C_IN / 8 + (section - 1) * (C_IN / 16) + kb * (Cfg::kGroups / 2)) + within;
                            bulk_copy(dst + g * rows * 16, a.x_planes + (group * t_pad + kTcPad + t0 - left) * 8,
                                      rows * 16, x_full + xs);
                        }
                    } else {
#pragma unroll 1
                    for (int p = 0; p < 2; ++p) {
#pragma unroll 1
                        for (int g = 0; g < Cfg::kGroups; ++g) {
                            const size_t row0 =
                                ((size_t)b * item_groups + (size_t)p * plane_groups + kb * Cfg::kGroups + g) *
                                    t_pad + kTcPad + t0 - left;
                            bulk_copy(dst + (p * Cfg::kGroups + g) * rows * 16,
                                      a.x_planes + row0 * 8, rows * 16, x_full + xs);
                        }
                    }
                    }
                    for (int tap = 0; tap < a.k; ++tap) {
                        const uint32_t ws = wcount % NW, wphase = (wcount / NW) & 1;
                        ++wcount;
                        if (a.debug) mark = clock64();
                        mbar_wait(w_empty + ws, wphase ^ 1);
                        if (a.debug) wait_w += clock64() - mark;
                        mbar_expect_tx(w_full + ws, Cfg::kWSlab);
                        bulk_copy(w_slabs + ws * Cfg::kWSlab,
                                  reinterpret_cast<const uint8_t*>(a.w_slabs) +
                                      (size_t)((nt * a.k + tap) * Cfg::kBlocks + kb) * Cfg::kWSlab,
                                  Cfg::kWSlab, w_full + ws);
                    }
                }
            }
            if (a.debug) {
                long long* d = a.debug + ((size_t)blockIdx.x * 10 + 0) * 4;
                d[0] = clock64() - begin; d[1] = wait_x; d[2] = wait_w;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(128, N);
            constexpr uint32_t idesc_wide = instr_desc(128, Cfg::kCols);
            uint32_t xcount = 0, wcount = 0, tcount = 0;
            long long wait_x = 0, wait_w = 0, wait_acc = 0, begin = a.debug ? clock64() : 0, mark = 0;
            const uint32_t x_plane = Cfg::kGroups * rows * 16;      // bytes between hi and lo
            // 8-row groups of the A operand: consecutive rows, or one group per frame
            const uint32_t a_sbo = MODE == kFrames ? a.frame_length * 16 : 128;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const uint32_t as = tcount % AS, aphase = (tcount / AS) & 1;
                ++tcount;
                if (a.debug) mark = clock64();
                mbar_wait(acc_empty + as, aphase ^ 1);
                if (a.debug) wait_acc += clock64() - mark;
                tc_fence_after();
                const uint32_t d_base = tmem_base + as * Cfg::kStageCols;
                for (int kb = 0; kb < Cfg::kBlocks; ++kb) {
                    const uint32_t xs = xcount % Cfg::kXStages, xphase = (xcount / Cfg::kXStages) & 1;
                    ++xcount;
                    if (a.debug) mark = clock64();
                    mbar_wait(x_full + xs, xphase);
                    if (a.debug) wait_x += clock64() - mark;
                    const uint32_t x_addr = smem_u32(x_slabs + xs * Cfg::kXSlab);
                    for (int tap = 0; tap < a.k; ++tap) {
                        const uint32_t ws = wcount % NW, wphase = (wcount / NW) & 1;
                        ++wcount;
                        if (a.debug) mark = clock64();
                        mbar_wait(w_full + ws, wphase);
                        if (a.debug) wait_w += clock64() - mark;
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(w_slabs + ws * Cfg::kWSlab);
                        const bool first = kb == 0 && tap == 0;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const uint32_t row = s * 128 + tap * a.dilation;
                            const uint32_t d = d_base + s * Cfg::kCols;
                            if constexpr (F8) {
                                constexpr uint32_t idesc0 = instr_desc_format0(128, N);
                                const uint32_t x8 = x_addr + Cfg::kGroups * rows * 16;
                                const uint32_t w8 = w_addr + Cfg::kGroups * N * 16;
#pragma unroll
                                for (int kk = 0; kk < KB / 32; ++kk) {
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {
                                        // fp16 x fp16: channels 32 kk + 16 h .. + 15 (two 8-channel row groups)
                                        const uint64_t a16 = smem_desc(
                                            x_addr + ((4 * kk + 2 * h) * rows + row) * 16, rows * 16, a_sbo);
                                        const uint64_t b16 = smem_desc(w_addr + (4 * kk + 2 * h) * N * 16, N * 16, 128);
                                        tc_mma(d, a16, b16, idesc0, !(first && kk == 0 && h == 0));
                                    }
                                    // e4m3(x) e4m3(w low) + e4m3(x low) e4m3(w): channels 32 kk .. + 31
                                    const uint32_t xa8 = x8 + (2 * kk * rows + row) * 16;
                                    const uint32_t wa8 = w8 + 2 * kk * N * 16;
                                    constexpr uint32_t w_half = (Cfg::kGroups / 2) * N * 16;
                                    const uint32_t x_half = (Cfg::kGroups / 2) * rows * 16;
                                    tc_mma_f8(d, smem_desc(xa8, rows * 16, a_sbo), smem_desc(wa8, N * 16, 128), idesc0, 1);
                                    tc_mma_f8(d, smem_desc(xa8 + x_half, rows * 16, a_sbo),
                                              smem_desc(wa8 + w_half, N * 16, 128), idesc0, 1);
                                }
                            } else {
#pragma unroll
                            for (int kk = 0; kk < KB / 16; ++kk) {
                                const uint32_t xa = x_addr + (2 * kk * rows + row) * 16;
                                const uint64_t a_hi = smem_desc(xa, rows * 16, a_sbo);
                                const uint64_t a_lo = smem_desc(xa + x_plane, rows * 16, a_sbo);
                                if constexpr (CONCAT) {
                                    // slab [k group][hi rows 0..N) | lo rows N..2N)][8]
                                    const uint32_t wa = w_addr + 2 * kk * (2 * N) * 16;
                                    const uint64_t b_both = smem_desc(wa, 2 * N * 16, 128);
                                    tc_mma(d, a_hi, b_both, idesc_wide, !(first && kk == 0));
                                    tc_mma(d, a_lo, b_both, idesc, 1);
                                } else {
                                    // slab [plane][k group][N][8]
                                    constexpr uint32_t w_plane = Cfg::kGroups * N * 16;
                                    const uint32_t wa = w_addr + 2 * kk * N * 16;
                                    const uint64_t b_hi = smem_desc(wa, N * 16, 128);
                                    const uint64_t b_lo = smem_desc(wa + w_plane, N * 16, 128);
                                    tc_mma(d, a_hi, b_hi, idesc, !(first && kk == 0));
                                    tc_mma(d, a_lo, b_hi, idesc, 1);
                                    tc_mma(d, a_hi, b_lo, idesc, 1);
                                }
                            }
                            }
                        }
                        tc_commit(w_empty + ws);
                    }
                    tc_commit(x_empty + xs);
                }
                tc_commit(acc_full + as);
            }
            if (a.debug) {
                long long* d = a.debug + ((size_t)blockIdx.x * 10 + 1) * 4;
                d[0] = clock64() - begin; d[1] = wait_x; d[2] = wait_w; d[3] = wait_acc;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===== epilogue: warps 4..11; warp w owns TMEM lanes 32 * (w % 4) .. + 31 and
        // every other 16-column chunk of the tile =====
        const int quad = warp & 3;
        const int half = (warp - kEpilogueWarp0) >> 2;
        constexpr int kW = 16;                    // columns per chunk
        // Rarely used epilogue options are compiled only into the variants that take them: the loop
        // over a tile's chunks is fully unrolled and every extra branch in it costs registers and
        // scheduling freedom in all of them (C = 128 / 32, k = 3 ran 20 - 45 % slower with these as
        // run-time flags, profiles/r2_tc_epilogue_regression.txt)
        constexpr bool kBiasBatch = MODE == kConv && C_IN == 128 && N == 256;   // HiFi-GAN's input conv
        constexpr bool kOutF8 = MODE == kConv ? F8 : UP == 8;   // "fp16 + 2 x fp8" operand for the next layer
        constexpr int kPerSub = N / kW;
        constexpr int kChunks = S * kPerSub;      // (subtile, 16-channel) chunks per tile
        static_assert(kChunks % 2 == 0, "chunks are split between two warp sets");
        uint32_t tcount = 0;
        const int groups_out = a.c_out / 8;
        const int t_out = a.valid ? a.t_len - span : a.t_len;   // output rows per item
        const int out_row = a.out_row > 0 ? a.out_row : t_out;  // fp32 row length
        const size_t item_elements = (size_t)a.c_out * out_row;    // < 2^31 (launcher)
        const float relu_floor = a.relu ? 0.f : -INFINITY;
        long long wait_cycles = 0, start_cycles = a.debug ? clock64() : 0;
        // residual chunks in flight (kConv / kFrames epilogue): they outlive a tile
        float res[(S * (N / kW) / 2) < 4 ? (S * (N / kW) / 2) : 4][kW];
        // this thread's output row of subtile s of the tile at position of_tb of its item, or -1
        auto row_at = [&](int of_tb, int s) {
            const int within = quad * 32 + lane;
            if constexpr (MODE == kFrames) {
                const int frame = of_tb * kFramesPerTile + within / kFrameRows, t = within % kFrameRows;
                return (frame < a.frames && t < a.frame_valid) ? frame * a.frame_length + t : -1;
            } else {
                const int t = of_tb * advance + s * 128 + within;
                return t < t_out ? t : -1;
            }
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles, rest = tile / n_tiles;
            const int b = rest / tiles_per_item;
            const int tb = rest % tiles_per_item;
            const uint32_t as = tcount % AS, aphase = (tcount / AS) & 1;
            ++tcount;
            auto row_of = [&](int s) { return row_at(tb, s); };
            auto load = [&](int s, int c0, uint32_t (&raw)[kW]) {
                const uint32_t address =
                    tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::kStageCols + s * Cfg::kCols + c0;
                tc_load16(address, raw);
                if constexpr (F8) {
#pragma unroll
                    for (int i = 0; i < kW; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * a.f8_unscale);
                }
                if constexpr (CONCAT) {
                    uint32_t other[kW];
                    tc_load16(address + N, other);
#pragma unroll
                    for (int i = 0; i < kW; ++i)
                        raw[i] = __float_as_uint(__uint_as_float(raw[i]) + __uint_as_float(other[i]));
                }
            };
            if constexpr (MODE == kTranspose) {
                // column n = o * UP + q is output sample UP * i + q of channel o;
                // a thread writes UP contiguous floats per o
                mbar_wait(acc_full + as, aphase);
                tc_fence_after();
                const int t_up = UP * a.t_len;
                // chunks are taken in runs that cover 8 output channels (8 / (16 / UP) chunks), so that
                // the run's thread also holds what a 16-byte row of the output's planes needs
                constexpr int kRun = UP >= 8 ? 8 * UP / kW : 1;
                static_assert(kChunks % (2 * kRun) == 0, "chunk runs are split between two warp sets");
                const int out_pad = tc_padded_length_device(t_up);
                const int groups_out = a.c_out / 8;
#pragma unroll 1
                for (int run = half; run < kChunks / kRun; run += 2) {
                    float values[kRun * kW];     // [channel within the run][phase]
                    const int s = run * kRun / kPerSub, c_run = (run * kRun % kPerSub) * kW;
                    const int i = row_of(s);
                    const int o_run = (nt * N + c_run) / UP;
#pragma unroll
                    for (int part = 0; part < kRun; ++part) {
                        const int c0 = c_run + part * kW;
                        uint32_t raw[kW];
                        load(s, c0, raw);
                        const int o0 = (nt * N + c0) / UP;
#pragma unroll
                        for (int j = 0; j < kW / UP; ++j) {
                            const float bias = bias_smem[o0 + j];
#pragma unroll
                            for (int q = 0; q < UP; ++q)
                                values[part * kW + j * UP + q] = __uint_as_float(raw[j * UP + q]) + bias;
                            if (i >= 0) {
                                float* dst = a.out + ((size_t)b * a.c_out + o0 + j) * t_up + (size_t)UP * i;
                                const float* v = values + part * kW + j * UP;
                                if constexpr (UP % 4 == 0) {
#pragma unroll
                                    for (int q = 0; q < UP; q += 4)
                                        *reinterpret_cast<float4*>(dst + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                                } else {
#pragma unroll
                                    for (int q = 0; q < UP; q += 2)
                                        *reinterpret_cast<float2*>(dst + q) = make_float2(v[q], v[q + 1]);
                                }
                            }
                        }
                    }
                    if (kOutF8 && a.out_planes && a.out_f8 && i >= 0) {
                        // the run's 8 channels: one fp16 row and half a row of each e4m3 section per sample
                        const size_t item = (size_t)b * (groups_out * 2);
                        const int half_row = (o_run / 8) & 1;
#pragma unroll
                        for (int q = 0; q < UP; ++q) {
                            uint32_t main[4], coarse[2], low[2];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                uint32_t z, l;
                                split_pair_f8(leaky01(values[(2 * e) * UP + q], a.out_slope),
                                              leaky01(values[(2 * e + 1) * UP + q], a.out_slope), main[e], z, l);
                                if (e % 2 == 0) { coarse[e / 2] = z; low[e / 2] = l; }
                                else { coarse[e / 2] |= z << 16; low[e / 2] |= l << 16; }
                            }
                            const size_t at = kTcPad + (size_t)UP * i + q;
                            uint4* rows = reinterpret_cast<uint4*>(a.out_planes);
                            rows[(item + o_run / 8) * out_pad + at] = make_uint4(main[0], main[1], main[2], main[3]);
                            uint2* coarse_row = reinterpret_cast<uint2*>(rows + (item + groups_out + o_run / 16) * out_pad + at);
                            uint2* low_row = reinterpret_cast<uint2*>(
                                rows + (item + groups_out + groups_out / 2 + o_run / 16) * out_pad + at);
                            coarse_row[half_row] = make_uint2(coarse[0], coarse[1]);
                            low_row[half_row] = make_uint2(low[0], low[1]);
                        }
                    } else if (a.out_planes && i >= 0) {
                        // planes of lrelu(y): per 8 channels and output sample one 16-byte row per plane
                        constexpr int kChannels = kRun * kW / UP;        // 8 (UP = 8: one run) or 8 (UP = 2: one chunk)
                        static_assert(kChannels == 8, "a run holds 8 output channels");
#pragma unroll
                        for (int q = 0; q < UP; ++q) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                split_pair(leaky01(values[(2 * e) * UP + q], a.out_slope),
                                           leaky01(values[(2 * e + 1) * UP + q], a.out_slope), hi[e], lo[e]);
                            const size_t row_hi =
                                ((size_t)(b * 2) * groups_out + o_run / 8) * out_pad + kTcPad + (size_t)UP * i + q;
                            const size_t row_lo = row_hi + (size_t)groups_out * out_pad;
                            *reinterpret_cast<uint4*>(a.out_planes + row_hi * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(a.out_planes + row_lo * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + as);
                continue;
            }
            // Side inputs do not depend on the accumulators.  The residual of kDepth chunks is always
            // in flight, across tile boundaries too (the first chunks of the next tile are requested
            // while the last ones of this tile are processed), so its DRAM latency is paid once per
            // launch; the MRF partial sum is fetched one chunk ahead.  The chunk loop is unrolled in
            // groups of kDepth: fully unrolled, the larger variants' epilogue (160 KB of SASS)
            // stalled on instruction fetch for a quarter of its cycles (ncu, profiles/r2_tc_epilogue_regression.txt)
            constexpr int kMine = kChunks / 2;                 // chunks of this warp set per tile
            constexpr int kDepth = kMine < 4 ? kMine : 4;      // residual chunks in flight
            constexpr int kGroup = kMine % kDepth == 0 ? kDepth : kMine;   // chunks per unrolled group
            static_assert(kGroup % 2 == 0 || kGroup == kMine, "the accumulate buffers alternate");
            // chunk c of every tile uses buffer c % kDepth: the pipeline runs across tiles when that
            // agrees with the rotation, i.e. when kDepth divides kMine (all but the 160-column head)
            constexpr bool kAcross = kMine % kDepth == 0;
            auto fetch = [&](const float* source, int of_tile, int chunk, float (&r)[kW]) {
                const int s = chunk / kPerSub, c0 = (chunk % kPerSub) * kW;
                const int of_nt = of_tile % n_tiles, of_rest = of_tile / n_tiles;
                const int of_b = of_rest / tiles_per_item, of_tb = of_rest % tiles_per_item;
                const int t = row_at(of_tb, s);
                // one 64-bit base per chunk, 32-bit element offsets within the item (the integer
                // address arithmetic was half of the epilogue's instructions)
                if (source != nullptr && t >= 0 && of_tile < num_tiles) {
                    const float* from = source + (size_t)of_b * item_elements +
                                        ((uint32_t)(of_nt * N + c0) * (uint32_t)out_row + (uint32_t)t);
#pragma unroll
                    for (int i = 0; i < kW; ++i, from += out_row) r[i] = *from;
                } else {
#pragma unroll
                    for (int i = 0; i < kW; ++i) r[i] = 0.f;
                }
            };
            if (!kAcross || tile == (int)blockIdx.x) {
#pragma unroll
                for (int d = 0; d < kDepth; ++d) fetch(a.residual, tile, half + 2 * d, res[d]);
            }
            // the MRF partial sum this launch adds to (accum_mode 2)
            const float* accum_in = a.accum_mode == 2 ? a.accum : nullptr;
            float acc[2][kW];
            fetch(accum_in, tile, half, acc[0]);
            const long long wait_start = a.debug ? clock64() : 0;
            mbar_wait(acc_full + as, aphase);
            if (a.debug) wait_cycles += clock64() - wait_start;
            tc_fence_after();
            // groups of kGroup chunks: the outer loop is a real loop, the inner one is unrolled (its body
            // keeps the indentation it had as a single loop)
#pragma unroll 1
            for (int group = 0; group < kMine / kGroup; ++group) {
#pragma unroll
            for (int g_index = 0; g_index < kGroup; ++g_index) {
                const int mine = group * kGroup + g_index;
                const int d = g_index % kDepth;     // static: the residual buffer of this chunk
                const int chunk = half + 2 * mine;
                const int s = chunk / kPerSub, c0 = (chunk % kPerSub) * kW;
                const int t = row_of(s);
                if (mine + 1 < kMine) fetch(accum_in, tile, chunk + 2, acc[(d + 1) & 1]);
                uint32_t raw[kW];
                load(s, c0, raw);
                float (&r)[kW] = res[d];
                float v[kW];
                const int c_first = nt * N + c0;
                float bias_of[kW];
#pragma unroll
                for (int q = 0; q < kW / 4; ++q) {
                    const float4 four = *reinterpret_cast<const float4*>(bias_smem + c_first + 4 * q);
                    bias_of[4 * q] = four.x; bias_of[4 * q + 1] = four.y;
                    bias_of[4 * q + 2] = four.z; bias_of[4 * q + 3] = four.w;
                }
#pragma unroll
                for (int i = 0; i < kW; ++i) {
                    float y = __uint_as_float(raw[i]) + r[i] + bias_of[i];
                    if constexpr (kBiasBatch) {
                        if (a.bias_batch) y += a.bias_batch[(size_t)b * a.c_out + c_first + i];
                    }
                    v[i] = fmaxf(y, relu_floor);
                }
                int row = t;
                if (a.pool) {
                    // MaxPool1d(2) over row pairs (2 i, 2 i + 1): adjacent lanes; even lanes store
#pragma unroll
                    for (int i = 0; i < kW; ++i)
                        v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
                    row = (t >= 0 && (lane & 1) == 0 && t + 1 < t_out) ? t / 2 : -1;
                }
                if (row >= 0) {
                    const uint32_t first = (uint32_t)c_first * (uint32_t)out_row + (uint32_t)row;
                    if (a.out) {
                        float* to = a.out + (size_t)b * item_elements + first;
#pragma unroll
                        for (int i = 0; i < kW; ++i, to += out_row) *to = v[i];
                    }
                    if (a.accum_mode) {
                        float* to = a.accum + (size_t)b * item_elements + first;
#pragma unroll
                        for (int i = 0; i < kW; ++i, to += out_row) {
                            const float total = fmaf(v[i], a.accum_scale, acc[d & 1][i]);
                            *to = total;
                            if (a.planes_from_accum) v[i] = total;   // the planes below are those of the sum
                        }
                    }
                    if (kOutF8 && a.out_planes && a.out_f8) {
                        // [C / 8 fp16 rows | C / 16 coarse rows | C / 16 low rows] per item
                        const int out_pad = tc_padded_length_device(t_out);
                        uint32_t main[8], coarse[4], low[4];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            uint32_t z, l;
                            split_pair_f8(leaky01(v[2 * e], a.out_slope), leaky01(v[2 * e + 1], a.out_slope), main[e], z, l);
                            if (e % 2 == 0) { coarse[e / 2] = z; low[e / 2] = l; }
                            else { coarse[e / 2] |= z << 16; low[e / 2] |= l << 16; }
                        }
                        uint4* rows = reinterpret_cast<uint4*>(a.out_planes) +
                                      (size_t)b * (groups_out * 2) * out_pad + kTcPad + t;
                        const uint32_t group = c_first / 8, pad = out_pad;
                        rows[group * pad] = make_uint4(main[0], main[1], main[2], main[3]);
                        rows[(group + 1) * pad] = make_uint4(main[4], main[5], main[6], main[7]);
                        rows[(groups_out + group / 2) * pad] = make_uint4(coarse[0], coarse[1], coarse[2], coarse[3]);
                        rows[(groups_out + groups_out / 2 + group / 2) * pad] = make_uint4(low[0], low[1], low[2], low[3]);
                    } else if (a.out_planes) {
                        const int out_pad = tc_padded_length_device(t_out);
                        uint4* rows = reinterpret_cast<uint4*>(a.out_planes) +
                                      (size_t)b * (groups_out * 2) * out_pad + kTcPad + t;
                        const uint32_t pad = out_pad;
#pragma unroll
                        for (int g = 0; g < kW / 8; ++g) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                split_pair(leaky01(v[g * 8 + 2 * e], a.out_slope),
                                           leaky01(v[g * 8 + 2 * e + 1], a.out_slope), hi[e], lo[e]);
                            const uint32_t group = c_first / 8 + g;
                            rows[group * pad] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            rows[(groups_out + group) * pad] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
                // the buffer is free: request the chunk kDepth further on, in this tile or the next
                {
                    const int ahead = mine + kDepth;
                    const bool same = ahead < kMine;
                    if (same || kAcross)
                        fetch(a.residual, same ? tile : tile + (int)gridDim.x,
                              half + 2 * (same ? ahead : ahead - kMine), r);
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);
        }
        if (a.debug && lane == 0) {
            long long* d = a.debug + ((size_t)blockIdx.x * 10 + warp - 2) * 4;
            d[0] = clock64() - start_cycles;
            d[1] = wait_cycles;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "n"(Cfg::kAlloc) : "memory");
    }
}

// fp32 (B, C, T) -> hi/lo planes of lrelu(x); rows outside [0, T) are zeroed
__global__ void __launch_bounds__(128) planes_from_f32_kernel(
    const float* __restrict__ x, __nv_bfloat16* __restrict__ planes,
    int channels, int t_len, int t_pad, float slope, int source_channels) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= t_pad) return;
    const int g = blockIdx.y, b = blockIdx.z;
    const int groups = channels / 8;
    const int t = row - kTcPad;
    uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    if (t >= 0 && t < t_len) {
        const float* src = x + ((size_t)b * source_channels + g * 8) * t_len + t;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = g * 8 + 2 * e;       // channels past the source's are zero
            const float y0 = c < source_channels ? __ldg(src + (size_t)(2 * e) * t_len) : 0.f;
            const float y1 = c + 1 < source_channels ? __ldg(src + (size_t)(2 * e + 1) * t_len) : 0.f;
            split_pair(leaky01(y0, slope), leaky01(y1, slope), hi[e], lo[e]);
        }
    }
    const size_t row_hi = ((size_t)(b * 2) * groups + g) * t_pad + row;
    const size_t row_lo = row_hi + (size_t)groups * t_pad;
    *reinterpret_cast<uint4*>(planes + row_hi * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(planes + row_lo * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// fp32 (B, C, T) -> the "fp16 + 2 x fp8" operand of lrelu(x) (conv1d_tc.cuh); one thread per row and
// 16 channels; rows outside [0, T) are zeroed
__global__ void __launch_bounds__(128) planes_f8_from_f32_kernel(
    const float* __restrict__ x, uint4* __restrict__ rows, int channels, int t_len, int t_pad, float slope) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= t_pad) return;
    const int g = blockIdx.y, b = blockIdx.z;       // g: 16-channel group
    const int groups = channels / 8;
    const int t = row - kTcPad;
    uint32_t main[8] = {0, 0, 0, 0, 0, 0, 0, 0}, coarse[4] = {0, 0, 0, 0}, low[4] = {0, 0, 0, 0};
    if (t >= 0 && t < t_len) {
        const float* src = x + ((size_t)b * channels + g * 16) * t_len + t;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            uint32_t z, l;
            split_pair_f8(leaky01(__ldg(src + (size_t)(2 * e) * t_len), slope),
                          leaky01(__ldg(src + (size_t)(2 * e + 1) * t_len), slope), main[e], z, l);
            if (e % 2 == 0) { coarse[e / 2] = z; low[e / 2] = l; }
            else { coarse[e / 2] |= z << 16; low[e / 2] |= l << 16; }
        }
    }
    const size_t item = (size_t)b * (groups * 2);
    rows[(item + 2 * g) * t_pad + row] = make_uint4(main[0], main[1], main[2], main[3]);
    rows[(item + 2 * g + 1) * t_pad + row] = make_uint4(main[4], main[5], main[6], main[7]);
    rows[(item + groups + g) * t_pad + row] = make_uint4(coarse[0], coarse[1], coarse[2], coarse[3]);
    rows[(item + groups + groups / 2 + g) * t_pad + row] = make_uint4(low[0], low[1], low[2], low[3]);
}

// the same operand -> fp32 (B, C, T): main / s_m + low / s_xl (tests)
__global__ void __launch_bounds__(128) f32_from_planes_f8_kernel(
    const uint8_t* __restrict__ planes, float* __restrict__ x, int channels, int t_len, int t_pad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_len) return;
    const int g = blockIdx.y, b = blockIdx.z;       // g: 8-channel group
    const int groups = channels / 8;
    const size_t item = (size_t)b * (groups * 2);
    const __half* main = reinterpret_cast<const __half*>(planes + ((item + g) * t_pad + kTcPad + t) * 16);
    const uint8_t* low = planes + ((item + groups + groups / 2 + g / 2) * t_pad + kTcPad + t) * 16 + (g % 2) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const __half_raw raw = __nv_cvt_fp8_to_halfraw(low[e], __NV_E4M3);
        x[((size_t)b * channels + g * 8 + e) * t_len + t] =
            __half2float(main[e]) / kF8ScaleMain + __half2float(__half(raw)) / kF8ScaleXLow;
    }
}

// planes -> fp32 (B, C, T): hi + lo (tests / debugging)
__global__ void __launch_bounds__(128) f32_from_planes_kernel(
    const __nv_bfloat16* __restrict__ planes, float* __restrict__ x,
    int channels, int t_len, int t_pad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_len) return;
    const int g = blockIdx.y, b = blockIdx.z;
    const int groups = channels / 8;
    const size_t row_hi = ((size_t)(b * 2) * groups + g) * t_pad + kTcPad + t;
    const size_t row_lo = row_hi + (size_t)groups * t_pad;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        x[((size_t)b * channels + g * 8 + e) * t_len + t] =
            __bfloat162float(planes[row_hi * 8 + e]) + __bfloat162float(planes[row_lo * 8 + e]);
}

// Zero the pad rows [0, kTcPad) and [kTcPad + T, t_pad) of every (b, plane, group)
__global__ void __launch_bounds__(128) zero_plane_pads_kernel(
    __nv_bfloat16* __restrict__ planes, int t_len, int t_pad) {
    uint4* rows = reinterpret_cast<uint4*>(planes) + (size_t)blockIdx.x * t_pad;
    const int tail = t_pad - kTcPad - t_len;
    for (int i = threadIdx.x; i < kTcPad + tail; i += blockDim.x)
        rows[i < kTcPad ? i : t_len + i] = make_uint4(0, 0, 0, 0);
}

// Conv1d weight (C_out, C_in, K) fp32, already folded, -> slabs
//   [n tile][tap][c_in / KB] x  plain:  [plane][KB / 8][N][8]
//                               concat: [KB / 8][hi rows N | lo rows N][8]
__global__ void pack_tc_weight_kernel(
    const float* __restrict__ w, __nv_bfloat16* __restrict__ slabs,
    int c_out, int c_in, int k, int kb_size, int n_tile, int concat) {
    const size_t total = (size_t)c_out * c_in * k;
    const int groups = kb_size / 8, blocks = c_in / kb_size;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t rest = idx;
        const int e = rest % 8; rest /= 8;
        const int col = rest % n_tile; rest /= n_tile;
        const int g = rest % groups; rest /= groups;
        const int kb = rest % blocks; rest /= blocks;
        const int tap = rest % k; rest /= k;
        const int nt = (int)rest;
        const int c = kb * kb_size + g * 8 + e, o = nt * n_tile + col;
        const float value = w[((size_t)o * c_in + c) * k + tap];
        const __nv_bfloat16 hi = __float2bfloat16_rn(value);
        const __nv_bfloat16 lo = __float2bfloat16_rn(value - __bfloat162float(hi));
        const size_t slab = ((size_t)(nt * k + tap) * blocks + kb) * (size_t)(2 * groups * n_tile * 8);
        if (concat) {
            slabs[slab + ((size_t)g * 2 * n_tile + col) * 8 + e] = hi;
            slabs[slab + ((size_t)g * 2 * n_tile + n_tile + col) * 8 + e] = lo;
        } else {
            slabs[slab + ((size_t)g * n_tile + col) * 8 + e] = hi;
            slabs[slab + (size_t)groups * n_tile * 8 + ((size_t)g * n_tile + col) * 8 + e] = lo;
        }
    }
}

// Conv1d weight (C_out, C_in, K) fp32 -> "fp16 + 2 x fp8" slabs (conv1d_tc.cuh):
// [n tile][tap][c_in / KB] x { [KB / 8][N][8] fp16(w s_main) | [KB / 16][N][16] e4m3((w s_main - fp16) s_low) | e4m3(w s) }
__global__ void pack_tc_weight_f8_kernel(
    const float* __restrict__ w, uint8_t* __restrict__ slabs, int c_out, int c_in, int k, int kb_size,
    int n_tile, float scale, float scale_main, float scale_low) {
    const size_t total = (size_t)c_out * c_in * k;
    const int blocks = c_in / kb_size;
    const size_t slab_bytes = (size_t)kb_size * n_tile * 4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t rest = idx;
        const int cl = rest % kb_size; rest /= kb_size;      // channel within the K block
        const int col = rest % n_tile; rest /= n_tile;
        const int kb = rest % blocks; rest /= blocks;
        const int tap = rest % k; rest /= k;
        const int nt = (int)rest;
        const int c = kb * kb_size + cl, o = nt * n_tile + col;
        const float value = w[((size_t)o * c_in + c) * k + tap];
        const float main = value * scale_main;
        const __half high = __float2half_rn(main);
        const float low = main - __half2float(high);
        uint8_t* slab = slabs + ((size_t)(nt * k + tap) * blocks + kb) * slab_bytes;
        reinterpret_cast<__half*>(slab)[((size_t)(cl / 8) * n_tile + col) * 8 + cl % 8] = high;
        uint8_t* low_section = slab + (size_t)kb_size * n_tile * 2;
        uint8_t* high_section = low_section + (size_t)kb_size * n_tile;
        const size_t at = ((size_t)(cl / 16) * n_tile + col) * 16 + cl % 16;
        low_section[at] = (uint8_t)__nv_cvt_float_to_fp8(low * scale_low, __NV_SATFINITE, __NV_E4M3);
        high_section[at] = (uint8_t)__nv_cvt_float_to_fp8(value * scale, __NV_SATFINITE, __NV_E4M3);
    }
}

// ConvTranspose1d weight (C_in, C_out, 2 UP), already folded ->
// [n tile][tap 0..2][c_in / KB][plane][KB / 8][N_TILE][8] bf16 with column
// n = o * UP + q and taps x[i-1], x[i], x[i+1] (conv_transpose1d.cu has the algebra):
//   tap 1: w[c, o, q + UP/2];  tap 0: q < UP/2 ? w[c, o, q + 3 UP/2] : 0;
//   tap 2: q >= UP/2 ? w[c, o, q - UP/2] : 0
__global__ void pack_tc_transpose_weight_kernel(
    const float* __restrict__ w, __nv_bfloat16* __restrict__ slabs,
    int c_in, int c_out, int up, int kb_size, int n_tile) {
    const int n_total = up * c_out;
    const size_t total = (size_t)3 * c_in * n_total;
    const int groups = kb_size / 8, blocks = c_in / kb_size, half = up / 2, k = 2 * up;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t rest = idx;
        const int e = rest % 8; rest /= 8;
        const int col = rest % n_tile; rest /= n_tile;
        const int g = rest % groups; rest /= groups;
        const int kb = rest % blocks; rest /= blocks;
        const int tap = rest % 3; rest /= 3;
        const int nt = (int)rest;
        const int c = kb * kb_size + g * 8 + e;
        const int n = nt * n_tile + col, o = n / up, q = n % up;
        int j = -1;
        if (tap == 1) j = q + half;
        else if (tap == 0 && q < half) j = q + half + up;
        else if (tap == 2 && q >= half) j = q - half;
        const float value = j >= 0 ? w[((size_t)c * c_out + o) * k + j] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(value);
        const __nv_bfloat16 lo = __float2bfloat16_rn(value - __bfloat162float(hi));
        const size_t slab = (size_t)((nt * 3 + tap) * blocks + kb) * 2;
        const size_t plane = (size_t)groups * n_tile * 8;
        const size_t inner = ((size_t)g * n_tile + col) * 8 + e;
        slabs[slab * plane + inner] = hi;
        slabs[(slab + 1) * plane + inner] = lo;
    }
}

long long* g_tc_debug = nullptr;

int sm_count() {
    static int count = 0;
    if (!count) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
    }
    return count;
}

template <int C_IN, int N, int S, int KB, int NW, int AS, int MODE = kConv, int UP = 0, bool CONCAT = false,
          int XS = 2, bool F8 = false>
int launch_variant(const TcConvArgs& a, int n_tiles, cudaStream_t stream) {
    using Cfg = TcConfig<C_IN, N, S, KB, NW, AS, MODE, CONCAT, XS, F8>;
    auto kernel = conv1d_tc_kernel<C_IN, N, S, KB, NW, AS, MODE, UP, CONCAT, XS, F8>;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem),
            "conv1d_tc smem attribute"));
        configured = true;
    }
    const int tiles_per_item =
        MODE == kFrames ? ceil_div(a.frames, kFramesPerTile) : ceil_div(a.t_len, Cfg::kTile);
    const int num_tiles = tiles_per_item * a.batch * n_tiles;
    const int grid = min(num_tiles, sm_count());
    TcConvArgs args = a;
    if (!args.debug) args.debug = g_tc_debug;
    PMN_REQUIRE((size_t)a.c_out * (size_t)(a.out_row > 0 ? a.out_row : a.t_len) < ((size_t)1 << 31) &&
                    (size_t)a.c_out / 4 * tc_padded_length(a.t_len) < ((size_t)1 << 31),
                "conv1d_tc: an item of the output exceeds 2^31 elements");
    PMN_REQUIRE(!a.bias_batch || (MODE == kConv && C_IN == 128 && N == 256),
                "conv1d_tc: a per-item bias is compiled into the 128 -> 512 variant only");
    PMN_REQUIRE(!(a.out_planes && a.out_f8) || (MODE == kConv ? F8 : UP == 8),
                "conv1d_tc: this variant does not write the fp8 operand form");
    const char* name = MODE == kTranspose ? "conv_transpose1d_tc_kernel" : "conv1d_tc_kernel";
    LaunchScope scope(name, stream);
    kernel<<<grid, kThreads, Cfg::kSmem, stream>>>(
        args, tc_padded_length(a.t_len), tiles_per_item, n_tiles, num_tiles);
    return launched(name);
}

}  // namespace

void tc_set_debug_counters(long long* counters) { g_tc_debug = counters; }

// Tile plan of every supported (C_in, C_out): K block, N tile and operand form.
// Must agree with the template arguments in launch_conv1d_tc.
bool tc_conv_plan(int c_in, int c_out, bool frames, TcPlan* plan) {
    struct Entry { int c_in, c_out; bool frames; TcPlan plan; };
    static const Entry table[] = {
        {256, 256, false, {32, 256, false}}, {128, 128, false, {64, 128, false}},
        // HiFi-GAN's input convolution (113 -> 512, k = 7) with its input channels padded to 128
        {128, 512, false, {32, 256, false}},
        {64, 64, false, {64, 64, true}},     {32, 32, false, {32, 32, true}},
        // penn FCNF0++ blocks 1..5 (valid convolutions, k = 32)
        {256, 32, false, {64, 32, true}},    {32, 128, false, {32, 128, false}},
        {128, 256, false, {32, 256, false}}, {256, 512, true, {32, 256, false}},
        // block 0 as a 32-"channel" (tap) 1x1 conv over im2col rows; head 2048 -> 1440
        {32, 256, false, {32, 256, false}},  {2048, 1440, false, {64, 160, false}},
        // penn block 1 folded by 4 in time (pitch.cu): 1024 -> 128, k = 9; also block 3 frame-major
        {1024, 128, false, {64, 128, false}},
        // penn blocks 4 and 5 frame-major (pitch.cu): K = 32 taps x C_in, k = 1
        {4096, 256, false, {32, 256, false}}, {8192, 512, false, {32, 256, false}},
    };
    for (const Entry& entry : table) {
        if (entry.c_in == c_in && entry.c_out == c_out && entry.frames == frames) {
            if (plan) *plan = entry.plan;
            return true;
        }
    }
    return false;
}

// "fp16 + 2 x fp8" variants: the pitch network's folded block 1 and the generator's C = 256 / 128
// residual blocks.  C = 256 takes 128 columns x 256 rows per tile (two tiles per time window): with 256
// columns x 128 rows a weight slab would serve 512 MMA cycles, 64 B per cycle and SM from L2
bool tc_f8_plan(int c_in, int c_out, TcPlan* plan) {
    TcPlan found;
    if (c_in == 1024 && c_out == 128) found = {64, 128, false};
    else if (c_in == 256 && c_out == 256) found = {32, 128, false};
    else if (c_in == 128 && c_out == 128) found = {64, 128, false};
    else return false;
    if (plan) *plan = found;
    return true;
}

bool tc_supported(int c_in, int c_out, int k, int dilation) {
    return tc_conv_plan(c_in, c_out, false, nullptr) && k >= 1 && k <= 32 &&
           (k - 1) * dilation <= 2 * kMaxHalo;
}

namespace {
// PMN_TCW: 0 never, 1 (default) where it is the faster kernel, 2 wherever it applies
int tcw_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* flag = getenv("PMN_TCW");
        mode = flag && flag[0] >= '0' && flag[0] <= '2' ? flag[0] - '0' : 1;
    }
    return mode;
}
}  // namespace

int launch_conv1d_tc(const TcConvArgs& a, cudaStream_t stream) {
    PMN_REQUIRE(a.x_planes && a.w_slabs, "conv1d_tc: null input");
    PMN_REQUIRE(a.out_slope >= 0.f && a.out_slope <= 1.f, "conv1d_tc: the output slope lies in [0, 1]");
    PMN_REQUIRE(!a.f8x2 || (tc_f8_plan(a.c_in, a.c_out, nullptr) && a.frame_length == 0),
                "conv1d_tc: no fp8 form for these channel counts");
    if (a.f8x2) {
        PMN_REQUIRE(a.out || a.out_planes || (a.accum && a.accum_mode), "conv1d_tc: no output");
        PMN_REQUIRE(a.batch > 0 && a.t_len > 0, "conv1d_tc: empty input");
        PMN_REQUIRE(a.k >= 1 && a.k <= 32 && (a.k - 1) * a.dilation <= 2 * kMaxHalo,
                    "conv1d_tc: receptive field too wide");
        PMN_REQUIRE(a.valid || a.k % 2 == 1, "conv1d_tc: same padding needs an odd kernel");
        if (a.c_in == 1024) return launch_variant<1024, 128, 2, 64, 2, 2, kConv, 0, false, 2, true>(a, 1, stream);
        if (a.c_in == 256) return launch_variant<256, 128, 2, 32, 4, 2, kConv, 0, false, 2, true>(a, 2, stream);
        return launch_variant<128, 128, 2, 64, 2, 2, kConv, 0, false, 2, true>(a, 1, stream);
    }
    if (tcw_mode() && tcw_applies(a) && (tcw_mode() == 2 || tcw_preferred(a)))
        return launch_conv1d_tcw(a, a.w_slabs + tc_plain_weight_elements(a.c_out, a.c_in, a.k), stream);
    PMN_REQUIRE(a.out || a.out_planes || (a.accum && a.accum_mode), "conv1d_tc: no output");
    PMN_REQUIRE(a.batch > 0 && a.t_len > 0, "conv1d_tc: empty input");
    PMN_REQUIRE(a.k >= 1 && a.k <= 32 && (a.k - 1) * a.dilation <= 2 * kMaxHalo,
                "conv1d_tc: receptive field too wide");
    PMN_REQUIRE(a.valid || a.k % 2 == 1, "conv1d_tc: same padding needs an odd kernel");
    const bool frames = a.frame_length > 0;
    PMN_REQUIRE(tc_conv_plan(a.c_in, a.c_out, frames, nullptr), "conv1d_tc: unsupported channel counts");
    if (frames) {
        PMN_REQUIRE(a.valid && a.frames > 0 && a.frame_length <= kMaxFrameLength &&
                        a.frame_valid > 0 && a.frame_valid <= kFrameRows && a.batch == 1,
                    "conv1d_tc: bad frame-mode arguments");
        return launch_variant<256, 256, 1, 32, 2, 2, kFrames>(a, 2, stream);
    }
    if (a.c_in == 256 && a.c_out == 256) return launch_variant<256, 256, 1, 32, 3, 2>(a, 1, stream);
    if (a.c_in == 128 && a.c_out == 128) return launch_variant<128, 128, 2, 64, 2, 2>(a, 1, stream);
    if (a.c_in == 128 && a.c_out == 512) return launch_variant<128, 256, 1, 32, 4, 2>(a, 2, stream);
    if (a.c_in == 64 && a.c_out == 64) return launch_variant<64, 64, 2, 64, 4, 2, kConv, 0, true>(a, 1, stream);
    if (a.c_in == 32 && a.c_out == 32) return launch_variant<32, 32, 4, 32, 8, 2, kConv, 0, true>(a, 1, stream);
    if (a.c_in == 256 && a.c_out == 32) return launch_variant<256, 32, 2, 64, 8, 2, kConv, 0, true>(a, 1, stream);
    if (a.c_in == 32 && a.c_out == 128) return launch_variant<32, 128, 2, 32, 4, 2>(a, 1, stream);
    if (a.c_in == 32 && a.c_out == 256) return launch_variant<32, 256, 1, 32, 4, 2>(a, 1, stream);
    if (a.c_in == 2048) return launch_variant<2048, 160, 1, 64, 3, 2>(a, 9, stream);
    if (a.c_in == 1024) return launch_variant<1024, 128, 2, 64, 2, 2>(a, 1, stream);
    // K = 4096 / 8192 with 256 columns: a 128-row tile streams its whole 4 MB weight from L2 in 98 k
    // MMA cycles (43 B per cycle and SM: the L2 limit), so a tile takes 256 rows (two subtiles per
    // weight slab) and a single accumulator stage -- the exposed epilogue is 2 % of such a tile
    if (a.c_in == 4096) return launch_variant<4096, 256, 2, 32, 3, 1>(a, 1, stream);
    if (a.c_in == 8192) return launch_variant<8192, 256, 2, 32, 3, 1>(a, 2, stream);
    return launch_variant<128, 256, 1, 32, 4, 2>(a, 1, stream);
}

namespace {
int transpose_n_tile(int c_in) { return c_in >= 256 ? 256 : c_in; }  // = min(256, UP * c_out)
int transpose_k_block(int c_in) { return c_in >= 256 ? 32 : 64; }
}  // namespace

bool tc_transpose_supported(int c_in, int c_out, int k, int stride) {
    if (k != 2 * stride || 2 * c_out != c_in) return false;
    return (stride == 8 && (c_in == 512 || c_in == 256)) || (stride == 2 && (c_in == 128 || c_in == 64));
}

size_t tc_transpose_weight_elements(int c_in, int c_out, int stride) {
    return (size_t)2 * 3 * c_in * stride * c_out;
}

int launch_conv_transpose1d_tc(const TcConvArgs& args, int stride, cudaStream_t stream) {
    PMN_REQUIRE(args.x_planes && args.w_slabs && args.out, "conv_transpose1d_tc: null pointer");
    PMN_REQUIRE(args.batch > 0 && args.t_len > 0, "conv_transpose1d_tc: empty input");
    PMN_REQUIRE(tc_transpose_supported(args.c_in, args.c_out, 2 * stride, stride),
                "conv_transpose1d_tc: unsupported shape");
    TcConvArgs a = args;
    a.k = 3;
    a.dilation = 1;
    a.valid = false;
    a.residual = nullptr; a.accum = nullptr; a.accum_mode = 0;   // out_planes: planes of lrelu(out, out_slope) or null
    const int n_tiles = stride * a.c_out / transpose_n_tile(a.c_in);
    switch (a.c_in) {
        case 512: return launch_variant<512, 256, 1, 32, 4, 2, kTranspose, 8>(a, n_tiles, stream);
        case 256: return launch_variant<256, 256, 1, 32, 4, 2, kTranspose, 8>(a, n_tiles, stream);
        case 128: return launch_variant<128, 128, 2, 64, 2, 2, kTranspose, 2>(a, n_tiles, stream);
        default: return launch_variant<64, 64, 2, 64, 4, 2, kTranspose, 2>(a, n_tiles, stream);
    }
}

int launch_pack_tc_transpose_weight(
    const float* w, __nv_bfloat16* slabs, int c_in, int c_out, int stride, cudaStream_t stream) {
    PMN_REQUIRE(w && slabs && tc_transpose_supported(c_in, c_out, 2 * stride, stride),
                "pack_tc_transpose_weight: bad argument");
    const size_t total = (size_t)3 * c_in * stride * c_out;
    const int blocks = (int)min((size_t)4096, (total + 255) / 256);
    LaunchScope scope("pack_tc_transpose_weight_kernel", stream);
    pack_tc_transpose_weight_kernel<<<blocks, 256, 0, stream>>>(
        w, slabs, c_in, c_out, stride, transpose_k_block(c_in), transpose_n_tile(c_in));
    return launched("pack_tc_transpose_weight_kernel");
}

int launch_planes_from_f32(
    const float* x, __nv_bfloat16* planes, int batch, int channels, int t_len, float slope,
    cudaStream_t stream, int source_channels, bool f8) {
    PMN_REQUIRE(x && planes && channels % 8 == 0 && batch > 0 && t_len > 0, "planes_from_f32: bad argument");
    PMN_REQUIRE(source_channels >= 0 && source_channels <= channels, "planes_from_f32: bad channel count");
    PMN_REQUIRE(slope >= 0.f && slope <= 1.f, "planes_from_f32: the slope lies in [0, 1]");
    const int t_pad = tc_padded_length(t_len);
    if (f8) {
        PMN_REQUIRE(channels % 16 == 0 && (source_channels == 0 || source_channels == channels),
                    "planes_from_f32: the fp8 form takes whole 16-channel groups");
        dim3 grid(ceil_div(t_pad, 128), channels / 16, batch);
        LaunchScope scope("planes_f8_from_f32_kernel", stream);
        planes_f8_from_f32_kernel<<<grid, 128, 0, stream>>>(
            x, reinterpret_cast<uint4*>(planes), channels, t_len, t_pad, slope);
        return launched("planes_f8_from_f32_kernel");
    }
    dim3 grid(ceil_div(t_pad, 128), channels / 8, batch);
    LaunchScope scope("planes_from_f32_kernel", stream);
    planes_from_f32_kernel<<<grid, 128, 0, stream>>>(
        x, planes, channels, t_len, t_pad, slope, source_channels ? source_channels : channels);
    return launched("planes_from_f32_kernel");
}

int launch_f32_from_planes(
    const __nv_bfloat16* planes, float* x, int batch, int channels, int t_len, cudaStream_t stream, bool f8) {
    PMN_REQUIRE(x && planes && channels % 8 == 0 && batch > 0 && t_len > 0, "f32_from_planes: bad argument");
    dim3 grid(ceil_div(t_len, 128), channels / 8, batch);
    if (f8) {
        PMN_REQUIRE(channels % 16 == 0, "f32_from_planes: the fp8 form takes whole 16-channel groups");
        LaunchScope scope("f32_from_planes_f8_kernel", stream);
        f32_from_planes_f8_kernel<<<grid, 128, 0, stream>>>(
            reinterpret_cast<const uint8_t*>(planes), x, channels, t_len, tc_padded_length(t_len));
        return launched("f32_from_planes_f8_kernel");
    }
    LaunchScope scope("f32_from_planes_kernel", stream);
    f32_from_planes_kernel<<<grid, 128, 0, stream>>>(planes, x, channels, t_len, tc_padded_length(t_len));
    return launched("f32_from_planes_kernel");
}

int launch_zero_plane_pads(
    __nv_bfloat16* planes, int batch, int channels, int t_len, cudaStream_t stream) {
    PMN_REQUIRE(planes && channels % 8 == 0 && batch > 0 && t_len > 0, "zero_plane_pads: bad argument");
    LaunchScope scope("zero_plane_pads_kernel", stream);
    zero_plane_pads_kernel<<<batch * 2 * (channels / 8), 128, 0, stream>>>(
        planes, t_len, tc_padded_length(t_len));
    return launched("zero_plane_pads_kernel");
}

int tc_f8_weight_shift_of(const float* w, size_t numel, cudaStream_t stream, int* shift) {
    PMN_REQUIRE(w && shift && numel > 0, "tc_f8_weight_shift_of: bad argument");
    std::vector<float> host(numel);
    PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "sync"));
    PMN_TRY(check_cuda(cudaMemcpy(host.data(), w, numel * sizeof(float), cudaMemcpyDeviceToHost), "read weight"));
    float largest = 0.f;
    for (float value : host) largest = fmaxf(largest, fabsf(value));
    *shift = tc_f8_weight_shift(largest);
    return PMN_OK;
}

int launch_pack_tc_weight_f8(
    const float* w, void* slabs, int c_out, int c_in, int k, int weight_shift, cudaStream_t stream) {
    TcPlan plan;
    PMN_REQUIRE(w && slabs && tc_f8_plan(c_in, c_out, &plan), "pack_tc_weight_f8: bad argument");
    PMN_REQUIRE(weight_shift >= 0 && weight_shift <= 16, "pack_tc_weight_f8: bad weight scale");
    const size_t total = (size_t)c_out * c_in * k;
    const int blocks = (int)min((size_t)2048, (total + 255) / 256);
    // one factor for the three products: s_m s_wm = s_x s_wl' = s_xl s_w, hence s_wm = s_w s_xl / s_m
    // and, per unit of the scaled weight, s_wl = s_wl' / s_wm = s_m / s_x
    const float scale = (float)(1 << weight_shift);
    LaunchScope scope("pack_tc_weight_f8_kernel", stream);
    pack_tc_weight_f8_kernel<<<blocks, 256, 0, stream>>>(
        w, static_cast<uint8_t*>(slabs), c_out, c_in, k, plan.k_block, plan.n_tile, scale,
        scale * kF8ScaleXLow / kF8ScaleMain, kF8ScaleMain / kF8ScaleX);
    return launched("pack_tc_weight_f8_kernel");
}

int launch_pack_tc_weight(
    const float* w, __nv_bfloat16* slabs, int c_out, int c_in, int k, bool frames, cudaStream_t stream) {
    TcPlan plan;
    PMN_REQUIRE(w && slabs && tc_conv_plan(c_in, c_out, frames, &plan), "pack_tc_weight: bad argument");
    if (!frames && tcw_shape_supported(c_in, c_out, k))
        PMN_TRY(launch_pack_tcw_weight(w, slabs + tc_plain_weight_elements(c_out, c_in, k), c_in, k, stream));
    const size_t total = (size_t)c_out * c_in * k;
    const int blocks = (int)min((size_t)2048, (total + 255) / 256);
    LaunchScope scope("pack_tc_weight_kernel", stream);
    pack_tc_weight_kernel<<<blocks, 256, 0, stream>>>(
        w, slabs, c_out, c_in, k, plan.k_block, plan.n_tile, plan.concat ? 1 : 0);
    return launched("pack_tc_weight_kernel");
}

}  // namespace pmn

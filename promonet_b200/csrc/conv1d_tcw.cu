// Narrow (C = 32 / 64) dilated "same" Conv1d on tcgen05 with the WEIGHTS on the M side.
//
// conv1d_tc.cu puts 128 time steps on the M side of every MMA and the C output channels on the N
// side; an M = 128 MMA never costs less than the ~64 cycles its 4 KB A operand takes to come out
// of shared memory, so with N = 32 / 64 columns the tensor pipe idles three quarters / half of the
// time (ncu: 24 % / 47 % of active cycles at C = 32 / 64, k = 11).  Here the roles are exchanged:
//     D[(q, o), n] = sum_g sum_c W[g G + q][o, c] * x[c, t0 - h d + n + g G d]
// M = 128 rows hold G = 128 / C taps of the weight stacked on top of each other (hi and lo planes
// as two A operands), N = 240 columns are time steps, K = input channels, and the activations'
// K-major planes (conv1d_tc.cuh) are read as the B operand, a tap-group shift again being a 16-byte
// row offset of the descriptor.  Every MMA is M = 128 x N = 240 at the tensor rate whatever C is.
// Row block q of the accumulator holds tap g G + q of every group g, i.e. the partial sum that
// belongs q d columns to the LEFT:
//     y[o, t0 + m] = sum_q D[(q, o), m + q d],   m < 224,
// so the epilogue reads row block q at column offset q d (free in TMEM), stages the G partial tiles
// of a 32-column chunk in shared memory, and after a 128-thread barrier the same four warps add
// them with TIME on the lanes -- which is the mapping every global access wants (fp32 rows and the
// 16-byte plane rows are contiguous along time).  bf16 x 3 products and fp32 accumulation as in
// conv1d_tc.cu; the results agree to fp32 rounding (the order of the tap sum differs).
//
// Persistent, warp-specialised CTA: warp 0 bulk-copy producer, warp 1 MMA issuer, warps 2-17
// epilogue (four sets of four warps taking the chunks in turn: with two sets the epilogue, a chain of
// TMEM load -> shared store -> barrier -> shared load -> global store per chunk, took 2.5 x the MMAs
// of a C = 32 tile), accumulators double-buffered in TMEM.
#include "conv1d_tc.cuh"
#include "tc_ptx.cuh"

namespace pmn {

namespace {

using namespace tc;

constexpr int kMaxSets = 4;                      // epilogue warp sets, at most
constexpr int kOut = 224;                        // output time steps per tile
constexpr int kChunk = 32;                       // columns per epilogue chunk
constexpr int kChunks = kOut / kChunk;           // 7
constexpr int kShiftMax = 16;                    // >= (G - 1) dilation
constexpr int kColumns = kOut + kShiftMax;       // N of every MMA (240: a multiple of 16)
constexpr int kKB = 32;                          // input channels per activation slab
constexpr int kRowsMax = kColumns + 56;          // + (groups - 1) G d <= 50
constexpr int kXSlab = 2 * (kKB / 8) * kRowsMax * 16;
constexpr int kWSlab = 2 * (kKB / 8) * 128 * 16;     // one tap group x one K block, both planes
constexpr int kXStages = 2, kWStages = 4, kAccStages = 2;
constexpr int kAccStride = 256;                  // TMEM columns between accumulator stages
constexpr int kPitch = kChunk + 4;               // staging row pitch (floats): 16-byte rows, conflict-free
constexpr int kStage = 128 * kPitch;             // floats of one staged chunk (kMaxSets of them)
constexpr int kBarriers = 2 * kXStages + 2 * kWStages + 2 * kAccStages;
constexpr int kSmem = kXStages * kXSlab + kWStages * kWSlab + kMaxSets * kStage * 4 + kBarriers * 8 + 16 + 64 * 4 + 128;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

// SETS epilogue warp sets of four warps; with two sets each has two staging buffers (one barrier per
// chunk), with four sets one (a second barrier before the buffer is overwritten).  Measured
// (profiles/r2_narrow_layers.txt): the C = 32 epilogue (four partial tiles per output) wants four
// sets, the C = 64 one (two partials, 16 channels per thread) spills and slows down with them.
template <int C, int SETS>
__global__ void __launch_bounds__(64 + SETS * 128, 1) conv1d_tcw_kernel(
    TcConvArgs a, const __nv_bfloat16* __restrict__ w_slabs, int t_pad, int tiles_per_item, int num_tiles) {
    constexpr int kSets = SETS;
    constexpr int kBuffers = kMaxSets / SETS;    // staging buffers per set
    constexpr int G = 128 / C;                   // taps stacked on the M side
    constexpr int kBlocks = C / kKB;
    constexpr int kPerThread = C / 4;            // channels a thread combines per chunk
    extern __shared__ uint8_t smem_raw[];
    // 128-byte alignment by pointer arithmetic on the __shared__ array: through an integer cast the
    // compiler loses the address space and emits generic LD / ST for every access to the buffers
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint8_t* x_slabs = smem;
    uint8_t* w_stage = x_slabs + kXStages * kXSlab;
    float* staging = reinterpret_cast<float*>(w_stage + kWStages * kWSlab);   // [set][buffer][128][kPitch]
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kMaxSets * kStage);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + kXStages;
    uint64_t* w_full = x_empty + kXStages;
    uint64_t* w_empty = w_full + kWStages;
    uint64_t* acc_full = w_empty + kWStages;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);
    float* bias_smem = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int groups = (a.k + G - 1) / G;                      // tap groups
    const int rows = kColumns + (groups - 1) * G * a.dilation; // staged window
    const int left = (a.k - 1) / 2 * a.dilation;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kXStages; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4 * kSets); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) bias_smem[threadIdx.x] = (a.bias && threadIdx.x < C) ? a.bias[threadIdx.x] : 0.f;
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
        // ===== producer =====
        if (lane == 0) {
            uint32_t xcount = 0, wcount = 0;
            const uint32_t x_bytes = 2 * (kKB / 8) * rows * 16;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int b = tile / tiles_per_item;
                const int t0 = (tile % tiles_per_item) * kOut;
                for (int kb = 0; kb < kBlocks; ++kb) {
                    const uint32_t xs = xcount % kXStages, xphase = (xcount / kXStages) & 1;
                    ++xcount;
                    mbar_wait(x_empty + xs, xphase ^ 1);
                    mbar_expect_tx(x_full + xs, x_bytes);
                    uint8_t* dst = x_slabs + xs * kXSlab;
#pragma unroll 1
                    for (int p = 0; p < 2; ++p) {
#pragma unroll 1
                        for (int g = 0; g < kKB / 8; ++g) {
                            const size_t row0 =
                                ((size_t)(b * 2 + p) * (C / 8) + kb * (kKB / 8) + g) * t_pad + kTcPad + t0 - left;
                            bulk_copy(dst + (p * (kKB / 8) + g) * rows * 16, a.x_planes + row0 * 8,
                                      rows * 16, x_full + xs);
                        }
                    }
                    for (int group = 0; group < groups; ++group) {
                        const uint32_t ws = wcount % kWStages, wphase = (wcount / kWStages) & 1;
                        ++wcount;
                        mbar_wait(w_empty + ws, wphase ^ 1);
                        mbar_expect_tx(w_full + ws, kWSlab);
                        bulk_copy(w_stage + ws * kWSlab,
                                  reinterpret_cast<const uint8_t*>(w_slabs) + (size_t)(kb * groups + group) * kWSlab,
                                  kWSlab, w_full + ws);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(128, kColumns);
            uint32_t xcount = 0, wcount = 0, tcount = 0;
            const uint32_t x_plane = (kKB / 8) * rows * 16;       // bytes between the hi and lo planes
            constexpr uint32_t w_plane = (kKB / 8) * 128 * 16;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const uint32_t as = tcount % kAccStages, aphase = (tcount / kAccStages) & 1;
                ++tcount;
                mbar_wait(acc_empty + as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + as * kAccStride;
                for (int kb = 0; kb < kBlocks; ++kb) {
                    const uint32_t xs = xcount % kXStages, xphase = (xcount / kXStages) & 1;
                    ++xcount;
                    mbar_wait(x_full + xs, xphase);
                    const uint32_t x_addr = smem_u32(x_slabs + xs * kXSlab);
                    for (int group = 0; group < groups; ++group) {
                        const uint32_t ws = wcount % kWStages, wphase = (wcount / kWStages) & 1;
                        ++wcount;
                        mbar_wait(w_full + ws, wphase);
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(w_stage + ws * kWSlab);
                        const uint32_t row = group * G * a.dilation;   // window row of this group's column 0
#pragma unroll
                        for (int kk = 0; kk < kKB / 16; ++kk) {
                            const uint32_t xb = x_addr + (2 * kk * rows + row) * 16;
                            const uint64_t b_hi = smem_desc(xb, rows * 16, 128);
                            const uint64_t b_lo = smem_desc(xb + x_plane, rows * 16, 128);
                            const uint32_t wa = w_addr + 2 * kk * 128 * 16;
                            const uint64_t a_hi = smem_desc(wa, 128 * 16, 128);
                            const uint64_t a_lo = smem_desc(wa + w_plane, 128 * 16, 128);
                            tc_mma(d, a_hi, b_hi, idesc, !(kb == 0 && group == 0 && kk == 0));
                            tc_mma(d, a_lo, b_hi, idesc, 1);
                            tc_mma(d, a_hi, b_lo, idesc, 1);
                        }
                        tc_commit(w_empty + ws);
                    }
                    tc_commit(x_empty + xs);
                }
                tc_commit(acc_full + as);
            }
        }
    } else {
        // ===== epilogue: kSets sets of four warps, chunk c on set c % kSets =====
        const int quad = warp & 3;                       // TMEM lanes 32 quad .. + 31
        const int set = (warp - 2) >> 2;
        const int member = (warp - 2) & 3;               // which quarter of the channels this warp combines
        const int block = quad * 32 / C;                 // row block q of this warp's TMEM lanes
        const int my_row = quad * 32 + lane;             // accumulator row (q, o) this thread stages
        const int shift = block * a.dilation;
        const int groups_out = C / 8;
        const int out_pad = tc_padded_length_device(a.t_len);
        float* const set_staging = staging + set * kBuffers * kStage;
        int buffer = 0;
        const int c_first = member * kPerThread;
        const bool accumulate = a.accum_mode == 2;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_item;
            const int t0 = (tile % tiles_per_item) * kOut;
            const uint32_t as = tcount % kAccStages, aphase = (tcount / kAccStages) & 1;
            ++tcount;
            // side inputs of a chunk do not depend on the accumulators: they are fetched one chunk ahead
            float res[kPerThread], acc[kPerThread];
            auto fetch = [&](int chunk) {
                const int t = t0 + chunk * kChunk + lane;
                const bool valid = t < a.t_len;
                // one 64-bit address per array, then pointer increments: computing every element's
                // address from scratch took several integer instructions per load
                const size_t idx = ((size_t)b * C + c_first) * a.t_len + t;
                if (valid && a.residual != nullptr) {
                    const float* from = a.residual + idx;
#pragma unroll
                    for (int i = 0; i < kPerThread; ++i, from += a.t_len) res[i] = *from;
                } else {
#pragma unroll
                    for (int i = 0; i < kPerThread; ++i) res[i] = 0.f;
                }
                if (accumulate) {
                    if (valid) {
                        const float* from = a.accum + idx;
#pragma unroll
                        for (int i = 0; i < kPerThread; ++i, from += a.t_len) acc[i] = *from;
                    } else {
#pragma unroll
                        for (int i = 0; i < kPerThread; ++i) acc[i] = 0.f;
                    }
                }
            };
            if (set < kChunks) fetch(set);
            mbar_wait(acc_full + as, aphase);
            tc_fence_after();
            bool released = false;
            for (int chunk = set; chunk < kChunks; chunk += kSets) {
                float* const stage = set_staging + buffer * kStage;
                if (kBuffers > 1) buffer ^= 1;
                // phase A: this warp's 32 accumulator rows x 32 columns (shifted by q d) -> shared memory
                {
                    const uint32_t address = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccStride +
                                             chunk * kChunk + shift;
                    uint32_t raw[32];
                    tc_load32(address, raw);
                    uint4* target = reinterpret_cast<uint4*>(stage + my_row * kPitch);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        target[i] = make_uint4(raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]);
                }
                if (chunk + kSets >= kChunks) {
                    // last read of this tile's accumulators by this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + as);
                    released = true;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
                // phase B: time on the lanes; this warp combines kPerThread channels of the chunk
                const int t = t0 + chunk * kChunk + lane;
                float v[kPerThread];
#pragma unroll
                for (int i = 0; i < kPerThread; ++i) {
                    float sum = stage[(c_first + i) * kPitch + lane];
#pragma unroll
                    for (int q = 1; q < G; ++q) sum += stage[(q * C + c_first + i) * kPitch + lane];
                    v[i] = sum + res[i] + bias_smem[c_first + i];
                }
                // the staged chunk has been read: the set may overwrite it (with two buffers the
                // next chunk's barrier orders that)
                if (kBuffers == 1) asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
                float previous[kPerThread];
                if (accumulate) {
#pragma unroll
                    for (int i = 0; i < kPerThread; ++i) previous[i] = acc[i];
                }
                if (chunk + kSets < kChunks) fetch(chunk + kSets);
                if (t < a.t_len) {
                    const size_t idx = ((size_t)b * C + c_first) * a.t_len + t;
                    if (a.out) {
                        float* to = a.out + idx;
#pragma unroll
                        for (int i = 0; i < kPerThread; ++i, to += a.t_len) *to = v[i];
                    }
                    if (a.accum_mode) {
                        float* to = a.accum + idx;
#pragma unroll
                        for (int i = 0; i < kPerThread; ++i, to += a.t_len) {
                            const float total = accumulate ? fmaf(v[i], a.accum_scale, previous[i]) : v[i] * a.accum_scale;
                            *to = total;
                            if (a.planes_from_accum) v[i] = total;   // the planes below are those of the sum
                        }
                    }
                    if (a.out_planes) {
                        uint4* rows = reinterpret_cast<uint4*>(a.out_planes) +
                                      (size_t)(b * 2) * groups_out * out_pad + kTcPad + t;
                        const uint32_t pad = out_pad;
#pragma unroll
                        for (int g = 0; g < kPerThread / 8; ++g) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                split_pair(leaky(v[g * 8 + 2 * e], a.out_slope),
                                           leaky(v[g * 8 + 2 * e + 1], a.out_slope), hi[e], lo[e]);
                            const uint32_t group = c_first / 8 + g;
                            rows[group * pad] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            rows[(groups_out + group) * pad] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
            }
            if (!released) {
                // a set without a chunk in this tile (kSets > kChunks) still owes its arrival
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + as);
            }
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

// Conv1d weight (C, C, K) fp32, already folded -> slabs [C / 32 K blocks][tap groups] x
// [plane][4 k-groups][128 rows (q, o)][8]: row (q, o) of group g is tap g G + q (zero past K - 1)
__global__ void pack_tcw_weight_kernel(
    const float* __restrict__ w, __nv_bfloat16* __restrict__ slabs, int channels, int k) {
    const int stack = 128 / channels;
    const int groups = (k + stack - 1) / stack, blocks = channels / kKB;
    const size_t total = (size_t)blocks * groups * 4 * 128 * 8;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t rest = idx;
        const int e = rest % 8; rest /= 8;
        const int row = rest % 128; rest /= 128;
        const int kg = rest % 4; rest /= 4;
        const int group = rest % groups; rest /= groups;
        const int kb = (int)rest;
        const int tap = group * stack + row / channels, o = row % channels;
        const int c = kb * kKB + kg * 8 + e;
        const float value = tap < k ? w[((size_t)o * channels + c) * k + tap] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(value);
        const __nv_bfloat16 lo = __float2bfloat16_rn(value - __bfloat162float(hi));
        const size_t slab = (size_t)(kb * groups + group) * (kWSlab / 2);
        const size_t inner = ((size_t)kg * 128 + row) * 8 + e;
        slabs[slab + inner] = hi;
        slabs[slab + (size_t)4 * 128 * 8 + inner] = lo;
    }
}

int tcw_sm_count() {
    static int count = 0;
    if (!count) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
    }
    return count;
}

template <int C, int SETS>
int launch_tcw(const TcConvArgs& a, const __nv_bfloat16* slabs, cudaStream_t stream) {
    auto kernel = conv1d_tcw_kernel<C, SETS>;
    constexpr int kThreads = 64 + SETS * 128;
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem),
            "conv1d_tcw smem attribute"));
        configured = true;
    }
    const int tiles_per_item = ceil_div(a.t_len, kOut);
    const int num_tiles = tiles_per_item * a.batch;
    const int grid = min(num_tiles, tcw_sm_count());
    LaunchScope scope("conv1d_tcw_kernel", stream);
    kernel<<<grid, kThreads, kSmem, stream>>>(a, slabs, tc_padded_length(a.t_len), tiles_per_item, num_tiles);
    return launched("conv1d_tcw_kernel");
}

}  // namespace

bool tcw_shape_supported(int c_in, int c_out, int k) {
    return c_in == c_out && (c_in == 32 || c_in == 64) && k % 2 == 1 && k >= 1 && k <= 11;
}

size_t tcw_weight_elements(int channels, int k) {
    const int stack = 128 / channels;
    return (size_t)(channels / kKB) * ((k + stack - 1) / stack) * (kWSlab / 2);
}

// Where the kernel is the faster of the two on a B200 (profiles/r2_narrow_layers.txt, batch 32 at the
// benchmark's lengths; second table: after conv1d_tc_kernel's epilogue lost half of its instructions):
// the long kernels, where conv1d_tc_kernel pays most for its idle tensor pipe (k = 11: 1.12 - 1.56 x),
// and the planes-only launches of the middle one (k = 7: 1.18 - 1.20 x); with a residual and an fp32
// output the time-on-M kernel's epilogue is now the cheaper one up to k = 7 (0.81 - 0.96), and at
// k = 3 it wins everywhere (0.68 - 0.94)
bool tcw_preferred(const TcConvArgs& a) {
    const bool planes_only = a.out == nullptr && a.accum_mode == 0 && a.residual == nullptr;
    if (a.c_in == 32 || a.c_in == 64) return a.k >= 9 || (a.k >= 5 && planes_only);
    return false;
}

bool tcw_applies(const TcConvArgs& a) {
    if (!tcw_shape_supported(a.c_in, a.c_out, a.k)) return false;
    const int stack = 128 / a.c_in;
    return !a.valid && !a.relu && !a.pool && !a.f8x2 && a.bias_batch == nullptr && a.frame_length == 0 &&
           a.item_groups == 0 && a.plane_groups == 0 &&
           a.out_row == 0 && a.debug == nullptr && a.dilation >= 1 && (stack - 1) * a.dilation <= kShiftMax &&
           (a.k - 1) / 2 * a.dilation <= kTcPad &&
           (((a.k + stack - 1) / stack) - 1) * stack * a.dilation <= kRowsMax - kColumns;
}

int launch_conv1d_tcw(const TcConvArgs& a, const __nv_bfloat16* slabs, cudaStream_t stream) {
    PMN_REQUIRE(a.x_planes && slabs, "conv1d_tcw: null input");
    PMN_REQUIRE(a.out || a.out_planes || (a.accum && a.accum_mode), "conv1d_tcw: no output");
    PMN_REQUIRE(a.batch > 0 && a.t_len > 0, "conv1d_tcw: empty input");
    PMN_REQUIRE(tcw_applies(a), "conv1d_tcw: unsupported arguments");
    return a.c_in == 32 ? launch_tcw<32, 4>(a, slabs, stream) : launch_tcw<64, 2>(a, slabs, stream);
}

int launch_pack_tcw_weight(const float* w, __nv_bfloat16* slabs, int channels, int k, cudaStream_t stream) {
    PMN_REQUIRE(w && slabs && tcw_shape_supported(channels, channels, k), "pack_tcw_weight: bad argument");
    const size_t total = tcw_weight_elements(channels, k) / 2;
    const int blocks = (int)min((size_t)1024, (total + 255) / 256);
    LaunchScope scope("pack_tcw_weight_kernel", stream);
    pack_tcw_weight_kernel<<<blocks, 256, 0, stream>>>(w, slabs, channels, k);
    return launched("pack_tcw_weight_kernel");
}

}  // namespace pmn

// FARGAN generator (config/fargan.py): promonet/model/fargan.py, rows F1/F2.
//
// The vocoder is strictly sequential: 4 subframes per frame, each consuming the
// 512 samples before it (pitch lookback) and three GRU states.  Per subframe the
// network is ~11 dependent matrix-vector layers (2.25 M MAC per utterance), so
// the cost is latency, not FLOPs.  Design:
//
//  * the conditioning MLP (fargan.py:139-160) has no recurrence: it runs for all
//    frames at once as three k=1 convolutions (conv1d.cu) before the loop;
//  * the recurrent part is ONE persistent cooperative kernel.  Utterances are
//    independent, so the batch is cut into groups of 16 and each group gets 64
//    CTAs that never talk to another group.  Every CTA keeps its slice of every
//    layer's weights (4 output units per layer, 140 KB) resident in shared
//    memory for the whole utterance; activations (16 utterances x <= 1152
//    values) are exchanged through L2 in a [k][utterance] layout.  There is no
//    barrier between the 11 layers of a subframe: every exchanged value is its
//    own flag.  A buffer holds a NaN sentinel until its producer stores the
//    value, consumers spin on the loads that fetch their operands anyway (one
//    L2 round trip per layer instead of an atomic, a polled counter and the
//    operand loads: 48.7 -> 25.6 us per subframe on 32 x 5 s).  Buffers rotate
//    over three slots by subframe and every CTA re-arms its slice of the slot
//    of subframe n + 1 during subframe n, after it has seen the first layer of
//    subframe n complete everywhere (so nobody still reads that slot) and a
//    whole subframe before anybody polls it again;
//  * the subframe input (conditioning slice, previous subframe, pitch lookback)
//    is rebuilt by every CTA from the sample history, so the "previous input"
//    state of FramewiseConv (fargan.py:349-364) never leaves shared memory.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <new>
#include <vector>

#include "fargan.cuh"
#include "features.cuh"
#include "tensor_store.cuh"

namespace pmn {

namespace {

constexpr int kHop = 256;
constexpr int kSub = 64;            // FARGAN_SUBFRAME_SIZE
constexpr int kSubframes = 4;       // FARGAN_SUBFRAMES
constexpr int kHistory = 512;       // NUM_PREVIOUS_SAMPLES
constexpr int kLookback = kSub + 4;
constexpr int kInput = 2 * kSub + kSub + kLookback;  // 260: cond slice, previous, lookback
constexpr int kFeatures = 113, kGlobal = 258, kCondIn = kFeatures + kGlobal;  // 371
constexpr int kCond = 2 * kHop;     // 512
constexpr int kGroupItems = 16;     // utterances per CTA group
constexpr int kGroupCtas = 64;
constexpr int kUnits = kHop / kGroupCtas;  // 4 output units per CTA per layer
constexpr int kThreads = 256;            // 512 measured the same (profiles/r2_fargan_breakdown.txt); the side
                                      // sums below want the registers
constexpr int kParts = kThreads / kGroupItems;  // 16 K-partitions

// Per-CTA weight image (floats), every block stored [k][rows]
constexpr int kWFw = 0;                                   // 520 x 4
constexpr int kWFwGlu = kWFw + 2 * kInput * kUnits;       // 256 x 4
constexpr int kWGru = kWFwGlu + kHop * kUnits;            // 3 x (384 x 12 + 256 x 12)
constexpr int kGruIh = (kHop + 2 * kSub) * 3 * kUnits;    // 4608
constexpr int kGruHh = kHop * 3 * kUnits;                 // 3072
constexpr int kWGlu = kWGru + 3 * (kGruIh + kGruHh);      // 3 x 256 x 4
constexpr int kWSkip = kWGlu + 3 * kHop * kUnits;         // 1152 x 4
constexpr int kWSkipGlu = kWSkip + (4 * kHop + 2 * kSub) * kUnits;
constexpr int kWOut = kWSkipGlu + kHop * kUnits;          // 256 x 1
constexpr int kWeights = kWOut + kHop;                    // 35104
static_assert(kWFwGlu % 4 == 0 && kWGru % 4 == 0 && kGruIh % 4 == 0 && kGruHh % 4 == 0 && kWGlu % 4 == 0 &&
              kWSkip % 4 == 0 && kWSkipGlu % 4 == 0 && kWOut % 4 == 0 && (kInput * kUnits) % 4 == 0 &&
              (kSub * kUnits) % 4 == 0, "weight blocks are read with 16-byte loads");

constexpr int kStage = 2 * kHop * kGroupItems;            // staged activations: 512 x 16 floats
constexpr int kScratch = (kParts / 2) * 24 * kGroupItems; // K-partition partial sums
constexpr int kSmemFloats = kWeights + 2 * kInput * kGroupItems + kStage + kScratch;
constexpr int kSmemBytes = kSmemFloats * 4 + 64;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// Activations of one group in global memory, all [k][16]; three slots each (slot n % 3 is
// written and read in subframe n, h also read in n + 1)
constexpr int kSlots = 3;
struct GroupState {
    float* fw[kSlots];        // tanh(fwconv)            256
    float* fwg[kSlots];       // after GLU               256
    float* h[3][kSlots];      // GRU states
    float* g[3][kSlots];      // GRU GLU outputs         256 each
    float* skip[kSlots];      // tanh(skip dense)        256
    float* skipg[kSlots];     // after GLU               256
    float* history;           // (512 + T) samples; sentinel until produced
    unsigned int* barrier;    // start-up only
};
constexpr int kExchanged = 2 + 3 + 3 + 2;         // buffers per slot
constexpr unsigned int kSentinel = 0x7fc0dead;    // a quiet NaN no computation produces

__device__ __forceinline__ float sentinel() { return __uint_as_float(kSentinel); }
__device__ __forceinline__ bool pending(float v) { return __float_as_uint(v) == kSentinel; }

// L2 load that the compiler may not hoist out of a polling loop
__device__ __forceinline__ float4 load_l2(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float load_l2(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// All 64 CTAs of a group arrive; generation-free monotonic counter
__device__ __forceinline__ void group_barrier(unsigned int* counter, unsigned int& target) {
    __syncthreads();
    target += kGroupCtas;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// Copy a [256][16] activation from L2 (never L1: other CTAs wrote it) into shared memory, waiting
// for every value to be produced.  A thread's four loads are issued together and only the ones
// that came back with the sentinel are repeated: one L2 round trip when the producers are done
// (a loop that waited on each load in turn cost four).
__device__ __forceinline__ void stage(float* dst, const float* src, int rows) {
    constexpr int kPer = kHop * kGroupItems / 4 / kThreads;   // 4
    const float4* s = reinterpret_cast<const float4*>(src) + threadIdx.x;
    float4* d = reinterpret_cast<float4*>(dst) + threadIdx.x;
    float4 v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) v[j] = load_l2(s + j * kThreads);
    bool again;
    do {
        again = false;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            if (pending(v[j].x) || pending(v[j].y) || pending(v[j].z) || pending(v[j].w)) {
                v[j] = load_l2(s + j * kThreads);
                again = true;
            }
        }
    } while (again);
#pragma unroll
    for (int j = 0; j < kPer; ++j) d[j * kThreads] = v[j];
}
// acc[r] += sum_k w[k][r] * x[k][b] over this thread's K partition.  The ROWS weights of a k are
// contiguous and 16-byte aligned (every block of the image starts at a multiple of 4 floats):
// they are read as ROWS / 4 vector loads -- with scalar loads the dot products were bound by the
// issue of shared-memory loads (one per multiply-add: 480 per thread and GRU cell)
template <int ROWS>
__device__ __forceinline__ void dot(
    float (&acc)[ROWS], const float* __restrict__ w, const float* __restrict__ x, int k_count,
    int part, int b) {
    if constexpr (ROWS % 4 == 0) {
#pragma unroll 4
        for (int k = part; k < k_count; k += kParts) {
            const float a = x[k * kGroupItems + b];
            const float4* row = reinterpret_cast<const float4*>(w + k * ROWS);
#pragma unroll
            for (int q = 0; q < ROWS / 4; ++q) {
                const float4 v = row[q];
                acc[4 * q] = fmaf(v.x, a, acc[4 * q]);
                acc[4 * q + 1] = fmaf(v.y, a, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v.z, a, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v.w, a, acc[4 * q + 3]);
            }
        }
    } else {
        for (int k = part; k < k_count; k += kParts) {
            const float a = x[k * kGroupItems + b];
            const float* row = w + k * ROWS;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(row[r], a, acc[r]);
        }
    }
}

// The same with the activation read straight from L2 (values a subframe old: complete, no polling);
// all of a thread's loads are independent, so the product costs one L2 latency -- spent in the
// shadow of an exchange the CTA is waiting for anyway
template <int ROWS>
__device__ __forceinline__ void dot_global(
    float (&acc)[ROWS], const float* __restrict__ w, const float* __restrict__ x, int part, int b) {
    constexpr int kSteps = kHop / kParts;
    float a[kSteps];
#pragma unroll
    for (int i = 0; i < kSteps; ++i) a[i] = __ldcg(x + (part + i * kParts) * kGroupItems + b);
#pragma unroll
    for (int i = 0; i < kSteps; ++i) {
        const float4* row = reinterpret_cast<const float4*>(w + (part + i * kParts) * ROWS);
#pragma unroll
        for (int q = 0; q < ROWS / 4; ++q) {
            const float4 v = row[q];
            acc[4 * q] = fmaf(v.x, a[i], acc[4 * q]);
            acc[4 * q + 1] = fmaf(v.y, a[i], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v.z, a[i], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v.w, a[i], acc[4 * q + 3]);
        }
    }
}

// Sum the K partitions: results land in scratch[r][b] (r < ROWS), valid after the sync
template <int ROWS>
__device__ __forceinline__ void reduce(float (&acc)[ROWS], float* scratch, int part, int b) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 16);
    __syncthreads();  // previous users of scratch are done
    if ((part & 1) == 0) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) scratch[((part >> 1) * ROWS + r) * kGroupItems + b] = acc[r];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ROWS * kGroupItems; idx += kThreads) {
        float sum = 0.f;
#pragma unroll
        for (int p = 0; p < kParts / 2; ++p) sum += scratch[p * ROWS * kGroupItems + idx];
        scratch[idx] = sum;  // idx < ROWS * 16 <= the p = 0 slab: each thread rewrites its own slot
    }
    __syncthreads();
}

// profiling aid (PMN_FARGAN_DEBUG=1): cycles thread 0 of CTA 0 spends waiting for operands
// (subframe input, every stage call) against the whole frame loop
__device__ long long g_fargan_cycles[8];

__global__ void __launch_bounds__(kThreads, 1) fargan_kernel(
    const float* __restrict__ weights,   // (64, kWeights) per-CTA images
    const float* __restrict__ cond,      // (B, 512, F) tanh conditioning
    const float* __restrict__ features,  // (B, 114, F): row 113 = pitch period
    const float* __restrict__ previous,  // (B, 512) or null (zeros)
    GroupState* __restrict__ groups, float* __restrict__ audio, int batch, int frames, int debug) {
    extern __shared__ __align__(16) float smem[];
    float* w = smem;
    // the two feature buffers, addressed by offset from the __shared__ array: taken from a pointer
    // array the compiler no longer knows their address space and reads them with generic loads
    float* const input0 = smem + kWeights;
    float* staged = smem + kWeights + 2 * kInput * kGroupItems;
    float* scratch = staged + kStage;

    const int tid = threadIdx.x;
    const int cta = blockIdx.x % kGroupCtas;
    const int group = blockIdx.x / kGroupCtas;
    const int b = tid & (kGroupItems - 1);
    const int part = tid >> 4;
    const GroupState gs = groups[group];
    // The 10 x 3 exchange buffers lie back to back in GroupState's member order (fargan_forward fills
    // them so): buffer (member, slot) by arithmetic.  Indexing the pointer arrays with a run-time slot
    // put the struct into local memory and a local load in front of every exchange access.
    float* const exchange = gs.fw[0];
    __builtin_assume(__isGlobal(exchange));      // pointers read from memory are generic otherwise
    __builtin_assume(__isGlobal(gs.history));
    auto buffer_of = [&](int member, int slot_index) {
        return exchange + (size_t)(member * kSlots + slot_index) * (kHop * kGroupItems);
    };
    enum { kFw = 0, kFwg = 1, kH = 2, kG = 5, kSkip = 8, kSkipg = 9 };
    const int samples = frames * kHop;
    unsigned int target = 0;

    for (int i = tid; i < kWeights; i += kThreads) w[i] = weights[(size_t)cta * kWeights + i];
    for (int i = tid; i < 2 * kInput * kGroupItems; i += kThreads) input0[i] = 0.f;  // state3 = 0
    // Start-up: every exchanged slot holds the sentinel, except the GRU states "of subframe -1"
    // (slot 2), which start at zero (fargan.py:406-415); history[0:512] = previous samples, the rest
    // sentinel.  The 64 CTAs share the fill and meet once at a barrier.
    {
        for (int buffer = 0; buffer < kExchanged * kSlots; ++buffer) {
            // h[s][2] are buffers 2 * kSlots + s * kSlots + 2
            const bool zero = buffer >= 2 * kSlots && buffer < 5 * kSlots && (buffer - 2 * kSlots) % kSlots == 2;
            for (int i = cta * kThreads + tid; i < kHop * kGroupItems; i += kGroupCtas * kThreads)
                exchange[(size_t)buffer * (kHop * kGroupItems) + i] = zero ? 0.f : sentinel();
        }
        const size_t total = (size_t)(kHistory + samples) * kGroupItems;
        for (size_t i = (size_t)cta * kThreads + tid; i < total; i += (size_t)kGroupCtas * kThreads) {
            float v = sentinel();
            if (i < (size_t)kHistory * kGroupItems) {
                const int k = (int)(i / kGroupItems), col = (int)(i % kGroupItems);
                const int it = group * kGroupItems + col;
                v = (previous && it < batch) ? previous[(size_t)it * kHistory + k] : 0.f;
                if (pending(v)) v = __uint_as_float(0x7fc00000u);   // a caller's NaN must not look unproduced
            }
            gs.history[i] = v;
        }
    }
    group_barrier(gs.barrier, target);

    int current = 0;  // input[current] = this subframe's features, input[current ^ 1] = previous
    const bool timing = debug && blockIdx.x == 0 && tid == 0;
    long long wait_input = 0, wait_stage = 0, t_dot = 0, t_reduce = 0, t_input = 0, mark = 0, mark2 = 0;
    auto tick2 = [&]() { if (timing) mark2 = clock64(); };
    auto tock2 = [&](long long& sum) { if (timing) sum += clock64() - mark2; };
    const long long loop_start = timing ? clock64() : 0;
    auto tick = [&]() { if (timing) mark = clock64(); };
    auto tock = [&](long long& sum) { if (timing) sum += clock64() - mark; };
    for (int n = 0; n < frames * kSubframes; ++n) {
        const int f = n / kSubframes, sub = n % kSubframes;
        const int slot = n % kSlots, before = (n + kSlots - 1) % kSlots, after = (n + 1) % kSlots;
        float* in = input0 + current * (kInput * kGroupItems);
        const float* state3 = input0 + (current ^ 1) * (kInput * kGroupItems);
        const float* hist = gs.history + (size_t)n * kSub * kGroupItems;  // 512-sample window

        // ---- subframe input: cond[:, sub::4] (fargan.py:109-113), previous 64, lookback 68 ----
        tick2();
        {
            // a thread's history samples are loaded together; the ones still pending (the previous
            // subframe's, just being produced) are loaded again until they arrive
            constexpr int kMine = (kInput * kGroupItems + kThreads - 1) / kThreads;   // 17
            const float* address[kMine];
            float value[kMine];
#pragma unroll
            for (int j = 0; j < kMine; ++j) {
                const int idx = tid + j * kThreads;
                address[j] = nullptr;
                value[j] = 0.f;
                if (idx >= kInput * kGroupItems) continue;
                const int k = idx / kGroupItems, col = idx % kGroupItems;
                const int it = group * kGroupItems + col;
                if (it >= batch) continue;
                if (k < 2 * kSub) {
                    value[j] = __ldg(cond + ((size_t)it * kCond + 4 * k + sub) * frames + f);
                } else if (k < 3 * kSub) {
                    address[j] = hist + (size_t)(kHistory - kSub + (k - 2 * kSub)) * kGroupItems + col;
                } else {
                    const int period = (int)rintf(__ldg(features + ((size_t)it * 114 + 113) * frames + f));
                    int index = kHistory - period + (k - 3 * kSub) - 2;   // fargan.py:233-239
                    if (index >= kHistory) index -= period;
                    index = max(index, 0);
                    address[j] = hist + (size_t)index * kGroupItems + col;
                }
                if (address[j]) value[j] = load_l2(address[j]);
            }
            bool again;
            tick();
            do {
                again = false;
#pragma unroll
                for (int j = 0; j < kMine; ++j) {
                    if (address[j] && pending(value[j])) {
                        value[j] = load_l2(address[j]);
                        again = true;
                    }
                }
            } while (again);
            tock(wait_input);
#pragma unroll
            for (int j = 0; j < kMine; ++j) {
                const int idx = tid + j * kThreads;
                if (idx < kInput * kGroupItems) in[idx] = value[j];
            }
        }
        __syncthreads();

        tock2(t_input);
        // ---- framewise conv: tanh(W [in; state3]) (fargan.py:349-364) ----
        {
            float acc[kUnits] = {};
            tick2();
            dot<kUnits>(acc, w + kWFw, in, kInput, part, b);
            tock2(t_dot);
            tick2();
            dot<kUnits>(acc, w + kWFw + kInput * kUnits, state3, kInput, part, b);
            tock2(t_dot);
            tick2();
            reduce<kUnits>(acc, scratch, part, b);
            tock2(t_reduce);
            if (tid < kUnits * kGroupItems)
                __stcg(buffer_of(kFw, slot) + (size_t)(cta * kUnits + tid / kGroupItems) * kGroupItems + (tid % kGroupItems),
                       tanhf(scratch[tid]));
        }

        // Side sums: everything of the GRU cells and of the skip layer that does not depend on this
        // subframe's exchanges -- the recurrent halves W_hh h of the three cells (h of the previous
        // subframe, read straight from L2), the lookback / previous-subframe columns of their input
        // halves and of the skip layer -- is computed here, while the other CTAs' framewise outputs
        // travel, instead of on the critical path of the later layers
        const float* lookback = in + (3 * kSub + 2) * kGroupItems;  // pitch_lookback[:, 2:-2]
        const float* last = in + 2 * kSub * kGroupItems;             // previous subframe
        float gh[3][3 * kUnits] = {}, gi_side[3][3 * kUnits] = {}, skip_acc[kUnits] = {}, h_previous[3] = {};
        tick2();
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const float* wih = w + kWGru + s * (kGruIh + kGruHh);
            dot_global<3 * kUnits>(gh[s], wih + kGruIh, buffer_of(kH + s, before), part, b);
            dot<3 * kUnits>(gi_side[s], wih + kHop * 3 * kUnits, lookback, kSub, part, b);
            dot<3 * kUnits>(gi_side[s], wih + (kHop + kSub) * 3 * kUnits, last, kSub, part, b);
            if (tid < kUnits * kGroupItems)
                h_previous[s] = __ldcg(buffer_of(kH + s, before) + (size_t)(cta * kUnits + tid / kGroupItems) * kGroupItems +
                                       (tid % kGroupItems));
        }
        dot<kUnits>(skip_acc, w + kWSkip + 4 * kHop * kUnits, lookback, kSub, part, b);
        dot<kUnits>(skip_acc, w + kWSkip + (4 * kHop + kSub) * kUnits, last, kSub, part, b);
        tock2(t_dot);

        // GLU: x * sigmoid(W x) for this CTA's 4 units (fargan.py:375-388)
        auto glu = [&](const float* wg, const float* x_global, float* out_global) {
            tick();
            stage(staged, x_global, kHop);
            __syncthreads();
            tock(wait_stage);
            float acc[kUnits] = {};
            tick2();
            dot<kUnits>(acc, wg, staged, kHop, part, b);
            tock2(t_dot);
            tick2();
            reduce<kUnits>(acc, scratch, part, b);
            tock2(t_reduce);
            if (tid < kUnits * kGroupItems) {
                const int unit = cta * kUnits + tid / kGroupItems, col = tid % kGroupItems;
                const float x = staged[unit * kGroupItems + col];
                __stcg(out_global + (size_t)unit * kGroupItems + col, x * sigmoid(scratch[tid]));
            }
        };

        glu(w + kWFwGlu, buffer_of(kFw, slot), buffer_of(kFwg, slot));
        // fw of this subframe is complete everywhere (glu staged all of it): every CTA has finished
        // subframe n - 1, so nobody reads the slot of subframe n + 1 any more (last written in
        // n - 2, last read in n - 1): re-arm this CTA's slice of it, a whole subframe ahead of its
        // next readers
        if (tid < kUnits * kGroupItems) {
            const size_t at = (size_t)(cta * kUnits + tid / kGroupItems) * kGroupItems + (tid % kGroupItems);
#pragma unroll
            for (int buffer = 0; buffer < kExchanged; ++buffer) __stcg(buffer_of(buffer, after) + at, sentinel());
            __threadfence();
        }

        // ---- three GRU cells + GLUs (fargan.py:267-309) ----
        // The input of cell s (fwg, g1, g2) is also a column block of the skip layer (blocks 3, 0, 1 of
        // [g1, g2, g3, fw | lookback, previous]): its share of the skip sum is taken while it is staged
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const float* wih = w + kWGru + s * (kGruIh + kGruHh);
            const float* x_global = s == 0 ? buffer_of(kFwg, slot) : buffer_of(kG + s - 1, slot);
            tick();
            stage(staged, x_global, kHop);
            __syncthreads();
            tock(wait_stage);
            tick2();
            dot<3 * kUnits>(gi_side[s], wih, staged, kHop, part, b);
            dot<kUnits>(skip_acc, w + kWSkip + (s == 0 ? 3 : s - 1) * kHop * kUnits, staged, kHop, part, b);
            tock2(t_dot);
            float both[6 * kUnits];
#pragma unroll
            for (int r = 0; r < 3 * kUnits; ++r) { both[r] = gi_side[s][r]; both[3 * kUnits + r] = gh[s][r]; }
            tick2();
            reduce<6 * kUnits>(both, scratch, part, b);
            tock2(t_reduce);
            if (tid < kUnits * kGroupItems) {
                const int u = tid / kGroupItems, col = tid % kGroupItems;
                const int unit = cta * kUnits + u;
                auto at = [&](int row) { return scratch[row * kGroupItems + col]; };
                const float r = sigmoid(at(3 * u) + at(3 * kUnits + 3 * u));
                const float z = sigmoid(at(3 * u + 1) + at(3 * kUnits + 3 * u + 1));
                const float c = tanhf(at(3 * u + 2) + r * at(3 * kUnits + 3 * u + 2));
                __stcg(buffer_of(kH + s, slot) + (size_t)unit * kGroupItems + col, (1.f - z) * c + z * h_previous[s]);
            }
            glu(w + kWGlu + s * kHop * kUnits, buffer_of(kH + s, slot), buffer_of(kG + s, slot));
        }

        // ---- skip: tanh(W [g1, g2, g3, fw, lookback, previous]) then GLU (fargan.py:311-325) ----
        {
            tick();
            stage(staged, buffer_of(kG + 2, slot), kHop);
            __syncthreads();
            tock(wait_stage);
            tick2();
            dot<kUnits>(skip_acc, w + kWSkip + 2 * kHop * kUnits, staged, kHop, part, b);
            tock2(t_dot);
            tick2();
            reduce<kUnits>(skip_acc, scratch, part, b);
            tock2(t_reduce);
            if (tid < kUnits * kGroupItems)
                __stcg(buffer_of(kSkip, slot) + (size_t)(cta * kUnits + tid / kGroupItems) * kGroupItems + (tid % kGroupItems),
                       tanhf(scratch[tid]));
        }
        glu(w + kWSkipGlu, buffer_of(kSkip, slot), buffer_of(kSkipg, slot));

        // ---- output: tanh(W skip), one of the 64 samples per CTA (fargan.py:327-329) ----
        {
            tick();
            stage(staged, buffer_of(kSkipg, slot), kHop);
            __syncthreads();
            tock(wait_stage);
            float acc[1] = {};
            tick2();
            dot<1>(acc, w + kWOut, staged, kHop, part, b);
            tock2(t_dot);
            tick2();
            reduce<1>(acc, scratch, part, b);
            tock2(t_reduce);
            if (tid < kGroupItems) {
                const float y = tanhf(scratch[tid]);
                __stcg(gs.history + (size_t)(kHistory + n * kSub + cta) * kGroupItems + tid, y);
                const int it = group * kGroupItems + tid;
                if (it < batch) audio[(size_t)it * samples + n * kSub + cta] = y;
            }
        }
        __syncthreads();   // `staged`, `scratch` and `in` are rewritten by the next subframe
        current ^= 1;
    }
    if (timing) {
        g_fargan_cycles[0] = clock64() - loop_start;
        g_fargan_cycles[1] = wait_input;
        g_fargan_cycles[2] = wait_stage;
        g_fargan_cycles[3] = frames * kSubframes;
        g_fargan_cycles[4] = t_dot; g_fargan_cycles[5] = t_reduce; g_fargan_cycles[6] = t_input;
    }
}

// x (B, 371, F) = [features[:113]; speaker embedding; ratios broadcast over frames]
__global__ void __launch_bounds__(128) cond_input_kernel(
    const float* __restrict__ features, const float* __restrict__ speaker_embedding,
    const int64_t* __restrict__ speakers, const float* __restrict__ sbr, const float* __restrict__ lr,
    float* __restrict__ out, int frames, int num_speakers) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    if (f >= frames) return;
    float v;
    if (c < kFeatures) {
        v = features[((size_t)b * 114 + c) * frames + f];
    } else if (c < kFeatures + 256) {
        int64_t speaker = speakers[b];
        speaker = speaker < 0 ? 0 : (speaker >= num_speakers ? num_speakers - 1 : speaker);
        v = speaker_embedding[(size_t)speaker * 256 + (c - kFeatures)];
    } else {
        v = c == kFeatures + 256 ? sbr[b] : lr[b];
    }
    out[((size_t)b * kCondIn + c) * frames + f] = v;
}

}  // namespace

}  // namespace pmn

struct pmn_fargan {
    pmn::TensorStore store;
    bool finalized = false;
    float ppg_threshold = 0.85f;
    float* images = nullptr;         // (64, kWeights)
    float* cond_weight[3] = {};      // packed (C_in, 1, C_out)
};

namespace pmn {

namespace {

int find(const pmn_fargan* g, const std::string& name, const Tensor** out) {
    return g->store.find(name, out);
}

// Host copy of `<prefix>.weight`, folding weight norm (Linear: per output row)
int host_weight(pmn_fargan* g, const std::string& prefix, int rows, int cols,
                std::vector<float>* out, cudaStream_t stream) {
    out->resize((size_t)rows * cols);
    const float* source;
    float* folded = nullptr;
    if (g->store.has(prefix + ".weight")) {
        const Tensor* plain;
        PMN_TRY(find(g, prefix + ".weight", &plain));
        if (plain->numel() != out->size()) return fail(PMN_ERR_STATE, "bad shape at " + prefix);
        source = plain->data;
    } else {
        const Tensor *wg, *wv;
        PMN_TRY(find(g, prefix + ".weight_g", &wg));
        PMN_TRY(find(g, prefix + ".weight_v", &wv));
        if (wv->numel() != out->size() || wg->numel() != (size_t)rows)
            return fail(PMN_ERR_STATE, "bad weight_g/weight_v shapes at " + prefix);
        PMN_TRY(check_cuda(cudaMalloc(&folded, out->size() * 4), "cudaMalloc"));
        int status = launch_weight_norm_fold(wv->data, wg->data, folded, rows, cols, stream);
        if (status != PMN_OK) { cudaFree(folded); return status; }
        source = folded;
    }
    int status = check_cuda(cudaStreamSynchronize(stream), "sync");
    if (status == PMN_OK)
        status = check_cuda(cudaMemcpy(out->data(), source, out->size() * 4, cudaMemcpyDeviceToHost), "D2H weight");
    if (folded) cudaFree(folded);
    return status;
}

struct Workspace {
    float *features, *cond_in, *cond_a, *cond_b, *state;
    GroupState* groups;       // device array
    unsigned int* barriers;
    size_t group_floats, bytes;
};

Workspace carve(void* base, int batch, int frames) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += align_up(bytes, 256);
        return r;
    };
    const int groups = ceil_div(batch, kGroupItems);
    w.features = (float*)take((size_t)batch * 114 * frames * 4);
    w.cond_in = (float*)take((size_t)batch * kCondIn * frames * 4);
    w.cond_a = (float*)take((size_t)batch * kCond * frames * 4);
    w.cond_b = (float*)take((size_t)batch * kCond * frames * 4);
    // per group: (fw, fwg, 3 h, 3 g, skip, skipg) x 3 slots x 256 x 16, history (512 + T) x 16
    w.group_floats = (size_t)kExchanged * kSlots * kHop * kGroupItems +
                     (size_t)(kHistory + frames * kHop) * kGroupItems;
    w.state = (float*)take((size_t)groups * w.group_floats * 4);
    w.groups = (GroupState*)take((size_t)groups * sizeof(GroupState));
    w.barriers = (unsigned int*)take((size_t)groups * 128);
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}

}  // namespace

pmn_fargan* fargan_create() { return new (std::nothrow) pmn_fargan(); }
void fargan_destroy(pmn_fargan* g) { delete g; }

int fargan_set_tensor(pmn_fargan* g, const char* name, const float* data, const int64_t* shape,
                      int ndim, cudaStream_t stream) {
    if (g->finalized) return fail(PMN_ERR_STATE, "set_tensor after finalize");
    return g->store.set(name, data, shape, ndim, stream);
}

int fargan_finalize(pmn_fargan* g, cudaStream_t stream) {
    if (g->finalized) return fail(PMN_ERR_STATE, "fargan already finalized");
    const Tensor* t;
    PMN_TRY(find(g, "speaker_embedding.weight", &t));
    PMN_TRY(find(g, "pitch_embedding.weight", &t));
    PMN_TRY(find(g, "pitch_distribution", &t));
    if (g->store.has("ppg_threshold")) {
        PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "sync"));
        PMN_TRY(check_cuda(
            cudaMemcpy(&g->ppg_threshold, g->store.data("ppg_threshold"), 4, cudaMemcpyDeviceToHost),
            "read ppg_threshold"));
    }
    // conditioning MLP as k = 1 convolutions
    const int cond_out[3] = {kCondIn, kCondIn, kCond};
    for (int i = 0; i < 3; ++i) {
        PMN_TRY(find(g, "model.conditioning_network." + std::to_string(2 * i) + ".weight", &t));
        if (t->numel() != (size_t)cond_out[i] * kCondIn) return fail(PMN_ERR_STATE, "bad conditioning shape");
        PMN_TRY(g->store.alloc(t->numel(), &g->cond_weight[i]));
        PMN_TRY(launch_pack_conv1d_weight(t->data, g->cond_weight[i], cond_out[i], kCondIn, 1, stream));
    }
    // per-CTA weight images
    const std::string net = "model.subframe_network.";
    std::vector<float> fw, fwglu, skip, skipglu, out, ih[3], hh[3], glu[3];
    PMN_TRY(host_weight(g, net + "framewise_convolution.model.0", kHop, 2 * kInput, &fw, stream));
    PMN_TRY(host_weight(g, net + "framewise_convolution.model.2.gate", kHop, kHop, &fwglu, stream));
    for (int s = 0; s < 3; ++s) {
        const std::string cell = net + "gru" + std::to_string(s + 1);
        const Tensor *wi, *wh;
        PMN_TRY(find(g, cell + ".weight_ih", &wi));
        PMN_TRY(find(g, cell + ".weight_hh", &wh));
        if (wi->numel() != (size_t)3 * kHop * (kHop + 2 * kSub) || wh->numel() != (size_t)3 * kHop * kHop)
            return fail(PMN_ERR_STATE, "bad GRU shape at " + cell);
        ih[s].resize(wi->numel());
        hh[s].resize(wh->numel());
        PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "sync"));
        PMN_TRY(check_cuda(cudaMemcpy(ih[s].data(), wi->data, wi->numel() * 4, cudaMemcpyDeviceToHost), "D2H"));
        PMN_TRY(check_cuda(cudaMemcpy(hh[s].data(), wh->data, wh->numel() * 4, cudaMemcpyDeviceToHost), "D2H"));
        PMN_TRY(host_weight(g, cell + "_glu.gate", kHop, kHop, &glu[s], stream));
    }
    PMN_TRY(host_weight(g, net + "skip_dense", kHop, 4 * kHop + 2 * kSub, &skip, stream));
    PMN_TRY(host_weight(g, net + "skip_glu.gate", kHop, kHop, &skipglu, stream));
    PMN_TRY(host_weight(g, net + "output_layer", kSub, kHop, &out, stream));

    std::vector<float> images((size_t)kGroupCtas * kWeights);
    for (int cta = 0; cta < kGroupCtas; ++cta) {
        float* image = images.data() + (size_t)cta * kWeights;
        // dense layers: rows cta * 4 .. + 3, stored [k][4]
        auto dense = [&](float* dst, const std::vector<float>& weight, int cols) {
            for (int k = 0; k < cols; ++k)
                for (int u = 0; u < kUnits; ++u)
                    dst[k * kUnits + u] = weight[(size_t)(cta * kUnits + u) * cols + k];
        };
        dense(image + kWFw, fw, 2 * kInput);
        dense(image + kWFwGlu, fwglu, kHop);
        for (int s = 0; s < 3; ++s) {
            float* base = image + kWGru + s * (kGruIh + kGruHh);
            // GRU: rows (gate * 256 + unit), stored [k][unit * 3 + gate]
            auto gates = [&](float* dst, const std::vector<float>& weight, int cols) {
                for (int k = 0; k < cols; ++k)
                    for (int u = 0; u < kUnits; ++u)
                        for (int gate = 0; gate < 3; ++gate)
                            dst[k * 3 * kUnits + 3 * u + gate] =
                                weight[(size_t)(gate * kHop + cta * kUnits + u) * cols + k];
            };
            gates(base, ih[s], kHop + 2 * kSub);
            gates(base + kGruIh, hh[s], kHop);
            dense(image + kWGlu + s * kHop * kUnits, glu[s], kHop);
        }
        dense(image + kWSkip, skip, 4 * kHop + 2 * kSub);
        dense(image + kWSkipGlu, skipglu, kHop);
        for (int k = 0; k < kHop; ++k) image[kWOut + k] = out[(size_t)cta * kHop + k];
    }
    PMN_TRY(g->store.alloc(images.size(), &g->images));
    PMN_TRY(check_cuda(cudaMemcpy(g->images, images.data(), images.size() * 4, cudaMemcpyHostToDevice), "H2D images"));
    g->finalized = true;
    return PMN_OK;
}

size_t fargan_workspace_bytes(int batch, int frames) { return carve(nullptr, batch, frames).bytes; }

int fargan_forward(
    pmn_fargan* g, const float* loudness, int rows, const float* pitch, const float* periodicity,
    const float* ppg, const int64_t* speakers, const float* sbr, const float* lr,
    const float* previous_samples, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!g->finalized) return fail(PMN_ERR_STATE, "fargan not finalized");
    PMN_REQUIRE(audio && workspace && speakers && sbr && lr, "fargan: null pointer");
    PMN_REQUIRE(batch > 0 && frames > 0, "fargan: empty batch");
    int device = 0, sms = 0, cooperative = 0;
    PMN_TRY(check_cuda(cudaGetDevice(&device), "cudaGetDevice"));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, device);
    if (!cooperative) return fail(PMN_ERR_STATE, "fargan: device lacks cooperative launch");
    const int max_items = sms / kGroupCtas * kGroupItems;  // utterances per cooperative launch
    PMN_REQUIRE(max_items > 0, "fargan: fewer than 64 SMs");
    Workspace w = carve(workspace, batch, frames);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "fargan: workspace too small");

    // features (B, 114, F) incl. pitch period; conditioning for every frame
    PMN_TRY(launch_features(
        loudness, rows, pitch, periodicity, ppg, g->store.data("pitch_distribution"),
        g->store.data("pitch_embedding.weight"), g->ppg_threshold, true, w.features,
        batch, frames, stream));
    {
        dim3 grid(ceil_div(frames, 128), kCondIn, batch);
        LaunchScope scope("cond_input_kernel", stream);
        cond_input_kernel<<<grid, 128, 0, stream>>>(
            w.features, g->store.data("speaker_embedding.weight"), speakers, sbr, lr,
            w.cond_in, frames, 109);
        PMN_TRY(launched("cond_input_kernel"));
    }
    const float* x = w.cond_in;
    float* outs[3] = {w.cond_a, w.cond_b, w.cond_a};
    const int cond_out[3] = {kCondIn, kCondIn, kCond};
    for (int i = 0; i < 3; ++i) {
        Conv1dArgs a;
        a.x = x; a.weight = g->cond_weight[i]; a.out = outs[i];
        a.batch = batch; a.c_in = kCondIn; a.c_out = cond_out[i];
        a.t_in = a.t_out = frames; a.k = 1; a.out_act = 1;
        PMN_TRY(launch_conv1d(a, stream));
        x = outs[i];
    }

    // group state pointers
    const int groups = ceil_div(batch, kGroupItems);
    std::vector<GroupState> host(groups);
    for (int i = 0; i < groups; ++i) {
        float* p = w.state + (size_t)i * w.group_floats;
        const size_t unit = (size_t)kHop * kGroupItems;
        GroupState& s = host[i];
        // buffer order = member order of GroupState (the kernel walks &fw[0] as 30 pointers)
        float** all = &s.fw[0];
        for (int buffer = 0; buffer < kExchanged * kSlots; ++buffer) all[buffer] = p + buffer * unit;
        s.history = p + (size_t)kExchanged * kSlots * unit;
        s.barrier = w.barriers + (size_t)i * 32;
    }
    PMN_TRY(check_cuda(
        cudaMemcpyAsync(w.groups, host.data(), groups * sizeof(GroupState), cudaMemcpyHostToDevice, stream),
        "upload group state"));
    PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "sync"));  // `host` must outlive the copy
    PMN_TRY(check_cuda(cudaMemsetAsync(w.barriers, 0, (size_t)groups * 128, stream), "memset barriers"));

    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(fargan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes),
            "fargan smem attribute"));
        configured = true;
    }
    const int groups_per_launch = max_items / kGroupItems;
    for (int first = 0; first < groups; first += groups_per_launch) {
        const int count = std::min(groups_per_launch, groups - first);
        const float* weights = g->images;
        const float* cond = x;
        const float* feats = w.features;
        const float* prev = previous_samples ? previous_samples + (size_t)first * kGroupItems * kHistory : nullptr;
        GroupState* gs = w.groups + first;
        float* out = audio + (size_t)first * kGroupItems * frames * kHop;
        int items = std::min(batch - first * kGroupItems, count * kGroupItems);
        // cond / features are indexed by absolute utterance: offset the bases instead
        cond += (size_t)first * kGroupItems * kCond * frames;
        feats += (size_t)first * kGroupItems * 114 * frames;
        int frames_arg = frames;
        static int debug = -1;
        if (debug < 0) {
            const char* flag = getenv("PMN_FARGAN_DEBUG");
            debug = flag && flag[0] == '1';
        }
        void* args[] = {&weights, &cond, &feats, &prev, &gs, &out, &items, &frames_arg, &debug};
        LaunchScope scope("fargan_kernel", stream);
        PMN_TRY(check_cuda(
            cudaLaunchCooperativeKernel(
                (void*)fargan_kernel, dim3(count * kGroupCtas), dim3(kThreads), args, kSmemBytes, stream),
            "fargan cooperative launch"));
        if (debug) {
            long long host[8];
            cudaStreamSynchronize(stream);
            cudaMemcpyFromSymbol(host, g_fargan_cycles, sizeof(host));
            fprintf(stderr, "fargan CTA 0: %lld subframes, %lld cycles each: %lld building the subframe input (%lld of "
                            "them waiting), %lld waiting in the stage calls, %lld in dot products, %lld in reductions\n",
                    host[3], host[0] / host[3], host[6] / host[3], host[1] / host[3], host[2] / host[3],
                    host[4] / host[3], host[5] / host[3]);
        }
    }
    return PMN_OK;
}

}  // namespace pmn

// Viterbi decoding (torbi.from_probabilities as called at
// promonet/preprocess/harmonics.py:270-276 and inside penn's 'viterbi' decoder,
// promonet/preprocess/core.py:64-81).
//
//   delta_0[j] = log pi[j] + log o_0[j]
//   delta_t[j] = max_i (delta_{t-1}[i] + log A[i, j]) + log o_t[j],  psi_t[j] = argmax_i (lowest i on ties)
//   path: backtrace from argmax_j delta_{T-1}[j]
//
// The recurrence is sequential in t, so one CTA walks one utterance with delta
// double-buffered in shared memory.  Pitch transition matrices are banded, so a
// prepass stores log A column by column as band[k][j] = log A[lo_j + k, j] for
// the rows lo_j .. hi_j that are not -inf: thread j then reads band[k * S + j]
// (coalesced over j) and delta[lo_j + k] (conflict-free).  A dense matrix is the
// same code with lo = 0.  fp32 add/compare only after the logs, so the indices
// are bit-exact against the CPU recurrence on the same log inputs.
#include <math.h>

#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "spectral.cuh"
#include "tc_ptx.cuh"

namespace pmn {

namespace {

constexpr int kThreads = 1024;

// Banded fast path (viterbi_cluster_kernel below)
constexpr int kClusterSize = 8;
constexpr int kSplit = 4;                       // threads per state
constexpr int kClusterThreads = 736;      // 23 warps: 4 threads x 180 states, 88 registers each
constexpr int kClusterSmem = 200 * 1024;
constexpr int kPair = 4;                        // utterances a cluster decodes at once, at most (below)
// Bands of at most kSplit * kRegisterBand rows (penn's pitch transition: 181) stay in REGISTERS:
// thread (state, part) keeps rows part, part + kSplit, ... of its column for the whole utterance,
// so a band entry costs one conflict-free shared-memory load (the scores) instead of two loads
// with two-way bank conflicts.  Measured before: 5.7 of the 7.4 kcycles of a frame were this scan.
constexpr int kRegisterBand = 46;

__host__ __device__ inline int cluster_slice(int states) {
    return (states + kClusterSize - 1) / kClusterSize;
}
// floats of shared memory the fast path needs for a given band width
__host__ __device__ inline size_t cluster_floats(int states, int max_width) {
    const int slice = cluster_slice(states);
    return (((size_t)max_width * (slice | 1) + 3) & ~(size_t)3) + 2 * kPair * (size_t)kClusterSize * slice + 16;
}


// Per column j: first and one-past-last row with a finite log-probability
__global__ void band_range_kernel(
    const float* __restrict__ transition, bool log_probs, int states,
    int* __restrict__ lo, int* __restrict__ width, int* __restrict__ max_width) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= states) return;
    int first = states, last = 0;
    for (int i = 0; i < states; ++i) {
        const float value = transition[(size_t)i * states + j];
        const bool live = log_probs ? value > -INFINITY : value > 0.f;
        if (live) { first = min(first, i); last = i + 1; }
    }
    if (first >= last) { first = 0; last = 1; }
    lo[j] = first;
    width[j] = last - first;
    atomicMax(max_width, last - first);
}

__global__ void band_fill_kernel(
    const float* __restrict__ transition, bool log_probs, int states,
    const int* __restrict__ lo, const int* __restrict__ width, const int* __restrict__ max_width,
    float* __restrict__ band) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (j >= states || k >= *max_width) return;
    float value = -INFINITY;
    if (k < width[j]) {
        value = transition[(size_t)(lo[j] + k) * states + j];
        if (!log_probs) value = logf(value);
    }
    band[(size_t)k * states + j] = value;
}

__global__ void __launch_bounds__(kThreads) viterbi_kernel(
    const float* __restrict__ observation, const int* __restrict__ batch_frames,
    const float* __restrict__ initial, bool log_probs,
    const float* __restrict__ band, const int* __restrict__ lo, const int* __restrict__ width,
    const int* __restrict__ max_width_ptr,
    short* __restrict__ psi, int* __restrict__ indices, int frames, int states) {
    if (max_width_ptr != nullptr &&
        cluster_floats(states, *max_width_ptr) * sizeof(float) <= (size_t)kClusterSmem &&
        kSplit * cluster_slice(states) <= kClusterThreads)
        return;  // the banded cluster kernel decoded this batch
    extern __shared__ float delta[];  // [2][states]
    __shared__ float best_value[32];
    __shared__ int best_index[32];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int length = batch_frames ? min(batch_frames[b], frames) : frames;
    const float* obs = observation + (size_t)b * frames * states;
    short* back = psi + (size_t)b * frames * states;
    int* path = indices + (size_t)b * frames;
    if (length <= 0) {
        for (int t = tid; t < frames; t += kThreads) path[t] = 0;
        return;
    }

    for (int j = tid; j < states; j += kThreads) {
        const float o = log_probs ? obs[j] : logf(obs[j]);
        const float p = log_probs ? initial[j] : logf(initial[j]);
        delta[j] = p + o;
    }
    __syncthreads();

    int current = 0;
    for (int t = 1; t < length; ++t) {
        const float* previous = delta + current * states;
        float* next = delta + (current ^ 1) * states;
        const float* row = obs + (size_t)t * states;
        for (int j = tid; j < states; j += kThreads) {
            const int first = lo[j], count = width[j];
            float best = -INFINITY;
            int arg = 0;
            const float* column = band + j;
            const float* source = previous + first;
            for (int k = 0; k < count; ++k) {
                const float value = source[k] + column[(size_t)k * states];
                if (value > best) { best = value; arg = first + k; }
            }
            const float o = log_probs ? row[j] : logf(row[j]);
            next[j] = best + o;
            back[(size_t)t * states + j] = (short)arg;
        }
        current ^= 1;
        __syncthreads();
    }

    // argmax over the final scores, lowest index on ties
    const float* final_scores = delta + current * states;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = tid; j < states; j += kThreads) {
        const float value = final_scores[j];
        if (value > best) { best = value; arg = j; }
    }
    for (int offset = 16; offset > 0; offset >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, best, offset);
        const int other_arg = __shfl_xor_sync(0xffffffffu, arg, offset);
        if (other > best || (other == best && other_arg < arg)) { best = other; arg = other_arg; }
    }
    if ((tid & 31) == 0) { best_value[tid >> 5] = best; best_index[tid >> 5] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            if (best_value[w] > best || (best_value[w] == best && best_index[w] < arg)) {
                best = best_value[w];
                arg = best_index[w];
            }
        }
        int state = arg == 0x7fffffff ? 0 : arg;
        if (!(best > -INFINITY)) state = 0;  // CPU recurrence keeps index 0 when nothing beats -inf
        for (int t = length - 1; t >= 0; --t) {
            path[t] = state;
            if (t > 0) state = back[(size_t)t * states + state];
        }
        for (int t = length; t < frames; ++t) path[t] = 0;
    }
}

// ---------------------------------------------------------------------------
// Banded fast path: one cluster of 8 CTAs per utterance.
//
// Each CTA owns a contiguous slice of the next-states and keeps the band of the
// log-transition matrix for its columns resident in shared memory for the whole
// utterance (181 x 180 floats for penn's pitch transition), so the per-frame
// work is shared-memory only: every state is scanned by 4 threads (a quarter of
// the band each, combined with two shuffles, lowest index winning ties), the new
// scores are published in the CTA's own shared memory, and after one cluster
// barrier every CTA gathers the full score vector from its peers through
// distributed shared memory.  Exits immediately when the band does not fit
// (dense transition matrices): viterbi_kernel then does the work.
// ---------------------------------------------------------------------------

// A B200 runs 15 clusters of 8 CTAs at once at one CTA per SM (cudaOccupancyMaxActiveClusters;
// profiles/debug/cluster_probe.cu), so a batch of 32 utterances at one per cluster takes three
// waves, the last with two clusters.  A cluster therefore decodes up to kPair utterances at once,
// as many as it takes to fit the batch into one wave (launch_viterbi): their recurrences are
// independent, the frame of one is scanned while the scores of another travel, and the band
// registers serve all of them (measured per frame and utterance: 5.1 kcycles alone, 4.0 in pairs).

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kClusterThreads, 1)
viterbi_cluster_kernel(
    const float* __restrict__ observation, const int* __restrict__ batch_frames,
    const float* __restrict__ initial, bool log_probs,
    const float* __restrict__ band, const int* __restrict__ lo, const int* __restrict__ width,
    const int* __restrict__ max_width_ptr,
    short* __restrict__ psi, int* __restrict__ indices, int frames, int states, int batch, int per_cluster,
    long long* __restrict__ debug) {
    namespace cg = cooperative_groups;
    long long stamps[4];
    if (debug) stamps[0] = clock64();
    const int max_width = *max_width_ptr;
    const int slice = cluster_slice(states);
    if (cluster_floats(states, max_width) * sizeof(float) > (size_t)kClusterSmem ||
        kSplit * slice > kClusterThreads)
        return;  // uniform across the grid: the general kernel handles it
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / kClusterSize) * per_cluster;      // first utterance of this cluster
    const int items = min(per_cluster, batch - b0);
    const int tid = threadIdx.x;
    const int pitch = slice | 1;                 // odd row pitch: conflict-free column reads

    extern __shared__ float smem[];
    float* band_s = smem;                        // [max_width][pitch]
    // every CTA keeps the WHOLE score vector of the previous and of the current frame of each of
    // its utterances.  A CTA publishes each new score straight into all eight copies with st.async
    // (a remote shared-memory store that reports its bytes to an mbarrier of the destination CTA),
    // so a frame needs no cluster barrier and no fence: a CTA starts frame t + 1 as soon as the
    // transaction barrier of frame t has counted the bytes of all `states` scores.  The two
    // buffers cannot be overwritten early: nobody can finish frame t + 1 (and write into the
    // buffer frame t was computed from) before every CTA has sent all its frame-t scores,
    // i.e. has finished reading that buffer.
    const int whole = kClusterSize * slice;
    float* full = band_s + (((size_t)max_width * pitch + 3) & ~(size_t)3);   // [kPair][2][whole] (+ 16)
    __shared__ uint64_t landed[kPair][2];               // landed[u][p]: scores of a frame of parity p

    const int j0 = rank * slice;
    const bool in_registers = max_width <= kSplit * kRegisterBand;
    if (!in_registers) {
        for (int idx = tid; idx < max_width * slice; idx += kClusterThreads) {
            const int k = idx / slice, jl = idx % slice;
            band_s[k * pitch + jl] = j0 + jl < states ? band[(size_t)k * states + j0 + jl] : -INFINITY;
        }
    }
    for (int idx = tid; idx < kPair * 2 * whole + 16; idx += kClusterThreads) full[idx] = -INFINITY;  // padded states
    if (tid == 0) {
        for (int i = 0; i < kPair * 2; ++i) tc::mbar_init(&landed[0][0] + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int jl = tid / kSplit, part = tid % kSplit;
    const int j = j0 + jl;
    const bool owner = jl < slice && j < states;
    int first = 0, count = 0;
    if (owner) { first = lo[j]; count = width[j]; }
    const int chunk = (count + kSplit - 1) / kSplit;
    const int begin = min(part * chunk, count), end = min(begin + chunk, count);
    float band_r[kRegisterBand];
    if (in_registers) {
#pragma unroll
        for (int i = 0; i < kRegisterBand; ++i) {
            const int k = kSplit * i + part;
            band_r[i] = (owner && k < count) ? band[(size_t)k * states + j] : -INFINITY;
        }
    }
    // the kSplit threads of a state share the publishing: thread `part` writes the copies of
    // CTAs part * (kClusterSize / kSplit) ...  (Measured alternatives, profiles/
    // r2_viterbi_breakdown.txt: four states per 16-byte st.async double the remote packets of a
    // warp, +33 %; one bulk copy of the slice per peer after a CTA barrier, +7 %.)
    constexpr int kPeersPerThread = kClusterSize / kSplit;
    uint32_t copies[kPeersPerThread], barriers[kPeersPerThread];   // shared::cluster addresses
#pragma unroll
    for (int i = 0; i < kPeersPerThread; ++i) {
        const uint32_t peer = part * kPeersPerThread + i;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                     : "=r"(copies[i]) : "r"(tc::smem_u32(full + j0 + jl)), "r"(peer));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                     : "=r"(barriers[i]) : "r"(tc::smem_u32(&landed[0][0])), "r"(peer));
    }
    // score of frame t (parity t & 1) of utterance u into every copy
    auto publish = [&](int u, float score, int parity) {
#pragma unroll
        for (int i = 0; i < kPeersPerThread; ++i)
            asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
                         ::"r"(copies[i] + (uint32_t)((u * 2 + parity) * whole) * 4u), "r"(__float_as_uint(score)),
                           "r"(barriers[i] + (uint32_t)(u * 2 + parity) * 8u) : "memory");
    };
    // all `states` scores of frame t of utterance u have landed in this CTA's copy.  Only the
    // threads that read the scores wait (a thread that reads nothing could otherwise still be
    // polling for frame t when the same barrier completes frame t + 2; a reader cannot lag that
    // far, because frame t + 2 cannot complete anywhere without its frame t + 1 score)
    auto await = [&](int u, int t) {
        if (tid == 0) tc::mbar_expect_tx(&landed[u][t & 1], (uint32_t)states * 4u);
        if (owner || tid == 0) tc::mbar_wait(&landed[u][t & 1], (t >> 1) & 1);
    };
    __shared__ int length[kPair];          // uniform per CTA: kept out of the registers
    __shared__ const float* obs[kPair];
    __shared__ short* back[kPair];
    float next_observation[kPair];
    if (tid < kPair) {
        const int u = tid, b = b0 + u;
        length[u] = u < items ? (batch_frames ? min(batch_frames[b], frames) : frames) : 0;
        obs[u] = observation + (size_t)(u < items ? b : b0) * frames * states;
        back[u] = psi + (size_t)(u < items ? b : b0) * frames * states;
    }
    __syncthreads();
    int longest = 0;
#pragma unroll
    for (int u = 0; u < kPair; ++u) longest = max(longest, length[u]);
    auto load_observation = [&](int u, int t) {
        if (!owner || t >= length[u]) return 0.f;
        const float o = obs[u][(size_t)t * states + j];
        return log_probs ? o : logf(o);
    };

    cluster.sync();      // every copy is initialised and every barrier exists before the first store
    if (debug) stamps[1] = clock64();
#pragma unroll
    for (int u = 0; u < kPair; ++u) {
        if (length[u] > 0 && owner)
            publish(u, (log_probs ? initial[j] : logf(initial[j])) + load_observation(u, 0), 0);
        next_observation[u] = load_observation(u, 1);    // one frame ahead of its use
    }

    for (int t = 1; t < longest; ++t) {
#pragma unroll
        for (int u = 0; u < kPair; ++u) {
            if (t >= length[u]) continue;                // uniform across the cluster
            await(u, t - 1);
            const float observed = next_observation[u];
            next_observation[u] = load_observation(u, t + 1);
            const float* previous = full + (size_t)(u * 2 + ((t - 1) & 1)) * whole;
            float best = -INFINITY;
            int arg = 0x7fffffff;
            if (in_registers) {
                // rows part, part + kSplit, ...: the lanes of a warp read 11 consecutive scores per
                // step (conflict-free); rows past the band carry -inf and never win
                const float* source = previous + first + part;
                int winner = -1;
                constexpr int kGroup = 8;     // loads issued together, ahead of their compare chain
                float scores[kGroup];
#pragma unroll
                for (int i0 = 0; i0 < kRegisterBand; i0 += kGroup) {
#pragma unroll
                    for (int i = 0; i < kGroup; ++i)
                        if (i0 + i < kRegisterBand) scores[i] = source[kSplit * (i0 + i)];
#pragma unroll
                    for (int i = 0; i < kGroup; ++i) {
                        if (i0 + i < kRegisterBand) {
                            const float value = scores[i] + band_r[i0 + i];
                            if (value > best) { best = value; winner = i0 + i; }
                        }
                    }
                }
                if (winner >= 0) arg = first + kSplit * winner + part;
            } else if (owner) {
                const float* column = band_s + jl;
                const float* source = previous + first;
                for (int k = begin; k < end; ++k) {
                    const float value = source[k] + column[k * pitch];
                    if (value > best) { best = value; arg = first + k; }
                }
            }
            // combine the kSplit parts of a state (adjacent lanes), lowest index on ties
#pragma unroll
            for (int offset = 1; offset < kSplit; offset <<= 1) {
                const float other = __shfl_xor_sync(0xffffffffu, best, offset);
                const int other_arg = __shfl_xor_sync(0xffffffffu, arg, offset);
                if (other > best || (other == best && other_arg < arg)) { best = other; arg = other_arg; }
            }
            if (owner) {
                publish(u, best + observed, t & 1);
                if (part == 0) back[u][(size_t)t * states + j] = (short)(arg == 0x7fffffff ? 0 : arg);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kPair; ++u)
        if (length[u] > 0) await(u, length[u] - 1);

    // final argmax and backtrace of utterance u on CTA u of the cluster (every CTA holds the whole
    // vector)
    if (debug) stamps[2] = clock64();
    __threadfence();
    cluster.sync();
    if (rank >= items) return;
    const int mine = rank;
    const int my_length = length[mine];
    int* path = indices + (size_t)(b0 + mine) * frames;
    // The backtrace is a chain of `length` dependent loads.  Followed through global memory
    // by one thread it costs an L2 round trip per frame; instead all threads stage the
    // back-pointers of a block of frames in the shared memory the band no longer needs, and
    // the walker steps through them there.
    __shared__ int walker_state;
    if (tid == 0) {
        const float* last = full + (size_t)(mine * 2 + ((my_length - 1) & 1)) * whole;
        int state = 0;
        float best = -INFINITY;
        if (my_length > 0) {
            for (int k = 0; k < whole; ++k)
                if (k < states && last[k] > best) { best = last[k]; state = k; }
        }
        walker_state = state;
        for (int t = max(my_length, 0); t < frames; ++t) path[t] = 0;
    }
    short* staged = reinterpret_cast<short*>(band_s);
    const short* mine_back = back[mine];
    const int block_frames = max(1, (int)(((size_t)max_width * pitch * sizeof(float)) / ((size_t)states * sizeof(short))));
    const bool wide = ((size_t)states * sizeof(short)) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(mine_back) & 15) == 0;   // 16-byte loads
    for (int hi = my_length - 1; hi >= 0; hi -= block_frames) {
        const int lo_frame = max(hi - block_frames + 1, 0);
        __syncthreads();                                  // the walker is done with the last block
        const size_t offset = (size_t)lo_frame * states;
        const int shorts = (hi - lo_frame + 1) * states;
        if (wide) {
            const uint4* source = reinterpret_cast<const uint4*>(mine_back + offset);
            uint4* target = reinterpret_cast<uint4*>(staged);
            for (int idx = tid; idx < shorts / 8; idx += kClusterThreads) target[idx] = __ldcg(source + idx);
        } else {
            for (int idx = tid; idx < shorts; idx += kClusterThreads) staged[idx] = __ldcg(mine_back + offset + idx);
        }
        __syncthreads();
        if (tid == 0) {
            int state = walker_state;
            for (int t = hi; t >= lo_frame; --t) {
                path[t] = state;
                if (t > 0) state = staged[(size_t)(t - lo_frame) * states + state];
            }
            walker_state = state;
        }
    }
    if (debug && blockIdx.x == 0 && tid == 0) {
        stamps[3] = clock64();
        for (int i = 0; i < 4; ++i) debug[i] = stamps[i] - stamps[0];
    }
}

struct Workspace {
    float* band;
    int *lo, *width, *max_width;
    short* psi;
    size_t bytes;
};

Workspace carve(void* base, int batch, int frames, int states) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += align_up(bytes, 256);
        return r;
    };
    w.band = (float*)take((size_t)states * states * sizeof(float));
    w.lo = (int*)take((size_t)states * sizeof(int));
    w.width = (int*)take((size_t)states * sizeof(int));
    w.max_width = (int*)take(sizeof(int));
    w.psi = (short*)take((size_t)batch * frames * states * sizeof(short));
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}

}  // namespace

size_t viterbi_workspace_bytes(int batch, int frames, int states) {
    return carve(nullptr, batch, frames, states).bytes;
}

int launch_viterbi(
    const float* observation, const int* batch_frames, const float* transition,
    const float* initial, bool log_probs, int* indices, int batch, int frames, int states,
    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    PMN_REQUIRE(observation && transition && initial && indices && workspace, "viterbi: null pointer");
    PMN_REQUIRE(batch > 0 && frames > 0 && states > 0, "viterbi: empty input");
    PMN_REQUIRE(states <= 32767, "viterbi: more than 32767 states");
    PMN_REQUIRE((size_t)2 * states * sizeof(float) <= 200 * 1024, "viterbi: too many states for shared memory");
    Workspace w = carve(workspace, batch, frames, states);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "viterbi: workspace too small");
    PMN_TRY(check_cuda(cudaMemsetAsync(w.max_width, 0, sizeof(int), stream), "viterbi memset"));
    {
        LaunchScope scope("band_range_kernel", stream);
        band_range_kernel<<<ceil_div(states, 128), 128, 0, stream>>>(
            transition, log_probs, states, w.lo, w.width, w.max_width);
        PMN_TRY(launched("band_range_kernel"));
    }
    {
        dim3 grid(ceil_div(states, 128), states);
        LaunchScope scope("band_fill_kernel", stream);
        band_fill_kernel<<<grid, 128, 0, stream>>>(
            transition, log_probs, states, w.lo, w.width, w.max_width, w.band);
        PMN_TRY(launched("band_fill_kernel"));
    }
    const size_t smem = (size_t)2 * states * sizeof(float);
    static bool configured = false;
    if (!configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
            "viterbi smem attribute"));
        configured = true;
    }
    static bool cluster_configured = false;
    if (!cluster_configured) {
        PMN_TRY(check_cuda(
            cudaFuncSetAttribute(
                viterbi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterSmem),
            "viterbi cluster smem attribute"));
        cluster_configured = true;
    }
    {
        // banded fast path (returns at once if the band does not fit in shared memory)
        LaunchScope scope("viterbi_cluster_kernel", stream);
        // more utterances than clusters that run at once (8 GPCs x 2 clusters of 8 SMs): two per cluster
        static long long* debug = nullptr;
        static bool debug_checked = false;
        if (!debug_checked) {
            debug_checked = true;
            const char* flag = getenv("PMN_VITERBI_DEBUG");
            if (flag && flag[0] == '1') cudaMalloc(&debug, 4 * sizeof(long long));
        }
        // as many utterances per cluster as it takes to run the batch in one wave of clusters
        static int wave = 0;
        if (!wave) {
            cudaLaunchConfig_t config = {};
            config.gridDim = dim3(kClusterSize);
            config.blockDim = dim3(kClusterThreads);
            config.dynamicSmemBytes = kClusterSmem;
            cudaLaunchAttribute attribute;
            attribute.id = cudaLaunchAttributeClusterDimension;
            attribute.val.clusterDim.x = kClusterSize;
            attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
            config.attrs = &attribute;
            config.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&wave, viterbi_cluster_kernel, &config) != cudaSuccess || wave < 1) {
                cudaGetLastError();
                wave = 15;      // measured on a B200
            }
        }
        const char* pair_flag = getenv("PMN_VITERBI_PAIR");   // profiling aid: force the count
        const int per_cluster = pair_flag ? std::max(1, std::min(kPair, atoi(pair_flag)))
                                          : std::max(1, std::min(kPair, (batch + wave - 1) / wave));
        const int clusters = (batch + per_cluster - 1) / per_cluster;
        viterbi_cluster_kernel<<<clusters * kClusterSize, kClusterThreads, kClusterSmem, stream>>>(
            observation, batch_frames, initial, log_probs, w.band, w.lo, w.width, w.max_width,
            w.psi, indices, frames, states, batch, per_cluster, debug);
        PMN_TRY(launched("viterbi_cluster_kernel"));
        if (debug) {
            // profiling aid (PMN_VITERBI_DEBUG=1): cycles of CTA 0 at the start of the frame loop,
            // at its end and at the end of the backtrace
            long long host[4];
            cudaStreamSynchronize(stream);
            cudaMemcpy(host, debug, sizeof(host), cudaMemcpyDeviceToHost);
            fprintf(stderr, "viterbi CTA 0: frame loop starts at %lld, ends at %lld, kernel ends at %lld cycles\n",
                    host[1], host[2], host[3]);
        }
    }
    LaunchScope scope("viterbi_kernel", stream);
    viterbi_kernel<<<batch, kThreads, smem, stream>>>(
        observation, batch_frames, initial, log_probs, w.band, w.lo, w.width, w.max_width, w.psi,
        indices, frames, states);
    return launched("viterbi_kernel");
}

}  // namespace pmn

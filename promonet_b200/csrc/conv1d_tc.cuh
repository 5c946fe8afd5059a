// tcgen05 dilated Conv1d (conv1d_tc.cu): operand layout and launchers
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace pmn {

// Activations for the tensor-core path are stored as two bf16 "planes" (hi, lo)
// with a = hi + lo to ~16 mantissa bits, channel-chunked and time-major:
//   planes[b][plane][c / 8][kTcPad + t][c % 8],   t in [-kTcPad, t_pad - kTcPad)
// Rows outside [0, T) are zero: they are the convolution's zero padding.
// A (time x channel) operand tile is then a set of contiguous runs, one per
// 8-channel group, and a tap shift is a row shift (16 B) of the operand address.
constexpr int kTcPad = 32;       // >= max halo (k-1)/2 * dilation = 25
constexpr int kTcTimeAlign = 512;

inline int tc_padded_length(int t_len) {
    return kTcPad + (t_len + kTcTimeAlign - 1) / kTcTimeAlign * kTcTimeAlign + kTcPad;
}
__device__ __forceinline__ int tc_padded_length_device(int t_len) {
    return kTcPad + (t_len + kTcTimeAlign - 1) / kTcTimeAlign * kTcTimeAlign + kTcPad;
}
inline size_t tc_planes_elements(int batch, int channels, int t_len) {
    return (size_t)batch * 2 * channels * tc_padded_length(t_len) + 8192;   // + kTcPlanesSlack
}
// Weight slabs (bf16 hi + lo = the bytes of the fp32 tensor): see pack_tc_weight_kernel
struct TcPlan {
    int k_block;   // input channels per shared-memory slab
    int n_tile;    // output channels per tile
    bool concat;   // hi and lo rows interleaved per K group (one wide MMA for a_hi)
};
bool tc_conv_plan(int c_in, int c_out, bool frames, TcPlan* plan);
// Narrow square layers (C = 32 / 64, odd k <= 11) also run with the weights on the M side of the
// MMAs (conv1d_tcw.cu); their slabs in that kernel's format follow the plain ones in the buffer
bool tcw_shape_supported(int c_in, int c_out, int k);
size_t tcw_weight_elements(int channels, int k);
inline size_t tc_plain_weight_elements(int c_out, int c_in, int k) { return (size_t)2 * c_out * c_in * k; }
inline size_t tc_weight_elements(int c_out, int c_in, int k) {
    return tc_plain_weight_elements(c_out, c_in, k) +
           (tcw_shape_supported(c_in, c_out, k) ? tcw_weight_elements(c_in, k) : 0);
}
bool tc_supported(int c_in, int c_out, int k, int dilation);

struct TcConvArgs {
    const __nv_bfloat16* x_planes = nullptr;  // lrelu already applied by the producer
    const __nv_bfloat16* w_slabs = nullptr;
    const float* bias = nullptr;              // (C_out) or null
    const float* bias_batch = nullptr;        // (B, C_out) added per item, or null (conv1d_tc_kernel only)
    const float* residual = nullptr;          // (B, C_out, T) fp32 or null
    float* out = nullptr;                     // (B, C_out, T) fp32 or null
    __nv_bfloat16* out_planes = nullptr;      // planes of lrelu(y, out_slope) or null
    float* accum = nullptr;                   // (B, C_out, T) fp32 or null
    int accum_mode = 0;                       // 0 unused, 1 store, 2 add
    float accum_scale = 1.f;
    bool planes_from_accum = false;           // out_planes = planes of lrelu(the value stored to accum)
    int batch = 0, c_in = 0, c_out = 0, t_len = 0;
    int k = 1, dilation = 1;
    bool valid = false;                       // false: "same" padding (odd k); true: no padding
    bool relu = false;                        // max(y, 0) before any store
    bool pool = false;                        // MaxPool1d(2) over row pairs: out row t / 2, out_row = pooled length
    int out_row = 0;                          // fp32 row length of out / residual / accum (0 = output length)
    // Frame mode (valid conv over frames laid end to end, a few output rows per
    // frame): frames of frame_length rows each, the first frame_valid outputs kept
    int frames = 0, frame_length = 0, frame_valid = 0;
    float out_slope = 1.f;
    // Operand addressing in 8-channel groups (0 = the default planes[b][plane][c / 8] layout):
    // item b starts item_groups groups after item b - 1 and the lo plane plane_groups groups after
    // the hi plane.  The frame-major pitch layers (pitch.cu) read planes[plane][position][c / 8]
    // with the batch index as the position: item_groups = C / 8, plane_groups = positions C / 8
    int item_groups = 0, plane_groups = 0;
    // "fp16 + 2 x fp8" operands (tc_f8_* below): x_planes / w_slabs are in the three-section format;
    // all three products carry the same power-of-two scale, which the epilogue takes out again
    bool f8x2 = false;
    float f8_unscale = 1.f;                   // tc_f8_unscale(weight_shift)
    bool out_f8 = false;                      // out_planes in the three-section format (any kernel mode)
    // Optional (gridDim.x, 10 warps, 4) cycle counters: [0] total, [1..3] barrier waits
    long long* debug = nullptr;
};

int launch_conv1d_tc(const TcConvArgs& args, cudaStream_t stream);
// conv1d_tcw.cu: the same convolution for the narrow square layers (launch_conv1d_tc dispatches to
// it when tcw_applies; PMN_TCW=0 in the environment keeps everything on conv1d_tc_kernel)
bool tcw_applies(const TcConvArgs& args);     // the kernel can run it
bool tcw_preferred(const TcConvArgs& args);   // ... and is the faster one (PMN_TCW=2: whenever it applies)
int launch_conv1d_tcw(const TcConvArgs& args, const __nv_bfloat16* slabs, cudaStream_t stream);
int launch_pack_tcw_weight(const float* w, __nv_bfloat16* slabs, int channels, int k, cudaStream_t stream);
// rows a kernel may read past the end of a planes buffer (conv1d_tcw's window of the last tile):
// every planes allocation ends with this many spare elements
constexpr size_t kTcPlanesSlack = 8192;
// Every following launch writes its cycle counters to `counters` (device, SMs x 40 int64); null = off
void tc_set_debug_counters(long long* counters);

// LeakyReLU'd planes (B, 2, C_in, .) -> ConvTranspose1d(kernel 2 stride, padding
// stride / 2) fp32 (B, C_out, stride * T) on the tensor cores.  Uses x_planes,
// w_slabs (launch_pack_tc_transpose_weight), bias, out, batch, c_in, c_out, t_len.
bool tc_transpose_supported(int c_in, int c_out, int k, int stride);
size_t tc_transpose_weight_elements(int c_in, int c_out, int stride);
int launch_conv_transpose1d_tc(const TcConvArgs& args, int stride, cudaStream_t stream);
int launch_pack_tc_transpose_weight(
    const float* w, __nv_bfloat16* slabs, int c_in, int c_out, int stride, cudaStream_t stream);

// fp32 (B, C, T) -> planes of lrelu(x, slope); also writes the zero pad rows.  source_channels
// (0 = channels): channels of x; the planes' channels beyond them are zero (an operand padded
// to a multiple of the kernel's K block)
int launch_planes_from_f32(
    const float* x, __nv_bfloat16* planes, int batch, int channels, int t_len, float slope,
    cudaStream_t stream, int source_channels = 0, bool f8 = false);

// planes -> fp32 (hi + lo; f8: the fp16 section + the low section, both unscaled), for tests
int launch_f32_from_planes(
    const __nv_bfloat16* planes, float* x, int batch, int channels, int t_len, cudaStream_t stream,
    bool f8 = false);

// Zero only the pad rows of a planes buffer (the kernels never write them)
int launch_zero_plane_pads(
    __nv_bfloat16* planes, int batch, int channels, int t_len, cudaStream_t stream);

// One residual pair y = x + c2(lrelu(c1(lrelu(x)))) (hifigan.py:198-210) in one kernel
// (conv_pair_tc.cu): x, out, accum are fp32 (B, C, T); w1 / w2 are the slabs of
// launch_pack_tc_weight; C in {32, 64, 128}, odd k <= 11, (k - 1) dilation <= 50.
// out / accum must not alias x.  Bit-identical to two conv1d_tc_kernel launches (where launch_conv1d_tc
// takes conv1d_tcw_kernel instead, the two agree to fp32 rounding: another order of the tap sum).
bool tc_pair_supported(int channels, int k, int dilation);
int launch_conv_pair_tc(
    const float* x, const __nv_bfloat16* w1, const float* bias1, const __nv_bfloat16* w2,
    const float* bias2, float* out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float slope, cudaStream_t stream);

// Profiling aid: per-CTA cycle counters of every following pair launch (null = off) and the
// kernel variant to launch (-1 = default)
void tc_pair_set_debug(long long* counters, int variant);

// fp32-grade products at two thirds of the tensor cycles, for operands of a bounded range (|x| < 56:
// LayerNorm outputs of the pitch network, LeakyReLU'd activations of the generator), with
// x_m = fp16(x s_m), w_m = fp16(w s_wm):
//   x w s ~ x_m w_m  [kind::f16]  + e4m3(x s_x) e4m3((w s_wm - w_m) s_wl)
//                                 + e4m3((x s_m - x_m) s_xl) e4m3(w s_w)   [kind::f8f6f4, K = 32]
// All scales are powers of two chosen so that the three products carry the SAME factor
// s = s_m s_wm = s_x s_wl' = s_xl' s_w (primes: per unit of x / w), so they accumulate in one TMEM
// accumulator and the epilogue multiplies by 1 / s (exact).  The two corrections are 2^-11 of the
// main term: the three mantissa bits of e4m3 are enough for them.
// Operand format, in 16-byte rows like the planes: [C / 8 groups][t_pad][8] fp16 of x s_m, then
// [C / 16][t_pad][16] e4m3 of x s_x, then the same of (x - x_m / s_m) s_xl
// (device side: tc::split_pair_f8, whose constants mirror these).
constexpr float kF8ScaleMain = 128.f;       // s_m: |x| < 512 in fp16 (saturating)
constexpr float kF8ScaleX = 8.f;            // s_x: |x| <= 56 before the e4m3 of x saturates
constexpr float kF8ScaleXLow = 16384.f;     // s_xl = 2^14: (x - x_m / s_m) <= 2^-11 |x|
// weight slabs [n tile][tap][c_in / KB] x { [KB / 8][N][8] fp16 | [KB / 16][N][16] e4m3 low | e4m3 high }
// with s_w = 2^weight_shift (|w| s_w <= 256), s_wm = s_w s_xl / s_m, s_wl' = s_w s_xl / s_x
bool tc_f8_plan(int c_in, int c_out, TcPlan* plan);     // shapes the fp8 form is built for, their tiling
int launch_pack_tc_weight_f8(
    const float* w, void* slabs, int c_out, int c_in, int k, int weight_shift, cudaStream_t stream);
inline float tc_f8_unscale(int weight_shift) {
    return 1.f / (kF8ScaleXLow * (float)(1 << weight_shift));
}
// ... for a weight tensor on the device (synchronises the stream: model set-up and parity entries only)
int tc_f8_weight_shift_of(const float* w, size_t numel, cudaStream_t stream, int* shift);
// the largest shift with max |w| 2^shift <= 256
inline int tc_f8_weight_shift(float largest) {
    int shift = 0;
    while (shift < 16 && largest * (float)(2 << shift) <= 256.f) ++shift;
    return shift;
}

// folded fp32 weight (C_out, C_in, K) -> hi/lo slabs
int launch_pack_tc_weight(
    const float* w, __nv_bfloat16* slabs, int c_out, int c_in, int k, bool frames, cudaStream_t stream);

}  // namespace pmn

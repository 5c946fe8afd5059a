// Launchers of the training-step kernels (train_conv.cu, train_ops.cu, train_spectral.cu)
#pragma once

#include "common.cuh"

namespace pmn {

// Activation applied to an operand while it is loaded (pmn_operand_act)
enum OperandAct {
    kActNone = 0,
    kActLrelu = 1,      // lrelu(value, slope)
    kActLreluMask = 2,  // value * (companion > 0 ? 1 : slope)   (backward of an output LeakyReLU)
    kActTanhMask = 3    // value * (1 - companion^2)            (backward of an output tanh)
};

// Activation applied in the epilogue (pmn_output_act)
enum OutputAct { kOutNone = 0, kOutLrelu = 1, kOutTanh = 2 };

struct ConvGemmArgs {
    pmn_conv_geometry g;           // geometry of the FORWARD convolution
    int transposed = 0;            // 0: gather input positions (forward); 1: gather output positions (dgrad)
    const float* a = nullptr;      // gathered tensor
    const float* a_companion = nullptr;
    int a_act = kActNone;
    float a_slope = 1.f;
    const float* wmat = nullptr;   // (out channels, gathered channels * kh * kw)
    const float* bias = nullptr;   // (out channels)
    const float* bias2 = nullptr;  // (batch, out channels)
    int out_act = kOutNone;
    float out_slope = 1.f;
    const float* mask_src = nullptr;  // same shape as out: value *= lrelu'(mask_src)
    float mask_slope = 1.f;
    const float* residual = nullptr;  // same shape as out
    float alpha = 1.f;
    int accumulate = 0;            // out += ...
    float* out = nullptr;
};

int launch_conv_gemm(const ConvGemmArgs& args, cudaStream_t stream);

// train_conv_tc.cu: the same contract on tcgen05 (tf32 operands); args.wmat is the
// tap-major packed weight (rows, taps, conv_tc_channel_pad(reduction channels))
int launch_conv_gemm_tc(const ConvGemmArgs& args, cudaStream_t stream);
int conv_tc_channel_pad(int channels);
size_t conv_tc_packed_floats(int rows, int reduce, int taps);
int launch_extract_grouped(
    const float* dense, float* gw, int dim0, int dim1, int taps, int groups, cudaStream_t stream);
int launch_prepare_weights(const pmn_weight_desc* table, int layers, int max_dim0, cudaStream_t stream);
void set_train_tc_debug(long long* counters);
int launch_pack_weight_taps(
    const float* w, float* out, int d0, int d1, int taps, int transposed, cudaStream_t stream);

struct ConvWgradArgs {
    pmn_conv_geometry g;
    const float* dy = nullptr;     // (B, c_out, h_out, w_out)
    const float* dy_companion = nullptr;
    int dy_act = kActNone;
    float dy_slope = 1.f;
    const float* x = nullptr;      // (B, c_in, h_in, w_in)
    const float* x_companion = nullptr;
    int x_act = kActNone;
    float x_slope = 1.f;
    float* gw = nullptr;           // (c_out, c_in, kh, kw), accumulated atomically
    float* gbias = nullptr;        // (c_out) or null, accumulated atomically
};

int launch_conv_wgrad(const ConvWgradArgs& args, cudaStream_t stream);
int launch_conv_wgrad_tc(const ConvWgradArgs& args, cudaStream_t stream);   // train_conv_tc.cu

int launch_transpose_weight(
    const float* w, float* wt, int dim0, int dim1, int taps, cudaStream_t stream);

int launch_weight_norm_backward(
    const float* v, const float* g, const float* gw, float* gv, float* gg, int dim0, int inner,
    cudaStream_t stream);
int launch_weight_norm_backward_table(
    const pmn_weight_norm_desc* table, int layers, int max_dim0, cudaStream_t stream);

// train_ops.cu
int launch_reflect_pad(
    const float* x, float* out, int rows, int t_in, int left, int right, cudaStream_t stream);
int launch_reflect_pad_backward(
    const float* gout, float* gx, int rows, int t_in, int left, int right, int accumulate,
    cudaStream_t stream);
int launch_axpby(float a, const float* x, float b, float* y, int64_t n, cudaStream_t stream);
int launch_mse_to_target(
    const float* x, int64_t n, float target, float weight, float* loss, float* grad,
    cudaStream_t stream);
int launch_l1_mean(
    const float* fake, const float* real, int64_t n, float weight, float* loss, float* gfake,
    int accumulate, cudaStream_t stream);
int launch_adamw(
    float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
    float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
    const float* step_device, cudaStream_t stream);
int launch_adamw_peer(
    const float* const* grad_peers, float* const* param_peers, int world, int rank, float* exp_avg,
    float* exp_avg_sq, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps,
    float weight_decay, int step, const float* step_device, cudaStream_t stream);
int launch_row_sum(
    const float* x, float* out, int rows, int cols, int accumulate, cudaStream_t stream);
int launch_channel_sum(
    const float* x, float* out, int batch, int channels, int inner, int accumulate,
    cudaStream_t stream);
int launch_copy_columns(
    const float* src, int src_width, int src_offset, float* dst, int dst_width, int dst_offset,
    int64_t rows, int cols, int accumulate, cudaStream_t stream);
int launch_embedding_backward(
    const float* gout, const int64_t* index, float* gtable, int batch, int channels, int frames,
    int rows, int out_channels, int channel_offset, cudaStream_t stream);
int launch_pitch_bins(
    const float* pitch, const float* edges, int64_t* bins, int n, int num_edges, float fmin,
    float fmax, cudaStream_t stream);

int launch_global_features(
    const float* speaker_embedding, const int64_t* speakers, const float* sbr, const float* lr,
    float* out, int batch, int speaker_channels, int num_speakers, cudaStream_t stream);

// train_spectral.cu
size_t stft_train_frames(int samples);
int launch_stft_train(
    const float* audio, int batch, int samples, int window_kind, float eps, int layout,
    float* spectrum, float* magnitude, cudaStream_t stream);
int launch_stft_train_backward(
    const float* gmagnitude, const float* spectrum, int batch, int samples, int window_kind,
    float eps, int layout, float* gaudio, int accumulate, cudaStream_t stream);
int launch_mel_loss(
    const float* magnitude, const float* target_mels, int batch, int frames, float loss_weight,
    float grad_weight, float* loss, float* gmagnitude, cudaStream_t stream);

int launch_dft_basis(float* out, int n_fft, cudaStream_t stream);
int launch_spectral_convergence(
    const float* spec, int batch, int bins, int frames, float weight, float* sums, float* loss,
    float* gspec, cudaStream_t stream);
int launch_frame_overlap_add(
    const float* gframes, float* gsignal, int batch, int n_fft, int frames, int hop, int samples,
    cudaStream_t stream);

}  // namespace pmn

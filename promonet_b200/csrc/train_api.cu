// extern "C" surface of the training-step operators (include/promonet_b200.h)
#include "features.cuh"
#include "train.cuh"

using namespace pmn;

extern "C" {

int pmn_conv_gemm(
    const pmn_conv_geometry* geometry, int transposed,
    const float* a, const float* a_companion, int a_act, float a_slope,
    const float* wmat, const float* bias, const float* bias2,
    int out_act, float out_slope, const float* mask_src, float mask_slope,
    const float* residual, float alpha, int accumulate, float* out, void* stream) {
    PMN_REQUIRE(geometry, "conv_gemm: null geometry");
    PMN_REQUIRE(a_act >= 0 && a_act <= 3 && out_act >= 0 && out_act <= 2, "conv_gemm: bad activation");
    ConvGemmArgs args;
    args.g = *geometry;
    args.transposed = transposed != 0;
    args.a = a; args.a_companion = a_companion; args.a_act = a_act; args.a_slope = a_slope;
    args.wmat = wmat; args.bias = bias; args.bias2 = bias2;
    args.out_act = out_act; args.out_slope = out_slope;
    args.mask_src = mask_src; args.mask_slope = mask_slope;
    args.residual = residual; args.alpha = alpha; args.accumulate = accumulate != 0;
    args.out = out;
    return launch_conv_gemm(args, (cudaStream_t)stream);
}

int pmn_conv_gemm_tc(
    const pmn_conv_geometry* geometry, int transposed,
    const float* a, const float* a_companion, int a_act, float a_slope,
    const float* wpacked, const float* bias, const float* bias2,
    int out_act, float out_slope, const float* mask_src, float mask_slope,
    const float* residual, float alpha, int accumulate, float* out, void* stream) {
    PMN_REQUIRE(geometry, "conv_gemm_tc: null geometry");
    PMN_REQUIRE(a_act >= 0 && a_act <= 3 && out_act >= 0 && out_act <= 2, "conv_gemm_tc: bad activation");
    PMN_REQUIRE(((uintptr_t)wpacked & 15) == 0, "conv_gemm_tc: packed weights must be 16-byte aligned");
    ConvGemmArgs args;
    args.g = *geometry;
    args.transposed = transposed != 0;
    args.a = a; args.a_companion = a_companion; args.a_act = a_act; args.a_slope = a_slope;
    args.wmat = wpacked; args.bias = bias; args.bias2 = bias2;
    args.out_act = out_act; args.out_slope = out_slope;
    args.mask_src = mask_src; args.mask_slope = mask_slope;
    args.residual = residual; args.alpha = alpha; args.accumulate = accumulate != 0;
    args.out = out;
    return launch_conv_gemm_tc(args, (cudaStream_t)stream);
}

int pmn_conv_tc_channel_pad(int channels) { return conv_tc_channel_pad(channels); }

size_t pmn_conv_tc_packed_floats(int rows, int reduce, int taps) {
    if (rows <= 0 || reduce <= 0 || taps <= 0) return 0;
    return conv_tc_packed_floats(rows, reduce, taps);
}

void pmn_debug_train_tc_counters(void* counters) { set_train_tc_debug(static_cast<long long*>(counters)); }

int pmn_pack_weight_taps(
    const float* w, float* out, int d0, int d1, int taps, int transposed, void* stream) {
    return launch_pack_weight_taps(w, out, d0, d1, taps, transposed, (cudaStream_t)stream);
}

int pmn_conv_wgrad(
    const pmn_conv_geometry* geometry,
    const float* dy, const float* dy_companion, int dy_act, float dy_slope,
    const float* x, const float* x_companion, int x_act, float x_slope,
    float* gw, float* gbias, void* stream) {
    PMN_REQUIRE(geometry, "conv_wgrad: null geometry");
    PMN_REQUIRE(dy_act >= 0 && dy_act <= 3 && x_act >= 0 && x_act <= 3, "conv_wgrad: bad activation");
    ConvWgradArgs args;
    args.g = *geometry;
    args.dy = dy; args.dy_companion = dy_companion; args.dy_act = dy_act; args.dy_slope = dy_slope;
    args.x = x; args.x_companion = x_companion; args.x_act = x_act; args.x_slope = x_slope;
    args.gw = gw; args.gbias = gbias;
    return launch_conv_wgrad(args, (cudaStream_t)stream);
}

int pmn_conv_wgrad_tc(
    const pmn_conv_geometry* geometry,
    const float* dy, const float* dy_companion, int dy_act, float dy_slope,
    const float* x, const float* x_companion, int x_act, float x_slope,
    float* gw, float* gbias, void* stream) {
    PMN_REQUIRE(geometry, "conv_wgrad_tc: null geometry");
    ConvWgradArgs args;
    args.g = *geometry;
    args.dy = dy; args.dy_companion = dy_companion; args.dy_act = dy_act; args.dy_slope = dy_slope;
    args.x = x; args.x_companion = x_companion; args.x_act = x_act; args.x_slope = x_slope;
    args.gw = gw; args.gbias = gbias;
    return launch_conv_wgrad_tc(args, (cudaStream_t)stream);
}

int pmn_extract_grouped(
    const float* dense, float* gw, int dim0, int dim1, int taps, int groups, void* stream) {
    return launch_extract_grouped(dense, gw, dim0, dim1, taps, groups, (cudaStream_t)stream);
}

int pmn_prepare_weights(const pmn_weight_desc* table, int layers, int max_dim0, void* stream) {
    return launch_prepare_weights(table, layers, max_dim0, (cudaStream_t)stream);
}

int pmn_transpose_weight(
    const float* w, float* wt, int dim0, int dim1, int taps, void* stream) {
    return launch_transpose_weight(w, wt, dim0, dim1, taps, (cudaStream_t)stream);
}

int pmn_weight_norm_backward(
    const float* v, const float* g, const float* gw, float* gv, float* gg, int dim0, int inner,
    void* stream) {
    return launch_weight_norm_backward(v, g, gw, gv, gg, dim0, inner, (cudaStream_t)stream);
}

int pmn_weight_norm_backward_table(
    const pmn_weight_norm_desc* table, int layers, int max_dim0, void* stream) {
    return launch_weight_norm_backward_table(table, layers, max_dim0, (cudaStream_t)stream);
}

int pmn_reflect_pad(
    const float* x, float* out, int rows, int t_in, int left, int right, void* stream) {
    return launch_reflect_pad(x, out, rows, t_in, left, right, (cudaStream_t)stream);
}

int pmn_reflect_pad_backward(
    const float* gout, float* gx, int rows, int t_in, int left, int right, int accumulate,
    void* stream) {
    return launch_reflect_pad_backward(gout, gx, rows, t_in, left, right, accumulate, (cudaStream_t)stream);
}

int pmn_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream) {
    return launch_axpby(a, x, b, y, n, (cudaStream_t)stream);
}

int pmn_mse_to_target(
    const float* x, int64_t n, float target, float weight, float* loss, float* grad, void* stream) {
    return launch_mse_to_target(x, n, target, weight, loss, grad, (cudaStream_t)stream);
}

int pmn_l1_mean(
    const float* fake, const float* real, int64_t n, float weight, float* loss, float* gfake,
    int accumulate, void* stream) {
    return launch_l1_mean(fake, real, n, weight, loss, gfake, accumulate, (cudaStream_t)stream);
}

int pmn_adamw(
    float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
    float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
    const float* step_device, void* stream) {
    return launch_adamw(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
        step_device, (cudaStream_t)stream);
}

int pmn_adamw_peer(
    const float* const* grad_peers, float* const* param_peers, int world, int rank, float* exp_avg,
    float* exp_avg_sq, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps,
    float weight_decay, int step, const float* step_device, void* stream) {
    return launch_adamw_peer(
        grad_peers, param_peers, world, rank, exp_avg, exp_avg_sq, begin, end, lr, beta1, beta2, eps,
        weight_decay, step, step_device, (cudaStream_t)stream);
}

int pmn_row_sum(const float* x, float* out, int rows, int cols, int accumulate, void* stream) {
    return launch_row_sum(x, out, rows, cols, accumulate, (cudaStream_t)stream);
}

int pmn_features(
    const float* loudness, int loudness_rows, const float* pitch, const float* periodicity,
    const float* ppg, const float* pitch_distribution, const float* pitch_embedding,
    float ppg_threshold, float* features, int batch, int frames, void* stream) {
    PMN_REQUIRE(pitch_distribution && pitch_embedding, "features: null table");
    return launch_features(
        loudness, loudness_rows, pitch, periodicity, ppg, pitch_distribution, pitch_embedding,
        ppg_threshold, false, features, batch, frames, (cudaStream_t)stream);
}

int pmn_pitch_bins(
    const float* pitch, const float* edges, int64_t* bins, int n, int num_edges,
    float fmin, float fmax, void* stream) {
    return launch_pitch_bins(pitch, edges, bins, n, num_edges, fmin, fmax, (cudaStream_t)stream);
}

int pmn_embedding_backward(
    const float* gout, const int64_t* index, float* gtable, int batch, int channels, int frames,
    int rows, int out_channels, int channel_offset, void* stream) {
    return launch_embedding_backward(
        gout, index, gtable, batch, channels, frames, rows, out_channels, channel_offset,
        (cudaStream_t)stream);
}

int pmn_global_features(
    const float* speaker_embedding, const int64_t* speakers, const float* spectral_balance_ratios,
    const float* loudness_ratios, float* out, int batch, int speaker_channels, int num_speakers,
    void* stream) {
    return launch_global_features(
        speaker_embedding, speakers, spectral_balance_ratios, loudness_ratios, out, batch,
        speaker_channels, num_speakers, (cudaStream_t)stream);
}

int pmn_stft_magnitude(
    const float* audio, int batch, int samples, int window_kind, float eps, int layout,
    float* spectrum, float* magnitude, void* stream) {
    return launch_stft_train(
        audio, batch, samples, window_kind, eps, layout, spectrum, magnitude, (cudaStream_t)stream);
}

int pmn_stft_magnitude_backward(
    const float* gmagnitude, const float* spectrum, int batch, int samples, int window_kind,
    float eps, int layout, float* gaudio, int accumulate, void* stream) {
    return launch_stft_train_backward(
        gmagnitude, spectrum, batch, samples, window_kind, eps, layout, gaudio, accumulate,
        (cudaStream_t)stream);
}

int pmn_mel_loss(
    const float* magnitude, const float* target_mels, int batch, int frames, float loss_weight,
    float grad_weight, float* loss, float* gmagnitude, void* stream) {
    return launch_mel_loss(
        magnitude, target_mels, batch, frames, loss_weight, grad_weight, loss, gmagnitude,
        (cudaStream_t)stream);
}

int pmn_channel_sum(
    const float* x, float* out, int batch, int channels, int inner, int accumulate, void* stream) {
    return launch_channel_sum(x, out, batch, channels, inner, accumulate, (cudaStream_t)stream);
}

int pmn_copy_columns(
    const float* src, int src_width, int src_offset, float* dst, int dst_width, int dst_offset,
    int64_t rows, int cols, int accumulate, void* stream) {
    return launch_copy_columns(
        src, src_width, src_offset, dst, dst_width, dst_offset, rows, cols, accumulate,
        (cudaStream_t)stream);
}

int pmn_dft_basis(float* out, int n_fft, void* stream) {
    return launch_dft_basis(out, n_fft, (cudaStream_t)stream);
}

int pmn_spectral_convergence(
    const float* spec, int batch, int bins, int frames, float weight, float* sums, float* loss,
    float* gspec, void* stream) {
    return launch_spectral_convergence(
        spec, batch, bins, frames, weight, sums, loss, gspec, (cudaStream_t)stream);
}

int pmn_frame_overlap_add(
    const float* gframes, float* gsignal, int batch, int n_fft, int frames, int hop, int samples,
    void* stream) {
    return launch_frame_overlap_add(gframes, gsignal, batch, n_fft, frames, hop, samples, (cudaStream_t)stream);
}

int pmn_grid_sample(
    const float* sequence, const float* grid, float* out, int items, int channels, int t_in,
    int t_out, int nearest, int renormalize, void* stream) {
    return launch_grid_sample(
        sequence, grid, out, items, channels, t_in, t_out, nearest != 0, renormalize != 0,
        (cudaStream_t)stream);
}

}  // extern "C"

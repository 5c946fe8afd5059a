// Launchers of the HBM-bound generator kernels (features.cu)
#pragma once

#include "common.cuh"

namespace pmn {

int launch_features(
    const float* loudness, int rows, const float* pitch, const float* periodicity,
    const float* ppg, const float* pitch_distribution, const float* pitch_embedding,
    float threshold, bool with_period, float* out, int batch, int frames,
    cudaStream_t stream);

int launch_speaker_bias(
    const float* speaker_embedding, const int64_t* speakers, const float* sbr,
    const float* lr, const float* weight, const float* bias, float* out,
    int batch, int speaker_channels, int c_out, int num_speakers, cudaStream_t stream);

int launch_head(
    const float* x, const float* weight, float* out, int batch, int channels,
    int t_len, float slope, cudaStream_t stream);

int launch_grid_sample(
    const float* sequence, const float* grid, float* out, int items, int channels, int t_in,
    int t_out, bool nearest, bool renormalize, cudaStream_t stream);

}  // namespace pmn

// Generator handle entry points (generator.cu)
#pragma once

#include "common.cuh"

namespace pmn {

pmn_generator* generator_create();
void generator_destroy(pmn_generator* g);
int generator_set_tensor(
    pmn_generator* g, const char* name, const float* data, const int64_t* shape, int ndim,
    cudaStream_t stream);
int generator_finalize(pmn_generator* g, int math, cudaStream_t stream);
int generator_set_pair_mask(pmn_generator* g, unsigned mask);
int generator_set_f8(pmn_generator* g, bool enabled);
size_t generator_workspace_bytes(const pmn_generator* g, int batch, int frames);
int generator_features(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, float* features, int batch, int frames,
    cudaStream_t stream);
int generator_forward(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, const int64_t* speakers,
    const float* sbr, const float* lr, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace pmn

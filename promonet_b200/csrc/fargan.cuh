// FARGAN generator handle (fargan.cu)
#pragma once

#include "common.cuh"

struct pmn_fargan;

namespace pmn {

pmn_fargan* fargan_create();
void fargan_destroy(pmn_fargan* g);
int fargan_set_tensor(pmn_fargan* g, const char* name, const float* data, const int64_t* shape,
                      int ndim, cudaStream_t stream);
int fargan_finalize(pmn_fargan* g, cudaStream_t stream);
size_t fargan_workspace_bytes(int batch, int frames);
int fargan_forward(
    pmn_fargan* g, const float* loudness, int rows, const float* pitch, const float* periodicity,
    const float* ppg, const int64_t* speakers, const float* sbr, const float* lr,
    const float* previous_samples, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace pmn

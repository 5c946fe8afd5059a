// Named fp32 device tensors owned by a model handle (the state_dict of the
// reference module, loaded entry by entry through pmn_*_set_tensor)
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace pmn {

struct Tensor {
    float* data = nullptr;
    std::vector<int64_t> shape;
    size_t numel() const {
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        return n;
    }
};

struct TensorStore {
    std::map<std::string, Tensor> tensors;
    std::vector<void*> owned;  // extra allocations made while finalizing

    ~TensorStore() {
        for (auto& item : tensors) cudaFree(item.second.data);
        for (void* p : owned) cudaFree(p);
    }

    // Copy `data` (device, fp32) into library-owned memory under `name`
    int set(const char* name, const float* data, const int64_t* shape, int ndim, cudaStream_t stream) {
        Tensor t;
        for (int i = 0; i < ndim; ++i) {
            if (shape[i] <= 0)
                return fail(PMN_ERR_ARGUMENT, std::string("set_tensor: empty dimension in ") + name);
            t.shape.push_back(shape[i]);
        }
        const size_t bytes = t.numel() * sizeof(float);
        PMN_TRY(check_cuda(cudaMalloc(&t.data, bytes), "cudaMalloc"));
        const int status = check_cuda(
            cudaMemcpyAsync(t.data, data, bytes, cudaMemcpyDeviceToDevice, stream), "set_tensor copy");
        if (status != PMN_OK) {
            cudaFree(t.data);
            return status;
        }
        auto old = tensors.find(name);
        if (old != tensors.end()) {
            cudaFree(old->second.data);
            tensors.erase(old);
        }
        tensors.emplace(name, std::move(t));
        return PMN_OK;
    }

    int find(const std::string& name, const Tensor** out) const {
        auto it = tensors.find(name);
        if (it == tensors.end()) return fail(PMN_ERR_STATE, "missing tensor: " + name);
        *out = &it->second;
        return PMN_OK;
    }

    bool has(const std::string& name) const { return tensors.count(name) != 0; }
    const float* data(const std::string& name) const { return tensors.at(name).data; }

    int alloc(size_t floats, float** out) {
        PMN_TRY(check_cuda(cudaMalloc(out, floats * sizeof(float)), "cudaMalloc"));
        owned.push_back(*out);
        return PMN_OK;
    }
};

}  // namespace pmn

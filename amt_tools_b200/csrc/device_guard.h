// Makes a device current for the duration of a scope and restores the caller's device afterwards: creating, using or
// destroying a plan / resampler / pipeline for cuda:k must not move the calling thread to cuda:k.
#pragma once

#include <cuda_runtime.h>

namespace amtfeat {

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t status = cudaSuccess;   // of switching to `device` (an invalid ordinal is reported here)
    explicit DeviceGuard(int device) {
        if (device < 0) return;
        status = cudaGetDevice(&prev);
        if (status != cudaSuccess || prev == device) return;
        status = cudaSetDevice(device);
        changed = status == cudaSuccess;
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

}  // namespace amtfeat

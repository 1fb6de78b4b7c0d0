// Shared plumbing for the C-ABI implementation: error reporting, CUDA checks, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "errors.h"

namespace mz {

int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MZ_CUDA(call)                                                           \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return mz::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// true if p points to device (or managed) memory of the current context
bool is_device_ptr(const void *p);

// RAII device buffer (never throws; check ok())
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        return alloc(count);
    }
    cudaError_t upload(const T *src, size_t count, cudaStream_t s = 0)
    {
        cudaError_t e = ensure(count);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyDefault, s);
    }
};

// pinned host staging buffer
template <class T>
struct PinBuf {
    T *p = nullptr;
    size_t n = 0;
    ~PinBuf() { release(); }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocHost((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

}  // namespace mz

// abi_types.h -- the part of the library internals that does not depend on the kernel headers: error helper, CUDA call
// check, the struct behind svo_mem_t.  (abi_internal.h adds the context; svo_builder.cu needs only this.)
#pragma once
#include "../../include/svo_b200.h"
#include <cuda_runtime.h>

void svo_fail(int code, const char *fmt, ...);

#define CU_CHECK(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) svo_fail((int)_e, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
    } while (0)

struct svo_mem_s {
    void *dptr;
    size_t bytes;
    int device;
    svo_ctx_t ctx = nullptr;        // the context the buffer was allocated in (svo_free drains that one, not the current one)
};


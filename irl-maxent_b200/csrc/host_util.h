// host_util.h -- error reporting and cached device workspaces (host side).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/irl_maxent_b200.h"

namespace irlb200 {

// record a message for irlb200_last_error() and return `code`
int fail(int code, const char *msg);
int fail_cuda(cudaError_t e, const char *what);
int device_count_impl();

// Cached scratch allocations, one block per (device, slot, stream), grown on demand and reused by later
// calls on the same stream (slot 0: predecessor weights of the streamed forward pass, slot 1: iterate
// buffers + barrier state of the cooperative kernels).  Launches on one stream are ordered, so reuse
// within a stream is safe; different streams get different blocks.
int workspace(int slot, size_t bytes, void **out, cudaStream_t stream);

}  // namespace irlb200

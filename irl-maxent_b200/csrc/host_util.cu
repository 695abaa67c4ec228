// host_util.cu -- error strings, device probing, workspace cache.
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "host_util.h"

namespace irlb200 {

static thread_local char g_err[512] = "";

int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return IRLB200_ECUDA;
}

int device_count_impl() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

namespace {
constexpr int kSlots = 4, kMaxDev = 16;
struct Slot { void *ptr = nullptr; size_t bytes = 0; };
// one block per (device, slot, STREAM): calls issued on different streams (or from different host threads
// on their own streams) never share barrier words, iterate buffers or weight scratch
std::map<std::tuple<int, int, cudaStream_t>, Slot> g_ws;
std::mutex g_mu;
}  // namespace

int workspace(int slot, size_t bytes, void **out, cudaStream_t stream) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    if (dev >= kMaxDev || slot >= kSlots) return fail(IRLB200_EINVAL, "workspace: bad slot/device");
    std::lock_guard<std::mutex> lk(g_mu);
    Slot &s = g_ws[std::make_tuple(dev, slot, stream)];
    if (s.bytes < bytes) {
        if (s.ptr) {
            // earlier launches on this stream may still use the old block: drain that stream only
            cudaStreamSynchronize(stream);
            cudaFree(s.ptr);
            s.ptr = nullptr;
            s.bytes = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        e = cudaMalloc(&s.ptr, want);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(workspace)");
        s.bytes = want;
    }
    *out = s.ptr;
    return IRLB200_OK;
}

}  // namespace irlb200

extern "C" int irlb200_version(void) { return 100; }
extern "C" const char *irlb200_last_error(void) { return irlb200::g_err; }
extern "C" int irlb200_device_count(void) { return irlb200::device_count_impl(); }

// api_common.cu -- library identity, error text and the launch counter of the C ABI.
#include <atomic>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"

namespace {
std::atomic<uint64_t> g_launches{0};
thread_local char g_err[256] = "";
}  // namespace

namespace acb {
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}
int cuda_fail() {
    cudaError_t e = cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s", e == cudaSuccess ? "unknown CUDA failure" : cudaGetErrorString(e));
    return AC_E_CUDA;
}
int launched() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString(e));
        return AC_E_CUDA;
    }
    return AC_OK;
}
}  // namespace acb

extern "C" {
const char* ac_version(void) { return "avatarcraft_b200 0.1 (sm_100a)"; }
const char* ac_last_cuda_error(void) { return g_err; }
uint64_t ac_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void ac_launch_count_add(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}

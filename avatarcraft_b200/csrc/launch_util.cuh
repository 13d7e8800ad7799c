// launch_util.cuh -- launch bookkeeping shared by the translation units of libavatarcraft_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acb {
int sm_count();          // SMs of the current device (148 on B200), cached
int launched();          // counts one kernel launch; returns AC_OK or AC_E_CUDA (cudaGetLastError)
int cuda_fail();         // records cudaGetLastError() text, returns AC_E_CUDA
}  // namespace acb

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device): function attributes are per
// context, so a process that drives several GPUs must set them on each.
#define ACB_SET_MAX_SMEM(func, bytes)                                                                          \
    do {                                                                                                       \
        static bool acb_done_[64] = {};                                                                        \
        int acb_dev_ = 0;                                                                                      \
        if (cudaGetDevice(&acb_dev_) == cudaSuccess && acb_dev_ >= 0 && acb_dev_ < 64 && !acb_done_[acb_dev_]) { \
            cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));             \
            acb_done_[acb_dev_] = true;                                                                        \
        }                                                                                                      \
    } while (0)

// launch_util.cuh -- launch bookkeeping shared by the translation units of libavatarcraft_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acb {
int sm_count();          // SMs of the current device (148 on B200), cached
int launched();          // counts one kernel launch; returns AC_OK or AC_E_CUDA (cudaGetLastError)
int cuda_fail();         // records cudaGetLastError() text, returns AC_E_CUDA
}  // namespace acb

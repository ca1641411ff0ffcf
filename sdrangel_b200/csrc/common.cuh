// common.cuh — shared host-side plumbing of libb200dsp: error reporting, device selection.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include "../../include/b200dsp.h"

namespace b200dsp {

int  b200_fail(int code, const char* fmt, ...);          // records the thread-local message, returns code
int  b200_cuda_check(cudaError_t e, const char* what, const char* file, int line);
int  b200_require_device();                               // 0 or B200DSP_ENODEV
int  b200_current_device();
int  b200_sm_count_of(int device);

#define B200_CUDA_CHECK(expr) ::b200dsp::b200_cuda_check((expr), #expr, __FILE__, __LINE__)

} // namespace b200dsp

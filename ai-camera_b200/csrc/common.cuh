// Shared helpers for the aicam CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/aicam.h"

namespace aicam {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define AICAM_CUDA_OK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::aicam::fail(AICAM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
  } while (0)

#define AICAM_CHECK_ARG(cond, msg)                                       \
  do {                                                                   \
    if (!(cond)) return ::aicam::fail(AICAM_ERR_INVALID_ARG, (msg));      \
  } while (0)

inline int last_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(AICAM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return AICAM_OK;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  cudaFuncSetAttribute is per device and
// overwrites the previous limit, so the library keeps a running maximum per (device, kernel) and only ever raises it:
// a second engine / tracker (smaller, or on another GPU of the same process) never lowers or misses the opt-in.
int ensure_dynamic_smem(const void* kernel, size_t bytes);
template <typename Fn>
inline int ensure_dynamic_smem(Fn* kernel, size_t bytes) { return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes); }
// SM count of the current device (cached per device)
int current_num_sms();

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace aicam

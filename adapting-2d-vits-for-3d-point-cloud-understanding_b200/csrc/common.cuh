// common.cuh - shared helpers for the p3tok kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "p3tok.h"

namespace p3tok {

// thread-local error string surfaced by p3tok_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);   // process-wide diagnostic counter (p3tok_kernel_launches)

#define P3_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::p3tok::set_error(__VA_ARGS__);      \
      return (code);                        \
    }                                       \
  } while (0)

#define P3_CUDA(expr)                                            \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return ::p3tok::cuda_fail(_e, #expr); \
  } while (0)

#define P3_LAUNCH_CHECK(name)                                       \
  do {                                                              \
    cudaError_t _e = cudaPeekAtLastError();   /* do not clear errors that belong to earlier, unrelated work */ \
    if (_e != cudaSuccess) return ::p3tok::cuda_fail(_e, name);     \
    ::p3tok::count_launch();                                        \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// float -> uint32 whose unsigned order equals the float order (finite values, -0 canonicalised away)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}

// |v|^2 = ((x*x)+(y*y))+(z*z), each op rounded on its own (torch.sum(v**2,-1) on 3 elements)
__device__ __forceinline__ float sq3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

}  // namespace p3tok

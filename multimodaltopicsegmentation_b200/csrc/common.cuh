// Shared helpers for the sm_100a kernels of libmts_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mts_b200.h"

namespace mts {

void set_error(const char *msg);  // defined in abi.cu (thread-local message for mts_last_error)

#define MTS_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) {                   \
      ::mts::set_error(msg);         \
      return (code);                 \
    }                                \
  } while (0)

#define MTS_LAUNCH_CHECK()                      \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) {                   \
      ::mts::set_error(cudaGetErrorString(e__)); \
      return (int)e__;                          \
    }                                           \
  } while (0)

#define MTS_CUDA(call)                          \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) {                   \
      ::mts::set_error(cudaGetErrorString(e__)); \
      return (int)e__;                          \
    }                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The reference evaluates torch.sigmoid in fp32; this is the same expression, IEEE division, accurate expf.
__device__ __forceinline__ float sigmoid_ref(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }

// fp32 -> nearest TF32-representable fp32 (low 13 mantissa bits zero), so that what tcgen05 kind::tf32 reads
// (it ignores those bits) is exactly the value we computed.  x = hi + lo + O(2^-23 |x|) with
// hi = tf32_rn(x), lo = tf32_rn(x - hi).
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace mts

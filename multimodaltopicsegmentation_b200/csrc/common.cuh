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

// Launch-site state that has to exist once per DEVICE: cudaFuncSetAttribute and the occupancy answers are per device, so a
// process that drives several GPUs must configure each of them (the package's model is one process per GPU, where slot 0 of
// these arrays is the only one used).  `MTS_PER_DEVICE(int, cap);` declares `cap` as a reference to the current device's slot
// of a function-static, zero-initialised array.
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
#define MTS_PER_DEVICE(type, name)                                  \
  static type name##_per_device[::mts::kMaxDevices] = {};           \
  type &name = name##_per_device[::mts::device_slot()]

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

// ---------------------------------------------------------------------------------------------------------
// Operand format of mts_gemm_tf32x3 (error-compensated TF32 GEMM: one TF32 product + one BF16 correction product).
//   x = hi + rest,  hi = tf32_trunc(x) = what kind::tf32 reads of the raw fp32 word,  rest = x - hi (exact, <= 13 bits)
//   A B^T ~= hi(A) hi(B)^T  [kind::tf32 on the RAW fp32 operands]
//          + bf16(A) bf16(rest B)^T + bf16(rest A) bf16(B)^T   [kind::f16/bf16: the correction terms carry 2^-11 of the
//            magnitude, so 8-bit mantissas leave a relative error of ~2^-19; bf16 runs K = 16 per MMA: 2 instead of 3
//            instruction streams per product]
//   "hi" array : fp32 [rows, Kp], the values themselves, zero beyond the logical width
//   "lo" array : same byte size, holding bf16 [rows, 2 Kp]: per 16-wide K block 32 values (64 bytes) --
//                A operand (side 0): [ bf16(x_k), k = 0..15 | bf16(rest x_k), k = 0..15 ],  B operand (side 1): halves swapped,
//                so that one K = 32 bf16 block of A times the same block of B (two K = 16 MMAs) sums both correction terms.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rest_exact(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t bf16x2_bits(float first, float second) {  // `first` at the lower address
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
  return r;
}
__device__ __forceinline__ uint16_t bf16_bits(float x) { return (uint16_t)(bf16x2_bits(x, 0.0f) & 0xFFFFu); }
// row = start of one operand row of the "lo" array (Kp floats = 2 Kp bf16); k = logical K index
__device__ __forceinline__ void corr_store1(float *row, int k, float x, int side) {
  uint16_t *b = reinterpret_cast<uint16_t *>(row) + (k >> 4) * 32 + (k & 15);
  b[side ? 16 : 0] = bf16_bits(x);
  b[side ? 0 : 16] = bf16_bits(tf32_rest_exact(x));
}
__device__ __forceinline__ void corr_store2(float *row, int k, float2 v, int side) {  // k even
  uint16_t *b = reinterpret_cast<uint16_t *>(row) + (k >> 4) * 32 + (k & 15);
  *reinterpret_cast<uint32_t *>(b + (side ? 16 : 0)) = bf16x2_bits(v.x, v.y);
  *reinterpret_cast<uint32_t *>(b + (side ? 0 : 16)) = bf16x2_bits(tf32_rest_exact(v.x), tf32_rest_exact(v.y));
}
__device__ __forceinline__ void corr_store4(float *row, int k, float4 v, int side) {  // k a multiple of 4
  uint16_t *b = reinterpret_cast<uint16_t *>(row) + (k >> 4) * 32 + (k & 15);
  *reinterpret_cast<uint2 *>(b + (side ? 16 : 0)) = make_uint2(bf16x2_bits(v.x, v.y), bf16x2_bits(v.z, v.w));
  *reinterpret_cast<uint2 *>(b + (side ? 0 : 16)) =
      make_uint2(bf16x2_bits(tf32_rest_exact(v.x), tf32_rest_exact(v.y)), bf16x2_bits(tf32_rest_exact(v.z), tf32_rest_exact(v.w)));
}

__device__ __forceinline__ void corr_store8(float *row, int k, float4 a, float4 b, int side) {  // k a multiple of 8
  uint16_t *p = reinterpret_cast<uint16_t *>(row) + (k >> 4) * 32 + (k & 15);
  *reinterpret_cast<uint4 *>(p + (side ? 16 : 0)) =
      make_uint4(bf16x2_bits(a.x, a.y), bf16x2_bits(a.z, a.w), bf16x2_bits(b.x, b.y), bf16x2_bits(b.z, b.w));
  *reinterpret_cast<uint4 *>(p + (side ? 0 : 16)) =
      make_uint4(bf16x2_bits(tf32_rest_exact(a.x), tf32_rest_exact(a.y)), bf16x2_bits(tf32_rest_exact(a.z), tf32_rest_exact(a.w)),
                 bf16x2_bits(tf32_rest_exact(b.x), tf32_rest_exact(b.y)), bf16x2_bits(tf32_rest_exact(b.z), tf32_rest_exact(b.w)));
}

// ---------------------------------------------------------------------------------------------------------
// Dropout on the attention probabilities (HF LongformerSelfAttention: nn.functional.dropout(attn_probs, p), inverted
// scaling).  The keep decision of probability (episode b, head, query i, key j) is a pure function of a 64-bit seed and
// those four indices -- a splitmix64 finaliser over  seed ^ (bh << 40 | i << 20 | j)  -- so the forward kernel and both
// backward kernels regenerate the same mask from the seed and no mask tensor exists.  u = top 24 bits / 2^24 is kept
// when u >= p.  The parity tests restate this hash in numpy on the checker side.  S <= 2^20.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float attn_keep_scale(uint64_t seed, uint32_t bh, uint32_t i, uint32_t j, uint32_t p24, float inv_keep) {
  uint64_t x = seed ^ (((uint64_t)bh << 40) | ((uint64_t)i << 20) | (uint64_t)j);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return ((uint32_t)(x >> 40) >= p24) ? inv_keep : 0.0f;
}

// host side of attn_keep_scale: the 24-bit threshold of a drop probability
inline uint32_t attn_drop_p24(float p) { return (uint32_t)((double)p * 16777216.0 + 0.5); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace mts

// Operand preparation for the 3xTF32 tensor-core GEMMs: fused early-fusion concat + time crop + hi/lo split.
// The reference materialises the concat on the host at load time (utils/load_datasets_precomputed.py:158-161);
// here the two modality tensors stay separate and the concat happens while the GEMM operand is written.
// Pure streaming kernels: 4(D1+D2) bytes read, 8 Kp bytes written per sentence.
#include <cuda_fp16.h>

#include "common.cuh"

namespace mts {

__global__ void __launch_bounds__(256) pack_rows_split_kernel(const float *__restrict__ src1, int64_t bstride1, int D1,
                                                              const float *__restrict__ src2, int64_t bstride2, int D2,
                                                              int B, int T, int Kp, float *__restrict__ hi,
                                                              float *__restrict__ lo) {
  const int64_t total = (int64_t)B * T * Kp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const int64_t row = idx / Kp;
    const int b = (int)(row / T), t = (int)(row % T);
    float v = 0.0f;
    if (k < D1) v = __ldg(src1 + (int64_t)b * bstride1 + (int64_t)t * D1 + k);
    else if (k < D1 + D2) v = __ldg(src2 + (int64_t)b * bstride2 + (int64_t)t * D2 + (k - D1));
    hi[idx] = v;                                // the tensor core reads the top 19 bits of the raw word
    corr_store1(lo + row * Kp, k, v, 0);        // bf16 correction operand (A side), see common.cuh
  }
}

// The same, 8 consecutive K elements per thread (two 16-byte loads, two 16-byte `hi` stores, two 16-byte halves of the
// correction operand): needs D1, D2 multiples of 8 and 16-byte aligned rows.  src2 may be NULL (D2 = 0).
__global__ void __launch_bounds__(256) pack_rows_split_v8_kernel(const float *__restrict__ src1, int64_t bstride1, int D1,
                                                                 const float *__restrict__ src2, int64_t bstride2, int D2,
                                                                 int B, int T, int Kp, float *__restrict__ hi,
                                                                 float *__restrict__ lo) {
  const int k8n = Kp >> 3;
  const int64_t total = (int64_t)B * T * k8n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % k8n) << 3;
    const int64_t row = idx / k8n;
    const int b = (int)(row / T), t = (int)(row % T);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    const float *p = nullptr;
    if (k < D1) p = src1 + (int64_t)b * bstride1 + (int64_t)t * D1 + k;
    else if (k < D1 + D2) p = src2 + (int64_t)b * bstride2 + (int64_t)t * D2 + (k - D1);
    if (p) {
      v0 = __ldcs(reinterpret_cast<const float4 *>(p));      // streamed once: do not keep in L2 ahead of the GEMM operands
      v1 = __ldcs(reinterpret_cast<const float4 *>(p) + 1);
    }
    if (hi) {   // (bf16 path: only the packed operand is wanted)
      float4 *h = reinterpret_cast<float4 *>(hi + row * Kp + k);
      h[0] = v0;
      h[1] = v1;
    }
    corr_store8(lo + row * Kp, k, v0, v1, 0);
  }
}

// bf16 sources (embeddings stored and transferred as bf16: half the host->device bytes of the path): the packed operand only
__global__ void __launch_bounds__(256) pack_rows_bf16in_kernel(const uint16_t *__restrict__ src1, int64_t bstride1, int D1,
                                                               const uint16_t *__restrict__ src2, int64_t bstride2, int D2,
                                                               int B, int T, int Kp, float *__restrict__ lo) {
  const int k8n = Kp >> 3;
  const int64_t total = (int64_t)B * T * k8n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % k8n) << 3;
    const int64_t row = idx / k8n;
    const int b = (int)(row / T), t = (int)(row % T);
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    const uint16_t *p = nullptr;
    if (k < D1) p = src1 + (int64_t)b * bstride1 + (int64_t)t * D1 + k;
    else if (k < D1 + D2) p = src2 + (int64_t)b * bstride2 + (int64_t)t * D2 + (k - D1);
    if (p) raw = __ldcs(reinterpret_cast<const uint4 *>(p));
    // a bf16 value is exact in TF32, so its remainder is zero: [bf16(x) x8 | 0 x8] at this 8-column half of the block
    uint16_t *dst = reinterpret_cast<uint16_t *>(lo + row * Kp) + (k >> 4) * 32 + (k & 15);
    *reinterpret_cast<uint4 *>(dst) = raw;
    *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void __launch_bounds__(256) split_tf32_v8_kernel(const float *__restrict__ src, int64_t ld, int rows, int cols,
                                                            int Kp, int side, float *__restrict__ hi,
                                                            float *__restrict__ lo) {
  const int k8n = Kp >> 3;
  const int64_t total = (int64_t)rows * k8n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % k8n) << 3;
    const int64_t r = idx / k8n;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (k < cols) {  // cols % 8 == 0: a group is entirely inside or entirely padding
      v0 = __ldg(reinterpret_cast<const float4 *>(src + r * ld + k));
      v1 = __ldg(reinterpret_cast<const float4 *>(src + r * ld + k) + 1);
    }
    float4 *h = reinterpret_cast<float4 *>(hi + r * Kp + k);
    h[0] = v0;
    h[1] = v1;
    corr_store8(lo + r * Kp, k, v0, v1, side);
  }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, int64_t ld, int rows, int cols,
                                                         int Kp, int side, float *__restrict__ hi, float *__restrict__ lo) {
  const int64_t total = (int64_t)rows * Kp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const int64_t r = idx / Kp;
    const float v = (k < cols) ? __ldg(src + r * ld + k) : 0.0f;
    hi[idx] = v;
    corr_store1(lo + r * Kp, k, v, side);
  }
}

// Transposed operand preparation for the weight-gradient GEMMs (dW = dY^T X is a contraction over TOKENS, so both
// operands must be K-major along the token axis for the tensor-core kernel):
//   out[c, r] = src[(r / T) * bstride + (r % T + shift) * ld + c]   for 0 <= r % T + shift < lengths[r / T],  else 0
// written as (hi, lo) TF32 halves [cols, Kp], columns r >= rows zero-filled.  shift = -1 / +1 yields the h_{t-1} /
// h_{t+1} operand of dW_hh without a shifted copy of the hidden states.  64 x 32 tiles through shared memory:
// coalesced reads along the source columns, coalesced writes along the token axis.
__global__ void __launch_bounds__(256) transpose_split_kernel(const float *__restrict__ src, int64_t bstride, int64_t ld,
                                                             int rows, int cols, int T, int shift,
                                                             const int32_t *__restrict__ lengths, int Kp, int side,
                                                             float *__restrict__ hi, float *__restrict__ lo) {
  // tile = 64 tokens (r) x 32 source columns (c).  Read: a warp takes 32 consecutive columns of one token (128
  // contiguous bytes).  Write: a thread takes 8 consecutive tokens of one column: two 16-byte stores of the fp32 operand
  // and two 16-byte halves of the packed correction operand.
  __shared__ float tile[64][33];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.0f;
    if (r < rows && c < cols) {
      const int b = r / T, t = r % T + shift;
      const int len = lengths ? lengths[b] : T;
      if (t >= 0 && t < len && t < T) v = __ldg(src + (int64_t)b * bstride + (int64_t)t * ld + c);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
  const int c = c0 + (threadIdx.x >> 3), rr = (threadIdx.x & 7) * 8, r = r0 + rr;
  if (c < cols && r < Kp) {  // Kp % 32 == 0 and r % 8 == 0: a group of 8 tokens is entirely inside [0, Kp) or outside
    const int cl = threadIdx.x >> 3;
    const float4 v0 = make_float4(tile[rr][cl], tile[rr + 1][cl], tile[rr + 2][cl], tile[rr + 3][cl]);
    const float4 v1 = make_float4(tile[rr + 4][cl], tile[rr + 5][cl], tile[rr + 6][cl], tile[rr + 7][cl]);
    float4 *h = reinterpret_cast<float4 *>(hi + (int64_t)c * Kp + r);
    h[0] = v0;
    h[1] = v1;
    corr_store8(lo + (int64_t)c * Kp, r, v0, v1, side);
  }
}

// Device-side collater (EncoderDataset.py:91-152 builds the padded batch on the host, every step): episodes live
// packed back to back in HBM; out[b, t, :] = t < lengths[ids[b]] ? src[offsets[ids[b]] + t, :] : pad.
__global__ void __launch_bounds__(256) gather_pad_kernel(const float *__restrict__ src, const int64_t *__restrict__ offsets,
                                                        const int32_t *__restrict__ lengths,
                                                        const int32_t *__restrict__ ids, int B, int T, int D, float pad,
                                                        float *__restrict__ out) {
  const int64_t total = (int64_t)B * T * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const int64_t row = idx / D;
    const int b = (int)(row / T), t = (int)(row % T);
    const int e = ids[b];
    out[idx] = (t < lengths[e]) ? __ldg(src + (offsets[e] + t) * D + c) : pad;
  }
}

// token rows between the padded [B*S, d] layout and the ragged [sum(len), d] layout (offsets int32 [B])
__global__ void __launch_bounds__(256) ragged_copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                          const int32_t *__restrict__ lengths,
                                                          const int32_t *__restrict__ offsets, int B, int S, int nv,
                                                          int to_padded, float pad) {
  const int64_t total = (int64_t)B * S * nv;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % nv);
    const int64_t row = idx / nv;
    const int b = (int)(row / S), t = (int)(row % S);
    const bool valid = t < lengths[b];
    const int64_t rrow = (int64_t)offsets[b] + t;
    if (to_padded) dst[idx] = valid ? __ldg(src + rrow * nv + c) : make_float4(pad, pad, pad, pad);
    else if (valid) dst[rrow * nv + c] = __ldg(src + idx);
  }
}

}  // namespace mts

using namespace mts;

static unsigned grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int mts_pack_rows_split(const float *src1, int64_t bstride1, int D1, const float *src2, int64_t bstride2,
                                   int D2, int B, int T, int Kp, float *hi, float *lo, void *stream) {
  MTS_REQUIRE(src1 && lo, MTS_E_BADARG, "pack_rows_split: null pointer");
  MTS_REQUIRE(D2 == 0 || src2, MTS_E_BADARG, "pack_rows_split: D2 > 0 without src2");
  MTS_REQUIRE(B > 0 && T > 0 && D1 > 0 && D2 >= 0, MTS_E_BADARG, "pack_rows_split: bad shape");
  MTS_REQUIRE(Kp % 32 == 0 && Kp >= D1 + D2, MTS_E_BADARG, "pack_rows_split: Kp must be a multiple of 32 and >= D1 + D2");
  const bool al16 = ((((uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0) &&
                    bstride1 % 4 == 0 && bstride2 % 4 == 0;
  MTS_REQUIRE(hi || (al16 && D1 % 8 == 0 && D2 % 8 == 0), MTS_E_UNSUPPORTED,
              "pack_rows_split: the packed-operand-only form needs widths that are multiples of 8 and 16-byte aligned rows");
  if (al16 && D1 % 8 == 0 && D2 % 8 == 0) {
    pack_rows_split_v8_kernel<<<grid_for((int64_t)B * T * (Kp / 8)), 256, 0, (cudaStream_t)stream>>>(
        src1, bstride1, D1, src2, bstride2, D2, B, T, Kp, hi, lo);
  } else {
    pack_rows_split_kernel<<<grid_for((int64_t)B * T * Kp), 256, 0, (cudaStream_t)stream>>>(src1, bstride1, D1, src2,
                                                                                           bstride2, D2, B, T, Kp, hi, lo);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

// Early-fusion concat + crop straight into the FP16-SPLIT operand of mts_gemm_f16x3: a warp owns one (episode, sentence) row of
// [x1 | x2], finds its largest magnitude, scales by the exact power of two that puts it into [2^13, 2^14) and writes
// pieces [row][2][K64] (fp16(x s), fp16(x s - piece 1), zero beyond D1 + D2) and row_scale [row] = 1 / s.  D1, D2 % 4 == 0,
// D1 + D2 <= 2048.
constexpr int PF_MAXV = 16;   // float4 per lane: D1 + D2 <= 2048 (8 serves widths <= 1024 with half the registers)
template <int MAXV>
__global__ void __launch_bounds__(256) pack_rows_f16_kernel(const float *__restrict__ src1, int64_t bstride1, int D1,
                                                            const float *__restrict__ src2, int64_t bstride2, int D2, int B, int T,
                                                            int K64, __half *__restrict__ pieces, float *__restrict__ row_scale) {
  const int lane = threadIdx.x & 31;
  const int nv1 = D1 >> 2, nv = (D1 + D2) >> 2;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int rows = B * T;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const int b = row / T, t = row % T;
    const float4 *p1 = reinterpret_cast<const float4 *>(src1 + (int64_t)b * bstride1 + (int64_t)t * D1);
    const float4 *p2 = src2 ? reinterpret_cast<const float4 *>(src2 + (int64_t)b * bstride2 + (int64_t)t * D2) : nullptr;
    float4 v[MAXV];
    float mx = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        v[i] = c < nv1 ? __ldg(p1 + c) : __ldg(p2 + (c - nv1));
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
      }
    }
    mx = warp_max(mx);
    const int ex = (int)((__float_as_uint(mx) >> 23) & 0xFFu);
    int sexp = ex == 0 ? 0 : 127 + 13 - ex;
    sexp = sexp > 110 ? 110 : (sexp < -110 ? -110 : sexp);
    const float sc = __uint_as_float((uint32_t)(127 + sexp) << 23);
    if (lane == 0) row_scale[row] = __uint_as_float((uint32_t)(127 - sexp) << 23);
    __half *q1 = pieces + (int64_t)row * 2 * K64, *q2 = q1 + K64;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 x = make_float4(v[i].x * sc, v[i].y * sc, v[i].z * sc, v[i].w * sc);
        const __half2 a0 = __floats2half2_rn(x.x, x.y), a1 = __floats2half2_rn(x.z, x.w);
        const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
        const __half2 b0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), b1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
        *reinterpret_cast<uint2 *>(q1 + 4 * c) = make_uint2(*reinterpret_cast<const uint32_t *>(&a0), *reinterpret_cast<const uint32_t *>(&a1));
        *reinterpret_cast<uint2 *>(q2 + 4 * c) = make_uint2(*reinterpret_cast<const uint32_t *>(&b0), *reinterpret_cast<const uint32_t *>(&b1));
      }
    }
    for (int c = D1 + D2 + lane; c < K64; c += 32) { q1[c] = __float2half(0.0f); q2[c] = __float2half(0.0f); }
  }
}

extern "C" int mts_pack_rows_f16(const float *src1, int64_t bstride1, int D1, const float *src2, int64_t bstride2, int D2, int B,
                                 int T, int K64, void *pieces, float *row_scale, void *stream) {
  MTS_REQUIRE(src1 && pieces && row_scale, MTS_E_BADARG, "pack_rows_f16: null pointer");
  MTS_REQUIRE(D2 == 0 || src2, MTS_E_BADARG, "pack_rows_f16: D2 > 0 without src2");
  MTS_REQUIRE(B > 0 && T > 0 && D1 > 0 && D2 >= 0, MTS_E_BADARG, "pack_rows_f16: bad shape");
  MTS_REQUIRE(K64 % 64 == 0 && K64 >= D1 + D2, MTS_E_BADARG, "pack_rows_f16: K64 must be a multiple of 64 and >= D1 + D2");
  MTS_REQUIRE(D1 % 4 == 0 && D2 % 4 == 0 && D1 + D2 <= 128 * PF_MAXV && bstride1 % 4 == 0 && bstride2 % 4 == 0 &&
                  ((((uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)pieces) & 15) == 0),
              MTS_E_UNSUPPORTED, "pack_rows_f16: widths must be multiples of 4 (sum <= 2048) and rows 16-byte aligned");
  const int64_t rows = (int64_t)B * T;
  const unsigned grid = (unsigned)((rows + 7) / 8 < (int64_t)kNumSMs * 16 ? (rows + 7) / 8 : (int64_t)kNumSMs * 16);
  if (D1 + D2 <= 1024)
    pack_rows_f16_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(src1, bstride1, D1, src2, bstride2, D2, B, T, K64,
                                                                     reinterpret_cast<__half *>(pieces), row_scale);
  else
    pack_rows_f16_kernel<PF_MAXV><<<grid, 256, 0, (cudaStream_t)stream>>>(src1, bstride1, D1, src2, bstride2, D2, B, T, K64,
                                                                           reinterpret_cast<__half *>(pieces), row_scale);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_pack_rows_bf16in(const void *src1, int64_t bstride1, int D1, const void *src2, int64_t bstride2, int D2,
                                    int B, int T, int Kp, float *lo, void *stream) {
  MTS_REQUIRE(src1 && lo, MTS_E_BADARG, "pack_rows_bf16in: null pointer");
  MTS_REQUIRE(D2 == 0 || src2, MTS_E_BADARG, "pack_rows_bf16in: D2 > 0 without src2");
  MTS_REQUIRE(B > 0 && T > 0 && D1 > 0 && D2 >= 0, MTS_E_BADARG, "pack_rows_bf16in: bad shape");
  MTS_REQUIRE(Kp % 32 == 0 && Kp >= D1 + D2, MTS_E_BADARG, "pack_rows_bf16in: Kp must be a multiple of 32 and >= D1 + D2");
  MTS_REQUIRE(D1 % 8 == 0 && D2 % 8 == 0 && bstride1 % 8 == 0 && bstride2 % 8 == 0 &&
                  ((((uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)lo) & 15) == 0),
              MTS_E_UNSUPPORTED, "pack_rows_bf16in: widths must be multiples of 8 and rows 16-byte aligned");
  pack_rows_bf16in_kernel<<<grid_for((int64_t)B * T * (Kp / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t *)src1, bstride1, D1, (const uint16_t *)src2, bstride2, D2, B, T, Kp, lo);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_split_tf32(const float *src, int64_t ld, int rows, int cols, int Kp, int side, float *hi, float *lo,
                              void *stream) {
  MTS_REQUIRE(src && hi && lo, MTS_E_BADARG, "split_tf32: null pointer");
  MTS_REQUIRE(rows > 0 && cols > 0 && Kp % 32 == 0 && Kp >= cols, MTS_E_BADARG, "split_tf32: bad shape");
  MTS_REQUIRE(side == 0 || side == 1, MTS_E_BADARG, "split_tf32: side must be 0 (A operand) or 1 (B operand)");
  if (((((uintptr_t)src | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0) && ld % 4 == 0 && cols % 8 == 0) {
    split_tf32_v8_kernel<<<grid_for((int64_t)rows * (Kp / 8)), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols, Kp, side,
                                                                                               hi, lo);
  } else {
    split_tf32_kernel<<<grid_for((int64_t)rows * Kp), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols, Kp, side, hi, lo);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_transpose_split(const float *src, int64_t bstride, int64_t ld, int rows, int cols, int T, int shift,
                                   const int32_t *lengths, int Kp, int side, float *hi, float *lo, void *stream) {
  MTS_REQUIRE(src && hi && lo, MTS_E_BADARG, "transpose_split: null pointer");
  MTS_REQUIRE(rows > 0 && cols > 0 && T > 0 && Kp % 32 == 0 && Kp >= rows, MTS_E_BADARG, "transpose_split: bad shape");
  MTS_REQUIRE(shift >= -1 && shift <= 1, MTS_E_BADARG, "transpose_split: shift must be -1, 0 or +1");
  MTS_REQUIRE((((uintptr_t)hi | (uintptr_t)lo) & 15) == 0, MTS_E_BADARG, "transpose_split: outputs must be 16-byte aligned");
  const dim3 grid((unsigned)((Kp + 63) / 64), (unsigned)((cols + 31) / 32));
  MTS_REQUIRE(grid.y <= 65535, MTS_E_UNSUPPORTED, "transpose_split: too many columns");
  transpose_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, bstride, ld, rows, cols, T, shift, lengths, Kp, side, hi, lo);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_gather_pad(const float *src, const int64_t *offsets, const int32_t *lengths, const int32_t *ids, int B,
                              int T, int D, float pad, float *out, void *stream) {
  MTS_REQUIRE(src && offsets && lengths && ids && out, MTS_E_BADARG, "gather_pad: null pointer");
  MTS_REQUIRE(B > 0 && T > 0 && D > 0, MTS_E_BADARG, "gather_pad: bad shape");
  gather_pad_kernel<<<grid_for((int64_t)B * T * D), 256, 0, (cudaStream_t)stream>>>(src, offsets, lengths, ids, B, T, D, pad, out);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_ragged_copy(const float *src, float *dst, const int32_t *lengths, const int32_t *offsets, int B, int S,
                               int d, int to_padded, float pad, void *stream) {
  MTS_REQUIRE(src && dst && lengths && offsets, MTS_E_BADARG, "ragged_copy: null pointer");
  MTS_REQUIRE(B > 0 && S > 0 && d > 0 && d % 4 == 0, MTS_E_BADARG, "ragged_copy: bad shape (d must be a multiple of 4)");
  MTS_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, MTS_E_BADARG, "ragged_copy: tensors must be 16-byte aligned");
  ragged_copy_kernel<<<grid_for((int64_t)B * S * (d / 4)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst), lengths, offsets, B, S, d / 4, to_padded, pad);
  MTS_LAUNCH_CHECK();
  return 0;
}

// Operand preparation for the 3xTF32 tensor-core GEMMs: fused early-fusion concat + time crop + hi/lo split.
// The reference materialises the concat on the host at load time (utils/load_datasets_precomputed.py:158-161);
// here the two modality tensors stay separate and the concat happens while the GEMM operand is written.
// Pure streaming kernels: 4(D1+D2) bytes read, 8 Kp bytes written per sentence.
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
    const float h = tf32_rn(v);
    hi[idx] = h;
    lo[idx] = tf32_rn(v - h);  // v - h is exact in fp32; rounding it to TF32 leaves an O(2^-23 |v|) residue
  }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, int64_t ld, int rows, int cols,
                                                         int Kp, float *__restrict__ hi, float *__restrict__ lo) {
  const int64_t total = (int64_t)rows * Kp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const int64_t r = idx / Kp;
    const float v = (k < cols) ? __ldg(src + r * ld + k) : 0.0f;
    const float h = tf32_rn(v);
    hi[idx] = h;
    lo[idx] = tf32_rn(v - h);
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
  MTS_REQUIRE(src1 && hi && lo, MTS_E_BADARG, "pack_rows_split: null pointer");
  MTS_REQUIRE(D2 == 0 || src2, MTS_E_BADARG, "pack_rows_split: D2 > 0 without src2");
  MTS_REQUIRE(B > 0 && T > 0 && D1 > 0 && D2 >= 0, MTS_E_BADARG, "pack_rows_split: bad shape");
  MTS_REQUIRE(Kp % 32 == 0 && Kp >= D1 + D2, MTS_E_BADARG, "pack_rows_split: Kp must be a multiple of 32 and >= D1 + D2");
  pack_rows_split_kernel<<<grid_for((int64_t)B * T * Kp), 256, 0, (cudaStream_t)stream>>>(src1, bstride1, D1, src2,
                                                                                         bstride2, D2, B, T, Kp, hi, lo);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_split_tf32(const float *src, int64_t ld, int rows, int cols, int Kp, float *hi, float *lo,
                              void *stream) {
  MTS_REQUIRE(src && hi && lo, MTS_E_BADARG, "split_tf32: null pointer");
  MTS_REQUIRE(rows > 0 && cols > 0 && Kp % 32 == 0 && Kp >= cols, MTS_E_BADARG, "split_tf32: bad shape");
  split_tf32_kernel<<<grid_for((int64_t)rows * Kp), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols, Kp, hi, lo);
  MTS_LAUNCH_CHECK();
  return 0;
}

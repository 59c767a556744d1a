// fp32 CUDA-core GEMM family (exact fp32 FMA accumulation).  Used for
//   * the weight-gradient / input-gradient products of the LSTM and head backward passes
//     (dW = dG^T X, dW_hh = dG^T H_prev, dX = dG W) where the reduction runs over sentences,
//   * validating the tcgen05 3xTF32 kernel (gemm_tc.cu) and shapes it rejects.
// C[M,N] (+)= op(A)[M,K] * op(B)[K,N]  with each operand either K-major (row = m, contiguous k) or
// MN-major (row = k, contiguous m).  128x128x16 tiles, 256 threads, 8x8 register micro-tiles.
// The B operand can be read with a row shift inside episode boundaries: this expresses h_{t-1} / h_{t+1}
// of the LSTM (dW_hh) without materialising a shifted copy of the hidden states.
#include "common.cuh"

namespace mts {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct ShiftMask {  // B row k -> k + shift, zero unless 0 <= (k % T) + shift < lengths[k / T]
  int shift;
  int T;
  const int32_t *lengths;
};

template <bool KMAJOR>
__device__ __forceinline__ void load_tile(const float *__restrict__ P, int64_t ld, int rows_total, int k_total,
                                          int row0, int k0, float (&reg)[8], int tid, const ShiftMask &sm,
                                          bool use_shift) {
  if (KMAJOR) {  // P[row*ld + k]: 128 rows x 16 k; thread -> (row = tid/4 + {0,64}, kq = (tid%4)*4)
    const int kq = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + (tid >> 2) + h * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kq + i;
        reg[h * 4 + i] = (row < rows_total && k < k_total) ? __ldg(P + (int64_t)row * ld + k) : 0.0f;
      }
    }
  } else {  // P[k*ld + row]: 16 k x 128 rows; thread -> (k = tid/32 + {0,8}, rq = (tid%32)*4)
    const int rq = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int k = k0 + (tid >> 5) + h * 8;
      bool ok = k < k_total;
      if (use_shift && ok) {
        const int bi = k / sm.T, t = k - bi * sm.T + sm.shift;
        ok = (t >= 0) && (t < sm.lengths[bi]);
        k += sm.shift;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row0 + rq + i;
        reg[h * 4 + i] = (ok && row < rows_total) ? __ldg(P + (int64_t)k * ld + row) : 0.0f;
      }
    }
  }
}

template <bool KMAJOR>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float (&reg)[8], int tid) {
  if (KMAJOR) {
    const int kq = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) S[kq + i][(tid >> 2) + h * 64] = reg[h * 4 + i];
  } else {
    const int rq = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<float4 *>(&S[(tid >> 5) + h * 8][rq]) =
          make_float4(reg[h * 4], reg[h * 4 + 1], reg[h * 4 + 2], reg[h * 4 + 3]);
  }
}

// grid: (ceil(N/BN), ceil(M/BM), splits)
template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float *__restrict__ A, int64_t lda,
                                                       const float *__restrict__ Bm, int64_t ldb,
                                                       const float *__restrict__ bias, float *__restrict__ C,
                                                       int64_t ldc, int M, int N, int K, int epilogue, int accumulate,
                                                       ShiftMask sm, int use_shift) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int splits = gridDim.z;
  const int kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[8];
  if (kbeg < kend) {
    load_tile<A_KMAJOR>(A, lda, M, kend, m0, kbeg, ra, tid, sm, false);
    load_tile<B_KMAJOR>(Bm, ldb, N, kend, n0, kbeg, rb, tid, sm, use_shift != 0);
  }
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    store_tile<A_KMAJOR>(As, ra, tid);
    store_tile<B_KMAJOR>(Bs, rb, tid);
    __syncthreads();
    if (k0 + BK < kend) {
      load_tile<A_KMAJOR>(A, lda, M, kend, m0, k0 + BK, ra, tid, sm, false);
      load_tile<B_KMAJOR>(Bm, ldb, N, kend, n0, k0 + BK, rb, tid, sm, use_shift != 0);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool first_split = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (splits > 1) {  // C was zeroed (or holds the running sum) by the host wrapper
        if (first_split && bias && epilogue >= 1) v += __ldg(bias + n);
        atomicAdd(C + (int64_t)m * ldc + n, v);
      } else {
        if (bias && epilogue >= 1) v += __ldg(bias + n);
        if (epilogue == 2) v = gelu_erf(v);
        if (accumulate) v += C[(int64_t)m * ldc + n];
        C[(int64_t)m * ldc + n] = v;
      }
    }
  }
}

// column sums: out[n] (+)= sum_m X[m, n]   (bias gradients)
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ X, int64_t ld, int M, int N,
                                                             float *__restrict__ partial) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.0f;
  for (int m = blockIdx.y; m < M; m += gridDim.y) acc += __ldg(X + (int64_t)m * ld + n);
  partial[(size_t)blockIdx.y * N + n] = acc;
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float *__restrict__ partial, int G, int N,
                                                           float *__restrict__ out, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = accumulate ? out[n] : 0.0f;
  for (int g = 0; g < G; ++g) acc += partial[(size_t)g * N + n];
  out[n] = acc;
}

}  // namespace mts

using namespace mts;

// layout bits: 1 = A is MN-major ("transposed"), 2 = B is MN-major.
//   0: C = A[M,K] B[N,K]^T   (NT: forward projections)      2: C = A[M,K] B[K,N]   (NN: dX = dG W)
//   3: C = A[K,M]^T B[K,N]   (TN: weight gradients)
extern "C" int mts_gemm_f32(const float *A, int64_t lda, const float *B, int64_t ldb, const float *bias, float *C,
                            int64_t ldc, int M, int N, int K, int layout, int epilogue, int accumulate, int splits,
                            int shift, int T, const int32_t *lengths, void *stream) {
  MTS_REQUIRE(A && B && C, MTS_E_BADARG, "gemm_f32: null pointer");
  MTS_REQUIRE(M > 0 && N > 0 && K > 0, MTS_E_BADARG, "gemm_f32: empty shape");
  MTS_REQUIRE(layout == 0 || layout == 2 || layout == 3, MTS_E_UNSUPPORTED, "gemm_f32: layout must be 0, 2 or 3");
  MTS_REQUIRE(shift == 0 || (lengths && T > 0 && (layout & 2)), MTS_E_BADARG, "gemm_f32: shift needs lengths, T, MN-major B");
  cudaStream_t st = (cudaStream_t)stream;
  if (splits < 1) splits = 1;
  if (splits > 1) {
    MTS_REQUIRE(epilogue != 2, MTS_E_UNSUPPORTED, "gemm_f32: GELU epilogue cannot be split over K");
    if (!accumulate) MTS_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st));
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  ShiftMask sm{shift, T > 0 ? T : 1, lengths};
  const int use_shift = shift != 0;
  if (layout == 0)
    gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, epilogue, accumulate, sm, 0);
  else if (layout == 2)
    gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, epilogue, accumulate, sm, use_shift);
  else
    gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, epilogue, accumulate, sm, use_shift);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t mts_colsum_ws_bytes(int M, int N) {
  const int G = M < 256 ? (M > 0 ? M : 1) : 256;
  return (int64_t)G * N * (int64_t)sizeof(float);
}

extern "C" int mts_colsum(const float *X, int64_t ld, int M, int N, float *out, int accumulate, void *ws, void *stream) {
  MTS_REQUIRE(X && out && ws, MTS_E_BADARG, "colsum: null pointer");
  MTS_REQUIRE(M > 0 && N > 0, MTS_E_BADARG, "colsum: empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int G = M < 256 ? M : 256;
  colsum_partial_kernel<<<dim3((N + 255) / 256, G), 256, 0, st>>>(X, ld, M, N, (float *)ws);
  colsum_final_kernel<<<(N + 255) / 256, 256, 0, st>>>((const float *)ws, G, N, out, accumulate);
  MTS_LAUNCH_CHECK();
  return 0;
}

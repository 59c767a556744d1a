// Linear-chain CRF kernels: Viterbi decode, forward algorithm (log partition), gold score, and the
// forward-backward gradient.  Restates models/CRF.py:148-240 of the reference as warp-per-episode scans.
//
// Layout: emissions [B, L, C] fp32; one warp owns one episode and walks it in chunks of 32 steps: the
// 32 lanes load the chunk's emissions coalesced, the sequential recurrence then runs on registers with the
// step's emission broadcast by shuffle, and the chunk's packed back-pointers go out as one coalesced store.
// Algorithmic HBM bytes per sentence: 4C (emissions) + 4 (path) [+ 4C alphas in training].
#include "common.cuh"

namespace mts {

constexpr float kImpossible = -1e4f;  // models/CRF.py:95 -- finite on purpose, takes part in fp32 adds

template <int C>
struct BpBits {
  static constexpr int kBits = (C <= 4) ? 2 : 3;
  static constexpr uint32_t kMask = (1u << kBits) - 1u;
  __host__ __device__ static constexpr uint32_t identity() {
    uint32_t m = 0;
    for (int x = 0; x < C; ++x) m |= (uint32_t)x << (x * kBits);
    return m;
  }
  __device__ static __forceinline__ uint32_t apply(uint32_t m, uint32_t x) { return (m >> (x * kBits)) & kMask; }
  // (a o b)(x) = a(b(x))
  __device__ static __forceinline__ uint32_t compose(uint32_t a, uint32_t b) {
    uint32_t r = 0;
#pragma unroll
    for (int x = 0; x < C; ++x) r |= apply(a, apply(b, x)) << (x * kBits);
    return r;
  }
};

// ---------------------------------------------------------------------------------------------------------
// Viterbi (CRF.py:172-216).  Bit-exact contract: acc[i][j] = s[j] + T[i][j] in fp32, first maximum over j wins,
// then + e[i]; masked steps leave the state untouched (x*1 + old*0 is exact for finite values).
// ---------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) crf_viterbi_kernel(const float *__restrict__ emis,
                                                          const int32_t *__restrict__ lengths,
                                                          const float *__restrict__ trans, int B, int L,
                                                          float *__restrict__ best_score, int32_t *__restrict__ paths,
                                                          uint32_t *__restrict__ bp_ws) {
  using BP = BpBits<C>;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int len = min(max(lengths[b], 0), L);
  constexpr int kStart = C - 2, kStop = C - 1;

  float T[C][C];
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) T[i][j] = __ldg(trans + i * C + j);

  float s[C];
#pragma unroll
  for (int i = 0; i < C; ++i) s[i] = (i == kStart) ? 0.0f : kImpossible;

  const float *e_b = emis + (size_t)b * L * C;
  uint32_t *bp_b = bp_ws + (size_t)b * L;

  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    float e[C];
    if (t < len) {
      if constexpr (C == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(e_b) + t);
        e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < C; ++i) e[i] = __ldg(e_b + (size_t)t * C + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i) e[i] = 0.0f;
    }
    const int nsteps = min(32, len - t0);
    uint32_t my_bp = 0;
#pragma unroll 4
    for (int k = 0; k < nsteps; ++k) {
      float ns[C];
      uint32_t bp = 0;
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float et = __shfl_sync(0xffffffffu, e[i], k);
        float best = __fadd_rn(s[0], T[i][0]);
        uint32_t arg = 0;
#pragma unroll
        for (int j = 1; j < C; ++j) {
          const float v = __fadd_rn(s[j], T[i][j]);
          if (v > best) { best = v; arg = j; }
        }
        ns[i] = __fadd_rn(best, et);
        bp |= arg << (i * BP::kBits);
      }
#pragma unroll
      for (int i = 0; i < C; ++i) s[i] = ns[i];
      if (lane == k) my_bp = bp;
    }
    if (t < len) bp_b[t] = my_bp;
  }

  // transition to STOP, arg-max (first maximum)
  float best = __fadd_rn(s[0], T[kStop][0]);
  uint32_t tag = 0;
#pragma unroll
  for (int i = 1; i < C; ++i) {
    const float v = __fadd_rn(s[i], T[kStop][i]);
    if (v > best) { best = v; tag = i; }
  }
  if (lane == 0) best_score[b] = best;
  __syncwarp();

  // Back-trace as a suffix scan of function composition: tag_t = (f_{t+1} o ... o f_{len-1})(tag_{len-1}),
  // where f_t is the packed back-pointer map of step t.  32 steps per warp-scan instead of a 32-deep chain.
  int32_t *p_b = paths + (size_t)b * L;
  uint32_t carry = tag;
  for (int t0 = ((len - 1) >> 5) << 5; t0 >= 0 && len > 0; t0 -= 32) {
    const int t = t0 + lane;
    uint32_t S = (t + 1 < len) ? bp_b[t + 1] : BP::identity();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t other = __shfl_down_sync(0xffffffffu, S, d);
      if (lane + d < 32) S = BP::compose(S, other);
    }
    const uint32_t my_tag = BP::apply(S, carry);
    if (t < len) p_b[t] = (int32_t)my_tag;
    carry = __shfl_sync(0xffffffffu, my_tag, 0);
  }
  for (int t = len + lane; t < L; t += 32) p_b[t] = -1;
}

// ---------------------------------------------------------------------------------------------------------
// Forward algorithm + gold score (CRF.py:218-240, 148-170).  Lane i < C owns state i.
// ---------------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ float lse_over(const float (&x)[C]) {
  float m = x[0];
#pragma unroll
  for (int j = 1; j < C; ++j) m = fmaxf(m, x[j]);
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < C; ++j) acc += expf(x[j] - m);
  return m + logf(acc);
}

template <int C>
__global__ void __launch_bounds__(128) crf_nll_fwd_kernel(const float *__restrict__ emis, const float *__restrict__ tags,
                                                          int64_t ldt, const int32_t *__restrict__ lengths,
                                                          const float *__restrict__ trans, int B, int L,
                                                          float *__restrict__ log_z, float *__restrict__ gold,
                                                          float *__restrict__ alphas) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int len = min(max(lengths[b], 0), L);
  constexpr int kStart = C - 2, kStop = C - 1;
  const int i = lane < C ? lane : C - 1;  // lanes >= C mirror lane C-1 (keeps shuffles convergent)

  float Trow[C];
#pragma unroll
  for (int j = 0; j < C; ++j) Trow[j] = __ldg(trans + i * C + j);

  const float *e_b = emis + (size_t)b * L * C;
  float *a_b = alphas ? alphas + (size_t)b * L * C : nullptr;
  float a = (i == kStart) ? 0.0f : kImpossible;

  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    float e[C];
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = (t < len) ? __ldg(e_b + (size_t)t * C + c) : 0.0f;
    const int nsteps = min(32, len - t0);
    for (int k = 0; k < nsteps; ++k) {
      float x[C];
      float ei = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v = __shfl_sync(0xffffffffu, e[c], k);
        if (c == i) ei = v;
      }
#pragma unroll
      for (int j = 0; j < C; ++j) x[j] = (__shfl_sync(0xffffffffu, a, j) + Trow[j]) + ei;
      a = lse_over<C>(x);
      if (a_b && lane < C) a_b[(size_t)(t0 + k) * C + lane] = a;
    }
  }
  {
    float x[C];
#pragma unroll
    for (int j = 0; j < C; ++j) x[j] = __shfl_sync(0xffffffffu, a, j) + __ldg(trans + kStop * C + j);
    const float z = lse_over<C>(x);
    if (lane == 0) log_z[b] = z;
  }

  // gold path score, lanes strided over time
  const float *y_b = tags + (size_t)b * ldt;
  float acc = 0.0f;
  for (int t = lane; t < len; t += 32) {
    const int y = (int)y_b[t];
    const int prev = (t == 0) ? kStart : (int)y_b[t - 1];
    acc += __ldg(trans + y * C + prev) + __ldg(e_b + (size_t)t * C + y);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const int last = (len > 0) ? (int)y_b[len - 1] : kStart;
    gold[b] = acc + __ldg(trans + kStop * C + last);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  Same chain of local soft-max Jacobians that autograd walks through the reference's T-step loop
// (well conditioned for any length: every weight vector is re-normalised from the saved alphas, instead of
// forming exp(alpha + beta - Z) whose fp32 error grows with the sequence length):
//   g_last[i]   = softmax_i(alpha_last[i] + T[stop][i])
//   w_t[i][j]   = softmax_j(alpha_{t-1}[j] + T[i][j])              (alpha_{-1} = START indicator)
//   d_emis[t,i] = scale (g_t[i] - 1[y_t = i]);  d_trans[i][j] += scale (g_t[i] w_t[i][j] - 1[y_t = i, y_{t-1} = j])
//   g_{t-1}[j]  = sum_i g_t[i] w_t[i][j]
// Equals autograd of mean_b(Z_b - gold_b) (tests/test_oracle_golden.py::test_crf pins the closed form).
// ---------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) crf_nll_bwd_kernel(const float *__restrict__ emis, const float *__restrict__ tags,
                                                          int64_t ldt, const int32_t *__restrict__ lengths,
                                                          const float *__restrict__ trans,
                                                          const float *__restrict__ alphas, int B, int L,
                                                          const float *__restrict__ scale_dev,
                                                          float *__restrict__ d_emis, float *__restrict__ d_trans) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int len = min(max(lengths[b], 0), L);
  constexpr int kStart = C - 2, kStop = C - 1;
  const float scale = __ldg(scale_dev + b);  // d loss / d logZ_b (= -d loss / d gold_b): grad_out / B for the mean
  const int i = lane < C ? lane : C - 1;     // lanes >= C mirror lane C-1 (keeps shuffles convergent)

  float Trow[C], dT[C];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    Trow[j] = __ldg(trans + i * C + j);  // T[i][j]: j -> i
    dT[j] = 0.0f;
  }
  const float *a_b = alphas + (size_t)b * L * C;
  const float *y_b = tags + (size_t)b * ldt;
  float *de_b = d_emis + (size_t)b * L * C;

  for (int t = len + (lane >> 3); t < L; t += 4)  // zero the padded tail (8 lanes x 4 rows per pass)
    for (int c = lane & 7; c < C; c += 8) de_b[(size_t)t * C + c] = 0.0f;
  if (len == 0) return;

  // g at the last valid step: soft-max over the transition to STOP
  float g;
  {
    float x[C];
#pragma unroll
    for (int j = 0; j < C; ++j) x[j] = a_b[(size_t)(len - 1) * C + j] + __ldg(trans + kStop * C + j);
    float m = x[0];
#pragma unroll
    for (int j = 1; j < C; ++j) m = fmaxf(m, x[j]);
    float s = 0.0f, mine = 0.0f;
#pragma unroll
    for (int j = 0; j < C; ++j) { const float ex = expf(x[j] - m); s += ex; if (j == i) mine = ex; }
    g = mine / s;
    const int y_last = (int)y_b[len - 1];
    if (lane < C) atomicAdd(d_trans + kStop * C + lane, scale * (g - (lane == y_last ? 1.0f : 0.0f)));
  }
  for (int t = len - 1; t >= 0; --t) {
    const int y = (int)y_b[t];
    const int yp = (t == 0) ? kStart : (int)y_b[t - 1];
    if (lane < C) de_b[(size_t)t * C + lane] = scale * (g - (lane == y ? 1.0f : 0.0f));
    float x[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float a_prev = (t == 0) ? ((j == kStart) ? 0.0f : kImpossible) : a_b[(size_t)(t - 1) * C + j];
      x[j] = a_prev + Trow[j];
    }
    float m = x[0];
#pragma unroll
    for (int j = 1; j < C; ++j) m = fmaxf(m, x[j]);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < C; ++j) { x[j] = expf(x[j] - m); s += x[j]; }
    const float gi = (lane < C) ? g / s : 0.0f;  // mirrored lanes must not be counted twice below
    float gn = 0.0f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float c = gi * x[j];  // g_t[i] w_t[i][j]
      dT[j] += c - ((i == y && j == yp && lane < C) ? 1.0f : 0.0f);
      // g_{t-1}[j] = sum over lanes i of c: butterfly over the first 8 lanes (C <= 8)
      float r = c;
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      r += __shfl_xor_sync(0xffffffffu, r, 4);
      if (j == i) gn = r;
    }
    g = gn;
  }
  if (lane < C) {
#pragma unroll
    for (int j = 0; j < C; ++j) atomicAdd(d_trans + lane * C + j, scale * dT[j]);
  }
}

template <int C>
static int launch_viterbi(const float *emis, const int32_t *lengths, const float *trans, int B, int L, float *best,
                          int32_t *paths, uint32_t *bp_ws, cudaStream_t st) {
  crf_viterbi_kernel<C><<<(B + 3) / 4, 128, 0, st>>>(emis, lengths, trans, B, L, best, paths, bp_ws);
  MTS_LAUNCH_CHECK();
  return 0;
}

}  // namespace mts

using namespace mts;

#define MTS_DISPATCH_C(C, CALL)                                              \
  switch (C) {                                                               \
    case 3: { constexpr int kC = 3; CALL; } break;                            \
    case 4: { constexpr int kC = 4; CALL; } break;                            \
    case 5: { constexpr int kC = 5; CALL; } break;                            \
    case 6: { constexpr int kC = 6; CALL; } break;                            \
    case 7: { constexpr int kC = 7; CALL; } break;                            \
    case 8: { constexpr int kC = 8; CALL; } break;                            \
    default: set_error("CRF: 3 <= C <= 8 required"); return MTS_E_UNSUPPORTED; \
  }

extern "C" int mts_crf_viterbi(const float *emis, const int32_t *lengths, const float *trans, int B, int L, int C,
                               float *best_score, int32_t *paths, uint32_t *bp_ws, void *stream) {
  MTS_REQUIRE(emis && lengths && trans && best_score && paths && bp_ws, MTS_E_BADARG, "crf_viterbi: null pointer");
  MTS_REQUIRE(B > 0 && L > 0, MTS_E_BADARG, "crf_viterbi: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_DISPATCH_C(C, return launch_viterbi<kC>(emis, lengths, trans, B, L, best_score, paths, bp_ws, st));
  return 0;
}

extern "C" int mts_crf_nll_fwd(const float *emis, const float *tags, int64_t ldt, const int32_t *lengths,
                               const float *trans, int B, int L, int C, float *log_z, float *gold, float *alphas,
                               void *stream) {
  MTS_REQUIRE(emis && tags && lengths && trans && log_z && gold, MTS_E_BADARG, "crf_nll_fwd: null pointer");
  MTS_REQUIRE(B > 0 && L > 0 && ldt >= L, MTS_E_BADARG, "crf_nll_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_DISPATCH_C(C, (crf_nll_fwd_kernel<kC><<<(B + 3) / 4, 128, 0, st>>>(emis, tags, ldt, lengths, trans, B, L, log_z,
                                                                         gold, alphas)));
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_crf_nll_bwd(const float *emis, const float *tags, int64_t ldt, const int32_t *lengths,
                               const float *trans, const float *alphas, int B, int L, int C,
                               const float *scale_dev, float *d_emis, float *d_trans, void *stream) {
  MTS_REQUIRE(emis && tags && lengths && trans && alphas && scale_dev && d_emis && d_trans, MTS_E_BADARG,
              "crf_nll_bwd: null pointer");
  MTS_REQUIRE(B > 0 && L > 0 && ldt >= L, MTS_E_BADARG, "crf_nll_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_CUDA(cudaMemsetAsync(d_trans, 0, sizeof(float) * C * C, st));
  MTS_DISPATCH_C(C, (crf_nll_bwd_kernel<kC><<<(B + 3) / 4, 128, 0, st>>>(emis, tags, ldt, lengths, trans, alphas,
                                                                         B, L, scale_dev, d_emis, d_trans)));
  MTS_LAUNCH_CHECK();
  return 0;
}

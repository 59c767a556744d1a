// Classification head, boundary thresholding and the segmentation losses
// (models/CRF.py:340-369 and models/focal_loss.py:38-57 of the reference), forward and backward.
// All of these are HBM-bound streaming kernels: one pass over the hidden states / logits.
#include "common.cuh"

namespace mts {

// ---------------------------------------------------------------------------------------------------------
// head forward: one warp per (b, t) row; coalesced float4 loads of the F features, shuffle reduction.
// ---------------------------------------------------------------------------------------------------------
template <int NOUT>
__global__ void __launch_bounds__(256) head_fwd_kernel(const float *__restrict__ feats, const float *__restrict__ w,
                                                       const float *__restrict__ bias,
                                                       const int32_t *__restrict__ lengths, int B, int T, int F,
                                                       float th, float *__restrict__ scores,
                                                       uint8_t *__restrict__ tags) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const float *x = feats + row * F;
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = 0.0f;
  if ((F & 3) == 0) {
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (int f = lane; f < (F >> 2); f += 32) {
      const float4 v = __ldg(x4 + f);
#pragma unroll
      for (int o = 0; o < NOUT; ++o) {
        const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + (size_t)o * F) + f);
        acc[o] += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
      }
    }
  } else {
    for (int f = lane; f < F; f += 32) {
      const float v = __ldg(x + f);
#pragma unroll
      for (int o = 0; o < NOUT; ++o) acc[o] += v * __ldg(w + (size_t)o * F + f);
    }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = warp_sum(acc[o]) + __ldg(bias + o);
  if (lane == 0) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o) scores[row * NOUT + o] = acc[o];
    if (tags) {
      const int b = (int)(row / T), t = (int)(row % T);
      uint8_t tag = 0xFF;
      if (t < lengths[b]) {
        float p;
        if (NOUT == 1) {
          p = sigmoid_ref(acc[0]);
        } else {  // torch.softmax: exp(x - max) / sum
          const float m = fmaxf(acc[0], acc[NOUT - 1]);
          const float e0 = expf(acc[0] - m), e1 = expf(acc[NOUT - 1] - m);
          p = __fdiv_rn(e1, e0 + e1);
        }
        tag = p > th ? 1 : 0;
      }
      tags[row] = tag;
    }
  }
}

// d_feats[r, f] = sum_o d_s[r, o] w[o, f]
template <int NOUT>
__global__ void __launch_bounds__(256) head_bwd_dx_kernel(const float *__restrict__ ds, const float *__restrict__ w,
                                                          int64_t rows, int F, float *__restrict__ dx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * F) return;
  const int64_t r = idx / F;
  const int f = (int)(idx % F);
  float acc = 0.0f;
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc += __ldg(ds + r * NOUT + o) * __ldg(w + (size_t)o * F + f);
  dx[idx] = acc;
}

// partial[g][o][f] = sum over rows r = g, g+G, ... of d_s[r,o] * feats[r,f]; column F carries d_bias.
template <int NOUT>
__global__ void __launch_bounds__(256) head_bwd_dw_partial_kernel(const float *__restrict__ ds,
                                                                  const float *__restrict__ feats, int64_t rows,
                                                                  int F, float *__restrict__ partial) {
  const int G = gridDim.x;
  for (int f = threadIdx.x; f <= F; f += blockDim.x) {
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.0f;
    for (int64_t r = blockIdx.x; r < rows; r += G) {
      const float x = (f < F) ? __ldg(feats + r * F + f) : 1.0f;
#pragma unroll
      for (int o = 0; o < NOUT; ++o) acc[o] += __ldg(ds + r * NOUT + o) * x;
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) partial[((size_t)blockIdx.x * NOUT + o) * (F + 1) + f] = acc[o];
  }
}

__global__ void head_bwd_dw_reduce_kernel(const float *__restrict__ partial, int G, int n_out, int F,
                                          float *__restrict__ dw, float *__restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out * (F + 1)) return;
  float acc = 0.0f;
  for (int g = 0; g < G; ++g) acc += partial[(size_t)g * n_out * (F + 1) + idx];
  const int o = idx / (F + 1), f = idx % (F + 1);
  if (f < F) dw[(size_t)o * F + f] = acc; else db[o] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// losses
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bce_logits(float z, float y) {
  return fmaxf(z, 0.0f) - z * y + log1pf(expf(-fabsf(z)));
}

__device__ __forceinline__ float focal_elem(float z, float y, float alpha, float gamma) {
  const float p = sigmoid_ref(z);
  const float ce = bce_logits(z, y);
  const float pt = p * y + (1.0f - p) * (1.0f - y);
  const float om = 1.0f - pt;
  const float mod = (gamma == 2.0f) ? om * om : powf(om, gamma);
  float loss = ce * mod;
  if (alpha >= 0.0f) loss *= alpha * y + (1.0f - alpha) * (1.0f - y);
  return loss;
}

__device__ __forceinline__ float focal_grad_elem(float z, float y, float alpha, float gamma) {
  const float p = sigmoid_ref(z);
  const float ce = bce_logits(z, y);
  const float pt = p * y + (1.0f - p) * (1.0f - y);
  const float om = 1.0f - pt;
  float mod, dmod;  // (1-pt)^gamma and gamma (1-pt)^(gamma-1)
  if (gamma == 2.0f) { mod = om * om; dmod = 2.0f * om; }
  else { mod = powf(om, gamma); dmod = (om > 0.0f) ? gamma * powf(om, gamma - 1.0f) : 0.0f; }
  const float dpt = (2.0f * y - 1.0f) * p * (1.0f - p);
  float g = (p - y) * mod - ce * dmod * dpt;
  if (alpha >= 0.0f) g *= alpha * y + (1.0f - alpha) * (1.0f - y);
  return g;
}

__device__ __forceinline__ float bce_elem(float z, float y) {
  const float p = sigmoid_ref(z);
  const float lp = fmaxf(logf(p), -100.0f), l1p = fmaxf(log1pf(-p), -100.0f);
  return -(y * lp + (1.0f - y) * l1p);
}
__device__ __forceinline__ float bce_grad_elem(float z, float y) {
  const float p = sigmoid_ref(z);
  const float v = p * (1.0f - p);
  return (p - y) * v / fmaxf(v, 1e-12f);  // BCELoss backward (eps 1e-12) chained with sigmoid'
}

// kind 0 focal, 1 bce: scores [B,T,1]; kind 2: scores [B,T,2], target -1 = ignore.
__global__ void __launch_bounds__(256) seg_loss_partial_kernel(const float *__restrict__ scores,
                                                               const float *__restrict__ target, int64_t ldt,
                                                               const int32_t *__restrict__ lengths, int B, int T,
                                                               int kind, float alpha, float gamma,
                                                               float *__restrict__ partial) {
  float acc = 0.0f, cnt = 0.0f;
  const int64_t n = (int64_t)B * T;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / T), t = (int)(idx % T);
    const float y = __ldg(target + (size_t)b * ldt + t);
    if (kind == 2) {
      if (y != -1.0f) {
        const float z0 = __ldg(scores + idx * 2), z1 = __ldg(scores + idx * 2 + 1);
        const float m = fmaxf(z0, z1);
        const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
        acc += lse - ((int)y == 1 ? z1 : z0);
        cnt += 1.0f;
      }
    } else if (t < lengths[b]) {
      const float z = __ldg(scores + idx);
      acc += (kind == 0) ? focal_elem(z, y, alpha, gamma) : bce_elem(z, y);
      cnt += 1.0f;
    }
  }
  __shared__ float sa[8], sc[8];
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = acc; sc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, c = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sa[i]; c += sc[i]; }
    partial[blockIdx.x] = a;
    partial[1024 + blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(1024) seg_loss_final_kernel(const float *__restrict__ partial, int nblocks,
                                                              float inv_count, float *__restrict__ loss_out) {
  __shared__ double sa[32], sc[32];
  double a = (threadIdx.x < nblocks) ? (double)partial[threadIdx.x] : 0.0;
  double c = (threadIdx.x < nblocks) ? (double)partial[1024 + threadIdx.x] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tc = 0.0;
    for (int i = 0; i < 32; ++i) { ta += sa[i]; tc += sc[i]; }
    const double inv = (inv_count > 0.0f) ? (double)inv_count : (tc > 0.0 ? 1.0 / tc : 0.0);
    loss_out[0] = (float)(ta * inv);
    loss_out[1] = (float)tc;
  }
}

__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float *__restrict__ scores,
                                                           const float *__restrict__ target, int64_t ldt,
                                                           const int32_t *__restrict__ lengths, int B, int T, int kind,
                                                           float alpha, float gamma, float inv_count,
                                                           const float *__restrict__ count_dev,
                                                           const float *__restrict__ grad_out,
                                                           float *__restrict__ d_scores) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * T) return;
  const int b = (int)(idx / T), t = (int)(idx % T);
  float inv = inv_count;
  if (inv <= 0.0f) { const float c = count_dev ? *count_dev : 0.0f; inv = c > 0.0f ? 1.0f / c : 0.0f; }
  const float go = __ldg(grad_out) * inv;
  const float y = __ldg(target + (size_t)b * ldt + t);
  if (kind == 2) {
    float g0 = 0.0f, g1 = 0.0f;
    if (y != -1.0f) {
      const float z0 = __ldg(scores + idx * 2), z1 = __ldg(scores + idx * 2 + 1);
      const float m = fmaxf(z0, z1);
      const float e0 = expf(z0 - m), e1 = expf(z1 - m);
      const float s = e0 + e1;
      g0 = go * (e0 / s - ((int)y == 0 ? 1.0f : 0.0f));
      g1 = go * (e1 / s - ((int)y == 1 ? 1.0f : 0.0f));
    }
    d_scores[idx * 2] = g0;
    d_scores[idx * 2 + 1] = g1;
  } else {
    float g = 0.0f;
    if (t < lengths[b]) {
      const float z = __ldg(scores + idx);
      g = go * ((kind == 0) ? focal_grad_elem(z, y, alpha, gamma) : bce_grad_elem(z, y));
    }
    d_scores[idx] = g;
  }
}

}  // namespace mts

using namespace mts;

extern "C" int mts_head_fwd(const float *feats, const float *w, const float *bias, const int32_t *lengths, int B, int T,
                            int F, int n_out, float th, float *scores, uint8_t *tags, void *stream) {
  MTS_REQUIRE(feats && w && bias && scores, MTS_E_BADARG, "head_fwd: null pointer");
  MTS_REQUIRE(!tags || lengths, MTS_E_BADARG, "head_fwd: tags requested without lengths");
  MTS_REQUIRE(B > 0 && T > 0 && F > 0, MTS_E_BADARG, "head_fwd: empty shape");
  MTS_REQUIRE(n_out == 1 || n_out == 2, MTS_E_UNSUPPORTED, "head_fwd: n_out must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * T;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (n_out == 1) head_fwd_kernel<1><<<grid, 256, 0, st>>>(feats, w, bias, lengths, B, T, F, th, scores, tags);
  else head_fwd_kernel<2><<<grid, 256, 0, st>>>(feats, w, bias, lengths, B, T, F, th, scores, tags);
  MTS_LAUNCH_CHECK();
  return 0;
}

static int head_bwd_groups(int64_t rows) { return (int)(rows < 2 * kNumSMs ? (rows > 0 ? rows : 1) : 2 * kNumSMs); }

extern "C" int64_t mts_head_bwd_ws_bytes(int B, int T, int F, int n_out) {
  return (int64_t)head_bwd_groups((int64_t)B * T) * n_out * (F + 1) * (int64_t)sizeof(float);
}

extern "C" int mts_head_bwd(const float *d_scores, const float *feats, const float *w, int B, int T, int F, int n_out,
                            float *d_feats, float *d_w, float *d_bias, void *ws, void *stream) {
  MTS_REQUIRE(d_scores && feats && w && d_w && d_bias && ws, MTS_E_BADARG, "head_bwd: null pointer");
  MTS_REQUIRE(n_out == 1 || n_out == 2, MTS_E_UNSUPPORTED, "head_bwd: n_out must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * T;
  const int G = head_bwd_groups(rows);
  float *partial = (float *)ws;
  if (n_out == 1) {
    if (d_feats) head_bwd_dx_kernel<1><<<(unsigned)((rows * F + 255) / 256), 256, 0, st>>>(d_scores, w, rows, F, d_feats);
    head_bwd_dw_partial_kernel<1><<<G, 256, 0, st>>>(d_scores, feats, rows, F, partial);
  } else {
    if (d_feats) head_bwd_dx_kernel<2><<<(unsigned)((rows * F + 255) / 256), 256, 0, st>>>(d_scores, w, rows, F, d_feats);
    head_bwd_dw_partial_kernel<2><<<G, 256, 0, st>>>(d_scores, feats, rows, F, partial);
  }
  head_bwd_dw_reduce_kernel<<<(n_out * (F + 1) + 255) / 256, 256, 0, st>>>(partial, G, n_out, F, d_w, d_bias);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_seg_loss_fwd(const float *scores, const float *target, int64_t ldt, const int32_t *lengths, int B,
                                int T, int kind, float alpha, float gamma, float inv_count, float *loss_out,
                                float *partial, void *stream) {
  MTS_REQUIRE(scores && target && lengths && loss_out && partial, MTS_E_BADARG, "seg_loss_fwd: null pointer");
  MTS_REQUIRE(kind >= 0 && kind <= 2, MTS_E_UNSUPPORTED, "seg_loss_fwd: kind must be 0 (focal), 1 (bce) or 2 (ce)");
  MTS_REQUIRE(B > 0 && T > 0 && ldt >= T, MTS_E_BADARG, "seg_loss_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)B * T;
  int nblocks = (int)((n + 255) / 256);
  if (nblocks > 1024) nblocks = 1024;
  seg_loss_partial_kernel<<<nblocks, 256, 0, st>>>(scores, target, ldt, lengths, B, T, kind, alpha, gamma, partial);
  seg_loss_final_kernel<<<1, 1024, 0, st>>>(partial, nblocks, inv_count, loss_out);
  MTS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mts_seg_loss_bwd(const float *scores, const float *target, int64_t ldt, const int32_t *lengths, int B,
                                 int T, int kind, float alpha, float gamma, float inv_count, const float *count_dev,
                                 const float *grad_out, float *d_scores, void *stream) {
  MTS_REQUIRE(scores && target && lengths && grad_out && d_scores, MTS_E_BADARG, "seg_loss_bwd: null pointer");
  MTS_REQUIRE(kind >= 0 && kind <= 2, MTS_E_UNSUPPORTED, "seg_loss_bwd: kind must be 0, 1 or 2");
  MTS_REQUIRE(inv_count > 0.0f || count_dev, MTS_E_BADARG, "seg_loss_bwd: need inv_count or a device count");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)B * T;
  seg_loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, target, ldt, lengths, B, T, kind, alpha,
                                                                  gamma, inv_count, count_dev, grad_out, d_scores);
  MTS_LAUNCH_CHECK();
  return 0;
}



// ---------------------------------------------------------------------------------------------------------
// On-device evaluation counts.  The reference's test_step walks every episode on the
// host through segeval (models/lightning_model.py:16-55, 607-637: Decimal arithmetic over Python lists).  Pk and
// WindowDiff are integer counts over prefix sums of the boundary flags; the host only forms the final ratios.
//   per episode b (n = len_b), after forcing the last unit to close a segment on both sides (:27-30):
//     seg_r[i], seg_h[i] = number of boundaries before position i   (exclusive prefix sums)
//     k = round-half-even(n / (2 * #reference segments)), at least 2                  (segeval's default window)
//     pk  = #{ i < n-k : (seg_r[i] == seg_r[i+k]) != (seg_h[i] == seg_h[i+k]) }
//     wd  = #{ i < n-k : seg_r[i+k] - seg_r[i] != seg_h[i+k] - seg_h[i] }
//   and tp / fp / fn of the raw boundary flags (F1 of the boundary class); `zero_last` first clears the last unit
//   of both sides (the end_boundary option, :609-611).
//   out[b][8] = {pk, wd, n - k, k, tp, fp, fn, #reference segments}
// One CTA per episode; prefix sums by warp 0 (shuffle scan, 32 per iteration), counting by all threads.
// ---------------------------------------------------------------------------------------------------------
namespace mts {

__global__ void __launch_bounds__(128) seg_metrics_kernel(const uint8_t *__restrict__ tags, int64_t ld_tags,
                                                          const float *__restrict__ target, int64_t ldt,
                                                          const int32_t *__restrict__ lengths, int T, int zero_last,
                                                          int32_t *__restrict__ out) {
  extern __shared__ int32_t seg[];  // seg_r[T + 1], seg_h[T + 1]
  __shared__ int32_t cnt[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int n = min(max(lengths[b], 0), T);
  int32_t *seg_r = seg, *seg_h = seg + (T + 1);
  const uint8_t *hb = tags + (int64_t)b * ld_tags;
  const float *rb = target + (int64_t)b * ldt;
  if (tid < 8) cnt[tid] = 0;
  __syncthreads();
  if (n == 0) {
    if (tid < 8) out[b * 8 + tid] = 0;
    return;
  }
  // raw-flag confusion counts
  int tp = 0, fp = 0, fn = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const bool last = (i == n - 1);
    const bool r = !(zero_last && last) && rb[i] != 0.0f;
    const bool h = !(zero_last && last) && hb[i] != 0;
    tp += r && h; fp += !r && h; fn += r && !h;
  }
  atomicAdd(&cnt[4], tp); atomicAdd(&cnt[5], fp); atomicAdd(&cnt[6], fn);
  // exclusive prefix sums of the flags with the last unit forced to 1
  if (tid < 32) {
    int carry_r = 0, carry_h = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      int r = 0, h = 0;
      if (i < n) {
        r = (i == n - 1) ? 1 : (rb[i] != 0.0f);
        h = (i == n - 1) ? 1 : (hb[i] != 0);
      }
      int sr = r, sh = h;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ur = __shfl_up_sync(0xffffffffu, sr, o), uh = __shfl_up_sync(0xffffffffu, sh, o);
        if (lane >= o) { sr += ur; sh += uh; }
      }
      if (i < n) { seg_r[i] = carry_r + sr - r; seg_h[i] = carry_h + sh - h; }
      carry_r += __shfl_sync(0xffffffffu, sr, 31);
      carry_h += __shfl_sync(0xffffffffu, sh, 31);
    }
    if (lane == 0) {
      seg_r[n] = carry_r; seg_h[n] = carry_h;
      const int den = 2 * carry_r;  // carry_r = number of reference segments (>= 1)
      int q = n / den;
      const int r2 = 2 * (n % den);
      if (r2 > den || (r2 == den && (q & 1))) ++q;
      cnt[3] = q > 1 ? q : 2;
      cnt[7] = carry_r;
    }
  }
  __syncthreads();
  const int k = cnt[3];
  int pk = 0, wd = 0;
  for (int i = tid; i + k < n; i += blockDim.x) {
    const int dr = seg_r[i + k] - seg_r[i], dh = seg_h[i + k] - seg_h[i];
    pk += (dr == 0) != (dh == 0);
    wd += dr != dh;
  }
  atomicAdd(&cnt[0], pk); atomicAdd(&cnt[1], wd);
  __syncthreads();
  if (tid == 0) cnt[2] = n - k;
  __syncthreads();
  if (tid < 8) out[b * 8 + tid] = cnt[tid];
}

}  // namespace mts

extern "C" int mts_seg_metrics(const uint8_t *tags, int64_t ld_tags, const float *target, int64_t ldt,
                               const int32_t *lengths, int B, int T, int zero_last, int32_t *out, void *stream) {
  MTS_REQUIRE(tags && target && lengths && out, MTS_E_BADARG, "seg_metrics: null pointer");
  MTS_REQUIRE(B > 0 && T > 0, MTS_E_BADARG, "seg_metrics: bad shape");
  const size_t smem = (size_t)2 * (T + 1) * sizeof(int32_t);
  MTS_REQUIRE(smem <= 200 * 1024, MTS_E_UNSUPPORTED, "seg_metrics: episodes longer than 25 000 sentences are not supported");
  if (smem > 48 * 1024)  // per call: the attribute is per device, and a process may drive several
    MTS_CUDA(cudaFuncSetAttribute(mts::seg_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mts::seg_metrics_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(tags, ld_tags, target, ldt, lengths, T, zero_last, out);
  MTS_LAUNCH_CHECK();
  return 0;
}

// LSTM recurrence, backward through time.  The reference has no hand-written backward: this must equal
// autograd through torch.nn.LSTM as called at models/NeuralArchitectures.py:113 (SURVEY.md section 8a').
//
// Produces dgx = d loss / d (gate pre-activations) for every valid step; the weight gradients are GEMMs
// over dgx (dW_ih = dgx^T X, dW_hh = dgx^T H_prev, db = colsum dgx) issued by the host wrapper.
//
// H == 256: persistent cluster kernel mirroring the forward one.  CTA r owns the cells of units
// [32r, 32r+32) x 8 episodes.  Per step:
//   cell phase   : dh = dy + sum of the 8 partial W_hh^T products received last step; gate derivatives;
//                  dp (4 x 32 x 8) -> shared memory and -> dgx in HBM;
//   matvec phase : partial dh_prev[k][e] = sum over this CTA's 128 gate rows of W_hh[row][k] dp[row][e]
//                  (W_hh slice register-resident: 32 rows x 4 k per thread), reduced across the 4 row
//                  slices by recursive halving, then pushed (st.async + mbarrier) to the CTA that owns unit k
//                  -- a reduce-scatter through distributed shared memory, 8 KB per CTA per step.
#include <cooperative_groups.h>

#include "cluster_utils.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace mts {

struct GateGrads {
  float dpi, dpf, dpg, dpo, dc_prev;
};

__device__ __forceinline__ GateGrads lstm_cell_bwd(float dh, float dc_in, float ig, float fg, float gg, float og,
                                                   float c, float c_prev) {
  const float tc = tanhf(c);
  const float d_o = dh * tc;
  const float dc = dc_in + dh * og * (1.0f - tc * tc);
  GateGrads r;
  r.dpi = dc * gg * ig * (1.0f - ig);
  r.dpf = dc * c_prev * fg * (1.0f - fg);
  r.dpg = dc * ig * (1.0f - gg * gg);
  r.dpo = d_o * og * (1.0f - og);
  r.dc_prev = dc * fg;
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// generic kernel: grid (B, 2, n_enc), 256 threads; smem dh[H], dc[H], dp[4H]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_bwd_generic_kernel(const float *__restrict__ dy,
                                                               const float *__restrict__ gates,
                                                               const float *__restrict__ w_hh,
                                                               const int32_t *__restrict__ lengths, int B, int T,
                                                               int H, int n_enc, float *__restrict__ dgx) {
  extern __shared__ float smem[];
  float *dh = smem, *dc = smem + H, *dp = smem + 2 * H;
  const int b = blockIdx.x, dir = blockIdx.y, enc = blockIdx.z;
  const int len = min(max(lengths[b], 0), T);
  const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * H * H;
  const int ycols = n_enc * 2 * H;
  const float *gs_b = gates + (((size_t)enc * 2 + dir) * B + b) * T * 5 * H;
  float *dgx_b = dgx + (size_t)enc * B * T * 8 * H + (size_t)b * T * 8 * H + (size_t)dir * 4 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) { dh[u] = 0.0f; dc[u] = 0.0f; }
  __syncthreads();
  for (int s = 0; s < len; ++s) {
    const int t = dir ? s : len - 1 - s;
    const int tp = dir ? t + 1 : t - 1;  // the step the forward pass visited just before t
    const bool has_prev = dir ? (tp < len) : (tp >= 0);
    const float *gs = gs_b + (size_t)t * 5 * H;
    for (int u = threadIdx.x; u < H; u += blockDim.x) {
      const float dht = dy[((size_t)b * T + t) * ycols + (size_t)enc * 2 * H + dir * H + u] + dh[u];
      const float c_prev = has_prev ? gs_b[(size_t)tp * 5 * H + 4 * H + u] : 0.0f;
      const GateGrads g = lstm_cell_bwd(dht, dc[u], gs[u], gs[H + u], gs[2 * H + u], gs[3 * H + u], gs[4 * H + u], c_prev);
      dc[u] = g.dc_prev;
      dp[u] = g.dpi; dp[H + u] = g.dpf; dp[2 * H + u] = g.dpg; dp[3 * H + u] = g.dpo;
      float *o = dgx_b + (size_t)t * 8 * H;
      o[u] = g.dpi; o[H + u] = g.dpf; o[2 * H + u] = g.dpg; o[3 * H + u] = g.dpo;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
      float acc = 0.0f;
      for (int r = 0; r < 4 * H; ++r) acc = fmaf(__ldg(W + (size_t)r * H + k), dp[r], acc);
      dh[k] = acc;
    }
    __syncthreads();
  }
  for (int t = len; t < T; ++t)
    for (int r = threadIdx.x; r < 4 * H; r += blockDim.x) dgx_b[(size_t)t * 8 * H + r] = 0.0f;
}

// ---------------------------------------------------------------------------------------------------------
// H = 256 cluster kernel (persistent over work items; NT tiles of 8 episodes interleaved per item, as forward)
// ---------------------------------------------------------------------------------------------------------
// dp buffer layout (float4 granules): [gate rs][row pair r2 = u>>1][episode pair xp = e>>1] ->
// (dp[u even][e even], dp[u odd][e even], dp[u even][e odd], dp[u odd][e odd]); one granule of padding per gate
// so that the 4 gate slices read by a quarter-warp fall into different banks.  Row pairs arrive as aligned
// register pairs for FFMA2.
constexpr int kDpGateStride4 = 16 * 4 + 1;          // granules per gate slice (16 row pairs x 4 episode pairs + pad)
constexpr int kDpFloats = 4 * kDpGateStride4 * 4;   // floats per buffer
__device__ __forceinline__ int dp_index(int gate, int u, int e) {
  return ((gate * kDpGateStride4 + (u >> 1) * 4 + (e >> 1)) << 2) + ((e & 1) << 1) + (u & 1);
}

template <int NT>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_bwd_cluster_kernel(const float *__restrict__ dy, const float *__restrict__ gates,
                            const float *__restrict__ w_hh_t, const int32_t *__restrict__ lengths,
                            const int32_t *__restrict__ order, int B, int T, int n_enc, int n_tiles,
                            float *__restrict__ dgx) {
  // dynamic shared memory (NT = 2 needs 49.7 KB, above the 48 KB static limit):
  //   recv  [NT][2][8][32][8]  partial dh from every CTA, double buffered
  //   dpbuf [NT][2][kDpFloats] double buffered: step s+1 writes while laggards still read step s
  //   full_bar [NT][2]
  extern __shared__ __align__(16) float smem_f[];
  float(*recv)[2][kCluster][kUnits][kBT] = reinterpret_cast<float(*)[2][kCluster][kUnits][kBT]>(smem_f);
  float(*dpbuf)[2][kDpFloats] = reinterpret_cast<float(*)[2][kDpFloats]>(smem_f + NT * 2 * kCluster * kUnits * kBT);
  uint64_t(*full_bar)[2] =
      reinterpret_cast<uint64_t(*)[2]>(smem_f + NT * 2 * kCluster * kUnits * kBT + NT * 2 * kDpFloats);

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int n_clusters = gridDim.x / kCluster;
  const int groups = (n_tiles + NT - 1) / NT;
  const int n_items = groups * 2 * n_enc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ycols = n_enc * 2 * kH;

  // cell role: (unit u of this CTA, episode slot e);  matvec role: (k-group kg: k = 4 kg .. 4 kg + 3, row slice rs = gate)
  const int u = tid >> 3, e = tid & 7;
  const int unit = rank * kUnits + u;
  const int rs = lane & 3, kg = warp * 8 + (lane >> 2);

  // my partial sums for k = 4 kg + rs go to CTA `warp`, unit row `lane` of its recv[.][.][rank]
  const uint32_t raddr0 = mapa(smem_u32(&recv[0][0][rank][lane][0]), (uint32_t)warp);
  const uint32_t rbar0 = mapa(smem_u32(&full_bar[0][0]), (uint32_t)warp);
  constexpr uint32_t kRecvBytes = kCluster * kUnits * kBT * 4;  // 8 KB

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int grp = item % groups;
    const int dir = (item / groups) & 1;
    const int enc = item / (2 * groups);

    // (W[row 2 r2][k], W[row 2 r2 + 1][k]) for my gate slice rs and my 4 k, read from the TRANSPOSED copy
    // W_hh^T [k][row] so that row pairs are adjacent in memory and land in aligned register pairs for FFMA2.
    float2 Wp[16][4];
    {
      const float *Wt = w_hh_t + ((size_t)enc * 2 + dir) * 4 * kH * kH;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 *col = reinterpret_cast<const float4 *>(Wt + (size_t)(kg * 4 + q) * 4 * kH + rs * kH + rank * kUnits);
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          const float4 v = __ldg(col + r4);
          Wp[2 * r4][q] = make_float2(v.x, v.y);
          Wp[2 * r4 + 1][q] = make_float2(v.z, v.w);
        }
      }
    }

    int bq[NT], len[NT];
    int nsteps = 0;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int tile = grp * NT + j;
      const int slot = tile * kBT + e;
      bq[j] = (tile < n_tiles && slot < B) ? (order ? order[slot] : slot) : -1;
      len[j] = (bq[j] >= 0) ? min(max(lengths[bq[j]], 0), T) : 0;
      nsteps = max(nsteps, len[j]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));

    for (int i = tid; i < NT * 2 * kCluster * kUnits * kBT; i += kThreads) (&recv[0][0][0][0][0])[i] = 0.0f;
    for (int i = tid; i < NT * 2 * kDpFloats; i += kThreads) (&dpbuf[0][0][0])[i] = 0.0f;
    if (tid == 0) {
#pragma unroll
      for (int j = 0; j < NT * 2; ++j) mbar_init(smem_u32(&full_bar[0][0] + j), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    const size_t ycol = (size_t)enc * 2 * kH + dir * kH + unit;
    const size_t gate_base = ((size_t)enc * 2 + dir) * B;
    const size_t dgx_enc = (size_t)enc * B * T * 8 * kH;
    const int gcol = dir * 4 * kH + unit;

    // time visited at backward step s by tile j's cell, and prefetch of its saved activations
    float ig[NT], fg[NT], gg[NT], og[NT], dyv[NT], c_cur[NT], c_nxt[NT], dc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      ig[j] = fg[j] = gg[j] = og[j] = dyv[j] = c_cur[j] = c_nxt[j] = dc[j] = 0.0f;
      if (len[j] > 0) {
        const int t0 = dir ? 0 : len[j] - 1;
        const float *gs = gates + ((gate_base + bq[j]) * T + t0) * 5 * kH + unit;
        ig[j] = __ldg(gs); fg[j] = __ldg(gs + kH); gg[j] = __ldg(gs + 2 * kH); og[j] = __ldg(gs + 3 * kH);
        c_cur[j] = __ldg(gs + 4 * kH);
        dyv[j] = __ldg(dy + ((size_t)bq[j] * T + t0) * ycols + ycol);
        if (len[j] > 1) {
          const int t1 = dir ? 1 : len[j] - 2;
          c_nxt[j] = __ldg(gates + ((gate_base + bq[j]) * T + t1) * 5 * kH + 4 * kH + unit);
        }
      }
    }

    for (int s = 0; s < nsteps; ++s) {
      const int p = s & 1;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (tid == 0 && s + 1 < nsteps) mbar_arrive_expect_tx(smem_u32(&full_bar[j][p ^ 1]), kRecvBytes);
        if (s > 0) mbar_wait(smem_u32(&full_bar[j][p]), ((s - 1) >> 1) & 1);

        // ---- cell phase --------------------------------------------------------------------------------------
        const bool active = s < len[j];
        GateGrads g{0.f, 0.f, 0.f, 0.f, 0.f};
        if (active) {
          float dht = dyv[j];
          if (s > 0) {
#pragma unroll
            for (int src = 0; src < kCluster; ++src) dht += recv[j][p][src][u][e];
          }
          g = lstm_cell_bwd(dht, dc[j], ig[j], fg[j], gg[j], og[j], c_cur[j], (s + 1 < len[j]) ? c_nxt[j] : 0.0f);
          dc[j] = g.dc_prev;
          const int t = dir ? s : len[j] - 1 - s;
          float *o = dgx + dgx_enc + ((size_t)bq[j] * T + t) * 8 * kH + gcol;
          o[0] = g.dpi; o[kH] = g.dpf; o[2 * kH] = g.dpg; o[3 * kH] = g.dpo;
        }
        float *dpb = &dpbuf[j][p][0];
        dpb[dp_index(0, u, e)] = g.dpi;
        dpb[dp_index(1, u, e)] = g.dpf;
        dpb[dp_index(2, u, e)] = g.dpg;
        dpb[dp_index(3, u, e)] = g.dpo;
        if (s + 1 < len[j]) {  // prefetch the next step's activations (and the cell state two steps ahead)
          const int tn = dir ? s + 1 : len[j] - 2 - s;
          const float *gs = gates + ((gate_base + bq[j]) * T + tn) * 5 * kH + unit;
          ig[j] = __ldg(gs); fg[j] = __ldg(gs + kH); gg[j] = __ldg(gs + 2 * kH); og[j] = __ldg(gs + 3 * kH);
          dyv[j] = __ldg(dy + ((size_t)bq[j] * T + tn) * ycols + ycol);
          c_cur[j] = c_nxt[j];
          if (s + 2 < len[j]) {
            const int t2 = dir ? s + 2 : len[j] - 3 - s;
            c_nxt[j] = __ldg(gates + ((gate_base + bq[j]) * T + t2) * 5 * kH + 4 * kH + unit);
          }
        }
        __syncthreads();
        if (s + 1 >= nsteps) continue;  // the last step's dh_prev has no consumer

        // ---- matvec phase: 32 rows (gate rs) x 4 k x 8 episodes, FFMA2 over row pairs ------------------------
        float acc[4][8];
        const float4 *dp4 = reinterpret_cast<const float4 *>(dpb) + rs * kDpGateStride4;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          float2 a2[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int x = 0; x < 4; ++x) a2[q][x] = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int r2 = 0; r2 < 16; ++r2) {
            const float4 da = dp4[r2 * 4 + pass * 2];      // episodes 4 pass + {0,1}, rows 2 r2 + {0,1}
            const float4 db = dp4[r2 * 4 + pass * 2 + 1];  // episodes 4 pass + {2,3}
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              a2[q][0] = __ffma2_rn(Wp[r2][q], make_float2(da.x, da.y), a2[q][0]);
              a2[q][1] = __ffma2_rn(Wp[r2][q], make_float2(da.z, da.w), a2[q][1]);
              a2[q][2] = __ffma2_rn(Wp[r2][q], make_float2(db.x, db.y), a2[q][2]);
              a2[q][3] = __ffma2_rn(Wp[r2][q], make_float2(db.z, db.w), a2[q][3]);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[q][pass * 4 + x] = a2[q][x].x + a2[q][x].y;
        }
        // reduce over the 4 row slices (gates); lane rs keeps k = 4 kg + rs
        float r1[2][8], out[8];
        const bool b1 = rs & 2, b0 = rs & 1;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const float keep = b1 ? acc[2 + q][x] : acc[q][x];
            const float send = b1 ? acc[q][x] : acc[2 + q][x];
            r1[q][x] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const float keep = b0 ? r1[1][x] : r1[0][x];
          const float send = b0 ? r1[0][x] : r1[1][x];
          out[x] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        const uint32_t boff = (uint32_t)(j * 2 + (p ^ 1)) * kRecvBytes, moff = (uint32_t)(j * 2 + (p ^ 1)) * 8;
        st_async_v4(raddr0 + boff, make_float4(out[0], out[1], out[2], out[3]), rbar0 + moff);
        st_async_v4(raddr0 + boff + 16, make_float4(out[4], out[5], out[6], out[7]), rbar0 + moff);
      }
    }
    // zero the padded tail of my (episode, unit) gate columns
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (bq[j] >= 0)
        for (int t = len[j]; t < T; ++t) {
          float *o = dgx + dgx_enc + ((size_t)bq[j] * T + t) * 8 * kH + gcol;
          o[0] = 0.0f; o[kH] = 0.0f; o[2 * kH] = 0.0f; o[3 * kH] = 0.0f;
        }
    cluster.sync();
  }
}

constexpr size_t bwd_smem_bytes(int nt) {
  return (size_t)nt * 2 * (kCluster * kUnits * kBT + kDpFloats) * sizeof(float) + (size_t)nt * 2 * sizeof(uint64_t);
}

template <typename K>
static int max_active_clusters_bwd(K kernel, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCluster * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kCluster;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 14;
  }
  return n;
}

}  // namespace mts

using namespace mts;

extern "C" int mts_lstm_rec_bwd(const float *dy, const float *gates, const float *w_hh, const float *w_hh_t,
                                const int32_t *lengths, const int32_t *order, int n_enc, int B, int T, int H, float *dgx,
                                void *stream) {
  MTS_REQUIRE(dy && gates && w_hh && w_hh_t && lengths && dgx, MTS_E_BADARG, "lstm_rec_bwd: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0 && H > 0, MTS_E_BADARG, "lstm_rec_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (H == kH) {
    constexpr size_t smem1 = bwd_smem_bytes(1), smem2 = bwd_smem_bytes(2);
    MTS_PER_DEVICE(int, cap);
    if (!cap) {
      MTS_CUDA(cudaFuncSetAttribute(lstm_bwd_cluster_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      cap = max_active_clusters_bwd(lstm_bwd_cluster_kernel<1>, smem1);
    }
    const int n_tiles = (B + kBT - 1) / kBT;
    const int items1 = n_tiles * 2 * n_enc;
    if (items1 <= cap) {
      lstm_bwd_cluster_kernel<1><<<(unsigned)(items1 * kCluster), kThreads, smem1, st>>>(dy, gates, w_hh_t, lengths, order,
                                                                                         B, T, n_enc, n_tiles, dgx);
    } else {
      const int items2 = ((n_tiles + 1) / 2) * 2 * n_enc;
      lstm_bwd_cluster_kernel<2><<<(unsigned)((items2 < cap ? items2 : cap) * kCluster), kThreads, smem2, st>>>(
          dy, gates, w_hh_t, lengths, order, B, T, n_enc, n_tiles, dgx);
    }
  } else {
    MTS_REQUIRE(H <= 2048, MTS_E_UNSUPPORTED, "lstm_rec_bwd: H > 2048 not supported by the generic kernel");
    const size_t smem = (size_t)6 * H * sizeof(float);
    lstm_bwd_generic_kernel<<<dim3(B, 2, n_enc), 256, smem, st>>>(dy, gates, w_hh, lengths, B, T, H, n_enc, dgx);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

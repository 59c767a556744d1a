// LSTM recurrence, forward.  Restates what torch.nn.LSTM does for the reference at
// models/NeuralArchitectures.py:113-115 (packed variable-length, bidirectional, zero initial state,
// gate rows i,f,g,o), given the hoisted input projection gx = X W_ih^T + b_ih + b_hh.
//
// H == 256: persistent thread-block-cluster kernel.
//   * one cluster of 8 CTAs per (tile of 8 episodes, direction, encoder); CTA r owns hidden units
//     [32r, 32r+32) = 128 gate rows of W_hh, held in REGISTERS for the whole sequence (128 regs/thread);
//   * per step every CTA computes its 128x8 pre-activations with fp32 FMAs (exact fp32 parity path),
//     applies the gates, and pushes its 32x8 slice of h_t into all 8 CTAs' shared memory with
//     st.async (DSMEM) completing on an mbarrier -- no cluster-wide barrier on the critical path;
//   * episodes are length-sorted into tiles on the host, so a tile runs max(len) steps.
//   HBM traffic per sentence-layer-direction: gx 4 KB in, h 1 KB out (+5 KB saved gates in training);
//   the kernel is latency/FMA bound, not HBM bound, at the reference's batch sizes.
// other H: a generic one-CTA-per-episode kernel (correctness path for small models and the golden fixtures).
#include <cooperative_groups.h>

#include "common.cuh"
#include "cluster_utils.cuh"

namespace cg = cooperative_groups;

namespace mts {

// ---------------------------------------------------------------------------------------------------------
// generic kernel: grid (B, 2, n_enc), 256 threads.  smem: h[H], c[H], pre[4H]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_fwd_generic_kernel(const float *__restrict__ gx,
                                                               const float *__restrict__ w_hh,
                                                               const int32_t *__restrict__ lengths, int B, int T,
                                                               int H, int n_enc, float *__restrict__ y,
                                                               float *__restrict__ gates) {
  extern __shared__ float smem[];
  float *h = smem, *c = smem + H, *pre = smem + 2 * H;
  const int b = blockIdx.x, dir = blockIdx.y, enc = blockIdx.z;
  const int len = min(max(lengths[b], 0), T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * H * H;
  const float *gx_e = gx + (size_t)enc * B * T * 8 * H;
  const int ycols = n_enc * 2 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) { h[u] = 0.0f; c[u] = 0.0f; }
  __syncthreads();
  for (int s = 0; s < len; ++s) {
    const int t = dir ? len - 1 - s : s;
    const float *g_row = gx_e + ((size_t)b * T + t) * 8 * H + (size_t)dir * 4 * H;
    for (int r = warp; r < 4 * H; r += nwarps) {
      float acc = 0.0f;
      for (int k = lane; k < H; k += 32) acc += __ldg(W + (size_t)r * H + k) * h[k];
      acc = warp_sum(acc);
      if (lane == 0) pre[r] = acc + __ldg(g_row + r);
    }
    __syncthreads();
    for (int u = threadIdx.x; u < H; u += blockDim.x) {
      const float ig = sigmoidf_acc(pre[u]), fg = sigmoidf_acc(pre[H + u]);
      const float gg = tanhf(pre[2 * H + u]), og = sigmoidf_acc(pre[3 * H + u]);
      const float cn = fg * c[u] + ig * gg;
      const float hn = og * tanhf(cn);
      c[u] = cn;
      h[u] = hn;
      y[((size_t)b * T + t) * ycols + (size_t)enc * 2 * H + dir * H + u] = hn;
      if (gates) {
        float *gs = gates + ((((size_t)enc * 2 + dir) * B + b) * T + t) * 5 * H;
        gs[u] = ig; gs[H + u] = fg; gs[2 * H + u] = gg; gs[3 * H + u] = og; gs[4 * H + u] = cn;
      }
    }
    __syncthreads();
  }
  for (int t = len; t < T; ++t)
    for (int u = threadIdx.x; u < H; u += blockDim.x)
      y[((size_t)b * T + t) * ycols + (size_t)enc * 2 * H + dir * H + u] = 0.0f;
}

// One work item = NT tiles of 8 episodes of one (direction, encoder), advanced in lock-step and INTERLEAVED:
// while tile 0's h_t is in flight through distributed shared memory, the CTA computes tile 1's step, and vice
// versa, so the exchange latency is hidden behind FMAs.  Clusters are persistent over work items.
template <bool SAVE, int NT>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    lstm_fwd_cluster_kernel(const float *__restrict__ gx, const float *__restrict__ w_hh,
                            const int32_t *__restrict__ lengths, const int32_t *__restrict__ order, int B, int T,
                            int n_enc, int n_tiles, float *__restrict__ y, float *__restrict__ gates) {
  __shared__ __align__(16) float hbuf[NT][2][kHBufFloats];
  __shared__ __align__(8) uint64_t full_bar[NT][2];

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int n_clusters = gridDim.x / kCluster;
  const int groups = (n_tiles + NT - 1) / NT;  // tile groups per (direction, encoder)
  const int n_items = groups * 2 * n_enc;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = lane & 7;                  // k-slice: k in [32 ks, 32 ks + 32); after the reduction: my episode slot
  const int u = warp * 4 + (lane >> 3);     // hidden unit inside this CTA, 0..31
  const int unit = rank * kUnits + u;       // hidden unit inside the layer, 0..255
  const int ycols = n_enc * 2 * kH;

  // h buffer layout (float4 granules): [half = e>>2][epair = (e>>1)&1][k2 = (k&31)>>1][ks = k>>5] ->
  // (h[k even][e even], h[k odd][e even], h[k even][e odd], h[k odd][e odd]).  A quarter-warp (8 k-slices) reads
  // 128 contiguous bytes per LDS.128: conflict-free, and k-pairs arrive as aligned register pairs for FFMA2.
  // Remote stores: the 4 lanes holding (unit pair) x (episode pair) share one granule; lane gi of the group
  // sends it to CTAs {2 gi, 2 gi + 1}.  (The mapped shared::cluster window is linear in the CTA-local offset,
  // so tile / buffer offsets just add.)
  const int gi = ((lane >> 3) & 1) * 2 + (lane & 1);
  const uint32_t dst0 = 2 * gi, dst1 = dst0 + 1;
  const uint32_t la = smem_u32(&hbuf[0][0][((((ks >> 2) * 2 + ((ks >> 1) & 1)) * 16 + (u >> 1)) * 8 + rank) * 4]);
  const uint32_t lb = smem_u32(&full_bar[0][0]);
  const uint32_t raddr0 = mapa(la, dst0), raddr1 = mapa(la, dst1);
  const uint32_t rbar0 = mapa(lb, dst0), rbar1 = mapa(lb, dst1);

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int grp = item % groups;
    const int dir = (item / groups) & 1;
    const int enc = item / (2 * groups);

    // ---- W_hh slice into registers: 4 gate rows of my unit x my 32 k -------------------------------------
    float Wr[4][32];
    {
      const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * kH * kH;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 *src = reinterpret_cast<const float4 *>(W + (size_t)(g * kH + unit) * kH + ks * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = __ldg(src + q);
          Wr[g][4 * q + 0] = v.x; Wr[g][4 * q + 1] = v.y; Wr[g][4 * q + 2] = v.z; Wr[g][4 * q + 3] = v.w;
        }
      }
    }

    // ---- my cells: (unit, episode slot ks) of each tile ----------------------------------------------------
    int bq[NT], len[NT];
    int nsteps = 0;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int tile = grp * NT + j;
      const int slot = tile * kBT + ks;
      bq[j] = (tile < n_tiles && slot < B) ? (order ? order[slot] : slot) : -1;
      len[j] = (bq[j] >= 0) ? min(max(lengths[bq[j]], 0), T) : 0;
      nsteps = max(nsteps, len[j]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));

    for (int i = tid; i < NT * 2 * kHBufFloats; i += kThreads) (&hbuf[0][0][0])[i] = 0.0f;
    if (tid == 0) {
#pragma unroll
      for (int j = 0; j < NT * 2; ++j) mbar_init(smem_u32(&full_bar[0][0] + j), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();  // every CTA's barriers and zeroed buffers exist before any remote store

    const size_t gx_enc = (size_t)enc * B * T * 8 * kH;
    const int gcol = dir * 4 * kH + unit;
    const size_t ycol = (size_t)enc * 2 * kH + dir * kH + unit;
    const size_t gate_base = ((size_t)enc * 2 + dir) * B;

    float c[NT], gxn[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      c[j] = 0.0f;
#pragma unroll
      for (int g = 0; g < 4; ++g) gxn[j][g] = 0.0f;
      if (len[j] > 0) {
        const int t0 = dir ? len[j] - 1 : 0;
        const float *g_row = gx + gx_enc + ((size_t)bq[j] * T + t0) * 8 * kH + gcol;
#pragma unroll
        for (int g = 0; g < 4; ++g) gxn[j][g] = __ldg(g_row + g * kH);
      }
    }

    for (int s = 0; s < nsteps; ++s) {
      const int p = s & 1;  // buffer holding h_{s-1}
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (tid == 0 && s + 1 < nsteps) mbar_arrive_expect_tx(smem_u32(&full_bar[j][p ^ 1]), kHBufFloats * 4);
        if (s > 0) mbar_wait(smem_u32(&full_bar[j][p]), ((s - 1) >> 1) & 1);

        const bool active = s < len[j];
        float gxc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) gxc[g] = gxn[j][g];
        if (s + 1 < len[j]) {  // prefetch the next step's input projection
          const int tn = dir ? len[j] - 2 - s : s + 1;
          const float *g_row = gx + gx_enc + ((size_t)bq[j] * T + tn) * 8 * kH + gcol;
#pragma unroll
          for (int g = 0; g < 4; ++g) gxn[j][g] = __ldg(g_row + g * kH);
        }

        // ---- 4 gate rows x 8 episodes x my 32 k, on the packed fp32 pipe ------------------------------------
        // Blackwell issues scalar FFMA at half rate; FFMA2 (fma.rn.f32x2) restores the full fp32 rate.  Pairs
        // run over consecutive k: (W[g][k], W[g][k+1]) * (h[k][e], h[k+1][e]) -> (even-k sum, odd-k sum), added
        // at the end; two passes of 4 episodes keep the accumulators at 32 registers.
        float acc[4][8];
        const float4 *hp = reinterpret_cast<const float4 *>(&hbuf[j][p][0]);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          float2 a2[4][4];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) a2[g][e] = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const float4 ha = hp[((pass * 2 + 0) * 16 + k2) * 8 + ks];  // episodes 4 pass + {0,1}, k = 2 k2 + {0,1}
            const float4 hb = hp[((pass * 2 + 1) * 16 + k2) * 8 + ks];  // episodes 4 pass + {2,3}
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float2 w2 = make_float2(Wr[g][2 * k2], Wr[g][2 * k2 + 1]);
              a2[g][0] = __ffma2_rn(w2, make_float2(ha.x, ha.y), a2[g][0]);
              a2[g][1] = __ffma2_rn(w2, make_float2(ha.z, ha.w), a2[g][1]);
              a2[g][2] = __ffma2_rn(w2, make_float2(hb.x, hb.y), a2[g][2]);
              a2[g][3] = __ffma2_rn(w2, make_float2(hb.z, hb.w), a2[g][3]);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[g][pass * 4 + e] = a2[g][e].x + a2[g][e].y;
        }
        // ---- reduce over the 8 k-slices by recursive halving; lane ks ends with episode slot ks -------------
        float r1[4][4], r2[4][2], pre[4];
        const bool b2 = ks & 4, b1 = ks & 2, b0 = ks & 1;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float keep = b2 ? acc[g][4 + e] : acc[g][e];
            const float send = b2 ? acc[g][e] : acc[g][4 + e];
            r1[g][e] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float keep = b1 ? r1[g][2 + e] : r1[g][e];
            const float send = b1 ? r1[g][e] : r1[g][2 + e];
            r2[g][e] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float keep = b0 ? r2[g][1] : r2[g][0];
          const float send = b0 ? r2[g][0] : r2[g][1];
          pre[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1) + gxc[g];
        }
        // ---- gates, cell, output ----------------------------------------------------------------------------
        float hn = 0.0f;
        if (active) {
          const float ig = sigmoid_fast(pre[0]), fg = sigmoid_fast(pre[1]);
          const float gg = tanh_fast(pre[2]), og = sigmoid_fast(pre[3]);
          c[j] = fmaf(fg, c[j], ig * gg);
          hn = og * tanh_fast(c[j]);
          const int t = dir ? len[j] - 1 - s : s;
          y[((size_t)bq[j] * T + t) * ycols + ycol] = hn;
          if (SAVE) {
            float *gs = gates + ((gate_base + bq[j]) * T + t) * 5 * kH + unit;
            gs[0] = ig; gs[kH] = fg; gs[2 * kH] = gg; gs[3 * kH] = og; gs[4 * kH] = c[j];
          }
        }
        if (s + 1 < nsteps) {  // broadcast h_s: 2 units x 2 episodes per 16-byte remote store, 2 stores per lane
          const int l0 = lane & ~9;  // group = lanes {l0, l0+8, l0+1, l0+9}: (even unit, odd unit) x (even ep, odd ep)
          float4 v;
          v.x = __shfl_sync(0xffffffffu, hn, l0);
          v.y = __shfl_sync(0xffffffffu, hn, l0 + 8);
          v.z = __shfl_sync(0xffffffffu, hn, l0 + 1);
          v.w = __shfl_sync(0xffffffffu, hn, l0 + 9);
          const uint32_t boff = (uint32_t)(j * 2 + (p ^ 1)) * (kHBufFloats * 4), moff = (uint32_t)(j * 2 + (p ^ 1)) * 8;
          st_async_v4(raddr0 + boff, v, rbar0 + moff);
          st_async_v4(raddr1 + boff, v, rbar1 + moff);
        }
      }
    }
    // zero the padded tail of my (episode, unit) columns
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (bq[j] >= 0)
        for (int t = len[j]; t < T; ++t) y[((size_t)bq[j] * T + t) * ycols + ycol] = 0.0f;
    cluster.sync();  // nobody moves on (or exits) while a peer may still address its shared memory
  }
}

// number of 8-CTA clusters of this kernel the device can hold at once (GPC granularity: ~14 on a B200)
template <typename K>
static int max_active_clusters(K kernel) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCluster * 64);
  cfg.blockDim = dim3(kThreads);
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

extern "C" int mts_lstm_rec_fwd(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order,
                                int n_enc, int B, int T, int H, float *y, float *gates, void *stream) {
  MTS_REQUIRE(gx && w_hh && lengths && y, MTS_E_BADARG, "lstm_rec_fwd: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0 && H > 0, MTS_E_BADARG, "lstm_rec_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (H == kH) {
    MTS_PER_DEVICE(int, cap);
    if (!cap) cap = max_active_clusters(lstm_fwd_cluster_kernel<false, 1>);
    const int n_tiles = (B + kBT - 1) / kBT;
    const int items1 = n_tiles * 2 * n_enc;
    // one tile per cluster while everything fits in a single wave; otherwise two interleaved tiles per cluster
    if (items1 <= cap) {
      const unsigned grid = (unsigned)(items1 * kCluster);
      if (gates) lstm_fwd_cluster_kernel<true, 1><<<grid, kThreads, 0, st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, y, gates);
      else lstm_fwd_cluster_kernel<false, 1><<<grid, kThreads, 0, st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, y, gates);
    } else {
      const int items2 = ((n_tiles + 1) / 2) * 2 * n_enc;
      const unsigned grid = (unsigned)((items2 < cap ? items2 : cap) * kCluster);
      if (gates) lstm_fwd_cluster_kernel<true, 2><<<grid, kThreads, 0, st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, y, gates);
      else lstm_fwd_cluster_kernel<false, 2><<<grid, kThreads, 0, st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, y, gates);
    }
  } else {
    MTS_REQUIRE(H <= 2048, MTS_E_UNSUPPORTED, "lstm_rec_fwd: H > 2048 not supported by the generic kernel");
    const size_t smem = (size_t)6 * H * sizeof(float);
    lstm_fwd_generic_kernel<<<dim3(B, 2, n_enc), 256, smem, st>>>(gx, w_hh, lengths, B, T, H, n_enc, y, gates);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

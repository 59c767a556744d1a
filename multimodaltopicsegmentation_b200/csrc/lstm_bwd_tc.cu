// LSTM recurrence, backward through time, on the 5th-generation tensor cores (H == 256).
// Same contract as lstm_bwd.cu: produces dgx = d loss / d (gate pre-activations) for every valid step from dy and
// the gates saved by the forward pass; equals autograd through torch.nn.LSTM (models/NeuralArchitectures.py:113).
//
// One cluster of 8 CTAs walks a tile of 16 episodes of one (direction, encoder) backwards.  CTA r owns the cells of
// hidden units [32r, 32r+32) and the matching 128 gate rows of W_hh.  Per step:
//   cell phase  (4 epilogue warps, thread = 1 unit x 4 episodes): dh = dy + sum of the 8 partial products received
//               from the cluster, gate derivatives, dp -> dgx in HBM and -> the B operand (hi/lo) in shared memory;
//   matvec      partial dh_prev[256 x 16] = W_hh[my 128 rows, :]^T dp[my 128 rows, 16]: 64 tcgen05.mma 128x16
//               (two M blocks x 4 gate blocks x (4 kind::tf32 K = 8 + 4 bf16 K = 16): TF32 product + packed bf16
//               correction product, common.cuh), the transposed weight slice resident in TENSOR MEMORY (raw fp32
//               2 x 128 columns, packed correction 2 x 96 columns + a 32 KB shared-memory tail), accumulators in TMEM;
//   reduce-scatter: each CTA reads its accumulators (tcgen05.ld) and pushes the 32 x 16 block of every owner CTA
//               into that CTA's receive buffer with st.async (DSMEM) completing on its mbarrier -- 16 KB per CTA per
//               step, the same volume as the forward all-gather.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cluster_utils.cuh"
#include "tcgen05_utils.cuh"

namespace cg = cooperative_groups;

namespace mts {

constexpr int TB_NB = 16;
constexpr int TB_THREADS = 160;
constexpr int TB_EPI = 128;
constexpr int TB_TAIL_BYTES = 2 * 128 * 128;            // W_lo^T, K-block 3, both M blocks
constexpr int TB_B_BYTES = 4 * TB_NB * 128;             // dp operand: 4 K-blocks (= gates) x 16 rows x 128 B
constexpr int TB_ROW = 20;                              // floats per received unit row: 16 episodes + 4 pad (bank spread)
constexpr int TB_RECV_FLOATS = 8 * 32 * TB_ROW;         // one receive buffer: [source CTA][unit][TB_ROW]
constexpr int TB_HI0 = 0, TB_HI1 = 128, TB_LO0 = 256, TB_LO1 = 352, TB_ACC0 = 448, TB_ACC1 = 464;
constexpr int TB_SMEM_USED = TB_TAIL_BYTES + 2 * TB_B_BYTES + 2 * TB_RECV_FLOATS * 4 + 512 + 1024;
constexpr int TB_SMEM = TB_SMEM_USED > 120 * 1024 ? TB_SMEM_USED : 120 * 1024;  // one CTA per SM (512 TMEM columns each)

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(TB_THREADS, 1)
    lstm_bwd_tc_kernel(const float *__restrict__ dy, const float *__restrict__ gates, const float *__restrict__ w_hh,
                       const int32_t *__restrict__ lengths, const int32_t *__restrict__ order, int B, int T, int n_enc,
                       int n_tiles, int ept, float *__restrict__ dgx) {
  // ept = episodes per tile (<= TB_NB), as in the forward kernel: a small batch is spread over all resident clusters.
  // Episode slot sl of a tile sits in operand row 4 * (sl % 4) + sl / 4, so that the four cell warps (warp w owns rows
  // 4w .. 4w + 3) share the present episodes evenly; rows of absent slots stay zero and are skipped.
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *wtail = smem;                                    // [2][128 rows x 128 B]
  uint8_t *bhi = wtail + TB_TAIL_BYTES;
  uint8_t *blo = bhi + TB_B_BYTES;
  float *recv = reinterpret_cast<float *>(blo + TB_B_BYTES);  // [2][8][32][TB_ROW]
  uint64_t *bars = reinterpret_cast<uint64_t *>(recv + 2 * TB_RECV_FLOATS);
  uint64_t *part_full = bars;      // [2]  the 8 partial blocks of step s-1 landed in recv[s & 1]
  uint64_t *b_ready = bars + 2;    //      dp operand of the step written
  uint64_t *acc_full = bars + 3;   //      the step's MMAs have completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
  int *len_s = reinterpret_cast<int *>(bars + 6);
  int *bq_s = len_s + TB_NB;

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int n_clusters = gridDim.x / kCluster;
  const int n_items = n_tiles * 2 * n_enc;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int ycols = n_enc * 2 * kH;

  if (tid == 0) {
    tc::bar_init(tc::s_u32(&part_full[0]), 1);
    tc::bar_init(tc::s_u32(&part_full[1]), 1);
    tc::bar_init(tc::s_u32(b_ready), TB_EPI);
    tc::bar_init(tc::s_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc::tmem_alloc<512>(tc::s_u32(tmem_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint32_t ph_part[2] = {0, 0}, ph_b = 0, ph_acc = 0;
  const int nf4 = ept < 4 ? ept : 4;   // groups of 4 operand rows (= cell warps) that hold at least one episode
  const int et = tid - 32;             // 0..127 over the epilogue warps
  const int q = warp & 3;              // TMEM lane quarter this warp may read
  const int cu = et & 31, cg4 = et >> 5;  // cell role: unit cu, episodes 4 cg4 .. 4 cg4 + 3
  int cur_dir = -1, cur_enc = -1;

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int tile = item % n_tiles;
    const int dir = (item / n_tiles) & 1;
    const int enc = item / (2 * n_tiles);
    const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * kH * kH;

    // ---- transposed weight slice on chip: A[m = hidden unit k][kk = my gate row], kk = gate * 32 + unit ----------
    if (dir != cur_dir || enc != cur_enc) {
      if (warp >= 1) {
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {
          const int m = mb * 128 + q * 32 + lane;              // hidden unit (row of W_hh^T) this thread loads
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            float hi[32], cr[32];  // cr: two 16-k blocks of [bf16(W) x16 | bf16(rest W) x16], packed 2 per word
            const float *col = W + (size_t)(kb * kH + rank * kUnits) * kH + m;  // W_hh[gate kb, unit 32 r + j][m]
#pragma unroll
            for (int j = 0; j < 32; ++j) hi[j] = __ldg(col + (size_t)j * kH);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x0 = hi[blk * 16 + 2 * j], x1 = hi[blk * 16 + 2 * j + 1];
                cr[blk * 16 + j] = __uint_as_float(bf16x2_bits(x0, x1));
                cr[blk * 16 + 8 + j] = __uint_as_float(bf16x2_bits(tf32_rest_exact(x0), tf32_rest_exact(x1)));
              }
            tc::tmem_st32(trow + (uint32_t)((mb ? TB_HI1 : TB_HI0) + kb * 32), hi);
            if (kb < 3) {
              tc::tmem_st32(trow + (uint32_t)((mb ? TB_LO1 : TB_LO0) + kb * 32), cr);
            } else {
              const int r = q * 32 + lane;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4 *>(wtail + mb * 16384 + r * 128 + ((i ^ (r & 7)) << 4)) =
                    make_float4(cr[4 * i], cr[4 * i + 1], cr[4 * i + 2], cr[4 * i + 3]);
            }
          }
        }
        tc::tmem_wait_st();
      }
      cur_dir = dir;
      cur_enc = enc;
    }
    if (tid < TB_NB) {
      const int sl = (tid >> 2) + 4 * (tid & 3);   // episode slot held by operand row `tid`
      const int slot = tile * ept + sl;
      const int bq = (sl < ept && slot < B) ? (order ? order[slot] : slot) : -1;
      bq_s[tid] = bq;
      len_s[tid] = (bq >= 0) ? min(max(lengths[bq], 0), T) : 0;
    }
    for (int idx = tid; idx < 2 * TB_B_BYTES / 16; idx += TB_THREADS)   // bhi and blo: rows of absent slots stay zero
      reinterpret_cast<float4 *>(bhi)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    int nsteps = 0;
#pragma unroll
    for (int e = 0; e < TB_NB; ++e) nsteps = max(nsteps, len_s[e]);
    cluster.sync();

    if (warp == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = tc::idesc_tf32(128, TB_NB), idesc_c = tc::idesc_bf16(128, TB_NB);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tail_a = tc::s_u32(wtail), bhi_a = tc::s_u32(bhi), blo_a = tc::s_u32(blo);
      const bool leader = tc::elect_one();
      for (int s = 0; s + 1 < nsteps; ++s) {   // the last step's product would feed nothing
        const int p = s & 1;
        if (leader) tc::bar_expect_tx(tc::s_u32(&part_full[p ^ 1]), 8 * 32 * nf4 * 16);   // nf4 float4 per (source, unit)
        tc::bar_wait_wd(tc::s_u32(b_ready), ph_b); ph_b ^= 1;
        tc::tc_fence_after();
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const uint32_t d_tmem = tb + (mb ? TB_ACC1 : TB_ACC0);
          const uint32_t a_hi = tb + (mb ? TB_HI1 : TB_HI0), a_lo = tb + (mb ? TB_LO1 : TB_LO0);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t bh = tc::desc_sw128(bhi_a + kb * (TB_NB * 128) + k * 32);
              const uint64_t bl = tc::desc_sw128(blo_a + kb * (TB_NB * 128) + k * 32);
              if (leader) {
                tc::umma_tf32_ts(d_tmem, a_hi + (uint32_t)(kb * 32 + k * 8), bh, idesc, (kb | k) != 0);
                if (kb < 3) tc::umma_bf16_ts(d_tmem, a_lo + (uint32_t)(kb * 32 + k * 8), bl, idesc_c, 1);
                else tc::umma_bf16_ss(d_tmem, tc::desc_sw128(tail_a + mb * 16384 + k * 32), bl, idesc_c, 1);
              }
            }
          }
        }
        if (leader) tc::umma_commit(tc::s_u32(acc_full));
        __syncwarp();
      }
    } else {
      // ===================== cell / reduce-scatter warps =====================
      const int unit = (int)rank * kUnits + cu;
      const size_t ycol = (size_t)enc * 2 * kH + dir * kH + unit;
      const size_t gate_base = ((size_t)enc * 2 + dir) * B;
      const size_t dgx_enc = (size_t)enc * B * T * 8 * kH;
      const int gcol = dir * 4 * kH + unit;
      int len[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { len[i] = len_s[4 * cg4 + i]; bq[i] = bq_s[4 * cg4 + i]; }
      const int n_i = min(4, max(0, (ept - cg4 + 3) >> 2));   // episodes this warp's threads carry (rows 4 cg4 + i, i < n_i)
      // saved activations of the step being visited (cur) and prefetched for the next one (nx)
      float ig[4], fg[4], gg[4], og[4], dyv[4], c_cur[4], c_prev[4], dc[4];
      float n_ig[4], n_fg[4], n_gg[4], n_og[4], n_dy[4], n_cp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ig[i] = fg[i] = gg[i] = og[i] = dyv[i] = c_cur[i] = c_prev[i] = dc[i] = 0.0f;
        if (len[i] > 0) {
          const int t0 = dir ? 0 : len[i] - 1;
          const float *gs = gates + ((gate_base + bq[i]) * T + t0) * 5 * kH + unit;
          ig[i] = __ldg(gs); fg[i] = __ldg(gs + kH); gg[i] = __ldg(gs + 2 * kH); og[i] = __ldg(gs + 3 * kH);
          c_cur[i] = __ldg(gs + 4 * kH);
          dyv[i] = __ldg(dy + ((size_t)bq[i] * T + t0) * ycols + ycol);
          if (len[i] > 1) {
            const int t1 = dir ? 1 : len[i] - 2;
            c_prev[i] = __ldg(gates + ((gate_base + bq[i]) * T + t1) * 5 * kH + 4 * kH + unit);
          }
        }
      }
      // reduce-scatter addressing: this thread reads TMEM lane (32 q + lane) of both accumulators: hidden units
      // 32 q + lane (owner CTA q) and 128 + 32 q + lane (owner CTA q + 4); it lands in row `lane` of my source slot there
      const uint32_t row_off = (uint32_t)(((int)rank * 32 + lane) * TB_ROW * 4);
      const uint32_t r_addr0 = mapa(tc::s_u32(recv) + row_off, (uint32_t)q);
      const uint32_t r_addr1 = mapa(tc::s_u32(recv) + row_off, (uint32_t)(q + 4));
      const uint32_t r_bar0 = mapa(tc::s_u32(&part_full[0]), (uint32_t)q);
      const uint32_t r_bar1 = mapa(tc::s_u32(&part_full[0]), (uint32_t)(q + 4));

      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        // ---- dh = dy + the 8 partial products of the previous step ------------------------------------------------
        float dh[4] = {dyv[0], dyv[1], dyv[2], dyv[3]};
        if (s > 0) {
          tc::bar_wait_wd(tc::s_u32(&part_full[p]), ph_part[p]); ph_part[p] ^= 1;
          const float *rb = recv + p * TB_RECV_FLOATS + cu * TB_ROW + 4 * cg4;
#pragma unroll
          for (int src = 0; src < 8 && n_i > 0; ++src) {
            const float4 v = *reinterpret_cast<const float4 *>(rb + src * 32 * TB_ROW);
            dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
          }
        }
        // ---- cell backward: 1 unit x 4 episodes --------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= n_i) break;   // warp-uniform
          const int e = 4 * cg4 + i;
          float dpi = 0.f, dpf = 0.f, dpg = 0.f, dpo = 0.f;
          if (s < len[i]) {
            const float tcv = tanh_fast(c_cur[i]);
            const float d_o = dh[i] * tcv;
            const float dcv = dc[i] + dh[i] * og[i] * (1.0f - tcv * tcv);
            dpi = dcv * gg[i] * ig[i] * (1.0f - ig[i]);
            dpf = dcv * c_prev[i] * fg[i] * (1.0f - fg[i]);
            dpg = dcv * ig[i] * (1.0f - gg[i] * gg[i]);
            dpo = d_o * og[i] * (1.0f - og[i]);
            dc[i] = dcv * fg[i];
            const int t = dir ? s : len[i] - 1 - s;
            float *o = dgx + dgx_enc + ((size_t)bq[i] * T + t) * 8 * kH + gcol;
            o[0] = dpi; o[kH] = dpf; o[2 * kH] = dpg; o[3 * kH] = dpo;
          }
          // B operand: row e (episode), K index = gate * 32 + unit -> K-block = gate.  Raw fp32 for the TF32 product,
          // and the packed correction operand (B side: per 16-k block [bf16(rest) x16 | bf16(dp) x16])
          const uint32_t off = tc::sw128_offset(e, cu);
          const int cb = (cu >> 4) * 64 + (cu & 15) * 2;  // byte offset of bf16(rest dp[cu]) inside the 128-byte row
          const uint32_t off_lo = (uint32_t)(e * 128 + ((((cb >> 4) ^ (e & 7)) << 4) | (cb & 15)));
          const uint32_t off_hb = (uint32_t)(e * 128 + (((((cb + 32) >> 4) ^ (e & 7)) << 4) | (cb & 15)));
          const float dpv[4] = {dpi, dpf, dpg, dpo};
#pragma unroll
          for (int gsel = 0; gsel < 4; ++gsel) {
            *reinterpret_cast<float *>(bhi + gsel * (TB_NB * 128) + off) = dpv[gsel];
            *reinterpret_cast<uint16_t *>(blo + gsel * (TB_NB * 128) + off_lo) = bf16_bits(tf32_rest_exact(dpv[gsel]));
            *reinterpret_cast<uint16_t *>(blo + gsel * (TB_NB * 128) + off_hb) = bf16_bits(dpv[gsel]);
          }
        }
        if (s + 1 < nsteps) {
          tc::fence_proxy_async();   // generic-proxy writes of the operand -> visible to tcgen05.mma
          tc::bar_arrive(tc::s_u32(b_ready));
        }
        // ---- prefetch the saved activations of the next visited step (independent of the recurrence) -------------
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          n_ig[i] = n_fg[i] = n_gg[i] = n_og[i] = n_dy[i] = n_cp[i] = 0.0f;
          if (s + 1 < len[i]) {
            const int tn = dir ? s + 1 : len[i] - 2 - s;
            const float *gs = gates + ((gate_base + bq[i]) * T + tn) * 5 * kH + unit;
            n_ig[i] = __ldg(gs); n_fg[i] = __ldg(gs + kH); n_gg[i] = __ldg(gs + 2 * kH); n_og[i] = __ldg(gs + 3 * kH);
            n_dy[i] = __ldg(dy + ((size_t)bq[i] * T + tn) * ycols + ycol);
            if (s + 2 < len[i]) {
              const int tp = dir ? s + 2 : len[i] - 3 - s;
              n_cp[i] = __ldg(gates + ((gate_base + bq[i]) * T + tp) * 5 * kH + 4 * kH + unit);
            }
          }
        }
        // ---- reduce-scatter of this step's partial products ----------------------------------------------------------
        if (s + 1 < nsteps) {
          tc::bar_wait_wd(tc::s_u32(acc_full), ph_acc); ph_acc ^= 1;
          tc::tc_fence_after();
          float d0[TB_NB], d1[TB_NB];
          tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + TB_ACC0, d0);
          tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + TB_ACC1, d1);
          tc::tc_fence_before();
          const uint32_t boff = (uint32_t)((p ^ 1) * TB_RECV_FLOATS * 4), moff = (uint32_t)((p ^ 1) * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j >= nf4) break;
            st_async_v4(r_addr0 + boff + 16 * j, make_float4(d0[4 * j], d0[4 * j + 1], d0[4 * j + 2], d0[4 * j + 3]), r_bar0 + moff);
            st_async_v4(r_addr1 + boff + 16 * j, make_float4(d1[4 * j], d1[4 * j + 1], d1[4 * j + 2], d1[4 * j + 3]), r_bar1 + moff);
          }
        }
        // ---- rotate the prefetched values in ------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ig[i] = n_ig[i]; fg[i] = n_fg[i]; gg[i] = n_gg[i]; og[i] = n_og[i]; dyv[i] = n_dy[i];
          c_cur[i] = c_prev[i];
          c_prev[i] = n_cp[i];
        }
      }
      // dgx of the padded tail: zeros (the weight-gradient GEMMs read every row)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (bq[i] >= 0)
          for (int t = len[i]; t < T; ++t) {
            float *o = dgx + dgx_enc + ((size_t)bq[i] * T + t) * 8 * kH + gcol;
            o[0] = 0.f; o[kH] = 0.f; o[2 * kH] = 0.f; o[3 * kH] = 0.f;
          }
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mts

using namespace mts;

extern "C" int mts_lstm_rec_bwd_tc(const float *dy, const float *gates, const float *w_hh, const int32_t *lengths,
                                   const int32_t *order, int n_enc, int B, int T, int H, float *dgx, void *stream) {
  // default: the fp16-split kernel (lstm_bwd_h3.cu); MTS_BWD_IMPL=tf32 keeps the TF32 + bf16 formulation of this file
  static const char *impl = getenv("MTS_BWD_IMPL");
  if (impl && impl[0] == 't') return mts_lstm_rec_bwd_tf32(dy, gates, w_hh, lengths, order, n_enc, B, T, H, dgx, stream);
  return mts_lstm_rec_bwd_h3(dy, gates, w_hh, lengths, order, n_enc, B, T, H, dgx, stream);
}

extern "C" int mts_lstm_rec_bwd_tf32(const float *dy, const float *gates, const float *w_hh, const int32_t *lengths,
                                     const int32_t *order, int n_enc, int B, int T, int H, float *dgx, void *stream) {
  MTS_REQUIRE(dy && gates && w_hh && lengths && dgx, MTS_E_BADARG, "lstm_rec_bwd_tc: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0, MTS_E_BADARG, "lstm_rec_bwd_tc: bad shape");
  MTS_REQUIRE(H == kH, MTS_E_UNSUPPORTED, "lstm_rec_bwd_tc: the tensor-core recurrence serves H == 256");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_PER_DEVICE(int, cap);
  if (!cap) {
    MTS_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * 64);
    cfg.blockDim = dim3(TB_THREADS);
    cfg.dynamicSmemBytes = TB_SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kCluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_bwd_tc_kernel, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 8;
    }
    cap = n;
  }
  int ept = TB_NB;   // see mts_lstm_rec_fwd_tc: few episodes per tile when the batch is smaller than 16 x the clusters
  {
    const int per_dir = cap / (2 * n_enc);
    if (per_dir >= 1 && (B + TB_NB - 1) / TB_NB <= per_dir) {
      const int want = (B + per_dir - 1) / per_dir;
      ept = want < 1 ? 1 : (want > TB_NB ? TB_NB : want);
    }
  }
  static const char *force_ept = getenv("MTS_REC_EPT");
  if (force_ept && atoi(force_ept) >= 1 && atoi(force_ept) <= TB_NB) ept = atoi(force_ept);
  const int n_tiles = (B + ept - 1) / ept;
  const int items = n_tiles * 2 * n_enc;
  const unsigned grid = (unsigned)((items < cap ? items : cap) * kCluster);
  lstm_bwd_tc_kernel<<<grid, TB_THREADS, TB_SMEM, st>>>(dy, gates, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, dgx);
  MTS_LAUNCH_CHECK();
  return 0;
}

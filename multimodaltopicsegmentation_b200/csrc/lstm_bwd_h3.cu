// LSTM recurrence, backward through time, on the 5th-generation tensor cores (H == 256) -- fp16-split operands.
// Same contract as lstm_bwd.cu / lstm_bwd_tc.cu: produces dgx = d loss / d (gate pre-activations) for every valid step from
// dy and the gates saved by the forward pass; equals autograd through torch.nn.LSTM (models/NeuralArchitectures.py:113).
//
// Why a second tensor-core formulation.  lstm_bwd_tc.cu computes the partial product W_hh[my rows]^T dp as one kind::tf32
// product (K = 8 per MMA) plus one bf16 correction product: 64 MMAs per step at ~18 cycles each -- the cell warps spend 41 %
// of a step waiting for them (ncu source page of profiles r02l).  Here both operands are split into fp16 pieces, as in the
// forward kernel lstm_rec_h3.cu:
//     W^T_row * 2^s = W1 + W2 + O(2^-22)   (per-row power-of-two scale: max |row| 2^s in [2^13, 2^14))
//     dp_col  * 2^k = D1 + D2 + O(2^-22)   (gradients have no fixed range: every step, every episode column of this CTA's
//                                            128 x 16 operand gets its own power-of-two scale from the column maximum --
//                                            one warp-wide integer max (redux.sync) per episode; the partial products
//                                            of different CTAs are only summed AFTER the epilogue has undone the scales)
//     W^T dp ~= W1 D1 + W2 D1 + W1 D2        3 kind::f16 products of K = 16 per MMA: 48 MMAs per step, issued as one
//                                            straight-line block with base + constant descriptors (11 cycles per MMA)
// Everything else follows lstm_bwd_tc.cu: one cluster of 8 CTAs per tile of <= 16 episodes of one (direction, encoder);
// CTA r owns hidden units [32 r, 32 r + 32) and the matching 128 gate rows; cell phase (thread = 1 unit x 4 episodes) ->
// operand in shared memory -> MMAs (transposed weight slice resident in TENSOR MEMORY: W1 2 x 64 columns, W2 2 x 64
// columns, no shared-memory tail) -> reduce-scatter of the 256 x 16 partial product to the owner CTAs with st.async.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cluster_utils.cuh"
#include "tcgen05_utils.cuh"

namespace cg = cooperative_groups;

namespace mts {
namespace hb {

constexpr int NB = 16;
constexpr int THREADS = 160;
constexpr int EPI = 128;
constexpr int PIECE_BYTES = 2 * NB * 128;          // one piece of dp: 2 k-blocks (64 fp16 each) x 16 rows x 128 B
constexpr int B_BYTES = 2 * PIECE_BYTES;           // [piece][k-block][row][128 B], K-major SWIZZLE_128B
constexpr int ROW = 20;                            // floats per received unit row: 16 episodes + 4 pad (bank spread)
constexpr int RECV_FLOATS = 8 * 32 * ROW;          // one receive buffer: [source CTA][unit][ROW]
constexpr uint32_t COL_W1 = 0, COL_W2 = 128, COL_ACC0 = 256, COL_ACC1 = 272;
constexpr int SMEM_USED = B_BYTES + 2 * RECV_FLOATS * 4 + 1024 + 1024;
constexpr int SMEM = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;  // one CTA per SM (512 TMEM columns each)

// kind::f16 with fp16 operands (format code 0), fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t h2_bits(float first, float second) {  // `first` at the lower address
  const __half2 h = __floats2half2_rn(first, second);
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 h2_floats(uint32_t bits) { return __half22float2(*reinterpret_cast<const __half2 *>(&bits)); }
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// tcgen05.mma kind::f16, A from tensor memory, B descriptor as (lo, hi) words: lo = base + constant (see lstm_rec_h3.cu)
__device__ __forceinline__ void umma_f16_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// exact power of two 2^s that puts a maximum magnitude (given as the bits of |x|) into [2^13, 2^14); 1 for zero / denormal
__device__ __forceinline__ int scale_exp(uint32_t abs_bits) {
  const int ex = (int)((abs_bits >> 23) & 0xFFu);
  int s = ex == 0 ? 0 : 127 + 13 - ex;
  return s > 110 ? 110 : (s < -110 ? -110 : s);
}
__device__ __forceinline__ float pow2f(int s) { return __uint_as_float((uint32_t)(127 + s) << 23); }

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(THREADS, 1)
    lstm_bwd_h3_kernel(const float *__restrict__ dy, const float *__restrict__ gates, const float *__restrict__ w_hh,
                       const int32_t *__restrict__ lengths, const int32_t *__restrict__ order, int B, int T, int n_enc,
                       int n_tiles, int ept, float *__restrict__ dgx) {
  // ept = episodes per tile (<= NB).  Episode slot sl of a tile sits in operand row 4 * (sl % 4) + sl / 4, so that the four
  // cell warps (warp w owns rows 4w .. 4w + 3) share the present episodes evenly; rows of absent slots stay zero.
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *bbuf = smem;                                         // [2 pieces][2 k-blocks][16 rows][128 B]
  float *recv = reinterpret_cast<float *>(bbuf + B_BYTES);      // [2][8][32][ROW]
  float *sd = recv + 2 * RECV_FLOATS;                           // [2 step parities][NB] 2^-k of the dp column of each operand row
  uint64_t *bars = reinterpret_cast<uint64_t *>(sd + 2 * NB);
  uint64_t *part_full = bars;      // [2]  the 8 partial blocks of step s-1 landed in recv[s & 1]
  uint64_t *b_ready = bars + 2;    //      dp operand of the step written
  uint64_t *acc_full = bars + 3;   // [2]  the step's MMAs into accumulator 0 (hidden units 0-127) / 1 (128-255) have completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5);
  int *len_s = reinterpret_cast<int *>(bars + 7);
  int *bq_s = len_s + NB;

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
    tc::bar_init(tc::s_u32(b_ready), EPI);
    tc::bar_init(tc::s_u32(&acc_full[0]), 1);
    tc::bar_init(tc::s_u32(&acc_full[1]), 1);
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
  const int q = warp & 3;              // TMEM lane quarter this warp may access
  const int cu = et & 31, cg4 = et >> 5;  // cell role: unit cu, episodes 4 cg4 .. 4 cg4 + 3
  int cur_dir = -1, cur_enc = -1;
  float rsw0 = 1.0f, rsw1 = 1.0f;      // 2^-s of the weight rows (hidden units 32 q + lane and 128 + 32 q + lane) this thread reads back

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int tile = item % n_tiles;
    const int dir = (item / n_tiles) & 1;
    const int enc = item / (2 * n_tiles);
    const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * kH * kH;

    // ---- transposed weight slice on chip: A[m = hidden unit][kk = my gate row], kk = gate * 32 + unit, as fp16 pieces ----
    if (dir != cur_dir || enc != cur_enc) {
      if (warp >= 1) {
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {
          const int m = mb * 128 + q * 32 + lane;              // hidden unit (row of W_hh^T) this thread loads
          const float *col0 = W + (size_t)(rank * kUnits) * kH + m;   // W_hh[gate kb, unit 32 r + j][m] = col0[(kb kH + j) kH]
          float mx = 0.0f;
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb)
#pragma unroll 8
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(__ldg(col0 + ((size_t)kb * kH + j) * kH)));
          const int sexp = scale_exp(__float_as_uint(mx));
          const float sc = pow2f(sexp);
          if (mb == 0) rsw0 = pow2f(-sexp); else rsw1 = pow2f(-sexp);
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            uint32_t w1[16], w2[16];   // 32 gate rows = 16 columns of two 16-bit values, row 2 c in the low half
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float a = __ldg(col0 + ((size_t)kb * kH + 2 * c) * kH) * sc;
              const float b = __ldg(col0 + ((size_t)kb * kH + 2 * c + 1) * kH) * sc;
              w1[c] = h2_bits(a, b);
              const float2 f = h2_floats(w1[c]);
              w2[c] = h2_bits(a - f.x, b - f.y);
            }
            tmem_st16u(trow + COL_W1 + (uint32_t)(mb * 64 + kb * 16), w1);
            tmem_st16u(trow + COL_W2 + (uint32_t)(mb * 64 + kb * 16), w2);
          }
        }
        tc::tmem_wait_st();
      }
      cur_dir = dir;
      cur_enc = enc;
    }
    if (tid < NB) {
      const int sl = (tid >> 2) + 4 * (tid & 3);   // episode slot held by operand row `tid`
      const int slot = tile * ept + sl;
      const int bq = (sl < ept && slot < B) ? (order ? order[slot] : slot) : -1;
      bq_s[tid] = bq;
      len_s[tid] = (bq >= 0) ? min(max(lengths[bq], 0), T) : 0;
      sd[tid] = sd[NB + tid] = 1.0f;
    }
    for (int idx = tid; idx < B_BYTES / 16; idx += THREADS)   // rows of absent slots stay zero
      reinterpret_cast<float4 *>(bbuf)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    int nsteps = 0;
#pragma unroll
    for (int e = 0; e < NB; ++e) nsteps = max(nsteps, len_s[e]);
    cluster.sync();

    if (warp == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = idesc_f16(128, NB);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const bool leader = tc::elect_one();
      const uint64_t dfull = tc::desc_sw128(tc::s_u32(bbuf));
      const uint32_t dlo = __shfl_sync(0xffffffffu, (uint32_t)dfull, 0), dhi = (uint32_t)(dfull >> 32);
      for (int s = 0; s + 1 < nsteps; ++s) {   // the last step's product would feed nothing
        const int p = s & 1;
        if (leader) tc::bar_expect_tx(tc::s_u32(&part_full[p ^ 1]), 8 * 32 * nf4 * 16);   // nf4 float4 per (source, unit)
        tc::bar_wait_wd(tc::s_u32(b_ready), ph_b); ph_b ^= 1;
        tc::tc_fence_after();
        if (leader) {
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const uint32_t d_tmem = tb + (mb ? COL_ACC1 : COL_ACC0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // k-steps of 16 gate rows
              const uint32_t d1 = dlo + (uint32_t)(((j >> 2) * (NB * 128) + (j & 3) * 32) >> 4);   // D1 piece
              const uint32_t d2 = d1 + (uint32_t)(PIECE_BYTES >> 4);                                 // D2 piece
              const uint32_t a1 = tb + COL_W1 + (uint32_t)(mb * 64 + 8 * j), a2 = a1 + (COL_W2 - COL_W1);
              umma_f16_ts_p(d_tmem, a1, d1, dhi, idesc, j != 0);   // W1 D1
              umma_f16_ts_p(d_tmem, a2, d1, dhi, idesc, 1);        // W2 D1
              umma_f16_ts_p(d_tmem, a1, d2, dhi, idesc, 1);        // W1 D2
            }
            // one commit per accumulator: the reduce-scatter of the first overlaps the MMAs of the second
            tc::umma_commit(tc::s_u32(&acc_full[mb]));
          }
        }
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
      // Everything of the cell derivative that depends only on the SAVED activations is formed off the critical path (here for
      // the first step, after the sends of a step for the next one, while the partial products are in flight):
      //   d_o = dh tcv, dcv = dc + dh fA, dp_i = dcv fI, dp_f = dcv fF, dp_g = dcv fG, dp_o = dh fO, dc' = dcv f
      float tcv[4], fA[4], fI[4], fF[4], fG[4], fO[4];
      auto cell_factors = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tcv[i] = tanh_fast(c_cur[i]);
          fA[i] = og[i] * (1.0f - tcv[i] * tcv[i]);
          fI[i] = gg[i] * ig[i] * (1.0f - ig[i]);
          fF[i] = c_prev[i] * fg[i] * (1.0f - fg[i]);
          fG[i] = ig[i] * (1.0f - gg[i] * gg[i]);
          fO[i] = tcv[i] * og[i] * (1.0f - og[i]);
        }
      };
      cell_factors();
      // reduce-scatter addressing: this thread reads TMEM lane (32 q + lane) of both accumulators: hidden units
      // 32 q + lane (owner CTA q) and 128 + 32 q + lane (owner CTA q + 4); it lands in row `lane` of my source slot there
      const uint32_t row_off = (uint32_t)(((int)rank * 32 + lane) * ROW * 4);
      const uint32_t r_addr0 = mapa(tc::s_u32(recv) + row_off, (uint32_t)q);
      const uint32_t r_addr1 = mapa(tc::s_u32(recv) + row_off, (uint32_t)(q + 4));
      const uint32_t r_bar0 = mapa(tc::s_u32(&part_full[0]), (uint32_t)q);
      const uint32_t r_bar1 = mapa(tc::s_u32(&part_full[0]), (uint32_t)(q + 4));
      // operand addressing of my unit: K index kk = gate * 32 + cu -> k-block gate >> 1, fp16 column (gate & 1) * 32 + cu
      const uint32_t bb = tc::s_u32(bbuf);
      const int colA = cu, colB = 32 + cu;   // gates 0 / 2 and gates 1 / 3

      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        // ---- dh = dy + the 8 partial products of the previous step ------------------------------------------------
        float dh[4] = {dyv[0], dyv[1], dyv[2], dyv[3]};
        if (s > 0) {
          tc::bar_wait_wd(tc::s_u32(&part_full[p]), ph_part[p]); ph_part[p] ^= 1;
          const float *rb = recv + p * RECV_FLOATS + cu * ROW + 4 * cg4;
#pragma unroll
          for (int src = 0; src < 8 && n_i > 0; ++src) {
            const float4 v = *reinterpret_cast<const float4 *>(rb + src * 32 * ROW);
            dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
          }
        }
        // ---- cell backward: 1 unit x 4 episodes --------------------------------------------------------------------
        float dpv[4][4];
        uint32_t mxb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float dpi = 0.f, dpf = 0.f, dpg = 0.f, dpo = 0.f;
          if (i < n_i && s < len[i]) {
            const float dcv = fmaf(dh[i], fA[i], dc[i]);
            dpi = dcv * fI[i];
            dpf = dcv * fF[i];
            dpg = dcv * fG[i];
            dpo = dh[i] * fO[i];
            dc[i] = dcv * fg[i];
            const int t = dir ? s : len[i] - 1 - s;
            float *o = dgx + dgx_enc + ((size_t)bq[i] * T + t) * 8 * kH + gcol;
            o[0] = dpi; o[kH] = dpf; o[2 * kH] = dpg; o[3 * kH] = dpo;
          }
          dpv[i][0] = dpi; dpv[i][1] = dpf; dpv[i][2] = dpg; dpv[i][3] = dpo;
          mxb[i] = __float_as_uint(fmaxf(fmaxf(fabsf(dpi), fabsf(dpf)), fmaxf(fabsf(dpg), fabsf(dpo))));
        }
        // column maxima over my CTA's 128 gate rows: bit patterns of non-negative floats order like integers
#pragma unroll
        for (int i = 0; i < 4; ++i) mxb[i] = __reduce_max_sync(0xffffffffu, mxb[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= n_i) break;   // warp-uniform
          const int e = 4 * cg4 + i;
          const int sexp = scale_exp(mxb[i]);
          const float sc = pow2f(sexp);
          // by step parity: a warp that runs ahead into the next cell phase must not overwrite the scales its slower
          // neighbours are still reading in this step's reduce-scatter (it cannot get two steps ahead: b_ready needs all)
          if (lane == 0) sd[p * NB + e] = pow2f(-sexp);
          // B operand rows: row e (episode), fp16 column of (gate, unit); pieces D1 = fp16(dp 2^k), D2 = fp16(dp 2^k - D1)
          const uint32_t rowb = bb + (uint32_t)(e * 128);
          const uint32_t offA = (uint32_t)((((colA >> 3) ^ (e & 7)) << 4) | ((colA & 7) << 1));
          const uint32_t offB = (uint32_t)((((colB >> 3) ^ (e & 7)) << 4) | ((colB & 7) << 1));
#pragma unroll
          for (int gsel = 0; gsel < 4; ++gsel) {
            const float x = dpv[i][gsel] * sc;
            const __half h1 = __float2half_rn(x);
            const __half h2 = __float2half_rn(x - __half2float(h1));
            const uint32_t a = rowb + (uint32_t)((gsel >> 1) * (NB * 128)) + ((gsel & 1) ? offB : offA);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(__half_as_ushort(h1)) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(a + PIECE_BYTES), "h"(__half_as_ushort(h2)) : "memory");
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
        // ---- reduce-scatter of this step's partial products (scales undone: 2^-s of my weight row x 2^-k of the column) ----
        if (s + 1 < nsteps) {
          const uint32_t boff = (uint32_t)((p ^ 1) * RECV_FLOATS * 4), moff = (uint32_t)((p ^ 1) * 8);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            tc::bar_wait_wd(tc::s_u32(&acc_full[mb]), ph_acc);
            tc::tc_fence_after();
            float d[NB];
            tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (mb ? COL_ACC1 : COL_ACC0), d);
            tc::tc_fence_before();
            const float rsw = mb ? rsw1 : rsw0;
            const uint32_t r_addr = (mb ? r_addr1 : r_addr0) + boff, r_bar = (mb ? r_bar1 : r_bar0) + moff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j >= nf4) break;
              const float4 k4 = *reinterpret_cast<const float4 *>(sd + p * NB + 4 * j);
              st_async_v4(r_addr + 16 * j, make_float4(d[4 * j] * (k4.x * rsw), d[4 * j + 1] * (k4.y * rsw), d[4 * j + 2] * (k4.z * rsw),
                                                       d[4 * j + 3] * (k4.w * rsw)), r_bar);
            }
          }
          ph_acc ^= 1;
        }
        // ---- rotate the prefetched values in ------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ig[i] = n_ig[i]; fg[i] = n_fg[i]; gg[i] = n_gg[i]; og[i] = n_og[i]; dyv[i] = n_dy[i];
          c_cur[i] = c_prev[i];
          c_prev[i] = n_cp[i];
        }
        cell_factors();
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

}  // namespace hb
}  // namespace mts

using namespace mts;

extern "C" int mts_lstm_rec_bwd_h3(const float *dy, const float *gates, const float *w_hh, const int32_t *lengths,
                                   const int32_t *order, int n_enc, int B, int T, int H, float *dgx, void *stream) {
  MTS_REQUIRE(dy && gates && w_hh && lengths && dgx, MTS_E_BADARG, "lstm_rec_bwd_h3: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0, MTS_E_BADARG, "lstm_rec_bwd_h3: bad shape");
  MTS_REQUIRE(H == kH, MTS_E_UNSUPPORTED, "lstm_rec_bwd_h3: the tensor-core recurrence serves H == 256");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_PER_DEVICE(int, cap);
  if (!cap) {
    MTS_CUDA(cudaFuncSetAttribute(hb::lstm_bwd_h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hb::SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * 64);
    cfg.blockDim = dim3(hb::THREADS);
    cfg.dynamicSmemBytes = hb::SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kCluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, hb::lstm_bwd_h3_kernel, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 8;
    }
    cap = n;
  }
  int ept = hb::NB;   // see mts_lstm_rec_fwd_tc: few episodes per tile when the batch is smaller than 16 x the clusters
  {
    const int per_dir = cap / (2 * n_enc);
    if (per_dir >= 1 && (B + hb::NB - 1) / hb::NB <= per_dir) {
      const int want = (B + per_dir - 1) / per_dir;
      ept = want < 1 ? 1 : (want > hb::NB ? hb::NB : want);
    }
  }
  static const char *force_ept = getenv("MTS_REC_EPT");
  if (force_ept && atoi(force_ept) >= 1 && atoi(force_ept) <= hb::NB) ept = atoi(force_ept);
  const int n_tiles = (B + ept - 1) / ept;
  const int items = n_tiles * 2 * n_enc;
  const unsigned grid = (unsigned)((items < cap ? items : cap) * kCluster);
  hb::lstm_bwd_h3_kernel<<<grid, hb::THREADS, hb::SMEM, st>>>(dy, gates, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, dgx);
  MTS_LAUNCH_CHECK();
  return 0;
}

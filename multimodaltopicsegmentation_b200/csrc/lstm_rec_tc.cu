// LSTM recurrence, forward, on the 5th-generation tensor cores (H == 256).
// Same contract as lstm_rec.cu (torch.nn.LSTM as called at models/NeuralArchitectures.py:113-115: packed
// variable-length, bidirectional, zero initial state, gate rows i,f,g,o; gx = X W_ih^T + b_ih + b_hh hoisted).
//
// One cluster of 8 CTAs advances a tile of 16 episodes of one (direction, encoder).  CTA r owns hidden units
// [32r, 32r+32) = 128 gate rows of W_hh and keeps them ON CHIP for the whole sequence, split for 3xTF32:
//     W = W_hi + rest,  W_hi = what kind::tf32 reads of the raw fp32 word (top 19 bits), rest = W - W_hi (exact)
//     W h ~= W_hi h_hi [kind::tf32] + bf16(W) bf16(rest h) + bf16(rest W) bf16(h) [kind::f16/bf16, one packed product:
//            the same error-compensation scheme as the GEMM, common.cuh] -- 64 instead of 96 MMAs per step
//     W raw fp32      : TENSOR MEMORY, 128 lanes x 256 columns                } the A operand of the .ts form of
//     W correction    : TENSOR MEMORY, 128 lanes x 224 columns (bf16 pairs:   } tcgen05.mma: measured ~14 cycles per
//                       per 16 k [bf16(W) | bf16(rest W)], k < 224)           } 128x16 MMA against ~38 when A is read
//     correction tail (k >= 224): shared memory, 16 KB                        } from shared memory
// (2 MB of operands per direction do not fit the shared memory of 8 SMs; TMEM holds all but 16 KB per CTA, the last
//  32 TMEM columns hold the accumulator.)
// Per step, per CTA:
//     pre[128 x 16] = W_hi h_hi + correction               32 + 32 tcgen05.mma 128x16, accumulator in TMEM
//     epilogue warps: tcgen05.ld -> + gx -> sigmoid/tanh -> shared-memory transpose -> cell update (c in
//     registers) -> h_t to HBM and, as raw fp32, straight into the B-operand buffers (K-block r) of all 8 CTAs
//     with st.async (DSMEM) completing on the receivers' mbarriers; the receivers only derive h_lo.
// The only per-step synchronisation is mbarrier-based (no cluster barrier).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cluster_utils.cuh"
#include "tcgen05_utils.cuh"

namespace cg = cooperative_groups;

namespace mts {

constexpr int TR_NB = 16;                 // episodes per tile (= MMA N)
constexpr int TR_SUB_THREADS = 160;       // one tile pipeline: warp 0 MMA issuer, warps 1..4 epilogue
constexpr int TR_EPI = 128;
constexpr int TR_TAIL_BYTES = 128 * 128;           // W_lo, k-block 7: 128 rows x 128 B, K-major SWIZZLE_128B
constexpr int TR_B_BYTES = 8 * TR_NB * 128;        // one h buffer: 8 k-blocks x 16 rows x 128 B = 16 KB
constexpr int TR_ACT_FLOATS = 4 * TR_NB * 32;
constexpr int TR_TMEM_COLS = 512;
constexpr int TR_WLO_COL = 256;                    // W_hi in columns [0, 256), W_lo (k < 224) in [256, 480)
constexpr int TR_ACC_COL = 480;                    // accumulator columns [480, 496) (+16 for the second pipeline)
constexpr int TR_SUB_BYTES = 3 * TR_B_BYTES + TR_ACT_FLOATS * 4 + 1024;  // per tile pipeline: bhi[2], blo, act, barriers
static_assert(TR_SUB_BYTES % 1024 == 0 && TR_TAIL_BYTES % 1024 == 0, "SWIZZLE_128B operand buffers need 1024-byte alignment");
// every CTA allocates all 512 TMEM columns, so two CTAs must never share an SM: ask for more than half its shared memory
template <int NT>
constexpr int tr_smem() {
  return (TR_TAIL_BYTES + NT * TR_SUB_BYTES + 1024) > 120 * 1024 ? (TR_TAIL_BYTES + NT * TR_SUB_BYTES + 1024) : 120 * 1024;
}

// Optional in-kernel timeline (profiling hook, off unless mts_debug_rec_profile() installs a buffer): CTA 0 writes
// clock64() stamps of the phases of steps [8, 8 + TR_PROF_STEPS) -- slots 0..4 by the MMA thread, 5..11 by epilogue
// thread 0 -- so that the per-step critical path can be read without a profiler.
constexpr int TR_PROF_STEPS = 4, TR_PROF_SLOTS = 12;
__device__ long long *g_tr_prof = nullptr;
#define TR_STAMP(slot)                                                                     \
  do {                                                                                     \
    if (prof && s >= 8 && s < 8 + TR_PROF_STEPS) prof[(s - 8) * TR_PROF_SLOTS + (slot)] = clock64(); \
  } while (0)

// NT = tile pipelines per CTA.  NT == 2 runs two independent tiles of 16 episodes of the same (direction, encoder)
// side by side -- each with its own MMA-issuing warp, 4 epilogue warps, operand buffers, barriers and accumulator
// columns -- against ONE resident copy of the weights: while one tile waits for its h exchange, the tensor pipe,
// the MUFU pipe and the DSMEM ports work for the other.  Used when there are more tiles than clusters.
template <bool SAVE, int NT>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(TR_SUB_THREADS * NT, 1)
    lstm_fwd_tc_kernel(const float *__restrict__ gx, const float *__restrict__ w_hh,
                       const int32_t *__restrict__ lengths, const int32_t *__restrict__ order, int B, int T, int n_enc,
                       int n_tiles, int ept, float *__restrict__ y, float *__restrict__ gates,
                       float *__restrict__ y_corr, int bf16_mode) {
  // bf16_mode (explicit precision switch): the step's product is bf16(W_hh) bf16(h) only -- the 32 kind::tf32 MMAs are
  // skipped and the packed operand of h carries [bf16(h) | 0] instead of [bf16(rest h) | bf16(h)], so that the packed
  // halves pair up W * h.  32 instead of 64 MMAs per step; tolerance stated in DESIGN.md.
  // ept = episodes per tile (<= TR_NB).  With fewer sequences than 16 x the resident clusters the host spreads them over
  // ALL clusters: the MMAs cost the same at any N <= 16, while the gate phase, the h exchange and the correction-operand
  // derivation shrink with the number of episodes a cluster carries.  Rows >= ept of the operand buffers stay zero.
  constexpr int THREADS = TR_SUB_THREADS * NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const int sub = (NT == 2 && warp >= 5) ? 1 : 0;         // tile pipeline this warp belongs to
  const int wr = warp - 5 * sub;                           // role inside the pipeline: 0 = MMA issuer, 1..4 = epilogue

  uint8_t *wtail = smem;                                   // shared by both pipelines
  uint8_t *sub_base = wtail + TR_TAIL_BYTES + sub * TR_SUB_BYTES;
  uint8_t *bhi = sub_base;                 // [2][TR_B_BYTES]
  uint8_t *blo = bhi + 2 * TR_B_BYTES;     // [TR_B_BYTES]
  float *act = reinterpret_cast<float *>(blo + TR_B_BYTES);  // [4][NB][32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(act + TR_ACT_FLOATS);
  uint64_t *h_full = bars;         // [2][2]  half g (K-block slots 4g..4g+3, 8 KB) of h_{s-1} landed in bhi[s & 1]
  uint64_t *lo_ready = bars + 4;   // [2]     the same half of blo derived
  uint64_t *acc_full = bars + 6;   //         the step's MMAs have completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TR_TAIL_BYTES + TR_ACT_FLOATS * 4 + 3 * TR_B_BYTES) + 14;  // pipeline 0's bars + 7
  int *len_s = reinterpret_cast<int *>(bars + 8);   // [NB]
  int *bq_s = len_s + TR_NB;                        // [NB]

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int n_clusters = gridDim.x / kCluster;
  const int groups = (n_tiles + NT - 1) / NT;              // tile groups per (direction, encoder)
  const int n_items = groups * 2 * n_enc;
  const int ycols = n_enc * 2 * kH;

  if (wr == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tc::bar_init(tc::s_u32(&h_full[i]), 1);
    tc::bar_init(tc::s_u32(&lo_ready[0]), TR_EPI);
    tc::bar_init(tc::s_u32(&lo_ready[1]), TR_EPI);
    tc::bar_init(tc::s_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc::tmem_alloc<TR_TMEM_COLS>(tc::s_u32(tmem_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_col = TR_ACC_COL + 16 * sub;

  // running mbarrier phases (the barriers are initialised once and live across work items)
  uint32_t ph_h[2] = {0, 0}, ph_lo = 0, ph_acc = 0;

  // epilogue-thread identities
  const int et = (wr - 1) * 32 + lane;  // 0..127 over the pipeline's epilogue warps
  const int q = warp & 3;               // TMEM lane quarter = gate index (i,f,g,o) this warp reads
  const int cj = et & 7, ce = et >> 3;  // cell mapping: units 4 cj .. 4 cj + 3 of episode slot ce

  int cur_dir = -1, cur_enc = -1;
  long long *prof = (blockIdx.x == 0 && (tid == 0 || tid == 32)) ? g_tr_prof : nullptr;

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int tile = (item % groups) * NT + sub;   // may be >= n_tiles for the second pipeline of the last group
    const int dir = (item / groups) & 1;
    const int enc = item / (2 * groups);
    const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * kH * kH;

    // ---- weights on chip (only when the (direction, encoder) changes) ----------------------------------------
    if (dir != cur_dir || enc != cur_enc) {
      if (sub == 0 && wr >= 1) {  // thread = one gate row of the tile: (gate q, unit lane)
        const int r = q * 32 + lane;
        const float *wrow = W + (size_t)(q * kH + rank * kUnits + lane) * kH;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        // K-blocks are stored in ARRIVAL order: slot kk holds the columns of hidden units [32 src, 32 src + 32) with
        // src = (rank - kk) % 8, the CTA whose h slice lands kk-th in this CTA's operand buffer (see the sends below),
        // so that every tensor-memory / shared-memory offset of the step loop is a compile-time constant.
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
          const int src_blk = ((int)rank - kk) & 7;
          float hi[32], cr[32];   // cr: 32 packed words = 64 bf16 = two 16-k blocks of [bf16(W) x16 | bf16(rest W) x16]
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(wrow + src_blk * 32) + i);
            hi[4 * i + 0] = x.x; hi[4 * i + 1] = x.y; hi[4 * i + 2] = x.z; hi[4 * i + 3] = x.w;
            const int blk = i >> 2, j = (i & 3) * 2;  // 16-k block, word pair inside its halves
            cr[blk * 16 + j] = __uint_as_float(bf16x2_bits(x.x, x.y));
            cr[blk * 16 + j + 1] = __uint_as_float(bf16x2_bits(x.z, x.w));
            cr[blk * 16 + 8 + j] = __uint_as_float(bf16x2_bits(tf32_rest_exact(x.x), tf32_rest_exact(x.y)));
            cr[blk * 16 + 8 + j + 1] = __uint_as_float(bf16x2_bits(tf32_rest_exact(x.z), tf32_rest_exact(x.w)));
          }
          tc::tmem_st32(trow + (uint32_t)(kk * 32), hi);  // raw fp32 words: the tensor core reads their top 19 bits
          if (kk < 7) {
            tc::tmem_st32(trow + (uint32_t)(TR_WLO_COL + kk * 32), cr);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4 *>(wtail + r * 128 + ((i ^ (r & 7)) << 4)) =
                  make_float4(cr[4 * i], cr[4 * i + 1], cr[4 * i + 2], cr[4 * i + 3]);
          }
        }
        tc::tmem_wait_st();
      }
      cur_dir = dir;
      cur_enc = enc;
    }
    // ---- tile bookkeeping, zero initial state ---------------------------------------------------------------
    const int st = tid - sub * TR_SUB_THREADS;  // thread index inside the pipeline
    if (st < TR_NB) {
      const int slot = tile * ept + st;
      const int bq = (tile < n_tiles && st < ept && slot < B) ? (order ? order[slot] : slot) : -1;
      bq_s[st] = bq;
      len_s[st] = (bq >= 0) ? min(max(lengths[bq], 0), T) : 0;
    }
    for (int idx = st; idx < TR_B_BYTES / 16; idx += TR_SUB_THREADS) {
      reinterpret_cast<float4 *>(bhi)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);  // h_{-1} = 0 (buffer 0)
      reinterpret_cast<float4 *>(bhi + TR_B_BYTES)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);  // rows >= ept are never sent
      reinterpret_cast<float4 *>(blo)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    int nsteps = 0;
#pragma unroll
    for (int e = 0; e < TR_NB; ++e) nsteps = max(nsteps, len_s[e]);
    cluster.sync();  // every CTA of the cluster is ready to receive

    const size_t gx_enc = (size_t)enc * B * T * 8 * kH;

    if (wr == 0) {
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
      constexpr uint32_t idesc = tc::idesc_tf32(128, TR_NB), idesc_c = tc::idesc_bf16(128, TR_NB);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t d_tmem = tb + acc_col;
      const uint32_t tail_a = tc::s_u32(wtail), blo_a = tc::s_u32(blo);
      const bool leader = tc::elect_one();
      // Every CTA sends its h slice to CTA (rank + i) % 8 at slot i, into K-block slot i of the receiver's buffer, so
      // the slots fill in order; they are tracked as two halves (slots 0-3, 4-7) with one mbarrier each, and the MMAs /
      // the h_lo derivation of the first half overlap the arrival of the second.  h arrives through the async proxy
      // (st.async from the peers): the mbarrier wait alone orders it before the MMAs.  h_lo is derived with ordinary
      // stores + a writer-side fence.proxy.async (measured faster than st.async-to-self: delivery adds ~440 cycles).
      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        TR_STAMP(0);
        if (leader && s + 1 < nsteps) {
          tc::bar_expect_tx(tc::s_u32(&h_full[(p ^ 1) * 2 + 0]), 4 * ept * 128);   // 4 K-block slots x ept rows x 128 B
          tc::bar_expect_tx(tc::s_u32(&h_full[(p ^ 1) * 2 + 1]), 4 * ept * 128);
        }
        const uint32_t bhi_a = tc::s_u32(bhi + p * TR_B_BYTES);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (s > 0) tc::bar_wait_wd(tc::s_u32(&h_full[p * 2 + g]), ph_h[p]);
          if (g == 0) TR_STAMP(1);
          tc::tc_fence_after();
#pragma unroll
          for (int kb = 4 * g; kb < 4 * g + 4; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t bd = tc::desc_sw128(bhi_a + kb * (TR_NB * 128) + k * 32);
              if (leader && !bf16_mode) tc::umma_tf32_ts(d_tmem, tb + (uint32_t)(kb * 32 + k * 8), bd, idesc, (kb | k) != 0);
            }
          }
        }
        TR_STAMP(2);
        if (s > 0 || bf16_mode) {  // at s == 0 both h and h_lo are the zero-filled buffers: the h_lo product contributes nothing
                                   // (bf16 mode: it is the only product, and it zeroes the accumulator)
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (s > 0) tc::bar_wait_wd(tc::s_u32(&lo_ready[g]), ph_lo);
            if (g == 1) TR_STAMP(3);
            tc::tc_fence_after();
#pragma unroll
            for (int kb = 4 * g; kb < 4 * g + 4; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // correction product: 64 bf16 per slot = 4 MMAs of K = 16; A = packed (W | rest W), B = packed (rest h | h)
                const uint64_t bd = tc::desc_sw128(blo_a + kb * (TR_NB * 128) + k * 32);
                if (leader) {
                  const uint32_t acc = (bf16_mode && (kb | k) == 0) ? 0u : 1u;
                  if (kb < 7) tc::umma_bf16_ts(d_tmem, tb + (uint32_t)(TR_WLO_COL + kb * 32 + k * 8), bd, idesc_c, acc);
                  else tc::umma_bf16_ss(d_tmem, tc::desc_sw128(tail_a + k * 32), bd, idesc_c, 1);
                }
              }
            }
          }
        }
        if (s > 0) { ph_h[p] ^= 1; ph_lo ^= 1; }
        if (leader) tc::umma_commit(tc::s_u32(acc_full));
        __syncwarp();
        TR_STAMP(4);
        // the next step's first MMA overwrites the accumulator: it is issued only after h_full[p ^ 1] completes,
        // i.e. after every epilogue thread of this CTA has read its accumulator rows and sent h_s.
      }
    } else {
      // ===================== epilogue warps =====================
      const int gcol = dir * 4 * kH + q * kH + (int)rank * kUnits + lane;   // my gate row inside a gx row
      float gxn[TR_NB];
#pragma unroll
      for (int e = 0; e < TR_NB; ++e) {
        const int len = len_s[e];
        gxn[e] = 0.0f;
        if (len > 0) {
          const int t0 = dir ? len - 1 : 0;
          gxn[e] = __ldg(gx + gx_enc + ((size_t)bq_s[e] * T + t0) * 8 * kH + gcol);
        }
      }
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const int my_len = len_s[ce], my_b = bq_s[ce];
      const size_t ycol = (size_t)enc * 2 * kH + dir * kH + rank * kUnits + 4 * cj;
      const size_t gate_base = ((size_t)enc * 2 + dir) * B;
      // remote addresses of my 16-byte granule in every CTA's bhi[0] and of their h_full[0]
      const uint32_t my_off = tc::sw128_offset(ce, 4 * cj);
      uint32_t raddr[kCluster], rbar[kCluster];
#pragma unroll
      for (int i = 0; i < kCluster; ++i) {  // slot i goes to CTA (rank + i) % 8: self first, then round the ring
        const uint32_t r = (rank + i) & 7;
        raddr[i] = mapa(tc::s_u32(bhi) + (uint32_t)(i * (TR_NB * 128)) + my_off, r);  // K-block slot i of that CTA
        rbar[i] = mapa(tc::s_u32(&h_full[i >> 2]), r);                               // half i / 4 of its buffer 0
      }

      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        TR_STAMP(5);
        // ---- derive h_lo, half by half as the halves land (4 float4 per thread per half) ----------------------------
        if (s > 0) {
          const float4 *src = reinterpret_cast<const float4 *>(bhi + p * TR_B_BYTES);
          float4 *dst = reinterpret_cast<float4 *>(blo);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            tc::bar_wait_wd(tc::s_u32(&h_full[p * 2 + g]), ph_h[p]);
            if (g == 0) TR_STAMP(6);
            float4 hv[4];   // all four loads first: the stores below may alias them as far as the compiler can tell
            const bool my_row = ((et >> 3) & 15) < ept;   // my granules all belong to episode row (et >> 3) & 15
            if (my_row) {
#pragma unroll
              for (int i = 0; i < 4; ++i) hv[i] = src[g * 512 + et + TR_EPI * i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!my_row) break;
              const int f4 = g * 512 + et + TR_EPI * i;   // physical 16-byte granule of the raw-h buffer
              const float4 v = hv[i];
              // granule -> (slot, episode row e, logical 4-k chunk c): the fp32 buffer is SWIZZLE_128B, chunk = phys ^ (e & 7)
              const int e = (f4 >> 3) & 15, c = (f4 & 7) ^ (e & 7);
              // the correction operand of this slot: 64 bf16 per row = two 16-k blocks of [bf16(rest) x16 | bf16(h) x16] (B side)
              const int blk = c >> 2, kk = (c & 3) * 4;              // 16-k block, first k inside it
              uint8_t *rowp = reinterpret_cast<uint8_t *>(dst) + (f4 >> 7) * (TR_NB * 128) + e * 128;
              const int off_lo = blk * 64 + kk * 2, off_hb = off_lo + 32;   // byte offsets inside the 128-byte row
              const uint2 hb = make_uint2(bf16x2_bits(v.x, v.y), bf16x2_bits(v.z, v.w));
              const uint2 rb = make_uint2(bf16x2_bits(tf32_rest_exact(v.x), tf32_rest_exact(v.y)), bf16x2_bits(tf32_rest_exact(v.z), tf32_rest_exact(v.w)));
              // B side: [bf16(rest h) | bf16(h)] pairs with A's [bf16(W) | bf16(rest W)]; bf16 mode: [bf16(h) | 0] -> W * h only
              *reinterpret_cast<uint2 *>(rowp + ((((off_lo >> 4) ^ (e & 7)) << 4) | (off_lo & 15))) = bf16_mode ? hb : rb;
              *reinterpret_cast<uint2 *>(rowp + ((((off_hb >> 4) ^ (e & 7)) << 4) | (off_hb & 15))) = bf16_mode ? make_uint2(0u, 0u) : hb;
            }
            tc::fence_proxy_async();   // generic-proxy writes of h_lo -> visible to tcgen05.mma
            tc::bar_arrive(tc::s_u32(&lo_ready[g]));
          }
          ph_h[p] ^= 1;
        }
        TR_STAMP(7);
        // ---- next step's input projection (independent of h) ---------------------------------------------------
        float gxc[TR_NB];
#pragma unroll
        for (int e = 0; e < TR_NB; ++e) {
          gxc[e] = gxn[e];
          const int len = len_s[e];
          if (s + 1 < len) {
            const int tn = dir ? len - 2 - s : s + 1;
            gxn[e] = __ldg(gx + gx_enc + ((size_t)bq_s[e] * T + tn) * 8 * kH + gcol);
          }
        }
        // ---- accumulator -> gate activations -> shared memory ---------------------------------------------------
        tc::bar_wait_wd(tc::s_u32(acc_full), ph_acc); ph_acc ^= 1;
        TR_STAMP(8);
        tc::tc_fence_after();
        float pre[TR_NB];
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col, pre);
        tc::tc_fence_before();
        TR_STAMP(9);
        // one exponential + one reciprocal per element for every gate: tanh(z) = 2 sigmoid(2 z) - 1  (the MUFU
        // pipe, 4 lanes/clk per sub-partition, bounds this phase: each epilogue warp has a sub-partition to itself)
        const float zs = (q == 2) ? -2.0f * 1.4426950408889634f : -1.4426950408889634f;
        const float oa = (q == 2) ? 2.0f : 1.0f, ob = (q == 2) ? -1.0f : 0.0f;
#pragma unroll
        // fully unrolled over 4, 8 or 16 columns (uniform choice): the MUFU latency needs the independent chains, a
        // branch per column group was measured 60 % slower
#define TR_ACT(NCOL)                                                          \
  _Pragma("unroll") for (int e = 0; e < (NCOL); ++e) {                        \
    const float z = pre[e] + gxc[e];                                          \
    const float sg = __fdividef(1.0f, 1.0f + exp2f(z * zs));                  \
    act[(q * TR_NB + e) * 32 + lane] = fmaf(sg, oa, ob);                      \
  }
        if (ept > 8) { TR_ACT(16) } else if (ept > 4) { TR_ACT(8) } else { TR_ACT(4) }
#undef TR_ACT
        if (sub == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        TR_STAMP(10);
        // ---- cell update: 4 units x 1 episode per thread -----------------------------------------------------
        const float4 ig = *reinterpret_cast<const float4 *>(act + (0 * TR_NB + ce) * 32 + 4 * cj);
        const float4 fg = *reinterpret_cast<const float4 *>(act + (1 * TR_NB + ce) * 32 + 4 * cj);
        const float4 gg = *reinterpret_cast<const float4 *>(act + (2 * TR_NB + ce) * 32 + 4 * cj);
        const float4 og = *reinterpret_cast<const float4 *>(act + (3 * TR_NB + ce) * 32 + 4 * cj);
        float4 hn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < my_len) {
          c[0] = fmaf(fg.x, c[0], ig.x * gg.x); c[1] = fmaf(fg.y, c[1], ig.y * gg.y);
          c[2] = fmaf(fg.z, c[2], ig.z * gg.z); c[3] = fmaf(fg.w, c[3], ig.w * gg.w);
          hn = make_float4(og.x * tanh_fast(c[0]), og.y * tanh_fast(c[1]), og.z * tanh_fast(c[2]), og.w * tanh_fast(c[3]));
        }
        if (s + 1 < nsteps && ce < ept) {  // h_s, raw fp32, into K-block `rank` of every CTA's next B buffer
          const uint32_t boff = (uint32_t)((p ^ 1) * TR_B_BYTES), moff = (uint32_t)((p ^ 1) * 16);
#pragma unroll
          for (int r = 0; r < kCluster; ++r) st_async_v4(raddr[r] + boff, hn, rbar[r] + moff);
        }
        TR_STAMP(11);
        if (s < my_len) {
          const int t = dir ? my_len - 1 - s : s;
          *reinterpret_cast<float4 *>(y + ((size_t)my_b * T + t) * ycols + ycol) = hn;
          // the next layer's GEMM takes y itself as its fp32 operand; its packed bf16 correction operand is written here
          if (y_corr) corr_store4(y_corr + ((size_t)my_b * T + t) * ycols, (int)ycol, hn, 0);
          if (SAVE) {
            float *gs = gates + ((gate_base + my_b) * T + t) * 5 * kH + rank * kUnits + 4 * cj;
            *reinterpret_cast<float4 *>(gs) = ig;
            *reinterpret_cast<float4 *>(gs + kH) = fg;
            *reinterpret_cast<float4 *>(gs + 2 * kH) = gg;
            *reinterpret_cast<float4 *>(gs + 3 * kH) = og;
            *reinterpret_cast<float4 *>(gs + 4 * kH) = make_float4(c[0], c[1], c[2], c[3]);
          }
        }
      }
      // zero the padded tail of my (episode, 4 units) columns
      if (my_b >= 0)
        for (int t = my_len; t < T; ++t) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4 *>(y + ((size_t)my_b * T + t) * ycols + ycol) = z;
          if (y_corr) corr_store4(y_corr + ((size_t)my_b * T + t) * ycols, (int)ycol, z, 0);
        }
    }
    // the MMA thread's phase counters must follow the epilogue's view of h_full (it skips nothing), and vice versa
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();  // nobody re-zeroes buffers (or exits) while a peer may still address its shared memory
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TR_TMEM_COLS>(tmem_base);
  }
  (void)THREADS;
}

template <typename K>
static int tc_max_active_clusters(K kernel, int threads, int smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCluster * 64);
  cfg.blockDim = dim3(threads);
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
    n = 8;
  }
  return n;
}

}  // namespace mts

using namespace mts;

// profiling hook: buf = device buffer of TR_PROF_STEPS * TR_PROF_SLOTS int64 (or NULL to switch the timeline off)
extern "C" int mts_debug_rec_profile(long long *buf) {
  MTS_CUDA(cudaMemcpyToSymbol(g_tr_prof, &buf, sizeof(buf)));
  return 0;
}

static int rec_fwd_tc(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order, int n_enc, int B, int T,
                      int H, float *y, float *gates, float *y_corr, void *stream, int bf16_mode);

// The tensor-core forward recurrence behind the public names: the fp16-split kernel (lstm_rec_h3.cu) unless the
// environment asks for this file's TF32 + bf16 formulation (MTS_REC_TC=tf32), which also stays callable by its own name.
static bool use_tf32_formulation() {
  static const char *v = getenv("MTS_REC_TC");
  return v && v[0] == 't';
}
extern "C" int mts_lstm_rec_fwd_tf32(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order,
                                     int n_enc, int B, int T, int H, float *y, float *gates, float *y_corr, void *stream) {
  return rec_fwd_tc(gx, w_hh, lengths, order, n_enc, B, T, H, y, gates, y_corr, stream, 0);
}
extern "C" int mts_lstm_rec_fwd_tc(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order,
                                   int n_enc, int B, int T, int H, float *y, float *gates, float *y_corr, void *stream) {
  if (use_tf32_formulation()) return rec_fwd_tc(gx, w_hh, lengths, order, n_enc, B, T, H, y, gates, y_corr, stream, 0);
  return mts_lstm_rec_fwd_h3(gx, w_hh, lengths, order, n_enc, B, T, H, y, gates, y_corr, 0, stream);
}
// the bf16 path (explicit precision switch): one bf16(W_hh) bf16(h) product per step, fp32 state and gates
extern "C" int mts_lstm_rec_fwd_tc_bf16(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order,
                                        int n_enc, int B, int T, int H, float *y, float *gates, float *y_corr, void *stream) {
  if (use_tf32_formulation()) return rec_fwd_tc(gx, w_hh, lengths, order, n_enc, B, T, H, y, gates, y_corr, stream, 1);
  return mts_lstm_rec_fwd_h3(gx, w_hh, lengths, order, n_enc, B, T, H, y, gates, y_corr, 1, stream);
}

static int rec_fwd_tc(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order, int n_enc, int B, int T,
                      int H, float *y, float *gates, float *y_corr, void *stream, int bf16_mode) {
  MTS_REQUIRE(gx && w_hh && lengths && y, MTS_E_BADARG, "lstm_rec_fwd_tc: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0, MTS_E_BADARG, "lstm_rec_fwd_tc: bad shape");
  MTS_REQUIRE(H == kH, MTS_E_UNSUPPORTED, "lstm_rec_fwd_tc: the tensor-core recurrence serves H == 256");
  MTS_REQUIRE(!y_corr || n_enc == 1, MTS_E_UNSUPPORTED, "lstm_rec_fwd_tc: the fused correction operand needs n_enc == 1");
  MTS_REQUIRE((n_enc * 2 * kH) % 4 == 0 && (((uintptr_t)y | (uintptr_t)gx | (uintptr_t)w_hh) & 15) == 0, MTS_E_BADARG,
              "lstm_rec_fwd_tc: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  MTS_PER_DEVICE(int, cap);
  if (!cap) {
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem<1>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem<1>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem<2>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem<2>()));
    cap = tc_max_active_clusters(lstm_fwd_tc_kernel<false, 2>, 2 * TR_SUB_THREADS, tr_smem<2>());
  }
  // episodes per tile: 16 when the tiles outnumber the resident clusters; otherwise as few as spreading the batch over
  // all clusters allows (a cluster's step gets shorter with fewer episodes, the MMAs cost the same)
  int ept = TR_NB;
  {
    const int per_dir = cap / (2 * n_enc);   // clusters one (direction, encoder) can have
    if (per_dir >= 1 && (B + TR_NB - 1) / TR_NB <= per_dir) {
      const int want = (B + per_dir - 1) / per_dir;
      ept = want < 1 ? 1 : (want > TR_NB ? TR_NB : want);
    }
  }
  static const char *force_ept = getenv("MTS_REC_EPT");
  if (force_ept && atoi(force_ept) >= 1 && atoi(force_ept) <= TR_NB) ept = atoi(force_ept);
  const int n_tiles = (B + ept - 1) / ept;
  const int items1 = n_tiles * 2 * n_enc;
  // one tile per cluster while everything fits in a single wave; otherwise two tile pipelines per cluster
  static const char *force = getenv("MTS_REC_NT");
  const bool two = force ? (force[0] == '2') : (items1 > cap);
  if (!two) {
    const unsigned grid = (unsigned)((items1 < cap ? items1 : cap) * kCluster);
    if (gates) lstm_fwd_tc_kernel<true, 1><<<grid, TR_SUB_THREADS, tr_smem<1>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr, bf16_mode);
    else lstm_fwd_tc_kernel<false, 1><<<grid, TR_SUB_THREADS, tr_smem<1>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr, bf16_mode);
  } else {
    const int items2 = ((n_tiles + 1) / 2) * 2 * n_enc;
    const unsigned grid = (unsigned)((items2 < cap ? items2 : cap) * kCluster);
    if (gates) lstm_fwd_tc_kernel<true, 2><<<grid, 2 * TR_SUB_THREADS, tr_smem<2>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr, bf16_mode);
    else lstm_fwd_tc_kernel<false, 2><<<grid, 2 * TR_SUB_THREADS, tr_smem<2>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr, bf16_mode);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

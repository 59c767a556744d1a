// LSTM recurrence, forward, fp16-split operands, CTA PAIRS sharing the h tile (tcgen05 cta_group::2).  H == 256.
// Same contract and arithmetic as lstm_rec_h3.cu (W h = W1 h1 + W2 h1 + W1 h2 as kind::f16 products, per-row power-of-two
// weight scale, the sender of h_t writes its fp16 pieces straight into the receivers' operand buffers).  What changes is
// who holds the h tile: the two CTAs of a pair (cluster ranks 2j, 2j+1) issue ONE M = 256 MMA stream (one thread of the
// even CTA) whose B operand -- the 16 episodes' h rows -- is split between them, 8 rows each, and read by the tensor
// cores of both SMs.  Every CTA therefore RECEIVES only half of the episodes' rows and every sender addresses 4 CTAs
// instead of 8: the distributed-shared-memory bytes per CTA and step are halved -- the exchange is what bounds
// lstm_rec_h3.cu, both at cfg1 and at large batches.
//   accumulator column c (= MMA N index): rows 0-7 live in the even CTA, 8-15 in the odd one; episode slot e of the tile
//   sits in column (e & 1) * 8 + (e >> 1), so that both halves fill evenly when a tile carries fewer than 16 episodes
//   K-slots are in SOURCE-RANK order in both CTAs (the M = 256 MMA pairs one A column with the same K index of both halves)
//   barriers: h_full[parity][half] per CTA as before; the odd CTA's MMA warp relays the completion of its halves to the
//   even CTA's peer_full[parity][half] (remote mbarrier arrive); the even CTA's commit multicasts to both acc_full
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cluster_utils.cuh"
#include "tcgen05_utils.cuh"

namespace cg = cooperative_groups;

namespace mts {

constexpr int HP_NB = 16;                  // episodes per tile (= MMA N)
constexpr int HP_SUB_THREADS = 160;        // one tile pipeline: warp 0 MMA issuer, warps 1..4 epilogue
constexpr int HP_EPI = 128;
constexpr int HP_ROWS = HP_NB / 2;           // episode rows of the h tile held by one CTA of a pair
constexpr int HP_PIECE_BYTES = 4 * HP_ROWS * 128;   // one piece (h1 or h2) of one h buffer: 4 k-blocks (64 fp16) x 8 rows x 128 B
constexpr int HP_B_BYTES = 2 * HP_PIECE_BYTES;    // 8 KB: [piece][k-block][row][128 B], K-major SWIZZLE_128B
constexpr int HP_ACT_FLOATS = 4 * HP_NB * 32;
constexpr int HP_TMEM_COLS = 512;
constexpr int HP_W2_COL = 128;             // W1 in columns [0, 128), W2 in [128, 256)
constexpr int HP_ACC_COL = 256;            // accumulators [256, 272) (+16 for the second pipeline)
constexpr int HP_SUB_BYTES = 2 * HP_B_BYTES + HP_ACT_FLOATS * 4 + 1024;  // per tile pipeline: h buffers [2], act, barriers
constexpr int HP_HEAD_BYTES = 1024;        // row scales (128 floats) + TMEM slot
static_assert(HP_SUB_BYTES % 1024 == 0, "SWIZZLE_128B operand buffers need 1024-byte alignment");
// every CTA allocates all 512 TMEM columns, so two CTAs must never share an SM: ask for more than half its shared memory
template <int NT>
constexpr int hp_smem() {
  return (HP_HEAD_BYTES + NT * HP_SUB_BYTES + 1024) > 120 * 1024 ? (HP_HEAD_BYTES + NT * HP_SUB_BYTES + 1024) : 120 * 1024;
}

// kind::f16 with fp16 operands (format code 0), fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t h2_bits(float first, float second) {  // `first` at the lower address
  const __half2 h = __floats2half2_rn(first, second);
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 h2_floats(uint32_t bits) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&bits));
}
__device__ __forceinline__ void st_async_v4u(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// tcgen05.mma kind::f16, A from tensor memory, B descriptor given as (lo, hi) words: the caller derives lo as base + constant
// (re-deriving the whole descriptor from the address cost 4 uniform instructions per MMA: 40 cycles per issued MMA, measured)
__device__ __forceinline__ void umma_f16_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {   // arrives on `bar` of every CTA in mask
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void bar_arrive_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// test_wait (non-blocking poll), not try_wait: a try_wait suspended on this barrier was only woken by its time limit when the
// completing arrive came from another CTA (+760 cycles per relay hop, measured)
__device__ __forceinline__ bool bar_try_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bar_wait_cluster_wd(uint32_t bar, uint32_t parity) {
  if (bar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!bar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem) {   // executed by one warp of BOTH CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
// accumulator column (= MMA N index) of episode slot e of a tile, and back
__device__ __forceinline__ int hp_slot_of_col(int c) { return (c & 7) * 2 + (c >> 3); }

// Optional in-kernel timeline (off unless mts_debug_rec_profile_h3p() installs a buffer): CTA 0 writes clock64() stamps of
// steps [8, 8 + HP_PROF_STEPS) -- slots 0..4 by the MMA warp, 5..11 by epilogue thread 0.
constexpr int HP_PROF_STEPS = 4, HP_PROF_SLOTS = 16;
__device__ long long *g_hp_prof = nullptr;
#define HP_STAMP(slot)                                                                     \
  do {                                                                                     \
    if (prof && s >= 8 && s < 8 + HP_PROF_STEPS) prof[(s - 8) * HP_PROF_SLOTS + (slot)] = clock64(); \
  } while (0)

// NT = tile pipelines per CTA (two independent tiles of the same (direction, encoder) against ONE resident copy of the
// weights: used when there are more tiles than clusters).
template <bool SAVE, int NT, bool BF16>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(HP_SUB_THREADS * NT, 1)
    lstm_fwd_h3p_kernel(const float *__restrict__ gx, const float *__restrict__ w_hh, const int32_t *__restrict__ lengths,
                       const int32_t *__restrict__ order, int B, int T, int n_enc, int n_tiles, int ept,
                       float *__restrict__ y, float *__restrict__ gates, float *__restrict__ y_corr) {
  constexpr bool bf16_mode = BF16;
  // BF16 (explicit precision switch): ONE bf16(W) bf16(h) product per step -- 16 MMAs, only the h1 piece is sent.
  // ept = episodes per tile (<= HP_NB): a small batch is spread over all resident clusters (the MMAs cost the same at any
  // N <= 16, the gate phase and the h exchange shrink with the episodes a cluster carries).  Rows >= ept stay zero.
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const int sub = (NT == 2 && warp >= 5) ? 1 : 0;         // tile pipeline this warp belongs to
  const int wr = warp - 5 * sub;                           // role inside the pipeline: 0 = MMA issuer, 1..4 = epilogue

  float *rscale_s = reinterpret_cast<float *>(smem);       // [128] 2^-s of my gate rows (shared by both pipelines)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 512);
  uint8_t *sub_base = smem + HP_HEAD_BYTES + sub * HP_SUB_BYTES;
  uint8_t *bbuf = sub_base;                                // [2][HP_B_BYTES]
  float *act = reinterpret_cast<float *>(bbuf + 2 * HP_B_BYTES);  // [4][NB][32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(act + HP_ACT_FLOATS);
  uint64_t *h_full = bars;          // [2][2]  half g (K-slots 4g .. 4g+3) of h_{s-1} landed in bbuf[s & 1]
  uint64_t *peer_full = bars + 4;   // [2][2]  even CTA only: the same half has landed in the odd CTA's buffer (relayed by its MMA warp)
  uint64_t *acc_full = bars + 16;   //         the step's MMAs have completed (multicast commit of the even CTA)
  int *len_s = reinterpret_cast<int *>(bars + 18);   // [NB]
  int *bq_s = len_s + HP_NB;                         // [NB]

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const bool even = (rank & 1u) == 0;                      // the CTA of the pair that issues the MMAs
  const int n_clusters = gridDim.x / kCluster;
  const int groups = (n_tiles + NT - 1) / NT;              // tile groups per (direction, encoder)
  const int n_items = groups * 2 * n_enc;
  const int ycols = n_enc * 2 * kH;

  if (wr == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) { tc::bar_init(tc::s_u32(&h_full[i]), 1); tc::bar_init(tc::s_u32(&peer_full[i]), 1); }
    tc::bar_init(tc::s_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_pair<HP_TMEM_COLS>(tc::s_u32(tmem_slot));   // pair-wide: the same columns in both CTAs' tensor memory
  tc::tc_fence_before();
  __syncthreads();
  cluster.sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_col = HP_ACC_COL + 16 * sub;

  uint32_t ph_h = 0, ph_acc = 0;   // ph_h: bit p = phase of the barriers of buffer p   // running mbarrier phases (the barriers live across work items)

  // epilogue-thread identities
  const int et = (wr - 1) * 32 + lane;  // 0..127 over the pipeline's epilogue warps
  const int q = warp & 3;               // TMEM lane quarter = gate index (i,f,g,o) this warp reads
  const int cj = et & 7, ce = et >> 3;  // cell mapping: units 4 cj .. 4 cj + 3 of episode slot ce

  int cur_dir = -1, cur_enc = -1;
  long long *prof = (blockIdx.x == 0 && (tid == 0 || tid == 32)) ? g_hp_prof : nullptr;

  for (int item = blockIdx.x / kCluster; item < n_items; item += n_clusters) {
    const int tile = (item % groups) * NT + sub;   // may be >= n_tiles for the second pipeline of the last group
    const int dir = (item / groups) & 1;
    const int enc = item / (2 * groups);
    const float *W = w_hh + ((size_t)enc * 2 + dir) * 4 * kH * kH;

    // ---- weights on chip (only when the (direction, encoder) changes) ----------------------------------------
    if (dir != cur_dir || enc != cur_enc) {
      if (sub == 0 && wr >= 1) {  // thread = one gate row of my 128: (gate q, unit lane)
        const float *wrow = W + (size_t)(q * kH + rank * kUnits + lane) * kH;
        float mx = 0.0f;
#pragma unroll 8
        for (int i = 0; i < kH / 4; ++i) {
          const float4 x = __ldg(reinterpret_cast<const float4 *>(wrow) + i);
          mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        }
        // exact power-of-two row scale: max |w| 2^s in [2^13, 2^14) -- inside fp16's range (65504) with both pieces of
        // the large entries normal; a zero / denormal row keeps s = 0
        const int ex = (int)((__float_as_uint(mx) >> 23) & 0xFFu);
        int sexp = ex == 0 ? 0 : 127 + 13 - ex;
        sexp = sexp > 110 ? 110 : sexp;
        const float sc = __uint_as_float((uint32_t)(127 + sexp) << 23);
        rscale_s[q * 32 + lane] = __uint_as_float((uint32_t)(127 - sexp) << 23);
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        // K-slots in SOURCE-RANK order: slot kk holds the columns of hidden units [32 kk, 32 kk + 32), the slice CTA kk sends
        // (the M = 256 MMA needs the same K order in both CTAs of the pair)
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
          const int src_blk = kk;
          uint32_t w1[16], w2[16];   // 32 units = 16 columns of two 16-bit values, unit 2 j in the low half
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(wrow + src_blk * 32) + i);
            const float a = x.x * sc, b = x.y * sc, c = x.z * sc, d = x.w * sc;
            if (bf16_mode) {
              w1[2 * i] = bf16x2_bits(a, b);
              w1[2 * i + 1] = bf16x2_bits(c, d);
              w2[2 * i] = w2[2 * i + 1] = 0u;
            } else {
              w1[2 * i] = h2_bits(a, b);
              w1[2 * i + 1] = h2_bits(c, d);
              const float2 f0 = h2_floats(w1[2 * i]), f1 = h2_floats(w1[2 * i + 1]);
              w2[2 * i] = h2_bits(a - f0.x, b - f0.y);
              w2[2 * i + 1] = h2_bits(c - f1.x, d - f1.y);
            }
          }
          tmem_st16u(trow + (uint32_t)(kk * 16), w1);
          tmem_st16u(trow + (uint32_t)(HP_W2_COL + kk * 16), w2);
        }
        tc::tmem_wait_st();
      }
      cur_dir = dir;
      cur_enc = enc;
    }
    // ---- tile bookkeeping, zero initial state ---------------------------------------------------------------
    const int st = tid - sub * HP_SUB_THREADS;  // thread index inside the pipeline
    if (st < HP_NB) {   // st = accumulator column; its episode slot inside the tile is hp_slot_of_col(st)
      const int es = hp_slot_of_col(st);
      const int slot = tile * ept + es;
      const int bq = (tile < n_tiles && es < ept && slot < B) ? (order ? order[slot] : slot) : -1;
      bq_s[st] = bq;
      len_s[st] = (bq >= 0) ? min(max(lengths[bq], 0), T) : 0;
    }
    for (int idx = st; idx < 2 * HP_B_BYTES / 16; idx += HP_SUB_THREADS)   // h_{-1} = 0; rows >= ept are never sent
      reinterpret_cast<float4 *>(bbuf)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    int nsteps = 0;
#pragma unroll
    for (int e = 0; e < HP_NB; ++e) nsteps = max(nsteps, len_s[e]);
    cluster.sync();  // every CTA of the cluster is ready to receive

    const size_t gx_enc = (size_t)enc * B * T * 8 * kH;

    if (wr == 0) {
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
      // M = 256: the pair's 2 x 128 gate rows; N = 16 episode columns, 8 rows of the B tile in each CTA
      constexpr uint32_t idesc = bf16_mode ? tc::idesc_bf16(256, HP_NB) : idesc_f16(256, HP_NB);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t d_tmem = tb + acc_col;
      const bool leader = tc::elect_one();
      const uint64_t dfull = tc::desc_sw128(tc::s_u32(bbuf));
      const uint32_t dlo0 = __shfl_sync(0xffffffffu, (uint32_t)dfull, 0), dhi = (uint32_t)(dfull >> 32);
      // episode columns with a sequence in MY half of the tile (slots of my parity below ept), and the bytes one half of
      // the K-slots (4 source CTAs) brings per step
      const int n_mine = min(HP_ROWS, max(0, (ept - (int)(rank & 1u) + 1) / 2));
      const uint32_t half_tx = (uint32_t)(4 * n_mine * (bf16_mode ? 64 : 128));
      const uint32_t peer_bar = mapa(tc::s_u32(&peer_full[0]), rank & ~1u);      // the even CTA's relay barriers (odd CTA only)
      const uint16_t pair_mask = (uint16_t)(3u << (rank & ~1u));
      // Every CTA arms and watches the mbarriers of its own buffer.  The EVEN CTA issues: a half's MMAs go once that half
      // has landed in both buffers of the pair -- its own barrier plus peer_full, which the ODD CTA's warp completes with a
      // remote arrive when its own barrier does.  h arrives through the async proxy (st.async): the mbarrier waits alone
      // order it before the MMAs.  The 24 MMAs of a half are one straight-line block (see lstm_rec_h3.cu).
      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        HP_STAMP(0);
        if (leader && s + 1 < nsteps) {
          tc::bar_expect_tx(tc::s_u32(&h_full[(p ^ 1) * 2 + 0]), half_tx);
          tc::bar_expect_tx(tc::s_u32(&h_full[(p ^ 1) * 2 + 1]), half_tx);
        }
        const uint32_t dlo = dlo0 + (uint32_t)(p * (HP_B_BYTES >> 4));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (s > 0) {
            tc::bar_wait_wd(tc::s_u32(&h_full[p * 2 + g]), (ph_h >> p) & 1u);
            if (even) bar_wait_cluster_wd(tc::s_u32(&peer_full[p * 2 + g]), (ph_h >> p) & 1u);
            else if (leader) bar_arrive_remote(peer_bar + (uint32_t)((p * 2 + g) * 8));
          }
          HP_STAMP(1 + g);
          tc::tc_fence_after();
          if (leader && even) {
#pragma unroll
            for (int i = 4 * g; i < 4 * g + 4; ++i) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint32_t d1 = dlo + (uint32_t)(((i >> 1) * (HP_ROWS * 128) + (i & 1) * 64 + k * 32) >> 4);   // h1 piece, slot i
                const uint32_t d2 = d1 + (uint32_t)(HP_PIECE_BYTES >> 4);                                          // h2 piece
                const uint32_t a1 = tb + (uint32_t)(16 * i + 8 * k), a2 = a1 + HP_W2_COL;
                umma_f16_ts_p(d_tmem, a1, d1, dhi, idesc, (i | k) != 0);                 // W1 h1
                if (!bf16_mode) {
                  umma_f16_ts_p(d_tmem, a2, d1, dhi, idesc, 1);                          // W2 h1
                  umma_f16_ts_p(d_tmem, a1, d2, dhi, idesc, 1);                          // W1 h2
                }
              }
            }
          }
          HP_STAMP(3 + g);
        }
        if (s > 0) ph_h ^= 1u << p;
        if (leader && even) umma2_commit_mc(tc::s_u32(acc_full), pair_mask);
        __syncwarp();
        // the relaying warp follows the issuing CTA step by step: with no episode in its half (1 episode per tile) nothing
        // else would hold it back, and a relay running two steps ahead would alias the parity of peer_full
        if (!even) { tc::bar_wait_wd(tc::s_u32(acc_full), ph_acc); ph_acc ^= 1; }
        HP_STAMP(5);
        // the next step's first MMA overwrites the accumulators of BOTH CTAs: it is issued only after the even CTA's
        // h_full[p ^ 1][0] completes, which takes sends from the epilogue threads of every CTA of the cluster (every CTA owns
        // units of the episodes in the even half), i.e. every CTA is past its tcgen05.ld of this step.
      }
    } else {
      // ===================== epilogue warps =====================
      const int gcol = dir * 4 * kH + q * kH + (int)rank * kUnits + lane;   // my gate row inside a gx row
      const float rs = rscale_s[q * 32 + lane];
      float gxn[HP_NB];
#pragma unroll
      for (int e = 0; e < HP_NB; ++e) {
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
      // Sends: lanes pair up (cj even / odd); the even lane sends the h1 granule (8 fp16 = units 4 cj .. 4 cj + 7), the odd
      // lane the h2 granule of the same 8 units.  My episode column ce lives in the CTAs of parity ce >> 3, row ce & 7 of their
      // tile; my units are K-slot `rank` there: 4 destinations.
      const int piece = cj & 1, chunk = cj >> 1;
      const int par = ce >> 3, row = ce & 7;
      uint32_t raddr[4], rbar[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t r = (uint32_t)(2 * j + par);
        const uint32_t off = (uint32_t)(piece * HP_PIECE_BYTES + (int)(rank >> 1) * (HP_ROWS * 128) + row * 128 +
                                        ((((int)(rank & 1u) * 4 + chunk) ^ row) << 4));
        raddr[j] = mapa(tc::s_u32(bbuf) + off, r);
        rbar[j] = mapa(tc::s_u32(&h_full[rank >> 2]), r);
      }
      const bool sender = hp_slot_of_col(ce) < ept && !(bf16_mode && piece);
      const int n0 = (ept + 1) / 2;   // active columns: [0, n0) in the even half, [8, 8 + ept / 2) in the odd half

      for (int s = 0; s < nsteps; ++s) {
        const int p = s & 1;
        // ---- next step's input projection (independent of h) ---------------------------------------------------
        float gxc[HP_NB];
#pragma unroll
        for (int e = 0; e < HP_NB; ++e) {
          gxc[e] = gxn[e];
          const int len = len_s[e];
          if (s + 1 < len) {
            const int tn = dir ? len - 2 - s : s + 1;
            gxn[e] = __ldg(gx + gx_enc + ((size_t)bq_s[e] * T + tn) * 8 * kH + gcol);
          }
        }
        // ---- accumulator -> gate activations -> shared memory ---------------------------------------------------
        tc::bar_wait_wd(tc::s_u32(acc_full), ph_acc); ph_acc ^= 1;
        HP_STAMP(10);
        tc::tc_fence_after();
        float pre[HP_NB];
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col, pre);
        tc::tc_fence_before();
        HP_STAMP(11);
        // one exponential + one reciprocal per element for every gate: tanh(z) = 2 sigmoid(2 z) - 1
        const float zs = (q == 2) ? -2.0f * 1.4426950408889634f : -1.4426950408889634f;
        const float oa = (q == 2) ? 2.0f : 1.0f, ob = (q == 2) ? -1.0f : 0.0f;
#define HP_ACT1(e)                                                            \
  {                                                                           \
    const float z = fmaf(pre[e], rs, gxc[e]);                                 \
    const float sg = __fdividef(1.0f, 1.0f + exp2f(z * zs));                  \
    act[(q * HP_NB + (e)) * 32 + lane] = fmaf(sg, oa, ob);                    \
  }
#define HP_ACT(N0)                                                            \
  _Pragma("unroll") for (int e = 0; e < (N0); ++e) { HP_ACT1(e) HP_ACT1(e + 8) }
        // fully unrolled over the columns in use of both halves (uniform choice): the MUFU pipe bounds this phase
        switch (n0) {
          case 1: HP_ACT(1) break;
          case 2: HP_ACT(2) break;
          case 3: HP_ACT(3) break;
          case 4: HP_ACT(4) break;
          case 5: HP_ACT(5) break;
          case 6: HP_ACT(6) break;
          case 7: HP_ACT(7) break;
          default: HP_ACT(8) break;
        }
#undef HP_ACT1
#undef HP_ACT
        if (sub == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        HP_STAMP(12);
        // ---- cell update: 4 units x 1 episode per thread -----------------------------------------------------
        const float4 ig = *reinterpret_cast<const float4 *>(act + (0 * HP_NB + ce) * 32 + 4 * cj);
        const float4 fg = *reinterpret_cast<const float4 *>(act + (1 * HP_NB + ce) * 32 + 4 * cj);
        const float4 gg = *reinterpret_cast<const float4 *>(act + (2 * HP_NB + ce) * 32 + 4 * cj);
        const float4 og = *reinterpret_cast<const float4 *>(act + (3 * HP_NB + ce) * 32 + 4 * cj);
        float4 hn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < my_len) {
          c[0] = fmaf(fg.x, c[0], ig.x * gg.x); c[1] = fmaf(fg.y, c[1], ig.y * gg.y);
          c[2] = fmaf(fg.z, c[2], ig.z * gg.z); c[3] = fmaf(fg.w, c[3], ig.w * gg.w);
          hn = make_float4(og.x * tanh_fast(c[0]), og.y * tanh_fast(c[1]), og.z * tanh_fast(c[2]), og.w * tanh_fast(c[3]));
        }
        HP_STAMP(13);
        if (s + 1 < nsteps) {  // h_s as fp16 pieces into K-slot i of CTA (rank + i) % 8, i = 0..7 (uniform branch)
          uint32_t p1a, p1b, p2a = 0u, p2b = 0u;
          if (bf16_mode) {
            p1a = bf16x2_bits(hn.x, hn.y);
            p1b = bf16x2_bits(hn.z, hn.w);
          } else {
            p1a = h2_bits(hn.x, hn.y);
            p1b = h2_bits(hn.z, hn.w);
            const float2 f0 = h2_floats(p1a), f1 = h2_floats(p1b);
            p2a = h2_bits(hn.x - f0.x, hn.y - f0.y);
            p2b = h2_bits(hn.z - f1.x, hn.w - f1.y);
          }
          // even lane keeps h1 and takes the partner's h1; odd lane keeps h2 and takes the partner's h2
          const uint32_t ga = piece ? p1a : p2a, gb = piece ? p1b : p2b;
          const uint32_t ra = __shfl_xor_sync(0xffffffffu, ga, 1), rb = __shfl_xor_sync(0xffffffffu, gb, 1);
          const uint4 gran = piece ? make_uint4(ra, rb, p2a, p2b) : make_uint4(p1a, p1b, ra, rb);
          if (sender) {
            const uint32_t boff = (uint32_t)((p ^ 1) * HP_B_BYTES), moff = (uint32_t)((p ^ 1) * 16);
#pragma unroll
            for (int r = 0; r < 4; ++r) st_async_v4u(raddr[r] + boff, gran, rbar[r] + moff);
          }
        }
        HP_STAMP(14);
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
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();  // nobody re-zeroes buffers (or exits) while a peer may still address its shared memory
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tmem_dealloc_pair<HP_TMEM_COLS>(tmem_base);
  }
}

template <typename K>
static int h3p_max_active_clusters(K kernel, int threads, int smem) {
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

// profiling hook: buf = device buffer of HP_PROF_STEPS * HP_PROF_SLOTS int64 (or NULL to switch the timeline off)
extern "C" int mts_debug_rec_profile_h3p(long long *buf) {
  MTS_CUDA(cudaMemcpyToSymbol(g_hp_prof, &buf, sizeof(buf)));
  return 0;
}

template <bool BF16>
static int launch_h3p(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order, int n_enc, int B, int T,
                     float *y, float *gates, float *y_corr, cudaStream_t st) {
  MTS_PER_DEVICE(int, cap);
  if (!cap) {
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_h3p_kernel<false, 1, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, hp_smem<1>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_h3p_kernel<true, 1, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, hp_smem<1>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_h3p_kernel<false, 2, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, hp_smem<2>()));
    MTS_CUDA(cudaFuncSetAttribute(lstm_fwd_h3p_kernel<true, 2, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, hp_smem<2>()));
    cap = h3p_max_active_clusters(lstm_fwd_h3p_kernel<false, 2, BF16>, 2 * HP_SUB_THREADS, hp_smem<2>());
  }
  // episodes per tile: 16 when the tiles outnumber the resident clusters; otherwise as few as spreading the batch over
  // all clusters allows (a cluster's step gets shorter with fewer episodes, the MMAs cost the same)
  int ept = HP_NB;
  {
    const int per_dir = cap / (2 * n_enc);   // clusters one (direction, encoder) can have
    if (per_dir >= 1 && (B + HP_NB - 1) / HP_NB <= per_dir) {
      const int want = (B + per_dir - 1) / per_dir;
      ept = want < 1 ? 1 : (want > HP_NB ? HP_NB : want);
    }
  }
  static const char *force_ept = getenv("MTS_REC_EPT");
  if (force_ept && atoi(force_ept) >= 1 && atoi(force_ept) <= HP_NB) ept = atoi(force_ept);
  const int n_tiles = (B + ept - 1) / ept;
  const int items1 = n_tiles * 2 * n_enc;
  // one tile per cluster while everything fits in a single wave; otherwise two tile pipelines per cluster
  static const char *force = getenv("MTS_REC_NT");
  const bool two = force ? (force[0] == '2') : (items1 > cap);
  if (!two) {
    const unsigned grid = (unsigned)((items1 < cap ? items1 : cap) * kCluster);
    if (gates) lstm_fwd_h3p_kernel<true, 1, BF16><<<grid, HP_SUB_THREADS, hp_smem<1>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr);
    else lstm_fwd_h3p_kernel<false, 1, BF16><<<grid, HP_SUB_THREADS, hp_smem<1>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr);
  } else {
    const int items2 = ((n_tiles + 1) / 2) * 2 * n_enc;
    const unsigned grid = (unsigned)((items2 < cap ? items2 : cap) * kCluster);
    if (gates) lstm_fwd_h3p_kernel<true, 2, BF16><<<grid, 2 * HP_SUB_THREADS, hp_smem<2>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr);
    else lstm_fwd_h3p_kernel<false, 2, BF16><<<grid, 2 * HP_SUB_THREADS, hp_smem<2>(), st>>>(gx, w_hh, lengths, order, B, T, n_enc, n_tiles, ept, y, gates, y_corr);
  }
  MTS_LAUNCH_CHECK();
  return 0;
}

// fp16-split tensor-core forward recurrence; same arguments as mts_lstm_rec_fwd_tc, H must be 256.  precision: 0 = three
// fp16 products (fp32-parity path), 1 = one bf16 product (the explicit bf16 path).
extern "C" int mts_lstm_rec_fwd_h3p(const float *gx, const float *w_hh, const int32_t *lengths, const int32_t *order, int n_enc,
                                   int B, int T, int H, float *y, float *gates, float *y_corr, int precision, void *stream) {
  MTS_REQUIRE(gx && w_hh && lengths && y, MTS_E_BADARG, "lstm_rec_fwd_h3p: null pointer");
  MTS_REQUIRE(n_enc >= 1 && B > 0 && T > 0, MTS_E_BADARG, "lstm_rec_fwd_h3p: bad shape");
  MTS_REQUIRE(H == kH, MTS_E_UNSUPPORTED, "lstm_rec_fwd_h3p: the tensor-core recurrence serves H == 256");
  MTS_REQUIRE(precision == 0 || precision == 1, MTS_E_BADARG, "lstm_rec_fwd_h3p: precision must be 0 (fp16 x3) or 1 (bf16)");
  MTS_REQUIRE(!y_corr || n_enc == 1, MTS_E_UNSUPPORTED, "lstm_rec_fwd_h3p: the fused correction operand needs n_enc == 1");
  MTS_REQUIRE((((uintptr_t)y | (uintptr_t)gx | (uintptr_t)w_hh) & 15) == 0, MTS_E_BADARG,
              "lstm_rec_fwd_h3p: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  return precision ? launch_h3p<true>(gx, w_hh, lengths, order, n_enc, B, T, y, gates, y_corr, st)
                   : launch_h3p<false>(gx, w_hh, lengths, order, n_enc, B, T, y, gates, y_corr, st);
}


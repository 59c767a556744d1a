// Banded (windowed) self-attention, forward, on the 5th-generation tensor cores.
// Same contract as band_attn_fwd_kernel (xfmr.cu): what the reference obtains from HF LongformerSelfAttention
// (models/RestrictedTransformerLayer.py:131 -> HF modeling_longformer.py:481-639, sliding chunks :758-867):
//   softmax over |i - j| <= w, j < len_b of (q_i / sqrt(hd)) . k_j in fp32, exact zeros for padded queries (:578).
//
// One work item = 128 queries of one head of one episode; a persistent CTA per SM walks the items.  Only the key
// tiles (64 keys) intersecting [q0 - w, q0 + 128 + w) x [0, len_b) are read.  Both contractions run as tcgen05.mma
// with the error-compensated operand scheme of the GEMM (common.cuh): x = hi + rest, hi = what kind::tf32 reads of
// the raw fp32 word,
//     S   = Q K^T  ~= hi(Q) hi(K)^T [kind::tf32]  +  bf16(Q) bf16(rest K)^T + bf16(rest Q) bf16(K)^T [kind::f16, packed]
//     O_t = P V    ~= hi(P) hi(V)   [kind::tf32]  +  bf16(P) bf16(rest V)   + bf16(rest P) bf16(V)   [kind::f16, packed]
// i.e. fp32-grade scores and outputs (relative error ~2^-19 per product).  The packed correction operands are
// derived ON CHIP from the raw fp32 rows (nothing but q, k, v is read from HBM, nothing but o written).
//
// Tensor memory (512 columns):  Q raw fp32 [0,HD)  (A operand of S, .ts form)
//                               S / P tiles [128,192) [192,256): the accumulator of Q K^T, overwritten IN PLACE by
//                                  the raw fp32 probabilities = the A operand of P V
//                               packed bf16 correction operand of P [256,320) [320,384)
//                               O_t accumulator [384, 384+HD)
// Shared memory: packed correction operand of Q (K-major SWIZZLE_128B planes of 32 head columns), the K tile (raw +
//   packed correction, K-major: N = keys, K = head dim) and the V tile (raw + packed correction, MN-major: N = head
//   dim, K = keys).  A K-major [rows][32 fp32] plane and an MN-major [K rows][32 fp32 of N] plane are the SAME bytes,
//   so K and V tiles are written by the same code.
//
// Warp roles (13 warps):
//   warps 0-3   softmax: thread = query row.  Loads Q (coalesced, transposed through shared memory) into tensor memory
//               and its correction operand into shared memory; per key tile: tcgen05.ld S, band / length mask, running
//               max (fp32), p = 2^((s - m) log2 e), row sums, tcgen05.st P (raw + packed correction), rescale factor
//               to shared memory.
//   warps 4-7   correction + epilogue: thread = query row.  O accumulates in tensor memory over the key tiles; the
//               running maximum is only moved when a tile exceeds it by more than RESCALE_TH (p stays <= e^TH, harmless
//               in fp32), so the rescale O *= alpha (tcgen05.ld / st) is rare.  At the end o / l, transposed through
//               shared memory, row-contiguous stores of o (+ the packed correction operand the next GEMM wants) and
//               the log-sum-exp.
//   warps 8-9   K producer, warps 10-11 V producer: global -> registers (prefetched one tile ahead of the buffer
//               hand-over) -> raw plane + correction plane, fence.proxy.async, mbarrier.
//   warp 12     TMEM allocation + the single MMA-issuing thread.
// All hand-overs are mbarriers; the tensor pipe, the MUFU pipe, the load path and the epilogue of consecutive tiles
// (and consecutive items) overlap.
#include <stdlib.h>

#include "tcgen05_utils.cuh"

namespace mts {
namespace atc {

constexpr int BQ = 128, KT = 64;
constexpr int THREADS = 13 * 32;
constexpr uint32_t COL_Q = 0, COL_S = 128, COL_PC = 256, COL_O = 384;
constexpr int TRS = 20;  // floats per row of a per-warp 32 x 16 transposition buffer (80 B: 16-byte aligned rows)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_TH = 5.5f;  // the reference maximum of a row moves only when a tile's maximum exceeds it by this much

template <int HD>
struct Cfg {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 128, "head dim: multiple of 16, at most 128");
  static constexpr int NP = (HD + 31) / 32;   // 32-column planes of a raw fp32 / packed-correction operand
  static constexpr int NPV = (HD + 63) / 64;  // 64-column planes of the MN-major bf16 correction operand of V
  static constexpr int C8 = HD / 8;           // 8-float units per row
  static constexpr int QC_BYTES = NP * BQ * 128;
  static constexpr int KH_BYTES = NP * KT * 128;
  static constexpr int VC_BYTES = NPV * 2 * KT * 128;
  static constexpr int OFF_QC = 0;
  static constexpr int OFF_KH = OFF_QC + QC_BYTES;
  static constexpr int OFF_KC = OFF_KH + KH_BYTES;
  static constexpr int OFF_VH = OFF_KC + KH_BYTES;
  static constexpr int OFF_VC = OFF_VH + KH_BYTES;
  static constexpr int OFF_TR = OFF_VC + VC_BYTES;
  static constexpr int TR_BYTES = 8 * 32 * TRS * 4;
  static constexpr int OFF_ALPHA = OFF_TR + TR_BYTES;        // [2][128]
  static constexpr int OFF_FIN = OFF_ALPHA + 2 * BQ * 4;     // [2][128] x {1 / l, lse}
  static constexpr int OFF_BAR = OFF_FIN + 2 * BQ * 2 * 4;
  static constexpr int USED = OFF_BAR + 16 * 8 + 16;
  // every CTA allocates all 512 TMEM columns, so two CTAs must never share an SM: ask for more than half its shared memory
  static constexpr int SMEM = (USED + 1024) > 120 * 1024 ? (USED + 1024) : 120 * 1024;
};

enum { B_QREADY = 0, B_QFREE, B_KFULL, B_KEMPTY, B_VFULL, B_VEMPTY, B_SFULL0, B_SFULL1, B_PREADY0, B_PREADY1, B_OFULL, B_OREADY,
       B_AREADY0, B_AREADY1, B_FREADY0, B_FREADY1, B_COUNT };
static_assert(B_COUNT <= 16, "barrier block holds 16 mbarriers");

// MN-major SWIZZLE_128B descriptor: LBO = byte stride between 128-byte chunks along N, SBO = between groups of 8 K rows
__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t B_MN_MAJOR = 1u << 16;  // instruction descriptor: B operand is MN-major

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 f4_bits(float4 v) {
  return make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
}
__device__ __forceinline__ uint4 bf16x8(float4 a, float4 b) {
  return make_uint4(bf16x2_bits(a.x, a.y), bf16x2_bits(a.z, a.w), bf16x2_bits(b.x, b.y), bf16x2_bits(b.z, b.w));
}
__device__ __forceinline__ float4 rest4(float4 v) {
  return make_float4(tf32_rest_exact(v.x), tf32_rest_exact(v.y), tf32_rest_exact(v.z), tf32_rest_exact(v.w));
}

// one work item, as every role derives it (same arithmetic everywhere keeps the roles' barrier phases in step)
struct Item {
  int b, head, q0, len, Sq, kbeg, kend, nt;
  int64_t row0;
  bool active;
};
__device__ __forceinline__ Item make_item(int item, int nheads, int n_qb, int S, int w, const int32_t *__restrict__ lengths,
                                          const int32_t *__restrict__ offsets) {
  Item it;
  it.head = item % nheads;
  const int rest = item / nheads;
  it.q0 = (rest % n_qb) * BQ;
  it.b = rest / n_qb;
  it.len = min(max(__ldg(lengths + it.b), 0), S);
  // ragged layout (offsets != NULL): episode b owns rows offsets[b] .. offsets[b] + len - 1 and nothing beyond
  it.row0 = offsets ? (int64_t)__ldg(offsets + it.b) : (int64_t)it.b * S;
  it.Sq = offsets ? it.len : S;
  it.active = it.q0 < it.len;
  it.kbeg = max(0, it.q0 - w);
  it.kend = min(it.len, it.q0 + BQ + w);
  it.nt = it.active ? (it.kend - it.kbeg + KT - 1) / KT : 0;
  return it;
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
    band_attn_fwd_tc_kernel(const float *__restrict__ qkv, int64_t ld, const int32_t *__restrict__ lengths,
                            const int32_t *__restrict__ offsets, int B, int S, int nheads, int w, float *__restrict__ out,
                            float *__restrict__ out_hi, float *__restrict__ out_lo, int Kp, float *__restrict__ lse) {
  using C = Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = tc::s_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
  float *alpha_s = reinterpret_cast<float *>(smem + C::OFF_ALPHA);
  float *fin_s = reinterpret_cast<float *>(smem + C::OFF_FIN);

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int d = nheads * HD;
  const int n_qb = (S + BQ - 1) / BQ;
  const int n_items = B * n_qb * nheads;
#define BAR(i) (sbase + (uint32_t)(C::OFF_BAR + 8 * (i)))

  if (threadIdx.x == 0) {
    tc::bar_init(BAR(B_QREADY), 128);
    tc::bar_init(BAR(B_QFREE), 1);
    tc::bar_init(BAR(B_KFULL), 64);
    tc::bar_init(BAR(B_KEMPTY), 1);
    tc::bar_init(BAR(B_VFULL), 64);
    tc::bar_init(BAR(B_VEMPTY), 1);
    tc::bar_init(BAR(B_SFULL0), 1);
    tc::bar_init(BAR(B_SFULL1), 1);
    tc::bar_init(BAR(B_PREADY0), 128);
    tc::bar_init(BAR(B_PREADY1), 128);
    tc::bar_init(BAR(B_OFULL), 1);
    tc::bar_init(BAR(B_OREADY), 128);
    tc::bar_init(BAR(B_AREADY0), 128);
    tc::bar_init(BAR(B_AREADY1), 128);
    tc::bar_init(BAR(B_FREADY0), 128);
    tc::bar_init(BAR(B_FREADY1), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tc::tmem_alloc<512>(tc::s_u32(tmem_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =============================== softmax warps: thread = query row ===============================
    const int r = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    float *tr = reinterpret_cast<float *>(smem + C::OFF_TR) + warp * 32 * TRS;
    const float scale = sqrtf((float)HD);
    uint32_t item_g = 0, tile_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      if (!it.active) continue;
      // ---- Q: coalesced loads, transpose through the warp's buffer, raw rows -> tensor memory, correction -> shared ----
      tc::bar_wait_wd(BAR(B_QFREE), (item_g & 1) ^ 1);  // every S = Q K^T of the previous item has completed
      tc::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < HD / 16; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = (lane >> 2) + 8 * i, gi = it.q0 + warp * 32 + rr;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gi < it.Sq) {
            v = __ldg(reinterpret_cast<const float4 *>(qkv + (it.row0 + gi) * ld + it.head * HD + 16 * c) + (lane & 3));
            v.x /= scale; v.y /= scale; v.z /= scale; v.w /= scale;  // query_vectors /= sqrt(head_dim) (HF :513)
          }
          *reinterpret_cast<float4 *>(tr + rr * TRS + 4 * (lane & 3)) = v;
        }
        __syncwarp();
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4 *>(tr + lane * TRS + 4 * j);
        __syncwarp();
        uint32_t raw[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          raw[4 * j] = __float_as_uint(x[j].x); raw[4 * j + 1] = __float_as_uint(x[j].y);
          raw[4 * j + 2] = __float_as_uint(x[j].z); raw[4 * j + 3] = __float_as_uint(x[j].w);
        }
        tmem_st16(trow + COL_Q + 16 * c, raw);
        // packed correction operand, A side: per 16 columns [bf16(x) x16 | bf16(rest) x16] = 4 chunks of 16 bytes
        const uint32_t rowa = sbase + C::OFF_QC + (uint32_t)((c >> 1) * (BQ * 128) + r * 128);
        const int cb = (c & 1) * 4, sw = r & 7;
        sts128(rowa + (((cb + 0) ^ sw) << 4), bf16x8(x[0], x[1]));
        sts128(rowa + (((cb + 1) ^ sw) << 4), bf16x8(x[2], x[3]));
        sts128(rowa + (((cb + 2) ^ sw) << 4), bf16x8(rest4(x[0]), rest4(x[1])));
        sts128(rowa + (((cb + 3) ^ sw) << 4), bf16x8(rest4(x[2]), rest4(x[3])));
      }
      tc::tmem_wait_st();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::bar_arrive(BAR(B_QREADY));

      // ---- key tiles ----------------------------------------------------------------------------------------------
      const int i = it.q0 + r;
      float m_run = -INFINITY, l_run = 0.0f;
#pragma unroll 1
      for (int t = 0; t < it.nt; ++t, ++tile_g) {
        const int buf = tile_g & 1;
        const int k0 = it.kbeg + t * KT;
        // valid tile-local key range of my row: band |i - j| <= w, j < kend (<= len), nothing for padded queries
        const int clo = max(0, i - w - k0);
        const int chi = (i < it.len) ? min(it.kend - k0, i + w + 1 - k0) : 0;
        tc::bar_wait_wd(BAR(B_SFULL0 + buf), (tile_g >> 1) & 1);
        tc::tc_fence_after();
        float s[KT];
        tmem_ld32_nowait(trow + COL_S + 64 * buf, s);
        tmem_ld32_nowait(trow + COL_S + 64 * buf + 32, s + 32);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < KT; ++c) {
          s[c] = (c >= clo && c < chi) ? s[c] : -INFINITY;
          mx = fmaxf(mx, s[c]);
        }
        // reference maximum: moved only when this tile exceeds it by more than RESCALE_TH (first valid tile: always)
        float alpha = 1.0f;
        if (mx > m_run + RESCALE_TH) {   // false while mx == -inf; true for the first finite mx (m_run == -inf)
          if (m_run != -INFINITY) { alpha = ex2f((m_run - mx) * LOG2E); l_run *= alpha; }
          m_run = mx;
        }
        alpha_s[buf * BQ + r] = alpha;
        tc::bar_arrive(BAR(B_AREADY0 + buf));   // the correction warps may rescale O while the exponentials run
        const float mb = (m_run == -INFINITY) ? 0.0f : -m_run * LOG2E;
        float psum = 0.0f;
#pragma unroll
        for (int blk = 0; blk < KT / 16; ++blk) {
          uint32_t ph[16], pc[16];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2f(fmaf(s[16 * blk + e], LOG2E, mb)), p1 = ex2f(fmaf(s[16 * blk + e + 1], LOG2E, mb));
            psum += p0 + p1;
            ph[e] = __float_as_uint(p0);
            ph[e + 1] = __float_as_uint(p1);
            pc[e >> 1] = bf16x2_bits(p0, p1);
            pc[8 + (e >> 1)] = bf16x2_bits(tf32_rest_exact(p0), tf32_rest_exact(p1));
          }
          tmem_st16(trow + COL_S + 64 * buf + 16 * blk, ph);    // raw fp32 probabilities, in place of the scores
          tmem_st16(trow + COL_PC + 64 * buf + 16 * blk, pc);   // [bf16(p) x16 | bf16(rest p) x16]
        }
        l_run += psum;
        if (t == it.nt - 1) {
          const bool live = (i < it.len) && (l_run > 0.0f);
          fin_s[((item_g & 1) * BQ + r) * 2] = live ? 1.0f / l_run : 0.0f;
          fin_s[((item_g & 1) * BQ + r) * 2 + 1] = live ? m_run + logf(l_run) : 0.0f;
          tc::bar_arrive(BAR(B_FREADY0 + (item_g & 1)));
        }
        tc::tmem_wait_st();
        tc::tc_fence_before();
        tc::bar_arrive(BAR(B_PREADY0 + buf));
      }
      ++item_g;
    }
  } else if (warp < 8) {
    // =============================== correction + epilogue warps: thread = query row ===============================
    // this warpgroup needs few registers and hands 40 per thread to the producers (same count: the CTA's pool is fixed)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const int q = warp - 4, r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float *tr = reinterpret_cast<float *>(smem + C::OFF_TR) + warp * 32 * TRS;
    uint32_t item_g = 0, tile_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      float inv = 0.0f, lse_v = 0.0f;
      if (it.active) {
#pragma unroll 1
        for (int t = 0; t < it.nt; ++t, ++tile_g) {
          // tile t's rescale factor is known as soon as its row maxima are (tile 0: 1 by construction, but its barrier
          // phase must still be consumed)
          tc::bar_wait_wd(BAR(B_AREADY0 + (tile_g & 1)), (tile_g >> 1) & 1);
          if (t > 0) {
            const float alpha = alpha_s[(tile_g & 1) * BQ + r];
            tc::bar_wait_wd(BAR(B_OFULL), (tile_g - 1) & 1);   // O holds tiles 0..t-1 completely
            tc::tc_fence_after();
            if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
              for (int c0 = 0; c0 < HD; c0 += 16) {
                float v[16];
                uint32_t u[16];
                tmem_ld16_nowait(trow + COL_O + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) u[e] = __float_as_uint(v[e] * alpha);
                tmem_st16(trow + COL_O + c0, u);
              }
              tc::tmem_wait_st();
            }
            tc::tc_fence_before();
          }
          // t == 0: this tile's P V overwrites O; my own reads of the previous item's O are behind me
          tc::bar_arrive(BAR(B_OREADY));
        }
        tc::bar_wait_wd(BAR(B_FREADY0 + (item_g & 1)), (item_g >> 1) & 1);
        inv = fin_s[((item_g & 1) * BQ + r) * 2];
        lse_v = fin_s[((item_g & 1) * BQ + r) * 2 + 1];
        tc::bar_wait_wd(BAR(B_OFULL), (tile_g - 1) & 1);   // the last P V of the item has completed
        tc::tc_fence_after();
        ++item_g;
      }
      // ---- o / l, transposed through the warp's buffer, row-contiguous stores (+ the next GEMM's correction operand) ----
      // (inactive item = a block of padded queries: exact zeros, HF :578; nothing exists there in the ragged layout)
      if (it.q0 < it.Sq) {
#pragma unroll 1
        for (int c = 0; c < HD / 16; ++c) {
          float v[16];
          if (it.active) {
            tmem_ld16_nowait(trow + COL_O + 16 * c, v);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4 *>(tr + lane * TRS + 4 * j) =
                make_float4(v[4 * j] * inv, v[4 * j + 1] * inv, v[4 * j + 2] * inv, v[4 * j + 3] * inv);
          __syncwarp();
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int rr = (lane >> 1) + 16 * h2, half = lane & 1, gi = it.q0 + q * 32 + rr;
            const float4 a = *reinterpret_cast<const float4 *>(tr + rr * TRS + 8 * half);
            const float4 b2 = *reinterpret_cast<const float4 *>(tr + rr * TRS + 8 * half + 4);
            if (gi < it.Sq) {
              const int col = it.head * HD + 16 * c + 8 * half;
              if (out) {
                float4 *dst = reinterpret_cast<float4 *>(out + (it.row0 + gi) * d + col);
                dst[0] = a;
                dst[1] = b2;
              }
              if (out_hi) {
                float4 *dst = reinterpret_cast<float4 *>(out_hi + (it.row0 + gi) * Kp + col);
                dst[0] = a;
                dst[1] = b2;
                corr_store8(out_lo + (it.row0 + gi) * Kp, col, a, b2, 0);
              }
            }
          }
          __syncwarp();
        }
        tc::tc_fence_before();
      }
      if (lse && it.q0 + r < S) lse[((int64_t)it.b * nheads + it.head) * S + it.q0 + r] = lse_v;
    }
  } else if (warp < 12) {
    // =============================== K (warps 8-9) and V (warps 10-11) producers ===============================
    // a producer thread keeps a whole tile's share (HD / 8 units of 8 floats) in registers between the prefetch and the
    // hand-over of the buffer: take the registers the correction warpgroup released
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int which = (warp < 10) ? 1 : 2;                 // column block of the qkv row: 1 = k, 2 = v
    const int tid_p = ((warp & 1) << 5) + lane;             // 0..63 inside my producer pair
    const uint32_t hi_base = sbase + (uint32_t)(which == 1 ? C::OFF_KH : C::OFF_VH);
    const uint32_t co_base = sbase + (uint32_t)(which == 1 ? C::OFF_KC : C::OFF_VC);
    const uint32_t full = BAR(which == 1 ? B_KFULL : B_VFULL), empty = BAR(which == 1 ? B_KEMPTY : B_VEMPTY);
    uint32_t tile_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      if (!it.active) continue;
      const float *gbase = qkv + it.row0 * ld + which * d + it.head * HD;
#pragma unroll 1
      for (int t = 0; t < it.nt; ++t, ++tile_g) {
        const int k0 = it.kbeg + t * KT;
        float4 va[C::C8], vb[C::C8];
#pragma unroll
        for (int j = 0; j < C::C8; ++j) {
          const int u = tid_p + 64 * j, row = u / C::C8, c8 = u % C::C8;
          va[j] = vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);   // keys beyond kend: zeros (P is 0 there; 0 * garbage could be NaN)
          if (k0 + row < it.kend) {
            const float4 *src = reinterpret_cast<const float4 *>(gbase + (int64_t)(k0 + row) * ld + 8 * c8);
            va[j] = __ldg(src);
            vb[j] = __ldg(src + 1);
          }
        }
        tc::bar_wait_wd(empty, (tile_g & 1) ^ 1);   // the MMAs that read the previous tile have completed
#pragma unroll
        for (int j = 0; j < C::C8; ++j) {
          const int u = tid_p + 64 * j, row = u / C::C8, c8 = u % C::C8;
          // raw fp32: plane of 32 columns, row = key, 128-byte rows, 16-byte chunks XOR-swizzled by the row
          const uint32_t hrow = hi_base + (uint32_t)((c8 >> 2) * (KT * 128) + row * 128);
          const int ci = (c8 & 3) * 2, sw = row & 7;
          sts128(hrow + ((ci ^ sw) << 4), f4_bits(va[j]));
          sts128(hrow + (((ci + 1) ^ sw) << 4), f4_bits(vb[j]));
          const uint4 xb = bf16x8(va[j], vb[j]), rb = bf16x8(rest4(va[j]), rest4(vb[j]));
          if (which == 1) {
            // K-major packed correction, B side: per 16 columns [bf16(rest) x16 | bf16(x) x16]
            const uint32_t crow = co_base + (uint32_t)((c8 >> 2) * (KT * 128) + row * 128);
            const int cb = ((c8 >> 1) & 1) * 4 + (c8 & 1);
            sts128(crow + ((cb ^ sw) << 4), rb);
            sts128(crow + (((cb + 2) ^ sw) << 4), xb);
          } else {
            // MN-major packed correction, B side: K' rows per 16 keys [rest v x16 | v x16], 64 head columns per 128-byte row
            const int rr = (row >> 4) * 32 + (row & 15);
            const uint32_t crow = co_base + (uint32_t)((c8 >> 3) * (2 * KT * 128) + rr * 128);
            const int ch = c8 & 7, sv = rr & 7;
            sts128(crow + ((ch ^ sv) << 4), rb);
            sts128(crow + 16 * 128 + ((ch ^ sv) << 4), xb);
          }
        }
        tc::fence_proxy_async();
        tc::bar_arrive(full);
      }
    }
  } else {
    // =============================== MMA issuer: warp-uniform control flow, one elected lane ===============================
    constexpr uint32_t id_s = tc::idesc_tf32(BQ, KT), id_sc = tc::idesc_bf16(BQ, KT);
    constexpr uint32_t id_o = tc::idesc_tf32(BQ, HD) | B_MN_MAJOR, id_oc = tc::idesc_bf16(BQ, HD) | B_MN_MAJOR;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = tc::elect_one();
    const uint32_t qc_a = sbase + C::OFF_QC, kh_a = sbase + C::OFF_KH, kc_a = sbase + C::OFF_KC;
    const uint32_t vh_a = sbase + C::OFF_VH, vc_a = sbase + C::OFF_VC;
    uint32_t item_g = 0, tile_g = 0;

    auto issue_qk = [&](uint32_t tg, bool last) {
      const uint32_t buf = tg & 1;
      tc::bar_wait_wd(BAR(B_KFULL), tg & 1);
      tc::tc_fence_after();
      const uint32_t d_s = tb + COL_S + 64 * buf;
#pragma unroll
      for (int ks = 0; ks < HD / 8; ++ks) {
        const uint64_t bd = tc::desc_sw128(kh_a + (uint32_t)((ks >> 2) * (KT * 128) + (ks & 3) * 32));
        if (leader) tc::umma_tf32_ts(d_s, tb + COL_Q + 8 * ks, bd, id_s, ks != 0);
      }
#pragma unroll
      for (int j = 0; j < HD / 8; ++j) {
        const uint64_t ad = tc::desc_sw128(qc_a + (uint32_t)((j >> 2) * (BQ * 128) + (j & 3) * 32));
        const uint64_t bd = tc::desc_sw128(kc_a + (uint32_t)((j >> 2) * (KT * 128) + (j & 3) * 32));
        if (leader) tc::umma_bf16_ss(d_s, ad, bd, id_sc, 1);
      }
      if (leader) {
        tc::umma_commit(BAR(B_SFULL0 + buf));
        tc::umma_commit(BAR(B_KEMPTY));
        if (last) tc::umma_commit(BAR(B_QFREE));
      }
      __syncwarp();
    };

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      if (!it.active) continue;
      tc::bar_wait_wd(BAR(B_QREADY), item_g & 1);
      tc::tc_fence_after();
      issue_qk(tile_g, it.nt == 1);
#pragma unroll 1
      for (int t = 0; t < it.nt; ++t, ++tile_g) {
        if (t + 1 < it.nt) issue_qk(tile_g + 1, t + 2 == it.nt);
        const uint32_t buf = tile_g & 1;
        tc::bar_wait_wd(BAR(B_PREADY0 + buf), (tile_g >> 1) & 1);
        tc::bar_wait_wd(BAR(B_VFULL), tile_g & 1);
        tc::bar_wait_wd(BAR(B_OREADY), tile_g & 1);   // O rescaled for this tile (t > 0) / free to be overwritten (t == 0)
        tc::tc_fence_after();
        const uint32_t d_o = tb + COL_O;
#pragma unroll
        for (int kg = 0; kg < KT / 8; ++kg) {
          const uint64_t bd = desc_sw128_mn(vh_a + (uint32_t)(kg * 1024), KT * 128);
          if (leader) tc::umma_tf32_ts(d_o, tb + COL_S + 64 * buf + 8 * kg, bd, id_o, (t | kg) != 0);
        }
#pragma unroll
        for (int j = 0; j < 2 * KT / 16; ++j) {
          const uint64_t bd = desc_sw128_mn(vc_a + (uint32_t)(j * 2048), 2 * KT * 128);
          if (leader) tc::umma_bf16_ts(d_o, tb + COL_PC + 64 * buf + 8 * j, bd, id_oc, 1);
        }
        if (leader) {
          tc::umma_commit(BAR(B_OFULL));
          tc::umma_commit(BAR(B_VEMPTY));
        }
        __syncwarp();
      }
      ++item_g;
    }
  }
#undef BAR
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int HD>
static int launch(const float *qkv, int64_t ld, const int32_t *lengths, const int32_t *offsets, int B, int S, int nheads, int w,
                  float *out, float *out_hi, float *out_lo, int Kp, float *lse, cudaStream_t st) {
  using C = Cfg<HD>;
  MTS_CUDA(cudaFuncSetAttribute(band_attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  const int n_items = B * ((S + BQ - 1) / BQ) * nheads;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  band_attn_fwd_tc_kernel<HD><<<grid, THREADS, C::SMEM, st>>>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse);
  MTS_LAUNCH_CHECK();
  return 0;
}

}  // namespace atc
}  // namespace mts

using namespace mts;

extern "C" int mts_band_attn_tc_supported(int hd) { return hd == 16 || hd == 32 || hd == 64 || hd == 112 || hd == 128; }

extern "C" int mts_band_attn_fwd_tc(const float *qkv, int64_t ld, const int32_t *lengths, const int32_t *offsets, int B, int S,
                                    int nheads, int hd, int w, float *out, float *out_hi, float *out_lo, int Kp, float *lse,
                                    void *stream) {
  MTS_REQUIRE(qkv && lengths && (out || out_hi), MTS_E_BADARG, "band_attn_fwd_tc: null pointer");
  MTS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), MTS_E_BADARG, "band_attn_fwd_tc: hi and lo go together");
  MTS_REQUIRE(B > 0 && S > 0 && nheads > 0 && hd > 0 && w >= 0, MTS_E_BADARG, "band_attn_fwd_tc: bad shape");
  MTS_REQUIRE(ld % 4 == 0 && ld >= 3 * nheads * hd, MTS_E_BADARG, "band_attn_fwd_tc: qkv row stride");
  MTS_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)out_hi & 15) == 0 && ((uintptr_t)out_lo & 15) == 0,
              MTS_E_BADARG, "band_attn_fwd_tc: buffers must be 16-byte aligned");
  MTS_REQUIRE(!out_hi || Kp == nheads * hd, MTS_E_UNSUPPORTED, "band_attn_fwd_tc: the split output needs Kp == model width");
  cudaStream_t st = (cudaStream_t)stream;
  switch (hd) {
    case 16: return atc::launch<16>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse, st);
    case 32: return atc::launch<32>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse, st);
    case 64: return atc::launch<64>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse, st);
    case 112: return atc::launch<112>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse, st);
    case 128: return atc::launch<128>(qkv, ld, lengths, offsets, B, S, nheads, w, out, out_hi, out_lo, Kp, lse, st);
    default: break;
  }
  set_error("band_attn_fwd_tc: head dim must be one of 16, 32, 64, 112, 128");
  return MTS_E_UNSUPPORTED;
}

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
//   dim, K = keys).  A K-major [rows][32 fp32] plane and an MN-major [K rows][32 fp32 of N] plane hold the same rows;
//   only the swizzle differs (tf32 MN-major operands exist in the SWIZZLE_128B_BASE32B layout only).
//
// Warp roles (13 working warps):
//   warps 0-3   softmax: thread = query row.  Per key tile: tcgen05.ld S, band / length mask, running max (fp32),
//               p = 2^((s - m) log2 e), row sums, tcgen05.st P (raw + packed correction), rescale factor to shared memory.
//   warps 4-7   correction, Q loader and epilogue: thread = query row.  Loads the NEXT item's Q during the current item's
//               last tile (coalesced, transposed through shared memory): raw rows into tensor memory, the packed
//               correction operand into shared memory.  O accumulates in tensor memory over the key tiles; the
//               running maximum is only moved when a tile exceeds it by more than RESCALE_TH (p stays <= e^TH, harmless
//               in fp32), so the rescale O *= alpha (tcgen05.ld / st) is rare.  At the end o / l, transposed through
//               shared memory, row-contiguous stores of o (+ the packed correction operand the next GEMM wants) and
//               the log-sum-exp.
//   warps 8-9   K producer, warps 10-11 V producer: global -> registers (prefetched one tile ahead of the buffer
//               hand-over) -> raw plane + correction plane, fence.proxy.async, mbarrier.
//   warp 12     TMEM allocation + the single MMA-issuing thread (warps 13-15 exist only because registers are allocated
//               per group of 4 warps: they release theirs to the producers and wait at the final barrier).
// All hand-overs are mbarriers; the tensor pipe, the MUFU pipe, the load path and the epilogue of consecutive tiles
// (and consecutive items) overlap.
#include <stdlib.h>

#include "tcgen05_utils.cuh"

namespace mts {
namespace atc {

constexpr int BQ = 128, KT = 64;
constexpr int THREADS = 16 * 32;   // 13 working warps; registers are allocated in units of 4 warps anyway
constexpr uint32_t COL_Q = 0, COL_S = 128, COL_PC = 256, COL_O = 384;
constexpr int TRS = 36;  // floats per row of a per-warp 32 x 32 transposition buffer (144 B: 16-byte aligned rows, conflict-free)
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
  static constexpr int TR_BYTES = 4 * 32 * TRS * 4;   // the four correction / Q-loader / epilogue warps
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

// MN-major SWIZZLE_128B descriptor (16-bit operands): LBO = byte stride between 128-byte chunks along N, SBO = between
// groups of 8 K rows (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major 32-bit (tf32) operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1, Swizzle<2,5,2> on the byte
// address: the 32-byte unit index, bits 5-6, is XORed with the K row index, bits 7-8): atoms of 4 K rows x 128 bytes
// (32 fp32 along N), SBO = byte stride between groups of 4 K rows, LBO = between 128-byte chunks along N.
__device__ __forceinline__ uint64_t desc_sw128b32_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
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

// Optional in-kernel timeline (off unless mts_debug_attn_profile() installs a buffer): CTA 0 stamps clock64() at the
// hand-over points of its first AP_ITEMS active items -- role 0 softmax, 1 correction, 2 K producer, 3 V producer, 4 MMA.
constexpr int AP_ITEMS = 6, AP_ROLES = 5, AP_SLOTS = 40;
__device__ long long *g_atc_prof = nullptr;
#define AP_STAMP(role, slot)                                                                                     \
  do {                                                                                                           \
    if (prof && item_g < (uint32_t)AP_ITEMS && (slot) < AP_SLOTS)                                                \
      prof[((int)item_g * AP_ROLES + (role)) * AP_SLOTS + (slot)] = clock64();                                    \
  } while (0)

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


// descriptor = (constant upper word, lower word = (address >> 4) | LBO << 16): stepping through an operand only adds to the
// lower word
__device__ __forceinline__ uint64_t desc_add(uint64_t base, uint32_t byte_off) { return base + (uint64_t)(byte_off >> 4); }

// S = Q K^T of one key tile: HD / 8 kind::tf32 MMAs (A = raw Q rows in tensor memory, B = raw K tile) + HD / 8 kind::f16
// MMAs over the packed correction operands.  Called by the whole MMA warp; one elected lane issues.  ONE copy in the
// instruction stream (the role code of five roles shares the instruction cache).
template <int HD>
__device__ __noinline__ void issue_qk(bool leader, uint32_t d_s, uint32_t q_tmem, uint64_t kh_desc, uint64_t qc_desc, uint64_t kc_desc) {
  constexpr uint32_t id_s = tc::idesc_tf32(BQ, KT), id_sc = tc::idesc_bf16(BQ, KT);
#pragma unroll
  for (int ks = 0; ks < HD / 8; ++ks) {
    const uint64_t bd = desc_add(kh_desc, (uint32_t)((ks >> 2) * (KT * 128) + (ks & 3) * 32));
    if (leader) tc::umma_tf32_ts(d_s, q_tmem + 8 * ks, bd, id_s, ks != 0);
  }
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    const uint64_t ad = desc_add(qc_desc, (uint32_t)((j >> 2) * (BQ * 128) + (j & 3) * 32));
    const uint64_t bd = desc_add(kc_desc, (uint32_t)((j >> 2) * (KT * 128) + (j & 3) * 32));
    if (leader) tc::umma_bf16_ss(d_s, ad, bd, id_sc, 1);
  }
}

// 32-column group g of a row: its width (the last group of HD = 112 is 16 wide)
template <int HD>
__device__ __forceinline__ constexpr int group_width(int g) { return (HD - 32 * g) >= 32 ? 32 : (HD - 32 * g); }

// Registers: launched with 128 per thread (512 threads); the four warpgroups then re-split the CTA's pool with setmaxnreg
// to 144 (softmax) + 112 (correction) + 176 (producers: a whole tile share in flight) + 80 (MMA issuer and its 3 idle warps).
template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
    band_attn_fwd_tc_kernel(const float *__restrict__ qkv, int64_t ld, const int32_t *__restrict__ lengths,
                            const int32_t *__restrict__ offsets, int B, int S, int nheads, int w, float *__restrict__ out,
                            float *__restrict__ out_hi, float *__restrict__ out_lo, int Kp, float *__restrict__ lse) {
  using C = Cfg<HD>;
  constexpr int NG = (HD + 31) / 32;   // 32-column groups of a row
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
  // first active work item of this CTA at or after index `from` (stride gridDim.x); n_items if there is none
  auto next_active = [&](int from, Item &it) {
    for (; from < n_items; from += gridDim.x) {
      it = make_item(from, nheads, n_qb, S, w, lengths, offsets);
      if (it.active) return from;
    }
    return n_items;
  };

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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int r = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    long long *prof = (blockIdx.x == 0 && threadIdx.x == 0) ? g_atc_prof : nullptr;
    uint32_t item_g = 0, tile_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      if (!it.active) continue;
      AP_STAMP(0, 0);
      const int i = it.q0 + r;
      float m_run = -INFINITY, l_run = 0.0f;
#pragma unroll 1
      for (int t = 0; t < it.nt; ++t, ++tile_g) {
        const int buf = tile_g & 1;
        const int k0 = it.kbeg + t * KT;
        // valid tile-local key range [clo, clo + nvalid) of my row: band |i - j| <= w, j < kend (<= len); empty for padded queries
        const int clo = max(0, i - w - k0);
        const int chi = (i < it.len) ? min(it.kend - k0, i + w + 1 - k0) : 0;
        const unsigned nvalid = (unsigned)max(chi - clo, 0);
        tc::bar_wait_wd(BAR(B_SFULL0 + buf), (tile_g >> 1) & 1);
        tc::tc_fence_after();
        AP_STAMP(0, 3 + 4 * t);
        float s[KT];
        tmem_ld32_nowait(trow + COL_S + 64 * buf, s);
        tmem_ld32_nowait(trow + COL_S + 64 * buf + 32, s + 32);
        tmem_wait_ld();
        AP_STAMP(0, 4 + 4 * t);
        // 16-key blocks that no row of this warp can see (outside the band of its 32 rows, or beyond the episode) are not
        // evaluated at all: their probabilities are stored as zeros.  Warp-uniform, so no divergence.
        unsigned live = 0;
#pragma unroll
        for (int blk = 0; blk < KT / 16; ++blk)
          if (__any_sync(0xffffffffu, clo < 16 * blk + 16 && chi > 16 * blk)) live |= 1u << blk;
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains (one warp per scheduler: ILP matters)
#pragma unroll
        for (int blk = 0; blk < KT / 16; ++blk) {
          if (!(live & (1u << blk))) continue;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int c = 16 * blk + e;
            s[c] = ((unsigned)(c - clo) < nvalid) ? s[c] : -INFINITY;
            mx4[e & 3] = fmaxf(mx4[e & 3], s[c]);
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // reference maximum: moved only when this tile exceeds it by more than RESCALE_TH (first valid tile: always)
        float alpha = 1.0f;
        if (mx > m_run + RESCALE_TH) {   // false while mx == -inf; true for the first finite mx (m_run == -inf)
          if (m_run != -INFINITY) { alpha = ex2f((m_run - mx) * LOG2E); l_run *= alpha; }
          m_run = mx;
        }
        alpha_s[buf * BQ + r] = alpha;
        tc::bar_arrive(BAR(B_AREADY0 + buf));   // the correction warps may rescale O while the exponentials run
        AP_STAMP(0, 5 + 4 * t);
        const float mb = (m_run == -INFINITY) ? 0.0f : -m_run * LOG2E;
        float ps4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int blk = 0; blk < KT / 16; ++blk) {
          uint32_t ph[16], pc[16];
          if (live & (1u << blk)) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float p0 = ex2f(fmaf(s[16 * blk + e], LOG2E, mb)), p1 = ex2f(fmaf(s[16 * blk + e + 1], LOG2E, mb));
              ps4[(e >> 1) & 3] += p0 + p1;
              ph[e] = __float_as_uint(p0);
              ph[e + 1] = __float_as_uint(p1);
              pc[e >> 1] = bf16x2_bits(p0, p1);
              pc[8 + (e >> 1)] = bf16x2_bits(tf32_rest_exact(p0), tf32_rest_exact(p1));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) ph[e] = pc[e] = 0u;
          }
          tmem_st16(trow + COL_S + 64 * buf + 16 * blk, ph);    // raw fp32 probabilities, in place of the scores
          tmem_st16(trow + COL_PC + 64 * buf + 16 * blk, pc);   // [bf16(p) x16 | bf16(rest p) x16]
        }
        l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        if (t == it.nt - 1) {
          const bool live = (i < it.len) && (l_run > 0.0f);
          fin_s[((item_g & 1) * BQ + r) * 2] = live ? 1.0f / l_run : 0.0f;
          fin_s[((item_g & 1) * BQ + r) * 2 + 1] = live ? m_run + logf(l_run) : 0.0f;
          tc::bar_arrive(BAR(B_FREADY0 + (item_g & 1)));
        }
        tc::tmem_wait_st();
        tc::tc_fence_before();
        tc::bar_arrive(BAR(B_PREADY0 + buf));
        AP_STAMP(0, 6 + 4 * t);
      }
      ++item_g;
    }
  } else if (warp < 8) {
    // =============================== correction, Q loader and epilogue warps: thread = query row ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int q = warp - 4, r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float *tr = reinterpret_cast<float *>(smem + C::OFF_TR) + q * 32 * TRS;
    const float inv_scale = 1.0f / sqrtf((float)HD);   // HF divides q by sqrt(head_dim) (:513); the product differs by <= 1 ulp
    long long *prof = (blockIdx.x == 0 && threadIdx.x == 128) ? g_atc_prof : nullptr;
    uint32_t item_g = 0, tile_g = 0;

    // Q of one item: coalesced loads (a warp instruction = 4 rows x 128 contiguous bytes), transposed through the warp's
    // buffer so that every lane holds ITS row, raw fp32 -> tensor memory, packed correction operand -> shared memory.
    // `idx` = position of the item in this CTA's sequence of active items.
    auto load_q = [&](const Item &qi, uint32_t idx) {
      tc::bar_wait_wd(BAR(B_QFREE), (idx & 1) ^ 1);   // every S = Q K^T of the previous item has completed
      tc::tc_fence_after();
      const float *qbase = qkv + qi.row0 * ld + qi.head * HD;
      const int rr0 = lane >> 3, c4 = lane & 7;
      float4 nx[8];
      auto fetch = [&](int g) {
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
          const int gi = qi.q0 + q * 32 + rr0 + 4 * i2;
          nx[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gi < qi.Sq && 32 * g + 4 * c4 < HD)
            nx[i2] = __ldg(reinterpret_cast<const float4 *>(qbase + (int64_t)gi * ld + 32 * g) + c4);
        }
      };
      fetch(0);
#pragma unroll 1   // (rolled: the instruction cache holds five roles' loops at once)
      for (int g = 0; g < NG; ++g) {
        const int wd = min(32, HD - 32 * g);
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2)
          *reinterpret_cast<float4 *>(tr + (rr0 + 4 * i2) * TRS + 4 * c4) =
              make_float4(nx[i2].x * inv_scale, nx[i2].y * inv_scale, nx[i2].z * inv_scale, nx[i2].w * inv_scale);
        __syncwarp();
        if (g + 1 < NG) fetch(g + 1);   // the next group's loads fly while this one is converted
        float4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4 *>(tr + lane * TRS + 4 * j);
        __syncwarp();
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {   // the two 16-column chunks of the group
          if (16 * hc < wd) {
            const int c = 2 * g + hc;
            uint32_t raw[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              raw[4 * j] = __float_as_uint(x[4 * hc + j].x); raw[4 * j + 1] = __float_as_uint(x[4 * hc + j].y);
              raw[4 * j + 2] = __float_as_uint(x[4 * hc + j].z); raw[4 * j + 3] = __float_as_uint(x[4 * hc + j].w);
            }
            tmem_st16(trow + COL_Q + 16 * c, raw);
            // packed correction operand, A side: per 16 columns [bf16(x) x16 | bf16(rest) x16] = 4 chunks of 16 bytes
            const uint32_t rowa = sbase + C::OFF_QC + (uint32_t)(g * (BQ * 128) + r * 128);
            const int cb = hc * 4, sw = r & 7;
            sts128(rowa + (((cb + 0) ^ sw) << 4), bf16x8(x[4 * hc], x[4 * hc + 1]));
            sts128(rowa + (((cb + 1) ^ sw) << 4), bf16x8(x[4 * hc + 2], x[4 * hc + 3]));
            sts128(rowa + (((cb + 2) ^ sw) << 4), bf16x8(rest4(x[4 * hc]), rest4(x[4 * hc + 1])));
            sts128(rowa + (((cb + 3) ^ sw) << 4), bf16x8(rest4(x[4 * hc + 2]), rest4(x[4 * hc + 3])));
          }
        }
      }
      tc::tmem_wait_st();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::bar_arrive(BAR(B_QREADY));
    };

    Item cur, nxt;
    int cur_i = next_active(blockIdx.x, cur);
    if (cur_i < n_items) load_q(cur, 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      float inv = 0.0f, lse_v = 0.0f;
      if (it.active) {
        const int nxt_i = next_active(item + gridDim.x, nxt);
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
#pragma unroll 1
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
          AP_STAMP(1, t);
        }
        // while the last tile's softmax and P V run: bring in the NEXT item's Q (its S products can then start before this
        // item's epilogue is over)
        if (nxt_i < n_items) load_q(nxt, item_g + 1);
        AP_STAMP(1, 19);
        tc::bar_wait_wd(BAR(B_FREADY0 + (item_g & 1)), (item_g >> 1) & 1);
        inv = fin_s[((item_g & 1) * BQ + r) * 2];
        lse_v = fin_s[((item_g & 1) * BQ + r) * 2 + 1];
        tc::bar_wait_wd(BAR(B_OFULL), (tile_g - 1) & 1);   // the last P V of the item has completed
        tc::tc_fence_after();
        AP_STAMP(1, 20);
      }
      // ---- o / l, transposed through the warp's buffer, row-contiguous stores (+ the next GEMM's correction operand) ----
      // (inactive item = a block of padded queries: exact zeros, HF :578; nothing exists there in the ragged layout)
      if (it.q0 < it.Sq) {
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
          const int wd = min(32, HD - 32 * g);
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0.0f;
          if (it.active) {
            if (wd == 32) tmem_ld32_nowait(trow + COL_O + 32 * g, v);
            else tmem_ld16_nowait(trow + COL_O + 32 * g, v);
            tmem_wait_ld();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4 *>(tr + lane * TRS + 4 * j) =
                make_float4(v[4 * j] * inv, v[4 * j + 1] * inv, v[4 * j + 2] * inv, v[4 * j + 3] * inv);
          __syncwarp();
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            const int rr = (lane >> 2) + 8 * i2, c8 = lane & 3, gi = it.q0 + q * 32 + rr;
            if (8 * c8 < wd && gi < it.Sq) {
              const float4 a = *reinterpret_cast<const float4 *>(tr + rr * TRS + 8 * c8);
              const float4 b2 = *reinterpret_cast<const float4 *>(tr + rr * TRS + 8 * c8 + 4);
              const int col = it.head * HD + 32 * g + 8 * c8;
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
      if (it.active) {
        AP_STAMP(1, 21);
        ++item_g;
      }
      if (lse && it.q0 + r < S) lse[((int64_t)it.b * nheads + it.head) * S + it.q0 + r] = lse_v;
    }
  } else if (warp < 12) {
    // =============================== K (warps 8-9) and V (warps 10-11) producers ===============================
    // a producer thread keeps a whole tile's share (up to 16 units of 8 floats) in registers between the prefetch and the
    // hand-over of the buffer: it takes the registers the correction and MMA warpgroups release
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int which = (warp < 10) ? 1 : 2;                 // column block of the qkv row: 1 = k, 2 = v
    const int wp = warp & 1;                               // my warp inside the producer pair
    // lane -> (row inside a group of 8 keys, unit inside a 32-column plane): a warp instruction touches 8 rows x 128
    // contiguous bytes in global memory, and its four 8-lane phases hit 8 distinct 16-byte bank groups in shared memory
    const int rl = lane & 7, cq = lane >> 3;
    constexpr int NJ = 4 * C::NP;                          // (8-key group, plane) pairs per warp
    const uint32_t hi_base = sbase + (uint32_t)(which == 1 ? C::OFF_KH : C::OFF_VH);
    const uint32_t co_base = sbase + (uint32_t)(which == 1 ? C::OFF_KC : C::OFF_VC);
    const uint32_t full = BAR(which == 1 ? B_KFULL : B_VFULL), empty = BAR(which == 1 ? B_KEMPTY : B_VEMPTY);
    // Every row this lane touches is (a multiple of 8) + rl, so all swizzle terms depend on (rl, cq) only: the addresses
    // inside a tile are per-thread constants plus compile-time offsets.
    //   raw plane     K (K-major, SWIZZLE_128B): 16-byte chunks 2 cq, 2 cq + 1 XOR rl
    //                 V (MN-major tf32, SWIZZLE_128B_BASE32B): 32-byte unit cq XOR (rl & 3)
    //   correction    K (K-major packed, B side: per 16 columns [bf16(rest) x16 | bf16(x) x16]): chunks cb, cb + 2 XOR rl
    //                 V (MN-major packed: K' rows per 16 keys [rest v x16 | v x16], 64 head columns per 128-byte row):
    //                   chunk (4 (plane & 1) + cq) XOR rl of K' row (key >> 4) * 32 + (key & 15) (+ 16 for the v half)
    const uint32_t h_off0 = hi_base + (uint32_t)(rl * 128 + (which == 1 ? (((2 * cq) ^ rl) << 4) : ((cq ^ (rl & 3)) << 5)));
    const uint32_t h_off1 = hi_base + (uint32_t)(rl * 128 + (which == 1 ? (((2 * cq + 1) ^ rl) << 4) : (((cq ^ (rl & 3)) << 5) + 16)));
    const int cbk = (cq >> 1) * 4 + (cq & 1);
    const uint32_t c_offA = co_base + (uint32_t)(rl * 128 + (which == 1 ? ((cbk ^ rl) << 4) : ((cq ^ rl) << 4)));
    const uint32_t c_offB = co_base + (uint32_t)(rl * 128 + (which == 1 ? (((cbk + 2) ^ rl) << 4) : (((4 + cq) ^ rl) << 4)));
    long long *prof = (blockIdx.x == 0 && (threadIdx.x == 256 || threadIdx.x == 320)) ? g_atc_prof : nullptr;
    uint32_t tile_g = 0, item_g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = make_item(item, nheads, n_qb, S, w, lengths, offsets);
      if (!it.active) continue;
      const float *gbase = qkv + it.row0 * ld + which * d + it.head * HD + 8 * cq;
#pragma unroll 1
      for (int t = 0; t < it.nt; ++t, ++tile_g) {
        const int k0 = it.kbeg + t * KT;
        const int nrows = it.kend - k0 - rl;   // my row 8 g + rl exists iff 8 g < nrows
        const float *tbase = gbase + (int64_t)(k0 + rl) * ld;
        float4 va[NJ], vb[NJ];
        AP_STAMP(1 + which, 3 * t);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int cmb = 2 * j + wp, g8 = cmb & 7, pl = cmb >> 3;
          va[j] = vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);   // keys beyond kend: zeros (P is 0 there; 0 * garbage could be NaN)
          if (4 * pl + cq < C::C8 && 8 * g8 < nrows) {
            const float4 *src = reinterpret_cast<const float4 *>(tbase + (int64_t)(8 * g8) * ld + 32 * pl);
            va[j] = __ldg(src);
            vb[j] = __ldg(src + 1);
          }
        }
        tc::bar_wait_wd(empty, (tile_g & 1) ^ 1);   // the MMAs that read the previous tile have completed
        AP_STAMP(1 + which, 3 * t + 1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int cmb = 2 * j + wp, g8 = cmb & 7, pl = cmb >> 3;
          if (4 * pl + cq >= C::C8) continue;   // (HD = 112: the last plane is half used)
          const uint32_t hofs = (uint32_t)(pl * (KT * 128) + g8 * 1024);
          sts128(h_off0 + hofs, f4_bits(va[j]));
          sts128(h_off1 + hofs, f4_bits(vb[j]));
          const uint4 xb = bf16x8(va[j], vb[j]), rb = bf16x8(rest4(va[j]), rest4(vb[j]));
          if (which == 1) {
            sts128(c_offA + hofs, rb);
            sts128(c_offB + hofs, xb);
          } else {
            const uint32_t cofs = (uint32_t)((pl >> 1) * (2 * KT * 128) + ((g8 >> 1) * 32 + (g8 & 1) * 8) * 128);
            const uint32_t ca = (pl & 1) ? c_offB : c_offA;
            sts128(ca + cofs, rb);
            sts128(ca + cofs + 16 * 128, xb);
          }
        }
        tc::fence_proxy_async();
        tc::bar_arrive(full);
        AP_STAMP(1 + which, 3 * t + 2);
      }
      ++item_g;
    }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");   // all four warps of the group (three of them only exist for this)
  }
  if (warp == 12) {
    // =============================== MMA issuer: warp-uniform control flow, one elected lane ===============================
    constexpr uint32_t id_o = tc::idesc_tf32(BQ, HD) | B_MN_MAJOR, id_oc = tc::idesc_bf16(BQ, HD) | B_MN_MAJOR;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = tc::elect_one();
    const uint32_t qc_a = sbase + C::OFF_QC, kh_a = sbase + C::OFF_KH, kc_a = sbase + C::OFF_KC;
    const uint32_t vh_a = sbase + C::OFF_VH, vc_a = sbase + C::OFF_VC;
    uint32_t item_g = 0, tile_g = 0;
    long long *prof = (blockIdx.x == 0 && lane == 0) ? g_atc_prof : nullptr;

    const uint64_t kh_desc = tc::desc_sw128(kh_a), qc_desc = tc::desc_sw128(qc_a), kc_desc = tc::desc_sw128(kc_a);
    const uint64_t vh_desc = desc_sw128b32_mn(vh_a, KT * 128), vc_desc = desc_sw128_mn(vc_a, 2 * KT * 128);
    // S = Q K^T of the tile with running index tg (of whichever item it belongs to); `last`: the item's last tile
    auto issue_qk = [&](uint32_t tg, bool last, int slot) {
      const uint32_t buf = tg & 1;
      tc::bar_wait_wd(BAR(B_KFULL), tg & 1);
      tc::tc_fence_after();
      AP_STAMP(4, slot);
      atc::issue_qk<HD>(leader, tb + COL_S + 64 * buf, tb + COL_Q, kh_desc, qc_desc, kc_desc);
      if (leader) {
        tc::umma_commit(BAR(B_SFULL0 + buf));
        tc::umma_commit(BAR(B_KEMPTY));
        if (last) tc::umma_commit(BAR(B_QFREE));
      }
      __syncwarp();
    };

    Item cur, nxt;
    int cur_i = next_active(blockIdx.x, cur);
    if (cur_i < n_items) {
      tc::bar_wait_wd(BAR(B_QREADY), 0);
      tc::tc_fence_after();
      AP_STAMP(4, 0);
      issue_qk(0, cur.nt == 1, 1);
    }
    while (cur_i < n_items) {
      const int nxt_i = next_active(cur_i + gridDim.x, nxt);
      bool nxt_started = false;
#pragma unroll 1
      for (int t = 0; t < cur.nt; ++t, ++tile_g) {
        // keep one S tile ahead of the softmax: the next tile of this item, or -- on the last tile -- the first tile of the
        // next item if its Q has already landed (otherwise after this tile's P V)
        if (t + 1 < cur.nt) {
          issue_qk(tile_g + 1, t + 2 == cur.nt, 2 + 6 * t);
        } else if (nxt_i < n_items && __all_sync(0xffffffffu, tc::bar_try_wait(BAR(B_QREADY), (item_g + 1) & 1))) {
          tc::tc_fence_after();
          issue_qk(tile_g + 1, nxt.nt == 1, 2 + 6 * t);
          nxt_started = true;
        }
        AP_STAMP(4, 3 + 6 * t);
        const uint32_t buf = tile_g & 1;
        tc::bar_wait_wd(BAR(B_PREADY0 + buf), (tile_g >> 1) & 1);
        AP_STAMP(4, 4 + 6 * t);
        tc::bar_wait_wd(BAR(B_VFULL), tile_g & 1);
        AP_STAMP(4, 5 + 6 * t);
        tc::bar_wait_wd(BAR(B_OREADY), tile_g & 1);   // O rescaled for this tile (t > 0) / free to be overwritten (t == 0)
        tc::tc_fence_after();
        AP_STAMP(4, 6 + 6 * t);
        const uint32_t d_o = tb + COL_O;
#pragma unroll
        for (int kg = 0; kg < KT / 8; ++kg) {
          const uint64_t bd = desc_add(vh_desc, (uint32_t)(kg * 1024));
          if (leader) tc::umma_tf32_ts(d_o, tb + COL_S + 64 * buf + 8 * kg, bd, id_o, (t | kg) != 0);
        }
#pragma unroll
        for (int j = 0; j < 2 * KT / 16; ++j) {
          const uint64_t bd = desc_add(vc_desc, (uint32_t)(j * 2048));
          if (leader) tc::umma_bf16_ts(d_o, tb + COL_PC + 64 * buf + 8 * j, bd, id_oc, 1);
        }
        if (leader) {
          tc::umma_commit(BAR(B_OFULL));
          tc::umma_commit(BAR(B_VEMPTY));
        }
        __syncwarp();
        AP_STAMP(4, 7 + 6 * t);
      }
      ++item_g;
      if (nxt_i < n_items && !nxt_started) {
        tc::bar_wait_wd(BAR(B_QREADY), item_g & 1);
        tc::tc_fence_after();
        issue_qk(tile_g, nxt.nt == 1, 1);
      }
      cur = nxt;
      cur_i = nxt_i;
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

// profiling hook: buf = device buffer of AP_ITEMS * AP_ROLES * AP_SLOTS int64 (or NULL to switch the timeline off)
extern "C" int mts_debug_attn_profile(long long *buf) {
  MTS_CUDA(cudaMemcpyToSymbol(atc::g_atc_prof, &buf, sizeof(buf)));
  return 0;
}

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

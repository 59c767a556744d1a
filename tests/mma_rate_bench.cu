// Micro-benchmark (not a test): issue-to-completion rate of small tcgen05.mma instructions, the recurrence's bound at
// large batch.  One CTA, one elected lane issues `n` MMAs into one accumulator, commits, waits; cycles / MMA.
//   A from tensor memory (.ts) or shared memory (.ss); M in {64, 128}; N in {16, 32, 64}; kind tf32 (K = 8).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I multimodaltopicsegmentation_b200/csrc \
//        -o tests/_bin/mma_rate_bench tests/mma_rate_bench.cu
#include <cstdio>
#include "tcgen05_utils.cuh"
using namespace mts;

template <int M, int N, bool TS>
__global__ void __launch_bounds__(128, 1) bench(int n, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (128 * 128 + 64 * 128) / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.001f * (i & 63);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { tc::bar_init(tc::s_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tc::tmem_alloc<512>(tc::s_u32(&slot));
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::idesc_tf32(M, N);
    const bool leader = tc::elect_one();
    const uint64_t adesc = tc::desc_sw128(tc::s_u32(smem));
    const uint64_t bdesc = tc::desc_sw128(tc::s_u32(smem + 128 * 128));
    uint32_t ph = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        if (leader) {
          if (TS) tc::umma_tf32_ts(tb + 256, tb + (uint32_t)((i & 31) * 8), bdesc + (uint64_t)((i & 3) * 2), idesc, i != 0);
          else tc::umma_tf32_ss(tb + 256, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, i != 0);
        }
      }
      const long long t1 = clock64();
      if (leader) tc::umma_commit(tc::s_u32(&bar));
      __syncwarp();
      tc::bar_wait(tc::s_u32(&bar), ph); ph ^= 1;
      const long long t2 = clock64();
      if (leader && rep == 2) { out[0] = (t1 - t0); out[1] = (t2 - t0); }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tb); }
}

template <int M, int N, bool TS>
static void run(const char *name) {
  long long *out; cudaMalloc(&out, 16);
  const int n = 256;
  cudaFuncSetAttribute(bench<M, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  bench<M, N, TS><<<1, 128, 64 * 1024>>>(n, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%-22s issue %6.1f cycles/MMA, to completion %6.1f cycles/MMA  [%s]\n", name, (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<128, 16, true>("ts M=128 N=16");
  run<64, 16, true>("ts M=64  N=16");
  run<128, 32, true>("ts M=128 N=32");
  run<128, 64, true>("ts M=128 N=64");
  run<64, 8, true>("ts M=64  N=8");
  run<128, 16, false>("ss M=128 N=16");
  run<64, 16, false>("ss M=64  N=16");
  run<128, 256, false>("ss M=128 N=256");
  return 0;
}

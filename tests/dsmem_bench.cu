// Micro-benchmark (not a test): all-gather of 2 KB per CTA inside an 8-CTA cluster, the h exchange of the recurrence
// kernels.  (a) st.async.v4 from 128 threads to the 8 CTAs (what lstm_fwd_tc_kernel does), (b) 8 bulk copies
// cp.async.bulk.shared::cluster.shared::cta of 2 KB issued by 8 threads.  Prints cycles per exchange.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tests/_bin/dsmem_bench tests/dsmem_bench.cu
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
  uint32_t o;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
  return o;
}
__device__ __forceinline__ void bar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void bar_expect(uint32_t b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(160, 1) bench(int iters, long long *out, float *sink) {
  __shared__ __align__(1024) float recv[2][8][512];   // [parity][sender slot][2 KB]
  __shared__ __align__(128) float stage[2][512];
  __shared__ __align__(8) uint64_t full[2];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int tid = threadIdx.x;
  if (tid == 0) { bar_init(s_u32(&full[0]), 1); bar_init(s_u32(&full[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  cluster.sync();
  uint32_t ph[2] = {0, 0};
  float acc = 0.f;
  long long t0 = 0;
  for (int it = 0; it < iters; ++it) {
    const int p = it & 1;
    if (it == 8) t0 = clock64();
    if (tid == 0) bar_expect(s_u32(&full[p]), 8 * 2048);
    __syncthreads();  // (the real kernel arms one step ahead; here every CTA arms before anyone can send: cluster-wide order below)
    cluster.sync();
    if (tid >= 32) {  // 128 "cell" threads
      const int t = tid - 32;
      const float4 v = make_float4(it + t, acc, rank, 1.f);
      if (MODE == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t dst = mapa(s_u32(&recv[p][rank][4 * t]), (rank + r) & 7);
          const uint32_t bar = mapa(s_u32(&full[p]), (rank + r) & 7);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                       "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(bar) : "memory");
        }
      } else {
        *reinterpret_cast<float4 *>(&stage[p][4 * t]) = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t < 8) {
          const uint32_t dst = mapa(s_u32(&recv[p][rank][0]), (rank + t) & 7);
          const uint32_t bar = mapa(s_u32(&full[p]), (rank + t) & 7);
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "r"(s_u32(&stage[p][0])), "r"(2048), "r"(bar) : "memory");
        }
      }
    }
    bar_wait(s_u32(&full[p]), ph[p]);
    ph[p] ^= 1;
    acc += recv[p][(tid >> 2) & 7][tid];
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / (iters - 8);
  sink[blockIdx.x * blockDim.x + tid] = acc;
}

int main() {
  long long *out; float *sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 64 * 160 * 4);
  const int iters = 2008;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) bench<0><<<64, 160>>>(iters, out, sink); else bench<1><<<64, 160>>>(iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("%s: %lld cycles per exchange (incl. one cluster.sync)  [%s]\n", mode == 0 ? "st.async x 8 per thread" : "8 bulk copies of 2 KB", h, cudaGetErrorString(e));
    }
  }
  // baseline: the loop with no transfer is not measurable here (the barrier would never complete); cluster.sync alone:
  return 0;
}

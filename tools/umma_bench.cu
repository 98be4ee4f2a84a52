// Micro-benchmark: tcgen05.mma (kind::f16, fp16, M=128, cta_group::1, SS, K-major SWIZZLE_NONE) throughput
// as a function of N, accumulator rotation and operand start alignment.  One CTA per SM, unrolled issue loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../wakeword_jupyterlab_b200/csrc/tc_common.cuh"
using namespace tc;

struct Cfg { int N, a_off, lbo_a, iters, swap; };

template <int NACC>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, c.N);
    const uint64_t ad = make_desc(smem_u32(smem) + c.a_off * 16, c.lbo_a, 128);
    const uint64_t bd = make_desc(smem_u32(smem) + 100 * 1024, 2048, 128);
    const uint64_t x = c.swap ? bd : ad, y = c.swap ? ad : bd;
    long long t0 = clock64();
    for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_f16(tm + (u % NACC) * (512 / NACC), x + u * 3, y + u * 3, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("N  n_acc a_off lbo_a  swap  cycles/MMA (ideal N/2)\n");
  for (int N : {32, 64, 128, 256})
    for (int n_acc : {1, 2, 4})
      for (int a_off : {0, 1})
        for (int lbo : {2048, 18688})
          for (int swap : {0, 1}) {
            if (N == 256 && n_acc == 4) continue;
            Cfg c{N, a_off, lbo, 4000, swap};
            if (n_acc == 1) bench<1><<<148, 128, 200 * 1024>>>(c, d);
            else if (n_acc == 2) bench<2><<<148, 128, 200 * 1024>>>(c, d);
            else bench<4><<<148, 128, 200 * 1024>>>(c, d);
            long long h = 0;
            cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%3d %d %d %5d %d  %7.1f  (%d)\n", N, n_acc, a_off, lbo, swap, (double)h / c.iters, N / 2);
          }
  return 0;
}

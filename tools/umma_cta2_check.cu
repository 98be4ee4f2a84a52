// Self-checking probe for the cta_group::2 form of conv12: one tcgen05.mma over a CTA PAIR,
// M = 256 (each CTA supplies its own 128 rows of A from its own shared memory), N = 128 with the B operand SPLIT between the
// two CTAs (each holds 64 of the 128 rows of B), K = 32 (two K = 16 instructions), fp16 -> fp32, K-major SWIZZLE_NONE operands
// exactly as the conv kernels lay them out.  Exact integer data; the host compares every accumulator bitwise and, if the
// assumed split (CTA r holds B rows 64 r .. 64 r + 63 = output columns 64 r ..) is wrong, says whether the swapped split
// matches instead.  Then times a burst of MMAs: with half of B per CTA an N = 128 instruction reads 4 + 2 KB per CTA, so it
// should run at the tensor rate (64 cycles) instead of the shared-memory-bound rate of cta_group::1.
// Run it FIRST when building 8.2.  Not linked into the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_cta2_check tools/umma_cta2_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../wakeword_jupyterlab_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int kK = 32, kNh = 64;                    // K, and the rows of B each CTA holds
constexpr int kABytes = 4 * 128 * 16;               // [kc 4][row 128][8 fp16]
constexpr int kBBytes = 4 * kNh * 16;               // [kc 4][n 64][8 fp16]

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all earlier MMAs of the pair -> one arrival on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
check(const __half* a_tiled, const __half* b_tiled, float* out, long long* cyc, int burst) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_s = smem;
  unsigned char* b_s = smem + kABytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + kBBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t rank = cluster_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // this CTA's 128 rows of A and its 64 rows of B (pre-tiled on the host, one block per rank)
  for (int i = tid * 16; i < kABytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(a_s + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a_tiled) + rank * kABytes + i);
  for (int i = tid * 16; i < kBBytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(b_s + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(b_tiled) + rank * kBBytes + i);
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (warp == 0) tmem_alloc2(slot, 256);            // one warp of EACH CTA of the pair
  tc_fence_before();
  cluster_sync_all();                               // both CTAs' operands, barriers and TMEM are ready
  tc_fence_after();
  const uint32_t tm = *slot;
  if (rank == 0 && tid == 0) {                      // only the leader CTA issues
    const uint32_t id = make_idesc(256, 128);
    const uint64_t ad = make_desc(smem_u32(a_s), 128 * 16, 128);
    const uint64_t bd = make_desc(smem_u32(b_s), kNh * 16, 128);
    umma2_f16(tm, ad, bd, id, 0);
    umma2_f16(tm, ad + (uint64_t)((2 * 128 * 16) >> 4), bd + (uint64_t)((2 * kNh * 16) >> 4), id, 1);
    umma2_commit_both(bar);
  }
  mbar_wait(bar, 0, 1);                             // each CTA waits on its own copy of the barrier
  tc_fence_after();
  uint32_t r[32];
  for (int h = 0; h < 4; ++h) {
    tmem_ld32_nowait(tm + ((uint32_t)(warp * 32) << 16) + h * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + h * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  // ---- burst timing: `burst` MMAs into the second half of the allocation
  if (rank == 0 && tid == 0) {
    const uint32_t id = make_idesc(256, 128);
    const uint64_t ad = make_desc(smem_u32(a_s), 128 * 16, 128);
    const uint64_t bd = make_desc(smem_u32(b_s), kNh * 16, 128);
    const long long c0 = clock64();
    for (int i = 0; i < burst; ++i) umma2_f16(tm + 128, ad, bd, id, 1);
    umma2_commit_both(bar + 1);
    mbar_wait(bar + 1, 0, 2);
    *cyc = clock64() - c0;
  } else if (rank == 1 && tid == 0) {
    mbar_wait(bar + 1, 0, 3);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tm, 256);
}

int main() {
  srand(3);
  std::vector<float> A(256 * kK), B(128 * kK);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : B) v = (float)(rand() % 7 - 3);
  std::vector<__half> at(2 * kABytes / 2), bt(2 * kBBytes / 2);
  for (int m = 0; m < 256; ++m)
    for (int k = 0; k < kK; ++k) at[(size_t)(m / 128) * (kABytes / 2) + ((k / 8) * 128 + m % 128) * 8 + k % 8] = __float2half(A[m * kK + k]);
  for (int n = 0; n < 128; ++n)      // assumed split: CTA r holds rows 64 r .. 64 r + 63 of B
    for (int k = 0; k < kK; ++k) bt[(size_t)(n / kNh) * (kBBytes / 2) + ((k / 8) * kNh + n % kNh) * 8 + k % 8] = __float2half(B[n * kK + k]);
  __half *da, *db; float* dout; long long* dcyc;
  cudaMalloc(&da, at.size() * 2); cudaMalloc(&db, bt.size() * 2); cudaMalloc(&dout, 256 * 128 * 4); cudaMalloc(&dcyc, 8);
  cudaMemcpy(da, at.data(), at.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 256 * 128 * 4);
  const int smem = kABytes + kBBytes + 64, burst = 2000;
  check<<<2, 128, smem>>>(da, db, dout, dcyc, burst);
  std::vector<float> got(256 * 128);
  cudaError_t e = cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long cyc = 0;
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  int bad = 0, bad_swapped = 0, shown = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      float ref = 0, ref_sw = 0;
      for (int k = 0; k < kK; ++k) { ref += A[m * kK + k] * B[n * kK + k]; ref_sw += A[m * kK + k] * B[((n + 64) % 128) * kK + k]; }
      const float g = got[m * 128 + n];
      if (g != ref) { ++bad; if (shown++ < 6) printf("  D[%d][%d] = %g, expected %g (swapped split would give %g)\n", m, n, g, ref, ref_sw); }
      if (g != ref_sw) ++bad_swapped;
    }
  printf("M = 256, N = 128 split 64 | 64 across the pair: %s (%d of %d differ; %d differ from the swapped split)\n",
         bad ? "FAIL" : "PASS", bad, 256 * 128, bad_swapped);
  printf("burst of %d MMAs (M 256, N 128, K 16): %.1f cycles each (tensor rate 64; cta_group::1 N = 128 measured 64.1 alone, ~89 under LSU load)\n",
         burst, (double)cyc / burst);
  return bad ? 2 : 0;
}

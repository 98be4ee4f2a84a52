// Self-checking probe for stage 1 of the tensor-core GEMM-FFT log-mel: the A operand of
// tcgen05.mma as an MN-major SWIZZLE_NONE *overlapping view* of ONE fp16 copy of the zero-padded clip.
//
//   clip copy:  rows of 64 samples, [row block r/8][column chunk c/8 (8)][r%8][8 fp16]  -> 128-byte core matrices
//   frame t  :  rows 8t .. 8t+31 (hop 512 = 8 rows, n_fft 2048 = 32 rows);  x_t[n1 + 64 n2] = S[8t + n2][n1]
//   A view   :  M = 128 = two frames x 64 n1, K = n2;  start = t * 1024 B, M-core stride (SBO) 128 B, K-core stride (LBO) 1 KB
//   D[(f, n1), j] = sum_n2 x_{t+f}[n1 + 64 n2] * B[j][n2]          (B: [N = 64][K = 32] fp16, K-major, as the conv kernels)
//
// All values are small integers, so the fp32 accumulators are exact and the comparison with the host loop is bitwise.
// Prints PASS / FAIL per frame pair and the cycles of the two MMAs.  Run it FIRST next round: it pins the descriptor
// encoding (a_major bit 15, LBO / SBO roles for MN-major) before any of the kernel is written around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mn_major_view_check tools/mn_major_view_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../wakeword_jupyterlab_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int kRows = 288;                 // 18,048 padded samples = 282 rows of 64, rounded up to whole row blocks
constexpr int kClipBytes = kRows * 64 * 2; // 36,864
constexpr int kN = 64, kK = 32;
constexpr int kBBytes = kN * kK * 2;       // [kc 4][n 64][8 fp16]

// instruction descriptor with A MN-major (bit 15), B K-major
__host__ __device__ constexpr uint32_t idesc_a_mn(int M, int N) { return make_idesc(M, N) | (1u << 15); }

__global__ void __launch_bounds__(128, 1) check(const __half* clip_tiled, const __half* b_tiled, int t0, float* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* a_s = smem;
  unsigned char* b_s = smem + kClipBytes;
  for (int i = threadIdx.x * 16; i < kClipBytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(a_s + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(clip_tiled) + i);
  for (int i = threadIdx.x * 16; i < kBBytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(b_s + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(b_tiled) + i);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    // A: MN-major, SBO = stride between core matrices along M (128 B), LBO = stride between core matrices along K (1 KB)
    const uint64_t ad = make_desc(smem_u32(a_s) + (uint32_t)t0 * 1024u, /*lbo*/ 1024, /*sbo*/ 128);
    // B: K-major [kc][n][8]: LBO = K-chunk stride (64 rows x 16 B), SBO = 8-row group stride (128 B)
    const uint64_t bd = make_desc(smem_u32(b_s), kN * 16, 128);
    const uint32_t id = idesc_a_mn(128, kN);
    const long long c0 = clock64();
    umma_f16(tm, ad, bd, id, 0);                                               // n2 = 0..15
    umma_f16(tm, ad + (2 * 1024 >> 4), bd + (2 * kN * 16 >> 4), id, 1);        // n2 = 16..31
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    *cyc = clock64() - c0;
  }
  __syncthreads();
  tc_fence_after();
  uint32_t r[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = 0; h < 2; ++h) {
    tmem_ld32_nowait(tm + ((uint32_t)(warp * 32) << 16) + h * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + h * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

int main() {
  std::vector<float> x(kRows * 64, 0.0f);
  srand(1);
  for (int i = 1024; i < 1024 + 16000; ++i) x[i] = (float)(rand() % 17 - 8);            // centre padding stays zero
  std::vector<__half> a(kRows * 64), b(kN * kK);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 64; ++c)
      a[(((r / 8) * 8 + c / 8) * 8 + r % 8) * 8 + c % 8] = __float2half(x[r * 64 + c]);
  std::vector<float> bm(kN * kK);
  for (int j = 0; j < kN; ++j)
    for (int k = 0; k < kK; ++k) {
      bm[j * kK + k] = (float)(rand() % 7 - 3);
      b[((k / 8) * kN + j) * 8 + k % 8] = __float2half(bm[j * kK + k]);
    }
  __half *da, *db; float* dout; long long* dcyc;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * kN * 4); cudaMalloc(&dcyc, 8);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kClipBytes + kBBytes;
  cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad_total = 0;
  for (int t0 : {0, 1, 2, 15, 30}) {                                             // frame pairs (t0, t0 + 1)
    check<<<1, 128, smem>>>(da, db, t0, dout, dcyc);
    std::vector<float> got(128 * kN);
    long long cyc = 0;
    cudaError_t e = cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int f = 0; f < 2; ++f)
      for (int n1 = 0; n1 < 64; ++n1)
        for (int j = 0; j < kN; ++j) {
          float ref = 0.0f;
          for (int n2 = 0; n2 < kK; ++n2) ref += x[(8 * (t0 + f) + n2) * 64 + n1] * bm[j * kK + n2];
          if (got[(f * 64 + n1) * kN + j] != ref) ++bad;
        }
    printf("frames %2d,%2d: %s (%d of %d differ), %lld cycles for the two MMAs\n", t0, t0 + 1, bad ? "FAIL" : "PASS", bad,
           128 * kN, cyc);
    bad_total += bad;
  }
  return bad_total ? 2 : 0;
}

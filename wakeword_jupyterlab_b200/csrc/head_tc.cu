// K4 on the tensor core: one LSTM layer at T = 1 over the whole batch (sm_100a, tcgen05 / TMEM).
//
// Replaces the two `self.lstm` layers of WakewordModel.forward (/root/reference/wakeword_training_script.py:175-180) like
// gated_dense_pipe_kernel in head.cu does (same inputs, same outputs, same gate arithmetic): per layer
//     G[B][3H] = X[B][K] W^T + b ;  h = sigmoid(G_o) tanh(sigmoid(G_i) tanh(G_g))
// The GEMM runs as tcgen05.mma.kind::tf32 with BOTH operands split into two TF32 terms (x = hi + lo, hi = the nearest
// TF32 value, lo = the nearest TF32 value of x - hi): hi*hi + lo*hi + hi*lo accumulated in fp32 leaves ~2^-22 of relative
// error per product, the accuracy class of an fp32 FMA chain of this length, and unlike an fp16
// split it needs no range scaling (TF32 keeps fp32's exponent).
//
// Work item = 128 clips x 64 hidden units x 3 gates (N = 192 accumulator columns: gate g of unit u at column 64 g + u, so a
// thread of the epilogue finds i, g, o of a unit in its own TMEM lane).  K is walked in chunks of 32: a stage holds the clip
// rows (hi | lo, 32 KB, written by 4 converter warps that read the fp32 activations as they are) and the weight rows (hi | lo,
// 48 KB, ONE bulk copy out of a layout packed on the device once per weight version); 12 instructions per stage (4 K = 8
// steps x 3 products), two stages, two accumulators (2 x 192 of the 512 TMEM columns): the epilogue of an item (8 warps,
// bias + gates + fp32 store) overlaps the instructions of the next.
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>

using namespace tc;

namespace {

constexpr int HT_ROWS = 128, HT_UNITS = 64, HT_N = 3 * HT_UNITS, HT_KC = 32;
constexpr int HT_A_PART = (HT_KC / 4) * HT_ROWS * 16;      // 16 KB: [k4 (8)][row (128)][4 x tf32]
constexpr int HT_B_PART = (HT_KC / 4) * HT_N * 16;         // 24 KB: [k4 (8)][row (192)][4 x tf32]
constexpr int HT_STAGE = 2 * HT_A_PART + 2 * HT_B_PART;    // 80 KB
constexpr int HT_NST = 2;
constexpr int HT_EPI_WARPS = 8, HT_CVT_WARPS = 4;
constexpr int HT_THREADS = (HT_EPI_WARPS + HT_CVT_WARPS + 2) * 32;
constexpr size_t HT_SMEM = (size_t)HT_NST * HT_STAGE + 16 * 8 + 16;

struct HeadTcParams {
  const float* x;        // [B][K]
  const float* wpack;    // [H / 64][K / 32][hi | lo][k4 (8)][row (192)][4]
  const float* bias;     // [3][H]
  float* out;            // [B][H]
  int B, K, H;
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// instruction descriptor (kind::tf32): D = f32 (bit 4), A = B = TF32 (format 2 at bits 7 and 10), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// nearest TF32 value (the tensor core itself truncates: rounding both terms here keeps the split unbiased)
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// [K][3][H] fp32 (head.cu's layout) -> the stage-ordered hi | lo operand of the kernel below
__global__ void head_tc_pack_kernel(const float* __restrict__ wt, float* __restrict__ out, int K, int H) {
  const int64_t total = (int64_t)(H / HT_UNITS) * (K / HT_KC) * 2 * (HT_KC / 4) * HT_N * 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int e = (int)(r & 3); r >>= 2;
    const int row = (int)(r % HT_N); r /= HT_N;
    const int k4 = (int)(r % (HT_KC / 4)); r /= (HT_KC / 4);
    const int part = (int)(r & 1); r >>= 1;
    const int kc = (int)(r % (K / HT_KC));
    const int ct = (int)(r / (K / HT_KC));
    const int g = row / HT_UNITS, u = row % HT_UNITS;
    const int k = kc * HT_KC + k4 * 4 + e;
    const float v = wt[((size_t)k * 3 + g) * H + ct * HT_UNITS + u];
    const float hi = tf32_rn(v);
    out[i] = part ? tf32_rn(v - hi) : hi;
  }
}

__global__ void __launch_bounds__(HT_THREADS, 1) gated_dense_tc_kernel(const HeadTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)HT_NST * HT_STAGE);
  uint64_t* full = bars;                  // [2] stage filled: 4 converter warps + the loader's expect_tx arrival
  uint64_t* empty = bars + 2;             // [2] stage consumed (tcgen05.commit)
  uint64_t* acc_full = bars + 4;          // [2] accumulator complete (tcgen05.commit)
  uint64_t* acc_empty = bars + 6;         // [2] accumulator drained (8 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < HT_NST; ++i) { mbar_init(full + i, HT_CVT_WARPS + 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, HT_EPI_WARPS); }
    fence_barrier_init();
  }
  constexpr int kMmaWarp = HT_EPI_WARPS + HT_CVT_WARPS + 1, kLoadWarp = HT_EPI_WARPS + HT_CVT_WARPS;
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int n_ct = p.H / HT_UNITS, n_kc = p.K / HT_KC;
  const int n_items = ((p.B + HT_ROWS - 1) / HT_ROWS) * n_ct;      // item = row tile * n_ct + column tile

  if (warp == kLoadWarp) {
    // ===================== weight loader (one thread): one 48 KB bulk copy per stage
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ct = item % n_ct;
        for (int kc = 0; kc < n_kc; ++kc, ++g) {
          const uint32_t s = g % HT_NST;
          mbar_wait(empty + s, ((g / HT_NST) & 1) ^ 1, 70);
          mbar_arrive_expect_tx(full + s, 2 * HT_B_PART);
          bulk_g2s(smem + (size_t)s * HT_STAGE + 2 * HT_A_PART,
                   p.wpack + ((size_t)ct * n_kc + kc) * (2 * HT_B_PART / 4), 2 * HT_B_PART, full + s);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp walks the loop, the elected lane issues)
    constexpr uint32_t idesc = make_idesc_tf32(HT_ROWS, HT_N);
    const uint64_t adesc0 = make_desc(smem_u32(smem), HT_ROWS * 16, 128);                      // K core matrices 2 KB apart
    const uint64_t bdesc0 = make_desc(smem_u32(smem + 2 * HT_A_PART), HT_N * 16, 128);         // 3 KB apart
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t a = it & 1;
      mbar_wait(acc_empty + a, ((it >> 1) & 1) ^ 1, 71);
      const uint32_t d = tmem_base + a * 256;
      for (int kc = 0; kc < n_kc; ++kc, ++g) {
        const uint32_t s = g % HT_NST;
        mbar_wait(full + s, (g / HT_NST) & 1, 72);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((s * HT_STAGE) >> 4), bd = bdesc0 + (uint64_t)((s * HT_STAGE) >> 4);
#pragma unroll
          for (int j = 0; j < HT_KC / 8; ++j) {
            const uint64_t aj = ad + (uint64_t)((2 * j * HT_ROWS * 16) >> 4), bj = bd + (uint64_t)((2 * j * HT_N * 16) >> 4);
            // small terms first: lo x hi, hi x lo, then hi x hi
            umma_tf32(d, aj + (HT_A_PART >> 4), bj, idesc, !(kc == 0 && j == 0));
            umma_tf32(d, aj, bj + (HT_B_PART >> 4), idesc, 1);
            umma_tf32(d, aj, bj, idesc, 1);
          }
          umma_commit(empty + s);
          if (kc == n_kc - 1) umma_commit(acc_full + a);
        }
        __syncwarp();
      }
    }
  } else if (warp >= HT_EPI_WARPS) {
    // ===================== converters (4 warps): thread = clip row; fp32 -> TF32 hi | lo, K-major core matrices
    const int r = (warp - HT_EPI_WARPS) * 32 + lane;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rt = item / n_ct;
      const int b = min(rt * HT_ROWS + r, p.B - 1);                 // rows past the batch repeat the last clip (never stored)
      const float4* __restrict__ xr = reinterpret_cast<const float4*>(p.x + (size_t)b * p.K);
      for (int kc = 0; kc < n_kc; ++kc, ++g) {
        const uint32_t s = g % HT_NST;
        float4 v[HT_KC / 4];
#pragma unroll
        for (int q = 0; q < HT_KC / 4; ++q) v[q] = __ldg(xr + kc * (HT_KC / 4) + q);
        mbar_wait_relaxed(empty + s, ((g / HT_NST) & 1) ^ 1, 73);
        unsigned char* ah = smem + (size_t)s * HT_STAGE + (size_t)r * 16;
#pragma unroll
        for (int q = 0; q < HT_KC / 4; ++q) {
          const float4 h = make_float4(tf32_rn(v[q].x), tf32_rn(v[q].y), tf32_rn(v[q].z), tf32_rn(v[q].w));
          *reinterpret_cast<float4*>(ah + (size_t)q * HT_ROWS * 16) = h;
          *reinterpret_cast<float4*>(ah + HT_A_PART + (size_t)q * HT_ROWS * 16) =
              make_float4(tf32_rn(v[q].x - h.x), tf32_rn(v[q].y - h.y), tf32_rn(v[q].z - h.z), tf32_rn(v[q].w - h.w));
        }
        fence_proxy_async();                                          // generic-proxy stores -> visible to the tensor core
        mbar_arrive_warp(full + s, lane);
      }
    }
  } else {
    // ===================== epilogue (8 warps): lane = clip row of the warp's TMEM quadrant, 32 units per warp
    const int q = warp & 3, half = warp >> 2;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int rt = item / n_ct, ct = item % n_ct;
      const uint32_t a = it & 1;
      const int b = rt * HT_ROWS + q * 32 + lane;
      mbar_wait_relaxed(acc_full + a, (it >> 1) & 1, 74);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256 + half * 32;
      const int j0 = ct * HT_UNITS + half * 32;
#pragma unroll 1
      for (int h16 = 0; h16 < 2; ++h16) {
        uint32_t ri[16], rg[16], ro[16];
        tmem_ld16_nowait(taddr + h16 * 16, ri);
        tmem_ld16_nowait(taddr + HT_UNITS + h16 * 16, rg);
        tmem_ld16_nowait(taddr + 2 * HT_UNITS + h16 * 16, ro);
        tmem_ld_wait();
        float hv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = j0 + h16 * 16 + u;
          const float gi = __uint_as_float(ri[u]) + __ldg(p.bias + j);
          const float gg = __uint_as_float(rg[u]) + __ldg(p.bias + p.H + j);
          const float go = __uint_as_float(ro[u]) + __ldg(p.bias + 2 * p.H + j);
          hv[u] = sigmoid_acc(go) * tanhf(sigmoid_acc(gi) * tanhf(gg));
        }
        if (b < p.B) {
          float4* o4 = reinterpret_cast<float4*>(p.out + (size_t)b * p.H + j0 + h16 * 16);
#pragma unroll
          for (int u = 0; u < 4; ++u) o4[u] = make_float4(hv[4 * u], hv[4 * u + 1], hv[4 * u + 2], hv[4 * u + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(acc_empty + a, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Layer `l` of the head on the tensor core.  Returns 1 when the shape is outside what the kernel handles (the caller then
// runs the fp32 kernel of head.cu), a negative error code on failure, WW_OK otherwise.
int ww_launch_gated_dense_tc(ww_ctx* c, int l, const float* x, float* out, int B, cudaStream_t st) {
  const int H = c->cfg.hidden_size, K = l == 0 ? 128 : H;
  const char* env = getenv("WW_HEAD_KERNEL");              // "fp32": the CUDA-core kernel (A/B runs, tests); read per call
  if ((env && strcmp(env, "fp32") == 0) || H % HT_UNITS || K % HT_KC || l >= 8) return 1;
  const size_t n_pack = (size_t)(H / HT_UNITS) * (K / HT_KC) * 2 * (HT_KC / 4) * HT_N * 4;
  if (!c->d_head_tc[l] || c->head_tc_version[l] != c->weights_version) {
    if (!c->d_head_tc[l]) WW_CHECK(c, cudaMalloc((void**)&c->d_head_tc[l], n_pack * sizeof(float)));
    head_tc_pack_kernel<<<(int)((n_pack + 255) / 256), 256, 0, st>>>(c->d_head_wt[l], c->d_head_tc[l], K, H);
    WW_LAUNCH_CHECK(c);
    c->head_tc_version[l] = c->weights_version;
  }
  HeadTcParams p;
  p.x = x; p.wpack = c->d_head_tc[l]; p.bias = c->d_head_b[l]; p.out = out; p.B = B; p.K = K; p.H = H;
  const int n_items = ((B + HT_ROWS - 1) / HT_ROWS) * (H / HT_UNITS);
  WW_CHECK(c, cudaFuncSetAttribute(gated_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_SMEM));
  gated_dense_tc_kernel<<<std::min(c->sm_count, n_items), HT_THREADS, HT_SMEM, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

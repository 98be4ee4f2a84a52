// K3a (tensor-core path): conv1 (CUDA cores, fp32) fused into conv2 (tcgen05 implicit GEMM) + bias + ReLU  (sm_100a)
//
// Replaces F.relu(self.conv1(x)) and F.relu(self.conv2(x)) of WakewordModel.forward
// (/root/reference/wakeword_training_script.py:170-171).  See conv3_tc.cu for the layout story.
//
// Work item = (clip, PAIR of 128-pixel tiles of the pixel-linear padded image).  Warp roles:
//   warps 0-15  producers (one pixel x 16 channels per thread): conv1 + ReLU in fp32 (packed FFMA2, weights read
//               from the constant bank as kernel parameters) for the 256 pixels and their 3x3 halo (2P + 2 more),
//               rounded to fp16 and stored as the K-major SWIZZLE_NONE A operand [chunk of 8 ch][pixel][16 B];
//               three A buffers so the producers run up to two items ahead of the tensor core; the log-mel rows an
//               item needs are staged one item ahead with cp.async;
//   warps 16-23 epilogue (two warps per TMEM lane quadrant, 32 output channels each): TMEM -> (D_hi + D_lo) * 2^-k +
//               bias, ReLU, zero the padding pixels, round to fp16, write the conv3 operand planes to HBM (512
//               contiguous bytes per warp store);
//   warps 24,25 MMA issuers (warp 24: even items / accumulator 0, warp 25: odd items / accumulator 1, so the ~100
//               cycle issue cost of each small MMA overlaps and the result stays deterministic).  Per 3x3 tap,
//               16-channel k-slice and tile:
//                 D[:, 0:128] += A x [W_hi ; W_lo]^T   (N = 128: both weight halves in one instruction; the
//               epilogue adds the two column halves; WW_CONV_FP16: N = 64, W_hi only)
//               -- the tap is only a start-address offset of the same shared-memory tile.
// conv2 weights (hi and lo stacked along N, 73,728 B) stay resident in shared memory.
#include "tc_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <type_traits>

using namespace tc;

#define C12_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && it < 48 && lane == 0) p.trace[it * 8 + (slot)] = clock64(); } while (0)

namespace {

constexpr int C12_THREADS = 832;
constexpr int W2_BYTES = 9 * 4 * 128 * 16;   // [tap][kc 4][n' 128 = 64 hi + 64 lo][8 fp16]
constexpr int NABUF_MAX = 3;                 // A-operand buffers (3 when shared memory allows, else 2)

struct Conv12Params {
  const float* logmel;            // [B][H][W]
  float w1[288];                  // conv1 weights [tap][cout] -- kernel parameters live in the constant bank, so
  float b1[32];                   // the FMAs read them as c[0][..] operands: no shared-memory traffic at all
  float b2[64];                   // conv2 bias, also via the constant bank
  const __half* w2s;              // stacked split weights * 2^k, canonical layout
  __half* act2;                   // [B][8 planes (chunks of 8 channels)][npix][8]
  float inv_scale;                // 2^-k
  int B, nabuf;
  Geom g;
  long long* trace;               // debug (WW_TC_TRACE=1): per-item role timestamps of CTA 0
};

// packed fp32x2 FMA (Blackwell): d = a * b + d on two lanes of a 64-bit register pair
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// SLOTS = ceil((256 + 2P + 2) / 256): halo'd pixels per producer thread (2 at the code preset, 3 at W = 161)
template <int NPASS, int SLOTS>
__global__ void __launch_bounds__(C12_THREADS, 1) conv12_kernel(const __grid_constant__ Conv12Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  unsigned char* w2s = smem;
  const uint32_t a_bytes = 4u * g.nsl2 * 16u;                  // one act1 buffer: 4 planes (chunks of 8 channels)
  unsigned char* a_buf0 = smem + W2_BYTES;
  const int NABUF = p.nabuf;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_buf0 + NABUF * a_bytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;      // [3]
  uint64_t* a_empty = bars + 4;     // [3]
  uint64_t* t_full = bars + 7;      // [2]
  uint64_t* t_empty = bars + 9;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int NL = 256 + 2 * g.P + 2;                            // pixels of an item incl. halo
  const int patch_rows = (NL + g.P - 1) / g.P + 3;
  const int patch_floats = (patch_rows * g.W + 3) & ~3;
  float* patch = reinterpret_cast<float*>(bars + 12);          // [2][patch_floats] log-mel rows of the current / next item

  // warp index through a shuffle: the compiler then knows it is warp-uniform, so role branches are uniform branches
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < NABUF_MAX; ++i) { mbar_init(a_full + i, 512); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 256); }
    fence_barrier_init();
  }
  if (warp == 24) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int items_per_clip = g.T2 >> 1;
  const int n_items = p.B * items_per_clip;

  if (warp < 16) {
    // ===================== conv1 producers
    const int ch0 = (tid >> 8) * 16;      // this thread's 16 of the 32 conv1 output channels
    auto stage_patch = [&](int item_, float* dst) {
      const int b_ = item_ / items_per_clip, tp_ = item_ - b_ * items_per_clip;
      const int pbase = 256 * tp_ - 1 - g.P - 1;
      const int r0 = (pbase >= 0 ? (int)__umulhi((uint32_t)pbase, g.magicP) : -1) - 2;   // image row of the first patch row
      const float* __restrict__ img = p.logmel + (size_t)b_ * g.H * g.W;
      for (int i = tid; i < patch_rows * g.W; i += 512) {
        const int pr = i / g.W;
        const int yy = r0 + pr;
        if (yy >= 0 && yy < g.H) {
          const uint32_t d = smem_u32(dst + i);
          const float* src = img + (size_t)yy * g.W + (i - pr * g.W);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if ((int)blockIdx.x < n_items) stage_patch(blockIdx.x, patch);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it % NABUF;
      mbar_wait_relaxed(a_empty + buf, ((it / NABUF) & 1) ^ 1, 10);
      if (warp == 0) C12_TRACE(0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 2, 512;" ::: "memory");                    // patch(it) visible; patch(it-1) no longer read
      if (item + (int)gridDim.x < n_items) stage_patch(item + gridDim.x, patch + ((it + 1) & 1) * patch_floats);
      const float* pt = patch + (it & 1) * patch_floats;
      const int tp = item % items_per_clip;
      const int pbase = 256 * tp - 1 - g.P - 1;
      const int r0 = (pbase >= 0 ? (int)__umulhi((uint32_t)pbase, g.magicP) : -1) - 2;
      unsigned char* ab = a_buf0 + buf * a_bytes;
#pragma unroll 1      // keep one copy of the conv1 body: the warp-specialised roles must share the instruction cache
      for (int u = 0; u < SLOTS; ++u) {
        const int l = (tid & 255) + u * 256;
        if (l < NL) {                                                    // warp-uniform except in one warp
          int y = 0, x = 0;
          const bool ok = pix_valid(pbase + l, g, y, x);
          float v[16];
          if (ok) {
            float in[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
              in[k] = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? pt[(yy - r0) * g.W + xx] : 0.0f;
            }
            // 16 channels as 8 packed fp32x2 accumulators: 72 FFMA2 instead of 144 FFMA
            unsigned long long acc[8];
            auto conv1 = [&](auto ch0c) {
              constexpr int CH0 = decltype(ch0c)::value;
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[c] = pack2(p.b1[CH0 + 2 * c], p.b1[CH0 + 2 * c + 1]);
#pragma unroll
              for (int k = 0; k < 9; ++k) {
                const unsigned long long in2 = pack2(in[k], in[k]);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  ffma2(acc[c], in2, pack2(p.w1[k * 32 + CH0 + 2 * c], p.w1[k * 32 + CH0 + 2 * c + 1]));
              }
            };
            if (ch0 == 0) conv1(std::integral_constant<int, 0>{});
            else conv1(std::integral_constant<int, 16>{});
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              unpack2(acc[c], v[2 * c], v[2 * c + 1]);
              v[2 * c] = fmaxf(v[2 * c], 0.0f);
              v[2 * c + 1] = fmaxf(v[2 * c + 1], 0.0f);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = 0.0f;
          }
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int kc = (ch0 >> 3) + k2;
            *reinterpret_cast<uint4*>(ab + ((size_t)kc * g.nsl2 + l) * 16) = cvt8(v + k2 * 8);
          }
        }
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full + buf);
      if (warp == 0) C12_TRACE(1);
    }
  } else if (warp >= 24) {
    // ===================== MMA issuers.  The whole warp runs the loop (descriptor math stays warp-uniform); the
    // tcgen05 instructions are guarded by elect.sync.
    if (warp == 24 && elect_one()) {
      mbar_arrive_expect_tx(w_full, W2_BYTES);
      bulk_g2s(w2s, p.w2s, W2_BYTES, w_full);
    }
    mbar_wait(w_full, 0, 20);
    constexpr uint32_t idesc128 = make_idesc(128, 128), idesc64 = make_idesc(128, 64);
    const uint64_t bdesc0 = make_desc(smem_u32(w2s), 2048, 128);
    const uint32_t lbo_a = (uint32_t)g.nsl2 * 16u;
    const uint64_t adesc0 = make_desc(smem_u32(a_buf0), lbo_a, 128);
    const uint32_t nsl = (uint32_t)g.nsl2;
    auto issue_items = [&](auto tbc) {
      constexpr int TB = decltype(tbc)::value;                           // accumulator buffer of this issuer
      const uint32_t d0 = tmem_base + TB * 256;
      int it = TB;
      for (int item = blockIdx.x + TB * gridDim.x; item < n_items; item += 2 * gridDim.x, it += 2) {
        const int buf = it % NABUF;
        mbar_wait(a_full + buf, (it / NABUF) & 1, 21);
        C12_TRACE(2);
        mbar_wait(t_empty + TB, ((it >> 1) & 1) ^ 1, 22);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)((buf * a_bytes) >> 4);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t row_off = (uint32_t)((g.P + 1) + (tap / 3 - 1) * g.P + (tap % 3 - 1));
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t bd = bdesc0 + (uint64_t)(((tap * 4 + 2 * j) * 2048) >> 4);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const uint64_t ad = adesc + (uint64_t)(2 * j * nsl + row_off + t * 128);
              const uint32_t d = d0 + t * 128;
              const uint32_t acc = (tap | j) != 0;
              // NPASS 2: a*W_hi -> cols 0..63 and a*W_lo -> cols 64..127 in one N = 128 instruction
              if (elect_one()) umma_f16(d, ad, bd, NPASS == 2 ? idesc128 : idesc64, acc);
            }
          }
        }
        if (elect_one()) {
          umma_commit(a_empty + buf);
          umma_commit(t_full + TB);
        }
        __syncwarp();
        C12_TRACE(3);
      }
    };
    if (warp == 24) issue_items(std::integral_constant<int, 0>{});
    else issue_items(std::integral_constant<int, 1>{});
  } else {
    // ===================== epilogue
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int hc = (warp - 16) >> 2;      // which 32 of the 64 output channels
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / items_per_clip, tp = item - b * items_per_clip;
      const int tb = it & 1;
      mbar_wait_relaxed(t_full + tb, (it >> 1) & 1, 30);
      if (warp == 16) C12_TRACE(4);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        float v[32];
        {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tb * 256 + t * 128 + hc * 32;
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(taddr, r0);
          if (NPASS == 2) tmem_ld32_nowait(taddr + 64, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) + (NPASS == 2 ? __uint_as_float(r1[i]) : 0.0f);
        }
        if (t == 1) {                       // both tiles are in registers: release the accumulator
          tc_fence_before();
          mbar_arrive(t_empty + tb);
          if (warp == 16) C12_TRACE(5);
        }
        const int s = 256 * tp + t * 128 + q * 32 + lane;
        int y, x;
        const bool ok = pix_valid(s - 1, g, y, x);
        uint4* dst = reinterpret_cast<uint4*>(p.act2) + (size_t)b * 8 * g.npix + s;
        const float inv_s = p.inv_scale;
        auto store_half = [&](auto hcc) {
          constexpr int HC = decltype(hcc)::value;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kc = HC * 4 + k4;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = ok ? fmaxf(fmaf(v[k4 * 8 + e], inv_s, p.b2[kc * 8 + e]), 0.0f) : 0.0f;
            dst[(size_t)kc * g.npix] = cvt8(o);
          }
        };
        if (hc == 0) store_half(std::integral_constant<int, 0>{});
        else store_half(std::integral_constant<int, 1>{});
      }
      if (warp == 16) C12_TRACE(6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 24) tmem_dealloc(tmem_base, 512);
}

size_t conv12_smem(const Geom& g, int nabuf) {
  const int NL = 256 + 2 * g.P + 2;
  const size_t patch_floats = (((NL + g.P - 1) / g.P + 3) * g.W + 3) & ~3;
  return (size_t)W2_BYTES + (size_t)nabuf * 4 * g.nsl2 * 16 + 16 * 8 + 2 * patch_floats * 4 + 64;
}

}  // namespace

// conv2 weights -> scaled fp16 hi/lo, stacked along N, UMMA canonical layout [tap][kc][n' = 64 hi + 64 lo][8]
int ww_conv12_tc_prepare(ww_ctx* c) {
  std::vector<float> w((size_t)64 * 32 * 9);       // [n][cin][tap]
  WW_CHECK(c, cudaMemcpy(w.data(), c->w["conv2.weight"], w.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<uint16_t> s((size_t)W2_BYTES / 2);
  const float sc = weight_scale(w);
  c->w2_inv_scale = 1.0f / sc;
  for (int tap = 0; tap < 9; ++tap)
    for (int kc = 0; kc < 4; ++kc)
      for (int n = 0; n < 64; ++n)
        for (int e = 0; e < 8; ++e) {
          const float v = w[((size_t)n * 32 + kc * 8 + e) * 9 + tap] * sc;
          const uint16_t hi = f2h(v), lo = f2h(v - h2f(hi));
          s[(((size_t)tap * 4 + kc) * 128 + n) * 8 + e] = hi;
          s[(((size_t)tap * 4 + kc) * 128 + 64 + n) * 8 + e] = lo;
        }
  if (!c->d_w2_split) WW_CHECK(c, cudaMalloc((void**)&c->d_w2_split, W2_BYTES));
  WW_CHECK(c, cudaMemcpy(c->d_w2_split, s.data(), W2_BYTES, cudaMemcpyHostToDevice));
  return WW_OK;
}

int ww_launch_conv12_tc(ww_ctx* c, const float* logmel, int B, const Geom& g, cudaStream_t st) {
  const int nabuf = conv12_smem(g, 3) <= 227 * 1024 ? 3 : 2;
  const size_t smem = conv12_smem(g, nabuf);
  const int slots = (256 + 2 * g.P + 2 + 255) / 256;
  if (smem > 227 * 1024 || slots > 3) {
    c->set_error("conv12_tc: frame count too large for the shared-memory tiles (use WW_CONV_FP32)");
    return WW_ERR_INVALID;
  }
  static size_t conf = 0;
  if (smem > conf) {
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  Conv12Params p;
  p.logmel = logmel; p.w2s = c->d_w2_split;
  memcpy(p.w1, c->h_w1t.data(), sizeof(p.w1));
  memcpy(p.b1, c->h_b1.data(), sizeof(p.b1));
  memcpy(p.b2, c->h_b2.data(), sizeof(p.b2));
  p.act2 = c->ws_act2_h; p.inv_scale = c->w2_inv_scale; p.B = B; p.nabuf = nabuf; p.g = g;
  static long long* d_trace = nullptr;
  const bool tracing = getenv("WW_TC_TRACE") != nullptr;
  if (tracing && !d_trace) { cudaMalloc((void**)&d_trace, 48 * 8 * 8); }
  if (tracing) cudaMemset(d_trace, 0, 48 * 8 * 8);
  p.trace = tracing ? d_trace : nullptr;
  const int grid = std::min(c->sm_count, B * (g.T2 / 2));
  ProfScope prof(c, WW_STAGE_CONV12, st);
  const bool fast = c->cfg.conv_mode == WW_CONV_FP16;
  if (slots <= 2) {
    if (fast) conv12_kernel<1, 2><<<grid, C12_THREADS, smem, st>>>(p);
    else conv12_kernel<2, 2><<<grid, C12_THREADS, smem, st>>>(p);
  } else {
    if (fast) conv12_kernel<1, 3><<<grid, C12_THREADS, smem, st>>>(p);
    else conv12_kernel<2, 3><<<grid, C12_THREADS, smem, st>>>(p);
  }
  WW_LAUNCH_CHECK(c);
  if (tracing) {
    long long h[48 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "conv12 trace (cycles rel. to item 0 producer start): prod_start prod_end mma_start mma_issued epi_start epi_tmem_free epi_end\n");
    for (int i = 30; i < 48; ++i) {
      fprintf(stderr, "item %2d:", i);
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %8lld", h[i * 8 + k] ? h[i * 8 + k] - h[0] : -1);
      fprintf(stderr, "\n");
    }
  }
  return WW_OK;
}

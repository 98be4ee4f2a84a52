// K3a (tensor-core path): conv1 (CUDA cores, fp32) fused into conv2 (tcgen05 implicit GEMM) + bias + ReLU  (sm_100a)
//
// Replaces F.relu(self.conv1(x)) and F.relu(self.conv2(x)) of WakewordModel.forward
// (/root/reference/wakeword_training_script.py:170-171).  See conv3_tc.cu for the layout story.
//
// Work item = (clip, 128-pixel tile of the pixel-linear padded image).  Warp roles:
//   warps 0-15 producers (one pixel x 16 channels per thread): conv1 + ReLU in fp32 for the tile and its 3x3 halo
//              (128 + 2P + 2 pixels), split to
//              bf16 hi/lo and stored as the K-major SWIZZLE_NONE A operand [chunk of 8 ch][pixel][16 B]
//              (double buffered, overlaps the MMAs of the previous tile);
//   warps 24,25 MMA issuers (warp 24: even tiles / accumulator 0, warp 25: odd tiles / accumulator 1, so the
//              ~100-cycle issue cost of each small MMA overlaps; one elected thread each): per 3x3 tap and 16-channel k-slice
//                 D[:, 0:128] += A_hi x [W_hi ; W_lo]^T   (N = 128: hi*hi and hi*lo in one instruction)
//                 D[:, 0:64 ] += A_lo x  W_hi^T           (N = 64)
//              -- the tap is only a start-address offset of the same shared-memory tile;
//   warps 16-23 epilogue (two warps per TMEM lane quadrant, 32 output channels each): TMEM -> D1 + D2 + bias, ReLU, zero the padding pixels, split hi/lo, write the conv3
//              operand planes to HBM (each warp store is 512 contiguous bytes).
// conv2 weights (hi and lo stacked along N, 73,728 B) stay resident in shared memory.
#include "tc_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <type_traits>

using namespace tc;

#define C12_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && it < 48 && lane == 0) p.trace[it * 8 + (slot)] = clock64(); } while (0)

namespace {

constexpr int C12_THREADS = 832;   // warps 0-15 producers, warps 16-23 epilogue, warps 24-25 MMA issuers (even / odd tiles)
constexpr int W2_BYTES = 9 * 4 * 128 * 16;   // [tap][kc 4][n' 128 = 64 hi + 64 lo][8 bf16]

struct Conv12Params {
  const float* logmel;            // [B][H][W]
  float w1[288];                  // conv1 weights [tap][cout] -- kernel parameters live in the constant bank, so
  float b1[32];                   // the FMAs read them as c[0][..] operands: no shared-memory traffic at all
  const __nv_bfloat16* w2s;       // stacked split weights, canonical layout
  float b2[64];                   // conv2 bias, also via the constant bank
  __nv_bfloat16* act2;            // [B][16 planes = chunk*2 + hl][npix][8]
  int B;
  Geom g;
  long long* trace;               // debug (WW_TC_TRACE=1): per-tile role timestamps of CTA 0
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

// SLOTS = ceil((128 + 2P + 2) / 256): halo'd pixels per producer thread (1 at the code preset, 2 at W = 161)
template <int NPASS, int SLOTS>
__global__ void __launch_bounds__(C12_THREADS, 1) conv12_kernel(const __grid_constant__ Conv12Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  unsigned char* w2s = smem;
  const uint32_t a_bytes = 8u * g.nsl2 * 16u;                  // one act1 buffer: 8 planes (kc*2 + hl)
  unsigned char* a_buf0 = smem + W2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_buf0 + 2 * a_bytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;      // [2]
  uint64_t* a_empty = bars + 3;     // [2]
  uint64_t* t_full = bars + 5;      // [2]
  uint64_t* t_empty = bars + 7;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int patch_floats = (((128 + 2 * g.P + 2 + g.P - 1) / g.P + 3) * g.W + 3) & ~3;
  float* patch = reinterpret_cast<float*>(bars + 10);        // [2][patch_floats] log-mel rows of the current / next tile

  // warp index through a shuffle: the compiler then knows it is warp-uniform, so role branches are uniform
  // branches and the MMA issue loop runs on the uniform datapath (no per-MMA R2UR waterfall)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 512);
      mbar_init(a_empty + i, 1);
      mbar_init(t_full + i, 1);
      mbar_init(t_empty + i, 256);
    }
    fence_barrier_init();
  }
  if (warp == 24) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int n_items = p.B * g.T2;
  const int NL = 128 + 2 * g.P + 2;

  if (warp < 16) {
    // ===================== conv1 producers
    // Software pipeline: the 3x3 input patches of the NEXT tile are loaded (global, L2 latency) before the
    // current tile is computed, so the loads overlap ~600 instructions of conv1 arithmetic.
    const int ch0 = (tid >> 8) * 16;      // this thread's 16 of the 32 conv1 output channels
    // The log-mel rows a tile needs (its pixels +- one image row) are contiguous in the [H][W] image: they are
    // staged into a double-buffered shared-memory patch with 4-byte cp.async one tile ahead, so the producers
    // never wait on HBM/L2 latency and no registers are spent on prefetching.
    const int patch_rows = (NL + g.P - 1) / g.P + 3;
    auto stage_patch = [&](int item_, float* dst) {
      const int b_ = item_ / g.T2, t2_ = item_ - b_ * g.T2;
      const int pbase = 128 * t2_ - 1 - g.P - 1;
      const int r0 = (pbase >= 0 ? (int)__umulhi((uint32_t)pbase, g.magicP) : -1) - 2;   // image row of the first patch row
      const float* __restrict__ img = p.logmel + (size_t)b_ * g.H * g.W;
      for (int i = tid; i < patch_rows * g.W; i += 512) {
        const int yy = r0 + i / g.W;
        if (yy >= 0 && yy < g.H) {
          const uint32_t d = smem_u32(dst + i);
          const float* src = img + (size_t)yy * g.W + (i - (i / g.W) * g.W);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if ((int)blockIdx.x < n_items) stage_patch(blockIdx.x, patch);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait_relaxed(a_empty + buf, ((it >> 1) & 1) ^ 1, 10);
      if (warp == 0) C12_TRACE(0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 2, 512;" ::: "memory");                    // patch(it) visible; patch(it-1) no longer read
      if (item + (int)gridDim.x < n_items) stage_patch(item + gridDim.x, patch + ((it + 1) & 1) * patch_floats);
      const float* pt = patch + (it & 1) * patch_floats;
      const int t2 = item % g.T2;
      const int pbase = 128 * t2 - 1 - g.P - 1;
      const int r0 = (pbase >= 0 ? (int)__umulhi((uint32_t)pbase, g.magicP) : -1) - 2;
      float in_c[SLOTS][9];
      bool ok_c[SLOTS];
#pragma unroll
      for (int u = 0; u < SLOTS; ++u) {
        const int l = (tid & 255) + u * 256;
        int y = 0, x = 0;
        ok_c[u] = (l < NL) && pix_valid(pbase + l, g, y, x);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
          in_c[u][k] = (ok_c[u] && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? pt[(yy - r0) * g.W + xx] : 0.0f;
        }
      }
      unsigned char* ab = a_buf0 + buf * a_bytes;
#pragma unroll
      for (int u = 0; u < SLOTS; ++u) {
        const int l = (tid & 255) + u * 256;
        if (l < NL) {
          float v[16];
          if (ok_c[u]) {
            // 16 channels as 8 packed fp32x2 accumulators: 72 FFMA2 instead of 144 FFMA; weights come from the
            // constant bank (kernel parameters) with compile-time offsets
            unsigned long long acc[8];
            auto conv1 = [&](auto ch0c) {
              constexpr int CH0 = decltype(ch0c)::value;
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[c] = pack2(p.b1[CH0 + 2 * c], p.b1[CH0 + 2 * c + 1]);
#pragma unroll
              for (int k = 0; k < 9; ++k) {
                const unsigned long long in2 = pack2(in_c[u][k], in_c[u][k]);
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
            uint4 hi, lo;
            split8(v + k2 * 8, hi, lo);
            *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 0) * g.nsl2 + l) * 16) = hi;
            *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 1) * g.nsl2 + l) * 16) = lo;
          }
        }
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full + buf);
      if (warp == 0) C12_TRACE(1);
    }
  } else if (warp >= 24) {
    // ===================== MMA issuers.  The whole warp runs the loop so that descriptors / addresses stay
    // warp-uniform (uniform registers feed UTCHMMA directly); only the elected lane issues.
    const bool leader = (lane == 0);
    if (leader && warp == 24) {
      mbar_arrive_expect_tx(w_full, W2_BYTES);
      bulk_g2s(w2s, p.w2s, W2_BYTES, w_full);
    }
    mbar_wait(w_full, 0, 20);
    constexpr uint32_t idesc128 = make_idesc(128, 128), idesc64 = make_idesc(128, 64);
    const uint64_t bdesc0 = make_desc(smem_u32(w2s), 2048, 128);
    const uint32_t lbo_a = 2u * g.nsl2 * 16u;
    const uint64_t adesc_b0 = make_desc(smem_u32(a_buf0), lbo_a, 128), adesc_b1 = make_desc(smem_u32(a_buf0 + a_bytes), lbo_a, 128);
    const uint32_t nsl = (uint32_t)g.nsl2;
    // One instantiation per issuer with a compile-time accumulator/buffer index: every descriptor is then a
    // function of kernel parameters and loop counters only, i.e. warp-uniform (uniform datapath, no R2UR).
    auto issue_tiles = [&](auto bufc) {
      constexpr int BUF = decltype(bufc)::value;
      const uint32_t d = tmem_base + BUF * 128;
      const uint64_t adesc = BUF ? adesc_b1 : adesc_b0;
      int it = BUF;
      for (int item = blockIdx.x + BUF * gridDim.x; item < n_items; item += 2 * gridDim.x, it += 2) {
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(a_full + BUF, par, 21);
        C12_TRACE(2);
        mbar_wait(t_empty + BUF, par ^ 1, 22);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t row_off = (uint32_t)((g.P + 1) + (tap / 3 - 1) * g.P + (tap % 3 - 1));
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t a_hi = adesc + (uint64_t)((4 * j + 0) * nsl + row_off);
            const uint64_t bd = bdesc0 + (uint64_t)(((tap * 4 + 2 * j) * 2048) >> 4);
            if (NPASS == 3) {
              const uint64_t a_lo = adesc + (uint64_t)((4 * j + 1) * nsl + row_off);
              if (elect_one()) {
                umma_bf16(d, a_hi, bd, idesc128, acc);   // hi*hi -> cols 0..63, hi*lo -> cols 64..127
                umma_bf16(d, a_lo, bd, idesc64, 1);      // lo*hi -> cols 0..63
              }
            } else {
              if (elect_one()) umma_bf16(d, a_hi, bd, idesc64, acc);
            }
            acc = 1;
          }
        }
        if (elect_one()) {
          umma_commit(a_empty + BUF);
          umma_commit(t_full + BUF);
        }
        __syncwarp();
        C12_TRACE(3);
      }
    };
    if (warp == 24) issue_tiles(std::integral_constant<int, 0>{});
    else issue_tiles(std::integral_constant<int, 1>{});
  } else {
    // ===================== epilogue
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int hc = (warp - 16) >> 2;      // which 32 of the 64 output channels
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / g.T2, t2 = item - b * g.T2;
      const int buf = it & 1;
      mbar_wait_relaxed(t_full + buf, (it >> 1) & 1, 30);
      if (warp == 16) C12_TRACE(4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + hc * 32;
      uint32_t r0[32], r1[32];
      float v[32];
      tmem_ld32_nowait(taddr, r0);
      if (NPASS == 3) tmem_ld32_nowait(taddr + 64, r1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) + (NPASS == 3 ? __uint_as_float(r1[i]) : 0.0f);
      tc_fence_before();
      mbar_arrive(t_empty + buf);
      if (warp == 16) C12_TRACE(5);
      const int s = 128 * t2 + q * 32 + lane;
      int y, x;
      const bool ok = pix_valid(s - 1, g, y, x);
      uint4* dst = reinterpret_cast<uint4*>(p.act2) + (size_t)b * 16 * g.npix + s;
      auto store_half = [&](auto hcc) {
        constexpr int HC = decltype(hcc)::value;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          constexpr int dummy = 0; (void)dummy;
          const int kc = HC * 4 + k4;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = ok ? fmaxf(v[k4 * 8 + e] + p.b2[kc * 8 + e], 0.0f) : 0.0f;
          uint4 hi, lo;
          split8(o, hi, lo);
          dst[(size_t)(kc * 2 + 0) * g.npix] = hi;
          dst[(size_t)(kc * 2 + 1) * g.npix] = lo;
        }
      };
      if (hc == 0) store_half(std::integral_constant<int, 0>{});
      else store_half(std::integral_constant<int, 1>{});
      if (warp == 16) C12_TRACE(6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 24) tmem_dealloc(tmem_base, 256);
}

size_t conv12_smem(const Geom& g) {
  const size_t patch_floats = (((128 + 2 * g.P + 2 + g.P - 1) / g.P + 3) * g.W + 3) & ~3;
  return (size_t)W2_BYTES + 2 * 8 * (size_t)g.nsl2 * 16 + 64 * 4 + 16 * 8 + 2 * patch_floats * 4 + 64;
}

}  // namespace

// conv2 weights -> bf16 hi/lo, stacked along N, UMMA canonical layout [tap][kc][n' = 64 hi + 64 lo][8]
int ww_conv12_tc_prepare(ww_ctx* c) {
  std::vector<float> w((size_t)64 * 32 * 9);       // [n][cin][tap]
  WW_CHECK(c, cudaMemcpy(w.data(), c->w["conv2.weight"], w.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<uint16_t> s((size_t)W2_BYTES / 2);
  for (int tap = 0; tap < 9; ++tap)
    for (int kc = 0; kc < 4; ++kc)
      for (int n = 0; n < 64; ++n)
        for (int e = 0; e < 8; ++e) {
          const float v = w[((size_t)n * 32 + kc * 8 + e) * 9 + tap];
          const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
          s[(((size_t)tap * 4 + kc) * 128 + n) * 8 + e] = hi;
          s[(((size_t)tap * 4 + kc) * 128 + 64 + n) * 8 + e] = lo;
        }
  if (!c->d_w2_split) WW_CHECK(c, cudaMalloc((void**)&c->d_w2_split, W2_BYTES));
  WW_CHECK(c, cudaMemcpy(c->d_w2_split, s.data(), W2_BYTES, cudaMemcpyHostToDevice));
  return WW_OK;
}

int ww_launch_conv12_tc(ww_ctx* c, const float* logmel, int B, const Geom& g, cudaStream_t st) {
  const size_t smem = conv12_smem(g);
  if (smem > 227 * 1024) {
    c->set_error("conv12_tc: frame count too large for the shared-memory tiles (use WW_CONV_FP32)");
    return WW_ERR_INVALID;
  }
  const int slots = (128 + 2 * g.P + 2 + 255) / 256;
  if (slots > 2) { c->set_error("conv12_tc: frame count too large (use WW_CONV_FP32)"); return WW_ERR_INVALID; }
  static size_t conf = 0;
  if (smem > conf) {
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  Conv12Params p;
  p.logmel = logmel; p.w2s = c->d_w2_split;
  memcpy(p.w1, c->h_w1t.data(), sizeof(p.w1));
  memcpy(p.b1, c->h_b1.data(), sizeof(p.b1));
  p.act2 = c->ws_act2_split;
  memcpy(p.b2, c->h_b2.data(), sizeof(p.b2)); p.B = B; p.g = g;
  static long long* d_trace = nullptr;
  const bool tracing = getenv("WW_TC_TRACE") != nullptr;
  if (tracing && !d_trace) { cudaMalloc((void**)&d_trace, 48 * 8 * 8); }
  if (tracing) cudaMemset(d_trace, 0, 48 * 8 * 8);
  p.trace = tracing ? d_trace : nullptr;
  const int grid = std::min(c->sm_count, B * g.T2);
  ProfScope prof(c, WW_STAGE_CONV12, st);
  const bool fast = c->cfg.conv_mode == WW_CONV_BF16;
  if (slots == 1) {
    if (fast) conv12_kernel<1, 1><<<grid, C12_THREADS, smem, st>>>(p);
    else conv12_kernel<3, 1><<<grid, C12_THREADS, smem, st>>>(p);
  } else {
    if (fast) conv12_kernel<1, 2><<<grid, C12_THREADS, smem, st>>>(p);
    else conv12_kernel<3, 2><<<grid, C12_THREADS, smem, st>>>(p);
  }
  WW_LAUNCH_CHECK(c);
  if (tracing) {
    long long h[48 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "conv12 trace (cycles rel. to tile 0 producer start): prod_start prod_end mma_start mma_issued epi_start epi_tmem_free epi_end\n");
    for (int i = 0; i < 24; ++i) {
      fprintf(stderr, "tile %2d:", i);
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %8lld", h[i * 8 + k] ? h[i * 8 + k] - h[0] : -1);
      fprintf(stderr, "\n");
    }
  }
  return WW_OK;
}

// K3a (tensor-core path): conv1 (CUDA cores, fp32) fused into conv2 (tcgen05 implicit GEMM) + bias + ReLU  (sm_100a)
//
// Replaces F.relu(self.conv1(x)) and F.relu(self.conv2(x)) of WakewordModel.forward
// (/root/reference/wakeword_training_script.py:170-171).  See conv3_tc.cu for the layout story.
//
// Work item = (clip, 128-pixel tile of the pixel-linear padded image).  Warp roles:
//   warps 0-7  producers: conv1 + ReLU in fp32 for the tile and its 3x3 halo (128 + 2P + 2 pixels), split to
//              bf16 hi/lo and stored as the K-major SWIZZLE_NONE A operand [chunk of 8 ch][pixel][16 B]
//              (double buffered, overlaps the MMAs of the previous tile);
//   warp 8     one thread issues the MMAs: per 3x3 tap and 16-channel k-slice
//                 D[:, 0:128] += A_hi x [W_hi ; W_lo]^T   (N = 128: hi*hi and hi*lo in one instruction)
//                 D[:, 0:64 ] += A_lo x  W_hi^T           (N = 64)
//              -- the tap is only a start-address offset of the same shared-memory tile;
//   warps 9-16 epilogue (two warps per TMEM lane quadrant, 32 output channels each): TMEM -> D1 + D2 + bias, ReLU, zero the padding pixels, split hi/lo, write the conv3
//              operand planes to HBM (each warp store is 512 contiguous bytes).
// conv2 weights (hi and lo stacked along N, 73,728 B) stay resident in shared memory.
#include "tc_common.cuh"

#include <algorithm>

using namespace tc;

namespace {

constexpr int C12_THREADS = 544;   // warps 0-7 producers, warp 8 MMA issuer, warps 9-16 epilogue
constexpr int W2_BYTES = 9 * 4 * 128 * 16;   // [tap][kc 4][n' 128 = 64 hi + 64 lo][8 bf16]

struct Conv12Params {
  const float* logmel;            // [B][H][W]
  const float* w1t;               // [9][32]
  const float* b1;                // [32]
  const __nv_bfloat16* w2s;       // stacked split weights, canonical layout
  const float* b2;                // [64]
  __nv_bfloat16* act2;            // [B][16 planes = chunk*2 + hl][npix][8]
  int B;
  Geom g;
};

template <int NPASS>
__global__ void __launch_bounds__(C12_THREADS, 1) conv12_kernel(Conv12Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  unsigned char* w2s = smem;
  const uint32_t a_bytes = 8u * g.nsl2 * 16u;                  // one act1 buffer: 8 planes (kc*2 + hl)
  unsigned char* a_buf0 = smem + W2_BYTES;
  float* w1s = reinterpret_cast<float*>(a_buf0 + 2 * a_bytes); // [9][32]
  float* b1s = w1s + 288;
  float* b2s = b1s + 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b2s + 64);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;      // [2]
  uint64_t* a_empty = bars + 3;     // [2]
  uint64_t* t_full = bars + 5;      // [2]
  uint64_t* t_empty = bars + 7;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 288; i += C12_THREADS) w1s[i] = p.w1t[i];
  if (tid < 32) b1s[tid] = p.b1[tid];
  if (tid < 64) b2s[tid] = p.b2[tid];
  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 256);
      mbar_init(a_empty + i, 1);
      mbar_init(t_full + i, 1);
      mbar_init(t_empty + i, 256);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.B * g.T2;
  const int NL = 128 + 2 * g.P + 2;

  if (warp < 8) {
    // ===================== conv1 producers
    // Software pipeline: the 3x3 input patches of the NEXT tile are loaded (global, L2 latency) before the
    // current tile is computed, so the loads overlap ~600 instructions of conv1 arithmetic.
    float in_n[2][9];
    bool ok_n[2] = {false, false};
    auto load_in = [&](int item_) {
      const int b_ = item_ / g.T2, t2_ = item_ - b_ * g.T2;
      const float* __restrict__ img = p.logmel + (size_t)b_ * g.H * g.W;
      const int pbase = 128 * t2_ - 1 - g.P - 1;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int l = tid + u * 256;
        int y = 0, x = 0;
        ok_n[u] = (l < NL) && pix_valid(pbase + l, g, y, x);
        if (ok_n[u]) {
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
            in_n[u][k] = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? __ldg(img + yy * g.W + xx) : 0.0f;
          }
        }
      }
    };
    if ((int)blockIdx.x < n_items) load_in(blockIdx.x);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      float in_c[2][9];
      bool ok_c[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ok_c[u] = ok_n[u];
#pragma unroll
        for (int k = 0; k < 9; ++k) in_c[u][k] = in_n[u][k];
      }
      if (item + (int)gridDim.x < n_items) load_in(item + gridDim.x);
      mbar_wait(a_empty + buf, ((it >> 1) & 1) ^ 1, 10);
      unsigned char* ab = a_buf0 + buf * a_bytes;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int l = tid + u * 256;
        if (l >= NL) break;
        float v[32];
        if (ok_c[u]) {
          const float* in = in_c[u];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = b1s[c];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 w = *reinterpret_cast<const float4*>(w1s + k * 32 + c4 * 4);
              v[c4 * 4 + 0] = fmaf(in[k], w.x, v[c4 * 4 + 0]);
              v[c4 * 4 + 1] = fmaf(in[k], w.y, v[c4 * 4 + 1]);
              v[c4 * 4 + 2] = fmaf(in[k], w.z, v[c4 * 4 + 2]);
              v[c4 * 4 + 3] = fmaf(in[k], w.w, v[c4 * 4 + 3]);
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.0f);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0.0f;
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4 hi, lo;
          split8(v + kc * 8, hi, lo);
          *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 0) * g.nsl2 + l) * 16) = hi;
          *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 1) * g.nsl2 + l) * 16) = lo;
        }
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full + buf);
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread)
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, W2_BYTES);
      bulk_g2s(w2s, p.w2s, W2_BYTES, w_full);
      mbar_wait(w_full, 0, 20);
      constexpr uint32_t idesc128 = make_idesc(128, 128), idesc64 = make_idesc(128, 64);
      const uint64_t bdesc0 = make_desc(smem_u32(w2s), 2048, 128);
      const uint32_t lbo_a = 2u * g.nsl2 * 16u;
      const uint64_t adesc0[2] = {make_desc(smem_u32(a_buf0), lbo_a, 128), make_desc(smem_u32(a_buf0 + a_bytes), lbo_a, 128)};
      const uint32_t nsl = (uint32_t)g.nsl2;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(a_full + buf, par, 21);
        mbar_wait(t_empty + buf, par ^ 1, 22);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * 128;
        uint32_t acc = 0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t row_off = (uint32_t)((g.P + 1) + (tap / 3 - 1) * g.P + (tap % 3 - 1));
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t a_hi = adesc0[buf] + (uint64_t)((4 * j + 0) * nsl + row_off);
            const uint64_t bd = bdesc0 + (uint64_t)(((tap * 4 + 2 * j) * 2048) >> 4);
            if (NPASS == 3) {
              const uint64_t a_lo = adesc0[buf] + (uint64_t)((4 * j + 1) * nsl + row_off);
              umma_bf16(d, a_hi, bd, idesc128, acc);   // hi*hi -> cols 0..63, hi*lo -> cols 64..127
              umma_bf16(d, a_lo, bd, idesc64, 1);      // lo*hi -> cols 0..63
            } else {
              umma_bf16(d, a_hi, bd, idesc64, acc);
            }
            acc = 1;
          }
        }
        umma_commit(a_empty + buf);
        umma_commit(t_full + buf);
      }
    }
  } else {
    // ===================== epilogue
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int hc = (warp - 9) >> 2;       // which 32 of the 64 output channels
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / g.T2, t2 = item - b * g.T2;
      const int buf = it & 1;
      mbar_wait(t_full + buf, (it >> 1) & 1, 30);
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
      const int s = 128 * t2 + q * 32 + lane;
      int y, x;
      const bool ok = pix_valid(s - 1, g, y, x);
      uint4* dst = reinterpret_cast<uint4*>(p.act2) + (size_t)b * 16 * g.npix + s;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int kc = hc * 4 + k4;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ok ? fmaxf(v[k4 * 8 + e] + b2s[kc * 8 + e], 0.0f) : 0.0f;
        uint4 hi, lo;
        split8(o, hi, lo);
        dst[(size_t)(kc * 2 + 0) * g.npix] = hi;
        dst[(size_t)(kc * 2 + 1) * g.npix] = lo;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
}

size_t conv12_smem(const Geom& g) { return (size_t)W2_BYTES + 2 * 8 * (size_t)g.nsl2 * 16 + (288 + 32 + 64) * 4 + 16 * 8 + 64; }

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
  static size_t conf = 0;
  if (smem > conf) {
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  Conv12Params p;
  p.logmel = logmel; p.w1t = c->d_convw_t[0]; p.b1 = c->w["conv1.bias"]; p.w2s = c->d_w2_split;
  p.b2 = c->w["conv2.bias"]; p.act2 = c->ws_act2_split; p.B = B; p.g = g;
  const int grid = std::min(c->sm_count, B * g.T2);
  ProfScope prof(c, WW_STAGE_CONV12, st);
  if (c->cfg.conv_mode == WW_CONV_BF16) conv12_kernel<1><<<grid, C12_THREADS, smem, st>>>(p);
  else conv12_kernel<3><<<grid, C12_THREADS, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

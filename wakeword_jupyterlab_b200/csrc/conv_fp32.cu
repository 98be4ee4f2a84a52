// K3 (exact-arithmetic path): 3x3 convolution + bias + ReLU on fp32 CUDA cores  (sm_100a)
//
// Replaces the cuDNN calls behind WakewordModel.forward's conv stack
// (/root/reference/wakeword_training_script.py:170-173): conv1 1->32, conv2 32->64, conv3 64->128,
// each 3x3 / stride 1 / zero padding 1 + ReLU, then AdaptiveAvgPool2d((1,1)).
// This is the WW_CONV_FP32 mode: true fp32 FMA arithmetic, used as the on-device cross-check of the
// tcgen05 split-precision kernels (conv_tc.cu) and for configurations those do not cover.
//
// Mapping: CTA = (clip, 8x32 output tile, COUT_TILE output channels); thread = one output pixel with
// COUT_TILE accumulators in registers; input tile (with halo) and the weight slab for CC input
// channels are staged in shared memory; weights are read as float4 broadcasts.
// The last layer never materialises its activation: bias+ReLU+sum over the tile is reduced in the
// CTA and written to a per-tile partial (deterministic; no atomics), which the head kernel sums.
#include "ctx.cuh"

namespace {

constexpr int TH = 8, TW = 32, kThreads = TH * TW;

// Packed fp32x2 FMA (Blackwell): two IEEE fma.rn per instruction on a 64-bit register pair.  A 3-register FFMA issues
// every other cycle per scheduler; FFMA2 is what reaches the fp32 peak.  Each lane is an ordinary fma.rn, so results
// are bit-identical to the scalar form.
typedef unsigned long long f32x2;
__device__ __forceinline__ void ffma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// EPI: 0 = bias + ReLU -> [B][COUT][H][W];  1 = bias + ReLU + sum over the tile -> [B][n_tiles][COUT] (POOL);
//      2 = data gradient (training): no bias, out = aux > 0 ? acc : 0 with aux = the forward activation of the layer
//          below (ReLU derivative), `wt` = the flipped / transposed weights.
constexpr int EPI_RELU = 0, EPI_POOL = 1, EPI_DGRAD = 2;

template <int CIN, int COUT, int COUT_TILE, int CC, int EPI>
__global__ void __launch_bounds__(kThreads)
conv3x3_relu_kernel(const float* __restrict__ in,     // [B][CIN][H][W]
                    const float* __restrict__ wt,     // [CIN][9][COUT]
                    const float* __restrict__ bias,   // [COUT]   (EPI_DGRAD: aux [B][COUT][H][W])
                    float* __restrict__ out,          // POOL ? [B][n_tiles][COUT] : [B][COUT][H][W]
                    int H, int W, int tiles_x, int n_tiles) {
  constexpr bool POOL = EPI == EPI_POOL;
  __shared__ float s_in[CC][TH + 2][TW + 2];
  __shared__ __align__(16) float s_w[CC][9][COUT_TILE];
  __shared__ float s_red[kThreads / 32][COUT_TILE];

  const int tid = threadIdx.x;
  const int tx = tid % TW, ty = tid / TW;
  const int tile = blockIdx.x;
  const int x0 = (tile % tiles_x) * TW, y0 = (tile / tiles_x) * TH;
  const int co0 = blockIdx.y * COUT_TILE;
  const int b = blockIdx.z;
  const float* inb = in + (size_t)b * CIN * H * W;

  f32x2 acc2[COUT_TILE / 2];
#pragma unroll
  for (int i = 0; i < COUT_TILE / 2; ++i) acc2[i] = 0ull;

  for (int c0 = 0; c0 < CIN; c0 += CC) {
    __syncthreads();
    for (int i = tid; i < CC * (TH + 2) * (TW + 2); i += kThreads) {
      const int c = i / ((TH + 2) * (TW + 2));
      const int r = i % ((TH + 2) * (TW + 2));
      const int yy = y0 + r / (TW + 2) - 1, xx = x0 + r % (TW + 2) - 1;
      float v = 0.0f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(inb + ((size_t)(c0 + c) * H + yy) * W + xx);
      s_in[c][r / (TW + 2)][r % (TW + 2)] = v;
    }
    for (int i = tid; i < CC * 9 * COUT_TILE; i += kThreads) {
      const int c = i / (9 * COUT_TILE);
      const int r = i % (9 * COUT_TILE);
      s_w[c][r / COUT_TILE][r % COUT_TILE] = __ldg(wt + ((size_t)(c0 + c) * 9 + r / COUT_TILE) * COUT + co0 + r % COUT_TILE);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CC; ++c) {
      float v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = s_in[c][ty + k / 3][tx + k % 3];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const f32x2 vv = pack2(v[k], v[k]);
#pragma unroll
        for (int co = 0; co < COUT_TILE; co += 4) {
          const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(&s_w[c][k][co]);    // (w0, w1), (w2, w3)
          ffma2(acc2[co / 2], vv, w4.x);
          ffma2(acc2[co / 2 + 1], vv, w4.y);
        }
      }
    }
  }
  float acc[COUT_TILE];
#pragma unroll
  for (int i = 0; i < COUT_TILE / 2; ++i) unpack2(acc2[i], acc[2 * i], acc[2 * i + 1]);

  const int y = y0 + ty, x = x0 + tx;
  const bool valid = (y < H) && (x < W);
  if (EPI == EPI_DGRAD) {
    if (valid) {
      const size_t o0 = ((size_t)b * COUT + co0) * H * W + (size_t)y * W + x;
#pragma unroll
      for (int co = 0; co < COUT_TILE; ++co)
        out[o0 + (size_t)co * H * W] = __ldg(bias + o0 + (size_t)co * H * W) > 0.0f ? acc[co] : 0.0f;
    }
  } else if (!POOL) {
    if (valid) {
      float* ob = out + ((size_t)b * COUT + co0) * H * W + (size_t)y * W + x;
#pragma unroll
      for (int co = 0; co < COUT_TILE; ++co) ob[(size_t)co * H * W] = fmaxf(acc[co] + __ldg(bias + co0 + co), 0.0f);
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int co = 0; co < COUT_TILE; ++co) {
      float v = valid ? fmaxf(acc[co] + __ldg(bias + co0 + co), 0.0f) : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][co] = v;
    }
    __syncthreads();
    if (tid < COUT_TILE) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += s_red[w][tid];
      out[((size_t)b * n_tiles + tile) * COUT + co0 + tid] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient (training step, WakewordTrainer.train_epoch, /root/reference/wakeword_training_script.py:247-257):
//   dW[co][ci][k] = sum_b sum_{y,x} dY[b][co][y][x] * X[b][ci][y + ky - 1][x + kx - 1]
// CTA = (32 output channels, 32 input channels, one slice of the batch); thread = 4 co x 1 ci x 9 taps (36 fp32
// accumulators).  Per 8x32 pixel tile the dY tile ([pixel][32 co]) and the X tile with halo are staged in shared
// memory; a thread walks each row with a sliding 3x3 window (3 new X loads + one 128-bit dY load per 36 FMAs).
// Partial sums per batch slice are written to part[slice][COUT][CIN][9] and reduced in fixed order by
// reduce_slices_kernel (deterministic, no atomics).
constexpr int WG_CO = 32, WG_CI = 32, WG_DY_PITCH = WG_CO + 4;
constexpr size_t WG_SMEM = (size_t)(TH * TW * WG_DY_PITCH + WG_CI * (TH + 2) * (TW + 2)) * sizeof(float);

__global__ void __launch_bounds__(kThreads)
conv3x3_wgrad_kernel(const float* __restrict__ x,      // [B][CIN][H][W]
                     const float* __restrict__ dy,     // [B][COUT][H][W]
                     float* __restrict__ part,         // [slices][COUT][CIN][9]
                     int B, int CIN, int COUT, int H, int W, int clips_per_slice) {
  extern __shared__ __align__(16) float wg_smem[];
  float (*s_dy)[WG_DY_PITCH] = reinterpret_cast<float (*)[WG_DY_PITCH]>(wg_smem);                      // [256 px][36]
  float (*s_x)[TH + 2][TW + 2] = reinterpret_cast<float (*)[TH + 2][TW + 2]>(wg_smem + TH * TW * WG_DY_PITCH);   // [32 ci][10][34]
  const int tid = threadIdx.x;
  const int cg = tid & 7, ci_l = tid >> 3;                        // 4 co per thread, one ci
  const int co0 = blockIdx.x * WG_CO, ci0 = blockIdx.y * WG_CI, slice = blockIdx.z;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  f32x2 acc2[2][9];                                                 // (co 0, co 1) and (co 2, co 3) per tap
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc2[a][k] = 0ull;
  const int b_lo = slice * clips_per_slice, b_hi = min(B, b_lo + clips_per_slice);
  for (int b = b_lo; b < b_hi; ++b) {
    for (int tile = 0; tile < tiles_x * tiles_y; ++tile) {
      const int x0 = (tile % tiles_x) * TW, y0 = (tile / tiles_x) * TH;
      __syncthreads();
      for (int i = tid; i < WG_CO * TH * TW; i += kThreads) {
        const int co = i / (TH * TW), pix = i % (TH * TW);
        const int yy = y0 + pix / TW, xx = x0 + pix % TW;
        float v = 0.0f;
        if (co0 + co < COUT && yy < H && xx < W) v = __ldg(dy + (((size_t)b * COUT + co0 + co) * H + yy) * W + xx);
        s_dy[pix][co] = v;
      }
      for (int i = tid; i < WG_CI * (TH + 2) * (TW + 2); i += kThreads) {
        const int c = i / ((TH + 2) * (TW + 2)), r = i % ((TH + 2) * (TW + 2));
        const int yy = y0 + r / (TW + 2) - 1, xx = x0 + r % (TW + 2) - 1;
        float v = 0.0f;
        if (ci0 + c < CIN && yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(x + (((size_t)b * CIN + ci0 + c) * H + yy) * W + xx);
        s_x[c][r / (TW + 2)][r % (TW + 2)] = v;
      }
      __syncthreads();
#pragma unroll 1
      for (int ty = 0; ty < TH; ++ty) {
        f32x2 w0[3], w1[3], w2[3];                                   // sliding 3x3 window (each value twice): columns tx-1, tx, tx+1
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float a0 = s_x[ci_l][ty + r][0], a1 = s_x[ci_l][ty + r][1];
          w0[r] = pack2(a0, a0); w1[r] = pack2(a1, a1);
        }
#pragma unroll 4
        for (int tx = 0; tx < TW; ++tx) {
#pragma unroll
          for (int r = 0; r < 3; ++r) { const float a2 = s_x[ci_l][ty + r][tx + 2]; w2[r] = pack2(a2, a2); }
          const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(&s_dy[ty * TW + tx][cg * 4]);   // (dy0, dy1), (dy2, dy3)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            ffma2(acc2[0][r * 3 + 0], d4.x, w0[r]); ffma2(acc2[1][r * 3 + 0], d4.y, w0[r]);
            ffma2(acc2[0][r * 3 + 1], d4.x, w1[r]); ffma2(acc2[1][r * 3 + 1], d4.y, w1[r]);
            ffma2(acc2[0][r * 3 + 2], d4.x, w2[r]); ffma2(acc2[1][r * 3 + 2], d4.y, w2[r]);
          }
#pragma unroll
          for (int r = 0; r < 3; ++r) { w0[r] = w1[r]; w1[r] = w2[r]; }
        }
      }
    }
  }
  float acc[4][9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { unpack2(acc2[0][k], acc[0][k], acc[1][k]); unpack2(acc2[1][k], acc[2][k], acc[3][k]); }
  const int ci = ci0 + ci_l;
  if (ci < CIN) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int co = co0 + cg * 4 + a;
      if (co < COUT) {
        float* o = part + (((size_t)slice * COUT + co) * CIN + ci) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) o[k] = acc[a][k];
      }
    }
  }
}

// out[i] = sum_s part[s][i]  (fixed order)
__global__ void reduce_slices_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int slices) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int k = 0; k < slices; ++k) s += part[(size_t)k * n + i];
  out[i] = s;
}

// bias gradient: db[co] = sum_b sum_p dY[b][co][p]; one CTA per (co, slice), then reduce_slices_kernel
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ part, int B,
                                                        int COUT, int HW, int clips_per_slice) {
  __shared__ float red[8];
  const int co = blockIdx.x, slice = blockIdx.y;
  const int b_lo = slice * clips_per_slice, b_hi = min(B, b_lo + clips_per_slice);
  float s = 0.0f;
  for (int b = b_lo; b < b_hi; ++b) {
    const float* p = dy + ((size_t)b * COUT + co) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) s += p[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    part[(size_t)slice * COUT + co] = t;
  }
}

// pooled[b][c] = mean_p act3[b][c][p]  (one warp per (b, c))
__global__ void __launch_bounds__(256) pool_mean_kernel(const float* __restrict__ act, float* __restrict__ pooled, int n_rows, int HW) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* p = act + (size_t)row * HW;
  float s = 0.0f;
  for (int i = lane; i < HW; i += 32) s += p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) pooled[row] = s / (float)HW;
}

// in place: act3 -> dY3 = (act3 > 0) * dpooled[b][c] / HW   (gradient of mean-pool after ReLU)
__global__ void __launch_bounds__(256) pool_relu_grad_kernel(float* __restrict__ act, const float* __restrict__ dpooled, int n_rows, int HW) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float* p = act + (size_t)row * HW;
  const float g = dpooled[row] / (float)HW;
  for (int i = lane; i < HW; i += 32) p[i] = p[i] > 0.0f ? g : 0.0f;
}

}  // namespace

// ---- training-step launchers (fp32 CUDA-core path; see train.cu) -------------------------------------------------
// forward with every activation kept: x [B][1][H][W] -> act1 [B][32][H][W], act2 [B][64][H][W], act3 [B][128][H][W]
int ww_train_conv_forward(ww_ctx* c, const float* x, int B, float* act1, float* act2, float* act3, float* pooled,
                          cudaStream_t st) {
  const int H = c->cfg.n_mels, W = c->W;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, n_tiles = tiles_x * tiles_y;
  dim3 block(kThreads);
  conv3x3_relu_kernel<1, 32, 32, 1, EPI_RELU><<<dim3(n_tiles, 1, B), block, 0, st>>>(
      x, c->d_convw_t[0], c->w["conv1.bias"], act1, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  conv3x3_relu_kernel<32, 64, 32, 8, EPI_RELU><<<dim3(n_tiles, 2, B), block, 0, st>>>(
      act1, c->d_convw_t[1], c->w["conv2.bias"], act2, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  conv3x3_relu_kernel<64, 128, 32, 8, EPI_RELU><<<dim3(n_tiles, 4, B), block, 0, st>>>(
      act2, c->d_convw_t[2], c->w["conv3.bias"], act3, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  pool_mean_kernel<<<(B * 128 * 32 + 255) / 256, 256, 0, st>>>(act3, pooled, B * 128, H * W);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

// backward of the conv stack.  act3 is overwritten with dY3, dact2 / dact1 receive dY2 / dY1 (already masked by the
// ReLU derivative); gradients land in gw{1,2,3} ([Cout][Cin][3][3]) and gb{1,2,3}.  wflip{2,3}: [Cout][9][Cin] weights
// for the data gradients (flipped taps, transposed channels); part: scratch for the sliced partial sums.
int ww_train_conv_backward(ww_ctx* c, const float* x, int B, const float* act1, const float* act2, float* act3,
                           const float* dpooled, float* dact2, float* dact1, const float* wflip3, const float* wflip2,
                           float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3, float* part,
                           int slices, cudaStream_t st) {
  const int H = c->cfg.n_mels, W = c->W, HW = H * W;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, n_tiles = tiles_x * tiles_y;
  const int cps = (B + slices - 1) / slices;
  dim3 block(kThreads);
  // per device, not per process: set on every call (cheap)
  WW_CHECK(c, cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
  auto wgrad = [&](const float* xin, const float* dy, int CIN, int COUT, float* gw, float* gb) -> int {
    conv3x3_wgrad_kernel<<<dim3((COUT + WG_CO - 1) / WG_CO, (CIN + WG_CI - 1) / WG_CI, slices), block, WG_SMEM, st>>>(
        xin, dy, part, B, CIN, COUT, H, W, cps);
    WW_LAUNCH_CHECK(c);
    const int n = COUT * CIN * 9;
    reduce_slices_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, gw, n, slices);
    WW_LAUNCH_CHECK(c);
    bias_grad_kernel<<<dim3(COUT, slices), 256, 0, st>>>(dy, part, B, COUT, HW, cps);
    WW_LAUNCH_CHECK(c);
    reduce_slices_kernel<<<(COUT + 255) / 256, 256, 0, st>>>(part, gb, COUT, slices);
    WW_LAUNCH_CHECK(c);
    return WW_OK;
  };
  pool_relu_grad_kernel<<<(B * 128 * 32 + 255) / 256, 256, 0, st>>>(act3, dpooled, B * 128, HW);
  WW_LAUNCH_CHECK(c);
  int rc;
  if ((rc = wgrad(act2, act3, 64, 128, gw3, gb3))) return rc;
  conv3x3_relu_kernel<128, 64, 32, 8, EPI_DGRAD><<<dim3(n_tiles, 2, B), block, 0, st>>>(
      act3, wflip3, act2, dact2, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  if ((rc = wgrad(act1, dact2, 32, 64, gw2, gb2))) return rc;
  conv3x3_relu_kernel<64, 32, 32, 8, EPI_DGRAD><<<dim3(n_tiles, 1, B), block, 0, st>>>(
      dact2, wflip2, act1, dact1, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  if ((rc = wgrad(x, dact1, 1, 32, gw1, gb1))) return rc;
  return WW_OK;
}

int ww_launch_conv_fp32(ww_ctx* c, const float* logmel, int B, cudaStream_t st) {
  const int H = c->cfg.n_mels, W = c->W;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int n_tiles = tiles_x * tiles_y;
  c->n_pool_part = n_tiles;
  dim3 block(kThreads);
  {
  ProfScope prof(c, WW_STAGE_CONV12, st);
  conv3x3_relu_kernel<1, 32, 32, 1, EPI_RELU><<<dim3(n_tiles, 1, B), block, 0, st>>>(
      logmel, c->d_convw_t[0], c->w["conv1.bias"], c->ws_act1, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  conv3x3_relu_kernel<32, 64, 32, 8, EPI_RELU><<<dim3(n_tiles, 2, B), block, 0, st>>>(
      c->ws_act1, c->d_convw_t[1], c->w["conv2.bias"], c->ws_act2, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  }
  ProfScope prof3(c, WW_STAGE_CONV3, st);
  conv3x3_relu_kernel<64, 128, 32, 8, EPI_POOL><<<dim3(n_tiles, 4, B), block, 0, st>>>(
      c->ws_act2, c->d_convw_t[2], c->w["conv3.bias"], c->pool_cur, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

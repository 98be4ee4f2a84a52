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

template <int CIN, int COUT, int COUT_TILE, int CC, bool POOL>
__global__ void __launch_bounds__(kThreads)
conv3x3_relu_kernel(const float* __restrict__ in,     // [B][CIN][H][W]
                    const float* __restrict__ wt,     // [CIN][9][COUT]
                    const float* __restrict__ bias,   // [COUT]
                    float* __restrict__ out,          // POOL ? [B][n_tiles][COUT] : [B][COUT][H][W]
                    int H, int W, int tiles_x, int n_tiles) {
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

  float acc[COUT_TILE];
#pragma unroll
  for (int i = 0; i < COUT_TILE; ++i) acc[i] = 0.0f;

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
#pragma unroll
        for (int co = 0; co < COUT_TILE; co += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&s_w[c][k][co]);
          acc[co + 0] = fmaf(v[k], w4.x, acc[co + 0]);
          acc[co + 1] = fmaf(v[k], w4.y, acc[co + 1]);
          acc[co + 2] = fmaf(v[k], w4.z, acc[co + 2]);
          acc[co + 3] = fmaf(v[k], w4.w, acc[co + 3]);
        }
      }
    }
  }

  const int y = y0 + ty, x = x0 + tx;
  const bool valid = (y < H) && (x < W);
  if (!POOL) {
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

}  // namespace

int ww_launch_conv_fp32(ww_ctx* c, const float* logmel, int B, cudaStream_t st) {
  const int H = c->cfg.n_mels, W = c->W;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int n_tiles = tiles_x * tiles_y;
  c->n_pool_part = n_tiles;
  dim3 block(kThreads);
  {
  ProfScope prof(c, WW_STAGE_CONV12, st);
  conv3x3_relu_kernel<1, 32, 32, 1, false><<<dim3(n_tiles, 1, B), block, 0, st>>>(
      logmel, c->d_convw_t[0], c->w["conv1.bias"], c->ws_act1, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  conv3x3_relu_kernel<32, 64, 32, 8, false><<<dim3(n_tiles, 2, B), block, 0, st>>>(
      c->ws_act1, c->d_convw_t[1], c->w["conv2.bias"], c->ws_act2, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  }
  ProfScope prof3(c, WW_STAGE_CONV3, st);
  conv3x3_relu_kernel<64, 128, 32, 8, true><<<dim3(n_tiles, 4, B), block, 0, st>>>(
      c->ws_act2, c->d_convw_t[2], c->w["conv3.bias"], c->pool_cur, H, W, tiles_x, n_tiles);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

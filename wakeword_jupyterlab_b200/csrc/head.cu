// K4: pooled features -> 2-layer LSTM (T = 1, zero state) -> Linear -> softmax -> threshold  (sm_100a)
//
// Replaces WakewordModel.forward after the pool (/root/reference/wakeword_training_script.py:175-182)
// and the scoring tail of predict_wakeword (wakeword_training.ipynb:886-891).
// The reference feeds the LSTM a length-1 sequence with h0 = c0 = 0, so per layer
//     g = W_ih x + b_ih + b_hh ;  c = sigmoid(g_i) tanh(g_g) ;  h = sigmoid(g_o) tanh(c)
// (gate rows i,f,g,o; weight_hh and the forget rows never reach the output -- SURVEY.md trap 2);
// eval-mode dropout is the identity.
//
// One launch per layer over the WHOLE batch: a register-tiled fp32 GEMM with the gate nonlinearity as epilogue.
// CTA tile = 64 clips x 64 hidden units x 3 gates, 256 threads, thread tile 4 clips x 4 units x 3 gates (48 FMAs
// per 4 LDS.128), K staged through shared memory in chunks of 32 (weights pre-transposed to [K][3][H] so the
// loads are coalesced).  Layer 0 finishes the global mean on the fly: x = sum of the conv kernel's per-group
// partial sums (fixed order) / (H*W).  A last small kernel does Linear + softmax + threshold (one warp per clip).
#include "ctx.cuh"

#include <algorithm>

namespace {

constexpr int TM = 64, TN = 64, KC = 32, kThreads = 256;

struct DenseParams {
  const float* x;          // [B][K] activations, or pool partials [B][n_part][128] when n_part > 0
  int n_part;
  float x_scale;
  const float* wt;         // [K][3][H]
  const float* bias;       // [3][H]
  float* out;              // [B][H]
  int B, K, H;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// packed fp32x2 FMA (two IEEE fma.rn per instruction; a 3-register FFMA only issues every other cycle)
typedef unsigned long long f32x2;
__device__ __forceinline__ void ffma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

__global__ void __launch_bounds__(kThreads, 3) gated_dense_kernel(DenseParams p) {
  __shared__ __align__(16) float xs[KC][TM];          // [k][clip]
  __shared__ __align__(16) float ws[KC][3][TN];       // [k][gate][unit]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;             // units tx*4.., clips ty*4..
  const int b0 = blockIdx.x * TM, j0 = blockIdx.y * TN;

  f32x2 acc2[3][4][2];                                 // [gate][clip][unit pair]
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc2[g][c][0] = 0ull; acc2[g][c][1] = 0ull; }

  for (int k0 = 0; k0 < p.K; k0 += KC) {
    __syncthreads();
    // ---- x tile: 64 clips x 32 k (global reads coalesced along k), stored transposed
    for (int i = tid; i < TM * KC; i += kThreads) {
      const int clip = i / KC, k = i % KC;
      float v = 0.0f;
      if (b0 + clip < p.B && k0 + k < p.K) {
        if (p.n_part > 0) {
          const float* pp = p.x + ((size_t)(b0 + clip) * p.n_part) * 128 + k0 + k;
          for (int t = 0; t < p.n_part; ++t) v += __ldg(pp + (size_t)t * 128);
          v *= p.x_scale;
        } else {
          v = __ldg(p.x + (size_t)(b0 + clip) * p.K + k0 + k);
        }
      }
      xs[k][clip] = v;
    }
    // ---- weight tile: 32 k x 3 gates x 64 units (coalesced along units)
    for (int i = tid; i < KC * 3 * TN; i += kThreads) {
      const int k = i / (3 * TN), r = i % (3 * TN), g = r / TN, u = r % TN;
      float v = 0.0f;
      if (k0 + k < p.K && j0 + u < p.H) v = __ldg(p.wt + ((size_t)(k0 + k) * 3 + g) * p.H + j0 + u);
      ws[k][g][u] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const f32x2 xx[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&ws[k][g][tx * 4]);     // (w0, w1), (w2, w3)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ffma2(acc2[g][c][0], xx[c], wv.x);
          ffma2(acc2[g][c][1], xx[c], wv.y);
        }
      }
    }
  }
  float acc[3][4][4];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      unpack2(acc2[g][c][0], acc[g][c][0], acc[g][c][1]);
      unpack2(acc2[g][c][1], acc[g][c][2], acc[g][c][3]);
    }
  // ---- gate epilogue: h = sigmoid(o) * tanh(sigmoid(i) * tanh(g))
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int b = b0 + ty * 4 + c;
    if (b >= p.B) continue;
    float h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + tx * 4 + u;
      float v = 0.0f;
      if (j < p.H) {
        const float gi = acc[0][c][u] + __ldg(p.bias + j);
        const float gg = acc[1][c][u] + __ldg(p.bias + p.H + j);
        const float go = acc[2][c][u] + __ldg(p.bias + 2 * p.H + j);
        v = sigmoidf_acc(go) * tanhf(sigmoidf_acc(gi) * tanhf(gg));
      }
      h[u] = v;
    }
    const int j = j0 + tx * 4;
    if (j + 3 < p.H) *reinterpret_cast<float4*>(p.out + (size_t)b * p.H + j) = make_float4(h[0], h[1], h[2], h[3]);
    else
      for (int u = 0; u < 4; ++u)
        if (j + u < p.H) p.out[(size_t)b * p.H + j + u] = h[u];
  }
}

// ---- pipelined variant (K % 32 == 0, H % 64 == 0, plain [B][K] input): the same tile shape and the same per-thread
// accumulation order (bit-identical results), but both tiles arrive by cp.async into a two-deep ring, so the L2 latency of
// chunk i+1 hides behind the FMAs of chunk i (one barrier per chunk).  x tile kept [clip][k] (pitch 36 floats: the two
// clip groups of a warp read banks 16 apart), read as one 128-bit load per clip and four k.
constexpr int XP = KC + 4;
constexpr size_t kPipeSmem = 2 * ((size_t)TM * XP + (size_t)KC * 3 * TN) * sizeof(float);

__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kThreads, 3) gated_dense_pipe_kernel(DenseParams p) {
  extern __shared__ __align__(16) float hs[];
  float* xs = hs;                                       // [2][TM][XP]
  float* ws = hs + 2 * TM * XP;                         // [2][KC][3][TN]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b0 = blockIdx.x * TM, j0 = blockIdx.y * TN;

  auto load_chunk = [&](int k0, int buf) {
    float* xb = xs + buf * TM * XP;
    float* wb = ws + buf * KC * 3 * TN;
    for (int i = tid; i < TM * (KC / 4); i += kThreads) {            // 64 clips x 8 pieces of 16 bytes
      const int clip = i >> 3, q = i & 7;
      const int b = min(b0 + clip, p.B - 1);                           // rows past the batch repeat the last clip (never stored)
      cp16(xb + clip * XP + q * 4, p.x + (size_t)b * p.K + k0 + q * 4);
    }
    for (int i = tid; i < KC * 3 * (TN / 4); i += kThreads) {          // 96 rows x 16 pieces
      const int row = i >> 4, q = i & 15;                              // row = k * 3 + g
      cp16(wb + row * TN + q * 4, p.wt + ((size_t)k0 * 3 + row) * p.H + j0 + q * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  f32x2 acc2[3][4][2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc2[g][c][0] = 0ull; acc2[g][c][1] = 0ull; }

  const int n_chunks = p.K / KC;
  load_chunk(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                     // chunk ch landed for everyone; everyone finished chunk ch-1
    if (ch + 1 < n_chunks) load_chunk((ch + 1) * KC, (ch + 1) & 1);
    const float* xb = xs + (ch & 1) * TM * XP + ty * 4 * XP;
    const float* wb = ws + (ch & 1) * KC * 3 * TN + tx * 4;
#pragma unroll 2
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 xv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xv[c] = *reinterpret_cast<const float4*>(xb + c * XP + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        f32x2 xx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float v = kk == 0 ? xv[c].x : kk == 1 ? xv[c].y : kk == 2 ? xv[c].z : xv[c].w;
          xx[c] = pack2(v, v);
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wb + ((k4 + kk) * 3 + g) * TN);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            ffma2(acc2[g][c][0], xx[c], wv.x);
            ffma2(acc2[g][c][1], xx[c], wv.y);
          }
        }
      }
    }
  }
  // ---- gate epilogue: h = sigmoid(o) * tanh(sigmoid(i) * tanh(g))
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int b = b0 + ty * 4 + c;
    if (b >= p.B) continue;
    float a[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g) { unpack2(acc2[g][c][0], a[g][0], a[g][1]); unpack2(acc2[g][c][1], a[g][2], a[g][3]); }
    float h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + tx * 4 + u;
      const float gi = a[0][u] + __ldg(p.bias + j);
      const float gg = a[1][u] + __ldg(p.bias + p.H + j);
      const float go = a[2][u] + __ldg(p.bias + 2 * p.H + j);
      h[u] = sigmoidf_acc(go) * tanhf(sigmoidf_acc(gi) * tanhf(gg));
    }
    *reinterpret_cast<float4*>(p.out + (size_t)b * p.H + j0 + tx * 4) = make_float4(h[0], h[1], h[2], h[3]);
  }
}

// finishes the global mean: x0[b][k] = (sum of the conv kernel's per-group partials, fixed order) * scale
__global__ void pool_finish_kernel(const float* __restrict__ part, int n_part, float scale, float* __restrict__ x0, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i >> 7;
    const int k = (int)(i & 127);
    const float* pp = part + (b * n_part) * 128 + k;
    float v = 0.0f;
    for (int t = 0; t < n_part; ++t) v += __ldg(pp + (size_t)t * 128);
    x0[i] = v * scale;
  }
}

struct FcParams {
  const float* h;          // [B][H]
  const float* fc_w;       // [n_classes][H]
  const float* fc_b;
  float* logits;           // [B][n_classes] or null
  float* prob1;            // [B] or null
  uint8_t* decision;       // [B] or null
  float threshold;
  int B, H, n_classes;
};

// Linear(H -> n_classes) + softmax + threshold: one warp per clip
__global__ void __launch_bounds__(256) fc_softmax_kernel(FcParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.B) return;
  const float* __restrict__ h = p.h + (size_t)warp * p.H;
  float l[16];
  for (int cls = 0; cls < p.n_classes; ++cls) {
    float s = 0.0f;
    for (int j = lane; j < p.H; j += 32) s = fmaf(__ldg(h + j), __ldg(p.fc_w + (size_t)cls * p.H + j), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    l[cls] = s + __ldg(p.fc_b + cls);
  }
  if (lane == 0) {
    float mx = l[0];
    for (int i = 1; i < p.n_classes; ++i) mx = fmaxf(mx, l[i]);
    float den = 0.0f;
    for (int i = 0; i < p.n_classes; ++i) den += expf(l[i] - mx);
    const float p1 = expf(l[p.n_classes > 1 ? 1 : 0] - mx) / den;
    if (p.logits) for (int i = 0; i < p.n_classes; ++i) p.logits[(size_t)warp * p.n_classes + i] = l[i];
    if (p.prob1) p.prob1[warp] = p1;
    if (p.decision) p.decision[warp] = (p1 >= p.threshold) ? 1 : 0;
  }
}

}  // namespace

// pool partials of clips [0, B) at c->ws_pool_part -> logits / prob1 / decision
int ww_launch_head(ww_ctx* c, int B, float* logits, float* prob1, uint8_t* decision, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  const int H = c->cfg.hidden_size;
  ProfScope prof(c, WW_STAGE_HEAD, st);
  const float* x = c->ws_pool_part;
  const bool pipe = (H % TN) == 0;                      // K is 128 or H: multiples of KC whenever H % 64 == 0
  if (pipe) {
    // per device, not per process: set on every call (cheap)
    WW_CHECK(c, cudaFuncSetAttribute(gated_dense_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
    // ws_h[1] (>= 128 floats per clip) holds the finished mean until layer 1 overwrites it with its own output
    const int64_t n = (int64_t)B * 128;
    pool_finish_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, st>>>(
        c->ws_pool_part, c->n_pool_part, 1.0f / (float)(c->cfg.n_mels * c->W), c->ws_h[1], n);
    WW_LAUNCH_CHECK(c);
    x = c->ws_h[1];
  }
  for (int l = 0; l < c->cfg.num_layers; ++l) {
    DenseParams p;
    p.x = x; p.n_part = (l == 0 && !pipe) ? c->n_pool_part : 0;
    p.x_scale = 1.0f / (float)(c->cfg.n_mels * c->W);
    p.wt = c->d_head_wt[l]; p.bias = c->d_head_b[l];
    p.out = c->ws_h[l & 1]; p.B = B; p.K = (l == 0) ? 128 : H; p.H = H;
    dim3 grid((B + TM - 1) / TM, (H + TN - 1) / TN);
    // the tensor-core kernel (head_tc.cu) where its tile shape applies, else the fp32 kernels of this file
    const int rc = pipe ? ww_launch_gated_dense_tc(c, l, p.x, p.out, B, st) : 1;
    if (rc < 0) return rc;
    if (rc == 1) {
      if (pipe) gated_dense_pipe_kernel<<<grid, kThreads, kPipeSmem, st>>>(p);
      else gated_dense_kernel<<<grid, kThreads, 0, st>>>(p);
      WW_LAUNCH_CHECK(c);
    }
    x = p.out;
  }
  FcParams f;
  f.h = x; f.fc_w = c->w["fc.weight"]; f.fc_b = c->w["fc.bias"];
  f.logits = logits; f.prob1 = prob1; f.decision = decision; f.threshold = c->cfg.threshold;
  f.B = B; f.H = H; f.n_classes = c->cfg.num_classes;
  fc_softmax_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(f);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

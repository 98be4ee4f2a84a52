// K4: pooled features -> 2-layer LSTM (T = 1, zero state) -> Linear -> softmax -> threshold  (sm_100a)
//
// Replaces WakewordModel.forward after the pool (/root/reference/wakeword_training_script.py:175-182)
// and the scoring tail of predict_wakeword (wakeword_training.ipynb:886-891).
// The reference feeds the LSTM a length-1 sequence with h0 = c0 = 0, so per layer
//     g = W_ih x + b_ih + b_hh ;  c = sigmoid(g_i) tanh(g_g) ;  h = sigmoid(g_o) tanh(c)
// (gate rows i,f,g,o; weight_hh and the forget rows never reach the output -- SURVEY.md trap 2);
// eval-mode dropout is the identity.  Weights are pre-transposed to [K][3][H] (gates i,g,o) so that
// consecutive threads (hidden units) read consecutive floats; activations of CPB clips sit in shared
// memory as [K][CPB] and are read as float4 broadcasts: 7 loads per 48 FMAs.
// Also finishes the global mean: sums the conv kernel's per-tile partial sums in a fixed order.
#include "ctx.cuh"

namespace {

constexpr int CPB = 8;    // clips per CTA

struct HeadParams {
  const float* pool_part;   // [B][n_part][128] partial sums of relu(conv3)
  int n_part;
  float inv_hw;
  int B, H, layers, n_classes;
  const float* wt[8];       // [K][3][H]
  const float* bias[8];     // [3][H]
  const float* fc_w;        // [n_classes][H]
  const float* fc_b;
  float* logits;            // [B][n_classes] or null
  float* prob1;             // [B] or null
  uint8_t* decision;        // [B] or null
  float threshold;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void head_kernel(HeadParams p) {
  extern __shared__ __align__(16) float sm[];
  const int H = p.H;
  const int Kmax = H > 128 ? H : 128;
  float* xs = sm;                   // [Kmax][CPB]
  float* hs = sm + Kmax * CPB;      // [Kmax][CPB]
  float* lg = hs + Kmax * CPB;      // [CPB][n_classes]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * CPB;

  for (int i = tid; i < 128 * CPB; i += blockDim.x) {
    const int clip = i >> 7, k = i & 127;
    float s = 0.0f;
    if (b0 + clip < p.B) {
      const float* pp = p.pool_part + ((size_t)(b0 + clip) * p.n_part) * 128 + k;
      for (int t = 0; t < p.n_part; ++t) s += __ldg(pp + (size_t)t * 128);
    }
    xs[k * CPB + clip] = s * p.inv_hw;
  }
  __syncthreads();

  int K = 128;
  for (int l = 0; l < p.layers; ++l) {
    const float* __restrict__ wt = p.wt[l];
    const float* __restrict__ bs = p.bias[l];
    for (int j = tid; j < H; j += blockDim.x) {
      float ai[CPB], ag[CPB], ao[CPB];
      const float bi = __ldg(bs + j), bg = __ldg(bs + H + j), bo = __ldg(bs + 2 * H + j);
#pragma unroll
      for (int c = 0; c < CPB; ++c) { ai[c] = bi; ag[c] = bg; ao[c] = bo; }
#pragma unroll 8
      for (int k = 0; k < K; ++k) {
        const float wi = __ldg(wt + ((size_t)k * 3 + 0) * H + j);
        const float wg = __ldg(wt + ((size_t)k * 3 + 1) * H + j);
        const float wo = __ldg(wt + ((size_t)k * 3 + 2) * H + j);
        const float4* xr = reinterpret_cast<const float4*>(xs + k * CPB);
#pragma unroll
        for (int c4 = 0; c4 < CPB / 4; ++c4) {
          const float4 x = xr[c4];
          ai[c4 * 4 + 0] = fmaf(wi, x.x, ai[c4 * 4 + 0]); ag[c4 * 4 + 0] = fmaf(wg, x.x, ag[c4 * 4 + 0]); ao[c4 * 4 + 0] = fmaf(wo, x.x, ao[c4 * 4 + 0]);
          ai[c4 * 4 + 1] = fmaf(wi, x.y, ai[c4 * 4 + 1]); ag[c4 * 4 + 1] = fmaf(wg, x.y, ag[c4 * 4 + 1]); ao[c4 * 4 + 1] = fmaf(wo, x.y, ao[c4 * 4 + 1]);
          ai[c4 * 4 + 2] = fmaf(wi, x.z, ai[c4 * 4 + 2]); ag[c4 * 4 + 2] = fmaf(wg, x.z, ag[c4 * 4 + 2]); ao[c4 * 4 + 2] = fmaf(wo, x.z, ao[c4 * 4 + 2]);
          ai[c4 * 4 + 3] = fmaf(wi, x.w, ai[c4 * 4 + 3]); ag[c4 * 4 + 3] = fmaf(wg, x.w, ag[c4 * 4 + 3]); ao[c4 * 4 + 3] = fmaf(wo, x.w, ao[c4 * 4 + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < CPB; ++c) {
        const float cc = sigmoidf_acc(ai[c]) * tanhf(ag[c]);
        hs[j * CPB + c] = sigmoidf_acc(ao[c]) * tanhf(cc);
      }
    }
    __syncthreads();
    float* t = xs; xs = hs; hs = t;
    K = H;
  }

  // ---- Linear(H -> n_classes): one warp per (clip, class)
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int q = warp; q < CPB * p.n_classes; q += nwarps) {
    const int clip = q / p.n_classes, cls = q % p.n_classes;
    float s = 0.0f;
    for (int j = lane; j < H; j += 32) s = fmaf(xs[j * CPB + clip], __ldg(p.fc_w + (size_t)cls * H + j), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) lg[q] = s + __ldg(p.fc_b + cls);
  }
  __syncthreads();
  for (int clip = tid; clip < CPB; clip += blockDim.x) {
    const int b = b0 + clip;
    if (b >= p.B) continue;
    const float* l = lg + clip * p.n_classes;
    float mx = l[0];
    for (int i = 1; i < p.n_classes; ++i) mx = fmaxf(mx, l[i]);
    float den = 0.0f;
    for (int i = 0; i < p.n_classes; ++i) den += expf(l[i] - mx);
    const float p1 = expf(l[p.n_classes > 1 ? 1 : 0] - mx) / den;
    if (p.logits) for (int i = 0; i < p.n_classes; ++i) p.logits[(size_t)b * p.n_classes + i] = l[i];
    if (p.prob1) p.prob1[b] = p1;
    if (p.decision) p.decision[b] = (p1 >= p.threshold) ? 1 : 0;
  }
}

}  // namespace

int ww_launch_head(ww_ctx* c, int B, float* logits, float* prob1, uint8_t* decision, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  HeadParams p;
  p.pool_part = c->ws_pool_part; p.n_part = c->n_pool_part;
  p.inv_hw = 1.0f / (float)(c->cfg.n_mels * c->W);
  p.B = B; p.H = c->cfg.hidden_size; p.layers = c->cfg.num_layers; p.n_classes = c->cfg.num_classes;
  for (int l = 0; l < 8; ++l) { p.wt[l] = c->d_head_wt[l]; p.bias[l] = c->d_head_b[l]; }
  p.fc_w = c->w["fc.weight"]; p.fc_b = c->w["fc.bias"];
  p.logits = logits; p.prob1 = prob1; p.decision = decision; p.threshold = c->cfg.threshold;
  const int H = p.H, Kmax = H > 128 ? H : 128;
  size_t smem = ((size_t)2 * Kmax * CPB + CPB * p.n_classes) * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    WW_CHECK(c, cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int threads = H < 1024 ? H : 1024;
  if (threads < 64) threads = 64;
  ProfScope prof(c, WW_STAGE_HEAD, st);
  head_kernel<<<(B + CPB - 1) / CPB, threads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

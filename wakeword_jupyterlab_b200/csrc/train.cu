// Training step: forward + backward + Adam of the CNN+LSTM classifier  (sm_100a)
//
// Replaces the body of WakewordTrainer.train_epoch's loop (/root/reference/wakeword_training_script.py:247-257)
// with the optimiser of WakewordTrainer.__init__ (:225-226):
//     zero_grad; output = model(data); loss = CrossEntropyLoss(output, target);
//     clip_grad_norm_ (called BEFORE backward on zeroed gradients: a no-op, kept as such); loss.backward();
//     Adam(lr, betas (0.9, 0.999), eps 1e-8, weight_decay 1e-5 added to the gradient).step()
// The LSTM sees a length-1 sequence with zero state, so per layer only the i, g, o rows of weight_ih and the two biases
// receive data gradients; weight_hh and the forget-gate rows get zero data gradient but still move under weight decay,
// exactly as in the reference (every state_dict entry is an Adam parameter).
//
// Round-1 arithmetic: exact fp32 on CUDA cores (conv_fp32.cu kernels for the conv stack, a tiled SGEMM for the head);
// the tcgen05 data/weight-gradient kernels are the next step.  Gradients live in one flat buffer in
// state_dict order, so data parallelism is one all-reduce of that buffer (NCCL through the caller's ncclComm_t, or
// torch.distributed on the exposed pointer) between ww_train_backward and ww_train_apply.
#include "ctx.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>

int ww_train_conv_forward(ww_ctx* c, const float* x, int B, float* act1, float* act2, float* act3, float* pooled, cudaStream_t st);
int ww_train_conv_backward(ww_ctx* c, const float* x, int B, const float* act1, const float* act2, float* act3,
                           const float* dpooled, float* dact2, float* dact1, const float* wflip3, const float* wflip2,
                           float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3, float* part,
                           int slices, cudaStream_t st);

namespace {

constexpr int kSlices = 64;
constexpr int kFastSteps = 64;      // optimiser steps between two host-side weight preparations (tensor-core training loop)
constexpr int64_t kPartFloats = (int64_t)kSlices * 128 * 64 * 9;      // TrainState::part: conv partial sums / split-K scratch

// C[m][n] (+)= sum_k A(m,k) * B(k,n) with arbitrary element strides; 64x64 tile, 16-wide k steps, 4x4 per thread
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                            const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                            float* __restrict__ C, int64_t scm, int64_t scn, int M,
                                                            int N, int K, const float* __restrict__ bias_n, int k_slice,
                                                            int64_t c_slice) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  // split-K: slice z sums k in [z k_slice, (z + 1) k_slice) into its own dense [M][N] block of C (scratch), which
  // reduce_k_slices_kernel adds up in fixed order
  const int k_lo = blockIdx.z * k_slice;
  K = min(K, k_lo + k_slice);
  C += (int64_t)blockIdx.z * c_slice;
  for (int k0 = k_lo; k0 < K; k0 += 16) {
    __syncthreads();
    for (int i = tid; i < 16 * 64; i += 256) {
      const int k = i & 15, m = i >> 4;
      As[k][m] = (m0 + m < M && k0 + k < K) ? A[(int64_t)(m0 + m) * sam + (int64_t)(k0 + k) * sak] : 0.0f;
      const int n = i & 63, kb = i >> 6;
      Bs[kb][n] = (n0 + n < N && k0 + kb < K) ? B[(int64_t)(k0 + kb) * sbk + (int64_t)(n0 + n) * sbn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(int64_t)m * scm + (int64_t)n * scn] = acc[i][j] + (bias_n ? bias_n[n] : 0.0f);
    }
}

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// LSTM cell, T = 1, zero state: gates [B][4H] (rows i,f,g,o, bias already added) -> h [B][H]; keeps i, g, o, tanh(c)
// in `gates` itself (slots i, f <- tanh(c), g, o) for the backward pass.  drop (or null): multiplicative mask on h.
__global__ void lstm_cell_fwd_kernel(float* __restrict__ gates, float* __restrict__ h, const float* __restrict__ drop,
                                     int B, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float* g = gates + (int64_t)b * 4 * H;
  const float i = sigm(g[j]), gg = tanhf(g[2 * H + j]), o = sigm(g[3 * H + j]);
  const float tc = tanhf(i * gg);
  g[j] = i; g[H + j] = tc; g[2 * H + j] = gg; g[3 * H + j] = o;
  float v = o * tc;
  if (drop) v *= drop[idx];
  h[idx] = v;
}

// dh [B][H] (+ saved i, tanh(c), g, o in `gates`) -> dgates [B][4H] in place (forget rows: 0)
__global__ void lstm_cell_bwd_kernel(float* __restrict__ gates, const float* __restrict__ dh, const float* __restrict__ drop,
                                     int B, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float* g = gates + (int64_t)b * 4 * H;
  const float i = g[j], tc = g[H + j], gg = g[2 * H + j], o = g[3 * H + j];
  float d = dh[idx];
  if (drop) d *= drop[idx];
  const float d_o = d * tc, d_c = d * o * (1.0f - tc * tc);
  g[j] = d_c * gg * i * (1.0f - i);
  g[H + j] = 0.0f;
  g[2 * H + j] = d_c * i * (1.0f - gg * gg);
  g[3 * H + j] = d_o * o * (1.0f - o);
}

// C[m scm + n scn] = sum_z part[z][m][n] (+ bias[n]): the fixed-order reduction of a split-K product
__global__ void reduce_k_slices_kernel(const float* __restrict__ part, int slices, float* __restrict__ C, int64_t scm, int64_t scn,
                                       int M, int N, const float* __restrict__ bias_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.0f;
  for (int z = 0; z < slices; ++z) s += part[(int64_t)z * M * N + i];
  const int m = i / N, n = i - m * N;
  C[(int64_t)m * scm + (int64_t)n * scn] = s + (bias_n ? bias_n[n] : 0.0f);
}

// y[n] = sum_m X[m][n]: block = 32 columns x 32 row lanes (each sums rows r, r + 32, ..), then a shared-memory tree
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ X, float* __restrict__ y, float* __restrict__ y2,
                                                      int M, int N) {
  __shared__ float red[32][33];
  const int n = blockIdx.x * 32 + threadIdx.x, r = threadIdx.y;
  float s = 0.0f;
  if (n < N)
    for (int m = r; m < M; m += 32) s += X[(int64_t)m * N + n];
  red[r][threadIdx.x] = s;
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.0f;
    for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
    y[n] = t;
    if (y2) y2[n] = t;
  }
}

// per-row softmax cross entropy: loss_row[b], dlogits[b][c] = (softmax - onehot) / B
__global__ void ce_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ dlogits,
                               float* __restrict__ loss_row, int B, int C) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = logits + (int64_t)b * C;
  float mx = l[0];
  for (int c = 1; c < C; ++c) mx = fmaxf(mx, l[c]);
  float den = 0.0f;
  for (int c = 0; c < C; ++c) den += expf(l[c] - mx);
  const int y = (int)labels[b];
  loss_row[b] = (y >= 0 && y < C) ? -(l[y] - mx - logf(den)) : 0.0f;
  for (int c = 0; c < C; ++c) dlogits[(int64_t)b * C + c] = (expf(l[c] - mx) / den - (c == y ? 1.0f : 0.0f)) / (float)B;
}
__global__ void __launch_bounds__(256) mean_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
  __shared__ float red[256];
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) s += x[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] / (float)n;
}

// torch.optim.Adam (single-tensor path) with coupled weight decay, in its order of operations; ONE launch walks the flat
// gradient / moment buffers and finds each element's parameter tensor in a table (tab: n_tensors x {offset, count}).
struct AdamTable { float* ptr[48]; int64_t off[48]; int64_t cnt[48]; int n; };

__global__ void adam_kernel(const __grid_constant__ AdamTable tab, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n_flat, float lr, float b1, float b2, float eps, float wd, float bc1,
                            float bc2_sqrt, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_flat) return;
  int k = 0;
  while (k + 1 < tab.n && i >= tab.off[k + 1]) ++k;
  const int64_t j = i - tab.off[k];
  if (j >= tab.cnt[k]) return;                                         // alignment padding between tensors
  float* p = tab.ptr[k] + j;
  const float grad = fmaf(wd, *p, g[i] * grad_scale);
  const float mi = m[i] + (grad - m[i]) * (1.0f - b1);                 // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = fmaf(1.0f - b2, grad * grad, v[i] * b2);            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  *p = *p - (lr / bc1) * (mi / denom);
}

// scatter the i, g, o rows of a [4H][K] gradient and zero the f rows: handled by computing full [4H][K] (f rows of dgates are 0)

int sgemm(ww_ctx* c, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
          int64_t scm, int64_t scn, int M, int N, int K, const float* bias, cudaStream_t st, float* scratch = nullptr,
          int64_t scratch_floats = 0) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  // few output tiles and a long K (the weight gradients: K = batch): split K over blockIdx.z into the scratch buffer
  const int tiles = (int)(grid.x * grid.y);
  int slices = 1;
  if (scratch && K >= 512 && tiles < 2 * c->sm_count) {
    slices = std::min({(2 * c->sm_count + tiles - 1) / tiles, K / 128, 64});
    while (slices > 1 && (int64_t)slices * M * N > scratch_floats) --slices;
  }
  if (slices <= 1) {
    sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, scm, scn, M, N, K, bias, K, 0);
    WW_LAUNCH_CHECK(c);
    return WW_OK;
  }
  const int k_slice = (((K + slices - 1) / slices) + 15) & ~15;
  grid.z = (K + k_slice - 1) / k_slice;
  sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, scratch, N, 1, M, N, K, nullptr, k_slice, (int64_t)M * N);
  WW_LAUNCH_CHECK(c);
  reduce_k_slices_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(scratch, (int)grid.z, C, scm, scn, M, N, bias);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

std::vector<std::string> param_order(const ww_ctx* c) {
  std::vector<std::string> r = {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"};
  for (int l = 0; l < c->cfg.num_layers; ++l) {
    const std::string s = std::to_string(l);
    r.push_back("lstm.weight_ih_l" + s); r.push_back("lstm.weight_hh_l" + s);
    r.push_back("lstm.bias_ih_l" + s); r.push_back("lstm.bias_hh_l" + s);
  }
  r.push_back("fc.weight"); r.push_back("fc.bias");
  return r;
}

int ensure_train(ww_ctx* c, int B, bool fp32_conv = false) {
  TrainState& t = c->train;
  const int H = c->cfg.n_mels, W = c->W, HW = H * W, Hd = c->cfg.hidden_size, L = c->cfg.num_layers, C = c->cfg.num_classes;
  if (t.names.empty()) {
    t.names = param_order(c);
    int64_t off = 0;
    for (const std::string& n : t.names) {
      if (!c->w.count(n)) { c->set_error("train: weight not set: " + n); return WW_ERR_WEIGHTS; }
      int64_t cnt = 1;
      for (int64_t d : c->w_shape[n]) cnt *= d;
      t.offset[n] = off; t.count[n] = cnt;
      off += (cnt + 3) & ~(int64_t)3;
    }
    t.n_flat = off;
    WW_CHECK(c, cudaMalloc((void**)&t.grad, off * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.m, off * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.v, off * 4));
    WW_CHECK(c, cudaMemset(t.grad, 0, off * 4));
    WW_CHECK(c, cudaMemset(t.m, 0, off * 4));
    WW_CHECK(c, cudaMemset(t.v, 0, off * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.wflip3, (size_t)128 * 9 * 64 * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.wflip2, (size_t)64 * 9 * 32 * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.part, (size_t)kSlices * 128 * 64 * 9 * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.loss, 4));
  }
  if (B > t.cap) {
    float** bufs[] = {&t.pooled, &t.dpooled, &t.gates, &t.hbuf, &t.dh, &t.logits, &t.dlogits, &t.loss_row};
    for (float** b : bufs) { cudaFree(*b); *b = nullptr; }
    const size_t n = (size_t)B;
    WW_CHECK(c, cudaMalloc((void**)&t.pooled, n * 128 * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.dpooled, n * std::max(128, Hd) * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.gates, n * L * 4 * Hd * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.hbuf, n * L * Hd * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.dh, n * 2 * Hd * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.logits, n * C * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.dlogits, n * C * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.loss_row, n * 4));
    t.cap = B;
  }
  if (fp32_conv && B > t.cap32) {       // full fp32 activations: only the exact CUDA-core path keeps them
    float** bufs[] = {&t.act1, &t.act2, &t.act3, &t.dact2, &t.dact1};
    for (float** b : bufs) { cudaFree(*b); *b = nullptr; }
    const size_t n = (size_t)B;
    WW_CHECK(c, cudaMalloc((void**)&t.act1, n * 32 * HW * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.act2, n * 64 * HW * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.act3, n * 128 * HW * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.dact2, n * 64 * HW * 4));
    WW_CHECK(c, cudaMalloc((void**)&t.dact1, n * 32 * HW * 4));
    t.cap32 = B;
  }
  return WW_OK;
}

// [Cout][Cin][3][3] -> [Cout][9][Cin] with flipped taps: the data gradient is a convolution with these weights
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, int COUT, int CIN) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= COUT * CIN * 9) return;
  const int k = i % 9, ci = (i / 9) % CIN, co = i / (9 * CIN);
  wf[((size_t)co * 9 + (8 - k)) * CIN + ci] = w[i];
}

}  // namespace

extern "C" {

int64_t ww_train_n_params(ww_ctx* c) {
  if (!c) return -1;
  DeviceGuard dev_guard(c->device);
  if (ensure_train(c, 0)) return -1;
  return c->train.n_flat;
}

int ww_train_param_range(ww_ctx* c, const char* name, int64_t* offset, int64_t* count) {
  if (!c || !name) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  int rc = ensure_train(c, 0);
  if (rc) return rc;
  auto it = c->train.offset.find(name);
  if (it == c->train.offset.end()) { c->set_error(std::string("train: unknown parameter ") + name); return WW_ERR_INVALID; }
  if (offset) *offset = it->second;
  if (count) *count = c->train.count[name];
  return WW_OK;
}

float* ww_train_grad_buffer(ww_ctx* c) {
  if (!c) return nullptr;
  DeviceGuard dev_guard(c->device);
  if (ensure_train(c, 0)) return nullptr;
  return c->train.grad;
}

int ww_get_weights(ww_ctx* c, const char* name, float* dst) {
  if (!c || !name || !dst) return WW_ERR_INVALID;
  auto it = c->w.find(name);
  if (it == c->w.end()) { c->set_error(std::string("ww_get_weights: unknown parameter ") + name); return WW_ERR_INVALID; }
  size_t n = 1;
  for (int64_t d : c->w_shape[name]) n *= (size_t)d;
  DeviceGuard dev_guard(c->device);
  // the Adam kernels of the last ww_train_apply ran on the caller's stream, which a blocking copy on the legacy stream does
  // not wait for when that stream is non-blocking: order the copy behind them explicitly
  if (c->apply_event) WW_CHECK(c, cudaEventSynchronize(c->apply_event));
  WW_CHECK(c, cudaMemcpy(dst, it->second, n * 4, cudaMemcpyDefault));
  return WW_OK;
}

// Adam state of one parameter (torch.optim.Adam's exp_avg / exp_avg_sq) and the shared step counter: what
// optimizer.state_dict() holds in the reference's best_wakeword_model.pth (wakeword_training_script.py:326-334).
// Pointers may be device or host memory; NULL skips that moment.
int ww_train_get_moments(ww_ctx* c, const char* name, float* exp_avg, float* exp_avg_sq) {
  if (!c || !name) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  int rc = ensure_train(c, 0);
  if (rc) return rc;
  auto it = c->train.offset.find(name);
  if (it == c->train.offset.end()) { c->set_error(std::string("train: unknown parameter ") + name); return WW_ERR_INVALID; }
  const size_t bytes = (size_t)c->train.count[name] * 4;
  if (c->apply_event) WW_CHECK(c, cudaEventSynchronize(c->apply_event));
  if (exp_avg) WW_CHECK(c, cudaMemcpy(exp_avg, c->train.m + it->second, bytes, cudaMemcpyDefault));
  if (exp_avg_sq) WW_CHECK(c, cudaMemcpy(exp_avg_sq, c->train.v + it->second, bytes, cudaMemcpyDefault));
  return WW_OK;
}

int ww_train_set_moments(ww_ctx* c, const char* name, const float* exp_avg, const float* exp_avg_sq) {
  if (!c || !name) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  int rc = ensure_train(c, 0);
  if (rc) return rc;
  auto it = c->train.offset.find(name);
  if (it == c->train.offset.end()) { c->set_error(std::string("train: unknown parameter ") + name); return WW_ERR_INVALID; }
  const size_t bytes = (size_t)c->train.count[name] * 4;
  WW_CHECK(c, cudaDeviceSynchronize());
  if (exp_avg) WW_CHECK(c, cudaMemcpy(c->train.m + it->second, exp_avg, bytes, cudaMemcpyDefault));
  if (exp_avg_sq) WW_CHECK(c, cudaMemcpy(c->train.v + it->second, exp_avg_sq, bytes, cudaMemcpyDefault));
  return WW_OK;
}

int64_t ww_train_get_step(ww_ctx* c) { return c ? c->train.step : -1; }

int ww_train_set_step(ww_ctx* c, int64_t step) {
  if (!c || step < 0) return WW_ERR_INVALID;
  c->train.step = step;
  return WW_OK;
}

int ww_train_backward(ww_ctx* c, const float* x, const int64_t* labels, int B, const float* drop_lstm,
                      const float* drop_out, float* loss, float* logits_out, void* stream) {
  if (!c || !x || !labels || B <= 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  // conv stack: tcgen05 kernels (train_tc.cu) unless the context computes in exact fp32 or WW_TRAIN_KERNEL=fp32 asks for it
  const char* kern = getenv("WW_TRAIN_KERNEL");
  const bool tc = ww_train_tc_supported(c) && !(kern && kern[0] == 'f');
  TrainState& t = c->train;
  int rc;
  if (tc && t.fast_pending && t.tc_version == c->weights_version) {
    // the last ww_train_apply rebuilt every operand form this step reads ON THE DEVICE (ww_train_tc_repack): no host work
    // except the conv biases, which travel as kernel parameters
    if ((rc = ww_train_tc_sync_biases(c))) return rc;
  } else {
    if ((rc = ww_prepare_weights(c, st))) return rc;
    if (tc && t.tc_version != c->weights_version) {
      WW_CHECK(c, cudaStreamSynchronize(st));
      if ((rc = ww_train_tc_prepare(c))) return rc;
      t.tc_version = c->weights_version;
      t.fast_steps = 0;
      t.fast_drift = 0.0f;
    }
  }
  t.tc_last = tc;
  if ((rc = ensure_train(c, B, !tc))) return rc;
  const int Hd = c->cfg.hidden_size, L = c->cfg.num_layers, C = c->cfg.num_classes;
  auto G = [&](const std::string& n) { return t.grad + t.offset[n]; };
  if (!tc) {
    flip_weights_kernel<<<(128 * 64 * 9 + 255) / 256, 256, 0, st>>>(c->w["conv3.weight"], t.wflip3, 128, 64);
    WW_LAUNCH_CHECK(c);
    flip_weights_kernel<<<(64 * 32 * 9 + 255) / 256, 256, 0, st>>>(c->w["conv2.weight"], t.wflip2, 64, 32);
    WW_LAUNCH_CHECK(c);
  }
  WW_CHECK(c, cudaMemsetAsync(t.grad, 0, t.n_flat * 4, st));                      // optimizer.zero_grad()

  // ---- forward (train mode)
  if (tc) rc = ww_train_tc_forward(c, x, B, t.pooled, st);
  else rc = ww_train_conv_forward(c, x, B, t.act1, t.act2, t.act3, t.pooled, st);
  if (rc) return rc;
  const float* xin = t.pooled;
  int K = 128;
  for (int l = 0; l < L; ++l) {
    const std::string s = std::to_string(l);
    float* gates = t.gates + (size_t)l * B * 4 * Hd;
    float* h = t.hbuf + (size_t)l * B * Hd;
    // gates = x W_ih^T + b_ih (+ b_hh below)
    if ((rc = sgemm(c, xin, K, 1, c->w["lstm.weight_ih_l" + s], 1, K, gates, 4 * Hd, 1, B, 4 * Hd, K,
                    c->d_bias_sum[l], st))) return rc;
    const float* drop = (l + 1 < L) ? drop_lstm : drop_out;        // nn.LSTM dropout between layers, nn.Dropout after the last
    if (l + 1 < L && drop_lstm) drop = drop_lstm + (size_t)l * B * Hd;
    lstm_cell_fwd_kernel<<<(int)(((int64_t)B * Hd + 255) / 256), 256, 0, st>>>(gates, h, drop, B, Hd);
    WW_LAUNCH_CHECK(c);
    xin = h; K = Hd;
  }
  if ((rc = sgemm(c, xin, Hd, 1, c->w["fc.weight"], 1, Hd, t.logits, C, 1, B, C, Hd, c->w["fc.bias"], st))) return rc;
  ce_loss_kernel<<<(B + 127) / 128, 128, 0, st>>>(t.logits, labels, t.dlogits, t.loss_row, B, C);
  WW_LAUNCH_CHECK(c);
  mean_kernel<<<1, 256, 0, st>>>(t.loss_row, t.loss, B);
  WW_LAUNCH_CHECK(c);
  if (loss) WW_CHECK(c, cudaMemcpyAsync(loss, t.loss, 4, cudaMemcpyDefault, st));
  if (logits_out) WW_CHECK(c, cudaMemcpyAsync(logits_out, t.logits, (size_t)B * C * 4, cudaMemcpyDefault, st));

  // ---- backward: fc
  if ((rc = sgemm(c, t.dlogits, 1, C, xin, Hd, 1, G("fc.weight"), Hd, 1, C, Hd, B, nullptr, st, t.part, kPartFloats))) return rc;   // dW = dlogits^T h
  colsum_kernel<<<1, dim3(32, 32), 0, st>>>(t.dlogits, G("fc.bias"), nullptr, B, C);
  WW_LAUNCH_CHECK(c);
  float* dh = t.dh;
  if ((rc = sgemm(c, t.dlogits, C, 1, c->w["fc.weight"], Hd, 1, dh, Hd, 1, B, Hd, C, nullptr, st))) return rc;  // dh = dlogits W
  // ---- backward: LSTM layers
  for (int l = L - 1; l >= 0; --l) {
    const std::string s = std::to_string(l);
    float* gates = t.gates + (size_t)l * B * 4 * Hd;
    const float* inp = (l == 0) ? t.pooled : t.hbuf + (size_t)(l - 1) * B * Hd;
    const int Kin = (l == 0) ? 128 : Hd;
    const float* drop = (l + 1 < L) ? (drop_lstm ? drop_lstm + (size_t)l * B * Hd : nullptr) : drop_out;
    lstm_cell_bwd_kernel<<<(int)(((int64_t)B * Hd + 255) / 256), 256, 0, st>>>(gates, dh, drop, B, Hd);
    WW_LAUNCH_CHECK(c);
    // dW_ih [4H][K] = dgates^T inp ; db_ih = db_hh = colsum(dgates) ; d_inp [B][K] = dgates W_ih
    if ((rc = sgemm(c, gates, 1, 4 * Hd, inp, Kin, 1, G("lstm.weight_ih_l" + s), Kin, 1, 4 * Hd, Kin, B, nullptr, st, t.part,
                    kPartFloats))) return rc;
    colsum_kernel<<<(4 * Hd + 31) / 32, dim3(32, 32), 0, st>>>(gates, G("lstm.bias_ih_l" + s), G("lstm.bias_hh_l" + s), B, 4 * Hd);
    WW_LAUNCH_CHECK(c);
    float* dnext = (l == 0) ? t.dpooled : (dh == t.dh ? t.dh + (size_t)B * Hd : t.dh);
    if ((rc = sgemm(c, gates, 4 * Hd, 1, c->w["lstm.weight_ih_l" + s], Kin, 1, dnext, Kin, 1, B, Kin, 4 * Hd, nullptr, st))) return rc;
    dh = dnext;
  }
  // ---- backward: conv stack
  if (tc)
    return ww_train_tc_backward(c, B, t.dpooled, G("conv1.weight"), G("conv1.bias"), G("conv2.weight"), G("conv2.bias"),
                                G("conv3.weight"), G("conv3.bias"), st);
  return ww_train_conv_backward(c, x, B, t.act1, t.act2, t.act3, t.dpooled, t.dact2, t.dact1, t.wflip3, t.wflip2,
                                G("conv1.weight"), G("conv1.bias"), G("conv2.weight"), G("conv2.bias"),
                                G("conv3.weight"), G("conv3.bias"), t.part, kSlices, st);
}

int ww_train_reset(ww_ctx* c) {
  if (!c) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  int rc = ensure_train(c, 0);
  if (rc) return rc;
  c->train.step = 0;
  WW_CHECK(c, cudaDeviceSynchronize());
  WW_CHECK(c, cudaMemset(c->train.m, 0, c->train.n_flat * 4));
  WW_CHECK(c, cudaMemset(c->train.v, 0, c->train.n_flat * 4));
  return WW_OK;
}

int ww_train_apply(ww_ctx* c, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   void* stream) {
  if (!c) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_train(c, 0);
  if (rc) return rc;
  TrainState& t = c->train;
  t.step += 1;
  // bias corrections in double, like torch.optim.Adam's Python-side scalars
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)t.step));
  const double bc2 = 1.0 - pow((double)beta2, (double)t.step);
  AdamTable tab;
  tab.n = (int)t.names.size();
  if (tab.n > 48) { c->set_error("ww_train_apply: too many parameter tensors"); return WW_ERR_INVALID; }
  for (int k = 0; k < tab.n; ++k) { tab.ptr[k] = c->w[t.names[k]]; tab.off[k] = t.offset[t.names[k]]; tab.cnt[k] = t.count[t.names[k]]; }
  adam_kernel<<<(int)((t.n_flat + 255) / 256), 256, 0, st>>>(tab, t.grad, t.m, t.v, t.n_flat, lr, beta1, beta2, eps, weight_decay,
                                                             bc1, (float)sqrt(bc2), grad_scale);
  WW_LAUNCH_CHECK(c);
  c->weights_version++;
  c->weights_dirty = true;      // prepared (transposed / split) forms are rebuilt at the next forward
  // training loop on the tensor-core kernels: rebuild the forms the NEXT step reads right here, on the device, with the
  // power-of-two scales of the last host-side preparation.  Those leave 8x headroom above the largest weight of a layer; a
  // full preparation is redone every kFastSteps steps, or earlier when the steps taken since (Adam moves a weight by at
  // most ~4 lr per step) could have grown the smallest layer maximum by half.
  t.fast_pending = false;
  t.fast_drift += 4.0f * fabsf(lr);
  const char* fast_env = getenv("WW_TRAIN_FAST");       // WW_TRAIN_FAST=0: host-side preparation every step (A/B, tests)
  if (t.tc_last && t.tc_version + 1 == c->weights_version && t.fast_steps < kFastSteps && t.fast_drift < t.fast_room &&
      !(fast_env && fast_env[0] == '0')) {
    if ((rc = ww_train_tc_repack(c, st))) return rc;
    t.tc_version = c->weights_version;
    t.fast_steps++;
    t.fast_pending = true;
  }
  if (!c->apply_event) WW_CHECK(c, cudaEventCreateWithFlags(&c->apply_event, cudaEventDisableTiming));
  WW_CHECK(c, cudaEventRecord(c->apply_event, st));     // ww_get_weights / ww_train_get_moments wait for it
  return WW_OK;
}

// ncclAllReduce through the caller's communicator, resolved at run time so that libwakeword_b200.so does not link NCCL
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);

int ww_train_step(ww_ctx* c, const float* x, const int64_t* labels, int B, float* loss, float lr, void* nccl_comm,
                  int world_size, void* stream) {
  int rc = ww_train_backward(c, x, labels, B, nullptr, nullptr, loss, nullptr, stream);
  if (rc) return rc;
  float scale = 1.0f;
  if (nccl_comm) {
    static nccl_allreduce_fn fn = nullptr;
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
      if (!fn) { c->set_error("ww_train_step: libnccl.so.2 / ncclAllReduce not found"); return WW_ERR_INVALID; }
    }
    // ncclFloat32 = 7, ncclSum = 0
    const int nrc = fn(c->train.grad, c->train.grad, (size_t)c->train.n_flat, 7, 0, nccl_comm, (cudaStream_t)stream);
    if (nrc != 0) { c->set_error("ww_train_step: ncclAllReduce failed with code " + std::to_string(nrc)); return WW_ERR_CUDA; }
    scale = 1.0f / (float)std::max(world_size, 1);
  }
  return ww_train_apply(c, lr, 0.9f, 0.999f, 1e-8f, 1e-5f, scale, stream);
}

}  // extern "C"

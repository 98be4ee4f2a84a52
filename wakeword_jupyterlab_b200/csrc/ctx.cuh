// Internal context of libwakeword_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/wakeword_b200.h"

#define WW_CHECK(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return WW_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define WW_LAUNCH_CHECK(ctx)                                                             \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));         \
      return WW_ERR_CUDA;                                                                \
    }                                                                                    \
    (ctx)->launches++;                                                                   \
  } while (0)

// Every C entry runs on its context's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct ResampleTable {
  int orig, neu;      // as passed by the caller
  int o, n;           // reduced by gcd
  int width, taps;
  int offset;         // float offset into d_rs_kern
  int nz;             // columns of the compact table (max non-zero taps over the phases)
};

// Device-visible descriptor of one prepared resample ratio.
struct RsDesc {
  int orig, neu, o, n, width, taps, offset, nz;   // device table at `offset`: [n][nz] taps, [n] first tap, [n] count
};

struct ProfSlot { cudaEvent_t a, b; int stage; };

// training-step state (train.cu): flat gradient / Adam buffers in state_dict order + saved activations
struct TrainState {
  std::vector<std::string> names;
  std::map<std::string, int64_t> offset, count;
  int64_t n_flat = 0, step = 0;
  float *grad = nullptr, *m = nullptr, *v = nullptr;
  float *wflip3 = nullptr, *wflip2 = nullptr, *part = nullptr, *loss = nullptr;
  void* tc = nullptr;               // tensor-core backward state (train_tc.cu)
  uint64_t tc_version = 0;          // weights_version its operand forms were built from
  bool tc_last = false, fast_pending = false;   // last backward ran the tensor-core kernels; last apply rebuilt their forms on the device
  int fast_steps = 0;               // device-side rebuilds since the last host-side preparation
  float fast_drift = 0.0f, fast_room = 0.0f;    // bound on how far a weight has moved since then / how far it may (half the smallest layer maximum)
  int cap = 0, cap32 = 0;           // clips the head buffers / the fp32 conv activations are sized for
  float *act1 = nullptr, *act2 = nullptr, *act3 = nullptr, *dact2 = nullptr, *dact1 = nullptr;
  float *pooled = nullptr, *dpooled = nullptr, *gates = nullptr, *hbuf = nullptr, *dh = nullptr;
  float *logits = nullptr, *dlogits = nullptr, *loss_row = nullptr;
};

struct ww_ctx {
  ww_config cfg;
  // optional per-stage CUDA-event timing (ww_profile / ww_profile_read)
  bool prof_on = false;
  std::vector<ProfSlot> prof_slots;
  size_t prof_used = 0;
  int device = 0;
  int sm_count = 148;
  int W = 0;            // frames per clip
  int n_bins = 0;       // n_fft/2 + 1
  int chunk = 0;        // clips per work chunk
  int64_t launches = 0;
  std::string err;

  // ---- log-mel tables
  float* d_window = nullptr;     // [n_fft] periodic Hann (zero-padded to n_fft), fp32
  float2* d_twiddle = nullptr;   // [n_fft] exp(-2 pi i t / n_fft)
  int* d_mel_start = nullptr;    // [n_mels] first non-zero bin
  int* d_mel_len = nullptr;      // [n_mels] number of bins
  int* d_mel_off = nullptr;      // [n_mels] offset into d_mel_w
  float* d_mel_w = nullptr;      // packed non-zero weights
  int mel_nnz = 0;
  std::vector<int> h_mel_start, h_mel_len;      // host copies (task table of the tensor-core log-mel kernel)
  // ---- tensor-core log-mel (logmel_tc.cu): DFT matrices as fp16 hi / lo in UMMA layouts, twiddles, mel task table
  int tc_lm_ready = 0;                          // 0 = not tried, 1 = ready, -1 = configuration outside the kernel's domain
  int tc_lm_tasks = 0;
  __half *d_tc_f32 = nullptr, *d_tc_f64hi = nullptr, *d_tc_f64lo = nullptr;
  float2 *d_tc_tw = nullptr, *d_tc_rot = nullptr;
  uint32_t* d_tc_tasks = nullptr;

  // ---- resample tables
  std::vector<ResampleTable> rs_tables;
  float* d_rs_kern = nullptr;    // packed [n phases][taps] tables
  RsDesc* d_rs_desc = nullptr;
  // running sums of squares of the noise bank (double, [rows][len + 1]), rebuilt by every API call that augments with it:
  // the energy of a clip's noise segment is then two loads instead of a pass over the segment (augment.cu)
  double* d_bank_prefix = nullptr;
  size_t bank_prefix_cap = 0;
  const float* bank_prefix_src = nullptr;
  int bank_prefix_rows = 0;
  int64_t bank_prefix_len = 0;
  int rs_kern_floats = 0, rs_kern_cap = 0, rs_desc_cap = 0;

  // ---- weights (fp32 masters, reference layouts)
  std::map<std::string, float*> w;          // name -> device fp32 copy
  std::map<std::string, std::vector<int64_t>> w_shape;
  bool weights_dirty = true;
  uint64_t weights_version = 1;     // bumped wherever weights_dirty is set (prepared forms remember the version they were built from)
  // prepared forms
  float* d_convw_t[3] = {nullptr, nullptr, nullptr};   // [Cin][9][Cout] fp32 (fp32 conv path, conv1 everywhere)
  float* d_head_wt[8] = {};                            // per LSTM layer: [K][3H] gate-interleaved (i,g,o), fp32
  float* d_head_b[8] = {};                             // per layer: [3H] b_ih + b_hh (i,g,o)
  float* d_bias_sum[8] = {};                           // per layer: [4H] b_ih + b_hh in reference row order (training)
  float* d_head_tc[8] = {};                            // per layer: TF32 hi | lo operand of the tensor-core head (head_tc.cu)
  uint64_t head_tc_version[8] = {};                    // weights_version each was packed from
  TrainState train;
  cudaEvent_t apply_event = nullptr;                    // recorded after the Adam kernels of ww_train_apply
  __half* d_w2_split = nullptr;                        // conv2 weights * 2^k, fp16 hi/lo, UMMA canonical layout
  __half* d_w3_split = nullptr;                        // conv3 weights * 2^k, fp16 hi/lo, UMMA canonical layout
  __half* d_w1_split = nullptr;                        // conv1 weights * 2^k, fp16 hi/lo, K = 9 padded to 16, twice
  float w1_inv_scale = 1.0f, w2_inv_scale = 1.0f, w3_inv_scale = 1.0f;      // 2^-k per layer
  float act1_scale = 1.0f, act2_scale = 1.0f;                               // power-of-two scales of the stored fp16 activations
  size_t c12_smem_conf = 0, c3_smem_conf = 0;   // dynamic shared-memory opt-in already made for this context's device
  int act2_lo_shift = 0;                                                    // e4m3 copy of act2 = fp16 copy x 2^-shift (conv12_tc.cu)

  // ---- workspaces (chunk clips)
  float* ws_clips = nullptr;     // [chunk][n_samples] augmented clips
  float* ws_logmel = nullptr;    // [chunk][n_mels][W]
  float* ws_logmel_pad = nullptr;   // tc path: [chunk][npix_in] zero-padded pixel-linear log-mel (padding zeroed once)
  float* ws_act1 = nullptr;      // fp32 path: [chunk][32][H][W]
  float* ws_act2 = nullptr;      // fp32 path: [chunk][64][H][W]
  __half* ws_act2_h = nullptr;   // tc path: [chunk][8 channel chunks][NPIX][8] fp16
  uint8_t* ws_act2_8 = nullptr;  // tc path: [chunk][4 channel chunks][NPIX][16] e4m3 (operand of the W_lo pass)
  __half* tc_act1_out = nullptr;      // training forward only: conv12 also stores conv1's output planes here (else null)
  uint32_t* tc_relu_bits = nullptr;   // training forward only: conv3 also stores the sign bits of its output here (else null)
  float* ws_pool_part = nullptr; // [pool_cap_clips][n_part][128]: conv3 partial sums of the whole batch of a call
  float* pool_cur = nullptr;     // where the current chunk's conv launch writes its partials
  float* ws_h[2] = {nullptr, nullptr};   // [pool_cap_clips][hidden]: head layer outputs (ping-pong)
  int64_t pool_cap_clips = 0;
  int pool_part_cap = 0;
  float* ws_logits = nullptr;    // [chunk][num_classes] (when caller passes NULL)
  int n_pool_part = 0;
  bool ws_ready = false;
  unsigned int* d_scalar = nullptr;   // scratch word for ww_normalize
  float2* d_pv_spec = nullptr; size_t pv_spec_bytes = 0;              // phase vocoder: STFT scratch [clips][frames][1025]
  float* d_stream_cache = nullptr; size_t stream_cache_bytes = 0;      // streaming: mel energies of the unique frames
  float* d_stream_bmax = nullptr; size_t stream_bmax_bytes = 0;        // streaming: block maxima of |x|
  uint32_t* d_tc_mask = nullptr;      // conv3 tile validity masks [T3][4]
  std::vector<float> h_w1t, h_b1, h_b2;   // host copies of conv1 weights [9][32], conv1/conv2 bias: passed as kernel parameters
  // host staging for ww_score_host
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;          // H2D copies of ww_score_host overlap the kernels
  std::vector<cudaEvent_t> copy_events;
  void* d_host_in = nullptr; size_t d_host_in_bytes = 0;
  void* d_host_out = nullptr; size_t d_host_out_bytes = 0;
  void* d_host_aug = nullptr; size_t d_host_aug_bytes = 0;

  void set_error(const std::string& s) { err = s; }
};

// RAII: CUDA events around one stage's launches on the launching stream (only when profiling is on).
struct ProfScope {
  ww_ctx* c; cudaStream_t st; ProfSlot* slot = nullptr;
  ProfScope(ww_ctx* c_, int stage, cudaStream_t st_) : c(c_), st(st_) {
    if (!c->prof_on) return;
    if (c->prof_used == c->prof_slots.size()) {
      ProfSlot s; s.stage = stage;
      cudaEventCreate(&s.a); cudaEventCreate(&s.b);
      c->prof_slots.push_back(s);
    }
    slot = &c->prof_slots[c->prof_used++];
    slot->stage = stage;
    cudaEventRecord(slot->a, st);
  }
  ~ProfScope() { if (slot) cudaEventRecord(slot->b, st); }
};

// Output addressing of the log-mel kernel: pixel (mel m, frame t) of clip b -> ptr[b * stride + off + m * pitch + t]
struct LogmelOut { float* ptr; int pitch; int off; int64_t stride; };

// ---- stage launchers (each returns WW_OK / error code; all enqueue on `st`)
// `clips` is fp32, or int16 PCM (x = s / 32768) when pcm16 != 0; strides are in samples
int ww_launch_logmel(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, float* out, int B, int normalize,
                     cudaStream_t st);                        // out: plain [B][1][n_mels][W]
int ww_launch_logmel_ex(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B,
                        int normalize, cudaStream_t st);
// streaming frame reuse (logmel.cu): mode 1 = build the cache of raw mel energies of frames starting every cache_g
// samples; mode 2 = windows whose interior frames are read from that cache
struct StreamReuse {
  int mode; float* cache; int64_t n_cache; int cache_g; int64_t abs_start0; const float* blockmax; int bm_block;
  int64_t n_total;
};
int ww_launch_logmel_stream(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B,
                            int normalize, const StreamReuse* sp, cudaStream_t st);
int ww_launch_logmel_tc(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B, int normalize,
                        cudaStream_t st);    // WW_OK = launched, 1 = outside its domain (use the FFT kernel), < 0 = error
int ww_launch_blockmax(ww_ctx* c, const void* x, int pcm16, int64_t n, int block, float* out, int64_t n_blocks,
                       cudaStream_t st);
int ww_launch_normalize(ww_ctx* c, const float* in, float* out, int64_t n, cudaStream_t st);
int ww_prepare_bank_energy(ww_ctx* c, const float* bank, int bank_rows, int64_t bank_len, cudaStream_t st);
int ww_launch_augment(ww_ctx* c, const void* clips, int pcm16, const float* bank, int bank_rows, int64_t bank_len,
                      const ww_aug* p, float* out, int B, cudaStream_t st);
int ww_launch_conv_fp32(ww_ctx* c, const float* logmel, int B, cudaStream_t st);   // -> ws_pool_part
int ww_launch_conv_tc(ww_ctx* c, const float* in_pad, int B, cudaStream_t st);     // padded log-mel -> ws_pool_part
int ww_launch_pad_logmel(ww_ctx* c, const float* logmel, float* in_pad, int B, cudaStream_t st);
LogmelOut ww_conv_tc_logmel_out(const ww_ctx* c, float* in_pad);
size_t ww_conv_tc_inpad_floats_per_clip(const ww_ctx* c);
int ww_launch_head(ww_ctx* c, int B, float* logits, float* prob1, uint8_t* decision, cudaStream_t st);
int ww_launch_gated_dense_tc(ww_ctx* c, int l, const float* x, float* out, int B, cudaStream_t st);   // 1: shape not handled
int ww_prepare_weights(ww_ctx* c, cudaStream_t st);
int ww_conv_tc_prepare(ww_ctx* c, cudaStream_t st);
size_t ww_conv_tc_act2_bytes_per_clip(const ww_ctx* c);
int ww_conv_tc_groups(const ww_ctx* c);
// training step, conv stack on the tensor cores (train_tc.cu)
bool ww_train_tc_supported(const ww_ctx* c);
int ww_train_tc_prepare(ww_ctx* c);
int ww_train_tc_forward(ww_ctx* c, const float* x, int B, float* pooled, cudaStream_t st);
int ww_train_tc_backward(ww_ctx* c, int B, const float* dpooled, float* gw1, float* gb1, float* gw2, float* gb2, float* gw3,
                         float* gb3, cudaStream_t st);
void ww_train_tc_free(ww_ctx* c);
int ww_train_tc_repack(ww_ctx* c, cudaStream_t st);      // after an optimiser step: operand forms rebuilt on the device
int ww_train_tc_sync_biases(ww_ctx* c);                  // before the next forward: conv biases -> host copies (kernel parameters)

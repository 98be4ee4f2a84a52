// C ABI of libwakeword_b200.so: context, tables, weights, stage orchestration.  See include/wakeword_b200.h.
#include "ctx.cuh"

#include <errno.h>
#include <math.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <algorithm>
#include <numeric>
#include <utility>

namespace {

std::string g_create_error;

// ------------------------------------------------------------------ slaney mel filterbank (librosa.filters.mel)
const double kFsp = 200.0 / 3.0, kMinLogHz = 1000.0, kMinLogMel = 1000.0 / (200.0 / 3.0);
double hz_to_mel(double f) {
  const double logstep = log(6.4) / 27.0;
  return f >= kMinLogHz ? kMinLogMel + log(f / kMinLogHz) / logstep : f / kFsp;
}
double mel_to_hz(double m) {
  const double logstep = log(6.4) / 27.0;
  return m >= kMinLogMel ? kMinLogHz * exp(logstep * (m - kMinLogMel)) : kFsp * m;
}

// dense [n_mels][n_bins] float32, same rounding points as librosa (float32 weights, float64 area norm)
std::vector<float> mel_filterbank(int sr, int n_fft, int n_mels, double fmin, double fmax) {
  const int n_bins = n_fft / 2 + 1;
  std::vector<double> mel_f(n_mels + 2);
  const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
  const double step = (m1 - m0) / (n_mels + 1);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : m0 + step * i);
  const double val = 1.0 / (n_fft * (1.0 / sr));
  std::vector<float> w((size_t)n_mels * n_bins, 0.0f);
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < n_bins; ++k) {
      const double fk = k * val;
      const double lower = -(mel_f[i] - fk) / fd0, upper = (mel_f[i + 2] - fk) / fd1;
      const float w32 = (float)std::max(0.0, std::min(lower, upper));
      w[(size_t)i * n_bins + k] = (float)((double)w32 * enorm);
    }
  }
  return w;
}

template <typename T>
int upload(ww_ctx* c, T** dst, const std::vector<T>& src) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  WW_CHECK(c, cudaMalloc((void**)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
  if (!src.empty()) WW_CHECK(c, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return WW_OK;
}

int build_tables(ww_ctx* c) {
  const ww_config& g = c->cfg;
  const int N = g.n_fft;
  // periodic Hann of win_length, centred zero-padding to n_fft (librosa util.pad_center)
  std::vector<float> win(N, 0.0f);
  const int lpad = (N - g.win_length) / 2;
  for (int n = 0; n < g.win_length; ++n)
    win[lpad + n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / g.win_length));
  int rc = upload(c, &c->d_window, win);
  if (rc) return rc;
  std::vector<float2> tw(N);
  for (int t = 0; t < N; ++t) {
    const double a = -2.0 * M_PI * t / N;
    tw[t] = make_float2((float)cos(a), (float)sin(a));
  }
  if ((rc = upload(c, &c->d_twiddle, tw))) return rc;
  std::vector<float> fb = mel_filterbank(g.sample_rate, N, g.n_mels, g.fmin, g.fmax);
  const int n_bins = N / 2 + 1;
  std::vector<int> start(g.n_mels), len(g.n_mels), off(g.n_mels);
  std::vector<float> packed;
  for (int m = 0; m < g.n_mels; ++m) {
    int lo = n_bins, hi = -1;
    for (int k = 0; k < n_bins; ++k)
      if (fb[(size_t)m * n_bins + k] != 0.0f) { lo = std::min(lo, k); hi = std::max(hi, k); }
    if (hi < lo) { lo = 0; hi = -1; }
    start[m] = lo; len[m] = hi - lo + 1; off[m] = (int)packed.size();
    for (int k = lo; k <= hi; ++k) packed.push_back(fb[(size_t)m * n_bins + k]);
  }
  c->mel_nnz = (int)packed.size();
  c->h_mel_start = start; c->h_mel_len = len;
  if ((rc = upload(c, &c->d_mel_start, start))) return rc;
  if ((rc = upload(c, &c->d_mel_len, len))) return rc;
  if ((rc = upload(c, &c->d_mel_off, off))) return rc;
  if ((rc = upload(c, &c->d_mel_w, packed))) return rc;
  return WW_OK;
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

int expected_shape(const ww_ctx* c, const std::string& name, std::vector<int64_t>* shape) {
  const int H = c->cfg.hidden_size;
  if (name == "conv1.weight") { *shape = {32, 1, 3, 3}; return 1; }
  if (name == "conv2.weight") { *shape = {64, 32, 3, 3}; return 1; }
  if (name == "conv3.weight") { *shape = {128, 64, 3, 3}; return 1; }
  if (name == "conv1.bias") { *shape = {32}; return 1; }
  if (name == "conv2.bias") { *shape = {64}; return 1; }
  if (name == "conv3.bias") { *shape = {128}; return 1; }
  if (name == "fc.weight") { *shape = {c->cfg.num_classes, H}; return 1; }
  if (name == "fc.bias") { *shape = {c->cfg.num_classes}; return 1; }
  for (int l = 0; l < c->cfg.num_layers; ++l) {
    const std::string s = std::to_string(l);
    if (name == "lstm.weight_ih_l" + s) { *shape = {4 * H, l == 0 ? 128 : H}; return 1; }
    if (name == "lstm.weight_hh_l" + s) { *shape = {4 * H, H}; return 1; }
    if (name == "lstm.bias_ih_l" + s || name == "lstm.bias_hh_l" + s) { *shape = {4 * H}; return 1; }
  }
  return 0;
}

std::vector<std::string> required_weights(const ww_ctx* c) {
  std::vector<std::string> r = {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
                                "conv3.weight", "conv3.bias", "fc.weight", "fc.bias"};
  for (int l = 0; l < c->cfg.num_layers; ++l) {
    const std::string s = std::to_string(l);
    r.push_back("lstm.weight_ih_l" + s);
    r.push_back("lstm.bias_ih_l" + s);
    r.push_back("lstm.bias_hh_l" + s);   // weight_hh never reaches the output (T = 1, h0 = 0): optional
  }
  return r;
}

std::vector<float> fetch(ww_ctx* c, const std::string& name) {
  const std::vector<int64_t>& s = c->w_shape[name];
  size_t n = 1;
  for (int64_t d : s) n *= (size_t)d;
  std::vector<float> h(n);
  cudaMemcpy(h.data(), c->w[name], n * sizeof(float), cudaMemcpyDeviceToHost);
  return h;
}

int free_all(ww_ctx* c) {
  cudaFree(c->d_window); cudaFree(c->d_twiddle); cudaFree(c->d_mel_start); cudaFree(c->d_mel_len);
  cudaFree(c->d_mel_off); cudaFree(c->d_mel_w); cudaFree(c->d_rs_kern); cudaFree(c->d_rs_desc);
  cudaFree(c->d_tc_f32); cudaFree(c->d_tc_f64hi); cudaFree(c->d_tc_f64lo); cudaFree(c->d_tc_tw); cudaFree(c->d_tc_rot); cudaFree(c->d_tc_tasks);
  for (auto& kv : c->w) cudaFree(kv.second);
  for (int i = 0; i < 3; ++i) cudaFree(c->d_convw_t[i]);
  for (int i = 0; i < 8; ++i) { cudaFree(c->d_head_wt[i]); cudaFree(c->d_head_b[i]); cudaFree(c->d_bias_sum[i]); cudaFree(c->d_head_tc[i]); }
  cudaFree(c->d_bank_prefix);
  {
    TrainState& t = c->train;
    float* bufs[] = {t.grad, t.m, t.v, t.wflip3, t.wflip2, t.part, t.loss, t.act1, t.act2, t.act3, t.dact2, t.dact1, t.pooled,
                     t.dpooled, t.gates, t.hbuf, t.dh, t.logits, t.dlogits, t.loss_row};
    for (float* b : bufs) cudaFree(b);
    ww_train_tc_free(c);
  }
  cudaFree(c->d_w1_split); cudaFree(c->d_w2_split); cudaFree(c->d_w3_split); cudaFree(c->ws_logmel_pad);
  cudaFree(c->ws_clips); cudaFree(c->ws_logmel); cudaFree(c->ws_act1); cudaFree(c->ws_act2);
  cudaFree(c->ws_act2_h); cudaFree(c->ws_act2_8); cudaFree(c->ws_pool_part); cudaFree(c->ws_logits); cudaFree(c->ws_h[0]); cudaFree(c->ws_h[1]);
  cudaFree(c->d_pv_spec); cudaFree(c->d_scalar); cudaFree(c->d_stream_cache); cudaFree(c->d_stream_bmax); cudaFree(c->d_tc_mask); cudaFree(c->d_host_in); cudaFree(c->d_host_out); cudaFree(c->d_host_aug);
  for (ProfSlot& p : c->prof_slots) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
  if (c->apply_event) cudaEventDestroy(c->apply_event);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  return 0;
}

// workspaces are allocated at the first forward/score call so that log-mel-only contexts stay small
int ensure_workspaces(ww_ctx* c) {
  if (c->ws_ready) return WW_OK;
  const ww_config& g = c->cfg;
  const int H = g.n_mels, W = c->W;
  WW_CHECK(c, cudaMalloc((void**)&c->ws_clips, (size_t)c->chunk * g.n_samples * 4));
  WW_CHECK(c, cudaMalloc((void**)&c->ws_logmel, (size_t)c->chunk * H * W * 4));
  WW_CHECK(c, cudaMalloc((void**)&c->ws_logits, (size_t)c->chunk * g.num_classes * 4));
  if (g.conv_mode == WW_CONV_FP32) {
    WW_CHECK(c, cudaMalloc((void**)&c->ws_act1, (size_t)c->chunk * 32 * H * W * 4));
    WW_CHECK(c, cudaMalloc((void**)&c->ws_act2, (size_t)c->chunk * 64 * H * W * 4));
  } else {
    WW_CHECK(c, cudaMalloc((void**)&c->ws_act2_h, (size_t)c->chunk * ww_conv_tc_act2_bytes_per_clip(c)));
    WW_CHECK(c, cudaMalloc((void**)&c->ws_act2_8, (size_t)c->chunk * ww_conv_tc_act2_bytes_per_clip(c) / 2));
    const size_t pad_bytes = (size_t)c->chunk * ww_conv_tc_inpad_floats_per_clip(c) * 4;
    WW_CHECK(c, cudaMalloc((void**)&c->ws_logmel_pad, pad_bytes));
    WW_CHECK(c, cudaMemset(c->ws_logmel_pad, 0, pad_bytes));      // the padding stays zero: kernels only write pixels
  }
  c->ws_ready = true;
  return WW_OK;
}

int ensure_buffer(ww_ctx* c, void** buf, size_t* cap, size_t need);

// pool partials and head activations cover the WHOLE batch of a call (the head runs once per call, not per chunk)
int ensure_pool(ww_ctx* c, int64_t B) {
  if (B <= c->pool_cap_clips) return WW_OK;
  const int H = c->cfg.n_mels, W = c->W;
  const int tiles_fp32 = ((W + 31) / 32) * ((H + 7) / 8);
  const int tiles_tc = ((H + 2) * (W + 2) + 127) / 128 + 1;
  const size_t part_cap = (size_t)std::max(tiles_fp32, tiles_tc);
  const int64_t cap = std::max<int64_t>(B, c->chunk);
  size_t dummy = 0;
  int rc;
  if ((rc = ensure_buffer(c, (void**)&c->ws_pool_part, &dummy, (size_t)cap * part_cap * 128 * 4))) return rc;
  for (int i = 0; i < 2; ++i) {
    dummy = 0;
    if ((rc = ensure_buffer(c, (void**)&c->ws_h[i], &dummy, (size_t)cap * std::max(c->cfg.hidden_size, 128) * 4))) return rc;
  }
  c->pool_cap_clips = cap;
  c->pool_part_cap = (int)part_cap;
  return WW_OK;
}

int ensure_buffer(ww_ctx* c, void** buf, size_t* cap, size_t need) {
  if (*cap >= need) return WW_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr; *cap = 0;
  WW_CHECK(c, cudaMalloc(buf, need));
  *cap = need;
  return WW_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int ww_abi_version(void) { return WW_ABI_VERSION; }

const char* ww_last_error(const ww_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ww_n_frames(const ww_ctx* ctx) { return ctx ? ctx->W : 0; }
int64_t ww_kernel_launches(const ww_ctx* ctx) { return ctx ? ctx->launches : 0; }
int ww_conv_mode(const ww_ctx* ctx) { return ctx ? ctx->cfg.conv_mode : -1; }

int ww_set_threshold(ww_ctx* c, float threshold) {
  if (!c || !(threshold >= 0.0f && threshold <= 1.0f)) return WW_ERR_INVALID;
  c->cfg.threshold = threshold;          // read by the head kernel's parameters at the next launch
  return WW_OK;
}

int ww_profile(ww_ctx* c, int enable) {
  if (!c) return WW_ERR_INVALID;
  c->prof_on = enable != 0;
  return WW_OK;
}

int ww_profile_read(ww_ctx* c, int stage, double* total_ms, int64_t* n) {
  if (!c) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  WW_CHECK(c, cudaDeviceSynchronize());
  if (stage < 0) { c->prof_used = 0; return WW_OK; }
  double tot = 0.0; int64_t cnt = 0;
  for (size_t i = 0; i < c->prof_used; ++i) {
    if (c->prof_slots[i].stage != stage) continue;
    float ms = 0.0f;
    WW_CHECK(c, cudaEventElapsedTime(&ms, c->prof_slots[i].a, c->prof_slots[i].b));
    tot += ms; ++cnt;
  }
  if (total_ms) *total_ms = tot;
  if (n) *n = cnt;
  return WW_OK;
}

int ww_create(ww_ctx** out, int device, const ww_config* cfg) {
  if (!out || !cfg) { g_create_error = "ww_create: null argument"; return WW_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("ww_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return WW_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "ww_create: bad device index"; return WW_ERR_INVALID; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "ww_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                     ", this library is built for sm_100a only";
    return WW_ERR_ARCH;
  }
  const ww_config& g = *cfg;
  if (!is_pow2(g.n_fft) || g.n_fft < 256 || g.n_fft > 2048 || g.win_length <= 0 || g.win_length > g.n_fft ||
      g.hop_length <= 0 || g.n_samples <= 0 || g.n_mels <= 0 || g.n_mels > 256 || g.sample_rate <= 0 ||
      g.hidden_size <= 0 || g.hidden_size % 32 != 0 || g.hidden_size > 1024 || g.num_layers < 1 ||
      g.num_layers > 8 || g.num_classes < 1 || g.num_classes > 16 || g.fmax <= g.fmin ||
      g.conv_mode < WW_CONV_SPLIT2 || g.conv_mode > WW_CONV_FP16) {
    g_create_error = "ww_create: unsupported configuration";
    return WW_ERR_INVALID;
  }
  DeviceGuard dev_guard(device);
  ww_ctx* c = new ww_ctx();
  c->cfg = g;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->W = 1 + g.n_samples / g.hop_length;
  c->n_bins = g.n_fft / 2 + 1;
  c->chunk = g.chunk_clips > 0 ? g.chunk_clips : 8192;   // device-resident batches: fewer, longer launches (+1.5 % over 4,096)
  if ((size_t)g.n_mels * c->W * 4 + (size_t)(5 * g.n_fft + g.n_fft / 4 + 4 * (g.n_fft >> 5) + 48) * 8 > 216 * 1024) {
    g_create_error = "ww_create: n_mels x frames too large for the log-mel kernel's shared memory";
    delete c;
    return WW_ERR_INVALID;
  }
  int rc = build_tables(c);
  if (rc == WW_OK && cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    c->set_error("ww_create: stream creation failed");
    rc = WW_ERR_CUDA;
  }
  if (rc != WW_OK) {
    g_create_error = c->err;
    free_all(c);
    delete c;
    return rc;
  }
  *out = c;
  return WW_OK;
}

void ww_destroy(ww_ctx* c) {
  if (!c) return;
  DeviceGuard dev_guard(c->device);
  cudaDeviceSynchronize();
  free_all(c);
  delete c;
}

int ww_set_weights(ww_ctx* c, const char* name, const float* src, const int64_t* shape, int ndim) {
  if (!c || !name || !src || !shape) return WW_ERR_INVALID;
  std::vector<int64_t> want;
  if (!expected_shape(c, name, &want)) { c->set_error(std::string("ww_set_weights: unknown parameter ") + name); return WW_ERR_INVALID; }
  std::vector<int64_t> got(shape, shape + ndim);
  if (got != want) { c->set_error(std::string("ww_set_weights: shape mismatch for ") + name); return WW_ERR_INVALID; }
  size_t n = 1;
  for (int64_t d : want) n *= (size_t)d;
  DeviceGuard dev_guard(c->device);
  float*& dst = c->w[name];
  if (!dst) WW_CHECK(c, cudaMalloc((void**)&dst, n * sizeof(float)));
  WW_CHECK(c, cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDefault));
  c->w_shape[name] = want;
  c->weights_dirty = true;
  c->weights_version++;
  return WW_OK;
}

int ww_prepare_resample(ww_ctx* c, int orig, int neu) {
  if (!c || orig <= 0 || neu <= 0) return WW_ERR_INVALID;
  for (const ResampleTable& t : c->rs_tables)
    if (t.orig == orig && t.neu == neu) return WW_OK;
  DeviceGuard dev_guard(c->device);
  const int g = std::gcd(orig, neu);
  ResampleTable t;
  t.orig = orig; t.neu = neu; t.o = orig / g; t.n = neu / g;
  const double lpw = 6.0, rolloff = 0.99;
  const double base = std::min(t.o, t.n) * rolloff;
  t.width = (int)ceil(lpw * t.o / base);
  t.taps = 2 * t.width + t.o;
  t.offset = (c->rs_kern_floats + 3) & ~3;             // table rows are read as float4: 16-byte aligned base
  std::vector<float> k((size_t)t.n * t.taps);
  for (int p = 0; p < t.n; ++p)
    for (int i = 0; i < t.taps; ++i) {
      double tt = ((double)(-p) / t.n + (double)(i - t.width) / t.o) * base;
      tt = std::max(-lpw, std::min(lpw, tt));
      const double wdw = pow(cos(tt * M_PI / lpw / 2.0), 2.0);
      tt *= M_PI;
      const double s = (tt == 0.0) ? 1.0 : sin(tt) / tt;
      k[(size_t)p * t.taps + i] = (float)(s * wdw * (base / t.o));
    }
  // Per-phase range of non-zero taps, appended as int bit patterns ([n] first tap, [n] end tap).  Taps outside
  // the +-6 zero-crossing window are exactly 0.0f (their double value underflows), so skipping them leaves
  // every output bit-identical while cutting e.g. 111 taps to ~14 for 97 -> 100.
  // Device form: compact [n][nz] table of the non-zero taps + [n] first-tap index + [n] count (int bit patterns).
  {
    std::vector<int> lo(t.n), cnt(t.n);
    int nz = 1;
    for (int p = 0; p < t.n; ++p) {
      int l = t.taps, h = 0;
      for (int i = 0; i < t.taps; ++i)
        if (k[(size_t)p * t.taps + i] != 0.0f) { l = std::min(l, i); h = std::max(h, i + 1); }
      if (h <= l) { l = 0; h = 0; }
      lo[p] = l; cnt[p] = h - l;
      nz = std::max(nz, h - l);
    }
    nz = (nz + 3) & ~3;                                  // rows zero-padded to whole float4 groups
    // row pitch = 4 (mod 8) words: the 128-bit row loads of 8 consecutive phases then hit 8 different bank groups
    const int pitch = nz + ((nz & 4) ? 0 : 4);
    std::vector<float> comp((((size_t)t.n * pitch + 2 * t.n) + 3) & ~(size_t)3, 0.0f);
    for (int p = 0; p < t.n; ++p) {
      for (int i = 0; i < cnt[p]; ++i) comp[(size_t)p * pitch + i] = k[(size_t)p * t.taps + lo[p] + i];
      memcpy(&comp[(size_t)t.n * pitch + p], &lo[p], 4);
      memcpy(&comp[(size_t)t.n * pitch + t.n + p], &cnt[p], 4);
    }
    t.nz = nz;
    k.swap(comp);
  }
  const int need = t.offset + (int)k.size();
  if (need > c->rs_kern_cap) {
    int cap = std::max(need * 2, 1 << 16);
    float* nb = nullptr;
    WW_CHECK(c, cudaMalloc((void**)&nb, (size_t)cap * sizeof(float)));
    if (c->d_rs_kern) {
      WW_CHECK(c, cudaDeviceSynchronize());
      WW_CHECK(c, cudaMemcpy(nb, c->d_rs_kern, (size_t)c->rs_kern_floats * sizeof(float), cudaMemcpyDeviceToDevice));
      cudaFree(c->d_rs_kern);
    }
    c->d_rs_kern = nb; c->rs_kern_cap = cap;
  }
  WW_CHECK(c, cudaMemcpy(c->d_rs_kern + t.offset, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice));
  c->rs_kern_floats = need;
  c->rs_tables.push_back(t);
  if ((int)c->rs_tables.size() > c->rs_desc_cap) {
    WW_CHECK(c, cudaDeviceSynchronize());
    if (c->d_rs_desc) cudaFree(c->d_rs_desc);
    c->rs_desc_cap = std::max(64, (int)c->rs_tables.size() * 2);
    WW_CHECK(c, cudaMalloc((void**)&c->d_rs_desc, (size_t)c->rs_desc_cap * sizeof(RsDesc)));
  }
  std::vector<RsDesc> d(c->rs_tables.size());
  for (size_t i = 0; i < d.size(); ++i) {
    const ResampleTable& r = c->rs_tables[i];
    d[i] = RsDesc{r.orig, r.neu, r.o, r.n, r.width, r.taps, r.offset, r.nz};
  }
  WW_CHECK(c, cudaMemcpy(c->d_rs_desc, d.data(), d.size() * sizeof(RsDesc), cudaMemcpyHostToDevice));
  return WW_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// weight preparation (host side; weights total <= 4 MB)
int ww_prepare_weights(ww_ctx* c, cudaStream_t st) {
  if (!c->weights_dirty) return WW_OK;
  for (const std::string& n : required_weights(c))
    if (!c->w.count(n)) { c->set_error("forward: weight not set: " + n); return WW_ERR_WEIGHTS; }
  WW_CHECK(c, cudaStreamSynchronize(st));
  const int cins[3] = {1, 32, 64}, couts[3] = {32, 64, 128};
  for (int l = 0; l < 3; ++l) {
    std::vector<float> w = fetch(c, "conv" + std::to_string(l + 1) + ".weight");   // [Cout][Cin][3][3]
    std::vector<float> t((size_t)cins[l] * 9 * couts[l]);
    for (int co = 0; co < couts[l]; ++co)
      for (int ci = 0; ci < cins[l]; ++ci)
        for (int k = 0; k < 9; ++k) t[((size_t)ci * 9 + k) * couts[l] + co] = w[((size_t)co * cins[l] + ci) * 9 + k];
    int rc = upload(c, &c->d_convw_t[l], t);
    if (rc) return rc;
    if (l == 0) { c->h_w1t = t; c->h_b1 = fetch(c, "conv1.bias"); c->h_b2 = fetch(c, "conv2.bias"); }   // -> kernel parameters (constant bank)
  }
  const int H = c->cfg.hidden_size;
  for (int l = 0; l < c->cfg.num_layers; ++l) {
    const std::string s = std::to_string(l);
    const int K = l == 0 ? 128 : H;
    std::vector<float> w = fetch(c, "lstm.weight_ih_l" + s);    // [4H][K], rows i,f,g,o
    std::vector<float> bi = fetch(c, "lstm.bias_ih_l" + s), bh = fetch(c, "lstm.bias_hh_l" + s);
    std::vector<float> wt((size_t)K * 3 * H), b((size_t)3 * H);
    const int gate_row[3] = {0, 2, 3};                          // i, g, o
    for (int g = 0; g < 3; ++g)
      for (int j = 0; j < H; ++j) {
        const int row = gate_row[g] * H + j;
        b[(size_t)g * H + j] = bi[row] + bh[row];
        for (int k = 0; k < K; ++k) wt[((size_t)k * 3 + g) * H + j] = w[(size_t)row * K + k];
      }
    int rc = upload(c, &c->d_head_wt[l], wt);
    if (rc) return rc;
    if ((rc = upload(c, &c->d_head_b[l], b))) return rc;
    std::vector<float> bs((size_t)4 * H);
    for (int r = 0; r < 4 * H; ++r) bs[r] = bi[r] + bh[r];
    if ((rc = upload(c, &c->d_bias_sum[l], bs))) return rc;
  }
  if (c->cfg.conv_mode != WW_CONV_FP32) {
    int rc = ww_conv_tc_prepare(c, st);
    if (rc) return rc;
  }
  c->weights_dirty = false;
  return WW_OK;
}

namespace {

int pool_parts(const ww_ctx* c) {
  if (c->cfg.conv_mode == WW_CONV_FP32) return ((c->W + 31) / 32) * ((c->cfg.n_mels + 7) / 8);
  return ww_conv_tc_groups(c);
}

// conv stack of one chunk; its pool partials land at clip offset `pool_off` of the batch-wide buffer.
// fp32 mode: `logmel` is the plain [B][H][W] image; tensor-core modes: the padded image in ws_logmel_pad.
int conv_chunk(ww_ctx* c, const float* logmel, int B, int64_t pool_off, cudaStream_t st) {
  c->pool_cur = c->ws_pool_part + (size_t)pool_off * pool_parts(c) * 128;
  return (c->cfg.conv_mode == WW_CONV_FP32) ? ww_launch_conv_fp32(c, logmel, B, st) : ww_launch_conv_tc(c, logmel, B, st);
}

}  // namespace

extern "C" {

int ww_augment(ww_ctx* c, const float* clips, const float* bank, int bank_rows, int64_t bank_len,
               const ww_aug* p, float* out, int B, void* stream) {
  if (!c || !clips || !p || !out || B < 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  if (int rc = ww_prepare_bank_energy(c, bank, bank_rows, bank_len, (cudaStream_t)stream)) return rc;
  return ww_launch_augment(c, clips, 0, bank, bank_rows, bank_len, p, out, B, (cudaStream_t)stream);
}

int ww_augment_pcm16(ww_ctx* c, const int16_t* clips, const float* bank, int bank_rows, int64_t bank_len,
                     const ww_aug* p, float* out, int B, void* stream) {
  if (!c || !clips || !p || !out || B < 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  if (int rc = ww_prepare_bank_energy(c, bank, bank_rows, bank_len, (cudaStream_t)stream)) return rc;
  return ww_launch_augment(c, clips, 1, bank, bank_rows, bank_len, p, out, B, (cudaStream_t)stream);
}

int ww_normalize(ww_ctx* c, const float* in, float* out, int64_t n, void* stream) {
  if (!c || !in || !out || n < 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  return ww_launch_normalize(c, in, out, n, (cudaStream_t)stream);
}

int ww_logmel(ww_ctx* c, const float* clips, int64_t clip_stride, float* out, int B, int normalize, void* stream) {
  if (!c || !clips || !out || B < 0 || clip_stride <= 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  return ww_launch_logmel(c, clips, 0, clip_stride, out, B, normalize, (cudaStream_t)stream);
}

int ww_logmel_pcm16(ww_ctx* c, const int16_t* clips, int64_t clip_stride, float* out, int B, int normalize, void* stream) {
  if (!c || !clips || !out || B < 0 || clip_stride <= 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  return ww_launch_logmel(c, clips, 1, clip_stride, out, B, normalize, (cudaStream_t)stream);
}

int ww_forward(ww_ctx* c, const float* logmel, float* logits, int B, void* stream) {
  if (!c || !logmel || !logits || B < 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_workspaces(c);
  if (rc) return rc;
  if ((rc = ww_prepare_weights(c, st))) return rc;
  if ((rc = ensure_pool(c, B))) return rc;
  const size_t per = (size_t)c->cfg.n_mels * c->W;
  for (int b0 = 0; b0 < B; b0 += c->chunk) {
    const int nb = std::min(c->chunk, B - b0);
    const float* src = logmel + (size_t)b0 * per;
    if (c->cfg.conv_mode != WW_CONV_FP32) {
      if ((rc = ww_launch_pad_logmel(c, src, c->ws_logmel_pad, nb, st))) return rc;
      src = c->ws_logmel_pad;
    }
    if ((rc = conv_chunk(c, src, nb, b0, st))) return rc;
  }
  return ww_launch_head(c, B, logits, nullptr, nullptr, st);
}

// (augment) -> log-mel -> conv stack for clips [0, B) whose pool partials go to clip offset pool_off; the head
// (whole batch) runs only when run_head is set, over clips [0, pool_off + B).
static int score_impl(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, const float* bank, int bank_rows,
                      int64_t bank_len, const ww_aug* aug, int normalize, float* logits, float* prob1,
                      uint8_t* decision, int64_t B, cudaStream_t st, int64_t pool_off = 0, bool run_head = true,
                      const StreamReuse* sr = nullptr) {
  int rc = ensure_workspaces(c);
  if (rc) return rc;
  if ((rc = ww_prepare_weights(c, st))) return rc;
  // ensure_pool re-allocates without keeping contents: a caller that scores a batch in pieces (pool_off > 0) must
  // have sized the pool for the whole batch before its first piece
  if (pool_off > 0 && pool_off + B > c->pool_cap_clips) {
    c->set_error("score: pool partials not sized for the whole batch before a piece-wise call");
    return WW_ERR_INVALID;
  }
  if ((rc = ensure_pool(c, pool_off + B))) return rc;
  const int N = c->cfg.n_samples;
  for (int64_t b0 = 0; b0 < B; b0 += c->chunk) {
    const int nb = (int)std::min<int64_t>(c->chunk, B - b0);
    const void* src = static_cast<const char*>(clips) + b0 * clip_stride * (pcm16 ? 2 : 4);
    int64_t stride = clip_stride;
    int in16 = pcm16;
    if (aug) {
      ww_aug a = *aug;
      a.flags += b0; a.shift += b0; a.rs_orig += b0; a.rs_new += b0; a.crop_off += b0;
      a.noise_idx += b0; a.noise_off += b0; a.snr_db += b0; a.gain += b0;
      if ((rc = ww_launch_augment(c, src, in16, bank, bank_rows, bank_len, &a, c->ws_clips, nb, st))) return rc;
      src = c->ws_clips;
      stride = N;
      in16 = 0;
    }
    const bool tc = c->cfg.conv_mode != WW_CONV_FP32;
    const LogmelOut lo = tc ? ww_conv_tc_logmel_out(c, c->ws_logmel_pad)
                            : LogmelOut{c->ws_logmel, c->W, 0, (int64_t)c->cfg.n_mels * c->W};
    if (sr) {                                         // streaming windows with cached interior frames
      StreamReuse w = *sr;
      w.abs_start0 = b0 * clip_stride;
      if ((rc = ww_launch_logmel_stream(c, src, in16, stride, lo, nb, 1, &w, st))) return rc;
    } else if ((rc = ww_launch_logmel_ex(c, src, in16, stride, lo, nb, aug ? 0 : normalize, st))) return rc;
    if ((rc = conv_chunk(c, tc ? c->ws_logmel_pad : c->ws_logmel, nb, pool_off + b0, st))) return rc;
  }
  if (!run_head) return WW_OK;
  if (pool_off + B > 0x7fffffff) { c->set_error("score: batch too large"); return WW_ERR_INVALID; }
  return ww_launch_head(c, (int)(pool_off + B), logits, prob1, decision, st);
}

static int score_entry(ww_ctx* c, const void* clips, int pcm16, const float* bank, int bank_rows, int64_t bank_len,
                       const ww_aug* aug, int normalize, float* logits, float* prob1, uint8_t* decision, int B,
                       void* stream) {
  if (!c || !clips || B < 0) return WW_ERR_INVALID;
  if (aug && (!aug->flags || !aug->shift || !aug->rs_orig || !aug->rs_new || !aug->crop_off || !aug->noise_idx ||
              !aug->noise_off || !aug->snr_db || !aug->gain)) {
    c->set_error("ww_score: ww_aug has null arrays");
    return WW_ERR_INVALID;
  }
  DeviceGuard dev_guard(c->device);
  if (aug)
    if (int rc = ww_prepare_bank_energy(c, bank, bank_rows, bank_len, (cudaStream_t)stream)) return rc;
  return score_impl(c, clips, pcm16, c->cfg.n_samples, bank, bank_rows, bank_len, aug, normalize, logits, prob1,
                    decision, B, (cudaStream_t)stream);
}

int ww_score(ww_ctx* c, const float* clips, const float* bank, int bank_rows, int64_t bank_len, const ww_aug* aug,
             int normalize, float* logits, float* prob1, uint8_t* decision, int B, void* stream) {
  return score_entry(c, clips, 0, bank, bank_rows, bank_len, aug, normalize, logits, prob1, decision, B, stream);
}

int ww_score_pcm16(ww_ctx* c, const int16_t* clips, const float* bank, int bank_rows, int64_t bank_len,
                   const ww_aug* aug, int normalize, float* logits, float* prob1, uint8_t* decision, int B,
                   void* stream) {
  return score_entry(c, clips, 1, bank, bank_rows, bank_len, aug, normalize, logits, prob1, decision, B, stream);
}

static int stream_entry(ww_ctx* c, const void* audio, int pcm16, int64_t T, int hop_samples, float* prob1,
                        uint8_t* decision, int64_t n_win, void* stream) {
  if (!c || !audio || hop_samples <= 0 || n_win < 0) return WW_ERR_INVALID;
  if (n_win > 0 && (n_win - 1) * hop_samples + c->cfg.n_samples > T) {
    c->set_error("ww_score_stream: windows exceed the audio length");
    return WW_ERR_INVALID;
  }
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  // Frame reuse: windows k and k' share every STFT frame that starts at the same sample and
  // lies inside both, and peak normalisation only scales the power spectrum.  The mel energies of all frames starting
  // at multiples of g = gcd(window hop, frame hop) are computed once; a window then transforms only the frames that
  // touch its zero padding.  Used when it at least halves the number of transforms.
  const int N = c->cfg.n_fft, hop = c->cfg.hop_length, ns = c->cfg.n_samples;
  const int g = std::gcd(hop_samples, hop);
  const int t_lo = (N / 2 + hop - 1) / hop, t_hi = (ns - N / 2) / hop;
  const int64_t n_cache = T >= N ? (T - N) / g + 1 : 0;
  const int64_t n_interior = (int64_t)std::max(0, t_hi - t_lo + 1) * n_win;
  const bool reuse = getenv("WW_STREAM_NO_REUSE") == nullptr && n_win > 1 && (N / 2) % g == 0 && ns % hop_samples == 0 &&
                     T < ((int64_t)1 << 30) * (pcm16 ? 2 : 1) && n_cache * 2 <= n_interior;
  if (!reuse)
    return score_impl(c, audio, pcm16, hop_samples, nullptr, 0, 0, nullptr, 1, nullptr, prob1, decision, n_win, st);
  int rc;
  const int64_t n_bm = (T + hop_samples - 1) / hop_samples;
  if ((rc = ensure_buffer(c, (void**)&c->d_stream_cache, &c->stream_cache_bytes, (size_t)n_cache * c->cfg.n_mels * 4))) return rc;
  if ((rc = ensure_buffer(c, (void**)&c->d_stream_bmax, &c->stream_bmax_bytes, (size_t)n_bm * 4))) return rc;
  if ((rc = ww_launch_blockmax(c, audio, pcm16, T, hop_samples, c->d_stream_bmax, n_bm, st))) return rc;
  StreamReuse sr;
  sr.mode = 1; sr.cache = c->d_stream_cache; sr.n_cache = n_cache; sr.cache_g = g; sr.abs_start0 = 0;
  sr.blockmax = c->d_stream_bmax; sr.bm_block = hop_samples; sr.n_total = T;
  const int W = c->W;
  const int64_t n_blocks = (n_cache + W - 1) / W;
  for (int64_t blk0 = 0; blk0 < n_blocks; blk0 += (1 << 20)) {          // launches of at most 2^20 frame blocks
    const int nb = (int)std::min<int64_t>(1 << 20, n_blocks - blk0);
    StreamReuse s1 = sr;
    s1.cache = sr.cache + (size_t)blk0 * W * c->cfg.n_mels;
    s1.n_cache = n_cache - blk0 * W;
    s1.n_total = T - blk0 * W * g;
    const void* src = static_cast<const char*>(audio) + (size_t)blk0 * W * g * (pcm16 ? 2 : 4);
    if ((rc = ww_launch_logmel_stream(c, src, pcm16, (int64_t)W * g, LogmelOut{nullptr, 0, 0, 0}, nb, 0, &s1, st))) return rc;
  }
  sr.mode = 2;
  return score_impl(c, audio, pcm16, hop_samples, nullptr, 0, 0, nullptr, 1, nullptr, prob1, decision, n_win, st, 0, true, &sr);
}

int ww_score_stream(ww_ctx* c, const float* audio, int64_t T, int hop_samples, float* prob1, uint8_t* decision,
                    int64_t n_win, void* stream) {
  return stream_entry(c, audio, 0, T, hop_samples, prob1, decision, n_win, stream);
}

int ww_score_stream_pcm16(ww_ctx* c, const int16_t* audio, int64_t T, int hop_samples, float* prob1,
                          uint8_t* decision, int64_t n_win, void* stream) {
  return stream_entry(c, audio, 1, T, hop_samples, prob1, decision, n_win, stream);
}

constexpr double kH2DRamp = 1.3;     // growth of the host->device pieces (WW_H2D_RAMP overrides, for tuning)
constexpr int kH2DCap = 8192;        // largest piece on the host path (<= the work chunk)
constexpr int kH2DFirst = 384;       // clips in the first piece (its copy is exposed)

static int score_host_impl(ww_ctx* c, const void* clips_host, int pcm16, const float* bank_dev, int bank_rows,
                           int64_t bank_len, const ww_aug* aug_host, int normalize, float* logits_host,
                           float* prob1_host, uint8_t* decision_host, int B) {
  if (!c || !clips_host || B < 0) return WW_ERR_INVALID;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = c->own_stream;
  const int N = c->cfg.n_samples, C = c->cfg.num_classes;
  const size_t esz = pcm16 ? 2 : 4;
  int rc;
  if ((rc = ensure_buffer(c, &c->d_host_in, &c->d_host_in_bytes, (size_t)B * N * esz))) return rc;
  const size_t out_bytes = (size_t)B * (C * 4 + 4 + 1);
  if ((rc = ensure_buffer(c, &c->d_host_out, &c->d_host_out_bytes, out_bytes))) return rc;
  float* d_logits = (float*)c->d_host_out;
  float* d_prob = d_logits + (size_t)B * C;
  uint8_t* d_dec = (uint8_t*)(d_prob + B);
  ww_aug a_dev = {};
  if (aug_host) {
    if ((rc = ensure_buffer(c, &c->d_host_aug, &c->d_host_aug_bytes, (size_t)B * 9 * 4))) return rc;
    uint32_t* base = (uint32_t*)c->d_host_aug;
    const void* srcs[9] = {aug_host->flags, aug_host->shift, aug_host->rs_orig, aug_host->rs_new, aug_host->crop_off,
                           aug_host->noise_idx, aug_host->noise_off, aug_host->snr_db, aug_host->gain};
    for (int i = 0; i < 9; ++i)
      WW_CHECK(c, cudaMemcpyAsync(base + (size_t)i * B, srcs[i], (size_t)B * 4, cudaMemcpyHostToDevice, st));
    a_dev.flags = base; a_dev.shift = (int32_t*)(base + (size_t)B); a_dev.rs_orig = (int32_t*)(base + (size_t)2 * B);
    a_dev.rs_new = (int32_t*)(base + (size_t)3 * B); a_dev.crop_off = (int32_t*)(base + (size_t)4 * B);
    a_dev.noise_idx = (int32_t*)(base + (size_t)5 * B); a_dev.noise_off = (int32_t*)(base + (size_t)6 * B);
    a_dev.snr_db = (float*)(base + (size_t)7 * B); a_dev.gain = (float*)(base + (size_t)8 * B);
  }
  if ((rc = ww_prepare_weights(c, st))) return rc;
  if (aug_host && (rc = ww_prepare_bank_energy(c, bank_dev, bank_rows, bank_len, st))) return rc;
  // the conv partials of every piece stay live until the head runs with the last one: size the pool for the whole
  // batch now (growing it between pieces would drop the earlier pieces' partials)
  if ((rc = ensure_pool(c, B))) return rc;
  // Two streams: the copy engine moves piece i+1 host->device while the SMs score piece i.  Pieces ramp up
  // geometrically (384, 512, 768, 1024, ... clips up to the work chunk) so that only a fraction of a millisecond of copy is exposed
  // before the first kernel starts; copying a clip is faster than scoring it, so later copies stay ahead.
  if (!c->copy_stream) WW_CHECK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  char* d_in = (char*)c->d_host_in;
  const char* h_in = (const char*)clips_host;
  std::vector<std::pair<int, int>> pieces;          // (first clip, clips)
  // Piece k+1 is copied while piece k is scored, so it may only be as much larger as copying a clip is faster than scoring
  // it (~1.3x with int16 PCM over PCIe 5); a doubling ramp stalls the SMs at every step of the ramp (measured: 2 ms).
  static const int cap_env = getenv("WW_H2D_CAP") ? std::max(64, atoi(getenv("WW_H2D_CAP"))) : kH2DCap;
  const int piece_cap = std::min(c->chunk, cap_env);
  static const double ramp = getenv("WW_H2D_RAMP") ? std::max(1.05, atof(getenv("WW_H2D_RAMP"))) : (double)kH2DRamp;
  static const int first = getenv("WW_H2D_FIRST") ? std::max(64, atoi(getenv("WW_H2D_FIRST"))) : kH2DFirst;
  for (int b0 = 0, sz = std::min(first, piece_cap); b0 < B; ) {
    const int nb = std::min(sz, B - b0);
    pieces.emplace_back(b0, nb);
    b0 += nb;
    sz = std::min((((int)(sz * ramp) + 127) / 128) * 128, piece_cap);
  }
  const int n_pieces = (int)pieces.size();
  while ((int)c->copy_events.size() < n_pieces) {
    cudaEvent_t e;
    WW_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->copy_events.push_back(e);
  }
  for (int i = 0; i < n_pieces; ++i) {
    const int b0 = pieces[i].first, nb = pieces[i].second;
    WW_CHECK(c, cudaMemcpyAsync(d_in + (size_t)b0 * N * esz, h_in + (size_t)b0 * N * esz, (size_t)nb * N * esz,
                                cudaMemcpyHostToDevice, c->copy_stream));
    WW_CHECK(c, cudaEventRecord(c->copy_events[i], c->copy_stream));
  }
  for (int i = 0; i < n_pieces; ++i) {
    const int b0 = pieces[i].first, nb = pieces[i].second;
    WW_CHECK(c, cudaStreamWaitEvent(st, c->copy_events[i], 0));
    ww_aug a = a_dev;
    if (aug_host) {
      a.flags += b0; a.shift += b0; a.rs_orig += b0; a.rs_new += b0; a.crop_off += b0;
      a.noise_idx += b0; a.noise_off += b0; a.snr_db += b0; a.gain += b0;
    }
    // the head (whole batch) runs with the last piece
    rc = score_impl(c, d_in + (size_t)b0 * N * esz, pcm16, N, bank_dev, bank_rows, bank_len, aug_host ? &a : nullptr,
                    normalize, d_logits, d_prob, d_dec, nb, st, b0, i == n_pieces - 1);
    if (rc) return rc;
  }
  if (logits_host) WW_CHECK(c, cudaMemcpyAsync(logits_host, d_logits, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
  if (prob1_host) WW_CHECK(c, cudaMemcpyAsync(prob1_host, d_prob, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  if (decision_host) WW_CHECK(c, cudaMemcpyAsync(decision_host, d_dec, (size_t)B, cudaMemcpyDeviceToHost, st));
  WW_CHECK(c, cudaStreamSynchronize(st));
  return WW_OK;
}

int ww_score_host(ww_ctx* c, const float* clips_host, const float* bank_dev, int bank_rows, int64_t bank_len,
                  const ww_aug* aug_host, int normalize, float* logits_host, float* prob1_host,
                  uint8_t* decision_host, int B) {
  return score_host_impl(c, clips_host, 0, bank_dev, bank_rows, bank_len, aug_host, normalize, logits_host, prob1_host,
                         decision_host, B);
}

int ww_score_host_pcm16(ww_ctx* c, const int16_t* clips_host, const float* bank_dev, int bank_rows, int64_t bank_len,
                        const ww_aug* aug_host, int normalize, float* logits_host, float* prob1_host,
                        uint8_t* decision_host, int B) {
  return score_host_impl(c, clips_host, 1, bank_dev, bank_rows, bank_len, aug_host, normalize, logits_host, prob1_host,
                         decision_host, B);
}

// ---------------------------------------------------------------------------------------------
// Pinned host buffers next to the GPU.  On a two-socket box the pages of a plain cudaHostAlloc land on the NUMA node of the
// allocating thread; with eight ranks feeding eight GPUs from one node the host->device copies of ww_score_host share one
// socket's memory controllers and the inter-socket link (round 1: 185 GB/s aggregate, 0.62 scaling efficiency end to end).
// ww_host_alloc places the pages on the GPU's own node (sysfs numa_node of its PCI function): mbind(MPOL_BIND) on an
// anonymous mapping, or, where the container's seccomp profile refuses mbind, first touch from a thread moved onto a CPU
// of that node; then cudaHostRegister.  `how` reports what happened: 0 = plain pinned memory (node unknown / nothing
// worked), 1 = mbind, 2 = first touch on a local CPU.
namespace {

int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
  for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// CPUs of a NUMA node, from sysfs cpulist ("0-31,64-95")
std::vector<int> node_cpus(int node) {
  std::vector<int> out;
  const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return out;
  int a, b;
  char sep;
  while (fscanf(f, "%d", &a) == 1) {
    b = a;
    if (fscanf(f, "%c", &sep) == 1 && sep == '-') {
      if (fscanf(f, "%d", &b) != 1) b = a;
      if (fscanf(f, "%c", &sep) != 1) sep = 0;
    }
    for (int c = a; c <= b; ++c) out.push_back(c);
    if (sep != ',') break;
  }
  fclose(f);
  return out;
}

struct HostBuf { void* p; size_t bytes; int device; };
std::vector<HostBuf> g_host_bufs;

}  // namespace

int ww_host_numa_node(ww_ctx* c) { return c ? gpu_numa_node(c->device) : -1; }

void* ww_host_alloc(ww_ctx* c, size_t bytes, int* how) {
  if (how) *how = 0;
  if (!c || bytes == 0) return nullptr;
  DeviceGuard dev_guard(c->device);
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  const size_t len = (bytes + page - 1) / page * page;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) { c->set_error("ww_host_alloc: mmap failed"); return nullptr; }
  int method = 0;
  const int node = gpu_numa_node(c->device);
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_mbind
    if (syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, mask, (unsigned long)(8 * sizeof(mask)), 0u) == 0) method = 1;
#endif
    if (method == 0) {
      const std::vector<int> cpus = node_cpus(node);
      cpu_set_t old, want;
      CPU_ZERO(&want);
      for (int cpu : cpus) if (cpu < CPU_SETSIZE) CPU_SET(cpu, &want);
      if (!cpus.empty() && sched_getaffinity(0, sizeof(old), &old) == 0 && sched_setaffinity(0, sizeof(want), &want) == 0) {
        memset(p, 0, len);                         // first touch on a CPU of the GPU's node
        sched_setaffinity(0, sizeof(old), &old);
        method = 2;
      }
    }
  }
  if (method != 2) memset(p, 0, len);              // fault the pages in (under the mbind policy when it was accepted)
  if (cudaHostRegister(p, len, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    munmap(p, len);
    c->set_error("ww_host_alloc: cudaHostRegister failed");
    return nullptr;
  }
  g_host_bufs.push_back(HostBuf{p, len, c->device});
  if (how) *how = method;
  return p;
}

void ww_host_free(ww_ctx*, void* p) {          // the context may already be gone (finalizers at interpreter exit): not touched
  if (!p) return;
  for (size_t i = 0; i < g_host_bufs.size(); ++i)
    if (g_host_bufs[i].p == p) {
      DeviceGuard dev_guard(g_host_bufs[i].device);
      cudaHostUnregister(p);
      munmap(p, g_host_bufs[i].bytes);
      g_host_bufs.erase(g_host_bufs.begin() + i);
      return;
    }
}

}  // extern "C"

/*
 * wakeword_b200.h -- C ABI of libwakeword_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for ONE hot path of sarpel/wakeword-jupyterlab:
 *   batched 16 kHz clips -> (augment) -> log-mel -> CNN+LSTM score.
 *
 * The reference has no FFI of its own (it is pure Python); each entry point
 * below replaces the arithmetic behind one reference method (file:line under
 * /root/reference), and is what the reference-side ctypes binding in
 * INTEGRATION.md calls.  Conventions:
 *   - all `const float*` / `float*` data pointers are DEVICE pointers owned by
 *     the caller unless the name ends in `_host`;
 *   - every call enqueues on the caller's cudaStream_t (passed as void*) and does
 *     not synchronise, except ww_create/ww_destroy/ww_set_weights/ww_*_host;
 *   - return value: 0 = ok, negative = error (text via ww_last_error);
 *   - one context per (device, host thread); a context is not thread safe;
 *   - there is NO CPU fallback: every entry fails if the device is not sm_100.
 */
#ifndef WAKEWORD_B200_H_
#define WAKEWORD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WW_ABI_VERSION 3

/* error codes */
#define WW_OK 0
#define WW_ERR_INVALID (-1)  /* bad argument / unsupported configuration      */
#define WW_ERR_CUDA (-2)     /* CUDA runtime error (text in ww_last_error)    */
#define WW_ERR_WEIGHTS (-3)  /* forward requested before all weights were set */
#define WW_ERR_ARCH (-4)     /* device is not compute capability 10.x         */

/* conv2/conv3 arithmetic (ww_config.conv_mode).  The tensor-core modes keep activations as one fp16 value
 * (their rounding errors are independent per pixel and average out in the global mean) and differ in how the
 * weights -- whose rounding errors do NOT average out -- are represented. */
#define WW_CONV_SPLIT2 0 /* tcgen05, weights = fp16 hi + fp16 lo (2 MMA passes), fp32 accumulate: parity mode (default) */
#define WW_CONV_FP32 1   /* fp32 CUDA-core direct convolution (exact-arithmetic cross-check path)                       */
#define WW_CONV_FP16 2   /* tcgen05, single fp16 pass: "fast" mode (logits ~5e-5 relative on the golden weights)        */

/* augmentation flag bits (ww_aug.flags), applied in this order */
#define WW_AUG_NORM_IN (1u << 0)  /* peak normalise first: normalize_audio, wakeword_training_script.py:73-76 */
#define WW_AUG_SHIFT (1u << 1)    /* circular shift, np.roll: wakeword_training_script.py:106-108              */
#define WW_AUG_SPEED (1u << 2)    /* polyphase resample rs_orig->rs_new then crop/pad: :114-117 + :78-83       */
#define WW_AUG_NOISE (1u << 3)    /* snr_mixer: stock/ms_snsd/MS-SNSD/audiolib.py:55-71                        */
#define WW_AUG_GAIN (1u << 4)     /* multiply by gain                                                           */
#define WW_AUG_NORM_OUT (1u << 5) /* peak normalise last                                                        */

typedef struct ww_ctx ww_ctx;

/* Mirrors AudioConfig (wakeword_training_script.py:29-37) + ModelConfig (:39-43). */
typedef struct ww_config {
  int32_t sample_rate; /* SAMPLE_RATE 16000                      */
  int32_t n_samples;   /* int(SAMPLE_RATE * DURATION) = 16000    */
  int32_t n_fft;       /* N_FFT 2048 (power of two, 256..2048)   */
  int32_t win_length;  /* WIN_LENGTH 2048 (<= n_fft)             */
  int32_t hop_length;  /* HOP_LENGTH 512                         */
  int32_t n_mels;      /* N_MELS 80                              */
  float fmin;          /* FMIN 0                                 */
  float fmax;          /* FMAX 8000                              */
  int32_t hidden_size; /* HIDDEN_SIZE 256 (multiple of 32)       */
  int32_t num_layers;  /* NUM_LAYERS 2                           */
  int32_t num_classes; /* NUM_CLASSES 2                          */
  float threshold;     /* predict_wakeword threshold 0.8         */
  int32_t conv_mode;   /* WW_CONV_*                              */
  int32_t chunk_clips; /* clips per internal work chunk (0 = default) */
} ww_config;

/* Per-clip augmentation parameters, structure of DEVICE arrays of length B.
 * All values are drawn by the host (seed-supplied); the kernels only consume them. */
typedef struct ww_aug {
  const uint32_t* flags;    /* WW_AUG_* bits                                          */
  const int32_t* shift;     /* out[i] = in[(i - shift) mod N]                         */
  const int32_t* rs_orig;   /* speed change: resample orig -> new (speed = orig/new)  */
  const int32_t* rs_new;
  const int32_t* crop_off;  /* start offset when the resampled clip is longer than N  */
  const int32_t* noise_idx; /* row of the noise bank                                  */
  const int32_t* noise_off; /* first sample of the N-sample noise segment             */
  const float* snr_db;      /* snr argument of snr_mixer                              */
  const float* gain;        /* linear gain                                            */
} ww_aug;

/* ---- lifetime ---------------------------------------------------------------------- */
int ww_abi_version(void);
int ww_create(ww_ctx** out, int device, const ww_config* cfg);
void ww_destroy(ww_ctx* ctx);
const char* ww_last_error(const ww_ctx* ctx); /* ctx may be NULL: last ww_create error */
int ww_n_frames(const ww_ctx* ctx);           /* W = 1 + n_samples / hop_length        */

/* ---- weights: names are the reference state_dict keys:
 *      conv{1,2,3}.{weight,bias}, lstm.{weight,bias}_{ih,hh}_l{k}, fc.{weight,bias}.
 *      `src` may be a device or a host pointer (copied synchronously).
 *      Replaces: WakewordModel.__init__ parameters, wakeword_training_script.py:150-165. */
int ww_set_weights(ww_ctx* ctx, const char* name, const float* src, const int64_t* shape, int ndim);

/* ---- polyphase tables: must be called (host side) for every (orig,new) pair that appears
 *      in ww_aug.rs_orig/rs_new before ww_augment / ww_score use it. */
int ww_prepare_resample(ww_ctx* ctx, int rs_orig, int rs_new);

/* ---- K1: batched augmentation. clips[B][n_samples], noise_bank[bank_rows][bank_len]
 *      -> out[B][n_samples].  Replaces AudioProcessor.normalize_audio / pad_or_truncate /
 *      augment_audio (wakeword_training_script.py:73-83, :103-123) with the north-star stage set. */
int ww_augment(ww_ctx* ctx, const float* clips, const float* noise_bank, int bank_rows, int64_t bank_len,
               const ww_aug* p, float* out, int B, void* stream);

/* ---- peak normalise a vector of any length: out[i] = in[i] / max|in| (IEEE divide; all-zero input is
 *      passed through).  Replaces AudioProcessor.normalize_audio (wakeword_training_script.py:73-76). */
int ww_normalize(ww_ctx* ctx, const float* in, float* out, int64_t n, void* stream);

/* ---- K2: log-mel. clips: B rows of n_samples floats, row r starts at clips + r*clip_stride
 *      (clip_stride = n_samples for a packed batch; a smaller stride gives overlapping windows).
 *      normalize != 0 applies normalize_audio per row first.  out[B][1][n_mels][W] fp32 dB.
 *      Replaces AudioProcessor.audio_to_mel (wakeword_training_script.py:85-101). */
int ww_logmel(ww_ctx* ctx, const float* clips, int64_t clip_stride, float* out, int B, int normalize,
              void* stream);

/* ---- K3+K4: model forward. logmel[B][1][n_mels][W] -> logits[B][num_classes].
 *      Replaces WakewordModel.forward in eval mode (wakeword_training_script.py:167-184). */
int ww_forward(ww_ctx* ctx, const float* logmel, float* logits, int B, void* stream);

/* ---- fused scoring: (augment) -> log-mel -> forward -> softmax -> threshold.
 *      aug may be NULL (then normalize selects normalize_audio only).  Any of logits / prob1 /
 *      decision may be NULL.  prob1 = softmax(logits)[:,1]; decision = prob1 >= threshold.
 *      Replaces predict_wakeword (wakeword_training.ipynb:871-893) over a batch. */
int ww_score(ww_ctx* ctx, const float* clips, const float* noise_bank, int bank_rows, int64_t bank_len,
             const ww_aug* aug, int normalize, float* logits, float* prob1, uint8_t* decision, int B,
             void* stream);

/* ---- streaming: windows k = 0..n_win-1 of n_samples samples starting at k*hop_samples of
 *      audio[T]; each window is scored exactly like predict_wakeword (per-window peak
 *      normalise, per-window dB reference).  n_win = 1 + (T - n_samples) / hop_samples. */
int ww_score_stream(ww_ctx* ctx, const float* audio, int64_t T, int hop_samples, float* prob1,
                    uint8_t* decision, int64_t n_win, void* stream);

/* ---- host-buffer entry (what a Python/ctypes caller with numpy arrays uses): copies
 *      clips H2D, scores, copies results D2H, synchronises.  Pointers are HOST pointers
 *      (pinned memory gives full PCIe rate).  aug_host arrays are host arrays. */
int ww_score_host(ww_ctx* ctx, const float* clips_host, const float* noise_bank_dev, int bank_rows,
                  int64_t bank_len, const ww_aug* aug_host, int normalize, float* logits_host,
                  float* prob1_host, uint8_t* decision_host, int B);

/* ---- the reference's own augmentation stages: phase-vocoder time stretch / pitch shift and
 *      Gaussian noise.  Replace librosa.effects.time_stretch + pad_or_truncate (wakeword_training_script.py:114-117),
 *      librosa.effects.pitch_shift (:110-112) and `audio + np.random.normal(0, NOISE_FACTOR)` (:119-121), with librosa's
 *      defaults at those call sites (n_fft 2048, hop 512: the context must have n_fft = win_length = 2048).
 *      Per clip, host-drawn, DEVICE arrays: rate (float64: the time-step grid is computed in double like numpy's);
 *      rs_orig / rs_new = 0 or equal: time stretch, result cropped at crop_off / zero-padded to n_samples;
 *      otherwise pitch shift: the stretched signal is resampled rs_orig -> rs_new (ww_prepare_resample first; the rational
 *      stand-in of sr / rate -> sr, rate = 2^(-n_steps / 12)) and fitted to n_samples.  clips[B][n_samples] -> out[B][n_samples]. */
typedef struct ww_pvoc {
  const double* rate;
  const int32_t* rs_orig;
  const int32_t* rs_new;
  const int32_t* crop_off;
} ww_pvoc;
int ww_time_stretch(ww_ctx* ctx, const float* clips, const ww_pvoc* p, float* out, int B, void* stream);
/* x[i] += sigma * N(0, 1), i < n; counter-based Philox4x32-10 stream keyed by `seed` (reproducible, order independent) */
int ww_add_gaussian_noise(ww_ctx* ctx, float* x, int64_t n, float sigma, uint64_t seed, void* stream);

/* ---- pinned host buffers for the *_host entries, placed on the NUMA node of the context's GPU (sysfs numa_node of its
 *      PCI function) so that eight ranks on a two-socket box do not all copy out of one socket's memory.
 *      *how: 0 = plain pinned memory, 1 = mbind(MPOL_BIND), 2 = first touch from a CPU of that node.
 *      The reference's counterpart is the DataLoader's pin_memory=True (wakeword_training_script.py:461-463). */
void* ww_host_alloc(ww_ctx* ctx, size_t bytes, int* how);
void ww_host_free(ww_ctx* ctx, void* ptr);
int ww_host_numa_node(ww_ctx* ctx); /* -1 = unknown */

/* ---- int16 PCM inputs.  The reference's clips are 16-bit WAV files that
 *      librosa.load turns into float32 as s / 32768 (AudioProcessor.load_audio, wakeword_training_script.py:65-71;
 *      the synthetic recipe writes them with sf.write, :359-388).  These entries take the int16 samples as they are
 *      on disk and apply exactly that scaling on the device: results are bit-identical to the fp32 entries fed
 *      with s / 32768, at half the host->device and HBM input bytes.  n_samples must be even. */
int ww_augment_pcm16(ww_ctx* ctx, const int16_t* clips, const float* noise_bank, int bank_rows, int64_t bank_len,
                     const ww_aug* p, float* out, int B, void* stream);
int ww_logmel_pcm16(ww_ctx* ctx, const int16_t* clips, int64_t clip_stride, float* out, int B, int normalize,
                    void* stream);
int ww_score_pcm16(ww_ctx* ctx, const int16_t* clips, const float* noise_bank, int bank_rows, int64_t bank_len,
                   const ww_aug* aug, int normalize, float* logits, float* prob1, uint8_t* decision, int B,
                   void* stream);
int ww_score_stream_pcm16(ww_ctx* ctx, const int16_t* audio, int64_t T, int hop_samples, float* prob1,
                          uint8_t* decision, int64_t n_win, void* stream);
int ww_score_host_pcm16(ww_ctx* ctx, const int16_t* clips_host, const float* noise_bank_dev, int bank_rows,
                        int64_t bank_len, const ww_aug* aug_host, int normalize, float* logits_host,
                        float* prob1_host, uint8_t* decision_host, int B);

/* ---- training step (BASELINE config 5).  Replaces the body of the batch loop of
 *      WakewordTrainer.train_epoch (wakeword_training_script.py:247-257) with the optimiser of
 *      WakewordTrainer.__init__ (:225-226): CrossEntropyLoss (mean), backward, Adam with coupled weight decay.
 *      The context owns the fp32 master weights (ww_set_weights / ww_get_weights), one flat fp32 gradient buffer in
 *      state_dict order (conv1.weight, conv1.bias, ..., lstm.weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, ...,
 *      fc.weight, fc.bias; every entry padded to a multiple of 4 floats) and the Adam moments.
 *      ww_train_backward: logmel[B][1][n_mels][W] (device), labels int64[B] (device); drop_lstm [layers-1][B][hidden]
 *      and drop_out [B][hidden] are optional multiplicative dropout masks (values 0 or 1/(1-p), host-seeded like the
 *      augmentation parameters; NULL = no dropout); writes the mean loss (device or host float*) and optionally the
 *      train-mode logits, and leaves the gradients in the flat buffer.
 *      Data parallelism = one sum all-reduce of ww_train_grad_buffer()[0 .. ww_train_n_params()) between
 *      ww_train_backward and ww_train_apply(grad_scale = 1 / world_size).
 *      ww_train_step = backward + (ncclAllReduce on `nccl_comm` when it is not NULL) + Adam(lr, 0.9, 0.999, 1e-8,
 *      weight_decay 1e-5), the reference's optimiser settings. */
int ww_train_backward(ww_ctx* ctx, const float* logmel, const int64_t* labels, int B, const float* drop_lstm,
                      const float* drop_out, float* loss, float* logits_out, void* stream);
int ww_train_apply(ww_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   void* stream);
int ww_train_step(ww_ctx* ctx, const float* logmel, const int64_t* labels, int B, float* loss, float lr,
                  void* nccl_comm, int world_size, void* stream);
int ww_train_reset(ww_ctx* ctx);                       /* zero the Adam moments and the step counter */
int64_t ww_train_n_params(ww_ctx* ctx);                /* floats in the flat gradient buffer (with padding) */
float* ww_train_grad_buffer(ww_ctx* ctx);              /* device pointer of the flat gradient buffer */
int ww_train_param_range(ww_ctx* ctx, const char* name, int64_t* offset, int64_t* count);
int ww_get_weights(ww_ctx* ctx, const char* name, float* dst); /* dst: device or host pointer */
/* Adam state, for optimizer.state_dict() of best_wakeword_model.pth (wakeword_training_script.py:326-334) and for giving
 * every WakewordTrainer its own optimiser on a shared context: exp_avg / exp_avg_sq of one parameter (device or host
 * pointers, NULL = skip) and the step counter. */
int ww_train_get_moments(ww_ctx* ctx, const char* name, float* exp_avg, float* exp_avg_sq);
int ww_train_set_moments(ww_ctx* ctx, const char* name, const float* exp_avg, const float* exp_avg_sq);
int64_t ww_train_get_step(ww_ctx* ctx);
int ww_train_set_step(ww_ctx* ctx, int64_t step);

/* ---- per-stage device timing for benchmarks: CUDA events recorded on the launching stream around
 *      each stage's kernels while enabled.  ww_profile_read synchronises, returns the summed
 *      milliseconds and the number of timed launches of `stage`, and (stage < 0) resets the log. */
#define WW_STAGE_AUGMENT 0
#define WW_STAGE_LOGMEL 1
#define WW_STAGE_CONV12 2 /* conv1 + conv2 (+ReLU) */
#define WW_STAGE_CONV3 3  /* conv3 + ReLU + global mean: the dominant kernel */
#define WW_STAGE_HEAD 4
int ww_profile(ww_ctx* ctx, int enable);
int ww_profile_read(ww_ctx* ctx, int stage, double* total_ms, int64_t* n_launches);

/* ---- introspection for benchmarks / tests */
int64_t ww_kernel_launches(const ww_ctx* ctx); /* kernels launched by this context so far */
int ww_conv_mode(const ww_ctx* ctx);
/* decision threshold of the following ww_score* calls (predict_wakeword's `threshold` argument, ipynb:871) */
int ww_set_threshold(ww_ctx* ctx, float threshold);

#ifdef __cplusplus
}
#endif
#endif /* WAKEWORD_B200_H_ */

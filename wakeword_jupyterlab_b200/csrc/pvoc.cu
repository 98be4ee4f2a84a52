// The reference's own augmentations: phase-vocoder time stretch, pitch shift, Gaussian noise.
//
// Replaces the librosa calls of AudioProcessor.augment_audio (/root/reference/wakeword_training_script.py:110-121):
//   librosa.effects.time_stretch(y, rate)   = istft(phase_vocoder(stft(y), rate), length = round(len / rate))  (:116)
//   librosa.effects.pitch_shift(y, n_steps) = fix_length(resample(time_stretch(y, 2^(-n/12)), sr / rate -> sr)) (:112)
//   y + np.random.normal(0, NOISE_FACTOR)                                                                        (:120-121)
// with librosa's defaults at those call sites (n_fft 2048, hop 512, periodic Hann, center = True with zero padding,
// window sum-of-squares normalisation in the inverse).  oracle/pvoc.py restates the same algorithm and is cross-checked
// against torch.stft / torch.istft / torchaudio.functional.phase_vocoder.
//
// Two kernels per batch:
//   stft_kernel      one CTA per (clip, frame): window, 2048-point shared-memory FFT, 1025 bins -> global scratch
//   pv_istft_kernel  one CTA per clip, thread = bin: walks the output frames with the phase accumulator in a register
//                    (magnitude interpolation + wrapped phase advance), inverse FFT of every output frame in shared memory,
//                    windowed overlap-add into a shared-memory signal buffer, sum-of-squares normalisation, then either
//                    the reference's pad_or_truncate (time stretch) or the polyphase resample back to the input rate
//                    (pitch shift; table of ww_prepare_resample) straight out of that buffer.
// Host-drawn parameters only (rate, crop offset, resample ratio), like every other augmentation stage; the Gaussian
// noise stage draws on the device from a counter-based Philox4x32-10 stream keyed by a host seed.
#include "ctx.cuh"

#include <algorithm>

namespace {

constexpr int PV_NFFT = 2048, PV_HOP = 512, PV_BINS = PV_NFFT / 2 + 1, PV_THREADS = 1024;
constexpr int PV_MAX_OUT = 40000;            // longest stretched signal held in shared memory (rate >= 0.4 at 16,000 samples)

__device__ __forceinline__ int bitrev11(int v) { return (int)(__brev((unsigned)v) >> 21); }

// in-place radix-2 FFT of 2048 complex points in shared memory (input already in bit-reversed order), 1024 threads,
// one butterfly per thread and stage; SIGN = -1 forward, +1 inverse (unscaled).  tw[t] = exp(-2 pi i t / 2048).
template <int SIGN>
__device__ __forceinline__ void fft2048(float2* z, const float2* __restrict__ tw, int tid) {
#pragma unroll 1
  for (int s = 0; s < 11; ++s) {
    const int half = 1 << s, j = tid & (half - 1), i0 = ((tid >> s) << (s + 1)) + j, i1 = i0 + half;
    float2 w = __ldg(tw + (j << (10 - s)));
    if (SIGN > 0) w.y = -w.y;
    __syncthreads();
    const float2 a = z[i0], b = z[i1];
    const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
    z[i0] = make_float2(a.x + t.x, a.y + t.y);
    z[i1] = make_float2(a.x - t.x, a.y - t.y);
  }
  __syncthreads();
}

// spec[(b T + t) 1025 + k] = sum_n hann[n] x_b[512 t + n - 1024] exp(-2 pi i k n / 2048)
__global__ void __launch_bounds__(PV_THREADS) stft_kernel(const float* __restrict__ clips, int N, int T,
                                                           const float* __restrict__ win, const float2* __restrict__ tw,
                                                           float2* __restrict__ spec) {
  __shared__ float2 z[PV_NFFT];
  const int tid = threadIdx.x, b = blockIdx.x / T, t = blockIdx.x - b * T;
  const float* x = clips + (size_t)b * N;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int n = tid + r * PV_THREADS, i = PV_HOP * t + n - PV_NFFT / 2;
    const float v = (i >= 0 && i < N) ? __ldg(x + i) * __ldg(win + n) : 0.0f;
    z[bitrev11(n)] = make_float2(v, 0.0f);
  }
  fft2048<-1>(z, tw, tid);
  float2* o = spec + ((size_t)b * T + t) * PV_BINS;
  o[tid] = z[tid];
  if (tid == 0) o[PV_NFFT / 2] = z[PV_NFFT / 2];
}

struct PvParams {
  const float2* spec;            // [B][T][1025]
  const double* rate;            // [B]
  const int* rs_orig;            // [B] pitch mode: resample rs_orig -> rs_new after the stretch (equal / 0: time-stretch mode)
  const int* rs_new;
  const int* crop_off;           // [B] time-stretch mode: start offset when the stretched signal is longer than N
  float* out;                    // [B][N]
  int N, T;
  const float* win;
  const float2* tw;
  const RsDesc* rs_desc;
  int n_rs;
  const float* rs_kern;
};

__global__ void __launch_bounds__(PV_THREADS) pv_istft_kernel(const PvParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);                 // [2048]
  float* y = reinterpret_cast<float*>(z + PV_NFFT);                // [PV_MAX_OUT + n_fft + hop]: overlap-add buffer
  const int tid = threadIdx.x, b = blockIdx.x, N = p.N, T = p.T;
  const double rate = p.rate[b];
  const int L = (int)rint((double)N / rate);                       // librosa: int(round(len / rate))
  const int T_out = (int)ceil((double)T / rate);                   // len(np.arange(0, T, rate))
  const int padded = L + PV_NFFT;
  const int n_frames = min(T_out, (padded + PV_HOP - 1) / PV_HOP); // librosa.istft with `length`
  const int total = PV_NFFT + PV_HOP * (n_frames - 1);
  float* dst = p.out + (size_t)b * N;
  if (!(rate > 0.0) || L < 1 || total > PV_MAX_OUT + PV_NFFT + PV_HOP) {       // loud: NaN clip
    for (int i = tid; i < N; i += PV_THREADS) dst[i] = __int_as_float(0x7fc00000);
    return;
  }
  for (int i = tid; i < total; i += PV_THREADS) y[i] = 0.0f;
  const float2* S = p.spec + (size_t)b * T * PV_BINS;
  // thread = bin k (thread 0 also carries bin 1024); expected phase advance pi hop k / 1024
  const int nb = tid == 0 ? 2 : 1;
  float phase[2];
  double adv[2];
  for (int r = 0; r < nb; ++r) {
    const int k = r ? PV_NFFT / 2 : tid;
    const float2 c = S[k];
    phase[r] = atan2f(c.y, c.x);
    adv[r] = 3.14159265358979323846 * PV_HOP * (double)k / (double)(PV_BINS - 1);
  }
  for (int t = 0; t < n_frames; ++t) {
    const double step = (double)t * rate;
    const int i0 = (int)step;
    const float alpha = (float)(step - (double)i0);
    __syncthreads();                                               // z of the previous frame has been consumed
    for (int r = 0; r < nb; ++r) {
      const int k = r ? PV_NFFT / 2 : tid;
      const float2 c0 = i0 < T ? S[(size_t)i0 * PV_BINS + k] : make_float2(0.0f, 0.0f);
      const float2 c1 = i0 + 1 < T ? S[(size_t)(i0 + 1) * PV_BINS + k] : make_float2(0.0f, 0.0f);
      const float mag = (1.0f - alpha) * hypotf(c0.x, c0.y) + alpha * hypotf(c1.x, c1.y);
      float sn, cs;
      sincosf(phase[r], &sn, &cs);
      const float2 v = make_float2(mag * cs, mag * sn);
      // Hermitian extension for the inverse transform (bins 0 and 1024 contribute their real part only, like irfft)
      if (k == 0 || k == PV_NFFT / 2) z[bitrev11(k)] = make_float2(v.x, 0.0f);
      else { z[bitrev11(k)] = v; z[bitrev11(PV_NFFT - k)] = make_float2(v.x, -v.y); }
      // phase advance: wrapped difference of the two frames + expected advance, accumulated in double like numpy
      double dph = (double)atan2f(c1.y, c1.x) - (double)atan2f(c0.y, c0.x) - adv[r];
      dph -= 6.283185307179586476925 * rint(dph / 6.283185307179586476925);
      phase[r] = (float)remainder((double)phase[r] + adv[r] + dph, 6.283185307179586476925);
    }
    fft2048<+1>(z, p.tw, tid);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int n = tid + r * PV_THREADS;
      y[PV_HOP * t + n] += z[n].x * (1.0f / PV_NFFT) * __ldg(p.win + n);    // frames are sequential: no race
    }
  }
  __syncthreads();
  // window sum-of-squares normalisation (at most four frames overlap a sample), then drop the centre padding
  for (int i = tid; i < total; i += PV_THREADS) {
    float wss = 0.0f;
    const int t_hi = min(i / PV_HOP, n_frames - 1);
    for (int t = t_hi; t >= 0 && i - PV_HOP * t < PV_NFFT; --t) {
      const float w = __ldg(p.win + i - PV_HOP * t);
      wss = fmaf(w, w, wss);
    }
    if (wss > 1.17549435e-38f) y[i] /= wss;
  }
  __syncthreads();
  const float* ys = y + PV_NFFT / 2;                               // ys[i], 0 <= i < L (zero beyond the overlap-add span)
  const int avail = min(L, total - PV_NFFT / 2);
  const int ro = p.rs_orig ? p.rs_orig[b] : 0, rn = p.rs_new ? p.rs_new[b] : 0;
  if (ro <= 0 || ro == rn) {
    // time stretch + pad_or_truncate(N) with the host-drawn crop offset (:117, :78-83)
    const int crop = L > N ? p.crop_off[b] : 0;
    for (int i = tid; i < N; i += PV_THREADS) {
      const int s = i + crop;
      dst[i] = (s < avail) ? ys[s] : 0.0f;
    }
    return;
  }
  // pitch shift: polyphase resample ro -> rn of the stretched signal, then fix_length(N)
  int ri = -1;
  for (int i = 0; i < p.n_rs; ++i)
    if (p.rs_desc[i].orig == ro && p.rs_desc[i].neu == rn) ri = i;
  if (ri < 0) {
    for (int i = tid; i < N; i += PV_THREADS) dst[i] = __int_as_float(0x7fc00000);      // ratio not prepared: loud
    return;
  }
  const RsDesc d = p.rs_desc[ri];
  const int pitch = d.nz + ((d.nz & 4) ? 0 : 4);
  const float* kern = p.rs_kern + d.offset;
  const int* lo_t = reinterpret_cast<const int*>(kern + d.n * pitch);
  const int* cnt_t = lo_t + d.n;
  const long long out_len = ((long long)d.n * L + d.o - 1) / d.o;
  for (int j = tid; j < N; j += PV_THREADS) {
    float acc = 0.0f;
    if (j < out_len) {
      const int q = j / d.n, ph = j - q * d.n;
      const int x0 = q * d.o - d.width + lo_t[ph];
      const float* kr = kern + ph * pitch;
      const int k0 = x0 < 0 ? -x0 : 0, k1 = min(cnt_t[ph], avail - x0);
      for (int k = k0; k < k1; ++k) acc = fmaf(kr[k], ys[x0 + k], acc);
    }
    dst[j] = acc;
  }
}

// ---- Gaussian noise: Philox4x32-10 (counter = sample quad index, key = seed) + Box-Muller
__device__ __forceinline__ void philox_round(uint4& c, uint2& k) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
  c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
}
__global__ void gaussian_noise_kernel(float* __restrict__ x, int64_t n, float sigma, unsigned long long seed) {
  const int64_t quads = (n + 3) / 4;
  for (int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; qi < quads; qi += (int64_t)gridDim.x * blockDim.x) {
    uint4 c = make_uint4((unsigned)qi, (unsigned)(qi >> 32), 0u, 0u);
    uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    const unsigned u[4] = {c.x, c.y, c.z, c.w};
    float g[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)u[2 * h] + 1.0f) * 2.3283064365386963e-10f;        // (0, 1]
      const float u2 = (float)u[2 * h + 1] * 2.3283064365386963e-10f;
      const float rr = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      g[2 * h] = rr * cs; g[2 * h + 1] = rr * sn;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t i = 4 * qi + e;
      if (i < n) x[i] = fmaf(sigma, g[e], x[i]);
    }
  }
}

}  // namespace

extern "C" {

int ww_time_stretch(ww_ctx* c, const float* clips, const ww_pvoc* pv, float* out, int B, void* stream) {
  if (!c || !clips || !pv || !pv->rate || !pv->crop_off || !out || B < 0) return WW_ERR_INVALID;
  if (B == 0) return WW_OK;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = c->cfg.n_samples, T = 1 + N / PV_HOP;
  if (c->cfg.n_fft != PV_NFFT) {            // the window / twiddle tables of the context are those of n_fft
    c->set_error("ww_time_stretch: needs a context with n_fft = 2048 (librosa's default at the reference's call sites)");
    return WW_ERR_INVALID;
  }
  if (c->cfg.win_length != PV_NFFT) { c->set_error("ww_time_stretch: needs win_length = n_fft = 2048"); return WW_ERR_INVALID; }
  const size_t smem = (size_t)PV_NFFT * sizeof(float2) + (size_t)(PV_MAX_OUT + PV_NFFT + PV_HOP) * sizeof(float);
  WW_CHECK(c, cudaFuncSetAttribute(pv_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int chunk = 2048;                   // clips per pass: 2048 x 32 frames x 1025 bins x 8 B = 537 MB of scratch
  const size_t need = (size_t)std::min(B, chunk) * T * PV_BINS * sizeof(float2);
  if (need > c->pv_spec_bytes) {
    if (c->d_pv_spec) { WW_CHECK(c, cudaDeviceSynchronize()); cudaFree(c->d_pv_spec); c->d_pv_spec = nullptr; c->pv_spec_bytes = 0; }
    WW_CHECK(c, cudaMalloc((void**)&c->d_pv_spec, need));
    c->pv_spec_bytes = need;
  }
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    stft_kernel<<<nb * T, PV_THREADS, 0, st>>>(clips + (size_t)b0 * N, N, T, c->d_window, c->d_twiddle, c->d_pv_spec);
    WW_LAUNCH_CHECK(c);
    PvParams p;
    p.spec = c->d_pv_spec; p.rate = pv->rate + b0; p.rs_orig = pv->rs_orig ? pv->rs_orig + b0 : nullptr;
    p.rs_new = pv->rs_new ? pv->rs_new + b0 : nullptr; p.crop_off = pv->crop_off + b0; p.out = out + (size_t)b0 * N;
    p.N = N; p.T = T; p.win = c->d_window; p.tw = c->d_twiddle;
    p.rs_desc = c->d_rs_desc; p.n_rs = (int)c->rs_tables.size(); p.rs_kern = c->d_rs_kern;
    pv_istft_kernel<<<nb, PV_THREADS, smem, st>>>(p);
    WW_LAUNCH_CHECK(c);
  }
  return WW_OK;
}

int ww_add_gaussian_noise(ww_ctx* c, float* x, int64_t n, float sigma, uint64_t seed, void* stream) {
  if (!c || !x || n < 0) return WW_ERR_INVALID;
  if (n == 0) return WW_OK;
  DeviceGuard dev_guard(c->device);
  const int64_t quads = (n + 3) / 4;
  const int grid = (int)std::min<int64_t>((quads + 255) / 256, (int64_t)c->sm_count * 16);
  gaussian_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, sigma, (unsigned long long)seed);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

}  // extern "C"

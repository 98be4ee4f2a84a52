// K2: fused framing + Hann window + STFT power + mel filterbank + per-clip-max dB  (sm_100a)
//
// Replaces AudioProcessor.audio_to_mel (/root/reference/wakeword_training_script.py:85-101):
// librosa.feature.melspectrogram(center=True, pad_mode="constant", hann, power=2, slaney mel)
// followed by librosa.power_to_db(ref=np.max, amin=1e-10, top_db=80).
//
// One CTA per clip.  Two real frames are packed into one complex n_fft-point FFT
// (frame 2p -> real part, frame 2p+1 -> imaginary part) computed by a mixed-radix (8/4/2)
// Stockham autosort FFT in shared memory; the two spectra are separated by Hermitian symmetry,
// squared, and projected onto the sparse (banded) slaney filterbank.  The whole n_mels x W mel
// power of the clip stays in shared memory, so the per-clip max that the dB conversion needs
// (ref=np.max) never costs a second pass over HBM.
//
// HBM traffic per clip: n_samples*4 B in (each sample is re-read n_fft/hop times, from L1/L2)
// + n_mels*W*4 B out  = 74,240 B at the code preset.
#include "ctx.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kAmin = 1e-10f;
constexpr float kTopDb = 80.0f;

__device__ __forceinline__ int padi(int i) { return i + (i >> 5); }   // de-conflict strided Stockham stores

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

__device__ __forceinline__ void dft2(float2* v) {
  float2 t = v[0];
  v[0] = cadd(t, v[1]);
  v[1] = csub(t, v[1]);
}
__device__ __forceinline__ void dft4(float2* v) {
  float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
  float2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(t0, t2);
  v[2] = csub(t0, t2);
  v[1] = cadd(t1, t3);
  v[3] = csub(t1, t3);
}
__device__ __forceinline__ void dft8(float2* v) {
  float2 e[4] = {v[0], v[2], v[4], v[6]};
  float2 o[4] = {v[1], v[3], v[5], v[7]};
  dft4(e);
  dft4(o);
  const float h = 0.70710678118654752440f;
  float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * exp(-i pi/4)
  float2 o2 = mul_mi(o[2]);                                                   // * (-i)
  float2 o3 = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * exp(-3 i pi/4)
  v[0] = cadd(e[0], o[0]);  v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], o1);    v[5] = csub(e[1], o1);
  v[2] = cadd(e[2], o2);    v[6] = csub(e[2], o2);
  v[3] = cadd(e[3], o3);    v[7] = csub(e[3], o3);
}
template <int R> __device__ __forceinline__ void dftR(float2* v);
template <> __device__ __forceinline__ void dftR<2>(float2* v) { dft2(v); }
template <> __device__ __forceinline__ void dftR<4>(float2* v) { dft4(v); }
template <> __device__ __forceinline__ void dftR<8>(float2* v) { dft8(v); }

// One Stockham stage: N points, radix R, Ns = product of the radices already applied.
// v[r] = src[j + r*N/R] * T[r*k*(N/(Ns*R))],  k = j % Ns;  dst[(j/Ns)*Ns*R + k + r*Ns] = DFT_R(v)[r].
template <int R>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ src, float2* __restrict__ dst,
                                               const float2* __restrict__ tw, int N, int Ns, int tid) {
  const int nb = N / R;
  const int tstep = N / (Ns * R);
  for (int j = tid; j < nb; j += kThreads) {
    const int k = j & (Ns - 1);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[padi(j + r * nb)];
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[r * k * tstep]);
    }
    dftR<R>(v);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[padi(base + r * Ns)] = v[r];
  }
}

struct LogmelParams {
  const float* clips;
  int64_t clip_stride;
  float* out;
  int B, normalize;
  int n_samples, n_fft, log2n, hop, W, n_mels;
  const float* window;
  const float2* twiddle;
  const int* mel_start;
  const int* mel_len;
  const int* mel_off;
  const float* mel_w;
};

__device__ __forceinline__ float block_max(float v, float* red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) r = fmaxf(r, red[i]);
  return r;
}

__global__ void __launch_bounds__(kThreads) logmel_kernel(LogmelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft;
  const int npad = N + (N >> 5) + 8;
  float2* buf0 = reinterpret_cast<float2*>(smem_raw);
  float2* buf1 = buf0 + npad;
  float2* tw = buf1 + npad;                                  // [N]
  float* mel_s = reinterpret_cast<float*>(tw + N);           // [n_mels * W]
  __shared__ float red[kThreads / 32];

  const int tid = threadIdx.x;
  const int nbins = N / 2 + 1;
  for (int i = tid; i < N; i += kThreads) tw[i] = p.twiddle[i];

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* __restrict__ x = p.clips + (int64_t)b * p.clip_stride;
    float peak = 1.0f;
    if (p.normalize) {
      float m = 0.0f;
      for (int i = tid; i < p.n_samples; i += kThreads) m = fmaxf(m, fabsf(__ldg(x + i)));
      peak = block_max(m, red, tid);
      if (!(peak > 0.0f)) peak = 1.0f;     // reference would emit NaN (0/0); guarded, see DESIGN.md
    }
    __syncthreads();

    const int n_pairs = (p.W + 1) >> 1;
    for (int pr = 0; pr < n_pairs; ++pr) {
      const int t0 = 2 * pr, t1 = 2 * pr + 1;
      const int s0 = p.hop * t0 - (N >> 1);
      const int s1 = (t1 < p.W) ? p.hop * t1 - (N >> 1) : (1 << 30);
      // ---- first stage: radix 8 straight from global (frame + window), Ns = 1
      {
        const int nb = N >> 3;
        for (int j = tid; j < nb; j += kThreads) {
          float2 v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = j + r * nb;
            const float w = __ldg(p.window + n);
            const int i0 = s0 + n, i1 = s1 + n;
            float a = (i0 >= 0 && i0 < p.n_samples) ? __ldg(x + i0) : 0.0f;
            float c = (i1 >= 0 && i1 < p.n_samples) ? __ldg(x + i1) : 0.0f;
            if (p.normalize) { a = __fdiv_rn(a, peak); c = __fdiv_rn(c, peak); }
            v[r] = make_float2(w * a, w * c);
          }
          dft8(v);
#pragma unroll
          for (int r = 0; r < 8; ++r) buf0[padi(j * 8 + r)] = v[r];
        }
      }
      __syncthreads();
      // ---- remaining stages
      float2* src = buf0;
      float2* dst = buf1;
      int Ns = 8, rem = p.log2n - 3;
      while (rem > 0) {
        if (rem >= 3) { stockham_stage<8>(src, dst, tw, N, Ns, tid); Ns <<= 3; rem -= 3; }
        else if (rem == 2) { stockham_stage<4>(src, dst, tw, N, Ns, tid); Ns <<= 2; rem -= 2; }
        else { stockham_stage<2>(src, dst, tw, N, Ns, tid); Ns <<= 1; rem -= 1; }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
      }
      // src now holds Z[k] (natural order); dst is free -> power spectra of the two frames
      float* P0 = reinterpret_cast<float*>(dst);
      float* P1 = P0 + nbins;
      for (int k = tid; k < nbins; k += kThreads) {
        const float2 zk = src[padi(k)];
        const float2 zn = src[padi((N - k) & (N - 1))];
        const float ar = zk.x + zn.x, ai = zk.y - zn.y;
        const float br = zk.y + zn.y, bi = zk.x - zn.x;
        P0[k] = 0.25f * (ar * ar + ai * ai);
        P1[k] = 0.25f * (br * br + bi * bi);
      }
      __syncthreads();
      // ---- sparse mel projection: one warp per (frame, mel) pair
      const int warp = tid >> 5, lane = tid & 31;
      for (int q = warp; q < 2 * p.n_mels; q += kThreads / 32) {
        const int f = q >= p.n_mels;
        const int m = q - f * p.n_mels;
        const int t = t0 + f;
        if (t >= p.W) continue;
        const float* P = f ? P1 : P0;
        const int st = __ldg(p.mel_start + m), len = __ldg(p.mel_len + m), off = __ldg(p.mel_off + m);
        float acc = 0.0f;
        for (int i = lane; i < len; i += 32) acc = fmaf(__ldg(p.mel_w + off + i), P[st + i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) mel_s[m * p.W + t] = acc;
      }
      __syncthreads();
    }
    // ---- power_to_db(ref=max, amin, top_db)
    const int total = p.n_mels * p.W;
    float m = 0.0f;
    for (int i = tid; i < total; i += kThreads) m = fmaxf(m, mel_s[i]);
    const float ref = block_max(m, red, tid);
    // explicit _rn ops: an FMA contraction here would make the per-clip maximum land at +-1 ulp instead of 0 dB
    const float ref_db = __fmul_rn(10.0f, log10f(fmaxf(kAmin, ref)));
    float* __restrict__ o = p.out + (int64_t)b * total;
    for (int i = tid; i < total; i += kThreads) {
      float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(kAmin, mel_s[i]))), ref_db);
      o[i] = fmaxf(v, -kTopDb);
    }
    __syncthreads();
  }
}

}  // namespace

int ww_launch_logmel(ww_ctx* c, const float* clips, int64_t clip_stride, float* out, int B, int normalize,
                     cudaStream_t st) {
  if (B <= 0) return WW_OK;
  LogmelParams p;
  p.clips = clips; p.clip_stride = clip_stride; p.out = out; p.B = B; p.normalize = normalize;
  p.n_samples = c->cfg.n_samples; p.n_fft = c->cfg.n_fft; p.hop = c->cfg.hop_length; p.W = c->W;
  p.n_mels = c->cfg.n_mels;
  int l2 = 0; while ((1 << l2) < p.n_fft) ++l2;
  p.log2n = l2;
  p.window = c->d_window; p.twiddle = c->d_twiddle;
  p.mel_start = c->d_mel_start; p.mel_len = c->d_mel_len; p.mel_off = c->d_mel_off; p.mel_w = c->d_mel_w;
  const int N = p.n_fft;
  const int npad = N + (N >> 5) + 8;
  size_t smem = (size_t)(2 * npad + N) * sizeof(float2) + (size_t)p.n_mels * p.W * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    WW_CHECK(c, cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = c->sm_count * per_sm;
  if (grid > B) grid = B;
  ProfScope prof(c, WW_STAGE_LOGMEL, st);
  logmel_kernel<<<grid, kThreads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

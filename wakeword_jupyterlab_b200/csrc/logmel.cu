// K2: fused framing + Hann window + STFT power + mel filterbank + per-clip-max dB  (sm_100a)
//
// Replaces AudioProcessor.audio_to_mel (/root/reference/wakeword_training_script.py:85-101):
// librosa.feature.melspectrogram(center=True, pad_mode="constant", hann, power=2, slaney mel)
// followed by librosa.power_to_db(ref=np.max, amin=1e-10, top_db=80).
//
// One CTA (2 groups of 256 threads, 2 CTAs per SM) per clip; each group owns one in-place n_fft-point complex FFT
// buffer in shared memory and takes frame PAIRS round-robin: frame 2p is the real part, frame 2p+1 the
// imaginary part of one complex FFT (mixed radix 8/4/2 Stockham, register-staged so it runs in place);
// the two spectra are separated by Hermitian symmetry and squared in place ((P_even, P_odd) as one
// float2 per bin) and projected onto the banded slaney filterbank (2 lanes per mel band, both frames
// at once).  Twiddles come from compact per-stage tables (no strided shared-memory gathers); window,
// twiddles and the packed filterbank live in shared memory.  The n_mels x W mel power of the clip stays
// in shared memory, so the per-clip max that the dB conversion needs (ref=np.max) never costs a second
// pass over HBM.  Peak normalisation (normalize_audio) is folded in: |X/peak|^2 = |X|^2 / peak^2.
//
// HBM traffic per clip: n_samples*4 B in (frame overlap re-reads hit L2) + n_mels*W*4 B out
// = 74,240 B at the code preset.
#include "ctx.cuh"

namespace {

constexpr int kGroups = 2;
constexpr int kGT = 256;                   // threads per group (one FFT)
constexpr int kThreads = kGroups * kGT;
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

__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kGT) : "memory");
}

// In-place Stockham stage (radix R, Ns = product of the radices already applied) on one group's buffer.
// Butterfly j: v[r] = z[j + r*N/R] * w^(r),  w = exp(-2 pi i k / (Ns R)), k = j % Ns (compact table `ts[k]`);
// z[(j - k) * R + k + r * Ns] = DFT_R(v)[r].  Every thread reads its (<= 2) butterflies, the group syncs, then writes.
template <int R, int ITERS>
__device__ __forceinline__ void stockham_stage_i(float2* z, const float2* __restrict__ ts, int N, int Ns, int gt, int grp) {
  const int nb = N / R;
  float2 v[ITERS][R];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int j = gt + it * kGT;
    if (j < nb) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[it][r] = z[padi(j + r * nb)];
      const float2 w1 = ts[j & (Ns - 1)];
      float2 w = w1;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        v[it][r] = cmul(v[it][r], w);
        if (r + 1 < R) w = cmul(w, w1);
      }
      dftR<R>(v[it]);
    }
  }
  group_sync(grp);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int j = gt + it * kGT;
    if (j < nb) {
      const int k = j & (Ns - 1);
      const int base = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) z[padi(base + r * Ns)] = v[it][r];
    }
  }
  group_sync(grp);
}
template <int R>
__device__ __forceinline__ void stockham_stage(float2* z, const float2* __restrict__ ts, int N, int Ns, int gt, int grp) {
  // n_fft <= 2048 (checked in ww_create): a radix-8 stage has <= 256 butterflies, radix 4/2 at most 512
  if (R < 8 && N / R > kGT) stockham_stage_i<R, (R < 8 ? 2 : 1)>(z, ts, N, Ns, gt, grp);
  else stockham_stage_i<R, 1>(z, ts, N, Ns, gt, grp);
}

struct LogmelParams {
  const float* clips;
  int64_t clip_stride;
  float* out;
  int B, normalize;
  int n_samples, n_fft, log2n, hop, W, n_mels, mel_nnz;
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

__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(LogmelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft;
  const int npad = N + (N >> 5) + 8;
  float2* zbuf = reinterpret_cast<float2*>(smem_raw);          // [kGroups][npad]
  float2* tw = zbuf + kGroups * npad;                          // [N]      exp(-2 pi i t / N)
  float2* twc = tw + N;                                        // compact tables of the middle stages (< N/8 entries)
  float* win = reinterpret_cast<float*>(twc + N / 8);          // [N]
  float* melw = win + N;                                       // [mel_nnz]
  int* mst = reinterpret_cast<int*>(melw + p.mel_nnz);         // [3][n_mels] start, len, off
  float* mel_s = reinterpret_cast<float*>(mst + 3 * p.n_mels); // [n_mels * W]
  __shared__ float red[kThreads / 32];

  const int tid = threadIdx.x;
  const int grp = tid / kGT, gt = tid % kGT;
  const int nbins = N / 2 + 1;

  // ---- tables -> shared memory (once per CTA)
  for (int i = tid; i < N; i += kThreads) { tw[i] = p.twiddle[i]; win[i] = p.window[i]; }
  for (int i = tid; i < p.mel_nnz; i += kThreads) melw[i] = p.mel_w[i];
  for (int i = tid; i < p.n_mels; i += kThreads) {
    mst[i] = p.mel_start[i]; mst[p.n_mels + i] = p.mel_len[i]; mst[2 * p.n_mels + i] = p.mel_off[i];
  }
  {
    // compact per-stage twiddles: stage with Ns and radix R uses T[k * N/(Ns R)], k < Ns; stages are
    // (8, 8, ..., last) so Ns = 8, 64, 512, ...; the last stage (tstep == 1) reads `tw` directly.
    int off = 0, Ns = 8, rem = p.log2n - 3;
    while (rem > 0) {
      const int lg = rem >= 3 ? 3 : rem;
      const int tstep = N / (Ns << lg);
      if (tstep > 1) {
        for (int k = tid; k < Ns; k += kThreads) twc[off + k] = p.twiddle[k * tstep];
        off += Ns;
      }
      Ns <<= lg; rem -= lg;
    }
  }
  __syncthreads();

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* __restrict__ x = p.clips + (int64_t)b * p.clip_stride;
    float inv_peak = 1.0f;
    if (p.normalize) {
      float m = 0.0f;
      for (int i = tid; i < p.n_samples; i += kThreads) m = fmaxf(m, fabsf(__ldg(x + i)));
      const float peak = block_max(m, red, tid);
      inv_peak = (peak > 0.0f) ? 1.0f / peak : 1.0f;   // silent clip: reference gives NaN (0/0); guarded, see DESIGN.md
    }
    float2* z = zbuf + grp * npad;
    const int n_pairs = (p.W + 1) >> 1;
    for (int pr = grp; pr < n_pairs; pr += kGroups) {
      const int t0 = 2 * pr, t1 = 2 * pr + 1;
      const int s0 = p.hop * t0 - (N >> 1);
      const int s1 = (t1 < p.W) ? p.hop * t1 - (N >> 1) : (1 << 30);
      // ---- first stage: radix 8 straight from global (framing + window), Ns = 1
      {
        const int nb = N >> 3;
        for (int j = gt; j < nb; j += kGT) {
          float2 v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = j + r * nb;
            const float w = win[n];
            const int i0 = s0 + n, i1 = s1 + n;
            const float a = (i0 >= 0 && i0 < p.n_samples) ? __ldg(x + i0) : 0.0f;
            const float c = (i1 >= 0 && i1 < p.n_samples) ? __ldg(x + i1) : 0.0f;
            v[r] = make_float2(w * a, w * c);
          }
          dft8(v);
#pragma unroll
          for (int r = 0; r < 8; ++r) z[padi(j * 8 + r)] = v[r];
        }
      }
      group_sync(grp);
      // ---- remaining stages (in place)
      {
        int Ns = 8, rem = p.log2n - 3, off = 0;
        while (rem > 0) {
          const int lg = rem >= 3 ? 3 : rem;
          const int tstep = N / (Ns << lg);
          const float2* ts = (tstep > 1) ? (twc + off) : tw;
          if (lg == 3) stockham_stage<8>(z, ts, N, Ns, gt, grp);
          else if (lg == 2) stockham_stage<4>(z, ts, N, Ns, gt, grp);
          else stockham_stage<2>(z, ts, N, Ns, gt, grp);
          if (tstep > 1) off += Ns;
          Ns <<= lg; rem -= lg;
        }
      }
      // ---- Hermitian split + power, in place: z[k] <- (|X_even[k]|^2, |X_odd[k]|^2) for k <= N/2.
      // Bin k reads z[k] and z[N-k] and is the only reader of both, so no sync is needed before the store.
      for (int k = gt; k < nbins; k += kGT) {
        const float2 zk = z[padi(k)];
        const float2 zn = z[padi((N - k) & (N - 1))];
        const float ar = zk.x + zn.x, ai = zk.y - zn.y;
        const float br = zk.y + zn.y, bi = zk.x - zn.x;
        z[padi(k)] = make_float2(0.25f * (ar * ar + ai * ai), 0.25f * (br * br + bi * bi));
      }
      group_sync(grp);
      // ---- banded mel projection: 2 lanes per band, both frames at once
      for (int mb = 0; mb < p.n_mels; mb += kGT / 2) {
        const int m = mb + (gt >> 1), half = gt & 1;
        float a0 = 0.0f, a1 = 0.0f;
        if (m < p.n_mels) {
          const int st = mst[m], len = mst[p.n_mels + m], off = mst[2 * p.n_mels + m];
#pragma unroll 4
          for (int i = half; i < len; i += 2) {
            const float w = melw[off + i];
            const float2 pw = z[padi(st + i)];
            a0 = fmaf(w, pw.x, a0);
            a1 = fmaf(w, pw.y, a1);
          }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        if (m < p.n_mels && half == 0) {
          mel_s[m * p.W + t0] = a0 * inv_peak * inv_peak;
          if (t1 < p.W) mel_s[m * p.W + t1] = a1 * inv_peak * inv_peak;
        }
      }
      group_sync(grp);
    }
    __syncthreads();
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
  p.n_mels = c->cfg.n_mels; p.mel_nnz = c->mel_nnz;
  int l2 = 0; while ((1 << l2) < p.n_fft) ++l2;
  p.log2n = l2;
  p.window = c->d_window; p.twiddle = c->d_twiddle;
  p.mel_start = c->d_mel_start; p.mel_len = c->d_mel_len; p.mel_off = c->d_mel_off; p.mel_w = c->d_mel_w;
  const int N = p.n_fft;
  const int npad = N + (N >> 5) + 8;
  size_t smem = (size_t)(kGroups * npad + N + N / 8) * sizeof(float2) + (size_t)N * sizeof(float) +
                (size_t)p.mel_nnz * sizeof(float) + (size_t)3 * p.n_mels * sizeof(int) +
                (size_t)p.n_mels * p.W * sizeof(float) + 16;
  if (smem > 227 * 1024) { c->set_error("ww_logmel: configuration exceeds shared memory"); return WW_ERR_INVALID; }
  static size_t configured = 0;
  if (smem > configured) {
    WW_CHECK(c, cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int grid = c->sm_count * per_sm;
  if (grid > B) grid = B;
  ProfScope prof(c, WW_STAGE_LOGMEL, st);
  logmel_kernel<<<grid, kThreads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

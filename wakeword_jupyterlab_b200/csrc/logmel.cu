// K2: fused framing + Hann window + STFT power + mel filterbank + per-clip-max dB  (sm_100a)
//
// Replaces AudioProcessor.audio_to_mel (/root/reference/wakeword_training_script.py:85-101):
// librosa.feature.melspectrogram(center=True, pad_mode="constant", hann, power=2, slaney mel)
// followed by librosa.power_to_db(ref=np.max, amin=1e-10, top_db=80).
//
// One CTA (2 groups of 256 threads, 2 CTAs per SM) per clip; each group owns one in-place n_fft-point
// complex FFT buffer in shared memory and takes frame PAIRS round-robin: frame 2p is the real part,
// frame 2p+1 the imaginary part of one complex FFT (mixed radix 8/4/2 Stockham, register-staged so it
// runs in place; n_fft is a template parameter so every index is a shift/constant); the two spectra are
// separated by Hermitian symmetry and squared in place ((P_even, P_odd) as one float2 per bin) and
// projected onto the banded slaney filterbank (2 lanes per mel band, both frames at once).  Twiddles
// come from compact per-stage tables (no strided shared-memory gathers); window, twiddles and the
// packed filterbank live in shared memory.  The n_mels x W mel power of the clip stays in shared
// memory, so the per-clip max that the dB conversion needs (ref=np.max) never costs a second pass over
// HBM.  Peak normalisation (normalize_audio) is folded in: |X/peak|^2 = |X|^2 / peak^2.
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

// Physical slot of logical FFT element i in a buffer WRITTEN by the stage whose NS is NSW: an XOR swizzle (a bijection
// inside every aligned run of 16 elements, so the next stage's loads of 16 consecutive butterflies stay conflict-free and
// no padding is needed) that spreads the strided stores of the writing stage over all banks:
//   NSW = 1 (radix-8 from global, thread t writes 8t .. 8t+7):  i ^ ((i >> 4) & 7)       -> 16 lanes hit 16 bank pairs
//   NSW = 8 (thread (a, k) writes 64a + 8r + k):                 i ^ (((i >> 6) & 1) << 3) -> runs of a, a+1 are 16 banks apart
//   NSW >= 32: the stores of a warp are already consecutive, no swizzle.
template <int NSW> __device__ __forceinline__ int lay(int i) {
  if constexpr (NSW == 1) return i ^ ((i >> 4) & 7);
  else if constexpr (NSW == 8) return i ^ (((i >> 6) & 1) << 3);
  else return i;
}

// complex add / sub as ONE packed fp32x2 instruction (two IEEE adds; a scalar FADD only issues every other cycle)
__device__ __forceinline__ unsigned long long as_u64(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 as_f2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_f2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(as_u64(a)), "l"(as_u64(b)));
  return as_f2(r);
}
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

// Stockham stage (radix R, NS = product of the radices already applied) from one buffer of the group into the other
// (ping-pong: one barrier per stage).
// Butterfly j: v[r] = zin[j + r*N/R] * w^r,  w = exp(-2 pi i k / (NS R)), k = j % NS (compact table `ts[k]`);
// zout[(j - k) * R + k + r * NS] = DFT_R(v)[r].   zin has the layout of the previous stage (NS / its radix), zout lay<NS>.
template <int N, int R, int NS, int NSPREV>
__device__ __forceinline__ void stockham_stage(const float2* zin, float2* zout, const float2* __restrict__ ts, int gt, int grp) {
  constexpr int NB = N / R;                               // butterflies (multiple of 32)
  constexpr int ITERS = NB > kGT ? NB / kGT : 1;
  float2 v[ITERS][R];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int j = gt + it * kGT;
    if (NB >= kGT || j < NB) {
      if constexpr (NB % 128 == 0) {                      // the swizzle does not depend on r
        const int rb = lay<NSPREV>(j);
#pragma unroll
        for (int r = 0; r < R; ++r) v[it][r] = zin[rb + r * NB];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) v[it][r] = zin[lay<NSPREV>(j + r * NB)];
      }
      const float2 w1 = ts[j & (NS - 1)];
      float2 w = w1;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        v[it][r] = cmul(v[it][r], w);
        if (r + 1 < R) w = cmul(w, w1);
      }
      dftR<R>(v[it]);
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int j = gt + it * kGT;
    if (NB >= kGT || j < NB) {
      const int k = j & (NS - 1);
      const int base = (j - k) * R + k;
      if constexpr (NS == 8 && R == 8) {
        const int sw = ((base >> 6) & 1) << 3;            // base = 64a + k: bit 3 (= r & 1) flips for odd a
#pragma unroll
        for (int r = 0; r < R; ++r) zout[(base + r * NS) ^ sw] = v[it][r];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) zout[lay<NS>(base + r * NS)] = v[it][r];
      }
    }
  }
  group_sync(grp);
}

// (|X_even[k]|^2, |X_odd[k]|^2) of the two real frames packed into one complex FFT, from Z[k] and Z[N-k]
__device__ __forceinline__ float2 pair_power(float2 zk, float2 zn) {
  const float ar = zk.x + zn.x, ai = zk.y - zn.y;
  const float br = zk.y + zn.y, bi = zk.x - zn.x;
  return make_float2(0.25f * (ar * ar + ai * ai), 0.25f * (br * br + bi * bi));
}

// LAST Stockham stage fused with the Hermitian split, for N / R == 2 kGT (two butterflies per thread, NS == N / R):
// thread t takes the butterflies j and NS - j (t = 0: the two self-paired ones, 0 and NS / 2), so the outputs
// X[j + r NS] and their mirror bins X[N - (j + r NS)] = X[(NS - j) + (R-1-r) NS] are in the SAME thread: the spectrum is
// never stored, only the powers zout[k] = (|X_even[k]|^2, |X_odd[k]|^2), k <= N/2.
template <int N, int R, int NS, int NSPREV>
__device__ __forceinline__ void last_stage_power(const float2* zin, float2* zout, const float2* __restrict__ ts, int gt, int grp) {
  static_assert(N / R == 2 * kGT && NS == N / R && NSPREV >= 32, "fused last stage: two butterflies per thread");
  const bool t0 = gt == 0;
  const int j0 = gt, j1 = t0 ? NS / 2 : NS - gt;
  float2 a[R], b[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { a[r] = zin[j0 + r * NS]; b[r] = zin[j1 + r * NS]; }
  {
    const float2 wa1 = ts[j0], wb1 = ts[j1];
    float2 wa = wa1, wb = wb1;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      a[r] = cmul(a[r], wa);
      b[r] = cmul(b[r], wb);
      if (r + 1 < R) { wa = cmul(wa, wa1); wb = cmul(wb, wb1); }
    }
  }
  dftR<R>(a);
  dftR<R>(b);
  // a[r] = X[j0 + r NS], b[r] = X[j1 + r NS]
#pragma unroll
  for (int r = 0; r < R / 2; ++r) {
    // t > 0: mirror of j0 + r NS is j1 + (R-1-r) NS and vice versa;  t = 0: 0 + r NS <-> 0 + ((R-r) % R) NS, NS/2 + r NS <-> NS/2 + (R-1-r) NS
    const float2 ma = t0 ? a[(R - r) % R] : b[R - 1 - r];
    const float2 mb = t0 ? b[R - 1 - r] : a[R - 1 - r];
    zout[j0 + r * NS] = pair_power(a[r], ma);
    zout[j1 + r * NS] = pair_power(b[r], mb);
  }
  if (t0) zout[N / 2] = pair_power(a[R / 2], a[R / 2]);      // Nyquist bin (its own mirror)
  group_sync(grp);
}

template <int LOG2N> __host__ __device__ constexpr bool fused_last_stage() {
  constexpr int rem = LOG2N - 3, lg = rem % 3 == 0 ? 3 : rem % 3;
  return rem > 0 && ((1 << LOG2N) >> lg) == 2 * kGT;
}

// remaining stages after the first radix-8 one: radices 8, 8, ..., then 4 or 2; compact twiddle tables
// (stage table k -> T[k * N / (NS R)]) are packed one after the other in `twc`; the last stage (step 1) reads `tw`.
// returns the buffer that holds the spectrum
template <int N, int NS, int REM, int OFF, int NSPREV>
__device__ __forceinline__ float2* run_stages(float2* zin, float2* zout, const float2* tw, const float2* twc, int gt, int grp) {
  if constexpr (REM > 0) {
    constexpr int LG = REM >= 3 ? 3 : REM;
    constexpr int R = 1 << LG;
    constexpr int TSTEP = N / (NS * R);
    if constexpr (REM == LG && N / R == 2 * kGT) {
      last_stage_power<N, R, NS, NSPREV>(zin, zout, TSTEP > 1 ? twc + OFF : tw, gt, grp);
      return zout;                                          // holds the POWERS of bins 0 .. N/2
    } else {
      stockham_stage<N, R, NS, NSPREV>(zin, zout, TSTEP > 1 ? twc + OFF : tw, gt, grp);
    }
    return run_stages<N, NS * R, REM - LG, OFF + (TSTEP > 1 ? NS : 0), NS>(zout, zin, tw, twc, gt, grp);
  } else {
    static_assert(NSPREV >= 32, "the last stage must leave the spectrum unskewed");
    return zin;
  }
}

// input sample -> float: fp32 as is, int16 PCM as s / 32768 (what librosa.load returns for a 16-bit WAV)
__device__ __forceinline__ float ldin(const float* x, int i) { return __ldg(x + i); }
__device__ __forceinline__ float ldin(const int16_t* x, int i) { return (float)__ldg(x + i) * (1.0f / 32768.0f); }

struct LogmelParams {
  const void* clips;
  int64_t clip_stride;
  LogmelOut out;
  int B, normalize;
  int n_samples, hop, W, n_mels, mel_nnz;
  const float* window;
  const float2* twiddle;
  const int* mel_start;
  const int* mel_len;
  const int* mel_off;
  const float* mel_w;
  // Streaming frame reuse.  mode 0: plain.  mode 1 (cache build): "clip" b is a block of W
  // consecutive cache frames, frame t starts at sample (b W + t) cache_g of `clips`, no centre offset; the raw mel
  // energies go to cache[(b W + t)][n_mels].  mode 2 (window): only the frame pairs that touch the zero padding of the
  // window are transformed; the interior frames come from the cache (energies of the raw signal, scaled by 1/peak^2
  // like every other frame), the window peak from block maxima of |x|.
  int mode;
  float* cache;            // [n_cache][n_mels]
  int64_t n_cache;         // cache frames
  int cache_g;             // samples between consecutive cache frames
  int64_t abs_start0;      // mode 2: absolute sample index of window 0 of this launch
  const float* blockmax;   // mode 2: max |x| over consecutive blocks of bm_block samples
  int bm_block;
  int64_t n_total;         // mode 1: samples available from `clips`
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

template <int LOG2N, typename TIn>
__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const LogmelParams p) {
  constexpr int N = 1 << LOG2N;
  constexpr int NPAD = N + 8;
  constexpr int NBINS = N / 2 + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* zbuf = reinterpret_cast<float2*>(smem_raw);          // [kGroups][2][NPAD]: ping-pong FFT buffers
  float2* tw = zbuf + kGroups * 2 * NPAD;                      // [N]      exp(-2 pi i t / N)
  float2* twc = tw + N;                                        // compact tables of the middle stages (< N/8 entries)
  const float* __restrict__ win = p.window;                    // [N] read once per frame pair through L1
  float* melw = reinterpret_cast<float*>(twc + N / 8);         // [mel_nnz]
  int* mst = reinterpret_cast<int*>(melw + p.mel_nnz);         // [3][n_mels] start, len, off
  float* mel_s = reinterpret_cast<float*>(mst + 3 * p.n_mels); // [n_mels][MP], odd pitch MP = W | 1: the per-band stores of a warp hit different banks
  __shared__ float red[kThreads / 32];

  const int tid = threadIdx.x;
  const int grp = tid / kGT, gt = tid % kGT;
  const int n_samples = p.n_samples, hop = p.hop, W = p.W, n_mels = p.n_mels;
  const int MP = W | 1;

  // ---- tables -> shared memory (once per CTA)
  for (int i = tid; i < N; i += kThreads) tw[i] = p.twiddle[i];
  for (int i = tid; i < p.mel_nnz; i += kThreads) melw[i] = p.mel_w[i];
  for (int i = tid; i < n_mels; i += kThreads) {
    mst[i] = p.mel_start[i]; mst[n_mels + i] = p.mel_len[i]; mst[2 * n_mels + i] = p.mel_off[i];
  }
  {
    int off = 0, Ns = 8, rem = LOG2N - 3;
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
  // the eight window taps this thread applies in the first stage are the same for every frame: keep them in registers
  float wv[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) wv[r] = ((N >> 3) >= kGT || gt < (N >> 3)) ? __ldg(win + gt + r * (N >> 3)) : 0.0f;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const TIn* __restrict__ x = static_cast<const TIn*>(p.clips) + (int64_t)b * p.clip_stride;
    float inv_peak = 1.0f;
    if (p.mode == 2) {
      const int64_t j0 = (p.abs_start0 + (int64_t)b * p.clip_stride) / p.bm_block;
      float m = 0.0f;
      for (int i = tid; i < n_samples / p.bm_block; i += kThreads) m = fmaxf(m, __ldg(p.blockmax + j0 + i));
      const float peak = block_max(m, red, tid);
      inv_peak = (peak > 0.0f) ? 1.0f / peak : 1.0f;
    } else if (p.normalize) {
      float m = 0.0f;
      for (int i = tid; i < n_samples; i += kThreads) m = fmaxf(m, fabsf(ldin(x, i)));
      const float peak = block_max(m, red, tid);
      inv_peak = (peak > 0.0f) ? 1.0f / peak : 1.0f;   // silent clip: reference gives NaN (0/0); guarded, see DESIGN.md
    }
    const float scale = inv_peak * inv_peak;
    float2* za = zbuf + grp * 2 * NPAD;
    float2* zb = za + NPAD;
    const int n_pairs = (W + 1) >> 1;
    // frames [t_lo, t_hi] lie fully inside the window; mode 2 transforms only the pairs [0, pe_lo) and [pe_hi, n_pairs)
    const int t_lo = ((N >> 1) + hop - 1) / hop, t_hi = (n_samples - (N >> 1)) / hop;
    const int pe_lo = p.mode == 2 ? min((t_lo + 1) >> 1, n_pairs) : n_pairs;
    const int pe_hi = p.mode == 2 ? max(min((t_hi + 1) >> 1, n_pairs), pe_lo) : n_pairs;
    const int n_do = pe_lo + (n_pairs - pe_hi);
    const int f_hop = p.mode == 1 ? p.cache_g : hop, f_off = p.mode == 1 ? 0 : -(N >> 1);
    // valid sample indices relative to x: [0, limit)
    const long long left = (long long)p.n_total - (long long)b * p.clip_stride;
    const unsigned limit = p.mode == 1 ? (unsigned)(left < (1ll << 30) ? (left > 0 ? left : 0) : (1ll << 30)) : (unsigned)n_samples;
    // raw samples of this thread's first-stage inputs (8 per frame); fetched one pair AHEAD, before the mel projection of
    // the current pair, so the global-load latency hides behind it
    float ra[8], rc[8];
    auto fetch = [&](int e) {
      constexpr int NB = N >> 3;
      const int pr = e < pe_lo ? e : pe_hi + (e - pe_lo);
      const int t1 = 2 * pr + 1;
      const int s0 = f_hop * (2 * pr) + f_off;
      const int s1 = (t1 < W) ? f_hop * t1 + f_off : (1 << 30);
      if (NB >= kGT || gt < NB) {
        if (s0 >= 0 && t1 < W && (unsigned)(s1 + N) <= limit) {
          // both frames lie inside the clip (all but the first and last pair or two): no bounds checks
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = gt + r * NB;
            ra[r] = ldin(x, s0 + n);
            rc[r] = ldin(x, s1 + n);
          }
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int n = gt + r * NB;
            const int i0 = s0 + n, i1 = s1 + n;
            ra[r] = ((unsigned)i0 < limit) ? ldin(x, i0) : 0.0f;
            rc[r] = ((unsigned)i1 < limit) ? ldin(x, i1) : 0.0f;
          }
        }
      }
    };
    if (grp < n_do) fetch(grp);
    for (int e = grp; e < n_do; e += kGroups) {
      const int pr = e < pe_lo ? e : pe_hi + (e - pe_lo);
      const int t0 = 2 * pr, t1 = 2 * pr + 1;
      // ---- first stage: radix 8 on the prefetched samples (framing + window), NS = 1
      {
        constexpr int NB = N >> 3;
        if (NB >= kGT || gt < NB) {
          float2 v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) v[r] = make_float2(wv[r] * ra[r], wv[r] * rc[r]);
          dft8(v);
          const int wb = gt * 8, sw = (gt >> 1) & 7;       // lay<1>(8 gt + r) = 8 gt + (r ^ ((gt >> 1) & 7))
#pragma unroll
          for (int r = 0; r < 8; ++r) za[wb + (r ^ sw)] = v[r];
        }
      }
      group_sync(grp);
      // Each stage reads one buffer and writes the other.  The mel projection of the previous pair read the buffer the
      // spectrum ends in; it is next written two barriers from here at the earliest, so no barrier is needed after it.
      float2* z = run_stages<N, 8, LOG2N - 3, 0, 1>(za, zb, tw, twc, gt, grp);
      // ---- Hermitian split + power: z[k] <- (|X_even[k]|^2, |X_odd[k]|^2) for k <= N/2: done inside the last stage
      // when that stage has two butterflies per thread (n_fft 2048 and 1024), else in place here (bin k reads z[k] and
      // z[N-k] and is the only reader of both, so no sync is needed before the store).
      if constexpr (!fused_last_stage<LOG2N>()) {
#pragma unroll
        for (int k = gt; k < NBINS; k += kGT) z[k] = pair_power(z[k], z[(N - k) & (N - 1)]);
        group_sync(grp);
      }
      if (e + kGroups < n_do) fetch(e + kGroups);
      // ---- banded mel projection: 4 lanes per band (4 consecutive bins = one 32-byte run per band and load: fewer
      // bank conflicts than 2 x 16 bands per warp), both frames at once
      for (int mb = 0; mb < n_mels; mb += kGT / 4) {
        const int m = mb + (gt >> 2), q = gt & 3;
        float a0 = 0.0f, a1 = 0.0f;
        if (m < n_mels) {
          const int st = mst[m], len = mst[n_mels + m], off = mst[2 * n_mels + m];
#pragma unroll 4
          for (int i = q; i < len; i += 4) {
            const float w = melw[off + i];
            const float2 pw = z[st + i];
            a0 = fmaf(w, pw.x, a0);
            a1 = fmaf(w, pw.y, a1);
          }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        if (m < n_mels && q == 0) {
          mel_s[m * MP + t0] = a0 * scale;
          if (t1 < W) mel_s[m * MP + t1] = a1 * scale;
        }
      }
      if (((LOG2N - 3 + 2) / 3) < 2) group_sync(grp);      // fewer than two ping-pong stages: the next pair would overwrite z
    }
    if (p.mode == 2) {
      // interior frames from the cache: frame t of this window starts at absolute sample a = start + hop t - N/2
      const int t_a = 2 * pe_lo, n_int = 2 * pe_hi - t_a;
      const int64_t a0 = p.abs_start0 + (int64_t)b * p.clip_stride - (N >> 1);
      for (int i = tid; i < n_int * n_mels; i += kThreads) {
        const int tt = i / n_mels, m = i - tt * n_mels, t = t_a + tt;
        if (t < W) {
          const int64_t idx = (a0 + (int64_t)hop * t) / p.cache_g;
          mel_s[m * MP + t] = __ldg(p.cache + idx * n_mels + m) * scale;
        }
      }
    }
    __syncthreads();
    if (p.mode == 1) {
      // cache build: raw mel energies, frame-major
      for (int i = tid; i < W * n_mels; i += kThreads) {
        const int t = i / n_mels, m = i - t * n_mels;
        const int64_t f = (int64_t)b * W + t;
        if (f < p.n_cache) p.cache[f * n_mels + m] = mel_s[m * MP + t];
      }
      __syncthreads();
      continue;
    }
    // ---- power_to_db(ref=max, amin, top_db)
    const int total = n_mels * W;
    float m = 0.0f;
    const uint32_t magicW = 0xffffffffu / (uint32_t)W + 1u;             // i / W == umulhi(i, magicW) for i, W < 2^16
    for (int i = tid; i < total; i += kThreads) m = fmaxf(m, mel_s[i + (int)__umulhi((uint32_t)i, magicW) * (MP - W)]);
    const float ref = block_max(m, red, tid);
    // explicit _rn ops: an FMA contraction here would make the per-clip maximum land at +-1 ulp instead of 0 dB
    const float ref_db = __fmul_rn(10.0f, log10f(fmaxf(kAmin, ref)));
    float* __restrict__ o = p.out.ptr + (int64_t)b * p.out.stride + p.out.off;
    for (int i = tid; i < total; i += kThreads) {
      const int m = (int)__umulhi((uint32_t)i, magicW);
      float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(kAmin, mel_s[i + m * (MP - W)]))), ref_db);
      o[m * p.out.pitch + (i - m * W)] = fmaxf(v, -kTopDb);
    }
    __syncthreads();
  }
}

template <int LOG2N, typename TIn>
int launch_tt(ww_ctx* c, const LogmelParams& p, size_t smem, int grid, cudaStream_t st) {
  // per device, not per process: set on every call (cheap)
  WW_CHECK(c, cudaFuncSetAttribute(logmel_kernel<LOG2N, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(c, WW_STAGE_LOGMEL, st);
  logmel_kernel<LOG2N, TIn><<<grid, kThreads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}
template <int LOG2N>
int launch_t(ww_ctx* c, const LogmelParams& p, int pcm16, size_t smem, int grid, cudaStream_t st) {
  return pcm16 ? launch_tt<LOG2N, int16_t>(c, p, smem, grid, st) : launch_tt<LOG2N, float>(c, p, smem, grid, st);
}

}  // namespace

int ww_launch_logmel(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, float* out, int B, int normalize,
                     cudaStream_t st) {
  return ww_launch_logmel_ex(c, clips, pcm16, clip_stride, LogmelOut{out, c->W, 0, (int64_t)c->cfg.n_mels * c->W}, B,
                             normalize, st);
}

int ww_launch_logmel_ex(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B,
                        int normalize, cudaStream_t st) {
  return ww_launch_logmel_stream(c, clips, pcm16, clip_stride, out, B, normalize, nullptr, st);
}

namespace {
__global__ void blockmax_kernel(const void* __restrict__ x, int pcm16, int64_t n, int block, float* __restrict__ out, int64_t n_blocks) {
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n_blocks) return;
  const int lane = threadIdx.x & 31;
  float m = 0.0f;
  for (int i = lane; i < block; i += 32) {
    const int64_t k = j * block + i;
    if (k < n) m = fmaxf(m, fabsf(pcm16 ? (float)static_cast<const int16_t*>(x)[k] * (1.0f / 32768.0f) : static_cast<const float*>(x)[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[j] = m;
}
}  // namespace

int ww_launch_blockmax(ww_ctx* c, const void* x, int pcm16, int64_t n, int block, float* out, int64_t n_blocks, cudaStream_t st) {
  if (n_blocks <= 0) return WW_OK;
  blockmax_kernel<<<(unsigned)((n_blocks + 7) / 8), 256, 0, st>>>(x, pcm16, n, block, out, n_blocks);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

// sp == nullptr: plain log-mel.  sp->mode 1: build the frame-energy cache (clips = the whole signal, B = blocks of W
// cache frames).  sp->mode 2: sliding windows whose interior frames come from the cache.
int ww_launch_logmel_stream(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B,
                            int normalize, const StreamReuse* sp, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  if (!sp) {             // plain log-mel: the tensor-core kernel where it applies (the reference's own preset)
    const int rc = ww_launch_logmel_tc(c, clips, pcm16, clip_stride, out, B, normalize, st);
    if (rc <= 0) return rc;
  }
  LogmelParams p;
  p.mode = sp ? sp->mode : 0;
  p.cache = sp ? sp->cache : nullptr; p.n_cache = sp ? sp->n_cache : 0; p.cache_g = sp ? sp->cache_g : 1;
  p.abs_start0 = sp ? sp->abs_start0 : 0; p.blockmax = sp ? sp->blockmax : nullptr; p.bm_block = sp ? sp->bm_block : 1;
  p.n_total = sp ? sp->n_total : 0;
  p.clips = clips; p.clip_stride = clip_stride; p.out = out; p.B = B; p.normalize = normalize;
  p.n_samples = c->cfg.n_samples; p.hop = c->cfg.hop_length; p.W = c->W;
  p.n_mels = c->cfg.n_mels; p.mel_nnz = c->mel_nnz;
  p.window = c->d_window; p.twiddle = c->d_twiddle;
  p.mel_start = c->d_mel_start; p.mel_len = c->d_mel_len; p.mel_off = c->d_mel_off; p.mel_w = c->d_mel_w;
  const int N = c->cfg.n_fft;
  const int npad = N + 8;
  size_t smem = (size_t)(kGroups * 2 * npad + N + N / 8) * sizeof(float2) +
                (size_t)p.mel_nnz * sizeof(float) + (size_t)3 * p.n_mels * sizeof(int) +
                (size_t)p.n_mels * (p.W | 1) * sizeof(float) + 16;
  if (smem > 227 * 1024) { c->set_error("ww_logmel: configuration exceeds shared memory"); return WW_ERR_INVALID; }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  int grid = c->sm_count * per_sm;
  if (grid > B) grid = B;
  switch (N) {
    case 256: return launch_t<8>(c, p, pcm16, smem, grid, st);
    case 512: return launch_t<9>(c, p, pcm16, smem, grid, st);
    case 1024: return launch_t<10>(c, p, pcm16, smem, grid, st);
    case 2048: return launch_t<11>(c, p, pcm16, smem, grid, st);
    default: c->set_error("ww_logmel: n_fft must be 256, 512, 1024 or 2048"); return WW_ERR_INVALID;
  }
}

// K1: batched augmentation / conditioning  (sm_100a)
//
// Stage set of the north star, consuming HOST-drawn (seed-supplied) parameters only:
//   NORM_IN  peak normalise           AudioProcessor.normalize_audio  wakeword_training_script.py:73-76
//   SHIFT    circular shift (np.roll) augment_audio                   wakeword_training_script.py:106-108
//   SPEED    polyphase windowed-sinc resample + crop/zero-pad         :114-117 (speed stage) + :78-83
//   NOISE    snr_mixer(clean, bank segment, snr)                      stock/ms_snsd/MS-SNSD/audiolib.py:55-71
//   GAIN     linear gain
//   NORM_OUT peak normalise
// One CTA (1024 threads) per clip.  Each thread keeps its 16 samples in REGISTERS through every element-wise stage
// (normalise, SNR mix, gain); only the gather stage goes through shared memory: shift and speed change are a
// single gather (the resampler reads its taps through the roll index map).  HBM traffic is one read of the clip
// (+ the noise segment, re-read from L2 for the later passes) and one write: 128 KB (192 KB) per clip.
// Index arithmetic (roll source index, (q,p) phase decomposition, tap range, crop offset, output length) is
// integer-exact against oracle/augment.py; only exactly-zero taps of the polyphase table are skipped.
#include "ctx.cuh"

#include <algorithm>

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxPerThread = 16;     // n_samples <= kThreads * kMaxPerThread
constexpr int kTblWords = 2560;       // shared-memory room for one ratio's compact polyphase table

struct AugKParams {
  const float* clips;
  const float* bank;
  int bank_rows;
  int64_t bank_len;
  ww_aug a;
  float* out;
  int B, N;
  const RsDesc* rs_desc;
  int n_rs;
  const float* rs_kern;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// two independent block reductions at once (sum or max); ends with every thread holding both results
template <bool MAX>
__device__ __forceinline__ void block_reduce2(float& a, float& b, float* red, int tid) {
  a = MAX ? warp_max(a) : warp_sum(a);
  b = MAX ? warp_max(b) : warp_sum(b);
  __syncthreads();
  if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = b; }
  __syncthreads();
  const float ra = red[tid & 31], rb = red[32 + (tid & 31)];       // kThreads / 32 == 32 partials each
  a = MAX ? warp_max(ra) : warp_sum(ra);
  b = MAX ? warp_max(rb) : warp_sum(rb);
}

__global__ void __launch_bounds__(kThreads, 1) augment_kernel(AugKParams p) {
  extern __shared__ __align__(16) float cur[];   // [N] clip, then [kTblWords] polyphase table
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const int N = p.N;
  float* tbl = cur + ((N + 3) & ~3);

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const uint32_t flags = p.a.flags[b];
    const float* __restrict__ x = p.clips + (int64_t)b * N;
    float o[kMaxPerThread];          // this thread's samples i = tid + e * kThreads, live in registers from here on
    float m = 0.0f, dummy = 0.0f;
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) {
      const int i = tid + e * kThreads;
      o[e] = (i < N) ? __ldg(x + i) : 0.0f;
      m = fmaxf(m, fabsf(o[e]));
    }
    if (flags & WW_AUG_NORM_IN) {
      block_reduce2<true>(m, dummy, red, tid);
      if (m > 0.0f) {
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) o[e] = __fdiv_rn(o[e], m);
      }
    }

    if (flags & (WW_AUG_SHIFT | WW_AUG_SPEED)) {
      // ---- one gather through shared memory: out[i] = sum_k kern[ph][k] * rolled[x0 + k],  rolled[t] = in[(t - shift) mod N]
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) {
        const int i = tid + e * kThreads;
        if (i < N) cur[i] = o[e];
      }
      int s = (flags & WW_AUG_SHIFT) ? p.a.shift[b] % N : 0;
      if (s < 0) s += N;
      if (flags & WW_AUG_SPEED) {
        const int orig = p.a.rs_orig[b], neu = p.a.rs_new[b];
        int found = -1;
        for (int i = 0; i < p.n_rs; ++i)
          if (p.rs_desc[i].orig == orig && p.rs_desc[i].neu == neu) { found = i; break; }
        if (found < 0) {
          __syncthreads();
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = __int_as_float(0x7fc00000);   // loud: NaN clip
        } else {
          const RsDesc d = p.rs_desc[found];
          // compact polyphase table of this ratio -> shared memory ([n][nz] taps, [n] first tap, [n] count)
          const int tbl_words = d.n * d.nz + 2 * d.n;
          const bool in_smem = tbl_words <= kTblWords;
          const float* __restrict__ gk = p.rs_kern + d.offset;
          if (in_smem)
            for (int i = tid; i < tbl_words; i += kThreads) tbl[i] = __ldg(gk + i);
          __syncthreads();
          const float* kern = in_smem ? tbl : gk;
          const int* lo_t = reinterpret_cast<const int*>(kern + d.n * d.nz);
          const int* cnt_t = lo_t + d.n;
          const int out_len = (d.n * N + d.o - 1) / d.o;                             // ceil(n*N/o), < 2^31
          const int crop = (out_len > N) ? p.a.crop_off[b] : 0;
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) {
            const int i = tid + e * kThreads;
            float acc0 = 0.0f, acc1 = 0.0f;
            const int j = i + crop;                                                  // resampled-domain index
            if (i < N && j < out_len) {
              const int q = j / d.n, ph = j - q * d.n;                               // (q, p) phase decomposition
              const int x0 = q * d.o - d.width + lo_t[ph];                           // source index of the first non-zero tap
              const float* kr = kern + ph * d.nz;
              const int k0 = x0 < 0 ? -x0 : 0;
              const int k1 = min(cnt_t[ph], N - x0);
              int src = x0 + k0 - s;
              if (src < 0) src += N;
              int k = k0;
              if (src + (k1 - k0) <= N) {                                            // no wrap inside the tap window
                const float* sp = cur + src - k0;
                for (; k + 1 < k1; k += 2) {                                         // two independent chains
                  acc0 = fmaf(kr[k], sp[k], acc0);
                  acc1 = fmaf(kr[k + 1], sp[k + 1], acc1);
                }
                if (k < k1) acc0 = fmaf(kr[k], sp[k], acc0);
              } else {
                for (; k < k1; ++k) {
                  acc0 = fmaf(kr[k], cur[src], acc0);
                  src = (src + 1 >= N) ? src + 1 - N : src + 1;
                }
              }
            }
            o[e] = acc0 + acc1;
          }
        }
      } else {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) {
          const int i = tid + e * kThreads;
          int src = i - s;
          if (src < 0) src += N;
          o[e] = (i < N) ? cur[src] : 0.0f;                                         // bit-exact copy (np.roll)
        }
      }
    }

    if (flags & WW_AUG_NOISE) {
      const float* __restrict__ nz = p.bank + (int64_t)p.a.noise_idx[b] * p.bank_len + p.a.noise_off[b];
      const float snr = p.a.snr_db[b];
      const float target = 0.0562341325190349f;            // 10 ** (-25 / 20)
      float n[kMaxPerThread];
      float sc = 0.0f, sn = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) {
        const int i = tid + e * kThreads;
        n[e] = (i < N) ? __ldg(nz + i) : 0.0f;
        sc = fmaf(o[e], o[e], sc);
        sn = fmaf(n[e], n[e], sn);
      }
      block_reduce2<false>(sc, sn, red, tid);
      const float scalarclean = target / sqrtf(sc / (float)N), scalarnoise = target / sqrtf(sn / (float)N);
      // the reference re-measures both RMS values after scaling (audiolib.py:60,65)
      float sc2 = 0.0f, sn2 = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) {
        o[e] *= scalarclean;
        n[e] *= scalarnoise;
        sc2 = fmaf(o[e], o[e], sc2);
        sn2 = fmaf(n[e], n[e], sn2);
      }
      block_reduce2<false>(sc2, sn2, red, tid);
      const float rc2 = sqrtf(sc2 / (float)N), rn2 = sqrtf(sn2 / (float)N);
      const float noisescalar = sqrtf(rc2 / exp10f(snr / 20.0f) / rn2);          // audiolib.py:68 (sqrt quirk kept)
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) o[e] = o[e] + n[e] * noisescalar;
    }
    if (flags & WW_AUG_GAIN) {
      const float g = p.a.gain[b];
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) o[e] *= g;
    }
    if (flags & WW_AUG_NORM_OUT) {
      float mo = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) mo = fmaxf(mo, fabsf(o[e]));
      block_reduce2<true>(mo, dummy, red, tid);
      if (mo > 0.0f) {
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) o[e] = __fdiv_rn(o[e], mo);
      }
    }
    float* __restrict__ dst = p.out + (int64_t)b * N;
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) {
      const int i = tid + e * kThreads;
      if (i < N) dst[i] = o[e];
    }
    __syncthreads();      // cur / tbl are reused by the next clip
  }
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* out) {
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}
__global__ void divide_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, const unsigned int* peak_bits) {
  const float peak = __uint_as_float(*peak_bits);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = peak > 0.0f ? __fdiv_rn(x[i], peak) : x[i];
}

}  // namespace

int ww_launch_normalize(ww_ctx* c, const float* in, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return WW_OK;
  if (!c->d_scalar) WW_CHECK(c, cudaMalloc((void**)&c->d_scalar, sizeof(unsigned int)));
  WW_CHECK(c, cudaMemsetAsync(c->d_scalar, 0, sizeof(unsigned int), st));
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
  absmax_kernel<<<grid, 256, 0, st>>>(in, n, c->d_scalar);
  WW_LAUNCH_CHECK(c);
  divide_kernel<<<grid, 256, 0, st>>>(in, out, n, c->d_scalar);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

int ww_launch_augment(ww_ctx* c, const float* clips, const float* bank, int bank_rows, int64_t bank_len,
                      const ww_aug* a, float* out, int B, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  AugKParams p;
  p.clips = clips; p.bank = bank; p.bank_rows = bank_rows; p.bank_len = bank_len;
  p.a = *a; p.out = out; p.B = B; p.N = c->cfg.n_samples;
  p.rs_desc = c->d_rs_desc; p.n_rs = (int)c->rs_tables.size(); p.rs_kern = c->d_rs_kern;
  if (p.N > kThreads * kMaxPerThread) {
    c->set_error("ww_augment: n_samples too large (max 16384 samples per clip)");
    return WW_ERR_INVALID;
  }
  size_t smem = (size_t)(((p.N + 3) & ~3) + kTblWords) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    WW_CHECK(c, cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = std::min(c->sm_count, B);
  ProfScope prof(c, WW_STAGE_AUGMENT, st);
  augment_kernel<<<grid, kThreads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  return WW_OK;
}

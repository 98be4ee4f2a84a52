// K1: batched augmentation / conditioning  (sm_100a)
//
// Stage set of the north star, consuming HOST-drawn (seed-supplied) parameters only:
//   NORM_IN  peak normalise           AudioProcessor.normalize_audio  wakeword_training_script.py:73-76
//   SHIFT    circular shift (np.roll) augment_audio                   wakeword_training_script.py:106-108
//   SPEED    polyphase windowed-sinc resample + crop/zero-pad         :114-117 (speed stage) + :78-83
//   NOISE    snr_mixer(clean, bank segment, snr)                      stock/ms_snsd/MS-SNSD/audiolib.py:55-71
//   GAIN     linear gain
//   NORM_OUT peak normalise
// Input clips are fp32 or int16 PCM (x = s / 32768, the value librosa.load returns for a 16-bit WAV).
//
// Persistent kernel, one CTA (1024 threads) per SM, software-pipelined over clips with cp.async:
//   * while clip b is processed, the raw samples of clip b + grid stream into the other half of a double-buffered
//     shared-memory stage, and its scalar parameters / resample-table lookup are fetched by a few threads;
//   * the polyphase table and the noise segment of clip b are requested at the top of the iteration and waited for
//     only where they are consumed (after the normalise pass / after the gather), so no global latency is exposed.
// Each thread owns samples tid + 1024 e and keeps them in REGISTERS through every element-wise stage; only the
// gather goes through shared memory: shift and speed change are a single gather (the resampler reads its taps through
// the roll index map) whose taps come as 128-bit rows of a compact, zero-padded per-phase table (row pitch = 4 mod 8
// words, so the 128-bit row loads of consecutive phases are bank-conflict free; the sample loads of a warp fall into
// one or two 128-byte lines and are broadcast).  HBM traffic is one read of the clip and one write (the noise bank stays in L2): 128 KB per clip
// (96 KB from int16 PCM).
// Index arithmetic (roll source index, (q,p) phase decomposition, tap range, crop offset, output length) is
// integer-exact against oracle/augment.py; only exactly-zero taps of the polyphase table are skipped or added.
#include "ctx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxPerThread = 16;     // n_samples <= kThreads * kMaxPerThread
constexpr int kTblWords = 2560;       // shared-memory room for one ratio's compact polyphase table
constexpr int kPad = 32;              // zero floats in front of and behind every float sample buffer (fixed-phase gather)

struct AugKParams {
  const void* clips;                  // fp32 or int16
  const float* bank;
  int bank_rows;
  int64_t bank_len;
  ww_aug a;
  float* out;
  int B, N;
  const RsDesc* rs_desc;
  int n_rs;
  const float* rs_kern;
  const double* bank_prefix;          // [bank_rows][bank_len + 1] running sums of squares of the bank, or null
};

// per-clip scalars, fetched one iteration ahead into shared memory
struct ClipPrm {
  uint32_t flags;
  int shift, crop, noise_idx, noise_off;
  float snr, gain;
  float noise_energy;                 // sum of squares of the clip's noise segment (pipelined kernel, from bank_prefix)
  int rs;                             // index into rs_desc, -1 = ratio not prepared
  RsDesc d;
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
// two independent block reductions at once (sum or max); ends with every thread holding both results.
// Every call site has its OWN 64-float scratch row, rewritten one clip (several barriers) later, so one barrier suffices.
template <bool MAX>
__device__ __forceinline__ void block_reduce2(float& a, float& b, float* red, int tid) {
  a = MAX ? warp_max(a) : warp_sum(a);
  b = MAX ? warp_max(b) : warp_sum(b);
  if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = b; }
  __syncthreads();
  const float ra = red[tid & 31], rb = red[32 + (tid & 31)];       // kThreads / 32 == 32 partials each
  a = MAX ? warp_max(ra) : warp_sum(ra);
  b = MAX ? warp_max(rb) : warp_sum(rb);
}

// block maximum of non-negative floats (they order like their bit patterns): one redux.sync per warp, twice
__device__ __forceinline__ float block_max_nonneg(float v, float* red, int tid) {
  const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
  if ((tid & 31) == 0) red[tid >> 5] = __uint_as_float(w);
  __syncthreads();
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(red[tid & 31])));   // kThreads / 32 == 32 partials
}

// x / m for a whole clip with ONE correctly rounded reciprocal: q = x r, e = x - m q (exact, FMA), result = q + e r.
// With r = RN(1 / m) this is the correctly rounded quotient (Markstein's theorem) unless the significand of m is all
// ones or the quotient is subnormal; those cases take the IEEE divide.  3 FMA-pipe instructions instead of ~12.
struct ClipDiv {
  float m, r;
  bool fast;
  __device__ __forceinline__ float operator()(float x) const {
    const float q = x * r;
    if (!fast || (q != 0.0f && fabsf(q) < 1e-30f)) return __fdiv_rn(x, m);
    return fmaf(fmaf(-m, q, x), r, q);
  }
  // x * RN(1 / m): within 1 ulp of the quotient.  Used where the samples pass through an inexact stage (resampling, noise
  // mix) before or after the division anyway, so that nothing downstream is bit-comparable with numpy in the first place.
  __device__ __forceinline__ float approx(float x) const { return x * r; }
};
__device__ __forceinline__ ClipDiv make_clip_div(float m) {
  ClipDiv d;
  d.m = m;
  d.r = __fdiv_rn(1.0f, m);                            // correctly rounded reciprocal
  d.fast = (__float_as_uint(m) & 0x7fffffu) != 0x7fffffu && m > 1e-30f && m < 1e30f;
  return d;
}

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// `bytes` (multiple of 4) from global to shared, 16-byte chunks when both sides allow it; all threads take part
__device__ __forceinline__ void stage_bytes(void* dst, const void* src, int bytes, int tid) {
  if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)bytes) & 15) == 0) {
    for (int i = tid * 16; i < bytes; i += kThreads * 16)
      cp_async16(static_cast<char*>(dst) + i, static_cast<const char*>(src) + i);
  } else {
    for (int i = tid * 4; i < bytes; i += kThreads * 4)
      cp_async4(static_cast<char*>(dst) + i, static_cast<const char*>(src) + i);
  }
}

__device__ __forceinline__ float cvt_in(float v) { return v; }
__device__ __forceinline__ float cvt_in(int16_t v) { return (float)v * (1.0f / 32768.0f); }

// Sample index of register slot e (0..15) of thread tid.
__device__ __forceinline__ int slot_index(int tid, int e) { return tid + e * kThreads; }

// Scalars of clip b -> shared memory; called by threads 0..63 (warp 0: the plain fields; warp 1: parallel search of
// the prepared resample ratios).
__device__ __forceinline__ void fetch_prm(const AugKParams& p, ClipPrm& q, int b, int tid) {
  const uint32_t flags = p.a.flags[b];
  if (tid == 0) {
    q.flags = flags; q.shift = p.a.shift[b]; q.crop = p.a.crop_off[b]; q.noise_idx = p.a.noise_idx[b];
    q.noise_off = p.a.noise_off[b]; q.snr = p.a.snr_db[b]; q.gain = p.a.gain[b];
  }
  if (tid == 32) q.rs = -1;
  asm volatile("bar.sync 3, 64;" ::: "memory");         // the default lands before a match overwrites it
  if ((flags & WW_AUG_SPEED) && tid >= 32) {
    const int orig = p.a.rs_orig[b], neu = p.a.rs_new[b];
    for (int i = tid - 32; i < p.n_rs; i += 32)
      if (p.rs_desc[i].orig == orig && p.rs_desc[i].neu == neu) { q.rs = i; q.d = p.rs_desc[i]; }
  }
}

// NZ4 groups of four taps, fully unrolled: acc0 takes the even taps, acc1 the odd ones, in tap order
template <int NZ4>
__device__ __forceinline__ void taps4(const float4* __restrict__ kr4, const float* __restrict__ sp, float& acc0, float& acc1) {
#pragma unroll
  for (int k4 = 0; k4 < NZ4; ++k4) {
    const float4 w = kr4[k4];
    acc0 = fmaf(w.x, sp[4 * k4], acc0);
    acc1 = fmaf(w.y, sp[4 * k4 + 1], acc1);
    acc0 = fmaf(w.z, sp[4 * k4 + 2], acc0);
    acc1 = fmaf(w.w, sp[4 * k4 + 3], acc1);
  }
}

// Fixed-phase gather.  Thread t < S (S = n * floor(kThreads / n)) owns the outputs i = t + m S: their phase
// ph = (t + crop) mod n is the same for every m, so the thread reads its tap row ONCE into registers and every output costs
// NZ4 * 4 sample loads and FMAs, nothing else (the consecutive outputs of a warp read consecutive-ish samples: one or two
// wavefronts per load).  `xr` is the ROLLED clip with kPad zeros on either side, so the edges of the clip and the wrap point
// of the roll need no special case; summation order = taps4 (even taps -> acc0, odd taps -> acc1).
template <int NZ4>
__device__ __forceinline__ void gather_fixed_phase(const float* __restrict__ xr, const float* __restrict__ tbl, int pitch,
                                                   const RsDesc& d, int crop, int out_len, int N, int S, int tid,
                                                   float (&o)[kMaxPerThread]) {
  const int* lo_t = reinterpret_cast<const int*>(tbl + d.n * pitch);
  const int j0 = tid + crop;
  const int qq0 = j0 / d.n, ph = j0 - qq0 * d.n;
  float w[4 * NZ4];
  const float4* kr4 = reinterpret_cast<const float4*>(tbl + ph * pitch);
#pragma unroll
  for (int k4 = 0; k4 < NZ4; ++k4) {
    const float4 v = kr4[k4];
    w[4 * k4] = v.x; w[4 * k4 + 1] = v.y; w[4 * k4 + 2] = v.z; w[4 * k4 + 3] = v.w;
  }
  const float* sp = xr + (qq0 * d.o - d.width + lo_t[ph]);      // >= xr - width: inside the leading pad
  const int xstep = (S / d.n) * d.o;
#pragma unroll
  for (int e = 0; e < kMaxPerThread; ++e) {
    const int i = tid + e * S;
    float acc0 = 0.0f, acc1 = 0.0f;
    // outputs past the clip or past the resampled length read the leading zero pad instead (4 NZ4 <= 24 < kPad): no branch
    const float* sq = (i < N && i + crop < out_len) ? sp : xr - kPad;
#pragma unroll
    for (int k4 = 0; k4 < NZ4; ++k4) {
      acc0 = fmaf(w[4 * k4], sq[4 * k4], acc0);
      acc1 = fmaf(w[4 * k4 + 1], sq[4 * k4 + 1], acc1);
      acc0 = fmaf(w[4 * k4 + 2], sq[4 * k4 + 2], acc0);
      acc1 = fmaf(w[4 * k4 + 3], sq[4 * k4 + 3], acc1);
    }
    o[e] = acc0 + acc1;
    sp += xstep;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1) augment_kernel(const AugKParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red4[4][64];                      // one scratch row per reduction site
  __shared__ ClipPrm prm[2];
  const int tid = threadIdx.x;
  const int N = p.N;
  const int NP = (N + 7) & ~7;                      // every buffer is padded to whole 16-byte groups
  constexpr bool kInPlace = sizeof(TIn) == 4;        // fp32 input: the stage buffer doubles as the gather source
  // float sample buffers carry kPad zero floats on either side (written once, below): the fixed-phase gather reads past
  // the ends of the clip instead of clipping its tap window
  constexpr int kPadIn = kInPlace ? kPad : 0;        // pad of a stage buffer in elements of TIn
  const int SP = NP + 2 * kPadIn;
  TIn* stage0 = reinterpret_cast<TIn*>(smem_raw) + kPadIn;
  TIn* stage1 = stage0 + SP;
  float* curf = kInPlace ? nullptr : reinterpret_cast<float*>(stage1 + SP) + kPad;
  float* noise_s = kInPlace ? reinterpret_cast<float*>(stage1 + SP - kPadIn) : curf + NP + kPad;
  float* tbl = noise_s + NP;
  if (kInPlace) {
    for (int k = tid; k < 2 * (2 * kPad + NP - N); k += kThreads) {
      float* base = reinterpret_cast<float*>((k & 1) ? stage1 : stage0);
      const int r = k >> 1;
      base[r < kPad ? r - kPad : N + (r - kPad)] = 0.0f;
    }
  } else {
    for (int r = tid; r < 2 * kPad + NP - N; r += kThreads) curf[r < kPad ? r - kPad : N + (r - kPad)] = 0.0f;
  }

  int it = 0;
  // prologue: parameters and raw samples of the first clip
  if ((int)blockIdx.x < p.B) {
    if (tid < 64) fetch_prm(p, prm[0], blockIdx.x, tid);
    stage_bytes(stage0, static_cast<const TIn*>(p.clips) + (int64_t)blockIdx.x * N, N * (int)sizeof(TIn), tid);
  }
  cp_commit();

  for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
    cp_wait<0>();
    __syncthreads();                                  // stage[it&1] and prm[it&1] are complete and visible
    const ClipPrm& q = prm[it & 1];
    const uint32_t flags = q.flags;
    TIn* st = (it & 1) ? stage1 : stage0;
    float* cur = kInPlace ? reinterpret_cast<float*>(st) : curf;
    const bool do_speed = (flags & WW_AUG_SPEED) != 0;
    const bool rs_ok = do_speed && q.rs >= 0;
    const RsDesc d = q.d;
    const int pitch = d.nz + ((d.nz & 4) ? 0 : 4);    // table row pitch in words: 4 mod 8 (see ww_prepare_resample)
    const int tbl_words = rs_ok ? d.n * pitch + 2 * d.n : 0;
    const bool in_smem = tbl_words <= kTblWords;
    // ---- async requests: G1 = polyphase table of this clip, G2 = its noise segment, G3 = next clip's samples
    if (rs_ok && in_smem) stage_bytes(tbl, p.rs_kern + d.offset, ((tbl_words + 3) & ~3) * 4, tid);
    cp_commit();
    if (flags & WW_AUG_NOISE)
      stage_bytes(noise_s, p.bank + (int64_t)q.noise_idx * p.bank_len + q.noise_off, N * 4, tid);
    cp_commit();
    const int bn = b + gridDim.x;
    if (bn < p.B) stage_bytes((it & 1) ? stage0 : stage1, static_cast<const TIn*>(p.clips) + (int64_t)bn * N, N * (int)sizeof(TIn), tid);
    cp_commit();
    // ---- scalars of the next clip (64 threads; consumed after the __syncthreads at the top of the next iteration)
    if (tid < 64 && bn < p.B) fetch_prm(p, prm[(it + 1) & 1], bn, tid);

    // ---- load own samples from the stage, peak normalise
    float o[kMaxPerThread];
    float m = 0.0f;
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) {
      const int i = slot_index(tid, e);
      o[e] = (i < N) ? cvt_in(st[i]) : 0.0f;
      m = fmaxf(m, fabsf(o[e]));
    }
    // normalize_audio is the IEEE quotient wherever its result is bit-comparable with numpy (no resampling / noise mix on
    // the clip); clips that go through those stages take one multiply per sample instead (1 ulp)
    const bool inexact = (flags & (WW_AUG_SPEED | WW_AUG_NOISE)) != 0;
    if (flags & WW_AUG_NORM_IN) {
      m = block_max_nonneg(m, red4[0], tid);
      if (m > 0.0f) {
        const ClipDiv dv = make_clip_div(m);
        if (inexact && dv.fast) {
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = dv.approx(o[e]);
        } else {
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = dv(o[e]);
        }
      }
    }

    // fixed-phase gather (see gather_fixed_phase): thread t < S owns outputs t + m S from here on
    const int S = rs_ok ? d.n * (kThreads / max(d.n, 1)) : kThreads;
    const bool fixed = rs_ok && in_smem && d.n <= kThreads && d.nz >= 16 && d.nz <= 24 && d.width <= kPad - 8 &&
                       (N + S - 1) / S <= kMaxPerThread;
    const int ss = fixed ? S : kThreads;               // owner stride of the register tile after the gather
    if (flags & (WW_AUG_SHIFT | WW_AUG_SPEED)) {
      // ---- one gather through shared memory: out[i] = sum_k kern[ph][k] * rolled[x0 + k],  rolled[t] = in[(t - shift) mod N]
      int s = (flags & WW_AUG_SHIFT) ? q.shift % N : 0;
      if (s < 0) s += N;
      if (fixed) {
        if (kInPlace && !(flags & WW_AUG_NORM_IN)) __syncthreads();      // every thread has read its samples: roll in place
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) {
          const int i = slot_index(tid, e);
          int dpos = i + s;
          if (dpos >= N) dpos -= N;
          if (i < N) cur[dpos] = o[e];
        }
      } else if (!kInPlace || (flags & WW_AUG_NORM_IN)) {
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) {
          const int i = slot_index(tid, e);
          if (i < N) cur[i] = o[e];
        }
      }
      cp_wait<2>();                                    // G1 (table) landed; G2 / G3 may still be in flight
      __syncthreads();
      if (do_speed) {
        if (!rs_ok) {
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = __int_as_float(0x7fc00000);   // loud: NaN clip
        } else if (fixed) {
          const int out_len = (d.n * N + d.o - 1) / d.o;                             // ceil(n*N/o), < 2^31
          const int crop = (out_len > N) ? q.crop : 0;
          if (tid < S) {
            switch (d.nz >> 2) {
              case 4: gather_fixed_phase<4>(cur, tbl, pitch, d, crop, out_len, N, S, tid, o); break;
              case 5: gather_fixed_phase<5>(cur, tbl, pitch, d, crop, out_len, N, S, tid, o); break;
              default: gather_fixed_phase<6>(cur, tbl, pitch, d, crop, out_len, N, S, tid, o); break;
            }
          } else {
#pragma unroll
            for (int e = 0; e < kMaxPerThread; ++e) o[e] = 0.0f;
          }
        } else {
          const float* kern = in_smem ? tbl : p.rs_kern + d.offset;
          const int* lo_t = reinterpret_cast<const int*>(kern + d.n * pitch);
          const int* cnt_t = lo_t + d.n;
          const int out_len = (d.n * N + d.o - 1) / d.o;                             // ceil(n*N/o), < 2^31
          const int crop = (out_len > N) ? q.crop : 0;
          const uint32_t magic = 0xffffffffu / (uint32_t)d.n + 1u;                   // j / n == umulhi(j, magic) for j, n < 2^16
          const int nz4 = d.nz >> 2;                                                 // table rows are zero-padded to x4 taps
#pragma unroll 1
          for (int e = 0; e < kMaxPerThread; ++e) {                                   // rare fallback: kept small
            const int i = slot_index(tid, e);
            float acc0 = 0.0f, acc1 = 0.0f;
            const int j = i + crop;                                                  // resampled-domain index
            if (i < N && j < out_len) {
              const int qq = (int)__umulhi((uint32_t)j, magic), ph = j - qq * d.n;   // (q, p) phase decomposition
              const int x0 = qq * d.o - d.width + lo_t[ph];                          // source index of the first non-zero tap
              const float* kr = kern + ph * pitch;
              int src = x0 - s;
              if (src < 0) src += N;
              if ((unsigned)x0 <= (unsigned)(N - d.nz) && (unsigned)src <= (unsigned)(N - d.nz)) {
                // interior: no edge clipping, no wrap inside the (padded) tap window
                const float4* kr4 = reinterpret_cast<const float4*>(kr);
                const float* sp = cur + src;
                for (int k4 = 0; k4 < nz4; ++k4) {
                  const float4 w = kr4[k4];
                  acc0 = fmaf(w.x, sp[4 * k4], acc0);
                  acc1 = fmaf(w.y, sp[4 * k4 + 1], acc1);
                  acc0 = fmaf(w.z, sp[4 * k4 + 2], acc0);
                  acc1 = fmaf(w.w, sp[4 * k4 + 3], acc1);
                }
              } else {
                const int k0 = x0 < 0 ? -x0 : 0;
                const int k1 = min(cnt_t[ph], N - x0);
                src = x0 + k0 - s;
                if (src < 0) src += N;
                if (src < 0) src += N;
                for (int k = k0; k < k1; ++k) {
                  acc0 = fmaf(kr[k], cur[src], acc0);
                  src = (src + 1 >= N) ? src + 1 - N : src + 1;
                }
              }
            }
            const float v = acc0 + acc1;
#pragma unroll
            for (int k = 0; k < kMaxPerThread; ++k)                                  // static indices keep o[] in registers
              if (k == e) o[k] = v;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) {
          const int i = slot_index(tid, e);
          int src = i - s;
          if (src < 0) src += N;
          o[e] = (i < N) ? cur[src] : 0.0f;                                         // bit-exact copy (np.roll)
        }
      }
    }

    if (flags & WW_AUG_NOISE) {
      cp_wait<1>();                                    // G2 (noise segment) landed
      __syncthreads();
      const float snr = q.snr;
      const float target = 0.0562341325190349f;            // 10 ** (-25 / 20)
      float n[kMaxPerThread];
      float sc = 0.0f, sn = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) {
        const int i = tid + e * ss;
        n[e] = (tid < ss && i < N) ? noise_s[i] : 0.0f;
        sc = fmaf(o[e], o[e], sc);
        sn = fmaf(n[e], n[e], sn);
      }
      block_reduce2<false>(sc, sn, red4[1], tid);
      // every warp runs these scalar chains in lockstep with nothing to overlap them: MUFU approximations (2 ulp, against
      // the stage's 5e-5 tolerance) keep them a handful of instructions long
      const float inv_n = 1.0f / (float)N;
      const float scalarclean = target * rsqrtf(sc * inv_n), scalarnoise = target * rsqrtf(sn * inv_n);
      // the reference re-measures both RMS values after scaling (audiolib.py:60,65)
      float sc2 = 0.0f, sn2 = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) {
        o[e] *= scalarclean;
        n[e] *= scalarnoise;
        sc2 = fmaf(o[e], o[e], sc2);
        sn2 = fmaf(n[e], n[e], sn2);
      }
      block_reduce2<false>(sc2, sn2, red4[2], tid);
      // noisescalar = sqrt(rmsclean / 10^(snr/20) / rmsnoise), audiolib.py:68 (sqrt quirk kept)
      //             = (sc2 / sn2)^(1/4) * 2^(-snr log2(10) / 40)
      const float noisescalar = rsqrtf(rsqrtf(__fdividef(sc2, sn2))) * exp2f(snr * -0.0830482023721841f);
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) o[e] = o[e] + n[e] * noisescalar;
    }
    if (flags & WW_AUG_GAIN) {
      const float g = q.gain;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) o[e] *= g;
    }
    if (flags & WW_AUG_NORM_OUT) {
      float mo = 0.0f;
#pragma unroll
      for (int e = 0; e < kMaxPerThread; ++e) mo = fmaxf(mo, fabsf(o[e]));
      mo = block_max_nonneg(mo, red4[3], tid);
      if (mo > 0.0f) {
        const ClipDiv dv = make_clip_div(mo);
        if (inexact && dv.fast) {
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = dv.approx(o[e]);
        } else {
#pragma unroll
          for (int e = 0; e < kMaxPerThread; ++e) o[e] = dv(o[e]);
        }
      }
    }
    float* __restrict__ dst = p.out + (int64_t)b * N;
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) {
      const int i = tid + e * ss;
      if (tid < ss && i < N) dst[i] = o[e];
    }
    // the __syncthreads at the top of the next iteration orders this clip's shared-memory reads (cur, tbl, noise)
    // before the next clip's cp.async writes and in-place normalise
  }
  cp_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------------
// Pipelined form (default).  The lock-step kernel above keeps every warp in the same phase between block barriers: the
// gather saturates the shared-memory port while the issue slots idle, the element-wise passes do the opposite (ncu: LSU
// data pipe 54 %, issue 47 %).  Here the CTA is two ROLES that work on DIFFERENT clips:
//   C (conditioning, warps 0-15): in iteration k it takes the gathered clip k out of `res` into registers (32 samples per
//     thread), releases `res` at once and finishes the clip: peak normalise (the gather is linear, so the peak of the raw
//     clip is applied behind it: one rounding of difference, none for clips that are only rolled), noise mix, gain, peak
//     normalise, store.  Threaded through that pass in four batches of 8 samples per thread, it loads clip k + 2 from HBM
//     and stores it ROLLED into src[k & 1] -- which the gather of clip k has just released -- with the zero pads of the
//     fixed-phase gather around it, and copies its polyphase table;
//   G (gather, warps 16-23): res[i] = sum_k taps[ph][k] * src[k & 1][x0(i) + k] for clip k, nothing else: two adjacent
//     outputs per thread off one window fetched with 64-bit loads, both tap rows in registers (gather_role_pairs).
// so the port-bound gather of one clip runs under the latency-bound passes of its neighbours.  The roles meet at named
// barriers in producer / consumer form (bar.arrive by the producer, bar.sync by the consumer, count = both roles):
// src_full[2], res_full, res_empty (res_full(k) also says that src[k & 1] is free again).  The per-clip scalars live in a
// 4-entry ring that warp 0 fills three clips ahead.  The peak of the raw clip, the energy of the gathered clip and the
// energy of the noise segment share ONE reduction; the RMS values "re-measured after scaling" (audiolib.py:60,65) are
// scale^2 x the sums already known (mathematically equal; 1e-7 relative); the block sums are taken over 16 warps instead
// of 32 (1e-7 relative on the noise scalars).  Per element the arithmetic is the lock-step kernel's up to summation order.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRole = 512;            // threads of the conditioning role
constexpr int kRoleG = 256;           // threads of the gather role (port-bound: a few warps saturate it, and a short queue in front of
                                      // the shared-memory port keeps the latency of the other role's accesses down)
constexpr int kPipeThreads = kRole + kRoleG;
constexpr int kPerC = 32;             // samples per conditioning thread: n_samples <= kRole * kPerC
enum { BAR_C = 1, BAR_SRC_FULL = 2, BAR_RES_FULL = 4, BAR_RES_EMPTY = 5 };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// reductions over the conditioning role (16 warps); every call site has its own scratch row (see block_reduce2)
__device__ __forceinline__ float role_max_nonneg(float v, float* red, int t) {
  const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
  if ((t & 31) == 0) red[t >> 5] = __uint_as_float(w);
  bar_sync(BAR_C, kRole);
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(red[t & 15])));
}
__device__ __forceinline__ void role_sum2(float& a, float& b, float* red, int t) {
  a = warp_sum(a);
  b = warp_sum(b);
  if ((t & 31) == 0) { red[t >> 5] = a; red[16 + (t >> 5)] = b; }
  bar_sync(BAR_C, kRole);
  float ra = red[t & 15], rb = red[16 + (t & 15)];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { ra += __shfl_xor_sync(0xffffffffu, ra, o); rb += __shfl_xor_sync(0xffffffffu, rb, o); }
  a = ra; b = rb;
}

// max of a non-negative value and two sums in one barrier
__device__ __forceinline__ void role_reduce3(float& m, float& a, float& b, float* red, int t) {
  const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  a = warp_sum(a);
  b = warp_sum(b);
  if ((t & 31) == 0) { red[t >> 5] = a; red[16 + (t >> 5)] = b; red[32 + (t >> 5)] = __uint_as_float(w); }
  bar_sync(BAR_C, kRole);
  float ra = red[t & 15], rb = red[16 + (t & 15)];
  m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(red[32 + (t & 15)])));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { ra += __shfl_xor_sync(0xffffffffu, ra, o); rb += __shfl_xor_sync(0xffffffffu, rb, o); }
  a = ra; b = rb;
}

// max of a non-negative value and one sum in one barrier
__device__ __forceinline__ void role_reduce_max_sum(float& m, float& a, float* red, int t) {
  const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  a = warp_sum(a);
  if ((t & 31) == 0) { red[t >> 5] = a; red[32 + (t >> 5)] = __uint_as_float(w); }
  bar_sync(BAR_C, kRole);
  float ra = red[t & 15];
  m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(red[32 + (t & 15)])));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ra += __shfl_xor_sync(0xffffffffu, ra, o);
  a = ra;
}

// Gather role, two adjacent outputs per thread.  Outputs j and j + 1 read windows that start 0-2 samples apart, so ONE window
// of NW = nz + 4 samples serves both: it is fetched as 64-bit loads from an even start (the parity of the first tap's
// position and the offset of the second output's window are constants of the thread -- its two phases are fixed and the
// window advances by an even step -- and are folded into two tap rows shifted into place once per clip), which turns the
// gather's nz loads per output into (nz + 4) / 4: the LSU pipe, which paces this kernel, sees 3.4x fewer instructions.
// Thread u < S / 2 owns the outputs 2 u + e + m S (S = n c, c even).  Returns false (nothing written) when a tap row does not
// fit its shifted window; the caller then runs the one-output form for this thread's outputs.
template <int NW>
__device__ __forceinline__ bool gather_role_pairs(const float* __restrict__ xr, const float* __restrict__ tbl, int pitch,
                                                  const RsDesc& d, int crop, int out_len, int N, int S, int u,
                                                  float* __restrict__ res) {
  const int* lo_t = reinterpret_cast<const int*>(tbl + d.n * pitch);
  const int* cnt_t = lo_t + d.n;
  const int j0 = 2 * u + crop, j1 = j0 + 1;
  const int q0 = j0 / d.n, p0 = j0 - q0 * d.n;
  const int q1 = j1 / d.n, p1 = j1 - q1 * d.n;
  const int x0 = q0 * d.o - d.width + lo_t[p0], x1 = q1 * d.o - d.width + lo_t[p1];
  const int wb = min(x0, x1) & ~1;                      // even window start (two's complement: also for negative positions)
  const int s0 = x0 - wb, s1 = x1 - wb, c0 = cnt_t[p0], c1 = cnt_t[p1];
  if (s0 + c0 > NW || s1 + c1 > NW) return false;
  float w0[NW], w1[NW];
  const float* r0 = tbl + p0 * pitch - s0;
  const float* r1 = tbl + p1 * pitch - s1;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    w0[k] = (k >= s0 && k < s0 + c0) ? r0[k] : 0.0f;
    w1[k] = (k >= s1 && k < s1 + c1) ? r1[k] : 0.0f;
  }
  const float* wp = xr + wb;
  const int xstep = (S / d.n) * d.o;                     // even
  const int i_end = min(N, out_len - crop);              // first output that has no source
#pragma unroll 1
  for (int i0 = 2 * u; i0 < N; i0 += S, wp += xstep) {
    // outputs past the resampled length read the leading zero pad instead: no branch
    const float2* w2 = reinterpret_cast<const float2*>(i0 < i_end ? wp : xr - kPad);
    float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll
    for (int k2 = 0; k2 < NW / 2; ++k2) {
      const float2 x = w2[k2];
      a0 = fmaf(w0[2 * k2], x.x, a0);
      a1 = fmaf(w0[2 * k2 + 1], x.y, a1);
      b0 = fmaf(w1[2 * k2], x.x, b0);
      b1 = fmaf(w1[2 * k2 + 1], x.y, b1);
    }
    const float v0 = a0 + a1, v1 = (i0 + 1 < i_end) ? b0 + b1 : 0.0f;
    if (i0 + 1 < N) *reinterpret_cast<float2*>(res + i0) = make_float2(v0, v1);
    else res[i0] = v0;
  }
  return true;
}

#ifdef WW_AUG_TRACE
__device__ unsigned long long aug_trace[16];
#define AUG_T(var) const long long var = clock64()
#define AUG_ACC(slot, a, b) do { if (blockIdx.x == 0 && (tid == 0 || tid == kRole)) aug_trace[slot] += (unsigned long long)((b) - (a)); } while (0)
#else
#define AUG_T(var)
#define AUG_ACC(slot, a, b)
#endif

// scalars of clip b -> one slot of the per-clip ring, by one warp (lane 0 the plain fields, the warp the search of the
// prepared resample ratios): a two-level chain of global loads
__device__ __forceinline__ void fetch_clip_scalars(const AugKParams& p, ClipPrm& q, int b, int lane) {
  const uint32_t flags = __ldg(p.a.flags + b);
  if (lane == 0) {
    q.flags = flags; q.shift = p.a.shift[b]; q.crop = p.a.crop_off[b]; q.noise_idx = p.a.noise_idx[b];
    q.noise_off = p.a.noise_off[b]; q.snr = p.a.snr_db[b]; q.gain = p.a.gain[b]; q.rs = -1;
    if (p.bank_prefix && (flags & WW_AUG_NOISE)) {
      const double* pp = p.bank_prefix + (int64_t)q.noise_idx * (p.bank_len + 1) + q.noise_off;
      q.noise_energy = (float)(pp[p.N] - pp[0]);
    }
  }
  __syncwarp();
  if (flags & WW_AUG_SPEED) {
    const int orig = p.a.rs_orig[b], neu = p.a.rs_new[b];
    for (int i = lane; i < p.n_rs; i += 32)
      if (p.rs_desc[i].orig == orig && p.rs_desc[i].neu == neu) { q.rs = i; q.d = p.rs_desc[i]; }
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kPipeThreads, 1) augment_pipe_kernel(const AugKParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red3[3][48];
  __shared__ ClipPrm prm[4];                          // ring: clip j's scalars live from one clip before its conditioning to its store
  const int tid = threadIdx.x;
  const int N = p.N;
  const int NP = (N + 7) & ~7;
  const int SP = NP + 2 * kPad;
  float* src0 = reinterpret_cast<float*>(smem_raw) + kPad;      // src[s] = src0 + s SP, kPad zeros on either side
  float* res = src0 - kPad + 2 * SP;                            // [NP]
  float* tbl0 = res + NP;                                       // [2][kTblWords]
  for (int k = tid; k < 2 * (2 * kPad + NP - N); k += kPipeThreads) {
    float* base = src0 + (k & 1) * SP;
    const int r = k >> 1;
    base[r < kPad ? r - kPad : N + (r - kPad)] = 0.0f;
  }
  __syncthreads();
  const int K = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // clips of this CTA

  if (tid >= kRole) {
    // ================================================= gather role
    const int t = tid - kRole;
    for (int k = 0; k < K; ++k) {
      const int s = k & 1;
      AUG_T(g0);
      bar_sync(BAR_SRC_FULL + s, kPipeThreads);
      AUG_T(g1);
      if (k >= 1) bar_sync(BAR_RES_EMPTY, kPipeThreads);
      AUG_T(g2);
      // The scalars of clip k + 2 (a two-level chain of global loads) are fetched HERE, by the first warp of the role that has
      // slack: the conditioning role has finished clip k - 2, whose ring slot this is, and reads the new entry after
      // res_full(k), which this warp arrives at after the gather below.
      if (t < 32 && k + 2 < K) fetch_clip_scalars(p, prm[(k + 2) & 3], (int)blockIdx.x + (k + 2) * (int)gridDim.x, t);
      AUG_ACC(0, g0, g1); AUG_ACC(1, g1, g2);
      const ClipPrm& q = prm[k & 3];
      const uint32_t flags = q.flags;
      const float* cur = src0 + s * SP;
      const float* tbl = tbl0 + s * kTblWords;
      if (!(flags & WW_AUG_SPEED)) {
        for (int i = t; i < N; i += kRoleG) res[i] = cur[i];      // the roll happened on the way in: bit-exact copy
      } else if (q.rs < 0) {
        for (int i = t; i < N; i += kRoleG) res[i] = __int_as_float(0x7fc00000);     // loud: NaN clip
      } else {
        const RsDesc d = q.d;
        const int pitch = d.nz + ((d.nz & 4) ? 0 : 4);
        const int tbl_words = d.n * pitch + 2 * d.n;
        const bool in_smem = tbl_words <= kTblWords;
        const int out_len = (d.n * N + d.o - 1) / d.o;                               // ceil(n*N/o), < 2^31
        const int crop = (out_len > N) ? q.crop : 0;
        // Two adjacent outputs per thread (gather_role_pairs).  A half-warp's 64-bit loads are conflict-free while its
        // windows span <= 32 words: lanes are 2 o / n words apart, so when the clip is being shortened (o > n) only the first
        // Lh lanes of every half-warp work (for n = 100 the idle lanes are the ones S = n c leaves idle anyway).
        const int Lh = d.o > d.n ? min(16, (15 * d.n) / d.o + 1) : 16;
        const int lanes = (kRoleG / 16) * Lh;                                     // threads that may own a pair
        const int cmul = ((2 * lanes) / d.n) & ~1;                                // periods per sweep (even: the window step stays even)
        const bool fixed = in_smem && cmul >= 2 && d.nz >= 16 && d.nz <= 24 && d.width <= kPad - 8;
        if (fixed) {
          const int S = d.n * cmul;
          const int u = (t >> 4) * Lh + (t & 15);
          if ((t & 15) < Lh && 2 * u < S) {
            bool done;
            switch (d.nz >> 2) {
              case 4: done = gather_role_pairs<20>(cur, tbl, pitch, d, crop, out_len, N, S, u, res); break;
              case 5: done = gather_role_pairs<24>(cur, tbl, pitch, d, crop, out_len, N, S, u, res); break;
              default: done = gather_role_pairs<28>(cur, tbl, pitch, d, crop, out_len, N, S, u, res); break;
            }
            if (!done) {
              // a tap row that does not fit its shifted window (irregular first-tap table): this thread's outputs one by one
              const int* lo_t = reinterpret_cast<const int*>(tbl + d.n * pitch);
              const int* cnt_t = lo_t + d.n;
              for (int i = 2 * u; i < N; i += S)
                for (int e = 0; e < 2 && i + e < N; ++e) {
                  const int j = i + e + crop;
                  float acc = 0.0f;
                  if (j < out_len) {
                    const int qq = j / d.n, ph = j - qq * d.n;
                    const int x0 = qq * d.o - d.width + lo_t[ph];
                    const float* kr = tbl + ph * pitch;
                    const int k0 = x0 < 0 ? -x0 : 0, k1 = min(cnt_t[ph], N - x0);
                    for (int kk = k0; kk < k1; ++kk) acc = fmaf(kr[kk], cur[x0 + kk], acc);
                  }
                  res[i + e] = acc;
                }
            }
          }
        } else {
          // rare fallback (very long tap rows / tables that do not fit): generic phase decomposition, taps from global
          const float* kern = in_smem ? tbl : p.rs_kern + d.offset;
          const int* lo_t = reinterpret_cast<const int*>(kern + d.n * pitch);
          const int* cnt_t = lo_t + d.n;
#pragma unroll 1
          for (int i = t; i < N; i += kRoleG) {
            float acc = 0.0f;
            const int j = i + crop;
            if (j < out_len) {
              const int qq = j / d.n, ph = j - qq * d.n;
              const int x0 = qq * d.o - d.width + lo_t[ph];
              const float* kr = kern + ph * pitch;
              const int k0 = x0 < 0 ? -x0 : 0;
              const int k1 = min(cnt_t[ph], N - x0);
              for (int kk = k0; kk < k1; ++kk) acc = fmaf(kr[kk], cur[x0 + kk], acc);
            }
            res[i] = acc;
          }
        }
      }
      AUG_T(g3);
      AUG_ACC(2, g2, g3);
      bar_arrive(BAR_RES_FULL, kPipeThreads);
    }
    return;
  }

  // =================================================== conditioning role
  const int t = tid;
  const int nt = N - t;                                  // sample t + off exists  <=>  off < nt
  const int esz = (int)sizeof(TIn);
  auto clip_ptr = [&](int j) {
    return static_cast<const TIn*>(p.clips) + (int64_t)((int)blockIdx.x + j * (int)gridDim.x) * N;
  };
  auto prefetch_clip = [&](int j) {                      // HBM -> L2, one 128-byte line per thread
    const char* g = reinterpret_cast<const char*>(clip_ptr(j));
    for (int l = t * 128; l < N * esz; l += kRole * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + l));
  };
  // sum of squares of the noise segment of a clip (thread partial): two batches of 16 loads
  auto noise_sumsq = [&](const ClipPrm& q) {
    const float* __restrict__ nb = p.bank + (int64_t)q.noise_idx * p.bank_len + q.noise_off + t;
    float sn = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float n[kPerC / 2];
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e) {
        const int off = (h * (kPerC / 2) + e) * kRole;
        n[e] = (off < nt) ? __ldg(nb + off) : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e) sn = fmaf(n[e], n[e], sn);
    }
    return sn;
  };

  // ---- conditioning of clip j into src[j & 1], in batches of kPB samples per thread that are threaded through the
  // finishing pass of clip j - 2 (issue a batch's loads, do something else, fold them into the peak and store them rolled):
  // the two chains of global / shared-memory latencies overlap instead of adding up, and the gather role never waits for it
  constexpr int kPB = 8;
  float pb[kPB];
  float m_prep = 0.0f;                                   // thread-partial peak of the clip being conditioned
  const TIn* __restrict__ px = nullptr;                  // its samples (+ t)
  float* pa = nullptr; float* pbw = nullptr; int lim = 0;  // rolled destination: element t + off -> pa[off], or pbw[off] from off >= lim on
  auto prep_begin = [&](int j) {
    const ClipPrm& qn = prm[j & 3];
    px = clip_ptr(j) + t;
    int sh = (qn.flags & WW_AUG_SHIFT) ? qn.shift % N : 0;
    if (sh < 0) sh += N;
    pa = src0 + (j & 1) * SP + t + sh;                   // rolled[(i + shift) mod N] = in[i]
    pbw = pa - N;
    lim = nt - sh;
    m_prep = 0.0f;
    // the clip's polyphase table travels by cp.async (no registers, no exposed latency): waited for in prep_end
    if ((qn.flags & WW_AUG_SPEED) && qn.rs >= 0) {
      const int pitch = qn.d.nz + ((qn.d.nz & 4) ? 0 : 4);
      const int tbl_words = qn.d.n * pitch + 2 * qn.d.n;
      if (tbl_words <= kTblWords) {
        const float* g = p.rs_kern + qn.d.offset;                                        // offset and size are whole float4s
        float* dst = tbl0 + (j & 1) * kTblWords;
        for (int i = t * 4; i < ((tbl_words + 3) & ~3); i += kRole * 4) cp_async16(dst + i, g + i);
      }
    }
    cp_commit();
  };
  auto prep_issue = [&](int bi) {
#pragma unroll
    for (int e = 0; e < kPB; ++e) {
      const int off = (bi * kPB + e) * kRole;
      pb[e] = (off < nt) ? cvt_in(__ldg(px + off)) : 0.0f;
    }
  };
  auto prep_consume = [&](int bi) {
#pragma unroll
    for (int e = 0; e < kPB; ++e) {
      const int off = (bi * kPB + e) * kRole;
      m_prep = fmaxf(m_prep, fabsf(pb[e]));
      if (off < nt) { float* d = (off >= lim) ? pbw : pa; d[off] = pb[e]; }
    }
  };
  auto prep_end = [&](int j) {
    cp_wait<0>();
    bar_arrive(BAR_SRC_FULL + (j & 1), kPipeThreads);
    if (j + 1 < K) prefetch_clip(j + 1);
  };
  static_assert(kPerC == 4 * kPB, "four conditioning batches per clip");

  if (t < 32)
    for (int j = 0; j < 2 && j < K; ++j) fetch_clip_scalars(p, prm[j], (int)blockIdx.x + j * (int)gridDim.x, t);
  bar_sync(BAR_C, kRole);
  float m_in = 0.0f, m_mid = 0.0f;                       // thread-partial peaks of clip k (being finished) and clip k + 1
  for (int j = 0; j < 2 && j < K; ++j) {                 // the first two clips: nothing to hide behind yet
    prep_begin(j);
    prep_issue(0); prep_consume(0); prep_issue(1); prep_consume(1); prep_issue(2); prep_consume(2); prep_issue(3); prep_consume(3);
    prep_end(j);
    if (j == 0) m_in = m_prep; else m_mid = m_prep;
  }

  float o[kPerC];
  for (int k = 0; k < K; ++k) {
    // ---------------- iteration k: finish clip k out of `res`; meanwhile condition clip k + 2 into src[k & 1], which the
    // gather of clip k has just released (res_full(k) says so: no barrier of its own)
    const int j = k + 2;
    const bool do_prep = j < K;
    const ClipPrm& q = prm[k & 3];
    const uint32_t flags = q.flags;
    AUG_T(c0);
    // the noise energy of clip k does not depend on the gather: taken while the gather role is still busy with the clip
    const bool have_energy = p.bank_prefix != nullptr;
    float sn = ((flags & WW_AUG_NOISE) && !have_energy) ? noise_sumsq(q) : 0.0f;
    AUG_T(c1);
    bar_sync(BAR_RES_FULL, kPipeThreads);
    AUG_T(c2);
    AUG_ACC(4, c0, c1); AUG_ACC(5, c1, c2);
#pragma unroll
    for (int e = 0; e < kPerC; ++e) o[e] = (e * kRole < nt) ? res[t + e * kRole] : 0.0f;
    if (k + 1 < K) bar_arrive(BAR_RES_EMPTY, kPipeThreads);
    AUG_T(c2a); AUG_ACC(8, c2, c2a);
    if (do_prep) { prep_begin(j); prep_issue(0); }
    const bool inexact = (flags & (WW_AUG_SPEED | WW_AUG_NOISE)) != 0;
    const float* __restrict__ nb = p.bank + (int64_t)q.noise_idx * p.bank_len + q.noise_off + t;
    // first half of the noise segment for the mix: requested before the reduction, consumed after it
    float n[kPerC / 2];
    if (flags & WW_AUG_NOISE) {
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e) n[e] = (e * kRole < nt) ? __ldg(nb + e * kRole) : 0.0f;
    }
    float sc = 0.0f, mred = 0.0f;
    if (flags & (WW_AUG_NORM_IN | WW_AUG_NOISE)) {
      // ONE reduction: peak of the raw clip (taken while it was conditioned), energy of the gathered clip, energy of the noise
      if (flags & WW_AUG_NOISE) {
#pragma unroll
        for (int e = 0; e < kPerC; ++e) sc = fmaf(o[e], o[e], sc);
      }
      mred = m_in;
      if (have_energy) { role_reduce_max_sum(mred, sc, red3[1], t); sn = q.noise_energy; }
      else role_reduce3(mred, sc, sn, red3[1], t);
    }
    AUG_T(c2b); AUG_ACC(9, c2a, c2b);
    if (do_prep) { prep_consume(0); prep_issue(1); }
    if ((flags & WW_AUG_NORM_IN) && mred > 0.0f) {
      const ClipDiv dv = make_clip_div(mred);
      if (inexact && dv.fast) {
#pragma unroll
        for (int e = 0; e < kPerC; ++e) o[e] = dv.approx(o[e]);
        sc *= dv.r * dv.r;                               // energy of the normalised clip
      } else {
#pragma unroll
        for (int e = 0; e < kPerC; ++e) o[e] = dv(o[e]);
        if (flags & WW_AUG_NOISE) {                      // IEEE-divide path on a noisy clip (peak with an all-ones significand): re-measure
          sc = 0.0f;
#pragma unroll
          for (int e = 0; e < kPerC; ++e) sc = fmaf(o[e], o[e], sc);
          float z = 0.0f;
          role_sum2(sc, z, red3[0], t);
        }
      }
    }
    if (flags & WW_AUG_NOISE) {
      const float target = 0.0562341325190349f;            // 10 ** (-25 / 20)
      const float inv_n = 1.0f / (float)N;
      const float scalarclean = target * rsqrtf(sc * inv_n), scalarnoise = target * rsqrtf(sn * inv_n);
      // the reference re-measures both RMS values after scaling (audiolib.py:60,65): sum (a x)^2 = a^2 sum x^2
      const float sc2 = sc * scalarclean * scalarclean, sn2 = sn * scalarnoise * scalarnoise;
      // noisescalar = sqrt(rmsclean / 10^(snr/20) / rmsnoise), audiolib.py:68 (sqrt quirk kept)
      const float noisescalar = rsqrtf(rsqrtf(__fdividef(sc2, sn2))) * exp2f(q.snr * -0.0830482023721841f);
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e) o[e] = fmaf(n[e] * scalarnoise, noisescalar, o[e] * scalarclean);
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e) {
        const int off = (kPerC / 2 + e) * kRole;
        n[e] = (off < nt) ? __ldg(nb + off) : 0.0f;
      }
      if (do_prep) { prep_consume(1); prep_issue(2); }
#pragma unroll
      for (int e = 0; e < kPerC / 2; ++e)
        o[kPerC / 2 + e] = fmaf(n[e] * scalarnoise, noisescalar, o[kPerC / 2 + e] * scalarclean);
    } else if (do_prep) {
      prep_consume(1); prep_issue(2);
    }
    AUG_T(c2c); AUG_ACC(10, c2b, c2c);
    if (flags & WW_AUG_GAIN) {
      const float g = q.gain;
#pragma unroll
      for (int e = 0; e < kPerC; ++e) o[e] *= g;
    }
    float mo = 0.0f;
    if (flags & WW_AUG_NORM_OUT) {
#pragma unroll
      for (int e = 0; e < kPerC; ++e) mo = fmaxf(mo, fabsf(o[e]));
      mo = role_max_nonneg(mo, red3[2], t);
    }
    AUG_T(c2d); AUG_ACC(11, c2c, c2d);
    if (do_prep) { prep_consume(2); prep_issue(3); }
    if ((flags & WW_AUG_NORM_OUT) && mo > 0.0f) {
      const ClipDiv dv = make_clip_div(mo);
      if (inexact && dv.fast) {
#pragma unroll
        for (int e = 0; e < kPerC; ++e) o[e] = dv.approx(o[e]);
      } else {
#pragma unroll
        for (int e = 0; e < kPerC; ++e) o[e] = dv(o[e]);
      }
    }
    float* __restrict__ dst = p.out + (int64_t)((int)blockIdx.x + k * (int)gridDim.x) * N + t;
#pragma unroll
    for (int e = 0; e < kPerC; ++e)
      if (e * kRole < nt) dst[e * kRole] = o[e];
    if (do_prep) { prep_consume(3); prep_end(j); }
    m_in = m_mid; m_mid = m_prep;
    AUG_T(c3);
    AUG_ACC(6, c2, c3);
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

namespace {
// running sums of squares of one bank row per CTA (double): P[0] = 0, P[i + 1] = P[i] + x[i]^2
__global__ void __launch_bounds__(1024) bank_prefix_kernel(const float* __restrict__ bank, int64_t len, double* __restrict__ out) {
  __shared__ double part[1024];
  const float* x = bank + (int64_t)blockIdx.x * len;
  double* P = out + (int64_t)blockIdx.x * (len + 1);
  const int tid = threadIdx.x;
  const int64_t chunk = (len + blockDim.x - 1) / blockDim.x;
  const int64_t lo = min((int64_t)tid * chunk, len), hi = min(lo + chunk, len);
  double s = 0.0;
  for (int64_t i = lo; i < hi; ++i) s += (double)x[i] * (double)x[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {       // inclusive scan of the per-thread totals
    const double v = tid >= o ? part[tid - o] : 0.0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  double run = part[tid] - s;
  if (tid == 0) P[0] = 0.0;
  for (int64_t i = lo; i < hi; ++i) { run += (double)x[i] * (double)x[i]; P[i + 1] = run; }
}
}  // namespace

// Called once by every API entry that augments with `bank`, on the stream its augment launches follow.
int ww_prepare_bank_energy(ww_ctx* c, const float* bank, int bank_rows, int64_t bank_len, cudaStream_t st) {
  c->bank_prefix_src = nullptr;
  if (!bank || bank_rows <= 0 || bank_len <= 0) return WW_OK;
  const char* env = getenv("WW_AUG_BANK_SUMS");          // "0": every clip measures its noise segment itself (tests); read per call
  if (env && env[0] == '0') return WW_OK;
  const size_t need = (size_t)bank_rows * (size_t)(bank_len + 1);
  if (need > c->bank_prefix_cap) {
    if (c->d_bank_prefix) { WW_CHECK(c, cudaDeviceSynchronize()); cudaFree(c->d_bank_prefix); c->d_bank_prefix = nullptr; }
    WW_CHECK(c, cudaMalloc((void**)&c->d_bank_prefix, need * sizeof(double)));
    c->bank_prefix_cap = need;
  }
  bank_prefix_kernel<<<bank_rows, 1024, 0, st>>>(bank, bank_len, c->d_bank_prefix);
  WW_LAUNCH_CHECK(c);
  c->bank_prefix_src = bank; c->bank_prefix_rows = bank_rows; c->bank_prefix_len = bank_len;
  return WW_OK;
}

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

template <typename TIn>
static int launch_augment_t(ww_ctx* c, const AugKParams& p, cudaStream_t st) {
  const int NP = (p.N + 7) & ~7;
  const int grid = std::min(c->sm_count, p.B);
  const char* kern = getenv("WW_AUGMENT_KERNEL");        // "lockstep": the one-clip-per-CTA kernel (A/B runs, tests); read per call
  if (kern && strcmp(kern, "lockstep") == 0) {
    // stage x2 (+ pads when the fp32 stage doubles as the gather source) | float copy + pads (int16 input) | noise | table
    const size_t smem = (sizeof(TIn) == 4 ? (size_t)2 * (NP + 2 * kPad) * 4 : (size_t)2 * NP * 2 + (size_t)(NP + 2 * kPad) * 4) +
                        (size_t)NP * 4 + (size_t)kTblWords * 4;
    // opt in on every launch: the attribute is per device, a process may drive several
    WW_CHECK(c, cudaFuncSetAttribute(augment_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(c, WW_STAGE_AUGMENT, st);
    augment_kernel<TIn><<<grid, kThreads, smem, st>>>(p);
    WW_LAUNCH_CHECK(c);
    return WW_OK;
  }
  // rolled source x2 (with pads) | gathered clip | polyphase table x2
  const size_t smem = ((size_t)2 * (NP + 2 * kPad) + NP + 2 * kTblWords) * 4;
  WW_CHECK(c, cudaFuncSetAttribute(augment_pipe_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(c, WW_STAGE_AUGMENT, st);
  augment_pipe_kernel<TIn><<<grid, kPipeThreads, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
#ifdef WW_AUG_TRACE
  {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, aug_trace, sizeof(h));
    const int K = (p.B + grid - 1) / grid;
    fprintf(stderr, "[aug trace] CTA 0, %d clips, cycles per clip: G wait src %llu, wait res_empty %llu, gather %llu | C prep %llu, wait res_full %llu, finish %llu (res load %llu, reduce3 %llu, mix %llu, max reduce %llu)\n",
            K, h[0] / K, h[1] / K, h[2] / K, h[4] / K, h[5] / K, h[6] / K, h[8] / K, h[9] / K, h[10] / K, h[11] / K);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(aug_trace, h, sizeof(h));
  }
#endif
  return WW_OK;
}

int ww_launch_augment(ww_ctx* c, const void* clips, int pcm16, const float* bank, int bank_rows, int64_t bank_len,
                      const ww_aug* a, float* out, int B, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  AugKParams p;
  p.clips = clips; p.bank = bank; p.bank_rows = bank_rows; p.bank_len = bank_len;
  p.a = *a; p.out = out; p.B = B; p.N = c->cfg.n_samples;
  p.rs_desc = c->d_rs_desc; p.n_rs = (int)c->rs_tables.size(); p.rs_kern = c->d_rs_kern;
  p.bank_prefix = (bank && c->bank_prefix_src == bank && c->bank_prefix_rows == bank_rows && c->bank_prefix_len == bank_len)
                      ? c->d_bank_prefix : nullptr;
  if (p.N > kThreads * kMaxPerThread) {
    c->set_error("ww_augment: n_samples too large (max 16384 samples per clip)");
    return WW_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(clips) & 3) || (pcm16 && (p.N & 1))) {
    c->set_error("ww_augment: clips must be 4-byte aligned (and n_samples even for int16 PCM)");
    return WW_ERR_INVALID;
  }
  return pcm16 ? launch_augment_t<int16_t>(c, p, st) : launch_augment_t<float>(c, p, st);
}

// K2 on the tensor core: framing + Hann + STFT power + mel + dB for the reference's own preset (n_fft = win = 2048, hop 512,
// 16,000 samples -> 80 x 32), the DFT as two tcgen05 GEMM stages (sm_100a).
//
// Replaces AudioProcessor.audio_to_mel (/root/reference/wakeword_training_script.py:85-101: librosa melspectrogram +
// power_to_db(ref=max)) like logmel.cu does; logmel.cu (shared-memory FFT) stays the kernel of every other configuration
// and of the streaming modes.
//
// Factorisation 2048 = 64 x 32, n = n1 + 64 n2, k = 32 k1 + k2 (n1, k1 < 64; n2, k2 < 32):
//     X[32 k1 + k2] = sum_n1 W64^(n1 k1) [ W2048^(n1 k2) sum_n2 x[n1 + 64 n2] W32^(n2 k2) ]
//   * the zero-padded clip (18,432 samples = 288 rows of 64) lives in shared memory ONCE as fp16 hi and lo copies of
//     x * 2^k / peak-normalised scale; frame t starts at row 8 t, so the A operand of a frame pair is an MN-major
//     SWIZZLE_NONE *view* of those bytes (start t * 1 KB, M-core stride 128 B, K-core stride 1 KB): nothing is framed or
//     windowed on CUDA cores (validated by tools/mn_major_view_check.cu and tools/gemm_fft_proto.cu on a B200);
//   * stage 1 (K = 32):  D1[(t, n1), (re | im of k2) x (hi | lo of F32)] = x_hi [F_hi | F_lo] + x_lo F_hi   (3 products);
//   * epilogue 1 (thread = (t, n1)): twiddle W2048^(n1 k2), the Hann window as the 3-tap filter
//     0.5 Y[k2] - 0.25 (Y[k2-1] + Y[k2+1]) over k2 (with the k1 carry W64^(+-n1) at the ends), fp16 hi/lo split ->
//     Y' = stage-3 A operand (MN-major, M = (frame, k2), K = n1 re | n1 im);
//   * stage 3 (K = 128, N = 64 = re | im of k1 < 32):  D3 = Y'_hi F64_hi + Y'_lo F64_hi + Y'_hi F64_lo;
//     bin 1024 (k1 = 32) is not computed: the launcher takes this kernel only when no mel filter touches it (true for the
//     reference's fmax = 8000 = Nyquist, where librosa's last filter ends at bin 1023);
//   * epilogue 3 (thread = (frame, k2)): power of bins 32 k1 + k2 -> shared-memory tile P4[bin] = (4 frames);
//   * mel warps: banded projection (2-16 lanes per band, one 128-bit load serves 4 frames), then per clip max -> dB -> store.
// One persistent CTA per SM, 16 warps (128 registers each: the epilogue-1 warps keep a whole Y' row in registers so that
// only their 16 stores sit between two stage-3 GEMMs): warp 0 MMA issuer | 1-3 converter (peak of the NEXT clip, then
// fp32 / int16 -> fp16 hi / lo in chunks of 2,048 padded samples that are released group by group, so conversion overlaps
// the GEMMs of the previous clip) | 4-11 epilogue 1 (two frame pairs in flight) | 12-15 epilogue 3 + mel + dB.
// Peak normalisation is a scale of the power spectrum: the clip is always transformed at unit peak (any input magnitude
// fits fp16 hi/lo), and `normalize = 0` multiplies the mel energies by peak^2 again.
#include "tc_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <vector>

using namespace tc;

namespace {

constexpr int kNfft = 2048, kN1 = 64, kN2 = 32, kW = 32, kSamples = 16000, kHop = 512, kBins = kNfft / 2 + 1;
constexpr int kRows = 288;                           // padded clip rows of 64 samples
constexpr int kChunks = 9;                           // converter chunks of 32 rows (2,048 padded samples)
constexpr int kClipBytes = kRows * 64 * 2;           // one fp16 copy of the padded clip: [row block][col chunk 8][row % 8][8]
constexpr int kF32Bytes = 4 * 128 * 16;              // stage-1 B: [kc 4][n 128 = (re, im) x 32 of F_hi | the same of F_lo][8]
constexpr int kN3 = 64;                              // stage-3 N: re of k1 0..31 | im of k1 0..31
constexpr int kF64Bytes = 16 * kN3 * 16;             // stage-3 B (hi or lo): [kc 16][n 64][8]
constexpr int kYBytes = 16 * 2048;                   // stage-3 A (hi or lo), MN-major: [m chunk 16][k row 128][8]
constexpr int kP4Bytes = 1024 * 16;                  // power tile: [bin 1024] x (4 frames) fp32
constexpr int kMelPitch = kW + 1;
constexpr int kMaxMels = 80, kMaxNnz = 2304, kMaxTasks = 1024;
constexpr float kFScale = 1024.0f, kYScale = 4096.0f;
constexpr int kWarps = 16, kThreads = kWarps * 32;      // 512 threads: 128 registers each (ptxas budgets the whole kernel by
                                                       // the smallest setmaxnreg value, so the roles share one budget instead)
constexpr int kConvWarps = 3;
constexpr bool kTcDefault = false;                     // which kernel ww_logmel takes where both apply (see profiles/: A/B)
constexpr unsigned kPollNs = 32;                       // sleep between mbarrier polls (see mbar_wait_sleep)
constexpr int kMelRounds = 4, kMelTaps = 8;            // lane-tasks per epilogue-3 thread; taps per lane-task (registers)
constexpr float kAmin = 1e-10f, kTopDb = 80.0f;

constexpr size_t kSmem = 2 * (size_t)kClipBytes + kF32Bytes + 2 * (size_t)kF64Bytes + 2 * (size_t)kYBytes + kP4Bytes +
                         (size_t)kMaxMels * kMelPitch * 4 + (size_t)kMaxNnz * 4 + (size_t)kMaxTasks * 4 +
                         (size_t)3 * kMaxMels * 4 + 64 * 8 + 128;

__host__ __device__ constexpr uint32_t idesc_a_mn(int M, int N) { return make_idesc(M, N) | (1u << 15); }

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
// v -> fp16 hi + fp16 lo (v = hi + lo to ~22 bits), two values at a time
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16(a, b);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16(a - h.x, b - h.y);
}
// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): one instruction for both components of a complex value
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 un2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void tmem_ld2_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 8 consecutive samples starting at a multiple of 8 (16-byte aligned rows: checked by the launcher)
__device__ __forceinline__ void ld8(const float* x, int i, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + i)), b = __ldg(reinterpret_cast<const float4*>(x + i + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const int16_t* x, int i, float* v) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + i));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = (float)(int16_t)(w[k] & 0xffffu) * (1.0f / 32768.0f);
    v[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f);
  }
}

struct TcLogmelParams {
  const void* clips;
  int64_t clip_stride;           // samples
  LogmelOut out;
  int B, normalize, n_mels, mel_nnz, n_tasks;
  const __half* f32;             // pre-tiled stage-1 B
  const __half* f64hi;           // pre-tiled stage-3 B
  const __half* f64lo;
  const float2* tw;              // [64][32] exp(-2 pi i n1 k2 / 2048)
  const float2* rot;             // [64]     exp(-2 pi i n1 / 64)
  const float* mel_w;            // packed non-zero filterbank weights
  const int* mel_start;          // [n_mels] first bin / bins / offset into mel_w
  const int* mel_len;
  const int* mel_off;
  const uint32_t* tasks;         // mel | lane j << 8 | lanes << 16, groups of `lanes` consecutive entries, lanes descending
  long long* trace;              // debug (WW_TC_TRACE=1): per-group role timestamps of CTA 0
};

#define LM_TRACE(G, slot) do { if (p.trace && blockIdx.x == 0 && lane == 0 && (G) < 40) p.trace[(G) * 16 + (slot)] = clock64(); } while (0)

// Banded mel projection of one power tile (4 frames) by the four epilogue-3 warps.  Every band has 2-16 lanes (entries of
// `tasks`, aligned groups); a thread owns the same kMelRounds lane-tasks for the whole kernel, so their bin offsets and
// weights (at most kMelTaps per lane) live in registers: per group a task is kMelTaps 128-bit loads (one serves the four
// frames), the FMAs, a shuffle reduction over the band's lanes and one store by lane 0: mel_s[mel][4 g + f].
struct MelTask {
  float w[kMelTaps];
  int base, step, cnt, lanes, dst;     // first bin, bin stride (= lanes), valid taps, lanes of the band, mel * kMelPitch or -1
};
__device__ __forceinline__ void mel_task_init(MelTask& k, const float* melw, const int* mst, const uint32_t* tasks, int n_tasks, int t) {
  const uint32_t e = t < n_tasks ? tasks[t] : (1u << 16);
  const int mel = e & 255, j = (e >> 8) & 255, lanes = (int)(e >> 16);
  const int st = mst[mel], len = t < n_tasks ? mst[kMaxMels + mel] : 0, off = mst[2 * kMaxMels + mel];
  k.base = st + j; k.step = lanes; k.lanes = lanes;
  k.cnt = len > j ? (len - j + lanes - 1) / lanes : 0;
  k.dst = t < n_tasks ? mel * kMelPitch : -1;          // every lane of the band: lanes j < 4 store one frame each
#pragma unroll
  for (int u = 0; u < kMelTaps; ++u) k.w[u] = u < k.cnt ? melw[off + j + u * lanes] : 0.0f;
}
// Two lane-tasks at once (their loads, FMAs and reductions interleave: the four warps that run this are latency bound).
// Reduction over the L lanes of a band by exchanging halves: after the xor-1 step a lane keeps two of the four frames, after
// the xor-2 step one (frame 2 (j & 1) + ((j >> 1) & 1)); xor 4 / 8 then sum whole values: 5 shuffles instead of 16.
__device__ __forceinline__ void mel_task_run2(const MelTask& ka, const MelTask& kb, const float4* __restrict__ p4,
                                              float* __restrict__ mel_s, int g, float pscale, int lane) {
  // unconditional loads (taps past the band have weight 0 and read a clamped, valid bin): all are in flight at once
  float4 va[kMelTaps], vb[kMelTaps];
#pragma unroll
  for (int u = 0; u < kMelTaps; ++u) va[u] = p4[min(ka.base + u * ka.step, 1023)];
#pragma unroll
  for (int u = 0; u < kMelTaps; ++u) vb[u] = p4[min(kb.base + u * kb.step, 1023)];
  float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b = a;
#pragma unroll
  for (int u = 0; u < kMelTaps; ++u) {
    a.x = fmaf(ka.w[u], va[u].x, a.x); a.y = fmaf(ka.w[u], va[u].y, a.y);
    a.z = fmaf(ka.w[u], va[u].z, a.z); a.w = fmaf(ka.w[u], va[u].w, a.w);
    b.x = fmaf(kb.w[u], vb[u].x, b.x); b.y = fmaf(kb.w[u], vb[u].y, b.y);
    b.z = fmaf(kb.w[u], vb[u].z, b.z); b.w = fmaf(kb.w[u], vb[u].w, b.w);
  }
  const bool odd = lane & 1, hi2 = lane & 2;          // lane groups are aligned to their size: lane bits = bits of j
  // xor 1: even lanes keep frames 0, 1, odd lanes frames 2, 3
  float a0 = odd ? a.z : a.x, a1 = odd ? a.w : a.y, b0 = odd ? b.z : b.x, b1 = odd ? b.w : b.y;
  a0 += __shfl_xor_sync(0xffffffffu, odd ? a.x : a.z, 1); a1 += __shfl_xor_sync(0xffffffffu, odd ? a.y : a.w, 1);
  b0 += __shfl_xor_sync(0xffffffffu, odd ? b.x : b.z, 1); b1 += __shfl_xor_sync(0xffffffffu, odd ? b.y : b.w, 1);
  // xor 2 (bands of 4+ lanes): lanes with bit 1 clear keep the first of their two frames, the others the second
  const float sa = __shfl_xor_sync(0xffffffffu, hi2 ? a0 : a1, 2), sb = __shfl_xor_sync(0xffffffffu, hi2 ? b0 : b1, 2);
  float ra = (hi2 ? a1 : a0) + sa, rb = (hi2 ? b1 : b0) + sb;
#pragma unroll
  for (int o = 4; o <= 8; o <<= 1) {
    const float ta = __shfl_xor_sync(0xffffffffu, ra, o), tb = __shfl_xor_sync(0xffffffffu, rb, o);
    if (o < ka.lanes) ra += ta;
    if (o < kb.lanes) rb += tb;
  }
  // stores: bands of 2 lanes hold two frames per lane (2 (j & 1), + 1), wider bands one frame in their lanes j < 4
  const int j = lane & 15;
  if (ka.dst >= 0) {
    float* ms = mel_s + ka.dst + 4 * g;
    if (ka.lanes == 2) { ms[2 * (j & 1)] = a0 * pscale; ms[2 * (j & 1) + 1] = a1 * pscale; }
    else if ((j & (ka.lanes - 1)) < 4) ms[2 * (j & 1) + ((j >> 1) & 1)] = ra * pscale;
  }
  if (kb.dst >= 0) {
    float* ms = mel_s + kb.dst + 4 * g;
    if (kb.lanes == 2) { ms[2 * (j & 1)] = b0 * pscale; ms[2 * (j & 1) + 1] = b1 * pscale; }
    else if ((j & (kb.lanes - 1)) < 4) ms[2 * (j & 1) + ((j >> 1) & 1)] = rb * pscale;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1) logmel_tc_kernel(const __grid_constant__ TcLogmelParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* xhi = smem;
  unsigned char* xlo = xhi + kClipBytes;
  unsigned char* f32 = xlo + kClipBytes;
  unsigned char* f64hi = f32 + kF32Bytes;
  unsigned char* f64lo = f64hi + kF64Bytes;
  unsigned char* yhi = f64lo + kF64Bytes;
  unsigned char* ylo = yhi + kYBytes;
  float4* p4 = reinterpret_cast<float4*>(ylo + kYBytes);
  float* mel_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(p4) + kP4Bytes);   // [n_mels][33]
  float* melw = mel_s + kMaxMels * kMelPitch;
  uint32_t* tasks = reinterpret_cast<uint32_t*>(melw + kMaxNnz);
  int* mst = reinterpret_cast<int*>(tasks + kMaxTasks);                                      // [3][kMaxMels]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mst + 3 * kMaxMels);
  uint64_t* x_full = bars;            // [9]  converter warps -> issuer (chunk of 32 rows converted)
  uint64_t* x_empty = bars + 9;       // [9]  stage-1 MMAs that read the chunk are done -> converter
  uint64_t* d1_full = bars + 18;      // [2]  pair h
  uint64_t* d1_empty = bars + 20;     // [2]
  uint64_t* y_full = bars + 22;       // [1]  8 epilogue-1 warps
  uint64_t* y_empty = bars + 23;      // [1]  stage-3 MMAs done
  uint64_t* d3_full = bars + 24;      // [2]
  uint64_t* d3_empty = bars + 26;     // [2]
  float2* scale_ring = reinterpret_cast<float2*>(bars + 32);   // [4] per clip: (1 / (2^k peak F), mel energy factor)
  float* red = reinterpret_cast<float*>(scale_ring + 4);       // [8] reduction scratch: [0,4) converter, [4,8) mel warps
  uint32_t* slot = reinterpret_cast<uint32_t*>(red + 8);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int n_mels = p.n_mels;

  // ---- one-time setup: tables, zero padding of the clip copies, barriers, TMEM
  for (int i = tid * 16; i < kF32Bytes; i += kThreads * 16)
    *reinterpret_cast<uint4*>(f32 + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f32) + i);
  for (int i = tid * 16; i < kF64Bytes; i += kThreads * 16) {
    *reinterpret_cast<uint4*>(f64hi + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64hi) + i);
    *reinterpret_cast<uint4*>(f64lo + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64lo) + i);
  }
  for (int i = tid * 16; i < kClipBytes; i += kThreads * 16) {     // centre padding (rows 0-15, 266-287) stays zero
    *reinterpret_cast<uint4*>(xhi + i) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(xlo + i) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < p.mel_nnz; i += kThreads) melw[i] = p.mel_w[i];
  for (int i = tid; i < p.n_tasks; i += kThreads) tasks[i] = p.tasks[i];
  for (int i = tid; i < n_mels; i += kThreads) {
    mst[i] = p.mel_start[i]; mst[kMaxMels + i] = p.mel_len[i]; mst[2 * kMaxMels + i] = p.mel_off[i];
  }
  if (tid == 0) {
    for (int i = 0; i < kChunks; ++i) { mbar_init(x_full + i, kConvWarps); mbar_init(x_empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full + i, 1); mbar_init(d1_empty + i, 4);
      mbar_init(d3_full + i, 1); mbar_init(d3_empty + i, 4);
    }
    mbar_init(y_full, 8); mbar_init(y_empty, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);          // D1: columns 128 h (h < 2); D3: columns 256 + 64 b (b < 2)
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  const int n_my = ((int)blockIdx.x < p.B) ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // clips of this CTA
  const uint32_t n_groups = (uint32_t)n_my * 8u;

  if (warp == 0) {
    // ===================== MMA issuer: S1(G + 1) is issued before S3(G), across clip boundaries too
    const uint64_t b1 = make_desc(smem_u32(f32), 128 * 16, 128);
    const uint64_t b3h = make_desc(smem_u32(f64hi), kN3 * 16, 128), b3l = make_desc(smem_u32(f64lo), kN3 * 16, 128);
    const uint64_t a3h = make_desc(smem_u32(yhi), 128, 2048), a3l = make_desc(smem_u32(ylo), 128, 2048);
    auto stage1 = [&](uint32_t G) {
      const uint32_t ci = G >> 3, g = G & 7;
      if (g == 0) mbar_wait_sleep(x_full + 0, ci & 1, 10, kPollNs);
      mbar_wait_sleep(x_full + g + 1, ci & 1, 11, kPollNs);                       // frames 4g .. 4g+3 read rows 32 g .. 32 g + 55
      LM_TRACE(G, 0);
#pragma unroll 1
      for (uint32_t h = 0; h < 2; ++h) {
        mbar_wait_sleep(d1_empty + h, (G & 1) ^ 1, 12, kPollNs);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t0 = 4 * g + 2 * h, d = tm + h * 128;
          const uint64_t ah = make_desc(smem_u32(xhi) + t0 * 1024u, 1024, 128), al = make_desc(smem_u32(xlo) + t0 * 1024u, 1024, 128);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            umma_f16(d, ah + (uint64_t)((s * 2 * 1024) >> 4), b1 + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 128), s);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            umma_f16(d, al + (uint64_t)((s * 2 * 1024) >> 4), b1 + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 64), 1);
          umma_commit(d1_full + h);
          if (h == 1) {
            umma_commit(x_empty + g);                              // chunk g is not read by any later group of this clip
            if (g == 7) umma_commit(x_empty + 8);
          }
        }
        __syncwarp();
      }
    };
    auto stage3 = [&](uint32_t G) {
      const uint32_t buf = G & 1;
      mbar_wait_sleep(y_full, G & 1, 13, kPollNs);
      LM_TRACE(G, 1);
      mbar_wait_sleep(d3_empty + buf, ((G >> 1) & 1) ^ 1, 14, kPollNs);
      LM_TRACE(G, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tm + 256 + buf * kN3;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint64_t a = prod == 1 ? a3l : a3h, b = prod == 2 ? b3l : b3h;
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_f16(d, a + (uint64_t)((s * 2 * 128) >> 4), b + (uint64_t)((s * 2 * kN3 * 16) >> 4), idesc_a_mn(128, kN3), (prod | s) != 0);
        }
        umma_commit(d3_full + buf);
        umma_commit(y_empty);
      }
      __syncwarp();
    };
    if (n_groups) stage1(0);
    for (uint32_t G = 0; G < n_groups; ++G) {
      if (G + 1 < n_groups) stage1(G + 1);
      stage3(G);
    }
  } else if (warp < 4) {
    // ===================== converter (runs ahead of the GEMMs by up to a clip): peak of the clip, then per chunk 8 samples ->
    // one 16-byte store into each fp16 copy.  Global loads are issued in batches so that a chunk costs one round trip.
    constexpr int CT = kConvWarps * 32;
    const int ct = tid - 32;                                       // 0 .. 95
    for (int ci = 0; ci < n_my; ++ci) {
      const int clip = (int)blockIdx.x + ci * (int)gridDim.x;
      const TIn* x = static_cast<const TIn*>(p.clips) + (int64_t)clip * p.clip_stride;
      float mx = 0.0f;
      for (int u0 = ct; u0 < kSamples / 8; u0 += 4 * CT) {
        float v[4][8];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int u = u0 + b * CT;
          if (u < kSamples / 8) ld8(x, 8 * u, v[b]);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[b][e] = 0.0f;
          }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int e = 0; e < 8; ++e) mx = fmaxf(mx, fabsf(v[b][e]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      named_sync(1, CT);                                           // red[0..2] of the previous clip has been read
      if (lane == 0) red[ct >> 5] = mx;
      named_sync(1, CT);
      const float peak = fmaxf(fmaxf(red[0], red[1]), red[2]);
      // x * xs has its peak in [2048, 4096): hi is an 11-bit fp16, lo the next 11 bits.  A silent clip transforms to zeros.
      float xs = 0.0f, inv1 = 0.0f;
      if (peak > 0.0f && peak < 3.0e38f) {
        int ex;
        frexpf(peak, &ex);                                         // peak = fr 2^ex, fr in [0.5, 1)
        xs = ldexpf(1.0f, min(12 - ex, 100));
        inv1 = 1.0f / (xs * peak * kFScale);                       // spectrum of the clip at unit peak
      }
      if (ct == 0) scale_ring[ci & 3] = make_float2(inv1, p.normalize ? 1.0f : peak * peak);
      for (int c = 0; c < kChunks; ++c) {
        // chunk c = padded samples [2048 c, 2048 c + 2048) = clip samples [2048 c - 1024, 2048 c + 1024): <= 256 units of 8
        const int s_lo = max(0, 2048 * c - 1024), s_hi = min(kSamples, 2048 * c + 1024);
        float v[3][8];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int u = s_lo / 8 + ct + b * CT;
          if (u < s_hi / 8) ld8(x, 8 * u, v[b]);
        }
        {                                                          // the converter runs far ahead: sleep between polls
          uint32_t spins = 0;
          while (!mbar_try_wait(x_empty + c, (ci & 1) ^ 1)) {
            __nanosleep(500);
            if (++spins > (1u << 22)) mbar_timeout(50);
          }
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int u = s_lo / 8 + ct + b * CT;
          if (u < s_hi / 8) {
            const int i = kNfft / 2 + 8 * u, r = i >> 6, c8 = (i & 63) >> 3;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(v[b][2 * e] * xs, v[b][2 * e + 1] * xs, hi[e], lo[e]);
            const uint32_t off = ((((r >> 3) * 8 + c8) * 8 + (r & 7)) * 8) * 2;
            *reinterpret_cast<uint4*>(xhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(xlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive_warp(x_full + c, lane);
      }
    }
  } else if (warp < 12) {
    // ===================== epilogue 1: thread = (frame 2h + f of the group, n1); TMEM lane quadrant = warp & 3
    const int q = warp & 3, h = (warp - 4) >> 2, m = q * 32 + lane, f = m >> 6, n1 = m & 63;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    const float2 rp = __ldg(p.rot + n1);                           // W64^(n1)
    // twiddles w_k = W2048^(n1 k) inside a chunk of 8 by the three-term recurrence w_{k+1} = 2 cos(t) w_k - w_{k-1} (two FMAs
    // per step, at most 7 steps from an exact-to-1-ulp restart), chunk anchors W2048^(8 c n1) by products of w8
    const float2 w1 = __ldg(p.tw + n1 * kN2 + 1), w8 = __ldg(p.tw + n1 * kN2 + 8);
    const float tc2 = 2.0f * w1.x;
    const uint32_t dbase = tm + h * 128 + lane_addr;
    const uint32_t off0 = (uint32_t)(2 * h + f) * 4u * 2048u + (uint32_t)(n1 >> 3) * 128u + (uint32_t)(n1 & 7) * 16u;
    for (uint32_t G = 0; G < n_groups; ++G) {
      mbar_wait_sleep(d1_full + h, G & 1, 20, kPollNs);
      if (warp == 4) LM_TRACE(G, 3);
      tc_fence_after();
      const float inv = scale_ring[(G >> 3) & 3].x;                // written by the converter before the clip's first chunk
      const float c1 = 0.5f * (kYScale / 32.0f) * inv, c2 = 0.25f * (kYScale / 32.0f) * inv;
      const unsigned long long pc1 = pk2(c1, c1), nc2 = pk2(-c2, -c2);
      // D1 columns: (re, im) of k2 at 2 k2, 2 k2 + 1 from the F_hi half, the same again at 64 + ... from the F_lo half
      unsigned long long p2, p1, y0;                                 // Y[k-2], Y[k-1] of the running 3-tap, Y[0] (packed re, im)
      {
        // left end: Y[-1] = W64^(-n1) Y[31] = W64^(-n1) W2048^(31 n1) D[31] = conj(w1) D[31]
        uint32_t a[2], b[2];
        tmem_ld2_nowait(dbase + 62, a);
        tmem_ld2_nowait(dbase + 126, b);
        tmem_ld_wait();
        const float dr = __uint_as_float(a[0]) + __uint_as_float(b[0]), di = __uint_as_float(a[1]) + __uint_as_float(b[1]);
        p2 = pk2(dr * w1.x + di * w1.y, di * w1.x - dr * w1.y);
      }
      uint32_t zrh[16], zrl[16], zih[16], zil[16];                  // the whole Y' row of this thread, packed fp16 pairs
      float zr_e = 0.0f, zi_e = 0.0f;
      // Hann 3-tap with one element of lag: z[k-1] = c1 Y[k-1] - c2 (Y[k-2] + Y[k]), both components per instruction
      auto emit = [&](int k, unsigned long long l, unsigned long long m_, unsigned long long r) {
        const float2 z = un2(fma2(add2(l, r), nc2, mul2(m_, pc1)));
        if (k & 1) {
          split2(zr_e, z.x, zrh[k >> 1], zrl[k >> 1]);
          split2(zi_e, z.y, zih[k >> 1], zil[k >> 1]);
        } else {
          zr_e = z.x; zi_e = z.y;
        }
      };
      // chunk c of the row (k2 = 8 c .. 8 c + 7) -> Y' (K rows n1 and 64 + n1 of m chunk 4 fslot + c)
      auto store_chunk = [&](int c) {
        const uint32_t off_r = off0 + c * 2048u, off_i = off_r + 8 * 128;
        *reinterpret_cast<uint4*>(yhi + off_r) = make_uint4(zrh[4 * c], zrh[4 * c + 1], zrh[4 * c + 2], zrh[4 * c + 3]);
        *reinterpret_cast<uint4*>(ylo + off_r) = make_uint4(zrl[4 * c], zrl[4 * c + 1], zrl[4 * c + 2], zrl[4 * c + 3]);
        *reinterpret_cast<uint4*>(yhi + off_i) = make_uint4(zih[4 * c], zih[4 * c + 1], zih[4 * c + 2], zih[4 * c + 3]);
        *reinterpret_cast<uint4*>(ylo + off_i) = make_uint4(zil[4 * c], zil[4 * c + 1], zil[4 * c + 2], zil[4 * c + 3]);
      };
      float2 wa = make_float2(1.0f, 0.0f);                           // anchor W2048^(8 c n1)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rh[8], rl[8];
        float2 wc = wa, wm = make_float2(wa.x * w1.x + wa.y * w1.y, wa.y * w1.x - wa.x * w1.y);   // w_0, w_{-1} = w_0 conj(w1)
        if (c < 3) wa = cmul(wa, w8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = 8 * c + e;
          if ((e & 3) == 0) {                                        // four k2 at a time (registers)
            tmem_ld8_nowait(dbase + 16 * c + 2 * e, rh);
            tmem_ld8_nowait(dbase + 64 + 16 * c + 2 * e, rl);
            tmem_ld_wait();
            if (k == 28) {                                           // every TMEM read of this pair is done
              tc_fence_before();
              mbar_arrive_warp(d1_empty + h, lane);
            }
          }
          const int e4 = e & 3;
          const float2 d = un2(add2(pk2(__uint_as_float(rh[2 * e4]), __uint_as_float(rh[2 * e4 + 1])),
                                    pk2(__uint_as_float(rl[2 * e4]), __uint_as_float(rl[2 * e4 + 1]))));
          const float2 yv = cmul(d, wc);
          const unsigned long long y = pk2(yv.x, yv.y);
          if (e < 7) {
            const float2 wn = make_float2(fmaf(tc2, wc.x, -wm.x), fmaf(tc2, wc.y, -wm.y));
            wm = wc; wc = wn;
          }
          if (k == 0) { y0 = y; p1 = y; }
          else { emit(k - 1, p2, p1, y); p2 = p1; p1 = y; }
          if (k == 16) {
            // chunks 0-2 are complete and wait in registers (the whole row would not fit): by now the stage-3 GEMM of the
            // previous group (1,450 cycles) has normally released Y'; only the last quarter of the row is computed after it
            if (warp == 4) LM_TRACE(G, 4);
            mbar_wait_sleep(y_empty, (G & 1) ^ 1, 21, kPollNs);
            if (warp == 4) LM_TRACE(G, 5);
            store_chunk(0); store_chunk(1);
          }
          if (k == 24) {
            store_chunk(2);
          }
        }
      }
      {
        const float2 y0v = un2(y0);
        const float2 y32 = cmul(y0v, rp);                            // Y[32] = W64^(n1) Y[0]
        emit(31, p2, p1, pk2(y32.x, y32.y));
      }
      store_chunk(3);
      fence_proxy_async();
      mbar_arrive_warp(y_full, lane);
      if (warp == 4) LM_TRACE(G, 6);
      if (warp == 11) LM_TRACE(G, 15);
    }
  } else {
    // ===================== epilogue 3 + mel: thread = (frame slot q of the group, k2 = lane) for the power tile; then the
    // same four warps project the tile onto the mel bands; per clip: max -> dB -> store
    const int q = warp & 3, mt = tid - 12 * 32;                    // mt 0 .. 127
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    const float inv = 32.0f / (kYScale * kFScale);
    float* pf = reinterpret_cast<float*>(p4) + q;                  // P4[bin].f[q]
    MelTask mk[kMelRounds];
#pragma unroll
    for (int r = 0; r < kMelRounds; ++r) mel_task_init(mk[r], melw, mst, tasks, p.n_tasks, r * 128 + mt);
    for (uint32_t G = 0; G < n_groups; ++G) {
      const uint32_t buf = G & 1, ci = G >> 3, g = G & 7;
      mbar_wait_sleep(d3_full + buf, (G >> 1) & 1, 30, kPollNs);
      if (warp == 12) LM_TRACE(G, 7);
      tc_fence_after();
      const uint32_t d = tm + 256 + buf * kN3 + lane_addr;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t cr[16], cim[16];
        tmem_ld16_nowait(d + 16 * half, cr);
        tmem_ld16_nowait(d + 32 + 16 * half, cim);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive_warp(d3_empty + buf, lane);
        }
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          const float re = __uint_as_float(cr[k1]) * inv, im = __uint_as_float(cim[k1]) * inv;
          pf[(32 * (k1 + 16 * half) + lane) * 4] = re * re + im * im;
        }
      }
      LM_TRACE(G, 8 + (warp - 12));
      named_sync(2, 128);                                          // the power tile of the group is complete
      if (warp == 12) LM_TRACE(G, 12);
      const float pscale = scale_ring[ci & 3].y;
#pragma unroll
      for (int r = 0; r < kMelRounds; r += 2) mel_task_run2(mk[r], mk[r + 1], p4, mel_s, (int)g, pscale, lane);
      if (warp == 12) LM_TRACE(G, 13);
      named_sync(2, 128);                                          // every band of the group is in mel_s; the tile is free
      if (warp == 12) LM_TRACE(G, 14);
      if (g == 7) {
        // ---- power_to_db(ref = max, amin, top_db) over the clip's [n_mels][32] tile (same operations as logmel.cu)
        const int clip = (int)blockIdx.x + (int)ci * (int)gridDim.x;
        const int total = n_mels * kW;
        float mx = 0.0f;
        for (int i = mt; i < total; i += 128) mx = fmaxf(mx, mel_s[i + (i >> 5)]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[4 + (mt >> 5)] = mx;
        named_sync(2, 128);
        const float ref = fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7]));
        const float ref_db = __fmul_rn(10.0f, log10f(fmaxf(kAmin, ref)));
        float* __restrict__ o = p.out.ptr + (int64_t)clip * p.out.stride + p.out.off;
        for (int i = mt; i < total; i += 128) {
          const int mm = i >> 5;
          const float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(kAmin, mel_s[i + mm]))), ref_db);
          o[mm * p.out.pitch + (i & 31)] = fmaxf(v, -kTopDb);
        }
        named_sync(2, 128);                                          // red / mel_s are rewritten by the next clip
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

void split_host(double v, uint16_t& hi, uint16_t& lo) {
  hi = f2h((float)v);
  lo = f2h((float)(v - (double)h2f(hi)));
}

}  // namespace

// DFT matrices (fp16 hi / lo, UMMA canonical layouts), twiddles and the mel task table: built once per context
static int logmel_tc_prepare(ww_ctx* c) {
  if (c->tc_lm_ready) return c->tc_lm_ready > 0 ? WW_OK : 1;
  c->tc_lm_ready = -1;
  const ww_config& g = c->cfg;
  if (g.n_fft != kNfft || g.win_length != kNfft || g.hop_length != kHop || g.n_samples != kSamples ||
      g.n_mels > kMaxMels || c->mel_nnz > kMaxNnz || c->W != kW)
    return 1;
  // no filter may touch bin 1024 (not computed); lanes per band by length
  std::vector<uint32_t> tasks;
  const int lane_opts[4] = {16, 8, 4, 2};
  for (int lo : lane_opts)
    for (int m = 0; m < g.n_mels; ++m) {
      const int len = c->h_mel_len[m];
      if (len > 0 && c->h_mel_start[m] + len > kBins - 1) return 1;
      const int want = len > 53 ? 16 : len > 26 ? 8 : len > 13 ? 4 : 2;
      if (len > kMelTaps * want) return 1;
      if (want != lo) continue;
      for (int j = 0; j < lo; ++j) tasks.push_back((uint32_t)m | ((uint32_t)j << 8) | ((uint32_t)lo << 16));
    }
  if ((int)tasks.size() > kMelRounds * 128) return 1;
  const double PI = 3.14159265358979323846;
  std::vector<uint16_t> f32(kF32Bytes / 2, 0), f64h(kF64Bytes / 2, 0), f64l(kF64Bytes / 2, 0);
  for (int k2 = 0; k2 < kN2; ++k2)
    for (int n2 = 0; n2 < kN2; ++n2) {
      const double a = -2 * PI * ((k2 * n2) % kN2) / kN2;
      uint16_t rh, rl, ih, il;
      split_host(cos(a) * kFScale, rh, rl);
      split_host(sin(a) * kFScale, ih, il);
      auto at = [&](int n) -> uint16_t& { return f32[((size_t)(n2 / 8) * 128 + n) * 8 + n2 % 8]; };
      at(2 * k2) = rh; at(2 * k2 + 1) = ih; at(64 + 2 * k2) = rl; at(65 + 2 * k2) = il;   // (re, im) adjacent: packed fp32 adds
    }
  for (int k1 = 0; k1 < 32; ++k1)
    for (int n1 = 0; n1 < kN1; ++n1) {
      const double a = -2 * PI * ((k1 * n1) % kN1) / kN1, fr = cos(a) * kFScale, fi = sin(a) * kFScale;
      auto put = [&](int n, int k, double v) {
        uint16_t h, l;
        split_host(v, h, l);
        const size_t e = ((size_t)(k / 8) * kN3 + n) * 8 + k % 8;
        f64h[e] = h; f64l[e] = l;
      };
      put(k1, n1, fr); put(k1, 64 + n1, -fi);           // re: Yr cos - Yi sin
      put(32 + k1, n1, fi); put(32 + k1, 64 + n1, fr);  // im: Yr sin + Yi cos
    }
  std::vector<float2> tw((size_t)kN1 * kN2), rot(kN1);
  for (int n1 = 0; n1 < kN1; ++n1) {
    for (int k2 = 0; k2 < kN2; ++k2)
      tw[(size_t)n1 * kN2 + k2] = make_float2((float)cos(-2 * PI * n1 * k2 / kNfft), (float)sin(-2 * PI * n1 * k2 / kNfft));
    rot[n1] = make_float2((float)cos(-2 * PI * n1 / kN1), (float)sin(-2 * PI * n1 / kN1));
  }
  auto up = [&](void** dst, const void* src, size_t bytes) -> int {
    WW_CHECK(c, cudaMalloc(dst, bytes));
    WW_CHECK(c, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return WW_OK;
  };
  int rc;
  if ((rc = up((void**)&c->d_tc_f32, f32.data(), kF32Bytes))) return rc;
  if ((rc = up((void**)&c->d_tc_f64hi, f64h.data(), kF64Bytes))) return rc;
  if ((rc = up((void**)&c->d_tc_f64lo, f64l.data(), kF64Bytes))) return rc;
  if ((rc = up((void**)&c->d_tc_tw, tw.data(), tw.size() * sizeof(float2)))) return rc;
  if ((rc = up((void**)&c->d_tc_rot, rot.data(), rot.size() * sizeof(float2)))) return rc;
  if ((rc = up((void**)&c->d_tc_tasks, tasks.data(), tasks.size() * 4))) return rc;
  c->tc_lm_tasks = (int)tasks.size();
  c->tc_lm_ready = 1;
  return WW_OK;
}

// Returns WW_OK when the tensor-core kernel was launched, 1 when this call is outside its domain (the caller then uses
// the shared-memory FFT kernel of logmel.cu), a negative error code otherwise.
int ww_launch_logmel_tc(ww_ctx* c, const void* clips, int pcm16, int64_t clip_stride, LogmelOut out, int B, int normalize,
                        cudaStream_t st) {
  const char* force = getenv("WW_LOGMEL_KERNEL");                   // "fft" / "tc": read per call (tests and A/B runs switch it)
  if (force ? force[0] == 'f' : !kTcDefault) return 1;
  const int align = pcm16 ? 8 : 4;                                  // 16-byte rows
  if ((clip_stride % align) != 0 || (reinterpret_cast<uintptr_t>(clips) & 15) != 0) return 1;
  int rc = logmel_tc_prepare(c);
  if (rc) return rc;
  TcLogmelParams p;
  p.clips = clips; p.clip_stride = clip_stride; p.out = out; p.B = B; p.normalize = normalize;
  p.n_mels = c->cfg.n_mels; p.mel_nnz = c->mel_nnz; p.n_tasks = c->tc_lm_tasks;
  p.f32 = c->d_tc_f32; p.f64hi = c->d_tc_f64hi; p.f64lo = c->d_tc_f64lo; p.tw = c->d_tc_tw; p.rot = c->d_tc_rot;
  p.mel_w = c->d_mel_w; p.mel_start = c->d_mel_start; p.mel_len = c->d_mel_len; p.mel_off = c->d_mel_off;
  p.tasks = c->d_tc_tasks;
  static long long* d_trace = nullptr;
  const bool tracing = getenv("WW_TC_TRACE") != nullptr;
  if (tracing && !d_trace) cudaMalloc((void**)&d_trace, 40 * 16 * 8);
  if (tracing) cudaMemset(d_trace, 0, 40 * 16 * 8);
  p.trace = tracing ? d_trace : nullptr;
  const int grid = std::min(c->sm_count, B);
  ProfScope prof(c, WW_STAGE_LOGMEL, st);
  if (pcm16) {
    WW_CHECK(c, cudaFuncSetAttribute(logmel_tc_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    logmel_tc_kernel<int16_t><<<grid, kThreads, kSmem, st>>>(p);
  } else {
    WW_CHECK(c, cudaFuncSetAttribute(logmel_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    logmel_tc_kernel<float><<<grid, kThreads, kSmem, st>>>(p);
  }
  WW_LAUNCH_CHECK(c);
  if (tracing) {
    long long h[40 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "logmel_tc trace (cycles): S1 ready | S3: y_full, d3_empty | ep1 w4: d1_full, at wait, y_empty, stored | ep3: d3_full, tile written w12..15, after barrier, mel done, after barrier | ep1 w11 stored\n");
    const long long t0 = h[0];
    for (int i = 0; i < 40; ++i) {
      fprintf(stderr, "grp %2d:", i);
      for (int k = 0; k < 16; ++k) fprintf(stderr, " %7lld", h[i * 16 + k] ? h[i * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return WW_OK;
}

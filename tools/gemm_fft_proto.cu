// Correctness prototype of the tensor-core GEMM-FFT log-mel front end: ONE CTA, no pipelining,
// one clip -> Hann-windowed STFT power [32 frames][1025 bins], compared on the host with a float64 DFT.
// It exercises every numerical step of the planned kernel on real tcgen05 hardware:
//   * the clip as ONE fp16 hi / lo copy (x 2^12) in the tiled layout, frames as overlapping MN-major views (stage 1 A operand)
//   * stage 1: 32-point DFT over n2 as x_hi x [F32_hi | F32_lo] (N = 128) + x_lo x F32_hi (N = 64), K = 32
//   * epilogue 1: hi + lo columns, twiddle W2048^(n1 k2), Hann as the 3-tap filter over k2 (with the k1 carry), 1/32, hi/lo split,
//     128-bit stores into the MN-major stage-3 operand
//   * stage 3: 64-point DFT over n1 as [Yr | Yi] x [[Fr, -Fi], [Fi, Fr]]^T, three operand products, K = 3 x 128
//   * epilogue 3: re / im of a bin in the same thread -> power
// This is a tool for the next round (not linked into the library, not a product path).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/gemm_fft_proto tools/gemm_fft_proto.cu
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../wakeword_jupyterlab_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int kNfft = 2048, kN1 = 64, kN2 = 32, kW = 32, kSamples = 16000;
constexpr int kRows = 288;                          // padded clip: 18,048 samples = 282 rows of 64, whole row blocks
constexpr int kClipBytes = kRows * 64 * 2;          // one fp16 copy, tiled [row block][col chunk 8][row%8][8]
constexpr int kF32Bytes = 4 * 128 * 16;             // stage-1 B: [kc 4][n 128 = re hi | im hi | re lo | im lo (32 each)][8]
constexpr int kYBytes = 16 * 2048;                  // stage-3 A (one of hi / lo): [m chunk 16][k row 128][8]  (MN-major)
constexpr int kF64Bytes = 16 * 128 * 16;            // stage-3 B (one of hi / lo): [kc 16][n 128 = re k1 | im k1][8]
constexpr float kXScale = 4096.0f, kFScale = 1024.0f, kYScale = 4096.0f;
constexpr size_t kSmem = 2 * kClipBytes + kF32Bytes + 2 * kYBytes + 2 * kF64Bytes + 64;

__host__ __device__ constexpr uint32_t idesc_a_mn(int M, int N) { return make_idesc(M, N) | (1u << 15); }

struct Params {
  const float* clip;             // [16000]
  const __half* f32;             // pre-tiled stage-1 B
  const __half* f64hi;           // pre-tiled stage-3 B (hi, lo)
  const __half* f64lo;
  const float2* tw;              // [64][32] exp(-2 pi i n1 k2 / 2048)
  const float2* rot;             // [64]     exp(-2 pi i n1 / 64)
  float* power;                  // [32][1025]
};

__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__global__ void __launch_bounds__(128, 1) proto(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* xhi = smem;
  unsigned char* xlo = xhi + kClipBytes;
  unsigned char* f32 = xlo + kClipBytes;
  unsigned char* yhi = f32 + kF32Bytes;
  unsigned char* ylo = yhi + kYBytes;
  unsigned char* f64hi = ylo + kYBytes;
  unsigned char* f64lo = f64hi + kF64Bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(f64lo + kF64Bytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- prologue: tables, and the clip as fp16 hi / lo (x 2^12), zero centre padding of 1024 samples on both sides
  for (int i = tid * 16; i < kF32Bytes; i += 128 * 16) *reinterpret_cast<uint4*>(f32 + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f32) + i);
  for (int i = tid * 16; i < kF64Bytes; i += 128 * 16) {
    *reinterpret_cast<uint4*>(f64hi + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64hi) + i);
    *reinterpret_cast<uint4*>(f64lo + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64lo) + i);
  }
  for (int i = tid; i < kRows * 64; i += 128) {
    const int s = i - kNfft / 2;
    const float v = (s >= 0 && s < kSamples) ? p.clip[s] * kXScale : 0.0f;
    __half h, l;
    split16(v, h, l);
    const int r = i >> 6, c = i & 63;
    const int e = (((r >> 3) * 8 + (c >> 3)) * 8 + (r & 7)) * 8 + (c & 7);
    reinterpret_cast<__half*>(xhi)[e] = h;
    reinterpret_cast<__half*>(xlo)[e] = l;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(slot, 256);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t d1 = tm, d3 = tm + 128;
  const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;

  for (int g = 0; g < kW / 4; ++g) {
    for (int h = 0; h < 2; ++h) {
      const int t0 = 4 * g + 2 * h;
      if (tid == 0) {
        const uint64_t bd = make_desc(smem_u32(f32), 128 * 16, 128);
        const uint64_t ah = make_desc(smem_u32(xhi) + (uint32_t)t0 * 1024u, 1024, 128);
        const uint64_t al = make_desc(smem_u32(xlo) + (uint32_t)t0 * 1024u, 1024, 128);
        for (int s = 0; s < 2; ++s)       // x_hi x [F_hi | F_lo]: all 128 columns
          umma_f16(d1, ah + (uint64_t)((s * 2 * 1024) >> 4), bd + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 128), s);
        for (int s = 0; s < 2; ++s)       // x_lo x F_hi: accumulates into columns 0 .. 63
          umma_f16(d1, al + (uint64_t)((s * 2 * 1024) >> 4), bd + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 64), 1);
        umma_commit(bar);
      }
      mbar_wait(bar, ph & 1, 1); ++ph;
      tc_fence_after();
      // ---- epilogue 1: this thread = (frame t0 + f, n1)
      const int f = tid >> 6, n1 = tid & 63;
      float2 y[kN2];
      {
        uint32_t c0[32], c1[32], c2[32], c3[32];
        tmem_ld32_nowait(d1 + lane_addr, c0);
        tmem_ld32_nowait(d1 + lane_addr + 32, c1);
        tmem_ld32_nowait(d1 + lane_addr + 64, c2);
        tmem_ld32_nowait(d1 + lane_addr + 96, c3);
        tmem_ld_wait();
        const float inv = 1.0f / (kXScale * kFScale);
#pragma unroll
        for (int k2 = 0; k2 < kN2; ++k2) {
          const float re = (__uint_as_float(c0[k2]) + __uint_as_float(c2[k2])) * inv;
          const float im = (__uint_as_float(c1[k2]) + __uint_as_float(c3[k2])) * inv;
          const float2 w = __ldg(p.tw + n1 * kN2 + k2);
          y[k2] = make_float2(re * w.x - im * w.y, re * w.y + im * w.x);
        }
      }
      const float2 rp = __ldg(p.rot + n1);                             // W64^(n1)
      const float2 ym1 = make_float2(y[31].x * rp.x + y[31].y * rp.y, y[31].y * rp.x - y[31].x * rp.y);   // y[31] * conj(rp)
      const float2 y32 = make_float2(y[0].x * rp.x - y[0].y * rp.y, y[0].x * rp.y + y[0].y * rp.x);       // y[0] * rp
      const int fslot = 2 * h + f;
      __align__(16) __half zr_h[kN2], zr_l[kN2], zi_h[kN2], zi_l[kN2];
#pragma unroll
      for (int k2 = 0; k2 < kN2; ++k2) {
        const float2 l = k2 == 0 ? ym1 : y[k2 - 1], r = k2 == kN2 - 1 ? y32 : y[k2 + 1];
        const float zr = (0.5f * y[k2].x - 0.25f * (l.x + r.x)) * (kYScale / 32.0f);
        const float zi = (0.5f * y[k2].y - 0.25f * (l.y + r.y)) * (kYScale / 32.0f);
        split16(zr, zr_h[k2], zr_l[k2]);
        split16(zi, zi_h[k2], zi_l[k2]);
      }
      // stage-3 A operand, MN-major: element (m = fslot * 32 + k2, k) at (m / 8) * 2048 + (k / 8) * 128 + (k % 8) * 16 + (m % 8) * 2
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mc = fslot * 4 + q;
        const int kr = n1, ki = 64 + n1;
        const uint32_t off_r = mc * 2048 + (kr >> 3) * 128 + (kr & 7) * 16, off_i = mc * 2048 + (ki >> 3) * 128 + (ki & 7) * 16;
        *reinterpret_cast<uint4*>(yhi + off_r) = *reinterpret_cast<const uint4*>(&zr_h[8 * q]);
        *reinterpret_cast<uint4*>(ylo + off_r) = *reinterpret_cast<const uint4*>(&zr_l[8 * q]);
        *reinterpret_cast<uint4*>(yhi + off_i) = *reinterpret_cast<const uint4*>(&zi_h[8 * q]);
        *reinterpret_cast<uint4*>(ylo + off_i) = *reinterpret_cast<const uint4*>(&zi_l[8 * q]);
      }
      fence_proxy_async();
      tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    // ---- stage 3: D3[(fslot, k2), (re k1 | im k1)] = Y' x F64^T, three operand products, K = 128 each
    if (tid == 0) {
      const uint64_t ah = make_desc(smem_u32(yhi), /*lbo: K cores*/ 128, /*sbo: M cores*/ 2048);
      const uint64_t al = make_desc(smem_u32(ylo), 128, 2048);
      const uint64_t bh = make_desc(smem_u32(f64hi), 128 * 16, 128);
      const uint64_t bl = make_desc(smem_u32(f64lo), 128 * 16, 128);
      const uint32_t id = idesc_a_mn(128, 128);
      for (int prod = 0; prod < 3; ++prod) {
        const uint64_t a = prod == 1 ? al : ah, b = prod == 2 ? bl : bh;
        for (int s = 0; s < 8; ++s)
          umma_f16(d3, a + (uint64_t)((s * 2 * 128) >> 4), b + (uint64_t)((s * 2 * 128 * 16) >> 4), id, (prod | s) != 0);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, ph & 1, 2); ++ph;
    tc_fence_after();
    {
      // ---- epilogue 3: this thread = (frame 4g + fslot, k2); columns k1 (re) and 64 + k1 (im)
      const int fslot = tid >> 5, k2 = tid & 31;
      uint32_t cr[32], ci[32], cr2[32], ci2[32];
      tmem_ld32_nowait(d3 + lane_addr, cr);
      tmem_ld32_nowait(d3 + lane_addr + 32, cr2);
      tmem_ld32_nowait(d3 + lane_addr + 64, ci);
      tmem_ld32_nowait(d3 + lane_addr + 96, ci2);
      tmem_ld_wait();
      const float inv = 32.0f / (kYScale * kFScale);       // epilogue 1 already removed the x and F32 scales
      float* out = p.power + (size_t)(4 * g + fslot) * (kNfft / 2 + 1);
#pragma unroll
      for (int k1 = 0; k1 <= 32; ++k1) {
        const int bin = 32 * k1 + k2;
        if (bin <= kNfft / 2) {
          const float re = __uint_as_float(k1 < 32 ? cr[k1] : cr2[0]) * inv, im = __uint_as_float(k1 < 32 ? ci[k1] : ci2[0]) * inv;
          out[bin] = re * re + im * im;
        }
      }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
  }
  if (tid < 32) tmem_dealloc(tm, 256);
}

static void split_host(double v, __half& hi, __half& lo) {
  hi = __float2half_rn((float)v);
  lo = __float2half_rn((float)(v - (double)__half2float(hi)));
}

int main() {
  const double PI = 3.14159265358979323846;
  // clip: tone + noise, peak-normalised (the reference's recipe shape)
  std::vector<float> clip(kSamples);
  srand(7);
  double peak = 0;
  for (int i = 0; i < kSamples; ++i) {
    const double t = i / 16000.0, u = (rand() + 0.5) / (RAND_MAX + 1.0), v = (rand() + 0.5) / (RAND_MAX + 1.0);
    const double nz = sqrt(-2 * log(u)) * cos(2 * PI * v);
    clip[i] = (float)(0.1 * nz + 0.3 * sin(2 * PI * 200 * t) + 0.2 * sin(2 * PI * 400 * t));
    peak = fmax(peak, fabs(clip[i]));
  }
  for (auto& c : clip) c = (float)(c / peak);
  // stage-1 B: [kc][n 128][8]: n = re hi (32) | im hi | re lo | im lo of F32[k2][n2] = exp(-2 pi i k2 n2 / 32), x 2^10
  std::vector<__half> f32(kF32Bytes / 2);
  for (int k2 = 0; k2 < kN2; ++k2)
    for (int n2 = 0; n2 < kN2; ++n2) {
      const double a = -2 * PI * ((k2 * n2) % kN2) / kN2;
      __half rh, rl, ih, il;
      split_host(cos(a) * kFScale, rh, rl);
      split_host(sin(a) * kFScale, ih, il);
      auto at = [&](int n) -> __half& { return f32[((n2 / 8) * 128 + n) * 8 + n2 % 8]; };
      at(k2) = rh; at(32 + k2) = ih; at(64 + k2) = rl; at(96 + k2) = il;
    }
  // stage-3 B: [kc 16][n 128][8]: n = k1 (re out) | 64 + k1 (im out); k = n1 (x Yr) | 64 + n1 (x Yi)
  std::vector<__half> f64h(kF64Bytes / 2), f64l(kF64Bytes / 2);
  for (int k1 = 0; k1 < kN1; ++k1)
    for (int n1 = 0; n1 < kN1; ++n1) {
      const double a = -2 * PI * ((k1 * n1) % kN1) / kN1, fr = cos(a) * kFScale, fi = sin(a) * kFScale;
      auto put = [&](int n, int k, double v) {
        __half h, l;
        split_host(v, h, l);
        const size_t e = ((size_t)(k / 8) * 128 + n) * 8 + k % 8;
        f64h[e] = h; f64l[e] = l;
      };
      put(k1, n1, fr); put(k1, 64 + n1, -fi);          // X_re = Fr Yr - Fi Yi
      put(64 + k1, n1, fi); put(64 + k1, 64 + n1, fr); // X_im = Fi Yr + Fr Yi
    }
  std::vector<float2> tw(kN1 * kN2), rot(kN1);
  for (int n1 = 0; n1 < kN1; ++n1) {
    for (int k2 = 0; k2 < kN2; ++k2) {
      const double a = -2 * PI * (n1 * k2) / kNfft;
      tw[n1 * kN2 + k2] = make_float2((float)cos(a), (float)sin(a));
    }
    rot[n1] = make_float2((float)cos(-2 * PI * n1 / kN1), (float)sin(-2 * PI * n1 / kN1));
  }
  Params p;
  float* dclip; __half *df32, *df64h, *df64l; float2 *dtw, *drot; float* dpow;
  cudaMalloc(&dclip, kSamples * 4); cudaMalloc(&df32, kF32Bytes); cudaMalloc(&df64h, kF64Bytes); cudaMalloc(&df64l, kF64Bytes);
  cudaMalloc(&dtw, tw.size() * 8); cudaMalloc(&drot, rot.size() * 8); cudaMalloc(&dpow, kW * (kNfft / 2 + 1) * 4);
  cudaMemcpy(dclip, clip.data(), kSamples * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(df32, f32.data(), kF32Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(df64h, f64h.data(), kF64Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(df64l, f64l.data(), kF64Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(drot, rot.data(), rot.size() * 8, cudaMemcpyHostToDevice);
  p.clip = dclip; p.f32 = df32; p.f64hi = df64h; p.f64lo = df64l; p.tw = dtw; p.rot = drot; p.power = dpow;
  cudaFuncSetAttribute(proto, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  proto<<<1, 128, kSmem>>>(p);
  std::vector<float> got(kW * (kNfft / 2 + 1));
  cudaError_t e = cudaMemcpy(got.data(), dpow, got.size() * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  // float64 reference: Hann-windowed DFT power of every frame
  std::vector<double> win(kNfft), xp(kRows * 64, 0.0);
  for (int n = 0; n < kNfft; ++n) win[n] = 0.5 - 0.5 * cos(2 * PI * n / kNfft);
  for (int i = 0; i < kSamples; ++i) xp[kNfft / 2 + i] = clip[i];
  double worst_rel_peak = 0, worst_rel_bin = 0, pmax_all = 0;
  for (int t : {0, 1, 2, 7, 16, 30, 31}) {
    std::vector<double> ref(kNfft / 2 + 1);
    double pmax = 0;
    for (int k = 0; k <= kNfft / 2; ++k) {
      double re = 0, im = 0;
      for (int n = 0; n < kNfft; ++n) {
        const double a = -2 * PI * ((long long)k * n % kNfft) / kNfft, v = win[n] * xp[512 * t + n];
        re += v * cos(a); im += v * sin(a);
      }
      ref[k] = re * re + im * im;
      pmax = fmax(pmax, ref[k]);
    }
    double wp = 0, wb = 0;
    for (int k = 0; k <= kNfft / 2; ++k) {
      const double d = fabs(got[(size_t)t * (kNfft / 2 + 1) + k] - ref[k]);
      wp = fmax(wp, d / pmax);
      if (ref[k] > 1e-6 * pmax) wb = fmax(wb, d / ref[k]);
    }
    printf("frame %2d: max |dP| / max P = %.2e, max |dP| / P over bins above -60 dB = %.2e  (1e-3 dB = 2.3e-4)\n", t, wp, wb);
    worst_rel_peak = fmax(worst_rel_peak, wp); worst_rel_bin = fmax(worst_rel_bin, wb); pmax_all = fmax(pmax_all, pmax);
  }
  const bool ok = worst_rel_bin < 2.3e-4;
  printf("%s: worst per-bin relative power error %.2e (peak-relative %.2e)\n", ok ? "PASS" : "FAIL", worst_rel_bin, worst_rel_peak);
  return ok ? 0 : 2;
}

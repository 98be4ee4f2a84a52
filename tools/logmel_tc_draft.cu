// DRAFT (compiles, NOT yet run on a GPU) of the pipelined tensor-core STFT core of DESIGN.md section 8.1: clips -> Hann-windowed
// STFT power [32 frames][1025 bins], one persistent warp-specialised CTA per SM.  The numerics, layouts and descriptors are those
// of tools/gemm_fft_proto.cu (validated on a B200); what is new and untested here is the pipelining:
//   warp 0       MMA issuer: stage 1 of group g+1 is issued before stage 3 of group g, so the tensor pipe has work while
//                epilogue 1 of group g is still splitting its Y'
//   warps 1-4    epilogue 1 (one per TMEM lane quadrant): D1 -> twiddle, Hann 3-tap, hi/lo split -> Y' (single buffer: only the
//                final stores wait for the stage-3 MMAs of the previous group)
//   warps 5-8    epilogue 3: D3 -> power -> global (the production kernel feeds the mel projection here)
//   warps 9-12   converter: clip (fp32) -> fp16 hi / lo copies in the tiled layout (whole clip; a row-block ring is the next step)
// The harness times a batch and checks a few clips against a float64 DFT.  Next round: build, run, fix, then move into
// wakeword_jupyterlab_b200/csrc/logmel.cu behind the existing ww_launch_logmel interface.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/logmel_tc_draft tools/logmel_tc_draft.cu
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../wakeword_jupyterlab_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int kNfft = 2048, kN1 = 64, kN2 = 32, kW = 32, kSamples = 16000, kBins = kNfft / 2 + 1;
constexpr int kRows = 288;
constexpr int kClipBytes = kRows * 64 * 2;          // one fp16 copy of the padded clip, tiled [row block][col chunk 8][row%8][8]
constexpr int kF32Bytes = 4 * 128 * 16;             // stage-1 B: [kc 4][n 128 = re hi | im hi | re lo | im lo][8]
constexpr int kN3 = 80;                             // stage-3 N: re of k1 0..39 | im of k1 0..39 (k1 <= 32 used)
constexpr int kF64Bytes = 16 * kN3 * 16;            // stage-3 B (hi or lo): [kc 16][n 80][8]
constexpr int kYBytes = 16 * 2048;                  // stage-3 A (hi or lo), MN-major: [m chunk 16][k row 128][8]
constexpr float kXScale = 4096.0f, kFScale = 1024.0f, kYScale = 4096.0f;
constexpr int kThreads = 13 * 32;
constexpr size_t kSmem = 2 * kClipBytes + kF32Bytes + 2 * kF64Bytes + 2 * kYBytes + 256;

__host__ __device__ constexpr uint32_t idesc_a_mn(int M, int N) { return make_idesc(M, N) | (1u << 15); }

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

struct Params {
  const float* clips;            // [B][16000]
  int B;
  const __half* f32;             // pre-tiled stage-1 B
  const __half* f64hi;           // pre-tiled stage-3 B
  const __half* f64lo;
  const float2* tw;              // [64][32] exp(-2 pi i n1 k2 / 2048)
  const float2* rot;             // [64]     exp(-2 pi i n1 / 64)
  float* power;                  // [B][32][1025]
};

__global__ void __launch_bounds__(kThreads, 1) stft_tc_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* xhi = smem;
  unsigned char* xlo = xhi + kClipBytes;
  unsigned char* f32 = xlo + kClipBytes;
  unsigned char* f64hi = f32 + kF32Bytes;
  unsigned char* f64lo = f64hi + kF64Bytes;
  unsigned char* yhi = f64lo + kF64Bytes;
  unsigned char* ylo = yhi + kYBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ylo + kYBytes);
  uint64_t* x_full = bars;            // [1] converter warps -> issuer
  uint64_t* x_empty = bars + 1;       // [1] stage-1 MMAs of the clip done -> converter
  uint64_t* d1_full = bars + 2;       // [2]
  uint64_t* d1_empty = bars + 4;      // [2]
  uint64_t* y_full = bars + 6;        // [1] 4 warps x 2 pairs
  uint64_t* y_empty = bars + 7;       // [1] stage-3 MMAs done
  uint64_t* d3_full = bars + 8;       // [2]
  uint64_t* d3_empty = bars + 10;     // [2]
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  for (int i = tid * 16; i < kF32Bytes; i += kThreads * 16)
    *reinterpret_cast<uint4*>(f32 + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f32) + i);
  for (int i = tid * 16; i < kF64Bytes; i += kThreads * 16) {
    *reinterpret_cast<uint4*>(f64hi + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64hi) + i);
    *reinterpret_cast<uint4*>(f64lo + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.f64lo) + i);
  }
  // centre padding (rows 0-15 and 266-287) is zero for every clip: written once
  for (int i = tid * 16; i < kClipBytes; i += kThreads * 16) {
    *reinterpret_cast<uint4*>(xhi + i) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(xlo + i) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    mbar_init(x_full, 4); mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(d1_full + i, 1); mbar_init(d1_empty + i, 4); mbar_init(d3_full + i, 1); mbar_init(d3_empty + i, 4); }
    mbar_init(y_full, 8); mbar_init(y_empty, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);          // D1: columns 128 b (b < 2); D3: columns 256 + 128 b
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;

  if (warp == 0) {
    // ===================== MMA issuer
    const uint64_t b1 = make_desc(smem_u32(f32), 128 * 16, 128);
    const uint64_t b3h = make_desc(smem_u32(f64hi), kN3 * 16, 128), b3l = make_desc(smem_u32(f64lo), kN3 * 16, 128);
    const uint64_t a3h = make_desc(smem_u32(yhi), 128, 2048), a3l = make_desc(smem_u32(ylo), 128, 2048);
    uint32_t ci = 0;
    for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x, ++ci) {
      mbar_wait(x_full, ci & 1, 10);
      tc_fence_after();
      auto stage1 = [&](int g, int h) {
        const uint32_t pc = ci * 16 + 2 * g + h, buf = pc & 1;
        mbar_wait(d1_empty + buf, ((pc >> 1) & 1) ^ 1, 11);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t0 = 4 * g + 2 * h, d = tm + buf * 128;
          const uint64_t ah = make_desc(smem_u32(xhi) + t0 * 1024u, 1024, 128), al = make_desc(smem_u32(xlo) + t0 * 1024u, 1024, 128);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            umma_f16(d, ah + (uint64_t)((s * 2 * 1024) >> 4), b1 + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 128), s);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            umma_f16(d, al + (uint64_t)((s * 2 * 1024) >> 4), b1 + (uint64_t)((s * 2 * 128 * 16) >> 4), idesc_a_mn(128, 64), 1);
          umma_commit(d1_full + buf);
          if (g == kW / 4 - 1 && h == 1) umma_commit(x_empty);     // the clip copies may be overwritten
        }
        __syncwarp();
      };
      auto stage3 = [&](int g) {
        const uint32_t gc = ci * 8 + g, buf = gc & 1;
        mbar_wait(y_full, gc & 1, 12);
        mbar_wait(d3_empty + buf, ((gc >> 1) & 1) ^ 1, 13);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tm + 256 + buf * 128;
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
      stage1(0, 0); stage1(0, 1);
      for (int g = 0; g < kW / 4; ++g) {
        if (g + 1 < kW / 4) { stage1(g + 1, 0); stage1(g + 1, 1); }
        stage3(g);
      }
    }
  } else if (warp <= 4) {
    // ===================== epilogue 1: thread = (frame t0 + f, n1), lane quadrant = warp & 3
    const int q = warp & 3, m = q * 32 + lane, f = m >> 6, n1 = m & 63;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    const float2 rp = __ldg(p.rot + n1);
    uint32_t ci = 0;
    for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x, ++ci) {
      for (int g = 0; g < kW / 4; ++g) {
        const uint32_t gc = ci * 8 + g;
        for (int h = 0; h < 2; ++h) {
          const uint32_t pc = ci * 16 + 2 * g + h, buf = pc & 1;
          mbar_wait(d1_full + buf, (pc >> 1) & 1, 20);
          tc_fence_after();
          float2 y[kN2];
          const uint32_t d = tm + buf * 128 + lane_addr;
          const float inv = 1.0f / (kXScale * kFScale);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r0[8], r1[8], r2[8], r3[8];
            tmem_ld8_nowait(d + 8 * c, r0);
            tmem_ld8_nowait(d + 32 + 8 * c, r1);
            tmem_ld8_nowait(d + 64 + 8 * c, r2);
            tmem_ld8_nowait(d + 96 + 8 * c, r3);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int k2 = 8 * c + e;
              const float re = (__uint_as_float(r0[e]) + __uint_as_float(r2[e])) * inv;
              const float im = (__uint_as_float(r1[e]) + __uint_as_float(r3[e])) * inv;
              const float2 w = __ldg(p.tw + n1 * kN2 + k2);
              y[k2] = make_float2(re * w.x - im * w.y, re * w.y + im * w.x);
            }
          }
          tc_fence_before();
          mbar_arrive_warp(d1_empty + buf, lane);                  // D1 buffer may be overwritten by the next-but-one pair
          const float2 ym1 = make_float2(y[31].x * rp.x + y[31].y * rp.y, y[31].y * rp.x - y[31].x * rp.y);   // W64^(-n1) Y[31]
          const float2 y32 = make_float2(y[0].x * rp.x - y[0].y * rp.y, y[0].x * rp.y + y[0].y * rp.x);       // W64^(+n1) Y[0]
          if (h == 0) mbar_wait(y_empty, (gc & 1) ^ 1, 21);       // stage 3 of the previous group has read Y'
          const int fslot = 2 * h + f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            __align__(16) __half zr_h[8], zr_l[8], zi_h[8], zi_l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int k2 = 8 * c + e;
              const float2 l = k2 == 0 ? ym1 : y[k2 - 1], r = k2 == kN2 - 1 ? y32 : y[k2 + 1];
              split16((0.5f * y[k2].x - 0.25f * (l.x + r.x)) * (kYScale / 32.0f), zr_h[e], zr_l[e]);
              split16((0.5f * y[k2].y - 0.25f * (l.y + r.y)) * (kYScale / 32.0f), zi_h[e], zi_l[e]);
            }
            const int mc = fslot * 4 + c, kr = n1, ki = 64 + n1;
            const uint32_t off_r = mc * 2048 + (kr >> 3) * 128 + (kr & 7) * 16, off_i = mc * 2048 + (ki >> 3) * 128 + (ki & 7) * 16;
            *reinterpret_cast<uint4*>(yhi + off_r) = *reinterpret_cast<const uint4*>(zr_h);
            *reinterpret_cast<uint4*>(ylo + off_r) = *reinterpret_cast<const uint4*>(zr_l);
            *reinterpret_cast<uint4*>(yhi + off_i) = *reinterpret_cast<const uint4*>(zi_h);
            *reinterpret_cast<uint4*>(ylo + off_i) = *reinterpret_cast<const uint4*>(zi_l);
          }
          fence_proxy_async();
          mbar_arrive_warp(y_full, lane);
        }
      }
    }
  } else if (warp <= 8) {
    // ===================== epilogue 3: thread = (frame 4g + fslot, k2); columns k1 (re) and 40 + k1 (im)
    const int q = warp & 3, fslot = q, k2 = lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    uint32_t ci = 0;
    for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x, ++ci) {
      for (int g = 0; g < kW / 4; ++g) {
        const uint32_t gc = ci * 8 + g, buf = gc & 1;
        mbar_wait(d3_full + buf, (gc >> 1) & 1, 30);
        tc_fence_after();
        uint32_t cr[32], ci_[32], cr2[8], ci2[8];
        const uint32_t d = tm + 256 + buf * 128 + lane_addr;
        tmem_ld32_nowait(d, cr);
        tmem_ld8_nowait(d + 32, cr2);
        tmem_ld32_nowait(d + 40, ci_);
        tmem_ld8_nowait(d + 72, ci2);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(d3_empty + buf, lane);
        const float inv = 32.0f / (kYScale * kFScale);
        float* out = p.power + ((size_t)clip * kW + 4 * g + fslot) * kBins;
#pragma unroll
        for (int k1 = 0; k1 <= 32; ++k1) {
          const int bin = 32 * k1 + k2;
          if (bin < kBins) {
            const float re = __uint_as_float(k1 < 32 ? cr[k1] : cr2[0]) * inv, im = __uint_as_float(k1 < 32 ? ci_[k1] : ci2[0]) * inv;
            out[bin] = re * re + im * im;
          }
        }
      }
    }
  } else {
    // ===================== converter: 8 consecutive samples -> one 16-byte store into each copy
    const int ct = tid - 9 * 32;                                   // 0 .. 127
    uint32_t ci = 0;
    for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x, ++ci) {
      mbar_wait(x_empty, (ci & 1) ^ 1, 40);
      const float* x = p.clips + (size_t)clip * kSamples;
      for (int u = ct; u < kSamples / 8; u += 128) {               // unit = 8 samples: padded index 1024 + 8 u
        const int i = kNfft / 2 + 8 * u, r = i >> 6, c8 = (i & 63) >> 3;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + 8 * u)), v1 = __ldg(reinterpret_cast<const float4*>(x + 8 * u + 4));
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split16(v[e] * kXScale, hi[e], lo[e]);
        const uint32_t off = ((((r >> 3) * 8 + c8) * 8 + (r & 7)) * 8) * 2;
        *reinterpret_cast<uint4*>(xhi + off) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(xlo + off) = *reinterpret_cast<const uint4*>(lo);
      }
      fence_proxy_async();
      mbar_arrive_warp(x_full, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static void split_host(double v, __half& hi, __half& lo) {
  hi = __float2half_rn((float)v);
  lo = __float2half_rn((float)(v - (double)__half2float(hi)));
}

int main(int argc, char** argv) {
  const double PI = 3.14159265358979323846;
  const int B = argc > 1 ? atoi(argv[1]) : 4096;
  std::vector<float> clips((size_t)B * kSamples);
  srand(11);
  for (int b = 0; b < B; ++b) {
    double peak = 0;
    float* c = clips.data() + (size_t)b * kSamples;
    for (int i = 0; i < kSamples; ++i) {
      const double t = i / 16000.0, u = (rand() + 0.5) / (RAND_MAX + 1.0), v = (rand() + 0.5) / (RAND_MAX + 1.0);
      const double nz = sqrt(-2 * log(u)) * cos(2 * PI * v);
      c[i] = (float)((b % 3 == 0) ? 0.1 * nz + 0.3 * sin(2 * PI * 200 * t) + 0.2 * sin(2 * PI * 400 * t) : 0.2 * nz);
      peak = fmax(peak, fabs(c[i]));
    }
    for (int i = 0; i < kSamples; ++i) c[i] = (float)(c[i] / peak);
  }
  std::vector<__half> f32(kF32Bytes / 2), f64h(kF64Bytes / 2, __float2half(0.0f)), f64l(kF64Bytes / 2, __float2half(0.0f));
  for (int k2 = 0; k2 < kN2; ++k2)
    for (int n2 = 0; n2 < kN2; ++n2) {
      const double a = -2 * PI * ((k2 * n2) % kN2) / kN2;
      __half rh, rl, ih, il;
      split_host(cos(a) * kFScale, rh, rl);
      split_host(sin(a) * kFScale, ih, il);
      auto at = [&](int n) -> __half& { return f32[((n2 / 8) * 128 + n) * 8 + n2 % 8]; };
      at(k2) = rh; at(32 + k2) = ih; at(64 + k2) = rl; at(96 + k2) = il;
    }
  for (int k1 = 0; k1 <= 32; ++k1)
    for (int n1 = 0; n1 < kN1; ++n1) {
      const double a = -2 * PI * ((k1 * n1) % kN1) / kN1, fr = cos(a) * kFScale, fi = sin(a) * kFScale;
      auto put = [&](int n, int k, double v) {
        __half h, l;
        split_host(v, h, l);
        const size_t e = ((size_t)(k / 8) * kN3 + n) * 8 + k % 8;
        f64h[e] = h; f64l[e] = l;
      };
      put(k1, n1, fr); put(k1, 64 + n1, -fi);
      put(40 + k1, n1, fi); put(40 + k1, 64 + n1, fr);
    }
  std::vector<float2> tw(kN1 * kN2), rot(kN1);
  for (int n1 = 0; n1 < kN1; ++n1) {
    for (int k2 = 0; k2 < kN2; ++k2) tw[n1 * kN2 + k2] = make_float2((float)cos(-2 * PI * n1 * k2 / kNfft), (float)sin(-2 * PI * n1 * k2 / kNfft));
    rot[n1] = make_float2((float)cos(-2 * PI * n1 / kN1), (float)sin(-2 * PI * n1 / kN1));
  }
  Params p;
  float* dclips; __half *df32, *df64h, *df64l; float2 *dtw, *drot; float* dpow;
  cudaMalloc(&dclips, clips.size() * 4); cudaMalloc(&df32, kF32Bytes); cudaMalloc(&df64h, kF64Bytes); cudaMalloc(&df64l, kF64Bytes);
  cudaMalloc(&dtw, tw.size() * 8); cudaMalloc(&drot, rot.size() * 8); cudaMalloc(&dpow, (size_t)B * kW * kBins * 4);
  cudaMemcpy(dclips, clips.data(), clips.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(df32, f32.data(), kF32Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(df64h, f64h.data(), kF64Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(df64l, f64l.data(), kF64Bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(drot, rot.data(), rot.size() * 8, cudaMemcpyHostToDevice);
  p.clips = dclips; p.B = B; p.f32 = df32; p.f64hi = df64h; p.f64lo = df64l; p.tw = dtw; p.rot = drot; p.power = dpow;
  cudaFuncSetAttribute(stft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = B < sms ? B : sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  stft_tc_kernel<<<grid, kThreads, kSmem>>>(p);                    // warm-up
  cudaEventRecord(e0);
  for (int it = 0; it < 3; ++it) stft_tc_kernel<<<grid, kThreads, kSmem>>>(p);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%d clips: %.3f ms per launch = %.2f M clips/s (power spectra only; the FFT kernel's log-mel runs at 7.8 M clips/s)\n", B, ms / 3,
         B / (ms / 3) / 1e3);
  // ---- check a few (clip, frame) pairs against a float64 Hann-windowed DFT
  std::vector<double> win(kNfft);
  for (int n = 0; n < kNfft; ++n) win[n] = 0.5 - 0.5 * cos(2 * PI * n / kNfft);
  double worst = 0;
  const int check_clips[3] = {0, B / 2, B - 1}, check_frames[4] = {0, 1, 17, 31};
  std::vector<float> got(kBins);
  for (int cc : check_clips)
    for (int t : check_frames) {
      cudaMemcpy(got.data(), dpow + ((size_t)cc * kW + t) * kBins, kBins * 4, cudaMemcpyDeviceToHost);
      std::vector<double> xp(kRows * 64, 0.0), ref(kBins);
      for (int i = 0; i < kSamples; ++i) xp[kNfft / 2 + i] = clips[(size_t)cc * kSamples + i];
      double pmax = 0;
      for (int k = 0; k < kBins; ++k) {
        double re = 0, im = 0;
        for (int n = 0; n < kNfft; ++n) {
          const double a = -2 * PI * ((long long)k * n % kNfft) / kNfft, v = win[n] * xp[512 * t + n];
          re += v * cos(a); im += v * sin(a);
        }
        ref[k] = re * re + im * im;
        pmax = fmax(pmax, ref[k]);
      }
      double wb = 0;
      for (int k = 0; k < kBins; ++k)
        if (ref[k] > 1e-6 * pmax) wb = fmax(wb, fabs(got[k] - ref[k]) / ref[k]);
      printf("clip %5d frame %2d: max |dP| / P over bins above -60 dB = %.2e\n", cc, t, wb);
      worst = fmax(worst, wb);
    }
  printf("%s: worst per-bin relative power error %.2e (1e-3 dB = 2.3e-4)\n", worst < 2.3e-4 ? "PASS" : "FAIL", worst);
  return worst < 2.3e-4 ? 0 : 2;
}

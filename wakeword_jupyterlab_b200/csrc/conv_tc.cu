// K3 (tensor-core path): conv1 -> conv2 -> conv3 -> ReLU -> global mean on tcgen05 / TMEM  (sm_100a)
//
// Replaces the cuDNN convolutions behind WakewordModel.forward
// (/root/reference/wakeword_training_script.py:170-173).  conv2 (32->64) and conv3 (64->128) are
// implicit GEMMs  D[pixel, cout] += A[pixel + tap_offset, cin] * W[cout, (tap, cin)]  issued as
// tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM):
//
//   * Activations live in a zero-padded, pixel-linear image: pixel (y, x) sits at padded index
//     p = (y+1)*P + (x+1) with pitch P = W+1 (the right pad of row y is the left pad of row y+1),
//     stored channel-chunk-major:  [chunk of 8 channels][pixel][8 x bf16 = 16 B].  That is exactly the
//     UMMA K-major SWIZZLE_NONE canonical layout (core matrix = 8 rows x 16 B, contiguous), with
//     SBO = 128 B (next 8 pixels) and LBO = the chunk-plane stride.  A 3x3 tap is then nothing but a
//     different start address of the SAME shared-memory tile:  start += ((ky-1)*P + (kx-1)) * 16 B.
//     No im2col is ever materialised; every activation byte is loaded into shared memory once per
//     tile group and read by the tensor core 9 (taps) x 3 (passes) times.
//   * fp32 parity (logits within 1e-4 relative): every operand is split bf16 hi + bf16 lo
//     (x = hi + lo + O(2^-17 x)); the product is accumulated as hi*hi + hi*lo + lo*hi in fp32
//     (3 MMA passes; SURVEY.md section 7 hard part 1).  WW_CONV_BF16 issues only hi*hi.
//   * conv12 kernel: conv1 (K = 9, CUDA cores, fp32) is computed by 4 producer warps straight into the
//     shared-memory A tile of conv2 (with halo), overlapped with the MMAs of the previous tile;
//     conv2 weights (74 KB split) stay resident in shared memory; the epilogue applies bias + ReLU,
//     zeroes the padding pixels, splits to hi/lo and writes the conv3 operand layout to HBM
//     (coalesced 512 B per warp store).
//   * conv3 kernel: G pixel tiles (G*128 pixels) of one clip share every weight stage; activations
//     are streamed by 1-D bulk copies (cp.async.bulk, mbarrier complete_tx) per 16-channel k-slice
//     ring, weights by an 8-stage ring of 8 KB (k-slice, tap) blocks; G accumulators (G*128 TMEM
//     columns) are drained by 4 epilogue warps: bias + ReLU + padding mask + sum over pixels
//     (register transpose-reduce), written as one deterministic partial per (clip, group).
//
// Warp roles are mbarrier-synchronised producer / MMA-issuer / epilogue; one thread issues MMAs.
#include "ctx.cuh"

#include <algorithm>

namespace {

// ------------------------------------------------------------------------------------------------
// PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails loudly) instead of hanging the GPU.
__device__ __noinline__ void mbar_timeout(int code) {
  printf("conv_tc watchdog: wait %d timed out (block %d, thread %d)\n", code, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) mbar_timeout(code);
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (stride between the two 8-element K chunks), [32,46) SBO>>4 (stride between
// 8-row groups), [46,48) version = 1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// split 8 floats into bf16 hi (round-to-nearest) and bf16 lo = bf16(x - hi)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const float r0 = v[2 * i] - __bfloat162float(h0), r1 = v[2 * i + 1] - __bfloat162float(h1);
    __nv_bfloat162 hh;
    hh.x = h0; hh.y = h1;
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = pack_bf16(r0, r1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Geom {
  int H, W, P;        // image rows (mels), cols (frames), pitch P = W + 1
  int npix;           // plane slots per clip (multiple of 128); slot s <-> padded index p = s - 1
  int T2, T3;         // conv2 tiles (npix / 128), conv3 tiles
  int G, n_groups;    // conv3: tiles per group, groups per clip
  int nsl2;           // conv12 A-tile slots   = round8(128 + 2P + 2)
  int nsl3;           // conv3 A-plane slots   = round8(G*128 + 2P + 2)
};

__device__ __forceinline__ bool pix_valid(int p, const Geom& g, int& y, int& x) {
  if (p < 0) return false;
  const int row = p / g.P;
  y = row - 1;
  x = p - row * g.P - 1;
  return (y >= 0) && (y < g.H) && (x >= 0);
}

// ------------------------------------------------------------------------------------------------
// conv1 + conv2
constexpr int C12_THREADS = 288;     // warps 0-3 producers, warp 4 MMA, warps 5-8 epilogue
constexpr int W2_BYTES = 9 * 2 * 4 * 64 * 16;   // [tap][hl][kc 4][n 64][8 bf16] = 73,728

struct Conv12Params {
  const float* logmel;            // [B][H][W]
  const float* w1t;               // [9][32]
  const float* b1;                // [32]
  const __nv_bfloat16* w2s;       // split, canonical layout
  const float* b2;                // [64]
  __nv_bfloat16* act2;            // [B][16 planes][npix][8]
  int B, npass;
  Geom g;
};

__global__ void __launch_bounds__(C12_THREADS, 1) conv12_kernel(Conv12Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  unsigned char* w2s = smem;                                   // W2_BYTES
  const uint32_t a_bytes = 8u * g.nsl2 * 16u;                  // one act1 buffer: 8 planes (kc*2+hl)
  unsigned char* a_buf0 = smem + W2_BYTES;
  float* w1s = reinterpret_cast<float*>(a_buf0 + 2 * a_bytes); // [9][32]
  float* b1s = w1s + 288;
  float* b2s = b1s + 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b2s + 64);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;      // [2]
  uint64_t* a_empty = bars + 3;     // [2]
  uint64_t* t_full = bars + 5;      // [2]
  uint64_t* t_empty = bars + 7;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 288; i += C12_THREADS) w1s[i] = p.w1t[i];
  if (tid < 32) b1s[tid] = p.b1[tid];
  if (tid < 64) b2s[tid] = p.b2[tid];
  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 128);
      mbar_init(a_empty + i, 1);
      mbar_init(t_full + i, 1);
      mbar_init(t_empty + i, 128);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.B * g.T2;
  const int NL = 128 + 2 * g.P + 2;

  if (warp < 4) {
    // ===================== conv1 producers: fp32 conv1 + ReLU -> bf16 hi/lo A tile (with halo)
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / g.T2, t2 = item - b * g.T2;
      const int buf = it & 1;
      mbar_wait(a_empty + buf, ((it >> 1) & 1) ^ 1, 10);
      unsigned char* ab = a_buf0 + buf * a_bytes;
      const float* __restrict__ img = p.logmel + (size_t)b * g.H * g.W;
      const int pbase = 128 * t2 - 1 - g.P - 1;
      for (int l = tid; l < NL; l += 128) {
        int y, x;
        float v[32];
        if (pix_valid(pbase + l, g, y, x)) {
          float in[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
            in[k] = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? __ldg(img + yy * g.W + xx) : 0.0f;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = b1s[c];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 w = *reinterpret_cast<const float4*>(w1s + k * 32 + c4 * 4);
              v[c4 * 4 + 0] = fmaf(in[k], w.x, v[c4 * 4 + 0]);
              v[c4 * 4 + 1] = fmaf(in[k], w.y, v[c4 * 4 + 1]);
              v[c4 * 4 + 2] = fmaf(in[k], w.z, v[c4 * 4 + 2]);
              v[c4 * 4 + 3] = fmaf(in[k], w.w, v[c4 * 4 + 3]);
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.0f);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0.0f;
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4 hi, lo;
          split8(v + kc * 8, hi, lo);
          *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 0) * g.nsl2 + l) * 16) = hi;
          *reinterpret_cast<uint4*>(ab + ((size_t)(kc * 2 + 1) * g.nsl2 + l) * 16) = lo;
        }
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full + buf);
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (one thread)
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, W2_BYTES);
      bulk_g2s(w2s, p.w2s, W2_BYTES, w_full);
      mbar_wait(w_full, 0, 20);
      const uint32_t idesc = make_idesc(128, 64);
      const uint32_t w2_addr = smem_u32(w2s);
      const uint32_t lbo_a = 2u * g.nsl2 * 16u;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(a_full + buf, par, 21);
        mbar_wait(t_empty + buf, par ^ 1, 22);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_buf0 + buf * a_bytes);
        const uint32_t d = tmem_base + buf * 64;
        uint32_t acc = 0;
        for (int tap = 0; tap < 9; ++tap) {
          const int row_off = (g.P + 1) + (tap / 3 - 1) * g.P + (tap % 3 - 1);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            for (int ps = 0; ps < p.npass; ++ps) {
              const int hla = (ps == 2), hlb = (ps == 1);
              const uint64_t ad = make_desc(a_addr + ((uint32_t)((2 * j) * 2 + hla) * g.nsl2 + row_off) * 16u, lbo_a, 128);
              const uint64_t bd = make_desc(w2_addr + (uint32_t)(((tap * 2 + hlb) * 4 + 2 * j) * 1024), 1024, 128);
              umma_bf16(d, ad, bd, idesc, acc);
              acc = 1;
            }
          }
        }
        umma_commit(a_empty + buf);
        umma_commit(t_full + buf);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> bias + ReLU + pad mask -> hi/lo -> act2 (HBM)
    const int q = warp & 3;       // TMEM lane quadrant this warp may access
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / g.T2, t2 = item - b * g.T2;
      const int buf = it & 1;
      mbar_wait(t_full + buf, (it >> 1) & 1, 30);
      tc_fence_after();
      float v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64;
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tc_fence_before();
      mbar_arrive(t_empty + buf);
      const int s = 128 * t2 + q * 32 + lane;
      int y, x;
      const bool ok = pix_valid(s - 1, g, y, x);
      uint4* dst = reinterpret_cast<uint4*>(p.act2) + (size_t)b * 16 * g.npix + s;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ok ? fmaxf(v[kc * 8 + e] + b2s[kc * 8 + e], 0.0f) : 0.0f;
        uint4 hi, lo;
        split8(o, hi, lo);
        dst[(size_t)(kc * 2 + 0) * g.npix] = hi;
        dst[(size_t)(kc * 2 + 1) * g.npix] = lo;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------
// conv3 + ReLU + global mean
constexpr int C3_THREADS = 192;     // warp 0 loader, warp 1 MMA, warps 2-5 epilogue
constexpr int C3_NST = 8;           // weight ring stages
constexpr int C3_STAGE_BYTES = 2 * 2 * 128 * 16;   // [hl][kc 2][n 128][8 bf16] = 8 KB

struct Conv3Params {
  const __nv_bfloat16* act2;      // [B][16 planes][npix][8]
  const __nv_bfloat16* w3s;       // [j 4][tap 9][hl][kc 2][n 128][8]
  const float* b3;                // [128]
  float* pool_part;               // [B][n_groups][128]
  int B, npass;
  Geom g;
};

__global__ void __launch_bounds__(C3_THREADS, 1) conv3_kernel(Conv3Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  const uint32_t plane_bytes = (uint32_t)g.nsl3 * 16u;
  unsigned char* a_s = smem;                                    // 16 planes: index (kc*2 + hl), kc = 0..7
  unsigned char* b_s = a_s + 16 * plane_bytes;                  // C3_NST stages
  float* b3s = reinterpret_cast<float*>(b_s + C3_NST * C3_STAGE_BYTES);
  float* scratch = b3s + 128;                                   // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 512);
  uint64_t* a_full = bars;                 // [4]
  uint64_t* a_empty = bars + 4;            // [4]
  uint64_t* b_full = bars + 8;             // [NST]
  uint64_t* b_empty = bars + 8 + C3_NST;   // [NST]
  uint64_t* t_full = bars + 8 + 2 * C3_NST;
  uint64_t* t_empty = t_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 128) b3s[tid] = p.b3[tid];
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < C3_NST; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    mbar_init(t_full, 1);
    mbar_init(t_empty, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.B * g.n_groups;

  if (warp == 0) {
    // ===================== loader (one thread): activation k-slices + weight ring
    if (lane == 0) {
      int it = 0;
      uint32_t bs = 0;        // running weight-stage counter
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int b = item / g.n_groups, grp = item - b * g.n_groups;
        const int n_t = min(g.G, g.T3 - grp * g.G);
        const uint32_t nload = (uint32_t)(n_t * 128 + 2 * g.P + 2) * 16u;
        const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.act2) +
                                    ((size_t)b * 16 * g.npix + (size_t)grp * g.G * 128) * 16;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(a_empty + j, (it & 1) ^ 1, 40);
          mbar_arrive_expect_tx(a_full + j, 4 * nload);
#pragma unroll
          for (int pl = 0; pl < 4; ++pl)      // planes (kc = 2j, 2j+1) x (hi, lo) are consecutive: index 4j + pl
            bulk_g2s(a_s + (size_t)(4 * j + pl) * plane_bytes, src0 + (size_t)(4 * j + pl) * g.npix * 16, nload, a_full + j);
          for (int tap = 0; tap < 9; ++tap, ++bs) {
            const uint32_t st = bs % C3_NST;
            mbar_wait(b_empty + st, ((bs / C3_NST) & 1) ^ 1, 41);
            mbar_arrive_expect_tx(b_full + st, C3_STAGE_BYTES);
            bulk_g2s(b_s + st * C3_STAGE_BYTES,
                     reinterpret_cast<const unsigned char*>(p.w3s) + (size_t)(j * 9 + tap) * C3_STAGE_BYTES,
                     C3_STAGE_BYTES, b_full + st);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, 128);
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      const uint32_t lbo_a = 2u * plane_bytes;
      int it = 0;
      uint32_t bs = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int grp = item % g.n_groups;
        const int n_t = min(g.G, g.T3 - grp * g.G);
        mbar_wait(t_empty, (it & 1) ^ 1, 50);
        uint32_t acc = 0;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(a_full + j, it & 1, 51);
          for (int tap = 0; tap < 9; ++tap, ++bs) {
            const uint32_t st = bs % C3_NST;
            mbar_wait(b_full + st, (bs / C3_NST) & 1, 52);
            tc_fence_after();
            const int row_off = (g.P + 1) + (tap / 3 - 1) * g.P + (tap % 3 - 1);
            for (int ps = 0; ps < p.npass; ++ps) {
              const int hla = (ps == 2), hlb = (ps == 1);
              const uint64_t bd = make_desc(b_addr + st * C3_STAGE_BYTES + hlb * 4096, 2048, 128);
              for (int i = 0; i < n_t; ++i) {
                const uint64_t ad =
                    make_desc(a_addr + (uint32_t)(4 * j + hla) * plane_bytes + (uint32_t)(i * 128 + row_off) * 16u, lbo_a, 128);
                umma_bf16(tmem_base + i * 128, ad, bd, idesc, acc);
              }
              acc = 1;
            }
            umma_commit(b_empty + st);
          }
          umma_commit(a_empty + j);
        }
        umma_commit(t_full);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> bias + ReLU + mask -> sum over pixels
    const int q = warp & 3;
    const int et = tid - 64;      // 0..127
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / g.n_groups, grp = item - b * g.n_groups;
      const int n_t = min(g.G, g.T3 - grp * g.G);
      mbar_wait(t_full, it & 1, 60);
      tc_fence_after();
      float pool[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int i = 0; i < n_t; ++i) {
        const int pidx = g.P + 128 * (grp * g.G + i) + q * 32 + lane;
        int y, x;
        const bool ok = pix_valid(pidx, g, y, x);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + i * 128 + cc * 32, v);
#pragma unroll
          for (int r = 0; r < 32; ++r) v[r] = ok ? fmaxf(v[r] + b3s[cc * 32 + r], 0.0f) : 0.0f;
          // transpose-reduce over the 32 lanes (pixels): lane L ends with the sum of channel cc*32 + L
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int r = 0; r < s; ++r) {
              const float send = up ? v[r] : v[r + s];
              const float keep = up ? v[r + s] : v[r];
              v[r] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
          }
          pool[cc] += v[0];
        }
      }
      tc_fence_before();
      mbar_arrive(t_empty);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) scratch[q * 128 + cc * 32 + lane] = pool[cc];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // fixed order 0..3 over quadrants -> deterministic
      const float s = scratch[et] + scratch[128 + et] + scratch[256 + et] + scratch[384 + et];
      p.pool_part[((size_t)b * g.n_groups + grp) * 128 + et] = s;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
inline uint16_t f2bf(float f) {        // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

Geom make_geom(const ww_ctx* c) {
  Geom g;
  g.H = c->cfg.n_mels;
  g.W = c->W;
  g.P = g.W + 1;
  g.T3 = (g.H * g.P + 127) / 128;
  const int need = 2 * g.P + 128 * g.T3 + 2;
  g.T2 = (need + 127) / 128;
  g.npix = g.T2 * 128;
  // tiles per group: as many as TMEM (4 x 128 columns) and shared memory allow, preferring an even split
  int best = 1;
  for (int G = 4; G >= 1; --G) {
    const int nsl3 = (G * 128 + 2 * g.P + 2 + 7) & ~7;
    const size_t smem = (size_t)16 * nsl3 * 16 + (size_t)C3_NST * C3_STAGE_BYTES + 4096;
    if (smem <= 227 * 1024) { best = G; break; }
  }
  // prefer the G (<= best) that minimises idle tile slots in the last group
  int bestG = best, waste = (best - g.T3 % best) % best;
  for (int G = best - 1; G >= 2; --G) {
    const int w = (G - g.T3 % G) % G;
    if (w < waste) { bestG = G; waste = w; }
  }
  g.G = bestG;
  g.n_groups = (g.T3 + g.G - 1) / g.G;
  g.nsl2 = (128 + 2 * g.P + 2 + 7) & ~7;
  g.nsl3 = (g.G * 128 + 2 * g.P + 2 + 7) & ~7;
  return g;
}

size_t conv12_smem(const Geom& g) { return (size_t)W2_BYTES + 2 * 8 * (size_t)g.nsl2 * 16 + (288 + 32 + 64) * 4 + 16 * 8 + 64; }
size_t conv3_smem(const Geom& g) {
  return (size_t)16 * g.nsl3 * 16 + (size_t)C3_NST * C3_STAGE_BYTES + (128 + 512) * 4 + (10 + 2 * C3_NST) * 8 + 64;
}

}  // namespace

size_t ww_conv_tc_act2_bytes_per_clip(const ww_ctx* c) {
  const Geom g = make_geom(c);
  return (size_t)16 * g.npix * 16;
}

int ww_conv_tc_groups(const ww_ctx* c) { return make_geom(c).n_groups; }

// bf16 hi/lo split of conv2 / conv3 weights into the UMMA canonical layouts
int ww_conv_tc_prepare(ww_ctx* c, cudaStream_t) {
  auto fetch = [&](const char* name, size_t n) {
    std::vector<float> h(n);
    cudaMemcpy(h.data(), c->w[name], n * sizeof(float), cudaMemcpyDeviceToHost);
    return h;
  };
  {
    std::vector<float> w = fetch("conv2.weight", (size_t)64 * 32 * 9);       // [n][cin][tap]
    std::vector<uint16_t> s((size_t)W2_BYTES / 2);
    for (int tap = 0; tap < 9; ++tap)
      for (int kc = 0; kc < 4; ++kc)
        for (int n = 0; n < 64; ++n)
          for (int e = 0; e < 8; ++e) {
            const float v = w[((size_t)n * 32 + kc * 8 + e) * 9 + tap];
            const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
            s[((((size_t)tap * 2 + 0) * 4 + kc) * 64 + n) * 8 + e] = hi;
            s[((((size_t)tap * 2 + 1) * 4 + kc) * 64 + n) * 8 + e] = lo;
          }
    if (!c->d_w2_split) WW_CHECK(c, cudaMalloc((void**)&c->d_w2_split, W2_BYTES));
    WW_CHECK(c, cudaMemcpy(c->d_w2_split, s.data(), W2_BYTES, cudaMemcpyHostToDevice));
  }
  {
    std::vector<float> w = fetch("conv3.weight", (size_t)128 * 64 * 9);
    const size_t bytes = (size_t)36 * C3_STAGE_BYTES;
    std::vector<uint16_t> s(bytes / 2);
    for (int j = 0; j < 4; ++j)
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < 128; ++n)
            for (int e = 0; e < 8; ++e) {
              const float v = w[((size_t)n * 64 + j * 16 + kc * 8 + e) * 9 + tap];
              const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
              const size_t blk = ((size_t)j * 9 + tap) * (C3_STAGE_BYTES / 2);
              s[blk + (((size_t)0 * 2 + kc) * 128 + n) * 8 + e] = hi;
              s[blk + (((size_t)1 * 2 + kc) * 128 + n) * 8 + e] = lo;
            }
    if (!c->d_w3_split) WW_CHECK(c, cudaMalloc((void**)&c->d_w3_split, bytes));
    WW_CHECK(c, cudaMemcpy(c->d_w3_split, s.data(), bytes, cudaMemcpyHostToDevice));
  }
  return WW_OK;
}

int ww_launch_conv_tc(ww_ctx* c, const float* logmel, int B, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  const Geom g = make_geom(c);
  const size_t s12 = conv12_smem(g), s3 = conv3_smem(g);
  if (s12 > 227 * 1024 || s3 > 227 * 1024) {
    c->set_error("conv_tc: frame count too large for the shared-memory tiles (use WW_CONV_FP32)");
    return WW_ERR_INVALID;
  }
  static size_t conf12 = 0, conf3 = 0;
  if (s12 > conf12) {
    WW_CHECK(c, cudaFuncSetAttribute(conv12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s12));
    conf12 = s12;
  }
  if (s3 > conf3) {
    WW_CHECK(c, cudaFuncSetAttribute(conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
    conf3 = s3;
  }
  const int npass = c->cfg.conv_mode == WW_CONV_BF16 ? 1 : 3;
  c->n_pool_part = g.n_groups;
  {
    Conv12Params p;
    p.logmel = logmel; p.w1t = c->d_convw_t[0]; p.b1 = c->w["conv1.bias"]; p.w2s = c->d_w2_split;
    p.b2 = c->w["conv2.bias"]; p.act2 = c->ws_act2_split; p.B = B; p.npass = npass; p.g = g;
    const int grid = std::min(c->sm_count, B * g.T2);
    ProfScope prof(c, WW_STAGE_CONV12, st);
    conv12_kernel<<<grid, C12_THREADS, s12, st>>>(p);
    WW_LAUNCH_CHECK(c);
  }
  {
    Conv3Params p;
    p.act2 = c->ws_act2_split; p.w3s = c->d_w3_split; p.b3 = c->w["conv3.bias"]; p.pool_part = c->ws_pool_part;
    p.B = B; p.npass = npass; p.g = g;
    const int grid = std::min(c->sm_count, B * g.n_groups);
    ProfScope prof(c, WW_STAGE_CONV3, st);
    conv3_kernel<<<grid, C3_THREADS, s3, st>>>(p);
    WW_LAUNCH_CHECK(c);
  }
  return WW_OK;
}

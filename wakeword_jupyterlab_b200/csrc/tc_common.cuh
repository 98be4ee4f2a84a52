// Shared pieces of the tcgen05 convolution kernels: PTX wrappers (mbarrier, bulk copy, TMEM, UMMA),
// descriptor builders, fp16 packing / hi-lo weight splitting and the pixel-linear image geometry.  sm_100a only.
#pragma once
#include "ctx.cuh"

#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>

#include <string.h>
#include <algorithm>

namespace tc {

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
// One arrival per WARP: every lane's earlier shared-memory / TMEM accesses are ordered before it by the __syncwarp.
// (A per-thread arrive wakes every suspended waiter of the CTA 32 times as often.)
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or ~20 us pass), so
// waiting warps neither burn issue slots nor hammer the shared-memory pipe with polls.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails loudly) instead of hanging the GPU.
static __device__ __noinline__ void mbar_timeout(int code) {
  printf("conv_tc watchdog: wait %d timed out (block %d, thread %d)\n", code, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 17)) mbar_timeout(code);      // ~2.6 s of 20 us suspensions
  }
}
// Polite wait: a failed try_wait comes back within a few cycles on this part (the suspend hint is only a hint), so a
// spinning warp issues ~1 shared-memory-pipe operation per loop trip and, with a dozen waiting warps per CTA, saturates
// the pipe the working warps need for LDS / STS / SHFL.  Sleeping `ns` between polls costs at most that much wake-up
// latency and frees the pipe.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, int code, unsigned ns) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++spins > (1u << 24)) mbar_timeout(code);
  }
}
// roles off the critical path (producers / epilogue warps) use the same hardware-suspended wait
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int code) { mbar_wait(bar, parity, code); }
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// same instruction with 8-bit operands (kind::f8f6f4, here e4m3 x e4m3 -> fp32): K = 32 per instruction at the cycle
// count of a K = 16 fp16 instruction.  The instruction descriptor has the same bit layout (format 0 = e4m3).
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// true on exactly one (converged) lane of the warp; ptxas understands elect.sync-guarded regions and keeps the
// guarded tcgen05 / bulk-copy instructions on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive columns: thread (lane) receives 32 columns of its TMEM lane.  No wait inside.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
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
}
// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (stride between the two 8-element K chunks of one MMA), [32,46) SBO>>4
// (stride between 8-row groups), [46,48) version = 1, [61,64) layout type = 0 (no swizzle).
// With no swizzle the start address only needs 16-byte alignment, which is what makes the shifted-view
// (3x3 tap = pointer offset) trick legal.  Adding (bytes >> 4) to the 64-bit value advances the start.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (kind::f16): D = f32 (bit 4), A = B = fp16 (format fields at bits 7 and 10 = 0), both
// K-major, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// two floats -> packed fp16x2 (round to nearest even, saturating to +-65504 instead of inf): `lo` in bits 0..15
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// two floats -> packed e4m3x2 (round to nearest, saturating to +-448), optionally with ReLU: `lo` in bits 0..7
template <bool RELU>
__device__ __forceinline__ uint32_t pack_e4m3(float lo, float hi) {
  unsigned short r;
  if (RELU) asm("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return (uint32_t)r;
}
// 16 floats -> one 16-byte core-matrix row of e4m3
template <bool RELU>
__device__ __forceinline__ uint4 cvt16_e4m3(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = pack_e4m3<RELU>(v[4 * i], v[4 * i + 1]) | (pack_e4m3<RELU>(v[4 * i + 2], v[4 * i + 3]) << 16);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// 8 floats -> one 16-byte core-matrix row of fp16
__device__ __forceinline__ uint4 cvt8(const float* v) {
  return make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
}

// ------------------------------------------------------------------------------------------------
// Pixel-linear zero-padded image: pixel (y, x) <-> padded index p = (y+1)*P + (x+1), pitch P = W + 1;
// activation planes store slot s = p + 1 (one leading slot so that tap offsets never go negative).
struct Geom {
  int H, W, P;        // image rows (mels), cols (frames), pitch
  int npix;           // plane slots per clip (multiple of 128)
  int T2, T3;         // conv2 tiles (npix / 128, even), conv3 tiles per clip (ceil((H+1)*P / 128)): the tape period
  int G, n_groups;    // conv3: tiles per group (<= 4); pool partials per clip (= T3: one per tile)
  int nsl2;           // conv12 A-buffer slots = round8(256 + 2P + 2) (two tiles + halo)
  int NL, NM;         // conv1 pixels per item (256 + 2P + 2) and conv1 M-tiles per item (ceil(NL / 128))
  int lead;           // zero floats in front of padded index 0 of the padded log-mel image (2P + 3)
  int npix_in;        // floats per clip of the padded log-mel image
  int patch_f;        // floats of one item's input patch = round4(NM*128 + 2P + 2)
  int nsl3;           // conv3 plane slots    = round8(G*128 + 2P + 2)
  int nst3;           // conv3 weight ring stages
  uint32_t magicP;    // ceil(2^32 / P): p / P == umulhi(p, magicP) for 0 <= p < 65536
};


__device__ __forceinline__ bool pix_valid(int p, const Geom& g, int& y, int& x) {
  if (p < 0) return false;
  const int row = (int)__umulhi((uint32_t)p, g.magicP);
  y = row - 1;
  x = p - row * g.P - 1;
  return (y >= 0) && (y < g.H) && (x >= 0);
}

// host: fp16 round-to-nearest-even (bit pattern) and back
inline uint16_t f2h(float f) {
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
inline float h2f(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
// host: e4m3 round-to-nearest, saturating (bit pattern) -- the lo half of the conv3 weights
inline uint8_t f2e4m3(float f) { return (uint8_t)__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E4M3); }

// Weight scale: the tensor-core operands are w * 2^k as fp16 hi + fp16 lo with k chosen so the largest |w| lands
// near 2^13.  hi is then far from fp16's subnormal range and lo (~2^-11 of hi) stays a normal number for every
// weight that matters; the epilogues multiply the fp32 accumulator by 2^-k (exact).
inline float weight_scale(const std::vector<float>& w) {
  float m = 0.0f;
  for (float v : w) m = std::max(m, fabsf(v));
  if (!(m > 0.0f) || !std::isfinite(m)) return 1.0f;
  int e;
  frexpf(m, &e);                       // m = f * 2^e, f in [0.5, 1)
  const int k = std::min(std::max(13 - e, -24), 24);
  return ldexpf(1.0f, k);
}

// conv3 weight stage = everything tap row tt needs from one PAIR of 16-channel slices (32 input channels):
//   [hi fp16 of slice 2p: tap 3][kc 2][cout 128][8] 12 KB | [hi fp16 of slice 2p+1] 12 KB | [lo e4m3 of the pair:
//   tap 3][kc16 2][cout 128][16] 12 KB   = 36 KB, the operand of 9 MMAs per accumulator half (1,152 tensor cycles).
// 6 stages per item; global memory holds them in consumption order (pair, tap row).
constexpr int C3_PART_BYTES = 3 * 2 * 128 * 16;          // 12 KB
constexpr int C3_STAGE_BYTES = 3 * C3_PART_BYTES;        // 36 KB
constexpr int C3_STAGES_PER_ITEM = 6;
constexpr int C3_NST_MAX = 4;                            // ring slots (as many as shared memory allows)
constexpr int C3_NST_MIN = 3;

inline size_t conv3_smem_bytes(int nsl3, int nst) {
  return (size_t)12 * nsl3 * 16 + (size_t)nst * C3_STAGE_BYTES + 128 * 4 + 256 * 4 + 32 * 8 + 64;
}

inline Geom make_geom(const ww_ctx* c) {
  Geom g;
  g.H = c->cfg.n_mels;
  g.W = c->W;
  g.P = g.W + 1;
  // conv3 walks the clips of a launch as ONE pixel-linear tape with period 128 * T3 slots per clip: clip b + 1's slot 0
  // follows clip b's slot 128 * T3 - 1, so a group of G tiles may straddle two clips and EVERY group is full.  (H+1)*P <=
  // 128 * T3 makes the seam safe: the valid outputs of a clip then read nothing past the next clip's P + 2 leading zero slots.
  g.T3 = ((g.H + 1) * g.P + 127) / 128;
  const int need = 2 * g.P + 128 * g.T3 + 2;
  g.T2 = (need + 127) / 128;
  g.T2 += g.T2 & 1;                   // conv12 works on tile pairs
  g.npix = g.T2 * 128;
  g.G = 1;
  for (int G = 4; G >= 1; --G) {      // as many tiles per weight pass as shared memory allows
    const int nsl3 = (G * 128 + 2 * g.P + 2 + 7) & ~7;
    if (conv3_smem_bytes(nsl3, C3_NST_MIN) <= 227 * 1024) { g.G = G; break; }
  }
  g.n_groups = g.T3;
  g.nsl2 = (256 + 2 * g.P + 2 + 7) & ~7;
  g.NL = 256 + 2 * g.P + 2;
  g.NM = (g.NL + 127) / 128;
  g.lead = 2 * g.P + 3;
  g.patch_f = (g.NM * 128 + 2 * g.P + 2 + 3) & ~3;
  g.npix_in = (128 * g.T2 + g.patch_f + 3) & ~3;     // last item starts at 128*T2 - 256 and reads patch_f floats
  g.nsl3 = (g.G * 128 + 2 * g.P + 2 + 7) & ~7;
  g.nst3 = C3_NST_MIN;
  while (g.nst3 < C3_NST_MAX && conv3_smem_bytes(g.nsl3, g.nst3 + 1) <= 227 * 1024) ++g.nst3;
  g.magicP = (uint32_t)((0x100000000ull + (uint64_t)g.P - 1) / (uint64_t)g.P);
  return g;
}

}  // namespace tc

int ww_launch_conv12_tc(ww_ctx* c, const float* in_pad, int B, const tc::Geom& g, cudaStream_t st);
int ww_launch_conv3_tc(ww_ctx* c, int B, const tc::Geom& g, cudaStream_t st);

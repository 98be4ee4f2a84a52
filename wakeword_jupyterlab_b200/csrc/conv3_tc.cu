// K3b (tensor-core path): conv3 (64->128, 3x3) + bias + ReLU + global mean on tcgen05 / TMEM  (sm_100a)
//
// Replaces F.relu(self.conv3(x)) and self.pool(x) of WakewordModel.forward
// (/root/reference/wakeword_training_script.py:172-173): 80 % of the model's FLOPs.
//
// Implicit GEMM  D[cout, pixel] += W[cout, (tap, cin)] * Act[pixel + tap_offset, cin]:
//   * A operand = weights, M = 128 = Cout (TMEM lane = output channel);
//   * B operand = activations, N = 256 pixels per instruction (two 128-pixel tiles), K = 16 channels.
//   Activations live in a zero-padded, pixel-linear image (pixel (y,x) at padded index (y+1)*P + (x+1),
//   pitch P = W+1) stored channel-chunk-major  [chunk of 8 channels][pixel][8 x fp16 = 16 B]  -- exactly the
//   UMMA K-major SWIZZLE_NONE canonical layout (core matrix = 8 pixels x 16 B contiguous, SBO = 128 B,
//   LBO = chunk-plane stride).  A 3x3 tap is therefore just a different START ADDRESS of the same shared
//   memory tile: start += ((ky-1)*P + (kx-1)) * 16 B.  No im2col exists anywhere; each activation byte is
//   copied to shared memory once per tile group and read by the tensor core 9 taps x NPASS passes times.
//   * fp32 parity (logits <= 1e-4 relative; SURVEY.md section 7 hard part 1).  Rounding errors of the ACTIVATIONS
//   are independent from pixel to pixel and average out in the global mean; rounding errors of the WEIGHTS are the
//   same at every pixel and do not.  So activations are a single fp16 value (11-bit significand) and only the
//   weights are split: W * 2^k = fp16 hi + lo, accumulated in fp32 as W_hi*a + W_lo*a.  The correction term only
//   needs a few bits, so W_lo and a second copy of the activations are e4m3 and that pass runs as kind::f8f6f4
//   (K = 32 per instruction at the cost of a K = 16 fp16 one): 1.5 passes instead of 2 (WW_CONV_SPLIT2: logits
//   1e-6 .. 8e-6 relative on the golden weights, same as an fp16 lo).  WW_CONV_FP16 issues W_hi*a only (5e-5).
//
// Work item = group of G <= 4 consecutive 128-pixel tiles of the TAPE formed by the clips of the launch (clip b + 1's
// padded image follows clip b's at a period of 128 * T3 slots: the zero rows under one image and above the next coincide),
// so that every weight stage fetched from L2 feeds 512 pixels and all 512 TMEM columns hold fp32 accumulators: every group
// is full (round 1 split the 21 tiles of a clip 4,4,4,3,3,3 and the 3-tile groups ran their odd tile as N = 128 MMAs at
// 84 % of the ideal rate against 92 %).  A group that straddles two clips loads its planes in two segments and its tiles
// report to their own clips: the pooling partials are per TILE, so a clip's result does not depend on where it sits.
//   warp 0      loader (one thread): activation planes of a 32-channel slice pair (4 fp16 + 2 e4m3 1-D cp.async.bulk;
//               two pairs = the whole K, each refilled for the next item as soon as it is consumed) and the weight
//               ring (3 stages x 36 KB = (slice pair, tap row): 9 MMAs per accumulator half);
//   warp 1      MMA issuer (one thread): descriptors are 64-bit adds on precomputed bases;
//   warps 2-9   epilogue: tcgen05.ld (lane = channel, 32 pixels per load) -> bias + ReLU + padding mask
//               (precomputed bit masks) -> per-thread sum over pixels -> one deterministic partial per
//               (clip, group, channel); the head kernel finishes the mean.
#include "tc_common.cuh"

#include <algorithm>
#include <stdlib.h>

using namespace tc;

#define C3_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && it < 48 && lane == 0) p.trace[it * 16 + (slot)] = clock64(); } while (0)

namespace {

constexpr int C3_THREADS = 320;     // warp 0 loader, warp 1 MMA, warps 2-9 epilogue

struct Conv3Params {
  const __half* act2;             // [B][8 planes][npix][8 fp16]
  const uint8_t* act2_8;          // [B][4 planes][npix][16 e4m3]
  const unsigned char* w3s;       // [pair 2][tap row 3] stages of 36 KB (tc_common.cuh), weights scaled by 2^k
  float inv_scale;                // 2^-k
  const float* b3;                // [128]
  const uint32_t* mask;           // [T3][4] validity bits of the 128 pixels of each tile
  float* pool_part;               // [B][T3][128]: one partial sum per tile
  uint32_t* relu_bits;            // BITS (training forward): [B][T3][4][128] bit r of word (t, k, ch) = output > 0 at pixel 32 k + r of tile t
  int B;
  Geom g;
  long long* trace;               // debug (WW_TC_TRACE=1): per-group role timestamps of CTA 0
};

template <int NPASS, bool BITS>
__global__ void __launch_bounds__(C3_THREADS, 1) conv3_kernel(Conv3Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Geom g = p.g;
  const uint32_t plane_bytes = (uint32_t)g.nsl3 * 16u;
  const int NST = g.nst3;
  unsigned char* a_s = smem;                                    // 8 fp16 activation planes (chunks of 8 channels)
  unsigned char* a8_s = a_s + 8 * plane_bytes;                  // 4 e4m3 activation planes (chunks of 16 channels)
  unsigned char* w_s = a8_s + 4 * plane_bytes;                  // weight ring
  float* b3s = reinterpret_cast<float*>(w_s + NST * C3_STAGE_BYTES);
  float* scratch = b3s + 128;                                   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 256);
  uint64_t* a_full = bars;                 // [2] per slice pair (32 input channels)
  uint64_t* a_empty = bars + 4;            // [2]
  uint64_t* w_full = bars + 8;             // [NST_MAX]
  uint64_t* w_empty = bars + 8 + C3_NST_MAX;   // [NST_MAX]
  uint64_t* t_full = bars + 8 + 2 * C3_NST_MAX;   // [2] accumulator halves: tiles (0,1) = TMEM cols 0..255, tiles (2,3) = 256..511
  uint64_t* t_empty = t_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, so role branches are uniform
  // branches and the MMA issue loop runs on the uniform datapath (no per-MMA R2UR waterfall)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid < 128) b3s[tid] = p.b3[tid];
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < C3_NST_MAX; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // Every CTA takes a CONTIGUOUS range of tape groups.
  const int n_tiles = p.B * g.T3;                       // < 2^31: a launch is one chunk of clips
  const int n_items = (n_tiles + g.G - 1) / g.G;
  const int item_lo = (int)((long long)n_items * blockIdx.x / gridDim.x);
  const int item_hi = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);
  auto item_tiles = [&](int item) { return min(g.G, n_tiles - item * g.G); };

  if (warp == 0) {
    // ===================== loader (one thread)
    if (lane == 0) {
      int it = 0;
      uint32_t st = 0, wpar = 1;        // weight ring position and the parity to wait for on w_empty
      constexpr uint32_t wbytes = NPASS == 2 ? C3_STAGE_BYTES : 2 * C3_PART_BYTES;
      for (int item = item_lo; item < item_hi; ++item, ++it) {
        const int tau0 = item * g.G;                                // first tape tile of the group
        const int b = tau0 / g.T3, t0 = tau0 - b * g.T3, n_t = item_tiles(item);
        // window = slots [128 t0, 128 t0 + nslots) of clip b; past slot 128 T3 the tape continues with clip b + 1's slot 0
        const uint32_t nslots = (uint32_t)(n_t * 128 + 2 * g.P + 2);
        uint32_t n1 = nslots, n2 = 0;
        if (128u * t0 + nslots > 128u * g.T3 && b + 1 < p.B) { n1 = 128u * (g.T3 - t0); n2 = nslots - n1; }
        const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.act2) + ((size_t)b * 8 * g.npix + (size_t)t0 * 128) * 16;
        const unsigned char* src8 = p.act2_8 + ((size_t)b * 4 * g.npix + (size_t)t0 * 128) * 16;
        const unsigned char* nxt0 = reinterpret_cast<const unsigned char*>(p.act2) + (size_t)(b + 1) * 8 * g.npix * 16;
        const unsigned char* nxt8 = p.act2_8 + (size_t)(b + 1) * 4 * g.npix * 16;
        for (int pr = 0; pr < 2; ++pr) {
          // activations of the pair: 4 fp16 planes (chunks of 8 channels) + 2 e4m3 planes (chunks of 16 channels)
          mbar_wait(a_empty + pr, (it & 1) ^ 1, 40);
          if (pr == 0) C3_TRACE(0);
          mbar_arrive_expect_tx(a_full + pr, (NPASS == 2 ? 6 : 4) * nslots * 16u);
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) {
            unsigned char* dst = a_s + (size_t)(4 * pr + pl) * plane_bytes;
            bulk_g2s(dst, src0 + (size_t)(4 * pr + pl) * g.npix * 16, n1 * 16u, a_full + pr);
            if (n2) bulk_g2s(dst + n1 * 16u, nxt0 + (size_t)(4 * pr + pl) * g.npix * 16, n2 * 16u, a_full + pr);
          }
          if (NPASS == 2) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              unsigned char* dst = a8_s + (size_t)(2 * pr + pl) * plane_bytes;
              bulk_g2s(dst, src8 + (size_t)(2 * pr + pl) * g.npix * 16, n1 * 16u, a_full + pr);
              if (n2) bulk_g2s(dst + n1 * 16u, nxt8 + (size_t)(2 * pr + pl) * g.npix * 16, n2 * 16u, a_full + pr);
            }
          }
          for (int tt = 0; tt < 3; ++tt) {
            mbar_wait(w_empty + st, wpar, 41);
            mbar_arrive_expect_tx(w_full + st, wbytes);
            bulk_g2s(w_s + st * C3_STAGE_BYTES, p.w3s + (size_t)(pr * 3 + tt) * C3_STAGE_BYTES, wbytes, w_full + st);
            if (++st == (uint32_t)NST) { st = 0; wpar ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer.  The whole warp runs the loop so that descriptors stay warp-uniform
    // (uniform registers feed UTCHMMA directly); only the elected lane issues, one straight-line region per stage visit.
    constexpr uint32_t idesc256 = make_idesc(128, 256), idesc128 = make_idesc(128, 128);
    const uint64_t wdesc0 = make_desc(smem_u32(w_s), 2048, 128);              // weights: kc stride 2 KB
    const uint64_t pdesc0 = make_desc(smem_u32(a_s), plane_bytes, 128);       // pixels: kc stride = 1 plane
    const uint64_t p8desc0 = make_desc(smem_u32(a8_s), plane_bytes, 128);     // e4m3 pixels: 16-channel chunk stride = 1 plane
    const uint32_t plane_u = plane_bytes >> 4;
    // The 9 (6 in fp16 mode) MMAs one stage feeds into accumulator half `h` (0: tiles 0,1   1: tiles 2,3): per tap of the
    // row, fp16 W_hi * a over the two 16-channel slices of pair `pr`, then e4m3 W_lo * a over its 32 channels.
    auto stage_mmas = [&](uint32_t slot, int pr, int tt, int h, uint32_t idesc, bool first) {
      const uint32_t d = tmem_base + h * 256;
#pragma unroll
      for (int tl = 0; tl < 3; ++tl) {
        const uint32_t row_off = (uint32_t)((g.P + 1) + (tt - 1) * g.P + (tl - 1)) + h * 256;   // tap (ky, kx) = (tt, tl)
        const uint64_t wd = wdesc0 + (uint64_t)((slot * C3_STAGE_BYTES + tl * 4096) >> 4);
        umma_f16(d, wd, pdesc0 + (uint64_t)(4 * pr * plane_u + row_off), idesc, !(first && tl == 0));
        umma_f16(d, wd + (C3_PART_BYTES >> 4), pdesc0 + (uint64_t)((4 * pr + 2) * plane_u + row_off), idesc, 1);
        if (NPASS == 2) umma_f8(d, wd + (2 * C3_PART_BYTES >> 4), p8desc0 + (uint64_t)(2 * pr * plane_u + row_off), idesc, 1);
      }
    };
    int it = 0, it1 = 0;      // it1 counts the groups that use accumulator half 1 (its barriers flip only then)
    uint32_t st = 0, wpar = 0;        // weight ring position and the parity to wait for on w_full
    auto ring_next = [&](uint32_t s_, uint32_t& par_) { if (++s_ == (uint32_t)NST) { s_ = 0; par_ ^= 1; } return s_; };
    for (int item = item_lo; item < item_hi; ++item, ++it) {
      const int n_t = item_tiles(item);
      // MMA shapes for this group: tiles (0,1) -> N = 256 or 128; tiles (2,3) -> N = 256, 128 or none
      const uint32_t id0 = n_t >= 2 ? idesc256 : idesc128;
      const uint32_t id1 = n_t >= 4 ? idesc256 : idesc128;
      const bool second = n_t >= 3;
      const uint32_t par = it & 1;
      // Stage order of an item: S0 S1 | S2 S3 | S4 S5.  The first two and the last two stages run the accumulator halves
      // one after the other (h0 of both stages, then h1 of both), the middle two interleave them: the epilogue of half 0
      // then overlaps the last h1 sweep, the epilogue of half 1 overlaps the first h0 sweep of the NEXT item, and each
      // has 2 x 1,152 tensor cycles to finish.
      // ---- head: S0, S1 (pair 0, tap rows 0, 1)
      mbar_wait(a_full + 0, par, 51);
      C3_TRACE(2);
      mbar_wait(t_empty + 0, par ^ 1, 50);
      C3_TRACE(1);
      {
        uint32_t s1 = st, p1 = wpar;
        for (int k = 0; k < 2; ++k) {
          mbar_wait(w_full + s1, p1, 52);
          tc_fence_after();
          if (elect_one()) stage_mmas(s1, 0, k, 0, id0, k == 0);
          __syncwarp();
          s1 = ring_next(s1, p1);
        }
        if (second) mbar_wait(t_empty + 1, (it1 & 1) ^ 1, 53);
        tc_fence_after();
        for (int k = 0; k < 2; ++k) {
          if (elect_one()) {
            if (second) stage_mmas(st, 0, k, 1, id1, k == 0);
            umma_commit(w_empty + st);
          }
          __syncwarp();
          st = ring_next(st, wpar);
        }
      }
      // ---- middle: S2 (pair 0, tap row 2), S3 (pair 1, tap row 0), halves interleaved
      for (int k = 2; k < 4; ++k) {
        if (k == 3) { mbar_wait(a_full + 1, par, 51); C3_TRACE(3); }
        mbar_wait(w_full + st, wpar, 52);
        tc_fence_after();
        if (elect_one()) {
          stage_mmas(st, k / 3, k % 3, 0, id0, false);
          if (second) stage_mmas(st, k / 3, k % 3, 1, id1, false);
          umma_commit(w_empty + st);
          if (k == 2) umma_commit(a_empty + 0);
        }
        __syncwarp();
        st = ring_next(st, wpar);
      }
      // ---- tail: S4, S5 (pair 1, tap rows 1, 2)
      {
        uint32_t s1 = st, p1 = wpar;
        for (int k = 1; k < 3; ++k) {
          mbar_wait(w_full + s1, p1, 52);
          tc_fence_after();
          if (elect_one()) {
            stage_mmas(s1, 1, k, 0, id0, false);
            if (k == 2) umma_commit(t_full + 0);
          }
          __syncwarp();
          s1 = ring_next(s1, p1);
        }
        for (int k = 1; k < 3; ++k) {
          if (elect_one()) {
            if (second) stage_mmas(st, 1, k, 1, id1, false);
            umma_commit(w_empty + st);
            if (k == 2) {
              if (second) umma_commit(t_full + 1);
              umma_commit(a_empty + 1);
            }
          }
          __syncwarp();
          st = ring_next(st, wpar);
        }
        if (second) ++it1;
      }
      C3_TRACE(6);
    }
  } else {
    // ===================== epilogue (8 warps): lane = output channel, columns = pixels
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int sub = (warp - 2) >> 2;        // which tile of each accumulator half this warp drains (the two warps that
                                            // share a lane quadrant split the half, so all 8 warps work on the ready half)
    const int ch = q * 32 + lane;
    const float bias = b3s[ch], inv_s = p.inv_scale;
    int it = 0, it1 = 0;
    for (int item = item_lo; item < item_hi; ++item, ++it) {
      const int tau0 = item * g.G;
      const int n_t = item_tiles(item);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h * 2 >= n_t) break;            // half h is unused by this group (no MMAs, no barrier flips)
        mbar_wait_relaxed(t_full + h, (h ? it1 : it) & 1, 60 + h);
        if (warp == 2 && h == 0) C3_TRACE(7);
        tc_fence_after();
        const int i = h * 2 + sub;
        if (i < n_t) {
          const int tau = tau0 + i;                         // tape tile -> (clip, tile of the clip)
          const int b = tau / g.T3, t = tau - b * g.T3;
          const uint4 m = __ldg(reinterpret_cast<const uint4*>(p.mask) + t);
          const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + i * 128;
          float sum = 0.0f;
#pragma unroll
          for (int cp = 0; cp < 2; ++cp) {
            uint32_t r0[32], r1[32];
            tmem_ld32_nowait(taddr + cp * 64, r0);
            tmem_ld32_nowait(taddr + cp * 64 + 32, r1);
            tmem_ld_wait();
            float s0 = 0.0f, s1 = 0.0f;
            uint32_t pos0 = 0u, pos1 = 0u;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float v0 = fmaxf(fmaf(__uint_as_float(r0[r]), inv_s, bias), 0.0f);
              const float v1 = fmaxf(fmaf(__uint_as_float(r1[r]), inv_s, bias), 0.0f);
              s0 += (mw[cp * 2] >> r) & 1u ? v0 : 0.0f;
              s1 += (mw[cp * 2 + 1] >> r) & 1u ? v1 : 0.0f;
              if (BITS) { pos0 |= (v0 > 0.0f ? 1u : 0u) << r; pos1 |= (v1 > 0.0f ? 1u : 0u) << r; }
            }
            sum += s0 + s1;
            if (BITS) {      // the ReLU derivative of the backward pass (padding pixels: 0)
              uint32_t* rb = p.relu_bits + (((size_t)b * g.T3 + t) * 4 + cp * 2) * 128 + ch;
              rb[0] = pos0 & mw[cp * 2];
              rb[128] = pos1 & mw[cp * 2 + 1];
            }
          }
          // one partial per (clip, tile, channel): a clip's mean does not depend on how its tiles fell into groups
          p.pool_part[((size_t)b * g.T3 + t) * 128 + ch] = sum;
        }
        tc_fence_before();
        mbar_arrive_warp(t_empty + h, lane);
        if (warp == 2 && h == 0) C3_TRACE(8);
        if (h) ++it1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t ww_conv_tc_act2_bytes_per_clip(const ww_ctx* c) {      // fp16 planes; the e4m3 copy is half of this
  const Geom g = make_geom(c);
  return (size_t)8 * g.npix * 16;
}

int ww_conv_tc_groups(const ww_ctx* c) { return make_geom(c).n_groups; }      // pool partials per clip (one per tile)

// where the log-mel kernel writes pixel (mel m, frame t) of clip b for the tensor-core path:
// ptr[b * stride + off + m * pitch + t]
LogmelOut ww_conv_tc_logmel_out(const ww_ctx* c, float* in_pad) {
  const Geom g = make_geom(c);
  return LogmelOut{in_pad, g.P, g.lead + g.P + 1, (int64_t)g.npix_in};
}

int ww_conv12_tc_prepare(ww_ctx* c);

// conv3 weights -> scaled fp16 hi/lo in the UMMA canonical layout [j][tap][hl][kc][cout][8]; tile validity masks
int ww_conv_tc_prepare(ww_ctx* c, cudaStream_t) {
  int rc = ww_conv12_tc_prepare(c);
  if (rc) return rc;
  std::vector<float> w((size_t)128 * 64 * 9);       // [cout][cin][tap]
  WW_CHECK(c, cudaMemcpy(w.data(), c->w["conv3.weight"], w.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<unsigned char> s((size_t)C3_STAGES_PER_ITEM * C3_STAGE_BYTES, 0);
  const float sc = weight_scale(w);
  c->w3_inv_scale = 1.0f / sc;
  for (int n = 0; n < 128; ++n)
    for (int ci = 0; ci < 64; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const float v = w[((size_t)n * 64 + ci) * 9 + tap] * sc;
        const uint16_t hi = f2h(v);
        const int pr = ci >> 5, c32 = ci & 31, tt = tap / 3, tl = tap % 3;
        unsigned char* stg = s.data() + (size_t)(pr * 3 + tt) * C3_STAGE_BYTES;
        // hi part of slice (c32 >> 4) of the pair: [tl][kc][cout][8 fp16], kc = 8-channel chunk within the slice
        unsigned char* hp = stg + (size_t)(c32 >> 4) * C3_PART_BYTES;
        memcpy(hp + (((size_t)tl * 2 + ((c32 >> 3) & 1)) * 128 + n) * 16 + (c32 & 7) * 2, &hi, 2);
        // lo part of the pair: [tl][kc16][cout][16 e4m3], kc16 = 16-channel chunk
        unsigned char* lp = stg + (size_t)2 * C3_PART_BYTES;
        lp[(((size_t)tl * 2 + (c32 >> 4)) * 128 + n) * 16 + (c32 & 15)] = f2e4m3(ldexpf(v - h2f(hi), c->act2_lo_shift));   // its operand is act2 x 2^-shift
      }
  if (!c->d_w3_split) WW_CHECK(c, cudaMalloc((void**)&c->d_w3_split, s.size()));
  WW_CHECK(c, cudaMemcpy(c->d_w3_split, s.data(), s.size(), cudaMemcpyHostToDevice));

  const Geom g = make_geom(c);
  std::vector<uint32_t> m((size_t)g.T3 * 4, 0u);
  for (int t = 0; t < g.T3; ++t)
    for (int r = 0; r < 128; ++r) {
      const int pidx = g.P + 128 * t + r;            // padded index of output row r of tile t
      const int row = pidx / g.P, y = row - 1, x = pidx - row * g.P - 1;
      if (y >= 0 && y < g.H && x >= 0) m[(size_t)t * 4 + r / 32] |= 1u << (r % 32);
    }
  if (c->d_tc_mask) cudaFree(c->d_tc_mask);
  WW_CHECK(c, cudaMalloc((void**)&c->d_tc_mask, m.size() * 4));
  WW_CHECK(c, cudaMemcpy(c->d_tc_mask, m.data(), m.size() * 4, cudaMemcpyHostToDevice));
  return WW_OK;
}

int ww_launch_conv3_tc(ww_ctx* c, int B, const Geom& g, cudaStream_t st) {
  const size_t smem = conv3_smem_bytes(g.nsl3, g.nst3);
  if (smem > 227 * 1024) {
    c->set_error("conv3_tc: frame count too large for the shared-memory tiles (use WW_CONV_FP32)");
    return WW_ERR_INVALID;
  }
  if (smem > c->c3_smem_conf) {      // per context (= per device), not per process
    WW_CHECK(c, cudaFuncSetAttribute(conv3_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv3_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WW_CHECK(c, cudaFuncSetAttribute(conv3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->c3_smem_conf = smem;
  }
  Conv3Params p;
  p.act2 = c->ws_act2_h; p.act2_8 = c->ws_act2_8; p.w3s = reinterpret_cast<const unsigned char*>(c->d_w3_split); p.inv_scale = c->w3_inv_scale / c->act2_scale; p.b3 = c->w["conv3.bias"]; p.mask = c->d_tc_mask;
  p.pool_part = c->pool_cur; p.relu_bits = c->tc_relu_bits; p.B = B; p.g = g;
  static long long* d_trace = nullptr;
  const bool tracing = getenv("WW_TC_TRACE") != nullptr;
  if (tracing && !d_trace) { cudaMalloc((void**)&d_trace, 48 * 16 * 8); }
  if (tracing) cudaMemset(d_trace, 0, 48 * 16 * 8);
  p.trace = tracing ? d_trace : nullptr;
  const int n_items = (int)(((long long)B * g.T3 + g.G - 1) / g.G);
  const int grid = std::min(c->sm_count, n_items);
  ProfScope prof(c, WW_STAGE_CONV3, st);
  const bool fp16 = c->cfg.conv_mode == WW_CONV_FP16;
  if (p.relu_bits) {
    if (fp16) conv3_kernel<1, true><<<grid, C3_THREADS, smem, st>>>(p);
    else conv3_kernel<2, true><<<grid, C3_THREADS, smem, st>>>(p);
  } else if (fp16) conv3_kernel<1, false><<<grid, C3_THREADS, smem, st>>>(p);
  else conv3_kernel<2, false><<<grid, C3_THREADS, smem, st>>>(p);
  WW_LAUNCH_CHECK(c);
  if (tracing) {
    long long h[48 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "conv3 trace: ld_start | mma: t_empty a_full0 a_full1 a_full2 a_full3 issued | epi: t_full t_free\n");
    for (int i = 0; i < 14; ++i) {
      fprintf(stderr, "grp %2d:", i);
      for (int k = 0; k < 9; ++k) fprintf(stderr, " %8lld", h[i * 16 + k] ? h[i * 16 + k] - h[0] : -1);
      fprintf(stderr, "\n");
    }
  }
  return WW_OK;
}

// in_pad: the chunk's log-mel images in the zero-padded pixel-linear layout (ww_conv_tc_logmel_out / ww_launch_pad_logmel)
int ww_launch_conv_tc(ww_ctx* c, const float* in_pad, int B, cudaStream_t st) {
  if (B <= 0) return WW_OK;
  const Geom g = make_geom(c);
  c->n_pool_part = g.n_groups;
  int rc = ww_launch_conv12_tc(c, in_pad, B, g, st);
  if (rc) return rc;
  return ww_launch_conv3_tc(c, B, g, st);
}

// Stage 1 + 2b: the frame-streaming kernels.
//
// Every kernel here reads each frame from HBM exactly once.  A frame is a flat stream of pixels
// cut into tiles (8192 px, or 2048 px for small frames); a launch is one resident wave of
// long-lived CTAs, each walking a contiguous run of (frame, tile) items whose bytes arrive in
// shared memory through 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) behind mbarriers.
//
//   count12_kernel<BITS>   above-noise counts only (no image output): producer warp + full/empty
//                          ring, compares on 16x2 SIMD lanes without extracting the pixels.
//   streamx_kernel<BITS>   uint16 difference image and/or decoded pixels: same skeleton; the
//                          previous frame's tile is carried in registers, so the difference
//                          costs no second read (one halo tile per run of frames).
//   stream_kernel<...>     general template: float32 / float64 difference outputs, 2048-px tiles,
//                          CTA-wide barrier per item.
//   stream_generic_kernel  shapes the TMA path cannot take (P % 32 != 0).
//
// Everything is integer arithmetic on values the reference holds as integer-valued float64
// (scripts/process_videos.py:670-674, :397-399, :759), so results are bit-exact.
#include <cstdlib>

#include "ff_detect_core.cuh"

namespace ff {
namespace {

constexpr int kStages = 4;

struct StreamParams {
  const uint8_t* frames;
  const uint8_t* halo;
  int64_t frame_bytes;
  int64_t px_per_frame;
  int n_frames;
  int tiles_per_frame;
  int64_t items_per_cta;   // (tile, frame) work items per CTA, tile-major order
  const int32_t* bg_dev;
  int empty_thr;
  int diff_thr;
  const uint8_t* skip;
  int32_t* partial;
  void* diff_out;
  uint16_t* decoded_out;
  int unit_frames;         // streamx_kernel with a retained difference: frames per unit (0: contiguous runs per CTA)
  int64_t n_units;         // ... and the number of (frame segment, tile) units, dealt to the CTAs round-robin
  int idle_ns;             // how long an idle detector warp sleeps between looks at the queue
  int slab_shift;          // range_kernel: 2^slab_shift consecutive items per slab (slabs are dealt to the CTAs round-robin)
  int slab_step_frames;    // (gridDim.x << slab_shift) = slab_step_frames * tiles_per_frame + slab_step_tiles:
  int slab_step_tiles;     // how far (frame, tile) moves from one slab of a CTA to its next
  int pdl;                 // host side only: launch with the programmatic-dependent-launch attribute
  uint32_t neg_one;        // 0xFFFFFFFF, as a value the compiler cannot see: "c - x" written as x * neg_one + c is an
                           // IMAD on the FMA pipe instead of one more instruction on the ALU pipe (streamx_kernel)
};

template <int BITS>
__device__ __forceinline__ void load_group(const uint8_t* stage, int g, int (&v)[8]) {
  if (BITS == 12) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage) + 3 * g;
    decode12x8(w[0], w[1], w[2], v);
  } else if (BITS == 16) {
    const uint4 q = reinterpret_cast<const uint4*>(stage)[g];
    v[0] = q.x & 0xFFFF; v[1] = q.x >> 16; v[2] = q.y & 0xFFFF; v[3] = q.y >> 16;
    v[4] = q.z & 0xFFFF; v[5] = q.z >> 16; v[6] = q.w & 0xFFFF; v[7] = q.w >> 16;
  } else {
    const uint2 q = reinterpret_cast<const uint2*>(stage)[g];
    v[0] = q.x & 0xFF; v[1] = (q.x >> 8) & 0xFF; v[2] = (q.x >> 16) & 0xFF; v[3] = q.x >> 24;
    v[4] = q.y & 0xFF; v[5] = (q.y >> 8) & 0xFF; v[6] = (q.y >> 16) & 0xFF; v[7] = q.y >> 24;
  }
}

template <int DIFF>
__device__ __forceinline__ void store_diff(void* base, int64_t px, const int (&d)[8]) {
  if (DIFF == FF_DIFF_U16) {
    uint4 q;
    q.x = (uint32_t)d[0] | ((uint32_t)d[1] << 16);
    q.y = (uint32_t)d[2] | ((uint32_t)d[3] << 16);
    q.z = (uint32_t)d[4] | ((uint32_t)d[5] << 16);
    q.w = (uint32_t)d[6] | ((uint32_t)d[7] << 16);
    __stcs(reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + px), q);
  } else if (DIFF == FF_DIFF_F32) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(base) + px);
    __stcs(o, make_float4((float)d[0], (float)d[1], (float)d[2], (float)d[3]));
    __stcs(o + 1, make_float4((float)d[4], (float)d[5], (float)d[6], (float)d[7]));
  } else if (DIFF == FF_DIFF_F64) {
    double2* o = reinterpret_cast<double2*>(static_cast<double*>(base) + px);
#pragma unroll
    for (int j = 0; j < 4; ++j) __stcs(o + j, make_double2((double)d[2 * j], (double)d[2 * j + 1]));
  }
}

// Pixel-pair access for the float difference outputs: lane l of a warp takes pair l of 32
// consecutive pairs, so one store instruction covers a contiguous 512-byte (f64) / 256-byte
// (f32) line instead of 32 scattered 16-byte pieces.
template <int BITS>
__device__ __forceinline__ void load_pair(const uint8_t* stage, int j, int (&v)[2]) {
  if (BITS == 12) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage);
    const int byte = 3 * j;
    const uint32_t lo = w[byte >> 2];
    const uint32_t hi = w[(byte >> 2) + 1];            // may read past the tile: stays inside smem
    const uint32_t le = __funnelshift_r(lo, hi, (byte & 3) * 8);   // b0 | b1<<8 | b2<<16 | ..
    const uint32_t t = __byte_perm(le, 0u, 0x4012);                // b0<<16 | b1<<8 | b2
    v[0] = (int)(t >> 12);
    v[1] = (int)(t & 0xFFFu);
  } else if (BITS == 16) {
    const uint32_t q = reinterpret_cast<const uint32_t*>(stage)[j];
    v[0] = q & 0xFFFF;
    v[1] = q >> 16;
  } else {
    const uint32_t q = reinterpret_cast<const uint16_t*>(stage)[j];
    v[0] = q & 0xFF;
    v[1] = q >> 8;
  }
}

template <int DIFF>
__device__ __forceinline__ void store_diff_pair(void* base, int64_t px, int d0, int d1) {
  if (DIFF == FF_DIFF_F32) {
    __stcs(reinterpret_cast<float2*>(static_cast<float*>(base) + px), make_float2((float)d0, (float)d1));
  } else if (DIFF == FF_DIFF_F64) {
    __stcs(reinterpret_cast<double2*>(static_cast<double*>(base) + px), make_double2((double)d0, (double)d1));
  }
}

// COUNT: emit per-(frame,tile) above-noise counts.  DIFF: retained difference dtype.
// DECODED: also write decoded uint16 pixels.  K: groups per thread per tile.
template <int BITS, bool COUNT, int DIFF, bool DECODED, int K>
__global__ void __launch_bounds__(kThreads) stream_kernel(const StreamParams p) {
  constexpr int kGroups = K * kThreads;
  constexpr int kStageBytes = kGroups * BITS;

  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);

  const int tid = threadIdx.x;

  uint64_t policy = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncthreads();

  int bg = 0, cthr = 0;
  griddep_wait();                      // the background scalar may come from the kernel right before this one
  if (COUNT || DIFF != FF_DIFF_NONE) {
    bg = __ldg(p.bg_dev);
    const int ethr = p.empty_thr >= 0 ? p.empty_thr : max(10, bg >> 1);
    cthr = bg + ethr;  // max(x-bg,0) > ethr  <=>  x > bg + ethr   (ethr >= 0)
  }
  const uint32_t c12 = (uint32_t)min(cthr, 4095);      // 12-bit pixels never exceed 4095
  const uint32_t nk_hi = ~((c12 << 20) | 0xFFFFFu);   // complemented bounds for add_gt
  const uint32_t nk_lo = ~((c12 << 8) | 0xFFu);
  const uint32_t ncthr = ~(uint32_t)cthr;
  // 16x2 SIMD (DPX) constants for the packed-12 / uint16-difference variant: every value on that
  // path fits a signed 16-bit lane (pixels and bg <= 4095, |differences| <= 4095).
  constexpr bool kSimd = (BITS == 12 && DIFF == FF_DIFF_U16);
  const uint32_t dup = 0x00010001u;
  const uint32_t s_ncthr = (uint32_t)(-(int)c12 & 0xFFFF) * dup;          // -cthr per lane
  const uint32_t s_nbg = (uint32_t)(-min(bg, 4095) & 0xFFFF) * dup;       // -bg per lane
  const int tm1 = min(max(p.diff_thr, 0) - 1, 8190);                      // thr-1 (thr > 4095 keeps nothing)
  const uint32_t s_k = (uint32_t)((1 - tm1) & 0xFFFF) * dup;              // relu(d - tm1) = relu((d-1) + (1-tm1))

  // Work item = 8-pixel group (16-byte vector stores) or, for float difference outputs, one
  // pixel pair per lane (see load_pair).  prev[] holds the background-subtracted previous frame
  // of this thread's pixels, two uint16 per register.
  constexpr bool kPair = (DIFF == FF_DIFF_F32 || DIFF == FF_DIFF_F64);
  constexpr int kItemPx = kPair ? 2 : kGroupPx;
  constexpr int kItems = K * kGroupPx / kItemPx;        // items per thread per tile
  uint32_t prev[kItems][kItemPx / 2];

  // The (tile, frame) grid is cut tile-major into equal runs, one per CTA, so the launch is
  // exactly one resident wave whatever the tile count; a run that crosses a tile boundary is
  // processed as two segments.  `git` numbers the stage-ring slots across segments.
  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  int64_t work = (int64_t)blockIdx.x * p.items_per_cta;
  const int64_t work_end = min(work + p.items_per_cta, total_work);
  uint32_t git = 0;
  while (work < work_end) {
  const int tile = (int)(work / p.n_frames);
  const int f_begin = (int)(work - (int64_t)tile * p.n_frames);
  const int f_end = (int)min((int64_t)p.n_frames, (int64_t)f_begin + (work_end - work));
  work += f_end - f_begin;

  const int64_t tile_px0 = (int64_t)tile * (kGroups * kGroupPx);
  const int tile_groups = (int)min((int64_t)kGroups, (p.px_per_frame - tile_px0) / kGroupPx);
  const uint32_t tile_bytes = (uint32_t)tile_groups * BITS;
  const int64_t tile_off = (int64_t)tile * kStageBytes;
  const int tile_items = tile_groups * (kGroupPx / kItemPx);

  // Frame feeding prev[] before the first frame of this segment (difference modes only).
  const uint8_t* halo_ptr = nullptr;
  if (DIFF != FF_DIFF_NONE) {
    int hf = f_begin - 1;
    if (p.skip != nullptr)
      while (hf >= 0 && p.skip[hf]) --hf;
    halo_ptr = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
  }
  const int has_halo = halo_ptr != nullptr ? 1 : 0;
  const int n_items = (f_end - f_begin) + has_halo;

  auto src_of = [&](int item) -> const uint8_t* {
    const int f = item - has_halo;
    const uint8_t* fr = f < 0 ? halo_ptr : p.frames + (int64_t)(f_begin + f) * p.frame_bytes;
    return fr + tile_off;
  };

  if (tid == 0) {   // every stage is free here: the previous segment drained the ring
    const int pre = min(kStages - 1, n_items);
    for (int i = 0; i < pre; ++i) {
      const uint32_t ps = (git + i) % kStages;
      mbar_arrive_expect_tx(&full[ps], tile_bytes);
      bulk_g2s(smem + ps * kStageBytes, src_of(i), tile_bytes, &full[ps], policy);
    }
  }

#pragma unroll
  for (int k = 0; k < kItems; ++k)
#pragma unroll
    for (int j = 0; j < kItemPx / 2; ++j) prev[k][j] = kSimd ? 0xFFFFFFFFu : 0u;   // SIMD path keeps ~sub
  bool have_prev = false;

  for (int it = 0; it < n_items; ++it, ++git) {
    const int s = git % kStages;
    const uint32_t parity = (git / kStages) & 1u;
    if (tid == 0) {
      const int nx = it + kStages - 1;  // its stage was drained in the previous iteration
      if (nx < n_items) {
        const uint32_t ns = (git + kStages - 1) % kStages;
        mbar_arrive_expect_tx(&full[ns], tile_bytes);
        bulk_g2s(smem + ns * kStageBytes, src_of(nx), tile_bytes, &full[ns], policy);
      }
    }
    mbar_wait(&full[s], parity);

    const uint8_t* stage = smem + s * kStageBytes;
    const bool is_halo = it < has_halo;
    const int f = f_begin + it - has_halo;
    const bool skipped = !is_halo && p.skip != nullptr && p.skip[f] != 0;
    const bool emit_diff = (DIFF != FF_DIFF_NONE) && !is_halo;
    const bool diff_valid = have_prev && !skipped;
    int cnt = 0;
    uint32_t acc2 = 0;   // SIMD path: two 16-bit counters

#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int g = tid + k * kThreads;
      if (g < tile_items) {
        if (BITS == 12 && COUNT && DIFF == FF_DIFF_NONE && !DECODED) {   // count in place, no extraction
          const uint32_t* w = reinterpret_cast<const uint32_t*>(stage) + 3 * g;
          count12x8(w[0], w[1], w[2], nk_hi, nk_lo, cnt);
          continue;
        }
        if (kSimd) {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(stage) + 3 * g;
          uint32_t x[4], o[4];
          decode12x8_16x2(w[0], w[1], w[2], x);
          const int64_t px = (int64_t)f * p.px_per_frame + tile_px0 + (int64_t)g * kGroupPx;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (COUNT) acc2 += __viaddmin_s16x2_relu(x[j], s_ncthr, dup);           // [x > cthr] per lane
            const uint32_t sub2 = __viaddmax_s16x2_relu(x[j], s_nbg, 0x80008000u);  // max(x - bg, 0) (see kLaneMin2)
            const uint32_t e2 = __viaddmax_s16x2(sub2, prev[k][j], 0x80008000u);    // sub + ~prev = d - 1
            const uint32_t r2 = __viaddmax_s16x2_relu(e2, s_k, 0x80008000u);        // relu(d - (thr-1))
            const uint32_t m2 = __vimin_s16x2_relu(r2, dup);                         // [d >= thr]
            o[j] = diff_valid ? r2 + m2 * (uint32_t)tm1 : 0u;                       // d where d >= thr, else 0
            if (!skipped) prev[k][j] = ~sub2;
          }
          if (DECODED && !is_halo)
            __stcs(reinterpret_cast<uint4*>(p.decoded_out + px), make_uint4(x[0], x[1], x[2], x[3]));
          if (emit_diff)
            __stcs(reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.diff_out) + px), make_uint4(o[0], o[1], o[2], o[3]));
          continue;
        }
        int v[kItemPx];
        if (kPair) {
          int v2[2];
          load_pair<BITS>(stage, g, v2);
          v[0] = v2[0];
          v[1] = v2[1];
        } else {
          int v8[8];
          load_group<BITS>(stage, g, v8);
#pragma unroll
          for (int j = 0; j < kItemPx; ++j) v[j] = v8[j];
        }
        const int64_t px = (int64_t)f * p.px_per_frame + tile_px0 + (int64_t)g * kItemPx;
        if (COUNT) {
#pragma unroll
          for (int j = 0; j < kItemPx; ++j) add_gt(cnt, (uint32_t)v[j], ncthr);
        }
        if (DECODED && !is_halo) {
          if (kPair) {
            __stcs(reinterpret_cast<uint32_t*>(p.decoded_out + px), (uint32_t)v[0] | ((uint32_t)v[1] << 16));
          } else {
            uint4 q;
            q.x = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
            q.y = (uint32_t)v[2 % kItemPx] | ((uint32_t)v[3 % kItemPx] << 16);
            q.z = (uint32_t)v[4 % kItemPx] | ((uint32_t)v[5 % kItemPx] << 16);
            q.w = (uint32_t)v[6 % kItemPx] | ((uint32_t)v[7 % kItemPx] << 16);
            __stcs(reinterpret_cast<uint4*>(p.decoded_out + px), q);
          }
        }
        if (DIFF != FF_DIFF_NONE) {
          int d[kItemPx];
          uint32_t cur[kItemPx / 2];
#pragma unroll
          for (int j = 0; j < kItemPx / 2; ++j) {
            const int s0 = max(v[2 * j] - bg, 0);
            const int s1 = max(v[2 * j + 1] - bg, 0);
            int d0 = s0 - (int)(prev[k][j] & 0xFFFFu);
            int d1 = s1 - (int)(prev[k][j] >> 16);
            d[2 * j] = (diff_valid && d0 >= p.diff_thr) ? d0 : 0;
            d[2 * j + 1] = (diff_valid && d1 >= p.diff_thr) ? d1 : 0;
            cur[j] = (uint32_t)s0 | ((uint32_t)s1 << 16);
          }
          if (emit_diff) {
            if (kPair) {
              store_diff_pair<DIFF>(p.diff_out, px, d[0], d[1]);
            } else {
              int d8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) d8[j] = d[j % kItemPx];
              store_diff<DIFF>(p.diff_out, px, d8);
            }
          }
          if (!skipped) {
#pragma unroll
            for (int j = 0; j < kItemPx / 2; ++j) prev[k][j] = cur[j];
          }
        }
      }
    }
    if (kSimd) cnt += (int)(acc2 & 0xFFFFu) + (int)(acc2 >> 16);
    if (DIFF != FF_DIFF_NONE && !skipped) have_prev = true;

    if (COUNT && !is_halo) {   // one partial count per (frame, tile, warp); ff_detect sums them
      cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
      if ((tid & 31) == 0)
        p.partial[((int64_t)f * p.tiles_per_frame + tile) * kWarpsPerCta + (tid >> 5)] = cnt;
    }
    __syncthreads();  // stage s drained by every thread
  }
  }  // segments
}

// ---- count-only kernel (packed 12-bit is the headline configuration; 8- and 16-bit share it) -----
// No output but the above-noise counts, so the pixels are never extracted.  Compared with the
// general template: a dedicated producer warp feeds a full/empty mbarrier ring (consumer warps
// never meet at a CTA-wide barrier and can run up to kCountStages items apart); every thread
// owns 32 consecutive pixels = 48 bytes = three conflict-free LDS.128 (stride 48 B: each
// quarter-warp covers all 32 banks); and the compare runs on 16x2 SIMD lanes (DPX):
//   A lanes = (b0<<8 | b1) = hi<<4 | nibble   hi > c  <=>  A > (c<<4 | 15)   (junk below the field
//                                                                              is dominated)
//   B lanes = (b1<<8 | b2) & 0x0FFF = lo       lo > c  directly
// 4 PRMT + 2 LOP3 + 6 DPX + 2 IADD3 per 8 pixels (the carry-chain form needs 20).
// 16-bit pixels already are 16x2 lanes (unsigned compare: max, add, min); 8-bit pixels are widened
// with two PRMTs per word.  Those two read 16-byte pieces t, t+256, ... (conflict-free LDS.128).
constexpr int kCountThreads = kThreads + 32;              // 8 consumer warps + 1 producer warp
constexpr int kCountCtasPerSm = 2;

#ifndef FF_COUNT_FMA_ADDS
#define FF_COUNT_FMA_ADDS 1
#endif
// acc + the number of the 8 packed-12 pixels in (w0, w1, w2) above the threshold, per 16-bit lane.  With
// FF_COUNT_FMA_ADDS (default) the four flag words are added as x * one + acc (one = 1 at run time: IMADs on the
// otherwise idle FMA pipe instead of two 3-input adds on the ALU pipe, which ncu shows 70 % busy; measured
// C3 1.1053 -> 1.1014 ms, C2 0.5649 -> 0.5644 ms; -DFF_COUNT_FMA_ADDS=0 builds the 3-input adds).
__device__ __forceinline__ uint32_t count12x8_simd(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t kA2,
                                                   uint32_t nkA2, uint32_t nc2, uint32_t acc = 0, uint32_t one = 1) {
  const uint32_t one2 = 0x00010001u;
  const uint32_t a0 = __byte_perm(w0, w1, 0x3401);                 // b1 b0 | b4 b3   (LSB first)
  const uint32_t a1 = __byte_perm(w1, w2, 0x5623);                 // b7 b6 | b10 b9
  const uint32_t b0 = __byte_perm(w0, w1, 0x4512) & 0x0FFF0FFFu;   // b2 b1 | b5 b4
  const uint32_t b1 = __byte_perm(w1, w2, 0x6734) & 0x0FFF0FFFu;   // b8 b7 | b11 b10
  // [a > kA] = min(max(a, kA) - kA, 1) on unsigned lanes;  [b > c] = relu(min(b - c, 1)) on signed lanes
  const uint32_t fa0 = __viaddmin_u16x2(__vimax3_u16x2(a0, kA2, kA2), nkA2, one2);
  const uint32_t fa1 = __viaddmin_u16x2(__vimax3_u16x2(a1, kA2, kA2), nkA2, one2);
  const uint32_t fb0 = __viaddmin_s16x2_relu(b0, nc2, one2);
  const uint32_t fb1 = __viaddmin_s16x2_relu(b1, nc2, one2);
#if FF_COUNT_FMA_ADDS
  // (inline PTX: written in C the compiler factors `one` out and is back to 3-input adds)
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(fa0), "r"(one));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(fa1), "r"(one));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(fb0), "r"(one));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(fb1), "r"(one));
  return acc;
#else
  return acc + (fa0 + fa1) + (fb0 + fb1);
#endif
}

template <int BITS, int kCountStages>
__global__ void __launch_bounds__(kCountThreads) count12_kernel(const StreamParams p) {
  constexpr int kTileBytes = 4 * kThreads * BITS;
  constexpr int kMaxPx = BITS == 8 ? 255 : (BITS == 12 ? 4095 : 65535);
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCountStages * kTileBytes);
  uint64_t* empty = full + kCountStages;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kCountStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWarpsPerCta);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  const int64_t work0 = (int64_t)blockIdx.x * p.items_per_cta;
  const int n_items = (int)(min(work0 + p.items_per_cta, total_work) - work0);
  if (n_items <= 0) return;
  // Item order is frame-major (tile fastest): nothing is carried from frame to frame here, so
  // every CTA reads ONE contiguous span of the clip (measured +2.5 % over the tile-major order
  // the difference kernels need).
  int f = (int)(work0 / p.tiles_per_frame);
  int tile = (int)(work0 - (int64_t)f * p.tiles_per_frame);
  auto advance = [&]() {
    if (++tile == p.tiles_per_frame) { tile = 0; ++f; }
  };
  const int64_t groups_per_frame = p.px_per_frame / kGroupPx;

  if (warp == kWarpsPerCta) {            // ---- producer: one elected lane drives the TMA ring
    if ((tid & 31) == 0) {
      const uint64_t policy = policy_evict_first();
      for (int it = 0; it < n_items; ++it) {
        const int s = it % kCountStages;
        mbar_wait(&empty[s], ((it / kCountStages) & 1) ^ 1);   // first pass: fresh barriers pass at once
        const int64_t g0 = (int64_t)tile * (4 * kThreads);
        const uint32_t bytes = (uint32_t)min((int64_t)(4 * kThreads), groups_per_frame - g0) * (uint32_t)BITS;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(smem + s * kTileBytes, p.frames + (int64_t)f * p.frame_bytes + (int64_t)tile * kTileBytes,
                 bytes, &full[s], policy);
        advance();
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------------
  griddep_wait();                      // the background scalar may come from the kernel right before this one
  const int bg = __ldg(p.bg_dev);
  const int ethr = p.empty_thr >= 0 ? p.empty_thr : max(10, bg >> 1);
  const uint32_t c = (uint32_t)min((int64_t)bg + ethr, (int64_t)kMaxPx);       // no pixel exceeds kMaxPx
  const uint32_t kA = BITS == 12 ? ((c << 4) | 15u) : c;
  const uint32_t kA2 = kA * 0x00010001u;
  const uint32_t nkA2 = ((0x10000u - kA) & 0xFFFFu) * 0x00010001u;
  const uint32_t nc2 = ((0x10000u - c) & 0xFFFFu) * 0x00010001u;
  const uint32_t one_rt = p.neg_one * p.neg_one;       // 1, unknown to the compiler (FF_COUNT_FMA_ADDS)
  const int my_group = tid * 4;
  // running values instead of a modulo, a 64-bit min and a 64-bit index product per item (see range_kernel)
  const int last_groups = (int)(groups_per_frame - (int64_t)(p.tiles_per_frame - 1) * (4 * kThreads));
  int32_t* out = p.partial + ((int64_t)f * p.tiles_per_frame + tile) * kWarpsPerCta + warp;
  int s = 0;
  uint32_t ph = 0;

  for (int it = 0; it < n_items; ++it) {
    mbar_wait(&full[s], ph);
    const int tile_groups = tile + 1 == p.tiles_per_frame ? last_groups : 4 * kThreads;
    uint32_t acc = 0;
    if (BITS == 12) {
      if (my_group < tile_groups) {        // groups per frame are a multiple of 4: all four or none
        const uint4* q = reinterpret_cast<const uint4*>(smem + s * kTileBytes + tid * 48);
        const uint4 q0 = q[0], q1 = q[1], q2 = q[2];
        acc = count12x8_simd(q0.x, q0.y, q0.z, kA2, nkA2, nc2, acc, one_rt);
        acc = count12x8_simd(q0.w, q1.x, q1.y, kA2, nkA2, nc2, acc, one_rt);
        acc = count12x8_simd(q1.z, q1.w, q2.x, kA2, nkA2, nc2, acc, one_rt);
        acc = count12x8_simd(q2.y, q2.z, q2.w, kA2, nkA2, nc2, acc, one_rt);
      }
    } else {
      const uint32_t one2 = 0x00010001u;
      const int pieces = tile_groups * BITS / 16;                      // 16-byte pieces in this tile
      const uint4* q = reinterpret_cast<const uint4*>(smem + s * kTileBytes);
#pragma unroll
      for (int k = 0; k < kTileBytes / 16 / kThreads; ++k) {
        const int i = tid + k * kThreads;
        if (i < pieces) {
          const uint4 v = q[i];
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (BITS == 16) {
              acc += __viaddmin_u16x2(__vimax3_u16x2(w4[j], kA2, kA2), nkA2, one2);
            } else {
              acc += __viaddmin_s16x2_relu(__byte_perm(w4[j], 0u, 0x4140), nc2, one2) +
                     __viaddmin_s16x2_relu(__byte_perm(w4[j], 0u, 0x4342), nc2, one2);
            }
          }
        }
      }
    }
    int cnt = (int)(acc & 0xFFFFu) + (int)(acc >> 16);
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);      // also orders every lane's smem reads before the release
    if ((tid & 31) == 0) {
      mbar_arrive(&empty[s]);
      *out = cnt;
    }
    out += kWarpsPerCta;                            // frame-major items: (f, tile) -> the next eight counts
    advance();
    if (++s == kCountStages) {
      s = 0;
      ph ^= 1u;
    }
  }
}

int sm_count_cached();

// The streaming kernels are ONE wave of long-lived CTAs, `ctas` per SM by design (bytes in flight per SM are
// tuned).  The hardware would happily put more of them on an SM when another small kernel (prep, the merge of
// the previous clip on its side stream, a PDL predecessor) occupies some other SM at launch time - and the
// doubled-up SMs then decide the kernel's time (measured: C4 uint16 1.55 -> 2.26 ms next to a 2-CTA merge).
// Asking for at least 1/(ctas+1) of an SM's shared memory makes `ctas` per SM the only possible placement.
constexpr int kSmemPerSm = 227 * 1024;
constexpr int pinned_smem(int needed, int ctas) {
  return needed > kSmemPerSm / (ctas + 1) + 1024 ? needed : kSmemPerSm / (ctas + 1) + 1024;
}

template <int BITS, int kCountStages>
int launch_count12(StreamParams p, cudaStream_t st) {
  constexpr int kSmem = pinned_smem(kCountStages * (4 * kThreads * BITS) + 2 * kCountStages * 8, kCountCtasPerSm);   // 2 per SM
  auto kern = count12_kernel<BITS, kCountStages>;
  static PerDeviceInt cache;
  int ctas = 1;
  int rc = cache.get([&](int* v) -> int {
    FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int occ = 0;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCountThreads, kSmem));
    // Bytes in flight per SM decide the achieved bandwidth, and MORE is not better: measured on
    // B200 (C2/C3, GB/s) 24 KB 5770, 48 KB 6740, 72 KB 7030-7160, 120 KB 6620, 144 KB 6500,
    // 168 KB 6320 (profiles/r01_count12_sweep.txt).  2 CTAs x (4-1) stages x 12 KB = 72 KB.
    int cap = kCountCtasPerSm;
    if (const char* e = getenv("FF_COUNT12_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;   // tuning knob
    if (occ > cap) occ = cap;
    *v = occ > 0 ? occ : 1;
    return FF_OK;
  }, &ctas);
  if (rc != FF_OK) return rc;
  const int64_t wave = (int64_t)sm_count_cached() * ctas;
  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  p.items_per_cta = (total_work + wave - 1) / wave;
  if (p.items_per_cta < 1) p.items_per_cta = 1;
  const int64_t grid = (total_work + p.items_per_cta - 1) / p.items_per_cta;
  return launch_kernel(kern, dim3((unsigned)grid), dim3(kCountThreads), kSmem, st, p.pdl != 0, p);
}

// ---- the whole range in ONE kernel: counts + empty-frame decision + detection + exit min + truncation ----
// count12_kernel's ring and compare, plus:
//   * the (frame, tile) items, in frame-major order, are cut into SLABS of 2^k consecutive items and the
//     slabs are dealt to the CTAs round-robin: all CTAs sweep through the clip together (a sliding window of
//     a few MB over all DRAM channels) and - what matters - the frames that hold a flame, which are
//     consecutive in time, are spread over ALL CTAs instead of landing in a handful of them;
//   * a SEGMENT is the part of a slab that lies in one frame.  Every consumer warp keeps the above-noise
//     count of its segment and adds it to the segment's word in SHARED memory ({warps arrived | count}, one
//     shared-memory atomic per warp and segment - not one partial-count store per warp and tile).  The warp
//     that arrives last hands (frame, count) to the CTA's DETECTOR WARP through a shared-memory queue.
//     Consumers never touch global memory besides the tiles: a global atomic whose result is needed costs
//     1.5-3 us under a saturated memory system, and putting it into the streaming loop (first version)
//     slowed the kernel by 20-150 %;
//   * the detector warp (warp 9) drains the queue, up to 32 segments at once, one per lane: it adds
//     {1 segment, count} to the frame's 64-bit word in global memory; the lane whose segment completes
//     the frame knows the frame's count, writes count_out and decides is_empty_frame
//     (scripts/process_videos.py:759-763).  Non-empty frames are then resolved by the whole warp
//     (detect_one_frame) - all of it while the consumers keep streaming;
//   * the detector warp of the CTA that finishes last truncates the range at its first exit frame and
//     publishes the block to the peers of a range-sharded run (range_tail).
// No partial-count array, no detect / truncate launches; the frames are still read from HBM exactly once
// (the two centre rows a detection needs come back through L2: 3 KB per flame frame).
constexpr int kDetWarps = 4;                              // detector warps per CTA (a detection is a chain of latencies: ~30 us)
constexpr int kFusedThreads = kThreads + 32 + 32 * kDetWarps;   // 8 consumer warps + producer warp + detector warps
constexpr int kQueue = 128;                               // (frame, count) entries waiting for the detector warps
constexpr int kSegRing = 16;                              // segment words in flight (warps drift < ring depth items apart)

// An ITEM is kItemKB KiB of a frame's stored bytes whatever the bit depth (8-bit: 12288 / 24576 pixels, 12-bit:
// 8192 / 16384, 16-bit: 6144 / 12288): the per-item costs - barrier round trip, warp reduction, lane-0
// bookkeeping - are paid per byte moved, not per pixel (8-KB items of 8-bit pixels ran at 0.82 of the copy rate).
template <int BITS, int kCountStages, int kItemKB>
__global__ void __launch_bounds__(kFusedThreads) range_kernel(const StreamParams p, const DetectParams d) {
  constexpr int kTileBytes = kItemKB * 1024;
  constexpr int kTileGroups = kTileBytes / BITS;          // 8-pixel groups per item
  static_assert(kTileBytes % (48 * kThreads) == 0, "an item is a whole number of 48-byte thread slices");
  constexpr int kMaxPx = BITS == 8 ? 255 : (BITS == 12 ? 4095 : 65535);
  static_assert(kCountStages < kSegRing, "segment ring must outlast the tile ring");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCountStages * kTileBytes);
  uint64_t* empty = full + kCountStages;
  volatile int* q_f = reinterpret_cast<volatile int*>(empty + kCountStages);   // queue: frame (-1 = slot empty) ...
  volatile int* q_cnt = q_f + kQueue;                                           // ... and the segment's count
  int* q_ctl = const_cast<int*>(q_cnt) + kQueue;          // [0] entries claimed, [1] consumer warps done, [2] detector warps done,
                                                          // [4 + w] next entry of detector warp w (entry e belongs to warp e % kDetWarps)
  unsigned* seg = reinterpret_cast<unsigned*>(q_ctl + 4 + kDetWarps);           // [kSegRing] {warps arrived : 8 | count : 24}
  uint8_t* det_smem = reinterpret_cast<uint8_t*>(seg + kSegRing);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kCountStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWarpsPerCta);
    }
    fence_mbar_init();
    q_ctl[0] = q_ctl[1] = q_ctl[2] = 0;
    for (int w = 0; w < kDetWarps; ++w) q_ctl[4 + w] = w;
  }
  if (tid < kQueue) q_f[tid] = -1;
  if (tid < kSegRing) seg[tid] = 0;
  __syncthreads();

  const int T = p.tiles_per_frame;
  const int sh = p.slab_shift;
  const int64_t S = (int64_t)1 << sh;
  const int64_t total_work = (int64_t)T * p.n_frames;
  const int64_t n_slabs = (total_work + S - 1) >> sh;        // gridDim.x <= n_slabs: every CTA has work
  const int64_t groups_per_frame = p.px_per_frame / kGroupPx;
  // (frame, tile) of this CTA's first slab; later slabs are reached by stepping - no divisions in the loops
  int f0, tile0;
  {
    const uint32_t first = (uint32_t)((int64_t)blockIdx.x << sh);   // < 2^31, see the launcher
    f0 = (int)(first / (uint32_t)T);
    tile0 = (int)(first - (uint32_t)f0 * (uint32_t)T);
  }

  if (warp == kWarpsPerCta) {            // ---- producer: one elected lane drives the TMA ring
    if (lane == 0) {
      const uint64_t policy = policy_evict_first();
      uint32_t it = 0;
      int fs = f0, ts = tile0;
      for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const int64_t i0 = slab << sh, i1 = min(i0 + S, total_work);
        int64_t f = fs;
        int tile = ts;
        fs += p.slab_step_frames;
        ts += p.slab_step_tiles;
        if (ts >= T) { ts -= T; ++fs; }
        for (int64_t i = i0; i < i1; ++i, ++it) {
          const int s = it % kCountStages;
          mbar_wait(&empty[s], ((it / kCountStages) & 1) ^ 1);   // first pass: fresh barriers pass at once
          const int64_t g0 = (int64_t)tile * kTileGroups;
          const uint32_t bytes = (uint32_t)min((int64_t)kTileGroups, groups_per_frame - g0) * (uint32_t)BITS;
          mbar_arrive_expect_tx(&full[s], bytes);
          bulk_g2s(smem + s * kTileBytes, p.frames + f * p.frame_bytes + (int64_t)tile * kTileBytes, bytes, &full[s], policy);
          if (++tile == T) { tile = 0; ++f; }
        }
      }
    }
    return;
  }

  griddep_wait();                        // scalars, first-exit word and workspace come from prep_kernel
  const int bg_raw = __ldg(p.bg_dev);

  if (warp > kWarpsPerCta) {             // ---- detector warps
    const int dw = warp - kWarpsPerCta - 1;
    const int thr_floor = d.threshold_dev != nullptr ? __ldg(d.threshold_dev) : d.threshold_floor;
    const RowSpan rs = centre_row_span<BITS>(d.height, d.width);
    uint8_t* mine = det_smem + (size_t)dw * detect_warp_smem(d.width, BITS);
    int* prof = reinterpret_cast<int*>(mine);
    uint8_t* raw_cur = mine + (size_t)d.width * sizeof(int);
    uint8_t* raw_pri = raw_cur + d.raw_stride;
    unsigned long long* arrive = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(d.ws) + kWorkspaceHeader);
    volatile int* ctl = q_ctl;
    const unsigned full_mask = 0xFFFFFFFFu;
    int head = dw;                       // my next entry; mine are dw, dw + kDetWarps, ...
    // An idle detector warp must not poll often: 8 of them per SM looking at the queue every 100 ns took 17 % of
    // the issue slots from the consumers (C2 step 0.619 -> 0.584 ms).  Back off while idle, be quick after work.
    int idle_ns = 250;
    for (;;) {
      // the leading run of my filled queue slots, one per lane (kQueue / kDetWarps = 32 of mine fit the ring)
      const int slot = (head + lane * kDetWarps) % kQueue;
      const int fq = q_f[slot];
      const unsigned ready = __ballot_sync(full_mask, fq >= 0);
      const int n = __ffs((int)~ready) - 1;               // number of leading ones (-1: all 32)
      const int take = n < 0 ? 32 : n;
      if (take == 0) {
        int stop = 0;
        if (lane == 0) stop = ctl[1] == kWarpsPerCta && ctl[0] <= head;     // consumers done, nothing of mine claimed
        if (__shfl_sync(full_mask, stop, 0)) break;
        __nanosleep(idle_ns);
        idle_ns = min(2 * idle_ns, p.idle_ns);
        continue;
      }
      idle_ns = 250;
      bool need = false;
      if (lane < take) {
        const int cnt = q_cnt[slot];
        q_f[slot] = -1;
        // this segment joins its frame: the lane that completes the frame owns the frame's decision
        const int64_t s_lo = ((int64_t)fq * T) >> sh, s_hi = ((int64_t)(fq + 1) * T - 1) >> sh;
        const unsigned expected = (unsigned)(s_hi - s_lo + 1);
        unsigned long long* word = arrive + (int64_t)fq * kArriveStride;
        unsigned long long old = 0;
        if (expected > 1) old = atomicAdd(word, (1ull << 32) | (unsigned)cnt);     // (a frame inside one slab needs no word)
        if ((unsigned)(old >> 32) + 1u == expected) {
          if (expected > 1) *word = 0;                     // left zero for the next launch
          const int total = (int)((unsigned)old + (unsigned)cnt);
          if (d.count_out != nullptr) d.count_out[fq] = total;
          const bool skipped = d.skip != nullptr && d.skip[fq] != 0;
          need = !skipped && (int64_t)total >= d.min_signal_count;
          if (!need) d.pos_out[fq] = FF_POS_NONE;
        }
      }
      head += take * kDetWarps;
      __syncwarp();
      if (lane == 0) ctl[4 + dw] = head;                   // my slots may be refilled
      unsigned pending = __ballot_sync(full_mask, need);
      while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1;
        const int f = __shfl_sync(full_mask, fq, src);
        int skip_it = 0;
        if (lane == 0) skip_it = behind_global_exit(d, f);
        if (__shfl_sync(full_mask, skip_it, 0)) {          // another rank saw the exit before this frame: it will be dropped
          if (lane == 0) d.pos_out[f] = FF_POS_NONE;
          continue;
        }
        const int pos = detect_one_frame<BITS>(d, rs, f, false, bg_raw, thr_floor, prof, raw_cur, raw_pri, lane);
        if (lane == 0) commit_position(d, f, pos);
      }
    }
    // the detector warp that leaves last finishes the CTA's part of the range
    int last = 0;
    if (lane == 0) {
      __threadfence();
      last = atomicAdd(q_ctl + 2, 1) == kDetWarps - 1;
    }
    if (__shfl_sync(full_mask, last, 0)) range_tail(d, lane);
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------------
  const int ethr = p.empty_thr >= 0 ? p.empty_thr : max(10, bg_raw >> 1);
  const uint32_t c = (uint32_t)min((int64_t)bg_raw + ethr, (int64_t)kMaxPx);       // no pixel exceeds kMaxPx
  const uint32_t kA = BITS == 12 ? ((c << 4) | 15u) : c;
  const uint32_t kA2 = kA * 0x00010001u;
  const uint32_t nkA2 = ((0x10000u - kA) & 0xFFFFu) * 0x00010001u;
  const uint32_t nc2 = ((0x10000u - c) & 0xFFFFu) * 0x00010001u;
  const uint32_t one_rt = p.neg_one * p.neg_one;       // 1, unknown to the compiler (FF_COUNT_FMA_ADDS)

  // The item loop is counted in ALU instructions (ncu: ALU pipe 70 % busy at 85 % of the DRAM pin rate; a third of
  // them were the item header): 32-bit trip count, ring slot and phase as running values, the ragged last tile
  // from a compare.
  unsigned seg_no = 0;                   // segments this warp has finished (the same sequence in every warp)
  int fs = f0, ts = tile0;
  int s = 0;                             // ring slot and its phase
  uint32_t ph = 0;
  const int last_groups = (int)(groups_per_frame - (int64_t)(T - 1) * kTileGroups);
  for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
    const int64_t i0 = slab << sh;
    const int n_it = (int)(min(i0 + S, total_work) - i0);
    int f = fs;
    int tile = ts;
    fs += p.slab_step_frames;
    ts += p.slab_step_tiles;
    if (ts >= T) { ts -= T; ++fs; }
    int wcnt = 0;
    for (int k_it = 0; k_it < n_it; ++k_it) {
      mbar_wait(&full[s], ph);
      const bool last_of_frame = tile + 1 == T;
      const int tile_groups = last_of_frame ? last_groups : kTileGroups;
      uint32_t acc = 0;
      if (BITS == 12) {
#pragma unroll
        for (int c4 = 0; c4 < kTileBytes / (48 * kThreads); ++c4) {
          const int g = (c4 * kThreads + tid) * 4;       // groups per frame are a multiple of 4: all four or none
          if (g < tile_groups) {
            const uint4* qq = reinterpret_cast<const uint4*>(smem + s * kTileBytes + (c4 * kThreads + tid) * 48);
            const uint4 q0 = qq[0], q1 = qq[1], q2 = qq[2];
            acc = count12x8_simd(q0.x, q0.y, q0.z, kA2, nkA2, nc2, acc, one_rt);
            acc = count12x8_simd(q0.w, q1.x, q1.y, kA2, nkA2, nc2, acc, one_rt);
            acc = count12x8_simd(q1.z, q1.w, q2.x, kA2, nkA2, nc2, acc, one_rt);
            acc = count12x8_simd(q2.y, q2.z, q2.w, kA2, nkA2, nc2, acc, one_rt);
          }
        }
      } else {
        const uint32_t one2 = 0x00010001u;
        const int pieces = tile_groups * BITS / 16;                      // 16-byte pieces in this tile
        const uint4* qq = reinterpret_cast<const uint4*>(smem + s * kTileBytes);
#pragma unroll
        for (int k = 0; k < kTileBytes / 16 / kThreads; ++k) {
          const int j0 = tid + k * kThreads;
          if (j0 < pieces) {
            const uint4 v = qq[j0];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (BITS == 16) {
                acc += __viaddmin_u16x2(__vimax3_u16x2(w4[j], kA2, kA2), nkA2, one2);
              } else {
                acc += __viaddmin_s16x2_relu(__byte_perm(w4[j], 0u, 0x4140), nc2, one2) +
                       __viaddmin_s16x2_relu(__byte_perm(w4[j], 0u, 0x4342), nc2, one2);
              }
            }
          }
        }
      }
      int cnt = (int)(acc & 0xFFFFu) + (int)(acc >> 16);
      cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);      // also orders every lane's smem reads before the release
      if (lane == 0) {
        mbar_arrive(&empty[s]);
        wcnt += cnt;
        if (last_of_frame || k_it + 1 == n_it) {      // the segment ends: leaving the frame, or the slab
          unsigned* word = seg + (seg_no & (kSegRing - 1));
          ++seg_no;
          const unsigned old = atomicAdd(word, (1u << 24) | (unsigned)wcnt);
          if ((old >> 24) == kWarpsPerCta - 1) {      // last warp of the CTA: the segment goes to the detector warp
            *word = 0;
            const int total = (int)((old & 0xFFFFFFu) + (unsigned)wcnt);
            const int slot = atomicAdd(q_ctl, 1);
            while (slot - ((volatile int*)q_ctl)[4 + slot % kDetWarps] >= kQueue) __nanosleep(64);
            q_cnt[slot % kQueue] = total;
            __threadfence_block();
            q_f[slot % kQueue] = f;
          }
          wcnt = 0;
        }
      }
      if (last_of_frame) { tile = 0; ++f; } else { ++tile; }
      if (++s == kCountStages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(q_ctl + 1, 1);
  }
}

template <int BITS, int kCountStages, int kItemKB>
int launch_range_fused(StreamParams p, const DetectParams& d, cudaStream_t st) {
  const size_t smem = (size_t)kCountStages * (kItemKB * 1024) + 2 * kCountStages * 8 + 2 * kQueue * 4 +
                      (4 + kDetWarps) * 4 + kSegRing * 4 + kDetWarps * detect_warp_smem(d.width, BITS);
  auto kern = range_kernel<BITS, kCountStages, kItemKB>;
  p.tiles_per_frame = (int)((p.px_per_frame / kGroupPx + (kItemKB * 1024 / BITS) - 1) / (kItemKB * 1024 / BITS));
  // launch configuration per device, recomputed only when the frame width (detector smem) changes
  static std::mutex m;
  static size_t have_smem[64] = {};
  static int have_occ[64] = {};
  int dev = 0, occ = 0;
  FF_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return FF_ERR_INVALID;
  {
    std::lock_guard<std::mutex> g(m);
    if (have_smem[dev] != smem) {
      FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // the SM's shared-memory / L1 split at its maximum, like the small kernels that run NEXT TO this one
      // (merge of the previous clip, prep): CTAs of kernels that ask for different splits cannot share an SM
      FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFusedThreads, smem));
      int cap = kCountCtasPerSm;             // bytes in flight per SM: see launch_count12
      if (const char* e = getenv("FF_COUNT12_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;   // tuning knob
      if (occ > cap) occ = cap;
      if (occ < 1) occ = 1;
      have_occ[dev] = occ;
      have_smem[dev] = smem;
    }
    occ = have_occ[dev];
  }
  const int64_t wave = (int64_t)sm_count_cached() * occ;
  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  // slab size: as large as possible (fewer segments per frame, longer contiguous reads) while every CTA
  // still gets >= 64 slabs, which bounds the imbalance of the round-robin deal to 1.6 %
  int64_t slab = total_work / (wave * 64);
  static const int slab_env = getenv("FF_RANGE_SLAB") ? atoi(getenv("FF_RANGE_SLAB")) : 0;       // tuning knob
  if (slab_env > 0) slab = slab_env;
  if (slab > 32 && slab_env <= 0) slab = 32;
  if (slab < 1) slab = 1;
  int shift = 0;
  while ((2ll << shift) <= slab) ++shift;                  // largest power of two <= slab
  p.slab_shift = shift;
  static const int idle_env = getenv("FF_RANGE_IDLE_NS") ? atoi(getenv("FF_RANGE_IDLE_NS")) : 0;     // tuning knob
  p.idle_ns = idle_env > 0 ? idle_env : 4000;
  const int64_t n_slabs = (total_work + (1ll << shift) - 1) >> shift;
  const int64_t grid = n_slabs < wave ? n_slabs : wave;
  if ((grid << shift) > 0x7FFFFFFFll) return FF_ERR_UNSUPPORTED;
  p.slab_step_frames = (int)((grid << shift) / p.tiles_per_frame);
  p.slab_step_tiles = (int)((grid << shift) % p.tiles_per_frame);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  FF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p, d));
  return FF_OK;
}

// ---- difference image (uint16 / float32 / float64) and / or decoded uint16 pixels (8-bit, packed 12-bit, 16-bit input) ---
// Same skeleton as count12_kernel (producer warp, full/empty ring, no CTA-wide barrier), same
// pixel ownership as the general template (thread t owns groups t, t+256, t+512, t+768 of the
// tile, so every warp-level 16-byte store covers 512 contiguous bytes) and all arithmetic on
// 16x2 SIMD lanes.  The previous frame is carried as Pn = 0x4000 - sub per lane, which turns the
// lane-wise difference into a plain 32-bit add (E = sub + Pn = 0x4000 + d, no lane can overflow).
// The ALU pipe, not HBM, is what held this kernel back (ncu, round 1: ALU 71 % busy -> 0.85 of the copy rate;
// round 2: 64 % with math-pipe throttle at 2.25 warps per scheduler -> 0.90), so the inner loop is counted in ALU
// instructions per 8 pixels: adds and the carry update are IMADs on the FMA pipe, no literal-zero operands
// (kLaneMin2), no per-word selects, and the item header is strength-reduced by hand.  With that the ring depth
// became the tuning parameter, sharply and differently per variant (launch_streamx_tuned).
// float32 / float64 images are the same integer lanes widened on the way out (lane_to_float / lane_to_double:
// exact), with lane pairs swapping words so that every store instruction of a warp writes whole 32-byte sectors.
// Item order: tile-major with a halo item per segment when the difference is retained (the
// carry needs consecutive frames of one tile); frame-major otherwise.
// 8-bit pixels are widened to 16x2 lanes with two PRMTs per word and share the signed-lane
// arithmetic with 12-bit (every value fits 13 bits).  Full-range 16-bit pixels do not: there the
// lanes are unsigned and every subtraction is "max, then subtract" (sub = max(x,bg) - bg,
// relu(d) = max(sub,prev) - prev, ...), which cannot borrow across lanes.
constexpr int kOutThreads = kThreads + 32;
// "max(., 0)" is written as the RELU form of VIADDMNMX with the most negative lane value as its third operand:
// a literal zero there is materialised into a register by ptxas again and again (one PRMT per use: 5 extra ALU
// instructions per 8 pixels), any other value travels as an immediate.
constexpr uint32_t kLaneMin2 = 0x80008000u;

// Output stores are streaming (st.global.cs); plain and write-through stores measured the same
// (C4 uint16 difference 0.920 / 0.925 / 0.926 of the copy rate).
__device__ __forceinline__ void st_out(uint4* p, uint4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_out(float4* p, float4 v) { __stcs(p, v); }

// One 16-bit lane of a 16x2 word as float32, exactly: the lane becomes the low mantissa bits of 2^23 (one PRMT),
// and 2^23 is subtracted again (one FADD) - no integer-to-float conversion unit involved.
template <int HI>
__device__ __forceinline__ float lane_to_float(uint32_t w) {
  return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, HI ? 0x7632 : 0x7610)), -8388608.0f);
}
template <int HI>
__device__ __forceinline__ double lane_to_double(uint32_t w) {       // the same with 2^52
  return __dadd_rn(__hiloint2double(0x43300000, (int)(HI ? w >> 16 : w & 0xFFFFu)), -4503599627370496.0);
}

template <int BITS>
__device__ __forceinline__ void load_lanes(const uint8_t* stage, int g, uint32_t (&x)[4]) {
  if (BITS == 12) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage) + 3 * g;
    decode12x8_16x2(w[0], w[1], w[2], x);
  } else if (BITS == 16) {
    const uint4 q = reinterpret_cast<const uint4*>(stage)[g];
    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
  } else {
    const uint2 q = reinterpret_cast<const uint2*>(stage)[g];
    x[0] = __byte_perm(q.x, 0u, 0x4140);   // b0 0 b1 0  (LSB first)
    x[1] = __byte_perm(q.x, 0u, 0x4342);
    x[2] = __byte_perm(q.y, 0u, 0x4140);
    x[3] = __byte_perm(q.y, 0u, 0x4342);
  }
}

// (Sixteen consumer warps instead of eight - each thread two groups of a tile - were tried for latency hiding and
// are slower, 0.93 against 0.97 on C4: the per-item header is paid per warp, and the kernel executed 27 % more
// instructions; profiles/r02_streamx_stage_sweep.txt.)
template <int BITS, bool COUNT, int DIFF, bool DECODED, int STAGES>
__global__ void __launch_bounds__(kOutThreads) streamx_kernel(const StreamParams p) {
  constexpr int NW = kWarpsPerCta;                         // consumer warps
  constexpr int kCons = 32 * NW;                           // consumer threads
  constexpr int kPerThread = 4 * kThreads / kCons;         // groups per thread and tile
  constexpr int kTileBytes = 4 * kThreads * BITS;          // 1024 groups of 8 pixels
  constexpr bool kUnsigned = BITS == 16;
  constexpr int kMaxPx = BITS == 8 ? 255 : (BITS == 12 ? 4095 : 65535);
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kTileBytes);
  uint64_t* empty = full + STAGES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  constexpr int kTileGroups = 4 * kThreads;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  int64_t work = (int64_t)blockIdx.x * p.items_per_cta;
  const int64_t work_end = min(work + p.items_per_cta, total_work);
  // Retained difference: the carry needs consecutive frames of ONE tile, so the work is cut into UNITS of
  // unit_frames frames of one tile, numbered segment-major, and the units are dealt to the CTAs round-robin:
  // at any moment all CTAs are in the same few frames - the reads of the whole GPU fall into a couple of
  // contiguous frames and so do its writes, instead of 148 read and 148 write streams all over the clip.
  const bool units = DIFF && p.unit_frames > 0;
  int64_t unit = blockIdx.x;
  if (!units && work >= work_end) return;
  const int64_t groups_per_frame = p.px_per_frame / kGroupPx;
  const bool producer = warp == NW;
  if (producer && (tid & 31) != 0) return;
  const uint64_t policy = producer ? policy_evict_first() : 0;

  // consumer constants
  int bg = 0;
  uint32_t nc2 = 0, nbg2 = 0, k2 = 0, c2 = 0, bg2 = 0, t2 = 0;
  int tm1 = 0;
  const uint32_t one2 = 0x00010001u;
  if (!producer) griddep_wait();       // the background scalar may come from the kernel right before this one
  if (!producer && (COUNT || DIFF)) {
    bg = min(__ldg(p.bg_dev), kMaxPx);                   // a larger bg zeroes everything anyway
    const int ethr = p.empty_thr >= 0 ? p.empty_thr : max(10, __ldg(p.bg_dev) >> 1);
    const uint32_t c = (uint32_t)min((int64_t)__ldg(p.bg_dev) + ethr, (int64_t)kMaxPx);   // no pixel exceeds kMaxPx
    if (kUnsigned) {
      c2 = c * one2;
      nc2 = ((0x10000u - c) & 0xFFFFu) * one2;
      bg2 = (uint32_t)bg * one2;
      tm1 = min(max(p.diff_thr - 1, 0), 65535);          // lanes hold relu(d): thr 0 and 1 coincide there
      t2 = (uint32_t)tm1 * one2;
    } else {
      nc2 = ((0x10000u - c) & 0xFFFFu) * one2;             // -c per lane
      nbg2 = ((0x10000u - (uint32_t)bg) & 0xFFFFu) * one2;  // -bg per lane
      tm1 = min(max(p.diff_thr, 0) - 1, 8190);             // thr-1 (thr > 4095 keeps nothing)
      k2 = ((uint32_t)(-(0x4000 + tm1)) & 0xFFFFu) * one2;  // relu(E + k) = relu(d - (thr-1))
    }
  }
  const uint32_t neg1 = p.neg_one;
  uint32_t pn[kPerThread][4];       // carry, 16x2: 0x4000 - sub (signed lanes) or sub itself (unsigned lanes)
#pragma unroll
  for (int k = 0; k < kPerThread; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) pn[k][j] = kUnsigned ? 0u : 0x40004000u;
  int s = 0;                 // ring slot and its phase, carried across segments
  uint32_t ph = 0;
  const int last_tile = p.tiles_per_frame - 1;
  const int last_groups = (int)(groups_per_frame - (int64_t)last_tile * kTileGroups);

  while (units ? unit < p.n_units : work < work_end) {
    // ---- one segment: consecutive items that share the carry (DIFF) or simply the rest of the run
    int tile, f0, n_seg;
    const uint8_t* halo_ptr = nullptr;
    if (DIFF) {
      if (units) {
        const int seg = (int)(unit / p.tiles_per_frame);
        tile = (int)(unit - (int64_t)seg * p.tiles_per_frame);
        f0 = seg * p.unit_frames;
        n_seg = min(p.unit_frames, p.n_frames - f0);
        unit += gridDim.x;
      } else {
        tile = (int)(work / p.n_frames);
        f0 = (int)(work - (int64_t)tile * p.n_frames);
        n_seg = (int)min((int64_t)(p.n_frames - f0), work_end - work);
      }
      int hf = f0 - 1;
      if (p.skip != nullptr)
        while (hf >= 0 && p.skip[hf]) --hf;
      halo_ptr = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
    } else {
      f0 = (int)(work / p.tiles_per_frame);
      tile = (int)(work - (int64_t)f0 * p.tiles_per_frame);
      n_seg = (int)(work_end - work);
    }
    work += n_seg;
    const int has_halo = halo_ptr != nullptr ? 1 : 0;
    const int n_items = n_seg + has_halo;
    bool have_prev = false;

    // Running state of the item loop, strength-reduced by hand (the item header - divisions, 64-bit products, the
    // ring slot from a modulo - was a quarter of the consumers' instructions): DIFF items are consecutive frames of
    // one tile (the halo item counts as frame f0 - 1), the others consecutive tiles in memory order.
    int f = DIFF ? f0 - has_halo : f0;
    int t = tile;
    int64_t px0 = (int64_t)f * p.px_per_frame + (int64_t)t * (kTileGroups * kGroupPx);
    int64_t pidx = ((int64_t)f * p.tiles_per_frame + t) * kWarpsPerCta + warp;
    const uint8_t* src = p.frames + (int64_t)f * p.frame_bytes + (int64_t)t * kTileBytes;      // producer

    for (int it = 0; it < n_items; ++it) {
      const bool is_halo = it < has_halo;
      const int tile_groups = t == last_tile ? last_groups : kTileGroups;

      if (producer) {
        mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t bytes = (uint32_t)tile_groups * (uint32_t)BITS;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(smem + s * kTileBytes, is_halo ? halo_ptr + (int64_t)t * kTileBytes : src, bytes, &full[s], policy);
        if (DIFF) {
          src += p.frame_bytes;
          ++f;
        } else {
          src += bytes;
          if (++t == p.tiles_per_frame) t = 0;
        }
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
        continue;
      }

      mbar_wait(&full[s], ph);
      const uint8_t* stage = smem + s * kTileBytes;
      const bool skipped = DIFF && !is_halo && p.skip != nullptr && p.skip[f] != 0;
      const bool diff_valid = have_prev && !skipped;
      // a frame without a valid difference (first of the clip, skipped) stores zeros: its threshold term is pushed
      // below every possible lane value instead of selecting per output word (signed lanes)
      const uint32_t kd2 = diff_valid ? k2 : 0x80018001u;
      uint32_t acc2 = 0;
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int g = tid + k * kCons;
        // (tiles end on a multiple of four groups: the two lanes of an exchanging pair are inside or outside together)
        const uint32_t in_mask = (DIFF == FF_DIFF_F32 || DIFF == FF_DIFF_F64) ? __ballot_sync(0xFFFFFFFFu, g < tile_groups) : 0u;
        if (g < tile_groups) {
          uint32_t x[4];
          load_lanes<BITS>(stage, g, x);
          if (COUNT) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc2 += kUnsigned ? __viaddmin_u16x2(__vimax3_u16x2(x[j], c2, c2), nc2, one2)   // [x > c], unsigned
                                : __viaddmin_s16x2_relu(x[j], nc2, one2);                       // [x > c], signed
          }
          if (DECODED && !is_halo)
            st_out(reinterpret_cast<uint4*>(p.decoded_out + px0 + (int64_t)g * kGroupPx), make_uint4(x[0], x[1], x[2], x[3]));
          if (DIFF) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (kUnsigned) {
                const uint32_t sub2 = __vimax3_u16x2(x[j], bg2, bg2) - bg2;                 // max(x - bg, 0)
                const uint32_t rd2 = __vimax3_u16x2(sub2, pn[k][j], pn[k][j]) - pn[k][j];   // relu(d)
                const uint32_t r2 = __vimax3_u16x2(rd2, t2, t2) - t2;                        // relu(relu(d) - (thr-1))
                const uint32_t m2 = __vimin3_u16x2(r2, one2, one2);                          // [d >= thr]
                o[j] = r2 + m2 * (uint32_t)tm1;                                              // d where d >= thr, else 0
                if (!skipped) pn[k][j] = sub2;
              } else {
                const uint32_t sub2 = __viaddmax_s16x2_relu(x[j], nbg2, kLaneMin2);  // max(x - bg, 0)
                const uint32_t e2 = sub2 + pn[k][j];                                // 0x4000 + d per lane
                const uint32_t r2 = __viaddmax_s16x2_relu(e2, kd2, kLaneMin2);       // relu(d - (thr-1))
                const uint32_t m2 = __vimin_s16x2_relu(r2, one2);                   // [d >= thr]
                o[j] = r2 + m2 * (uint32_t)tm1;                                     // d where d >= thr, else 0
                if (!skipped) pn[k][j] = sub2 * neg1 + 0x40004000u;                  // 0x4000 - sub per lane
              }
            }
            if (!is_halo) {
              if (kUnsigned && !diff_valid) o[0] = o[1] = o[2] = o[3] = 0u;
              if (DIFF == FF_DIFF_U16) {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.diff_out) + px0 + (int64_t)g * kGroupPx);
                st_out(dst, make_uint4(o[0], o[1], o[2], o[3]));
              } else if (DIFF == FF_DIFF_F64) {
                // float64 image (the reference's dtype): 64 bytes per thread.  Lanes 2i and 2i+1 swap two words so
                // that store instruction q of the warp writes sector q of every pair's 128 bytes in full: the even
                // lane the first two pixels of the sector, the odd lane the other two.
                const bool odd = (tid & 1) != 0;
                const uint32_t r0 = __shfl_xor_sync(in_mask, odd ? o[0] : o[1], 1);
                const uint32_t r1 = __shfl_xor_sync(in_mask, odd ? o[2] : o[3], 1);
                const uint32_t a[4] = {odd ? r0 : o[0], odd ? r1 : o[2], odd ? o[1] : r0, odd ? o[3] : r1};
                double2* pair = reinterpret_cast<double2*>(static_cast<double*>(p.diff_out) + px0 + (int64_t)(g & ~1) * kGroupPx) + (odd ? 1 : 0);
#pragma unroll
                for (int q = 0; q < 4; ++q) __stcs(pair + 2 * q, make_double2(lane_to_double<0>(a[q]), lane_to_double<1>(a[q])));
              } else {                  // float32 image: the integer lanes widened exactly (lane_to_float)
                // Lanes 2i and 2i+1 hold 16 consecutive pixels = 64 bytes of output: they swap halves so that each of
                // the two store instructions of the warp writes whole 32-byte sectors (two 16-byte stores per thread
                // straight from its own lanes leave every sector half-written twice: 0.859 against 0.875 on C4).
                const bool odd = (tid & 1) != 0;
                const uint32_t r0 = __shfl_xor_sync(in_mask, odd ? o[0] : o[2], 1);
                const uint32_t r1 = __shfl_xor_sync(in_mask, odd ? o[1] : o[3], 1);
                const uint32_t a0 = odd ? r0 : o[0], a1 = odd ? r1 : o[1];          // pixels 4 (tid & 1) .. + 3 of the pair
                const uint32_t b0 = odd ? o[2] : r0, b1 = odd ? o[3] : r1;          // pixels 8 + 4 (tid & 1) ..
                float4* pair = reinterpret_cast<float4*>(static_cast<float*>(p.diff_out) + px0 + (int64_t)(g & ~1) * kGroupPx) + (odd ? 1 : 0);
                st_out(pair, make_float4(lane_to_float<0>(a0), lane_to_float<1>(a0), lane_to_float<0>(a1), lane_to_float<1>(a1)));
                st_out(pair + 2, make_float4(lane_to_float<0>(b0), lane_to_float<1>(b0), lane_to_float<0>(b1), lane_to_float<1>(b1)));
              }
            }
          }
        }
      }
      if (DIFF && !skipped) have_prev = true;
      int cnt = (int)(acc2 & 0xFFFFu) + (int)(acc2 >> 16);
      if (COUNT) cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
      else __syncwarp();
      if ((tid & 31) == 0) {
        mbar_arrive(&empty[s]);
        if (COUNT && !is_halo) p.partial[pidx] = cnt;
      }
      if (DIFF) {
        ++f;
        px0 += p.px_per_frame;
        pidx += (int64_t)p.tiles_per_frame * kWarpsPerCta;
      } else {
        px0 += tile_groups * kGroupPx;
        pidx += kWarpsPerCta;
        if (++t == p.tiles_per_frame) {
          t = 0;
          ++f;
        }
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
}

int sm_count_cached();

template <int BITS, bool COUNT, int DIFF, bool DECODED, int STAGES>
int launch_streamx(StreamParams p, int ctas_cap, cudaStream_t st) {
  constexpr int kNeeded = STAGES * (4 * kThreads * BITS) + 2 * STAGES * 8;
  auto kern = streamx_kernel<BITS, COUNT, DIFF, DECODED, STAGES>;
  static PerDeviceInt cache;
  int ctas = 1;
  int rc = cache.get([&](int* v) -> int {
    FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerSm));
    FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kOutThreads, kNeeded));
    if (occ > ctas_cap) occ = ctas_cap;
    *v = occ > 0 ? occ : 1;
    return FF_OK;
  }, &ctas);
  if (rc != FF_OK) return rc;
  const int smem = pinned_smem(kNeeded, ctas);       // exactly `ctas` CTAs fit an SM, whatever else runs
  const int64_t wave = (int64_t)sm_count_cached() * ctas;
  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  p.items_per_cta = (total_work + wave - 1) / wave;
  if (p.items_per_cta < 1) p.items_per_cta = 1;
  int64_t grid = (total_work + p.items_per_cta - 1) / p.items_per_cta;
  static const bool sync_units = getenv("FF_STREAMX_SYNC") == nullptr || atoi(getenv("FF_STREAMX_SYNC")) != 0;   // tuning knob
  p.unit_frames = 0;
  p.neg_one = 0xFFFFFFFFu;
  if (DIFF && sync_units && p.n_frames >= 96) {
    // frames per unit: long enough that the extra halo item per unit stays below ~2 %, and such that the
    // units divide evenly over the wave (every CTA the same number of units)
    int best_nseg = 1;
    double best_cost = 1e30;
    for (int nseg = (p.n_frames + 511) / 512; nseg <= p.n_frames / 48; ++nseg) {
      const int len = (p.n_frames + nseg - 1) / nseg;
      const int64_t units = (int64_t)p.tiles_per_frame * ((p.n_frames + len - 1) / len);
      const int64_t rounds = (units + wave - 1) / wave;
      const double cost = (double)(rounds * wave) / (double)units * (1.0 + 1.0 / len);      // imbalance x halo overhead
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best_nseg = nseg;
      }
      if (cost < 1.012) break;
    }
    p.unit_frames = (p.n_frames + best_nseg - 1) / best_nseg;
    p.n_units = (int64_t)p.tiles_per_frame * ((p.n_frames + p.unit_frames - 1) / p.unit_frames);
    grid = p.n_units < wave ? p.n_units : wave;
  }
  return launch_kernel(kern, dim3((unsigned)grid), dim3(kOutThreads), smem, st, p.pdl != 0, p);
}

template <int BITS, bool COUNT, int DIFF, bool DECODED>
int launch_streamx_tuned(const StreamParams& p, cudaStream_t st) {
  // One CTA per SM with a 5-deep ring (48 KB of reads in flight per SM at 12 bits) is the measured
  // optimum for the read+write variants - fewer concurrent DRAM streams beat more parallelism (C4
  // uint16 diff: 0.85 of the copy rate at 3 CTAs x 4 stages, 0.92 at 1 x 5;
  // profiles/r01_stream12_sweep.txt).
  // The ring depth is tuned per variant (profiles/r02_streamx_stage_sweep.txt; one stage more or less costs 3-10 %):
  // the more of the traffic is writes, the fewer reads want to be in flight.
  constexpr int kDefaultStages = DIFF == FF_DIFF_F32 ? 3 : DIFF == FF_DIFF_F64 ? 2
                                 : DIFF == FF_DIFF_U16 ? (BITS == 16 ? 4 : 5) : 4;
  static const int stages = getenv("FF_STREAM12_STAGES") ? atoi(getenv("FF_STREAM12_STAGES")) : kDefaultStages;   // tuning knobs
  static const int ctas = getenv("FF_STREAM12_CTAS") ? atoi(getenv("FF_STREAM12_CTAS")) : 1;
  switch (stages) {
    case 2: return launch_streamx<BITS, COUNT, DIFF, DECODED, 2>(p, ctas, st);
    case 3: return launch_streamx<BITS, COUNT, DIFF, DECODED, 3>(p, ctas, st);
    case 4: return launch_streamx<BITS, COUNT, DIFF, DECODED, 4>(p, ctas, st);
    case 6: return launch_streamx<BITS, COUNT, DIFF, DECODED, 6>(p, ctas, st);
    default: return launch_streamx<BITS, COUNT, DIFF, DECODED, 5>(p, ctas, st);
  }
}

// Shapes the TMA path cannot take (P % 32 != 0): one thread per pixel pair, plain loads,
// the previous frame is simply re-read.  Correct for every even P; not a performance path.
template <int BITS, int DIFF>
__global__ void stream_generic_kernel(const StreamParams p, int64_t pairs_per_frame) {
  const int64_t total = (int64_t)p.n_frames * pairs_per_frame;
  const int bg = __ldg(p.bg_dev);
  const int ethr = p.empty_thr >= 0 ? p.empty_thr : max(10, bg >> 1);
  const int cthr = bg + ethr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / pairs_per_frame);
    const int64_t q = (i - (int64_t)f * pairs_per_frame) * 2;
    const uint8_t* fr = p.frames + (int64_t)f * p.frame_bytes;
    const bool two = q + 1 < p.px_per_frame;
    const int v0 = load_px_generic<BITS>(fr, q);
    const int v1 = two ? load_px_generic<BITS>(fr, q + 1) : 0;
    const int c = (v0 > cthr ? 1 : 0) + ((two && v1 > cthr) ? 1 : 0);
    if (c) atomicAdd(p.partial + f, c);
    if (p.decoded_out != nullptr) {
      p.decoded_out[(int64_t)f * p.px_per_frame + q] = (uint16_t)v0;
      if (two) p.decoded_out[(int64_t)f * p.px_per_frame + q + 1] = (uint16_t)v1;
    }
    if (DIFF != FF_DIFF_NONE) {
      const bool skipped = p.skip != nullptr && p.skip[f] != 0;
      int hf = f - 1;
      if (p.skip != nullptr)
        while (hf >= 0 && p.skip[hf]) --hf;
      const uint8_t* pr = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
      int d0 = 0, d1 = 0;
      if (pr != nullptr && !skipped) {
        d0 = max(v0 - bg, 0) - max(load_px_generic<BITS>(pr, q) - bg, 0);
        if (d0 < p.diff_thr) d0 = 0;
        if (two) {
          d1 = max(v1 - bg, 0) - max(load_px_generic<BITS>(pr, q + 1) - bg, 0);
          if (d1 < p.diff_thr) d1 = 0;
        }
      }
      const int64_t o = (int64_t)f * p.px_per_frame + q;
      if (DIFF == FF_DIFF_U16) {
        static_cast<uint16_t*>(p.diff_out)[o] = (uint16_t)d0;
        if (two) static_cast<uint16_t*>(p.diff_out)[o + 1] = (uint16_t)d1;
      } else if (DIFF == FF_DIFF_F32) {
        static_cast<float*>(p.diff_out)[o] = (float)d0;
        if (two) static_cast<float*>(p.diff_out)[o + 1] = (float)d1;
      } else {
        static_cast<double*>(p.diff_out)[o] = (double)d0;
        if (two) static_cast<double*>(p.diff_out)[o + 1] = (double)d1;
      }
    }
  }
}

// Decode-only generic kernel (no background known): ff_unpack on odd shapes.
__global__ void unpack12_generic_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out,
                                        int64_t n_px) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_px;
       q += (int64_t)gridDim.x * blockDim.x)
    out[q] = (uint16_t)load_px_generic<12>(in, q);
}

int sm_count_cached() {
  static PerDeviceInt cache;
  int n = 148;
  cache.get([](int* v) -> int {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    *v = sms;
    return FF_OK;
  }, &n);
  return n;
}

template <int BITS, bool COUNT, int DIFF, bool DECODED, int K>
int launch_stream(StreamParams p, cudaStream_t st) {
  auto kern = stream_kernel<BITS, COUNT, DIFF, DECODED, K>;
  constexpr int kStageBytes = K * kThreads * BITS;
  constexpr int kNeeded = kStages * kStageBytes + kStages * 8;
  constexpr bool kCapped = DIFF == FF_DIFF_F64 || DIFF == FF_DIFF_F32;
  static PerDeviceInt cache;
  int ctas = 1;
  int rc = cache.get([&](int* v) -> int {
    FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kCapped ? kSmemPerSm : kNeeded));
    int occ = 0;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, kNeeded));
    // Write-heavy variants run faster with FEWER resident CTAs (fewer concurrent DRAM streams):
    // measured on C4, float64 difference 0.86 -> 0.95 of the copy rate at 1 CTA/SM, float32
    // 0.82 -> 0.87 at 2 (profiles/r01_count12_sweep.txt).  FF_STREAM_CTAS overrides (tuning knob).
    int cap = DIFF == FF_DIFF_F64 ? 1 : (DIFF == FF_DIFF_F32 ? 2 : occ);
    if (const char* e = getenv("FF_STREAM_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;
    if (occ > cap) occ = cap;
    *v = occ > 0 ? occ : 1;
    return FF_OK;
  }, &ctas);
  if (rc != FF_OK) return rc;
  const int smem = kCapped ? pinned_smem(kNeeded, ctas) : kNeeded;     // see pinned_smem
  // One resident wave: the (tile, frame) items are split evenly over SMs x CTAs/SM long-lived CTAs.
  const int64_t wave = (int64_t)sm_count_cached() * ctas;
  const int64_t total_work = (int64_t)p.tiles_per_frame * p.n_frames;
  p.items_per_cta = (total_work + wave - 1) / wave;
  if (p.items_per_cta < 1) p.items_per_cta = 1;
  const int64_t grid = (total_work + p.items_per_cta - 1) / p.items_per_cta;
  if (grid > 0x7FFFFFFF) return FF_ERR_UNSUPPORTED;
  return launch_kernel(kern, dim3((unsigned)grid), dim3(kThreads), smem, st, p.pdl != 0, p);
}

template <int BITS, bool COUNT, int DIFF, bool DECODED>
int launch_stream_k(const StreamParams& p, int k, cudaStream_t st) {
  // 8192-px tiles reach this template only with float difference outputs: everything else at
  // that tile size is served by count12_kernel / streamx_kernel.
  if (k == 4) {
    if constexpr (DIFF == FF_DIFF_F32 || DIFF == FF_DIFF_F64) return launch_stream<BITS, COUNT, DIFF, DECODED, 4>(p, st);
    else return FF_ERR_UNSUPPORTED;
  }
  return launch_stream<BITS, COUNT, DIFF, DECODED, 1>(p, st);
}

template <int BITS>
int dispatch_stream(const StreamParams& p, int k, int diff, bool decoded, cudaStream_t st) {
  if (decoded) {
    if (BITS != 12) return FF_ERR_UNSUPPORTED;
    switch (diff) {
      case FF_DIFF_NONE: return launch_stream_k<12, true, FF_DIFF_NONE, true>(p, k, st);
      case FF_DIFF_U16: return launch_stream_k<12, true, FF_DIFF_U16, true>(p, k, st);
      case FF_DIFF_F32: return launch_stream_k<12, true, FF_DIFF_F32, true>(p, k, st);
      case FF_DIFF_F64: return launch_stream_k<12, true, FF_DIFF_F64, true>(p, k, st);
    }
    return FF_ERR_INVALID;
  }
  switch (diff) {
    case FF_DIFF_NONE: return launch_stream_k<BITS, true, FF_DIFF_NONE, false>(p, k, st);
    case FF_DIFF_U16: return launch_stream_k<BITS, true, FF_DIFF_U16, false>(p, k, st);
    case FF_DIFF_F32: return launch_stream_k<BITS, true, FF_DIFF_F32, false>(p, k, st);
    case FF_DIFF_F64: return launch_stream_k<BITS, true, FF_DIFF_F64, false>(p, k, st);
  }
  return FF_ERR_INVALID;
}

template <int BITS>
int dispatch_generic(const StreamParams& p, int diff, cudaStream_t st) {
  const int64_t pairs = (p.px_per_frame + 1) / 2;
  const int64_t total = pairs * p.n_frames;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  FF_CUDA_TRY(cudaMemsetAsync(p.partial, 0, sizeof(int32_t) * (size_t)p.n_frames, st));
  switch (diff) {
    case FF_DIFF_NONE: stream_generic_kernel<BITS, FF_DIFF_NONE><<<(unsigned)blocks, 256, 0, st>>>(p, pairs); break;
    case FF_DIFF_U16: stream_generic_kernel<BITS, FF_DIFF_U16><<<(unsigned)blocks, 256, 0, st>>>(p, pairs); break;
    case FF_DIFF_F32: stream_generic_kernel<BITS, FF_DIFF_F32><<<(unsigned)blocks, 256, 0, st>>>(p, pairs); break;
    case FF_DIFF_F64: stream_generic_kernel<BITS, FF_DIFF_F64><<<(unsigned)blocks, 256, 0, st>>>(p, pairs); break;
    default: return FF_ERR_INVALID;
  }
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

}  // namespace

// True when a range with these outputs runs as ONE kernel (range_kernel): counts-only work on the TMA path.
bool range_is_fused(int64_t px, int diff_dtype, bool decoded, bool profiles) {
  const Tiling t = choose_tiling(px);
  if (getenv("FF_RANGE_UNFUSED")) return false;       // tuning / cross-check knob: the three-kernel sequence
  // Frames of at least 16384 pixels (BASELINE config 1: 512 x 64) take the range kernel as well: its items are
  // 12 KiB of stored bytes whatever the frame size, and a small clip is launch-latency bound - two launches
  // instead of three.  FF_RANGE_FUSE_SMALL=0 keeps them on the three-kernel form (cross-check knob).
  const char* small = getenv("FF_RANGE_FUSE_SMALL");
  const bool big_enough = t.k == 4 || ((small == nullptr || atoi(small) != 0) && px >= 16384);
  return t.fast && big_enough && diff_dtype == FF_DIFF_NONE && !decoded && !profiles;
}

// range_kernel for the frames described by `d` (whose ws / hooks / truncate say what the last CTA does).
int range_fused_impl(const DetectParams& d, int bits, int32_t empty_thr, bool pdl, cudaStream_t st) {
  const Tiling t = choose_tiling(d.px_per_frame);
  if (!range_is_fused(d.px_per_frame, FF_DIFF_NONE, false, d.profile_out != nullptr) || d.ws == nullptr)
    return FF_ERR_INVALID;
  if (((reinterpret_cast<uintptr_t>(d.frames) | reinterpret_cast<uintptr_t>(d.halo)) & 15u) != 0) return FF_ERR_ALIGNMENT;
  StreamParams p{};
  p.frames = d.frames;
  p.halo = d.halo;
  p.frame_bytes = d.frame_bytes;
  p.px_per_frame = d.px_per_frame;
  p.n_frames = d.n_frames;
  p.tiles_per_frame = t.tiles_per_frame;
  p.items_per_cta = 1;
  p.bg_dev = d.bg_dev;
  p.empty_thr = empty_thr;
  p.diff_thr = d.diff_thr;
  p.skip = d.skip;
  p.pdl = pdl ? 1 : 0;
  p.neg_one = 0xFFFFFFFFu;
  // Ring geometry, measured (profiles/r02_range_item_sweep.txt; GB/s of algorithmic bytes at C2):
  //   12-bit  12 KiB x 4 stages 6870   24 KiB x 3 stages 6700        -> 12 KiB (72 KiB in flight per SM)
  //   16-bit  12 KiB x 4        7140   24 KiB x 3        6820        -> 12 KiB
  //    8-bit  12 KiB x 4        6140   24 KiB x 3        6770        -> 24 KiB (96 KiB in flight per SM)
  // FF_RANGE_ITEM_KB / FF_COUNT12_STAGES override (tuning knobs).
  static const int stages = getenv("FF_COUNT12_STAGES") ? atoi(getenv("FF_COUNT12_STAGES")) : 4;
  static const int item_env = getenv("FF_RANGE_ITEM_KB") ? atoi(getenv("FF_RANGE_ITEM_KB")) : 0;
  const int item_kb = item_env > 0 ? item_env : (bits == 8 ? 24 : 12);
  if (item_kb == 24) {
    if (bits == 16) return launch_range_fused<16, 3, 24>(p, d, st);
    if (bits == 8) return launch_range_fused<8, 3, 24>(p, d, st);
    return launch_range_fused<12, 3, 24>(p, d, st);
  }
  if (bits == 16) return launch_range_fused<16, 4, 12>(p, d, st);
  if (bits == 8) return launch_range_fused<8, 4, 12>(p, d, st);
  switch (stages) {
    case 3: return launch_range_fused<12, 3, 12>(p, d, st);
    case 6: return launch_range_fused<12, 6, 12>(p, d, st);
    default: return launch_range_fused<12, 4, 12>(p, d, st);
  }
}

int stream_frames_impl(const void* frames, const void* halo, int64_t n_frames, int height, int width,
                       int bits, const int32_t* bg_dev, int32_t empty_thr, int32_t diff_thr,
                       const uint8_t* skip, int32_t* partial, void* diff_out, int diff_dtype,
                       uint16_t* decoded_out, cudaStream_t st, bool pdl) {
  if (frames == nullptr || bg_dev == nullptr || partial == nullptr) return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width <= 0 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  if (diff_dtype < FF_DIFF_NONE || diff_dtype > FF_DIFF_F64) return FF_ERR_INVALID;
  if ((diff_dtype != FF_DIFF_NONE) != (diff_out != nullptr)) return FF_ERR_INVALID;
  if (diff_dtype == FF_DIFF_U16 && diff_thr < 0) return FF_ERR_UNSUPPORTED;
  if (decoded_out != nullptr && bits != 12) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;  // frames must start on a byte triple

  const Tiling t = choose_tiling(px);
  StreamParams p{};
  p.frames = static_cast<const uint8_t*>(frames);
  p.halo = static_cast<const uint8_t*>(halo);
  p.frame_bytes = frame_bytes_of(px, bits);
  p.px_per_frame = px;
  p.n_frames = (int)n_frames;
  p.tiles_per_frame = t.tiles_per_frame;
  p.items_per_cta = 1;
  p.bg_dev = bg_dev;
  p.empty_thr = empty_thr;
  p.diff_thr = diff_thr;
  p.skip = skip;
  p.partial = partial;
  p.diff_out = diff_out;
  p.decoded_out = decoded_out;
  p.pdl = pdl ? 1 : 0;
  p.neg_one = 0xFFFFFFFFu;

  const bool aligned = ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(halo) |
                         reinterpret_cast<uintptr_t>(diff_out) | reinterpret_cast<uintptr_t>(decoded_out)) & 15u) == 0;
  if (t.fast && aligned) {
    if (diff_dtype == FF_DIFF_NONE && decoded_out == nullptr && t.k == 4) {      // counts only
      static const int stages = getenv("FF_COUNT12_STAGES") ? atoi(getenv("FF_COUNT12_STAGES")) : 4;   // tuning knob
      if (bits == 16) return launch_count12<16, 4>(p, st);      // 2 CTAs x 3 x 16 KB in flight
      if (bits == 8) return launch_count12<8, 6>(p, st);        // 2 CTAs x 5 x 8 KB in flight
      switch (stages) {
        case 2: return launch_count12<12, 2>(p, st);
        case 3: return launch_count12<12, 3>(p, st);
        case 6: return launch_count12<12, 6>(p, st);
        case 8: return launch_count12<12, 8>(p, st);
        default: return launch_count12<12, 4>(p, st);
      }
    }
    if (t.k == 4 && diff_dtype == FF_DIFF_U16) {          // uint16 difference image (+ decoded pixels, 12-bit)
      const bool dec = decoded_out != nullptr;
      if (bits == 12)
        return dec ? launch_streamx_tuned<12, true, true, true>(p, st) : launch_streamx_tuned<12, true, true, false>(p, st);
      if (!dec) return bits == 16 ? launch_streamx_tuned<16, true, true, false>(p, st)
                                  : launch_streamx_tuned<8, true, true, false>(p, st);
    }
    // float32 / float64 difference images: the same 16x2 arithmetic, lanes widened on the way out.
    // FF_STREAMX_F32=0 / FF_STREAMX_F64=0 select the general template (pixel pair per lane, scalar arithmetic).
    static const int f32_mode = getenv("FF_STREAMX_F32") ? atoi(getenv("FF_STREAMX_F32")) : 1;     // tuning knobs
    if (t.k == 4 && diff_dtype == FF_DIFF_F32 && decoded_out == nullptr && f32_mode != 0)
      return bits == 12 ? launch_streamx_tuned<12, true, FF_DIFF_F32, false>(p, st)
                        : (bits == 16 ? launch_streamx_tuned<16, true, FF_DIFF_F32, false>(p, st)
                                      : launch_streamx_tuned<8, true, FF_DIFF_F32, false>(p, st));
    static const int f64_mode = getenv("FF_STREAMX_F64") ? atoi(getenv("FF_STREAMX_F64")) : 1;     // tuning knob
    if (t.k == 4 && diff_dtype == FF_DIFF_F64 && decoded_out == nullptr && f64_mode != 0)
      return bits == 12 ? launch_streamx_tuned<12, true, FF_DIFF_F64, false>(p, st)
                        : (bits == 16 ? launch_streamx_tuned<16, true, FF_DIFF_F64, false>(p, st)
                                      : launch_streamx_tuned<8, true, FF_DIFF_F64, false>(p, st));
    if (t.k == 4 && bits == 12 && diff_dtype == FF_DIFF_NONE)        // counts + decoded (counts alone: count12)
      return launch_streamx_tuned<12, true, false, true>(p, st);
    switch (bits) {
      case 8: return dispatch_stream<8>(p, t.k, diff_dtype, decoded_out != nullptr, st);
      case 12: return dispatch_stream<12>(p, t.k, diff_dtype, decoded_out != nullptr, st);
      default: return dispatch_stream<16>(p, t.k, diff_dtype, decoded_out != nullptr, st);
    }
  }
  if (t.fast && !aligned) return FF_ERR_ALIGNMENT;
  switch (bits) {
    case 8: return dispatch_generic<8>(p, diff_dtype, st);
    case 12: return dispatch_generic<12>(p, diff_dtype, st);
    default: return dispatch_generic<16>(p, diff_dtype, st);
  }
}

int unpack_impl(const void* packed, void* out, int64_t n_frames, int height, int width, int bits,
                cudaStream_t st) {
  if (packed == nullptr || out == nullptr) return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width <= 0 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  const int64_t px = (int64_t)height * width;
  if (bits == 8 || bits == 16) {  // already byte-addressable: a plain device copy
    FF_CUDA_TRY(cudaMemcpyAsync(out, packed, (size_t)(n_frames * px * bits / 8), cudaMemcpyDeviceToDevice, st));
    return FF_OK;
  }
  if (bits != 12) return FF_ERR_UNSUPPORTED;
  if ((n_frames * px) & 1) return FF_ERR_UNSUPPORTED;
  const Tiling t = choose_tiling(px);
  const bool aligned = ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (t.fast && aligned) {
    StreamParams p{};
    p.frames = static_cast<const uint8_t*>(packed);
    p.frame_bytes = frame_bytes_of(px, 12);
    p.px_per_frame = px;
    p.n_frames = (int)n_frames;
    p.tiles_per_frame = t.tiles_per_frame;
    p.decoded_out = static_cast<uint16_t*>(out);
    if (t.k == 4) return launch_streamx_tuned<12, false, false, true>(p, st);
    return launch_stream<12, false, FF_DIFF_NONE, true, 1>(p, st);
  }
  const int64_t n_px = n_frames * px;
  int64_t blocks = (n_px + 255) / 256;
  const int64_t cap = (int64_t)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  unpack12_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<const uint8_t*>(packed),
                                                           static_cast<uint16_t*>(out), n_px);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

}  // namespace ff

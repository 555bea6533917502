// Stage 3 + 4 device code shared by detect_kernel (csrc/ff_detect.cu) and the fused range kernel
// (csrc/ff_stream.cu): one warp resolves one frame's centre-row profile, and the last CTA of a
// launch finishes the range (local truncation, publication to the peers of a range-sharded run).
#pragma once
#include <climits>

#include "ff_common.cuh"

namespace ff {

struct DetectParams {
  const uint8_t* frames;
  const uint8_t* halo;
  int64_t frame_bytes;
  int64_t px_per_frame;
  int n_frames;
  int64_t first_frame;
  int height, width;
  const int32_t* bg_dev;
  const int32_t* partial;
  int partials_per_frame;
  int64_t min_signal_count;
  int method;
  int use_diff;
  int diff_thr;
  int threshold_floor;
  const int32_t* threshold_dev;   // non-null: read the threshold method's bound from the device (prep_kernel wrote it)
  int grad2_bound;
  int min_run;
  int exit_margin;
  const uint8_t* skip;
  int32_t* pos_out;
  int32_t* count_out;
  int32_t* first_exit;
  int32_t* profile_out;
  int raw_stride;  // bytes reserved per staged row (multiple of 16)
  // tail of the launch (last CTA): truncation against *first_exit and the hooks of a range-sharded run
  RangeWorkspace* ws;             // nullptr: no tail
  int truncate;
  RangeHooks hooks;
};

// bytes of shared memory one detecting warp needs: int profile[W] + two staged raw rows
__host__ __device__ inline int detect_row_stride(int width, int bits) {
  const int row_bytes = (bits == 12) ? (width / 2 + 2) * 3 : width * (bits / 8);
  return (row_bytes + 3 + 15) & ~15;
}
__host__ __device__ inline size_t detect_warp_smem(int width, int bits) {
  return (size_t)width * sizeof(int) + 2 * (size_t)detect_row_stride(width, bits);
}

#if defined(__CUDACC__)
// Stage bytes [lo, lo+n) of `src` into `dst` (per-warp shared memory), whole words when the
// global address allows it.
__device__ __forceinline__ void warp_copy_bytes(uint8_t* dst, const uint8_t* __restrict__ src, int n, int lane) {
  if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
    const int nw = n >> 2;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    for (int i = lane; i < nw; i += 32) d32[i] = __ldg(s32 + i);
    for (int i = (nw << 2) + lane; i < n; i += 32) dst[i] = __ldg(src + i);
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = __ldg(src + i);
  }
}

// Where the centre row of a frame sits in the stored bytes.
struct RowSpan {
  int64_t q0;       // flat pixel index of the row's first pixel
  int64_t qbase;    // pixel held by staged byte 0
  int64_t byte_lo;
  int nbytes;
};
template <int BITS>
__device__ __forceinline__ RowSpan centre_row_span(int height, int W) {
  RowSpan r;
  r.q0 = (int64_t)(height / 2) * W;
  int64_t byte_hi;
  if (BITS == 12) {
    r.byte_lo = (r.q0 >> 1) * 3;
    byte_hi = ((r.q0 + W - 1) >> 1) * 3 + 3;
  } else {
    r.byte_lo = r.q0 * (BITS / 8);
    byte_hi = (r.q0 + W) * (BITS / 8);
  }
  r.nbytes = (int)(byte_hi - r.byte_lo);
  r.qbase = (BITS == 12) ? (r.q0 & ~(int64_t)1) : r.q0;
  return r;
}

// The latest non-skipped frame before f (scripts/process_videos.py:469, :1462, :1443-1445), or the halo.
__device__ __forceinline__ const uint8_t* prior_frame_of(const DetectParams& p, int f) {
  int hf = f - 1;
  if (p.skip != nullptr)
    while (hf >= 0 && p.skip[hf]) --hf;
  return hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
}

// One warp, one frame: stage the centre row (and the prior frame's), build the int32 profile and
// resolve the detection_method.  `empty` frames only get their profile written (profile_out).
// Returns the position (same value in every lane).  prof / raw_cur / raw_pri: this warp's shared memory.
template <int BITS>
__device__ int detect_one_frame(const DetectParams& p, const RowSpan& rs, int f, bool empty, int bg, int thr_floor,
                                int* prof, uint8_t* raw_cur, uint8_t* raw_pri, int lane) {
  const unsigned full = 0xFFFFFFFFu;
  const int W = p.width;
  const uint8_t* prior = p.use_diff ? prior_frame_of(p, f) : nullptr;
  const bool have_profile = !p.use_diff || prior != nullptr;
  int pos = FF_POS_NONE;
  if (!have_profile || (empty && p.profile_out == nullptr)) return pos;

  const uint8_t* cur_row = p.frames + (int64_t)f * p.frame_bytes + rs.byte_lo;
  const uint8_t* pri_row = p.use_diff ? prior + rs.byte_lo : cur_row;
  // Rows that start on a word boundary and hold whole 8-pixel groups (every Photron shape): each lane pulls its
  // groups of both rows straight from global memory - all loads independent, ONE round trip - and decodes in
  // registers.  (A detection is a chain of latencies inside a kernel that saturates the memory system; the
  // staged byte-wise path below costs several dependent round trips.)
  const bool fast = (W & 7) == 0 && ((reinterpret_cast<uintptr_t>(cur_row) | reinterpret_cast<uintptr_t>(pri_row)) & 3u) == 0 &&
                    (BITS != 12 || (rs.q0 & 7) == 0) && (reinterpret_cast<uintptr_t>(prof) & 15u) == 0;
  __syncwarp();
  if (fast) {
    constexpr int kWords = BITS / 4;                       // 32-bit words per 8-pixel group: 2 / 3 / 4
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(cur_row);
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(pri_row);
    const int groups = W >> 3;
#pragma unroll 4
    for (int g = lane; g < groups; g += 32) {
      uint32_t a[kWords], b[kWords];
#pragma unroll
      for (int k = 0; k < kWords; ++k) a[k] = __ldg(cw + g * kWords + k);
      if (p.use_diff) {
#pragma unroll
        for (int k = 0; k < kWords; ++k) b[k] = __ldg(pw + g * kWords + k);
      }
      int c8[8], q8[8];
      if (BITS == 12) {
        decode12x8(a[0], a[1], a[2], c8);
        if (p.use_diff) decode12x8(b[0], b[1], b[2], q8);
      } else if (BITS == 16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          c8[2 * k] = a[k] & 0xFFFF;
          c8[2 * k + 1] = a[k] >> 16;
          if (p.use_diff) {
            q8[2 * k] = b[k] & 0xFFFF;
            q8[2 * k + 1] = b[k] >> 16;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c8[k] = (a[k >> 2] >> (8 * (k & 3))) & 0xFF;
          if (p.use_diff) q8[k] = (b[k >> 2] >> (8 * (k & 3))) & 0xFF;
        }
      }
      int v8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int v = max(c8[k] - bg, 0);
        if (p.use_diff) {
          v -= max(q8[k] - bg, 0);
          if (v < p.diff_thr) v = 0;
        }
        v8[k] = v;
      }
      int4* dst = reinterpret_cast<int4*>(prof + 8 * g);
      dst[0] = make_int4(v8[0], v8[1], v8[2], v8[3]);
      dst[1] = make_int4(v8[4], v8[5], v8[6], v8[7]);
      if (p.profile_out != nullptr) {
        int* po = p.profile_out + (int64_t)f * W + 8 * g;
#pragma unroll
        for (int k = 0; k < 8; ++k) po[k] = v8[k];
      }
    }
  } else {
    warp_copy_bytes(raw_cur, cur_row, rs.nbytes, lane);
    if (p.use_diff) warp_copy_bytes(raw_pri, pri_row, rs.nbytes, lane);
    __syncwarp();
    for (int x = lane; x < W; x += 32) {
      const int64_t q = rs.q0 + x - rs.qbase;
      int v = max(load_px_generic<BITS>(raw_cur, q) - bg, 0);
      if (p.use_diff) {
        v -= max(load_px_generic<BITS>(raw_pri, q) - bg, 0);
        if (v < p.diff_thr) v = 0;
      }
      prof[x] = v;
      if (p.profile_out != nullptr) p.profile_out[(int64_t)f * W + x] = v;
    }
  }
  __syncwarp();
  if (empty) return pos;

  const int nchunk = (W + 31) >> 5;
  if (p.method == FF_METHOD_HALF_MAXIMUM) {
    // first arg-max: maximise (value, -x)
    long long best = LLONG_MIN;
    for (int x = lane; x < W; x += 32) {
      const long long key = ((long long)prof[x] << 32) | (unsigned)(0x7FFFFFFF - x);
      best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long other = __shfl_xor_sync(full, best, o);
      best = other > best ? other : best;
    }
    const int peak = (int)(best >> 32);
    const int k = 0x7FFFFFFF - (int)(best & 0xFFFFFFFFll);
    if (peak > 0) {
      for (int c = (k + 1) >> 5; c < nchunk; ++c) {
        const int x = (c << 5) + lane;
        const bool below = x > k && x < W && 2 * prof[x] < peak;
        const unsigned bits = __ballot_sync(full, below);
        if (bits) {
          pos = (c << 5) + __ffs(bits) - 1;
          break;
        }
      }
    }
  } else if (p.method == FF_METHOD_GRADIENT) {
    // np.gradient: central (f[i+1]-f[i-1])/2, one-sided at both ends; compare 2*g.
    long long best = LLONG_MAX;
    for (int x = lane; x < W; x += 32) {
      int g2;
      if (x == 0) g2 = 2 * (prof[1] - prof[0]);
      else if (x == W - 1) g2 = 2 * (prof[W - 1] - prof[W - 2]);
      else g2 = prof[x + 1] - prof[x - 1];
      const long long key = ((long long)g2 << 32) | (unsigned)x;
      best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long other = __shfl_xor_sync(full, best, o);
      best = other < best ? other : best;
    }
    const int g2min = (int)(best >> 32);
    if (g2min < p.grad2_bound) pos = (int)(best & 0xFFFFFFFFll);
  } else {
    // rightmost run of (p > T) with length >= min_run, scanned right to left
    bool in_run = false;
    int run_end = -1, run_len = 0;
    for (int c = nchunk - 1; c >= 0 && pos < 0; --c) {
      const int x = (c << 5) + lane;
      const unsigned w = __ballot_sync(full, x < W && prof[x] > thr_floor);
      int hi = 31;  // next bit to examine
      while (hi >= 0) {
        if (in_run) {
          const unsigned shifted = w << (31 - hi);
          const int ones = __clz((int)~shifted);  // leading ones from bit hi downward
          run_len += ones;
          hi -= ones;
          if (hi >= 0) {  // a zero bit ended the run inside this word
            if (run_len >= p.min_run) {
              pos = run_end;
              break;
            }
            in_run = false;
          }
        } else {
          const unsigned masked = hi == 31 ? w : (w & ((2u << hi) - 1u));
          if (!masked) break;
          hi = 31 - __clz((int)masked);
          run_end = (c << 5) + hi;
          run_len = 0;
          in_run = true;
        }
      }
    }
    if (pos < 0 && in_run && run_len >= p.min_run) pos = run_end;
  }
  return pos;
}

// Lane 0 of the detecting warp: store the position; an exit frame (scripts/process_videos.py:1488-1489)
// goes into the range's first-exit min and - in a range-sharded run - into every rank's global exit
// word, which the host-streamed path polls between chunks to stop uploading (:1494 across ranks).
__device__ __forceinline__ void commit_position(const DetectParams& p, int f, int pos) {
  p.pos_out[f] = pos;
  if (pos >= 0 && pos >= p.width - p.exit_margin) {
    const int fg = (int)(p.first_frame + f);
    atomicMin(p.first_exit, fg);
    if (p.hooks.table != nullptr) {
      const PeerTable* t = p.hooks.table;
      const int64_t off = t->exit_off + (p.hooks.epoch & 1);
      if (fg < ld_relaxed_sys(t->base[p.hooks.rank] + off))
        for (int r = 0; r < p.hooks.world; ++r) red_min_sys(t->base[r] + off, fg);
    }
  }
}

// Range-sharded runs: has ANY rank already seen the flame leave before frame f of this range?  Then the frame
// will be dropped by the merge whatever its position is (README.md:145-149; the reference never looks at it,
// scripts/process_videos.py:1494) and the detecting warp need not resolve it.  Reads this rank's copy of the
// clip-global exit word, which the peers' detecting warps keep up to date (commit_position).
__device__ __forceinline__ bool behind_global_exit(const DetectParams& p, int f) {
  if (p.hooks.table == nullptr) return false;
  const PeerTable* t = p.hooks.table;
  return (int64_t)ld_relaxed_sys(t->base[p.hooks.rank] + t->exit_off + (p.hooks.epoch & 1)) < p.first_frame + f;
}

// Called by one whole warp of every CTA when the CTA has written all its results.  The warp of the
// CTA that arrives last finishes the range: frames at or after the range's first exit frame are
// dropped (README.md:145-149) and, in a range-sharded run, this rank's block is published to its
// peers (st.release.sys of the epoch into their flag rows; merge_ranges_kernel waits for it).
__device__ __forceinline__ void range_tail(const DetectParams& p, int lane) {
  if (p.ws == nullptr) return;
  unsigned last = 0;
  if (lane == 0) {
    __threadfence();
    last = atomicAdd(&p.ws->tail_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  last = __shfl_sync(0xFFFFFFFFu, last, 0);
  if (!last) return;
  __threadfence();
  if (lane == 0) p.ws->tail_ticket = 0;              // left zero for the next launch
  if (p.truncate) {
    const int64_t fe = atomicMin(p.first_exit, INT_MAX);     // coherent read
    int64_t i0 = fe - p.first_frame;
    if (i0 < 0) i0 = 0;
    for (int64_t i = i0 + lane; i < p.n_frames; i += 32) p.pos_out[i] = FF_POS_DROPPED;
  }
  if (p.hooks.table != nullptr && (p.hooks.flags & FF_HOOK_PUBLISH)) {
    __syncwarp();                                    // the truncation stores of all lanes come first
    __threadfence_system();                          // ONE system-scope fence, then relaxed flag stores, one lane per peer
    const PeerTable* t = p.hooks.table;              // (a release store per peer would repeat the fence world times)
    for (int r = lane; r < p.hooks.world; r += 32) st_relaxed_sys(t->base[r] + t->flags_off + p.hooks.rank, p.hooks.epoch);
  }
}
#endif  // __CUDACC__

}  // namespace ff

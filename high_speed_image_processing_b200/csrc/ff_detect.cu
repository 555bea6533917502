// Stage 2a (background reduction), stage 3 (warp-per-profile detection) and stage 4
// (first-exit min + truncation).
//
// ff_detect gives every frame one warp.  The warp stages the packed centre row of its frame
// (and of the prior frame when the profile is a frame difference) in shared memory, builds
// the int32 profile there, and resolves the detection_method with warp primitives:
//   half_maximum  shuffle arg-max (first maximum) + ballot/ffs scan for the first sample
//                 right of the peak with 2*p < peak
//   gradient      shuffle arg-min of the doubled central difference (np.gradient semantics)
//   threshold     ballot scan from the right for the rightmost run of p > T
// All quantities are integers (see ff_stream.cu), so the answers are bit-exact.
#include <climits>
#include "ff_common.cuh"

namespace ff {
namespace {

constexpr int kDetectWarps = 8;

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
  int grad2_bound;
  int min_run;
  int exit_margin;
  const uint8_t* skip;
  int32_t* pos_out;
  int32_t* count_out;
  int32_t* first_exit;
  int32_t* profile_out;
  int raw_stride;  // bytes reserved per staged row (multiple of 4)
};

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

template <int BITS>
__global__ void __launch_bounds__(kDetectWarps * 32) detect_kernel(const DetectParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int W = p.width;
  int* prof = reinterpret_cast<int*>(smem) + (size_t)warp * W;
  uint8_t* raw_base = smem + (size_t)kDetectWarps * W * sizeof(int) + (size_t)warp * 2 * p.raw_stride;
  uint8_t* raw_cur = raw_base;
  uint8_t* raw_pri = raw_base + p.raw_stride;
  const unsigned full = 0xFFFFFFFFu;

  const int bg = __ldg(p.bg_dev);
  const int row = p.height / 2;
  // Flat pixel index of the centre row inside a frame, and the byte span that holds it.
  const int64_t q0 = (int64_t)row * W;
  int64_t byte_lo, byte_hi;
  if (BITS == 12) {
    byte_lo = (q0 >> 1) * 3;
    byte_hi = ((q0 + W - 1) >> 1) * 3 + 3;
  } else {
    byte_lo = q0 * (BITS / 8);
    byte_hi = (q0 + W) * (BITS / 8);
  }
  const int nbytes = (int)(byte_hi - byte_lo);
  const int64_t qbase = (BITS == 12) ? (q0 & ~(int64_t)1) : q0;  // pixel held by raw byte 0

  // Frames are dealt to warps in strides: warp w looks at frames w, w + nw, w + 2 nw, ... 32 at a
  // time, one candidate per lane.  Every lane sums its candidate's partial counts (the empty-frame
  // decision of :759-763) in parallel; only the frames that need a profile - the few percent that
  // hold a flame, and they are consecutive, so the stride spreads them over all warps - are then
  // handled one by one by the whole warp.
  const int nw = gridDim.x * kDetectWarps;
  const int w = blockIdx.x * kDetectWarps + warp;
  const bool vec = (p.partials_per_frame & 3) == 0 && (reinterpret_cast<uintptr_t>(p.partial) & 15u) == 0;
  for (int64_t base = w; base < p.n_frames; base += (int64_t)32 * nw) {
    const int64_t fc = base + (int64_t)lane * nw;
    const bool valid = fc < p.n_frames;
    int cnt_c = 0;
    bool need_c = false;
    if (valid) {
      // above-noise pixel count (is_empty_frame, scripts/process_videos.py:759-763)
      const int32_t* pc = p.partial + fc * p.partials_per_frame;
      if (vec) {
        const int4* p4 = reinterpret_cast<const int4*>(pc);
#pragma unroll 8
        for (int t = 0; t < p.partials_per_frame / 4; ++t) {
          const int4 v = __ldg(p4 + t);
          cnt_c += (v.x + v.y) + (v.z + v.w);
        }
      } else {
        for (int t = 0; t < p.partials_per_frame; ++t) cnt_c += __ldg(pc + t);
      }
      if (p.count_out != nullptr) p.count_out[fc] = cnt_c;
      const bool skipped_c = p.skip != nullptr && p.skip[fc] != 0;
      const bool empty_c = (int64_t)cnt_c < p.min_signal_count;
      // without a prior frame there is no difference profile: only frame 0 of a range without halo
      // (or a range that starts with skipped frames) - resolved exactly in the per-frame path below
      need_c = !skipped_c && (!empty_c || p.profile_out != nullptr);
      if (!need_c) p.pos_out[fc] = FF_POS_NONE;
    }
    unsigned pending = __ballot_sync(full, need_c);
    while (pending) {
    const int src = __ffs((int)pending) - 1;
    pending &= pending - 1;
    const int f = (int)(base + (int64_t)src * nw);
    const int cnt = __shfl_sync(full, cnt_c, src);
    const bool skipped = false;
    const bool empty = (int64_t)cnt < p.min_signal_count;

    // prior frame: the latest non-skipped frame before f (:469, :1462, :1443-1445)
    const uint8_t* prior = nullptr;
    if (p.use_diff) {
      int hf = f - 1;
      if (p.skip != nullptr)
        while (hf >= 0 && p.skip[hf]) --hf;
      prior = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
    }
    const bool have_profile = !skipped && (!p.use_diff || prior != nullptr);

    int pos = FF_POS_NONE;
    if (have_profile && (!empty || p.profile_out != nullptr)) {
      __syncwarp();
      warp_copy_bytes(raw_cur, p.frames + (int64_t)f * p.frame_bytes + byte_lo, nbytes, lane);
      if (p.use_diff) warp_copy_bytes(raw_pri, prior + byte_lo, nbytes, lane);
      __syncwarp();
      for (int x = lane; x < W; x += 32) {
        const int64_t q = q0 + x - qbase;
        int v = max(load_px_generic<BITS>(raw_cur, q) - bg, 0);
        if (p.use_diff) {
          v -= max(load_px_generic<BITS>(raw_pri, q) - bg, 0);
          if (v < p.diff_thr) v = 0;
        }
        prof[x] = v;
        if (p.profile_out != nullptr) p.profile_out[(int64_t)f * W + x] = v;
      }
      __syncwarp();

      if (!empty) {
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
            const unsigned w = __ballot_sync(full, x < W && prof[x] > p.threshold_floor);
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
      }
    }
    if (lane == 0) {
      p.pos_out[f] = pos;
      if (pos >= 0 && pos >= W - p.exit_margin)  // scripts/process_videos.py:1488-1489
        atomicMin(p.first_exit, (int)(p.first_frame + f));
    }
    }  // pending frames of this round
  }
}

template <int BITS>
__global__ void background_kernel(const uint8_t* __restrict__ frame, int64_t n_px, int height, int width,
                                  int32_t* bg_max, uint16_t* centerline) {
  int m = 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_px; q += (int64_t)gridDim.x * blockDim.x)
    m = max(m, load_px_generic<BITS>(frame, q));
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(bg_max, m);
  if (centerline != nullptr && blockIdx.x == 0) {
    const int64_t q0 = (int64_t)(height / 2) * width;
    for (int x = threadIdx.x; x < width; x += blockDim.x)
      centerline[x] = (uint16_t)load_px_generic<BITS>(frame, q0 + x);
  }
}

__global__ void truncate_kernel(int32_t* pos, int64_t n, int64_t first_frame, const int32_t* first_exit) {
  const int64_t fe = *first_exit;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (first_frame + i >= fe) pos[i] = FF_POS_DROPPED;
}

}  // namespace

int detect_impl(const void* frames, const void* halo, int64_t n_frames, int64_t first_frame, int height,
                int width, int bits, const int32_t* bg_dev, const int32_t* partial, int64_t min_signal_count,
                int method, int use_frame_diff, int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound,
                int32_t min_run_px, int32_t exit_margin_px, const uint8_t* skip, int32_t* pos_out,
                int32_t* count_out, int32_t* first_exit, int32_t* profile_out, cudaStream_t st) {
  if (frames == nullptr || bg_dev == nullptr || partial == nullptr || pos_out == nullptr || first_exit == nullptr)
    return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width <= 0 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (first_frame < 0 || first_frame + n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  if (method < FF_METHOD_THRESHOLD || method > FF_METHOD_HALF_MAXIMUM) return FF_ERR_INVALID;
  if (method == FF_METHOD_GRADIENT && width < 2) return FF_ERR_INVALID;  // np.gradient needs 2 samples
  if (min_run_px < 1) return FF_ERR_INVALID;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;

  DetectParams p{};
  p.frames = static_cast<const uint8_t*>(frames);
  p.halo = static_cast<const uint8_t*>(halo);
  p.frame_bytes = frame_bytes_of(px, bits);
  p.px_per_frame = px;
  p.n_frames = (int)n_frames;
  p.first_frame = first_frame;
  p.height = height;
  p.width = width;
  p.bg_dev = bg_dev;
  p.partial = partial;
  p.partials_per_frame = choose_tiling(px).partials_per_frame;
  p.min_signal_count = min_signal_count;
  p.method = method;
  p.use_diff = use_frame_diff ? 1 : 0;
  p.diff_thr = diff_thr;
  p.threshold_floor = threshold_floor;
  p.grad2_bound = grad2_bound;
  p.min_run = min_run_px;
  p.exit_margin = exit_margin_px;
  p.skip = skip;
  p.pos_out = pos_out;
  p.count_out = count_out;
  p.first_exit = first_exit;
  p.profile_out = profile_out;
  const int row_bytes = (bits == 12) ? (width / 2 + 2) * 3 : width * (bits / 8);
  p.raw_stride = (row_bytes + 3 + 15) & ~15;

  const size_t smem = (size_t)kDetectWarps * width * sizeof(int) + (size_t)kDetectWarps * 2 * p.raw_stride;
  if (smem > 200 * 1024) return FF_ERR_UNSUPPORTED;  // W beyond ~5k columns: not a Photron sensor
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024)
      FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one resident wave of warps; each takes every nw-th frame
    int occ = 1, sms = 148, dev = 0;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDetectWarps * 32, smem));
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (int64_t)sms * (occ > 0 ? occ : 1);
    const int64_t enough = (n_frames + kDetectWarps - 1) / kDetectWarps;
    if (blocks > enough) blocks = enough;
    kern<<<(unsigned)blocks, kDetectWarps * 32, smem, st>>>(p);
    FF_CUDA_TRY(cudaGetLastError());
    return FF_OK;
  };
  switch (bits) {
    case 8: return launch(detect_kernel<8>);
    case 12: return launch(detect_kernel<12>);
    default: return launch(detect_kernel<16>);
  }
}

int background_impl(const void* frame0, int height, int width, int bits, int32_t* bg_max, uint16_t* centerline,
                    cudaStream_t st) {
  if (frame0 == nullptr || bg_max == nullptr) return FF_ERR_INVALID;
  if (height <= 0 || width <= 0) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;
  FF_CUDA_TRY(cudaMemsetAsync(bg_max, 0, sizeof(int32_t), st));
  int64_t blocks = (px + 1023) / 1024;
  if (blocks > 296) blocks = 296;
  const uint8_t* f = static_cast<const uint8_t*>(frame0);
  switch (bits) {
    case 8: background_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(f, px, height, width, bg_max, centerline); break;
    case 12: background_kernel<12><<<(unsigned)blocks, 256, 0, st>>>(f, px, height, width, bg_max, centerline); break;
    default: background_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(f, px, height, width, bg_max, centerline); break;
  }
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

int truncate_impl(int32_t* pos, int64_t n_frames, int64_t first_frame, const int32_t* first_exit, cudaStream_t st) {
  if (pos == nullptr || first_exit == nullptr || n_frames < 0) return FF_ERR_INVALID;
  if (n_frames == 0) return FF_OK;
  int64_t blocks = (n_frames + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  truncate_kernel<<<(unsigned)blocks, 256, 0, st>>>(pos, n_frames, first_frame, first_exit);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

}  // namespace ff

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
#include "ff_detect_core.cuh"

namespace ff {
namespace {

constexpr int kDetectWarps = 8;

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

  griddep_wait();                       // partial counts / scalars come from the kernels before this one
  const int bg = __ldg(p.bg_dev);
  const int thr_floor = p.threshold_dev != nullptr ? __ldg(p.threshold_dev) : p.threshold_floor;
  const RowSpan rs = centre_row_span<BITS>(p.height, W);

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
      // (or a range that starts with skipped frames) - resolved exactly in the per-frame path
      need_c = !skipped_c && (!empty_c || p.profile_out != nullptr);
      if (!need_c) p.pos_out[fc] = FF_POS_NONE;
    }
    unsigned pending = __ballot_sync(full, need_c);
    while (pending) {
      const int src = __ffs((int)pending) - 1;
      pending &= pending - 1;
      const int f = (int)(base + (int64_t)src * nw);
      const int cnt = __shfl_sync(full, cnt_c, src);
      const bool empty = (int64_t)cnt < p.min_signal_count;
      int skip_it = 0;
      if (lane == 0) skip_it = p.profile_out == nullptr && behind_global_exit(p, f);
      if (__shfl_sync(full, skip_it, 0)) {                 // another rank saw the exit before this frame: it will be dropped
        if (lane == 0) p.pos_out[f] = FF_POS_NONE;
        continue;
      }
      const int pos = detect_one_frame<BITS>(p, rs, f, empty, bg, thr_floor, prof, raw_cur, raw_pri, lane);
      if (lane == 0) commit_position(p, f, pos);
    }
  }
  // last CTA of the launch: local truncation, publication to the peers
  if (p.ws != nullptr) {
    __syncthreads();
    if (warp == 0) range_tail(p, lane);
  }
}

// ---- prep_kernel: everything a range needs before its frames are streamed -----------------------
//   * (range-sharded runs) wait until every peer has finished reading the block this epoch will
//     overwrite (their acks of epoch-2), see csrc/ff_exchange.cu;
//   * the range's first-exit word = FF_NO_EXIT;
//   * the clip's background scalar = max of frame 0 (scripts/process_videos.py:1357-1358), the centre
//     row of frame 0 (:1361-1362) and - for the threshold method - the float64 statistics of that row
//     and the flame threshold max(mean + 5 std, 2 max) (:1363-1370), evaluated in NumPy's own order
//     of operations (pairwise summation of add.reduce) so that the threshold is the bit-identical
//     float64; the host repeats the computation with NumPy itself as a cross-check.
// The kernel lets its dependents start at once (PDL): the streaming kernel behind it brings its
// CTAs up and its producer warps begin to fetch tiles while this one runs.
struct PrepParams {
  const uint8_t* frame0;     // nullptr: scalars are already known
  int64_t n_px;
  int height, width;
  int32_t* scalars;          // int32[16]: [0] bg, [1] floor(flame threshold), [2] status; float64 at byte 16: mean, std, max, threshold
  uint16_t* centerline;      // uint16[W], nullable
  int want_stats;
  int32_t* first_exit;       // nullable
  RangeWorkspace* ws;
  RangeHooks hooks;
  // NumPy's pairwise summation of W values as tables (a function of W only, built on the host):
  // leaves [off, off + len) and the additions that combine their sums, in the order of the recursion
  int n_leaves;
  uint16_t leaf_off[96];
  uint8_t leaf_len[96];      // <= 128
  uint8_t op_a[96], op_b[96];   // sum[op_a[k]] += sum[op_b[k]], k = 0 .. n_leaves - 2; the total ends in sum[0]
};

constexpr int kPrepThreads = 256;
constexpr int kMaxLeaves = 96;      // W <= 4096: at most 64 leaves of NumPy's pairwise summation

// np.add.reduce over n float64 values = pairwise_sum(a, n) (numpy/_core/src/umath/loops_utils.h.src):
// n < 8 plain loop; n <= 128 eight strided accumulators combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
// plus the n % 8 tail; else split at n2 = n/2 - (n/2) % 8.  The leaves (n <= 128) are summed by groups of
// eight threads, one accumulator each; one thread then adds the leaf sums in the order of the recursion.
// Host side: the recursion as tables (PrepParams::leaf_*, op_*).  Returns the number of leaves.
int pairwise_tables(int n, uint16_t* off, uint8_t* len, uint8_t* op_a, uint8_t* op_b) {
  int n_leaves = 0, n_ops = 0;
  // post-order walk; every node returns the index of the leaf slot that accumulates its sum (its first leaf)
  struct Walk {
    uint16_t* off; uint8_t* len; uint8_t* a; uint8_t* b; int* nl; int* no;
    int go(int o, int m) {
      if (m <= 128) {
        off[*nl] = (uint16_t)o;
        len[*nl] = (uint8_t)m;
        return (*nl)++;
      }
      int n2 = m / 2;
      n2 -= n2 % 8;
      const int l = go(o, n2);
      const int r = go(o + n2, m - n2);
      a[*no] = (uint8_t)l;
      b[*no] = (uint8_t)r;
      ++*no;
      return l;
    }
  } w{off, len, op_a, op_b, &n_leaves, &n_ops};
  w.go(0, n);
  return n_leaves;
}

// Whole CTA: np.add.reduce(xs[0:n]) with the tables in p.  leaf_sum: shared array; result valid in every thread.
__device__ double cta_pairwise_sum(const PrepParams& p, const double* xs, double* leaf_sum) {
  const int gid = threadIdx.x >> 3, j = threadIdx.x & 7;
  for (int it = 0; it * (kPrepThreads / 8) < p.n_leaves; ++it) {
    const int L = it * (kPrepThreads / 8) + gid;
    const bool valid = L < p.n_leaves;
    const int off = valid ? p.leaf_off[L] : 0;
    const int len = valid ? p.leaf_len[L] : 0;
    double r = 0.0;
    const int body = len - (len % 8);
    if (len >= 8) {
      r = xs[off + j];
      for (int i = 8; i < body; i += 8) r = __dadd_rn(r, xs[off + i + j]);
    }
    r = __dadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 4));
    if (valid && j == 0) {
      if (len < 8) {
        r = 0.0;
        for (int i = 0; i < len; ++i) r = __dadd_rn(r, xs[off + i]);
      } else {
        for (int i = body; i < len; ++i) r = __dadd_rn(r, xs[off + i]);
      }
      leaf_sum[L] = r;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k + 1 < p.n_leaves; ++k) leaf_sum[p.op_a[k]] = __dadd_rn(leaf_sum[p.op_a[k]], leaf_sum[p.op_b[k]]);
  __syncthreads();
  const double total = leaf_sum[0];
  __syncthreads();          // leaf_sum is reused by the next reduction
  return total;
}

template <int BITS>
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const PrepParams p) {
  griddep_launch_dependents();
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_max[kPrepThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x;

  if (blockIdx.x == 0) {
    if (p.hooks.table != nullptr && (p.hooks.flags & FF_HOOK_WAIT) && tid < p.hooks.world) {
      // the block (and exit word) of this epoch's parity was last used two epochs ago
      const PeerTable* t = p.hooks.table;
      const int32_t* ack = t->base[p.hooks.rank] + t->acks_off + tid;
      const long long t0 = clock64();
      while (ld_acquire_sys(ack) < p.hooks.epoch - 2) {
        if (clock64() - t0 > p.hooks.spin_limit) {
          atomicExch(t->base[p.hooks.rank] + t->status_off, 0x100 + tid);
          break;
        }
      }
    }
    __syncthreads();
    if (tid == 0 && p.first_exit != nullptr) *p.first_exit = FF_NO_EXIT;
  }
  if (p.frame0 == nullptr) return;

  // ---- max of frame 0: every CTA a slice, the last one to finish combines --------------------------
  int m = 0;
  const bool words = BITS == 12 && (p.n_px & 7) == 0 && (reinterpret_cast<uintptr_t>(p.frame0) & 3u) == 0;
  if (words) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p.frame0);
    for (int64_t g = (int64_t)blockIdx.x * kPrepThreads + tid; g < p.n_px / 8; g += (int64_t)gridDim.x * kPrepThreads) {
      int v[8];
      decode12x8(__ldg(w + 3 * g), __ldg(w + 3 * g + 1), __ldg(w + 3 * g + 2), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m = max(m, v[k]);
    }
  } else {
    for (int64_t q = (int64_t)blockIdx.x * kPrepThreads + tid; q < p.n_px; q += (int64_t)gridDim.x * kPrepThreads)
      m = max(m, load_px_generic<BITS>(p.frame0, q));
  }
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((tid & 31) == 0) s_max[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    int mm = 0;
    for (int k = 0; k < kPrepThreads / 32; ++k) mm = max(mm, s_max[k]);
    p.ws->prep_max[blockIdx.x] = mm;
    __threadfence();
    s_last = atomicAdd(&p.ws->prep_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) {
    int mm = 0;
    for (unsigned k = 0; k < gridDim.x; ++k) mm = max(mm, ((volatile int32_t*)p.ws->prep_max)[k]);
    p.scalars[0] = mm;
    p.ws->prep_ticket = 0;
  }
  const int W = p.width;
  const int64_t q0 = (int64_t)(p.height / 2) * W;
  if (p.centerline != nullptr)
    for (int x = tid; x < W; x += kPrepThreads) p.centerline[x] = (uint16_t)load_px_generic<BITS>(p.frame0, q0 + x);
  if (!p.want_stats) return;

  // ---- float64 statistics of the centre row, NumPy's order of operations ------------------------------
  double* xs = reinterpret_cast<double*>(smem);                 // [W]
  double* leaf_sum = xs + W;                                    // [kMaxLeaves]
  int mx = 0;
  for (int x = tid; x < W; x += kPrepThreads) {
    const int v = load_px_generic<BITS>(p.frame0, q0 + x);
    xs[x] = (double)v;
    mx = max(mx, v);
  }
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  if ((tid & 31) == 0) s_max[tid >> 5] = mx;
  __syncthreads();
  mx = 0;
  for (int k = 0; k < kPrepThreads / 32; ++k) mx = max(mx, s_max[k]);
  const double mean = __ddiv_rn(cta_pairwise_sum(p, xs, leaf_sum), (double)W);
  for (int x = tid; x < W; x += kPrepThreads) {
    const double d = __dsub_rn(xs[x], mean);                    // arr - arrmean
    xs[x] = __dmul_rn(d, d);                                    // multiply(x, x)
  }
  __syncthreads();
  const double var = __ddiv_rn(cta_pairwise_sum(p, xs, leaf_sum), (double)W);
  if (tid == 0) {
    const double sd = __dsqrt_rn(var);
    const double a = __dadd_rn(mean, __dmul_rn(5.0, sd));       // mean + 5 * std           (:1367)
    const double b = __dmul_rn((double)mx, 2.0);                // centerline_max * 2.0      (:1369)
    const double thr = b > a ? b : a;                           // max(a, b)
    double fl = floor(thr);
    if (fl > 2147483647.0) fl = 2147483647.0;
    p.scalars[1] = (int32_t)fl;
    double* out = reinterpret_cast<double*>(p.scalars + 4);
    out[0] = mean;
    out[1] = sd;
    out[2] = (double)mx;
    out[3] = thr;
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

int make_detect_params(DetectParams* out, const void* frames, const void* halo, int64_t n_frames,
                       int64_t first_frame, int height, int width, int bits, const int32_t* bg_dev,
                       const int32_t* partial, int64_t min_signal_count, int method, int use_frame_diff,
                       int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px,
                       int32_t exit_margin_px, const uint8_t* skip, int32_t* pos_out, int32_t* count_out,
                       int32_t* first_exit, int32_t* profile_out) {
  if (frames == nullptr || bg_dev == nullptr || pos_out == nullptr || first_exit == nullptr) return FF_ERR_INVALID;
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
  p.raw_stride = detect_row_stride(width, bits);
  *out = p;
  return FF_OK;
}

int launch_detect(const DetectParams& p, int bits, bool pdl, cudaStream_t st) {
  if (p.partial == nullptr) return FF_ERR_INVALID;
  const size_t smem = detect_warp_smem(p.width, bits) * kDetectWarps;
  if (smem > 200 * 1024) return FF_ERR_UNSUPPORTED;  // W beyond ~3.5k columns: not a Photron sensor
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024)
      FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one resident wave of warps; each takes every nw-th frame
    int occ = 1, sms = 148, dev = 0;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDetectWarps * 32, smem));
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (int64_t)sms * (occ > 0 ? occ : 1);
    const int64_t enough = ((int64_t)p.n_frames + kDetectWarps - 1) / kDetectWarps;
    if (blocks > enough) blocks = enough;
    return launch_kernel(kern, dim3((unsigned)blocks), dim3(kDetectWarps * 32), smem, st, pdl, p);
  };
  switch (bits) {
    case 8: return launch(detect_kernel<8>);
    case 12: return launch(detect_kernel<12>);
    default: return launch(detect_kernel<16>);
  }
}

int detect_impl(const void* frames, const void* halo, int64_t n_frames, int64_t first_frame, int height,
                int width, int bits, const int32_t* bg_dev, const int32_t* partial, int64_t min_signal_count,
                int method, int use_frame_diff, int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound,
                int32_t min_run_px, int32_t exit_margin_px, const uint8_t* skip, int32_t* pos_out,
                int32_t* count_out, int32_t* first_exit, int32_t* profile_out, cudaStream_t st) {
  DetectParams p{};
  const int rc = make_detect_params(&p, frames, halo, n_frames, first_frame, height, width, bits, bg_dev, partial,
                                    min_signal_count, method, use_frame_diff, diff_thr, threshold_floor, grad2_bound,
                                    min_run_px, exit_margin_px, skip, pos_out, count_out, first_exit, profile_out);
  if (rc != FF_OK) return rc;
  return launch_detect(p, bits, false, st);
}

// Launches prep_kernel.  frame0 may be null (scalars known: only the first-exit word and the peers' acks).
int prep_impl(const void* frame0, int height, int width, int bits, int32_t* scalars, uint16_t* centerline,
              int want_stats, int32_t* first_exit, RangeWorkspace* ws, const RangeHooks& hooks, cudaStream_t st) {
  if (height <= 0 || width <= 0) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;
  if (frame0 != nullptr && (scalars == nullptr || ws == nullptr)) return FF_ERR_INVALID;
  if (frame0 == nullptr && first_exit == nullptr && hooks.table == nullptr) return FF_OK;      // nothing to do
  if (want_stats && width > 4096) return FF_ERR_UNSUPPORTED;
  PrepParams p{};
  p.frame0 = static_cast<const uint8_t*>(frame0);
  p.n_px = px;
  p.height = height;
  p.width = width;
  p.scalars = scalars;
  p.centerline = centerline;
  p.want_stats = frame0 != nullptr && want_stats;
  p.first_exit = first_exit;
  p.ws = ws;
  p.hooks = hooks;
  if (p.want_stats) p.n_leaves = pairwise_tables(width, p.leaf_off, p.leaf_len, p.op_a, p.op_b);
  int64_t blocks = 1;
  if (frame0 != nullptr) {
    blocks = (px + kPrepThreads * 32 - 1) / (kPrepThreads * 32);        // ~32 pixels per thread
    if (blocks > kPrepMaxCtas) blocks = kPrepMaxCtas;
    if (blocks < 1) blocks = 1;
  }
  const size_t smem = p.want_stats ? (size_t)width * 8 + kMaxLeaves * 8 : 0;
  static const bool once = (cudaFuncSetAttribute(prep_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                            cudaFuncSetAttribute(prep_kernel<12>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                            cudaFuncSetAttribute(prep_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
  (void)once;       // same shared-memory split as the range kernel it runs next to (PDL)
  switch (bits) {
    case 8: return launch_kernel(prep_kernel<8>, dim3((unsigned)blocks), dim3(kPrepThreads), smem, st, false, p);
    case 12: return launch_kernel(prep_kernel<12>, dim3((unsigned)blocks), dim3(kPrepThreads), smem, st, false, p);
    default: return launch_kernel(prep_kernel<16>, dim3((unsigned)blocks), dim3(kPrepThreads), smem, st, false, p);
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

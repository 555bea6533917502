// C-ABI entry points of libflamefront.so (declared in include/flamefront.h) and the
// host-resident streaming driver (ff_process_host).
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ff_common.cuh"

namespace ff {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  std::snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int stream_frames_impl(const void*, const void*, int64_t, int, int, int, const int32_t*, int32_t, int32_t,
                       const uint8_t*, int32_t*, void*, int, uint16_t*, cudaStream_t);
int unpack_impl(const void*, void*, int64_t, int, int, int, cudaStream_t);
int detect_impl(const void*, const void*, int64_t, int64_t, int, int, int, const int32_t*, const int32_t*, int64_t,
                int, int, int32_t, int32_t, int32_t, int32_t, int32_t, const uint8_t*, int32_t*, int32_t*, int32_t*,
                int32_t*, cudaStream_t);
int background_impl(const void*, int, int, int, int32_t*, uint16_t*, cudaStream_t);
int truncate_impl(int32_t*, int64_t, int64_t, const int32_t*, cudaStream_t);
int head_lines_impl(const void*, const void*, int64_t, int, int, int, const int32_t*, const int32_t*, int64_t, int32_t,
                    const double*, int, const uint8_t*, double*, uint8_t*, int32_t*, cudaStream_t);
int head_track_impl(const double*, const uint8_t*, int64_t, int64_t, int, int32_t, int32_t, int32_t, double, double,
                    int32_t, int32_t, int32_t, int32_t*, int32_t*, int32_t*, cudaStream_t);
int64_t head_track_scratch_len(int64_t);
void stream_copy(uint8_t* dst, const uint8_t* src, size_t n);      // ff_hostcopy.cpp
int frame_subtract_background_impl(const void*, int, int64_t, double, double*, cudaStream_t);
int frame_difference_impl(const void*, const void*, int, int64_t, double, double*, cudaStream_t);
int frame_three_difference_impl(const void*, const void*, const void*, int, int64_t, double, double*, cudaStream_t);
int frame_count_above_impl(const void*, int, int64_t, double, int64_t*, cudaStream_t);
int head_images_impl(const void*, const void*, int64_t, int, int, int, int32_t, int32_t, int32_t, int, const double*,
                     int, const uint8_t*, double*, double*, double*, double*, double*, double*, uint8_t*, cudaStream_t);

}  // namespace ff

using namespace ff;

// ---- parallel host copy: page cache / pageable memory -> pinned bounce buffer ---------------
// A memory-mapped .mraw file is pageable memory; cudaMemcpyAsync from it degrades to the driver's
// single-threaded staged copy (measured 11 GB/s on the B200 box, against 55 GB/s from pinned
// memory).  The streaming driver therefore moves pageable sources through its own pinned bounce
// buffers, filled by a small pool of threads that split every chunk between them, while the
// previous chunk's DMA and kernels are in flight.
class CopyPool {
 public:
  explicit CopyPool(int n_threads) {
    for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_job_.notify_all();
    for (auto& t : workers_) t.join();
  }
  // Blocking: returns when dst[0, bytes) is filled.  The caller takes part in the copy.
  void copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
    const size_t part = 2u << 20;
    {
      std::lock_guard<std::mutex> g(m_);
      dst_ = dst;
      src_ = src;
      bytes_ = bytes;
      part_ = part;
      n_parts_ = (bytes + part - 1) / part;
      next_.store(0);
      done_ = 0;
      ++gen_;
    }
    cv_job_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [this] { return done_ == n_parts_; });
  }

 private:
  void work() {
    size_t mine = 0;
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= n_parts_) break;
      const size_t a = i * part_;
      const size_t n = a + part_ <= bytes_ ? part_ : bytes_ - a;
      stream_copy(dst_ + a, src_ + a, n);
      ++mine;
    }
    if (mine) {
      std::lock_guard<std::mutex> g(m_);
      done_ += mine;
      if (done_ == n_parts_) cv_done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_job_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_job_, cv_done_;
  uint8_t* dst_ = nullptr;
  const uint8_t* src_ = nullptr;
  size_t bytes_ = 0, part_ = 0, n_parts_ = 0, done_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// ---- host-resident streaming context -----------------------------------------------------
struct ff_host_ctx {
  int device = 0;
  int64_t chunk_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t compute_stream = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  uint8_t* stage[2] = {nullptr, nullptr};
  int64_t stage_bytes = 0;
  int32_t* partial[2] = {nullptr, nullptr};
  int64_t partial_elems = 0;
  int32_t* pos_dev = nullptr;
  int32_t* cnt_dev = nullptr;
  uint8_t* skip_dev = nullptr;
  int64_t frames_cap = 0;
  int32_t* scalars_dev = nullptr;   // [0] bg, [1] first_exit
  int32_t* scalars_host = nullptr;  // pinned: [0] bg, [1] init value, [2..3] per-buffer first_exit readback
  // pageable sources only
  uint8_t* bounce[2] = {nullptr, nullptr};     // pinned, same layout as stage[]
  int64_t bounce_bytes = 0;
  cudaEvent_t bounce_free[2] = {nullptr, nullptr};
  CopyPool* pool = nullptr;
};

static int ctx_release(ff_host_ctx* c) {
  if (c == nullptr) return FF_OK;
  cudaSetDevice(c->device);
  for (int i = 0; i < 2; ++i) {
    if (c->stage[i]) cudaFree(c->stage[i]);
    if (c->bounce[i]) cudaFreeHost(c->bounce[i]);
    if (c->bounce_free[i]) cudaEventDestroy(c->bounce_free[i]);
    if (c->partial[i]) cudaFree(c->partial[i]);
    if (c->copied[i]) cudaEventDestroy(c->copied[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
  }
  if (c->pos_dev) cudaFree(c->pos_dev);
  if (c->cnt_dev) cudaFree(c->cnt_dev);
  if (c->skip_dev) cudaFree(c->skip_dev);
  if (c->scalars_dev) cudaFree(c->scalars_dev);
  if (c->scalars_host) cudaFreeHost(c->scalars_host);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->compute_stream) cudaStreamDestroy(c->compute_stream);
  delete c->pool;
  delete c;
  return FF_OK;
}

static int ctx_reserve_bounce(ff_host_ctx* c, int64_t bytes) {
  if (bytes > c->bounce_bytes) {
    for (int i = 0; i < 2; ++i) {
      if (c->bounce[i]) FF_CUDA_TRY(cudaFreeHost(c->bounce[i]));
      c->bounce[i] = nullptr;
      FF_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->bounce[i]), (size_t)bytes, cudaHostAllocDefault));
      if (c->bounce_free[i] == nullptr)
        FF_CUDA_TRY(cudaEventCreateWithFlags(&c->bounce_free[i], cudaEventDisableTiming));
    }
    c->bounce_bytes = bytes;
  }
  if (c->pool == nullptr) {
    int n = (int)std::thread::hardware_concurrency() / 2 - 1;       // the calling thread copies too
    if (const char* e = getenv("FF_HOST_COPY_THREADS")) n = atoi(e) - 1;
    if (n < 0) n = 0;
    if (n > 15) n = 15;
    c->pool = new (std::nothrow) CopyPool(n);
    if (c->pool == nullptr) return FF_ERR_INVALID;
  }
  return FF_OK;
}

static int ctx_reserve(ff_host_ctx* c, int64_t stage_bytes, int64_t partial_elems, int64_t frames) {
  if (stage_bytes > c->stage_bytes) {
    for (int i = 0; i < 2; ++i) {
      if (c->stage[i]) FF_CUDA_TRY(cudaFree(c->stage[i]));
      c->stage[i] = nullptr;
      FF_CUDA_TRY(cudaMalloc(&c->stage[i], (size_t)stage_bytes));
    }
    c->stage_bytes = stage_bytes;
  }
  if (partial_elems > c->partial_elems) {
    for (int i = 0; i < 2; ++i) {
      if (c->partial[i]) FF_CUDA_TRY(cudaFree(c->partial[i]));
      c->partial[i] = nullptr;
      FF_CUDA_TRY(cudaMalloc(&c->partial[i], sizeof(int32_t) * (size_t)partial_elems));
    }
    c->partial_elems = partial_elems;
  }
  if (frames > c->frames_cap) {
    if (c->pos_dev) FF_CUDA_TRY(cudaFree(c->pos_dev));
    if (c->cnt_dev) FF_CUDA_TRY(cudaFree(c->cnt_dev));
    if (c->skip_dev) FF_CUDA_TRY(cudaFree(c->skip_dev));
    c->pos_dev = c->cnt_dev = nullptr;
    c->skip_dev = nullptr;
    FF_CUDA_TRY(cudaMalloc(&c->pos_dev, sizeof(int32_t) * (size_t)frames));
    FF_CUDA_TRY(cudaMalloc(&c->cnt_dev, sizeof(int32_t) * (size_t)frames));
    FF_CUDA_TRY(cudaMalloc(&c->skip_dev, (size_t)frames));
    c->frames_cap = frames;
  }
  return FF_OK;
}

extern "C" {

int ff_abi_version(void) { return FF_ABI_VERSION; }

const char* ff_strerror(int status) {
  switch (status) {
    case FF_OK: return "ok";
    case FF_ERR_INVALID: return "invalid argument";
    case FF_ERR_UNSUPPORTED: return "unsupported bit depth, shape or dtype";
    case FF_ERR_CUDA: return "CUDA runtime error";
    case FF_ERR_NO_DEVICE: return "no CUDA device";
    case FF_ERR_ALIGNMENT: return "pointer must be 16-byte aligned";
    default: return "unknown status";
  }
}

const char* ff_last_cuda_error(void) { return g_cuda_err; }

int ff_device_count(int* count) {
  if (count == nullptr) return FF_ERR_INVALID;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaGetDeviceCount");
    *count = 0;
    return FF_ERR_NO_DEVICE;
  }
  *count = n;
  return FF_OK;
}

int ff_device_sm_count(int device, int* sm_count) {
  if (sm_count == nullptr) return FF_ERR_INVALID;
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaDeviceGetAttribute");
    return FF_ERR_NO_DEVICE;
  }
  *sm_count = n;
  return FF_OK;
}

int ff_partial_len(int64_t n_frames, int height, int width, int bits, int64_t* n_elems, int* partials_per_frame) {
  if (n_frames < 0 || height <= 0 || width <= 0) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const Tiling t = choose_tiling((int64_t)height * width);
  if (n_elems) *n_elems = n_frames * t.partials_per_frame;
  if (partials_per_frame) *partials_per_frame = t.partials_per_frame;
  return FF_OK;
}

int ff_unpack(const void* packed_dev, void* out_dev, int64_t n_frames, int height, int width, int bits,
              void* stream) {
  return unpack_impl(packed_dev, out_dev, n_frames, height, width, bits, static_cast<cudaStream_t>(stream));
}

int ff_background(const void* frame0_dev, int height, int width, int bits, int32_t* bg_max_dev,
                  uint16_t* centerline_dev, void* stream) {
  return background_impl(frame0_dev, height, width, bits, bg_max_dev, centerline_dev,
                         static_cast<cudaStream_t>(stream));
}

int ff_stream_frames(const void* frames_dev, const void* halo_dev, int64_t n_frames, int height, int width,
                     int bits, const int32_t* bg_dev, int32_t empty_thr, int32_t diff_thr,
                     const uint8_t* skip_dev, int32_t* partial_dev, void* diff_out_dev, int diff_dtype,
                     uint16_t* decoded_out_dev, void* stream) {
  return stream_frames_impl(frames_dev, halo_dev, n_frames, height, width, bits, bg_dev, empty_thr, diff_thr,
                            skip_dev, partial_dev, diff_out_dev, diff_dtype, decoded_out_dev,
                            static_cast<cudaStream_t>(stream));
}

int ff_detect(const void* frames_dev, const void* halo_dev, int64_t n_frames, int64_t first_frame, int height,
              int width, int bits, const int32_t* bg_dev, const int32_t* partial_dev, int64_t min_signal_count,
              int method, int use_frame_diff, int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound,
              int32_t min_run_px, int32_t exit_margin_px, const uint8_t* skip_dev, int32_t* pos_out_dev,
              int32_t* count_out_dev, int32_t* first_exit_dev, int32_t* profile_out_dev, void* stream) {
  return detect_impl(frames_dev, halo_dev, n_frames, first_frame, height, width, bits, bg_dev, partial_dev,
                     min_signal_count, method, use_frame_diff, diff_thr, threshold_floor, grad2_bound, min_run_px,
                     exit_margin_px, skip_dev, pos_out_dev, count_out_dev, first_exit_dev, profile_out_dev,
                     static_cast<cudaStream_t>(stream));
}

int ff_truncate(int32_t* pos_dev, int64_t n_frames, int64_t first_frame, const int32_t* first_exit_dev,
                void* stream) {
  return truncate_impl(pos_dev, n_frames, first_frame, first_exit_dev, static_cast<cudaStream_t>(stream));
}

int ff_head_lines(const void* frames_dev, const void* halo_dev, int64_t n_frames, int height, int width, int bits,
                  const int32_t* bg_dev, const int32_t* partial_dev, int64_t min_signal_count, int32_t diff_thr,
                  const double* gauss_weights_host, int radius, const uint8_t* skip_dev, double* lines_out_dev,
                  uint8_t* flags_out_dev, int32_t* scratch_dev, void* stream) {
  return head_lines_impl(frames_dev, halo_dev, n_frames, height, width, bits, bg_dev, partial_dev, min_signal_count,
                         diff_thr, gauss_weights_host, radius, skip_dev, lines_out_dev, flags_out_dev, scratch_dev,
                         static_cast<cudaStream_t>(stream));
}

int ff_head_track_scratch_len(int64_t n_frames, int64_t* n_elems) {
  if (n_elems == nullptr || n_frames <= 0) return FF_ERR_INVALID;
  *n_elems = head_track_scratch_len(n_frames);
  return FF_OK;
}

int ff_head_track(const double* lines_dev, const uint8_t* flags_dev, int64_t n_frames, int64_t first_frame, int width,
                  int32_t edge_margin_px, int32_t max_displacement_px, int32_t search_window_px,
                  double min_gradient_strength, double sobel_threshold_fraction, int32_t exit_margin_px,
                  int32_t last_frame_in, int32_t last_pos_in, int32_t* out_dev, int32_t* stop_dev,
                  int32_t* scratch_dev, void* stream) {
  return head_track_impl(lines_dev, flags_dev, n_frames, first_frame, width, edge_margin_px, max_displacement_px,
                         search_window_px, min_gradient_strength, sobel_threshold_fraction, exit_margin_px,
                         last_frame_in, last_pos_in, out_dev, stop_dev, scratch_dev,
                         static_cast<cudaStream_t>(stream));
}

int ff_frame_subtract_background(const void* image_dev, int px_type, int64_t n_px, double background,
                                 double* out_dev, void* stream) {
  return frame_subtract_background_impl(image_dev, px_type, n_px, background, out_dev, static_cast<cudaStream_t>(stream));
}

int ff_frame_difference(const void* current_dev, const void* prior_dev, int px_type, int64_t n_px, double threshold,
                        double* out_dev, void* stream) {
  return frame_difference_impl(current_dev, prior_dev, px_type, n_px, threshold, out_dev, static_cast<cudaStream_t>(stream));
}

int ff_frame_three_difference(const void* prev_dev, const void* curr_dev, const void* next_dev, int px_type,
                              int64_t n_px, double threshold, double* out_dev, void* stream) {
  return frame_three_difference_impl(prev_dev, curr_dev, next_dev, px_type, n_px, threshold, out_dev,
                                     static_cast<cudaStream_t>(stream));
}

int ff_frame_count_above(const void* frame_dev, int px_type, int64_t n_px, double threshold, int64_t* count_dev,
                         void* stream) {
  return frame_count_above_impl(frame_dev, px_type, n_px, threshold, count_dev, static_cast<cudaStream_t>(stream));
}

int ff_head_images(const void* frames_dev, const void* halo_dev, int64_t n_frames, int height, int width, int bits,
                   int32_t bg, int32_t bg_halo, int32_t diff_thr, int morphology_size, const double* gauss_weights_host,
                   int radius, const uint8_t* skip_dev, double* sub_out_dev, double* diff_out_dev, double* opened_out_dev,
                   double* blurred_out_dev, double* sobel_out_dev, double* gradient_out_dev, uint8_t* state_out_dev,
                   void* stream) {
  return head_images_impl(frames_dev, halo_dev, n_frames, height, width, bits, bg, bg_halo, diff_thr, morphology_size,
                          gauss_weights_host, radius, skip_dev, sub_out_dev, diff_out_dev, opened_out_dev,
                          blurred_out_dev, sobel_out_dev, gradient_out_dev, state_out_dev,
                          static_cast<cudaStream_t>(stream));
}

int ff_host_ctx_create(int device, int64_t chunk_bytes, ff_host_ctx** ctx_out) {
  if (ctx_out == nullptr || chunk_bytes <= 0) return FF_ERR_INVALID;
  *ctx_out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return FF_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(device));
  ff_host_ctx* c = new (std::nothrow) ff_host_ctx();
  if (c == nullptr) return FF_ERR_INVALID;
  c->device = device;
  c->chunk_bytes = chunk_bytes;
  int rc = FF_OK;
  auto guard = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == FF_OK) {
      set_cuda_error(e, what);
      rc = FF_ERR_CUDA;
    }
  };
  guard(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
  guard(cudaStreamCreateWithFlags(&c->compute_stream, cudaStreamNonBlocking), "cudaStreamCreate(compute)");
  for (int i = 0; i < 2; ++i) {
    guard(cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming), "cudaEventCreate");
    guard(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming), "cudaEventCreate");
  }
  guard(cudaMalloc(&c->scalars_dev, 4 * sizeof(int32_t)), "cudaMalloc(scalars)");
  guard(cudaMallocHost(&c->scalars_host, 4 * sizeof(int32_t)), "cudaMallocHost(scalars)");
  if (rc != FF_OK) {
    ctx_release(c);
    return rc;
  }
  *ctx_out = c;
  return FF_OK;
}

int ff_host_ctx_destroy(ff_host_ctx* ctx) { return ctx_release(ctx); }

int ff_host_upload(ff_host_ctx* c, const void* src_host, void* dst_dev, int64_t bytes) {
  if (c == nullptr || src_host == nullptr || dst_dev == nullptr || bytes < 0) return FF_ERR_INVALID;
  if (bytes == 0) return FF_OK;
  FF_CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t cs = c->copy_stream;
  cudaPointerAttributes attr{};
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess)
    pageable = !(attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
  else
    cudaGetLastError();
  const uint8_t* src = static_cast<const uint8_t*>(src_host);
  uint8_t* dst = static_cast<uint8_t*>(dst_dev);
  if (!pageable) {
    FF_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, cs));
    FF_CUDA_TRY(cudaStreamSynchronize(cs));
    return FF_OK;
  }
  const int64_t piece = c->chunk_bytes > 0 ? c->chunk_bytes : (64 << 20);
  int rc = ctx_reserve_bounce(c, piece);
  if (rc != FF_OK) return rc;
  bool used[2] = {false, false};
  int b = 0;
  for (int64_t off = 0; off < bytes; off += piece, b ^= 1) {
    const int64_t n = off + piece <= bytes ? piece : bytes - off;
    if (used[b]) FF_CUDA_TRY(cudaEventSynchronize(c->bounce_free[b]));
    c->pool->copy(c->bounce[b], src + off, (size_t)n);      // overlaps the previous piece's DMA
    FF_CUDA_TRY(cudaMemcpyAsync(dst + off, c->bounce[b], (size_t)n, cudaMemcpyHostToDevice, cs));
    FF_CUDA_TRY(cudaEventRecord(c->bounce_free[b], cs));
    used[b] = true;
  }
  FF_CUDA_TRY(cudaStreamSynchronize(cs));
  return FF_OK;
}

int ff_process_host(ff_host_ctx* c, const void* frames_host, const void* halo_host, int64_t n_frames,
                    int64_t first_frame, int height, int width, int bits, int32_t bg, int32_t empty_thr,
                    int64_t min_signal_count, int method, int use_frame_diff, int32_t diff_thr,
                    int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px, int32_t exit_margin_px,
                    const uint8_t* skip_host, int32_t* pos_out_host, int32_t* count_out_host,
                    int64_t* frames_done_out, int32_t* first_exit_out) {
  if (c == nullptr || frames_host == nullptr || pos_out_host == nullptr) return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width <= 0 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;
  FF_CUDA_TRY(cudaSetDevice(c->device));

  const int64_t fb = frame_bytes_of(px, bits);
  // Staging buffer = [halo slot][chunk frames]; the halo slot is padded so frames stay 16-B aligned.
  const int64_t halo_slot = (fb + 15) & ~(int64_t)15;
  int64_t chunk_frames = c->chunk_bytes / fb;
  if (chunk_frames < 1) chunk_frames = 1;
  if (chunk_frames > n_frames) chunk_frames = n_frames;
  const Tiling tl = choose_tiling(px);
  int rc = ctx_reserve(c, halo_slot + chunk_frames * fb, chunk_frames * tl.partials_per_frame, n_frames);
  if (rc != FF_OK) return rc;

  cudaStream_t cs = c->copy_stream, ks = c->compute_stream;
  int32_t* bg_dev = c->scalars_dev;
  int32_t* exit_dev = c->scalars_dev + 1;
  c->scalars_host[0] = bg;
  c->scalars_host[1] = FF_NO_EXIT;
  c->scalars_host[2] = c->scalars_host[3] = FF_NO_EXIT;
  FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_dev, c->scalars_host, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ks));
  FF_CUDA_TRY(cudaMemsetAsync(c->cnt_dev, 0, sizeof(int32_t) * (size_t)n_frames, ks));
  const uint8_t* skip_dev = nullptr;
  if (skip_host != nullptr) {
    FF_CUDA_TRY(cudaMemcpyAsync(c->skip_dev, skip_host, (size_t)n_frames, cudaMemcpyHostToDevice, ks));
    skip_dev = c->skip_dev;
  }

  const uint8_t* src = static_cast<const uint8_t*>(frames_host);
  const int64_t n_chunks = (n_frames + chunk_frames - 1) / chunk_frames;
  // Pinned (or registered) sources are DMA'd in place; anything else goes through the bounce buffers.
  cudaPointerAttributes attr{};
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, frames_host) == cudaSuccess)
    pageable = !(attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
  else
    cudaGetLastError();       // unregistered memory reports an error on old drivers: clear it
  if (pageable) {
    rc = ctx_reserve_bounce(c, halo_slot + chunk_frames * fb);
    if (rc != FF_OK) return rc;
  }
  bool bounce_used[2] = {false, false};
  bool used[2] = {false, false};
  int64_t frames_done = 0;
  int32_t seen_exit = FF_NO_EXIT;

  for (int64_t ci = 0; ci < n_chunks; ++ci) {
    const int b = (int)(ci & 1);
    const int64_t a = ci * chunk_frames;
    const int64_t e = (a + chunk_frames < n_frames) ? a + chunk_frames : n_frames;
    if (used[b]) {  // chunk ci-2 must be finished before its buffer is overwritten
      FF_CUDA_TRY(cudaEventSynchronize(c->done[b]));
      if (c->scalars_host[2 + b] < seen_exit) seen_exit = c->scalars_host[2 + b];
    }
    // Chunk ci-1 may already be finished too: peek without blocking.
    if (used[b ^ 1] && cudaEventQuery(c->done[b ^ 1]) == cudaSuccess && c->scalars_host[2 + (b ^ 1)] < seen_exit)
      seen_exit = c->scalars_host[2 + (b ^ 1)];
    if (seen_exit != FF_NO_EXIT) break;  // the reference loop breaks at the exit frame (:1494)

    uint8_t* halo_dst = c->stage[b] + (halo_slot - fb);
    uint8_t* frames_dst = c->stage[b] + halo_slot;
    const uint8_t* halo_dev = nullptr;
    // Halo = the latest non-skipped frame before the chunk (the reference's prior-frame carry,
    // scripts/process_videos.py:469,1462 with :1443-1445): usually frame a-1, contiguous with
    // the chunk in host memory, so one copy moves both.
    int64_t h = a - 1;
    if (skip_host != nullptr)
      while (h >= 0 && skip_host[h]) --h;
    if (pageable) {
      // bounce[b] is free once the DMA that last read it has finished (chunk ci-2)
      if (bounce_used[b]) FF_CUDA_TRY(cudaEventSynchronize(c->bounce_free[b]));
      uint8_t* bh = c->bounce[b] + (halo_slot - fb);
      uint8_t* bf = c->bounce[b] + halo_slot;
      size_t bytes;
      const uint8_t* from;
      if (h >= 0 && h == a - 1) {                 // halo contiguous with the chunk: one parallel copy
        c->pool->copy(bh, src + h * fb, (size_t)((e - a + 1) * fb));
        halo_dev = halo_dst;
        from = bh;
        bytes = (size_t)((e - a + 1) * fb);
      } else {
        const void* hsrc = h >= 0 ? static_cast<const void*>(src + h * fb) : halo_host;
        if (hsrc != nullptr) {
          std::memcpy(bh, hsrc, (size_t)fb);
          halo_dev = halo_dst;
        }
        c->pool->copy(bf, src + a * fb, (size_t)((e - a) * fb));
        from = hsrc != nullptr ? bh : bf;
        bytes = (size_t)((e - a) * fb) + (hsrc != nullptr ? (size_t)fb : 0);
      }
      FF_CUDA_TRY(cudaMemcpyAsync(c->stage[b] + (from - c->bounce[b]), from, bytes, cudaMemcpyHostToDevice, cs));
      FF_CUDA_TRY(cudaEventRecord(c->bounce_free[b], cs));
      bounce_used[b] = true;
    } else if (h >= 0 && h == a - 1) {
      FF_CUDA_TRY(cudaMemcpyAsync(halo_dst, src + h * fb, (size_t)((e - a + 1) * fb), cudaMemcpyHostToDevice, cs));
      halo_dev = halo_dst;
    } else {
      const void* hsrc = h >= 0 ? static_cast<const void*>(src + h * fb) : halo_host;
      if (hsrc != nullptr) {
        FF_CUDA_TRY(cudaMemcpyAsync(halo_dst, hsrc, (size_t)fb, cudaMemcpyHostToDevice, cs));
        halo_dev = halo_dst;
      }
      FF_CUDA_TRY(cudaMemcpyAsync(frames_dst, src + a * fb, (size_t)((e - a) * fb), cudaMemcpyHostToDevice, cs));
    }
    FF_CUDA_TRY(cudaEventRecord(c->copied[b], cs));
    FF_CUDA_TRY(cudaStreamWaitEvent(ks, c->copied[b], 0));

    const uint8_t* skip_chunk = skip_dev ? skip_dev + a : nullptr;
    rc = stream_frames_impl(frames_dst, halo_dev, e - a, height, width, bits, bg_dev, empty_thr, diff_thr, skip_chunk,
                            c->partial[b], nullptr, FF_DIFF_NONE, nullptr, ks);
    if (rc != FF_OK) return rc;
    rc = detect_impl(frames_dst, halo_dev, e - a, first_frame + a, height, width, bits, bg_dev, c->partial[b],
                     min_signal_count, method, use_frame_diff, diff_thr, threshold_floor, grad2_bound, min_run_px,
                     exit_margin_px, skip_chunk, c->pos_dev + a, c->cnt_dev + a, exit_dev, nullptr, ks);
    if (rc != FF_OK) return rc;
    FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 2 + b, exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
    FF_CUDA_TRY(cudaEventRecord(c->done[b], ks));
    // The next copy into this buffer waits for these kernels on the device side as well.
    FF_CUDA_TRY(cudaStreamWaitEvent(cs, c->done[b], 0));
    used[b] = true;
    frames_done = e;
  }

  // Frames never copied are, by construction, at or beyond the exit frame.
  if (frames_done < n_frames)
    FF_CUDA_TRY(cudaMemsetAsync(c->pos_dev + frames_done, 0xFF, sizeof(int32_t) * (size_t)(n_frames - frames_done), ks));
  rc = truncate_impl(c->pos_dev, n_frames, first_frame, exit_dev, ks);
  if (rc != FF_OK) return rc;
  FF_CUDA_TRY(cudaMemcpyAsync(pos_out_host, c->pos_dev, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, ks));
  if (count_out_host != nullptr)
    FF_CUDA_TRY(cudaMemcpyAsync(count_out_host, c->cnt_dev, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, ks));
  FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 1, exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
  FF_CUDA_TRY(cudaStreamSynchronize(ks));
  FF_CUDA_TRY(cudaStreamSynchronize(cs));
  if (frames_done_out) *frames_done_out = frames_done;
  if (first_exit_out) *first_exit_out = c->scalars_host[1];
  return FF_OK;
}

}  // extern "C"

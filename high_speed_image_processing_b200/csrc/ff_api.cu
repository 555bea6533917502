// C-ABI entry points of libflamefront.so (declared in include/flamefront.h) and the
// host-resident streaming driver (ff_process_host).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ff_internal.h"

namespace ff {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  std::snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int head_lines_impl(const void*, const void*, int64_t, int, int, int, const int32_t*, const int32_t*, int64_t, int32_t,
                    int, const double*, int, const uint8_t*, double*, uint8_t*, int32_t*, cudaStream_t);
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
    // Every job carries its own counters: a worker that wakes up late for job N still holds job N
    // (whose parts are all taken) and can neither copy a part of job N+1 nor count towards it.
    auto job = std::make_shared<Job>();
    job->dst = dst;
    job->src = src;
    job->bytes = bytes;
    job->part = 2u << 20;
    job->n_parts = (bytes + job->part - 1) / job->part;
    {
      std::lock_guard<std::mutex> g(m_);
      current_ = job;
      ++gen_;
    }
    cv_job_.notify_all();
    work(*job);
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return job->done >= job->n_parts; });
  }

 private:
  struct Job {
    uint8_t* dst = nullptr;
    const uint8_t* src = nullptr;
    size_t bytes = 0, part = 0, n_parts = 0;
    std::atomic<size_t> next{0};
    size_t done = 0;          // guarded by m_
  };
  void work(Job& j) {
    size_t mine = 0;
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.n_parts) break;
      const size_t a = i * j.part;
      const size_t n = a + j.part <= j.bytes ? j.part : j.bytes - a;
      stream_copy(j.dst + a, j.src + a, n);
      ++mine;
    }
    if (mine) {
      std::lock_guard<std::mutex> g(m_);
      j.done += mine;
      if (j.done >= j.n_parts) cv_done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_job_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        job = current_;
      }
      work(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_job_, cv_done_;
  std::shared_ptr<Job> current_;
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
  int32_t* scalars_dev = nullptr;   // int32[32]: [0..15] clip-scalar block ([0] bg), [16] first_exit
  int32_t* scalars_host = nullptr;  // pinned: [0] bg, [1] final first_exit, [2..3] per-buffer first_exit readback,
                                    // [4..5] per-buffer readback of the clip-global exit word (range-sharded runs)
  void* ws = nullptr;               // ff_process_range workspace for one chunk (kept zero-filled)
  int64_t ws_bytes = 0;
  int copy_threads = -1;            // -1: FF_HOST_COPY_THREADS or half the cores
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
  if (c->ws) cudaFree(c->ws);
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
    if (c->copy_threads > 0) n = c->copy_threads - 1;
    if (n < 0) n = 0;
    if (n > 15) n = 15;
    c->pool = new (std::nothrow) CopyPool(n);
    if (c->pool == nullptr) return FF_ERR_INVALID;
  }
  return FF_OK;
}

static int ctx_reserve(ff_host_ctx* c, int64_t stage_bytes, int64_t partial_elems, int64_t frames, int64_t ws_bytes) {
  if (ws_bytes > c->ws_bytes) {
    if (c->ws) FF_CUDA_TRY(cudaFree(c->ws));
    c->ws = nullptr;
    FF_CUDA_TRY(cudaMalloc(&c->ws, (size_t)ws_bytes));
    FF_CUDA_TRY(cudaMemset(c->ws, 0, (size_t)ws_bytes));
    c->ws_bytes = ws_bytes;
  }
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
      if (partial_elems > 0) FF_CUDA_TRY(cudaMalloc(&c->partial[i], sizeof(int32_t) * (size_t)partial_elems));
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
                  int morphology_size, const double* gauss_weights_host, int radius, const uint8_t* skip_dev,
                  double* lines_out_dev, uint8_t* flags_out_dev, int32_t* scratch_dev, void* stream) {
  return head_lines_impl(frames_dev, halo_dev, n_frames, height, width, bits, bg_dev, partial_dev, min_signal_count,
                         diff_thr, morphology_size, gauss_weights_host, radius, skip_dev, lines_out_dev, flags_out_dev,
                         scratch_dev, static_cast<cudaStream_t>(stream));
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
  guard(cudaMalloc(&c->scalars_dev, 32 * sizeof(int32_t)), "cudaMalloc(scalars)");
  guard(cudaMallocHost(&c->scalars_host, 8 * sizeof(int32_t)), "cudaMallocHost(scalars)");
  if (rc != FF_OK) {
    ctx_release(c);
    return rc;
  }
  *ctx_out = c;
  return FF_OK;
}

int ff_host_ctx_destroy(ff_host_ctx* ctx) { return ctx_release(ctx); }

int ff_host_ctx_set_copy_threads(ff_host_ctx* c, int n_threads) {
  if (c == nullptr || n_threads < 1) return FF_ERR_INVALID;
  if (c->pool != nullptr && n_threads != c->copy_threads) {      // takes effect for the next pageable source
    delete c->pool;
    c->pool = nullptr;
  }
  c->copy_threads = n_threads > 16 ? 16 : n_threads;
  return FF_OK;
}

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

int ff_process_host_range(ff_host_ctx* c, const ff_host_args* ha) {
  if (c == nullptr || ha == nullptr || ha->frames_host == nullptr) return FF_ERR_INVALID;
  const int64_t n_frames = ha->n_frames, first_frame = ha->first_frame;
  const int height = ha->height, width = ha->width, bits = ha->bits;
  const bool to_block = ha->pos_block_dev != nullptr;
  if (!to_block && ha->pos_out_host == nullptr) return FF_ERR_INVALID;
  if (to_block && (ha->count_block_dev == nullptr || ha->first_exit_block_dev == nullptr)) return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width <= 0 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;
  FF_CUDA_TRY(cudaSetDevice(c->device));
  const uint8_t* skip_host = ha->skip_host;

  const int64_t fb = frame_bytes_of(px, bits);
  // Staging buffer = [halo slot][chunk frames]; the halo slot is padded so frames stay 16-B aligned.
  const int64_t halo_slot = (fb + 15) & ~(int64_t)15;
  int64_t chunk_frames = c->chunk_bytes / fb;
  if (chunk_frames < 1) chunk_frames = 1;
  if (chunk_frames > n_frames) chunk_frames = n_frames;
  const Tiling tl = choose_tiling(px);
  const bool fused = range_is_fused(px, FF_DIFF_NONE, false, false);
  int rc = ctx_reserve(c, halo_slot + chunk_frames * fb, fused ? 0 : chunk_frames * tl.partials_per_frame,
                       n_frames, range_workspace_bytes(chunk_frames));
  if (rc != FF_OK) return rc;

  cudaStream_t cs = c->copy_stream, ks = c->compute_stream;
  int32_t* scal_dev = c->scalars_dev;
  int32_t* pos_dev = to_block ? ha->pos_block_dev : c->pos_dev;
  int32_t* cnt_dev = to_block ? ha->count_block_dev : c->cnt_dev;
  int32_t* exit_dev = to_block ? ha->first_exit_block_dev : c->scalars_dev + 16;
  RangeHooks hooks = hooks_from(ha->hooks);
  const int hook_flags = hooks.flags;
  const int32_t* global_exit_dev = (ha->hooks != nullptr && hooks.table != nullptr) ? ha->hooks->exit_word_dev : nullptr;
  c->scalars_host[0] = ha->bg;
  c->scalars_host[1] = FF_NO_EXIT;
  for (int k = 2; k < 6; ++k) c->scalars_host[k] = FF_NO_EXIT;
  FF_CUDA_TRY(cudaMemcpyAsync(scal_dev, c->scalars_host, sizeof(int32_t), cudaMemcpyHostToDevice, ks));
  // first-exit word = FF_NO_EXIT; in a range-sharded run: wait until the peers no longer read the block
  hooks.flags = hook_flags & FF_HOOK_WAIT;
  rc = prep_impl(nullptr, height, width, bits, nullptr, nullptr, 0, exit_dev, nullptr, hooks, ks);
  if (rc != FF_OK) return rc;
  hooks.flags = 0;                       // the chunks only propagate exit frames; the block is published at the end
  FF_CUDA_TRY(cudaMemsetAsync(cnt_dev, 0, sizeof(int32_t) * (size_t)n_frames, ks));
  const uint8_t* skip_dev = nullptr;
  if (skip_host != nullptr) {
    FF_CUDA_TRY(cudaMemcpyAsync(c->skip_dev, skip_host, (size_t)n_frames, cudaMemcpyHostToDevice, ks));
    skip_dev = c->skip_dev;
  }
  int32_t seen_exit = FF_NO_EXIT;        // smallest exit frame this rank knows of (own chunks, and the peers')
  if (global_exit_dev != nullptr) {      // another rank may already have seen the flame leave
    FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 4, global_exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
    FF_CUDA_TRY(cudaStreamSynchronize(ks));
    seen_exit = c->scalars_host[4];
  }

  const uint8_t* src = static_cast<const uint8_t*>(ha->frames_host);
  const int64_t n_chunks = (n_frames + chunk_frames - 1) / chunk_frames;
  // Pinned (or registered) sources are DMA'd in place; anything else goes through the bounce buffers.
  cudaPointerAttributes attr{};
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, ha->frames_host) == cudaSuccess)
    pageable = !(attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
  else
    cudaGetLastError();       // unregistered memory reports an error on old drivers: clear it
  if (pageable) {
    rc = ctx_reserve_bounce(c, halo_slot + chunk_frames * fb);
    if (rc != FF_OK) return rc;
  }
  bool bounce_used[2] = {false, false};
  bool used[2] = {false, false};
  int64_t frames_done = 0, bytes_up = 0;

  for (int64_t ci = 0; ci < n_chunks; ++ci) {
    const int b = (int)(ci & 1);
    const int64_t a = ci * chunk_frames;
    const int64_t e = (a + chunk_frames < n_frames) ? a + chunk_frames : n_frames;
    if (used[b]) {  // chunk ci-2 must be finished before its buffer is overwritten
      FF_CUDA_TRY(cudaEventSynchronize(c->done[b]));
      seen_exit = std::min(seen_exit, std::min(c->scalars_host[2 + b], c->scalars_host[4 + b]));
    }
    // Chunk ci-1 may already be finished too: peek without blocking.
    if (used[b ^ 1] && cudaEventQuery(c->done[b ^ 1]) == cudaSuccess)
      seen_exit = std::min(seen_exit, std::min(c->scalars_host[2 + (b ^ 1)], c->scalars_host[4 + (b ^ 1)]));
    // The reference loop breaks at the exit frame (:1494): nothing at or behind it is copied, whichever
    // rank found it.  (An exit in one of this range's own chunks always lies before the next chunk.)
    if ((int64_t)seen_exit <= first_frame + a) break;

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
        const void* hsrc = h >= 0 ? static_cast<const void*>(src + h * fb) : ha->halo_host;
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
      bytes_up += (int64_t)bytes;
    } else if (h >= 0 && h == a - 1) {
      FF_CUDA_TRY(cudaMemcpyAsync(halo_dst, src + h * fb, (size_t)((e - a + 1) * fb), cudaMemcpyHostToDevice, cs));
      halo_dev = halo_dst;
      bytes_up += (e - a + 1) * fb;
    } else {
      const void* hsrc = h >= 0 ? static_cast<const void*>(src + h * fb) : ha->halo_host;
      if (hsrc != nullptr) {
        FF_CUDA_TRY(cudaMemcpyAsync(halo_dst, hsrc, (size_t)fb, cudaMemcpyHostToDevice, cs));
        halo_dev = halo_dst;
        bytes_up += fb;
      }
      FF_CUDA_TRY(cudaMemcpyAsync(frames_dst, src + a * fb, (size_t)((e - a) * fb), cudaMemcpyHostToDevice, cs));
      bytes_up += (e - a) * fb;
    }
    FF_CUDA_TRY(cudaEventRecord(c->copied[b], cs));
    FF_CUDA_TRY(cudaStreamWaitEvent(ks, c->copied[b], 0));

    RangeJob j{};
    j.frames = frames_dst;
    j.halo = halo_dev;
    j.n_frames = e - a;
    j.first_frame = first_frame + a;
    j.height = height;
    j.width = width;
    j.bits = bits;
    j.method = ha->method;
    j.use_frame_diff = ha->use_frame_diff;
    j.min_run_px = ha->min_run_px;
    j.exit_margin_px = ha->exit_margin_px;
    j.diff_thr = ha->diff_thr;
    j.grad2_bound = ha->grad2_bound;
    j.empty_thr = ha->empty_thr;
    j.threshold_floor = ha->threshold_floor;
    j.min_signal_count = ha->min_signal_count;
    j.skip = skip_dev ? skip_dev + a : nullptr;
    j.scalars = scal_dev;
    j.pos_out = pos_dev + a;
    j.count_out = cnt_dev + a;
    j.first_exit = exit_dev;
    j.diff_dtype = FF_DIFF_NONE;
    j.partial = c->partial[b];
    j.ws = static_cast<RangeWorkspace*>(c->ws);
    j.hooks = hooks;
    rc = process_range_impl(j, ks);
    if (rc != FF_OK) return rc;
    FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 2 + b, exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
    if (global_exit_dev != nullptr)
      FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 4 + b, global_exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
    FF_CUDA_TRY(cudaEventRecord(c->done[b], ks));
    // (the next copy into stage[b] is issued only after the host has waited for done[b] above, so the
    // copy of chunk ci+1 into the other buffer overlaps these kernels)
    used[b] = true;
    frames_done = e;
  }

  // Frames never copied are, by construction, at or beyond the exit frame.
  if (frames_done < n_frames)
    FF_CUDA_TRY(cudaMemsetAsync(pos_dev + frames_done, 0xFF, sizeof(int32_t) * (size_t)(n_frames - frames_done), ks));
  if (to_block) {
    // range-sharded: truncation happens in the merge, against the clip-global exit frame
    if (hooks.table != nullptr && (hook_flags & FF_HOOK_PUBLISH)) {
      hooks.flags = FF_HOOK_PUBLISH;
      rc = publish_impl(hooks, ks);
      if (rc != FF_OK) return rc;
    }
  } else {
    rc = truncate_impl(pos_dev, n_frames, first_frame, exit_dev, ks);
    if (rc != FF_OK) return rc;
  }
  if (ha->pos_out_host != nullptr)
    FF_CUDA_TRY(cudaMemcpyAsync(ha->pos_out_host, pos_dev, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, ks));
  if (ha->count_out_host != nullptr)
    FF_CUDA_TRY(cudaMemcpyAsync(ha->count_out_host, cnt_dev, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, ks));
  FF_CUDA_TRY(cudaMemcpyAsync(c->scalars_host + 1, exit_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ks));
  FF_CUDA_TRY(cudaStreamSynchronize(ks));
  FF_CUDA_TRY(cudaStreamSynchronize(cs));
  if (ha->frames_done_out) *ha->frames_done_out = frames_done;
  if (ha->first_exit_out) *ha->first_exit_out = c->scalars_host[1];
  if (ha->bytes_uploaded_out) *ha->bytes_uploaded_out = bytes_up;
  return FF_OK;
}

int ff_process_host(ff_host_ctx* c, const void* frames_host, const void* halo_host, int64_t n_frames,
                    int64_t first_frame, int height, int width, int bits, int32_t bg, int32_t empty_thr,
                    int64_t min_signal_count, int method, int use_frame_diff, int32_t diff_thr,
                    int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px, int32_t exit_margin_px,
                    const uint8_t* skip_host, int32_t* pos_out_host, int32_t* count_out_host,
                    int64_t* frames_done_out, int32_t* first_exit_out) {
  ff_host_args a{};
  a.frames_host = frames_host;
  a.halo_host = halo_host;
  a.n_frames = n_frames;
  a.first_frame = first_frame;
  a.height = height;
  a.width = width;
  a.bits = bits;
  a.bg = bg;
  a.empty_thr = empty_thr;
  a.min_signal_count = min_signal_count;
  a.method = method;
  a.use_frame_diff = use_frame_diff;
  a.diff_thr = diff_thr;
  a.threshold_floor = threshold_floor;
  a.grad2_bound = grad2_bound;
  a.min_run_px = min_run_px;
  a.exit_margin_px = exit_margin_px;
  a.skip_host = skip_host;
  a.pos_out_host = pos_out_host;
  a.count_out_host = count_out_host;
  a.frames_done_out = frames_done_out;
  a.first_exit_out = first_exit_out;
  return ff_process_host_range(c, &a);
}

}  // extern "C"

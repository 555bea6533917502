// Internal entry points shared by the translation units of libflamefront (not part of the C-ABI).
#pragma once
#include "ff_detect_core.cuh"

namespace ff {

// ff_stream.cu
int stream_frames_impl(const void* frames, const void* halo, int64_t n_frames, int height, int width, int bits,
                       const int32_t* bg_dev, int32_t empty_thr, int32_t diff_thr, const uint8_t* skip,
                       int32_t* partial, void* diff_out, int diff_dtype, uint16_t* decoded_out, cudaStream_t st,
                       bool pdl = false);
int unpack_impl(const void* packed, void* out, int64_t n_frames, int height, int width, int bits, cudaStream_t st);
bool range_is_fused(int64_t px_per_frame, int diff_dtype, bool decoded, bool profiles);
int range_fused_impl(const DetectParams& d, int bits, int32_t empty_thr, bool pdl, cudaStream_t st);

// ff_detect.cu
int make_detect_params(DetectParams* out, const void* frames, const void* halo, int64_t n_frames,
                       int64_t first_frame, int height, int width, int bits, const int32_t* bg_dev,
                       const int32_t* partial, int64_t min_signal_count, int method, int use_frame_diff,
                       int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px,
                       int32_t exit_margin_px, const uint8_t* skip, int32_t* pos_out, int32_t* count_out,
                       int32_t* first_exit, int32_t* profile_out);
int launch_detect(const DetectParams& p, int bits, bool pdl, cudaStream_t st);
int detect_impl(const void* frames, const void* halo, int64_t n_frames, int64_t first_frame, int height, int width,
                int bits, const int32_t* bg_dev, const int32_t* partial, int64_t min_signal_count, int method,
                int use_frame_diff, int32_t diff_thr, int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px,
                int32_t exit_margin_px, const uint8_t* skip, int32_t* pos_out, int32_t* count_out, int32_t* first_exit,
                int32_t* profile_out, cudaStream_t st);
int prep_impl(const void* frame0, int height, int width, int bits, int32_t* scalars, uint16_t* centerline,
              int want_stats, int32_t* first_exit, RangeWorkspace* ws, const RangeHooks& hooks, cudaStream_t st);
int background_impl(const void* frame0, int height, int width, int bits, int32_t* bg_max, uint16_t* centerline,
                    cudaStream_t st);
int truncate_impl(int32_t* pos, int64_t n_frames, int64_t first_frame, const int32_t* first_exit, cudaStream_t st);

// ff_exchange.cu
int publish_impl(const RangeHooks& hooks, cudaStream_t st);

// ff_range.cu: one contiguous frame range, frames on the device (the body of ff_process_range)
struct RangeJob {
  const void* frames;
  const void* halo;
  const void* frame0;        // non-null: background scalar (and the threshold method's bound) come from prep_kernel
  int64_t n_frames, first_frame;
  int height, width, bits;
  int method, use_frame_diff, min_run_px, exit_margin_px;
  int32_t diff_thr, grad2_bound, empty_thr, threshold_floor;
  int64_t min_signal_count;
  const uint8_t* skip;
  int32_t* scalars;          // int32[16] clip-scalar block ([0] background scalar)
  uint16_t* centerline;      // uint16[W] or null
  int32_t* pos_out;
  int32_t* count_out;
  int32_t* first_exit;
  int init_first_exit;       // set *first_exit = FF_NO_EXIT before the frames are looked at
  int truncate;
  void* diff_out;
  int diff_dtype;
  uint16_t* decoded_out;
  int32_t* profile_out;
  int32_t* partial;          // scratch of the three-kernel form (ff_partial_len), unused by the fused kernel
  RangeWorkspace* ws;
  RangeHooks hooks;
};
int process_range_impl(const RangeJob& j, cudaStream_t st);

}  // namespace ff

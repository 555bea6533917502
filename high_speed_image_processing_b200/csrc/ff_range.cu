// ff_process_range: stages 2-4 for one contiguous, device-resident frame range in one call
// (the frame loop of scripts/process_videos.py:1441-1516; per-clip scalars of :1356-1370).
#include <cstdlib>

#include "ff_internal.h"

namespace ff {

int process_range_impl(const RangeJob& j, cudaStream_t st) {
  if (j.frames == nullptr || j.scalars == nullptr || j.pos_out == nullptr || j.first_exit == nullptr || j.ws == nullptr)
    return FF_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(j.ws) & 15u) != 0 || (reinterpret_cast<uintptr_t>(j.scalars) & 15u) != 0)
    return FF_ERR_ALIGNMENT;
  const bool from_frame0 = j.frame0 != nullptr;
  const int want_stats = from_frame0 && j.method == FF_METHOD_THRESHOLD;
  DetectParams d{};
  int rc = make_detect_params(&d, j.frames, j.halo, j.n_frames, j.first_frame, j.height, j.width, j.bits, j.scalars,
                              j.partial, j.min_signal_count, j.method, j.use_frame_diff, j.diff_thr,
                              j.threshold_floor, j.grad2_bound, j.min_run_px, j.exit_margin_px, j.skip, j.pos_out,
                              j.count_out, j.first_exit, j.profile_out);
  if (rc != FF_OK) return rc;
  d.threshold_dev = want_stats ? j.scalars + 1 : nullptr;
  d.ws = j.ws;
  d.truncate = j.truncate;
  d.hooks = j.hooks;

  const bool need_prep = from_frame0 || j.init_first_exit || (j.hooks.table != nullptr && (j.hooks.flags & FF_HOOK_WAIT));
  if (need_prep) {
    rc = prep_impl(j.frame0, j.height, j.width, j.bits, j.scalars, j.centerline, want_stats,
                   j.init_first_exit ? j.first_exit : nullptr, j.ws, j.hooks, st);
    if (rc != FF_OK) return rc;
  }
  // Programmatic dependent launch behind prep_kernel: the range kernel's CTAs come up and its producer warps
  // fetch tiles while prep still runs (C2 step 0.577 -> 0.572 ms).  Only for the range kernel, whose CTAs fill
  // every SM to its resource limit: a one-CTA-per-SM stream kernel launched early is placed around prep's
  // CTAs, ends up unevenly spread, and the doubled-up SMs decide its time (C4 uint16: 3.1 -> 4.9 ms).
  // FF_PDL=0 / 1 forces it off / on for both.
  static const int pdl_env = getenv("FF_PDL") != nullptr ? atoi(getenv("FF_PDL")) : -1;
  const int64_t px = (int64_t)j.height * j.width;
  if (range_is_fused(px, j.diff_dtype, j.decoded_out != nullptr, j.profile_out != nullptr))
    return range_fused_impl(d, j.bits, j.empty_thr, need_prep && pdl_env != 0, st);
  if (j.partial == nullptr) return FF_ERR_INVALID;
  rc = stream_frames_impl(j.frames, j.halo, j.n_frames, j.height, j.width, j.bits, j.scalars, j.empty_thr, j.diff_thr,
                          j.skip, j.partial, j.diff_out, j.diff_dtype, j.decoded_out, st, need_prep && pdl_env == 1);
  if (rc != FF_OK) return rc;
  return launch_detect(d, j.bits, false, st);
}

}  // namespace ff

using namespace ff;

extern "C" {

int ff_process_range_plan(int64_t n_frames, int height, int width, int bits, int diff_dtype, int want_decoded,
                          int want_profiles, int64_t* workspace_bytes, int64_t* partial_elems, int* fused) {
  if (n_frames < 0 || height <= 0 || width <= 0) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  const int64_t px = (int64_t)height * width;
  const bool f = range_is_fused(px, diff_dtype, want_decoded != 0, want_profiles != 0);
  if (workspace_bytes) *workspace_bytes = range_workspace_bytes(n_frames);
  if (partial_elems) *partial_elems = f ? 0 : n_frames * choose_tiling(px).partials_per_frame;
  if (fused) *fused = f ? 1 : 0;
  return FF_OK;
}

int ff_process_range(const ff_range_args* a, void* stream) {
  if (a == nullptr) return FF_ERR_INVALID;
  if (a->diff_dtype < FF_DIFF_NONE || a->diff_dtype > FF_DIFF_F64) return FF_ERR_INVALID;
  RangeJob j{};
  j.frames = a->frames_dev;
  j.halo = a->halo_dev;
  j.frame0 = a->frame0_dev;
  j.n_frames = a->n_frames;
  j.first_frame = a->first_frame;
  j.height = a->height;
  j.width = a->width;
  j.bits = a->bits;
  j.method = a->method;
  j.use_frame_diff = a->use_frame_diff;
  j.min_run_px = a->min_run_px;
  j.exit_margin_px = a->exit_margin_px;
  j.diff_thr = a->diff_thr;
  j.grad2_bound = a->grad2_bound;
  j.empty_thr = a->empty_thr;
  j.threshold_floor = a->threshold_floor;
  j.min_signal_count = a->min_signal_count;
  j.skip = a->skip_dev;
  j.scalars = a->scalars_dev;
  j.centerline = a->centerline_dev;
  j.pos_out = a->pos_out_dev;
  j.count_out = a->count_out_dev;
  j.first_exit = a->first_exit_dev;
  j.init_first_exit = a->init_first_exit;
  j.truncate = a->truncate;
  j.diff_out = a->diff_out_dev;
  j.diff_dtype = a->diff_dtype;
  j.decoded_out = a->decoded_out_dev;
  j.profile_out = a->profile_out_dev;
  j.partial = a->partial_dev;
  j.ws = static_cast<RangeWorkspace*>(a->workspace_dev);
  j.hooks = hooks_from(a->hooks);
  return process_range_impl(j, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

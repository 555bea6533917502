/*
 * flamefront.h - C-ABI of libflamefront.so, the B200 (sm_100a) implementation of the
 * per-frame flame-front path of Nadexterbrown/High-Speed-Image-Processing.
 *
 * The reference has no FFI: its seams are plain Python calls.
 * Each entry point below names the reference interface it replaces (paths are relative
 * to the reference repository root).  The ctypes binding a maintainer would add is in
 * INTEGRATION.md; the binding this repo ships is
 * high_speed_image_processing_b200/_cabi.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary;
 *   - every function returns 0 (FF_OK) or a negative FF_ERR_* code, never throws;
 *   - "dev" pointers are device memory owned by the caller (e.g. a torch tensor's
 *     data_ptr()); "host" pointers are host memory owned by the caller;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all
 *     device-resident calls are asynchronous on that stream;
 *   - frames are stored exactly as in a Photron .mraw file: frame-major, row-major,
 *     8-bit, little-endian 16-bit, or packed 12-bit (3 bytes = 2 pixels:
 *     p0 = b0<<4 | b1>>4, p1 = (b1&15)<<8 | b2), `bits` selects which (the CIH "Color Bit");
 *   - positions are int32: >= 0 detected column, FF_POS_NONE (-1) no detection / empty /
 *     skipped frame, FF_POS_DROPPED (-2) removed by flame-exit truncation.
 */
#ifndef FLAMEFRONT_H_
#define FLAMEFRONT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FF_ABI_VERSION 6

/* status codes */
#define FF_OK                 0
#define FF_ERR_INVALID       -1   /* bad argument (NULL pointer, non-positive size, ...) */
#define FF_ERR_UNSUPPORTED   -2   /* bit depth / shape / dtype this build cannot handle   */
#define FF_ERR_CUDA          -3   /* a CUDA runtime call failed; see ff_last_cuda_error() */
#define FF_ERR_NO_DEVICE     -4   /* no CUDA device / driver                              */
#define FF_ERR_ALIGNMENT     -5   /* pointer not aligned as the entry point requires      */

/* detection_method switch (README.md:55,62,132-141; VideoSourceConfig.detection_method) */
#define FF_METHOD_THRESHOLD     0
#define FF_METHOD_GRADIENT      1
#define FF_METHOD_HALF_MAXIMUM  2

/* dtype of the retained full-frame difference image */
#define FF_DIFF_NONE  0
#define FF_DIFF_U16   1   /* lossless when diff_thr >= 0: values are integers in [0,65535] */
#define FF_DIFF_F32   2
#define FF_DIFF_F64   3   /* the reference's own dtype (float64)                        */

/* element type of a decoded frame handed to the frame-level operators */
#define FF_PX_U8   0
#define FF_PX_U16  1
#define FF_PX_F64  3

#define FF_POS_NONE     (-1)
#define FF_POS_DROPPED  (-2)
#define FF_NO_EXIT      2147483647   /* value of *first_exit when the flame never exits */

/* ---- library ------------------------------------------------------------------- */
int         ff_abi_version(void);
const char* ff_strerror(int status);
const char* ff_last_cuda_error(void);          /* thread-local text of the last CUDA failure */
int         ff_device_count(int* count);
int         ff_device_sm_count(int device, int* sm_count);

/* ---- scratch sizing ---------------------------------------------------------------
 * ff_stream_frames writes partial counts of above-noise pixels (one per frame, tile and warp
 * of the streaming kernel) into a caller-owned int32 scratch array; ff_detect sums them.
 * ff_partial_len returns its length in int32 elements and the number of partial counts per
 * frame for a given frame shape.                                                        */
int ff_partial_len(int64_t n_frames, int height, int width, int bits,
                   int64_t* n_elems, int* partials_per_frame);

/* ---- stage 1: decode ------------------------------------------------------------------
 * Replaces pyMRAW.load_video's 12-bit unpack (call site src/photron/video.py:332) and the
 * per-frame copy in PhotonVideo.__getitem__ (src/photron/video.py:575-582).
 * packed_dev: n_frames*H*W*bits/8 bytes; out_dev: uint16[n_frames*H*W] (uint8 for bits=8). */
int ff_unpack(const void* packed_dev, void* out_dev, int64_t n_frames, int height, int width,
              int bits, void* stream);

/* ---- stage 2a: background-frame reduction ------------------------------------------------
 * Replaces `background_scalar = float(np.max(video[0]))` (scripts/process_videos.py:1357-1358)
 * and the centre-row extraction at :1361-1362.  The float64 mean/std/max and the flame
 * threshold (:1363-1370) stay on the host, computed from centerline_dev's W values.
 * bg_max_dev: int32[1]; centerline_dev: uint16[W] (row H/2 of the frame), nullable.       */
int ff_background(const void* frame0_dev, int height, int width, int bits,
                  int32_t* bg_max_dev, uint16_t* centerline_dev, void* stream);

/* ---- stage 2b: fused front end ----------------------------------------------------------
 * One HBM read per frame.  Replaces, per frame, subtract_scalar_background
 * (scripts/process_videos.py:670-674, called at :1455 and :380), the pixel count inside
 * is_empty_frame (:759, called at :1458-1459) and, when diff_dtype != FF_DIFF_NONE, the
 * full-frame difference `d = sub_t - sub_{t-1}; d[d < thr] = 0` (:397-399, :677-701) with the
 * prior-frame carry of :469/:1462.
 *   frames_dev   n_frames frames, base 16-byte aligned
 *   halo_dev     the frame preceding frames_dev[0] (same encoding) or NULL
 *   bg_dev       int32[1] background scalar (output of ff_background)
 *   empty_thr    pixel is "signal" iff max(x-bg,0) > empty_thr; pass < 0 to derive
 *                floor(max(10, bg/2)) from *bg_dev on the device (:1458)
 *   diff_thr     ceil(frame_diff_threshold); differences < diff_thr become 0
 *   skip_dev     uint8[n_frames], nonzero = frame listed in skip_frames (:1443-1445), nullable
 *   partial_dev  int32[ff_partial_len] scratch, fully overwritten
 *   diff_out_dev [n_frames,H,W] of diff_dtype, nullable iff diff_dtype == FF_DIFF_NONE;
 *                frames without a prior (first frame, skipped frames) are written as zeros
 *   decoded_out_dev uint16[n_frames,H,W] decoded pixels (12-bit input only), nullable      */
int ff_stream_frames(const void* frames_dev, const void* halo_dev, int64_t n_frames,
                     int height, int width, int bits,
                     const int32_t* bg_dev, int32_t empty_thr, int32_t diff_thr,
                     const uint8_t* skip_dev, int32_t* partial_dev,
                     void* diff_out_dev, int diff_dtype, uint16_t* decoded_out_dev,
                     void* stream);

/* ---- stage 3 + 4a: warp-per-profile detection and first-exit min ---------------------------
 * Replaces the centre-row profile extraction (scripts/process_videos.py:375-376,417-418), the
 * empty-frame decision (:761-763), the detection_method switch (README.md:132-141; gradient =
 * HEAD Method A, :413,427-430) and the exit test (:1488-1494).
 *   first_frame       clip-global index of frames_dev[0] (exit frames are reported globally)
 *   min_signal_count  a frame is empty iff its above-noise count < min_signal_count
 *   use_frame_diff    profile = centre row of the difference image (1) or of the
 *                     background-subtracted image (0)   (VideoSourceConfig.use_frame_diff)
 *   threshold_floor   threshold method: pixel is "high" iff p > threshold_floor
 *                     ( = floor(centerline_flame_threshold), :1367-1370 )
 *   grad2_bound       gradient method: valid iff 2*g_min < grad2_bound
 *                     ( = ceil(-2*min_gradient_strength), :174,428 )
 *   min_run_px        threshold method: minimum run length (>= 1)
 *   exit_margin_px    exit iff pos >= W - exit_margin_px (README.md:146; HEAD :193 uses 15)
 *   pos_out_dev       int32[n_frames]
 *   count_out_dev     int32[n_frames] above-noise pixel count per frame, nullable
 *   first_exit_dev    int32[1]; atomicMin'ed with the global index of every exit frame; the
 *                     caller initialises it to FF_NO_EXIT (or a previous chunk's value)
 *   profile_out_dev   int32[n_frames,W], nullable                                            */
int ff_detect(const void* frames_dev, const void* halo_dev, int64_t n_frames, int64_t first_frame,
              int height, int width, int bits,
              const int32_t* bg_dev, const int32_t* partial_dev, int64_t min_signal_count,
              int method, int use_frame_diff, int32_t diff_thr,
              int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px,
              int32_t exit_margin_px, const uint8_t* skip_dev,
              int32_t* pos_out_dev, int32_t* count_out_dev, int32_t* first_exit_dev,
              int32_t* profile_out_dev, void* stream);

/* ---- stage 4b: truncation ----------------------------------------------------------------
 * README.md:145-149: everything from the first exit frame on is dropped.  Frames whose
 * global index (first_frame + i) >= *first_exit_dev get FF_POS_DROPPED.  In a multi-GPU
 * run *first_exit_dev holds the all-reduced (min) value.                                   */
int ff_truncate(int32_t* pos_dev, int64_t n_frames, int64_t first_frame,
                const int32_t* first_exit_dev, void* stream);

/* ---- stages 2-4 in one call: a contiguous frame range, frames on the device -------------------------
 * The call SURVEY.md section 8b proposes for seam B3: replaces the frame loop of process_video_source
 * (scripts/process_videos.py:1441-1516) for one contiguous frame range, including the per-clip scalars of
 * :1356-1370 when frame0_dev is given.  Nothing here waits for the host.
 *
 *   prep kernel     (frame0_dev != NULL) background scalar = max of frame 0, its centre row, and - for the
 *                   threshold method - the float64 mean / std / max of that row and max(mean + 5 std, 2 max)
 *                   in NumPy's own order of operations (pairwise add.reduce), i.e. the bit-identical
 *                   threshold; first-exit word = FF_NO_EXIT; in a range-sharded run the wait for the peers
 *   range kernel    counts-only ranges (no image output, H*W % 32 == 0, H*W >= 128 Ki pixels): ONE kernel
 *                   streams the frames once (TMA ring), counts above-noise pixels, decides is_empty_frame,
 *                   resolves the detection_method on the non-empty frames (a detector warp per CTA), takes
 *                   the exit-frame min, truncates and publishes to the peers.  The kernel starts while
 *                   the prep kernel still runs (programmatic dependent launch).
 *   otherwise       ff_stream_frames -> ff_detect, with the same tail in ff_detect's last CTA.
 * ff_process_range_plan tells the caller what to allocate: workspace_bytes (zero-filled ONCE by its owner;
 * every call leaves it zero-filled again, so it can be reused without clearing), partial_elems (int32
 * scratch of the three-kernel form, 0 when the range runs fused) and whether it runs fused.
 *
 *   scalars_dev     int32[16] clip-scalar block.  [0] background scalar, [1] floor(flame threshold);
 *                   float64 at byte 16: centre-row mean, std, max, flame threshold (threshold method only).
 *                   Written by the prep kernel when frame0_dev != NULL, else [0] is read (ff_background's
 *                   output) and threshold_floor is taken from the argument.
 *   centerline_dev  uint16[W] centre row of frame 0 for the host's own float64 statistics, nullable
 *   empty_thr       as ff_stream_frames (< 0: derived from the background scalar on the device)
 *   init_first_exit set *first_exit_dev = FF_NO_EXIT first (else the caller did, or carries a value in)
 *   truncate        drop positions at/after *first_exit_dev (single-GPU; range-sharded runs truncate in
 *                   the merge against the global exit frame)
 *   hooks           range-sharded runs: from ff_exchange_begin (NULL or zeroed otherwise)                */
#define FF_HOOK_WAIT     1   /* before the block is written: wait until the peers no longer read its previous use */
#define FF_HOOK_PUBLISH  2   /* when the block is complete: tell the peers (flag = epoch)                          */
typedef struct ff_range_hooks {
  void*    table_dev;     /* device table of the peers' exchange buffers (owned by the ff_exchange)       */
  int32_t  epoch, world, rank, flags;
  int64_t  spin_limit;    /* GPU clock ticks a device-side wait may take before it raises the status word */
  int32_t* exit_word_dev; /* this rank's copy of the clip-global exit-frame min (host-streamed path polls it) */
} ff_range_hooks;

typedef struct ff_range_args {
  const void* frames_dev;
  const void* halo_dev;
  const void* frame0_dev;
  int64_t n_frames, first_frame;
  int32_t height, width, bits;
  int32_t method, use_frame_diff, min_run_px, exit_margin_px;
  int32_t diff_thr, grad2_bound, empty_thr, threshold_floor;
  int64_t min_signal_count;
  const uint8_t* skip_dev;
  int32_t*  scalars_dev;
  uint16_t* centerline_dev;
  int32_t*  pos_out_dev;
  int32_t*  count_out_dev;
  int32_t*  first_exit_dev;
  int32_t   init_first_exit, truncate;
  void*     diff_out_dev;
  int32_t   diff_dtype, reserved;
  uint16_t* decoded_out_dev;
  int32_t*  profile_out_dev;
  int32_t*  partial_dev;
  void*     workspace_dev;
  const ff_range_hooks* hooks;
} ff_range_args;

int ff_process_range_plan(int64_t n_frames, int height, int width, int bits, int diff_dtype, int want_decoded,
                          int want_profiles, int64_t* workspace_bytes, int64_t* partial_elems, int* fused);
int ff_process_range(const ff_range_args* args, void* stream);

/* ---- stage 4 across GPUs: the one exchange step of a range-sharded clip -----------------------------
 * Replaces the pickled `comm.gather` + sort of per-rank result lists (scripts/process_videos.py:
 * 1533-1541) and turns the per-rank `break` (:1494) into the global truncation of README.md:145-149.
 * Every rank owns contiguous_range(total, rank, world) (the contiguous variant of
 * MPIVideoProcessor.distribute_indices, src/photron/parallel.py:101-113) and lets ff_detect write
 * into a RANGE BLOCK of int32:  { first_exit, 0, 0, 0 | pos[cap] | counts[cap] }  (ff_range_block_len).
 *
 * Transport 1 - gathered: the caller all-gathers the blocks (one all_gather_into_tensor) and
 * ff_merge_ranges finishes: global exit = min of the headers, truncation, de-padding into
 * pos_out_dev[total] / count_out_dev[total] (nullable) / first_exit_out_dev[1].
 *
 * Transport 2 - peer memory (one box, NVLink): ff_exchange_* keeps the blocks where they were written;
 * peers map every rank's exchange buffer through CUDA IPC.  No collective and no barrier on the data path:
 *   - the LAST CTA of ff_process_range's kernel publishes the finished block (epoch flag in every peer's
 *     buffer, st.release.sys) - the compute kernel publishes, not the merge;
 *   - ff_exchange_finish launches the merge kernel (wait for the flags, pull the blocks over NVLink, global
 *     exit min, truncation, de-padding).  It depends on nothing but the flags, so the caller may launch it
 *     on a SIDE stream and go on with the next clip on the main stream;
 *   - the merge acknowledges to the peers that their blocks are no longer read; blocks are double-buffered by
 *     epoch parity and the prep kernel of epoch e waits for the acks of e-2 before anything is overwritten;
 *   - every exit frame a detecting warp sees is min-reduced into an exit word in EVERY rank's buffer
 *     (red.min.sys); ff_process_host_range polls its rank's copy between chunks and stops uploading frames
 *     that lie behind it - the reference's `break` (scripts/process_videos.py:1494) across ranks.
 *   ff_exchange_create      allocates the local blocks + flag / ack / exit words (cap_frames per block)
 *   ff_exchange_get_handle  writes ff_exchange_handle_bytes() bytes to exchange with the peers
 *   ff_exchange_open_peers  handles = world * handle_bytes, rank-major (own slot ignored)
 *   ff_exchange_begin       next epoch: this epoch's pos/count/first_exit device pointers and the hooks to hand
 *                           to ff_process_range / ff_process_host_range (which wait, fill, publish)
 *   ff_exchange_acquire /   for a block filled by hand instead: wait for the peers + reset the header before
 *   ff_exchange_publish     writing it / publish it afterwards (both asynchronous on `stream`)
 *   ff_exchange_finish      the merge kernel; every rank calls it once per ff_exchange_begin
 *   ff_exchange_status      synchronises `stream`; *status_out != 0 means a wait timed out after
 *                           FF_EXCHANGE_TIMEOUT_S seconds (environment, default 20): 1+r = rank r never
 *                           published (the merge then wrote FF_POS_NONE / FF_NO_EXIT), 0x100+r = rank r never
 *                           acknowledged.  Reading clears it.                                              */
int ff_range_block_len(int64_t cap_frames, int64_t* n_elems);
int ff_merge_ranges(const int32_t* gathered_dev, int world, int64_t block_cap_frames, int64_t total_frames,
                    int32_t* pos_out_dev, int32_t* count_out_dev, int32_t* first_exit_out_dev, void* stream);
typedef struct ff_exchange ff_exchange;
int ff_exchange_create(int device, int rank, int world, int64_t cap_frames, ff_exchange** out);
int ff_exchange_handle_bytes(void);
int ff_exchange_get_handle(ff_exchange* x, void* handle_out);
int ff_exchange_open_peers(ff_exchange* x, const void* handles);
int ff_exchange_begin(ff_exchange* x, int32_t** pos_dev, int32_t** count_dev, int32_t** first_exit_dev,
                      ff_range_hooks* hooks_out);
int ff_exchange_acquire(ff_exchange* x, void* stream);
int ff_exchange_publish(ff_exchange* x, void* stream);
int ff_exchange_finish(ff_exchange* x, int64_t total_frames, int32_t* pos_out_dev, int32_t* count_out_dev,
                       int32_t* first_exit_out_dev, void* stream);
int ff_exchange_status(ff_exchange* x, int32_t* status_out, void* stream);
int ff_exchange_destroy(ff_exchange* x);

/* ---- HEAD-parity detector (the code the reference executes at HEAD) ------------------------------
 * ff_head_lines replaces, per non-empty frame, the image pipeline of FlameDetector.detect
 * (scripts/process_videos.py:397-418): thresholded frame difference -> grey_opening k x k
 * (morphology_size = FlameDetectorConfig.morphology_kernel_size, :170; odd, 1..7; 3 takes the fast kernel) ->
 * gaussian_filter(sigma) -> sobel(axis=1) and np.gradient(axis=1), evaluated only on the band of
 * rows that reaches the centre row, in float64 with SciPy's operation order (bit-identical).
 *   gauss_weights_host  2*radius+1 float64 taps (scipy _gaussian_kernel1d), HOST pointer
 *   lines_out_dev       float64[n_frames,2,W]: [.,0,:] Sobel centre row, [.,1,:] gradient centre row
 *   flags_out_dev       uint8[n_frames]: 0 not processed (skipped/empty), 1 lines valid,
 *                       2 processed but no prior frame (detect() ran without a difference image)
 *   scratch_dev         int32[n_frames + 4] scratch (list of the frames that reached the detector)
 * ff_head_track replaces the sequential part (:317-348 search bounds, :420-465 candidate
 * selection) and the exit stop (:1488-1494).
 *   last_frame_in/last_pos_in  tracker state carried in (-1/-1 = no detection yet)
 *   out_dev   int32[n_frames,5]: final, pos_min_gradient, pos_rightmost_sobel, search_start,
 *             search_end; all -1 for frames that were not processed or lie after the exit frame
 *   stop_dev  int32[3]: exit frame (global index) or FF_NO_EXIT, last detection frame, last position
 *   scratch_dev  int32[ff_head_track_scratch_len(n_frames)] or NULL.  The walk over the frames is
 *             sequential only through (last frame, last position); it runs as speculative walks of
 *             16-frame segments.  With scratch the segments' start states are chained and checked in
 *             parallel; without it one warp validates the segments one after the other.  Same results. */
int ff_head_lines(const void* frames_dev, const void* halo_dev, int64_t n_frames, int height, int width,
                  int bits, const int32_t* bg_dev, const int32_t* partial_dev, int64_t min_signal_count,
                  int32_t diff_thr, int morphology_size, const double* gauss_weights_host, int radius,
                  const uint8_t* skip_dev, double* lines_out_dev, uint8_t* flags_out_dev, int32_t* scratch_dev,
                  void* stream);
int ff_head_track_scratch_len(int64_t n_frames, int64_t* n_elems);
int ff_head_track(const double* lines_dev, const uint8_t* flags_dev, int64_t n_frames, int64_t first_frame,
                  int width, int32_t edge_margin_px, int32_t max_displacement_px, int32_t search_window_px,
                  double min_gradient_strength, double sobel_threshold_fraction, int32_t exit_margin_px,
                  int32_t last_frame_in, int32_t last_pos_in, int32_t* out_dev, int32_t* stop_dev,
                  int32_t* scratch_dev, void* stream);

/* ---- frame-level operators (the per-frame seam of scripts/process_videos.py) ----------------------
 * What a caller gets who uses the reference's frame functions instead of its driver loop.  Inputs
 * are DECODED frames ([H,W] arrays as PhotonVideo.__getitem__ returns them, src/photron/video.py:559-584)
 * of px_type FF_PX_U8 / FF_PX_U16 / FF_PX_F64, n_px = H*W elements; outputs are float64 like the
 * reference's (one IEEE operation per NumPy operation, so values are bit-identical).
 *   ff_frame_subtract_background  subtract_scalar_background (:670-674): x.astype(f64) - bg; x[x<0] = 0
 *   ff_frame_difference           subtract_prior_frame (:677-701): d = cur - prior; d[d < thr] = 0
 *   ff_frame_three_difference     three_frame_difference (:704-740): min(|cur-prev|, |next-cur|), thresholded
 *   ff_frame_count_above          np.sum(frame > noise_threshold) of is_empty_frame (:759); count_dev int64[1]  */
int ff_frame_subtract_background(const void* image_dev, int px_type, int64_t n_px, double background,
                                 double* out_dev, void* stream);
int ff_frame_difference(const void* current_dev, const void* prior_dev, int px_type, int64_t n_px,
                        double threshold, double* out_dev, void* stream);
int ff_frame_three_difference(const void* prev_dev, const void* curr_dev, const void* next_dev, int px_type,
                              int64_t n_px, double threshold, double* out_dev, void* stream);
int ff_frame_count_above(const void* frame_dev, int px_type, int64_t n_px, double threshold,
                         int64_t* count_dev, void* stream);

/* ff_head_images: every full-frame intermediate FlameDetector.detect returns in its
 * FlameDetectionResult (scripts/process_videos.py:197-217, computed at :380-413), for n_frames frames
 * stored as in the .mraw file (or a single decoded frame: bits = 16 / 8):
 *   sub_out       frame_subtracted = max(x - bg, 0)                                  (:380)
 *   diff_out      frame_diff = sub - prior_sub, values < diff_thr zeroed             (:397-399)
 *   opened_out    noise_removed = grey_opening(frame_diff, size=(k,k)), k odd <= 7   (:403-404)
 *   blurred_out   gaussian_filter(noise_removed, sigma)                               (:407)
 *   sobel_out     sobel(blurred, axis=1)                                              (:410)
 *   gradient_out  np.gradient(blurred, axis=1)                                        (:413)
 * each float64[n_frames,H,W], nullable; float64 stages in SciPy's operation order, mode 'reflect'.
 *   bg / bg_halo        integer background scalar of the frames / of the halo frame (detect() keeps
 *                       its prior frame as it was subtracted at ITS call, :469); both >= 0
 *   gauss_weights_host  2*radius+1 float64 taps (scipy _gaussian_kernel1d), HOST pointer
 *   state_out_dev       uint8[n_frames], nullable: 0 listed in skip_dev (outputs zeroed), 1 all images
 *                       valid, 2 no prior frame (only sub_out is meaningful; the reference returns None
 *                       for the others, :386-393)                                                      */
int ff_head_images(const void* frames_dev, const void* halo_dev, int64_t n_frames, int height, int width,
                   int bits, int32_t bg, int32_t bg_halo, int32_t diff_thr, int morphology_size,
                   const double* gauss_weights_host, int radius, const uint8_t* skip_dev,
                   double* sub_out_dev, double* diff_out_dev, double* opened_out_dev, double* blurred_out_dev,
                   double* sobel_out_dev, double* gradient_out_dev, uint8_t* state_out_dev, void* stream);

/* ---- host-resident clips: chunked H2D streaming -----------------------------------------------
 * The end-to-end form of stages 2b-4 for a clip that lives in host memory (pinned, or the
 * mmapped .mraw file): replaces the whole frame loop of process_video_source
 * (scripts/process_videos.py:1441-1516) for one video / one contiguous frame range.
 * Frames are copied in chunks on a copy stream, double-buffered against the kernels; once a
 * finished chunk has reported an exit frame no further chunks are copied (the reference
 * `break`s at :1494).  Pageable sources go through the threaded bounce buffers described at
 * ff_host_upload.  Blocking: returns when pos_out_host/count_out_host are complete.
 *   ctx               from ff_host_ctx_create (owns staging buffers, streams, events)
 *   frames_host       n_frames frames; halo_host the frame before them or NULL
 *   bg                background scalar by value (host already synchronised on it)
 *   frames_done_out   number of leading frames actually processed (== n_frames unless an
 *                     exit stopped the copy early); frames beyond it hold FF_POS_DROPPED
 *   first_exit_out    global index of the first exit frame or FF_NO_EXIT                     */
typedef struct ff_host_ctx ff_host_ctx;
int ff_host_ctx_create(int device, int64_t chunk_bytes, ff_host_ctx** ctx_out);
int ff_host_ctx_destroy(ff_host_ctx* ctx);
/* Threads that fill the pinned bounce buffers from a pageable source (1..16; default FF_HOST_COPY_THREADS
 * or half the cores) - e.g. cores / ranks when several ranks share a host.                             */
int ff_host_ctx_set_copy_threads(ff_host_ctx* ctx, int n_threads);
/* Blocking host -> device copy of `bytes` bytes for the device-resident entry points.  Pinned or
 * registered sources are DMA'd in place; pageable ones (e.g. np.memmap of the .mraw file, the
 * array the reference gets from pyMRAW at src/photron/video.py:332) are moved through the
 * context's pinned bounce buffers, filled by the copy threads while the previous piece's DMA is in
 * flight - 40 GB/s from the page cache against 11 GB/s for a plain cudaMemcpy of pageable memory on
 * the measured box.                                                                             */
int ff_host_upload(ff_host_ctx* ctx, const void* src_host, void* dst_dev, int64_t bytes);
int ff_process_host(ff_host_ctx* ctx,
                    const void* frames_host, const void* halo_host, int64_t n_frames,
                    int64_t first_frame, int height, int width, int bits,
                    int32_t bg, int32_t empty_thr, int64_t min_signal_count,
                    int method, int use_frame_diff, int32_t diff_thr,
                    int32_t threshold_floor, int32_t grad2_bound, int32_t min_run_px,
                    int32_t exit_margin_px, const uint8_t* skip_host,
                    int32_t* pos_out_host, int32_t* count_out_host,
                    int64_t* frames_done_out, int32_t* first_exit_out);
/* The same for one rank's range of a range-sharded clip: results may stay on the device in the rank's range
 * block (pos_block_dev / count_block_dev / first_exit_block_dev from ff_exchange_begin; the host arrays are
 * then optional), `hooks` ties the range to the peers - wait before the block is written, exit frames
 * propagated to every rank, block published when the last chunk is done - and uploading stops at the first
 * chunk that starts at or behind the smallest exit frame ANY rank has seen.
 *   bytes_uploaded_out   host -> device bytes actually moved (the early stop is what keeps it small)     */
typedef struct ff_host_args {
  const void* frames_host;
  const void* halo_host;
  int64_t n_frames, first_frame;
  int32_t height, width, bits;
  int32_t bg, empty_thr;
  int32_t method, use_frame_diff, diff_thr, threshold_floor, grad2_bound, min_run_px, exit_margin_px;
  int64_t min_signal_count;
  const uint8_t* skip_host;
  int32_t* pos_out_host;
  int32_t* count_out_host;
  int32_t* pos_block_dev;
  int32_t* count_block_dev;
  int32_t* first_exit_block_dev;
  const ff_range_hooks* hooks;
  int64_t* frames_done_out;
  int32_t* first_exit_out;
  int64_t* bytes_uploaded_out;
} ff_host_args;
int ff_process_host_range(ff_host_ctx* ctx, const ff_host_args* args);

#ifdef __cplusplus
}
#endif
#endif /* FLAMEFRONT_H_ */

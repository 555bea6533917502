// the reference's HEAD detector (scripts/process_videos.py:350-465) on the GPU.
//
// head_band_kernel  - per non-empty frame, on the (2*(R+3)+1)-row band around the centre row:
//     thresholded frame difference (:397-399) -> 3x3 grey opening (:404) -> Gaussian sigma
//     (:407, scipy radius R = int(4*sigma+0.5)) -> Sobel(axis=1) and np.gradient(axis=1)
//     (:410-413); only the centre row of the last two is ever read (:417-418).
//     The morphology is exact integer work; the Gaussian/Sobel/gradient are float64 in SciPy,
//     so they are evaluated here in float64 with explicit round-to-nearest mul/add (no FMA) in
//     the exact operation order of scipy.ndimage's NI_Correlate1D (centre tap first, then
//     symmetric pairs outermost -> innermost) - positions come out bit-identical.
//     Boundaries: scipy mode 'reflect'.  Every stage is symmetric, so evaluating the stages on
//     the reflect-EXTENDED band equals reflecting each stage's output.
// head_flags_kernel - empty-frame decision per frame + the list of frames with work.
// head_track_*      - the sequential part (:317-348, :420-465): velocity-constrained search
//     window from the last detected position, arg-min gradient / rightmost Sobel, max of the
//     candidates, stop at the exit frame (:1488-1494) - run as a speculative parallel walk
//     (full-width answers for every frame, 16-frame segments walked at once, validated in
//     order); head_track_generic_kernel is the plain sequential walk kept as its cross-check.
#include <climits>
#include <cstdlib>
#include <type_traits>

#include "ff_common.cuh"

namespace ff {
namespace {

constexpr int kMaxRadius = 8;
constexpr int kHeadThreads = 256;
constexpr int kHeadTileW = 256;

struct HeadBandParams {
  const uint8_t* frames;
  const uint8_t* halo;
  int64_t frame_bytes;
  int n_frames;
  int height, width;
  const int32_t* bg_dev;
  const int32_t* partial;
  int partials_per_frame;
  int64_t min_signal_count;
  int diff_thr;
  const uint8_t* skip;
  double w[2 * kMaxRadius + 1];
  int radius;
  int morph;       // half size of the k x k grey opening (k = 2 * morph + 1; the reference's default k = 3)
  int halo_px;     // band half height / column halo: radius + 2 * morph + 1
  double* lines;   // [n][2][W]  (sobel row, gradient row)
  uint8_t* flags;  // [n]  0 = not processed (skipped / empty), 1 = lines valid, 2 = processed, no prior frame
  int32_t* scratch;  // [0] number of frames with flag 1, [4 + k] their indices (any order)
};

// One warp per frame: the empty-frame decision (:759-763, sum of the streaming kernel's partial
// counts), the flag, and the list of frames the band kernel has work for - so that kernel runs as
// persistent CTAs over the ~5 % of frames that hold a flame instead of launching (and retiring)
// a shared-memory-heavy CTA for every frame of the clip.
__global__ void __launch_bounds__(kHeadThreads) head_flags_kernel(const HeadBandParams p) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (kHeadThreads / 32) + (threadIdx.x >> 5);
  if (f >= p.n_frames) return;
  const bool skipped = p.skip != nullptr && p.skip[f] != 0;
  int cnt = 0;
  if (!skipped)
    for (int t = lane; t < p.partials_per_frame; t += 32) cnt += __ldg(p.partial + (int64_t)f * p.partials_per_frame + t);
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  if (lane != 0) return;
  const bool empty = (int64_t)cnt < p.min_signal_count;
  int flag = 0;
  if (!skipped && !empty) {
    int hf = f - 1;
    if (p.skip != nullptr)
      while (hf >= 0 && p.skip[hf]) --hf;
    flag = (hf >= 0 || p.halo != nullptr) ? 1 : 2;
  }
  p.flags[f] = (uint8_t)flag;
  if (flag == 1) p.scratch[4 + atomicAdd(p.scratch, 1)] = f;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if ((unsigned)i < (unsigned)n) return i;      // interior: no division
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// The float64 stages shared by both band kernels.  `band` holds the opened difference image as uint16,
// band row i / band column j at band[i * stride + j + col0]; rows HALO-1-R .. HALO+1+R and columns
// [2 morph, LW - 2 morph) must be valid (HALO = R + 2 morph + 1).  G0 = Gaussian along rows (axis 0) for band rows HALO-1, HALO, HALO+1; BL =
// Gaussian along columns (axis 1), valid columns [HALO-1, LW-HALO+1); then Sobel(axis=1) and
// np.gradient(axis=1) of the centre row.  scipy NI_Correlate1D order: centre tap first, then symmetric
// pairs outermost -> innermost, explicit round-to-nearest mul/add (no FMA).  Ends with a CTA barrier.
__device__ __forceinline__ void band_float_stages(const HeadBandParams& p, const uint16_t* band, int stride, int col0,
                                                  int LW, int tw, int x_begin, int f, double* g0, double* bl) {
  const int tid = threadIdx.x;
  const int R = p.radius, HALO = p.halo_px, W = p.width, M2 = 2 * p.morph;
  for (int e = tid; e < 3 * LW; e += kHeadThreads) {
    const int b = e / LW, j = e - b * LW;
    double tmp = 0.0;
    if (j >= M2 && j < LW - M2) {
      const uint16_t* col = band + (HALO - 1 + b) * stride + j + col0;
      tmp = __dmul_rn((double)col[0], p.w[R]);
      for (int jj = -R; jj < 0; ++jj) {
        const double pair = __dadd_rn((double)col[jj * stride], (double)col[-jj * stride]);
        tmp = __dadd_rn(tmp, __dmul_rn(pair, p.w[R + jj]));
      }
    }
    g0[e] = tmp;
  }
  __syncthreads();
  for (int e = tid; e < 3 * LW; e += kHeadThreads) {
    const int b = e / LW, j = e - b * LW;
    double tmp = 0.0;
    if (j >= HALO - 1 && j < LW - HALO + 1) {
      const double* g = g0 + b * LW;
      tmp = __dmul_rn(g[j], p.w[R]);
      for (int jj = -R; jj < 0; ++jj) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(g[j + jj], g[j - jj]), p.w[R + jj]));
    }
    bl[e] = tmp;
  }
  __syncthreads();
  double* out_s = p.lines + ((int64_t)f * 2 + 0) * W;
  double* out_g = p.lines + ((int64_t)f * 2 + 1) * W;
  for (int t = tid; t < tw; t += kHeadThreads) {
    const int j = HALO + t;
    const int x = x_begin + t;
    double s3[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double* v = bl + b * LW;
      // correlate1d([-1,0,1]): tmp = in[0]*0; tmp += (in[-1] - in[+1]) * (-1)
      s3[b] = __dadd_rn(__dmul_rn(v[j], 0.0), __dmul_rn(__dsub_rn(v[j - 1], v[j + 1]), -1.0));
    }
    // correlate1d([1,2,1]) along rows: tmp = in[0]*2; tmp += (in[-1] + in[+1]) * 1
    out_s[x] = __dadd_rn(__dmul_rn(s3[1], 2.0), __dmul_rn(__dadd_rn(s3[0], s3[2]), 1.0));
    const double* v = bl + 1 * LW;
    double g;
    if (x == 0) g = __ddiv_rn(__dsub_rn(v[j + 1], v[j]), 1.0);
    else if (x == W - 1) g = __ddiv_rn(__dsub_rn(v[j], v[j - 1]), 1.0);
    else g = __ddiv_rn(__dsub_rn(v[j + 1], v[j - 1]), 2.0);
    out_g[x] = g;
  }
  __syncthreads();      // the next work item reuses the shared-memory band
}

template <int BITS>
__global__ void __launch_bounds__(kHeadThreads) head_band_kernel(const HeadBandParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x;
  const int W = p.width, H = p.height;
  const int HALO = p.halo_px;
  const int MH = p.morph;
  const int NB = 2 * HALO + 1;            // band rows
  const int c = H / 2;
  const int tiles_x = (W + kHeadTileW - 1) / kHeadTileW;
  const int n_work = __ldg(p.scratch) * tiles_x;
  const int bg = __ldg(p.bg_dev);

  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int f = __ldg(p.scratch + 4 + work / tiles_x);
  const int x_begin = (work % tiles_x) * kHeadTileW;
  const int tw = min(kHeadTileW, W - x_begin);
  const int LW = tw + 2 * HALO;           // band columns held by this CTA
  int hf = f - 1;
  if (p.skip != nullptr)
    while (hf >= 0 && p.skip[hf]) --hf;
  const uint8_t* prior = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;

  // ---- shared-memory carve-up -----------------------------------------------------------------
  uint16_t* bufA = reinterpret_cast<uint16_t*>(smem);                    // [NB][LW]
  uint16_t* bufB = bufA + NB * LW;                                        // [NB][LW]
  const size_t u16_bytes = ((size_t)2 * NB * LW * sizeof(uint16_t) + 15) & ~(size_t)15;
  double* g0 = reinterpret_cast<double*>(smem + u16_bytes);               // [3][LW]
  double* bl = g0 + 3 * LW;                                               // [3][LW]

  const uint8_t* cur = p.frames + (int64_t)f * p.frame_bytes;

  // ---- D: thresholded difference on the reflect-extended band ----------------------------------
  // (warps walk rows, lanes walk columns: coalesced byte reads, no index division)
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < NB; i += kHeadThreads / 32) {
    const int64_t rowq = (int64_t)reflect_idx(c - HALO + i, H) * W;
    for (int j = lane; j < LW; j += 32) {
      const int64_t q = rowq + reflect_idx(x_begin - HALO + j, W);
      int d = max(load_px_generic<BITS>(cur, q) - bg, 0) - max(load_px_generic<BITS>(prior, q) - bg, 0);
      if (d < p.diff_thr) d = 0;
      bufA[i * LW + j] = (uint16_t)d;   // diff_thr >= 0 is enforced by the launcher: 0 <= d <= 65535
    }
  }
  __syncthreads();
  // ---- E = k x k minimum (valid rows [MH,NB-MH), cols [MH,LW-MH)); k = 2 MH + 1 (:403-404) -------
  for (int i = warp; i < NB; i += kHeadThreads / 32) {
    const bool row_ok = i >= MH && i < NB - MH;
    for (int j = lane; j < LW; j += 32) {
      unsigned m = 0;
      if (row_ok && j >= MH && j < LW - MH) {
        m = 0xFFFFu;
        for (int di = -MH; di <= MH; ++di)
          for (int dj = -MH; dj <= MH; ++dj) m = min(m, (unsigned)bufA[(i + di) * LW + j + dj]);
      }
      bufB[i * LW + j] = (uint16_t)m;
    }
  }
  __syncthreads();
  // ---- NR = k x k maximum of E (valid rows [2MH,NB-2MH), cols [2MH,LW-2MH)) ------------------------
  for (int i = warp; i < NB; i += kHeadThreads / 32) {
    const bool row_ok = i >= 2 * MH && i < NB - 2 * MH;
    for (int j = lane; j < LW; j += 32) {
      unsigned m = 0;
      if (row_ok && j >= 2 * MH && j < LW - 2 * MH) {
        for (int di = -MH; di <= MH; ++di)
          for (int dj = -MH; dj <= MH; ++dj) m = max(m, (unsigned)bufB[(i + di) * LW + j + dj]);
      }
      bufA[i * LW + j] = (uint16_t)m;
    }
  }
  __syncthreads();
  band_float_stages(p, bufA, LW, 0, LW, tw, x_begin, f, g0, bl);
  }
}

// ---- fast band kernel: rows that start on an 8-pixel boundary (W % 8 == 0, 4-byte aligned frames) --
// The general kernel above is issue-bound on per-pixel address arithmetic and byte loads.  Here stage D
// works on aligned groups of 8 pixels (three 32-bit loads per frame for packed 12-bit), decodes them to
// 16x2 words and takes the thresholded difference with DPX three-input min/max (10 instructions per pixel
// PAIR); the 3x3 opening never goes back to shared memory between its four passes (a lane walks down one
// word column with the last three rows in registers and trades edge pixels by shuffle); the Gaussian is
// unrolled for the radius (template) and shares one pass over the opened rows between its three output
// rows.  The band buffer starts at the aligned column x_begin - 16 (band column j lives at buffer column
// j + 16 - HALO) and is 288 columns wide.
constexpr int kBandPad = 16;
constexpr int kBandLWA = kHeadTileW + 2 * kBandPad;      // 288
constexpr int kBandGroups = kBandLWA / 8;                // 36 groups of 8 pixels per band row
constexpr int kBandWords = kBandLWA / 2;                 // 144 16x2 words per band row

template <int BITS>
__device__ __forceinline__ void load8_16x2(const uint8_t* __restrict__ base, int64_t q0, uint32_t (&x)[4]) {
  if (BITS == 12) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (q0 >> 1) * 3);
    decode12x8_16x2(__ldg(w), __ldg(w + 1), __ldg(w + 2), x);
  } else if (BITS == 16) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + q0 * 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldg(w + k);
  } else {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + q0);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
    x[0] = __byte_perm(w0, 0u, 0x4140);
    x[1] = __byte_perm(w0, 0u, 0x4342);
    x[2] = __byte_perm(w1, 0u, 0x4140);
    x[3] = __byte_perm(w1, 0u, 0x4342);
  }
}

// (double)n for 0 <= n < 2^32 without the conversion pipe: 2^52 + n is exact, and so is the subtraction
__device__ __forceinline__ double u2d(uint32_t n) {
  return __dsub_rn(__hiloint2double(0x43300000, (int)n), 4503599627370496.0);
}

// Sobel(axis=1) and np.gradient(axis=1) of the centre row from the three blurred rows `bl`
__device__ __forceinline__ void band_lines_out(const HeadBandParams& p, const double* bl, int LW, int HALO, int tw,
                                               int x_begin, int f) {
  const int W = p.width;
  double* out_s = p.lines + ((int64_t)f * 2 + 0) * W;
  double* out_g = p.lines + ((int64_t)f * 2 + 1) * W;
  for (int t = threadIdx.x; t < tw; t += kHeadThreads) {
    const int j = HALO + t;
    const int x = x_begin + t;
    double s3[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double* v = bl + b * LW;
      // correlate1d([-1,0,1]): tmp = in[0]*0; tmp += (in[-1] - in[+1]) * (-1)
      s3[b] = __dadd_rn(__dmul_rn(v[j], 0.0), __dmul_rn(__dsub_rn(v[j - 1], v[j + 1]), -1.0));
    }
    // correlate1d([1,2,1]) along rows: tmp = in[0]*2; tmp += (in[-1] + in[+1]) * 1
    out_s[x] = __dadd_rn(__dmul_rn(s3[1], 2.0), __dmul_rn(__dadd_rn(s3[0], s3[2]), 1.0));
    const double* v = bl + 1 * LW;
    double g;
    if (x == 0) g = __ddiv_rn(__dsub_rn(v[j + 1], v[j]), 1.0);
    else if (x == W - 1) g = __ddiv_rn(__dsub_rn(v[j], v[j - 1]), 1.0);
    else g = __ddiv_rn(__dsub_rn(v[j + 1], v[j - 1]), 2.0);
    out_g[x] = g;
  }
}

// The float64 stages with the radius known at compile time: one thread per column builds all three
// row-blurred values from ONE pass over the 2R+3 opened rows (a symmetric pair of uint16 is summed as an
// integer before its single conversion: exact, so the same double as scipy's in[-k] + in[+k]); taps come
// straight from the constant bank.  Same operation order as band_float_stages.
template <int R>
__device__ __forceinline__ void band_float_stages_fast(const HeadBandParams& p, const uint16_t* band, int off, int LW,
                                                       int tw, int x_begin, int f, double* g0, double* bl) {
  constexpr int HALO = R + 3;
  const int tid = threadIdx.x;
  for (int j = 2 + tid; j < LW - 2; j += kHeadThreads) {
    const uint16_t* col = band + (HALO - 1 - R) * kBandLWA + j + off;
    uint32_t v[2 * R + 3];
#pragma unroll
    for (int k = 0; k < 2 * R + 3; ++k) v[k] = col[k * kBandLWA];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double tmp = __dmul_rn(u2d(v[R + b]), p.w[R]);
#pragma unroll
      for (int jj = -R; jj < 0; ++jj) tmp = __dadd_rn(tmp, __dmul_rn(u2d(v[R + b + jj] + v[R + b - jj]), p.w[R + jj]));
      g0[b * LW + j] = tmp;
    }
  }
  __syncthreads();
  const int cols = tw + 2;
  for (int e = tid; e < 3 * cols; e += kHeadThreads) {
    const int b = (e >= cols) + (e >= 2 * cols);
    const int j = HALO - 1 + e - b * cols;
    const double* g = g0 + b * LW + j;
    double tmp = __dmul_rn(g[0], p.w[R]);
#pragma unroll
    for (int jj = -R; jj < 0; ++jj) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(g[jj], g[-jj]), p.w[R + jj]));
    bl[b * LW + j] = tmp;
  }
  __syncthreads();
  band_lines_out(p, bl, LW, HALO, tw, x_begin, f);
  __syncthreads();      // the next work item reuses the shared-memory band
}

// RT = the Gaussian radius when it is one of the instantiated ones (2, 4, 6, 8: sigma 0.5 .. 2), else 0
template <int BITS, int RT>
__global__ void __launch_bounds__(kHeadThreads) head_band_fast_kernel(const HeadBandParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x;
  const int W = p.width, H = p.height;
  const int R = RT ? RT : p.radius;
  const int HALO = R + 3;
  const int NB = 2 * HALO + 1;
  const int off = kBandPad - HALO;          // >= 5: kMaxRadius + 3 = 11 <= kBandPad
  const int c = H / 2;
  const int tiles_x = (W + kHeadTileW - 1) / kHeadTileW;
  const int n_work = __ldg(p.scratch) * tiles_x;
  const uint32_t bg = (uint32_t)__ldg(p.bg_dev);
  const uint32_t bg2 = bg | (bg << 16);
  const uint32_t tm1 = (uint32_t)max(p.diff_thr, 1) - 1u;      // d >= 0 always: thresholds 0 and 1 keep the same pixels
  const uint32_t t2 = tm1 | (tm1 << 16);
  uint16_t* bufA = reinterpret_cast<uint16_t*>(smem);                    // [NB][kBandLWA]  difference
  uint16_t* bufB = bufA + NB * kBandLWA;                                  // [NB][kBandLWA]  opened difference
  uint32_t* A32 = reinterpret_cast<uint32_t*>(bufA);
  uint32_t* B32 = reinterpret_cast<uint32_t*>(bufB);
  const int LWmax = kHeadTileW + 2 * HALO;
  double* g0 = reinterpret_cast<double*>(smem + (size_t)2 * NB * kBandLWA * sizeof(uint16_t));   // [3][LW]
  double* bl = g0 + 3 * LWmax;                                            // [3][LW]

  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int f = __ldg(p.scratch + 4 + work / tiles_x);
    const int x_begin = (work % tiles_x) * kHeadTileW;
    const int tw = min(kHeadTileW, W - x_begin);      // multiple of 8
    const int LW = tw + 2 * HALO;
    const int groups = (tw + 2 * kBandPad) / 8;
    const int words = groups * 4;
    int hf = f - 1;
    if (p.skip != nullptr)
      while (hf >= 0 && p.skip[hf]) --hf;
    const uint8_t* prior = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
    const uint8_t* cur = p.frames + (int64_t)f * p.frame_bytes;

    // ---- D: thresholded difference on 16x2 words, 8 pixels per task; groups outside the image are mirrored below
    for (int t = tid; t < NB * kBandGroups; t += kHeadThreads) {
      const int i = t / kBandGroups, g = t - i * kBandGroups;
      const int xg = x_begin - kBandPad + 8 * g;
      if (g >= groups || xg < 0 || xg + 8 > W) continue;
      const int64_t q0 = (int64_t)reflect_idx(c - HALO + i, H) * W + xg;
      uint32_t a[4], b[4], o[4];
      load8_16x2<BITS>(cur, q0, a);
      load8_16x2<BITS>(prior, q0, b);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t sa = __vimax3_u16x2(a[k], bg2, bg2) - bg2;          // max(a - bg, 0) per lane
        const uint32_t sb = __vimax3_u16x2(b[k], bg2, bg2) - bg2;
        const uint32_t rd = __vimax3_u16x2(sa, sb, sb) - sb;               // relu(d)
        const uint32_t r2 = __vimax3_u16x2(rd, t2, t2) - t2;               // relu(relu(d) - (thr-1))
        const uint32_t m2 = __vimin3_u16x2(r2, 0x00010001u, 0x00010001u);  // [d >= thr]
        o[k] = r2 + m2 * tm1;                                              // d where d >= thr, else 0
      }
      *reinterpret_cast<uint4*>(bufA + i * kBandLWA + 8 * g) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // scipy 'reflect' left of column 0 / right of column W-1: the mirrored pixel is one of this tile's
    // own (at most 16 columns inside the image, or anywhere in an image narrower than that)
    if (x_begin < kBandPad || x_begin + tw + kBandPad > W) {
      for (int t = tid; t < NB * 2 * kBandPad; t += kHeadThreads) {
        const int i = t / (2 * kBandPad), k = t - i * (2 * kBandPad);
        const int x = k < kBandPad ? k - kBandPad : W + k - kBandPad;          // 16 columns either side
        if (x < x_begin - kBandPad || x >= x_begin + tw + kBandPad) continue;    // not this tile's side
        bufA[i * kBandLWA + x - x_begin + kBandPad] = bufA[i * kBandLWA + reflect_idx(x, W) - x_begin + kBandPad];
      }
      __syncthreads();
    }
    // ---- 3x3 opening in registers: a lane owns one 16x2 word column and walks down the band rows, holding
    // the last three difference rows (vertical min), exchanging with its neighbours by shuffle (horizontal
    // min -> erosion row), then the same on the last three erosion rows with max.  A warp's first and last
    // lane only feed their neighbours, so warps overlap by two words; opened rows [2, NB-2) x words
    // [1, words-1) are written - a superset of what the Gaussian reads.
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int n_chunks = (words - 2 + 29) / 30;
      if (warp < n_chunks) {
        const int w = 30 * warp + lane;
        const bool st_ok = lane >= 1 && lane <= 30 && w <= words - 2;
        const uint32_t* src = A32 + min(w, words - 1);
        uint32_t* dst = B32 + w;
        uint32_t d0 = src[0], d1 = src[kBandWords];
        uint32_t e0 = 0u, e1 = 0u;
        for (int i = 2; i < NB; ++i) {
          const uint32_t d2 = src[i * kBandWords];
          const uint32_t v = __vimin3_u16x2(d0, d1, d2);                     // rows i-2 .. i
          d0 = d1;
          d1 = d2;
          const uint32_t vl = __shfl_up_sync(0xFFFFFFFFu, v, 1), vr = __shfl_down_sync(0xFFFFFFFFu, v, 1);
          const uint32_t e2 = __vimin3_u16x2(__byte_perm(vl, v, 0x5432), v, __byte_perm(v, vr, 0x5432));   // erosion row i-1
          if (i >= 4) {
            const uint32_t x = __vimax3_u16x2(e0, e1, e2);                   // erosion rows i-3 .. i-1
            const uint32_t xl = __shfl_up_sync(0xFFFFFFFFu, x, 1), xr = __shfl_down_sync(0xFFFFFFFFu, x, 1);
            const uint32_t o = __vimax3_u16x2(__byte_perm(xl, x, 0x5432), x, __byte_perm(x, xr, 0x5432));  // opened row i-2
            if (st_ok) dst[(i - 2) * kBandWords] = o;
          }
          e0 = e1;
          e1 = e2;
        }
      }
      __syncthreads();
    }
    if (RT) band_float_stages_fast<RT ? RT : 2>(p, bufB, off, LW, tw, x_begin, f, g0, bl);
    else band_float_stages(p, bufB, kBandLWA, off, LW, tw, x_begin, f, g0, bl);
  }
}

struct HeadTrackParams {
  const double* lines;
  const uint8_t* flags;
  int n_frames;
  int64_t first_frame;
  int width;
  int edge_margin, max_disp, window, exit_margin;
  double min_strength, sobel_frac;
  int last_frame_in, last_pos_in;  // tracker state carried in from an earlier range (-1 = none)
  int32_t* out;                    // [n][5]: final, min_gradient, rightmost_sobel, search_start, search_end
  int32_t* stop;                   // [3]: exit frame (global) or FF_NO_EXIT, last frame, last pos
  int32_t* seg;                    // [n_seg][8] per-segment states of the chained speculation, or nullptr
  int32_t* seg_hdr;                // [0] first segment whose speculative walk ends with a detection / carried state
  int32_t* out2;                   // [n][5] rows re-computed by head_track_fixup_kernel
};

// Monotone map from the bits of a non-NaN double to an unsigned integer (-0.0 == +0.0): the
// float64 comparisons of the reference run on these keys (integer compares and REDUX instead of
// FP64 compare chains); ties stay ties.
__device__ __forceinline__ unsigned long long order_key(unsigned long long bits) {
  if (bits == 0x8000000000000000ull) bits = 0ull;
  return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}

// The sequential walk, one CTA, frame by frame, straight from the reference (:317-348, :420-465,
// :1488-1494): the cross-check for the speculative tracker below (FF_TRACK_SEQUENTIAL=1).
__global__ void __launch_bounds__(kHeadThreads) head_track_generic_kernel(const HeadTrackParams p) {
  __shared__ double s_min[kHeadThreads / 32];
  __shared__ int s_arg[kHeadThreads / 32];
  __shared__ double s_amax[kHeadThreads / 32];
  __shared__ int s_right[kHeadThreads / 32];
  __shared__ int s_last_f, s_last_p, s_stop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.width;
  if (tid == 0) {
    s_last_f = p.last_frame_in;
    s_last_p = p.last_pos_in;
    s_stop = 0;
    p.stop[0] = FF_NO_EXIT;
  }
  __syncthreads();
  for (int f = 0; f < p.n_frames; ++f) {
    const int fl = p.flags[f];
    if (fl == 0) continue;
    const int gf = (int)(p.first_frame + f);
    int s0, s1;
    if (s_last_p < 0) {
      s0 = p.edge_margin;
      s1 = W - p.edge_margin;
    } else {
      s0 = s_last_p;
      s1 = min(W - p.edge_margin, s_last_p + p.max_disp * max(1, gf - s_last_f) + p.window);
    }
    int pos_a = -1, pos_b = -1;
    if (fl == 1 && s1 > s0 && s0 >= 0) {     // non-empty search slice (:424)
      s1 = min(s1, W);
      const double* sob = p.lines + ((int64_t)f * 2 + 0) * W;
      const double* grd = p.lines + ((int64_t)f * 2 + 1) * W;
      double mn = 1.0 / 0.0, amax = -1.0;
      int arg = INT_MAX;
      for (int x = s0 + tid; x < s1; x += kHeadThreads) {
        const double g = grd[x];
        if (g < mn) { mn = g; arg = x; }
        amax = fmax(amax, fabs(sob[x]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(0xFFFFFFFFu, mn, o);
        const int oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (om < mn || (om == mn && oa < arg)) { mn = om; arg = oa; }
        amax = fmax(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
      }
      if (lane == 0) { s_min[warp] = mn; s_arg[warp] = arg; s_amax[warp] = amax; }
      __syncthreads();
      mn = s_min[0]; arg = s_arg[0]; amax = s_amax[0];
#pragma unroll
      for (int k = 1; k < kHeadThreads / 32; ++k) {
        if (s_min[k] < mn || (s_min[k] == mn && s_arg[k] < arg)) { mn = s_min[k]; arg = s_arg[k]; }
        amax = fmax(amax, s_amax[k]);
      }
      if (mn < -p.min_strength) pos_a = arg;                                  // :427-430
      int right = -1;
      if (amax > p.min_strength) {                                            // :434-440
        const double thr = __dmul_rn(amax, p.sobel_frac);
        for (int x = s0 + tid; x < s1; x += kHeadThreads)
          if (fabs(sob[x]) > thr) right = x;
      }
      right = __reduce_max_sync(0xFFFFFFFFu, right);
      if (lane == 0) s_right[warp] = right;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kHeadThreads / 32; ++k) pos_b = max(pos_b, s_right[k]);
    }
    const int final_pos = max(pos_a, pos_b);                                  // :452-465
    __syncthreads();   // everyone has read s_last_* / reduction scratch
    if (tid == 0) {
      int32_t* o = p.out + (int64_t)f * 5;
      o[0] = final_pos; o[1] = pos_a; o[2] = pos_b; o[3] = s0; o[4] = s1;
      if (final_pos >= 0) { s_last_f = gf; s_last_p = final_pos; }
      if (final_pos >= 0 && final_pos >= W - p.exit_margin) {                 // :1488-1494
        p.stop[0] = gf;
        s_stop = 1;
      }
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (tid == 0) {
    p.stop[1] = s_last_f;
    p.stop[2] = s_last_p;
  }
}

// ---- speculative parallel tracker ---------------------------------------------------------------
// The walk is sequential only through its state (last detection frame, last position).  The clip
// is cut into segments of 32 consecutive frames, one warp each, spread over the whole GPU
// (head_track_spec_kernel): segment 0 starts from the true state, every other segment from
// "nothing detected yet", and each warp walks its segment SPECULATIVELY, writing its results.
// One warp then validates the segments in order (head_track_commit_kernel): it re-runs the first
// active frame(s) of a segment from the true state until its state equals the state the
// speculation had at the same point - from there on the two walks are the same deterministic
// function of (state, data), so the rest of the segment stands as computed - and finally clears
// everything after the exit frame.  The flame front moves a few pixels per frame and dominates its
// neighbourhood, so a walk started from scratch locks onto it within a frame or two; if it never
// does, validation simply degrades into the sequential walk.  Results are bit-identical to the
// sequential kernel either way (tests run both).  Lines are read straight from global memory
// (they sit in L2 after head_band_kernel).
constexpr int kSegFrames = 16;               // frames per segment (<= 32: one lane per frame); a speculative
                                             // walk is a latency chain of this many frames
constexpr int kSpecWarpsPerCta = 8;
constexpr int kSegInts = 8;                  // int32 per segment in the chained-speculation scratch
constexpr int kMaxRepair = 8;                // frames a segment re-runs from its guess before it gives up

struct TrackState { int last_f, last_p; };
struct FrameResult { int final_pos, pos_a, pos_b, s0, s1; };

__device__ __forceinline__ FrameResult track_frame(const HeadTrackParams& p, int f, int fl, TrackState st, int lane) {
  const unsigned fullmask = 0xFFFFFFFFu;
  const int W = p.width;
  const int gf = (int)(p.first_frame + f);
  FrameResult r;
  if (st.last_p < 0) {
    r.s0 = p.edge_margin;
    r.s1 = W - p.edge_margin;
  } else {
    r.s0 = st.last_p;
    r.s1 = min(W - p.edge_margin, st.last_p + p.max_disp * max(1, gf - st.last_f) + p.window);
  }
  r.pos_a = r.pos_b = -1;
  if (fl == 1 && r.s1 > r.s0 && r.s0 >= 0) {     // non-empty search slice (:424)
    r.s1 = min(r.s1, W);
    const int s0 = r.s0, s1 = r.s1;
    const unsigned long long* sob = reinterpret_cast<const unsigned long long*>(p.lines + (int64_t)f * 2 * W);
    const unsigned long long* grd = sob + W;
    unsigned long long mn = ~0ull, amax = 0ull;    // key(+inf) < ~0;  |x| keys start at 0
    int arg = INT_MAX;
    const bool small = s1 - s0 <= 128;
    unsigned long long a4[4];
    if (small) {                                   // all loads issued up front, window held in registers
      unsigned long long g4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = s0 + lane + 32 * j;
        const bool in = x < s1;
        g4[j] = in ? __ldg(grd + x) : 0x7FF8000000000000ull;
        a4[j] = in ? (__ldg(sob + x) & 0x7FFFFFFFFFFFFFFFull) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = s0 + lane + 32 * j;
        const unsigned long long g = x < s1 ? order_key(g4[j]) : ~0ull;
        if (g < mn) { mn = g; arg = x; }
        amax = a4[j] > amax ? a4[j] : amax;
      }
    } else {
      // wide window (nothing detected yet): 128 columns per round, the 8 loads of a round in flight together
      for (int x0 = s0; x0 < s1; x0 += 128) {
        unsigned long long g4[4], b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = x0 + lane + 32 * j;
          const bool in = x < s1;
          g4[j] = in ? __ldg(grd + x) : 0x7FF8000000000000ull;
          b4[j] = in ? (__ldg(sob + x) & 0x7FFFFFFFFFFFFFFFull) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int x = x0 + lane + 32 * j;
          const unsigned long long g = x < s1 ? order_key(g4[j]) : ~0ull;
          if (g < mn) { mn = g; arg = x; }
          amax = b4[j] > amax ? b4[j] : amax;
        }
      }
    }
    // first minimum: smallest key, then smallest index (np.argmin)
    const unsigned mn_hi = __reduce_min_sync(fullmask, (unsigned)(mn >> 32));
    const unsigned lo_c = (unsigned)(mn >> 32) == mn_hi ? (unsigned)mn : 0xFFFFFFFFu;
    const unsigned mn_lo = __reduce_min_sync(fullmask, lo_c);
    const bool mine = (unsigned)(mn >> 32) == mn_hi && (unsigned)mn == mn_lo;
    arg = (int)__reduce_min_sync(fullmask, mine ? (unsigned)arg : 0x7FFFFFFFu);
    const unsigned long long mn_key = ((unsigned long long)mn_hi << 32) | mn_lo;
    const unsigned am_hi = __reduce_max_sync(fullmask, (unsigned)(amax >> 32));
    const unsigned am_lo = __reduce_max_sync(fullmask, (unsigned)(amax >> 32) == am_hi ? (unsigned)amax : 0u);
    if (mn_key < order_key((unsigned long long)__double_as_longlong(-p.min_strength))) r.pos_a = arg;   // :427-430
    const double amax_d = __longlong_as_double((long long)(((unsigned long long)am_hi << 32) | am_lo));
    if (amax_d > p.min_strength) {                                                        // :434-440
      const unsigned long long thr = (unsigned long long)__double_as_longlong(__dmul_rn(amax_d, p.sobel_frac));
      int right = -1;
      if (small) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (a4[j] > thr) right = s0 + lane + 32 * j;      // out-of-window slots hold 0 and thr >= 0
      } else {
        for (int x0 = s0; x0 < s1; x0 += 128) {
          unsigned long long b4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = x0 + lane + 32 * j;
            b4[j] = x < s1 ? (__ldg(sob + x) & 0x7FFFFFFFFFFFFFFFull) : 0ull;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b4[j] > thr) right = x0 + lane + 32 * j;
        }
      }
      r.pos_b = __reduce_max_sync(fullmask, right);
    }
  }
  r.final_pos = max(r.pos_a, r.pos_b);                                                   // :452-465
  return r;
}

// Search result of every active frame for the state "nothing detected yet" (whole width between
// the edge margins) - one warp per frame, all frames at once.  The speculative walks start in that
// state and stay in it through frames without a detection (flame entering, burnt gas after the
// exit), so without this a walk would pay a 1000-column search again and again.
__global__ void __launch_bounds__(kSpecWarpsPerCta * 32) head_track_fullwidth_kernel(const HeadTrackParams p) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * kSpecWarpsPerCta + (threadIdx.x >> 5);
  if (f >= p.n_frames) return;
  const int fl = (int)p.flags[f];
  if (fl == 0) return;
  TrackState none;
  none.last_f = -1;
  none.last_p = -1;
  const FrameResult r = track_frame(p, f, fl, none, lane);
  if (lane == 0) {
    int32_t* o = p.out + (int64_t)f * 5;
    o[0] = r.final_pos; o[1] = r.pos_a; o[2] = r.pos_b; o[3] = r.s0; o[4] = r.s1;
  }
}

__global__ void __launch_bounds__(kSpecWarpsPerCta * 32) head_track_spec_kernel(const HeadTrackParams p) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * kSpecWarpsPerCta + (threadIdx.x >> 5);
  const int f_lo = seg * kSegFrames;
  if (f_lo >= p.n_frames) return;
  const unsigned fullmask = 0xFFFFFFFFu;
  const int fme = f_lo + lane;
  const int flme = (lane < kSegFrames && fme < p.n_frames) ? (int)p.flags[fme] : 0;
  unsigned act = __ballot_sync(fullmask, flme != 0);
  TrackState st;
  st.last_f = seg == 0 ? p.last_frame_in : -1;
  st.last_p = seg == 0 ? p.last_pos_in : -1;
  while (act) {
    const int l = __ffs((int)act) - 1;
    act &= act - 1;
    const int f = f_lo + l;
    const int fl = __shfl_sync(fullmask, flme, l);
    if (st.last_p < 0) {            // still nothing detected: the full-width answer is already in out[f]
      const int fp = __ldcg(p.out + (int64_t)f * 5);
      if (fp >= 0) {
        st.last_f = (int)(p.first_frame + f);
        st.last_p = fp;
      }
      continue;
    }
    const FrameResult r = track_frame(p, f, fl, st, lane);
    if (r.final_pos >= 0) {
      st.last_f = (int)(p.first_frame + f);
      st.last_p = r.final_pos;
    }
    if (lane == 0) {
      int32_t* o = p.out + (int64_t)f * 5;
      o[0] = r.final_pos; o[1] = r.pos_a; o[2] = r.pos_b; o[3] = r.s0; o[4] = r.s1;
    }
  }
  if (p.seg != nullptr && lane == 0) {       // E1: where this walk ended (segment 0: from the true state)
    int32_t* e = p.seg + (int64_t)seg * kSegInts;
    e[0] = st.last_f;
    e[1] = st.last_p;
    e[4] = 0;                                // frames repaired by the fix-up pass
    e[5] = p.n_frames;                       // exit frame inside this segment (fix-up pass)
    if (st.last_p >= 0) atomicMin(p.seg_hdr, seg);      // first segment that ends with a state
  }
}

// ---- chained speculation --------------------------------------------------------------------------
// A walk started from "nothing detected yet" differs from the true walk only until it has locked onto
// the front - typically its first active frame.  head_track_fixup_kernel repairs exactly those frames
// for ALL segments at once: segment s guesses its true start state G_s = the end state of the nearest
// earlier speculative walk that ended with a state (detection or carried-in), re-runs its leading
// active frames from G_s until its state meets the speculation's, writes the repaired rows (saving
// the speculation's in out2) and records where it ended (E2) and the first exit frame it saw.  The
// guess is right for every segment up to the first one whose repaired walk did NOT end in the state
// its successors guessed (or that gave up after kMaxRepair frames) - found for all segments in
// parallel by head_track_resolve_kernel, which then needs no sequential work at all unless such a
// segment exists before the exit frame; from there it puts the speculation's rows back and falls
// back to validating segment by segment.
//   seg[s] = { E1.f, E1.p, E2.f, E2.p, n_fixed (0: not re-run), exit frame in s or n_frames, G.f, G.p }
__global__ void __launch_bounds__(kSpecWarpsPerCta * 32) head_track_fixup_kernel(const HeadTrackParams p) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * kSpecWarpsPerCta + (threadIdx.x >> 5);
  const int f_lo = seg * kSegFrames;
  if (f_lo >= p.n_frames) return;
  const unsigned fullmask = 0xFFFFFFFFu;
  const int W = p.width;
  const int fme = f_lo + lane;
  const int flme = (lane < kSegFrames && fme < p.n_frames) ? (int)p.flags[fme] : 0;
  const unsigned act_all = __ballot_sync(fullmask, flme != 0);
  if (!act_all) return;
  int32_t* e = p.seg + (int64_t)seg * kSegInts;
  // G: nearest earlier segment whose speculative walk ended with a state (none before seg_hdr[0])
  TrackState cur;
  cur.last_f = -1;
  cur.last_p = -1;
  const int first_with_state = __ldcg(p.seg_hdr);
  for (int t0 = seg - 1; t0 >= first_with_state; t0 -= 32) {
    const int t = t0 - lane;
    const int lp = t >= 0 ? __ldcg(p.seg + (int64_t)t * kSegInts + 1) : -1;
    const unsigned m = __ballot_sync(fullmask, lp >= 0);
    if (m) {
      const int l = __ffs((int)m) - 1;               // lowest lane = nearest segment
      cur.last_p = __shfl_sync(fullmask, lp, l);
      cur.last_f = __ldcg(p.seg + (int64_t)(t0 - l) * kSegInts + 0);
      break;
    }
  }
  int my_final = flme != 0 ? __ldcg(p.out + (int64_t)fme * 5) : -1;      // the speculation's result
  int n_fixed = 0;
  // Segment 0 started from the true state; without a state before, the speculation IS the walk; and a
  // guess that is an exit detection of THIS range means the walk has ended before this segment if the
  // guess holds (a carried-in state stops nothing, whatever its position).
  if (seg != 0 && cur.last_p >= 0 && (cur.last_p < W - p.exit_margin || cur.last_f < p.first_frame)) {
    const TrackState guess = cur;
    const int spec_final = my_final;
    TrackState spec;
    spec.last_f = -1;
    spec.last_p = -1;
    unsigned act = act_all;
    bool exit_hit = false;
    while (act && n_fixed < kMaxRepair && !(spec.last_f == cur.last_f && spec.last_p == cur.last_p)) {
      const int l = __ffs((int)act) - 1;
      act &= act - 1;
      const int f = f_lo + l;
      const int fl = __shfl_sync(fullmask, flme, l);
      const FrameResult r = track_frame(p, f, fl, cur, lane);
      const int sf = __shfl_sync(fullmask, spec_final, l);
      if (sf >= 0) {
        spec.last_f = (int)(p.first_frame + f);
        spec.last_p = sf;
      }
      if (r.final_pos >= 0) {
        cur.last_f = (int)(p.first_frame + f);
        cur.last_p = r.final_pos;
      }
      if (lane < 5) {                        // keep the speculation's row, then replace it
        int32_t* o = p.out + (int64_t)f * 5 + lane;
        p.out2[(int64_t)f * 5 + lane] = *o;
        *o = lane == 0 ? r.final_pos : lane == 1 ? r.pos_a : lane == 2 ? r.pos_b : lane == 3 ? r.s0 : r.s1;
      }
      if (lane == l) my_final = r.final_pos;
      ++n_fixed;
      if (r.final_pos >= 0 && r.final_pos >= W - p.exit_margin) {      // the walk ends here if the guess holds
        exit_hit = true;
        break;
      }
    }
    if (lane == 0) {
      const bool met = spec.last_f == cur.last_f && spec.last_p == cur.last_p;
      const bool own = e[1] >= 0;
      if (met || exit_hit) {                 // from here on the walks coincide (or nothing follows): they end alike
        e[2] = own ? e[0] : guess.last_f;
        e[3] = own ? e[1] : guess.last_p;
      } else if (!act) {                     // every active frame re-run
        e[2] = cur.last_f;
        e[3] = cur.last_p;
      } else {                               // gave up after kMaxRepair frames: the segment is not resolved
        e[2] = -2;
        e[3] = -2;
      }
      e[6] = guess.last_f;
      e[7] = guess.last_p;
    }
  }
  const unsigned ext = __ballot_sync(fullmask, my_final >= 0 && my_final >= W - p.exit_margin);      // :1488-1494
  if (lane == 0) {
    e[4] = n_fixed;
    e[5] = ext ? f_lo + __ffs((int)ext) - 1 : p.n_frames;
  }
}

struct WalkEnd { int exit_f; TrackState cur; };

// Validation of the speculative walks one segment after the other, by ONE warp, from segment
// `seg_begin` on with the true state `cur`: re-run the first active frame(s) of a segment until the
// state equals the state the speculation had at the same point; the rest of the segment stands.
// Returns the range-local exit frame (or n_frames) and the state at the end / at the exit.
__device__ WalkEnd commit_walk(const HeadTrackParams& p, int seg_begin, TrackState cur, int lane) {
  const unsigned fullmask = 0xFFFFFFFFu;
  const int W = p.width;
  int exit_f = p.n_frames;
  const int n_seg = (p.n_frames + kSegFrames - 1) / kSegFrames;
  const bool vec_flags = (reinterpret_cast<uintptr_t>(p.flags) & 15u) == 0;
  for (int seg0 = seg_begin & ~31; seg0 < n_seg && exit_f == p.n_frames; seg0 += 32) {
    // lane L summarises segment seg0 + L: bit j of m_act = frame j reached the detector, bit j of
    // m_one = it has a difference image (flag 1).  One round of loads per 32 segments keeps the
    // long empty stretches of a clip off the sequential path.
    unsigned m_act = 0, m_one = 0;
    if (seg0 + lane >= seg_begin) {
      const int fs = (seg0 + lane) * kSegFrames;
      if (vec_flags && fs + kSegFrames <= p.n_frames) {
        static_assert(kSegFrames == 16 || kSegFrames == 32, "flag masks are read as one or two 16-byte pieces");
        uint32_t w8[8] = {};
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(p.flags + fs));
        w8[0] = v0.x; w8[1] = v0.y; w8[2] = v0.z; w8[3] = v0.w;
        if (kSegFrames == 32) {
          const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(p.flags + fs) + 1);
          w8[4] = v1.x; w8[5] = v1.y; w8[6] = v1.z; w8[7] = v1.w;
        }
#pragma unroll
        for (int j = 0; j < kSegFrames; ++j) {
          const uint32_t fl = (w8[j >> 2] >> (8 * (j & 3))) & 0xFFu;
          if (fl != 0) m_act |= 1u << j;
          if (fl == 1) m_one |= 1u << j;
        }
      } else {
        for (int j = 0; j < kSegFrames; ++j) {
          const int f = fs + j;
          const int fl = f < p.n_frames ? (int)p.flags[f] : 0;
          if (fl != 0) m_act |= 1u << j;
          if (fl == 1) m_one |= 1u << j;
        }
      }
    }
    unsigned busy = __ballot_sync(fullmask, m_act != 0);
    while (busy && exit_f == p.n_frames) {
      const int sl = __ffs((int)busy) - 1;
      busy &= busy - 1;
      const int seg = seg0 + sl;
      const int f_lo = seg * kSegFrames;
      const int fme = f_lo + lane;
      unsigned act = __shfl_sync(fullmask, m_act, sl);
      const unsigned one = __shfl_sync(fullmask, m_one, sl);
      const int flme = ((act >> lane) & 1u) ? (((one >> lane) & 1u) ? 1 : 2) : 0;
      if (!act) continue;
      // what the speculation produced for my frame (before anything is overwritten)
      const int spec_final = flme != 0 ? p.out[(int64_t)fme * 5] : -1;
      TrackState spec;              // the speculation's state BEFORE the next active frame
      spec.last_f = seg == 0 ? p.last_frame_in : -1;
      spec.last_p = seg == 0 ? p.last_pos_in : -1;
      // ---- re-run from the true state until it meets the speculation's state -------------------------
      while (act && !(spec.last_f == cur.last_f && spec.last_p == cur.last_p)) {
        const int l = __ffs((int)act) - 1;
        act &= act - 1;
        const int f = f_lo + l;
        const int fl = __shfl_sync(fullmask, flme, l);
        const FrameResult r = track_frame(p, f, fl, cur, lane);      // first: its loads overlap spec_final's
        const int sf = __shfl_sync(fullmask, spec_final, l);
        if (sf >= 0) {                                   // the speculation's state after this frame
          spec.last_f = (int)(p.first_frame + f);
          spec.last_p = sf;
        }
        if (r.final_pos >= 0) {
          cur.last_f = (int)(p.first_frame + f);
          cur.last_p = r.final_pos;
        }
        if (lane == 0) {
          int32_t* o = p.out + (int64_t)f * 5;
          o[0] = r.final_pos; o[1] = r.pos_a; o[2] = r.pos_b; o[3] = r.s0; o[4] = r.s1;
        }
        if (r.final_pos >= 0 && r.final_pos >= W - p.exit_margin) {                      // :1488-1494
          exit_f = f;
          act = 0;
        }
      }
      if (exit_f != p.n_frames || !act) continue;
      // ---- the remaining frames of the segment stand as speculated ----------------------------------------
      const bool mine = ((act >> lane) & 1u) != 0;
      const unsigned det = __ballot_sync(fullmask, mine && spec_final >= 0);
      const unsigned ext = __ballot_sync(fullmask, mine && spec_final >= 0 && spec_final >= W - p.exit_margin);
      unsigned upto = det;                                // detections up to and including the exit frame
      if (ext) {
        const int le = __ffs((int)ext) - 1;
        exit_f = f_lo + le;
        upto &= (2u << le) - 1u;
      }
      if (upto) {
        const int ll = 31 - __clz((int)upto);
        cur.last_f = (int)(p.first_frame + f_lo + ll);
        cur.last_p = __shfl_sync(fullmask, spec_final, ll);
      }
    }   // busy segments of this group
  }   // groups of 32 segments
  WalkEnd w;
  w.exit_f = exit_f;
  w.cur = cur;
  return w;
}

// frames after the exit frame were never reached by the reference loop (:1494)
__device__ __forceinline__ void clear_after_exit(const HeadTrackParams& p, int exit_f) {
  const int64_t first_dead = (int64_t)exit_f + 1;
  for (int64_t i = first_dead * 5 + threadIdx.x; i < (int64_t)p.n_frames * 5; i += blockDim.x) p.out[i] = -1;
}

__global__ void __launch_bounds__(256) head_track_commit_kernel(const HeadTrackParams p) {
  __shared__ int s_exit_f;          // range-local index of the exit frame, or n_frames
  if (threadIdx.x < 32) {
    TrackState cur;
    cur.last_f = p.last_frame_in;
    cur.last_p = p.last_pos_in;
    const WalkEnd w = commit_walk(p, 0, cur, threadIdx.x);
    if (threadIdx.x == 0) {
      s_exit_f = w.exit_f;
      p.stop[0] = w.exit_f == p.n_frames ? FF_NO_EXIT : (int)(p.first_frame + w.exit_f);
      p.stop[1] = w.cur.last_f;
      p.stop[2] = w.cur.last_p;
    }
  }
  __syncthreads();
  clear_after_exit(p, s_exit_f);
}

// Chained speculation, final step: everything in parallel unless a repaired walk missed its guess.
__global__ void __launch_bounds__(256) head_track_resolve_kernel(const HeadTrackParams p) {
  __shared__ int s_u, s_exit_f, s_last_seg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned fullmask = 0xFFFFFFFFu;
  const int n_seg = (p.n_frames + kSegFrames - 1) / kSegFrames;
  if (tid == 0) {
    s_u = n_seg;
    s_exit_f = p.n_frames;
    s_last_seg = -1;
  }
  __syncthreads();
  // ---- first segment whose repaired walk did not end in the state its successors guessed -------------
  for (int sg = tid; sg < n_seg; sg += blockDim.x) {
    const int4 lo = __ldcg(reinterpret_cast<const int4*>(p.seg + (int64_t)sg * kSegInts));
    const int4 hi = __ldcg(reinterpret_cast<const int4*>(p.seg + (int64_t)sg * kSegInts) + 1);
    if (hi.x > 0) {                               // re-run from a guess (hi.z, hi.w)
      const bool own = lo.y >= 0;                 // the speculative walk ended with a state: later guesses use E1
      const int tf = own ? lo.x : hi.z, tp = own ? lo.y : hi.w;
      if (lo.z != tf || lo.w != tp) atomicMin(&s_u, sg);
    }
  }
  __syncthreads();
  const int u = s_u;                              // the rows of segments 0..u-1 are the true walk's
  // ---- exit frame and last state among them -----------------------------------------------------------
  for (int sg = tid; sg < u; sg += blockDim.x) {
    const int4 lo = __ldcg(reinterpret_cast<const int4*>(p.seg + (int64_t)sg * kSegInts));
    const int4 hi = __ldcg(reinterpret_cast<const int4*>(p.seg + (int64_t)sg * kSegInts) + 1);
    if (hi.y < p.n_frames) atomicMin(&s_exit_f, hi.y);
    if ((hi.x > 0 ? lo.w : lo.y) >= 0) atomicMax(&s_last_seg, sg);
  }
  __syncthreads();
  int exit_f = s_exit_f;
  const bool fallback = exit_f == p.n_frames && u < n_seg;
  if (fallback) {
    // a guess missed before any exit: segments from u on get the speculation's rows back ...
    for (int sg = u + warp; sg < n_seg; sg += blockDim.x / 32) {
      const int n_fixed = p.seg[(int64_t)sg * kSegInts + 4];
      if (n_fixed == 0) continue;
      const int fme = sg * kSegFrames + lane;
      const int flme = (lane < kSegFrames && fme < p.n_frames) ? (int)p.flags[fme] : 0;
      const unsigned act = __ballot_sync(fullmask, flme != 0);
      if (flme != 0 && __popc(act & ((1u << lane) - 1u)) < n_fixed) {
#pragma unroll
        for (int k = 0; k < 5; ++k) p.out[(int64_t)fme * 5 + k] = p.out2[(int64_t)fme * 5 + k];
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    TrackState cur;
    cur.last_f = p.last_frame_in;
    cur.last_p = p.last_pos_in;
    if (exit_f < p.n_frames) {
      cur.last_f = (int)(p.first_frame + exit_f);
      cur.last_p = p.out[(int64_t)exit_f * 5];
    } else if (!fallback) {
      const int ls = s_last_seg;
      if (ls >= 0) {
        const int32_t* e = p.seg + (int64_t)ls * kSegInts;
        const bool ran = e[4] > 0;
        cur.last_f = ran ? e[2] : e[0];
        cur.last_p = ran ? e[3] : e[1];
      }
    } else {                                      // ... and are validated one after the other from the state
      cur.last_f = p.seg[(int64_t)u * kSegInts + 6];          // segment u guessed (right, as all before it hold)
      cur.last_p = p.seg[(int64_t)u * kSegInts + 7];
      const WalkEnd w = commit_walk(p, u, cur, lane);
      exit_f = w.exit_f;
      cur = w.cur;
    }
    if (lane == 0) {
      s_exit_f = exit_f;
      p.stop[0] = exit_f == p.n_frames ? FF_NO_EXIT : (int)(p.first_frame + exit_f);
      p.stop[1] = cur.last_f;
      p.stop[2] = cur.last_p;
    }
  }
  __syncthreads();
  clear_after_exit(p, s_exit_f);
}

}  // namespace

int64_t head_track_scratch_len(int64_t n_frames) {
  return ((5 * n_frames + 3) & ~(int64_t)3) + kSegInts * ((n_frames + kSegFrames - 1) / kSegFrames) + 8;
}

int head_lines_impl(const void* frames, const void* halo, int64_t n_frames, int height, int width, int bits,
                    const int32_t* bg_dev, const int32_t* partial, int64_t min_signal_count, int32_t diff_thr,
                    int morphology_size, const double* gauss_weights_host, int radius, const uint8_t* skip,
                    double* lines_out, uint8_t* flags_out, int32_t* scratch, cudaStream_t st) {
  if (frames == nullptr || bg_dev == nullptr || partial == nullptr || gauss_weights_host == nullptr ||
      lines_out == nullptr || flags_out == nullptr || scratch == nullptr)
    return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width < 2 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  if (radius < 0 || radius > kMaxRadius) return FF_ERR_UNSUPPORTED;
  if (morphology_size < 1 || morphology_size > 7 || (morphology_size & 1) == 0) return FF_ERR_UNSUPPORTED;
  if (diff_thr < 0) return FF_ERR_UNSUPPORTED;       // band is held as uint16
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;

  HeadBandParams p{};
  p.frames = static_cast<const uint8_t*>(frames);
  p.halo = static_cast<const uint8_t*>(halo);
  p.frame_bytes = frame_bytes_of(px, bits);
  p.height = height;
  p.width = width;
  p.bg_dev = bg_dev;
  p.partial = partial;
  p.partials_per_frame = choose_tiling(px).partials_per_frame;
  p.min_signal_count = min_signal_count;
  p.diff_thr = diff_thr;
  p.skip = skip;
  p.radius = radius;
  p.morph = (morphology_size - 1) / 2;
  p.halo_px = radius + 2 * p.morph + 1;
  for (int i = 0; i < 2 * radius + 1; ++i) p.w[i] = gauss_weights_host[i];
  p.lines = lines_out;
  p.flags = flags_out;

  const int halo_px = p.halo_px;
  const int nb = 2 * halo_px + 1;
  const int lw = kHeadTileW + 2 * halo_px;
  const size_t smem = (((size_t)2 * nb * lw * sizeof(uint16_t) + 15) & ~(size_t)15) + (size_t)6 * lw * sizeof(double);
  p.n_frames = (int)n_frames;
  p.scratch = scratch;
  FF_CUDA_TRY(cudaMemsetAsync(scratch, 0, 4 * sizeof(int32_t), st));
  const int warps = kHeadThreads / 32;
  head_flags_kernel<<<(unsigned)((n_frames + warps - 1) / warps), kHeadThreads, 0, st>>>(p);
  FF_CUDA_TRY(cudaGetLastError());
  const int tiles_x = (width + kHeadTileW - 1) / kHeadTileW;
  const bool fast = morphology_size == 3 && diff_thr <= 65535 && (width % 8) == 0 && (p.frame_bytes % 4) == 0 && getenv("FF_BAND_GENERAL") == nullptr &&
                    (reinterpret_cast<uintptr_t>(frames) % 4) == 0 &&
                    (halo == nullptr || (reinterpret_cast<uintptr_t>(halo) % 4) == 0);
  const size_t smem_general = smem;
  const size_t smem_fast = (size_t)2 * nb * kBandLWA * sizeof(uint16_t) + (size_t)6 * lw * sizeof(double);
  auto launch = [&](auto kern, size_t smem) -> int {
    if (smem > 48 * 1024) FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kHeadThreads, smem));
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (int64_t)sms * (occ > 0 ? occ : 1);          // persistent CTAs over the active list
    if (grid > n_frames * tiles_x) grid = n_frames * tiles_x;
    kern<<<(unsigned)grid, kHeadThreads, smem, st>>>(p);
    FF_CUDA_TRY(cudaGetLastError());
    return FF_OK;
  };
  if (fast) {
    auto by_radius = [&](auto bits_tag) -> int {
      constexpr int B = decltype(bits_tag)::value;
      switch (radius) {
        case 2: return launch(head_band_fast_kernel<B, 2>, smem_fast);
        case 4: return launch(head_band_fast_kernel<B, 4>, smem_fast);
        case 6: return launch(head_band_fast_kernel<B, 6>, smem_fast);
        case 8: return launch(head_band_fast_kernel<B, 8>, smem_fast);
        default: return launch(head_band_fast_kernel<B, 0>, smem_fast);
      }
    };
    switch (bits) {
      case 8: return by_radius(std::integral_constant<int, 8>{});
      case 12: return by_radius(std::integral_constant<int, 12>{});
      default: return by_radius(std::integral_constant<int, 16>{});
    }
  }
  switch (bits) {
    case 8: return launch(head_band_kernel<8>, smem_general);
    case 12: return launch(head_band_kernel<12>, smem_general);
    default: return launch(head_band_kernel<16>, smem_general);
  }
}

int head_track_impl(const double* lines, const uint8_t* flags, int64_t n_frames, int64_t first_frame, int width,
                    int32_t edge_margin_px, int32_t max_displacement_px, int32_t search_window_px,
                    double min_gradient_strength, double sobel_threshold_fraction, int32_t exit_margin_px,
                    int32_t last_frame_in, int32_t last_pos_in, int32_t* out, int32_t* stop, int32_t* scratch,
                    cudaStream_t st) {
  if (lines == nullptr || flags == nullptr || out == nullptr || stop == nullptr) return FF_ERR_INVALID;
  if (n_frames <= 0 || width < 2 || n_frames > 0x7FFFFFFF || first_frame < 0) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaMemsetAsync(out, 0xFF, sizeof(int32_t) * 5 * (size_t)n_frames, st));
  HeadTrackParams p{};
  p.lines = lines;
  p.flags = flags;
  p.n_frames = (int)n_frames;
  p.first_frame = first_frame;
  p.width = width;
  p.edge_margin = edge_margin_px;
  p.max_disp = max_displacement_px;
  p.window = search_window_px;
  p.exit_margin = exit_margin_px;
  p.min_strength = min_gradient_strength;
  p.sobel_frac = sobel_threshold_fraction;
  p.last_frame_in = last_frame_in;
  p.last_pos_in = last_pos_in;
  p.out = out;
  p.stop = stop;
  // one segment's worth of frames (e.g. FlameDetector.detect: a single frame) gains nothing from
  // speculation: one launch of the plain walk instead of four
  if (getenv("FF_TRACK_SEQUENTIAL") == nullptr && n_frames > kSegFrames) {       // speculative parallel walk
    const int64_t n_seg = (n_frames + kSegFrames - 1) / kSegFrames;
    const unsigned seg_ctas = (unsigned)((n_seg + kSpecWarpsPerCta - 1) / kSpecWarpsPerCta);
    const bool chained = scratch != nullptr && getenv("FF_TRACK_UNCHAINED") == nullptr;
    if (chained) {
      if (reinterpret_cast<uintptr_t>(scratch) % 16 != 0) return FF_ERR_ALIGNMENT;
      const int64_t rows = (5 * n_frames + 3) & ~(int64_t)3;          // segment records start 16-byte aligned
      p.out2 = scratch;
      p.seg = scratch + rows;
      p.seg_hdr = p.seg + kSegInts * n_seg;
      FF_CUDA_TRY(cudaMemsetAsync(p.seg_hdr, 0x7F, 4 * sizeof(int32_t), st));
    }
    head_track_fullwidth_kernel<<<(unsigned)((n_frames + kSpecWarpsPerCta - 1) / kSpecWarpsPerCta), kSpecWarpsPerCta * 32, 0, st>>>(p);
    FF_CUDA_TRY(cudaGetLastError());
    head_track_spec_kernel<<<seg_ctas, kSpecWarpsPerCta * 32, 0, st>>>(p);
    FF_CUDA_TRY(cudaGetLastError());
    if (chained) {
      head_track_fixup_kernel<<<seg_ctas, kSpecWarpsPerCta * 32, 0, st>>>(p);
      FF_CUDA_TRY(cudaGetLastError());
      head_track_resolve_kernel<<<1, 256, 0, st>>>(p);
    } else {
      head_track_commit_kernel<<<1, 256, 0, st>>>(p);
    }
    FF_CUDA_TRY(cudaGetLastError());
    return FF_OK;
  }
  head_track_generic_kernel<<<1, kHeadThreads, 0, st>>>(p);      // sequential walk (cross-check)
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

}  // namespace ff

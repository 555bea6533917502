// The reference's per-frame operator seam on the GPU: what a caller of
// scripts/process_videos.py gets when it uses the frame-level functions instead of the driver loop.
//
// frame_op_kernel      - subtract_scalar_background (:670-674), subtract_prior_frame (:677-701),
//                        three_frame_difference (:704-740): float64 element-wise, one IEEE
//                        operation per NumPy operation, so every value is bit-identical.
// frame_count_kernel   - np.sum(frame > noise_threshold) inside is_empty_frame (:759).
// head_images_kernel   - every FULL-FRAME intermediate FlameDetector.detect returns for
//                        visualisation (:380-413): background-subtracted frame, thresholded
//                        difference, k x k grey opening, Gaussian blur, Sobel(axis=1),
//                        np.gradient(axis=1).  Same float64 operation order as head_band_kernel
//                        (scipy NI_Correlate1D: centre tap, then symmetric pairs outermost ->
//                        innermost, no FMA), evaluated tile by tile on the reflect-extended
//                        neighbourhood - every stage is symmetric, so the stages commute with
//                        scipy's mode='reflect' extension.
#include "ff_common.cuh"

namespace ff {
namespace {

constexpr int kOpThreads = 256;

template <class T>
__device__ __forceinline__ double as_f64(T v) { return (double)v; }

// OP 0: max(a - s, 0) with NumPy's `x[x < 0] = 0` (NaN stays NaN)
// OP 1: d = a - b; d[d < s] = 0
// OP 2: m = minimum(|b - a|, |c - b|); m[m < s] = 0      (a = prev, b = curr, c = next)
template <class T, int OP>
__global__ void __launch_bounds__(kOpThreads) frame_op_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                              const T* __restrict__ c, int64_t n, double s,
                                                              double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kOpThreads;
  for (int64_t i = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; i < n; i += stride) {
    double v;
    if (OP == 0) {
      v = __dsub_rn(as_f64(a[i]), s);
      if (v < 0.0) v = 0.0;
    } else if (OP == 1) {
      v = __dsub_rn(as_f64(a[i]), as_f64(b[i]));
      if (v < s) v = 0.0;
    } else {
      const double d1 = fabs(__dsub_rn(as_f64(b[i]), as_f64(a[i])));
      const double d2 = fabs(__dsub_rn(as_f64(c[i]), as_f64(b[i])));
      // np.minimum propagates NaN; fmin would drop it
      v = (d1 != d1) ? d1 : ((d2 != d2) ? d2 : (d1 < d2 ? d1 : d2));
      if (v < s) v = 0.0;
    }
    out[i] = v;
  }
}

template <class T>
__global__ void __launch_bounds__(kOpThreads) frame_count_kernel(const T* __restrict__ a, int64_t n, double thr,
                                                                 unsigned long long* __restrict__ count) {
  __shared__ int s_part[kOpThreads / 32];
  int cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * kOpThreads;
  for (int64_t i = (int64_t)blockIdx.x * kOpThreads + threadIdx.x; i < n; i += stride) cnt += as_f64(a[i]) > thr;
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long total = 0;
#pragma unroll
    for (int k = 0; k < kOpThreads / 32; ++k) total += (unsigned long long)s_part[k];
    if (total) atomicAdd(count, total);
  }
}

int op_grid(int64_t n) {
  int64_t g = (n + kOpThreads - 1) / kOpThreads;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int OP>
int launch_frame_op(const void* a, const void* b, const void* c, int px_type, int64_t n, double s, double* out,
                    cudaStream_t st) {
  const int g = op_grid(n);
  switch (px_type) {
    case FF_PX_U8:
      frame_op_kernel<uint8_t, OP><<<g, kOpThreads, 0, st>>>(static_cast<const uint8_t*>(a), static_cast<const uint8_t*>(b),
                                                             static_cast<const uint8_t*>(c), n, s, out);
      break;
    case FF_PX_U16:
      frame_op_kernel<uint16_t, OP><<<g, kOpThreads, 0, st>>>(static_cast<const uint16_t*>(a), static_cast<const uint16_t*>(b),
                                                              static_cast<const uint16_t*>(c), n, s, out);
      break;
    case FF_PX_F64:
      frame_op_kernel<double, OP><<<g, kOpThreads, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b),
                                                            static_cast<const double*>(c), n, s, out);
      break;
    default:
      return FF_ERR_UNSUPPORTED;
  }
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

// ---- full-frame detector images ---------------------------------------------------------------
constexpr int kImgMaxRadius = 8;
constexpr int kImgMaxMorphHalf = 3;       // morphology_kernel_size 1, 3, 5, 7
constexpr int kImgTileH = 16;
constexpr int kImgTileW = 64;
constexpr int kImgThreads = 256;

struct HeadImagesParams {
  const uint8_t* frames;
  const uint8_t* halo;
  int64_t frame_bytes;
  int n_frames;
  int height, width;
  int bg, bg_halo;
  int diff_thr;
  int morph_half;
  const uint8_t* skip;
  double w[2 * kImgMaxRadius + 1];
  int radius;
  double* sub;        // each [n][H][W], nullable
  double* diff;
  double* opened;
  double* blurred;
  double* sobel;
  double* gradient;
  uint8_t* state;     // [n]: 0 skipped, 1 all images valid, 2 no prior frame (only `sub` is meaningful)
};

__device__ __forceinline__ int reflect_px(int i, int n) {
  if ((unsigned)i < (unsigned)n) return i;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

template <int BITS>
__global__ void __launch_bounds__(kImgThreads) head_images_kernel(const HeadImagesParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = p.width, H = p.height;
  const int R = p.radius, mh = p.morph_half;
  const int HALO = R + 2 * mh + 1;          // opening (2*mh) + Gaussian (R) + Sobel / gradient (1)
  const int NB = kImgTileH + 2 * HALO;      // rows of the integer band
  const int LW = kImgTileW + 2 * HALO;      // columns of the band
  const int GR = kImgTileH + 2;             // rows of the float64 stages (tile +- 1 for the Sobel smoothing)
  uint16_t* bufA = reinterpret_cast<uint16_t*>(smem);                     // [NB][LW]
  uint16_t* bufB = bufA + NB * LW;                                         // [NB][LW]
  const size_t u16_bytes = ((size_t)2 * NB * LW * sizeof(uint16_t) + 15) & ~(size_t)15;
  double* g0 = reinterpret_cast<double*>(smem + u16_bytes);                // [GR][LW]
  double* bl = g0 + GR * LW;                                               // [GR][LW]

  const int tiles_x = (W + kImgTileW - 1) / kImgTileW;
  const int tiles_y = (H + kImgTileH - 1) / kImgTileH;
  const int64_t per_frame = (int64_t)tiles_x * tiles_y;
  const int64_t n_work = per_frame * p.n_frames;
  const int64_t frame_px = (int64_t)H * W;

  for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int f = (int)(work / per_frame);
    const int t = (int)(work - (int64_t)f * per_frame);
    const int y0 = (t / tiles_x) * kImgTileH, x0 = (t % tiles_x) * kImgTileW;
    const int th = min(kImgTileH, H - y0), tw = min(kImgTileW, W - x0);
    const bool skipped = p.skip != nullptr && p.skip[f] != 0;
    int hf = f - 1;
    if (p.skip != nullptr)
      while (hf >= 0 && p.skip[hf]) --hf;
    const uint8_t* prior = hf >= 0 ? p.frames + (int64_t)hf * p.frame_bytes : p.halo;
    const int bg_prior = hf >= 0 ? p.bg : p.bg_halo;
    const bool full = !skipped && prior != nullptr;
    if (t == 0 && tid == 0 && p.state != nullptr) p.state[f] = skipped ? 0 : (full ? 1 : 2);
    const uint8_t* cur = p.frames + (int64_t)f * p.frame_bytes;
    const int64_t obase = (int64_t)f * frame_px;

    if (!full) {
      // no difference image (:386-393): the background-subtracted frame is all detect() produces
      for (int e = tid; e < th * kImgTileW; e += kImgThreads) {
        const int ty = e / kImgTileW, tx = e - ty * kImgTileW;
        if (tx >= tw) continue;
        const int64_t q = (int64_t)(y0 + ty) * W + x0 + tx;
        if (p.sub != nullptr) p.sub[obase + q] = skipped ? 0.0 : (double)max(load_px_generic<BITS>(cur, q) - p.bg, 0);
        if (p.diff != nullptr) p.diff[obase + q] = 0.0;
        if (p.opened != nullptr) p.opened[obase + q] = 0.0;
        if (p.blurred != nullptr) p.blurred[obase + q] = 0.0;
        if (p.sobel != nullptr) p.sobel[obase + q] = 0.0;
        if (p.gradient != nullptr) p.gradient[obase + q] = 0.0;
      }
      continue;
    }

    // ---- D: thresholded difference on the reflect-extended band (:397-399) ------------------------
    for (int i = warp; i < NB; i += kImgThreads / 32) {
      const int gy = y0 - HALO + i;
      const int64_t rowq = (int64_t)reflect_px(gy, H) * W;
      const bool row_in = i >= HALO && i < HALO + th;
      for (int j = lane; j < LW; j += 32) {
        const int64_t q = rowq + reflect_px(x0 - HALO + j, W);
        const int s = max(load_px_generic<BITS>(cur, q) - p.bg, 0);
        int d = s - max(load_px_generic<BITS>(prior, q) - bg_prior, 0);
        if (d < p.diff_thr) d = 0;
        bufA[i * LW + j] = (uint16_t)d;      // diff_thr >= 0 (launcher): 0 <= d <= 65535
        if (row_in && j >= HALO && j < HALO + tw) {
          if (p.sub != nullptr) p.sub[obase + q] = (double)s;
          if (p.diff != nullptr) p.diff[obase + q] = (double)d;
        }
      }
    }
    __syncthreads();
    // ---- E = k x k minimum (grey erosion), valid [mh, NB-mh) x [mh, LW-mh) ------------------------
    for (int i = warp; i < NB; i += kImgThreads / 32) {
      const bool row_ok = i >= mh && i < NB - mh;
      for (int j = lane; j < LW; j += 32) {
        unsigned m = 0;
        if (row_ok && j >= mh && j < LW - mh) {
          m = 0xFFFFu;
          for (int di = -mh; di <= mh; ++di)
            for (int dj = -mh; dj <= mh; ++dj) m = min(m, (unsigned)bufA[(i + di) * LW + j + dj]);
        }
        bufB[i * LW + j] = (uint16_t)m;
      }
    }
    __syncthreads();
    // ---- NR = k x k maximum of E (grey dilation) = the opening (:404), valid [2mh, ..-2mh) ---------
    for (int i = warp; i < NB; i += kImgThreads / 32) {
      const bool row_ok = i >= 2 * mh && i < NB - 2 * mh;
      const bool row_in = i >= HALO && i < HALO + th;
      for (int j = lane; j < LW; j += 32) {
        unsigned m = 0;
        if (row_ok && j >= 2 * mh && j < LW - 2 * mh) {
          for (int di = -mh; di <= mh; ++di)
            for (int dj = -mh; dj <= mh; ++dj) m = max(m, (unsigned)bufB[(i + di) * LW + j + dj]);
        }
        bufA[i * LW + j] = (uint16_t)m;
        if (p.opened != nullptr && row_in && j >= HALO && j < HALO + tw)
          p.opened[obase + (int64_t)(y0 + i - HALO) * W + x0 + j - HALO] = (double)m;
      }
    }
    __syncthreads();
    // ---- G0 = Gaussian along axis 0 for band rows HALO-1 .. HALO+TH (:407) --------------------------
    for (int e = tid; e < GR * LW; e += kImgThreads) {
      const int b = e / LW, j = e - b * LW;
      double tmp = 0.0;
      if (j >= 2 * mh && j < LW - 2 * mh) {
        const int row = HALO - 1 + b;
        tmp = __dmul_rn((double)bufA[row * LW + j], p.w[R]);
        for (int jj = -R; jj < 0; ++jj) {
          const double pair = __dadd_rn((double)bufA[(row + jj) * LW + j], (double)bufA[(row - jj) * LW + j]);
          tmp = __dadd_rn(tmp, __dmul_rn(pair, p.w[R + jj]));
        }
      }
      g0[e] = tmp;
    }
    __syncthreads();
    // ---- BL = Gaussian along axis 1, valid columns [HALO-1, LW-HALO+1) --------------------------------
    for (int e = tid; e < GR * LW; e += kImgThreads) {
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
    // ---- outputs: blurred, Sobel(axis=1) (:410), np.gradient(axis=1) (:413) ---------------------------
    for (int e = tid; e < th * kImgTileW; e += kImgThreads) {
      const int ty = e / kImgTileW, tx = e - ty * kImgTileW;
      if (tx >= tw) continue;
      const int b = ty + 1, j = HALO + tx;
      const int x = x0 + tx;
      const int64_t o = obase + (int64_t)(y0 + ty) * W + x;
      const double* v = bl + b * LW;
      if (p.blurred != nullptr) p.blurred[o] = v[j];
      if (p.sobel != nullptr) {
        double s3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double* r = bl + (b - 1 + k) * LW;
          // correlate1d([-1,0,1]): tmp = in[0]*0; tmp += (in[-1] - in[+1]) * (-1)
          s3[k] = __dadd_rn(__dmul_rn(r[j], 0.0), __dmul_rn(__dsub_rn(r[j - 1], r[j + 1]), -1.0));
        }
        // correlate1d([1,2,1]) along axis 0: tmp = in[0]*2; tmp += (in[-1] + in[+1]) * 1
        p.sobel[o] = __dadd_rn(__dmul_rn(s3[1], 2.0), __dmul_rn(__dadd_rn(s3[0], s3[2]), 1.0));
      }
      if (p.gradient != nullptr) {
        double g;
        if (x == 0) g = __ddiv_rn(__dsub_rn(v[j + 1], v[j]), 1.0);
        else if (x == W - 1) g = __ddiv_rn(__dsub_rn(v[j], v[j - 1]), 1.0);
        else g = __ddiv_rn(__dsub_rn(v[j + 1], v[j - 1]), 2.0);
        p.gradient[o] = g;
      }
    }
    __syncthreads();      // the next work item reuses the shared-memory band
  }
}

}  // namespace

int frame_subtract_background_impl(const void* image, int px_type, int64_t n_px, double background, double* out,
                                   cudaStream_t st) {
  if (image == nullptr || out == nullptr || n_px <= 0) return FF_ERR_INVALID;
  return launch_frame_op<0>(image, nullptr, nullptr, px_type, n_px, background, out, st);
}

int frame_difference_impl(const void* current, const void* prior, int px_type, int64_t n_px, double threshold,
                          double* out, cudaStream_t st) {
  if (current == nullptr || prior == nullptr || out == nullptr || n_px <= 0) return FF_ERR_INVALID;
  return launch_frame_op<1>(current, prior, nullptr, px_type, n_px, threshold, out, st);
}

int frame_three_difference_impl(const void* prev, const void* curr, const void* next, int px_type, int64_t n_px,
                                double threshold, double* out, cudaStream_t st) {
  if (prev == nullptr || curr == nullptr || next == nullptr || out == nullptr || n_px <= 0) return FF_ERR_INVALID;
  return launch_frame_op<2>(prev, curr, next, px_type, n_px, threshold, out, st);
}

int frame_count_above_impl(const void* frame, int px_type, int64_t n_px, double threshold, int64_t* count,
                           cudaStream_t st) {
  if (frame == nullptr || count == nullptr || n_px <= 0) return FF_ERR_INVALID;
  if (px_type != FF_PX_U8 && px_type != FF_PX_U16 && px_type != FF_PX_F64) return FF_ERR_UNSUPPORTED;
  FF_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int64_t), st));
  auto* c = reinterpret_cast<unsigned long long*>(count);
  const int g = op_grid(n_px);
  switch (px_type) {
    case FF_PX_U8: frame_count_kernel<uint8_t><<<g, kOpThreads, 0, st>>>(static_cast<const uint8_t*>(frame), n_px, threshold, c); break;
    case FF_PX_U16: frame_count_kernel<uint16_t><<<g, kOpThreads, 0, st>>>(static_cast<const uint16_t*>(frame), n_px, threshold, c); break;
    default: frame_count_kernel<double><<<g, kOpThreads, 0, st>>>(static_cast<const double*>(frame), n_px, threshold, c); break;
  }
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

int head_images_impl(const void* frames, const void* halo, int64_t n_frames, int height, int width, int bits,
                     int32_t bg, int32_t bg_halo, int32_t diff_thr, int morphology_size,
                     const double* gauss_weights_host, int radius, const uint8_t* skip, double* sub_out,
                     double* diff_out, double* opened_out, double* blurred_out, double* sobel_out,
                     double* gradient_out, uint8_t* state_out, cudaStream_t st) {
  if (frames == nullptr || gauss_weights_host == nullptr) return FF_ERR_INVALID;
  if (n_frames <= 0 || height <= 0 || width < 2 || n_frames > 0x7FFFFFFF) return FF_ERR_INVALID;
  if (bits != 8 && bits != 12 && bits != 16) return FF_ERR_UNSUPPORTED;
  if (radius < 0 || radius > kImgMaxRadius) return FF_ERR_UNSUPPORTED;
  if (morphology_size < 1 || (morphology_size & 1) == 0 || morphology_size > 2 * kImgMaxMorphHalf + 1)
    return FF_ERR_UNSUPPORTED;                        // odd sizes only: symmetric windows
  if (diff_thr < 0 || bg < 0 || bg_halo < 0) return FF_ERR_UNSUPPORTED;   // the band is held as uint16
  const int64_t px = (int64_t)height * width;
  if (bits == 12 && (px & 1)) return FF_ERR_UNSUPPORTED;

  HeadImagesParams p{};
  p.frames = static_cast<const uint8_t*>(frames);
  p.halo = static_cast<const uint8_t*>(halo);
  p.frame_bytes = frame_bytes_of(px, bits);
  p.n_frames = (int)n_frames;
  p.height = height;
  p.width = width;
  p.bg = bg;
  p.bg_halo = bg_halo;
  p.diff_thr = diff_thr;
  p.morph_half = (morphology_size - 1) / 2;
  p.skip = skip;
  p.radius = radius;
  for (int i = 0; i < 2 * radius + 1; ++i) p.w[i] = gauss_weights_host[i];
  p.sub = sub_out;
  p.diff = diff_out;
  p.opened = opened_out;
  p.blurred = blurred_out;
  p.sobel = sobel_out;
  p.gradient = gradient_out;
  p.state = state_out;

  const int halo_px = radius + 2 * p.morph_half + 1;
  const int nb = kImgTileH + 2 * halo_px, lw = kImgTileW + 2 * halo_px;
  const size_t smem = (((size_t)2 * nb * lw * sizeof(uint16_t) + 15) & ~(size_t)15) +
                      (size_t)2 * (kImgTileH + 2) * lw * sizeof(double);
  const int64_t tiles = (int64_t)((width + kImgTileW - 1) / kImgTileW) * ((height + kImgTileH - 1) / kImgTileH);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) FF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kImgThreads, smem));
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (int64_t)sms * (occ > 0 ? occ : 1);
    if (grid > tiles * n_frames) grid = tiles * n_frames;
    kern<<<(unsigned)grid, kImgThreads, smem, st>>>(p);
    FF_CUDA_TRY(cudaGetLastError());
    return FF_OK;
  };
  switch (bits) {
    case 8: return launch(head_images_kernel<8>);
    case 12: return launch(head_images_kernel<12>);
    default: return launch(head_images_kernel<16>);
  }
}

}  // namespace ff

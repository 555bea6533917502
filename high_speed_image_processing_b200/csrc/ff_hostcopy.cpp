// Host copy into the pinned bounce buffers of the streaming driver (ff_api.cu, CopyPool).
//
// The destination is only ever read by the DMA engine, so it is written with non-temporal stores:
// no read-for-ownership of the destination lines and no cache pollution - per byte copied the memory
// system moves source read + destination write (+ the DMA's read) instead of four transfers.
// glibc's memcpy switches to streaming stores only above a size threshold that the 2-MiB pieces of
// the copy threads stay under.  FF_HOST_COPY_NT=0 selects plain memcpy.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ff {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head) {
    const size_t h = head < n ? head : n;
    std::memcpy(dst, src, h);
    dst += h;
    src += h;
    n -= h;
  }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) std::memcpy(dst + i, src + i, n - i);
}
#endif

// 0 = memcpy, 1 = AVX2 streaming stores; decided once.
static int copy_mode() {
  static const int mode = [] {
    const char* e = std::getenv("FF_HOST_COPY_NT");
    if (e != nullptr && e[0] == '0') return 0;
#if defined(__x86_64__)
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
  }();
  return mode;
}

void stream_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__x86_64__)
  if (copy_mode() == 1 && n >= 4096) {
    copy_nt_avx2(dst, src, n);
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

}  // namespace ff

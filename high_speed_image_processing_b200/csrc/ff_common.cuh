// Shared device/host helpers for libflamefront (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/flamefront.h"

namespace ff {

// ---- error plumbing --------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);

#define FF_CUDA_TRY(expr)                                        \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) {                                     \
      ::ff::set_cuda_error(_e, #expr);                           \
      return FF_ERR_CUDA;                                        \
    }                                                            \
  } while (0)

// ---- launch configuration computed once per device (thread-safe; entry points may be called
// from several host threads, one per GPU) ----------------------------------------------------
struct PerDeviceInt {
  std::mutex m;
  bool have[64] = {};
  int value[64] = {};
  // init(int* v) -> FF status; runs once per device under the lock.
  template <class Init>
  int get(Init&& init, int* out) {
    int dev = 0;
    FF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FF_ERR_INVALID;
    std::lock_guard<std::mutex> g(m);
    if (!have[dev]) {
      const int rc = init(&value[dev]);
      if (rc != FF_OK) return rc;
      have[dev] = true;
    }
    *out = value[dev];
    return FF_OK;
  }
};

// ---- tiling shared by ff_partial_len / ff_stream_frames / ff_detect ---------------------
// A "group" is 8 consecutive pixels = `bits` bytes of the packed stream.  A tile is
// kThreads*K groups; the streaming kernel assigns one CTA to (tile, frame-chunk).
constexpr int kThreads = 256;
constexpr int kGroupPx = 8;

constexpr int kWarpsPerCta = kThreads / 32;

struct Tiling {
  int k;                   // groups per thread per tile (1 or 4)
  int tile_px;             // kThreads * k * 8
  int tiles_per_frame;     // ceil(P / tile_px)
  int partials_per_frame;  // int32 above-noise counts the streaming kernel emits per frame:
                           // one per (tile, warp) on the TMA path, one per frame otherwise
  bool fast;               // TMA-staged path usable (P % 32 == 0)
};

inline Tiling choose_tiling(int64_t px_per_frame) {
  Tiling t;
  t.fast = (px_per_frame % 32) == 0 && px_per_frame > 0;
  t.k = (px_per_frame >= 16 * kThreads * 4 * kGroupPx) ? 4 : 1;   // >= 16 big tiles -> K=4
  t.tile_px = kThreads * t.k * kGroupPx;
  if (!t.fast) {        // generic kernel accumulates one count per frame
    t.k = 0;
    t.tile_px = 0;
    t.tiles_per_frame = 1;
    t.partials_per_frame = 1;
    return t;
  }
  t.tiles_per_frame = (int)((px_per_frame + t.tile_px - 1) / t.tile_px);
  t.partials_per_frame = t.tiles_per_frame * kWarpsPerCta;
  return t;
}

inline int64_t frame_bytes_of(int64_t px, int bits) { return px * bits / 8; }

// ---- range-sharded runs: what the compute kernels need to know about the peer exchange ---------
// (filled by ff_exchange_begin; see csrc/ff_exchange.cu for the protocol)
constexpr int kMaxRanks = 64;
struct PeerTable {                    // lives in device memory, owned by the ff_exchange
  int32_t* base[kMaxRanks];           // every rank's exchange buffer (own entry = local memory)
  int64_t flags_off, acks_off, exit_off, status_off;   // int32 offsets inside such a buffer
};
struct RangeHooks {                   // passed to kernels by value (24 bytes)
  const PeerTable* table;             // nullptr: single-GPU run, every hook is a no-op
  int32_t epoch, world, rank;
  int32_t flags;                      // FF_HOOK_WAIT: prep_kernel waits for the peers' acks of epoch-2;
                                      // FF_HOOK_PUBLISH: the last CTA of the range publishes the block.
                                      // Exit frames are propagated to the peers' exit words in any case.
  long long spin_limit;               // clock64 ticks before a wait gives up and raises the status word
};
inline RangeHooks no_hooks() {
  RangeHooks h{};
  return h;
}
inline RangeHooks hooks_from(const ff_range_hooks* h) {
  RangeHooks r{};
  if (h != nullptr && h->table_dev != nullptr) {
    r.table = static_cast<const PeerTable*>(h->table_dev);
    r.epoch = h->epoch;
    r.world = h->world;
    r.rank = h->rank;
    r.flags = h->flags;
    r.spin_limit = h->spin_limit;
  }
  return r;
}

// ---- workspace of ff_process_range (zero-filled by its owner once; every kernel that uses a word
// leaves it zero again) ---------------------------------------------------------------------------
constexpr int kPrepMaxCtas = 64;
struct RangeWorkspace {
  unsigned int prep_ticket;           // CTAs of prep_kernel that delivered their partial maximum
  unsigned int tail_ticket;           // CTAs of the fused / detect kernel that are done
  unsigned int pad[2];
  int32_t prep_max[kPrepMaxCtas];     // per-CTA maxima of frame 0
  // followed by one 64-bit word per frame, {arrivals:32 | above-noise count:32}, each on its OWN 128-byte
  // line: all CTAs sweep through the clip together, so the words of neighbouring frames are hit at the same
  // time - packed 16 to a line they serialise in one L2 slice (measured: C3 range kernel 1.79 -> 1.2 ms)
};
constexpr int64_t kWorkspaceHeader = 512;
constexpr int64_t kArriveStride = 16;      // in 64-bit words
inline int64_t range_workspace_bytes(int64_t n_frames) {
  return kWorkspaceHeader + 8 * kArriveStride * (n_frames > 0 ? n_frames : 0);
}

#if defined(__CUDACC__)
// ---- mbarrier + bulk async copy (TMA 1-D) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// L2 eviction policy for read-once streams.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- 12-bit decode: 12 bytes (3 LE words) -> 8 pixels -------------------------------------
// bytes b0..b11; pair j = (b[3j], b[3j+1], b[3j+2]) -> hi = b0<<4 | b1>>4, lo = (b1&15)<<8 | b2.
__device__ __forceinline__ void decode12x8(uint32_t w0, uint32_t w1, uint32_t w2, int (&v)[8]) {
  const uint32_t t0 = __byte_perm(w0, 0u, 0x4012);  // 0  b0 b1 b2
  const uint32_t t1 = __byte_perm(w0, w1, 0x3345);  // b3 b3 b4 b5 (top byte masked below)
  const uint32_t t2 = __byte_perm(w1, w2, 0x2234);  // b6 b6 b7 b8
  const uint32_t t3 = __byte_perm(w2, 0u, 0x4123);  // 0  b9 b10 b11
  v[0] = (int)(t0 >> 12);
  v[1] = (int)(t0 & 0xFFFu);
  v[2] = (int)((t1 >> 12) & 0xFFFu);
  v[3] = (int)(t1 & 0xFFFu);
  v[4] = (int)((t2 >> 12) & 0xFFFu);
  v[5] = (int)(t2 & 0xFFFu);
  v[6] = (int)(t3 >> 12);
  v[7] = (int)(t3 & 0xFFFu);
}

// 12 bytes -> four 16x2 words, pixel 2j in the LOW half and pixel 2j+1 in the HIGH half of x[j]
// (the byte order of a little-endian uint16 pair, so x[j] can be stored as is).  Each triple is
// permuted to (b1 b2 | b0 b1): the high half then holds lo-pixel | junk nibble, the low half
// hi-pixel << 4, and one shift + mask pair cleans both lanes at once.
__device__ __forceinline__ void decode12x8_16x2(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&x)[4]) {
  const uint32_t p0 = __byte_perm(w0, 0u, 0x1201);
  const uint32_t p1 = __byte_perm(w0, w1, 0x4534);
  const uint32_t p2 = __byte_perm(w1, w2, 0x3423);
  const uint32_t p3 = __byte_perm(w2, 0u, 0x2312);
  x[0] = ((p0 >> 4) & 0x00000FFFu) | (p0 & 0x0FFF0000u);
  x[1] = ((p1 >> 4) & 0x00000FFFu) | (p1 & 0x0FFF0000u);
  x[2] = ((p2 >> 4) & 0x00000FFFu) | (p2 & 0x0FFF0000u);
  x[3] = ((p3 >> 4) & 0x00000FFFu) | (p3 & 0x0FFF0000u);
}

// cnt += (x > k) for unsigned x, k, written as carry-out + add-with-carry (2 SASS instructions
// instead of the compare / add / select triple the compiler emits for the C expression):
// x > k  <=>  x + ~k carries out of 32 bits.  Callers pass nk = ~k.  (add.cc, not sub.cc: the
// carry flag after a PTX subtraction is the hardware's NOT-borrow, which must not feed addc.)
__device__ __forceinline__ void add_gt(int& cnt, uint32_t x, uint32_t nk) {
  asm("{\n.reg .u32 d;\nadd.cc.u32 d, %1, %2;\naddc.u32 %0, %0, 0;\n}" : "+r"(cnt) : "r"(x), "r"(nk));
}

// Count how many of the 8 packed-12 pixels in 3 LE words exceed c (0 <= c <= 4095) WITHOUT
// extracting them: each byte triple is permuted to the top 24 bits of a word, T = hi<<20 | lo<<8 | g
// (g = don't-care low byte).  hi > c <=> T > (c<<20 | 0xFFFFF);  lo > c <=> (T & 0xFFFFF) > (c<<8 | 0xFF):
// whatever sits below a field is dominated by the all-ones padding of the bound.
// nk_hi / nk_lo are the complemented bounds (see add_gt).
__device__ __forceinline__ void count12x8(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t nk_hi, uint32_t nk_lo,
                                          int& cnt) {
  const uint32_t t0 = __byte_perm(w0, 0u, 0x0124);  // b0 b1 b2 0
  const uint32_t t1 = __byte_perm(w0, w1, 0x3455);  // b3 b4 b5 g
  const uint32_t t2 = __byte_perm(w1, w2, 0x2344);  // b6 b7 b8 g
  const uint32_t t3 = __byte_perm(w2, 0u, 0x1234);  // b9 b10 b11 0
  add_gt(cnt, t0, nk_hi);
  add_gt(cnt, t0 & 0xFFFFFu, nk_lo);
  add_gt(cnt, t1, nk_hi);
  add_gt(cnt, t1 & 0xFFFFFu, nk_lo);
  add_gt(cnt, t2, nk_hi);
  add_gt(cnt, t2 & 0xFFFFFu, nk_lo);
  add_gt(cnt, t3, nk_hi);
  add_gt(cnt, t3 & 0xFFFFFu, nk_lo);
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor on the stream still runs; griddep_wait() blocks until that predecessor has
// completed and its memory is visible (no-op without the attribute).  The predecessor allows the
// early start with griddep_launch_dependents().
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- system-scope flags in peer memory (range exchange, csrc/ff_exchange.cu) --------------------
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t ld_relaxed_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int32_t* p, int32_t v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_min_sys(int32_t* p, int32_t v) {
  asm volatile("red.relaxed.sys.global.min.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pixel out of a packed buffer whose first byte holds flat pixel 0 (any bit depth).
template <int BITS>
__device__ __forceinline__ int load_px_generic(const uint8_t* __restrict__ base, int64_t q) {
  if (BITS == 8) return base[q];
  if (BITS == 16) return (int)base[2 * q] | ((int)base[2 * q + 1] << 8);
  const uint8_t* t = base + (q >> 1) * 3;
  return (q & 1) ? (((int)(t[1] & 15) << 8) | (int)t[2]) : (((int)t[0] << 4) | ((int)t[1] >> 4));
}

// Kernel launch with the optional PDL attribute (the kernel may start while its predecessor on the
// stream still runs and must griddep_wait() before it touches that predecessor's outputs).
template <class Kern, class Params>
inline int launch_kernel(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  FF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  return FF_OK;
}
#endif  // __CUDACC__

}  // namespace ff

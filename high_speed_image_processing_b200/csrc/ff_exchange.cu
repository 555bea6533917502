// Stage 4 across GPUs: the one exchange step of a range-sharded clip.
//
// The reference gathers pickled per-rank result lists on rank 0 and sorts them
// (scripts/process_videos.py:1533-1541), and every MPI rank `break`s on its own exit frame
// (:1494).  Here each rank's range kernel writes straight into a RANGE BLOCK
//
//     int32 block[4 + 2*cap] = { first_exit, 0, 0, 0 | pos[cap] | counts[cap] }
//
// and ONE kernel per rank finishes the clip: global exit frame = min over the blocks' headers,
// truncation against it (README.md:145-149), and de-padding of the per-rank arrays into the
// contiguous pos[total] / counts[total] every rank returns.  Two transports feed that kernel:
//
//   gathered   the blocks were all-gathered into one local buffer by the process group
//              (one NCCL all_gather_into_tensor; gloo in the CPU tests)       -> ff_merge_ranges
//   peer       ff_exchange_*: the blocks stay where they were written; every rank maps its peers'
//              exchange buffers over NVLink (CUDA IPC) and the finishing kernel pulls the blocks
//              directly.  No collective and no barrier on the data path; the protocol is four
//              kinds of words in every rank's buffer, all written with system-scope stores:
//
//       flags[r]  epoch of the latest block rank r has COMPLETED.  Written by the last CTA of
//                 r's range kernel (range_tail) - the compute kernel publishes, not the merge.
//       acks[r]   epoch of the latest merge rank r has finished, i.e. r no longer reads my
//                 block of that epoch.  Blocks are double-buffered by epoch parity, so before
//                 epoch e overwrites the block of e-2 the prep kernel waits for acks >= e-2.
//       exit[2]   per epoch parity: min over ALL ranks of the exit frames seen so far in this
//                 clip (red.min.sys from the detecting warps).  The host-streamed path polls it
//                 between chunks and stops uploading frames that lie behind it - the
//                 cross-rank form of the reference's `break` (:1494).  Reset by the merge.
//       status    non-zero when a wait timed out.
//
//   The merge kernel depends on nothing but the flags, so it runs on a side stream while the
//   main stream already streams the next clip: the exchange is off the critical path.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ff_internal.h"

namespace ff {
namespace {

constexpr int kHdr = 4;   // int32 header words per block

struct MergeParams {
  const int32_t* gathered;            // [world][stride], or nullptr for the peer transport
  int64_t stride;                     // int32 elements per block
  int64_t slot_off;                   // peer transport: offset of this epoch's block in every buffer
  int64_t cap;                        // capacity (frames) of a block's pos / counts arrays
  int world;
  int rank;
  int64_t total;
  int32_t* pos_out;
  int32_t* count_out;
  int32_t* first_exit_out;
  // peer transport only
  const PeerTable* table;
  int32_t epoch;
  int publish;                        // this rank's flag is still to be written (block filled by hand)
  unsigned int* ticket;               // CTAs of this kernel that are done (left zero)
  long long spin_limit;
};

// Owner rank and local offset of global frame i under contiguous_range(total, r, world):
// the first total % world ranks hold one extra frame (src/photron/parallel.py:101-113).
__device__ __forceinline__ void owner_of(int64_t i, int64_t base, int64_t extra, int& r, int64_t& off) {
  const int64_t boundary = extra * (base + 1);
  if (i < boundary) {
    r = (int)(i / (base + 1));
    off = i - (int64_t)r * (base + 1);
  } else {
    const int64_t j = i - boundary;
    r = (int)(extra + j / base);
    off = j - (j / base) * base;
  }
}

// Small CTAs (64 threads): on its side stream this kernel runs NEXT TO the next clip's range kernel, whose two
// CTAs per SM leave ~5 K registers free - a CTA that does not fit there would wait for a range CTA to leave
// and, worse, a spinning merge CTA that got in first would keep a range CTA out (measured at N=2 with
// 256-thread CTAs: +90 us per step).
// (64 registers by launch bound: 2 warps x 2048 registers fit the 5.6 K the range kernel leaves free.)
constexpr int kMergeThreads = 64;
static_assert(kMergeThreads >= kMaxRanks, "one thread per rank polls a flag");

template <bool PEER>
__global__ void __launch_bounds__(kMergeThreads, 16) merge_ranges_kernel(const MergeParams p) {
  __shared__ int s_fe;
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const PeerTable* t = p.table;
  if (tid == 0) s_ok = 1;
  __syncthreads();
  if (PEER) {
    if (p.publish && blockIdx.x == 0 && tid < p.world) {
      __threadfence_system();
      st_release_sys(t->base[tid] + t->flags_off + p.rank, p.epoch);
    }
    // Arrive: wait until every rank has published this epoch to me.
    if (tid < p.world) {
      const int32_t* flag = t->base[p.rank] + t->flags_off + tid;
      const long long t0 = clock64();
      while (ld_acquire_sys(flag) < p.epoch) {
        if (clock64() - t0 > p.spin_limit) {
          s_ok = 0;
          atomicExch(t->base[p.rank] + t->status_off, 1 + tid);
          break;
        }
      }
    }
    __syncthreads();
  }
  const bool ok = s_ok != 0;
  if (tid == 0) s_fe = FF_NO_EXIT;
  __syncthreads();
  if (ok && tid < p.world) {
    const int32_t* b = PEER ? t->base[tid] + p.slot_off : p.gathered + (int64_t)tid * p.stride;
    atomicMin(&s_fe, PEER ? __ldcv(b) : b[0]);
  }
  __syncthreads();
  const int64_t fe = s_fe;
  if (blockIdx.x == 0 && tid == 0) *p.first_exit_out = s_fe;

  const int64_t base = p.total / p.world, extra = p.total % p.world;
  // Four elements per thread and round, all loads issued before the first store: the peers' blocks are read over
  // NVLink - or over PCIe where a box gives its GPUs no NVLink path - and a thread that waits for one remote load at a
  // time makes the merge a chain of round trips (measured on such a box at N = 2: merges of ~0.7 ms, longer than the
  // step they are meant to hide behind).
  constexpr int kUnroll = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + tid; i0 < p.total; i0 += kUnroll * stride) {
    int32_t v[kUnroll], c[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      v[u] = FF_POS_NONE;
      c[u] = 0;
      if (i < p.total && ok) {        // (a peer that never arrived: defined outputs, the status word tells)
        int r;
        int64_t off;
        owner_of(i, base, extra, r, off);
        const int32_t* b = PEER ? t->base[r] + p.slot_off : p.gathered + (int64_t)r * p.stride;
        const int32_t* src_pos = b + kHdr + off;
        v[u] = PEER ? __ldcv(src_pos) : *src_pos;
        if (p.count_out != nullptr) c[u] = PEER ? __ldcv(src_pos + p.cap) : src_pos[p.cap];
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < p.total) {
        p.pos_out[i] = (ok && i >= fe) ? FF_POS_DROPPED : v[u];
        if (p.count_out != nullptr) p.count_out[i] = c[u];
      }
    }
  }
  if (PEER) {
    // last CTA: nobody writes this epoch's exit word any more (every rank has published) - reset it
    // for epoch + 2 - and tell the peers that their blocks of this epoch are no longer read here.
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      if (tid == 0) {
        *p.ticket = 0;
        t->base[p.rank][t->exit_off + (p.epoch & 1)] = FF_NO_EXIT;
      }
      __syncthreads();
      if (tid < p.world) {
        __threadfence_system();
        st_release_sys(t->base[tid] + t->acks_off + p.rank, p.epoch);
      }
    }
  }
}

// Block filled by hand (results that came back from somewhere else): wait for the peers' acks and
// reset the header (what prep_kernel does for ff_process_range) / publish the finished block.
__global__ void publish_kernel(const RangeHooks h) {
  __threadfence_system();
  const PeerTable* t = h.table;
  for (int r = threadIdx.x; r < h.world; r += blockDim.x) st_relaxed_sys(t->base[r] + t->flags_off + h.rank, h.epoch);
}

// The small kernels of the exchange run next to a range kernel that configures its SMs for maximum shared
// memory; they ask for the same split so that their CTAs can be placed on those SMs (once per process).
template <class Kern>
void share_sm_with_range_kernel(Kern kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

int merge_grid(int64_t total) {
  int64_t blocks = (total + 1023) / 1024;
  if (blocks > 74) blocks = 74;       // few, small CTAs: this kernel shares the GPU with the next clip's range kernel
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int publish_impl(const RangeHooks& hooks, cudaStream_t st) {
  if (hooks.table == nullptr) return FF_OK;
  static const bool once = (share_sm_with_range_kernel(publish_kernel), true);
  (void)once;
  publish_kernel<<<1, 32, 0, st>>>(hooks);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

}  // namespace ff

using namespace ff;

// ---- peer-memory exchange context ---------------------------------------------------------------
struct ff_exchange {
  int device = 0, rank = 0, world = 1;
  int64_t cap = 0;                 // frames per block
  int64_t stride = 0;              // int32 elements per block
  // one cudaMalloc, mapped by the peers: [block slot 0 | block slot 1 | flags[kMaxRanks] | acks[kMaxRanks] |
  //                                       exit[2] pad | status | ticket ...]
  int32_t* local = nullptr;
  int64_t flags_off = 0, acks_off = 0, exit_off = 0, status_off = 0, ticket_off = 0, alloc_elems = 0;
  int32_t* peer_base[kMaxRanks] = {nullptr};
  bool opened[kMaxRanks] = {false};
  bool peers_ready = false;
  PeerTable* table_dev = nullptr;  // local device memory
  int32_t epoch = 0;               // epoch of the block handed out last
  int32_t* status_host = nullptr;  // pinned readback
  long long spin_limit = 0;
};

static int upload_table(ff_exchange* x) {
  PeerTable t{};
  for (int r = 0; r < x->world; ++r) t.base[r] = x->peer_base[r];
  t.flags_off = x->flags_off;
  t.acks_off = x->acks_off;
  t.exit_off = x->exit_off;
  t.status_off = x->status_off;
  FF_CUDA_TRY(cudaMemcpy(x->table_dev, &t, sizeof(t), cudaMemcpyHostToDevice));
  return FF_OK;
}

static RangeHooks hooks_of(const ff_exchange* x, int flags) {
  RangeHooks h{};
  h.table = x->table_dev;
  h.epoch = x->epoch;
  h.world = x->world;
  h.rank = x->rank;
  h.flags = flags;
  h.spin_limit = x->spin_limit;
  return h;
}

extern "C" {

int ff_merge_ranges(const int32_t* gathered_dev, int world, int64_t block_cap_frames, int64_t total_frames,
                    int32_t* pos_out_dev, int32_t* count_out_dev, int32_t* first_exit_out_dev, void* stream) {
  if (gathered_dev == nullptr || pos_out_dev == nullptr || first_exit_out_dev == nullptr) return FF_ERR_INVALID;
  if (world < 1 || world > kMaxRanks) return FF_ERR_UNSUPPORTED;
  if (total_frames < 0 || block_cap_frames < 0) return FF_ERR_INVALID;
  if ((total_frames + world - 1) / world > block_cap_frames) return FF_ERR_INVALID;
  MergeParams p{};
  p.gathered = gathered_dev;
  p.cap = block_cap_frames;
  p.stride = kHdr + 2 * block_cap_frames;
  p.world = world;
  p.total = total_frames;
  p.pos_out = pos_out_dev;
  p.count_out = count_out_dev;
  p.first_exit_out = first_exit_out_dev;
  merge_ranges_kernel<false><<<merge_grid(total_frames), kMergeThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

int ff_range_block_len(int64_t cap_frames, int64_t* n_elems) {
  if (cap_frames < 0 || n_elems == nullptr) return FF_ERR_INVALID;
  *n_elems = kHdr + 2 * cap_frames;
  return FF_OK;
}

int ff_exchange_create(int device, int rank, int world, int64_t cap_frames, ff_exchange** out) {
  if (out == nullptr) return FF_ERR_INVALID;
  *out = nullptr;
  if (world < 1 || world > kMaxRanks) return FF_ERR_UNSUPPORTED;
  if (rank < 0 || rank >= world || cap_frames < 1) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(device));
  ff_exchange* x = new (std::nothrow) ff_exchange();
  if (x == nullptr) return FF_ERR_INVALID;
  x->device = device;
  x->rank = rank;
  x->world = world;
  x->cap = cap_frames;
  x->stride = (kHdr + 2 * cap_frames + 3) & ~(int64_t)3;      // keep slot 1 16-byte aligned
  x->flags_off = 2 * x->stride;
  x->acks_off = x->flags_off + kMaxRanks;
  x->exit_off = x->acks_off + kMaxRanks;
  x->status_off = x->exit_off + 4;
  x->ticket_off = x->status_off + 4;
  x->alloc_elems = x->ticket_off + 4;
  // A peer that never arrives must not hang the GPU for ever: give up after FF_EXCHANGE_TIMEOUT_S
  // seconds (default 20; ranks may legitimately reach a step seconds apart, e.g. after file I/O).
  const char* env_t = getenv("FF_EXCHANGE_TIMEOUT_S");
  const double timeout_s = env_t ? atof(env_t) : 20.0;
  int khz = 1900000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);       // clock64 ticks at the SM clock
  x->spin_limit = (long long)((timeout_s > 0.01 ? timeout_s : 0.01) * 1e3 * (double)(khz > 0 ? khz : 1900000));
  cudaError_t e = cudaMalloc(&x->local, sizeof(int32_t) * (size_t)x->alloc_elems);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, sizeof(int32_t) * (size_t)x->alloc_elems);
  const int32_t no_exit[2] = {FF_NO_EXIT, FF_NO_EXIT};
  if (e == cudaSuccess) e = cudaMemcpy(x->local + x->exit_off, no_exit, sizeof(no_exit), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&x->table_dev, sizeof(PeerTable));
  if (e == cudaSuccess) e = cudaMallocHost(&x->status_host, sizeof(int32_t));
  if (e != cudaSuccess) {
    set_cuda_error(e, "ff_exchange_create");
    if (x->local) cudaFree(x->local);
    if (x->table_dev) cudaFree(x->table_dev);
    delete x;
    return FF_ERR_CUDA;
  }
  *x->status_host = 0;
  x->peer_base[rank] = x->local;
  if (world == 1) {
    x->peers_ready = true;
    const int rc = upload_table(x);
    if (rc != FF_OK) return rc;
  }
  *out = x;
  return FF_OK;
}

int ff_exchange_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int ff_exchange_get_handle(ff_exchange* x, void* handle_out) {
  if (x == nullptr || handle_out == nullptr) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  FF_CUDA_TRY(cudaIpcGetMemHandle(&h, x->local));
  std::memcpy(handle_out, &h, sizeof(h));
  return FF_OK;
}

int ff_exchange_open_peers(ff_exchange* x, const void* handles) {
  if (x == nullptr || handles == nullptr) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  const char* hb = static_cast<const char*>(handles);
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank || x->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    FF_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer_base[r] = static_cast<int32_t*>(ptr);
    x->opened[r] = true;
  }
  const int rc = upload_table(x);
  if (rc != FF_OK) return rc;
  x->peers_ready = true;
  return FF_OK;
}

int ff_exchange_begin(ff_exchange* x, int32_t** pos_dev, int32_t** count_dev, int32_t** first_exit_dev,
                      ff_range_hooks* hooks_out) {
  if (x == nullptr || pos_dev == nullptr || count_dev == nullptr || first_exit_dev == nullptr) return FF_ERR_INVALID;
  if (!x->peers_ready) return FF_ERR_INVALID;
  x->epoch += 1;
  int32_t* blk = x->local + (int64_t)(x->epoch & 1) * x->stride;
  *first_exit_dev = blk;
  *pos_dev = blk + kHdr;
  *count_dev = blk + kHdr + x->cap;
  if (hooks_out != nullptr) {
    hooks_out->table_dev = x->table_dev;
    hooks_out->epoch = x->epoch;
    hooks_out->world = x->world;
    hooks_out->rank = x->rank;
    hooks_out->flags = FF_HOOK_WAIT | FF_HOOK_PUBLISH;
    hooks_out->spin_limit = x->spin_limit;
    hooks_out->exit_word_dev = x->local + x->exit_off + (x->epoch & 1);
  }
  return FF_OK;
}

int ff_exchange_acquire(ff_exchange* x, void* stream) {
  if (x == nullptr || x->epoch < 1) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  int32_t* blk = x->local + (int64_t)(x->epoch & 1) * x->stride;
  return prep_impl(nullptr, 1, 1, 8, nullptr, nullptr, 0, blk, nullptr, hooks_of(x, FF_HOOK_WAIT),
                   static_cast<cudaStream_t>(stream));
}

int ff_exchange_publish(ff_exchange* x, void* stream) {
  if (x == nullptr || x->epoch < 1) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  return publish_impl(hooks_of(x, FF_HOOK_PUBLISH), static_cast<cudaStream_t>(stream));
}

int ff_exchange_finish(ff_exchange* x, int64_t total_frames, int32_t* pos_out_dev, int32_t* count_out_dev,
                       int32_t* first_exit_out_dev, void* stream) {
  if (x == nullptr || pos_out_dev == nullptr || first_exit_out_dev == nullptr) return FF_ERR_INVALID;
  if (!x->peers_ready) return FF_ERR_INVALID;
  if (x->epoch < 1 || total_frames < 0) return FF_ERR_INVALID;
  if ((total_frames + x->world - 1) / x->world > x->cap) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  MergeParams p{};
  p.gathered = nullptr;
  p.slot_off = (int64_t)(x->epoch & 1) * x->stride;
  p.stride = x->stride;
  p.cap = x->cap;
  p.world = x->world;
  p.rank = x->rank;
  p.total = total_frames;
  p.pos_out = pos_out_dev;
  p.count_out = count_out_dev;
  p.first_exit_out = first_exit_out_dev;
  p.table = x->table_dev;
  p.epoch = x->epoch;
  p.publish = 0;
  p.ticket = reinterpret_cast<unsigned int*>(x->local + x->ticket_off);
  p.spin_limit = x->spin_limit;
  static const bool once = (share_sm_with_range_kernel(merge_ranges_kernel<true>), true);
  (void)once;
  merge_ranges_kernel<true><<<merge_grid(total_frames), kMergeThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

int ff_exchange_status(ff_exchange* x, int32_t* status_out, void* stream) {
  if (x == nullptr || status_out == nullptr) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FF_CUDA_TRY(cudaMemcpyAsync(x->status_host, x->local + x->status_off, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FF_CUDA_TRY(cudaMemsetAsync(x->local + x->status_off, 0, sizeof(int32_t), st));     // reported once
  FF_CUDA_TRY(cudaStreamSynchronize(st));
  *status_out = *x->status_host;
  return FF_OK;
}

int ff_exchange_destroy(ff_exchange* x) {
  if (x == nullptr) return FF_OK;
  cudaSetDevice(x->device);
  for (int r = 0; r < x->world; ++r)
    if (x->opened[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
  if (x->local) cudaFree(x->local);
  if (x->table_dev) cudaFree(x->table_dev);
  if (x->status_host) cudaFreeHost(x->status_host);
  delete x;
  return FF_OK;
}

}  // extern "C"

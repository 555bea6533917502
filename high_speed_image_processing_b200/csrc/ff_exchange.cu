// Stage 4 across GPUs: the one exchange step of a range-sharded clip.
//
// The reference gathers pickled per-rank result lists on rank 0 and sorts them
// (scripts/process_videos.py:1533-1541), and every MPI rank `break`s on its own exit frame
// (:1494).  Here each rank's ff_detect writes straight into a RANGE BLOCK
//
//     int32 block[4 + 2*cap] = { first_exit, n_local, first_frame, epoch | pos[cap] | counts[cap] }
//
// and ONE kernel per rank finishes the clip: global exit frame = min over the blocks' headers,
// truncation against it (README.md:145-149), and de-padding of the per-rank arrays into the
// contiguous pos[total] / counts[total] every rank returns.  Two transports feed that kernel:
//
//   gathered   the blocks were all-gathered into one local buffer by the process group
//              (one NCCL all_gather_into_tensor; gloo in the CPU tests)       -> ff_merge_ranges
//   peer       ff_exchange_*: the blocks stay where ff_detect wrote them; every rank maps its
//              peers' blocks over NVLink (CUDA IPC) and the finishing kernel pulls them
//              directly.  A monotonically increasing epoch written into each peer's flag row
//              (st.release.sys / ld.acquire.sys) replaces the collective's barrier, and blocks
//              are double-buffered so no second barrier is needed before the next step.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ff_common.cuh"

namespace ff {
namespace {

constexpr int kMaxRanks = 64;
constexpr int kHdr = 4;   // int32 header words per block

struct MergeParams {
  const int32_t* gathered;            // [world][stride], or nullptr for the peer transport
  const int32_t* peer[kMaxRanks];     // peer transport: block of rank r for this epoch
  int64_t stride;                     // int32 elements per block
  int64_t cap;                        // capacity (frames) of a block's pos / counts arrays
  int world;
  int rank;
  int64_t total;
  int32_t* pos_out;
  int32_t* count_out;
  int32_t* first_exit_out;
  // peer transport only
  int32_t* my_flags;                  // [world] epochs published to me
  int32_t* peer_flags[kMaxRanks];     // flag rows of the peers (their my_flags)
  int32_t epoch;
  int32_t* status;                    // device word: set non-zero if a peer never arrived
  long long spin_limit;               // clock64 ticks before giving up
};

__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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

template <bool PEER>
__global__ void __launch_bounds__(256) merge_ranges_kernel(const MergeParams p) {
  __shared__ int s_fe;
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  if (PEER) {
    // Publish: my block for this epoch is complete (ff_detect ran earlier on this stream).
    if (blockIdx.x == 0 && tid < p.world) {
      __threadfence_system();
      st_release_sys(p.peer_flags[tid] + p.rank, p.epoch);
    }
    // Arrive: wait until every rank has published this epoch to me.
    if (tid == 0) s_ok = 1;
    __syncthreads();
    if (tid < p.world) {
      const long long t0 = clock64();
      while (ld_acquire_sys(p.my_flags + tid) < p.epoch) {
        if (clock64() - t0 > p.spin_limit) {
          s_ok = 0;
          atomicExch(p.status, 1 + tid);
          break;
        }
      }
    }
    __syncthreads();
    if (!s_ok) return;
  }
  if (tid == 0) s_fe = FF_NO_EXIT;
  __syncthreads();
  if (tid < p.world) {
    const int32_t* b = PEER ? p.peer[tid] : p.gathered + (int64_t)tid * p.stride;
    atomicMin(&s_fe, PEER ? __ldcv(b) : b[0]);
  }
  __syncthreads();
  const int64_t fe = s_fe;
  if (blockIdx.x == 0 && tid == 0) *p.first_exit_out = s_fe;

  const int64_t base = p.total / p.world, extra = p.total % p.world;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int r;
    int64_t off;
    owner_of(i, base, extra, r, off);
    const int32_t* b = PEER ? p.peer[r] : p.gathered + (int64_t)r * p.stride;
    const int32_t* src_pos = b + kHdr + off;
    const int32_t* src_cnt = b + kHdr + p.cap + off;
    const int32_t v = PEER ? __ldcv(src_pos) : *src_pos;
    p.pos_out[i] = i >= fe ? FF_POS_DROPPED : v;
    if (p.count_out != nullptr) p.count_out[i] = PEER ? __ldcv(src_cnt) : *src_cnt;
  }
}

int merge_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148) blocks = 148;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace ff

using namespace ff;

// ---- peer-memory exchange context ---------------------------------------------------------------
struct ff_exchange {
  int device = 0, rank = 0, world = 1;
  int64_t cap = 0;                 // frames per block
  int64_t stride = 0;              // int32 elements per block
  // one cudaMalloc: [block slot 0 | block slot 1 | flags[kMaxRanks] | status | no_exit constant]
  int32_t* local = nullptr;
  int64_t flags_off = 0, status_off = 0, const_off = 0, alloc_elems = 0;
  int32_t* peer_base[kMaxRanks] = {nullptr};
  bool opened[kMaxRanks] = {false};
  bool peers_ready = false;
  int32_t epoch = 0;               // epoch of the block handed out last
  int32_t* status_host = nullptr;  // pinned readback
};

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
  merge_ranges_kernel<false><<<merge_grid(total_frames), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
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
  x->status_off = x->flags_off + kMaxRanks;
  x->const_off = x->status_off + 4;
  x->alloc_elems = x->const_off + 4;
  cudaError_t e = cudaMalloc(&x->local, sizeof(int32_t) * (size_t)x->alloc_elems);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, sizeof(int32_t) * (size_t)x->alloc_elems);
  const int32_t no_exit = FF_NO_EXIT;
  if (e == cudaSuccess) e = cudaMemcpy(x->local + x->const_off, &no_exit, sizeof(no_exit), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMallocHost(&x->status_host, sizeof(int32_t));
  if (e != cudaSuccess) {
    set_cuda_error(e, "ff_exchange_create");
    if (x->local) cudaFree(x->local);
    delete x;
    return FF_ERR_CUDA;
  }
  *x->status_host = 0;
  x->peer_base[rank] = x->local;
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
  x->peers_ready = true;
  return FF_OK;
}

int ff_exchange_begin(ff_exchange* x, int32_t** pos_dev, int32_t** count_dev, int32_t** first_exit_dev,
                      void* stream) {
  if (x == nullptr || pos_dev == nullptr || count_dev == nullptr || first_exit_dev == nullptr) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  x->epoch += 1;
  int32_t* blk = x->local + (int64_t)(x->epoch & 1) * x->stride;
  FF_CUDA_TRY(cudaMemcpyAsync(blk, x->local + x->const_off, sizeof(int32_t), cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)));
  *first_exit_dev = blk;
  *pos_dev = blk + kHdr;
  *count_dev = blk + kHdr + x->cap;
  return FF_OK;
}

int ff_exchange_finish(ff_exchange* x, int64_t total_frames, int32_t* pos_out_dev, int32_t* count_out_dev,
                       int32_t* first_exit_out_dev, void* stream) {
  if (x == nullptr || pos_out_dev == nullptr || first_exit_out_dev == nullptr) return FF_ERR_INVALID;
  if (!x->peers_ready && x->world > 1) return FF_ERR_INVALID;
  if (x->epoch < 1 || total_frames < 0) return FF_ERR_INVALID;
  if ((total_frames + x->world - 1) / x->world > x->cap) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  MergeParams p{};
  p.gathered = nullptr;
  const int64_t slot = (int64_t)(x->epoch & 1) * x->stride;
  for (int r = 0; r < x->world; ++r) {
    p.peer[r] = x->peer_base[r] + slot;
    p.peer_flags[r] = x->peer_base[r] + x->flags_off;
  }
  p.stride = x->stride;
  p.cap = x->cap;
  p.world = x->world;
  p.rank = x->rank;
  p.total = total_frames;
  p.pos_out = pos_out_dev;
  p.count_out = count_out_dev;
  p.first_exit_out = first_exit_out_dev;
  p.my_flags = x->local + x->flags_off;
  p.epoch = x->epoch;
  p.status = x->local + x->status_off;
  // A peer that never arrives must not hang the GPU for ever: give up after FF_EXCHANGE_TIMEOUT_S
  // seconds (default 20; ranks may legitimately reach this step seconds apart, e.g. after file I/O).
  const char* env_t = getenv("FF_EXCHANGE_TIMEOUT_S");
  const double timeout_s = env_t ? atof(env_t) : 20.0;
  p.spin_limit = (long long)((timeout_s > 0.01 ? timeout_s : 0.01) * 1.9e9);
  merge_ranges_kernel<true><<<merge_grid(total_frames), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FF_CUDA_TRY(cudaGetLastError());
  return FF_OK;
}

int ff_exchange_status(ff_exchange* x, int32_t* status_out, void* stream) {
  if (x == nullptr || status_out == nullptr) return FF_ERR_INVALID;
  FF_CUDA_TRY(cudaSetDevice(x->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FF_CUDA_TRY(cudaMemcpyAsync(x->status_host, x->local + x->status_off, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
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
  if (x->status_host) cudaFreeHost(x->status_host);
  delete x;
  return FF_OK;
}

}  // extern "C"

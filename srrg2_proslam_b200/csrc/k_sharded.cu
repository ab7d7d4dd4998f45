// k_sharded.cu -- multi-GPU entry points of the C ABI: one process per GPU, query rows of the exhaustive
// Hamming sweep sharded over the ranks of an NCCL communicator, the train set replicated.
//
// Reference path: CorrespondenceFinderDescriptorBasedBruteforce::compute, the pair loop of
//   .../correspondence_finders/correspondence_finder_descriptor_based_bruteforce_impl.cpp:32-74
// is a per-query-row reduction (best, second best, argmin), so rows shard without any exchange during the sweep; the only
// exchange step is ONE all-gather of 3 x int32 per row afterwards (ncclAllGather on the context's stream, NVLink / NVSwitch
// between the GPUs of a box).  Frames of the stage-1 / stereo pipeline shard the same way with no collective at all
// (pslam_shard_frames gives the ranges).
//
// libnccl is resolved at run time (dlopen of the soname a host process already has loaded, e.g. the one PyTorch ships, else
// the system's): the library itself links neither NCCL nor libcuda, and single-GPU users never touch it.
#include <dlfcn.h>

#include <mutex>

#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

// the few NCCL symbols this file uses, with the ABI of nccl.h (2.x): ncclResult_t / ncclDataType_t are ints, ncclComm_t an
// opaque pointer, ncclUniqueId 128 bytes
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void* id128) = nullptr;
  int (*CommInitRank)(void** comm, int nranks, PslamNcclId id, int rank) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t stream) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_INT32 = 2;  // ncclInt32

NcclApi* nccl_api(pslam_ctx* ctx) {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;  // the copy the process already uses
    if (!api.handle)
      for (const char* n : names)
        if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (api.handle) {
      api.GetUniqueId = (int (*)(void*)) dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, PslamNcclId, int)) dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*)) dlsym(api.handle, "ncclCommDestroy");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t)) dlsym(api.handle, "ncclAllGather");
      api.GetErrorString = (const char* (*) (int) ) dlsym(api.handle, "ncclGetErrorString");
    }
  });
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) {
    pslam_set_error(ctx, PSLAM_E_CUDA, "NCCL is not available in this process (libnccl.so.2 could not be loaded)", cudaSuccess);
    return nullptr;
  }
  return &api;
}

int nccl_fail(pslam_ctx* ctx, NcclApi* api, const char* what, int r) {
  char msg[256];
  snprintf(msg, sizeof msg, "%s: %s", what, api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
  return pslam_set_error(ctx, PSLAM_E_CUDA, msg, cudaSuccess);
}

// gathered: [world][3][per] (rank-major) -> best / second / idx [n] in query order
__global__ void unshard_best2_kernel(const int* __restrict__ gathered, int world, int per, int n, int* __restrict__ best,
                                     int* __restrict__ second, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = i / per, j = i - r * per;
  const int* src = gathered + (size_t) r * 3 * per;
  best[i] = src[j];
  second[i] = src[per + j];
  idx[i] = src[2 * per + j];
}

// ---- exchange step without a collective: peer-to-peer tables ---------------------------------------------------------------
constexpr int P2P_FLAGS = 64;
// after the merge kernel of this stream has completed (its stores to the peers' tables are performed): tell every peer
__global__ void p2p_signal_kernel(int* const* __restrict__ peers, int world, int rank, size_t flags_offset, int epoch) {
  __threadfence_system();
  if ((int) threadIdx.x < world) {
    volatile int* f = peers[threadIdx.x] + flags_offset;
    f[rank] = epoch;
  }
}
// wait until every source rank has written this epoch into OUR table, then hand the table out
// The wait is BOUNDED (~4 s of SM clocks): a rank that never signals (crashed process, mismatched call sequence) must not
// leave this GPU spinning forever; the caller sees PSLAM_FLAG_P2P_TIMEOUT at the next status check and a table with stale rows.
__global__ void p2p_wait_copy_kernel(const int* __restrict__ table, size_t flags_offset, int world, int epoch, int cap_rows, int parity,
                                     int n, int* __restrict__ best, int* __restrict__ second, int* __restrict__ idx,
                                     int* __restrict__ status_flags) {
  if ((int) threadIdx.x < world) {
    const volatile int* f = table + flags_offset;
    const long long t0 = clock64();
    while (f[threadIdx.x] < epoch) {
      if (clock64() - t0 > 8000000000LL) {
        atomicOr(status_flags, PSLAM_FLAG_P2P_TIMEOUT);
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t base = (size_t) parity * 3 * cap_rows + i;
  best[i] = __ldcv(table + base);
  second[i] = __ldcv(table + base + cap_rows);
  idx[i] = __ldcv(table + base + 2 * (size_t) cap_rows);
}

}  // namespace

extern "C" {

int pslam_p2p_table_export(pslam_ctx* ctx, int max_rows, PslamIpcHandle* mine) {
  if (!ctx || max_rows <= 0 || !mine) return PSLAM_E_INVALID;
  static_assert(sizeof(PslamIpcHandle) == sizeof(cudaIpcMemHandle_t), "PslamIpcHandle must be a cudaIpcMemHandle_t");
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->d_p2p_table) return pslam_set_error(ctx, PSLAM_E_INVALID, "p2p table already exported (release it first)", cudaSuccess);
  const int cap = (max_rows + 255) / 256 * 256;
  const size_t words = 2 * 3 * (size_t) cap + P2P_FLAGS;
  PSLAM_CUDA_TRY(ctx, cudaMalloc((void**) &ctx->d_p2p_table, words * sizeof(int)));
  PSLAM_CUDA_TRY(ctx, cudaMemset(ctx->d_p2p_table, 0, words * sizeof(int)));
  ctx->p2p_cap_rows = cap;
  ctx->p2p_epoch = 0;
  cudaIpcMemHandle_t h;
  PSLAM_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->d_p2p_table));
  memcpy(mine, &h, sizeof(h));
  return PSLAM_OK;
}

int pslam_p2p_table_import(pslam_ctx* ctx, int rank, int world, const PslamIpcHandle* all) {
  if (!ctx || !all || world <= 0 || world > P2P_FLAGS || rank < 0 || rank >= world) return PSLAM_E_INVALID;
  if (!ctx->d_p2p_table) return pslam_set_error(ctx, PSLAM_E_INVALID, "p2p: export this rank's table first", cudaSuccess);
  if (ctx->d_p2p_peers) return pslam_set_error(ctx, PSLAM_E_INVALID, "p2p: peers already imported (release the table first)", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ctx->p2p_peer[r] = ctx->d_p2p_table;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, all + r, sizeof(h));
    void* p = nullptr;
    PSLAM_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->p2p_peer[r] = (int*) p;
  }
  PSLAM_CUDA_TRY(ctx, cudaMalloc((void**) &ctx->d_p2p_peers, sizeof(int*) * P2P_FLAGS));
  PSLAM_CUDA_TRY(ctx, cudaMemcpy(ctx->d_p2p_peers, ctx->p2p_peer, sizeof(int*) * (size_t) world, cudaMemcpyHostToDevice));
  ctx->p2p_world = world;
  ctx->p2p_rank = rank;
  return PSLAM_OK;
}

int pslam_p2p_table_release(pslam_ctx* ctx) {
  if (!ctx) return PSLAM_E_INVALID;
  if (!ctx->d_p2p_table) return PSLAM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->p2p_world; ++r)
    if (r != ctx->p2p_rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  if (ctx->d_p2p_peers) cudaFree(ctx->d_p2p_peers);
  cudaFree(ctx->d_p2p_table);
  ctx->d_p2p_table = nullptr;
  ctx->d_p2p_peers = nullptr;
  ctx->p2p_world = 0;
  memset(ctx->p2p_peer, 0, sizeof(ctx->p2p_peer));
  return PSLAM_OK;
}

int pslam_bf_best2_sharded_p2p_dev(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                                   const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second, int32_t* d_best_idx) {
  if (!ctx || n_fixed < 0 || n_moving <= 0) return PSLAM_E_INVALID;
  if (!ctx->d_p2p_peers || ctx->p2p_world <= 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "p2p: pslam_p2p_table_export / _import first", cudaSuccess);
  if (n_fixed > ctx->p2p_cap_rows) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "p2p: more rows than the exported table holds", cudaSuccess);
  if (n_fixed == 0) return PSLAM_OK;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int rank = ctx->p2p_rank, world = ctx->p2p_world;
  int b = 0, e = 0;
  pslam_shard_rows(n_fixed, rank, world, 256, &b, &e);
  // Every rank runs the same sequence of calls, so `epoch` advances in lock step.  Tables are double-buffered by epoch parity: a
  // peer can only write epoch e + 1 after it has seen OUR flag of epoch e, which we raise after having copied epoch e - 1 out.
  const int epoch = ++ctx->p2p_epoch, parity = epoch & 1;
  const size_t flags_offset = 2 * 3 * (size_t) ctx->p2p_cap_rows;
  const int rc = pslam_k_bf_best2_p2p(ctx, e - b, d_desc_fixed + 8 * (size_t) b, n_moving, d_desc_moving, b, ctx->p2p_cap_rows, parity,
                                      world, ctx->d_p2p_peers);
  if (rc) return rc;
  p2p_signal_kernel<<<1, P2P_FLAGS, 0, ctx->stream>>>(ctx->d_p2p_peers, world, rank, flags_offset, epoch);
  PSLAM_LAUNCH_CHECK(ctx, "p2p_signal_kernel");
  p2p_wait_copy_kernel<<<(n_fixed + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_p2p_table, flags_offset, world, epoch, ctx->p2p_cap_rows,
                                                                     parity, n_fixed, d_best, d_second, d_best_idx, ctx->d_flags);
  PSLAM_LAUNCH_CHECK(ctx, "p2p_wait_copy_kernel");
  return PSLAM_OK;
}

int pslam_shard_frames(int n_frames, int rank, int world, int* begin, int* end) {
  if (n_frames < 0 || world <= 0 || rank < 0 || rank >= world || !begin || !end) return PSLAM_E_INVALID;
  const int q = n_frames / world, r = n_frames % world;
  *begin = rank * q + (rank < r ? rank : r);
  *end = *begin + q + (rank < r ? 1 : 0);
  return PSLAM_OK;
}

int pslam_shard_rows(int n_rows, int rank, int world, int align, int* begin, int* end) {
  if (n_rows < 0 || world <= 0 || rank < 0 || rank >= world || align <= 0 || !begin || !end) return PSLAM_E_INVALID;
  long long per = ((long long) n_rows + world - 1) / world;
  per = (per + align - 1) / align * align;
  const long long b = (long long) rank * per < n_rows ? (long long) rank * per : n_rows;
  *begin = (int) b;
  *end = (int) (b + per < n_rows ? b + per : n_rows);
  return (int) per;  // rows per shard (the padded all-gather block)
}

int pslam_nccl_unique_id(pslam_ctx* ctx, PslamNcclId* id) {
  if (!ctx || !id) return PSLAM_E_INVALID;
  NcclApi* api = nccl_api(ctx);
  if (!api) return PSLAM_E_CUDA;
  const int r = api->GetUniqueId(id);
  return r ? nccl_fail(ctx, api, "ncclGetUniqueId", r) : PSLAM_OK;
}

int pslam_nccl_comm_create(pslam_ctx* ctx, const PslamNcclId* id, int rank, int world, void** comm) {
  if (!ctx || !id || !comm || world <= 0 || rank < 0 || rank >= world) return PSLAM_E_INVALID;
  NcclApi* api = nccl_api(ctx);
  if (!api) return PSLAM_E_CUDA;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int r = api->CommInitRank(comm, world, *id, rank);
  return r ? nccl_fail(ctx, api, "ncclCommInitRank", r) : PSLAM_OK;
}

int pslam_nccl_comm_destroy(pslam_ctx* ctx, void* comm) {
  if (!ctx || !comm) return PSLAM_E_INVALID;
  NcclApi* api = nccl_api(ctx);
  if (!api) return PSLAM_E_CUDA;
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int r = api->CommDestroy(comm);
  return r ? nccl_fail(ctx, api, "ncclCommDestroy", r) : PSLAM_OK;
}

int pslam_bf_best2_sharded_dev(pslam_ctx* ctx, void* nccl_comm, int rank, int world, int n_fixed, const uint32_t* d_desc_fixed,
                               int n_moving, const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second,
                               int32_t* d_best_idx) {
  if (!ctx || world <= 0 || rank < 0 || rank >= world || n_fixed < 0 || n_moving < 0 || (world > 1 && !nccl_comm))
    return PSLAM_E_INVALID;
  if (n_fixed == 0) return PSLAM_OK;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (world == 1 && !nccl_comm) return pslam_k_bf_best2(ctx, n_fixed, d_desc_fixed, n_moving, d_desc_moving, d_best, d_second, d_best_idx);
  int b = 0, e = 0;
  const int per = pslam_shard_rows(n_fixed, rank, world, 256, &b, &e);
  // staging at the END of the scratch buffer (the sweep carves its partial tables from the front):
  // [3][per] of this rank | [world][3][per] gathered
  const size_t need = sizeof(int) * 3 * (size_t) per * ((size_t) world + 1) + 512;
  if (need + ((size_t) 16 << 20) > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bf_best2_sharded: scratch too small for the gathered table", cudaSuccess);
  int* mine = reinterpret_cast<int*>(ctx->d_scratch + ((ctx->scratch_bytes - need) & ~(size_t) 255));
  int* gathered = mine + 3 * (size_t) per;
  if (e - b < per) PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(mine, 0, sizeof(int) * 3 * (size_t) per, ctx->stream));
  if (e > b) {
    const int rc = pslam_k_bf_best2(ctx, e - b, d_desc_fixed + 8 * (size_t) b, n_moving, d_desc_moving, mine, mine + per,
                                    mine + 2 * (size_t) per);
    if (rc) return rc;
  }
  NcclApi* api = nccl_api(ctx);
  if (!api) return PSLAM_E_CUDA;
  const int r = api->AllGather(mine, gathered, 3 * (size_t) per, NCCL_INT32, nccl_comm, ctx->stream);
  if (r) return nccl_fail(ctx, api, "ncclAllGather", r);
  unshard_best2_kernel<<<(n_fixed + 255) / 256, 256, 0, ctx->stream>>>(gathered, world, per, n_fixed, d_best, d_second, d_best_idx);
  PSLAM_LAUNCH_CHECK(ctx, "unshard_best2_kernel");
  return PSLAM_OK;
}

}  // extern "C"

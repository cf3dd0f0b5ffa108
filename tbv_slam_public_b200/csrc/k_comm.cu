// k_comm.cu — multi-GPU exchange of the hot path behind the C-ABI (tbv_comm_init / tbv_allgather_constraints; 8e).
//
// What it replaces: in the serial program every accepted loop candidate is appended to the pose graph in candidate order
// (ScanContextClosure::SearchAndAddConstraint, tbv_slam/src/tbv_slam/loopclosure.cpp:658-724 -> ApplyConstratins :261-318 ->
// PoseGraph::AddConstraintThSafe).  When the candidate list is sharded over GPUs (id_from mod world) each rank registers its
// share and this file gives every rank the complete list again: ONE ncclAllGather of fixed-size blocks on the context's stream — the count
// travels inside the block (record 0), so there is no second collective — followed by a device merge of the per-rank lists (each ascending
// by candidate index) into global candidate order.  Nothing here synchronises the host; the host-output entry points copy the merged
// records back afterwards.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy the host process already has loaded if any), so the library has no
// link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "tbv_comm.cuh"

static_assert(sizeof(tbv_constraint) == 128, "tbv_constraint is exchanged as raw 128-byte records");
static_assert(sizeof(ncclUniqueId) == TBV_COMM_ID_BYTES, "TBV_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

namespace tbv {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommCuDevice)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// The function table is immutable after the first successful load (call_once): shared read-only by every context, no launch state.
const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the host's own copy first (torch bundles one)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
    bool all = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) all = false; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
    api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
    api.CommCuDevice = reinterpret_cast<decltype(api.CommCuDevice)>(sym("ncclCommCuDevice"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    ok = all;
  });
  return ok ? &api : nullptr;
}

#define TBV_NCCL(api, call)                                                                          \
  do {                                                                                               \
    ncclResult_t _r = (call);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      tbv::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(_r)); \
      return TBV_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)
}  // namespace

struct CommState {
  ncclComm_t comm = nullptr;
  bool owned = false;          // created by tbv_comm_init_rank (destroyed with the context) vs borrowed from the host (tbv_comm_init)
  int world = 1, rank = 0;
  int capacity = 0;            // records per rank the buffers are sized for
  DevBuf<tbv_constraint> send; // [capacity + 1]
  DevBuf<tbv_constraint> recv; // [world][capacity + 1]
  DevBuf<tbv_constraint> all;  // [world * capacity] merged, candidate order
  DevBuf<int> n_all;           // [1]
  tbv_constraint* host_stage = nullptr;   // pinned, [world * capacity]: the merged records cross PCIe into pinned memory (pageable targets
  size_t host_stage_n = 0;                // are staged by the driver in small chunks: 10x slower for the 1 MB of an 8-rank batch)
};

static CommState* comm_state(tbv_ctx* ctx, bool create) {
  if (!ctx->comm && create) ctx->comm = new CommState();
  return static_cast<CommState*>(ctx->comm);
}

int comm_world(const tbv_ctx* ctx) { return ctx->comm ? static_cast<const CommState*>(ctx->comm)->world : 1; }
int comm_rank(const tbv_ctx* ctx) { return ctx->comm ? static_cast<const CommState*>(ctx->comm)->rank : 0; }

int comm_reserve(tbv_ctx* ctx, int capacity) {
  CommState* S = comm_state(ctx, true);
  if (capacity < 1) capacity = 1;
  if (capacity <= S->capacity) return TBV_OK;
  // the three buffers are sized together: the merge kernel addresses recv as [world][S->capacity + 1]
  S->send.release(); S->recv.release(); S->all.release();
  S->capacity = 0;
  int rc;
  if ((rc = S->send.reserve((size_t)capacity + 1)) || (rc = S->recv.reserve((size_t)S->world * ((size_t)capacity + 1))) ||
      (rc = S->all.reserve((size_t)S->world * capacity)) || (rc = S->n_all.reserve(1)))
    return rc;
  S->capacity = capacity;
  return TBV_OK;
}
tbv_constraint* comm_send_records(tbv_ctx* ctx) { return comm_state(ctx, true)->send.p + 1; }
int* comm_send_count(tbv_ctx* ctx) { return reinterpret_cast<int*>(comm_state(ctx, true)->send.p); }
tbv_constraint* comm_all(tbv_ctx* ctx) { return comm_state(ctx, true)->all.p; }
// Merged records -> `dst` (any host memory) through the context's pinned staging buffer; synchronises the stream.  The count and the
// records cross in ONE round when the whole merged area is small (<= 4 MB: every batch the loop-closure flow produces) — one stream
// synchronisation per call instead of count-then-records; larger areas fetch the count first and then exactly the valid records.
// *n_out = number of merged records (may exceed dst_capacity: the caller reports TBV_ERR_CAPACITY; only dst_capacity are copied).
int comm_fetch_all(tbv_ctx* ctx, tbv_constraint* dst, int dst_capacity, int* n_out) {
  CommState* S = comm_state(ctx, true);
  const size_t area = (size_t)S->world * (size_t)S->capacity;       // records the merge can have produced
  if (S->host_stage_n < area + 1) {
    if (S->host_stage) cudaFreeHost(S->host_stage);
    S->host_stage = nullptr; S->host_stage_n = 0;
    TBV_CUDA(cudaHostAlloc((void**)&S->host_stage, (area + 1) * sizeof(tbv_constraint), cudaHostAllocDefault));
    S->host_stage_n = area + 1;
  }
  int* count = reinterpret_cast<int*>(S->host_stage + area);         // the slot after the records holds the count
  const size_t guess = area < (size_t)dst_capacity ? area : (size_t)dst_capacity;
  const bool one_round = guess * sizeof(tbv_constraint) <= ((size_t)4 << 20);
  TBV_CUDA(cudaMemcpyAsync(count, S->n_all.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (one_round && guess > 0) TBV_CUDA(cudaMemcpyAsync(S->host_stage, S->all.p, guess * sizeof(tbv_constraint), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  const int n = *count;
  *n_out = n;
  const size_t take = (size_t)(n < dst_capacity ? n : dst_capacity);
  if (take == 0) return TBV_OK;
  if (!one_round) {
    TBV_CUDA(cudaMemcpyAsync(S->host_stage, S->all.p, take * sizeof(tbv_constraint), cudaMemcpyDeviceToHost, ctx->stream));
    TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  memcpy(dst, S->host_stage, take * sizeof(tbv_constraint));
  return TBV_OK;
}
int* comm_n_all(tbv_ctx* ctx) { return comm_state(ctx, true)->n_all.p; }

void comm_release(tbv_ctx* ctx) {
  CommState* S = static_cast<CommState*>(ctx->comm);
  if (!S) return;
  if (S->comm && S->owned) {
    const NcclApi* api = nccl_api();
    if (api) api->CommDestroy(S->comm);
  }
  S->send.release(); S->recv.release(); S->all.release(); S->n_all.release();
  if (S->host_stage) cudaFreeHost(S->host_stage);
  delete S;
  ctx->comm = nullptr;
}

// Merge of `world` rank blocks, each [stride] records = header + payload ascending by candidate, into one list ascending by candidate
// (ties — which a sharded candidate list never produces — by rank).  8 lanes move one 128-byte record (one uint4 each, coalesced); the
// output position of record (r, i) is the number of records of all blocks that sort before it: `world` binary searches.
__global__ void __launch_bounds__(256)
k_merge_constraints(const tbv_constraint* __restrict__ blocks, int world, int stride, int capacity, tbv_constraint* __restrict__ out,
                    int* __restrict__ n_out) {
  __shared__ int s_cnt[64];
  for (int r = threadIdx.x; r < world; r += blockDim.x) {
    int c = *reinterpret_cast<const int*>(blocks + (size_t)r * stride);
    s_cnt[r] = c < 0 ? 0 : (c > capacity ? capacity : c);
  }
  __syncthreads();
  int total = 0;
  for (int r = 0; r < world; r++) total += s_cnt[r];
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = total;
  const int sub = threadIdx.x & 7;
  const int per_block = blockDim.x >> 3;
  for (int e = blockIdx.x * per_block + (threadIdx.x >> 3); e < total; e += gridDim.x * per_block) {
    int r = 0, i = e;                                  // e-th record in (rank, index) enumeration
    while (i >= s_cnt[r]) { i -= s_cnt[r]; r++; }
    const tbv_constraint* mine = blocks + (size_t)r * stride + 1 + i;
    const int key = mine->candidate;
    int pos = 0;
    for (int q = 0; q < world; q++) {
      if (q == r) { pos += i; continue; }
      const tbv_constraint* lst = blocks + (size_t)q * stride + 1;
      int lo = 0, hi = s_cnt[q];                       // first index whose candidate is > key (q < r) or >= key (q > r)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int c = lst[mid].candidate;
        if (q < r ? c <= key : c < key) lo = mid + 1; else hi = mid;
      }
      pos += lo;
    }
    reinterpret_cast<uint4*>(out + pos)[sub] = reinterpret_cast<const uint4*>(mine)[sub];
  }
}

// The exchange on explicit buffers and an explicit stream: send [capacity + 1] (header + records), recv [world][capacity + 1],
// all [world * capacity], n_all [1].  tbv_loopdb_submit_sharded gives every batch in flight its own set and runs the exchange on a
// second stream, so that the collective of one batch overlaps the registration of the next.
int comm_allgather_merge_on(tbv_ctx* ctx, cudaStream_t stream, const tbv_constraint* send, tbv_constraint* recv, tbv_constraint* all,
                            int* n_all, int capacity) {
  CommState* S = comm_state(ctx, true);
  TBV_REQUIRE(capacity >= 1 && send && all && n_all, "bad exchange buffers");
  TBV_REQUIRE(S->world <= 64, "more than 64 ranks");
  // Only the used part of a block travels: header + `capacity` records; the receive blocks lie back to back with that stride.
  const int stride = capacity + 1;
  const tbv_constraint* blocks = send;
  if (S->world > 1) {
    const NcclApi* api = nccl_api();
    TBV_REQUIRE(api && S->comm && recv, "no NCCL communicator on this context (tbv_comm_init / tbv_comm_init_rank)");
    TBV_NCCL(api, api->AllGather(send, recv, (size_t)stride * sizeof(tbv_constraint), ncclChar, S->comm, stream));
    if (ctx->prof.on && stream == ctx->stream) prof_mark(ctx, "nccl_all_gather");   // not one of this library's kernels: timed, not counted as a launch
    blocks = recv;
  }
  int grid = (S->world * capacity + 31) / 32;
  if (grid > 4 * ctx->sm_count) grid = 4 * ctx->sm_count;
  if (grid < 1) grid = 1;
  k_merge_constraints<<<grid, 256, 0, stream>>>(blocks, S->world, stride, capacity, all, n_all);
  ctx->launches++;
  if (ctx->prof.on && stream == ctx->stream) prof_mark(ctx, "k_merge_constraints");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

int comm_allgather_merge(tbv_ctx* ctx, int capacity) {
  CommState* S = comm_state(ctx, true);
  TBV_REQUIRE(capacity >= 1 && capacity <= S->capacity, "exchange buffers are smaller than the requested capacity");
  return comm_allgather_merge_on(ctx, ctx->stream, S->send.p, S->recv.p, S->all.p, S->n_all.p, capacity);
}

}  // namespace tbv

using namespace tbv;

extern "C" {

int tbv_comm_unique_id(void* unique_id) {
  TBV_REQUIRE(unique_id, "null pointer");
  const NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return TBV_ERR_CUDA; }
  ncclUniqueId id;
  TBV_NCCL(api, api->GetUniqueId(&id));
  memcpy(unique_id, &id, sizeof(id));
  return TBV_OK;
}

int tbv_comm_init_rank(tbv_ctx* ctx, const void* unique_id, int world, int rank) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && unique_id && world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad arguments");
  const NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return TBV_ERR_CUDA; }
  comm_release(ctx);
  CommState* S = comm_state(ctx, true);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  TBV_NCCL(api, api->CommInitRank(&S->comm, world, id, rank));
  S->owned = true; S->world = world; S->rank = rank;
  return TBV_OK;
}

int tbv_comm_init(tbv_ctx* ctx, void* nccl_comm) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && nccl_comm, "null pointer");
  const NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return TBV_ERR_CUDA; }
  ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
  int world = 0, rank = 0, dev = -1;
  TBV_NCCL(api, api->CommCount(comm, &world));
  TBV_NCCL(api, api->CommUserRank(comm, &rank));
  TBV_NCCL(api, api->CommCuDevice(comm, &dev));
  TBV_REQUIRE(dev == ctx->device, "the communicator belongs to another CUDA device than the context");
  TBV_REQUIRE(world >= 1 && world <= 64, "communicators of 1..64 ranks are supported");
  comm_release(ctx);
  CommState* S = comm_state(ctx, true);
  S->comm = comm; S->owned = false; S->world = world; S->rank = rank;
  return TBV_OK;
}

int tbv_comm_world(tbv_ctx* ctx, int* world, int* rank) {
  TBV_REQUIRE(ctx, "null context");
  if (world) *world = comm_world(ctx);
  if (rank) *rank = comm_rank(ctx);
  return TBV_OK;
}

int tbv_comm_destroy(tbv_ctx* ctx) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx, "null context");
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  comm_release(ctx);
  return TBV_OK;
}

int tbv_allgather_constraints_dev(tbv_ctx* ctx, const tbv_constraint* local_dev, const int* n_local_dev, int capacity,
                                  const tbv_constraint** all_dev, const int** n_all_dev) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && local_dev && n_local_dev && capacity >= 1 && all_dev && n_all_dev, "bad arguments");
  int rc = comm_reserve(ctx, capacity);
  if (rc) return rc;
  TBV_CUDA(cudaMemcpyAsync(comm_send_count(ctx), n_local_dev, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  TBV_CUDA(cudaMemcpyAsync(comm_send_records(ctx), local_dev, (size_t)capacity * sizeof(tbv_constraint), cudaMemcpyDeviceToDevice, ctx->stream));
  if ((rc = comm_allgather_merge(ctx, capacity))) return rc;
  *all_dev = comm_all(ctx);
  *n_all_dev = comm_n_all(ctx);
  return TBV_OK;
}

int tbv_allgather_constraints(tbv_ctx* ctx, const tbv_constraint* local_dev, const int* n_local_dev, int capacity, tbv_constraint* all,
                              int all_capacity, int* n_all) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && n_all && all_capacity >= 0 && (all || all_capacity == 0), "bad arguments");
  const tbv_constraint* all_dev = nullptr;
  const int* n_all_dev = nullptr;
  int rc = tbv_allgather_constraints_dev(ctx, local_dev, n_local_dev, capacity, &all_dev, &n_all_dev);
  if (rc) return rc;
  if ((rc = comm_fetch_all(ctx, all, all_capacity, n_all))) return rc;
  if (*n_all > all_capacity) { set_error("tbv_allgather_constraints: %d records gathered, room for %d", *n_all, all_capacity); rc = TBV_ERR_CAPACITY; }
  return rc;
}

}  // extern "C"

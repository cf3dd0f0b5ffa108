// tbv_common.cuh — shared host/device plumbing of libtbv_b200.so (context, error handling, small device helpers).
// Product code: never includes anything from oracle/.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tbv_b200.h"

namespace tbv {

void set_error(const char* fmt, ...);

#define TBV_CUDA(call)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      tbv::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return TBV_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define TBV_REQUIRE(cond, msg)                   \
  do {                                           \
    if (!(cond)) {                               \
      tbv::set_error("invalid argument: %s", msg); \
      return TBV_ERR_INVALID;                    \
    }                                            \
  } while (0)

// Stream-ordered allocation for the host-pointer entry points: while an AllocScope is alive on this thread, DevBuf takes its memory
// from the device's default memory pool with cudaMallocAsync on the scope's stream and gives it back with cudaFreeAsync — a few
// microseconds once the pool is warm (tbv_create sets the pool's release threshold so that it keeps its memory), instead of the
// cudaMalloc / cudaFree pair (100+ us each, and cudaFree synchronises the device) these calls used to pay per temporary buffer.
// Long-lived state allocated outside a scope keeps plain cudaMalloc.
extern thread_local cudaStream_t g_alloc_stream;
struct AllocScope {
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t s) : prev(g_alloc_stream) { g_alloc_stream = s; }
  ~AllocScope() { g_alloc_stream = prev; }
};

// growable device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t pool_stream = nullptr;  // non-null: p came from cudaMallocAsync on this stream
  int reserve(size_t count) {
    if (count <= n) return TBV_OK;
    release();
    cudaError_t e;
    if (g_alloc_stream) {
      e = cudaMallocAsync((void**)&p, count * sizeof(T), g_alloc_stream);
      if (e == cudaSuccess) pool_stream = g_alloc_stream;
    } else {
      e = cudaMalloc((void**)&p, count * sizeof(T));
    }
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
      return TBV_ERR_CUDA;
    }
    n = count;
    return TBV_OK;
  }
  void release() {
    if (p) {
      if (pool_stream) cudaFreeAsync(p, pool_stream); else cudaFree(p);
    }
    p = nullptr;
    n = 0;
    pool_stream = nullptr;
  }
};

// A filtered cloud resident on the device, struct-of-arrays, `cap` entries per scan.
struct DevCloud {
  int cap = 0, batch = 0;
  DevBuf<float> x, y;
  DevBuf<uint8_t> inten;
  DevBuf<uint16_t> az, rg;
  DevBuf<int> count;
  int reserve(int batch_, int cap_) {
    batch = batch_;
    cap = cap_;
    const size_t n = (size_t)batch_ * cap_;
    int rc;
    if ((rc = x.reserve(n))) return rc;
    if ((rc = y.reserve(n))) return rc;
    if ((rc = inten.reserve(n))) return rc;
    if ((rc = az.reserve(n))) return rc;
    if ((rc = rg.reserve(n))) return rc;
    return count.reserve(batch_);
  }
  void release() { x.release(); y.release(); inten.release(); az.release(); rg.release(); count.release(); }
};

struct FilterState {  // device-resident result of the last k-strongest call
  int batch = 0, n_az = 0, n_range = 0, k = 0;
  DevBuf<uint8_t> polar;       // staging for host-input calls
  DevBuf<double2> cs_table;    // [n_az] (cos theta, sin theta) computed on the host with glibc
  DevBuf<double> th_table;     // [n_az] atan2(sin theta, cos theta), host glibc: the azimuth as Compensate's atan2 sees it
  int cs_n_az = 0;
  DevCloud filtered, peaks;
  DevCloud tmp_f, tmp_p;       // split scans only (small batches): per-block segments, moved into filtered / peaks by the scan's last CTA
  DevBuf<int2> seg_tot;        // [batch][split] points / peaks of every block
  DevBuf<int> seg_done;        // [batch] arrival counters, zero between launches
};

struct SmemOptIn {  // largest dynamic shared-memory size a kernel has been opted into on this context's device
  const void* func;
  size_t bytes;
};

struct Prof {  // optional per-launch device timing: one event after every kernel launch (tbv_profile_begin/_end)
  bool on = false;
  int n = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
};

}  // namespace tbv

struct tbv_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  int sm_count = 148;             // multiprocessors of `device` (queried in tbv_create)
  int smem_optin_max = 232448;    // largest opt-in shared memory per block of `device` (queried in tbv_create; 227 KB on sm_100)
  tbv::Prof prof;
  std::vector<tbv::SmemOptIn> smem_optin;  // per context (= per device): no process-global launch state
  tbv::FilterState filt;
  unsigned long long filt_epoch = 0;   // bumped by every filter call that writes the context's clouds on the context's stream (see tbv_odom_set_overlap)
  void* cells_scratch = nullptr;  // tbv::CellsScratch (k_cells.cu)
  void* reg_scratch = nullptr;    // tbv::RegScratch (k_register.cu)
  void* comm = nullptr;           // tbv::CommState (k_comm.cu): NCCL communicator + exchange buffers, or null
};

// Every C-ABI entry point makes its context's device current first: contexts on different devices may live in one process
// (one host thread per context, or one thread alternating between them).
#define TBV_ENTER(c)                             \
  do {                                           \
    const tbv_ctx* _tbv_c = (c);                 \
    if (_tbv_c) cudaSetDevice(_tbv_c->device);   \
  } while (0)

namespace tbv {
// Opt kernel `func` into `bytes` of dynamic shared memory on this context's device (a per-device attribute).  Always set on first use:
// the 48 KB default covers static + dynamic together, so a kernel with static shared memory needs the opt-in below 48 KB of dynamic too.
// The record is kept per context, so a second context on another device gets its own opt-in.
template <typename F>
inline int ensure_dyn_smem(tbv_ctx* ctx, F* func, size_t bytes) {
  const void* f = reinterpret_cast<const void*>(func);
  for (SmemOptIn& o : ctx->smem_optin)
    if (o.func == f) {
      if (o.bytes >= bytes) return TBV_OK;
      TBV_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      o.bytes = bytes;
      return TBV_OK;
    }
  TBV_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  ctx->smem_optin.push_back({f, bytes});
  return TBV_OK;
}
void prof_mark(tbv_ctx* ctx, const char* name);  // k_misc.cu
inline void launched(tbv_ctx* ctx, const char* name) {  // bookkeeping after every kernel launch of this library
  ctx->launches++;
  if (ctx->prof.on) prof_mark(ctx, name);
}
// implemented in k_filter.cu
// mot_dev != nullptr ([batch][3] previous frame-to-frame motion): both clouds are motion-compensated as they are emitted
// stream != nullptr: the kernel is launched there instead of on the context's stream (the caller orders it)
int filter_kstrongest_dev(tbv_ctx* ctx, const uint8_t* polar_dev, int n_az, int n_range, size_t row_stride, int batch,
                          const tbv_filter_params* params, int want_peaks, const double* mot_dev = nullptr, int ccw = 0, cudaStream_t stream = nullptr);
// the compensation of the fused filter as its own launch (same arithmetic, same bits): both clouds of the last filter call, in place
int compensate_polar_clouds_dev(tbv_ctx* ctx, const double* mot_dev /*[batch][3]*/, int ccw, int want_peaks);
// batched cv::rotate(ROTATE_90_COUNTERCLOCKWISE) on the device (k_misc.cu): src [batch][rows][cols] -> dst [batch][cols][rows]
int rotate90ccw_dev(tbv_ctx* ctx, const uint8_t* src_dev, int rows, int cols, int batch, uint8_t* dst_dev, cudaStream_t stream = nullptr);
int compensate_clouds_dev(tbv_ctx* ctx, DevCloud& cloud, const double* mot_dev /*[batch][3]*/, int ccw);
// Hash of the device pointers of a context's grow-on-demand scratch (they move when another caller on the same context asks for more):
// a captured graph of the odometry step is valid only while these are what they were at capture time.
inline uint64_t fp_mix(uint64_t h, const void* p) { return (h ^ (uint64_t)reinterpret_cast<uintptr_t>(p)) * 0x9E3779B97F4A7C15ull; }
uint64_t cells_fingerprint(tbv_ctx* ctx);   // k_cells.cu
uint64_t reg_fingerprint(tbv_ctx* ctx);     // k_register.cu
void cells_release(tbv_ctx* ctx);
void reg_release(tbv_ctx* ctx);
void comm_release(tbv_ctx* ctx);
}  // namespace tbv

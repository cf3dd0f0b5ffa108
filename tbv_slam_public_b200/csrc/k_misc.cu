// k_misc.cu — context lifecycle, error string, pinned-memory helpers, 90-degree rotation of range-major scans.
#include <mutex>

#include "tbv_common.cuh"

namespace tbv {
thread_local cudaStream_t g_alloc_stream = nullptr;
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

void prof_mark(tbv_ctx* ctx, const char* name) {
  Prof& P = ctx->prof;
  if ((int)P.ev.size() <= P.n) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    P.ev.push_back(e);
    P.names.push_back(name);
  }
  P.names[P.n] = name;
  cudaEventRecord(P.ev[P.n], ctx->stream);
  P.n++;
}

// dst(i, j) = src(j, W-1-i): cv::rotate(ROTATE_90_COUNTERCLOCKWISE) as used at radar_driver.cpp:80-84.
// 32x32 shared-memory tile transpose so both the read and the write are coalesced.
__global__ void k_rotate90ccw(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst) {
  __shared__ uint8_t tile[32][33];
  src += (size_t)blockIdx.z * H * W;   // blockIdx.z: image of a batch
  dst += (size_t)blockIdx.z * H * W;
  const int j0 = blockIdx.y * 32, c0 = blockIdx.x * 32;  // src rows j, src cols c
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int j = j0 + dy, c = c0 + threadIdx.x;
    if (j < H && c < W) tile[dy][threadIdx.x] = src[(size_t)j * W + c];
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, j = j0 + threadIdx.x;  // dst row i = W-1-c, dst col j
    if (j < H && c < W) dst[(size_t)(W - 1 - c) * H + j] = tile[threadIdx.x][dy];
  }
}
// The same rotation, 64 x 64-byte tiles, 4 bytes per thread per access (H and W multiples of 4, 4-byte aligned images): a byte-per-thread
// transpose issues 4x the load / store instructions for the same bytes and runs at a fraction of the HBM rate.
__global__ void __launch_bounds__(256) k_rotate90ccw_v4(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst) {
  __shared__ __align__(4) uint8_t tile[64][68];   // tile[c][j]: source column c, source row j (row stride 68: 4-byte aligned, spreads banks)
  src += (size_t)blockIdx.z * H * W;
  dst += (size_t)blockIdx.z * H * W;
  const int j0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int jl = ty + 16 * it, j = j0 + jl, c = c0 + 4 * tx;
    if (j < H && c < W) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(src + (size_t)j * W + c);
      tile[4 * tx + 0][jl] = v.x; tile[4 * tx + 1][jl] = v.y; tile[4 * tx + 2][jl] = v.z; tile[4 * tx + 3][jl] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int cl = ty + 16 * it, c = c0 + cl, j = j0 + 4 * tx;   // dst row i = W-1-c, dst cols j .. j+3
    if (c < W && j < H) *reinterpret_cast<uint32_t*>(dst + (size_t)(W - 1 - c) * H + j) = *reinterpret_cast<const uint32_t*>(&tile[cl][4 * tx]);
  }
}

int rotate90ccw_dev(tbv_ctx* ctx, const uint8_t* src_dev, int rows, int cols, int batch, uint8_t* dst_dev, cudaStream_t stream) {
  cudaStream_t st = stream ? stream : ctx->stream;
  const bool v4 = rows % 4 == 0 && cols % 4 == 0 && ((reinterpret_cast<uintptr_t>(src_dev) | reinterpret_cast<uintptr_t>(dst_dev)) & 3u) == 0;
  for (int b0 = 0; b0 < batch; b0 += 65535) {   // gridDim.z limit
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    const uint8_t* s = src_dev + (size_t)b0 * rows * cols;
    uint8_t* d = dst_dev + (size_t)b0 * rows * cols;
    if (v4) {
      k_rotate90ccw_v4<<<dim3((cols + 63) / 64, (rows + 63) / 64, nb), 256, 0, st>>>(s, rows, cols, d);
    } else {
      k_rotate90ccw<<<dim3((cols + 31) / 32, (rows + 31) / 32, nb), dim3(32, 8), 0, st>>>(s, rows, cols, d);
    }
    launched(ctx, "k_rotate90ccw");
  }
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}
}  // namespace tbv

using namespace tbv;

extern "C" {

const char* tbv_last_error(void) { return g_err.c_str(); }
int tbv_version(void) { return 100; }

tbv_ctx* tbv_create(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s): libtbv_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range [0,%d)", device, n);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  tbv_ctx* ctx = new tbv_ctx();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (ctx->sm_count <= 0) ctx->sm_count = 148;
  cudaDeviceGetAttribute(&ctx->smem_optin_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (ctx->smem_optin_max <= 0) ctx->smem_optin_max = 232448;
  // the context's stream takes the highest priority: side streams of the library (the filter stream of an overlapped odometry step) fill
  // the gaps its kernels leave instead of competing with them
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    delete ctx;
    return nullptr;
  }
  {  // keep freed blocks in the default pool: the host-pointer entry points allocate their temporaries from it on every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  return ctx;
}

void tbv_destroy(tbv_ctx* ctx) {
  TBV_ENTER(ctx);
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  FilterState& F = ctx->filt;
  F.polar.release(); F.cs_table.release(); F.th_table.release(); F.filtered.release(); F.peaks.release(); F.tmp_f.release(); F.tmp_p.release(); F.seg_tot.release(); F.seg_done.release();
  cells_release(ctx);
  reg_release(ctx);
  comm_release(ctx);
  for (cudaEvent_t e : ctx->prof.ev) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int tbv_profile_begin(tbv_ctx* ctx) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx, "null context");
  ctx->prof.n = 0;
  ctx->prof.on = true;
  prof_mark(ctx, "begin");
  return TBV_OK;
}

int tbv_profile_end(tbv_ctx* ctx, int capacity, const char** names, float* ms, int* n) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && n, "null pointer");
  Prof& P = ctx->prof;
  P.on = false;
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  int m = 0;
  for (int i = 1; i < P.n; i++) {
    float t = 0.f;
    TBV_CUDA(cudaEventElapsedTime(&t, P.ev[i - 1], P.ev[i]));
    if (m < capacity) {
      if (names) names[m] = P.names[i];
      if (ms) ms[m] = t;
    }
    m++;
  }
  *n = m;
  return TBV_OK;
}

void* tbv_stream(tbv_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int tbv_synchronize(tbv_ctx* ctx) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx, "null context");
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

long long tbv_launch_count(tbv_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* tbv_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    set_error("cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void tbv_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int tbv_rotate90ccw(tbv_ctx* ctx, const uint8_t* src, int rows, int cols, uint8_t* dst) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && src && dst && rows > 0 && cols > 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  const size_t n = (size_t)rows * cols;
  DevBuf<uint8_t> a, b;
  int rc;
  if ((rc = a.reserve(n)) || (rc = b.reserve(n))) { a.release(); b.release(); return rc; }
  cudaError_t e = cudaMemcpyAsync(a.p, src, n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    e = rotate90ccw_dev(ctx, a.p, rows, cols, 1, b.p) == TBV_OK ? cudaSuccess : cudaErrorUnknown;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(dst, b.p, n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  a.release(); b.release();
  if (e != cudaSuccess) { set_error("tbv_rotate90ccw: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  return TBV_OK;
}

}  // extern "C"

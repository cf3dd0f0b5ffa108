// k_cfar.cu — K1b: cell-averaging CFAR per azimuth.
//
// Replaces AzimuthCACFAR::getFilteredPointCloud + getMean + getCAScalingFactor
// (cfear_radarodometry/src/cfear_radarodometry/cfar.cpp:12-16, 28-83) as called by radarDriver::Process (radar_driver.cpp:52-56).
//
// The reference re-sums I^2 over the trailing / leading window for every candidate bin (O(window) per bin).  Here one warp
// owns one azimuth row: the row is staged in shared memory, an inclusive prefix sum of I^2 is built once (integers: exact,
// order-independent; < 2^32 for rows up to 8192 bins), and every bin's two window sums are two subtractions.  The means are
// the same doubles the reference gets (integer sum / count).  Detections are recorded as one bit per bin; a second kernel
// turns the bitmaps into the ordered cloud (azimuth-major, range ascending — the reference's push_back order).
#include <cmath>

#include "tbv_common.cuh"

namespace tbv {

constexpr int CF_WARPS = 4;

__global__ void __launch_bounds__(CF_WARPS * 32)
cfar_rows(const uint8_t* __restrict__ polar, int total_rows, int n_az, int n_range, size_t row_stride, int window, int guard,
          double range_res, double static_threshold, double min_distance, double max_distance, double scaling_factor, int words_per_row,
          uint32_t* __restrict__ bitmap, int* __restrict__ row_cnt) {
  extern __shared__ uint32_t s_pref[];  // [CF_WARPS][n_range + 1] exclusive prefix of I^2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  uint32_t* pref = s_pref + (size_t)warp * (n_range + 1);
  for (int row = blockIdx.x * CF_WARPS + warp; row < total_rows; row += gridDim.x * CF_WARPS) {
    const int scan = row / n_az, az = row - scan * n_az;
    const uint8_t* rp = polar + ((size_t)scan * n_az + az) * row_stride;
    // prefix sums, 32 bins per step
    uint32_t running = 0;
    for (int base = 0; base < n_range; base += 32) {
      const int r = base + lane;
      const uint32_t v = r < n_range ? (uint32_t)rp[r] : 0u;
      uint32_t inc = v * v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
      }
      if (r < n_range) pref[r] = running + inc - v * v;
      running += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) pref[n_range] = running;
    __syncwarp();
    int cnt = 0;
    for (int base = 0; base < n_range; base += 32) {
      const int r = base + lane;
      bool det = false;
      if (r < n_range) {
        const double range = range_res * double(r);
        const double intensity = double(rp[r]);
        if (range > min_distance && range < max_distance && intensity > static_threshold) {
          const int ts = max(0, r - guard - window), te = r - guard;          // trailing window [ts, te)
          const int fs = r + guard, fe = min(n_range, r + guard + window);    // leading window  [fs, fe)
          // an empty window is 0/0 = NaN in the reference: the bin is rejected
          if (te > ts && fe > fs && te >= 0 && fs <= n_range) {
            const double tmean = (double)(pref[te] - pref[ts]) / (double)(te - ts);
            const double fmean = (double)(pref[fe] - pref[fs]) / (double)(fe - fs);
            const double mean = (tmean + fmean) / 2.0;
            const double threshold = scaling_factor * mean;
            det = intensity * intensity > threshold;
          }
        }
      }
      const uint32_t bal = __ballot_sync(FULL, det);
      if (lane == 0) bitmap[(size_t)row * words_per_row + (base >> 5)] = bal;
      cnt += __popc(bal);
    }
    if (lane == 0) row_cnt[row] = cnt;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
cfar_emit(const uint8_t* __restrict__ polar, int n_az, int n_range, size_t row_stride, int words_per_row, const uint32_t* __restrict__ bitmap,
          const int* __restrict__ row_cnt, double range_res, const double2* __restrict__ cs_table, int cap, float* __restrict__ ox,
          float* __restrict__ oy, uint8_t* __restrict__ oi, uint16_t* __restrict__ oaz, uint16_t* __restrict__ org, int* __restrict__ ocount) {
  extern __shared__ int s_off[];  // [n_az + 1]
  const int scan = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const int* cnts = row_cnt + (size_t)scan * n_az;
  if (warp == 0) {
    int running = 0;
    for (int base = 0; base < n_az; base += 32) {
      const int v = (base + lane < n_az) ? cnts[base + lane] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
      }
      if (base + lane < n_az) s_off[base + lane] = running + inc - v;
      running += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) { s_off[n_az] = running; ocount[scan] = running; }
  }
  __syncthreads();
  const size_t cbase = (size_t)scan * cap;
  for (int az = warp; az < n_az; az += nwarps) {
    const uint8_t* rp = polar + ((size_t)scan * n_az + az) * row_stride;
    const uint32_t* bm = bitmap + ((size_t)scan * n_az + az) * words_per_row;
    const double2 cs = cs_table[az];
    int off = s_off[az];
    for (int w = 0; w < words_per_row; w++) {
      const uint32_t bal = bm[w];
      if (bal & (1u << lane)) {
        const int r = (w << 5) + lane;
        const int q = off + __popc(bal & ((1u << lane) - 1u));
        if (q < cap) {
          const double range = range_res * double(r);
          ox[cbase + q] = (float)(range * cs.x);   // cfar.cpp:63-64: no half-bin offset
          oy[cbase + q] = (float)(range * cs.y);
          oi[cbase + q] = rp[r];
          oaz[cbase + q] = (uint16_t)az;
          org[cbase + q] = (uint16_t)r;
        }
      }
      off += __popc(bal);
    }
  }
}

int ensure_cs_table(tbv_ctx* ctx, int n_az);  // k_filter.cu
int fetch_cloud(tbv_ctx* ctx, const DevCloud& d, int batch, tbv_points* out);

}  // namespace tbv

using namespace tbv;

extern "C" int tbv_filter_cacfar(tbv_ctx* ctx, const uint8_t* polar, int n_az, int n_range, size_t row_stride, int batch,
                                 const tbv_cfar_params* p, tbv_points* out) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && polar && p && out, "null pointer");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(n_az > 0 && n_range > 0 && batch > 0 && row_stride >= (size_t)n_range, "bad image shape");
  TBV_REQUIRE(n_range <= 8192 && n_az <= 4096, "image too large (n_range <= 8192, n_az <= 4096)");
  TBV_REQUIRE(p->window_size >= 1 && p->nb_guard_cells >= 0 && out->capacity > 0, "bad CFAR parameters");
  FilterState& F = ctx->filt;
  const size_t bytes = (size_t)batch * n_az * row_stride;
  const int words = (n_range + 31) / 32;
  const int total_rows = batch * n_az;
  DevBuf<uint32_t> bitmap;
  DevBuf<int> row_cnt;
  DevCloud cloud;
  int rc;
  auto cleanup = [&]() { bitmap.release(); row_cnt.release(); cloud.release(); };
  if ((rc = F.polar.reserve(bytes)) || (rc = bitmap.reserve((size_t)total_rows * words)) || (rc = row_cnt.reserve(total_rows)) ||
      (rc = cloud.reserve(batch, out->capacity)) || (rc = ensure_cs_table(ctx, n_az))) { cleanup(); return rc; }
  cudaError_t e = cudaMemcpyAsync(F.polar.p, polar, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { cleanup(); set_error("tbv_filter_cacfar: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  const double N = (double)(p->window_size * 2);                                  // cfar.cpp:32
  const double scaling = N * (std::pow(p->false_alarm_rate, -1. / N) - 1.);       // cfar.cpp:12-16 (host libm)
  const size_t smem = (size_t)CF_WARPS * (n_range + 1) * sizeof(uint32_t);
  if ((rc = ensure_dyn_smem(ctx, cfar_rows, smem))) { cleanup(); return rc; }
  const int grid = (total_rows + CF_WARPS - 1) / CF_WARPS;
  cfar_rows<<<grid < ctx->sm_count * 8 ? grid : ctx->sm_count * 8, CF_WARPS * 32, smem, ctx->stream>>>(F.polar.p, total_rows, n_az, n_range, row_stride, p->window_size,
                                                                                  p->nb_guard_cells, p->range_resolution, p->static_threshold,
                                                                                  p->min_distance, p->max_distance, scaling, words, bitmap.p, row_cnt.p);
  launched(ctx, "cfar_rows");
  cfar_emit<<<batch, 256, (n_az + 1) * sizeof(int), ctx->stream>>>(F.polar.p, n_az, n_range, row_stride, words, bitmap.p, row_cnt.p,
                                                                    p->range_resolution, F.cs_table.p, out->capacity, cloud.x.p, cloud.y.p,
                                                                    cloud.inten.p, cloud.az.p, cloud.rg.p, cloud.count.p);
  launched(ctx, "cfar_emit");
  e = cudaGetLastError();
  if (e != cudaSuccess) { cleanup(); set_error("tbv_filter_cacfar: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  rc = fetch_cloud(ctx, cloud, batch, out);
  cleanup();
  return rc;
}

// k_coral.cu — CorAl radar alignment quality, batched over scan pairs.
//
// Replaces CorAlRadarQuality (coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:8-229: GetNearby, Covariance,
// ComputeEntropy, constructor) as TBV calls it from ScanLearningInterface::getCorAlQualityMeasure
// (coral_alignment_quality/src/alignment_checker/alignmentinterface.cpp:437-454): for every point of the two peaks clouds (moved
// into a common frame), the radius-1 m neighbours in its own cloud and in both clouds -> sample covariances -> differential
// entropies 1/2 log(2 pi e det + 1e-8); the quality is (mean joint entropy, mean separate entropy, overlap).  In the reference this
// is 20.5 of the 23 ms spent verifying one loop candidate.
//
// One CTA per pair, everything in shared memory: both clouds are transformed (pcl::transformPointCloud arithmetic: double, narrowed
// per coordinate), bucketed on a common uniform grid (cell >= radius) with a STABLE counting sort (waves of 512 points, ranks from
// thread ids — the order inside a bucket is the cloud order, so every sum below is deterministic), then one thread per query point
// scans the 3 x 3 bucket block of both clouds.  The neighbour SETS are exact (float L2, strict <, FLANN's radius test); the
// covariances come from moments about the query point (|d| < radius, nothing cancels) instead of the reference's mean-subtracted
// matrix product — within a few ulp; the per-pair sums use a fixed tree.
#include <cfloat>
#include <cmath>

#include "tbv_common.cuh"

namespace tbv {

constexpr int CQ_THREADS = 512;
constexpr int CQ_CAP = 4096;     // points per cloud
constexpr int CQ_GN = 88;        // grid cells per axis at most (+1) -> <= 89 * 89 buckets
constexpr int CQ_NB = 8192;
constexpr int CQ_SMEM = 2 * CQ_CAP * 4 * 2      // px, py [2][CAP]
                        + 2 * CQ_CAP * 2        // original index of every sorted point [2][CAP]
                        + 2 * CQ_CAP * 2        // bucket of every point in cloud order [2][CAP]
                        + 2 * (CQ_NB + 1) * 4   // bucket starts [2][NB + 1]
                        + CQ_CAP * 2;           // wave slots

struct CoralPair {
  int src_first, n_src, ref_first, n_ref;  // slices of the concatenated cloud arrays
  double Ts[6], Tr[6];                     // r00 r01 tx r10 r11 ty of src->GetAffine() * Toffset and ref->GetAffine()
  long long pp_first;                      // first row of this pair in per_point, or -1
};

__device__ __forceinline__ float cq_min(float v) { for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d)); return v; }
__device__ __forceinline__ float cq_max(float v) { for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d)); return v; }

__global__ void __launch_bounds__(CQ_THREADS, 1)
k_coral(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ I, const CoralPair* __restrict__ pairs, float radius,
        int weight_res_intensity, int overlap_req, tbv_coral_result* __restrict__ results, double* __restrict__ per_point) {
  extern __shared__ __align__(16) uint8_t cq_smem[];
  float* px = reinterpret_cast<float*>(cq_smem);                 // [2][CAP]
  float* py = px + 2 * CQ_CAP;                                   // [2][CAP]
  uint16_t* oidx = reinterpret_cast<uint16_t*>(py + 2 * CQ_CAP); // [2][CAP]
  uint16_t* bidx = oidx + 2 * CQ_CAP;                            // [2][CAP]
  int* A = reinterpret_cast<int*>(bidx + 2 * CQ_CAP);            // [2][NB + 1]
  uint16_t* slot = reinterpret_cast<uint16_t*>(A + 2 * (CQ_NB + 1));
  __shared__ float s_red[4][CQ_THREADS / 32];
  __shared__ double s_acc[3][CQ_THREADS / 32];
  __shared__ int s_cnt[CQ_THREADS / 32], s_scan[CQ_THREADS / 32];
  __shared__ float s_grid[3];  // minx, miny, cell
  __shared__ int s_dim[2];
  const CoralPair P = pairs[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tbv_coral_result* out = results + blockIdx.x;
  const int n[2] = {P.n_src, P.n_ref};
  const int first[2] = {P.src_first, P.ref_first};
  const int merged = P.n_src + P.n_ref;
  if (P.n_src > CQ_CAP || P.n_ref > CQ_CAP || P.n_src <= 0 || P.n_ref <= 0) {  // the reference asserts both clouds non-empty (:117)
    if (tid == 0) { out->joint = 0; out->sep = 0; out->overlap = 0; out->count_valid = 0; out->merged_size = merged; out->valid = (P.n_src > CQ_CAP || P.n_ref > CQ_CAP) ? -1 : 0; }
    return;
  }
  // pcl::transformPointCloud(cloud, out, Affine3d): (float)(m00*x + m01*y + m02*z + m03), left to right, z = 0
  auto transform = [&](int c, int i, float& tx, float& ty) {
    const double* T = c == 0 ? P.Ts : P.Tr;
    const double x = (double)X[first[c] + i], y = (double)Y[first[c] + i];
    tx = (float)(((T[0] * x + T[1] * y) + 0.0 * 0.0) + T[2]);
    ty = (float)(((T[3] * x + T[4] * y) + 0.0 * 0.0) + T[5]);
  };
  // ---- bounding box of both clouds -> common grid --------------------------------------------------------------------
  {
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (int c = 0; c < 2; c++)
      for (int i = tid; i < n[c]; i += CQ_THREADS) {
        float a, b;
        transform(c, i, a, b);
        mnx = fminf(mnx, a); mxx = fmaxf(mxx, a); mny = fminf(mny, b); mxy = fmaxf(mxy, b);
      }
    mnx = cq_min(mnx); mny = cq_min(mny); mxx = cq_max(mxx); mxy = cq_max(mxy);
    if (lane == 0) { s_red[0][warp] = mnx; s_red[1][warp] = mny; s_red[2][warp] = mxx; s_red[3][warp] = mxy; }
    __syncthreads();
    if (warp == 0) {
      const bool in = lane < CQ_THREADS / 32;
      mnx = cq_min(in ? s_red[0][lane] : FLT_MAX); mny = cq_min(in ? s_red[1][lane] : FLT_MAX);
      mxx = cq_max(in ? s_red[2][lane] : -FLT_MAX); mxy = cq_max(in ? s_red[3][lane] : -FLT_MAX);
      if (lane == 0) {
        const float ext = fmaxf(mxx - mnx, mxy - mny);
        float cell = fmaxf(radius, ext / (float)CQ_GN);
        if (!(cell > 0.f) || !isfinite(cell)) cell = 1.f;
        s_grid[0] = mnx; s_grid[1] = mny; s_grid[2] = cell;
        s_dim[0] = min(CQ_GN + 1, (int)floorf((mxx - mnx) / cell) + 1);
        s_dim[1] = min(CQ_GN + 1, (int)floorf((mxy - mny) / cell) + 1);
      }
    }
    __syncthreads();
  }
  const float gminx = s_grid[0], gminy = s_grid[1], cell = s_grid[2];
  const int nx = s_dim[0], ny = s_dim[1], nb = nx * ny;
  auto bucket_of = [&](float v, float mn, int nn) { const int b = (int)floorf((v - mn) / cell); return b < 0 ? 0 : (b >= nn ? nn - 1 : b); };
  // ---- per cloud: histogram, scan, stable scatter --------------------------------------------------------------------
  for (int c = 0; c < 2; c++) {
    int* Ac = A + c * (CQ_NB + 1);
    for (int k = tid; k <= nb; k += CQ_THREADS) Ac[k] = 0;
    __syncthreads();
    for (int i = tid; i < n[c]; i += CQ_THREADS) {
      float a, b;
      transform(c, i, a, b);
      const int bk = bucket_of(b, gminy, ny) * nx + bucket_of(a, gminx, nx);
      bidx[c * CQ_CAP + i] = (uint16_t)bk;
      atomicAdd(&Ac[bk + 1], 1);
    }
    __syncthreads();
    {  // inclusive scan over Ac[0..nb]: Ac[k] = first sorted position of bucket k
      const int chunk = (nb + 1 + CQ_THREADS - 1) / CQ_THREADS;
      const int k0 = min(nb + 1, tid * chunk), k1 = min(nb + 1, k0 + chunk);
      int sum = 0;
      for (int k = k0; k < k1; k++) sum += Ac[k];
      int inc = sum;
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
      if (lane == 31) s_scan[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        const int v = lane < CQ_THREADS / 32 ? s_scan[lane] : 0;
        int iv = v;
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iv, d); if (lane >= d) iv += t; }
        if (lane < CQ_THREADS / 32) s_scan[lane] = iv - v;
      }
      __syncthreads();
      int run = s_scan[warp] + inc - sum;
      for (int k = k0; k < k1; k++) { run += Ac[k]; Ac[k] = run; }
    }
    __syncthreads();
    // stable scatter in waves of CQ_THREADS consecutive points (the scheme of the fused cells kernel): slots of one bucket drawn by
    // one wave are contiguous, the rank inside them is the number of smaller thread ids; Ac[k] is the cursor of bucket k
    for (int i0 = 0; i0 < n[c]; i0 += CQ_THREADS) {
      const int i = i0 + tid;
      const bool on = i < n[c];
      int bk = 0, base = 0;
      float a = 0.f, b = 0.f;
      if (on) { bk = bidx[c * CQ_CAP + i]; base = Ac[bk]; transform(c, i, a, b); }
      __syncthreads();
      if (on) slot[atomicAdd(&Ac[bk], 1)] = (uint16_t)tid;
      __syncthreads();
      int rank = 0;
      if (on) { const int end = Ac[bk]; for (int q = base; q < end; q++) rank += (slot[q] < tid); }
      __syncthreads();
      if (on) { const int pos = base + rank; px[c * CQ_CAP + pos] = a; py[c * CQ_CAP + pos] = b; oidx[c * CQ_CAP + pos] = (uint16_t)i; }
    }
    __syncthreads();
    {  // every cursor ended at the start of the next bucket: shift back (Ac[k] <- Ac[k - 1], Ac[0] <- 0), chunk by chunk from the top
      const int chunk = (nb + 1 + CQ_THREADS - 1) / CQ_THREADS;
      const int k0 = min(nb + 1, tid * chunk), k1 = min(nb + 1, k0 + chunk);
      const int prev = k0 >= 1 && k0 <= nb ? Ac[k0 - 1] : 0;
      __syncthreads();
      for (int k = k1 - 1; k > k0; k--) Ac[k] = Ac[k - 1];
      if (k0 < k1) Ac[k0] = prev;
    }
    __syncthreads();
  }
  // ---- one thread per query point (sorted order; src points first) -----------------------------------------------------
  const float r2 = (float)((double)radius * (double)radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius * radius)
  const double two_pi_e = 2.0 * M_PI * exp(1.0);
  double acc_j = 0.0, acc_s = 0.0, acc_w = 0.0;
  int acc_n = 0;
  for (int m = tid; m < merged; m += CQ_THREADS) {
    const int c = m < P.n_src ? 0 : 1;
    const int pos = c == 0 ? m : m - P.n_src;
    const float qx = px[c * CQ_CAP + pos], qy = py[c * CQ_CAP + pos];
    const double dqx = (double)qx, dqy = (double)qy;
    const int bx = bucket_of(qx, gminx, nx), by = bucket_of(qy, gminy, ny);
    int cnt[2] = {0, 0};
    double S[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};  // sum ex, ey, ex ex, ex ey, ey ey about the query, per cloud
    for (int d = 0; d < 2; d++) {
      const int* Ad = A + d * (CQ_NB + 1);
      const float* dx_ = px + d * CQ_CAP;
      const float* dy_ = py + d * CQ_CAP;
      for (int yy = max(by - 1, 0); yy <= min(by + 1, ny - 1); yy++) {
        const int j0 = Ad[yy * nx + max(bx - 1, 0)], j1 = Ad[yy * nx + min(bx + 1, nx - 1) + 1];
        for (int j = j0; j < j1; j++) {
          const float fx = dx_[j], fy = dy_[j];
          const float ddx = qx - fx, ddy = qy - fy;
          float dist = ddx * ddx;  // FLANN L2_Simple over the two dimensions of pcl::PointXY
          dist = dist + ddy * ddy;
          if (dist < r2) {
            const double ex = (double)fx - dqx, ey = (double)fy - dqy;
            cnt[d]++;
            S[d][0] += ex; S[d][1] += ey; S[d][2] += ex * ex; S[d][3] += ex * ey; S[d][4] += ey * ey;
          }
        }
      }
    }
    const int own = c, oth = 1 - c;
    double sep_e = 100.0, joint_e = 100.0;  // sep_res_ / joint_res_ initial values (:120-122)
    bool valid = false;
    if (cnt[oth] >= overlap_req && cnt[own] > 2) {  // :136,158 and Covariance's rows <= 2 test (:34)
      auto det_of = [](int rows, double sx, double sy, double sxx, double sxy, double syy) {
        const double nn = (double)rows;
        const double mx = sx / nn, my = sy / nn;
        const double den = (double)(float)rows - 1.0;  // "float n = x.rows(); cov = covSum*1.0/(n-1.0)" (:43-44)
        const double c00 = (sxx - nn * mx * mx) / den, c01 = (sxy - nn * mx * my) / den, c11 = (syy - nn * my * my) / den;
        return c00 * c11 - c01 * c01;
      };
      const double det_s = det_of(cnt[own], S[own][0], S[own][1], S[own][2], S[own][3], S[own][4]);
      const double det_j = det_of(cnt[0] + cnt[1], S[0][0] + S[1][0], S[0][1] + S[1][1], S[0][2] + S[1][2], S[0][3] + S[1][3], S[0][4] + S[1][4]);
      if (!(isnan(det_s) || isnan(det_j))) {
        const double se = 1.0 / 2.0 * log(two_pi_e * det_s + 0.00000001);
        const double je = 1.0 / 2.0 * log(two_pi_e * det_j + 0.00000001);
        if (!(isnan(se) || isnan(je))) { sep_e = se; joint_e = je; valid = true; }
      }
    }
    const int orig = (c == 0 ? 0 : P.n_src) + (int)oidx[c * CQ_CAP + pos];
    if (per_point && P.pp_first >= 0) {
      double* r = per_point + (size_t)(P.pp_first + orig) * 3;
      r[0] = sep_e; r[1] = joint_e; r[2] = valid ? 1.0 : 0.0;
    }
    if (valid) {
      const double w = weight_res_intensity ? (double)I[first[c] + (int)oidx[c * CQ_CAP + pos]] : 1.0;
      acc_w += w; acc_j += w * joint_e; acc_s += w * sep_e; acc_n++;
    }
  }
  // ---- fixed-shape reduction ---------------------------------------------------------------------------------------------
  for (int d = 16; d > 0; d >>= 1) {
    acc_j += __shfl_xor_sync(0xffffffffu, acc_j, d); acc_s += __shfl_xor_sync(0xffffffffu, acc_s, d);
    acc_w += __shfl_xor_sync(0xffffffffu, acc_w, d); acc_n += __shfl_xor_sync(0xffffffffu, acc_n, d);
  }
  if (lane == 0) { s_acc[0][warp] = acc_j; s_acc[1][warp] = acc_s; s_acc[2][warp] = acc_w; s_cnt[warp] = acc_n; }
  __syncthreads();
  if (tid == 0) {
    double j = 0, s = 0, w = 0;
    int cv = 0;
    for (int k = 0; k < CQ_THREADS / 32; k++) { j += s_acc[0][k]; s += s_acc[1][k]; w += s_acc[2][k]; cv += s_cnt[k]; }
    if (cv > 0) { s /= w; j /= w; }  // :188-192
    out->joint = j; out->sep = s;
    out->count_valid = cv; out->merged_size = merged;
    out->overlap = cv / ((double)merged);
    out->valid = out->overlap < 0.1 ? 0 : 1;  // :195-202
  }
}

}  // namespace tbv

using namespace tbv;

extern "C" int tbv_coral_quality_batch(tbv_ctx* ctx, int n_clouds, const float* const* x, const float* const* y, const float* const* intensity,
                                       const int* n_points, int n_pairs, const int* src_cloud, const int* ref_cloud, const double* T_src,
                                       const double* T_offset, const double* T_ref, const tbv_coral_params* params, tbv_coral_result* results,
                                       double* per_point) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && x && y && intensity && n_points && src_cloud && ref_cloud && T_src && T_ref && params && results && n_clouds >= 1 && n_pairs >= 0,
              "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(params->radius > 0, "radius must be positive");
  if (n_pairs == 0) return TBV_OK;
  cudaSetDevice(ctx->device);
  std::vector<int> first(n_clouds + 1, 0);
  for (int c = 0; c < n_clouds; c++) {
    TBV_REQUIRE(n_points[c] >= 0, "negative cloud size");
    first[c + 1] = first[c] + n_points[c];
  }
  const int total = first[n_clouds];
  std::vector<float> hx(total ? total : 1), hy(total ? total : 1), hi(total ? total : 1);
  for (int c = 0; c < n_clouds; c++) {
    if (!n_points[c]) continue;
    memcpy(&hx[first[c]], x[c], n_points[c] * sizeof(float));
    memcpy(&hy[first[c]], y[c], n_points[c] * sizeof(float));
    memcpy(&hi[first[c]], intensity[c], n_points[c] * sizeof(float));
  }
  auto aff = [](const double* v, double T[6]) {  // vectorToAffine3d (registration.cpp:129-135), planar part
    const double c = std::cos(v[2]), s = std::sin(v[2]);
    T[0] = c; T[1] = -s; T[2] = v[0]; T[3] = s; T[4] = c; T[5] = v[1];
  };
  std::vector<CoralPair> hp(n_pairs);
  long long pp = 0;
  for (int p = 0; p < n_pairs; p++) {
    TBV_REQUIRE(src_cloud[p] >= 0 && src_cloud[p] < n_clouds && ref_cloud[p] >= 0 && ref_cloud[p] < n_clouds, "pair indexes a missing cloud");
    CoralPair& P = hp[p];
    P.src_first = first[src_cloud[p]]; P.n_src = n_points[src_cloud[p]];
    P.ref_first = first[ref_cloud[p]]; P.n_ref = n_points[ref_cloud[p]];
    double A[6], B[6] = {1, 0, 0, 0, 1, 0};
    aff(T_src + 3 * p, A);
    if (T_offset) aff(T_offset + 3 * p, B);
    // src->GetAffine() * Toffset (AlignmentQuality.cpp:105): Eigen transform product, linear = Ra Rb, t = Ra tb + ta
    P.Ts[0] = A[0] * B[0] + A[1] * B[3]; P.Ts[1] = A[0] * B[1] + A[1] * B[4]; P.Ts[2] = (A[0] * B[2] + A[1] * B[5]) + A[2];
    P.Ts[3] = A[3] * B[0] + A[4] * B[3]; P.Ts[4] = A[3] * B[1] + A[4] * B[4]; P.Ts[5] = (A[3] * B[2] + A[4] * B[5]) + A[5];
    aff(T_ref + 3 * p, P.Tr);
    P.pp_first = per_point ? pp : -1;
    pp += P.n_src + P.n_ref;
  }
  DevBuf<float> dx, dy, di;
  DevBuf<CoralPair> dp;
  DevBuf<tbv_coral_result> dr;
  DevBuf<double> dpp;
  auto cleanup = [&]() { dx.release(); dy.release(); di.release(); dp.release(); dr.release(); dpp.release(); };
  int rc;
  if ((rc = dx.reserve(hx.size())) || (rc = dy.reserve(hy.size())) || (rc = di.reserve(hi.size())) || (rc = dp.reserve(n_pairs)) || (rc = dr.reserve(n_pairs)) ||
      (per_point && (rc = dpp.reserve((size_t)(pp ? pp : 1) * 3)))) { cleanup(); return rc; }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(dx.p, hx.data(), hx.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy.p, hy.data(), hy.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(di.p, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dp.p, hp.data(), hp.size() * sizeof(CoralPair), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && ensure_dyn_smem(ctx, k_coral, CQ_SMEM) != TBV_OK) { cleanup(); return TBV_ERR_CUDA; }
  if (e == cudaSuccess) {
    k_coral<<<n_pairs, CQ_THREADS, CQ_SMEM, st>>>(dx.p, dy.p, di.p, dp.p, (float)params->radius, params->weight_res_intensity,
                                                  params->overlap_req > 0 ? params->overlap_req : 1, dr.p, per_point ? dpp.p : nullptr);
    launched(ctx, "k_coral");
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(results, dr.p, n_pairs * sizeof(tbv_coral_result), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && per_point && pp) e = cudaMemcpyAsync(per_point, dpp.p, (size_t)pp * 3 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_coral_quality_batch: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  for (int p = 0; p < n_pairs; p++)
    if (results[p].valid < 0) { set_error("tbv_coral_quality_batch: pair %d has a cloud of more than %d points", p, CQ_CAP); return TBV_ERR_CAPACITY; }
  return TBV_OK;
}

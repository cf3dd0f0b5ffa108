// k_cells.cu — K3: oriented surface points ("cells") from a filtered cloud.
//
// Replaces MapPointNormal::MapPointNormal / ComputeNormals (cfear_radarodometry/src/cfear_radarodometry/pointnormal.cpp:65-90,
// 265-297), cell::cell (:7-36) and cell::ComputeNormal (:37-63), including the PCL VoxelGrid down-sampling and the
// FLANN radius search they call.  No kd-tree: the voxel grid that produces the sample points doubles as the spatial
// index (points are counting-sorted by voxel; the neighbours of a sample lie in at most a few contiguous runs of that
// order).  Kernels, all batched over scans:
//   C1 bounding box + voxel index per point + voxel histogram        (1 CTA / scan)
//   C2 exclusive scan over voxels, list of non-empty voxels = samples (1 CTA / scan)
//   C3 scatter points into voxel order
//   C4 per sample: order its points by cloud index, float centroid    (1 warp / sample)
//   C5 per sample: radius neighbours, sort by (d2, index), weighted mean / covariance in neighbour order,
//      2x2 symmetric eigen-solve, validity, planarity                 (1 warp / sample)
//   C6 ordered compaction of the valid cells                          (1 CTA / scan)
// Arithmetic follows the reference operation by operation (float voxel / distance math, double statistics, no FMA
// contraction: the library is built with -fmad=false).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "tbv_cells.cuh"

namespace tbv {

constexpr int C5_WARPS = 8;

struct GridInfo {
  int min_bx, min_by, div_x, div_y, n_vox, ok;
  float inv_leaf;
};

struct CellsScratch {
  int batch = 0, cap_pts = 0, vox_cap = 0, max_samples = 0;
  DevBuf<GridInfo> grid;
  DevBuf<int> vox_idx, vox_start, vox_fill, sample_vox, sorted_raw, sorted, err;
  DevBuf<float> cx, cy;
  DevBuf<float> sx, sy, si;     // [batch][cap_pts] voxel-sorted copy of the cloud (C4 -> C5)
  DevBuf<double> cand;          // [batch][CELL_FIELDS][max_samples]
  DevBuf<uint8_t> cand_valid;   // [batch][max_samples]
  DevBuf<long long> dbg;
  DevBuf<uint16_t> vidx16, order16;  // fused path, scans larger than CFU_PCAP points only
  void release() {
    vidx16.release(); order16.release();
    grid.release(); vox_idx.release(); vox_start.release(); vox_fill.release(); sample_vox.release(); sorted_raw.release();
    sorted.release(); err.release(); cx.release(); cy.release(); cand.release(); cand_valid.release(); sx.release(); sy.release(); si.release();
  }
};

__device__ __forceinline__ float warp_min(float v) {
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// ---- C1 -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
c1_grid(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ count, int cap_pts, float leaf, int vox_cap,
        GridInfo* __restrict__ grid, int* __restrict__ vox_idx, int* __restrict__ vox_cnt /*[batch][vox_cap+1]*/) {
  __shared__ float s_mn[2][8], s_mx[2][8];
  __shared__ GridInfo s_g;
  const int scan = blockIdx.x;
  const int n = min(count[scan], cap_pts);
  const float* px = x + (size_t)scan * cap_pts;
  const float* py = y + (size_t)scan * cap_pts;
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = px[i], b = py[i];
    mnx = fminf(mnx, a); mxx = fmaxf(mxx, a); mny = fminf(mny, b); mxy = fmaxf(mxy, b);
  }
  mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[0][warp] = mnx; s_mn[1][warp] = mny; s_mx[0][warp] = mxx; s_mx[1][warp] = mxy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      mnx = fminf(mnx, s_mn[0][w]); mny = fminf(mny, s_mn[1][w]); mxx = fmaxf(mxx, s_mx[0][w]); mxy = fmaxf(mxy, s_mx[1][w]);
    }
    GridInfo g;
    const float inv = 1.0f / leaf;  // PCL: inverse_leaf_size_ = 1 / leaf_size_
    g.inv_leaf = inv;
    g.ok = 0; g.min_bx = g.min_by = 0; g.div_x = g.div_y = 1; g.n_vox = 0;
    if (n > 0) {
      const long long dx = (long long)((mxx - mnx) * inv) + 1, dy = (long long)((mxy - mny) * inv) + 1;
      g.min_bx = (int)floorf(mnx * inv);
      g.min_by = (int)floorf(mny * inv);
      const int max_bx = (int)floorf(mxx * inv), max_by = (int)floorf(mxy * inv);
      g.div_x = max_bx - g.min_bx + 1;
      g.div_y = max_by - g.min_by + 1;
      const long long nv = (long long)g.div_x * (long long)g.div_y;
      g.n_vox = nv > 0x7fffffffLL ? 0x7fffffff : (int)nv;
      g.ok = (dx * dy <= 0x7fffffffLL) && (nv <= (long long)vox_cap);
    }
    s_g = g;
    grid[scan] = g;
  }
  __syncthreads();
  const GridInfo g = s_g;
  if (!g.ok) return;
  int* vi = vox_idx + (size_t)scan * cap_pts;
  int* vc = vox_cnt + (size_t)scan * (vox_cap + 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // PCL voxel_grid.hpp: ijk = (int)(floor(p * inverse_leaf_size) - (float)min_b)
    const int ijk0 = (int)(floorf(px[i] * g.inv_leaf) - (float)g.min_bx);
    const int ijk1 = (int)(floorf(py[i] * g.inv_leaf) - (float)g.min_by);
    const int idx = ijk0 + ijk1 * g.div_x;
    vi[i] = idx;
    atomicAdd(&vc[idx], 1);
  }
}

// ---- C2 -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
c2_scan(const GridInfo* __restrict__ grid, int vox_cap, int max_samples, int* __restrict__ vox_cnt_start /* in: counts, out: starts */,
        int* __restrict__ sample_vox, int* __restrict__ n_samples, int* __restrict__ err) {
  __shared__ int s_a[32], s_b[32];
  const int scan = blockIdx.x;
  const GridInfo g = grid[scan];
  if (!g.ok) {
    if (threadIdx.x == 0) { n_samples[scan] = 0; if (g.n_vox > 0) err[scan] = TBV_ERR_CAPACITY; }
    return;
  }
  int* vc = vox_cnt_start + (size_t)scan * (vox_cap + 1);
  const int nv = g.n_vox;
  const int chunk = (nv + blockDim.x - 1) / blockDim.x;
  const int b0 = min(nv, (int)threadIdx.x * chunk), b1 = min(nv, b0 + chunk);
  int sum = 0, ne = 0;
  for (int v = b0; v < b1; v++) { const int c = vc[v]; sum += c; ne += (c > 0); }
  // block exclusive scan of (sum, ne)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int is = sum, ie = ne;
  for (int d = 1; d < 32; d <<= 1) {
    const int ts = __shfl_up_sync(0xffffffffu, is, d), te = __shfl_up_sync(0xffffffffu, ie, d);
    if (lane >= d) { is += ts; ie += te; }
  }
  if (lane == 31) { s_a[warp] = is; s_b[warp] = ie; }
  __syncthreads();
  if (warp == 0) {
    int a = s_a[lane], b = s_b[lane];
    int ia = a, ib = b;
    for (int d = 1; d < 32; d <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
      if (lane >= d) { ia += ta; ib += tb; }
    }
    s_a[lane] = ia - a;
    s_b[lane] = ib - b;
    if (lane == 31) {
      n_samples[scan] = min(ib, max_samples);
      if (ib > max_samples) err[scan] = TBV_ERR_CAPACITY;
    }
  }
  __syncthreads();
  int off = s_a[warp] + is - sum, soff = s_b[warp] + ie - ne;
  int* sv = sample_vox + (size_t)scan * max_samples;
  for (int v = b0; v < b1; v++) {
    const int c = vc[v];
    vc[v] = off;
    off += c;
    if (c > 0) { if (soff < max_samples) sv[soff] = v; soff++; }
  }
  if (b1 == nv && b0 < nv) vc[nv] = off;          // total
  if (nv == 0 && threadIdx.x == 0) vc[0] = 0;
}

// ---- C3 -------------------------------------------------------------------------------------------------------
__global__ void c3_scatter(const GridInfo* __restrict__ grid, const int* __restrict__ count, int cap_pts, int vox_cap,
                           const int* __restrict__ vox_idx, const int* __restrict__ vox_start, int* __restrict__ vox_fill,
                           int* __restrict__ sorted_raw) {
  const int scan = blockIdx.y;
  if (!grid[scan].ok) return;
  const int n = min(count[scan], cap_pts);
  const int* vi = vox_idx + (size_t)scan * cap_pts;
  const int* vs = vox_start + (size_t)scan * (vox_cap + 1);
  int* vf = vox_fill + (size_t)scan * vox_cap;
  int* sr = sorted_raw + (size_t)scan * cap_pts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = vi[i];
    sr[vs[v] + atomicAdd(&vf[v], 1)] = i;
  }
}

// ---- C4 -------------------------------------------------------------------------------------------------------
// Eight lanes per sample: order the voxel's points by cloud index ([DEV-1]: PCL's std::sort leaves the intra-voxel order
// to libstdc++'s introsort; ascending index is used here and in the oracle's default mode), write them out in that order
// as a voxel-sorted struct-of-arrays copy of the cloud (x, y, intensity as float) — what C5 streams, coalesced — then
// accumulate the centroid in float exactly like pcl::CentroidPoint (sum in order, then divide by (float)n).
constexpr int C4_LANES = 8;
__global__ void __launch_bounds__(128)
c4_centroids(const GridInfo* __restrict__ grid, const int* __restrict__ n_samples, int max_samples, int cap_pts, int vox_cap,
             const float* __restrict__ x, const float* __restrict__ y, const uint8_t* __restrict__ inten_u8, const float* __restrict__ inten_f32,
             const int* __restrict__ sample_vox, const int* __restrict__ vox_start, const int* __restrict__ sorted_raw,
             float* __restrict__ sx_out, float* __restrict__ sy_out, float* __restrict__ si_out, float* __restrict__ cx, float* __restrict__ cy) {
  const int scan = blockIdx.y;
  if (!grid[scan].ok) return;
  const int ns = n_samples[scan];
  const int lane = threadIdx.x & 31, sub = lane & (C4_LANES - 1);
  const unsigned gmask = ((1u << C4_LANES) - 1u) << (lane & ~(C4_LANES - 1));
  const int group = threadIdx.x / C4_LANES, groups_per_cta = 128 / C4_LANES;
  const float* px = x + (size_t)scan * cap_pts;
  const float* py = y + (size_t)scan * cap_pts;
  const uint8_t* pi8 = inten_u8 ? inten_u8 + (size_t)scan * cap_pts : nullptr;
  const float* pif = inten_f32 ? inten_f32 + (size_t)scan * cap_pts : nullptr;
  const int* vs = vox_start + (size_t)scan * (vox_cap + 1);
  const int* sr = sorted_raw + (size_t)scan * cap_pts;
  float* ox = sx_out + (size_t)scan * cap_pts;
  float* oy = sy_out + (size_t)scan * cap_pts;
  float* oi = si_out + (size_t)scan * cap_pts;
  for (int s = blockIdx.x * groups_per_cta + group; s < ns; s += gridDim.x * groups_per_cta) {
    const int v = sample_vox[(size_t)scan * max_samples + s];
    const int s0 = vs[v], n = vs[v + 1] - s0;
    for (int e = sub; e < n; e += C4_LANES) {
      const int my = sr[s0 + e];
      int rank = 0;
      for (int j = 0; j < n; j++) rank += (sr[s0 + j] < my);
      ox[s0 + rank] = px[my];
      oy[s0 + rank] = py[my];
      oi[s0 + rank] = pi8 ? (float)pi8[my] : pif[my];
    }
    __syncwarp(gmask);
    if (sub == 0) {
      float sx = 0.f, sy = 0.f;
      for (int j = 0; j < n; j++) {
        sx += ox[s0 + j];
        sy += oy[s0 + j];
      }
      const float fn = (float)n;
      cx[(size_t)scan * max_samples + s] = sx / fn;
      cy[(size_t)scan * max_samples + s] = sy / fn;
    }
    __syncwarp(gmask);
  }
}

// ---- C5 -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double eig_hypot(double x, double y) {  // Eigen 3.3 numext::hypot
  double ax = fabs(x), ay = fabs(y), p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.0) return 0.0;
  return p * sqrt(1.0 + qp * qp);
}
__device__ __forceinline__ void make_givens(double p, double q, double& c, double& s) {  // Eigen JacobiRotation::makeGivens
  if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
  else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) {
    double t = q / p; double u = sqrt(1.0 + t * t); if (p < 0.0) u = -u;
    c = 1.0 / u; s = -t * c;
  } else {
    double t = p / q; double u = sqrt(1.0 + t * t); if (q < 0.0) u = -u;
    s = -1.0 / u; c = -t * s;
  }
}
// Eigen 3.3.7 SelfAdjointEigenSolver<Matrix2d>::compute: scale, (trivial) tridiagonalisation, implicit QR with
// Wilkinson shift, ascending sort.  evec[row][col].
__device__ void self_adjoint_eig2(double m00, double m10, double m11, double eval[2], double evec[2][2]) {
  double scale = fmax(fabs(m00), fmax(fabs(m10), fabs(m11)));
  if (scale == 0.0) scale = 1.0;
  double d0 = m00 / scale, d1 = m11 / scale, sub = m10 / scale;
  double Q00 = 1, Q01 = 0, Q10 = 0, Q11 = 1;
  const double precision = 2.0 * DBL_EPSILON, considerAsZero = DBL_MIN;
  int iter = 0;
  while (true) {
    if (fabs(sub) <= (fabs(d0) + fabs(d1)) * precision || fabs(sub) <= considerAsZero) sub = 0.0;
    if (sub == 0.0) break;
    iter++;
    if (iter > 60) break;
    const double td = (d0 - d1) * 0.5, e = sub;
    double mu = d1;
    if (td == 0.0) mu -= fabs(e);
    else if (e != 0.0) {
      const double e2 = e * e;
      const double h = eig_hypot(td, e);
      if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    const double xx = d0 - mu, z = sub;
    double c, s;
    make_givens(xx, z, c, s);
    const double sdk = s * d0 + c * sub;
    const double dkp1 = s * sub + c * d1;
    d0 = c * (c * d0 - s * sub) - s * (c * sub - s * d1);
    d1 = s * sdk + c * dkp1;
    sub = c * sdk - s * dkp1;
    double xi = Q00, yi = Q01;
    Q00 = c * xi - s * yi; Q01 = s * xi + c * yi;
    xi = Q10; yi = Q11;
    Q10 = c * xi - s * yi; Q11 = s * xi + c * yi;
  }
  if (d1 < d0) {
    double t = d0; d0 = d1; d1 = t;
    t = Q00; Q00 = Q01; Q01 = t;
    t = Q10; Q10 = Q11; Q11 = t;
  }
  eval[0] = d0 * scale; eval[1] = d1 * scale;
  evec[0][0] = Q00; evec[0][1] = Q01; evec[1][0] = Q10; evec[1][1] = Q11;
}

// Sixteen lanes per sample point (a window row holds ~13 points: a full warp would idle).  The neighbour SET is exact (float
// distance test of the FLANN radius search, strict <); the statistics follow cell::cell's formulas (normalised weights, mean,
// covariance about the mean; pointnormal.cpp:13-33) with per-lane partial sums in window-scan order and a fixed xor-tree
// across the 16 lanes: deterministic, and within a few ulp of the reference's neighbour-order sums (cells are
// tolerance-parity, the neighbour counts are exact).  Points are streamed from the voxel-sorted copy written by C4:
// consecutive lanes read consecutive floats.
constexpr int C5_LANES = 16;
constexpr int C5_MAXP = 4;      // window points cached per lane on the fast path (16 x 4 = 64 points)
constexpr int C5_MAXROWS = 4;   // window rows on the fast path
__global__ void __launch_bounds__(C5_WARPS * 32)
c5_cells(const GridInfo* __restrict__ grid, const int* __restrict__ n_samples, int max_samples, int cap_pts, int vox_cap,
         const float* __restrict__ sx_in, const float* __restrict__ sy_in, const float* __restrict__ si_in,
         const int* __restrict__ vox_start, const float* __restrict__ cx, const float* __restrict__ cy,
         float radius, int weight_intensity, double* __restrict__ cand) {
  const int scan = blockIdx.y;
  const GridInfo g = grid[scan];
  if (!g.ok) return;
  const int ns = n_samples[scan];
  const int lane = threadIdx.x & 31, sub = lane & (C5_LANES - 1);
  const unsigned gmask = 0xffffu << (lane & ~(C5_LANES - 1));
  const int group = threadIdx.x / C5_LANES, groups_per_cta = C5_WARPS * 32 / C5_LANES;
  const float* px = sx_in + (size_t)scan * cap_pts;
  const float* py = sy_in + (size_t)scan * cap_pts;
  const float* pi = si_in + (size_t)scan * cap_pts;
  const int* vs = vox_start + (size_t)scan * (vox_cap + 1);
  const float r2 = (float)((double)radius * (double)radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius*radius)
  const float eps = 1e-3f;
  double* cnd = cand + (size_t)scan * CELL_FIELDS * max_samples;
  const size_t st = max_samples;

  for (int s = blockIdx.x * groups_per_cta + group; s < ns; s += gridDim.x * groups_per_cta) {
    const float qx = cx[(size_t)scan * max_samples + s], qy = cy[(size_t)scan * max_samples + s];
    // voxel window that certainly contains every point with d2 < r2
    int ix0 = (int)(floorf((qx - radius - eps) * g.inv_leaf) - (float)g.min_bx);
    int ix1 = (int)(floorf((qx + radius + eps) * g.inv_leaf) - (float)g.min_bx);
    int iy0 = (int)(floorf((qy - radius - eps) * g.inv_leaf) - (float)g.min_by);
    int iy1 = (int)(floorf((qy + radius + eps) * g.inv_leaf) - (float)g.min_by);
    ix0 = max(ix0, 0); iy0 = max(iy0, 0); ix1 = min(ix1, g.div_x - 1); iy1 = min(iy1, g.div_y - 1);
    const int nrow = iy1 - iy0 + 1;
    // row runs of the window in the voxel-sorted order; all their bounds are fetched before any point is touched
    int ra[C5_MAXROWS], rl[C5_MAXROWS];
#pragma unroll
    for (int r = 0; r < C5_MAXROWS; r++) {
      ra[r] = 0; rl[r] = 0;
      if (r < nrow) {
        ra[r] = vs[(iy0 + r) * g.div_x + ix0];
        rl[r] = vs[(iy0 + r) * g.div_x + ix1 + 1] - ra[r];
      }
    }
    const int p1 = rl[0], p2 = p1 + rl[1], p3 = p2 + rl[2], total = p3 + rl[3];
    int cnt = 0;
    double wsum = 0.0, u0 = 0.0, u1 = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    if (nrow <= C5_MAXROWS && total <= C5_LANES * C5_MAXP) {
      // ---- fast path: the whole window (<= 64 points) is read ONCE into registers, lane `sub` holds points sub, sub+16, ...
      // of the concatenated rows; the three passes then run out of registers
      float fx[C5_MAXP], fy[C5_MAXP];
      double wv[C5_MAXP];
      bool in[C5_MAXP];
#pragma unroll
      for (int k = 0; k < C5_MAXP; k++) {
        const int t = sub + k * C5_LANES;
        in[k] = false; fx[k] = 0.f; fy[k] = 0.f; wv[k] = 0.0;
        if (t < total) {
          const int j = t < p1 ? ra[0] + t : (t < p2 ? ra[1] + (t - p1) : (t < p3 ? ra[2] + (t - p2) : ra[3] + (t - p3)));
          fx[k] = px[j]; fy[k] = py[j];
          const float dx = qx - fx[k], dy = qy - fy[k];
          float d = dx * dx;        // FLANN L2_Simple: result += diff*diff per dimension (z contributes +0)
          d = d + dy * dy;
          if (d < r2) {
            in[k] = true;
            wv[k] = weight_intensity ? fmax((double)pi[j] - 60.0, 0.0) : 1.0;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < C5_MAXP; k++)
        if (in[k]) { cnt++; wsum += wv[k]; }
      for (int d = C5_LANES / 2; d > 0; d >>= 1) { cnt += __shfl_xor_sync(gmask, cnt, d); wsum += __shfl_xor_sync(gmask, wsum, d); }
      if (cnt < 6) {  // pointnormal.cpp:291
        if (sub == 0) cnd[CF_NS * st + s] = 0.0;
        continue;
      }
#pragma unroll
      for (int k = 0; k < C5_MAXP; k++)
        if (in[k]) {
          wv[k] = wv[k] / wsum;
          u0 += wv[k] * (double)fx[k];
          u1 += wv[k] * (double)fy[k];
        }
      for (int d = C5_LANES / 2; d > 0; d >>= 1) { u0 += __shfl_xor_sync(gmask, u0, d); u1 += __shfl_xor_sync(gmask, u1, d); }
#pragma unroll
      for (int k = 0; k < C5_MAXP; k++)
        if (in[k]) {
          const double d0 = (double)fx[k] - u0, d1 = (double)fy[k] - u1;
          const double xw0 = wv[k] * d0, xw1 = wv[k] * d1;
          c00 += d0 * xw0; c01 += d0 * xw1; c10 += d1 * xw0; c11 += d1 * xw1;
        }
    } else {
      // ---- general path (large or tall windows): three streaming passes over the window rows
      for (int iy = iy0; iy <= iy1; iy++) {
        const int a = vs[iy * g.div_x + ix0], b = vs[iy * g.div_x + ix1 + 1];
        for (int j = a + sub; j < b; j += C5_LANES) {
          const float dx = qx - px[j], dy = qy - py[j];
          float d = dx * dx;
          d = d + dy * dy;
          if (d < r2) {
            cnt++;
            wsum += weight_intensity ? fmax((double)pi[j] - 60.0, 0.0) : 1.0;
          }
        }
      }
      for (int d = C5_LANES / 2; d > 0; d >>= 1) { cnt += __shfl_xor_sync(gmask, cnt, d); wsum += __shfl_xor_sync(gmask, wsum, d); }
      if (cnt < 6) {
        if (sub == 0) cnd[CF_NS * st + s] = 0.0;
        continue;
      }
      for (int iy = iy0; iy <= iy1; iy++) {
        const int a = vs[iy * g.div_x + ix0], b = vs[iy * g.div_x + ix1 + 1];
        for (int j = a + sub; j < b; j += C5_LANES) {
          const float fx = px[j], fy = py[j];
          const float dx = qx - fx, dy = qy - fy;
          float d = dx * dx;
          d = d + dy * dy;
          if (d < r2) {
            const double w = (weight_intensity ? fmax((double)pi[j] - 60.0, 0.0) : 1.0) / wsum;
            u0 += w * (double)fx;
            u1 += w * (double)fy;
          }
        }
      }
      for (int d = C5_LANES / 2; d > 0; d >>= 1) { u0 += __shfl_xor_sync(gmask, u0, d); u1 += __shfl_xor_sync(gmask, u1, d); }
      for (int iy = iy0; iy <= iy1; iy++) {
        const int a = vs[iy * g.div_x + ix0], b = vs[iy * g.div_x + ix1 + 1];
        for (int j = a + sub; j < b; j += C5_LANES) {
          const float fx = px[j], fy = py[j];
          const float dx = qx - fx, dy = qy - fy;
          float d = dx * dx;
          d = d + dy * dy;
          if (d < r2) {
            const double w = (weight_intensity ? fmax((double)pi[j] - 60.0, 0.0) : 1.0) / wsum;
            const double d0 = (double)fx - u0, d1 = (double)fy - u1;
            const double xw0 = w * d0, xw1 = w * d1;
            c00 += d0 * xw0; c01 += d0 * xw1; c10 += d1 * xw0; c11 += d1 * xw1;
          }
        }
      }
    }
    for (int d = C5_LANES / 2; d > 0; d >>= 1) {
      c00 += __shfl_xor_sync(gmask, c00, d); c01 += __shfl_xor_sync(gmask, c01, d);
      c10 += __shfl_xor_sync(gmask, c10, d); c11 += __shfl_xor_sync(gmask, c11, d);
    }
    if (sub == 0) {
      cnd[CF_U0 * st + s] = u0; cnd[CF_U1 * st + s] = u1;
      cnd[CF_C00 * st + s] = c00; cnd[CF_C01 * st + s] = c01; cnd[CF_C10 * st + s] = c10; cnd[CF_C11 * st + s] = c11;
      cnd[CF_SUMI * st + s] = wsum;
      cnd[CF_NS * st + s] = (double)cnt;
    }
  }
}

// ---- C6 -------------------------------------------------------------------------------------------------------
// One thread per sample: cell::ComputeNormal (pointnormal.cpp:37-63) — 2x2 eigen-solve, validity, planarity, normal
// orientation — then the ordered compaction of the valid cells (the reference's push_back order = sample order).
__global__ void __launch_bounds__(256)
c6_compact(const GridInfo* __restrict__ grid, const int* __restrict__ n_samples, int max_samples, const double* __restrict__ cand,
           double origin_x, double origin_y, double* __restrict__ out, int out_cap, int* __restrict__ out_count, int* __restrict__ err) {
  __shared__ int s_w[8];
  __shared__ int s_running;
  const int scan = blockIdx.x;
  const int ns = grid[scan].ok ? n_samples[scan] : 0;
  const double* cnd = cand + (size_t)scan * CELL_FIELDS * max_samples;
  const size_t st = max_samples;
  double* o = out + (size_t)scan * CELL_FIELDS * out_cap;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < ns; base += 256) {
    const int s = base + threadIdx.x;
    bool v = false;
    double rec[CELL_FIELDS];
    if (s < ns) {
      const double N = cnd[CF_NS * st + s];
      if (N >= 6.0) {
        const double u0 = cnd[CF_U0 * st + s], u1 = cnd[CF_U1 * st + s];
        const double c00 = cnd[CF_C00 * st + s], c01 = cnd[CF_C01 * st + s], c10 = cnd[CF_C10 * st + s], c11 = cnd[CF_C11 * st + s];
        const double wsum = cnd[CF_SUMI * st + s];
        double eval[2], evec[2][2];
        self_adjoint_eig2(c00, c10, c11, eval, evec);
        double n0 = evec[0][0], n1 = evec[1][0];
        const double lambda_min = eval[0], lambda_max = eval[1];
        const double condition_number = fabs(lambda_max / lambda_min);
        const double determinant = lambda_max * lambda_min;
        v = (condition_number <= 10000) && (determinant > 0.00001) && lambda_min > 0 && lambda_max > 0;
        const double pox = origin_x - u0, poy = origin_y - u1;
        if (n0 * pox + n1 * poy < 0) { n0 = -n0; n1 = -n1; }
        rec[CF_U0] = u0; rec[CF_U1] = u1; rec[CF_C00] = c00; rec[CF_C01] = c01; rec[CF_C10] = c10; rec[CF_C11] = c11;
        rec[CF_SCALE] = log(1.0 + condition_number / 2);
        rec[CF_N0] = n0; rec[CF_N1] = n1; rec[CF_O0] = evec[0][1]; rec[CF_O1] = evec[1][1];
        rec[CF_LMIN] = lambda_min; rec[CF_LMAX] = lambda_max; rec[CF_SUMI] = wsum; rec[CF_AVGI] = wsum / N; rec[CF_NS] = N;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = s_running;
    for (int w = 0; w < warp; w++) off += s_w[w];
    if (v) {
      const int q = off + __popc(bal & ((1u << lane) - 1u));
      if (q < out_cap)
#pragma unroll
        for (int f = 0; f < CELL_FIELDS; f++) o[(size_t)f * out_cap + q] = rec[f];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += s_w[w];
      s_running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_count[scan] = min(s_running, out_cap);
    if (s_running > out_cap) err[scan] = TBV_ERR_CAPACITY;
  }
}

// ---- fused path: C1..C6 in ONE launch, one CTA per scan, everything but the input cloud in shared memory -----------------------
// The six kernels above exchange per-scan tables through global memory and are latency-bound on them (ncu: C5 stalls on L1/L2
// loads at 37 % occupancy).  For the scan sizes of the odometry / loop paths all of those tables fit the 227 KB of one SM:
//   A      u16[n_vox + 1]   voxel histogram -> voxel start offsets (packed pairs, 32-bit shared atomics)
//   svox   u16[samples]     non-empty voxels in ascending voxel order (= PCL's sample order)
//   cx, cy f32[samples]     float centroids (pcl::CentroidPoint arithmetic, ascending cloud index inside a voxel [DEV-1])
//   px, py, pw f32[points]  voxel-sorted copy of the cloud;  vidx, order u16[points]
// Scans with more than CFU_PCAP points keep the five point arrays in global scratch instead (same code, generic pointers).
// The neighbour SET of a sample is exact (float distance test of the FLANN radius search).  The statistics are accumulated in ONE
// pass as weighted moments about the sample point q (|x - q| <= radius, so nothing cancels): W = sum w, S = sum w d, SS = sum w d d^T,
// then u = q + S / W and cov = SS / W - (S / W)(S / W)^T — algebraically cell::cell's normalised-weight mean / covariance
// (pointnormal.cpp:13-33), within a few ulp of it (cells are tolerance-parity).
constexpr int CFU_THREADS = 1024;
constexpr int CFU_VMAX = 22528;  // voxels of the scratch grid at 3 m leaves, extent = 1.05 x range + 8 m: Oxford (165 m) 15.4 k, MulRan (200 m) 22.2 k
constexpr int CFU_SMAX = 4096;   // samples per scan
constexpr int CFU_PCAP = 8192;   // points per scan held in shared memory
constexpr int CFU_CHUNK = (CFU_VMAX + 1 + CFU_THREADS - 1) / CFU_THREADS;  // entries of A per thread (23)
constexpr int CFU_OFF_A = 0;
constexpr int CFU_OFF_SVOX = ((CFU_VMAX + 2) * 2 + 15) & ~15;
constexpr int CFU_OFF_CX = CFU_OFF_SVOX + CFU_SMAX * 2;
constexpr int CFU_OFF_CY = CFU_OFF_CX + CFU_SMAX * 4;
constexpr int CFU_OFF_PX = CFU_OFF_CY + CFU_SMAX * 4;
constexpr int CFU_OFF_PY = CFU_OFF_PX + CFU_PCAP * 4;
constexpr int CFU_OFF_PW = CFU_OFF_PY + CFU_PCAP * 4;
constexpr int CFU_OFF_VIDX = CFU_OFF_PW + CFU_PCAP * 4;
constexpr int CFU_OFF_ORDER = CFU_OFF_VIDX + CFU_PCAP * 2;
constexpr int CFU_SMEM = CFU_OFF_ORDER + CFU_PCAP * 2;
constexpr int CFU_CAND_FIELDS = 8;  // u0 u1 c00 c01 c11 W N (+1 pad)

template <int L5>
__global__ void __launch_bounds__(CFU_THREADS, 1)
cells_fused(const float* __restrict__ x, const float* __restrict__ y, const uint8_t* __restrict__ inten_u8, const float* __restrict__ inten_f32,
            const int* __restrict__ count, int cap_pts, float leaf, float radius, int weight_intensity, int vox_cap, int max_samples,
            float* g_px, float* g_py, float* g_pw, uint16_t* g_vidx, uint16_t* g_order, double* cand, size_t cand_stride, double origin_x,
            double origin_y, double* __restrict__ out, int out_cap, int* __restrict__ out_count, int* __restrict__ n_samples_out,
            int* __restrict__ err, long long* __restrict__ dbg) {
#define CFU_TICK(i) do { if (dbg && threadIdx.x == 0) dbg[(size_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
  CFU_TICK(0);
  extern __shared__ __align__(16) uint8_t cfu_smem[];
  uint16_t* A = reinterpret_cast<uint16_t*>(cfu_smem + CFU_OFF_A);
  uint32_t* A32 = reinterpret_cast<uint32_t*>(cfu_smem + CFU_OFF_A);
  uint16_t* s_svox = reinterpret_cast<uint16_t*>(cfu_smem + CFU_OFF_SVOX);
  float* s_cx = reinterpret_cast<float*>(cfu_smem + CFU_OFF_CX);
  float* s_cy = reinterpret_cast<float*>(cfu_smem + CFU_OFF_CY);
  __shared__ float s_red[4][32];
  __shared__ int s_wa[32], s_wb[32];
  __shared__ GridInfo s_g;
  __shared__ int s_ns, s_running, s_heavy_n;
  const int scan = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = min(count[scan], cap_pts);
  const bool big = n > CFU_PCAP;
  if (tid == 0) s_heavy_n = 0;
  float* px = big ? g_px + (size_t)scan * cap_pts : reinterpret_cast<float*>(cfu_smem + CFU_OFF_PX);
  float* py = big ? g_py + (size_t)scan * cap_pts : reinterpret_cast<float*>(cfu_smem + CFU_OFF_PY);
  float* pw = big ? g_pw + (size_t)scan * cap_pts : reinterpret_cast<float*>(cfu_smem + CFU_OFF_PW);
  uint16_t* vidx = big ? g_vidx + (size_t)scan * cap_pts : reinterpret_cast<uint16_t*>(cfu_smem + CFU_OFF_VIDX);
  uint16_t* order = big ? g_order + (size_t)scan * cap_pts : reinterpret_cast<uint16_t*>(cfu_smem + CFU_OFF_ORDER);
  const float* gx = x + (size_t)scan * cap_pts;
  const float* gy = y + (size_t)scan * cap_pts;

  // ---- P0: bounding box -> voxel grid (pcl::VoxelGrid::applyFilter) ---------------------------------------------------------------
  {
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (int i = tid; i < n; i += CFU_THREADS) {
      const float a = gx[i], b = gy[i];
      mnx = fminf(mnx, a); mxx = fmaxf(mxx, a); mny = fminf(mny, b); mxy = fmaxf(mxy, b);
    }
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    if (lane == 0) { s_red[0][warp] = mnx; s_red[1][warp] = mny; s_red[2][warp] = mxx; s_red[3][warp] = mxy; }
    __syncthreads();
    if (warp == 0) {
      mnx = warp_min(s_red[0][lane]); mny = warp_min(s_red[1][lane]); mxx = warp_max(s_red[2][lane]); mxy = warp_max(s_red[3][lane]);
      if (lane == 0) {
        GridInfo g;
        const float inv = 1.0f / leaf;  // PCL: inverse_leaf_size_ = 1 / leaf_size_
        g.inv_leaf = inv;
        g.ok = 0; g.min_bx = g.min_by = 0; g.div_x = g.div_y = 1; g.n_vox = 0;
        if (n > 0) {
          const long long dx = (long long)((mxx - mnx) * inv) + 1, dy = (long long)((mxy - mny) * inv) + 1;
          g.min_bx = (int)floorf(mnx * inv);
          g.min_by = (int)floorf(mny * inv);
          const int max_bx = (int)floorf(mxx * inv), max_by = (int)floorf(mxy * inv);
          g.div_x = max_bx - g.min_bx + 1;
          g.div_y = max_by - g.min_by + 1;
          const long long nv = (long long)g.div_x * (long long)g.div_y;
          g.n_vox = nv > 0x7fffffffLL ? 0x7fffffff : (int)nv;
          g.ok = (dx * dy <= 0x7fffffffLL) && (nv <= (long long)vox_cap);
        }
        s_g = g;
      }
    }
    __syncthreads();
  }
  const GridInfo g = s_g;
  if (!g.ok) {
    if (tid == 0) {
      n_samples_out[scan] = 0;
      out_count[scan] = 0;
      err[scan] = g.n_vox > 0 ? TBV_ERR_CAPACITY : TBV_OK;
    }
    return;
  }
  const int nv = g.n_vox, nv1 = nv + 1;
  for (int k = tid; k < (nv1 + 2) / 2; k += CFU_THREADS) A32[k] = 0u;
  __syncthreads();

  CFU_TICK(1);
  // ---- P1: voxel index per point + histogram: count of voxel v is kept at A[v + 1] -------------------------------------------------
  for (int i = tid; i < n; i += CFU_THREADS) {
    // PCL voxel_grid.hpp: ijk = (int)(floor(p * inverse_leaf_size) - (float)min_b)
    const int ijk0 = (int)(floorf(gx[i] * g.inv_leaf) - (float)g.min_bx);
    const int ijk1 = (int)(floorf(gy[i] * g.inv_leaf) - (float)g.min_by);
    const int v = ijk0 + ijk1 * g.div_x;
    vidx[i] = (uint16_t)v;
    atomicAdd(&A32[(v + 1) >> 1], 1u << (((v + 1) & 1) << 4));
  }
  __syncthreads();

  // ---- P2: inclusive scan (A[k] = first sorted position of voxel k, A[nv] = n) + ascending list of the non-empty voxels -----------
  const int chunk = (nv1 + CFU_THREADS - 1) / CFU_THREADS;
  const int k0 = min(nv1, tid * chunk), k1 = min(nv1, k0 + chunk);
  {
    int sum = 0, ne = 0;
    for (int k = k0; k < k1; k++) { const int c = A[k]; sum += c; ne += (c > 0); }
    int is = sum, ie = ne;
    for (int d = 1; d < 32; d <<= 1) {
      const int ts = __shfl_up_sync(0xffffffffu, is, d), te = __shfl_up_sync(0xffffffffu, ie, d);
      if (lane >= d) { is += ts; ie += te; }
    }
    if (lane == 31) { s_wa[warp] = is; s_wb[warp] = ie; }
    __syncthreads();
    if (warp == 0) {
      const int a = s_wa[lane], b = s_wb[lane];
      int ia = a, ib = b;
      for (int d = 1; d < 32; d <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
      }
      s_wa[lane] = ia - a;
      s_wb[lane] = ib - b;
      if (lane == 31) {
        s_ns = min(ib, max_samples);
        n_samples_out[scan] = min(ib, max_samples);
        err[scan] = ib > max_samples ? TBV_ERR_CAPACITY : TBV_OK;
      }
    }
    __syncthreads();
    int run = s_wa[warp] + is - sum, soff = s_wb[warp] + ie - ne;
    for (int k = k0; k < k1; k++) {
      const int c = A[k];
      run += c;
      A[k] = (uint16_t)run;
      if (c > 0) { if (soff < max_samples) s_svox[soff] = (uint16_t)(k - 1); soff++; }
    }
  }
  __syncthreads();
  const int ns = s_ns;

  CFU_TICK(2);
  // ---- P3: scatter the points into voxel order (A[v] doubles as the cursor of voxel v), then shift A back -------------------------
  // STABLE: inside a voxel the points end up in ascending cloud index ([DEV-1], the order pcl::CentroidPoint adds them in), with
  // no sort.  The cloud is taken in waves of 1024 consecutive points, one per thread.  In a wave every thread (1) reads its voxel's
  // cursor, (2) draws a slot from it with an atomic and leaves its thread id there — the slots a wave draws for one voxel are
  // contiguous, in arrival order — (3) counts the smaller thread ids among its voxel's slots of this wave: that is its stable
  // rank, (4) stores the point at cursor + rank.  Waves are separated by barriers, so earlier waves (smaller indices) lie below.
  {
    const uint8_t* pi8 = inten_u8 ? inten_u8 + (size_t)scan * cap_pts : nullptr;
    const float* pif = inten_f32 ? inten_f32 + (size_t)scan * cap_pts : nullptr;
    for (int i0 = 0; i0 < n; i0 += CFU_THREADS) {
      const int i = i0 + tid;
      const bool on = i < n;
      int v = 0, base = 0, sh = 0;
      float fx = 0.f, fy = 0.f, fI = 0.f;
      if (on) {
        v = vidx[i];
        base = A[v];
        fx = gx[i]; fy = gy[i];
        fI = pi8 ? (float)pi8[i] : pif[i];
        sh = (v & 1) << 4;
      }
      __syncthreads();
      if (on) {
        const uint32_t old = atomicAdd(&A32[v >> 1], 1u << sh);
        order[(old >> sh) & 0xffffu] = (uint16_t)tid;
      }
      __syncthreads();
      int rank = 0;
      if (on) {
        const int end = A[v];
        for (int q = base; q < end; q++) rank += (order[q] < tid);
      }
      __syncthreads();
      if (on) {
        const int pos = base + rank;
        px[pos] = fx;
        py[pos] = fy;
        // cell::cell's weight max(I - 60, 0) (pointnormal.cpp:16-19): the float difference is exact for 30 <= I < 2^25 and its
        // sign is right below that, so the float holds the reference's double weight exactly
        pw[pos] = weight_intensity ? fmaxf(fI - 60.f, 0.f) : 1.f;
      }
    }
  }
  __syncthreads();
  {
    uint16_t keep[CFU_CHUNK];  // every cursor ended at the start of the next voxel: A[k] <- A[k - 1], A[0] <- 0
#pragma unroll
    for (int q = 0; q < CFU_CHUNK; q++) {
      const int k = k0 + q;
      keep[q] = (q < chunk && k < k1 && k >= 1) ? A[k - 1] : (uint16_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CFU_CHUNK; q++) {
      const int k = k0 + q;
      if (q < chunk && k < k1) A[k] = keep[q];
    }
  }
  __syncthreads();

  CFU_TICK(3);
  // ---- P4: float centroid of every sample (pcl::CentroidPoint: sum in cloud order, divide by (float)n) ---------------------------------
  {
    for (int s = tid; s < ns; s += CFU_THREADS) {
      const int v = s_svox[s];
      const int s0 = A[v], m = A[v + 1] - s0;
      float sx = 0.f, sy = 0.f;
      int j = 0;
      for (; j + 4 <= m; j += 4) {  // loads up front, the adds stay in cloud order
        const float x0 = px[s0 + j], x1 = px[s0 + j + 1], x2 = px[s0 + j + 2], x3 = px[s0 + j + 3];
        const float y0 = py[s0 + j], y1 = py[s0 + j + 1], y2 = py[s0 + j + 2], y3 = py[s0 + j + 3];
        sx += x0; sx += x1; sx += x2; sx += x3;
        sy += y0; sy += y1; sy += y2; sy += y3;
      }
      for (; j < m; j++) { sx += px[s0 + j]; sy += py[s0 + j]; }
      const float fn = (float)m;
      s_cx[s] = sx / fn;
      s_cy[s] = sy / fn;
    }
    __syncthreads();
    if (tid == 0) s_heavy_n = 0;
    __syncthreads();
  }

  CFU_TICK(4);
  // ---- P5: per sample, radius neighbours (exact set) and their weighted moments about the sample -----------------------------------
  double* cnd = cand + (size_t)scan * cand_stride;
  const size_t cst = max_samples;
  {
    const float r2 = (float)((double)radius * (double)radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius*radius)
    const float eps = 1e-3f;
    // window of sample s -> rows [iy0, iy1] x voxels [ix0, ix1]; returns the number of points in the window
    auto window = [&](float qx, float qy, int& ix0, int& ix1, int& iy0, int& iy1) -> int {
      ix0 = (int)(floorf((qx - radius - eps) * g.inv_leaf) - (float)g.min_bx);
      ix1 = (int)(floorf((qx + radius + eps) * g.inv_leaf) - (float)g.min_bx);
      iy0 = (int)(floorf((qy - radius - eps) * g.inv_leaf) - (float)g.min_by);
      iy1 = (int)(floorf((qy + radius + eps) * g.inv_leaf) - (float)g.min_by);
      ix0 = max(ix0, 0); iy0 = max(iy0, 0); ix1 = min(ix1, g.div_x - 1); iy1 = min(iy1, g.div_y - 1);
      int total = 0;
      for (int iy = iy0; iy <= iy1; iy++) total += (int)A[iy * g.div_x + ix1 + 1] - (int)A[iy * g.div_x + ix0];
      return total;
    };
    auto moments = [&](int s, float qx, float qy, int ix0, int ix1, int iy0, int iy1, int sub, int lanes, unsigned gmask) {
      const double dqx = (double)qx, dqy = (double)qy;
      int cnt = 0;
      double W = 0.0, Sx = 0.0, Sy = 0.0, Sxx = 0.0, Sxy = 0.0, Syy = 0.0;
      auto visit = [&](int j) {
        const float fx = px[j], fy = py[j];
        const float dx = qx - fx, dy = qy - fy;
        float d = dx * dx;  // FLANN L2_Simple: result += diff*diff per dimension (z contributes +0)
        d = d + dy * dy;
        if (d < r2) {
          const double w = (double)pw[j];
          const double ex = (double)fx - dqx, ey = (double)fy - dqy;
          const double wx = w * ex, wy = w * ey;
          cnt++;
          W += w; Sx += wx; Sy += wy;
          Sxx = __fma_rn(wx, ex, Sxx); Sxy = __fma_rn(wx, ey, Sxy); Syy = __fma_rn(wy, ey, Syy);
        }
      };
      const int nrow = iy1 - iy0 + 1;
      if (nrow <= 4) {
        // the window's rows as ONE index range [0, total): a row holds only a few points, so lanes stride over the
        // concatenation instead of over each row
        int ra[4], cum[4];
        int total = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          ra[r] = 0;
          if (r < nrow) {
            ra[r] = A[(iy0 + r) * g.div_x + ix0];
            total += (int)A[(iy0 + r) * g.div_x + ix1 + 1] - ra[r];
          }
          cum[r] = total;
        }
        for (int t = sub; t < total; t += lanes) {
          const int j = t < cum[0] ? ra[0] + t : (t < cum[1] ? ra[1] + (t - cum[0]) : (t < cum[2] ? ra[2] + (t - cum[1]) : ra[3] + (t - cum[2])));
          visit(j);
        }
      } else {
        for (int iy = iy0; iy <= iy1; iy++) {
          const int a = A[iy * g.div_x + ix0], b = A[iy * g.div_x + ix1 + 1];
          for (int j = a + sub; j < b; j += lanes) visit(j);
        }
      }
      for (int d = lanes >> 1; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(gmask, cnt, d);
        W += __shfl_xor_sync(gmask, W, d); Sx += __shfl_xor_sync(gmask, Sx, d); Sy += __shfl_xor_sync(gmask, Sy, d);
        Sxx += __shfl_xor_sync(gmask, Sxx, d); Sxy += __shfl_xor_sync(gmask, Sxy, d); Syy += __shfl_xor_sync(gmask, Syy, d);
      }
      if (sub == 0) {  // raw moments; P6 finishes them (only for samples with >= 6 neighbours, pointnormal.cpp:291)
        cnd[6 * cst + s] = (double)cnt;
        if (cnt >= 6) {
          cnd[0 * cst + s] = W; cnd[1 * cst + s] = Sx; cnd[2 * cst + s] = Sy;
          cnd[3 * cst + s] = Sxx; cnd[4 * cst + s] = Sxy; cnd[5 * cst + s] = Syy;
        }
      }
    };
    {
      const int sub = lane & (L5 - 1);
      const unsigned gmask = (L5 == 32) ? 0xffffffffu : (((1u << L5) - 1u) << (lane & ~(L5 - 1)));
      for (int s = tid / L5; s < ns; s += CFU_THREADS / L5) {
        const float qx = s_cx[s], qy = s_cy[s];
        int ix0, ix1, iy0, iy1;
        const int total = window(qx, qy, ix0, ix1, iy0, iy1);
        if (L5 < 32 && total > 24 * L5) {  // dense window: queue it for a whole warp
          if (sub == 0) vidx[atomicAdd(&s_heavy_n, 1)] = (uint16_t)s;
          continue;
        }
        moments(s, qx, qy, ix0, ix1, iy0, iy1, sub, L5, gmask);
      }
    }
    __syncthreads();
    const int nh = s_heavy_n;
    for (int h = warp; h < nh; h += CFU_THREADS / 32) {
      const int s = vidx[h];
      const float qx = s_cx[s], qy = s_cy[s];
      int ix0, ix1, iy0, iy1;
      window(qx, qy, ix0, ix1, iy0, iy1);
      moments(s, qx, qy, ix0, ix1, iy0, iy1, lane, 32, 0xffffffffu);
    }
  }
  if (tid == 0) s_running = 0;
  __syncthreads();  // also makes the candidate records (global) of this CTA visible to all of its threads

  CFU_TICK(5);
  // ---- P6: cell::ComputeNormal per candidate (pointnormal.cpp:37-63) + ordered compaction of the valid cells -----------------------
  // ordered position of every flagged thread of this round among all flagged so far (ascending thread id); 3 barriers
  auto ordered_slot = [&](bool flag) -> int {
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_wa[warp] = __popc(bal);
    __syncthreads();
    int off = s_running;
    for (int w = 0; w < warp; w++) off += s_wa[w];
    const int q = off + __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; w++) t += s_wa[w];
      s_running += t;
    }
    __syncthreads();
    return q;
  };
  // (a) samples with >= 6 neighbours, in sample order (about half of them): the eigen-solves then run on full warps
  uint16_t* clist = order;  // free since P4
  for (int base = 0; base < ns; base += CFU_THREADS) {
    const int s = base + tid;
    const bool c = s < ns && cnd[6 * cst + s] >= 6.0;
    const int q = ordered_slot(c);
    if (c) clist[q] = (uint16_t)s;
  }
  const int n_cand = s_running;
  __syncthreads();
  if (tid == 0) s_running = 0;
  __syncthreads();
  // (b) eigen-solve, validity, ordered output
  double* o = out + (size_t)scan * CELL_FIELDS * out_cap;
  for (int base = 0; base < n_cand; base += CFU_THREADS) {
    const int c = base + tid;
    bool v = false;
    double rec[CELL_FIELDS];
    if (c < n_cand) {
      const int s = clist[c];
      const double N = cnd[6 * cst + s];
      const double wsum = cnd[0 * cst + s];
      // mean / covariance from the moments about the sample point q: u = q + S / W, cov = SS / W - (S / W)(S / W)^T
      const double mx = cnd[1 * cst + s] / wsum, my = cnd[2 * cst + s] / wsum;
      const double u0 = (double)s_cx[s] + mx, u1 = (double)s_cy[s] + my;
      const double c00 = cnd[3 * cst + s] / wsum - mx * mx, c01 = cnd[4 * cst + s] / wsum - mx * my, c11 = cnd[5 * cst + s] / wsum - my * my;
      double eval[2], evec[2][2];
      self_adjoint_eig2(c00, c01, c11, eval, evec);
      double n0 = evec[0][0], n1 = evec[1][0];
      const double lambda_min = eval[0], lambda_max = eval[1];
      const double condition_number = fabs(lambda_max / lambda_min);
      const double determinant = lambda_max * lambda_min;
      v = (condition_number <= 10000) && (determinant > 0.00001) && lambda_min > 0 && lambda_max > 0;
      const double pox = origin_x - u0, poy = origin_y - u1;
      if (n0 * pox + n1 * poy < 0) { n0 = -n0; n1 = -n1; }
      rec[CF_U0] = u0; rec[CF_U1] = u1; rec[CF_C00] = c00; rec[CF_C01] = c01; rec[CF_C10] = c01; rec[CF_C11] = c11;
      rec[CF_SCALE] = log(1.0 + condition_number / 2);
      rec[CF_N0] = n0; rec[CF_N1] = n1; rec[CF_O0] = evec[0][1]; rec[CF_O1] = evec[1][1];
      rec[CF_LMIN] = lambda_min; rec[CF_LMAX] = lambda_max; rec[CF_SUMI] = wsum; rec[CF_AVGI] = wsum / N; rec[CF_NS] = N;
    }
    const int q = ordered_slot(v);
    if (v && q < out_cap)
#pragma unroll
      for (int f = 0; f < CELL_FIELDS; f++) o[(size_t)f * out_cap + q] = rec[f];
  }
  if (tid == 0) {
    out_count[scan] = min(s_running, out_cap);
    if (s_running > out_cap) err[scan] = TBV_ERR_CAPACITY;
  }
  CFU_TICK(6);
#undef CFU_TICK
}

// AoS <-> field-major conversion kernels
__global__ void k_cells_aos_to_soa(const double* __restrict__ aos, int n, double* __restrict__ soa, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int f = 0; f < CELL_FIELDS; f++) soa[(size_t)f * cap + i] = aos[(size_t)i * CELL_FIELDS + f];
}
__global__ void k_cells_soa_to_aos(const double* __restrict__ soa, int cap, int n, double* __restrict__ aos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int f = 0; f < CELL_FIELDS; f++) aos[(size_t)i * CELL_FIELDS + f] = soa[(size_t)f * cap + i];
}

// --------------------------------------------------------------------------------------------------------------
static CellsScratch* scratch(tbv_ctx* ctx) {
  if (!ctx->cells_scratch) ctx->cells_scratch = new CellsScratch();
  return (CellsScratch*)ctx->cells_scratch;
}
void cells_release(tbv_ctx* ctx) {
  if (!ctx->cells_scratch) return;
  CellsScratch* s = (CellsScratch*)ctx->cells_scratch;
  s->release();
  delete s;
  ctx->cells_scratch = nullptr;
}

uint64_t cells_fingerprint(tbv_ctx* ctx) {
  if (!ctx->cells_scratch) return 0;
  const CellsScratch& S = *(CellsScratch*)ctx->cells_scratch;
  uint64_t h = 1;
  for (const void* p : {(const void*)S.cand.p, (const void*)S.err.p, (const void*)S.sx.p, (const void*)S.sy.p, (const void*)S.si.p, (const void*)S.vidx16.p,
                        (const void*)S.order16.p, (const void*)S.dbg.p})
    h = fp_mix(h, p);
  return h;
}

int cells_build_dev(tbv_ctx* ctx, const float* x, const float* y, const uint8_t* inten_u8, const float* inten_f32, const int* count_dev,
                    int cap_pts, int batch, const CellsParams& par, int cell_cap, CellStore& out) {
  TBV_REQUIRE(par.radius > 0.f && par.downsample_factor > 0.0, "radius and downsample_factor must be positive");
  CellsScratch& S = *scratch(ctx);
  const float leaf = (float)((double)par.radius / par.downsample_factor);  // pointnormal.cpp:279
  // voxel grid scratch sized from the largest extent the caller vouches for
  const double ext = par.max_extent > 0 ? par.max_extent : 400.0;
  long long side = (long long)(2.0 * ext / leaf) + 4;
  long long vox_cap_ll = side * side;
  TBV_REQUIRE(vox_cap_ll <= (1ll << 24), "voxel grid too fine for the given extent (more than 2^24 voxels)");
  const int vox_cap = (int)vox_cap_ll;
  const int max_samples = par.max_samples > 0 ? par.max_samples : cell_cap;
  cudaStream_t st = ctx->stream;
  if (vox_cap <= CFU_VMAX && max_samples <= CFU_SMAX && cap_pts <= 65535) {
    // ---- fused path: one launch, one CTA per scan --------------------------------------------------------------------------
    int rc;
    if ((rc = S.err.reserve(batch)) || (rc = out.reserve(batch, cell_cap)))
      return rc;
    // per scan: the candidate records of P5/P6, or (before P5) 12 bytes per point as scratch for voxels of more than 1024 points
    const size_t cand_stride = std::max((size_t)CFU_CAND_FIELDS * max_samples, ((size_t)cap_pts * 12 + 7) / 8);
    if ((rc = S.cand.reserve((size_t)batch * cand_stride))) return rc;
    if (cap_pts > CFU_PCAP) {  // global home of the point arrays of oversized scans
      if ((rc = S.sx.reserve((size_t)batch * cap_pts)) || (rc = S.sy.reserve((size_t)batch * cap_pts)) || (rc = S.si.reserve((size_t)batch * cap_pts)) ||
          (rc = S.vidx16.reserve((size_t)batch * cap_pts)) || (rc = S.order16.reserve((size_t)batch * cap_pts)))
        return rc;
    }
#ifdef TBV_DEV_TIMERS
    const bool want_dbg = true;
#else
    const bool want_dbg = false;   // per-phase clocks: development builds only
#endif
    if (want_dbg) {
      if ((rc = S.dbg.reserve((size_t)batch * 8))) return rc;
    }
    // 2 lanes per sample in the neighbourhood phase (1 / 2 / 4 / 8 / 16 were measured on a B200; 2 is the fastest)
    if ((rc = ensure_dyn_smem(ctx, cells_fused<2>, CFU_SMEM))) return rc;
    cells_fused<2><<<batch, CFU_THREADS, CFU_SMEM, st>>>(x, y, inten_u8, inten_f32, count_dev, cap_pts, leaf, par.radius, par.weight_intensity, vox_cap, max_samples,
                                                         S.sx.p, S.sy.p, S.si.p, S.vidx16.p, S.order16.p, S.cand.p, cand_stride, par.origin[0], par.origin[1], out.f64.p, cell_cap,
                                                         out.count.p, out.n_samples.p, S.err.p, S.dbg.p);
    if (rc) return rc;
    launched(ctx, "cells_fused");
    TBV_CUDA(cudaGetLastError());
    if (want_dbg) {  // debug only: per-phase clocks averaged over the scans of this launch
      std::vector<long long> h((size_t)batch * 8);
      TBV_CUDA(cudaMemcpyAsync(h.data(), S.dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
      TBV_CUDA(cudaStreamSynchronize(st));
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int b = 0; b < batch; b++)
        for (int i = 0; i < 6; i++) acc[i] += (double)(h[(size_t)b * 8 + i + 1] - h[(size_t)b * 8 + i]);
      fprintf(stderr, "cells_fused phases (cycles/scan): P0 %.0f  P1-2 %.0f  P3 %.0f  P4 %.0f  P5 %.0f  P6 %.0f\n", acc[0] / batch, acc[1] / batch,
              acc[2] / batch, acc[3] / batch, acc[4] / batch, acc[5] / batch);
    }
    return TBV_OK;
  }
  int rc;
  if ((rc = S.grid.reserve(batch)) || (rc = S.err.reserve(batch)) || (rc = S.vox_idx.reserve((size_t)batch * cap_pts)) ||
      (rc = S.vox_start.reserve((size_t)batch * (vox_cap + 1))) || (rc = S.vox_fill.reserve((size_t)batch * vox_cap)) ||
      (rc = S.sample_vox.reserve((size_t)batch * max_samples)) || (rc = S.sorted_raw.reserve((size_t)batch * cap_pts)) ||
      (rc = S.sx.reserve((size_t)batch * cap_pts)) || (rc = S.sy.reserve((size_t)batch * cap_pts)) ||
      (rc = S.si.reserve((size_t)batch * cap_pts)) || (rc = S.cx.reserve((size_t)batch * max_samples)) ||
      (rc = S.cy.reserve((size_t)batch * max_samples)) || (rc = S.cand.reserve((size_t)batch * CELL_FIELDS * max_samples)) ||
      (rc = S.cand_valid.reserve((size_t)batch * max_samples)))
    return rc;
  if ((rc = out.reserve(batch, cell_cap))) return rc;
  TBV_CUDA(cudaMemsetAsync(S.vox_start.p, 0, (size_t)batch * (vox_cap + 1) * sizeof(int), st));
  TBV_CUDA(cudaMemsetAsync(S.vox_fill.p, 0, (size_t)batch * vox_cap * sizeof(int), st));
  TBV_CUDA(cudaMemsetAsync(S.err.p, 0, (size_t)batch * sizeof(int), st));
  c1_grid<<<batch, 256, 0, st>>>(x, y, count_dev, cap_pts, leaf, vox_cap, S.grid.p, S.vox_idx.p, S.vox_start.p);
  launched(ctx, "c1_grid");
  c2_scan<<<batch, 1024, 0, st>>>(S.grid.p, vox_cap, max_samples, S.vox_start.p, S.sample_vox.p, out.n_samples.p, S.err.p);
  launched(ctx, "c2_scan");
  {
    dim3 g((cap_pts + 255) / 256 < 64 ? (cap_pts + 255) / 256 : 64, batch);
    c3_scatter<<<g, 256, 0, st>>>(S.grid.p, count_dev, cap_pts, vox_cap, S.vox_idx.p, S.vox_start.p, S.vox_fill.p, S.sorted_raw.p);
    launched(ctx, "c3_scatter");
  }
  {
    const int g4 = 128 / C4_LANES, gx = (max_samples + g4 - 1) / g4 < 48 ? (max_samples + g4 - 1) / g4 : 48;
    c4_centroids<<<dim3(gx, batch), 128, 0, st>>>(S.grid.p, out.n_samples.p, max_samples, cap_pts, vox_cap, x, y, inten_u8, inten_f32, S.sample_vox.p,
                                                  S.vox_start.p, S.sorted_raw.p, S.sx.p, S.sy.p, S.si.p, S.cx.p, S.cy.p);
    launched(ctx, "c4_centroids");
    const int gp = C5_WARPS * 32 / C5_LANES, g5 = (max_samples + gp - 1) / gp < 48 ? (max_samples + gp - 1) / gp : 48;
    c5_cells<<<dim3(g5, batch), C5_WARPS * 32, 0, st>>>(S.grid.p, out.n_samples.p, max_samples, cap_pts, vox_cap, S.sx.p, S.sy.p, S.si.p,
                                                        S.vox_start.p, S.cx.p, S.cy.p, par.radius, par.weight_intensity, S.cand.p);
    launched(ctx, "c5_cells");
  }
  c6_compact<<<batch, 256, 0, st>>>(S.grid.p, out.n_samples.p, max_samples, S.cand.p, par.origin[0], par.origin[1], out.f64.p, cell_cap, out.count.p,
                                    S.err.p);
  launched(ctx, "c6_compact");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

int cells_errors(tbv_ctx* ctx, int batch, std::vector<int>& err) {
  CellsScratch& S = *scratch(ctx);
  err.resize(batch);
  TBV_CUDA(cudaMemcpyAsync(err.data(), S.err.p, batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}
const int* cells_err_dev(tbv_ctx* ctx) { return scratch(ctx)->err.p; }

int cells_upload(tbv_ctx* ctx, const tbv_cell* cells, int n, double* set_dev, int cap) {
  if (n <= 0) return TBV_OK;
  TBV_REQUIRE(n <= cap, "cell set larger than its device capacity");
  double* tmp = nullptr;
  TBV_CUDA(cudaMallocAsync((void**)&tmp, (size_t)n * sizeof(tbv_cell), ctx->stream));
  TBV_CUDA(cudaMemcpyAsync(tmp, cells, (size_t)n * sizeof(tbv_cell), cudaMemcpyHostToDevice, ctx->stream));
  k_cells_aos_to_soa<<<(n + 127) / 128, 128, 0, ctx->stream>>>(tmp, n, set_dev, cap);
  launched(ctx, "cells_aos_to_soa");
  TBV_CUDA(cudaFreeAsync(tmp, ctx->stream));
  return TBV_OK;
}
int cells_download(tbv_ctx* ctx, const double* set_dev, int cap, int n, tbv_cell* cells) {
  if (n <= 0) return TBV_OK;
  double* tmp = nullptr;
  TBV_CUDA(cudaMallocAsync((void**)&tmp, (size_t)n * sizeof(tbv_cell), ctx->stream));
  k_cells_soa_to_aos<<<(n + 127) / 128, 128, 0, ctx->stream>>>(set_dev, cap, n, tmp);
  launched(ctx, "cells_soa_to_aos");
  TBV_CUDA(cudaMemcpyAsync(cells, tmp, (size_t)n * sizeof(tbv_cell), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaFreeAsync(tmp, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

}  // namespace tbv

using namespace tbv;

extern "C" int tbv_build_cells(tbv_ctx* ctx, const float* x, const float* y, const float* intensity, int n, float radius,
                               double downsample_factor, int weight_intensity, const double origin[2], tbv_cell* cells, int cell_capacity,
                               int* n_cells, int* n_samples) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && x && y && intensity && origin && cells && n_cells, "null pointer");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(n >= 0 && cell_capacity > 0, "bad sizes");
  *n_cells = 0;
  if (n_samples) *n_samples = 0;
  if (n == 0) return TBV_OK;  // reference: exit(0) on an empty cloud (pointnormal.cpp:72-75); here: no cells
  // extent from the data (host side: the points are host-resident anyway)
  double ext = 1.0;
  for (int i = 0; i < n; i++) {
    ext = std::max(ext, (double)std::fabs(x[i]));
    ext = std::max(ext, (double)std::fabs(y[i]));
  }
  DevBuf<float> dx, dy, di;
  DevBuf<int> dc;
  int rc;
  auto cleanup = [&]() { dx.release(); dy.release(); di.release(); dc.release(); };
  if ((rc = dx.reserve(n)) || (rc = dy.reserve(n)) || (rc = di.reserve(n)) || (rc = dc.reserve(1))) { cleanup(); return rc; }
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(dx.p, x, n * sizeof(float), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dy.p, y, n * sizeof(float), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(di.p, intensity, n * sizeof(float), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dc.p, &n, sizeof(int), cudaMemcpyHostToDevice, st);
  CellsParams par;
  par.radius = radius; par.downsample_factor = downsample_factor; par.weight_intensity = weight_intensity;
  par.origin[0] = origin[0]; par.origin[1] = origin[1]; par.max_extent = ext + 1.0; par.max_samples = 0;
  CellStore store;
  // the number of samples is bounded by the number of points; candidates need that much room
  const int work_cap = n;
  int h_cnt = 0, h_ns = 0;
  std::vector<int> err;
  // Clouds of more than CFU_SMAX points first try the fused kernel with its sample capacity (a scan rarely occupies that many
  // voxels); only if that overflows is the general multi-kernel path run with room for one sample per point.
  for (int attempt = (n > CFU_SMAX ? 0 : 1); attempt < 2; attempt++) {
    par.max_samples = attempt == 0 ? CFU_SMAX : 0;
    rc = cells_build_dev(ctx, dx.p, dy.p, nullptr, di.p, dc.p, n, 1, par, work_cap, store);
    if (rc) { cleanup(); store.release(); return rc; }
    cudaMemcpyAsync(&h_cnt, store.count.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_ns, store.n_samples.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    rc = cells_errors(ctx, 1, err);
    if (rc) { cleanup(); store.release(); return rc; }
    if (!err[0]) break;
  }
  if (err[0]) { cleanup(); store.release(); set_error("tbv_build_cells: neighbourhood or grid capacity exceeded"); return TBV_ERR_CAPACITY; }
  *n_cells = h_cnt;
  if (n_samples) *n_samples = h_ns;
  const int ncopy = h_cnt < cell_capacity ? h_cnt : cell_capacity;
  rc = cells_download(ctx, store.set_ptr(0), store.cap, ncopy, cells);
  cleanup();
  store.release();
  if (rc) return rc;
  if (h_cnt > cell_capacity) { set_error("cell_capacity %d < %d cells", cell_capacity, h_cnt); return TBV_ERR_CAPACITY; }
  return TBV_OK;
}

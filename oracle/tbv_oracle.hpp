// =============================================================================================
// tbv_oracle.hpp — CPU ORACLE for the TBV / CFEAR hot path.   *** TEST INFRASTRUCTURE ONLY ***
//
// A dependency-free C++17 restatement of the reference algorithm (dan11003/tbv_slam_public @ 90f17c59)
// for the path  k-strongest / CA-CFAR filtering -> motion compensation -> oriented surface points
// ("cells") -> scan-to-keyframes registration (association + robust LM) -> keyframe fuser.
// Every function cites the reference file:line it restates (paths relative to /root/reference).
//
// Who may use this: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// The shipped product (tbv_slam_public_b200/, libtbv_b200.so) never includes, links or calls it.
//
// PARITY UNPINNED: the reference cannot be built here (needs ROS1/PCL/FLANN/OpenCV/Eigen/Ceres, none
// installed, no network) and its own tests hold no golden vectors for this path.
// Third-party arithmetic that the reference delegates to un-vendored libraries is restated from the
// published algorithms of the pinned versions (PCL 1.10 VoxelGrid / FLANN 1.9.1 radius + 1-NN search,
// Eigen 3.3.7 SelfAdjointEigenSolver, Ceres 2.1.0 trust-region LM + loss functions + corrector).
// Deliberate deviations are tagged  [DEV-n]  and listed in DESIGN.md.
//
// Build:  g++ -O3 -std=c++17 -ffp-contract=off   (no -ffast-math, no -march=native: the reference
// builds with plain -O3 on x86-64, i.e. SSE2 and no FMA contraction; cfear_radarodometry/CMakeLists.txt:4-5,32-33)
// =============================================================================================
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <utility>
#include <vector>

namespace tbv_oracle {

// ---------------------------------------------------------------------------------------------
// Basic types
// ---------------------------------------------------------------------------------------------
struct PointXYZI {  // pcl::PointXYZI as used by the reference (x,y,z,intensity floats)
  float x = 0, y = 0, z = 0, intensity = 0;
};
typedef std::vector<PointXYZI> Cloud;

typedef std::pair<uint8_t, int> intensity_range;  // radar_filters.h: typedef std::pair<uchar,int>

// One emitted point with its polar provenance — "filtered point indices" of the parity contract.
struct PolarPoint {
  uint16_t azimuth, range;
  uint8_t intensity;
  float x, y;
};

// =============================================================================================
// (a1) StructuredKStrongest::FilterKstrongest            radar_filters.cpp:198-237
// =============================================================================================
// Per azimuth: ascending vector of (intensity, range) kept via lower_bound + insert, erase(begin) when
// size > k.  The 400x3768 zero "sparse_filrered_" image the reference also fills (:203,231-235) is
// read by nothing on this path and is not restated (result-neutral; noted in BASELINE.md §2.5).
inline void FilterKstrongest(const uint8_t* img, int nb_azimuths, int nb_ranges, size_t row_stride,
                             int z_min, int k_strongest,
                             std::vector<std::vector<intensity_range>>& dense_filtered) {
  const uint8_t u_zmin = (uint8_t)z_min;  // :212  uchar(z_min_)
  dense_filtered.assign(nb_azimuths, std::vector<intensity_range>());
  for (int bearing = 0; bearing < nb_azimuths; bearing++) {
    std::vector<intensity_range>& v = dense_filtered[bearing];
    const uint8_t* row = img + (size_t)bearing * row_stride;
    for (int range = 0; range < nb_ranges; range++) {
      const uint8_t intensity = row[range];
      if (intensity < u_zmin) continue;  // :217
      if (v.empty()) {
        v.push_back(std::make_pair(intensity, range));
      } else {
        const intensity_range p = std::make_pair(intensity, range);
        auto it = std::lower_bound(v.cbegin(), v.cend(), p);  // :225 (std::pair operator<)
        v.insert(it, p);
        if (v.size() > (size_t)k_strongest) v.erase(v.begin());  // :227-228
      }
    }
  }
}

// =============================================================================================
// (a2) StructuredKStrongest::AxialNonMaxSupress          radar_filters.cpp:238-298
// =============================================================================================
// score(r') = sum of the 7 raw bytes flat[b*Nr + r'-3 .. r'+3] (uint16), memoised per azimuth only for
// r' within +-3 of a kept bin that lies in the guard band 3 <= r < Nr-3; scores never computed read as 0
// (unordered_map::operator[]).  The reference indexes cv::Mat::at<uchar>(bearing, r_nn) with r_nn possibly
// <0 or >=cols, i.e. it walks the contiguous image buffer into the neighbouring row; outside the whole
// image buffer that is undefined behaviour in the reference —  [DEV-3] bytes outside the buffer read as 0.
inline void AxialNonMaxSupress(const uint8_t* img, int nb_azimuths, int nb_ranges, size_t row_stride,
                               const std::vector<std::vector<intensity_range>>& dense_filtered,
                               std::vector<std::vector<intensity_range>>& dense_filtered_peaks) {
  const int window_size = 3;
  dense_filtered_peaks.assign(nb_azimuths, std::vector<intensity_range>());
  const long long total = (long long)(nb_azimuths - 1) * (long long)row_stride + nb_ranges;  // bytes in buffer
  std::vector<int> stamp(nb_ranges + 2 * window_size + 2, -1);  // memo validity per azimuth (index r_n + window)
  std::vector<uint16_t> score(nb_ranges + 2 * window_size + 2, 0);
  for (int bearing = 0; bearing < nb_azimuths; bearing++) {
    const long long base = (long long)bearing * (long long)row_stride;
    for (auto&& ir : dense_filtered[bearing]) {
      const int masked_range = ir.second;
      if (masked_range < window_size || masked_range >= nb_ranges - window_size) continue;  // :251
      for (int r_n = masked_range - window_size; r_n <= masked_range + window_size; r_n++) {
        if (stamp[r_n + window_size] != bearing) {  // :255 score.find(r_n)==score.end()
          uint16_t s = 0;
          for (int r_nn = r_n - window_size; r_nn <= r_n + window_size; r_nn++) {
            const long long flat = base + r_nn;  // :260 at<uchar>(bearing, r_nn) on a contiguous buffer
            s += (uint16_t)((flat >= 0 && flat < total) ? img[flat] : 0);
          }
          score[r_n + window_size] = s;
          stamp[r_n + window_size] = bearing;
        }
      }
    }
    auto get = [&](int r) -> uint16_t {  // score[r] with default-0 for never-computed keys (:271-276)
      const int i = r + window_size;
      if (i < 0 || i >= (int)stamp.size()) return 0;
      return stamp[i] == bearing ? score[i] : (uint16_t)0;
    };
    for (auto&& ir : dense_filtered[bearing]) {
      const int masked_range = ir.second;
      bool largest = true;
      const uint16_t pthis = get(masked_range);
      for (int i = 1; i <= window_size; i++) {
        const uint16_t pnext = get(masked_range + i);
        const uint16_t pprev = get(masked_range - i);
        if (pprev > pthis || pthis < pnext) {  // :282
          largest = false;
          break;
        }
      }
      if (largest) {
        const uint8_t intensity = img[base + masked_range];  // :292
        dense_filtered_peaks[bearing].push_back(std::make_pair(intensity, masked_range));
      }
    }
  }
}

// =============================================================================================
// (a3) StructuredKStrongest::getPeaksFilteredPointCloud   radar_filters.cpp:309-337
// =============================================================================================
// range_res arrives as radarDriver::Parameters::range_res (float, radar_driver.h:40-45) and is widened to
// double by the StructuredKStrongest ctor (radar_filters.h:86) — the caller passes the widened value.
inline int MinRangeBin(double min_distance, double range_res) { return (int)std::ceil(min_distance / range_res); }  // :315

inline void ToPointCloud(const std::vector<std::vector<intensity_range>>& vek, int nb_azimuths,
                         double min_distance, double range_res, Cloud& out, std::vector<PolarPoint>* polar = nullptr) {
  const int min_range_bin = MinRangeBin(min_distance, range_res);
  for (int bearing = 0; bearing < nb_azimuths; bearing++) {
    const double theta = (double(bearing + 1) / nb_azimuths) * 2. * M_PI;  // :317
    if (vek[bearing].empty()) continue;
    const double cos_t = std::cos(theta);
    const double sin_t = std::sin(theta);
    const double range_res_half = range_res / 2.0;
    for (auto&& ir : vek[bearing]) {
      const int range = ir.second;
      if (range > min_range_bin) {  // :327
        PointXYZI p;
        p.x = (float)((range_res_half + range_res * range) * cos_t);  // :329
        p.y = (float)((range_res_half + range_res * range) * sin_t);  // :330
        p.intensity = ir.first;
        p.z = 0;
        out.push_back(p);
        if (polar) polar->push_back({(uint16_t)bearing, (uint16_t)range, ir.first, p.x, p.y});
      }
    }
  }
}

// Convenience: radarDriver::Process k-strongest branch  radar_driver.cpp:57-61
struct KStrongestOutput {
  Cloud cloud, cloud_peaks;
  std::vector<PolarPoint> polar, polar_peaks;
};
inline void StructuredKStrongest(const uint8_t* img, int nb_azimuths, int nb_ranges, size_t row_stride,
                                 float z_min, int k_strongest, float min_distance, float range_res,
                                 KStrongestOutput& o, bool want_peaks = true) {
  std::vector<std::vector<intensity_range>> dense, peaks;
  FilterKstrongest(img, nb_azimuths, nb_ranges, row_stride, (int)z_min, k_strongest, dense);  // float -> int (radar_filters.h:86)
  ToPointCloud(dense, nb_azimuths, (double)min_distance, (double)range_res, o.cloud, &o.polar);
  if (want_peaks) {
    AxialNonMaxSupress(img, nb_azimuths, nb_ranges, row_stride, dense, peaks);
    ToPointCloud(peaks, nb_azimuths, (double)min_distance, (double)range_res, o.cloud_peaks, &o.polar_peaks);
  }
}

// MulRan-style input: MONO8 range-major image rotated 90 deg CCW on receipt (radar_driver.cpp:80-84):
// dst(i, j) = src(j, W-1-i) for src with H rows (range) x W cols (azimuth)  ->  dst has W rows x H cols.
inline void Rotate90CCW(const uint8_t* src, int H, int W, std::vector<uint8_t>& dst) {
  dst.resize((size_t)H * W);
  for (int i = 0; i < W; i++)
    for (int j = 0; j < H; j++) dst[(size_t)i * H + j] = src[(size_t)j * W + (W - 1 - i)];
}

// =============================================================================================
// (a4) AzimuthCACFAR                                      cfar.cpp:12-83, radar_driver.cpp:52-56
// =============================================================================================
struct CFARParams {
  int window_size = 40;
  double false_alarm_rate = 0.01;
  int nb_guard_cells = 10;
  double range_resolution = 0.0438;
  double static_threshold = 20;  // z_min
  double min_distance = 2.5;
  double max_distance = 400.0;  // hard-coded at radar_driver.cpp:54
};
inline double CAScalingFactor(double false_alarm_rate, int window_size) {  // cfar.cpp:12-16
  const double N = window_size;
  return N * (std::pow(false_alarm_rate, -1. / N) - 1.);
}
inline double CFARMean(const uint8_t* az, int start_idx, int end_idx) {  // cfar.cpp:73-83 (size_t i = start_idx)
  double sum = 0., N = 0.;
  // reference declares `size_t i = start_idx` and compares `i < end_idx` (int -> size_t): a negative
  // end_idx converts to a huge value; start_idx is clamped >= 0 by the caller and end_idx = r-g is < 0 only
  // when r < g, where start_idx = 0 and the loop would run off the row.  [DEV-4] such bins yield 0/0 = NaN
  // (rejected) here; they have range < min_distance for every shipped parameter set, so never reach the test.
  if (end_idx < 0) return std::numeric_limits<double>::quiet_NaN();
  for (int i = start_idx; i < end_idx; i++) {
    sum += std::pow(double(az[i]), 2.);
    N += 1.;
  }
  return sum / N;
}
inline void AzimuthCACFAR(const uint8_t* img, int rows, int cols, size_t row_stride, const CFARParams& p, Cloud& out,
                          std::vector<PolarPoint>* polar = nullptr) {
  const double scaling_factor = CAScalingFactor(p.false_alarm_rate, p.window_size * 2);  // cfar.cpp:32
  for (int azimuth_nb = 0; azimuth_nb < rows; azimuth_nb++) {
    const uint8_t* az = img + (size_t)azimuth_nb * row_stride;
    const double theta = (double(azimuth_nb + 1) / rows) * 2. * M_PI;
    for (int range_bin = 0; range_bin < cols; range_bin++) {
      const double range = p.range_resolution * double(range_bin);
      const double intensity = double(az[range_bin]);
      if (range > p.min_distance && range < p.max_distance && intensity > p.static_threshold) {
        const int trailing_window_start = std::max(0, range_bin - p.nb_guard_cells - p.window_size);
        const int trailing_window_end = range_bin - p.nb_guard_cells;
        const double trailing_mean = CFARMean(az, trailing_window_start, trailing_window_end);
        const int forwarding_window_start = range_bin + p.nb_guard_cells;
        const int forwarding_window_end = std::min(cols, range_bin + p.nb_guard_cells + p.window_size);
        const double forwarding_mean = CFARMean(az, forwarding_window_start, forwarding_window_end);
        const double mean = (trailing_mean + forwarding_mean) / 2.0;
        const double threshold = scaling_factor * mean;
        const double squared_intensity = std::pow(intensity, 2.);
        if (squared_intensity > threshold) {  // NaN threshold compares false
          PointXYZI q;
          q.x = (float)(range * std::cos(theta));
          q.y = (float)(range * std::sin(theta));
          q.intensity = (float)intensity;
          out.push_back(q);
          if (polar) polar->push_back({(uint16_t)azimuth_nb, (uint16_t)range_bin, az[range_bin], q.x, q.y});
        }
      }
    }
  }
}

// =============================================================================================
// Planar rigid transforms with the reference's Eigen semantics
// =============================================================================================
// The reference carries Eigen::Affine3d that are planar by construction (vectorToAffine3d,
// registration.cpp:129-135: Translation * Rz(theta)); only the top-left 2x2 and (x,y) take part.
struct Affine2 {
  double r00 = 1, r01 = 0, r10 = 0, r11 = 1, tx = 0, ty = 0;
};
inline Affine2 vectorToAffine(double x, double y, double theta) {  // registration.cpp:129-149 (AngleAxis about Z)
  Affine2 T;
  const double c = std::cos(theta), s = std::sin(theta);
  T.r00 = c; T.r01 = -s; T.r10 = s; T.r11 = c; T.tx = x; T.ty = y;
  return T;
}
inline Affine2 Mul(const Affine2& A, const Affine2& B) {  // Eigen Transform product: linear = Ra*Rb, t = Ra*tb + ta
  Affine2 C;
  C.r00 = A.r00 * B.r00 + A.r01 * B.r10;
  C.r01 = A.r00 * B.r01 + A.r01 * B.r11;
  C.r10 = A.r10 * B.r00 + A.r11 * B.r10;
  C.r11 = A.r10 * B.r01 + A.r11 * B.r11;
  C.tx = (A.r00 * B.tx + A.r01 * B.ty) + A.tx;
  C.ty = (A.r10 * B.tx + A.r11 * B.ty) + A.ty;
  return C;
}
inline Affine2 Inverse(const Affine2& A) {  // Eigen Affine inverse: general (cofactor) inverse of the linear part, t' = -Linv*t
  Affine2 I;
  const double det = A.r00 * A.r11 - A.r01 * A.r10;
  const double invdet = 1.0 / det;
  I.r00 = A.r11 * invdet;
  I.r01 = -A.r01 * invdet;
  I.r10 = -A.r10 * invdet;
  I.r11 = A.r00 * invdet;
  I.tx = -(I.r00 * A.tx + I.r01 * A.ty);
  I.ty = -(I.r10 * A.tx + I.r11 * A.ty);
  return I;
}
// Affine3dToVectorXYeZ (utils.cpp:115-122): yaw = T.linear().eulerAngles(0,1,2)[2]; for an exactly planar
// rotation Eigen 3.3's formula reduces to atan2(R10, R11).
inline void AffineToVector(const Affine2& T, double par[3]) {
  par[0] = T.tx;
  par[1] = T.ty;
  par[2] = std::atan2(T.r10, T.r11);
}
inline void Apply(const Affine2& T, double x, double y, double& ox, double& oy) {
  ox = (T.r00 * x + T.r01 * y) + T.tx;
  oy = (T.r10 * x + T.r11 * y) + T.ty;
}

// =============================================================================================
// (a5) Compensate                                          utils.cpp:96-113, utils.h:28-32, utils.cpp:130-146
// =============================================================================================
inline double GetRelTimeStamp(const double x, const double y, const bool ccw) {
  double a = std::atan2(y, x);
  double d = ((a > 0.00001 ? a : (2 * M_PI + a)) / (2 * M_PI));
  return ccw ? -(d - 0.5) : (d - 0.5);
}
inline void Compensate(Cloud& cloud, const double mot[3], bool ccw) {
  for (size_t i = 0; i < cloud.size(); i++) {
    const PointXYZI p = cloud[i];
    const double d = GetRelTimeStamp(p.x, p.y, ccw);
    const double s_1 = std::sin(d * mot[2]);
    const double c_1 = std::cos(d * mot[2]);
    const double tx = d * mot[0], ty = d * mot[1];
    const double px = p.x, py = p.y;
    cloud[i].x = (float)((c_1 * px + (-s_1) * py) + tx);  // R*peig + t  (utils.cpp:103)
    cloud[i].y = (float)((s_1 * px + c_1 * py) + ty);
  }
}

// =============================================================================================
// PCL 1.10 VoxelGrid<PointXYZI> (downsample_all_data = true) restated     pointnormal.cpp:277-280
// =============================================================================================
enum VoxelOrder {
  VOXEL_ORDER_STABLE = 0,    // [DEV-1] points of a voxel summed in ascending cloud index
  VOXEL_ORDER_STD_SORT = 1   // PCL's std::sort (libstdc++ introsort, unstable) order within a voxel
};
struct VoxelGridResult {
  Cloud centroids;                   // ascending voxel index
  int min_b[3], max_b[3], div_b[3];  // PCL min_b_/max_b_/div_b_
  float inv_leaf;
};
inline bool VoxelGrid(const Cloud& in, float leaf, VoxelGridResult& out, VoxelOrder order = VOXEL_ORDER_STABLE) {
  out.centroids.clear();
  if (in.empty()) return false;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones()/leaf_size_
  out.inv_leaf = inv;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};  // getMinMax3D
  for (const auto& p : in) {
    const float v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], v[a]);
      mx[a] = std::max(mx[a], v[a]);
    }
  }
  long long dxyz = 1;
  for (int a = 0; a < 3; a++) dxyz *= (long long)((mx[a] - mn[a]) * inv) + 1;
  if (dxyz > (long long)std::numeric_limits<int32_t>::max()) return false;  // "leaf size too small" -> PCL copies input; treated as failure
  for (int a = 0; a < 3; a++) {
    out.min_b[a] = (int)std::floor(mn[a] * inv);
    out.max_b[a] = (int)std::floor(mx[a] * inv);
    out.div_b[a] = out.max_b[a] - out.min_b[a] + 1;
  }
  const int mul[3] = {1, out.div_b[0], out.div_b[0] * out.div_b[1]};
  struct cloud_point_index_idx {
    unsigned int idx, cloud_point_index;
    bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
  };
  std::vector<cloud_point_index_idx> index_vector;
  index_vector.reserve(in.size());
  for (size_t i = 0; i < in.size(); i++) {
    const int ijk0 = (int)(std::floor(in[i].x * inv) - (float)out.min_b[0]);
    const int ijk1 = (int)(std::floor(in[i].y * inv) - (float)out.min_b[1]);
    const int ijk2 = (int)(std::floor(in[i].z * inv) - (float)out.min_b[2]);
    const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    index_vector.push_back({(unsigned)idx, (unsigned)i});
  }
  if (order == VOXEL_ORDER_STD_SORT)
    std::sort(index_vector.begin(), index_vector.end());
  else
    std::stable_sort(index_vector.begin(), index_vector.end());
  size_t first = 0;
  while (first < index_vector.size()) {
    size_t last = first + 1;
    while (last < index_vector.size() && index_vector[last].idx == index_vector[first].idx) ++last;
    float sx = 0, sy = 0, sz = 0, si = 0;  // pcl::CentroidPoint accumulators (AccumulatorXYZ, AccumulatorIntensity): float sums
    for (size_t li = first; li < last; ++li) {
      const PointXYZI& p = in[index_vector[li].cloud_point_index];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
    }
    const float n = (float)(last - first);
    PointXYZI c;
    c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
    out.centroids.push_back(c);
    first = last;
  }
  return true;
}

// =============================================================================================
// FLANN 1.9.1 radius search as used through pcl::search::KdTree::radiusSearchT   pointnormal.cpp:269,291
// =============================================================================================
// Set = { i : (dx*dx + dy*dy) + dz*dz  <  (float)(r*r) } in float (L2_Simple, strict <), returned sorted
// by (distance, index) (RadiusResultSet + std::sort of DistanceIndex).  The kd-tree only prunes; the set and
// order are those of exhaustive search, which is what is restated — over a uniform bucket grid for speed.
class RadiusSearcher {
 public:
  void Build(const Cloud& c, float radius) {
    cloud_ = &c;
    cell_ = radius > 0 ? radius : 1.0f;
    minx_ = miny_ = FLT_MAX;
    float maxx = -FLT_MAX, maxy = -FLT_MAX;
    for (const auto& p : c) {
      minx_ = std::min(minx_, p.x); miny_ = std::min(miny_, p.y);
      maxx = std::max(maxx, p.x); maxy = std::max(maxy, p.y);
    }
    nx_ = (int)std::floor((maxx - minx_) / cell_) + 1;
    ny_ = (int)std::floor((maxy - miny_) / cell_) + 1;
    if ((long long)nx_ * ny_ > (1 << 24)) { nx_ = ny_ = 1; cell_ = FLT_MAX; }
    start_.assign((size_t)nx_ * ny_ + 1, 0);
    for (const auto& p : c) start_[Bucket(p.x, p.y) + 1]++;
    for (size_t i = 1; i < start_.size(); i++) start_[i] += start_[i - 1];
    order_.resize(c.size());
    std::vector<int> fill(start_.begin(), start_.end() - 1);
    for (size_t i = 0; i < c.size(); i++) order_[fill[Bucket(c[i].x, c[i].y)]++] = (int)i;
  }
  // returns number of neighbours; indices sorted by (dist, index)
  int Search(const PointXYZI& q, double radius, std::vector<int>& idx, std::vector<float>& sqd) const {
    const float r2 = (float)(radius * radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius*radius)
    tmp_.clear();
    int bx0 = Clamp((int)std::floor((q.x - (float)radius - minx_) / cell_) - 1, nx_);
    int bx1 = Clamp((int)std::floor((q.x + (float)radius - minx_) / cell_) + 1, nx_);
    int by0 = Clamp((int)std::floor((q.y - (float)radius - miny_) / cell_) - 1, ny_);
    int by1 = Clamp((int)std::floor((q.y + (float)radius - miny_) / cell_) + 1, ny_);
    for (int by = by0; by <= by1; by++)
      for (int bx = bx0; bx <= bx1; bx++) {
        const size_t b = (size_t)by * nx_ + bx;
        for (int s = start_[b]; s < start_[b + 1]; s++) {
          const int i = order_[s];
          const PointXYZI& p = (*cloud_)[i];
          const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
          float d = 0;  // L2_Simple: result += diff*diff per dimension
          d += dx * dx; d += dy * dy; d += dz * dz;
          if (d < r2) tmp_.push_back(std::make_pair(d, i));
        }
      }
    std::sort(tmp_.begin(), tmp_.end());  // DistanceIndex::operator<: (dist, index)
    idx.resize(tmp_.size()); sqd.resize(tmp_.size());
    for (size_t k = 0; k < tmp_.size(); k++) { sqd[k] = tmp_[k].first; idx[k] = tmp_[k].second; }
    return (int)tmp_.size();
  }
 private:
  static int Clamp(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }
  size_t Bucket(float x, float y) const {
    int bx = Clamp((int)std::floor((x - minx_) / cell_), nx_);
    int by = Clamp((int)std::floor((y - miny_) / cell_), ny_);
    return (size_t)by * nx_ + bx;
  }
  const Cloud* cloud_ = nullptr;
  float cell_ = 1, minx_ = 0, miny_ = 0;
  int nx_ = 1, ny_ = 1;
  std::vector<int> start_, order_;
  mutable std::vector<std::pair<float, int>> tmp_;
};

// =============================================================================================
// Eigen 3.3.7 SelfAdjointEigenSolver<Matrix2d>::compute (iterative path) restated   pointnormal.cpp:39-45
// =============================================================================================
// Reads the lower triangle (m00, m10, m11) only.  Scale by max|coeff|, (trivial) tridiagonalisation,
// implicit symmetric QR steps with Wilkinson shift, ascending sort with eigenvector swap, rescale.
struct Eig2 {
  double eval[2];     // ascending
  double evec[2][2];  // evec[row][col]; column j belongs to eval[j]
};
namespace detail {
inline double eig_hypot(double x, double y) {  // Eigen numext::hypot (3.3.x generic implementation)
  double ax = std::fabs(x), ay = std::fabs(y), p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.0) return 0.0;
  return p * std::sqrt(1.0 + qp * qp);
}
inline void makeGivens(double p, double q, double& c, double& s) {  // Eigen JacobiRotation::makeGivens (real)
  if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
  else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
  else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p; double u = std::sqrt(1.0 + t * t); if (p < 0.0) u = -u;
    c = 1.0 / u; s = -t * c;
  } else {
    double t = p / q; double u = std::sqrt(1.0 + t * t); if (q < 0.0) u = -u;
    s = -1.0 / u; c = -t * s;
  }
}
}  // namespace detail
inline Eig2 SelfAdjointEig2(double m00, double m10, double m11) {
  Eig2 R;
  double scale = std::max(std::fabs(m00), std::max(std::fabs(m10), std::fabs(m11)));
  if (scale == 0.0) scale = 1.0;
  double diag[2] = {m00 / scale, m11 / scale};
  double sub = m10 / scale;
  double Q[2][2] = {{1, 0}, {0, 1}};
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  const double considerAsZero = std::numeric_limits<double>::min();
  const int maxIterations = 30;
  int iter = 0;
  while (true) {
    if (std::fabs(sub) <= (std::fabs(diag[0]) + std::fabs(diag[1])) * precision || std::fabs(sub) <= considerAsZero) sub = 0.0;
    if (sub == 0.0) break;
    iter++;
    if (iter > maxIterations * 2) break;
    // tridiagonal_qr_step(start=0, end=1)
    const double td = (diag[0] - diag[1]) * 0.5;
    const double e = sub;
    double mu = diag[1];
    if (td == 0.0) mu -= std::fabs(e);
    else if (e != 0.0) {
      const double e2 = e * e;
      const double h = detail::eig_hypot(td, e);
      if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    const double x = diag[0] - mu;
    const double z = sub;
    double c, s;
    detail::makeGivens(x, z, c, s);
    const double sdk = s * diag[0] + c * sub;
    const double dkp1 = s * sub + c * diag[1];
    diag[0] = c * (c * diag[0] - s * sub) - s * (c * sub - s * diag[1]);
    diag[1] = s * sdk + c * dkp1;
    sub = c * sdk - s * dkp1;
    for (int i = 0; i < 2; i++) {  // Q = Q * G  (apply_rotation_in_the_plane with j.transpose())
      const double xi = Q[i][0], yi = Q[i][1];
      Q[i][0] = c * xi - s * yi;
      Q[i][1] = s * xi + c * yi;
    }
  }
  if (diag[1] < diag[0]) {  // ascending sort + column swap
    std::swap(diag[0], diag[1]);
    std::swap(Q[0][0], Q[0][1]);
    std::swap(Q[1][0], Q[1][1]);
  }
  R.eval[0] = diag[0] * scale; R.eval[1] = diag[1] * scale;
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) R.evec[i][j] = Q[i][j];
  return R;
}

// =============================================================================================
// (a7,a8) cell::cell + cell::ComputeNormal                 pointnormal.cpp:7-63, pointnormal.h:66-73
// =============================================================================================
struct Cell {
  double u[2] = {0, 0};                    // u_
  double cov[2][2] = {{0.1, 0}, {0, 0.1}}; // cov_ (row, col)
  double scale = 0;                        // scale_ ("planarity")
  double snormal[2] = {0, 0}, orth_normal[2] = {0, 0};
  double lambda_min = 0, lambda_max = 0;
  double sum_intensity = 0, avg_intensity = 0;
  uint64_t Nsamples = 0;
  bool valid = false;
};
inline Cell MakeCell(const Cloud& input, const std::vector<int>& nn, bool weight_intensity, const double origin[2]) {
  Cell c;
  const size_t N = nn.size();
  c.Nsamples = N;
  std::vector<double> w(N), x0(N), x1(N);
  for (size_t i = 0; i < N; i++) {
    x0[i] = input[nn[i]].x;
    x1[i] = input[nn[i]].y;
    w[i] = weight_intensity ? std::max(input[nn[i]].intensity - 60.0, 0.0) : 1.0;  // :15 (60 hard-coded)
  }
  double sum = 0;  // w.sum() — weights are integer-valued so any summation order is exact  (:18)
  for (size_t i = 0; i < N; i++) sum += w[i];
  c.sum_intensity = sum;
  c.avg_intensity = sum / N;
  for (size_t i = 0; i < N; i++) w[i] = w[i] / sum;  // :21
  for (size_t i = 0; i < N; i++) {  // :23-24 sequential
    c.u[0] += w[i] * x0[i];
    c.u[1] += w[i] * x1[i];
  }
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
  for (size_t i = 0; i < N; i++) {  // :26-33  cov_ = x^T * (w .* x); [DEV-2] summed in neighbour order
    const double d0 = x0[i] - c.u[0], d1 = x1[i] - c.u[1];
    const double xw0 = w[i] * d0, xw1 = w[i] * d1;
    c00 += d0 * xw0; c01 += d0 * xw1; c10 += d1 * xw0; c11 += d1 * xw1;
  }
  c.cov[0][0] = c00; c.cov[0][1] = c01; c.cov[1][0] = c10; c.cov[1][1] = c11;
  // ComputeNormal :37-63
  const Eig2 es = SelfAdjointEig2(c.cov[0][0], c.cov[1][0], c.cov[1][1]);
  c.snormal[0] = es.evec[0][0]; c.snormal[1] = es.evec[1][0];
  c.orth_normal[0] = es.evec[0][1]; c.orth_normal[1] = es.evec[1][1];
  c.lambda_min = es.eval[0];
  c.lambda_max = es.eval[1];
  const double condition_number = std::fabs(c.lambda_max / c.lambda_min);
  const double determinant = c.lambda_max * c.lambda_min;
  const double det_tolerance = 0.00001;
  const bool cov_reasonable = (condition_number <= 10000) && (determinant > det_tolerance) && c.lambda_min > 0 && c.lambda_max > 0;
  c.scale = std::log(1.0 + condition_number / 2);
  const double pox = origin[0] - c.u[0], poy = origin[1] - c.u[1];
  if (c.snormal[0] * pox + c.snormal[1] * poy < 0) { c.snormal[0] = -c.snormal[0]; c.snormal[1] = -c.snormal[1]; }
  c.valid = cov_reasonable;
  return c;
}

// =============================================================================================
// (a6, a9) MapPointNormal                                  pointnormal.cpp:65-90, 151-162, 238-254, 265-297
// =============================================================================================
class MapPointNormal {
 public:
  MapPointNormal() {}
  MapPointNormal(const Cloud& cld, float radius, const double origin[2], bool weight_intensity, double downsample_factor = 1.0,
                 VoxelOrder order = VOXEL_ORDER_STABLE)
      : radius_(radius) {
    ComputeNormals(cld, origin, weight_intensity, downsample_factor, order);
    ComputeSearchTreeFromCells();
  }
  // from precomputed cells (e.g. produced by the GPU path) — used by parity tests
  explicit MapPointNormal(const std::vector<Cell>& cs, float radius) : cells(cs), radius_(radius) { ComputeSearchTreeFromCells(); }

  size_t GetSize() const { return cells.size(); }
  const Cell& GetCell(size_t i) const { return cells[i]; }

  // GetClosestIdx :238-254 — 1-NN over cell means narrowed to float (pcl::PointXY), float squared L2
  // ((0+dx*dx)+dy*dy, FLANN L2_Simple), accepted iff d2 < d*d (float vs double compare).  The reference takes
  // the global 1-NN and then tests its distance; a neighbour that passes the test lies within d <= gcell_ of
  // the query, i.e. inside the 3x3 bucket block around it, and nothing outside that block can be closer than
  // an accepted one — so searching the block gives the same answer.  Exact float ties -> lowest index
  // (FLANN's tie order is traversal-dependent; ties between distinct cell means do not occur in practice).
  int GetClosestIdx(double px, double py, double d) const {
    if (cells.empty()) return -1;
    if (d > (double)gcell_) return GetClosestIdxBrute(px, py, d);
    const float qx = (float)px, qy = (float)py;
    int best = -1;
    float bestd = FLT_MAX;
    const int cbx = (int)std::floor((qx - gminx_) / gcell_), cby = (int)std::floor((qy - gminy_) / gcell_);
    for (int by = cby - 1; by <= cby + 1; by++) {
      if (by < 0 || by >= gny_) continue;
      for (int bx = cbx - 1; bx <= cbx + 1; bx++) {
        if (bx < 0 || bx >= gnx_) continue;
        const size_t b = (size_t)by * gnx_ + bx;
        for (int s = gstart_[b]; s < gstart_[b + 1]; s++) {
          const int i = gorder_[s];
          const float dx = qx - mx_[i], dy = qy - my_[i];
          float dd = 0;
          dd += dx * dx; dd += dy * dy;
          if (dd < bestd || (dd == bestd && i < best)) { bestd = dd; best = i; }
        }
      }
    }
    if (best >= 0 && (double)bestd < d * d) return best;
    return -1;
  }
  // exhaustive variant (used by tests to validate the bucket search)
  int GetClosestIdxBrute(double px, double py, double d) const {
    const float qx = (float)px, qy = (float)py;
    int best = -1;
    float bestd = FLT_MAX;
    for (size_t i = 0; i < cells.size(); i++) {
      const float dx = qx - mx_[i], dy = qy - my_[i];
      float dd = 0;
      dd += dx * dx; dd += dy * dy;
      if (dd < bestd) { bestd = dd; best = (int)i; }
    }
    if (best >= 0 && (double)bestd < d * d) return best;
    return -1;
  }

  std::vector<Cell> cells;
  int n_samples = 0;  // voxel-grid sample points examined (diagnostic)

 private:
  void ComputeNormals(const Cloud& input, const double origin[2], bool weight_intensity, double downsample_factor, VoxelOrder order) {
    if (input.empty()) return;  // reference: exit(0) (pointnormal.cpp:72-75); the oracle returns an empty map
    RadiusSearcher kdt_input;
    kdt_input.Build(input, radius_);
    VoxelGridResult vg;
    const float leaf = (float)(radius_ / downsample_factor);  // :279 float/double -> double -> float argument
    if (!VoxelGrid(input, leaf, vg, order)) return;
    n_samples = (int)vg.centroids.size();
    std::vector<int> idx;
    std::vector<float> sqd;
    for (size_t i = 0; i < vg.centroids.size(); i++) {
      if (kdt_input.Search(vg.centroids[i], radius_, idx, sqd) >= 6) {  // :291
        Cell c = MakeCell(input, idx, weight_intensity, origin);
        if (c.valid) cells.push_back(c);
      }
    }
  }
  void ComputeSearchTreeFromCells() {  // :151-162
    const size_t n = cells.size();
    mx_.resize(n); my_.resize(n);
    gminx_ = gminy_ = FLT_MAX;
    float maxx = -FLT_MAX, maxy = -FLT_MAX;
    for (size_t i = 0; i < n; i++) {
      mx_[i] = (float)cells[i].u[0];  // pcl::PointXY
      my_[i] = (float)cells[i].u[1];
      gminx_ = std::min(gminx_, mx_[i]); gminy_ = std::min(gminy_, my_[i]);
      maxx = std::max(maxx, mx_[i]); maxy = std::max(maxy, my_[i]);
    }
    gcell_ = 4.0f;
    if (n == 0) { gnx_ = gny_ = 1; gminx_ = gminy_ = 0; gstart_.assign(2, 0); gorder_.clear(); return; }
    gnx_ = (int)std::floor((maxx - gminx_) / gcell_) + 1;
    gny_ = (int)std::floor((maxy - gminy_) / gcell_) + 1;
    gstart_.assign((size_t)gnx_ * gny_ + 1, 0);
    auto bucket = [&](float x, float y) {
      int bx = std::min(gnx_ - 1, std::max(0, (int)std::floor((x - gminx_) / gcell_)));
      int by = std::min(gny_ - 1, std::max(0, (int)std::floor((y - gminy_) / gcell_)));
      return (size_t)by * gnx_ + bx;
    };
    for (size_t i = 0; i < n; i++) gstart_[bucket(mx_[i], my_[i]) + 1]++;
    for (size_t i = 1; i < gstart_.size(); i++) gstart_[i] += gstart_[i - 1];
    gorder_.resize(n);
    std::vector<int> fill(gstart_.begin(), gstart_.end() - 1);
    for (size_t i = 0; i < n; i++) gorder_[fill[bucket(mx_[i], my_[i])]++] = (int)i;
  }
  float radius_ = 0;
  std::vector<float> mx_, my_;
  float gcell_ = 4.0f, gminx_ = 0, gminy_ = 0;
  int gnx_ = 1, gny_ = 1;
  std::vector<int> gstart_, gorder_;
};

}  // namespace tbv_oracle

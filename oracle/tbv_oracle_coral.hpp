// tbv_oracle_coral.hpp — TEST INFRASTRUCTURE: CPU restatement of CorAl's radar alignment quality.
//
// Follows coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:8-229 (CorAlRadarQuality: GetNearby, Covariance,
// ComputeEntropy, constructor) as TBV calls it — ScanLearningInterface::getCorAlQualityMeasure,
// coral_alignment_quality/src/alignment_checker/alignmentinterface.cpp:437-454: peaks clouds, radius 1.0, ent_cfg = any,
// weight_res_intensity = false, output_overlap = true.  Third-party arithmetic restated: pcl::transformPointCloud (PCL 1.10, double
// transform, float result), pcl::KdTreeFLANN<PointXY>::radiusSearch (FLANN 1.9.1: float L2, strict <, results sorted by (distance,
// index)), Eigen reductions as left-to-right sums ([DEV-5]).  Never included by the product.
#pragma once
#include <cmath>
#include <vector>

#include "tbv_oracle.hpp"
#include "tbv_oracle_loop.hpp"

namespace tbv_oracle {

struct CoralParams {
  double radius = 1.0;           // alignmentinterface.cpp:444
  int weight_res_intensity = 0;  // alignmentinterface.cpp:444
  int overlap_req = 1;           // AlignmentQuality.h:244
};
struct CoralResult {
  double joint = 0, sep = 0, overlap = 0;  // quality_ = {joint_, sep_, overlap_}  (AlignmentQuality.cpp:193-199)
  int count_valid = 0, merged_size = 0, valid = 0;
};

namespace coral_detail {
// CorAlRadarQuality::Covariance (AlignmentQuality.cpp:30-51): rows are (x, y) doubles
inline bool Covariance(std::vector<double>& x, double cov[4], double mean[2]) {
  const size_t rows = x.size() / 2;
  if (rows <= 2) return false;
  double sx = 0, sy = 0;
  for (size_t i = 0; i < rows; i++) { sx += x[2 * i]; sy += x[2 * i + 1]; }
  mean[0] = sx / (double)rows; mean[1] = sy / (double)rows;
  for (size_t i = 0; i < rows; i++) { x[2 * i] = x[2 * i] - mean[0]; x[2 * i + 1] = x[2 * i + 1] - mean[1]; }
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0;  // covSum = x^T x
  for (size_t i = 0; i < rows; i++) {
    c00 += x[2 * i] * x[2 * i]; c01 += x[2 * i] * x[2 * i + 1];
    c10 += x[2 * i + 1] * x[2 * i]; c11 += x[2 * i + 1] * x[2 * i + 1];
  }
  const float n = (float)rows;                  // "float n = x.rows()"
  const double den = n - 1.0;
  cov[0] = c00 * 1.0 / den; cov[1] = c01 * 1.0 / den; cov[2] = c10 * 1.0 / den; cov[3] = c11 * 1.0 / den;
  return true;
}
// CorAlRadarQuality::ComputeEntropy (AlignmentQuality.cpp:79-98)
inline bool ComputeEntropy(const double cs[4], const double cj[4], double& sep_entropy, double& joint_entropy) {
  const double det_j = cj[0] * cj[3] - cj[1] * cj[2];
  const double det_s = cs[0] * cs[3] - cs[1] * cs[2];
  if (std::isnan(det_s) || std::isnan(det_j)) return false;
  sep_entropy = 1.0 / 2.0 * std::log(2.0 * M_PI * std::exp(1.0) * det_s + 0.00000001);
  joint_entropy = 1.0 / 2.0 * std::log(2.0 * M_PI * std::exp(1.0) * det_j + 0.00000001);
  return !(std::isnan(sep_entropy) || std::isnan(joint_entropy));
}
}  // namespace coral_detail

// CorAlRadarQuality::CorAlRadarQuality (AlignmentQuality.cpp:99-229).  src_local / ref_local: the scans' peaks clouds in their own
// frames; Tsrc = src->GetAffine() * Toffset, Tref = ref->GetAffine() (:105-106).  per_point (optional, merged_size x 3): sep, joint,
// valid for every point, src points first (sep_res_, joint_res_, sep_valid before the weighting loop).
inline CoralResult CorAlRadarQuality(const Cloud& src_local, const Cloud& ref_local, const Affine2& Tsrc, const Affine2& Tref, const CoralParams& par,
                                     std::vector<double>* per_point = nullptr) {
  using namespace coral_detail;
  Cloud src, ref;
  TransformCloud(src_local, Tsrc, src);
  TransformCloud(ref_local, Tref, ref);
  for (auto& p : src) p.z = 0;   // pcl3dto2d (Utils.cpp:200-213): the kd-trees are over PointXY
  for (auto& p : ref) p.z = 0;
  RadiusSearcher kd_src, kd_ref;
  kd_src.Build(src, (float)par.radius);
  kd_ref.Build(ref, (float)par.radius);
  const size_t merged_size = src.size() + ref.size();
  std::vector<double> sep_res(merged_size, 100.0), joint_res(merged_size, 100.0), intensity(merged_size, 0.0);
  std::vector<char> sep_valid(merged_size, 0);
  for (size_t i = 0; i < src.size(); i++) intensity[i] = src[i].intensity;
  for (size_t i = 0; i < ref.size(); i++) intensity[src.size() + i] = ref[i].intensity;
  std::vector<int> is, ir;
  std::vector<float> ds, dr;
  auto nearby = [&](const PointXYZI& q, std::vector<double>& msrc, std::vector<double>& mref, std::vector<double>& merged) {  // GetNearby :8-29
    kd_src.Search(q, par.radius, is, ds);
    kd_ref.Search(q, par.radius, ir, dr);
    msrc.clear(); mref.clear(); merged.clear();
    for (int k : is) { msrc.push_back(src[k].x); msrc.push_back(src[k].y); merged.push_back(src[k].x); merged.push_back(src[k].y); }
    for (int k : ir) { mref.push_back(ref[k].x); mref.push_back(ref[k].y); merged.push_back(ref[k].x); merged.push_back(ref[k].y); }
  };
  std::vector<double> msrc, mref, mj;
  long index = -1;
  for (const auto& q : src) {  // :131-151
    index++;
    nearby(q, msrc, mref, mj);
    if ((int)(mref.size() / 2) < par.overlap_req) continue;
    double cs[4], cj[4], ms[2], mjn[2];
    if (Covariance(msrc, cs, ms) && Covariance(mj, cj, mjn)) {
      double se, je;
      if (ComputeEntropy(cs, cj, se, je)) { sep_res[index] = se; joint_res[index] = je; sep_valid[index] = 1; }
    }
  }
  for (const auto& q : ref) {  // :153-172
    index++;
    nearby(q, msrc, mref, mj);
    if ((int)(msrc.size() / 2) < par.overlap_req) continue;
    double cs[4], cj[4], ms[2], mjn[2];
    if (Covariance(mref, cs, ms) && Covariance(mj, cj, mjn)) {
      double se, je;
      if (ComputeEntropy(cs, cj, se, je)) { sep_res[index] = se; joint_res[index] = je; sep_valid[index] = 1; }
    }
  }
  if (per_point) {
    per_point->assign(merged_size * 3, 0.0);
    for (size_t i = 0; i < merged_size; i++) { (*per_point)[3 * i] = sep_res[i]; (*per_point)[3 * i + 1] = joint_res[i]; (*per_point)[3 * i + 2] = sep_valid[i]; }
  }
  CoralResult R;
  double w_sum = 0, joint = 0, sep = 0;
  int count_valid = 0;
  for (size_t i = 0; i < merged_size; i++) {  // :174-185
    if (sep_valid[i]) {
      const double w = par.weight_res_intensity ? intensity[i] : 1.0;
      w_sum += w;
      joint_res[i] = w * joint_res[i];
      sep_res[i] = w * sep_res[i];
      joint += joint_res[i]; sep += sep_res[i];
      count_valid++;
    }
  }
  if (count_valid > 0) { sep /= w_sum; joint /= w_sum; }
  R.joint = joint; R.sep = sep;
  R.count_valid = count_valid; R.merged_size = (int)merged_size;
  R.overlap = count_valid / ((double)merged_size);
  R.valid = !(R.overlap < 0.1);
  return R;
}

}  // namespace tbv_oracle

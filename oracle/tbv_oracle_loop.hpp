// =============================================================================================
// tbv_oracle_loop.hpp — CPU ORACLE, part 3: Scan-Context descriptor / keys / distance / candidate search and
// pose-graph residual + Jacobian + normal-equation assembly.      *** TEST INFRASTRUCTURE ONLY ***
// See tbv_oracle.hpp for the rules on who may use this and for the "parity unpinned" statement.
// Eigen reductions (mean(), norm(), dot()) are restated as plain left-to-right sums  [DEV-5].
// =============================================================================================
#pragma once
#include "tbv_oracle.hpp"

namespace tbv_oracle {

// =============================================================================================
// (a17) RSCManager::MakeRadarCloudContext + xy2theta
//       place_recognition_radar/src/place_recognition_radar/RadarScancontext.cpp:59-131, Scancontext.cpp:62-77
// =============================================================================================
struct SCParams {
  int PC_NUM_RING = 40, PC_NUM_SECTOR = 120;
  double PC_MAX_RADIUS = 80.0;
  double SEARCH_RATIO = 0.1;
  int NUM_CANDIDATES_FROM_TREE = 10;
  int N_CANDIDATES = 1;
  double odom_sigma_error = 0.05;
  bool odometry_coupled_closure = true, augment_sc = true;
  double no_point = 0;
  int desc_function = 0;  // 0 = "sum", 1 = "max"
  double desc_divider = 1;
  double DISTANCE_EXCLUDE_RECENT = 10.0;
};
inline float xy2theta(const float& _x, const float& _y) {  // float atan overload, product in double, narrowed on return
  if (_x >= 0 & _y >= 0) return (float)((180 / M_PI) * std::atan(_y / _x));
  if (_x < 0 & _y >= 0) return (float)(180 - ((180 / M_PI) * std::atan(_y / (-_x))));
  if (_x < 0 & _y < 0) return (float)(180 + ((180 / M_PI) * std::atan(_y / _x)));
  if (_x >= 0 & _y < 0) return (float)(360 - ((180 / M_PI) * std::atan((-_y) / _x)));
  return 0;
}
// desc is column-major (Eigen MatrixXd default): desc[col * rings + row]
inline void MakeRadarCloudContext(const Cloud& cloud, const SCParams& par, std::vector<double>& desc) {
  const int R = par.PC_NUM_RING, S = par.PC_NUM_SECTOR;
  const int NO_POINT = -1000;
  desc.assign((size_t)R * S, (double)NO_POINT);
  for (size_t pt_idx = 0; pt_idx < cloud.size(); pt_idx++) {
    const float px = cloud[pt_idx].x, py = cloud[pt_idx].y, pi = cloud[pt_idx].intensity;
    const float azim_range = std::sqrt(px * px + py * py);
    const float azim_angle = xy2theta(px, py);
    if (azim_range > par.PC_MAX_RADIUS) continue;
    const int ring_idx = std::max(std::min(R, int(std::ceil((azim_range / par.PC_MAX_RADIUS) * R))), 1);
    const int sctor_idx = std::max(std::min(S, int(std::ceil((azim_angle / 360.0) * S))), 1);
    double& d = desc[(size_t)(sctor_idx - 1) * R + (ring_idx - 1)];
    if (d == NO_POINT) d = pi;
    else if (par.desc_function == 0) d += pi;
    else d = std::max(d, (double)pi);
  }
  // "Divison before no_point check, thus no_point is not set" (:113-125): untouched bins end up -1000/divider
  for (auto& d : desc) d = d / par.desc_divider;
  for (auto& d : desc) if (d == NO_POINT) d = par.no_point;
}
// pcl::transformPointCloud(cloud, out, Eigen::Affine3d) (PCL 1.10 detail::Transformer<double>::se3): each coordinate is
// computed in double as m(r,0)*x + m(r,1)*y + m(r,2)*z + m(r,3), left to right, then cast to float.
inline void TransformCloud(const Cloud& in, const Affine2& T, Cloud& out) {
  out.resize(in.size());
  for (size_t i = 0; i < in.size(); i++) {
    const double x = in[i].x, y = in[i].y, z = in[i].z;
    out[i].x = (float)(T.r00 * x + T.r01 * y + 0.0 * z + T.tx);
    out[i].y = (float)(T.r10 * x + T.r11 * y + 0.0 * z + T.ty);
    out[i].z = (float)(0.0 * x + 0.0 * y + 1.0 * z + 0.0);
    out[i].intensity = in[i].intensity;
  }
}
inline void TranslateCloud(const Cloud& in, double tx, double ty, Cloud& out) {  // augmentations, RadarScancontext.cpp:162-179
  Affine2 T;
  T.tx = tx; T.ty = ty;
  TransformCloud(in, T, out);
}

// (a18) ring key (row means, narrowed to float) / sector key (column means)   Scancontext.cpp:239-268, :103-107
inline void MakeRingkey(const std::vector<double>& desc, int R, int S, std::vector<float>& key) {
  key.resize(R);
  for (int r = 0; r < R; r++) {
    double s = 0;
    for (int c = 0; c < S; c++) s += desc[(size_t)c * R + r];
    key[r] = (float)(s / S);
  }
}
inline void MakeSectorkey(const std::vector<double>& desc, int R, int S, std::vector<double>& key) {
  key.resize(S);
  for (int c = 0; c < S; c++) {
    double s = 0;
    for (int r = 0; r < R; r++) s += desc[(size_t)c * R + r];
    key[c] = s / R;
  }
}

// (a20) distDirectSC / fastAlignUsingVkey / distanceBtnScanContext            Scancontext.cpp:80-189
// sc2 shifted right by `shift` columns: shifted.col((c + shift) % S) = sc2.col(c)
inline double distDirectSC(const double* sc1, const double* sc2, int R, int S, int shift) {
  int num_eff_cols = 0;
  double sum_sector_similarity = 0;
  for (int col = 0; col < S; col++) {
    const double* c1 = sc1 + (size_t)col * R;
    const double* c2 = sc2 + (size_t)((col - shift + S) % S) * R;
    double n1 = 0, n2 = 0, dot = 0;
    for (int r = 0; r < R; r++) { n1 += c1[r] * c1[r]; n2 += c2[r] * c2[r]; dot += c1[r] * c2[r]; }
    n1 = std::sqrt(n1); n2 = std::sqrt(n2);
    if (n1 == 0 | n2 == 0) continue;
    sum_sector_similarity = sum_sector_similarity + dot / (n1 * n2);
    num_eff_cols = num_eff_cols + 1;
  }
  num_eff_cols = std::max(num_eff_cols, 1);
  return 1.0 - sum_sector_similarity / num_eff_cols;
}
inline int fastAlignUsingVkey(const std::vector<double>& vkey1, const std::vector<double>& vkey2) {
  const int S = (int)vkey1.size();
  int argmin = 0;
  double mn = 10000000;
  for (int shift = 0; shift < S; shift++) {
    double s = 0;
    for (int c = 0; c < S; c++) {
      const double d = vkey1[c] - vkey2[(c - shift + S) % S];
      s += d * d;
    }
    const double cur = std::sqrt(s);
    if (cur < mn) { argmin = shift; mn = cur; }
  }
  return argmin;
}
inline std::pair<double, int> distanceBtnScanContext(const std::vector<double>& sc1, const std::vector<double>& sc2, int R, int S, double search_ratio) {
  std::vector<double> vkey1, vkey2;
  MakeSectorkey(sc1, R, S, vkey1);
  MakeSectorkey(sc2, R, S, vkey2);
  const int argmin_vkey_shift = fastAlignUsingVkey(vkey1, vkey2);
  const int SEARCH_RADIUS = (int)std::round(0.5 * search_ratio * S);
  std::vector<int> space{argmin_vkey_shift};
  for (int ii = 1; ii < SEARCH_RADIUS + 1; ii++) {
    space.push_back((argmin_vkey_shift + ii + S) % S);
    space.push_back((argmin_vkey_shift - ii + S) % S);
  }
  std::sort(space.begin(), space.end());
  int argmin_shift = 0;
  double min_sc_dist = 10000000;
  for (int num_shift : space) {
    const double cur = distDirectSC(sc1.data(), sc2.data(), R, S, num_shift);
    if (cur < min_sc_dist) { argmin_shift = num_shift; min_sc_dist = cur; }
  }
  return std::make_pair(min_sc_dist, argmin_shift);
}

// (a19) RSCManager state, ExcludeAndUpdateLikelihood, OdometryNNSearch, detectLoopClosureID
//       RadarScancontext.cpp:133-221, 251-345
struct SCCandidate {
  double min_dist, min_dist_sc, min_dist_odom;
  float yaw_diff_rad;
  int nn_idx, argmin_shift;
  int aug_idx;        // which query (0 = identity, 1..4 = lateral offsets)
  double aug_xy[2];   // Taug translation
};
class RSCManager {
 public:
  explicit RSCManager(const SCParams& p) : par(p) {}
  struct Query { std::vector<double> desc; std::vector<float> ring_key; double off[2]; };

  void makeAndSaveScancontextAndKeysRadarCloud(const Cloud& cloud, const Affine2& Todom) {
    std::vector<double> sc;
    MakeRadarCloudContext(cloud, par, sc);
    std::vector<float> rk;
    MakeRingkey(sc, par.PC_NUM_RING, par.PC_NUM_SECTOR, rk);
    polarcontexts_.push_back(sc);
    polarcontext_invkeys_mat_.push_back(rk);
    current_and_augments_.clear();
    current_and_augments_.push_back({sc, rk, {0.0, 0.0}});
    ExcludeAndUpdateLikelihood(Todom);
    if (par.augment_sc) {
      const double aug[4][2] = {{0.0, -2.0}, {0.0, 2.0}, {0.0, -4.0}, {0.0, 4.0}};
      for (int a = 0; a < 4; a++) {
        Cloud augmented;
        TranslateCloud(cloud, aug[a][0], aug[a][1], augmented);
        std::vector<double> sca;
        MakeRadarCloudContext(augmented, par, sca);
        std::vector<float> rka;
        MakeRingkey(sca, par.PC_NUM_RING, par.PC_NUM_SECTOR, rka);
        current_and_augments_.push_back({sca, rka, {aug[a][0], aug[a][1]}});
      }
    }
  }
  void ExcludeAndUpdateLikelihood(const Affine2& Todom) {  // :183-221
    odom_poses_.push_back(Todom);
    if (odom_poses_.size() <= 2) NUM_EXCLUDE_RECENT = 2;
    else {
      double distance = 0.0;
      NUM_EXCLUDE_RECENT = 0;
      Affine2 Tprev = odom_poses_.back();
      for (int i = (int)odom_poses_.size() - 1; i >= 0 && distance < par.DISTANCE_EXCLUDE_RECENT; i--) {
        const Affine2 d = Mul(Inverse(Tprev), odom_poses_[i]);
        distance = distance + std::sqrt(d.tx * d.tx + d.ty * d.ty);
        Tprev = odom_poses_[i];
        NUM_EXCLUDE_RECENT++;
      }
    }
    double tpx = Todom.tx, tpy = Todom.ty;
    double odom_trav_distance = 0;
    const int idx_current = (int)odom_poses_.size() - 1;
    odom_similarity.assign(idx_current, 0.0);
    for (int i = idx_current - 1; i >= 0; i--) {
      const double tix = odom_poses_[i].tx, tiy = odom_poses_[i].ty;
      odom_trav_distance += std::sqrt((tpx - tix) * (tpx - tix) + (tpy - tiy) * (tpy - tiy));
      tpx = tix; tpy = tiy;
      const double odom_est_distance = std::sqrt((Todom.tx - tix) * (Todom.tx - tix) + (Todom.ty - tiy) * (Todom.ty - tiy));
      const double error = std::max(odom_est_distance - 5.0, 0.0);
      const double rel_error = error / odom_trav_distance;
      const double probability = std::exp(-rel_error * rel_error / (2 * par.odom_sigma_error * par.odom_sigma_error));
      odom_similarity[i] = 1.0 - probability;
    }
  }
  static double L2norm(const std::vector<float>& v1, const std::vector<float>& v2) {  // :251-258 (float accumulator, double err)
    float l2 = 0;
    for (size_t i = 0; i < v1.size(); i++) {
      const double err = (v1[i] - v2[i]);
      l2 += err * err;
    }
    return l2;
  }
  void OdometryNNSearch(std::vector<size_t>& candidate_indexes, const std::vector<float>& current_key) const {  // :259-284
    std::vector<float> curr_key = current_key;
    curr_key.push_back(0.0);
    const int idx_current = (int)polarcontext_invkeys_mat_.size() - 1;
    std::vector<std::pair<double, int>> dist_idx_vek;
    for (int idx = 0; idx < idx_current - 1 - NUM_EXCLUDE_RECENT; idx++) {
      std::vector<float> k = polarcontext_invkeys_mat_[idx];
      k.push_back(10 * odom_similarity[idx]);  // double -> float push_back
      dist_idx_vek.push_back(std::make_pair(L2norm(curr_key, k), idx));
    }
    std::sort(dist_idx_vek.begin(), dist_idx_vek.end());  // sorted insertion by (dist, idx) == full sort on the pair order
    for (size_t i = 0; i < dist_idx_vek.size() && (int)i < par.NUM_CANDIDATES_FROM_TREE; i++) candidate_indexes.push_back(dist_idx_vek[i].second);
  }
  std::vector<SCCandidate> detectLoopClosureID() const {  // :286-345 (odometry_coupled_closure path)
    std::vector<SCCandidate> similar;
    if ((int)polarcontext_invkeys_mat_.size() < NUM_EXCLUDE_RECENT + 1) return similar;
    for (size_t q = 0; q < current_and_augments_.size(); q++) {
      const Query& query = current_and_augments_[q];
      std::vector<size_t> cand;
      OdometryNNSearch(cand, query.ring_key);
      for (size_t ci = 0; ci < cand.size(); ci++) {
        const int cand_idx = (int)cand[ci];
        const std::pair<double, int> r = distanceBtnScanContext(query.desc, polarcontexts_[cand_idx], par.PC_NUM_RING, par.PC_NUM_SECTOR, par.SEARCH_RATIO);
        const double d_odom = par.odometry_coupled_closure ? odom_similarity[cand_idx] : 0;
        const double d = par.odometry_coupled_closure ? r.first + d_odom : r.first;
        const double unit = 360.0 / double(par.PC_NUM_SECTOR);
        const float deg = (float)(r.second * unit);
        const float align = (float)(deg * M_PI / 180.0);  // deg2rad(float)
        similar.push_back({d, r.first, d_odom, align, cand_idx, r.second, (int)q, {query.off[0], query.off[1]}});
        std::stable_sort(similar.begin(), similar.end(), [](const SCCandidate& a, const SCCandidate& b) { return a.min_dist < b.min_dist; });
        if ((int)similar.size() > par.N_CANDIDATES) similar.erase(similar.end() - 1);
      }
    }
    return similar;
  }
  SCParams par;
  int NUM_EXCLUDE_RECENT = 0;
  std::vector<std::vector<double>> polarcontexts_;
  std::vector<std::vector<float>> polarcontext_invkeys_mat_;
  std::vector<Affine2> odom_poses_;
  std::vector<double> odom_similarity;
  std::vector<Query> current_and_augments_;
};

// =============================================================================================
// (a21) PoseGraph3dErrorTerm + CeresLeastSquares::AddConstraintType
//       tbv_slam/include/tbv_slam/ceresoptimizer.h:51-95, tbv_slam/src/tbv_slam/ceresoptimizer.cpp:18-108
// =============================================================================================
// Node: p (x,y,z), q stored in Eigen coefficient order (x,y,z,w).  Residual (6) = L * [ q_a^-1 (p_b - p_a) - p_m ;
// 2 vec(q_m * (q_a^-1 q_b)^-1) ] with L = llt(Info).matrixL().  The AutoDiff Jacobian w.r.t. the ambient
// (p 3, q 4) blocks is restated analytically and multiplied by EigenQuaternionParameterization's 4x3 plus-Jacobian.
struct PGNode { double p[3]; double q[4]; };                 // q = (x,y,z,w)
struct PGConstraint { int id_begin, id_end; double p[3]; double q[4]; int type; double info[36]; };  // type 0 = odometry, 1 = loop
struct PGParams {
  double odom_vxx = 0.01, odom_vyy = 0.01, odom_vtt = 0.001, loop_scaling = 500000;
  bool replace_cov_by_identity = true;
  double loop_cauchy = 0.1;
};
namespace detail {
inline void QuatRotate(const double q[4], const double v[3], double o[3]) {  // Eigen Quaternion::_transformVector
  const double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  const double uvx = ux + ux, uvy = uy + uy, uvz = uz + uz;
  o[0] = v[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  o[1] = v[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  o[2] = v[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
inline void QuatMul(const double p[4], const double q[4], double o[4]) {  // Hamilton product, (x,y,z,w) storage
  o[0] = p[3] * q[0] + p[0] * q[3] + p[1] * q[2] - p[2] * q[1];
  o[1] = p[3] * q[1] - p[0] * q[2] + p[1] * q[3] + p[2] * q[0];
  o[2] = p[3] * q[2] + p[0] * q[1] - p[1] * q[0] + p[2] * q[3];
  o[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
}
// 6x6 Cholesky L (lower) of a symmetric positive definite matrix (row-major), Eigen LLT semantics (lower triangle read)
inline bool Chol6(const double A[36], double L[36]) {
  for (int i = 0; i < 36; i++) L[i] = 0;
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0)) return false;
    L[j * 6 + j] = std::sqrt(d);
    for (int i = j + 1; i < 6; i++) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = v / L[j * 6 + j];
    }
  }
  return true;
}
}  // namespace detail

// sqrt information per constraint (ceresoptimizer.cpp:83-100)
inline bool PGSqrtInformation(const PGConstraint& c, const PGParams& par, double L[36]) {
  double I[36];
  const double loop_scale_factor = (c.type == 1) ? 1.0 / par.loop_scaling : 1.0;
  if (par.replace_cov_by_identity) {
    const double diag[6] = {1.0 / par.odom_vxx, 1.0 / par.odom_vyy, 1, 1, 1, 1.0 / par.odom_vtt};
    for (int i = 0; i < 36; i++) I[i] = 0;
    for (int i = 0; i < 6; i++) I[i * 6 + i] = 1.0 * diag[i] * loop_scale_factor;
  } else {
    for (int i = 0; i < 36; i++) I[i] = c.info[i] * loop_scale_factor;
  }
  return detail::Chol6(I, L);
}

// residual (6) and tangent-space Jacobians Ja, Jb (6x6 row-major each: columns = [p(3) | rotation tangent(3)])
// BEFORE the loss correction.
inline void PGEvaluate(const PGNode& A, const PGNode& B, const PGConstraint& c, const double L[36], double r[6], double Ja[36], double Jb[36]) {
  using namespace detail;
  const double qa_inv[4] = {-A.q[0], -A.q[1], -A.q[2], A.q[3]};
  const double d[3] = {B.p[0] - A.p[0], B.p[1] - A.p[1], B.p[2] - A.p[2]};
  double p_ab[3];
  QuatRotate(qa_inv, d, p_ab);
  double q_ab[4];
  QuatMul(qa_inv, B.q, q_ab);
  const double q_ab_conj[4] = {-q_ab[0], -q_ab[1], -q_ab[2], q_ab[3]};
  double dq[4];
  QuatMul(c.q, q_ab_conj, dq);
  double e[6] = {p_ab[0] - c.p[0], p_ab[1] - c.p[1], p_ab[2] - c.p[2], 2.0 * dq[0], 2.0 * dq[1], 2.0 * dq[2]};
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += L[i * 6 + k] * e[k];
    r[i] = s;
  }
  if (!Ja || !Jb) return;
  // --- ambient derivatives of e -------------------------------------------------------------
  // p_ab = Rot(u, w; d) = d + 2w (u x d) + 2 u x (u x d), with u = -a_v, w = a_w
  const double u[3] = {qa_inv[0], qa_inv[1], qa_inv[2]}, w = qa_inv[3];
  double Rm[3][3];  // d p_ab / d d  = I + 2w [u]x + 2 [u]x^2
  {
    const double ux[3][3] = {{0, -u[2], u[1]}, {u[2], 0, -u[0]}, {-u[1], u[0], 0}};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double uu = 0;
        for (int k = 0; k < 3; k++) uu += ux[i][k] * ux[k][j];
        Rm[i][j] = (i == j ? 1.0 : 0.0) + 2.0 * w * ux[i][j] + 2.0 * uu;
      }
  }
  double dP_dw[3];  // 2 (u x d)
  dP_dw[0] = 2.0 * (u[1] * d[2] - u[2] * d[1]);
  dP_dw[1] = 2.0 * (u[2] * d[0] - u[0] * d[2]);
  dP_dw[2] = 2.0 * (u[0] * d[1] - u[1] * d[0]);
  double dP_du[3][3];  // -2w [d]x + 2 ((u.d) I + u d^T - 2 d u^T)
  {
    const double dx[3][3] = {{0, -d[2], d[1]}, {d[2], 0, -d[0]}, {-d[1], d[0], 0}};
    const double ud = u[0] * d[0] + u[1] * d[1] + u[2] * d[2];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) dP_du[i][j] = -2.0 * w * dx[i][j] + 2.0 * ((i == j ? ud : 0.0) + u[i] * d[j] - 2.0 * d[i] * u[j]);
  }
  // delta = m * conj(b) * a  (conj(conj(a) b) = conj(b) a): linear in a and in b
  const double cb[4] = {-B.q[0], -B.q[1], -B.q[2], B.q[3]};
  double M[4];
  QuatMul(c.q, cb, M);
  // d delta / d a = Lmat(M); rows x,y,z
  const double LM[3][4] = {{M[3], -M[2], M[1], M[0]}, {M[2], M[3], -M[0], M[1]}, {-M[1], M[0], M[3], M[2]}};
  // d delta / d b = Lmat(m) * Rmat(a) * diag(-1,-1,-1,1)
  const double* m = c.q;
  const double Lm[4][4] = {{m[3], -m[2], m[1], m[0]}, {m[2], m[3], -m[0], m[1]}, {-m[1], m[0], m[3], m[2]}, {-m[0], -m[1], -m[2], m[3]}};
  const double* a = A.q;
  const double Ra[4][4] = {{a[3], a[2], -a[1], a[0]}, {-a[2], a[3], a[0], a[1]}, {a[1], -a[0], a[3], a[2]}, {-a[0], -a[1], -a[2], a[3]}};
  double dD_db[3][4];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += Lm[i][k] * Ra[k][j];
      dD_db[i][j] = s * (j < 3 ? -1.0 : 1.0);
    }
  // ambient e-Jacobians: E_pa (6x3), E_qa (6x4), E_pb (6x3), E_qb (6x4)
  double Epa[6][3] = {{0}}, Eqa[6][4] = {{0}}, Epb[6][3] = {{0}}, Eqb[6][4] = {{0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { Epa[i][j] = -Rm[i][j]; Epb[i][j] = Rm[i][j]; }
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Eqa[i][j] = -dP_du[i][j];  // u = -a_v
    Eqa[i][3] = dP_dw[i];
    for (int j = 0; j < 4; j++) { Eqa[3 + i][j] = 2.0 * LM[i][j]; Eqb[3 + i][j] = 2.0 * dD_db[i][j]; }
  }
  // EigenQuaternionParameterization::ComputeJacobian (4x3), x = (x,y,z,w)
  auto plusJ = [](const double x[4], double J[4][3]) {
    J[0][0] = x[3];  J[0][1] = x[2];  J[0][2] = -x[1];
    J[1][0] = -x[2]; J[1][1] = x[3];  J[1][2] = x[0];
    J[2][0] = x[1];  J[2][1] = -x[0]; J[2][2] = x[3];
    J[3][0] = -x[0]; J[3][1] = -x[1]; J[3][2] = -x[2];
  };
  double Pa[4][3], Pb[4][3];
  plusJ(A.q, Pa);
  plusJ(B.q, Pb);
  double Ea[6][6], Eb[6][6];
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 3; j++) { Ea[i][j] = Epa[i][j]; Eb[i][j] = Epb[i][j]; }
    for (int j = 0; j < 3; j++) {
      double sa = 0, sb = 0;
      for (int k = 0; k < 4; k++) { sa += Eqa[i][k] * Pa[k][j]; sb += Eqb[i][k] * Pb[k][j]; }
      Ea[i][3 + j] = sa; Eb[i][3 + j] = sb;
    }
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double sa = 0, sb = 0;
      for (int k = 0; k < 6; k++) { sa += L[i * 6 + k] * Ea[k][j]; sb += L[i * 6 + k] * Eb[k][j]; }
      Ja[i * 6 + j] = sa; Jb[i * 6 + j] = sb;
    }
}

// Normal-equation assembly: H = sum J^T J (robustified), g = sum J^T r, cost = sum 0.5 rho(|r|^2).
// fixed_node's columns are dropped (SetParameterBlockConstant on the first node, ceresoptimizer.cpp:37-39).
// Output format: Hdiag [n_nodes][36], Hoff [n_con][36] (block (begin,end) of constraint c, row-major, = Ja^T Jb), g [n_nodes][6].
inline double PGAssemble(const std::vector<PGNode>& nodes, const std::vector<PGConstraint>& cons, const PGParams& par, int fixed_node,
                         double* Hdiag, double* Hoff, double* g, double* residuals_out /*6 per constraint, corrected*/) {
  const size_t N = nodes.size();
  for (size_t i = 0; i < N * 36; i++) Hdiag[i] = 0;
  for (size_t i = 0; i < N * 6; i++) g[i] = 0;
  double cost = 0;
  // odometry first, then loop constraints (AddConstraintType order, :34-35)
  for (int pass = 0; pass < 2; pass++)
    for (size_t ci = 0; ci < cons.size(); ci++) {
      const PGConstraint& c = cons[ci];
      if ((c.type == 1) != (pass == 1)) continue;
      double L[36], r[6], Ja[36], Jb[36];
      PGSqrtInformation(c, par, L);
      PGEvaluate(nodes[c.id_begin], nodes[c.id_end], c, L, r, Ja, Jb);
      double sq = 0;
      for (int i = 0; i < 6; i++) sq += r[i] * r[i];
      if (c.type == 1) {  // CauchyLoss(0.1) + corrector
        const double a = par.loop_cauchy, b = a * a, cc = 1.0 / b;
        const double sum = 1.0 + sq * cc, inv = 1.0 / sum;
        const double rho0 = b * std::log(sum), rho1 = std::max(std::numeric_limits<double>::min(), inv);
        cost += 0.5 * rho0;
        const double s1 = std::sqrt(rho1);  // rho'' < 0 -> plain sqrt(rho') scaling
        for (int i = 0; i < 36; i++) { Ja[i] *= s1; Jb[i] *= s1; }
        for (int i = 0; i < 6; i++) r[i] *= s1;
      } else {
        cost += 0.5 * sq;
      }
      if (residuals_out) for (int i = 0; i < 6; i++) residuals_out[ci * 6 + i] = r[i];
      const bool fa = (c.id_begin == fixed_node), fb = (c.id_end == fixed_node);
      double* Ha = Hdiag + (size_t)c.id_begin * 36;
      double* Hb = Hdiag + (size_t)c.id_end * 36;
      double* Ho = Hoff + ci * 36;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
          double aa = 0, bb = 0, ab = 0;
          for (int k = 0; k < 6; k++) { aa += Ja[k * 6 + i] * Ja[k * 6 + j]; bb += Jb[k * 6 + i] * Jb[k * 6 + j]; ab += Ja[k * 6 + i] * Jb[k * 6 + j]; }
          if (!fa) Ha[i * 6 + j] += aa;
          if (!fb) Hb[i * 6 + j] += bb;
          Ho[i * 6 + j] = (fa || fb) ? 0.0 : ab;
        }
      for (int i = 0; i < 6; i++) {
        double ga = 0, gb = 0;
        for (int k = 0; k < 6; k++) { ga += Ja[k * 6 + i] * r[k]; gb += Jb[k * 6 + i] * r[k]; }
        if (!fa) g[(size_t)c.id_begin * 6 + i] += ga;
        if (!fb) g[(size_t)c.id_end * 6 + i] += gb;
      }
    }
  return cost;
}

}  // namespace tbv_oracle

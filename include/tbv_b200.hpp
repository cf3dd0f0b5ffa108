// tbv_b200.hpp — C++14 host mirror of the reference's class surface for the hot path, over the C-ABI of tbv_b200.h.
//
// The reference calls its hot path through C++ classes.  This header gives the same names, argument meaning and error
// behaviour on top of libtbv_b200.so, with std containers where the reference uses PCL / Eigen types (neither exists in this image):
//   StructuredKStrongest   cfear_radarodometry/include/cfear_radarodometry/radar_filters.h:86-110
//   AzimuthCACFAR          cfear_radarodometry/include/cfear_radarodometry/cfar.h:28-42
//   MapPointNormal         cfear_radarodometry/include/cfear_radarodometry/pointnormal.h:108-200
//   n_scan_normal_reg      cfear_radarodometry/include/cfear_radarodometry/n_scan_normal.h:33-75 (Registration, registration.h:76)
//   OdometryKeyframeFuser  cfear_radarodometry/include/cfear_radarodometry/odometrykeyframefuser.h:90-213 (batched over sequences)
// A maintainer of the reference replaces the bodies of those classes with these calls; tests/cpp/test_host_mirror.cpp
// exercises the header against the CPU oracle the way the reference's own tests would.  Header-only; link with -ltbv_b200.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tbv_b200.h"

namespace tbv_b200 {

struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };   // pcl::PointXYZI
typedef std::vector<PointXYZI> PointCloud;                        // pcl::PointCloud<pcl::PointXYZI>
struct Pose2 { double x = 0, y = 0, yaw = 0; };                   // an Eigen::Affine3d of the planar problem, as Affine3dToVectorXYeZ reads it
typedef std::array<double, 36> Matrix6d;                          // row-major

enum costmetric { P2P = TBV_P2P, P2L = TBV_P2L, P2D = TBV_P2D };                                                  // registration.h:55
enum losstype { None = TBV_LOSS_NONE, Huber = TBV_LOSS_HUBER, Cauchy = TBV_LOSS_CAUCHY, SoftLOne = TBV_LOSS_SOFTLONE,
                Combined = TBV_LOSS_COMBINED, Tukey = TBV_LOSS_TUKEY };                                            // registration.h:60
enum weightoption { Uniform = TBV_W_UNIFORM, Sim_N = TBV_W_SIM_N, Sim_direciton = TBV_W_SIM_DIRECTION, Sim_scale = TBV_W_SIM_SCALE,
                    Combined_weights = TBV_W_COMBINED };                                                           // registration.h:50

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) { if (rc != TBV_OK) throw Error(rc, tbv_last_error()); }

class Context {   // one GPU context = one stream; distinct contexts may be used from distinct threads (odometry thread + loop thread)
 public:
  explicit Context(int device = 0) : h_(tbv_create(device)) { if (!h_) throw Error(TBV_ERR_NO_GPU, tbv_last_error()); }
  ~Context() { tbv_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  tbv_ctx* get() const { return h_; }
 private:
  tbv_ctx* h_;
};

// radarDriver::Process's k-strongest branch (radar_driver.cpp:57-61): the constructor filters, getPeaksFilteredPointCloud appends the
// requested cloud.  image: CV_8UC1, rows = azimuths, `step` bytes between rows (cv::Mat::step).
class StructuredKStrongest {
 public:
  StructuredKStrongest(Context& ctx, const uint8_t* image, int rows, int cols, size_t step, int z_min, int k_strongest, double min_distance,
                       double range_res) {
    const int cap = rows * k_strongest;
    for (auto* c : {&f_, &p_}) { c->x.resize(cap); c->y.resize(cap); c->i.resize(cap); }
    tbv_points f{cap, &f_.n, nullptr, nullptr, f_.i.data(), f_.x.data(), f_.y.data()};
    tbv_points p{cap, &p_.n, nullptr, nullptr, p_.i.data(), p_.x.data(), p_.y.data()};
    tbv_filter_params par{(float)z_min, k_strongest, (float)min_distance, (float)range_res};
    check(tbv_filter_kstrongest(ctx.get(), image, rows, cols, step, 1, &par, &f, &p));
  }
  void getPeaksFilteredPointCloud(PointCloud& output_pointcloud, bool peaks = false) const {
    const Soa& c = peaks ? p_ : f_;
    for (int k = 0; k < c.n; k++) { PointXYZI q; q.x = c.x[k]; q.y = c.y[k]; q.z = 0; q.intensity = c.i[k]; output_pointcloud.push_back(q); }
  }
 private:
  struct Soa { std::vector<float> x, y; std::vector<uint8_t> i; int n = 0; };
  Soa f_, p_;
};

class AzimuthCACFAR {
 public:
  AzimuthCACFAR(Context& ctx, int window_size = 40, double false_alarm_rate = 0.01, int nb_guard_cells = 5, double range_resolution = 0.0438,
                double static_threshold = 60.0, double min_distance = 2.5, double max_distance = 200.0)
      : ctx_(ctx), par_{window_size, false_alarm_rate, nb_guard_cells, range_resolution, static_threshold, min_distance, max_distance} {}
  void getFilteredPointCloud(const uint8_t* image, int rows, int cols, size_t step, PointCloud& output_pointcloud) const {
    const int cap = rows * cols;
    std::vector<float> x(cap), y(cap);
    std::vector<uint8_t> I(cap);
    int n = 0;
    tbv_points o{cap, &n, nullptr, nullptr, I.data(), x.data(), y.data()};
    check(tbv_filter_cacfar(ctx_.get(), image, rows, cols, step, 1, &par_, &o));
    for (int k = 0; k < n; k++) { PointXYZI q; q.x = x[k]; q.y = y[k]; q.intensity = I[k]; output_pointcloud.push_back(q); }
  }
 private:
  Context& ctx_;
  tbv_cfar_params par_;
};

// MapPointNormal(cloud, radius, origin, weight_intensity, raw) (pointnormal.h:118): the oriented surface points of a filtered cloud.
class MapPointNormal {
 public:
  static double& downsample_factor() { static double v = 1.0; return v; }   // pointnormal.cpp:5: a static of the reference class as well
  MapPointNormal(Context& ctx, const PointCloud& cld, float radius, const std::array<double, 2>& origin = {0.0, 0.0}, bool weight_intensity = false) {
    if (cld.empty()) return;         // the reference exits on an empty cloud (pointnormal.cpp:72-75); here: no cells
    std::vector<float> x(cld.size()), y(cld.size()), I(cld.size());
    for (size_t k = 0; k < cld.size(); k++) { x[k] = cld[k].x; y[k] = cld[k].y; I[k] = cld[k].intensity; }
    cells.resize(cld.size());
    int n = 0;
    check(tbv_build_cells(ctx.get(), x.data(), y.data(), I.data(), (int)cld.size(), radius, downsample_factor(), weight_intensity ? 1 : 0, origin.data(),
                          cells.data(), (int)cells.size(), &n, nullptr));
    cells.resize(n);
  }
  size_t GetSize() const { return cells.size(); }
  const tbv_cell& GetCell(size_t i) const { return cells[i]; }
  std::vector<tbv_cell> cells;
};
typedef std::shared_ptr<MapPointNormal> MapNormalPtr;

// n_scan_normal_reg (n_scan_normal.h:33-75): scans.back() is the moving scan in its own frame, the others are fixed; Tsrc in/out.
class n_scan_normal_reg {
 public:
  explicit n_scan_normal_reg(Context& ctx, costmetric cost = P2L, losstype loss = Huber, double loss_limit = 0.1, weightoption opt = Uniform)
      : ctx_(ctx), par_{cost, loss, opt, loss_limit, 1.0, 0.0, 8, 20} {}
  void SetParameters(unsigned max_itr_association, unsigned max_itr_solver) { par_.max_itr_association = (int)max_itr_association; par_.max_itr_solver = (int)max_itr_solver; }
  void SetD2dPar(double cov_scale, double regularization) { par_.cov_scale = cov_scale; par_.regularization = regularization; }
  bool Register(std::vector<MapNormalPtr>& scans, std::vector<Pose2>& Tsrc, std::vector<Matrix6d>* reg_cov = nullptr) {
    std::vector<const tbv_cell*> ptr; std::vector<int> n; std::vector<double> T;
    pack(scans, Tsrc, ptr, n, T);
    check(tbv_register(ctx_.get(), (int)scans.size(), ptr.data(), n.data(), T.data(), &par_, &summary_));
    for (size_t k = 0; k < Tsrc.size(); k++) Tsrc[k] = Pose2{T[3 * k], T[3 * k + 1], T[3 * k + 2]};
    itr_ = (size_t)summary_.itrs;
    score_ = summary_.score;
    if (reg_cov) {   // n_scan_normal.cpp:171-175: the same fixed covariance for every scan
      Matrix6d c; c.fill(0.0); c[0] = 0.1 * 0.1; c[7] = 0.1 * 0.1; c[35] = 0.01 * 0.01;
      reg_cov->assign(scans.size(), c);
    }
    return summary_.success != 0;
  }
  bool GetCost(std::vector<MapNormalPtr>& scans, std::vector<Pose2>& Tsrc, double& score, std::vector<double>& residuals) {
    std::vector<const tbv_cell*> ptr; std::vector<int> n; std::vector<double> T;
    pack(scans, Tsrc, ptr, n, T);
    const int cap = 2 * n.back() * (int)(scans.size() - 1) + 8;
    residuals.assign(cap, 0.0);
    double sc = 0, cost = 0; int nres = 0;
    check(tbv_get_cost(ctx_.get(), (int)scans.size(), ptr.data(), n.data(), T.data(), &par_, (int)itr_, &sc, &cost, &nres, residuals.data(), cap));
    residuals.resize(nres > 0 ? nres : 0);
    if (nres <= 1) return false;    // "too few residuals" (n_scan_normal.cpp:203-206)
    score = cost;                   // problem_->Evaluate's total cost (:208)
    score_ = sc;                    // score_ = score / max(residuals, 1) (:209)
    return true;
  }
  double getScore() const { return score_; }
  bool GetCovarianceScaler(double& cov_scale) const {   // n_scan_normal.cpp:433-439: one free 3-parameter block
    if (summary_.num_residuals - 3 == 0) return false;
    cov_scale = summary_.final_cost / (summary_.num_residuals - 3);
    return true;
  }
  tbv_reg_summary summary_{};
  size_t itr_ = 0;
 private:
  static void pack(const std::vector<MapNormalPtr>& scans, const std::vector<Pose2>& Tsrc, std::vector<const tbv_cell*>& ptr, std::vector<int>& n,
                   std::vector<double>& T) {
    if (scans.size() != Tsrc.size() || scans.size() < 2) throw Error(TBV_ERR_INVALID, "Register: need >= 2 scans and as many poses");
    for (size_t k = 0; k < scans.size(); k++) {
      ptr.push_back(scans[k]->cells.data()); n.push_back((int)scans[k]->cells.size());
      T.push_back(Tsrc[k].x); T.push_back(Tsrc[k].y); T.push_back(Tsrc[k].yaw);
    }
  }
  Context& ctx_;
  tbv_reg_params par_;
  double score_ = 0;
};

// radarDriver::Process + OdometryKeyframeFuser::processFrame (odometrykeyframefuser.cpp:143-259) for n_seq independent sequences in
// lock-step: one call = one frame of every sequence (scans back to back in one host buffer); all state stays on the device.
class OdometryKeyframeFuser {
 public:
  OdometryKeyframeFuser(Context& ctx, int n_seq, int n_az, int n_range, const tbv_odom_params& par)
      : n_seq_(n_seq), h_(tbv_odom_create(ctx.get(), n_seq, n_az, n_range, &par)) { if (!h_) throw Error(TBV_ERR_INVALID, tbv_last_error()); }
  ~OdometryKeyframeFuser() { tbv_odom_destroy(h_); }
  OdometryKeyframeFuser(const OdometryKeyframeFuser&) = delete;
  OdometryKeyframeFuser& operator=(const OdometryKeyframeFuser&) = delete;
  std::vector<tbv_odom_out> pointcloudCallback(const uint8_t* polar_scans) {
    std::vector<tbv_odom_out> out(n_seq_);
    check(tbv_odom_step(h_, polar_scans, out.data()));
    return out;
  }
 private:
  int n_seq_;
  tbv_odom* h_;
};

// RSCManager (place_recognition_radar/include/place_recognition_radar/RadarScancontext.h:31-131): the keyframe database (descriptors,
// ring keys, odometry poses) lives on the host, as in the reference; every computation runs on the GPU — the descriptor + keys of the
// keyframe and its four lateral augmentations in one launch, the ring-key search with the odometry likelihood, and all descriptor
// distances of the <= 50 (query, candidate) pairs in one launch.  The candidate bookkeeping is the reference's (RadarScancontext.cpp:
// 286-345: running sort by distance, keep N_CANDIDATES).
struct candidate {        // RadarScancontext.h: struct candidate
  double min_dist = 0, min_dist_sc = 0, min_dist_odom = 0;
  float yaw_diff_rad = 0;
  int nn_idx = -1, argmin_shift = 0;
  int aug_idx = 0;        // which query: 0 = the keyframe itself, 1..4 = the lateral offsets of Taug
  double aug_xy[2] = {0, 0};
};
inline tbv_sc_params default_sc_params() {   // TBV-8 offline settings (tbv_slam/src/tbv_slam_offline.cpp:81-101)
  return tbv_sc_params{40, 120, 80.0, 0.1, 10, 3, 0.05, 1, 1, 0.0, 0, 1000.0, 10.0};
}
class RSCManager {
 public:
  explicit RSCManager(Context& ctx, const tbv_sc_params& par = default_sc_params()) : ctx_(ctx), par_(par) {}
  void makeAndSaveScancontextAndKeysRadarCloud(const PointCloud& cloud, const Pose2& Todom) {
    static const double aug[5][2] = {{0, 0}, {0, -2}, {0, 2}, {0, -4}, {0, 4}};   // RadarScancontext.cpp:162-166 (+ the cloud itself)
    const int nq = par_.augment_sc ? 5 : 1, R = par_.num_ring, S = par_.num_sector;
    std::vector<float> x(cloud.size()), y(cloud.size()), I(cloud.size());
    for (size_t k = 0; k < cloud.size(); k++) { x[k] = cloud[k].x; y[k] = cloud[k].y; I[k] = cloud[k].intensity; }
    q_desc_.assign((size_t)nq * R * S, 0.0); q_keys_.assign((size_t)nq * R, 0.f); q_off_.assign(&aug[0][0], &aug[0][0] + 2 * nq);
    check(tbv_sc_make(ctx_.get(), x.data(), y.data(), I.data(), (int)cloud.size(), &par_, nq, q_off_.data(), q_desc_.data(), q_keys_.data(), nullptr));
    polarcontexts_.insert(polarcontexts_.end(), q_desc_.begin(), q_desc_.begin() + (size_t)R * S);
    ringkeys_.insert(ringkeys_.end(), q_keys_.begin(), q_keys_.begin() + R);
    odom_.push_back(Todom.x); odom_.push_back(Todom.y); odom_.push_back(Todom.yaw);
  }
  std::vector<candidate> detectLoopClosureID() {
    std::vector<candidate> similar;
    const int R = par_.num_ring, S = par_.num_sector, n_db = (int)(ringkeys_.size() / R), nq = (int)(q_keys_.size() / R), want = par_.num_candidates_from_tree;
    if (n_db == 0) return similar;
    std::vector<int> cur(nq, n_db - 1), ci((size_t)nq * want), ne(nq);
    std::vector<double> cs((size_t)nq * want);
    check(tbv_sc_search(ctx_.get(), ringkeys_.data(), odom_.data(), n_db, nq, q_keys_.data(), cur.data(), &par_, ci.data(), cs.data(), ne.data()));
    if (n_db < ne[0] + 1) return similar;
    std::vector<int> pq, pc, uniq;        // (query, candidate) pairs; the candidates' descriptors are gathered once
    std::vector<double> psim;
    for (int q = 0; q < nq; q++)
      for (int t = 0; t < want; t++)
        if (ci[(size_t)q * want + t] >= 0) { pq.push_back(q); pc.push_back(ci[(size_t)q * want + t]); psim.push_back(cs[(size_t)q * want + t]); }
    if (pq.empty()) return similar;
    uniq = pc;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<double> cdesc((size_t)uniq.size() * R * S);
    for (size_t u = 0; u < uniq.size(); u++) std::copy(polarcontexts_.begin() + (size_t)uniq[u] * R * S, polarcontexts_.begin() + (size_t)(uniq[u] + 1) * R * S, cdesc.begin() + u * R * S);
    std::vector<int> cpos(pc.size()), shift(pc.size());
    for (size_t k = 0; k < pc.size(); k++) cpos[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), pc[k]) - uniq.begin());
    std::vector<double> dist(pc.size());
    check(tbv_sc_distance_batch(ctx_.get(), q_desc_.data(), nq, cdesc.data(), (int)uniq.size(), (int)pc.size(), pq.data(), cpos.data(), &par_, dist.data(), shift.data()));
    const double unit = 360.0 / (double)S;
    for (size_t k = 0; k < pc.size(); k++) {
      candidate c;
      c.min_dist_sc = dist[k];
      c.min_dist_odom = par_.odometry_coupled_closure ? psim[k] : 0.0;
      c.min_dist = par_.odometry_coupled_closure ? dist[k] + c.min_dist_odom : dist[k];
      const float deg = (float)(shift[k] * unit);
      c.yaw_diff_rad = (float)((double)deg * M_PI / 180.0);
      c.nn_idx = pc[k]; c.argmin_shift = shift[k]; c.aug_idx = pq[k];
      c.aug_xy[0] = q_off_[2 * pq[k]]; c.aug_xy[1] = q_off_[2 * pq[k] + 1];
      similar.push_back(c);
      std::stable_sort(similar.begin(), similar.end(), [](const candidate& a, const candidate& b) { return a.min_dist < b.min_dist; });
      if ((int)similar.size() > par_.n_candidates) similar.pop_back();
    }
    return similar;
  }
  size_t size() const { return odom_.size() / 3; }
 private:
  Context& ctx_;
  tbv_sc_params par_;
  std::vector<double> polarcontexts_, odom_, q_desc_, q_off_;
  std::vector<float> ringkeys_, q_keys_;
};

// CorAlRadarQuality / CFEARQuality for a batch of candidate pairs (coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:99-229, 330-352):
// clouds / cell sets are given once and indexed by the pairs; pair p compares src at T_src[p] (* T_offset[p]) with ref at T_ref[p].
inline std::vector<tbv_coral_result> CorAlRadarQuality(Context& ctx, const std::vector<PointCloud>& clouds, const std::vector<int>& src, const std::vector<int>& ref,
                                                       const std::vector<Pose2>& T_src, const std::vector<Pose2>& T_ref, double radius = 1.0,
                                                       bool weight_res_intensity = false) {
  std::vector<std::vector<float>> x(clouds.size()), y(clouds.size()), I(clouds.size());
  std::vector<const float*> px, py, pi;
  std::vector<int> n;
  for (size_t c = 0; c < clouds.size(); c++) {
    for (const PointXYZI& q : clouds[c]) { x[c].push_back(q.x); y[c].push_back(q.y); I[c].push_back(q.intensity); }
    px.push_back(x[c].data()); py.push_back(y[c].data()); pi.push_back(I[c].data()); n.push_back((int)clouds[c].size());
  }
  std::vector<double> ts, tr;
  for (size_t k = 0; k < src.size(); k++) { ts.insert(ts.end(), {T_src[k].x, T_src[k].y, T_src[k].yaw}); tr.insert(tr.end(), {T_ref[k].x, T_ref[k].y, T_ref[k].yaw}); }
  tbv_coral_params par{radius, weight_res_intensity ? 1 : 0, 1};
  std::vector<tbv_coral_result> out(src.size());
  check(tbv_coral_quality_batch(ctx.get(), (int)clouds.size(), px.data(), py.data(), pi.data(), n.data(), (int)src.size(), src.data(), ref.data(), ts.data(), nullptr, tr.data(),
                                &par, out.data(), nullptr));
  return out;
}

// ---- CeresLeastSquares (tbv_slam/include/tbv_slam/ceresoptimizer.h:25-49, tbv_slam/src/tbv_slam/ceresoptimizer.cpp:13-62) ----------------------------
// Solve() = BuildOptimizationProblem + ceres::Solve(LEVENBERG_MARQUARDT, SPARSE_NORMAL_CHOLESKY, max_num_iterations 200), in place on the node
// poses.  Evaluation (tbv_pgo_assemble) and the damped linear solve (tbv_pgo_solve_step) run on the device; this class is the trust-region
// bookkeeping of Ceres 2.1.0's TrustRegionMinimizer with its default options (initial radius 1e4, accept rho > 1e-3, radius update by
// max(1/3, 1 - (2 rho - 1)^3), halving / quartering / ... on consecutive rejections, function / gradient / parameter tolerances 1e-6 / 1e-10 / 1e-8).
// Ceres' Jacobi column scaling is not applied: same fixed point, different iterates.  The backend is a template parameter so that the loop can
// be exercised without a GPU (tests/cpp/test_pgo_host.cpp plugs the oracle in); the product type is CeresLeastSquares = ...<DevicePoseGraph>.
struct Pose3d {                                   // types.h:46-81, q stored (x, y, z, w) like Eigen::Quaterniond::coeffs()
  double p[3] = {0, 0, 0};
  double q[4] = {0, 0, 0, 1};
};
struct Constraint3d {                             // types.h:155-190 (members the optimiser reads)
  unsigned long id_begin = 0, id_end = 0;         // ROW of the node in the vector handed to the optimiser
  Pose3d t_be;
  Matrix6d information{};                         // used when replace_cov_by_identity == 0
  int type = 0;                                   // 0 odometry, 1 loop_appearance; others are not optimised (ceresoptimizer.cpp:34-35)
};
inline tbv_pgo_params default_pgo_params() { return tbv_pgo_params{0.01, 0.01, 0.001, 500000.0, 1, 0.1}; }   // ceresoptimizer.cpp:18-27

struct DevicePoseGraph {                          // the two device calls
  Context* ctx;
  explicit DevicePoseGraph(Context& c) : ctx(&c) {}
  void assemble(int n, const double* nodes, int m, const int* ids, const double* meas, const double* info, const tbv_pgo_params& par, int fixed,
                double* cost, double* Hd, double* Ho, double* g) const {
    check(tbv_pgo_assemble(ctx->get(), n, nodes, m, ids, meas, info, &par, fixed, cost, Hd, Ho, g, nullptr));
  }
  void solve(int n, int m, const int* ids, const double* Hd, const double* Ho, const double* g, int fixed, double radius, int max_iters, double rel_tol,
             double* delta, int* iters) const {
    check(tbv_pgo_solve_step(ctx->get(), n, m, ids, Hd, Ho, g, fixed, radius, max_iters, rel_tol, delta, iters, nullptr));
  }
};

template <class Backend>
class CeresLeastSquaresT {
 public:
  struct Options {                                // ceres::Solver::Options defaults except max_num_iterations (ceresoptimizer.cpp:53)
    int max_num_iterations = 200;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32, min_relative_decrease = 1e-3;
    int cg_max_iterations = 20000;
    double cg_relative_tolerance = 1e-10;
  };
  struct Summary {
    double initial_cost = 0, final_cost = 0;
    int iterations = 0, num_successful_steps = 0, cg_iterations = 0;
    std::string termination = "max_num_iterations";   // NO_CONVERGENCE in Ceres' words; the others are CONVERGENCE
    bool IsSolutionUsable() const { return termination != "min_trust_region_radius"; }
  };

  CeresLeastSquaresT(Backend backend, std::vector<Pose3d>& nodes, const std::vector<Constraint3d>& constraints,
                     const tbv_pgo_params& par = default_pgo_params(), int fixed_node = 0)
      : be_(backend), nodes_(nodes), par_(par), fixed_(fixed_node) {
    for (const Constraint3d& c : constraints) {
      if (c.type != 0 && c.type != 1) continue;
      if (c.id_begin >= nodes.size() || c.id_end >= nodes.size()) throw Error(TBV_ERR_INVALID, "constraint references a missing node");
      ids_.insert(ids_.end(), {(int)c.id_begin, (int)c.id_end, c.type});
      meas_.insert(meas_.end(), c.t_be.p, c.t_be.p + 3);
      meas_.insert(meas_.end(), c.t_be.q, c.t_be.q + 4);
      info_.insert(info_.end(), c.information.begin(), c.information.end());
    }
  }
  Options options;
  Summary summary_;

  // x (+) delta: p += dp, q = exp(dr) * q  (ceres::EigenQuaternionParameterization::Plus)
  static void Plus(const double* x, const double* d, double* out) {
    for (int k = 0; k < 3; k++) out[k] = x[k] + d[k];
    const double n = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    const double s = n > 0.0 ? std::sin(n) / n : 1.0, dw = std::cos(n);
    const double dx = s * d[3], dy = s * d[4], dz = s * d[5];
    const double qx = x[3], qy = x[4], qz = x[5], qw = x[6];
    out[3] = dw * qx + qw * dx + (dy * qz - dz * qy);
    out[4] = dw * qy + qw * dy + (dz * qx - dx * qz);
    out[5] = dw * qz + qw * dz + (dx * qy - dy * qx);
    out[6] = dw * qw - (dx * qx + dy * qy + dz * qz);
  }

  void Solve() {
    const int n = (int)nodes_.size(), m = (int)(ids_.size() / 3);
    Summary S;
    std::vector<double> x(7 * (size_t)n), xc(7 * (size_t)n), Hd(36 * (size_t)n), Ho(36 * (size_t)std::max(m, 1)), g(6 * (size_t)n), Hdc(Hd.size()),
        Hoc(Ho.size()), gc(g.size()), delta(6 * (size_t)n), Hx(6 * (size_t)n);
    for (int i = 0; i < n; i++) { std::copy(nodes_[i].p, nodes_[i].p + 3, &x[7 * i]); std::copy(nodes_[i].q, nodes_[i].q + 4, &x[7 * i + 3]); }
    const double* info = par_.replace_cov_by_identity ? nullptr : info_.data();
    double cost = 0;
    be_.assemble(n, x.data(), m, ids_.data(), meas_.data(), info, par_, fixed_, &cost, Hd.data(), Ho.data(), g.data());
    S.initial_cost = S.final_cost = cost;
    auto max_abs = [](const std::vector<double>& v) { double a = 0; for (double e : v) a = std::max(a, std::fabs(e)); return a; };
    auto norm = [](const std::vector<double>& v) { double a = 0; for (double e : v) a += e * e; return std::sqrt(a); };
    double radius = options.initial_trust_region_radius, decrease = 2.0;
    if (max_abs(g) <= options.gradient_tolerance) S.termination = "gradient_tolerance";
    else
      for (int it = 0; it < options.max_num_iterations; it++) {
        S.iterations++;
        int cg = 0;
        be_.solve(n, m, ids_.data(), Hd.data(), Ho.data(), g.data(), fixed_, radius, options.cg_max_iterations, options.cg_relative_tolerance, delta.data(), &cg);
        S.cg_iterations += cg;
        // model_cost_change = -delta^T (g + H delta / 2), H in the block layout of tbv_pgo_assemble
        for (int i = 0; i < n; i++)
          for (int a = 0; a < 6; a++) {
            double acc = 0;
            for (int b = 0; b < 6; b++) acc += Hd[36 * (size_t)i + 6 * a + b] * delta[6 * (size_t)i + b];
            Hx[6 * (size_t)i + a] = acc;
          }
        for (int c = 0; c < m; c++) {
          const int ia = ids_[3 * c], ib = ids_[3 * c + 1];
          const double* B = &Ho[36 * (size_t)c];
          for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
              Hx[6 * (size_t)ia + a] += B[6 * a + b] * delta[6 * (size_t)ib + b];
              Hx[6 * (size_t)ib + a] += B[6 * b + a] * delta[6 * (size_t)ia + b];
            }
        }
        double model_change = 0;
        bool finite = true;
        for (size_t k = 0; k < delta.size(); k++) { model_change -= delta[k] * (g[k] + 0.5 * Hx[k]); finite = finite && std::isfinite(delta[k]); }
        auto reject = [&]() { radius /= decrease; decrease *= 2.0; return radius < options.min_trust_region_radius; };
        if (!finite || !(model_change > 0.0)) {
          if (reject()) { S.termination = "min_trust_region_radius"; break; }
          continue;
        }
        if (norm(delta) <= options.parameter_tolerance * (norm(x) + options.parameter_tolerance)) { S.termination = "parameter_tolerance"; break; }
        for (int i = 0; i < n; i++) Plus(&x[7 * (size_t)i], &delta[6 * (size_t)i], &xc[7 * (size_t)i]);
        double cost_c = 0;
        be_.assemble(n, xc.data(), m, ids_.data(), meas_.data(), info, par_, fixed_, &cost_c, Hdc.data(), Hoc.data(), gc.data());
        const double rho = (cost - cost_c) / model_change;
        if (rho > options.min_relative_decrease) {
          const double dcost = cost - cost_c;
          S.num_successful_steps++;
          x.swap(xc); Hd.swap(Hdc); Ho.swap(Hoc); g.swap(gc);
          cost = S.final_cost = cost_c;
          const double t = 2.0 * rho - 1.0;
          radius = std::min(radius / std::max(1.0 / 3.0, 1.0 - t * t * t), options.max_trust_region_radius);
          decrease = 2.0;
          if (max_abs(g) <= options.gradient_tolerance) { S.termination = "gradient_tolerance"; break; }
          if (std::fabs(dcost) <= options.function_tolerance * cost) { S.termination = "function_tolerance"; break; }
        } else if (reject()) {
          S.termination = "min_trust_region_radius";
          break;
        }
      }
    for (int i = 0; i < n; i++) { std::copy(&x[7 * i], &x[7 * i] + 3, nodes_[i].p); std::copy(&x[7 * i + 3], &x[7 * i] + 7, nodes_[i].q); }
    summary_ = S;
  }

 private:
  Backend be_;
  std::vector<Pose3d>& nodes_;
  tbv_pgo_params par_;
  int fixed_;
  std::vector<int> ids_;
  std::vector<double> meas_, info_;
};
typedef CeresLeastSquaresT<DevicePoseGraph> CeresLeastSquares;

// ---- OdometryKeyframeFuser with the reference's OWN interface: point clouds in (odometrykeyframefuser.cpp:143-259, 395-411) -------------------------------
// The batched OdometryKeyframeFuser above fuses the k-strongest filter into the frame (tbv_odom_step) and is the fast path.  This one takes the CLOUD,
// whatever filter produced it (CA-CFAR: radar_driver.cpp:52-56, or a caller's own): processFrame as host bookkeeping — poses as planar affine
// matrices multiplied in Eigen's operation order, keyframe rule, sliding window — over three device calls per frame: tbv_compensate, tbv_build_cells,
// tbv_register.  The backend is a template parameter so the bookkeeping can be checked without a GPU (tests/cpp/test_points_fuser.cpp plugs the
// oracle in); the product type is PointCloudOdometryFuser = ...<DeviceOdometryPrimitives>.
struct Affine2d {                                  // planar Eigen::Affine3d: 2x2 linear part + translation
  double r00 = 1, r01 = 0, r10 = 0, r11 = 1, tx = 0, ty = 0;
  static Affine2d FromPose(const Pose2& p) {       // vectorToAffine3d (registration.cpp:129-135)
    const double c = std::cos(p.yaw), s = std::sin(p.yaw);
    Affine2d T; T.r00 = c; T.r01 = -s; T.r10 = s; T.r11 = c; T.tx = p.x; T.ty = p.y;
    return T;
  }
  Pose2 ToPose() const { return Pose2{tx, ty, std::atan2(r10, r11)}; }   // Affine3dToVectorXYeZ
  Affine2d operator*(const Affine2d& b) const {
    Affine2d c;
    c.r00 = r00 * b.r00 + r01 * b.r10; c.r01 = r00 * b.r01 + r01 * b.r11;
    c.r10 = r10 * b.r00 + r11 * b.r10; c.r11 = r10 * b.r01 + r11 * b.r11;
    c.tx = (r00 * b.tx + r01 * b.ty) + tx; c.ty = (r10 * b.tx + r11 * b.ty) + ty;
    return c;
  }
  Affine2d inverse() const {                       // Eigen's general affine inverse: cofactors / determinant, t' = -L^-1 t
    const double inv = 1.0 / (r00 * r11 - r01 * r10);
    Affine2d i;
    i.r00 = r11 * inv; i.r01 = -r01 * inv; i.r10 = -r10 * inv; i.r11 = r00 * inv;
    i.tx = -(i.r00 * tx + i.r01 * ty); i.ty = -(i.r10 * tx + i.r11 * ty);
    return i;
  }
};

struct DeviceOdometryPrimitives {
  Context* ctx;
  explicit DeviceOdometryPrimitives(Context& c) : ctx(&c) {}
  void compensate(std::vector<float>& x, std::vector<float>& y, const double mot[3], bool ccw) const {
    if (!x.empty()) check(tbv_compensate(ctx->get(), x.data(), y.data(), (int)x.size(), mot, ccw ? 1 : 0));
  }
  std::vector<tbv_cell> build_cells(const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& intensity, float radius,
                                    double downsample_factor, bool weight_intensity) const {
    std::vector<tbv_cell> cells(std::max<size_t>(x.size(), 1));
    const double origin[2] = {0.0, 0.0};
    int n = 0;
    if (!x.empty())
      check(tbv_build_cells(ctx->get(), x.data(), y.data(), intensity.data(), (int)x.size(), radius, downsample_factor, weight_intensity ? 1 : 0, origin,
                            cells.data(), (int)cells.size(), &n, nullptr));
    cells.resize(n);
    return cells;
  }
  void register_scans(const std::vector<const tbv_cell*>& scans, const std::vector<int>& n_cells, std::vector<double>& T, const tbv_reg_params& par,
                      tbv_reg_summary& summary) const {
    check(tbv_register(ctx->get(), (int)scans.size(), scans.data(), n_cells.data(), T.data(), &par, &summary));
  }
};

template <class Backend>
class PointCloudOdometryFuserT {
 public:
  PointCloudOdometryFuserT(Backend backend, const tbv_odom_params& par) : be_(backend), par_(par) {}   // par.filter is not used here
  struct Keyframe { Affine2d pose; std::shared_ptr<std::vector<tbv_cell>> cells; };

  static bool KeyFrameBasedFuse(const Affine2d& diff, bool use_keyframe, double min_keyframe_dist, double min_keyframe_rot_deg) {   // :62-73
    if (!use_keyframe) return true;
    const double yaw = std::atan2(diff.r10, diff.r11), tnorm = std::sqrt(diff.tx * diff.tx + diff.ty * diff.ty);
    return tnorm > min_keyframe_dist || std::fabs(yaw) > (min_keyframe_rot_deg * M_PI / 180.0);
  }
  static bool AccelerationVelocitySanityCheck(const Affine2d& prev, const Affine2d& cur) {                                             // :76-94
    const double dt = 0.25, vel_limit = 200, acc_limit = 200;
    const double vx = cur.tx / dt, vy = cur.ty / dt, ax = (cur.tx - prev.tx) / (dt * dt), ay = (cur.ty - prev.ty) / (dt * dt);
    return !(std::sqrt(ax * ax + ay * ay) > acc_limit || std::sqrt(vx * vx + vy * vy) > vel_limit);
  }

  // cloud (and the optional peaks cloud) are compensated in place, as the reference does; returns Tcurrent
  Pose2 pointcloudCallback(PointCloud& cloud, PointCloud* cloud_peaks = nullptr) {
    const Affine2d TprevMot = Tmot_;
    if (par_.compensate) {
      const Pose2 m = TprevMot.ToPose();
      const double mot[3] = {m.x, m.y, m.yaw};
      compensate_cloud(cloud, mot);
      if (cloud_peaks) compensate_cloud(*cloud_peaks, mot);
    }
    std::vector<float> x(cloud.size()), y(cloud.size()), I(cloud.size());
    for (size_t k = 0; k < cloud.size(); k++) { x[k] = cloud[k].x; y[k] = cloud[k].y; I[k] = cloud[k].intensity; }
    auto cells = std::make_shared<std::vector<tbv_cell>>(be_.build_cells(x, y, I, (float)par_.res, par_.downsample_factor, par_.weight_intensity != 0));
    last_cells_ = cells;
    const Affine2d Tguess = par_.use_guess ? T_prev_ * TprevMot : T_prev_;
    updated = false; last_reg_ok = true; last_itrs = 0;
    if (keyframes_.empty()) {
      keyframes_.push_back(Keyframe{Affine2d(), cells});
      updated = true;
      return Tcurrent_.ToPose();
    }
    std::vector<const tbv_cell*> scans; std::vector<int> n; std::vector<double> T;
    for (const Keyframe& k : keyframes_) {
      const Pose2 p = k.pose.ToPose();
      scans.push_back(k.cells->data()); n.push_back((int)k.cells->size()); T.insert(T.end(), {p.x, p.y, p.yaw});
    }
    const Pose2 g = Tguess.ToPose();
    scans.push_back(cells->data()); n.push_back((int)cells->size()); T.insert(T.end(), {g.x, g.y, g.yaw});
    tbv_reg_summary s{};
    be_.register_scans(scans, n, T, par_.reg, s);                     // the reference ignores the result (shadowed `success`, :184-193)
    last_reg_ok = s.success != 0; last_itrs = s.itrs;
    Tcurrent_ = Affine2d::FromPose(Pose2{T[T.size() - 3], T[T.size() - 2], T[T.size() - 1]});
    const Affine2d Tmot_current = T_prev_.inverse() * Tcurrent_;
    if (!AccelerationVelocitySanityCheck(Tmot_, Tmot_current)) Tcurrent_ = Tguess;
    Tmot_ = T_prev_.inverse() * Tcurrent_;
    const Affine2d Tkeydiff = keyframes_.back().pose.inverse() * Tcurrent_;
    if (KeyFrameBasedFuse(Tkeydiff, par_.use_keyframe != 0, par_.min_keyframe_dist, par_.min_keyframe_rot_deg)) {
      keyframes_.push_back(Keyframe{Tcurrent_, cells});
      if (keyframes_.size() > (size_t)par_.submap_scan_size) keyframes_.erase(keyframes_.begin());
      updated = true;
    }
    T_prev_ = Tcurrent_;
    return Tcurrent_.ToPose();
  }
  const std::vector<Keyframe>& keyframes() const { return keyframes_; }
  size_t last_n_cells() const { return last_cells_ ? last_cells_->size() : 0; }
  bool updated = false, last_reg_ok = true;
  int last_itrs = 0;

 private:
  void compensate_cloud(PointCloud& c, const double mot[3]) {
    std::vector<float> x(c.size()), y(c.size());
    for (size_t k = 0; k < c.size(); k++) { x[k] = c[k].x; y[k] = c[k].y; }
    be_.compensate(x, y, mot, par_.radar_ccw != 0);
    for (size_t k = 0; k < c.size(); k++) { c[k].x = x[k]; c[k].y = y[k]; }
  }
  Backend be_;
  tbv_odom_params par_;
  std::vector<Keyframe> keyframes_;
  std::shared_ptr<std::vector<tbv_cell>> last_cells_;
  Affine2d Tcurrent_, T_prev_, Tmot_;
};
typedef PointCloudOdometryFuserT<DeviceOdometryPrimitives> PointCloudOdometryFuser;

// ---- simple graph hand-off (cfear_radarodometry/include/cfear_radarodometry/types.h:93-192, types.cpp:103-130) in the Boost-free .tbvg layout ---------
// Same content as the reference's simple_graph (vector<pair<RadarScan, vector<Constraint3d>>>), member for member; the byte layout is specified in
// tbv_slam_public_b200/graph_io.py (little endian, no padding) and is what SaveSimpleGraph / LoadSimpleGraph below write and read.
struct GraphConstraint {                          // Constraint3d with every serialised member
  unsigned long id_begin = 0, id_end = 0;
  Pose3d t_be;
  Matrix6d information{};
  int type = 0;                                   // ConstraintType: 0 odometry, 1 loop_appearance, 2 mini_loop, 3 candidate
  std::vector<std::pair<std::string, double>> quality;   // std::map order (sorted by key) on disk
  std::string info;
};
struct RadarScan {                                // serialised members of RadarScan + MapPointNormal
  Pose3d T, Tgt;
  bool has_Tgt_ = false;
  unsigned int idx_ = 0;
  unsigned long stamp_ = 0;
  std::array<double, 16> motion_{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};   // 4x4 row-major
  PointCloud cloud_peaks_, cloud_nopeaks_;
  std::vector<tbv_cell> cloud_normal_;
  std::vector<std::array<float, 2>> downsampled_;
  float radius_ = 3.0f;
  bool weight_intensity_ = true;
};
typedef std::vector<std::pair<RadarScan, std::vector<GraphConstraint>>> simple_graph;

namespace detail {
struct Writer {
  std::string out;
  void raw(const void* p, size_t n) { out.append(static_cast<const char*>(p), n); }
  template <class T> void put(T v) { raw(&v, sizeof(T)); }          // host is little endian (x86-64 / aarch64), as the layout requires
};
struct Reader {
  const std::string& in;
  size_t o = 0;
  explicit Reader(const std::string& s) : in(s) {}
  void raw(void* p, size_t n) {
    if (o + n > in.size()) throw Error(TBV_ERR_INVALID, "truncated .tbvg file");
    std::copy(in.data() + o, in.data() + o + n, static_cast<char*>(p));
    o += n;
  }
  template <class T> T get() { T v; raw(&v, sizeof(T)); return v; }
  template <class N> size_t count(size_t element_bytes) {           // an element count that the rest of the file can actually hold
    const size_t n = (size_t)get<N>();
    if (element_bytes && n > (in.size() - o) / element_bytes) throw Error(TBV_ERR_INVALID, "truncated .tbvg file");
    return n;
  }
};
inline void put_pose(Writer& w, const Pose3d& P) { w.raw(P.p, 24); w.raw(P.q, 32); }
inline void get_pose(Reader& r, Pose3d& P) { r.raw(P.p, 24); r.raw(P.q, 32); }
inline void put_cloud(Writer& w, const PointCloud& c) {
  w.put<uint32_t>((uint32_t)c.size());
  for (const PointXYZI& p : c) { w.put(p.x); w.put(p.y); w.put(p.z); w.put(p.intensity); }
}
inline void get_cloud(Reader& r, PointCloud& c) {
  c.resize(r.count<uint32_t>(16));
  for (PointXYZI& p : c) { p.x = r.get<float>(); p.y = r.get<float>(); p.z = r.get<float>(); p.intensity = r.get<float>(); }
}
}  // namespace detail

inline std::string SerializeSimpleGraph(const simple_graph& graph) {
  static_assert(sizeof(tbv_cell) == 128, "tbv_cell is 16 doubles");
  detail::Writer w;
  w.raw("TBVG", 4);
  w.put<uint32_t>(1);
  w.put<uint32_t>((uint32_t)graph.size());
  for (const auto& nc : graph) {
    const RadarScan& s = nc.first;
    detail::put_pose(w, s.T);
    detail::put_pose(w, s.Tgt);
    w.put<uint8_t>(s.has_Tgt_ ? 1 : 0);
    w.put<uint32_t>(s.idx_);
    w.put<uint64_t>((uint64_t)s.stamp_);
    w.raw(s.motion_.data(), 128);
    detail::put_cloud(w, s.cloud_peaks_);
    detail::put_cloud(w, s.cloud_nopeaks_);
    w.put<uint32_t>((uint32_t)s.cloud_normal_.size());
    if (!s.cloud_normal_.empty()) w.raw(s.cloud_normal_.data(), s.cloud_normal_.size() * sizeof(tbv_cell));
    w.put<uint32_t>((uint32_t)s.downsampled_.size());
    for (const auto& d : s.downsampled_) { w.put(d[0]); w.put(d[1]); }
    w.put<float>(s.radius_);
    w.put<uint8_t>(s.weight_intensity_ ? 1 : 0);
    w.put<uint32_t>((uint32_t)nc.second.size());
    for (const GraphConstraint& c : nc.second) {
      w.put<uint64_t>((uint64_t)c.id_begin);
      w.put<uint64_t>((uint64_t)c.id_end);
      detail::put_pose(w, c.t_be);
      w.raw(c.information.data(), 288);
      w.put<uint32_t>((uint32_t)c.type);
      std::vector<std::pair<std::string, double>> q = c.quality;
      std::sort(q.begin(), q.end(), [](const std::pair<std::string, double>& a, const std::pair<std::string, double>& b) { return a.first < b.first; });
      w.put<uint32_t>((uint32_t)q.size());
      for (const auto& kv : q) { w.put<uint16_t>((uint16_t)kv.first.size()); w.raw(kv.first.data(), kv.first.size()); w.put<double>(kv.second); }
      w.put<uint32_t>((uint32_t)c.info.size());
      w.raw(c.info.data(), c.info.size());
    }
  }
  return w.out;
}

inline simple_graph ParseSimpleGraph(const std::string& bytes) {
  detail::Reader r(bytes);
  char magic[4];
  r.raw(magic, 4);
  if (std::string(magic, 4) != "TBVG") throw Error(TBV_ERR_INVALID, "not a .tbvg file");
  if (r.get<uint32_t>() != 1) throw Error(TBV_ERR_INVALID, "unsupported .tbvg version");
  simple_graph graph(r.count<uint32_t>(1));
  for (auto& nc : graph) {
    RadarScan& s = nc.first;
    detail::get_pose(r, s.T);
    detail::get_pose(r, s.Tgt);
    s.has_Tgt_ = r.get<uint8_t>() != 0;
    s.idx_ = r.get<uint32_t>();
    s.stamp_ = (unsigned long)r.get<uint64_t>();
    r.raw(s.motion_.data(), 128);
    detail::get_cloud(r, s.cloud_peaks_);
    detail::get_cloud(r, s.cloud_nopeaks_);
    s.cloud_normal_.resize(r.count<uint32_t>(sizeof(tbv_cell)));
    if (!s.cloud_normal_.empty()) r.raw(s.cloud_normal_.data(), s.cloud_normal_.size() * sizeof(tbv_cell));
    s.downsampled_.resize(r.count<uint32_t>(8));
    for (auto& d : s.downsampled_) { d[0] = r.get<float>(); d[1] = r.get<float>(); }
    s.radius_ = r.get<float>();
    s.weight_intensity_ = r.get<uint8_t>() != 0;
    nc.second.resize(r.count<uint32_t>(1));
    for (GraphConstraint& c : nc.second) {
      c.id_begin = (unsigned long)r.get<uint64_t>();
      c.id_end = (unsigned long)r.get<uint64_t>();
      detail::get_pose(r, c.t_be);
      r.raw(c.information.data(), 288);
      c.type = (int)r.get<uint32_t>();
      c.quality.resize(r.count<uint32_t>(10));
      for (auto& kv : c.quality) {
        kv.first.resize(r.count<uint16_t>(1));
        if (!kv.first.empty()) r.raw(&kv.first[0], kv.first.size());
        kv.second = r.get<double>();
      }
      c.info.resize(r.count<uint32_t>(1));
      if (!c.info.empty()) r.raw(&c.info[0], c.info.size());
    }
  }
  if (r.o != bytes.size()) throw Error(TBV_ERR_INVALID, "trailing bytes in .tbvg file");
  return graph;
}

// The constraints the optimiser takes, with ids mapped to rows (CeresLeastSquares reads nodes by id; here: by position in the graph)
inline void GraphToOptimizerInput(const simple_graph& graph, std::vector<Pose3d>& nodes, std::vector<Constraint3d>& constraints) {
  nodes.clear();
  constraints.clear();
  std::vector<std::pair<unsigned int, size_t>> rows;
  for (size_t i = 0; i < graph.size(); i++) { nodes.push_back(graph[i].first.T); rows.emplace_back(graph[i].first.idx_, i); }
  std::sort(rows.begin(), rows.end());
  auto row_of = [&](unsigned long id) -> unsigned long {
    auto it = std::lower_bound(rows.begin(), rows.end(), std::make_pair((unsigned int)id, (size_t)0));
    if (it == rows.end() || it->first != id) throw Error(TBV_ERR_INVALID, "constraint references a missing node");
    return (unsigned long)it->second;
  };
  for (const auto& nc : graph)
    for (const GraphConstraint& c : nc.second) {
      if (c.type != 0 && c.type != 1) continue;
      Constraint3d o;
      o.id_begin = row_of(c.id_begin); o.id_end = row_of(c.id_end); o.t_be = c.t_be; o.information = c.information; o.type = c.type;
      constraints.push_back(o);
    }
}

}  // namespace tbv_b200

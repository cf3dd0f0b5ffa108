// oracle_capi.cpp — extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline /
// --impl reference).  *** TEST INFRASTRUCTURE ONLY *** — never linked into libtbv_b200.so.
//
// Cell records cross this boundary as 16 consecutive doubles:
//   [0]u0 [1]u1 [2]c00 [3]c01 [4]c10 [5]c11 [6]scale [7]n0 [8]n1 [9]o0 [10]o1 [11]lambda_min [12]lambda_max
//   [13]sum_intensity [14]avg_intensity [15]Nsamples
// Poses cross it as (x, y, theta).
#include <atomic>
#include <chrono>
#include <thread>

#include "tbv_oracle_loop.hpp"
#include "tbv_oracle_coral.hpp"
#include "tbv_oracle_reg.hpp"

using namespace tbv_oracle;

namespace {
void CellToRec(const Cell& c, double* r) {
  r[0] = c.u[0]; r[1] = c.u[1];
  r[2] = c.cov[0][0]; r[3] = c.cov[0][1]; r[4] = c.cov[1][0]; r[5] = c.cov[1][1];
  r[6] = c.scale;
  r[7] = c.snormal[0]; r[8] = c.snormal[1];
  r[9] = c.orth_normal[0]; r[10] = c.orth_normal[1];
  r[11] = c.lambda_min; r[12] = c.lambda_max;
  r[13] = c.sum_intensity; r[14] = c.avg_intensity;
  r[15] = (double)c.Nsamples;
}
Cell RecToCell(const double* r) {
  Cell c;
  c.u[0] = r[0]; c.u[1] = r[1];
  c.cov[0][0] = r[2]; c.cov[0][1] = r[3]; c.cov[1][0] = r[4]; c.cov[1][1] = r[5];
  c.scale = r[6];
  c.snormal[0] = r[7]; c.snormal[1] = r[8];
  c.orth_normal[0] = r[9]; c.orth_normal[1] = r[10];
  c.lambda_min = r[11]; c.lambda_max = r[12];
  c.sum_intensity = r[13]; c.avg_intensity = r[14];
  c.Nsamples = (uint64_t)r[15];
  c.valid = true;
  return c;
}
MapNormalPtr MapFromRecs(const double* recs, int n, float radius) {
  std::vector<Cell> cs(n);
  for (int i = 0; i < n; i++) cs[i] = RecToCell(recs + (size_t)i * 16);
  return MapNormalPtr(new MapPointNormal(cs, radius));
}
int CopyPolar(const std::vector<PolarPoint>& v, int cap, uint16_t* az, uint16_t* rg, uint8_t* inten, float* x, float* y) {
  const int n = (int)std::min((size_t)cap, v.size());
  for (int i = 0; i < n; i++) {
    if (az) az[i] = v[i].azimuth;
    if (rg) rg[i] = v[i].range;
    if (inten) inten[i] = v[i].intensity;
    if (x) x[i] = v[i].x;
    if (y) y[i] = v[i].y;
  }
  return (int)v.size();
}
}  // namespace

extern "C" {

// ---- filters --------------------------------------------------------------------------------
// returns number of filtered points; *n_peaks receives the number of peak points. Arrays hold up to `cap`.
int orc_kstrongest(const uint8_t* img, int n_az, int n_range, long row_stride, float z_min, int k, float min_distance, float range_res,
                   int cap, uint16_t* az, uint16_t* rg, uint8_t* inten, float* x, float* y,
                   int* n_peaks, uint16_t* paz, uint16_t* prg, uint8_t* pinten, float* px, float* py) {
  KStrongestOutput o;
  StructuredKStrongest(img, n_az, n_range, (size_t)row_stride, z_min, k, min_distance, range_res, o, n_peaks != nullptr);
  if (n_peaks) *n_peaks = CopyPolar(o.polar_peaks, cap, paz, prg, pinten, px, py);
  return CopyPolar(o.polar, cap, az, rg, inten, x, y);
}

int orc_cacfar(const uint8_t* img, int n_az, int n_range, long row_stride, int window_size, double false_alarm_rate, int nb_guard_cells,
               double range_res, double static_threshold, double min_distance, double max_distance,
               int cap, uint16_t* az, uint16_t* rg, uint8_t* inten, float* x, float* y) {
  CFARParams p;
  p.window_size = window_size; p.false_alarm_rate = false_alarm_rate; p.nb_guard_cells = nb_guard_cells;
  p.range_resolution = range_res; p.static_threshold = static_threshold; p.min_distance = min_distance; p.max_distance = max_distance;
  Cloud c;
  std::vector<PolarPoint> pol;
  AzimuthCACFAR(img, n_az, n_range, (size_t)row_stride, p, c, &pol);
  return CopyPolar(pol, cap, az, rg, inten, x, y);
}

void orc_rotate90ccw(const uint8_t* src, int H, int W, uint8_t* dst) {
  std::vector<uint8_t> d;
  Rotate90CCW(src, H, W, d);
  std::memcpy(dst, d.data(), d.size());
}

// ---- compensation ---------------------------------------------------------------------------
void orc_compensate(float* x, float* y, int n, const double mot[3], int ccw) {
  Cloud c(n);
  for (int i = 0; i < n; i++) { c[i].x = x[i]; c[i].y = y[i]; }
  Compensate(c, mot, ccw != 0);
  for (int i = 0; i < n; i++) { x[i] = c[i].x; y[i] = c[i].y; }
}

// ---- cells ----------------------------------------------------------------------------------
// returns number of valid cells (records written up to cap); *n_samples = voxel-grid sample points
int orc_build_cells(const float* x, const float* y, const float* intensity, int n, float radius, double downsample_factor,
                    int weight_intensity, const double origin[2], int voxel_order, int cap, double* recs, int* n_samples) {
  Cloud c(n);
  for (int i = 0; i < n; i++) { c[i].x = x[i]; c[i].y = y[i]; c[i].z = 0; c[i].intensity = intensity[i]; }
  MapPointNormal m(c, radius, origin, weight_intensity != 0, downsample_factor, (VoxelOrder)voxel_order);
  if (n_samples) *n_samples = m.n_samples;
  const int nc = (int)m.GetSize();
  for (int i = 0; i < nc && i < cap; i++) CellToRec(m.GetCell(i), recs + (size_t)i * 16);
  return nc;
}

// voxel-grid centroids only (x,y,intensity), returns count
int orc_voxel_centroids(const float* x, const float* y, const float* intensity, int n, float leaf, int voxel_order, int cap, float* cx, float* cy, float* ci) {
  Cloud c(n);
  for (int i = 0; i < n; i++) { c[i].x = x[i]; c[i].y = y[i]; c[i].z = 0; c[i].intensity = intensity[i]; }
  VoxelGridResult vg;
  if (!VoxelGrid(c, leaf, vg, (VoxelOrder)voxel_order)) return 0;
  const int m = (int)vg.centroids.size();
  for (int i = 0; i < m && i < cap; i++) { cx[i] = vg.centroids[i].x; cy[i] = vg.centroids[i].y; ci[i] = vg.centroids[i].intensity; }
  return m;
}

void orc_eig2(double m00, double m10, double m11, double eval[2], double evec[4]) {
  Eig2 e = SelfAdjointEig2(m00, m10, m11);
  eval[0] = e.eval[0]; eval[1] = e.eval[1];
  evec[0] = e.evec[0][0]; evec[1] = e.evec[0][1]; evec[2] = e.evec[1][0]; evec[3] = e.evec[1][1];
}

// 1-NN over cell means (bucket search and brute force) — a9
int orc_closest_idx(const double* recs, int n, float radius, double px, double py, double d, int brute) {
  MapNormalPtr m = MapFromRecs(recs, n, radius);
  return brute ? m->GetClosestIdxBrute(px, py, d) : m->GetClosestIdx(px, py, d);
}

// ---- registration ---------------------------------------------------------------------------
struct orc_reg_params {
  int cost, loss, weight_opt;
  double loss_limit, cov_scale, regularization;
  int max_itr_association, max_itr_solver;  // <=0: keep class defaults (8, 20)
};
struct orc_reg_summary {
  int success, itrs, lm_iterations, num_residuals, last_n_iterations, termination;
  double score, final_cost, last_relative_decrease;
};
static n_scan_normal_reg MakeReg(const orc_reg_params* p) {
  n_scan_normal_reg reg(p->cost, p->loss, p->loss_limit, p->weight_opt);
  reg.SetD2dPar(p->cov_scale, p->regularization);
  if (p->max_itr_association > 0 && p->max_itr_solver > 0) reg.SetParameters(p->max_itr_association, p->max_itr_solver);
  return reg;
}
// scans: n_scans cell-record arrays (last = moving scan); T: n_scans x 3 (x,y,theta) in/out
int orc_register(int n_scans, const double* const* recs, const int* n_cells, double* T, const orc_reg_params* p, orc_reg_summary* s) {
  std::vector<MapNormalPtr> scans;
  std::vector<Affine2> Tv;
  for (int i = 0; i < n_scans; i++) {
    scans.push_back(MapFromRecs(recs[i], n_cells[i], 0.f));
    Tv.push_back(vectorToAffine(T[3 * i], T[3 * i + 1], T[3 * i + 2]));
  }
  n_scan_normal_reg reg = MakeReg(p);
  const bool ok = reg.Register(scans, Tv);
  for (int i = 0; i < n_scans; i++) AffineToVector(Tv[i], T + 3 * i);
  if (s) {
    s->success = ok; s->itrs = (int)reg.itr_; s->lm_iterations = reg.total_lm_iterations_;
    s->num_residuals = reg.summary_.num_residuals; s->score = ok ? reg.getScore() : 0.0; s->final_cost = reg.summary_.final_cost;
    s->last_n_iterations = (int)reg.summary_.iterations.size();
    s->last_relative_decrease = reg.summary_.iterations.empty() ? 0.0 : reg.summary_.iterations.back().relative_decrease;
    s->termination = reg.summary_.termination_type;
  }
  return ok ? 1 : 0;
}
// GetCost (a14): returns #residuals (or -1 on failure); residuals written up to cap
int orc_get_cost(int n_scans, const double* const* recs, const int* n_cells, const double* T, const orc_reg_params* p, int itr,
                 double* score, double* cost, int cap, double* residuals) {
  std::vector<MapNormalPtr> scans;
  std::vector<Affine2> Tv;
  for (int i = 0; i < n_scans; i++) {
    scans.push_back(MapFromRecs(recs[i], n_cells[i], 0.f));
    Tv.push_back(vectorToAffine(T[3 * i], T[3 * i + 1], T[3 * i + 2]));
  }
  n_scan_normal_reg reg = MakeReg(p);
  reg.itr_ = (size_t)itr;
  double c = 0;
  std::vector<double> res;
  if (!reg.GetCost(scans, Tv, c, res)) return -1;
  if (cost) *cost = c;
  if (score) *score = reg.getScore();
  for (size_t i = 0; i < res.size() && (int)i < cap; i++) residuals[i] = res[i];
  return (int)res.size();
}
// OdometryKeyframeFuser::approximateCovarianceBySampling (odometrykeyframefuser.cpp:261-380), the sampling half: n^3 GetCost
// evaluations around T.back() on the grid linspace(-xy/2, xy/2, n)^2 x linspace(-yaw/2, yaw/2, n), theta-major then x then y (:293-296).
// samples: [n^3][4] = (dx, dy, dyaw, cost); a failed GetCost keeps the previous sample's cost (sample_cost is not reset, :283,304).
// Returns the number of samples.
static std::vector<double> orc_linspace(double start, double end, int num) {  // odometrykeyframefuser.cpp:498-524
  std::vector<double> v;
  if (num == 0) return v;
  if (num == 1) { v.push_back(start); return v; }
  const double delta = (end - start) / ((double)num - 1);
  for (int i = 0; i < num - 1; ++i) v.push_back(start + delta * i);
  v.push_back(end);
  return v;
}
int orc_cost_samples(int n_scans, const double* const* recs, const int* n_cells, const double* T, const orc_reg_params* p, int itr,
                     double xy_range, double yaw_range, int n_per_axis, double* samples) {
  std::vector<MapNormalPtr> scans;
  std::vector<Affine2> Tv;
  for (int i = 0; i < n_scans; i++) {
    scans.push_back(MapFromRecs(recs[i], n_cells[i], 0.f));
    Tv.push_back(vectorToAffine(T[3 * i], T[3 * i + 1], T[3 * i + 2]));
  }
  n_scan_normal_reg reg = MakeReg(p);
  const Affine2 Tbest = Tv.back();
  const std::vector<double> xy = orc_linspace(-xy_range * 0.5, xy_range * 0.5, n_per_axis);
  const std::vector<double> th = orc_linspace(-yaw_range * 0.5, yaw_range * 0.5, n_per_axis);
  double sample_cost = 0;
  std::vector<double> res;
  int k = 0;
  for (int it = 0; it < n_per_axis; it++)
    for (int ix = 0; ix < n_per_axis; ix++)
      for (int iy = 0; iy < n_per_axis; iy++) {
        Affine2 S;  // translation = (dx, dy) + T_best.translation ; linear = AngleAxis(dtheta, z) * T_best.linear
        const double c = std::cos(th[it]), sn = std::sin(th[it]);
        S.r00 = c * Tbest.r00 + (-sn) * Tbest.r10; S.r01 = c * Tbest.r01 + (-sn) * Tbest.r11;
        S.r10 = sn * Tbest.r00 + c * Tbest.r10;    S.r11 = sn * Tbest.r01 + c * Tbest.r11;
        S.tx = xy[ix] + Tbest.tx; S.ty = xy[iy] + Tbest.ty;
        Tv.back() = S;
        reg.itr_ = (size_t)itr;
        reg.GetCost(scans, Tv, sample_cost, res);
        samples[4 * k + 0] = xy[ix]; samples[4 * k + 1] = xy[iy]; samples[4 * k + 2] = th[it]; samples[4 * k + 3] = sample_cost;
        k++;
      }
  return k;
}
// One (target, source) pair: associate at the given poses with search radius chosen by `itr` (1 -> 2*radius_),
// then evaluate cost, gradient g = J^T r and H = J^T J (robustified, unscaled) at x = T_src.  assoc (optional,
// length n_src): target index per source cell or -1.
int orc_pair_normal_eq(const double* tgt, int n_tgt, const double T_tgt[3], const double* src, int n_src, const double T_src[3],
                       const orc_reg_params* p, int itr, double* cost, int* n_res, double H[9], double g[3], int* assoc) {
  std::vector<MapNormalPtr> scans{MapFromRecs(tgt, n_tgt, 0.f), MapFromRecs(src, n_src, 0.f)};
  n_scan_normal_reg reg = MakeReg(p);
  reg.itr_ = (size_t)itr;
  reg.problem_ = Problem();
  reg.problem_.loss = p->loss;
  reg.problem_.loss_limit = p->loss_limit;
  const Affine2 Tt = vectorToAffine(T_tgt[0], T_tgt[1], T_tgt[2]), Ts = vectorToAffine(T_src[0], T_src[1], T_src[2]);
  reg.AddScanPairCost(*scans[0], *scans[1], Tt, Ts, 0);
  if (assoc) {
    for (int i = 0; i < n_src; i++) assoc[i] = -1;
    for (const auto& b : reg.problem_.blocks) assoc[b.src_idx] = b.tar_idx;
  }
  std::vector<double> res, jac;
  double c = 0, grad[3];
  reg.problem_.Evaluate(T_src, &c, &res, grad, &jac);
  if (cost) *cost = c;
  if (n_res) *n_res = (int)res.size();
  for (int a = 0; a < 3; a++) {
    g[a] = grad[a];
    for (int b = 0; b < 3; b++) {
      double h = 0;
      for (size_t r = 0; r < res.size(); r++) h += jac[r * 3 + a] * jac[r * 3 + b];
      H[a * 3 + b] = h;
    }
  }
  return (int)reg.problem_.blocks.size();
}

// loop-candidate registration (a16). Returns success; Talign and Trevised as (x,y,theta)
int orc_loop_register(const double* from, int n_from, const double* to, int n_to, const double Tfrom[3], const double Tto[3],
                      double Talign[3], double Trevised[3], int* itrs, double* score) {
  Affine2 Ta, Tr;
  const bool ok = LoopRegister(MapFromRecs(from, n_from, 0.f), MapFromRecs(to, n_to, 0.f), vectorToAffine(Tfrom[0], Tfrom[1], Tfrom[2]),
                               vectorToAffine(Tto[0], Tto[1], Tto[2]), Ta, &Tr, itrs, score);
  if (ok && Talign) AffineToVector(Ta, Talign);
  if (Trevised) AffineToVector(Tr, Trevised);
  return ok ? 1 : 0;
}

// ---- odometry (a15): filter -> compensate -> cells -> register -> keyframe policy -------------
struct orc_odom_params {
  float z_min; int k_strongest; float min_distance; float range_res;
  int cost_type, loss_type, weight_opt;
  double loss_limit, covar_scale, regularization;
  int submap_scan_size, weight_intensity, use_guess, compensate, radar_ccw, use_keyframe;
  double res, min_keyframe_dist, min_keyframe_rot_deg, downsample_factor;
  int voxel_order;
};
struct OdomHandle {
  orc_odom_params p;
  OdometryKeyframeFuser* fuser;
};
static FuserParameters ToFuser(const orc_odom_params& p) {
  FuserParameters f;
  f.cost_type = p.cost_type; f.weight_opt = p.weight_opt; f.submap_scan_size = p.submap_scan_size;
  f.weight_intensity = p.weight_intensity != 0; f.use_guess = p.use_guess != 0; f.compensate = p.compensate != 0;
  f.radar_ccw = p.radar_ccw != 0; f.use_keyframe = p.use_keyframe != 0; f.res = p.res;
  f.min_keyframe_dist = p.min_keyframe_dist; f.min_keyframe_rot_deg = p.min_keyframe_rot_deg; f.loss_type = p.loss_type;
  f.loss_limit = p.loss_limit; f.covar_scale = p.covar_scale; f.regularization = p.regularization;
  f.downsample_factor = p.downsample_factor; f.voxel_order = p.voxel_order;
  return f;
}
void* orc_odom_create(const orc_odom_params* p) {
  OdomHandle* h = new OdomHandle();
  h->p = *p;
  h->fuser = new OdometryKeyframeFuser(ToFuser(*p));
  return h;
}
void orc_odom_destroy(void* hv) {
  OdomHandle* h = (OdomHandle*)hv;
  if (!h) return;
  delete h->fuser;
  delete h;
}
struct orc_odom_out {
  double pose[3];
  int n_points, n_cells, itrs, reg_ok, is_keyframe, n_keyframes;
  double ms_filter, ms_compensate_normals, ms_register;
};
static void OdomStep(OdomHandle* h, const uint8_t* img, int n_az, int n_range, long row_stride, orc_odom_out* o) {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  KStrongestOutput k;
  StructuredKStrongest(img, n_az, n_range, (size_t)row_stride, h->p.z_min, h->p.k_strongest, h->p.min_distance, h->p.range_res, k, true);
  auto t1 = clk::now();
  Affine2 T = h->fuser->processFrame(k.cloud, &k.cloud_peaks);
  auto t2 = clk::now();
  if (o) {
    AffineToVector(T, o->pose);
    o->n_points = (int)k.cloud.size(); o->n_cells = h->fuser->last_n_cells; o->itrs = h->fuser->last_itrs;
    o->reg_ok = h->fuser->last_reg_ok; o->is_keyframe = h->fuser->updated; o->n_keyframes = (int)h->fuser->keyframes_.size();
    o->ms_filter = std::chrono::duration<double, std::milli>(t1 - t0).count();
    o->ms_compensate_normals = 0;
    o->ms_register = std::chrono::duration<double, std::milli>(t2 - t1).count();
  }
}
void orc_odom_step(void* hv, const uint8_t* img, int n_az, int n_range, long row_stride, orc_odom_out* o) {
  OdomStep((OdomHandle*)hv, img, n_az, n_range, row_stride, o);
}
// current keyframe window (for parity checks): returns #keyframes; pose (x,y,theta) per keyframe and cell counts
int orc_odom_keyframes(void* hv, int cap, double* poses, int* n_cells) {
  OdomHandle* h = (OdomHandle*)hv;
  int n = (int)h->fuser->keyframes_.size();
  for (int i = 0; i < n && i < cap; i++) {
    AffineToVector(h->fuser->keyframes_[i].pose, poses + 3 * i);
    n_cells[i] = (int)h->fuser->keyframes_[i].normals->GetSize();
  }
  return n;
}
int orc_odom_keyframe_cells(void* hv, int kf, int cap, double* recs) {
  OdomHandle* h = (OdomHandle*)hv;
  if (kf < 0 || kf >= (int)h->fuser->keyframes_.size()) return -1;
  const MapPointNormal& m = *h->fuser->keyframes_[kf].normals;
  for (size_t i = 0; i < m.GetSize() && (int)i < cap; i++) CellToRec(m.GetCell(i), recs + i * 16);
  return (int)m.GetSize();
}

// Many independent sequences over a pool of scans, spread over n_threads host threads (the reference's own
// scaling model is process-level workers, tbv_slam/python/eval.py:44-47).  Sequence s processes pool scans
// first[s], first[s]+1, ..., first[s]+n_frames-1.  poses: [n_seq][n_frames][3].  Returns wall seconds.
double orc_odom_run(const orc_odom_params* p, const uint8_t* pool, int n_pool, int n_az, int n_range, int n_seq, const int* first,
                    int n_frames, int n_threads, double* poses) {
  using clk = std::chrono::steady_clock;
  std::atomic<int> next(0);
  const size_t scan_bytes = (size_t)n_az * n_range;
  auto worker = [&]() {
    for (;;) {
      const int s = next.fetch_add(1);
      if (s >= n_seq) break;
      OdomHandle h;
      h.p = *p;
      OdometryKeyframeFuser fuser(ToFuser(*p));
      h.fuser = &fuser;
      for (int f = 0; f < n_frames; f++) {
        const int idx = (first[s] + f) % n_pool;
        orc_odom_out o;
        OdomStep(&h, pool + scan_bytes * idx, n_az, n_range, n_range, &o);
        if (poses) std::memcpy(poses + ((size_t)s * n_frames + f) * 3, o.pose, 3 * sizeof(double));
      }
    }
  };
  auto t0 = clk::now();
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, n_threads); i++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(clk::now() - t0).count();
}

// Same, timing only the frames after the first n_warm of every sequence (the GPU arm's warm-up steps): each thread sums the
// wall time of its own timed frames; returns the maximum over threads (all threads are busy throughout), i.e. the time in
// which n_seq * (n_frames - n_warm) scans were registered.  thread_seconds (optional, n_threads entries) receives the sums.
double orc_odom_run_timed(const orc_odom_params* p, const uint8_t* pool, int n_pool, int n_az, int n_range, int n_seq, const int* first,
                          int n_warm, int n_frames, int n_threads, double* poses, double* thread_seconds) {
  using clk = std::chrono::steady_clock;
  std::atomic<int> next(0);
  const size_t scan_bytes = (size_t)n_az * n_range;
  const int nt = std::max(1, n_threads);
  std::vector<double> acc(nt, 0.0);
  auto worker = [&](int tid) {
    for (;;) {
      const int s = next.fetch_add(1);
      if (s >= n_seq) break;
      OdomHandle h;
      h.p = *p;
      OdometryKeyframeFuser fuser(ToFuser(*p));
      h.fuser = &fuser;
      for (int f = 0; f < n_frames; f++) {
        const int idx = (first[s] + f) % n_pool;
        orc_odom_out o;
        const auto t0 = clk::now();
        OdomStep(&h, pool + scan_bytes * idx, n_az, n_range, n_range, &o);
        if (f >= n_warm) acc[tid] += std::chrono::duration<double>(clk::now() - t0).count();
        if (poses) std::memcpy(poses + ((size_t)s * n_frames + f) * 3, o.pose, 3 * sizeof(double));
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nt; i++) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
  double mx = 0.0;
  for (int i = 0; i < nt; i++) { mx = std::max(mx, acc[i]); if (thread_seconds) thread_seconds[i] = acc[i]; }
  return mx;
}

// ---- Ceres-restated pieces exposed for unit tests ------------------------------------------------
void orc_loss(int loss, double loss_limit, double weight, double s, double rho[3]) { ceres_restated::ScaledLoss(loss, loss_limit, weight, s, rho); }

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

#include "oracle_capi_loop.inc"

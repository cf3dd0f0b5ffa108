// CPU test of tbv_b200::PointCloudOdometryFuserT (include/tbv_b200.hpp): processFrame's host bookkeeping with the oracle's primitives (compensation,
// surface points, registration through the oracle's C surface) standing in for the three device calls, against the oracle's own fused frame
// (orc_odom_step — the checker of tbv_odom_step) on the same scans.  Test infrastructure only.  Usage: test_points_fuser scans.bin n n_az n_range
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "tbv_b200.hpp"
#include "oracle_capi.cpp"   // the oracle's extern "C" surface (orc_*), compiled into this test

struct OraclePrimitives {
  void compensate(std::vector<float>& x, std::vector<float>& y, const double mot[3], bool ccw) const {
    if (!x.empty()) orc_compensate(x.data(), y.data(), (int)x.size(), mot, ccw ? 1 : 0);
  }
  std::vector<tbv_cell> build_cells(const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& I, float radius, double downsample,
                                    bool weight_intensity) const {
    static_assert(sizeof(tbv_cell) == 16 * sizeof(double), "a cell record is 16 doubles on both sides");
    std::vector<tbv_cell> cells(std::max<size_t>(x.size(), 1));
    const double origin[2] = {0, 0};
    int ns = 0;
    const int n = x.empty() ? 0 : orc_build_cells(x.data(), y.data(), I.data(), (int)x.size(), radius, downsample, weight_intensity ? 1 : 0, origin, 0 /*stable order*/,
                                                  (int)cells.size(), reinterpret_cast<double*>(cells.data()), &ns);
    cells.resize(n);
    return cells;
  }
  void register_scans(const std::vector<const tbv_cell*>& scans, const std::vector<int>& n_cells, std::vector<double>& T, const tbv_reg_params& par,
                      tbv_reg_summary& summary) const {
    static_assert(sizeof(orc_reg_params) == sizeof(tbv_reg_params) && sizeof(orc_reg_summary) == sizeof(tbv_reg_summary), "same records");
    std::vector<const double*> recs;
    for (const tbv_cell* c : scans) recs.push_back(reinterpret_cast<const double*>(c));
    orc_register((int)scans.size(), recs.data(), n_cells.data(), T.data(), reinterpret_cast<const orc_reg_params*>(&par), reinterpret_cast<orc_reg_summary*>(&summary));
  }
};

int main(int argc, char** argv) {
  if (argc != 5) return 2;
  const int n = std::atoi(argv[2]), n_az = std::atoi(argv[3]), n_range = std::atoi(argv[4]);
  std::vector<uint8_t> scans((size_t)n * n_az * n_range);
  std::ifstream f(argv[1], std::ios::binary);
  f.read(reinterpret_cast<char*>(scans.data()), (std::streamsize)scans.size());
  if (!f) { std::printf("cannot read %s\n", argv[1]); return 2; }

  tbv_odom_params par{};                                   // BASELINE config 2 (api.default_odom_params)
  par.filter = tbv_filter_params{60.0f, 40, 2.5f, 0.0438f};
  par.reg = tbv_reg_params{TBV_P2L, TBV_LOSS_HUBER, TBV_W_COMBINED, 0.1, 1.0, 1.0, 0, 0};
  par.submap_scan_size = 4; par.weight_intensity = 1; par.use_guess = 1; par.compensate = 1; par.radar_ccw = 0; par.use_keyframe = 1;
  par.res = 3.0; par.min_keyframe_dist = 1.5; par.min_keyframe_rot_deg = 5.0; par.downsample_factor = 1.0;
  orc_odom_params op{60.0f, 40, 2.5f, 0.0438f, TBV_P2L, TBV_LOSS_HUBER, TBV_W_COMBINED, 0.1, 1.0, 1.0, 4, 1, 1, 1, 0, 1, 3.0, 1.5, 5.0, 1.0, 0};
  void* ref = orc_odom_create(&op);
  tbv_b200::PointCloudOdometryFuserT<OraclePrimitives> fuser(OraclePrimitives(), par);

  int fails = 0, keyframes = 0;
  double worst = 0;
  const int cap = n_az * 40;
  std::vector<uint16_t> az(cap), rg(cap); std::vector<uint8_t> I(cap); std::vector<float> x(cap), y(cap);
  for (int i = 0; i < n; i++) {
    const uint8_t* img = scans.data() + (size_t)i * n_az * n_range;
    const int np = orc_kstrongest(img, n_az, n_range, n_range, 60.0f, 40, 2.5f, 0.0438f, cap, az.data(), rg.data(), I.data(), x.data(), y.data(),
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    tbv_b200::PointCloud cloud(np);
    for (int k = 0; k < np; k++) { cloud[k].x = x[k]; cloud[k].y = y[k]; cloud[k].intensity = (float)I[k]; }
    const tbv_b200::Pose2 pose = fuser.pointcloudCallback(cloud);
    orc_odom_out o;
    orc_odom_step(ref, img, n_az, n_range, n_range, &o);
    const double d = std::fmax(std::fmax(std::fabs(pose.x - o.pose[0]), std::fabs(pose.y - o.pose[1])), std::fabs(pose.yaw - o.pose[2]));
    worst = std::fmax(worst, d);
    const bool same = d < 1e-12 && fuser.updated == (o.is_keyframe != 0) && fuser.last_itrs == o.itrs && (int)fuser.keyframes().size() == o.n_keyframes &&
                      (int)fuser.last_n_cells() == o.n_cells && np == o.n_points;
    if (!same) {
      std::printf("frame %d differs: d=%.3e keyframe %d/%d itrs %d/%d window %zu/%d cells %zu/%d\n", i, d, (int)fuser.updated, o.is_keyframe, fuser.last_itrs, o.itrs,
                  fuser.keyframes().size(), o.n_keyframes, fuser.last_n_cells(), o.n_cells);
      fails++;
    }
    keyframes += fuser.updated;
  }
  orc_odom_destroy(ref);
  std::printf("%d frames, %d keyframes, worst pose difference %.3e\n", n, keyframes, worst);
  if (fails || keyframes < 3 || keyframes >= n) { std::printf("%d FAILED\n", fails); return 1; }
  std::printf("PASS\n");
  return 0;
}

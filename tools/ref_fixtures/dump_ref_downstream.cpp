// dump_ref_downstream.cpp — runs the REFERENCE's own classes on the committed synthetic scans and writes what they compute, so that the
// stages downstream of the filter can be pinned to the reference itself instead of to this repo's oracle.
//
// This file is NOT built in this repository (Eigen / PCL / FLANN / Ceres / ROS are absent from the build container).  It is
// compiled inside the reference's own docker image (tbv_slam/docker/Dockerfile) against an unmodified checkout of dan11003/tbv_slam_public
// by tools/make_ref_fixtures.sh, which then converts its text output into tests/golden/ref_downstream.npz.  It only CALLS the reference:
//   radar_filters.h   StructuredKStrongest(image, z_min, k, min_distance, range_res) + getPeaksFilteredPointCloud(cloud, peaks)
//   utils.h           Compensate(cloud, motion (x, y, yaw), ccw)
//   pointnormal.h     MapPointNormal(cloud, radius, origin, weight_intensity) + GetCells()
//   n_scan_normal.h   n_scan_normal_reg(cost, loss, loss_limit, weight option) + SetParameters + Register + GetCost + getScore, itr_
//   RadarScancontext.h RSCManager(pars) + makeAndSaveScancontextAndKeysRadarCloud + detectLoopClosureID
//
// Input  (argv[1]): scans.bin = int32 n, n_az, n_range; n * n_az * n_range bytes; n * 3 doubles (ground-truth x, y, yaw of every scan).
// Output (argv[2]): a text file, one record per line: "<tag> <scan or pair ids> <count> <values ...>" with 17 significant digits.
#include <cstdio>
#include <fstream>
#include <vector>

#include <cv_bridge/cv_bridge.h>
#include <opencv2/core.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "cfear_radarodometry/n_scan_normal.h"
#include "cfear_radarodometry/pointnormal.h"
#include "cfear_radarodometry/radar_filters.h"
#include "cfear_radarodometry/registration.h"
#include "cfear_radarodometry/utils.h"
#include "place_recognition_radar/RadarScancontext.h"

using namespace CFEAR_Radarodometry;
typedef pcl::PointCloud<pcl::PointXYZI> Cloud;

static void put(FILE* f, const char* tag, int a, int b, const std::vector<double>& v) {
  std::fprintf(f, "%s %d %d %zu", tag, a, b, v.size());
  for (double x : v) std::fprintf(f, " %.17g", x);
  std::fprintf(f, "\n");
}

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: %s scans.bin out.txt\n", argv[0]); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  int n = 0, n_az = 0, n_range = 0;
  in.read((char*)&n, 4); in.read((char*)&n_az, 4); in.read((char*)&n_range, 4);
  std::vector<unsigned char> bytes((size_t)n * n_az * n_range);
  in.read((char*)bytes.data(), (std::streamsize)bytes.size());
  std::vector<double> gt((size_t)n * 3);
  in.read((char*)gt.data(), (std::streamsize)(gt.size() * sizeof(double)));
  if (!in) { std::printf("short read\n"); return 1; }
  FILE* out = std::fopen(argv[2], "w");

  const int z_min = 60, k_strongest = 40;                 // CFEAR-3 (oxford_cfear-3)
  const double min_distance = 2.5, range_res = 0.0438, radius = 3.0;
  std::vector<MapNormalPtr> maps;
  std::vector<Cloud::Ptr> peaks_clouds;
  std::vector<Eigen::Affine3d> T;
  for (int s = 0; s < n; s++) {
    cv_bridge::CvImagePtr img(new cv_bridge::CvImage);
    img->encoding = "mono8";
    img->image = cv::Mat(n_az, n_range, CV_8UC1, bytes.data() + (size_t)s * n_az * n_range).clone();
    StructuredKStrongest filt(img, z_min, k_strongest, min_distance, range_res);      // radar_driver.cpp:57-61
    Cloud::Ptr cloud(new Cloud), peaks(new Cloud);
    filt.getPeaksFilteredPointCloud(cloud, false);
    filt.getPeaksFilteredPointCloud(peaks, true);
    T.push_back(vectorToAffine3d(gt[3 * s], gt[3 * s + 1], 0, 0, 0, gt[3 * s + 2]));
    // motion of the previous frame pair (odometrykeyframefuser.cpp:146-150); ground truth stands in for the estimate
    std::vector<double> mot(3, 0.0);
    if (s > 0) { std::vector<double> v; Affine3dToVectorXYeZ(T[s - 1].inverse() * T[s], v); mot = v; }
    Compensate(*cloud, mot, false);
    Compensate(*peaks, mot, false);
    std::vector<double> pts;
    for (const auto& p : cloud->points) { pts.push_back(p.x); pts.push_back(p.y); pts.push_back(p.intensity); }
    put(out, "cloud", s, 0, pts);
    pts.clear();
    for (const auto& p : peaks->points) { pts.push_back(p.x); pts.push_back(p.y); pts.push_back(p.intensity); }
    put(out, "peaks", s, 0, pts);
    MapNormalPtr m(new MapPointNormal(cloud, (float)radius, Eigen::Vector2d(0, 0), true, false));   // odometrykeyframefuser.cpp:161
    std::vector<double> cv;
    for (const cell& c : m->GetCells()) {
      const double rec[] = {c.u_(0), c.u_(1), c.cov_(0, 0), c.cov_(0, 1), c.cov_(1, 0), c.cov_(1, 1), c.scale_, c.snormal_(0), c.snormal_(1),
                            c.lambda_min, c.lambda_max, c.sum_intensity_, c.avg_intensity_, (double)c.Nsamples_};
      cv.insert(cv.end(), rec, rec + 14);
    }
    put(out, "cells", s, 0, cv);
    maps.push_back(m);
    peaks_clouds.push_back(peaks);
  }
  // ---- Register: scan s against up to 4 earlier scans, guess = ground truth displaced by (0.3 m, -0.2 m, 0.02 rad) ------------------------
  for (int s = 1; s < n; s++) {
    std::vector<MapNormalPtr> scans;
    std::vector<Eigen::Affine3d> Ts;
    for (int t = std::max(0, s - 4); t < s; t++) { scans.push_back(maps[t]); Ts.push_back(T[t]); }
    scans.push_back(maps[s]);
    Ts.push_back(T[s] * vectorToAffine3d(0.3, -0.2, 0, 0, 0, 0.02));
    for (int variant = 0; variant < 2; variant++) {       // 0: odometry preset (P2L, Huber 0.1, combined weights); 1: loop preset (uniform, 4 x 10)
      n_scan_normal_reg reg(P2L, Huber, 0.1, variant == 0 ? weightoption::Combined : weightoption::Uniform);
      if (variant == 1) reg.SetParameters(4, 10);
      std::vector<Eigen::Affine3d> Tio = Ts;
      std::vector<Matrix6d> cov;
      const bool ok = reg.Register(scans, Tio, cov, false);
      std::vector<double> v; Affine3dToVectorXYeZ(Tio.back(), v);
      put(out, variant == 0 ? "register" : "register_loop", s, (int)scans.size(), {(double)ok, v[0], v[1], v[2], reg.getScore(), (double)reg.itr_});
    }
    // GetCost as CFEARQuality calls it (AlignmentQuality.cpp:336-344): P2L, Huber 0.3, uniform, at the ground-truth poses
    n_scan_normal_reg q(P2L, Huber, 0.3, weightoption::Uniform);
    std::vector<MapNormalPtr> pair = {maps[s - 1], maps[s]};
    std::vector<Eigen::Affine3d> Tp = {T[s - 1], T[s]};
    double score = 0; std::vector<double> residuals;
    const bool ok = q.GetCost(pair, Tp, score, residuals);
    std::vector<double> rec = {(double)ok, score, (double)residuals.size()};
    rec.insert(rec.end(), residuals.begin(), residuals.end());
    put(out, "get_cost", s, 0, rec);
  }
  // ---- Scan Context: every scan enters the database, candidates after each insertion ------------------------------------------------------
  {
    PlaceRecognitionRadar::RSCManager::Parameters par;    // defaults of RadarScancontext.h:35-58
    par.prints = false;
    PlaceRecognitionRadar::RSCManager rsc(par);
    for (int s = 0; s < n; s++) {
      Cloud::Ptr c(new Cloud(*peaks_clouds[s]));
      rsc.makeAndSaveScancontextAndKeysRadarCloud(c, T[s], true);
      const Eigen::MatrixXd& d = rsc.polarcontexts_.back();
      std::vector<double> dv(d.data(), d.data() + d.size());                        // column-major 40 x 120
      put(out, "sc_desc", s, (int)d.rows(), dv);
      std::vector<PlaceRecognitionRadar::candidate> cand = rsc.detectLoopClosureID();
      std::vector<double> cv;
      for (const auto& k : cand) { cv.push_back(k.nn_idx); cv.push_back(k.argmin_shift); cv.push_back(k.min_dist); cv.push_back(k.min_dist_sc); cv.push_back(k.min_dist_odom); cv.push_back(k.yaw_diff_rad); }
      put(out, "sc_cand", s, (int)cand.size(), cv);
    }
  }
  std::fclose(out);
  return 0;
}

// tests/cpp/test_host_mirror.cpp — the C++ host mirror (include/tbv_b200.hpp over libtbv_b200.so) against the CPU oracle, written the
// way a test of the reference's own classes would read: build the filter, the surface points, register, ask for the cost, run the fuser.
// TEST INFRASTRUCTURE: includes oracle/ headers as the checker.  Usage: test_host_mirror <scans.bin> <n_scans> <n_az> <n_range>
// (scans.bin: n_scans x n_az x n_range u8, written by tests/test_cpp_host_gpu.py).  Exit code 0 and a final "PASS" line on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include <cuda_runtime_api.h>

#include "tbv_b200.hpp"
#include "tbv_oracle.hpp"
#include "tbv_oracle_reg.hpp"
#include "tbv_oracle_loop.hpp"
#include "tbv_oracle_coral.hpp"

namespace gpu = tbv_b200;
namespace cpu = tbv_oracle;

static int g_checks = 0;
#define EXPECT(cond, ...)                                                        \
  do {                                                                           \
    g_checks++;                                                                  \
    if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); std::exit(1); } \
  } while (0)

static double ang(double d) { return std::fabs(std::atan2(std::sin(d), std::cos(d))); }

int main(int argc, char** argv) {
  if (argc < 5) { std::printf("usage: %s scans.bin n_scans n_az n_range\n", argv[0]); return 2; }
  const int n_scans = std::atoi(argv[2]), n_az = std::atoi(argv[3]), n_range = std::atoi(argv[4]);
  const size_t scan_bytes = (size_t)n_az * n_range;
  std::vector<uint8_t> scans(scan_bytes * n_scans);
  { std::ifstream f(argv[1], std::ios::binary); f.read((char*)scans.data(), (std::streamsize)scans.size()); EXPECT(f.gcount() == (std::streamsize)scans.size(), "short read"); }

  gpu::Context ctx(0);
  const int z_min = 60, k = 40;
  const double min_distance = 2.5, range_res = 0.0438;
  const float radius = 3.0f;

  // ---- radarDriver::Process: both clouds, bit for bit -----------------------------------------------------------------------------
  std::vector<gpu::MapNormalPtr> g_maps;
  std::vector<cpu::MapNormalPtr> c_maps;
  for (int s = 0; s < n_scans; s++) {
    const uint8_t* img = scans.data() + s * scan_bytes;
    gpu::StructuredKStrongest filt(ctx, img, n_az, n_range, (size_t)n_range, z_min, k, min_distance, range_res);
    gpu::PointCloud cloud, peaks;
    filt.getPeaksFilteredPointCloud(cloud, false);
    filt.getPeaksFilteredPointCloud(peaks, true);
    cpu::KStrongestOutput ref;
    cpu::StructuredKStrongest(img, n_az, n_range, (size_t)n_range, (float)z_min, k, (float)min_distance, (float)range_res, ref, true);
    EXPECT(cloud.size() == ref.cloud.size() && peaks.size() == ref.cloud_peaks.size(), "scan %d: %zu/%zu points vs %zu/%zu", s, cloud.size(), peaks.size(),
           ref.cloud.size(), ref.cloud_peaks.size());
    for (size_t i = 0; i < cloud.size(); i++)
      EXPECT(std::memcmp(&cloud[i].x, &ref.cloud[i].x, 4) == 0 && std::memcmp(&cloud[i].y, &ref.cloud[i].y, 4) == 0 && cloud[i].intensity == ref.cloud[i].intensity,
             "scan %d point %zu differs", s, i);
    for (size_t i = 0; i < peaks.size(); i++)
      EXPECT(std::memcmp(&peaks[i].x, &ref.cloud_peaks[i].x, 4) == 0 && std::memcmp(&peaks[i].y, &ref.cloud_peaks[i].y, 4) == 0, "scan %d peak %zu differs", s, i);

    // ---- MapPointNormal: same cells in the same order; statistics within the tolerance of DESIGN.md ------------------------------
    gpu::MapNormalPtr gm(new gpu::MapPointNormal(ctx, cloud, radius, {0.0, 0.0}, true));
    const double origin[2] = {0, 0};
    cpu::MapNormalPtr cm(new cpu::MapPointNormal(ref.cloud, radius, origin, true));
    EXPECT(gm->GetSize() == cm->GetSize() && gm->GetSize() > 50, "scan %d: %zu cells vs %zu", s, gm->GetSize(), cm->GetSize());
    for (size_t i = 0; i < gm->GetSize(); i++) {
      const tbv_cell& a = gm->GetCell(i);
      const cpu::Cell& b = cm->GetCell(i);
      EXPECT(std::fabs(a.u[0] - b.u[0]) < 1e-10 && std::fabs(a.u[1] - b.u[1]) < 1e-10 && a.n_samples == (double)b.Nsamples, "scan %d cell %zu differs", s, i);
      EXPECT(std::fabs(a.snormal[0] - b.snormal[0]) < 1e-7 && std::fabs(a.snormal[1] - b.snormal[1]) < 1e-7, "scan %d cell %zu normal differs", s, i);
    }
    g_maps.push_back(gm);
    c_maps.push_back(cm);
  }

  // ---- n_scan_normal_reg::Register / GetCost: scan s against scans s-1 (and s-2) -----------------------------------------------------
  for (int s = 1; s < n_scans; s++) {
    std::vector<gpu::MapNormalPtr> gs;
    std::vector<cpu::MapNormalPtr> cs;
    std::vector<gpu::Pose2> gT;
    std::vector<cpu::Affine2> cT;
    for (int t = std::max(0, s - 2); t <= s; t++) {
      gs.push_back(g_maps[t]); cs.push_back(c_maps[t]);
      const double x = 1.77 * t, y = 1.77 * t, yaw = 0.785 - 0.0001 * t;       // rough poses along the synthetic trajectory; the last is the guess
      gT.push_back(gpu::Pose2{x, y, yaw}); cT.push_back(cpu::vectorToAffine(x, y, yaw));
    }
    gpu::n_scan_normal_reg greg(ctx, gpu::P2L, gpu::Huber, 0.1, gpu::Combined_weights);
    cpu::n_scan_normal_reg creg(cpu::P2L, cpu::Huber, 0.1, cpu::Combined_weights);
    std::vector<gpu::Matrix6d> cov;
    const bool gok = greg.Register(gs, gT, &cov);
    const bool cok = creg.Register(cs, cT);
    double cp[3];
    cpu::AffineToVector(cT.back(), cp);
    EXPECT(gok == cok && gok, "Register scan %d: success %d vs %d", s, (int)gok, (int)cok);
    EXPECT(greg.itr_ == creg.itr_, "Register scan %d: %zu association rounds vs %zu", s, greg.itr_, creg.itr_);
    EXPECT(std::fabs(gT.back().x - cp[0]) < 1e-5 && std::fabs(gT.back().y - cp[1]) < 1e-5 && ang(gT.back().yaw - cp[2]) < 1e-6, "Register scan %d: pose (%.9f %.9f %.9f) vs (%.9f %.9f %.9f)",
           s, gT.back().x, gT.back().y, gT.back().yaw, cp[0], cp[1], cp[2]);
    EXPECT(std::fabs(greg.getScore() - creg.getScore()) <= 1e-9 * std::fabs(creg.getScore()), "Register scan %d: score", s);
    EXPECT(cov.size() == gs.size() && cov[0][0] == 0.1 * 0.1 && cov[0][35] == 0.01 * 0.01, "reg_cov");
    double gscore = 0, cscore = 0;
    std::vector<double> gres, cres;
    EXPECT(greg.GetCost(gs, gT, gscore, gres) && creg.GetCost(cs, cT, cscore, cres), "GetCost scan %d failed", s);
    EXPECT(gres.size() == cres.size() && std::fabs(gscore - cscore) <= 1e-9 * std::fabs(cscore), "GetCost scan %d: %zu residuals cost %.12g vs %zu, %.12g", s, gres.size(), gscore,
           cres.size(), cscore);
    double gsc = 0;
    EXPECT(greg.GetCovarianceScaler(gsc) && gsc > 0, "GetCovarianceScaler");
  }

  // ---- OdometryKeyframeFuser: the whole frame loop -----------------------------------------------------------------------------------
  {
    tbv_odom_params op;
    std::memset(&op, 0, sizeof(op));
    op.filter = tbv_filter_params{(float)z_min, k, (float)min_distance, (float)range_res};
    op.reg = tbv_reg_params{TBV_P2L, TBV_LOSS_HUBER, TBV_W_COMBINED, 0.1, 1.0, 1.0, 0, 0};
    op.submap_scan_size = 4; op.weight_intensity = 1; op.use_guess = 1; op.compensate = 1; op.radar_ccw = 0; op.use_keyframe = 1;
    op.res = 3.0; op.min_keyframe_dist = 1.5; op.min_keyframe_rot_deg = 5.0; op.downsample_factor = 1.0;
    gpu::OdometryKeyframeFuser gf(ctx, 1, n_az, n_range, op);
    cpu::FuserParameters fp;
    fp.cost_type = cpu::P2L; fp.weight_opt = cpu::Combined_weights; fp.submap_scan_size = 4; fp.weight_intensity = true; fp.res = 3.0;
    fp.loss_type = cpu::Huber; fp.loss_limit = 0.1; fp.covar_scale = 1.0; fp.regularization = 1.0;
    cpu::OdometryKeyframeFuser cf(fp);
    for (int s = 0; s < n_scans; s++) {
      const uint8_t* img = scans.data() + s * scan_bytes;
      const std::vector<tbv_odom_out> o = gf.pointcloudCallback(img);
      cpu::KStrongestOutput ref;
      cpu::StructuredKStrongest(img, n_az, n_range, (size_t)n_range, (float)z_min, k, (float)min_distance, (float)range_res, ref, true);
      const cpu::Affine2 Tc = cf.processFrame(ref.cloud, &ref.cloud_peaks);
      double cp[3];
      cpu::AffineToVector(Tc, cp);
      EXPECT(o[0].status == TBV_OK, "frame %d status %d", s, o[0].status);
      EXPECT(std::fabs(o[0].pose[0] - cp[0]) < 1e-5 && std::fabs(o[0].pose[1] - cp[1]) < 1e-5 && ang(o[0].pose[2] - cp[2]) < 1e-6, "fuser frame %d pose (%.9f %.9f %.9f) vs (%.9f %.9f %.9f)", s,
             o[0].pose[0], o[0].pose[1], o[0].pose[2], cp[0], cp[1], cp[2]);
    }
  }
  // ---- RSCManager: descriptors + keys per keyframe, loop candidates; CorAl quality of consecutive keyframes ---------------------------
  {
    tbv_sc_params sp = gpu::default_sc_params();
    gpu::RSCManager gsc(ctx, sp);
    cpu::SCParams cp;
    cp.N_CANDIDATES = sp.n_candidates; cp.desc_divider = sp.desc_divider; cp.no_point = sp.no_point;
    cpu::RSCManager csc(cp);
    std::vector<gpu::PointCloud> gpeaks;
    std::vector<cpu::Cloud> cpeaks;
    for (int rep = 0; rep < 12; rep++) {          // the same 4 places visited three times, 30 m of odometry apart: revisits become candidates
      const int s = rep % n_scans;
      const uint8_t* img = scans.data() + s * scan_bytes;
      gpu::StructuredKStrongest filt(ctx, img, n_az, n_range, (size_t)n_range, z_min, k, min_distance, range_res);
      gpu::PointCloud peaks;
      filt.getPeaksFilteredPointCloud(peaks, true);
      cpu::KStrongestOutput ref;
      cpu::StructuredKStrongest(img, n_az, n_range, (size_t)n_range, (float)z_min, k, (float)min_distance, (float)range_res, ref, true);
      const double ox = 30.0 * rep, oy = 0.5 * rep, oyaw = 0.01 * rep;
      gsc.makeAndSaveScancontextAndKeysRadarCloud(peaks, gpu::Pose2{ox, oy, oyaw});
      csc.makeAndSaveScancontextAndKeysRadarCloud(ref.cloud_peaks, cpu::vectorToAffine(ox, oy, oyaw));
      const std::vector<gpu::candidate> gc = gsc.detectLoopClosureID();
      const std::vector<cpu::SCCandidate> cc = csc.detectLoopClosureID();
      EXPECT(gc.size() == cc.size(), "keyframe %d: %zu candidates vs %zu", rep, gc.size(), cc.size());
      for (size_t i = 0; i < gc.size(); i++)
        EXPECT(gc[i].nn_idx == cc[i].nn_idx && gc[i].argmin_shift == cc[i].argmin_shift && gc[i].aug_idx == cc[i].aug_idx &&
                   std::fabs(gc[i].min_dist - cc[i].min_dist) < 1e-9 && gc[i].yaw_diff_rad == cc[i].yaw_diff_rad,
               "keyframe %d candidate %zu: (%d, %d, %d, %.12g) vs (%d, %d, %d, %.12g)", rep, i, gc[i].nn_idx, gc[i].argmin_shift, gc[i].aug_idx, gc[i].min_dist,
               cc[i].nn_idx, cc[i].argmin_shift, cc[i].aug_idx, cc[i].min_dist);
      if (rep >= 8) EXPECT(!gc.empty(), "keyframe %d: a revisit produced no candidate", rep);
      if (rep < n_scans) { gpeaks.push_back(peaks); cpeaks.push_back(ref.cloud_peaks); }
    }
    std::vector<int> src, refi;
    std::vector<gpu::Pose2> Ts, Tr;
    for (int s = 1; s < n_scans; s++) { src.push_back(s); refi.push_back(s - 1); Ts.push_back(gpu::Pose2{1.77 * s, 1.77 * s, 0.785}); Tr.push_back(gpu::Pose2{1.77 * (s - 1), 1.77 * (s - 1), 0.785}); }
    const std::vector<tbv_coral_result> q = gpu::CorAlRadarQuality(ctx, gpeaks, src, refi, Ts, Tr);
    for (size_t p = 0; p < src.size(); p++) {
      const cpu::CoralResult r = cpu::CorAlRadarQuality(cpeaks[src[p]], cpeaks[refi[p]], cpu::vectorToAffine(Ts[p].x, Ts[p].y, Ts[p].yaw),
                                                        cpu::vectorToAffine(Tr[p].x, Tr[p].y, Tr[p].yaw), cpu::CoralParams());
      EXPECT(q[p].count_valid == r.count_valid && q[p].merged_size == r.merged_size && q[p].valid == r.valid, "CorAl pair %zu counts", p);
      EXPECT(std::fabs(q[p].joint - r.joint) < 1e-8 && std::fabs(q[p].sep - r.sep) < 1e-8 && q[p].overlap == r.overlap, "CorAl pair %zu: %.12g %.12g vs %.12g %.12g", p, q[p].joint,
             q[p].sep, r.joint, r.sep);
    }
  }
  // ---- multi-GPU exports from a C++ host: the candidate loop of ScanContextClosure::SearchAndAddConstraint (loopclosure.cpp:658-724), sharded ------
  // One rank here (a one-rank NCCL communicator is a real communicator): tbv_comm_unique_id -> tbv_comm_init_rank -> tbv_loopdb_register_sharded must
  // return exactly the records of the single-GPU tbv_loopdb_register, and tbv_allgather_constraints must return a device-resident share unchanged.
  {
    int cell_cap = 1;
    std::vector<const tbv_cell*> sets;
    std::vector<int> n_cells;
    for (auto& m : g_maps) { sets.push_back(m->cells.data()); n_cells.push_back((int)m->cells.size()); cell_cap = std::max(cell_cap, (int)m->cells.size()); }
    tbv_loopdb* db = tbv_loopdb_create(ctx.get(), n_scans, cell_cap);
    EXPECT(db != nullptr, "tbv_loopdb_create: %s", tbv_last_error());
    int first = -1;
    EXPECT(tbv_loopdb_add(db, n_scans, sets.data(), n_cells.data(), &first) == TBV_OK && first == 0, "tbv_loopdb_add: %s", tbv_last_error());
    std::vector<int> from, to;
    std::vector<double> Tf, Tt;
    for (int a = 0; a < n_scans; a++)
      for (int b = 0; b < n_scans; b++) {
        if (a == b) continue;
        from.push_back(a); to.push_back(b);
        Tf.insert(Tf.end(), {1.77 * a + 0.4, 1.77 * a - 0.3, 0.785 + 0.02});     // a perturbed guess for `from`
        Tt.insert(Tt.end(), {1.77 * b, 1.77 * b, 0.785});
      }
    const int n_cand = (int)from.size();
    // n_scan_normal_reg(P2L) with ctor defaults (Huber 0.1, uniform weights) + SetParameters(4, 10), loopclosure.cpp:56-57
    tbv_reg_params lp{gpu::P2L, gpu::Huber, gpu::Uniform, 0.1, 1.0, 0.0, 4, 10};
    std::vector<tbv_constraint> ref(n_cand), got(n_cand), low(n_cand);
    int n_ref = 0, n_got = 0, n_low = 0, world = 0, rank = -1;
    EXPECT(tbv_loopdb_register(db, n_cand, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, nullptr, &lp, 0.0, ref.data(), n_cand, &n_ref, nullptr) == TBV_OK,
           "tbv_loopdb_register: %s", tbv_last_error());
    EXPECT(n_ref >= n_cand / 2, "only %d of %d candidates accepted", n_ref, n_cand);
    unsigned char id[TBV_COMM_ID_BYTES];
    EXPECT(tbv_comm_unique_id(id) == TBV_OK, "tbv_comm_unique_id: %s", tbv_last_error());
    EXPECT(tbv_comm_init_rank(ctx.get(), id, 1, 0) == TBV_OK, "tbv_comm_init_rank: %s", tbv_last_error());
    EXPECT(tbv_comm_world(ctx.get(), &world, &rank) == TBV_OK && world == 1 && rank == 0, "tbv_comm_world");
    float ms[4] = {0, 0, 0, 0};
    EXPECT(tbv_loopdb_register_sharded(db, n_cand, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, &lp, 0.0, got.data(), n_cand, &n_got, ms) == TBV_OK,
           "tbv_loopdb_register_sharded: %s", tbv_last_error());
    EXPECT(n_got == n_ref && std::memcmp(got.data(), ref.data(), (size_t)n_ref * sizeof(tbv_constraint)) == 0, "sharded registration differs from the single-GPU call (%d vs %d)", n_got, n_ref);
    EXPECT(ms[3] > 0.f && ms[0] > 0.f, "phase timing missing");
    // the pipelined form: two batches in flight, a third refused, collected in submission order
    {
      std::vector<tbv_constraint> p0(n_cand), p1(n_cand);
      int n0 = 0, n1 = 0;
      EXPECT(tbv_loopdb_submit_sharded(db, n_cand, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, &lp, 0.0) == TBV_OK, "submit 0: %s", tbv_last_error());
      EXPECT(tbv_loopdb_submit_sharded(db, n_cand / 2, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, &lp, 0.0) == TBV_OK, "submit 1: %s", tbv_last_error());
      EXPECT(tbv_loopdb_submit_sharded(db, n_cand, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, &lp, 0.0) == TBV_ERR_INVALID, "a third batch in flight was accepted");
      EXPECT(tbv_loopdb_collect_sharded(db, p0.data(), n_cand, &n0, nullptr) == TBV_OK, "collect 0: %s", tbv_last_error());
      EXPECT(tbv_loopdb_collect_sharded(db, p1.data(), n_cand, &n1, ms) == TBV_OK, "collect 1: %s", tbv_last_error());
      EXPECT(n0 == n_ref && std::memcmp(p0.data(), ref.data(), (size_t)n_ref * sizeof(tbv_constraint)) == 0, "pipelined batch 0 differs (%d vs %d)", n0, n_ref);
      int n_half = 0;
      for (int i = 0; i < n_ref; i++) n_half += ref[i].candidate < n_cand / 2;
      EXPECT(n1 == n_half && std::memcmp(p1.data(), ref.data(), (size_t)n_half * sizeof(tbv_constraint)) == 0, "pipelined batch 1 differs (%d vs %d)", n1, n_half);
      EXPECT(tbv_loopdb_collect_sharded(db, p1.data(), n_cand, &n1, nullptr) == TBV_ERR_INVALID, "collect without a batch in flight");
    }
    // the low-level export: records left on the device by tbv_loopdb_register_dev, then the exchange
    tbv_constraint* d_rec = nullptr;
    int* d_n = nullptr;
    EXPECT(cudaMalloc((void**)&d_rec, (size_t)n_cand * sizeof(tbv_constraint)) == 0 && cudaMalloc((void**)&d_n, sizeof(int)) == 0, "cudaMalloc");
    EXPECT(tbv_loopdb_register_dev(db, n_cand, from.data(), to.data(), Tf.data(), Tt.data(), nullptr, nullptr, &lp, 0.0, d_rec, n_cand, d_n) == TBV_OK,
           "tbv_loopdb_register_dev: %s", tbv_last_error());
    EXPECT(tbv_allgather_constraints(ctx.get(), d_rec, d_n, n_cand, low.data(), n_cand, &n_low) == TBV_OK, "tbv_allgather_constraints: %s", tbv_last_error());
    EXPECT(n_low == n_ref && std::memcmp(low.data(), ref.data(), (size_t)n_ref * sizeof(tbv_constraint)) == 0, "tbv_allgather_constraints changed the records");
    cudaFree(d_rec); cudaFree(d_n);
    EXPECT(tbv_comm_destroy(ctx.get()) == TBV_OK, "tbv_comm_destroy");
    tbv_loopdb_destroy(db);
  }
  std::printf("PASS %d checks (%d scans %dx%d)\n", g_checks, n_scans, n_az, n_range);
  return 0;
}

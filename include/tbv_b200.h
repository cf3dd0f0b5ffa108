/* tbv_b200.h — C-ABI of the B200-native TBV / CFEAR hot path (libtbv_b200.so).
 *
 * The reference (dan11003/tbv_slam_public @ 90f17c59) has no FFI layer: its seams for this path are C++ classes
 * called in-process.  Every entry point below names the reference interface it replaces (paths relative to the
 * reference root); INTEGRATION.md shows the few-line adapter a maintainer would add at each call site.
 *
 * Conventions
 *   - plain C types only; caller-owned HOST buffers unless the name ends in _dev (then: device pointers);
 *   - int status: 0 = TBV_OK, <0 = error (never exit(), never throws); tbv_last_error() gives the message;
 *   - no global state: everything hangs off a tbv_ctx (one CUDA device, one stream); distinct contexts may be
 *     used from distinct threads (the reference's odometry thread + loop-closure thread);
 *   - poses are planar (x, y, theta[rad]) — the reference's Eigen::Affine3d are planar by construction
 *     (vectorToAffine3d, cfear_radarodometry/src/cfear_radarodometry/registration.cpp:129-135);
 *   - a "cell" (reference class `cell`, cfear_radarodometry/include/cfear_radarodometry/pointnormal.h:45-105)
 *     crosses the boundary as 16 doubles, see tbv_cell.
 */
#ifndef TBV_B200_H_
#define TBV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBV_OK 0
#define TBV_ERR_INVALID (-1)   /* bad argument */
#define TBV_ERR_CUDA (-2)      /* CUDA runtime error, see tbv_last_error */
#define TBV_ERR_CAPACITY (-3)  /* an internal or caller buffer is too small */
#define TBV_ERR_NO_GPU (-4)    /* no CUDA device: the product has no CPU fallback */

typedef struct tbv_ctx tbv_ctx;

/* ---- enums: values match the reference's enums ------------------------------------------------------------ */
/* cost_metric  (cfear_radarodometry/include/cfear_radarodometry/registration.h:55) */
enum { TBV_P2P = 0, TBV_P2L = 1, TBV_P2D = 2 };
/* loss_type    (registration.h:60) */
enum { TBV_LOSS_NONE = 0, TBV_LOSS_HUBER = 1, TBV_LOSS_CAUCHY = 2, TBV_LOSS_SOFTLONE = 3, TBV_LOSS_COMBINED = 4, TBV_LOSS_TUKEY = 5 };
/* weightoption (registration.h:50) */
enum { TBV_W_UNIFORM = 0, TBV_W_SIM_N = 1, TBV_W_SIM_DIRECTION = 2, TBV_W_SIM_SCALE = 3, TBV_W_COMBINED = 4 };

/* ---- records ------------------------------------------------------------------------------------------------ */
/* reference `cell` (pointnormal.h:66-73) */
typedef struct tbv_cell {
  double u[2];            /* u_ */
  double cov[4];          /* cov_ row-major: c00 c01 c10 c11 */
  double scale;           /* scale_ (planarity) */
  double snormal[2];      /* snormal_ */
  double orth_normal[2];  /* orth_normal */
  double lambda_min, lambda_max;
  double sum_intensity, avg_intensity;
  double n_samples;       /* Nsamples_ */
} tbv_cell;

/* k-strongest / CA-CFAR filter parameters — radarDriver::Parameters (radar_driver.h:40-64); z_min, range_res,
 * min_distance are float there and widened exactly as StructuredKStrongest's ctor does (radar_filters.h:86). */
typedef struct tbv_filter_params {
  float z_min;
  int k_strongest;
  float min_distance;
  float range_res;
} tbv_filter_params;

typedef struct tbv_cfar_params {   /* AzimuthCACFAR ctor, cfar.h:28-42; max_distance = 400 at radar_driver.cpp:54 */
  int window_size;
  double false_alarm_rate;
  int nb_guard_cells;
  double range_resolution;
  double static_threshold;
  double min_distance;
  double max_distance;
} tbv_cfar_params;

/* A filtered cloud in struct-of-arrays form; each array holds `capacity` entries per scan, scan b starts at
 * b*capacity.  Order = the reference's: azimuth-major, ascending (intensity, range) inside an azimuth.
 * Any array pointer may be NULL (not wanted). */
typedef struct tbv_points {
  int capacity;        /* entries per scan (>= n_az * k for the k-strongest filter) */
  int* count;          /* [batch] out: points per scan */
  uint16_t* azimuth;
  uint16_t* range;
  uint8_t* intensity;
  float* x;
  float* y;
} tbv_points;

/* n_scan_normal_reg configuration (n_scan_normal.h:33-35,53-55,72-75; registration.h:117-122) */
typedef struct tbv_reg_params {
  int cost;                 /* TBV_P2L ... */
  int loss;                 /* TBV_LOSS_HUBER ... */
  int weight_opt;           /* TBV_W_... */
  double loss_limit;        /* 0.1 */
  double cov_scale;         /* SetD2dPar */
  double regularization;    /* SetD2dPar */
  int max_itr_association;  /* SetParameters; default 8 */
  int max_itr_solver;       /* SetParameters; default 20 */
} tbv_reg_params;

typedef struct tbv_reg_summary {
  int success;              /* Register's bool */
  int itrs;                 /* timing key "itrs": association iteration counter at exit */
  int lm_iterations;        /* total LM iterations over all association rounds */
  int num_residuals;        /* summary_.num_residuals of the last solve */
  int last_n_iterations;    /* summary_.iterations.size() of the last solve */
  int termination;          /* 0 convergence, 1 no_convergence, 2 failure (last solve) */
  double score;             /* final_cost / num_residuals */
  double final_cost;
  double last_relative_decrease;
} tbv_reg_summary;

/* OdometryKeyframeFuser::Parameters subset on the path (odometrykeyframefuser.h:90-113) + filter + preset */
typedef struct tbv_odom_params {
  tbv_filter_params filter;
  tbv_reg_params reg;       /* max_itr_* <= 0 -> class defaults (8, 20) */
  int submap_scan_size;
  int weight_intensity;
  int use_guess, compensate, radar_ccw, use_keyframe;
  double res;               /* cell radius / voxel leaf */
  double min_keyframe_dist, min_keyframe_rot_deg;
  double downsample_factor;
  int cell_capacity;        /* cells kept per scan; <= 0 -> 1024 */
  int sample_capacity;      /* voxel-grid sample points examined per scan; <= 0 -> 4096 */
} tbv_odom_params;

typedef struct tbv_odom_out {   /* one per sequence per step */
  double pose[3];
  int n_points, n_cells, itrs, reg_ok, is_keyframe, n_keyframes;
  int lm_iterations, num_residuals;
  int status;               /* TBV_OK, or TBV_ERR_CAPACITY when a per-scan capacity was exceeded (results then unreliable) */
  int n_samples;            /* voxel-grid sample points examined for this scan */
  double score;
} tbv_odom_out;

/* ---- context -------------------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`.  Returns NULL (and sets tbv_last_error) when no GPU is present. */
tbv_ctx* tbv_create(int device);
void tbv_destroy(tbv_ctx* ctx);
const char* tbv_last_error(void);
/* the cudaStream_t all work of this context is enqueued on (for event timing by the caller) */
void* tbv_stream(tbv_ctx* ctx);
int tbv_synchronize(tbv_ctx* ctx);
/* number of kernel launches issued by this library since the context was created */
long long tbv_launch_count(tbv_ctx* ctx);
int tbv_version(void);
/* Per-launch device timing (the reference's CFEAR_Radarodometry::timing, statistics.h:38, keyed by kernel): after
 * tbv_profile_begin every kernel launch of this context is followed by an event record; tbv_profile_end synchronises and
 * returns, in launch order, the kernel names (static strings) and the milliseconds between consecutive events. */
int tbv_profile_begin(tbv_ctx* ctx);
int tbv_profile_end(tbv_ctx* ctx, int capacity, const char** names, float* ms, int* n);

/* ---- K1: k-strongest filter -------------------------------------------------------------------------------------
 * Replaces StructuredKStrongest::StructuredKStrongest + FilterKstrongest + AxialNonMaxSupress +
 * getPeaksFilteredPointCloud (cfear_radarodometry/src/cfear_radarodometry/radar_filters.cpp:198-337), as called by
 * radarDriver::Process (radar_driver.cpp:57-61).  polar: batch scans of n_az rows x n_range u8, row_stride bytes
 * between rows, scan b at polar + b*n_az*row_stride.  out_peaks may be NULL. */
int tbv_filter_kstrongest(tbv_ctx* ctx, const uint8_t* polar, int n_az, int n_range, size_t row_stride, int batch,
                          const tbv_filter_params* params, tbv_points* out_filtered, tbv_points* out_peaks);
/* Same with the scans already resident in device memory; results stay on the device (tbv_cloud_* accessors). */
int tbv_filter_kstrongest_dev(tbv_ctx* ctx, const uint8_t* polar_dev, int n_az, int n_range, size_t row_stride, int batch,
                              const tbv_filter_params* params, int want_peaks);
/* copies the device-resident result of the last *_dev filter call to host arrays */
int tbv_filter_fetch(tbv_ctx* ctx, tbv_points* out_filtered, tbv_points* out_peaks);

/* MulRan-style input (radarDriver::Callback, radar_driver.cpp:74-90): MONO8 range-major image [n_range][n_az]
 * rotated 90 deg CCW into the azimuth-major layout the filters expect.  Host in, host out. */
int tbv_rotate90ccw(tbv_ctx* ctx, const uint8_t* src, int rows, int cols, uint8_t* dst);

/* ---- K1b: CA-CFAR filter — AzimuthCACFAR::getFilteredPointCloud (cfar.cpp:35-83) -------------------------------- */
int tbv_filter_cacfar(tbv_ctx* ctx, const uint8_t* polar, int n_az, int n_range, size_t row_stride, int batch,
                      const tbv_cfar_params* params, tbv_points* out);

/* ---- K2: motion compensation — CFEAR_Radarodometry::Compensate (utils.cpp:96-113) ------------------------------- */
int tbv_compensate(tbv_ctx* ctx, float* x, float* y, int n, const double mot_xyt[3], int ccw);

/* ---- K3: oriented surface points — MapPointNormal ctor / ComputeNormals / cell (pointnormal.cpp:7-90, 265-297) --- */
/* returns TBV_OK and *n_cells (may exceed cell_capacity: then only cell_capacity records are written and the call
 * returns TBV_ERR_CAPACITY); *n_samples (optional) = number of voxel-grid sample points examined */
int tbv_build_cells(tbv_ctx* ctx, const float* x, const float* y, const float* intensity, int n, float radius,
                    double downsample_factor, int weight_intensity, const double origin[2], tbv_cell* cells, int cell_capacity,
                    int* n_cells, int* n_samples);

/* ---- K4: one (target, source) pair: correspondences + robustified cost, J^T J, J^T r ----------------------------
 * n_scan_normal_reg::AddScanPairCost (n_scan_normal.cpp:213-324) followed by one evaluation of the resulting
 * residual blocks at x = T_src (what ceres::Problem::Evaluate / the first LM iteration computes).
 * itr selects the search radius exactly like the reference's member itr_ (1 -> 2*radius_, else radius_ = 2.0).
 * assoc (optional, n_src ints): target index matched to each source cell or -1. */
int tbv_pair_normal_eq(tbv_ctx* ctx, const tbv_cell* tgt, int n_tgt, const double T_tgt[3], const tbv_cell* src, int n_src,
                       const double T_src[3], const tbv_reg_params* params, int itr, double* cost, int* n_res, double H[9],
                       double g[3], int32_t* assoc);

/* ---- K5: registration — n_scan_normal_reg::Register (n_scan_normal.cpp:82-185) -----------------------------------
 * scans[n_scans-1] is the moving scan (local frame), the others are fixed; T is [n_scans][3] in/out. */
int tbv_register(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, double* T,
                 const tbv_reg_params* params, tbv_reg_summary* summary);
/* GetCost (n_scan_normal.cpp:186-211): residuals (optional) receives up to res_capacity corrected residuals */
int tbv_get_cost(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, const double* T,
                 const tbv_reg_params* params, int itr, double* score, double* cost, int* n_res, double* residuals,
                 int res_capacity);
/* OdometryKeyframeFuser::approximateCovarianceBySampling (cfear_radarodometry/src/cfear_radarodometry/odometrykeyframefuser.cpp:261-380;
 * the same block in tbv_slam/src/tbv_slam/loopclosure.cpp:127-200).  Sampling half: the n_per_axis^3 GetCost evaluations around the last
 * scan's pose T[n_scans-1] (grid linspace(-xy_range/2, xy_range/2) x same x linspace(-yaw_range/2, yaw_range/2); theta-major, then x,
 * then y) run as independent problems in ONE launch.  samples: [n_per_axis^3][4] = dx, dy, dyaw, cost (problem_->Evaluate's total cost).
 * itr: the registration object's itr_ at the time of the call (after Register: > 1 -> search radius radius_). */
int tbv_cost_samples(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, const double* T,
                     const tbv_reg_params* params, int itr, double xy_range, double yaw_range, int n_per_axis, double* samples);
/* Fitting half (host arithmetic): least-squares quadric through the samples, Hessian, convexity test, covariance
 * 2 H^-1 * score_scale * cov_scaler in the reference's 6x6 layout (x, y, ., ., ., yaw; row-major).  score_scale = GetCovarianceScaler()
 * = final_cost / (num_residuals - 3) of the preceding Register (n_scan_normal.cpp:433-439).  *convex = 0: the sampling is not used. */
int tbv_cov_from_cost_samples(const double* samples, int n_samples, double score_scale, double cov_scaler, double cov6x6[36], int* convex);
/* Batched independent two-scan registrations — loopclosure::Register (tbv_slam/src/tbv_slam/loopclosure.cpp:35-97):
 * pair p registers `from` (moving, pose T_from[p]) against `to` (fixed, pose T_to[p]) with P2L / Huber 0.1 / uniform
 * weights / SetParameters(4,10) unless params says otherwise.  Cell sets are given once (n_sets), pairs index them.
 * Outputs per pair: T_revised (x,y,theta), T_align = T_revised^-1 * T_to, summary. */
int tbv_register_batch(tbv_ctx* ctx, int n_sets, const tbv_cell* const* sets, const int* n_cells, int n_pairs, const int* from_set,
                       const int* to_set, const double* T_from, const double* T_to, const tbv_reg_params* params,
                       double* T_revised, double* T_align, tbv_reg_summary* summaries);

/* ---- batched odometry: radarDriver::Process + OdometryKeyframeFuser::processFrame --------------------------------
 * (radar_driver.cpp:48-73, odometrykeyframefuser.cpp:143-259) for n_seq independent sequences advanced in lock-step:
 * one call consumes one polar scan per sequence and returns one pose per sequence.  All per-sequence state
 * (keyframe window, T_prev, Tmot) lives on the device. */
typedef struct tbv_odom tbv_odom;
tbv_odom* tbv_odom_create(tbv_ctx* ctx, int n_seq, int n_az, int n_range, const tbv_odom_params* params);
void tbv_odom_destroy(tbv_odom* od);
int tbv_odom_reset(tbv_odom* od);
/* host scans [n_seq][n_az][n_range] (pinned or pageable) -> out[n_seq]; includes H2D and D2H */
int tbv_odom_step(tbv_odom* od, const uint8_t* polar_host, tbv_odom_out* out);
/* range_major != 0: from now on every scan handed to tbv_odom_step / _step_dev / _submit is in the sensor's wire layout
 * [n_range][n_az] (MulRan and every non-Oxford dataset: radar_driver.cpp:74-90) and is rotated 90 deg CCW on the device as the
 * first launch of the step — cv::rotate(ROTATE_90_COUNTERCLOCKWISE) on receipt.  Default 0: azimuth-major [n_az][n_range] (Oxford). */
int tbv_odom_set_wire_layout(tbv_odom* od, int range_major);
/* enable != 0: overlapped steps for large batches.  The filter of a step (rotate on receipt + k-strongest + NMS + polar -> Cartesian)
 * depends on nothing but the step's scans, so it is launched on a second, low-priority stream into a second set of clouds and runs while
 * the registration of the PREVIOUS step drains; the motion compensation (which needs the previous step's pose) becomes its own small
 * launch on the context's stream — same arithmetic on the same floats, identical results.  Steps are not replayed from CUDA graphs in
 * this mode.  Contract for tbv_odom_step_dev: the scans handed over must be COMPLETE in device memory when the call is made (not merely
 * enqueued on the context's stream); tbv_odom_submit orders its own uploads.  Default 0. */
int tbv_odom_set_overlap(tbv_odom* od, int enable);
/* The step (7 kernel launches) is replayed from a CUDA graph once an input buffer has been seen twice — the upload buffers of
 * tbv_odom_step / _submit, or a caller's own ring of device buffers passed to tbv_odom_step_dev.  enable = 0 turns that off (direct
 * launches); default on.  Results are identical either way. */
int tbv_odom_set_graphs(tbv_odom* od, int enable);
/* scans already on the device; results stay on the device until tbv_odom_fetch */
int tbv_odom_step_dev(tbv_odom* od, const uint8_t* polar_dev);
int tbv_odom_fetch(tbv_odom* od, tbv_odom_out* out);
/* Software-pipelined host path: enqueue the upload of the next step's scans on a copy stream while the previous
 * step computes. tbv_odom_submit returns immediately; tbv_odom_collect blocks for the oldest submitted step. */
int tbv_odom_submit(tbv_odom* od, const uint8_t* polar_host_pinned);
int tbv_odom_collect(tbv_odom* od, tbv_odom_out* out);
/* current cells of sequence `seq` (the scan processed last) and of its keyframe window, for parity checks */
int tbv_odom_cells(tbv_odom* od, int seq, int keyframe /* -1 = current scan */, tbv_cell* cells, int capacity, int* n_cells,
                   double pose[3]);

/* ---- K6/K7: radar Scan Context ---------------------------------------------------------------------------------------
 * RSCManager::Parameters + SCManager constants (place_recognition_radar/include/place_recognition_radar/RadarScancontext.h:38-85,
 * Scancontext.h:100-118); TBV-8 offline values: 40 rings x 120 sectors, 80 m, sum / 1000, 10 from the ring-key search. */
typedef struct tbv_sc_params {
  int num_ring, num_sector;
  double max_radius;
  double search_ratio;              /* 0.1 */
  int num_candidates_from_tree;     /* 10 */
  int n_candidates;
  double odom_sigma_error;          /* 0.05 */
  int odometry_coupled_closure, augment_sc;
  double no_point;
  int desc_function;                /* 0 = "sum", 1 = "max" */
  double desc_divider;
  double distance_exclude_recent;   /* 10 m */
} tbv_sc_params;

/* RSCManager::MakeRadarCloudContext (RadarScancontext.cpp:59-131) for the cloud translated by each of n_offsets lateral
 * offsets (the augmentations of makeAndSaveScancontextAndKeysRadarCloud, :162-179; offset (0,0) = the cloud itself), plus
 * makeRingkeyFromScancontext / makeSectorkeyFromScancontext (Scancontext.cpp:239-268).
 * desc: [n_offsets][num_sector][num_ring] (column-major rings x sectors, Eigen's layout); ringkey: [n_offsets][num_ring] float;
 * sectorkey (optional): [n_offsets][num_sector]. */
int tbv_sc_make(tbv_ctx* ctx, const float* x, const float* y, const float* intensity, int n, const tbv_sc_params* params, int n_offsets,
                const double* offsets_xy, double* desc, float* ringkey, double* sectorkey);
/* SCManager::distanceBtnScanContext (Scancontext.cpp:157-189) for n_pairs (query, candidate) descriptor pairs. */
int tbv_sc_distance_batch(tbv_ctx* ctx, const double* desc_q, int n_q, const double* desc_c, int n_c, int n_pairs, const int* q_idx,
                          const int* c_idx, const tbv_sc_params* params, double* dist, int* shift);
/* RSCManager::ExcludeAndUpdateLikelihood + OdometryNNSearch (RadarScancontext.cpp:183-221, 259-284) for n_q queries; query q
 * belongs to keyframe q_current[q] and sees database entries [0, q_current[q] - 1 - n_exclude).  db_keys: [n_db][num_ring],
 * odom_xyt: [n_db][3].  Outputs: cand_idx / cand_odom_sim [n_q][num_candidates_from_tree] (-1 / 0 padded), n_exclude [n_q]. */
int tbv_sc_search(tbv_ctx* ctx, const float* db_keys, const double* odom_xyt, int n_db, int n_q, const float* q_keys, const int* q_current,
                  const tbv_sc_params* params, int* cand_idx, double* cand_odom_sim, int* n_exclude);

/* CFEARQuality (coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:330-352, called from
 * ScanLearningInterface::getCFEARQualityMeasure, alignmentinterface.cpp:457-475): GetCost of {ref (fixed), src} at the given poses with the
 * caller's params (reference: P2L, Huber 0.3, uniform weights; itr_ = 0) for a batch of pairs in one launch.
 * quality: [n_pairs][3] = score, number of residuals, (|src| + |ref|) / 2  ({0, 0, 0} when GetCost fails). */
int tbv_cfear_quality_batch(tbv_ctx* ctx, int n_sets, const tbv_cell* const* sets, const int* n_cells, int n_pairs, const int* src_set,
                            const int* ref_set, const double* T_src, const double* T_offset, const double* T_ref,
                            const tbv_reg_params* params, double* quality);

/* ---- CorAl alignment quality (next-row f-1): CorAlRadarQuality ------------------------------------------------------------
 * coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:8-229 as called by ScanLearningInterface::getCorAlQualityMeasure
 * (coral_alignment_quality/src/alignment_checker/alignmentinterface.cpp:437-454: peaks clouds, radius 1.0, entropy setting `any`,
 * weight_res_intensity false).  Batched over pairs, one CTA per pair: clouds are given once (local frames) and indexed by the pairs;
 * pair p compares cloud src_cloud[p] at pose T_src[p] * T_offset[p] with cloud ref_cloud[p] at pose T_ref[p] (poses: x, y, theta;
 * T_offset may be NULL = identity).  results[p] = quality_ {joint, sep, overlap} + counts; valid = (overlap >= 0.1).
 * per_point (optional): sum over pairs of merged_size rows of 3 doubles = sep, joint, valid of every point (src points first), the
 * reference's sep_res_ / joint_res_ / sep_valid before the weighting loop.  Clouds are limited to 4096 points (TBV_ERR_CAPACITY). */
typedef struct tbv_coral_params {
  double radius;             /* AlignmentQuality::parameters::radius (TBV: 1.0) */
  int weight_res_intensity;  /* weights = point intensity instead of 1 */
  int overlap_req;           /* AlignmentQuality.h:244 overlap_req_ = 1 */
} tbv_coral_params;
typedef struct tbv_coral_result {
  double joint, sep, overlap;
  int count_valid, merged_size, valid;
} tbv_coral_result;
int tbv_coral_quality_batch(tbv_ctx* ctx, int n_clouds, const float* const* x, const float* const* y, const float* const* intensity,
                            const int* n_points, int n_pairs, const int* src_cloud, const int* ref_cloud, const double* T_src,
                            const double* T_offset, const double* T_ref, const tbv_coral_params* params, tbv_coral_result* results,
                            double* per_point);

/* ---- K8: pose-graph normal equations ------------------------------------------------------------------------------------
 * CeresLeastSquares::BuildOptimizationProblem / AddConstraintType + PoseGraph3dErrorTerm (tbv_slam/src/tbv_slam/ceresoptimizer.cpp:
 * 28-108, tbv_slam/include/tbv_slam/ceresoptimizer.h:51-95) followed by one evaluation: robustified residuals, tangent-space
 * Jacobians (EigenQuaternionParameterization), and the block normal equations.  nodes: 7 doubles (p xyz, q xyzw); constraint c:
 * ids[3c..] = (id_begin, id_end, type: 0 odometry / 1 loop), meas[7c..] = t_be (p, q), info (optional) [36c..] row-major.
 * Outputs: H_diag [n_nodes][36], H_off [n_con][36] (block (begin,end) = Ja^T Jb), g [n_nodes][6], residuals (optional) [n_con][6]. */
typedef struct tbv_pgo_params {
  double odom_vxx, odom_vyy, odom_vtt, loop_scaling;   /* 0.01, 0.01, 0.001, 500000 (ceresoptimizer.cpp:18-27) */
  int replace_cov_by_identity;
  double loop_cauchy;                                  /* CauchyLoss(0.1) on loop constraints */
} tbv_pgo_params;
int tbv_pgo_assemble(tbv_ctx* ctx, int n_nodes, const double* nodes, int n_con, const int* ids, const double* meas, const double* info,
                     const tbv_pgo_params* params, int fixed_node, double* cost, double* H_diag, double* H_off, double* g, double* residuals);

/* ---- K8b: one trust-region step of the pose graph on the device (next-row f-2) -------------------------------------------
 * The linear solve inside ceres::Solve as CeresLeastSquares::SolveOptimizationProblem configures it (tbv_slam/src/tbv_slam/
 * ceresoptimizer.cpp:50-62: LEVENBERG_MARQUARDT + SPARSE_NORMAL_CHOLESKY, Ceres 2.1.0 is a system dependency, not vendored):
 * (H + D) delta = -g with D = clamp(diag(H), 1e-6, 1e32) / radius (min_lm_diagonal / max_lm_diagonal), H, g exactly the
 * tbv_pgo_assemble outputs (H_diag, H_off, g; the fixed node's rows and columns are held at zero).  Solved by conjugate
 * gradients preconditioned with the odometry chain (the block-tridiagonal part of H + D, factorised and applied by block cyclic
 * reduction) on one thread-block cluster of 8 CTAs, deterministic reductions, instead of a sparse Cholesky: ~5 CG iterations on a
 * 4 500-node graph; delta agrees with the direct solve to rel_tol * |g| in the residual.  delta: [n_nodes][6] tangent step
 * (p xyz | rotation), applied as p += dp, q = exp(dr) * q (EigenQuaternionParameterization::Plus).
 * iters / rel_residual (optional): CG iterations used and |b - A delta| / |b| reached.  Stops at max_iters or rel_tol. */
int tbv_pgo_solve_step(tbv_ctx* ctx, int n_nodes, int n_con, const int* ids, const double* H_diag, const double* H_off, const double* g,
                       int fixed_node, double radius, int max_iters, double rel_tol, double* delta, int* iters, double* rel_residual);

/* Same solve with an explicit LM damping vector: (H + diag(damping)) delta = -g, damping [n_nodes][6] > 0 (NULL: derived from
 * `radius` as above).  This is what a faithful LevenbergMarquardtStrategy needs: its diagonal is clamp(diag(S H S), 1e-6, 1e32) with the
 * Jacobi scaling S fixed at iteration 0 and is reused after a rejected step. */
int tbv_pgo_solve_damped(tbv_ctx* ctx, int n_nodes, int n_con, const int* ids, const double* H_diag, const double* H_off, const double* g,
                         const double* damping, int fixed_node, double radius, int max_iters, double rel_tol, double* delta, int* iters,
                         double* rel_residual);

/* CeresLeastSquares::Solve (tbv_slam/src/tbv_slam/ceresoptimizer.cpp:13-62, called from PoseGraph::ForceOptimize, posegraph.cpp:118-128)
 * with the whole Levenberg-Marquardt loop on the device: nodes [n_nodes][7] (p xyz | q xyzw) are optimised IN PLACE, as the reference
 * optimises RadarScan::T.p / T.q in place.  Every iteration = tbv_pgo_assemble's kernels at the candidate + one chain-preconditioned
 * solve + the trust-region bookkeeping of Ceres 2.1.0's TrustRegionMinimizer / LevenbergMarquardtStrategy (Jacobi scaling fixed at
 * iteration 0, LM diagonal reuse after a rejected step, tolerances tested on the candidate, radius update
 * radius /= max(1/3, 1 - (2 rho - 1)^3)); the graph, the blocks and all vectors stay in HBM, the host reads ~60 bytes per iteration.
 * options: NULL or fields <= 0 -> the reference's values (max_num_iterations 200, function / gradient / parameter tolerance
 * 1e-6 / 1e-10 / 1e-8, initial radius 1e4; CG: 20000 iterations, 1e-10). */
typedef struct tbv_pgo_options {
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance, initial_radius;
  int max_cg_iterations;
  double cg_rel_tol;
} tbv_pgo_options;
enum { TBV_PGO_MAX_ITERATIONS = 0, TBV_PGO_GRADIENT_TOLERANCE = 1, TBV_PGO_PARAMETER_TOLERANCE = 2, TBV_PGO_FUNCTION_TOLERANCE = 3,
       TBV_PGO_MIN_RADIUS = 4, TBV_PGO_FAILURE = 5 };
typedef struct tbv_pgo_summary {
  double initial_cost, final_cost;
  int iterations, successful_steps, cg_iterations, termination;   /* termination: TBV_PGO_* */
  float device_ms;                                                /* device time of the whole optimisation (events on the context's stream) */
} tbv_pgo_summary;
int tbv_pgo_optimize(tbv_ctx* ctx, int n_nodes, double* nodes, int n_con, const int* ids, const double* meas, const double* info,
                     const tbv_pgo_params* params, int fixed_node, const tbv_pgo_options* options, tbv_pgo_summary* summary);

/* ---- loop-closure keyframe database + sharded candidate registration ---------------------------------------------------
 * The loop-closure thread registers every Scan-Context candidate (from, to) with loopclosure::RegisterLoopCandidate ->
 * loopclosure::Register (tbv_slam/src/tbv_slam/loopclosure.cpp:320-364, 35-97) against cells kept in the pose graph's
 * RadarScan nodes (cloud_normal_, cfear_radarodometry/include/cfear_radarodometry/types.h).  tbv_loopdb keeps those cell
 * sets (and their 4 m search grids) resident in HBM, so a batch of candidates is ONE kernel launch with no cell upload;
 * candidates are independent, so a batch shards over GPUs by candidate with every rank holding the whole database
 * (SURVEY 8e); accepted constraints come back as fixed-size records ready for an all-gather.
 *
 * tbv_constraint mirrors Constraint3d (cfear_radarodometry/include/cfear_radarodometry/types.h:155-172) for a planar
 * pose: t_be = Talign = Trevised^-1 * Tto as (x, y, theta); cov = reg_cov of the moving scan (diag(0.1^2, 0.1^2, 0.01^2),
 * n_scan_normal.cpp:171-175) with its translation block rotated into the revised frame (loopclosure.cpp:93):
 * (xx, xy, yy, tt); the reference stores information = cov^-1.  128 bytes. */
typedef struct tbv_constraint {
  int32_t id_begin, id_end;   /* from, to (keyframe ids in the database) */
  int32_t type;               /* ConstraintType: 0 odometry, 1 loop_appearance (types.h:153) */
  int32_t candidate;          /* index of the candidate in the caller's (global) list: all-gather order key */
  double t_be[3];
  double cov[4];
  double score;               /* n_scan_normal_reg::getScore() = final_cost / num_residuals */
  double t_revised[3];        /* Trevised (world pose of `from` after registration) */
  int32_t itrs, num_residuals;
  double quality[2];          /* caller-defined (sc_sim, odom_bounds), copied from the candidate list */
} tbv_constraint;

typedef struct tbv_loopdb tbv_loopdb;
tbv_loopdb* tbv_loopdb_create(tbv_ctx* ctx, int max_keyframes, int cell_capacity);
void tbv_loopdb_destroy(tbv_loopdb* db);
/* appends n_sets cell sets (keyframes); ids are assigned consecutively from the current size, which is returned in
 * *first_id.  One H2D per set + one grid-build launch for the whole batch. */
int tbv_loopdb_add(tbv_loopdb* db, int n_sets, const tbv_cell* const* sets, const int* n_cells, int* first_id);
int tbv_loopdb_size(tbv_loopdb* db);
/* Registers n_cand candidates in one launch.  from/to: keyframe ids; T_from/T_to: [n_cand][3] initial world poses
 * (Tfrom = pose of `from`, Tto = Tfrom * guess, loopclosure.cpp:337-338); candidate_index (optional): global index
 * stored in the record (default: the local index); quality (optional): [n_cand][2] copied through.
 * Accepted = Register returned true (and score <= max_score when max_score > 0).  Accepted constraints are written in
 * candidate order to out[0..*n_out) (host), at most out_capacity; summaries (optional) [n_cand] receives every result. */
int tbv_loopdb_register(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                        const int* candidate_index, const double* quality, const tbv_reg_params* params, double max_score,
                        tbv_constraint* out, int out_capacity, int* n_out, tbv_reg_summary* summaries);
/* Same, results left on the device for a collective: out_dev [out_capacity] records, n_out_dev one int.  Enqueued on the
 * context's stream; no host synchronisation. */
int tbv_loopdb_register_dev(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                            const int* candidate_index, const double* quality, const tbv_reg_params* params, double max_score,
                            tbv_constraint* out_dev, int out_capacity, int* n_out_dev);

/* ---- multi-GPU: candidates sharded over ranks, accepted constraints all-gathered (SURVEY 8b "multi-GPU", 8e) --------------------------
 * Replaces, for a candidate list split over GPUs, the serial order in which ScanContextClosure::SearchAndAddConstraint
 * (tbv_slam/src/tbv_slam/loopclosure.cpp:658-724) hands accepted candidates to ApplyConstratins (:261-318) /
 * PoseGraph::AddConstraintThSafe: every rank ends up with ALL accepted constraints in global candidate order.
 * One process (or host thread) per GPU; each context carries one NCCL communicator.  Either the host owns the communicator
 * (tbv_comm_init: a C++/ROS host that already uses NCCL passes its ncclComm_t) or the library creates one from an
 * ncclUniqueId that the host distributes with whatever it has (MPI, a ROS parameter, torch.distributed's store):
 * rank 0 calls tbv_comm_unique_id, every rank calls tbv_comm_init_rank.  NCCL is loaded at run time (libnccl.so.2);
 * contexts without a communicator behave as world = 1 and never touch it. */
#define TBV_COMM_ID_BYTES 128                                    /* sizeof(ncclUniqueId) */
int tbv_comm_unique_id(void* unique_id /* TBV_COMM_ID_BYTES, out */);
int tbv_comm_init_rank(tbv_ctx* ctx, const void* unique_id, int world, int rank);   /* collective; the context owns the communicator */
int tbv_comm_init(tbv_ctx* ctx, void* nccl_comm /* ncclComm_t on the context's device; stays owned by the caller */);
int tbv_comm_world(tbv_ctx* ctx, int* world, int* rank);
int tbv_comm_destroy(tbv_ctx* ctx);
/* All-gather of accepted constraints.  local_dev: [capacity] records on the device, the first *n_local_dev valid, ascending by
 * `candidate` (what tbv_loopdb_register_dev writes); capacity: the same on every rank (>= the largest share).  ONE ncclAllGather
 * of (capacity + 1) 128-byte records per rank on the context's stream — the count travels in the block — then a device merge into
 * global candidate order.  all [all_capacity] (host) receives the *n_all merged records, identical on every rank. */
int tbv_allgather_constraints(tbv_ctx* ctx, const tbv_constraint* local_dev, const int* n_local_dev, int capacity, tbv_constraint* all,
                              int all_capacity, int* n_all);
/* Same, results left on the device (valid until the next exchange on this context); no host synchronisation. */
int tbv_allgather_constraints_dev(tbv_ctx* ctx, const tbv_constraint* local_dev, const int* n_local_dev, int capacity,
                                  const tbv_constraint** all_dev, const int** n_all_dev);
/* The sharded form of tbv_loopdb_register in one call: every rank passes the SAME global candidate list; rank r registers the
 * candidates with from mod world == r in one launch, packs its accepted constraints (candidate = global index), and the
 * exchange above returns all of them, in global candidate order, to every rank.  At world = 1 (no communicator) it equals
 * tbv_loopdb_register.  timing_ms (optional, [4]): device time of {H2D of the share + registration + packing, all-gather + merge,
 * D2H of the merged records, whole call} of this call in ms, from events on the context's stream. */
int tbv_loopdb_register_sharded(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                                const double* quality, const tbv_reg_params* params, double max_score, tbv_constraint* all,
                                int all_capacity, int* n_all, float* timing_ms);

/* The same in two halves, so that a host with a stream of candidate batches (one per new keyframe, loopclosure.cpp:658-724) keeps two
 * in flight: submit enqueues ONE H2D copy of the rank's share (pinned staging), the registration launch and the packing on the
 * context's stream and the exchange (ncclAllGather + merge + D2H of the merged records into pinned memory) on a second stream of the
 * database, and returns without waiting; collect blocks until the OLDEST submitted batch is on the host and copies it out.  Batches are
 * collected in submission order; at most two may be in flight (TBV_ERR_INVALID on a third submit).  Every rank submits the same batches
 * in the same order.  The exchange of batch i overlaps the registration of batch i + 1.  tbv_loopdb_register_sharded = submit + collect. */
int tbv_loopdb_submit_sharded(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                              const double* quality, const tbv_reg_params* params, double max_score);
int tbv_loopdb_collect_sharded(tbv_loopdb* db, tbv_constraint* all, int all_capacity, int* n_all, float* timing_ms);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) */
void* tbv_host_alloc(size_t bytes);
void tbv_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TBV_B200_H_ */

// k_odom.cu — the odometry pipeline: radarDriver::Process + OdometryKeyframeFuser::processFrame for n_seq independent
// sequences advanced in lock-step, all state resident on the device.
//
// Replaces (per sequence, per frame) radarDriver::Process (cfear_radarodometry/src/cfear_radarodometry/radar_driver.cpp:48-73)
// and OdometryKeyframeFuser::processFrame / KeyFrameBasedFuse / AccelerationVelocitySanityCheck / FormatScans /
// AddToReference (cfear_radarodometry/src/cfear_radarodometry/odometrykeyframefuser.cpp:62-94, 143-259, 470-494).
//
// One step = K1/K2 (k-strongest, clouds) -> compensation with the previous frame-to-frame motion -> K3 (cells) -> problem
// setup (Tguess = Tprev * Tmot, keyframe window) -> K4/K5 (registration) -> fuser update (sanity check, keyframe policy,
// window rotation).  Nothing returns to the host between the stages; the host only uploads the scans and reads back one
// tbv_odom_out per sequence.  Frame t of a sequence needs the pose of frame t-1, so the parallelism is ACROSS sequences
// (the reference's own scaling model: one worker process per sequence).
#include <cmath>

#include "tbv_reg.cuh"

namespace tbv {

constexpr int MAX_KF = 16;  // largest supported submap_scan_size on the device path

struct AffD {
  double r00, r01, r10, r11, tx, ty;
};
__device__ __forceinline__ AffD aff_identity() { return AffD{1, 0, 0, 1, 0, 0}; }
__device__ __forceinline__ AffD aff_from_vec(double x, double y, double th) {
  const double c = cos(th), s = sin(th);
  return AffD{c, -s, s, c, x, y};
}
__device__ __forceinline__ AffD aff_mul(const AffD& A, const AffD& B) {
  AffD C;
  C.r00 = A.r00 * B.r00 + A.r01 * B.r10;
  C.r01 = A.r00 * B.r01 + A.r01 * B.r11;
  C.r10 = A.r10 * B.r00 + A.r11 * B.r10;
  C.r11 = A.r10 * B.r01 + A.r11 * B.r11;
  C.tx = (A.r00 * B.tx + A.r01 * B.ty) + A.tx;
  C.ty = (A.r10 * B.tx + A.r11 * B.ty) + A.ty;
  return C;
}
__device__ __forceinline__ AffD aff_inv(const AffD& A) {
  AffD I;
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
__device__ __forceinline__ void aff_to_vec(const AffD& T, double v[3]) {  // utils.cpp:115-122
  v[0] = T.tx; v[1] = T.ty; v[2] = atan2(T.r10, T.r11);
}

struct SeqState {
  AffD T_prev, Tmot, Tcurrent, Tguess;
  AffD kf_pose[MAX_KF];
  int kf_slot[MAX_KF];  // store slot of keyframe i (0 = oldest)
  int n_kf;
  int frame;
};

struct OdomDevParams {
  int K;  // submap_scan_size
  int use_guess, use_keyframe;
  double min_keyframe_dist, min_keyframe_rot_rad;
};

__global__ void k_odom_reset(SeqState* st, int n_seq) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seq) return;
  SeqState z;
  z.T_prev = z.Tmot = z.Tcurrent = z.Tguess = aff_identity();
  for (int i = 0; i < MAX_KF; i++) { z.kf_pose[i] = aff_identity(); z.kf_slot[i] = i; }
  z.n_kf = 0;
  z.frame = 0;
  st[s] = z;
}

// motion vector for Compensate: Affine3dToVectorXYeZ(TprevMot) (utils.cpp:109-113)
__global__ void k_odom_motion(const SeqState* __restrict__ st, int n_seq, double* __restrict__ mot) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seq) return;
  double v[3];
  aff_to_vec(st[s].Tmot, v);
  mot[3 * s + 0] = v[0]; mot[3 * s + 1] = v[1]; mot[3 * s + 2] = v[2];
}

// FormatScans (:478-494) + Tguess (:164-168) -> registration problems
__global__ void k_odom_problems(SeqState* __restrict__ st, int n_seq, OdomDevParams P, RegProblem* __restrict__ problems,
                                int* __restrict__ fixed_set, double* __restrict__ fixed_pose) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seq) return;
  SeqState& S = st[s];
  S.Tguess = P.use_guess ? aff_mul(S.T_prev, S.Tmot) : S.T_prev;
  RegProblem pr;
  pr.n_fixed = S.n_kf;
  pr.fixed_first = s * P.K;
  pr.src_set = s * (P.K + 1) + P.K;
  pr.active = S.n_kf > 0 ? 1 : 0;
  aff_to_vec(S.Tguess, pr.src_pose);
  for (int i = 0; i < S.n_kf; i++) {
    fixed_set[s * P.K + i] = s * (P.K + 1) + S.kf_slot[i];
    double v[3];
    aff_to_vec(S.kf_pose[i], v);
    fixed_pose[(size_t)(s * P.K + i) * 3 + 0] = v[0];
    fixed_pose[(size_t)(s * P.K + i) * 3 + 1] = v[1];
    fixed_pose[(size_t)(s * P.K + i) * 3 + 2] = v[2];
  }
  problems[s] = pr;
}

// processFrame after Register (:186-249): one CTA per sequence (the CTA also copies the cells of a new keyframe)
__global__ void __launch_bounds__(256)
k_odom_update(SeqState* __restrict__ st, OdomDevParams P, const RegResult* __restrict__ results, const int* __restrict__ n_points,
              const double* __restrict__ cur_cells, const int* __restrict__ cur_count, int cell_cap, double* __restrict__ kf_cells,
              int* __restrict__ kf_count, const int* __restrict__ cells_err, const int* __restrict__ cur_samples,
              tbv_odom_out* __restrict__ outs, int* __restrict__ fused_set) {
  __shared__ int s_fuse_slot;
  const int s = blockIdx.x;
  if (threadIdx.x == 0) {
    SeqState& S = st[s];
    const RegResult r = results[s];
    tbv_odom_out o;
    memset(&o, 0, sizeof(o));
    int fuse_slot = -1;
    if (S.n_kf == 0) {  // first frame: AddToReference (:470-476), pose stays identity
      S.kf_pose[0] = aff_identity();
      S.kf_slot[0] = 0;
      S.n_kf = 1;
      fuse_slot = 0;
      o.reg_ok = 1;
      o.is_keyframe = 1;
    } else {
      S.Tcurrent = r.pose_updated ? aff_from_vec(r.pose[0], r.pose[1], r.pose[2]) : S.Tguess;
      const AffD Tmot_current = aff_mul(aff_inv(S.T_prev), S.Tcurrent);
      {  // AccelerationVelocitySanityCheck (:76-94): Tsensor = 0.25 s, limits 200 m/s, 200 m/s^2
        const double dt = 0.25, vel_limit = 200, acc_limit = 200;
        const double vx = Tmot_current.tx / dt, vy = Tmot_current.ty / dt;
        const double vel = sqrt(vx * vx + vy * vy);
        const double ax = (Tmot_current.tx - S.Tmot.tx) / (dt * dt), ay = (Tmot_current.ty - S.Tmot.ty) / (dt * dt);
        const double acc = sqrt(ax * ax + ay * ay);
        if (acc > acc_limit || vel > vel_limit) S.Tcurrent = S.Tguess;
      }
      S.Tmot = aff_mul(aff_inv(S.T_prev), S.Tcurrent);
      const AffD Tkeydiff = aff_mul(aff_inv(S.kf_pose[S.n_kf - 1]), S.Tcurrent);
      bool fuse = true;
      if (P.use_keyframe) {  // KeyFrameBasedFuse (:62-73)
        const double yaw = atan2(Tkeydiff.r10, Tkeydiff.r11);
        const double tnorm = sqrt(Tkeydiff.tx * Tkeydiff.tx + Tkeydiff.ty * Tkeydiff.ty);
        fuse = tnorm > P.min_keyframe_dist || fabs(yaw) > P.min_keyframe_rot_rad;
      }
      if (fuse) {
        if (S.n_kf < P.K) {
          fuse_slot = S.kf_slot[S.n_kf];  // unused slot (kf_slot is a permutation of 0..K-1)
          S.kf_pose[S.n_kf] = S.Tcurrent;
          S.n_kf++;
        } else {  // push_back + erase(begin): the oldest keyframe's slot is recycled
          fuse_slot = S.kf_slot[0];
          for (int i = 0; i + 1 < P.K; i++) { S.kf_slot[i] = S.kf_slot[i + 1]; S.kf_pose[i] = S.kf_pose[i + 1]; }
          S.kf_slot[P.K - 1] = fuse_slot;
          S.kf_pose[P.K - 1] = S.Tcurrent;
        }
        o.is_keyframe = 1;
      }
      S.T_prev = S.Tcurrent;
      o.reg_ok = r.success;
      o.itrs = r.itrs;
      o.lm_iterations = r.lm_iterations;
      o.num_residuals = r.num_residuals;
      o.score = r.score;
    }
    S.frame++;
    aff_to_vec(S.Tcurrent, o.pose);
    o.n_points = n_points[s];
    o.n_cells = cur_count[s];
    o.n_keyframes = S.n_kf;
    o.status = cells_err[s];
    o.n_samples = cur_samples[s];
    outs[s] = o;
    s_fuse_slot = fuse_slot;
    fused_set[s] = fuse_slot >= 0 ? s * (P.K + 1) + fuse_slot : -1;  // the set whose search grid must be rebuilt
  }
  __syncthreads();
  const int slot = s_fuse_slot;
  if (slot < 0) return;
  const int n = min(cur_count[s], cell_cap);
  const double* src = cur_cells + (size_t)s * CELL_FIELDS * cell_cap;
  double* dst = kf_cells + ((size_t)s * P.K + slot) * CELL_FIELDS * cell_cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {  // all 16 fields of a cell in flight at once
    double v[CELL_FIELDS];
#pragma unroll
    for (int f = 0; f < CELL_FIELDS; f++) v[f] = src[(size_t)f * cell_cap + i];
#pragma unroll
    for (int f = 0; f < CELL_FIELDS; f++) dst[(size_t)f * cell_cap + i] = v[f];
  }
  if (threadIdx.x == 0) kf_count[s * P.K + slot] = n;
}

}  // namespace tbv

using namespace tbv;

struct tbv_odom {
  tbv_ctx* ctx = nullptr;
  int n_seq = 0, n_az = 0, n_range = 0, K = 0, cell_cap = 0, sample_cap = 0;
  tbv_odom_params par;
  OdomDevParams dpar;
  RegParamsDev rpar;
  CellsParams cpar;
  DevBuf<SeqState> state;
  DevBuf<double> mot, fixed_pose;
  DevBuf<RegProblem> problems;
  DevBuf<int> fixed_set, fused_set;
  GridStore kf_grids;
  DevBuf<RegResult> results;
  DevBuf<SetView> views;
  DevBuf<tbv_odom_out> outs_dev;
  CellStore cur, kf;
  // host-input path: double-buffered uploads on a copy stream
  DevBuf<uint8_t> polar[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t uploaded[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
  tbv_odom_out* outs_host[2] = {nullptr, nullptr};  // pinned
  int n_submitted = 0, n_collected = 0;
  // CUDA graphs of the step, one per input buffer the caller keeps handing in (the two upload buffers of the submit / collect pipeline, or a
  // caller's own ring): the second step on a buffer is captured, every later one is a single cudaGraphLaunch instead of 7 kernel launches
  struct StepGraph { const uint8_t* key = nullptr; cudaGraphExec_t exec = nullptr; int launches = 0; int seen = 0; uint64_t fp = 0; };
  StepGraph graphs[4];
  int use_graphs = 1, steps_done = 0;
  int wire_range_major = 0;        // scans arrive [n_range][n_az] (MulRan wire layout) and are rotated on receipt (tbv_odom_set_wire_layout)
  DevBuf<uint8_t> rotated;         // [n_seq][n_az][n_range] azimuth-major copies of the step's scans
  // Overlapped steps (tbv_odom_set_overlap): the filter of step t + 1 does not depend on step t (only its compensation does), so it runs
  // on a second, low-priority stream into a second set of clouds and fills the SMs the registration of step t leaves idle as its CTAs
  // retire; the compensation is then its own small launch on the context's stream (same arithmetic, same bits).
  int overlap = 0, parity = 0;
  cudaStream_t filter_stream = nullptr;
  DevCloud alt_f, alt_p;                                  // the clouds the context's filter state does NOT point at right now
  cudaEvent_t ev_filtered[2] = {nullptr, nullptr};        // [parity] the filter of the step has written its clouds
  cudaEvent_t ev_free[2] = {nullptr, nullptr};            // [parity] the step that used these clouds has read them for the last time
  bool free_valid[2] = {false, false};
  cudaEvent_t ev_sync = nullptr;                          // orders the filter stream behind other users of the context's clouds
  unsigned long long seen_epoch = 0;                      // ctx->filt_epoch at this fuser's last step
  cudaEvent_t input_ready = nullptr;                      // set by tbv_odom_submit for the step being enqueued: its scans are uploaded
  cudaStream_t input_consumer = nullptr;                  // the stream on which that step read its scans (for the upload ring's reuse event)
};

static void odom_free(tbv_odom* od) {
  if (!od) return;
  cudaSetDevice(od->ctx->device);
  cudaStreamSynchronize(od->ctx->stream);
  if (od->copy_stream) cudaStreamSynchronize(od->copy_stream);
  if (od->filter_stream) { cudaStreamSynchronize(od->filter_stream); cudaStreamDestroy(od->filter_stream); }
  for (int i = 0; i < 2; i++) {
    if (od->ev_filtered[i]) cudaEventDestroy(od->ev_filtered[i]);
    if (od->ev_free[i]) cudaEventDestroy(od->ev_free[i]);
  }
  if (od->ev_sync) cudaEventDestroy(od->ev_sync);
  od->alt_f.release(); od->alt_p.release();
  od->state.release(); od->mot.release(); od->fixed_pose.release(); od->problems.release(); od->fixed_set.release();
  od->rotated.release();
  for (auto& g : od->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  od->results.release(); od->views.release(); od->fused_set.release(); od->kf_grids.release(); od->outs_dev.release(); od->cur.release(); od->kf.release();
  for (int i = 0; i < 2; i++) {
    od->polar[i].release();
    if (od->uploaded[i]) cudaEventDestroy(od->uploaded[i]);
    if (od->consumed[i]) cudaEventDestroy(od->consumed[i]);
    if (od->done[i]) cudaEventDestroy(od->done[i]);
    if (od->outs_host[i]) cudaFreeHost(od->outs_host[i]);
  }
  if (od->copy_stream) cudaStreamDestroy(od->copy_stream);
  delete od;
}

static int odom_init(tbv_odom* od) {
  tbv_ctx* ctx = od->ctx;
  const int n_seq = od->n_seq, K = od->K;
  int rc;
  if ((rc = od->state.reserve(n_seq)) || (rc = od->mot.reserve((size_t)n_seq * 3)) || (rc = od->fixed_pose.reserve((size_t)n_seq * K * 3)) ||
      (rc = od->problems.reserve(n_seq)) || (rc = od->fixed_set.reserve((size_t)n_seq * K)) || (rc = od->results.reserve(n_seq)) ||
      (rc = od->views.reserve((size_t)n_seq * (K + 1))) || (rc = od->outs_dev.reserve(n_seq)) || (rc = od->cur.reserve(n_seq, od->cell_cap)) ||
      (rc = od->kf.reserve(n_seq * K, od->cell_cap)) || (rc = od->fused_set.reserve(n_seq)) || (rc = od->kf_grids.reserve(n_seq * K, od->cell_cap)))
    return rc;
  TBV_CUDA(cudaMemsetAsync(od->kf.count.p, 0, (size_t)n_seq * K * sizeof(int), ctx->stream));
  TBV_CUDA(cudaMemsetAsync(od->kf_grids.hdr.p, 0, (size_t)n_seq * K * sizeof(CellGrid), ctx->stream));
  TBV_CUDA(cudaMemsetAsync(od->cur.count.p, 0, (size_t)n_seq * sizeof(int), ctx->stream));
  TBV_CUDA(cudaMemsetAsync(od->fixed_set.p, 0, (size_t)n_seq * K * sizeof(int), ctx->stream));
  TBV_CUDA(cudaMemsetAsync(od->fixed_pose.p, 0, (size_t)n_seq * K * 3 * sizeof(double), ctx->stream));
  std::vector<SetView> hv((size_t)n_seq * (K + 1));
  for (int s = 0; s < n_seq; s++) {
    for (int i = 0; i < K; i++)
      hv[(size_t)s * (K + 1) + i] = od->kf_grids.view(s * K + i, od->kf.set_ptr(s * K + i), od->cell_cap, od->kf.count.p + s * K + i, 0);
    hv[(size_t)s * (K + 1) + K] = SetView{od->cur.set_ptr(s), od->cell_cap, od->cur.count.p + s, 0, nullptr, nullptr, nullptr};
  }
  TBV_CUDA(cudaMemcpyAsync(od->views.p, hv.data(), hv.size() * sizeof(SetView), cudaMemcpyHostToDevice, ctx->stream));
  k_odom_reset<<<(n_seq + 127) / 128, 128, 0, ctx->stream>>>(od->state.p, n_seq);
  launched(ctx, "k_odom_reset");
  TBV_CUDA(cudaGetLastError());
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

// enqueue one step on ctx->stream; scans are on the device
static int odom_enqueue(tbv_odom* od, const uint8_t* polar_dev) {
  tbv_ctx* ctx = od->ctx;
  const int n_seq = od->n_seq;
  cudaStream_t st = ctx->stream;
  // previous frame-to-frame motion first: K2 compensates the points as it emits them (odometrykeyframefuser.cpp:146-150)
  int rc;
  od->input_consumer = st;
  if (od->overlap) {
    // ---- the filter on its own stream, into the other set of clouds; nothing in it depends on the previous step --------------------------
    FilterState& Fs = ctx->filt;
    std::swap(Fs.filtered, od->alt_f); std::swap(Fs.peaks, od->alt_p);
    const int pr = od->parity; od->parity ^= 1;
    cudaStream_t fs = ctx->prof.on ? st : od->filter_stream;   // a profile is taken with every launch on the context's stream
    od->input_consumer = fs;
    if (fs != st) {
      if (ctx->filt_epoch != od->seen_epoch) {   // somebody else filtered on this context since the last step: its clouds are one of the two sets
        TBV_CUDA(cudaEventRecord(od->ev_sync, st));                                          // this step is about to overwrite — wait for whatever
        TBV_CUDA(cudaStreamWaitEvent(fs, od->ev_sync, 0));                                   // reads them on the context's stream (one step without overlap)
        od->seen_epoch = ctx->filt_epoch;
      }
      if (od->free_valid[pr]) TBV_CUDA(cudaStreamWaitEvent(fs, od->ev_free[pr], 0));        // step t - 2 has read these clouds for the last time
      if (od->input_ready) TBV_CUDA(cudaStreamWaitEvent(fs, od->input_ready, 0));           // host-input pipeline: the scans are uploaded
    } else if (od->input_ready) {
      TBV_CUDA(cudaStreamWaitEvent(st, od->input_ready, 0));
    }
    if (od->wire_range_major) {   // radar_driver.cpp:80-84: cv::rotate(ROTATE_90_COUNTERCLOCKWISE) on receipt
      if ((rc = od->rotated.reserve((size_t)n_seq * od->n_az * od->n_range))) return rc;
      if ((rc = rotate90ccw_dev(ctx, polar_dev, od->n_range, od->n_az, n_seq, od->rotated.p, fs))) return rc;
      polar_dev = od->rotated.p;
    }
    if ((rc = filter_kstrongest_dev(ctx, polar_dev, od->n_az, od->n_range, (size_t)od->n_range, n_seq, &od->par.filter, 1, nullptr, od->par.radar_ccw, fs)))
      return rc;
    if (fs != st) {
      TBV_CUDA(cudaEventRecord(od->ev_filtered[pr], fs));
      TBV_CUDA(cudaStreamWaitEvent(st, od->ev_filtered[pr], 0));
    }
    if (od->par.compensate) {   // previous frame-to-frame motion (odometrykeyframefuser.cpp:146-150), then the compensation of both clouds
      k_odom_motion<<<(n_seq + 127) / 128, 128, 0, st>>>(od->state.p, n_seq, od->mot.p);
      launched(ctx, "k_odom_motion");
      if ((rc = compensate_polar_clouds_dev(ctx, od->mot.p, od->par.radar_ccw, 1))) return rc;
    }
  } else {
  if (od->wire_range_major) {   // radar_driver.cpp:80-84: cv::rotate(ROTATE_90_COUNTERCLOCKWISE) on receipt
    if ((rc = od->rotated.reserve((size_t)n_seq * od->n_az * od->n_range))) return rc;
    if ((rc = rotate90ccw_dev(ctx, polar_dev, od->n_range, od->n_az, n_seq, od->rotated.p))) return rc;
    polar_dev = od->rotated.p;
  }
  if (od->par.compensate) {
    k_odom_motion<<<(n_seq + 127) / 128, 128, 0, st>>>(od->state.p, n_seq, od->mot.p);
    launched(ctx, "k_odom_motion");
  }
  if ((rc = filter_kstrongest_dev(ctx, polar_dev, od->n_az, od->n_range, (size_t)od->n_range, n_seq, &od->par.filter, 1,
                                  od->par.compensate ? od->mot.p : nullptr, od->par.radar_ccw)))
    return rc;
  }
  FilterState& F = ctx->filt;
  if ((rc = cells_build_dev(ctx, F.filtered.x.p, F.filtered.y.p, F.filtered.inten.p, nullptr, F.filtered.count.p, F.filtered.cap, n_seq, od->cpar,
                            od->cell_cap, od->cur)))
    return rc;
  k_odom_problems<<<(n_seq + 127) / 128, 128, 0, st>>>(od->state.p, n_seq, od->dpar, od->problems.p, od->fixed_set.p, od->fixed_pose.p);
  launched(ctx, "k_odom_problems");
  if ((rc = register_launch(ctx, REG_MODE_REGISTER, 0, od->views.p, od->problems.p, od->fixed_set.p, od->fixed_pose.p, n_seq, od->K, od->cell_cap,
                            od->cell_cap, od->rpar, od->results.p, nullptr, false)))
    return rc;
  k_odom_update<<<n_seq, 256, 0, st>>>(od->state.p, od->dpar, od->results.p, F.filtered.count.p, od->cur.f64.p, od->cur.count.p, od->cell_cap,
                                       od->kf.f64.p, od->kf.count.p, cells_err_dev(ctx), od->cur.n_samples.p, od->outs_dev.p, od->fused_set.p);
  launched(ctx, "k_odom_update");
  TBV_CUDA(cudaGetLastError());
  if (od->overlap) {   // the step's clouds have been read for the last time (cells: points, update: counts)
    const int pr = od->parity ^ 1;
    TBV_CUDA(cudaEventRecord(od->ev_free[pr], st));
    od->free_valid[pr] = true;
  }
  // new keyframes get their search grid now (used by the registrations of the following frames)
  return cellgrid_build_launch(ctx, od->views.p, od->fused_set.p, n_seq, n_seq * (od->K + 1), od->cell_cap, od->cpar.max_extent);
}

// One step through a cached CUDA graph when the input buffer has been seen before (the launch-bound regime is the online one: one
// sequence per step, where the 7 launches of the step cost as much as its kernels); direct launches otherwise and while profiling.
static uint64_t step_fingerprint(tbv_odom* od) {   // every context-level buffer the step's kernels were handed
  tbv_ctx* ctx = od->ctx;
  const FilterState& F = ctx->filt;
  uint64_t h = fp_mix(cells_fingerprint(ctx), (const void*)(uintptr_t)reg_fingerprint(ctx));
  for (const DevCloud* c : {&F.filtered, &F.peaks})
    for (const void* p : {(const void*)c->x.p, (const void*)c->y.p, (const void*)c->inten.p, (const void*)c->az.p, (const void*)c->rg.p, (const void*)c->count.p})
      h = fp_mix(h, p);
  h = fp_mix(h, F.cs_table.p); h = fp_mix(h, F.th_table.p); h = fp_mix(h, od->rotated.p);
  return h;
}

static int odom_enqueue_graphed(tbv_odom* od, const uint8_t* polar_dev) {
  tbv_ctx* ctx = od->ctx;
  od->input_consumer = ctx->stream;
  if (!od->use_graphs || od->overlap || ctx->prof.on || od->steps_done < 1) {   // the first step allocates and sets kernel attributes: never captured;
                                                                                 // an overlapped step spans two streams that must not be joined per step
    od->steps_done++;
    return odom_enqueue(od, polar_dev);
  }
  od->steps_done++;
  tbv_odom::StepGraph* slot = nullptr;
  for (auto& g : od->graphs)
    if (g.key == polar_dev) { slot = &g; break; }
  if (slot && slot->exec) {
    if (slot->fp == step_fingerprint(od)) {
      TBV_CUDA(cudaGraphLaunch(slot->exec, ctx->stream));
      ctx->launches += slot->launches;
      return TBV_OK;
    }
    cudaGraphExecDestroy(slot->exec);   // a scratch buffer of the context moved since the capture (another caller grew it): capture again
    slot->exec = nullptr;
  }
  if (!slot) {   // first time this buffer is seen: remember it (replacing the least used entry) and launch directly
    slot = &od->graphs[0];
    for (auto& g : od->graphs)
      if (g.seen < slot->seen) slot = &g;
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
    slot->key = polar_dev; slot->seen = 1; slot->launches = 0;
    return odom_enqueue(od, polar_dev);
  }
  // second step on this buffer: capture it
  slot->seen++;
  const long long before = ctx->launches;
  cudaGraph_t graph = nullptr;
  TBV_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = odom_enqueue(od, polar_dev);
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc != TBV_OK || e != cudaSuccess || !graph) {   // not capturable in this state: run directly, stop trying
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->launches = before;
    od->use_graphs = 0;
    return odom_enqueue(od, polar_dev);
  }
  slot->fp = step_fingerprint(od);
  slot->launches = (int)(ctx->launches - before);   // counted once during the capture: that count stands for the launch below
  cudaError_t ei = cudaGraphInstantiate(&slot->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { slot->exec = nullptr; cudaGetLastError(); ctx->launches = before; od->use_graphs = 0; return odom_enqueue(od, polar_dev); }
  TBV_CUDA(cudaGraphLaunch(slot->exec, ctx->stream));
  return TBV_OK;
}

extern "C" {

tbv_odom* tbv_odom_create(tbv_ctx* ctx, int n_seq, int n_az, int n_range, const tbv_odom_params* params) {
  TBV_ENTER(ctx);
  if (!ctx || !params || n_seq <= 0 || n_az <= 0 || n_range <= 0) { set_error("tbv_odom_create: bad arguments"); return nullptr; }
  if (params->submap_scan_size < 1 || params->submap_scan_size > MAX_KF) { set_error("submap_scan_size must be in [1,%d]", MAX_KF); return nullptr; }
  if (!(params->res > 0) || !(params->downsample_factor > 0)) { set_error("res and downsample_factor must be positive"); return nullptr; }
  cudaSetDevice(ctx->device);
  tbv_odom* od = new tbv_odom();
  od->ctx = ctx; od->n_seq = n_seq; od->n_az = n_az; od->n_range = n_range; od->par = *params;
  od->K = params->submap_scan_size;
  od->cell_cap = params->cell_capacity > 0 ? params->cell_capacity : 1024;
  od->sample_cap = params->sample_capacity > 0 ? params->sample_capacity : 4096;
  od->dpar.K = od->K; od->dpar.use_guess = params->use_guess; od->dpar.use_keyframe = params->use_keyframe;
  od->dpar.min_keyframe_dist = params->min_keyframe_dist;
  od->dpar.min_keyframe_rot_rad = params->min_keyframe_rot_deg * M_PI / 180.0;  // odometrykeyframefuser.cpp:69
  od->rpar = to_dev(params->reg);
  od->cpar.radius = (float)params->res;  // MapPointNormal(cloud, par.res, ...) takes float radius (pointnormal.h:118)
  od->cpar.downsample_factor = params->downsample_factor;
  od->cpar.weight_intensity = params->weight_intensity;
  od->cpar.origin[0] = od->cpar.origin[1] = 0.0;
  od->cpar.max_extent = (double)n_range * (double)params->filter.range_res * 1.05 + 8.0;  // scan radius + compensation margin
  od->cpar.max_samples = od->sample_cap;
  if (odom_init(od) != TBV_OK) { odom_free(od); return nullptr; }
  return od;
}

void tbv_odom_destroy(tbv_odom* od) { odom_free(od); }

int tbv_odom_reset(tbv_odom* od) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od, "null handle");
  tbv_ctx* ctx = od->ctx;
  cudaSetDevice(ctx->device);
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (od->copy_stream) TBV_CUDA(cudaStreamSynchronize(od->copy_stream));
  if (od->filter_stream) TBV_CUDA(cudaStreamSynchronize(od->filter_stream));
  od->free_valid[0] = od->free_valid[1] = false;
  od->n_submitted = od->n_collected = 0;
  TBV_CUDA(cudaMemsetAsync(od->kf.count.p, 0, (size_t)od->n_seq * od->K * sizeof(int), ctx->stream));
  k_odom_reset<<<(od->n_seq + 127) / 128, 128, 0, ctx->stream>>>(od->state.p, od->n_seq);
  launched(ctx, "k_odom_reset");
  TBV_CUDA(cudaGetLastError());
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

int tbv_odom_set_graphs(tbv_odom* od, int enable) {
  TBV_REQUIRE(od, "null handle");
  od->use_graphs = enable != 0;
  return TBV_OK;
}

int tbv_odom_set_overlap(tbv_odom* od, int enable) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od, "null handle");
  tbv_ctx* ctx = od->ctx;
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (od->filter_stream) TBV_CUDA(cudaStreamSynchronize(od->filter_stream));
  if (enable && !od->filter_stream) {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    TBV_CUDA(cudaStreamCreateWithPriority(&od->filter_stream, cudaStreamNonBlocking, least));
    for (int i = 0; i < 2; i++) {
      TBV_CUDA(cudaEventCreateWithFlags(&od->ev_filtered[i], cudaEventDisableTiming));
      TBV_CUDA(cudaEventCreateWithFlags(&od->ev_free[i], cudaEventDisableTiming));
    }
    TBV_CUDA(cudaEventCreateWithFlags(&od->ev_sync, cudaEventDisableTiming));
  }
  od->seen_epoch = ctx->filt_epoch;
  od->free_valid[0] = od->free_valid[1] = false;
  od->overlap = enable != 0;
  return TBV_OK;
}

int tbv_odom_set_wire_layout(tbv_odom* od, int range_major) {
  TBV_REQUIRE(od, "null handle");
  if (od->wire_range_major != (range_major != 0))
    for (auto& g : od->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g = tbv_odom::StepGraph(); }   // captured steps had the other layout
  od->wire_range_major = range_major != 0;
  return TBV_OK;
}

int tbv_odom_step_dev(tbv_odom* od, const uint8_t* polar_dev) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && polar_dev, "null pointer");
  cudaSetDevice(od->ctx->device);
  return odom_enqueue_graphed(od, polar_dev);
}

int tbv_odom_fetch(tbv_odom* od, tbv_odom_out* out) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && out, "null pointer");
  TBV_CUDA(cudaMemcpyAsync(out, od->outs_dev.p, (size_t)od->n_seq * sizeof(tbv_odom_out), cudaMemcpyDeviceToHost, od->ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(od->ctx->stream));
  return TBV_OK;
}

static int odom_pipeline_init(tbv_odom* od) {
  if (od->copy_stream) return TBV_OK;
  TBV_CUDA(cudaStreamCreateWithFlags(&od->copy_stream, cudaStreamNonBlocking));
  const size_t bytes = (size_t)od->n_seq * od->n_az * od->n_range;
  for (int i = 0; i < 2; i++) {
    int rc = od->polar[i].reserve(bytes);
    if (rc) return rc;
    TBV_CUDA(cudaEventCreateWithFlags(&od->uploaded[i], cudaEventDisableTiming));
    TBV_CUDA(cudaEventCreateWithFlags(&od->consumed[i], cudaEventDisableTiming));
    TBV_CUDA(cudaEventCreateWithFlags(&od->done[i], cudaEventDisableTiming));
    TBV_CUDA(cudaHostAlloc((void**)&od->outs_host[i], (size_t)od->n_seq * sizeof(tbv_odom_out), cudaHostAllocDefault));
  }
  return TBV_OK;
}

int tbv_odom_submit(tbv_odom* od, const uint8_t* polar_host) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && polar_host, "null pointer");
  TBV_REQUIRE(od->n_submitted - od->n_collected < 2, "two steps are already in flight: call tbv_odom_collect first");
  cudaSetDevice(od->ctx->device);
  int rc = odom_pipeline_init(od);
  if (rc) return rc;
  const int b = od->n_submitted & 1;
  const size_t bytes = (size_t)od->n_seq * od->n_az * od->n_range;
  // the buffer may still be read by the step submitted two calls ago
  if (od->n_submitted >= 2) TBV_CUDA(cudaStreamWaitEvent(od->copy_stream, od->consumed[b], 0));
  TBV_CUDA(cudaMemcpyAsync(od->polar[b].p, polar_host, bytes, cudaMemcpyHostToDevice, od->copy_stream));
  TBV_CUDA(cudaEventRecord(od->uploaded[b], od->copy_stream));
  if (od->overlap) od->input_ready = od->uploaded[b];   // awaited by the stream that reads the scans (the filter stream of an overlapped step)
  else TBV_CUDA(cudaStreamWaitEvent(od->ctx->stream, od->uploaded[b], 0));
  rc = odom_enqueue_graphed(od, od->polar[b].p);
  od->input_ready = nullptr;
  if (rc) return rc;
  TBV_CUDA(cudaEventRecord(od->consumed[b], od->input_consumer));
  TBV_CUDA(cudaMemcpyAsync(od->outs_host[b], od->outs_dev.p, (size_t)od->n_seq * sizeof(tbv_odom_out), cudaMemcpyDeviceToHost, od->ctx->stream));
  TBV_CUDA(cudaEventRecord(od->done[b], od->ctx->stream));
  od->n_submitted++;
  return TBV_OK;
}

int tbv_odom_collect(tbv_odom* od, tbv_odom_out* out) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && out, "null pointer");
  TBV_REQUIRE(od->n_collected < od->n_submitted, "nothing in flight");
  const int b = od->n_collected & 1;
  TBV_CUDA(cudaEventSynchronize(od->done[b]));
  memcpy(out, od->outs_host[b], (size_t)od->n_seq * sizeof(tbv_odom_out));
  od->n_collected++;
  return TBV_OK;
}

int tbv_odom_step(tbv_odom* od, const uint8_t* polar_host, tbv_odom_out* out) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && polar_host && out, "null pointer");
  TBV_REQUIRE(od->n_submitted == od->n_collected, "steps are in flight: collect them before a synchronous step");
  int rc = tbv_odom_submit(od, polar_host);
  if (rc) return rc;
  return tbv_odom_collect(od, out);
}

int tbv_odom_cells(tbv_odom* od, int seq, int keyframe, tbv_cell* cells, int capacity, int* n_cells, double pose[3]) {
  TBV_ENTER(od ? od->ctx : nullptr);
  TBV_REQUIRE(od && cells && n_cells && seq >= 0 && seq < od->n_seq && capacity > 0, "bad arguments");
  tbv_ctx* ctx = od->ctx;
  cudaSetDevice(ctx->device);
  SeqState hs;
  TBV_CUDA(cudaMemcpyAsync(&hs, od->state.p + seq, sizeof(SeqState), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  const double* set;
  const int* cnt_dev;
  AffD T;
  if (keyframe < 0) {
    set = od->cur.set_ptr(seq); cnt_dev = od->cur.count.p + seq; T = hs.Tcurrent;
  } else {
    TBV_REQUIRE(keyframe < hs.n_kf, "keyframe index outside the window");
    const int slot = hs.kf_slot[keyframe];
    set = od->kf.set_ptr(seq * od->K + slot); cnt_dev = od->kf.count.p + seq * od->K + slot; T = hs.kf_pose[keyframe];
  }
  int n = 0;
  TBV_CUDA(cudaMemcpyAsync(&n, cnt_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_cells = n;
  if (pose) { pose[0] = T.tx; pose[1] = T.ty; pose[2] = std::atan2(T.r10, T.r11); }
  const int ncopy = n < capacity ? n : capacity;
  int rc = cells_download(ctx, set, od->cell_cap, ncopy, cells);
  if (rc) return rc;
  if (n > capacity) { set_error("capacity %d < %d cells", capacity, n); return TBV_ERR_CAPACITY; }
  return TBV_OK;
}

}  // extern "C"

// k_register.cu — K4 (correspondences + robustified cost / J^T J / J^T r) and K5 (association loop + Ceres-style
// Levenberg-Marquardt), one CTA per registration problem, everything on the device.
//
// Replaces n_scan_normal_reg::{Register, GetCost, AddScanPairCost, BuildOptimizationProblem, SolveOptimizationProblem}
// (cfear_radarodometry/src/cfear_radarodometry/n_scan_normal.cpp:82-211, 213-324, 342-389, 441-450), the cost functors
// P2LEfficientCost / P2PEfficientCost / P2DEfficientCost (include/cfear_radarodometry/n_scan_normal.h:180-255, 330-361),
// Registration::{Weights, GetWeight, GetLoss} (registration.h:88-101, registration.cpp:67-96), MapPointNormal::GetClosestIdx
// (pointnormal.cpp:238-254) and the ceres::Solve call (trust region / Levenberg-Marquardt, Ceres 2.1.0 defaults).
//
// Design
//   * a "problem" = up to max_fixed fixed cell sets (keyframes) + one moving set; the CTA keeps the whole outer
//     association loop and all LM iterations on chip: no host round trip per iteration, thousands of problems per launch
//     (odometry: one per sequence; loop closure: one per candidate);
//   * association: every fixed set has a 4 m search grid over its cell means (float: the reference searches a float kd-tree of
//     pcl::PointXY) whose bucket rows are sorted by x when the grid is built; the CTA stages the grid entries + one offset per bucket
//     row in shared memory, and a query is, per bucket row its square touches, a binary search for the square's left edge + a scan to
//     its right edge — the exhaustive 1-NN + radius test the kd-tree answers, lowest cell index on exact ties;
//   * every warp owns an equal share of the moving cells against ALL fixed sets and writes its accepted correspondences — exactly
//     what a cost functor captures: source mean, target mean and normal / sqrt-information in the world frame, loss weight — into its
//     own segment of the block scratch (tiles of 32 blocks, fields at constant offsets); the warp that writes a block evaluates it;
//   * one evaluation = every warp sweeps its segment lane-strided, accumulates cost, g = J^T r (3) and H = J^T J (6 unique)
//     of the loss-corrected residuals/Jacobians, fixed-shape warp-shuffle + cross-warp reduction (deterministic);
//   * lane 0 of warp 0 runs the trust-region logic on the 3x3 system: Jacobi scaling, LM diagonal clamp, Cholesky, model
//     cost change, step quality, radius update, the three tolerance tests — the same state machine as
//     ceres::internal::TrustRegionMinimizer with max_consecutive_nonmonotonic_steps = 0;
//   * the configuration every caller on the hot path uses (registration mode, P2L, Huber) has its own instantiation that carries no
//     other cost function, loss or mode: 64 registers per thread (4 CTAs per SM: one wave for 592 problems) with few spills.
// fp64 throughout (fp32 only where the reference uses float: the NN search), built with -fmad=false.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "tbv_reg.cuh"

namespace tbv {

constexpr int RG_THREADS = 256;            // threads per problem (CTA) of the standard launch; the single-fixed-scan launch uses RG_THREADS_SMALL
constexpr int RG_WARPS = RG_THREADS / 32;  // = the largest number of warps a CTA of this kernel has (sizes of the per-warp tables)
constexpr int RG_THREADS_SMALL = 128;
constexpr int NACC = 10;  // cost, g0..g2, H00,H01,H02,H11,H12,H22

struct Aff {
  double r00, r01, r10, r11, tx, ty;
};
__device__ __forceinline__ Aff vec_to_aff(double x, double y, double th) {  // registration.cpp:129-135
  Aff T;
  const double c = cos(th), s = sin(th);
  T.r00 = c; T.r01 = -s; T.r10 = s; T.r11 = c; T.tx = x; T.ty = y;
  return T;
}
__device__ __forceinline__ Aff aff_mul(const Aff& A, const Aff& B) {
  Aff C;
  C.r00 = A.r00 * B.r00 + A.r01 * B.r10;
  C.r01 = A.r00 * B.r01 + A.r01 * B.r11;
  C.r10 = A.r10 * B.r00 + A.r11 * B.r10;
  C.r11 = A.r10 * B.r01 + A.r11 * B.r11;
  C.tx = (A.r00 * B.tx + A.r01 * B.ty) + A.tx;
  C.ty = (A.r10 * B.tx + A.r11 * B.ty) + A.ty;
  return C;
}
__device__ __forceinline__ Aff aff_inv(const Aff& A) {
  Aff I;
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

// ---- Ceres 2.1.0 loss functions wrapped in ScaledLoss (registration.cpp:77-96, n_scan_normal.cpp:275) -------------------
__device__ __forceinline__ void loss_huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = fmax(DBL_MIN, a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}
__device__ __forceinline__ void loss_cauchy(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * log(sum);
  rho[1] = fmax(DBL_MIN, inv);
  rho[2] = -c * (inv * inv);
}
__device__ void scaled_loss(int loss, double limit, double w, double s, double rho[3]) {
  switch (loss) {
    case TBV_LOSS_HUBER: loss_huber(limit, s, rho); break;
    case TBV_LOSS_CAUCHY: loss_cauchy(limit, s, rho); break;
    case TBV_LOSS_SOFTLONE: {
      const double b = limit * limit, c = 1.0 / b;
      const double sum = 1.0 + s * c;
      const double tmp = sqrt(sum);
      rho[0] = 2.0 * b * (tmp - 1.0);
      rho[1] = fmax(DBL_MIN, 1.0 / tmp);
      rho[2] = -(c * rho[1]) / (2.0 * sum);
      break;
    }
    case TBV_LOSS_TUKEY: {
      const double a2 = limit * limit;
      if (s <= a2) {
        const double value = 1.0 - s / a2;
        const double value_sq = value * value;
        rho[0] = a2 / 3.0 * (1.0 - value_sq * value);
        rho[1] = value_sq;
        rho[2] = -2.0 / a2 * value;
      } else {
        rho[0] = a2 / 3.0; rho[1] = 0.0; rho[2] = 0.0;
      }
      break;
    }
    case TBV_LOSS_COMBINED: {  // ComposedLoss(Huber(1), Cauchy(1))
      double rg[3], rf[3];
      loss_cauchy(1.0, s, rg);
      loss_huber(1.0, rg[0], rf);
      rho[0] = rf[0];
      rho[1] = rf[1] * rg[1];
      rho[2] = rf[2] * rg[1] * rg[1] + rf[1] * rg[2];
      break;
    }
    default:  // ScaledLoss with a null inner loss
      rho[0] = w * s; rho[1] = w; rho[2] = 0.0;
      return;
  }
  rho[0] *= w; rho[1] *= w; rho[2] *= w;
}

// One residual block at x: loss-corrected residuals f[n] and Jacobian J[n][3], returns 0.5*rho(s).
// blk points at field 0 of the block (stride = field stride).
__device__ __forceinline__ double eval_block(int cost_type, int loss, double limit, const double* __restrict__ blk, size_t stride,
                                             double x0, double x1, double cy, double sy, double f[2], double J[6], int& n, bool want_jac) {
  const double sx = blk[0], sy_ = blk[stride], tx = blk[2 * stride], ty = blk[3 * stride];
  const double a4 = blk[4 * stride], a5 = blk[5 * stride], w = blk[7 * stride];
  const double mx = (cy * sx + (-sy) * sy_) + x0;
  const double my = (sy * sx + cy * sy_) + x1;
  const double dmx = (-sy) * sx + (-cy) * sy_;
  const double dmy = cy * sx + (-sy) * sy_;
  if (cost_type == TBV_P2L) {
    n = 1;
    const double v0 = mx - tx, v1 = my - ty;
    f[0] = v0 * a4 + v1 * a5;
    J[0] = a4; J[1] = a5; J[2] = dmx * a4 + dmy * a5;
  } else if (cost_type == TBV_P2P) {
    n = 2;
    f[0] = tx - mx; f[1] = ty - my;
    J[0] = -1.0; J[1] = 0.0; J[2] = -dmx; J[3] = 0.0; J[4] = -1.0; J[5] = -dmy;
  } else {  // P2D: L = [a4 0; a5 a6]
    n = 2;
    const double a6 = blk[6 * stride];
    const double e0 = mx - tx, e1 = my - ty;
    f[0] = a4 * e0 + 0.0 * e1;
    f[1] = a5 * e0 + a6 * e1;
    J[0] = a4; J[1] = 0.0; J[2] = a4 * dmx + 0.0 * dmy;
    J[3] = a5; J[4] = a6; J[5] = a5 * dmx + a6 * dmy;
  }
  double sq = 0.0;
  for (int r = 0; r < n; r++) sq += f[r] * f[r];
  double rho[3];
  scaled_loss(loss, limit, w, sq, rho);
  // ceres::internal::Corrector
  const double sqrt_rho1 = sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (sq == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1;
    alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
    const double alpha = 1.0 - sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq;
  }
  if (want_jac) {
    if (alpha_sq_norm == 0.0) {
      for (int i = 0; i < n * 3; i++) J[i] *= sqrt_rho1;
    } else {
      for (int c = 0; c < 3; c++) {
        double rtj = 0.0;
        for (int r = 0; r < n; r++) rtj += J[r * 3 + c] * f[r];
        for (int r = 0; r < n; r++) J[r * 3 + c] = sqrt_rho1 * (J[r * 3 + c] - alpha_sq_norm * f[r] * rtj);
      }
    }
  }
  for (int r = 0; r < n; r++) f[r] *= residual_scaling;
  return 0.5 * rho[0];
}

// ---- LM state machine (thread 0) ----------------------------------------------------------------------------------------
struct LMState {
  double x[3], x_norm, x_cost;
  double g[3], H[6];          // at x: unscaled gradient and J^T J (00 01 02 11 12 22)
  double scal[3];             // jacobian_scaling
  double radius, decrease_factor, diag[3];
  bool reuse_diagonal;
  double ev_minimum_cost, ev_current_cost, ev_reference_cost, ev_candidate_cost, ev_acc_ref, ev_acc_cand;
  int num_consecutive_invalid;
  // current iteration summary
  int it_iteration;
  bool it_successful;
  double it_cost, it_gmax, it_gnorm, it_rel_dec;
  // pushed summaries
  int n_pushed;
  double last_rel_dec, last_gmax, last_gnorm, min_pushed_cost, initial_cost;
  int termination;            // 0 convergence, 1 no convergence, 2 failure
  bool terminated;
  double minimum_cost, params[3];
  // pending step
  double cand[3], model_cost_change;
};

__device__ void lm_gradient_norms(LMState& S) {
  double gmax = 0.0, gn = 0.0;
  for (int c = 0; c < 3; c++) {
    const double pg = S.x[c] - (S.x[c] + (-S.g[c]));
    gmax = fmax(gmax, fabs(pg));
    gn += pg * pg;
  }
  S.it_gmax = gmax;
  S.it_gnorm = sqrt(gn);
}

// Dense 3x3 Cholesky solve H y = b (lower triangle h00 h10 h11 h20 h21 h22), every operation of the textbook loops in their order, written
// out on scalars: nothing is indexed at run time, so nothing lives in local memory.
__device__ __forceinline__ bool chol_solve3(double h00, double h10, double h11, double h20, double h21, double h22, double b0, double b1, double b2,
                                            double& y0, double& y1, double& y2) {
  if (!(h00 > 0.0) || !isfinite(h00)) return false;
  const double l00 = sqrt(h00);
  const double l10 = h10 / l00;
  const double l20 = h20 / l00;
  const double d1 = h11 - l10 * l10;
  if (!(d1 > 0.0) || !isfinite(d1)) return false;
  const double l11 = sqrt(d1);
  const double l21 = (h21 - l20 * l10) / l11;
  double d2 = h22 - l20 * l20;
  d2 -= l21 * l21;
  if (!(d2 > 0.0) || !isfinite(d2)) return false;
  const double l22 = sqrt(d2);
  const double z0 = b0 / l00;
  const double z1 = (b1 - l10 * z0) / l11;
  double v = b2 - l20 * z0;
  v -= l21 * z1;
  const double z2 = v / l22;
  y2 = z2 / l22;
  y1 = (z1 - l21 * y2) / l11;
  v = z0 - l10 * y1;
  v -= l20 * y2;
  y0 = v / l00;
  return isfinite(y0) && isfinite(y1) && isfinite(y2);
}

// after the evaluation at the initial point (acc = cost, g, H)
__device__ __noinline__ void lm_begin(LMState& S, const double x0[3], const double* acc) {
  for (int c = 0; c < 3; c++) { S.x[c] = x0[c]; S.params[c] = x0[c]; }
  S.x_norm = sqrt(S.x[0] * S.x[0] + S.x[1] * S.x[1] + S.x[2] * S.x[2]);
  S.radius = 1e4; S.decrease_factor = 2.0; S.reuse_diagonal = false;
  S.diag[0] = S.diag[1] = S.diag[2] = 0.0;
  S.ev_acc_ref = 0.0; S.ev_acc_cand = 0.0;
  S.num_consecutive_invalid = 0;
  S.n_pushed = 0; S.terminated = false; S.termination = 1;
  S.it_iteration = 0; S.it_rel_dec = 0.0;
  S.x_cost = acc[0];
  for (int c = 0; c < 3; c++) S.g[c] = acc[1 + c];
  for (int c = 0; c < 6; c++) S.H[c] = acc[4 + c];
  const double hd[3] = {S.H[0], S.H[3], S.H[5]};
  for (int c = 0; c < 3; c++) S.scal[c] = 1.0 / (1.0 + sqrt(hd[c]));
  S.it_cost = S.x_cost;
  lm_gradient_norms(S);
  S.initial_cost = S.x_cost;
  S.min_pushed_cost = S.x_cost;
  S.it_successful = true;
  S.minimum_cost = S.x_cost;
  S.ev_minimum_cost = S.ev_current_cost = S.ev_reference_cost = S.ev_candidate_cost = S.x_cost;
}

// FinalizeIterationAndCheckIfMinimizerCanContinue + ComputeTrustRegionStep (+ HandleInvalidStep loop).
// Returns true when S.cand must be evaluated; false when the solve is over.
__device__ __noinline__ bool lm_advance(LMState& S, int max_iterations) {
  if (S.terminated) return false;
  for (;;) {
    // ---- Finalize
    if (S.it_successful) {
      if (S.x_cost < S.minimum_cost || S.it_iteration == 0) {
        S.minimum_cost = fmin(S.minimum_cost, S.x_cost);
        S.params[0] = S.x[0]; S.params[1] = S.x[1]; S.params[2] = S.x[2];
      }
    }
    S.n_pushed++;
    S.last_rel_dec = S.it_rel_dec; S.last_gmax = S.it_gmax; S.last_gnorm = S.it_gnorm;
    S.min_pushed_cost = fmin(S.min_pushed_cost, S.it_cost);
    if (S.it_iteration >= max_iterations) { S.termination = 1; S.terminated = true; return false; }
    if (S.it_successful && S.it_gmax <= 1e-10) { S.termination = 0; S.terminated = true; return false; }
    if (S.radius < 1e-32) { S.termination = 0; S.terminated = true; return false; }
    // ---- next iteration
    S.it_iteration++;
    S.it_successful = false;
    S.it_rel_dec = 0.0;
    S.it_cost = 0.0;
    if (!S.reuse_diagonal) {
      S.diag[0] = fmin(fmax((S.scal[0] * S.scal[0]) * S.H[0], 1e-6), 1e32);
      S.diag[1] = fmin(fmax((S.scal[1] * S.scal[1]) * S.H[3], 1e-6), 1e32);
      S.diag[2] = fmin(fmax((S.scal[2] * S.scal[2]) * S.H[5], 1e-6), 1e32);
    }
    double st0 = 0.0, st1 = 0.0, st2 = 0.0;
    bool solved;
    {
      // Hs = D H D (symmetric: (c_a c_b) H_ab == (c_b c_a) H_ba bit for bit), rhs = D g, damped diagonal Hs_aa + lmd_a^2
      const double c0 = S.scal[0], c1 = S.scal[1], c2 = S.scal[2];
      const double lmd0 = sqrt(S.diag[0] / S.radius), lmd1 = sqrt(S.diag[1] / S.radius), lmd2 = sqrt(S.diag[2] / S.radius);
      solved = chol_solve3((c0 * c0) * S.H[0] + lmd0 * lmd0, (c0 * c1) * S.H[1], (c1 * c1) * S.H[3] + lmd1 * lmd1, (c0 * c2) * S.H[2],
                           (c1 * c2) * S.H[4], (c2 * c2) * S.H[5] + lmd2 * lmd2, c0 * S.g[0], c1 * S.g[1], c2 * S.g[2], st0, st1, st2);
    }
    S.reuse_diagonal = true;
    bool valid = false;
    const double c0 = S.scal[0], c1 = S.scal[1], c2 = S.scal[2];
    if (solved) {
      // the scaled system again (the same products, recomputed from shared memory rather than kept in registers across the solve)
      const double s00 = (c0 * c0) * S.H[0], s01 = (c0 * c1) * S.H[1], s02 = (c0 * c2) * S.H[2];
      const double s11 = (c1 * c1) * S.H[3], s12 = (c1 * c2) * S.H[4], s22 = (c2 * c2) * S.H[5];
      const double r0 = c0 * S.g[0], r1 = c1 * S.g[1], r2 = c2 * S.g[2];
      st0 *= -1.0; st1 *= -1.0; st2 *= -1.0;
      double lin = 0.0, quad = 0.0;
      {
        lin += st0 * r0;
        double hv = 0.0;
        hv += s00 * st0; hv += s01 * st1; hv += s02 * st2;
        quad += st0 * hv;
      }
      {
        lin += st1 * r1;
        double hv = 0.0;
        hv += s01 * st0; hv += s11 * st1; hv += s12 * st2;
        quad += st1 * hv;
      }
      {
        lin += st2 * r2;
        double hv = 0.0;
        hv += s02 * st0; hv += s12 * st1; hv += s22 * st2;
        quad += st2 * hv;
      }
      S.model_cost_change = -(lin + quad / 2.0);
      valid = S.model_cost_change > 0.0;
    }
    if (valid) {
      S.num_consecutive_invalid = 0;
      S.cand[0] = S.x[0] + st0 * c0; S.cand[1] = S.x[1] + st1 * c1; S.cand[2] = S.x[2] + st2 * c2;
      return true;
    }
    // ---- HandleInvalidStep
    S.num_consecutive_invalid++;
    if (S.num_consecutive_invalid >= 5) { S.termination = 2; S.terminated = true; return false; }
    S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diagonal = true;
    S.it_cost = S.x_cost; S.it_rel_dec = 0.0;
    S.it_gmax = S.last_gmax; S.it_gnorm = S.last_gnorm;
  }
}

// after the evaluation at S.cand (acc = cost, g, H there)
__device__ __noinline__ void lm_candidate(LMState& S, const double* acc) {
  double candidate_cost = acc[0];
  if (!isfinite(candidate_cost)) candidate_cost = DBL_MAX;
  {
    const double d0 = S.x[0] - S.cand[0], d1 = S.x[1] - S.cand[1], d2 = S.x[2] - S.cand[2];
    const double step_norm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double tol = 1e-8 * (S.x_norm + 1e-8);
    if (step_norm <= tol) { S.termination = 0; S.terminated = true; return; }
  }
  const double cost_change = S.x_cost - candidate_cost;
  if (fabs(cost_change) <= 1e-6 * S.x_cost) { S.termination = 0; S.terminated = true; return; }
  if (candidate_cost >= DBL_MAX) {
    S.it_rel_dec = -DBL_MAX;
  } else {
    const double rd = (S.ev_current_cost - candidate_cost) / S.model_cost_change;
    const double hist = (S.ev_reference_cost - candidate_cost) / (S.ev_acc_ref + S.model_cost_change);
    S.it_rel_dec = fmax(rd, hist);
  }
  if (S.it_rel_dec > 1e-3) {
    for (int c = 0; c < 3; c++) S.x[c] = S.cand[c];
    S.x_norm = sqrt(S.x[0] * S.x[0] + S.x[1] * S.x[1] + S.x[2] * S.x[2]);
    S.x_cost = acc[0];
    for (int c = 0; c < 3; c++) S.g[c] = acc[1 + c];
    for (int c = 0; c < 6; c++) S.H[c] = acc[4 + c];
    S.it_cost = S.x_cost;
    lm_gradient_norms(S);
    S.it_successful = true;
    const double q3 = 2.0 * S.it_rel_dec - 1.0;
    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - q3 * q3 * q3);  // pow(2 rho - 1, 3)
    S.radius = fmin(1e16, S.radius);
    S.decrease_factor = 2.0;
    S.reuse_diagonal = false;
    S.ev_current_cost = candidate_cost;
    S.ev_acc_cand += S.model_cost_change;
    S.ev_acc_ref += S.model_cost_change;
    if (S.ev_current_cost < S.ev_minimum_cost) {
      S.ev_minimum_cost = S.ev_current_cost; S.ev_candidate_cost = S.ev_current_cost; S.ev_acc_cand = 0.0;
      S.ev_reference_cost = S.ev_candidate_cost; S.ev_acc_ref = S.ev_acc_cand;
    } else if (S.ev_current_cost > S.ev_candidate_cost) {
      S.ev_candidate_cost = S.ev_current_cost; S.ev_acc_cand = 0.0;
    }
  } else {
    S.it_successful = false;
    S.it_cost = candidate_cost;
    S.it_gmax = S.last_gmax; S.it_gnorm = S.last_gnorm;
    S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diagonal = true;
  }
}

// ---- uniform grid over a fixed scan's cell means (the reference's kd-tree over pcl::PointXY, pointnormal.cpp:151-162) -----
// Buckets of side 4 m >= every search radius used (2*radius_ = 4 on the first association round, radius_ = 2 afterwards): an
// accepted neighbour (d2 < R2) lies in the 3x3 block around the query's bucket and nothing outside the block can be closer
// than an accepted one, so searching the block returns the reference's global 1-NN + distance test.
constexpr float GRID_CELL = 4.0f;

__device__ __forceinline__ int set_count(const SetView& s) { return s.n_ptr ? min(*s.n_ptr, s.cap) : min(s.n_val, s.cap); }

__device__ __forceinline__ int grid_coord(float v, float mn, int n) {  // bucket of a stored point: clamped
  const float f = floorf((v - mn) / GRID_CELL);
  if (!(f >= 0.f)) return 0;
  return f >= (float)n ? n - 1 : (int)f;
}

__global__ void __launch_bounds__(256)
k_cellgrid_build(const SetView* __restrict__ sets, const int* __restrict__ which, int n_sets, int bucket_cap, int sort_cap) {
  extern __shared__ int s_cnt[];  // [bucket_cap] <= GRID_CAP: the caller's bound on the buckets of one grid (larger grids: no grid),
                                  // then [sort_cap] float4 (16-byte aligned: bucket_cap is a multiple of 4): the entries before their final order
  __shared__ float s_red[4][8];
  __shared__ int s_part[256];
  __shared__ CellGrid s_g;
  const int bi = blockIdx.x;
  const int si = which ? which[bi] : bi;
  if (si < 0 || si >= n_sets) return;
  const SetView t = sets[si];
  if (!t.grid) return;
  const int n = set_count(t);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* U0 = t.f + (size_t)CF_U0 * t.cap;
  const double* U1 = t.f + (size_t)CF_U1 * t.cap;
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
  for (int i = tid; i < n; i += 256) {
    const float a = (float)U0[i], b = (float)U1[i];
    mnx = fminf(mnx, a); mxx = fmaxf(mxx, a); mny = fminf(mny, b); mxy = fmaxf(mxy, b);
  }
  for (int d = 16; d > 0; d >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
  }
  if (lane == 0) { s_red[0][warp] = mnx; s_red[1][warp] = mny; s_red[2][warp] = mxx; s_red[3][warp] = mxy; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; w++) {
      mnx = fminf(mnx, s_red[0][w]); mny = fminf(mny, s_red[1][w]); mxx = fmaxf(mxx, s_red[2][w]); mxy = fmaxf(mxy, s_red[3][w]);
    }
    CellGrid g;
    g.minx = mnx; g.miny = mny; g.nx = 1; g.ny = 1; g.ok = 0;
    if (n > 0 && n <= 65535) {
      const float fx = floorf((mxx - mnx) / GRID_CELL), fy = floorf((mxy - mny) / GRID_CELL);
      if (fx >= 0.f && fy >= 0.f && fx < 16384.f && fy < 16384.f) {
        g.nx = (int)fx + 1; g.ny = (int)fy + 1;
        g.ok = ((long long)g.nx * g.ny <= bucket_cap) ? 1 : 0;
      }
    }
    s_g = g;
    *t.grid = g;
  }
  __syncthreads();
  const CellGrid g = s_g;
  if (!g.ok) return;
  const int nb = g.nx * g.ny;
  for (int b = tid; b < nb; b += 256) s_cnt[b] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256)
    atomicAdd(&s_cnt[grid_coord((float)U1[i], g.miny, g.ny) * g.nx + grid_coord((float)U0[i], g.minx, g.nx)], 1);
  __syncthreads();
  const int chunk = (nb + 255) / 256;
  const int b0 = min(nb, tid * chunk), b1 = min(nb, b0 + chunk);
  int sum = 0;
  for (int b = b0; b < b1; b++) sum += s_cnt[b];
  {  // exclusive scan of the 256 per-thread sums: warp shuffles + one pass over the 8 warp totals
    int inc = sum;
    for (int d = 1; d < 32; d <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t2; }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < warp; k++) wbase += s_part[k];
    __syncthreads();
    s_part[tid] = wbase + inc - sum;
  }
  __syncthreads();
  int off = s_part[tid];
  for (int b = b0; b < b1; b++) {
    const int c = s_cnt[b];
    t.gstart[b] = (uint16_t)off;
    s_cnt[b] = off;
    off += c;
  }
  if (tid == 0) t.gstart[nb] = (uint16_t)n;
  __syncthreads();
  // Entries in bucket order (row-major).  When the set fits the staging area they are also put in ascending (x, cell index) order inside
  // every bucket — buckets of a row ascend in x, so every bucket ROW is then sorted by x, which is what the per-row binary search of
  // k_register's staged nearest-neighbour search needs (header ok = 2).  Larger sets keep the arbitrary order inside a bucket (ok = 1:
  // searched bucket by bucket from global memory).
  const bool sort_rows = n <= sort_cap;
  float4* s_ent = reinterpret_cast<float4*>(s_cnt + bucket_cap);
  for (int i = tid; i < n; i += 256) {
    const float a = (float)U0[i], b = (float)U1[i];
    const int pos = atomicAdd(&s_cnt[grid_coord(b, g.miny, g.ny) * g.nx + grid_coord(a, g.minx, g.nx)], 1);
    const float4 e = make_float4(a, b, __int_as_float(i), 0.f);
    if (sort_rows) s_ent[pos] = e; else t.gent[pos] = e;
  }
  if (!sort_rows) return;
  __syncthreads();
  for (int pos = tid; pos < n; pos += 256) {
    const float4 e = s_ent[pos];
    const int idx = __float_as_int(e.z);
    const int b = grid_coord(e.y, g.miny, g.ny) * g.nx + grid_coord(e.x, g.minx, g.nx);
    const int lo = t.gstart[b], hi = t.gstart[b + 1];          // written by this block before the barrier above
    const float key = (e.x == e.x) ? e.x : -FLT_MAX;            // a NaN mean (never produced by the cells kernel) sorts first: the order stays total
    int rank = 0;
    for (int k = lo; k < hi; k++) {
      const float4 o = s_ent[k];
      const float okey = (o.x == o.x) ? o.x : -FLT_MAX;
      rank += (okey < key || (okey == key && __float_as_int(o.z) < idx)) ? 1 : 0;
    }
    t.gent[lo + rank] = e;
  }
  if (tid == 0) t.grid->ok = 2;
}

// MapPointNormal::GetClosestIdx (pointnormal.cpp:238-254): float 1-NN over the cell means, accepted iff d2 < R*R.
// g = the set's grid header (staged in shared memory by the caller).  The bounds of the (at most three) bucket rows are
// fetched together before any entry is read; an entry is one 16-byte record (mean x, y, cell index).
__device__ __forceinline__ int nn_search(const SetView& t, const CellGrid& g, int n_tgt, float qx, float qy, double R) {
  float bestd = FLT_MAX;
  int best = -1;
  if (t.grid && g.ok) {
    const float fx = floorf((qx - g.minx) / GRID_CELL), fy = floorf((qy - g.miny) / GRID_CELL);
    const int cbx = !(fx >= -2.f) ? -2 : (fx > (float)(g.nx + 1) ? g.nx + 1 : (int)fx);
    const int cby = !(fy >= -2.f) ? -2 : (fy > (float)(g.ny + 1) ? g.ny + 1 : (int)fy);
    // Buckets that can hold a point closer than R: the 3 x 3 block around the query's bucket (R <= GRID_CELL), cut down to the
    // buckets the square [q - R, q + R] (plus a float-rounding margin) touches — one or two per axis once R = radius_ = 2 m.
    // A neighbour outside that square is farther than R and would be rejected by the final test anyway, so the result is the
    // same as the exhaustive 1-NN + radius test of the reference.
    const float Rm = (float)R + 1e-3f;
    const float lx = floorf((qx - Rm - g.minx) / GRID_CELL), hx = floorf((qx + Rm - g.minx) / GRID_CELL);
    const float ly = floorf((qy - Rm - g.miny) / GRID_CELL), hy = floorf((qy + Rm - g.miny) / GRID_CELL);
    int bx0 = max(cbx - 1, 0), bx1 = min(cbx + 1, g.nx - 1);
    int by0 = max(cby - 1, 0), by1 = min(cby + 1, g.ny - 1);
    if (lx > (float)bx0) bx0 = (int)fminf(lx, (float)g.nx);
    if (hx < (float)bx1) bx1 = (int)fmaxf(hx, -1.f);
    if (ly > (float)by0) by0 = (int)fminf(ly, (float)g.ny);
    if (hy < (float)by1) by1 = (int)fmaxf(hy, -1.f);
    if (bx0 <= bx1 && by0 <= by1) {
      int s0[3], s1[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int by = by0 + r;
        s0[r] = 0; s1[r] = 0;
        if (by <= by1) { s0[r] = t.gstart[by * g.nx + bx0]; s1[r] = t.gstart[by * g.nx + bx1 + 1]; }
      }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        for (int s = s0[r]; s < s1[r]; s += 2) {
          const float4 e0 = t.gent[s];
          const float4 e1 = t.gent[s + 1 < s1[r] ? s + 1 : s];   // odd tail: the same entry again (a tie with itself changes nothing)
          {
            const int i = __float_as_int(e0.z);
            const float dx = qx - e0.x, dy = qy - e0.y;
            float dd = dx * dx;   // FLANN L2_Simple, no contraction (-fmad=false)
            dd = dd + dy * dy;
            if (dd < bestd || (dd == bestd && i < best)) { bestd = dd; best = i; }
          }
          {
            const int i = __float_as_int(e1.z);
            const float dx = qx - e1.x, dy = qy - e1.y;
            float dd = dx * dx;
            dd = dd + dy * dy;
            if (dd < bestd || (dd == bestd && i < best)) { bestd = dd; best = i; }
          }
        }
      }
    }
  } else {  // no grid (huge extent / more than 65535 cells): exhaustive scan, first minimum wins
    const double* U0 = t.f + (size_t)CF_U0 * t.cap;
    const double* U1 = t.f + (size_t)CF_U1 * t.cap;
    for (int i = 0; i < n_tgt; i++) {
      const float dx = qx - (float)U0[i], dy = qy - (float)U1[i];
      float dd = dx * dx;
      dd = dd + dy * dy;
      if (dd < bestd) { bestd = dd; best = i; }
    }
  }
  return (best >= 0 && (double)bestd < R * R) ? best : -1;
}

// The specialised registration kernels reach the global-memory search only for sets too large to stage: out of line there, so that its
// state does not compete for registers with the staged loop every problem of the hot path runs.
__device__ __noinline__ int nn_search_cold(const SetView* t, const CellGrid* g, int n_tgt, float qx, float qy, double R) {
  return nn_search(*t, *g, n_tgt, qx, qy, R);
}

// The same search over a copy of the set's grid entries in SHARED memory (grids built with ok == 2: every bucket row ascends in x).
// row[r] = first entry of bucket row r (row[ny] = number of entries).  In every bucket row the square [q - R, q + R] touches, a binary
// search finds the first entry with x >= qx - R and the scan stops at the first x > qx + R: a query looks at the handful of cells inside
// its square, whatever the row holds (a wall along x puts dozens of cells into one bucket row; with one query per lane the warp would
// otherwise wait for the lane with the longest row).  The entries visited are a superset of those within R of the query — an entry
// outside the square is farther than R and could not be accepted — so the result is the exhaustive 1-NN + radius test of the
// reference: the closest entry overall, ties to the smaller cell index, kept iff d2 < R*R.
__device__ __forceinline__ void nn_visit(const float4 e, float qx, float qy, float& bestd, int& best) {
  const int i = __float_as_int(e.z);
  const float dx = qx - e.x, dy = qy - e.y;
  float dd = dx * dx;   // FLANN L2_Simple, no contraction (-fmad=false)
  dd = dd + dy * dy;
  if (dd < bestd || (dd == bestd && i < best)) { bestd = dd; best = i; }
}
__device__ __forceinline__ int nn_search_staged(const float4* __restrict__ ent, const uint16_t* __restrict__ row, const CellGrid& g, float qx, float qy,
                                                float Rm, double R2, bool live) {
  const float ly = floorf((qy - Rm - g.miny) / GRID_CELL), hy = floorf((qy + Rm - g.miny) / GRID_CELL);
  if (!(live && hy >= 0.f && ly <= (float)(g.ny - 1))) return -1;   // the square lies outside the grid rows, q is NaN, or an idle lane
  const int by0 = ly > 0.f ? (int)ly : 0, by1 = hy < (float)(g.ny - 1) ? (int)hy : g.ny - 1;
  const float xlo = qx - Rm, xhi = qx + Rm;
  float bestd = FLT_MAX;
  int best = -1;
  for (int r = by0; r <= by1; r++) {
    int lo = row[r];
    const int end = row[r + 1];
    int n = end - lo;
    while (n > 0) {   // lower bound of xlo in the row
      const int half = n >> 1;
      if (ent[lo + half].x < xlo) { lo += half + 1; n -= half + 1; } else n = half;
    }
    for (int j = lo; j < end; j++) {
      const float4 e = ent[j];
      if (e.x > xhi) break;
      nn_visit(e, qx, qy, bestd, best);
    }
  }
  return (best >= 0 && (double)bestd < R2) ? best : -1;
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
constexpr int RG_MAX_FIXED = 16;

// State of n_scan_normal_reg::Register's association loop (n_scan_normal.cpp:96-160), advanced by one thread.
struct OuterState {
  double par[3], prev_par[3], tsrc[3];   // parameters.back(), its previous value, Tsrc.back()
  double prev_score, final_cost, last_rel_dec;
  int total_lm, itr, pose_updated, success, num_residuals, last_n_iterations, termination;
  int go;                                // 1: run another association round
};

// After one ceres::Solve (state S): the bookkeeping of n_scan_normal.cpp:112-158 and the loop header's itr++ / condition.
__device__ __noinline__ void outer_advance(OuterState& O, const LMState& S, int num_residuals, int max_itr_association) {
  O.num_residuals = num_residuals;
  O.par[0] = S.params[0]; O.par[1] = S.params[1]; O.par[2] = S.params[2];
  O.final_cost = fmin(S.initial_cost, S.min_pushed_cost);  // SetSummaryFinalCost
  O.last_rel_dec = S.last_rel_dec;
  O.last_n_iterations = S.n_pushed;
  O.termination = S.termination;
  O.total_lm += O.last_n_iterations - 1;
  O.success = O.termination != 2;  // IsSolutionUsable
  if (O.success) { O.pose_updated = 1; O.tsrc[0] = O.par[0]; O.tsrc[1] = O.par[1]; O.tsrc[2] = O.par[2]; }
  const double current_score = O.final_cost;
  const double rel_improvement = (O.prev_score - current_score) / O.prev_score;
  O.go = 0;
  if (O.itr > 3) {  // min_itr_ = 3
    if (O.prev_score < current_score) {
      O.par[0] = O.prev_par[0]; O.par[1] = O.prev_par[1]; O.par[2] = O.prev_par[2];
      return;
    } else if (rel_improvement < 0.00001) {
      return;
    } else if (O.last_rel_dec < 0.00001 || O.last_n_iterations == 1) {
      return;
    }
  }
  O.prev_score = current_score;
  O.prev_par[0] = O.par[0]; O.prev_par[1] = O.par[1]; O.prev_par[2] = O.par[2];
  O.itr++;
  O.go = (O.itr <= max_itr_association && O.success) ? 1 : 0;
}

// Per-problem constants, written once by the set-up phase.  Every phase below is its own (not inlined) function that reads what it
// needs from here: nothing of one phase stays in registers during another, so each phase has the whole 64-register budget of the
// 4-CTAs-per-SM launch for its own loop and the serial trust-region code is not squeezed out of the register file by loop state
// it never uses (local-memory traffic on that path costs an L2 round trip per access: the L1 left beside 212 KB of shared memory is tiny).
constexpr int RG_TILE = BLK_FIELDS * 32;   // doubles per tile of 32 residual blocks (layout: rg_segment below)

struct RegCtx {
  RegParamsDev P;
  const double* sf;            // moving set, field-major
  size_t scap;
  double* blocks;              // this problem's residual blocks: one segment per warp, tiles of 32 blocks (rg_segment)
  size_t bstride;              // slots per problem in the assoc / residual scratch
  int* assoc;                  // evaluation mode: [fixed][source] -> matched target or -1
  double* residuals;           // evaluation mode, optional
  const double* fixed_pose;    // poses of this problem's fixed scans
  double src_pose[3];
  int n_src, n_fixed, jc, staged, mode, slot_cap, nres_per_block;
};

struct RegShared {
  unsigned long long stage_bar;   // mbarrier: completion of the bulk copies (TMA) that stage the fixed sets' grid entries
  double acc[NACC];
  double warp_acc[RG_WARPS][NACC];
  double ex[3], cs[2];
  int flag, n_blocks, warp_cnt[RG_WARPS];   // blocks in every warp's segment
  int cnt[RG_MAX_FIXED][RG_WARPS];          // ... split by fixed scan (evaluation mode: the reference's block order for the residual vector)
  Aff Tst[RG_MAX_FIXED], Ttar[RG_MAX_FIXED], TtarInv[RG_MAX_FIXED];   // source -> fixed frame at the current pose; fixed -> world (and back)
  SetView tgt[RG_MAX_FIXED];
  CellGrid grid[RG_MAX_FIXED];   // search-grid headers of the fixed sets
  int n_tgt[RG_MAX_FIXED];
  int ent_off[RG_MAX_FIXED], row_off[RG_MAX_FIXED];           // byte offsets into the dynamic shared memory (staged working set)
  RegCtx c;
  LMState lm;
  OuterState outer;
};

// Fast evaluation of one residual block for the losses whose second derivative is never positive (Huber, none): Ceres'
// Corrector then always takes its first branch (alpha = 0), so rho'' itself — one fp64 division per block in the generic
// code — is not needed, and an inlier block's sqrt(rho') is sqrt(w), precomputed at association time (field 8).  Every
// remaining operation is the generic path's, in the same order: results are bit-identical to eval_block.
template <int COST, bool HUBER>
__device__ __forceinline__ void eval_block_simple(const double* __restrict__ blk, size_t stride, double limit, double x0, double x1, double cy,
                                                  double sy, double* __restrict__ a) {
  const double sx = blk[0], sy_ = blk[stride], tx = blk[2 * stride], ty = blk[3 * stride];
  const double a4 = blk[4 * stride], a5 = blk[5 * stride], w = blk[7 * stride], sw = blk[8 * stride];
  const double mx = (cy * sx + (-sy) * sy_) + x0;
  const double my = (sy * sx + cy * sy_) + x1;
  const double dmx = (-sy) * sx + (-cy) * sy_;
  const double dmy = cy * sx + (-sy) * sy_;
  constexpr int n = (COST == TBV_P2L) ? 1 : 2;
  double f[2], J[6];
  if (COST == TBV_P2L) {
    const double v0 = mx - tx, v1 = my - ty;
    f[0] = v0 * a4 + v1 * a5;
    J[0] = a4; J[1] = a5; J[2] = dmx * a4 + dmy * a5;
  } else if (COST == TBV_P2P) {
    f[0] = tx - mx; f[1] = ty - my;
    J[0] = -1.0; J[1] = 0.0; J[2] = -dmx; J[3] = 0.0; J[4] = -1.0; J[5] = -dmy;
  } else {
    const double a6 = blk[6 * stride];
    const double e0 = mx - tx, e1 = my - ty;
    f[0] = a4 * e0 + 0.0 * e1;
    f[1] = a5 * e0 + a6 * e1;
    J[0] = a4; J[1] = 0.0; J[2] = a4 * dmx + 0.0 * dmy;
    J[3] = a5; J[4] = a6; J[5] = a5 * dmx + a6 * dmy;
  }
  double sq = 0.0;
#pragma unroll
  for (int r = 0; r < n; r++) sq += f[r] * f[r];
  double rho0, sqrt_rho1;
  const double b = limit * limit;
  if (HUBER && sq > b) {
    const double r = sqrt(sq);
    rho0 = (2.0 * limit * r - b) * w;
    sqrt_rho1 = sqrt(fmax(DBL_MIN, limit / r) * w);
  } else {
    rho0 = HUBER ? sq * w : w * sq;
    sqrt_rho1 = sw;
  }
  a[0] += 0.5 * rho0;
#pragma unroll
  for (int r = 0; r < n; r++) {
    const double j0 = J[r * 3 + 0] * sqrt_rho1, j1 = J[r * 3 + 1] * sqrt_rho1, j2 = J[r * 3 + 2] * sqrt_rho1, fr = f[r] * sqrt_rho1;
    a[1] += j0 * fr; a[2] += j1 * fr; a[3] += j2 * fr;
    a[4] += j0 * j0; a[5] += j0 * j1; a[6] += j0 * j2; a[7] += j1 * j1; a[8] += j1 * j2; a[9] += j2 * j2;
  }
}

// the blocks [0, n) of one warp's segment, lane-strided
template <int COST, bool HUBER>
__device__ __forceinline__ void eval_loop_simple(const double* __restrict__ seg, size_t bstride, int n, double limit, double x0, double x1, double cy,
                                                 double sy, int lane, double* __restrict__ a) {
  for (int k = lane; k < n; k += 32, seg += RG_TILE) eval_block_simple<COST, HUBER>(seg, bstride, limit, x0, x1, cy, sy, a);
}

// Every warp owns the source cells [j_begin, j_end) — an equal share of the moving set — against EVERY fixed scan, and a segment of the
// block arrays with room for all of them (n_fixed * jc blocks at warp * n_fixed * jc).  The warps therefore carry equal shares of the
// association whatever the match rates of the individual keyframes, the warp that writes a block is the warp that evaluates it, and
// an evaluation is a lane-strided sweep over the warp's own contiguous segment.  (The order in which blocks are summed is fixed by
// this layout, not by the data: results are bit-identical from run to run.  The reference's block order — fixed-scan-major, source
// index ascending — is only needed for the residual vector of the evaluation mode, which is re-derived from sh.cnt.)
// Inside a segment the blocks lie in tiles of 32: tile t holds field f of blocks 32t .. 32t+31 at doubles [t * RG_TILE + 32 f, +32) — a lane
// that sweeps the segment reads all fields of its block at CONSTANT offsets from one pointer (no per-field stride arithmetic, nothing to
// keep in registers but the pointer), and a warp's accesses stay fully coalesced (32 consecutive doubles per field).
__device__ __forceinline__ size_t rg_segment_doubles(int n_fixed, int jc) { return (size_t)((n_fixed * jc + 31) >> 5) * RG_TILE; }
__device__ __forceinline__ double* rg_segment(const RegCtx& c, int warp) { return c.blocks + (size_t)warp * rg_segment_doubles(c.n_fixed, c.jc); }

// ---- association at pose x with search radius R (AddScanPairCost for every fixed scan): the search runs out of shared memory; per
// accepted slot ONE round trip to the matched target cell's fields (all loads issued together).  One barrier per round.
// Returns the number of residual blocks of the problem (the same value in every thread).
// COST >= 0: the cost function is fixed at compile time (specialised kernel, registration mode only); COST < 0: read from the parameters.
template <int COST>
__device__ __forceinline__ int rg_associate(RegShared& sh, const uint8_t* __restrict__ rg_stage, const double* x, double R) {
  const RegCtx& c = sh.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const int n_fixed = c.n_fixed;
  if (tid < n_fixed) {
    const double* fp = c.fixed_pose + (size_t)tid * 3;
    const Aff Ttar = vec_to_aff(fp[0], fp[1], fp[2]);
    sh.Ttar[tid] = Ttar;
    sh.Tst[tid] = aff_mul(aff_inv(Ttar), vec_to_aff(x[0], x[1], x[2]));
  }
  __syncthreads();
  const float Rm = (float)R + 1e-3f;
  const double R2 = R * R;
  const int weight_opt = c.P.weight_opt, cost = (COST >= 0) ? COST : c.P.cost, mode = (COST >= 0) ? (int)REG_MODE_REGISTER : c.mode;
  const bool weighted = weight_opt != TBV_W_UNIFORM, staged = c.staged != 0;
  const double angle_outlier = c.P.angle_outlier;
  const double* __restrict__ sf = c.sf;
  const size_t scap = c.scap;
  constexpr size_t bstride = 32;
  const double2* s_u = reinterpret_cast<const double2*>(rg_stage);
  const int jc = c.jc, n_src = c.n_src;
  const int j_begin = min(n_src, warp * jc), j_end = min(n_src, j_begin + jc);
  double* seg = rg_segment(c, warp);
  int my_cnt = 0;
  for (int fi = 0; fi < n_fixed; fi++) {
    const Aff Tst = sh.Tst[fi];
    const double* __restrict__ tf = sh.tgt[fi].f;
    const size_t tcap = (size_t)sh.tgt[fi].cap;
    const int cnt0 = my_cnt;
    for (int jb = j_begin; jb < j_end; jb += 32) {
      const int j = jb + lane;
      const bool live = j < j_end;
      double ux = 0.0, uy = 0.0;
      if (live) {
        if (staged) { const double2 u = s_u[j]; ux = u.x; uy = u.y; }
        else { ux = sf[(size_t)CF_U0 * scap + j]; uy = sf[(size_t)CF_U1 * scap + j]; }
      }
      const float qx = (float)((Tst.r00 * ux + Tst.r01 * uy) + Tst.tx), qy = (float)((Tst.r10 * ux + Tst.r11 * uy) + Tst.ty);
      int ti = -1;
      if (staged) {
        ti = nn_search_staged(reinterpret_cast<const float4*>(rg_stage + sh.ent_off[fi]), reinterpret_cast<const uint16_t*>(rg_stage + sh.row_off[fi]),
                              sh.grid[fi], qx, qy, Rm, R2, live);
      } else if (live) {
        ti = (COST >= 0) ? nn_search_cold(&sh.tgt[fi], &sh.grid[fi], sh.n_tgt[fi], qx, qy, R) : nn_search(sh.tgt[fi], sh.grid[fi], sh.n_tgt[fi], qx, qy, R);
      }
      bool ok = false;
      double w = 1.0, tn0 = 0.0, tn1 = 0.0, tu0 = 0.0, tu1 = 0.0;
      if (ti >= 0) {
        // every value the decision, the weight and the block need, in one round trip
        const double sn0 = sf[(size_t)CF_N0 * scap + j], sn1 = sf[(size_t)CF_N1 * scap + j];
        tn0 = tf[(size_t)CF_N0 * tcap + ti]; tn1 = tf[(size_t)CF_N1 * tcap + ti];
        tu0 = tf[(size_t)CF_U0 * tcap + ti]; tu1 = tf[(size_t)CF_U1 * tcap + ti];
        double N1 = 0.0, N2 = 0.0, p1 = 0.0, p2 = 0.0;
        if (weighted) {
          N1 = sf[(size_t)CF_NS * scap + j]; p1 = sf[(size_t)CF_SCALE * scap + j];
          N2 = tf[(size_t)CF_NS * tcap + ti]; p2 = tf[(size_t)CF_SCALE * tcap + ti];
        }
        const double snx = Tst.r00 * sn0 + Tst.r01 * sn1;
        const double sny = Tst.r10 * sn0 + Tst.r11 * sn1;
        const double sim = fmax(snx * tn0 + sny * tn1, 0.0);
        if (sim > angle_outlier) {
          ok = true;
          if (weighted) {
            const double simN = 2 * fmin(N1, N2) / (N1 + N2);
            const double simP = 2 * fmin(p1, p2) / (p1 + p2);
            switch (weight_opt) {   // registration.cpp:67-75
              case TBV_W_SIM_N: w = simN; break;
              case TBV_W_SIM_DIRECTION: w = sim; break;
              case TBV_W_SIM_SCALE: w = simP; break;
              case TBV_W_COMBINED: w = simN + sim + simP; break;
              default: w = 1.0;
            }
          }
        } else {
          ti = -1;
        }
      }
      if (mode == REG_MODE_EVAL && live) c.assoc[(size_t)fi * c.slot_cap + j] = ti;   // read back only by tbv_pair_normal_eq
      // the accepted correspondences of this warp go to its own segment in (fixed scan, source index) order
      const unsigned bal = __ballot_sync(FULL, ok);
      if (ok) {
        const int m = my_cnt + __popc(bal & ((1u << lane) - 1u));
        double* b = seg + (size_t)(m >> 5) * RG_TILE + (m & 31);
        const Aff Ttar = sh.Ttar[fi];
        b[0 * bstride] = ux;
        b[1 * bstride] = uy;
        b[2 * bstride] = (Ttar.r00 * tu0 + Ttar.r01 * tu1) + Ttar.tx;
        b[3 * bstride] = (Ttar.r10 * tu0 + Ttar.r11 * tu1) + Ttar.ty;
        if (cost == TBV_P2L) {
          b[4 * bstride] = Ttar.r00 * tn0 + Ttar.r01 * tn1;
          b[5 * bstride] = Ttar.r10 * tn0 + Ttar.r11 * tn1;
        } else if ((COST < 0 || COST == TBV_P2D) && cost == TBV_P2D) {  // n_scan_normal.cpp:288-298
          const double regularization = c.P.regularization, cov_scale = c.P.cov_scale;
          const double c00 = tf[(size_t)CF_C00 * tcap + ti], c01 = tf[(size_t)CF_C01 * tcap + ti];
          const double c10 = tf[(size_t)CF_C10 * tcap + ti], c11 = tf[(size_t)CF_C11 * tcap + ti];
          const double R00 = Ttar.r00, R01 = Ttar.r01, R10 = Ttar.r10, R11 = Ttar.r11;
          const double RC00 = R00 * c00 + R01 * c10, RC01 = R00 * c01 + R01 * c11;
          const double RC10 = R10 * c00 + R11 * c10, RC11 = R10 * c01 + R11 * c11;
          const double M00 = RC00 * R00 + RC01 * R01, M01 = RC00 * R10 + RC01 * R11;
          const double M10 = RC10 * R00 + RC11 * R01, M11 = RC10 * R10 + RC11 * R11;
          const double t00 = (regularization + M00) * cov_scale, t01 = (0.0 + M01) * cov_scale;
          const double t10 = (0.0 + M10) * cov_scale, t11 = (regularization + M11) * cov_scale;
          const double det = t00 * t11 - t10 * t01;
          const double invdet = 1.0 / det;
          const double i00 = t11 * invdet, i10 = -t10 * invdet, i11 = t00 * invdet;
          const double l00 = sqrt(i00);
          const double l10 = i10 / l00;
          const double l11 = sqrt(i11 - l10 * l10);
          b[4 * bstride] = l00;
          b[5 * bstride] = l10;
          b[6 * bstride] = l11;
        }
        b[7 * bstride] = w;
        b[8 * bstride] = sqrt(w);
      }
      my_cnt += __popc(bal);
    }
    if (lane == 0) sh.cnt[fi][warp] = my_cnt - cnt0;
  }
  if (lane == 0) sh.warp_cnt[warp] = my_cnt;
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int wv = 0; wv < RG_WARPS; wv++) total += (wv < (int)(blockDim.x >> 5)) ? sh.warp_cnt[wv] : 0;
  if (tid == 0) sh.n_blocks = total;   // read again by thread 0 only (result record)
  return total;
}

// The same association for the specialised kernels (registration mode): source-cell-major — the moving cell's own fields are fetched once
// and held while it is matched against every fixed scan (their latency overlaps the first search) — and the per-round transform
// Tst = Ttar^-1 * T(x) reuses the cos / sin of x[2] that warp 0 has just published in sh.cs for the evaluation at the same point
// (vec_to_aff(x) evaluates the same two functions on the same argument) and the inverse fixed poses of the set-up phase: no
// trigonometry between the barriers of a round.  Blocks of a warp's segment are in (source index, fixed scan) order here; the sums are
// the same fixed-shape reductions.
template <int COST>
__device__ __forceinline__ int rg_associate_fast(RegShared& sh, const uint8_t* __restrict__ rg_stage, const double* x, double R) {
  const RegCtx& c = sh.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const int n_fixed = c.n_fixed;
  if (tid < n_fixed) {
    Aff Tx;
    Tx.r00 = sh.cs[0]; Tx.r01 = -sh.cs[1]; Tx.r10 = sh.cs[1]; Tx.r11 = sh.cs[0]; Tx.tx = x[0]; Tx.ty = x[1];
    sh.Tst[tid] = aff_mul(sh.TtarInv[tid], Tx);
  }
  __syncthreads();
  const float Rm = (float)R + 1e-3f;
  const double R2 = R * R;
  const int weight_opt = c.P.weight_opt;
  const bool weighted = weight_opt != TBV_W_UNIFORM, staged = c.staged != 0;
  const double angle_outlier = c.P.angle_outlier;
  const double* __restrict__ sf = c.sf;
  const size_t scap = c.scap;
  constexpr size_t bstride = 32;
  const int jc = c.jc, n_src = c.n_src;
  const int j_begin = min(n_src, warp * jc), j_end = min(n_src, j_begin + jc);
  double* seg = rg_segment(c, warp);
  int my_cnt = 0;
  for (int jb = j_begin; jb < j_end; jb += 32) {
    const int j = jb + lane;
    const bool live = j < j_end;
    double ux = 0.0, uy = 0.0, sn0 = 0.0, sn1 = 0.0, N1 = 0.0, p1 = 0.0;
    if (live) {
      ux = sf[(size_t)CF_U0 * scap + j]; uy = sf[(size_t)CF_U1 * scap + j];
      sn0 = sf[(size_t)CF_N0 * scap + j]; sn1 = sf[(size_t)CF_N1 * scap + j];
      if (weighted) { N1 = sf[(size_t)CF_NS * scap + j]; p1 = sf[(size_t)CF_SCALE * scap + j]; }
    }
    for (int fi = 0; fi < n_fixed; fi++) {
      const Aff& Tst = sh.Tst[fi];
      const double* __restrict__ tf = sh.tgt[fi].f;
      const size_t tcap = (size_t)sh.tgt[fi].cap;
      const float qx = (float)((Tst.r00 * ux + Tst.r01 * uy) + Tst.tx), qy = (float)((Tst.r10 * ux + Tst.r11 * uy) + Tst.ty);
      int ti = -1;
      if (staged) {
        ti = nn_search_staged(reinterpret_cast<const float4*>(rg_stage + sh.ent_off[fi]), reinterpret_cast<const uint16_t*>(rg_stage + sh.row_off[fi]),
                              sh.grid[fi], qx, qy, Rm, R2, live);
      } else if (live) {
        ti = nn_search_cold(&sh.tgt[fi], &sh.grid[fi], sh.n_tgt[fi], qx, qy, R);
      }
      bool ok = false;
      double w = 1.0, tn0 = 0.0, tn1 = 0.0, tu0 = 0.0, tu1 = 0.0;
      if (ti >= 0) {
        // every value of the target cell the decision, the weight and the block need, in one round trip
        tn0 = tf[(size_t)CF_N0 * tcap + ti]; tn1 = tf[(size_t)CF_N1 * tcap + ti];
        tu0 = tf[(size_t)CF_U0 * tcap + ti]; tu1 = tf[(size_t)CF_U1 * tcap + ti];
        double N2 = 0.0, p2 = 0.0;
        if (weighted) { N2 = tf[(size_t)CF_NS * tcap + ti]; p2 = tf[(size_t)CF_SCALE * tcap + ti]; }
        const double snx = Tst.r00 * sn0 + Tst.r01 * sn1;
        const double sny = Tst.r10 * sn0 + Tst.r11 * sn1;
        const double sim = fmax(snx * tn0 + sny * tn1, 0.0);
        if (sim > angle_outlier) {
          ok = true;
          if (weighted) {
            const double simN = 2 * fmin(N1, N2) / (N1 + N2);
            const double simP = 2 * fmin(p1, p2) / (p1 + p2);
            switch (weight_opt) {   // registration.cpp:67-75
              case TBV_W_SIM_N: w = simN; break;
              case TBV_W_SIM_DIRECTION: w = sim; break;
              case TBV_W_SIM_SCALE: w = simP; break;
              case TBV_W_COMBINED: w = simN + sim + simP; break;
              default: w = 1.0;
            }
          }
        }
      }
      const unsigned bal = __ballot_sync(FULL, ok);
      if (ok) {
        const int m = my_cnt + __popc(bal & ((1u << lane) - 1u));
        double* b = seg + (size_t)(m >> 5) * RG_TILE + (m & 31);
        const Aff& Ttar = sh.Ttar[fi];
        b[0 * bstride] = ux;
        b[1 * bstride] = uy;
        b[2 * bstride] = (Ttar.r00 * tu0 + Ttar.r01 * tu1) + Ttar.tx;
        b[3 * bstride] = (Ttar.r10 * tu0 + Ttar.r11 * tu1) + Ttar.ty;
        if (COST == TBV_P2L) {
          b[4 * bstride] = Ttar.r00 * tn0 + Ttar.r01 * tn1;
          b[5 * bstride] = Ttar.r10 * tn0 + Ttar.r11 * tn1;
        } else if (COST == TBV_P2D) {  // n_scan_normal.cpp:288-298
          const double c00 = tf[(size_t)CF_C00 * tcap + ti], c01 = tf[(size_t)CF_C01 * tcap + ti];
          const double c10 = tf[(size_t)CF_C10 * tcap + ti], c11 = tf[(size_t)CF_C11 * tcap + ti];
          const double R00 = Ttar.r00, R01 = Ttar.r01, R10 = Ttar.r10, R11 = Ttar.r11;
          const double RC00 = R00 * c00 + R01 * c10, RC01 = R00 * c01 + R01 * c11;
          const double RC10 = R10 * c00 + R11 * c10, RC11 = R10 * c01 + R11 * c11;
          const double M00 = RC00 * R00 + RC01 * R01, M01 = RC00 * R10 + RC01 * R11;
          const double M10 = RC10 * R00 + RC11 * R01, M11 = RC10 * R10 + RC11 * R11;
          const double t00 = (c.P.regularization + M00) * c.P.cov_scale, t01 = (0.0 + M01) * c.P.cov_scale;
          const double t10 = (0.0 + M10) * c.P.cov_scale, t11 = (c.P.regularization + M11) * c.P.cov_scale;
          const double det = t00 * t11 - t10 * t01;
          const double invdet = 1.0 / det;
          const double i00 = t11 * invdet, i10 = -t10 * invdet, i11 = t00 * invdet;
          const double l00 = sqrt(i00);
          const double l10 = i10 / l00;
          const double l11 = sqrt(i11 - l10 * l10);
          b[4 * bstride] = l00;
          b[5 * bstride] = l10;
          b[6 * bstride] = l11;
        }
        b[7 * bstride] = w;
        b[8 * bstride] = sqrt(w);
      }
      my_cnt += __popc(bal);
    }
  }
  if (lane == 0) sh.warp_cnt[warp] = my_cnt;
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int wv = 0; wv < RG_WARPS; wv++) total += (wv < (int)(blockDim.x >> 5)) ? sh.warp_cnt[wv] : 0;
  if (tid == 0) sh.n_blocks = total;   // read again by thread 0 only (result record)
  return total;
}

// evaluation mode: block k of this warp's segment -> its index in the reference's block order (fixed-scan-major, source index ascending)
__device__ __forceinline__ int rg_reference_index(const RegShared& sh, int warp, int k) {
  const int n_fixed = sh.c.n_fixed;
  int fi = 0, cum = 0;
  while (fi < n_fixed - 1 && k >= cum + sh.cnt[fi][warp]) { cum += sh.cnt[fi][warp]; fi++; }
  int q = k - cum;
  for (int f = 0; f < fi; f++)
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) q += sh.cnt[f][wv];
  for (int wv = 0; wv < warp; wv++) q += sh.cnt[fi][wv];
  return q;
}

// ---- evaluation at sh.ex (cos/sin in sh.cs): per-warp partial sums in sh.warp_acc (fixed shapes: deterministic); ends with a barrier,
// warp 0 combines them afterwards.
template <int COST, int LOSS>
__device__ __forceinline__ void rg_evaluate(RegShared& sh, int write_residuals) {
  const RegCtx& c = sh.c;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  double a[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) a[i] = 0.0;
  const double x0 = sh.ex[0], x1 = sh.ex[1], cy = sh.cs[0], sy = sh.cs[1];
  const int n_mine = sh.warp_cnt[warp];
  const double* seg = rg_segment(c, warp) + lane;   // this lane's block of tile 0; its later blocks follow at RG_TILE doubles each
  constexpr size_t bstride = 32;
  const double limit = c.P.loss_limit;
  if (COST >= 0) {   // specialised kernel: one loop, nothing else compiled in
    eval_loop_simple<COST, LOSS == TBV_LOSS_HUBER>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
  } else {
    const int cost = c.P.cost, loss = c.P.loss;
    const bool simple_loss = (loss == TBV_LOSS_HUBER || loss == TBV_LOSS_NONE) && c.mode == REG_MODE_REGISTER;
    if (simple_loss) {
      if (loss == TBV_LOSS_HUBER) {
        if (cost == TBV_P2L) eval_loop_simple<TBV_P2L, true>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
        else if (cost == TBV_P2P) eval_loop_simple<TBV_P2P, true>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
        else eval_loop_simple<TBV_P2D, true>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
      } else {
        if (cost == TBV_P2L) eval_loop_simple<TBV_P2L, false>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
        else if (cost == TBV_P2P) eval_loop_simple<TBV_P2P, false>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
        else eval_loop_simple<TBV_P2D, false>(seg, bstride, n_mine, limit, x0, x1, cy, sy, lane, a);
      }
    } else {
      for (int k = lane; k < n_mine; k += 32, seg += RG_TILE) {
        double f[2], J[6];
        int n;
        a[0] += eval_block(cost, loss, limit, seg, bstride, x0, x1, cy, sy, f, J, n, true);
        for (int r = 0; r < n; r++) {
          const double j0 = J[r * 3 + 0], j1 = J[r * 3 + 1], j2 = J[r * 3 + 2], fr = f[r];
          a[1] += j0 * fr; a[2] += j1 * fr; a[3] += j2 * fr;
          a[4] += j0 * j0; a[5] += j0 * j1; a[6] += j0 * j2; a[7] += j1 * j1; a[8] += j1 * j2; a[9] += j2 * j2;
        }
        if (write_residuals && c.residuals) {
          const int q = rg_reference_index(sh, warp, k);
          for (int r = 0; r < n; r++) c.residuals[(size_t)q * n + r] = f[r];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    double v = a[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    if (lane == 0) sh.warp_acc[warp][i] = v;
  }
  __syncthreads();
}

// warp 0, after rg_evaluate(): cross-warp sums in warp order -> sh.acc (visible to the warp after __syncwarp)
__device__ __forceinline__ void rg_combine(RegShared& sh, int lane) {
  if (lane < NACC) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < RG_WARPS; wv++) v += (wv < (int)(blockDim.x >> 5)) ? sh.warp_acc[wv][lane] : 0.0;
    sh.acc[lane] = v;
  }
  __syncwarp();
}
// warp 0: publish the next evaluation point chosen by lane 0 (x in sh.ex); cos and sin are computed by two different lanes
__device__ __forceinline__ void rg_publish_eval_point(RegShared& sh, int lane, bool go) {
  const int g = __shfl_sync(0xffffffffu, go ? 1 : 0, 0);
  if (g) {
    const double th = sh.ex[2];
    if (lane == 1) sh.cs[0] = cos(th);
    if (lane == 2) sh.cs[1] = sin(th);
  }
}

// ---- set-up: the problem's constants, L2 prefetch of its working set, staging of what the search reads into shared memory
// STAGE_SRC: also stage the moving set's means (the fixed-scan-major association of the generic kernel re-reads them for every fixed scan;
// the source-major association of the specialised kernels reads them once from global memory and leaves the room to the L1 cache).
// TMA bulk copy global -> shared memory with completion on an mbarrier (cp.async.bulk, SASS UBLKCP): source, destination and size are
// multiples of 16 bytes.
__device__ __forceinline__ uint32_t rg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rg_bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(rg_smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(rg_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void rg_bar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RG_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RG_WAIT_DONE;\n"
      "bra RG_WAIT_LOOP;\n"
      "RG_WAIT_DONE:\n"
      "}\n" ::"r"(rg_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <bool STAGE_SRC>
__device__ __forceinline__ void rg_setup(RegShared& sh, uint8_t* __restrict__ rg_stage, int stage_bytes, const SetView* __restrict__ sets,
                                      const int* __restrict__ fixed_set, int fixed_first) {
  const RegCtx& c = sh.c;
  const int tid = threadIdx.x;
  const int n_fixed = c.n_fixed, n_src = c.n_src;
  if (tid < n_fixed) {
    sh.tgt[tid] = sets[fixed_set[fixed_first + tid]];
    sh.n_tgt[tid] = set_count(sh.tgt[tid]);
    CellGrid g;
    g.minx = g.miny = 0.f; g.nx = g.ny = 1; g.ok = 0;
    if (sh.tgt[tid].grid) g = *sh.tgt[tid].grid;
    sh.grid[tid] = g;
    const double* fp = c.fixed_pose + (size_t)tid * 3;   // the fixed poses do not change during the call
    const Aff Ttar = vec_to_aff(fp[0], fp[1], fp[2]);
    sh.Ttar[tid] = Ttar;
    sh.TtarInv[tid] = aff_inv(Ttar);
  }
  __syncthreads();
  // Pull the problem's read-only working set into L2 with full-line prefetches (all lines in flight at once): the cell fields
  // the association reads (mean, normal, N, planarity) of the moving set and of every fixed set, and the fixed sets' grid
  // entries.  The association itself touches them through short dependent chains (query -> bucket rows -> entries -> target
  // fields) and would otherwise pay a DRAM round trip at every link the first time.
  {
    const int pf_fields[6] = {CF_U0, CF_U1, CF_N0, CF_N1, CF_NS, CF_SCALE};
    for (int f = -1; f < n_fixed; f++) {
      const double* base = (f < 0) ? c.sf : sh.tgt[f].f;
      const size_t cap = (f < 0) ? c.scap : (size_t)sh.tgt[f].cap;
      const int cnt = (f < 0) ? n_src : sh.n_tgt[f];
      const int lines = (cnt * 8 + 127) / 128;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const char* ptr = reinterpret_cast<const char*>(base + (size_t)pf_fields[k] * cap);
        for (int l = tid; l < lines; l += blockDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (size_t)l * 128));
      }
      if (f >= 0 && sh.tgt[f].gent) {
        const char* ptr = reinterpret_cast<const char*>(sh.tgt[f].gent);
        const int elines = (cnt * 16 + 127) / 128;
        for (int l = tid; l < elines; l += blockDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (size_t)l * 128));
      }
    }
  }
  // Stage what the nearest-neighbour search reads into shared memory when it fits (it does for the odometry and loop workloads:
  // ~38 KB for 450-cell sets and four keyframes): the moving set's means and, per fixed set, its grid entries (x-sorted bucket rows)
  // + one start offset per bucket row.  The search then runs out of shared memory and the association pays ONE global round trip per
  // slot (the matched target's mean / normal / N / planarity) instead of one per link of query -> bucket rows -> entries -> target fields.
  if (tid == 0) {
    int need = STAGE_SRC ? n_src * 16 : 0, ok = 1;
    for (int f = 0; f < n_fixed; f++) {
      if (!(sh.tgt[f].grid && sh.grid[f].ok == 2)) { ok = 0; break; }
      sh.ent_off[f] = need; need += sh.n_tgt[f] * 16;
      sh.row_off[f] = need; need += ((sh.grid[f].ny + 1) * 2 + 15) & ~15;
    }
    sh.c.staged = (ok && need <= stage_bytes) ? 1 : 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(rg_smem_u32(&sh.stage_bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (c.staged) {
    // the grid entries of every fixed set: ONE bulk copy (TMA) per set, all in flight together, completion on one mbarrier
    uint32_t bulk_bytes = 0;
    for (int f = 0; f < n_fixed; f++) bulk_bytes += (uint32_t)sh.n_tgt[f] * 16u;
    if (tid == 0 && bulk_bytes > 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(rg_smem_u32(&sh.stage_bar)), "r"(bulk_bytes) : "memory");
      for (int f = 0; f < n_fixed; f++)
        if (sh.n_tgt[f] > 0) rg_bulk_load(rg_stage + sh.ent_off[f], sh.tgt[f].gent, (uint32_t)sh.n_tgt[f] * 16u, &sh.stage_bar);
    }
    double2* su = reinterpret_cast<double2*>(rg_stage);
    if (STAGE_SRC)
      for (int j = tid; j < n_src; j += blockDim.x) su[j] = make_double2(c.sf[(size_t)CF_U0 * c.scap + j], c.sf[(size_t)CF_U1 * c.scap + j]);
    for (int f = 0; f < n_fixed; f++) {
      uint16_t* rw = reinterpret_cast<uint16_t*>(rg_stage + sh.row_off[f]);
      const int nt = sh.n_tgt[f], ny = sh.grid[f].ny, nx = sh.grid[f].nx;
      const uint16_t* __restrict__ gstart = sh.tgt[f].gstart;
      for (int k = tid; k <= ny; k += blockDim.x) rw[k] = k < ny ? gstart[k * nx] : (uint16_t)nt;   // one offset per bucket row (a strided gather)
    }
    if (bulk_bytes > 0) rg_bar_wait(&sh.stage_bar, 0);
  }
  __syncthreads();
}

// MIN_CTAS: resident CTAs per SM the register budget is set for (4 -> 64 registers).
// COST / LOSS >= 0: registration mode with that cost function and loss compiled in (the odometry and loop-closure configuration, P2L +
// Huber, runs a kernel that carries no other cost function, loss, or the evaluation mode); < 0: everything, selected at run time.
// THREADS: CTA size (RG_THREADS, or RG_THREADS_SMALL for launches of single-fixed-scan problems — loop-closure candidates — where
// eight 4-warp CTAs per SM keep 1 184 problems resident at once instead of 592).
template <int THREADS, int MIN_CTAS, int COST, int LOSS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
k_register(int mode, int eval_itr, const SetView* __restrict__ sets, const RegProblem* __restrict__ problems, const int* __restrict__ fixed_set,
           const double* __restrict__ fixed_pose, int max_fixed, int slot_cap, RegParamsDev P, RegResult* __restrict__ results,
           double* __restrict__ eval_out, int* __restrict__ assoc_all, double* __restrict__ blocks_all, int* __restrict__ n_blocks_all,
           double* __restrict__ residuals_all, unsigned long long* __restrict__ dbg, int stage_bytes) {
  __shared__ RegShared sh;
  extern __shared__ __align__(16) uint8_t rg_stage[];
#ifdef TBV_DEV_TIMERS   // per-phase cycle counters of development builds; a release build carries none of their registers
  long long t_assoc = 0, t_eval = 0, t_lm = 0, t_mark = 0, t_s0 = 0, t_s1 = 0, t_s2 = 0, t_s3 = 0, t_sub = 0, n_rounds = 0, n_evals = 0;
  const long long t_start = clock64();
#define RG_TIMER_MARK() t_mark = clock64()
#define RG_TIMER_ADD(acc) { const long long t_now = clock64(); acc += t_now - t_mark; t_mark = t_now; t_sub = t_now; }
#define RG_TIMER_SUB(acc) { const long long t_now = clock64(); acc += t_now - t_sub; t_sub = t_now; }
#define RG_TIMER_COUNT(n) n++
#else
  (void)dbg;
#define RG_TIMER_MARK()
#define RG_TIMER_ADD(acc)
#define RG_TIMER_SUB(acc)
#define RG_TIMER_COUNT(n)
#endif
  const int p = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!problems[p].active) {
    if (tid == 0) {
      RegResult* out = results + p;
      out->pose[0] = problems[p].src_pose[0]; out->pose[1] = problems[p].src_pose[1]; out->pose[2] = problems[p].src_pose[2];
      out->align[0] = out->align[1] = out->align[2] = 0.0;
      out->pose_updated = 0; out->success = 0; out->itrs = 0; out->lm_iterations = 0; out->num_residuals = 0; out->last_n_iterations = 0;
      out->termination = 0; out->score = 0.0; out->final_cost = 0.0; out->last_relative_decrease = 0.0;
      if (n_blocks_all) n_blocks_all[p] = 0;
    }
    return;
  }
  if (tid == 0) {
    const RegProblem prob = problems[p];
    const SetView src = sets[prob.src_set];
    RegCtx& c = sh.c;
    c.P = P;
    c.sf = src.f; c.scap = (size_t)src.cap;
    c.n_src = set_count(src);
    c.n_fixed = min(prob.n_fixed, RG_MAX_FIXED);
    c.jc = (c.n_src + (THREADS / 32) - 1) / (THREADS / 32);
    c.bstride = (size_t)max_fixed * (slot_cap + RG_WARPS);   // room for every warp's segment: n_fixed * RG_WARPS * jc <= n_fixed * (n_src + RG_WARPS - 1)
    c.blocks = blocks_all + (size_t)p * BLK_FIELDS * (c.bstride + 32 * RG_WARPS);   // >= RG_WARPS segments rounded up to whole tiles
    c.assoc = assoc_all + (size_t)p * c.bstride;
    c.residuals = residuals_all ? residuals_all + (size_t)p * 2 * c.bstride : nullptr;
    c.fixed_pose = fixed_pose + (size_t)prob.fixed_first * 3;
    c.src_pose[0] = prob.src_pose[0]; c.src_pose[1] = prob.src_pose[1]; c.src_pose[2] = prob.src_pose[2];
    c.staged = 0; c.mode = mode; c.slot_cap = slot_cap;
    c.nres_per_block = (((COST >= 0) ? COST : P.cost) == TBV_P2L) ? 1 : 2;
  }
  __syncthreads();
  rg_setup<(COST < 0)>(sh, rg_stage, stage_bytes, sets, fixed_set, problems[p].fixed_first);

  // =========================================================================================================
  if (COST < 0 && mode == REG_MODE_EVAL) {
    const double R = (eval_itr == 1) ? 2 * P.radius : P.radius;
    const int nblk = rg_associate<COST>(sh, rg_stage, sh.c.src_pose, R);
    if (tid == 0) {
      sh.ex[0] = sh.c.src_pose[0]; sh.ex[1] = sh.c.src_pose[1]; sh.ex[2] = sh.c.src_pose[2];
      sh.cs[0] = cos(sh.c.src_pose[2]); sh.cs[1] = sin(sh.c.src_pose[2]);
    }
    __syncthreads();
    rg_evaluate<COST, LOSS>(sh, 1);
    if (warp == 0) {
      rg_combine(sh, lane);
      if (lane == 0) {
        RegResult* out = results + p;
        out->pose[0] = sh.c.src_pose[0]; out->pose[1] = sh.c.src_pose[1]; out->pose[2] = sh.c.src_pose[2];
        out->align[0] = out->align[1] = out->align[2] = 0.0;
        out->pose_updated = 0; out->itrs = 0; out->lm_iterations = 0; out->last_n_iterations = 0; out->termination = 0;
        out->last_relative_decrease = 0.0;
        out->num_residuals = nblk * sh.c.nres_per_block;
        out->success = out->num_residuals > 1;
        out->final_cost = sh.acc[0];
        out->score = sh.acc[0] / (double)max(nblk * sh.c.nres_per_block, 1);
        n_blocks_all[p] = nblk;
        if (eval_out)
          for (int i = 0; i < NACC; i++) eval_out[(size_t)p * NACC + i] = sh.acc[i];
      }
    }
    return;
  }

  // ---- Register (n_scan_normal.cpp:82-185) — the LM state AND the outer loop's state live in shared memory and are
  // advanced by lane 0 of warp 0, which also owns the cross-warp reduction: two barriers per LM iteration (partial sums
  // ready / decision published); the other threads only keep what the hot loops need in registers.
  LMState& S = sh.lm;
  OuterState& O = sh.outer;
  if (tid == 0) {
    for (int k = 0; k < 3; k++) { O.par[k] = sh.c.src_pose[k]; O.prev_par[k] = sh.c.src_pose[k]; O.tsrc[k] = sh.c.src_pose[k]; }
    O.prev_score = DBL_MAX;
    O.total_lm = 0; O.itr = 1; O.pose_updated = 0; O.success = 1;
    O.num_residuals = 0; O.last_n_iterations = 0; O.termination = 1;
    O.final_cost = 0.0; O.last_rel_dec = 0.0;
    O.go = (1 <= P.max_itr_association) ? 1 : 0;
  }
  __syncthreads();
  while (O.go) {
    const double R = (O.itr == 1) ? 2 * P.radius : P.radius;
    if (warp == 0) {   // evaluation point of the first evaluation = parameters.back()
      if (lane == 0) { sh.ex[0] = O.par[0]; sh.ex[1] = O.par[1]; sh.ex[2] = O.par[2]; }
      __syncwarp();
      rg_publish_eval_point(sh, lane, true);
      __syncwarp();   // sh.cs is read by the first lanes of this warp at the top of the association (rg_associate_fast)
    }
    RG_TIMER_MARK();
    const int nblk = (COST >= 0) ? rg_associate_fast<COST>(sh, rg_stage, O.par, R) : rg_associate<COST>(sh, rg_stage, O.par, R);  // ends with a barrier: sh.ex / sh.cs are visible too
    RG_TIMER_ADD(t_assoc);
    RG_TIMER_COUNT(n_rounds);
    const int nres = nblk * sh.c.nres_per_block;
    if (nres <= 1) {  // BuildOptimizationProblem fails (:371-374): Register returns false, itr_ not advanced
      if (tid == 0) { O.num_residuals = nres; O.success = 0; }
      break;
    }
    // ---- ceres::Solve
    bool first = true;
    for (;;) {
      rg_evaluate<COST, LOSS>(sh, 0);
      RG_TIMER_ADD(t_eval);
      RG_TIMER_COUNT(n_evals);
      if (warp == 0) {
        rg_combine(sh, lane);
        RG_TIMER_SUB(t_s0);
        bool go = false;
        if (lane == 0) {
          if (first) lm_begin(S, O.par, sh.acc); else lm_candidate(S, sh.acc);
          RG_TIMER_SUB(t_s1);
          go = lm_advance(S, P.max_itr_solver);
          RG_TIMER_SUB(t_s2);
          if (go) { sh.ex[0] = S.cand[0]; sh.ex[1] = S.cand[1]; sh.ex[2] = S.cand[2]; }
          else outer_advance(O, S, nres, P.max_itr_association);
          sh.flag = go ? 1 : 0;
        }
        __syncwarp();
        rg_publish_eval_point(sh, lane, go);
        RG_TIMER_SUB(t_s3);
      }
      first = false;
      __syncthreads();
      RG_TIMER_ADD(t_lm);
      if (!sh.flag) break;
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (O.success && O.pose_updated) { O.tsrc[0] = O.par[0]; O.tsrc[1] = O.par[1]; O.tsrc[2] = O.par[2]; }
    RegResult* out = results + p;
    out->pose[0] = O.tsrc[0]; out->pose[1] = O.tsrc[1]; out->pose[2] = O.tsrc[2];
    out->pose_updated = O.pose_updated;
    out->success = O.success ? 1 : 0;
    out->itrs = O.itr;
    out->lm_iterations = O.total_lm;
    out->num_residuals = O.num_residuals;
    out->last_n_iterations = O.last_n_iterations;
    out->termination = O.termination;
    out->final_cost = O.final_cost;
    out->last_relative_decrease = O.last_rel_dec;
    out->score = O.success ? O.final_cost / (double)O.num_residuals : 0.0;
    // Talign = Trevised^-1 * Tto (loopclosure.cpp:73), Trevised = vectorToAffine(parameters)
    const double* fp = sh.c.fixed_pose;
    const Aff Tal = aff_mul(aff_inv(vec_to_aff(O.tsrc[0], O.tsrc[1], O.tsrc[2])), vec_to_aff(fp[0], fp[1], fp[2]));
    out->align[0] = Tal.tx; out->align[1] = Tal.ty; out->align[2] = atan2(Tal.r10, Tal.r11);
    if (n_blocks_all) n_blocks_all[p] = sh.n_blocks;
#ifdef TBV_DEV_TIMERS
    if (dbg) {
      atomicAdd(&dbg[0], (unsigned long long)t_assoc); atomicAdd(&dbg[1], (unsigned long long)t_eval); atomicAdd(&dbg[2], (unsigned long long)t_lm);
      atomicAdd(&dbg[3], (unsigned long long)(clock64() - t_start)); atomicAdd(&dbg[4], 1ull);
      atomicAdd(&dbg[5], (unsigned long long)t_s0); atomicAdd(&dbg[6], (unsigned long long)t_s1); atomicAdd(&dbg[7], (unsigned long long)t_s2); atomicAdd(&dbg[8], (unsigned long long)t_s3);
      atomicAdd(&dbg[9], (unsigned long long)n_rounds); atomicAdd(&dbg[10], (unsigned long long)n_evals);
      if (p < 2048) { dbg[16 + 4 * p] = (unsigned long long)(clock64() - t_start); dbg[17 + 4 * p] = (unsigned long long)n_evals; dbg[18 + 4 * p] = (unsigned long long)sh.c.n_src; dbg[19 + 4 * p] = (unsigned long long)sh.n_blocks; }
    }
#endif
  }
}

// --------------------------------------------------------------------------------------------------------------
RegParamsDev to_dev(const tbv_reg_params& p) {
  RegParamsDev d;
  d.cost = p.cost; d.loss = p.loss; d.weight_opt = p.weight_opt; d.loss_limit = p.loss_limit; d.cov_scale = p.cov_scale;
  d.regularization = p.regularization;
  const bool both = p.max_itr_association > 0 && p.max_itr_solver > 0;            // SetParameters sets both (n_scan_normal.h:53-55)
  d.max_itr_association = both ? p.max_itr_association : 8;                       // n_scan_normal.h:75
  d.max_itr_solver = both ? p.max_itr_solver : 20;                                // n_scan_normal.cpp:9
  d.angle_outlier = std::cos(M_PI / 6.0);                                         // n_scan_normal.cpp:216
  d.radius = 2.0;                                                                 // registration.h:122
  return d;
}

RegScratch* reg_scratch(tbv_ctx* ctx) {
  if (!ctx->reg_scratch) ctx->reg_scratch = new RegScratch();
  return (RegScratch*)ctx->reg_scratch;
}
uint64_t reg_fingerprint(tbv_ctx* ctx) {
  if (!ctx->reg_scratch) return 0;
  const RegScratch& S = *(RegScratch*)ctx->reg_scratch;
  uint64_t h = 2;
  for (const void* p : {(const void*)S.assoc.p, (const void*)S.blocks.p, (const void*)S.n_blocks.p, (const void*)S.residuals.p, (const void*)S.dbg.p})
    h = fp_mix(h, p);
  return h;
}
void reg_release(tbv_ctx* ctx) {
  if (!ctx->reg_scratch) return;
  RegScratch* s = (RegScratch*)ctx->reg_scratch;
  s->release();
  delete s;
  ctx->reg_scratch = nullptr;
}
int reg_scratch_reserve(tbv_ctx* ctx, int n_problems, int max_fixed, int slot_cap, bool want_residuals) {
  RegScratch& S = *reg_scratch(ctx);
  const size_t slots = (size_t)n_problems * max_fixed * (slot_cap + RG_WARPS);   // the kernel's bstride per problem
  int rc;
  if ((rc = S.assoc.reserve(slots)) || (rc = S.blocks.reserve((slots + (size_t)n_problems * 32 * RG_WARPS) * BLK_FIELDS)) || (rc = S.n_blocks.reserve(n_problems)))
    return rc;
  if (want_residuals && (rc = S.residuals.reserve(slots * 2))) return rc;
  return TBV_OK;
}

int register_launch(tbv_ctx* ctx, int mode, int eval_itr, const SetView* sets_dev, const RegProblem* problems_dev, const int* fixed_set_dev,
                    const double* fixed_pose_dev, int n_problems, int max_fixed, int slot_cap, int tgt_cap, const RegParamsDev& params,
                    RegResult* results_dev, double* eval_out_dev, bool want_residuals) {
  if (n_problems <= 0) return TBV_OK;
  TBV_REQUIRE(max_fixed >= 1 && slot_cap >= 1 && tgt_cap >= 1, "bad registration capacities");
  int rc = reg_scratch_reserve(ctx, n_problems, max_fixed, slot_cap, want_residuals);
  if (rc) return rc;
  RegScratch& S = *reg_scratch(ctx);
  TBV_REQUIRE(max_fixed <= RG_MAX_FIXED, "too many fixed scans per problem (at most 16)");
  // dynamic shared memory for the staged working set: 4 CTAs x (48 KB + 4.2 KB static + 1 KB reserved) fit one SM's 228 KB
  constexpr int RG_STAGE = 48 * 1024;
  unsigned long long* dbg = nullptr;  // per-phase cycle counters: development builds only (-DTBV_DEV_TIMERS)
#ifdef TBV_DEV_TIMERS
  if (!S.dbg.p) { if ((rc = S.dbg.reserve(16 + 4 * 2048))) return rc; }
  dbg = S.dbg.p;
  TBV_CUDA(cudaMemsetAsync(dbg, 0, (16 + 4 * 2048) * sizeof(unsigned long long), ctx->stream));
#endif
  // the configuration every caller on the hot path uses (odometry and loop closure: P2L, Huber) has its own kernel; it stages the fixed
  // scans' grids only (42 KB: 4 CTAs x (42 KB + 5.2 KB static + 1 KB reserved) stay under the 196 KB shared-memory configuration, which
  // leaves 60 KB of L1 to the stack lines instead of 28 KB)
  auto kernel = k_register<RG_THREADS, 4, -1, -1>;
  int stage = RG_STAGE, threads = RG_THREADS;
  if (mode == REG_MODE_REGISTER && params.cost == TBV_P2L && params.loss == TBV_LOSS_HUBER) {
    kernel = k_register<RG_THREADS, 4, TBV_P2L, TBV_LOSS_HUBER>; stage = 42 * 1024;
    // single-fixed-scan problems (loop-closure candidates) in numbers that would not be resident at 4 CTAs per SM: 4-warp CTAs, 8 per SM
    // (8 x (20 KB + 5.2 KB + 1 KB) of shared memory, 8 x 128 x 64 registers), one wave for up to 8 problems per SM
    if (max_fixed == 1 && n_problems > 4 * ctx->sm_count && (size_t)tgt_cap * 16 + 4096 <= 20 * 1024) {
      kernel = k_register<RG_THREADS_SMALL, 8, TBV_P2L, TBV_LOSS_HUBER>; stage = 20 * 1024; threads = RG_THREADS_SMALL;
    }
  }
  if ((rc = ensure_dyn_smem(ctx, kernel, stage))) return rc;
  kernel<<<n_problems, threads, stage, ctx->stream>>>(mode, eval_itr, sets_dev, problems_dev, fixed_set_dev, fixed_pose_dev, max_fixed, slot_cap, params,
                                                     results_dev, eval_out_dev, S.assoc.p, S.blocks.p, S.n_blocks.p,
                                                     want_residuals ? S.residuals.p : nullptr, dbg, stage);
  launched(ctx, "k_register");
  TBV_CUDA(cudaGetLastError());
#ifdef TBV_DEV_TIMERS
  {  // mean cycles per problem spent in association / evaluation / LM + barrier
    static unsigned long long h[16 + 4 * 2048];
    TBV_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    TBV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[4]) fprintf(stderr, "k_register cycles/problem: associate %.0f  evaluate %.0f  lm+barrier %.0f  total %.0f | combine %.0f begin/candidate %.0f advance %.0f publish %.0f | rounds %.2f evals %.2f\n", (double)h[0] / h[4], (double)h[1] / h[4],
                      (double)h[2] / h[4], (double)h[3] / h[4], (double)h[5] / h[4], (double)h[6] / h[4], (double)h[7] / h[4], (double)h[8] / h[4], (double)h[9] / h[4], (double)h[10] / h[4]);
    if (h[4] && n_problems <= 2048) {   // distribution over the problems of the launch: who is the slowest, and why
      std::vector<int> idx(n_problems);
      for (int i = 0; i < n_problems; i++) idx[i] = i;
      std::sort(idx.begin(), idx.end(), [&](int a, int b) { return h[16 + 4 * a] < h[16 + 4 * b]; });
      auto row = [&](int i) { fprintf(stderr, " [cyc %llu evals %llu n_src %llu blocks %llu]", h[16 + 4 * i], h[17 + 4 * i], h[18 + 4 * i], h[19 + 4 * i]); };
      fprintf(stderr, "k_register distribution: min"); row(idx[0]);
      fprintf(stderr, " p50"); row(idx[n_problems / 2]);
      fprintf(stderr, " p90"); row(idx[n_problems * 9 / 10]);
      fprintf(stderr, " p99"); row(idx[n_problems * 99 / 100]);
      fprintf(stderr, " max"); row(idx[n_problems - 1]);
      fprintf(stderr, "\n");
    }
  }
#endif
  return TBV_OK;
}

int cellgrid_build_launch(tbv_ctx* ctx, const SetView* sets_dev, const int* which_dev, int n_launch, int n_sets, int cell_cap, double max_extent) {
  if (n_launch <= 0) return TBV_OK;
  // shared-memory counters for as many buckets as a grid over cells within max_extent of the sensor can have (<= GRID_CAP): for the
  // odometry's 181 m that is 33 KB instead of 64 KB, so all 592 CTAs are resident at once instead of running as 1.33 waves
  int bucket_cap = GRID_CAP;
  if (max_extent > 0) {
    const long long side = (long long)(2.0 * max_extent / GRID_CELL) + 2;
    if (side * side < GRID_CAP) bucket_cap = (int)(side * side);
  }
  bucket_cap = (bucket_cap + 3) & ~3;   // the entry staging area behind the counters stays 16-byte aligned
  // + room to order the entries of a set of up to cell_cap cells inside their buckets (x-sorted bucket rows: the staged search of
  // k_register needs them); a set that does not fit keeps bucket order only and is searched from global memory
  const size_t cap_bytes = (size_t)ctx->smem_optin_max > 2048 ? (size_t)ctx->smem_optin_max - 2048 : 0;   // static shared memory of the kernel + reserve
  int sort_cap = cell_cap > 0 ? cell_cap : 0;
  if ((size_t)bucket_cap * sizeof(int) + (size_t)sort_cap * sizeof(float4) > cap_bytes)
    sort_cap = cap_bytes > (size_t)bucket_cap * sizeof(int) ? (int)((cap_bytes - (size_t)bucket_cap * sizeof(int)) / sizeof(float4)) : 0;
  const size_t smem = (size_t)bucket_cap * sizeof(int) + (size_t)sort_cap * sizeof(float4);
  {
    const int rc = ensure_dyn_smem(ctx, k_cellgrid_build, smem);
    if (rc) return rc;
  }
  k_cellgrid_build<<<n_launch, 256, smem, ctx->stream>>>(sets_dev, which_dev, n_sets, bucket_cap, sort_cap);
  launched(ctx, "k_cellgrid_build");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

int GridStore::reserve(int n_sets_, int cell_cap_) {
  n_sets = n_sets_;
  cell_cap = cell_cap_;
  int rc;
  if ((rc = hdr.reserve(n_sets_)) || (rc = start.reserve((size_t)n_sets_ * (GRID_CAP + 1))) || (rc = ent.reserve((size_t)n_sets_ * cell_cap_)))
    return rc;
  return TBV_OK;
}

// --------------------------------------------------------------------------------------------------------------
// C-ABI entry points with host-resident cell records
// --------------------------------------------------------------------------------------------------------------
namespace {
struct HostProblemSet {  // uploads n_sets cell sets, describes them as SetViews and builds their search grids
  tbv_ctx* ctx;
  std::vector<DevBuf<double>> bufs;
  DevBuf<SetView> views;
  GridStore grids;
  int max_n = 1;
  ~HostProblemSet() {
    for (auto& b : bufs) b.release();
    views.release();
    grids.release();
  }
  int upload(int n_sets, const tbv_cell* const* sets, const int* n_cells) {
    bufs.resize(n_sets);
    std::vector<SetView> hv(n_sets);
    for (int i = 0; i < n_sets; i++) {
      TBV_REQUIRE(n_cells[i] >= 0 && (n_cells[i] == 0 || sets[i]), "bad cell set");
      if (n_cells[i] > max_n) max_n = n_cells[i];
    }
    int rc = grids.reserve(n_sets, max_n);
    if (rc) return rc;
    for (int i = 0; i < n_sets; i++) {
      const int n = n_cells[i];
      const int cap = n > 0 ? n : 1;
      if ((rc = bufs[i].reserve((size_t)CELL_FIELDS * cap))) return rc;
      if ((rc = cells_upload(ctx, sets[i], n, bufs[i].p, cap))) return rc;
      hv[i] = grids.view(i, bufs[i].p, cap, nullptr, n);
    }
    if ((rc = views.reserve(n_sets))) return rc;
    TBV_CUDA(cudaMemcpyAsync(views.p, hv.data(), n_sets * sizeof(SetView), cudaMemcpyHostToDevice, ctx->stream));
    TBV_CUDA(cudaStreamSynchronize(ctx->stream));  // hv goes out of scope
    return cellgrid_build_launch(ctx, views.p, nullptr, n_sets, n_sets, max_n);
  }
};

template <typename T>
int to_device(tbv_ctx* ctx, DevBuf<T>& d, const std::vector<T>& h) {
  int rc = d.reserve(h.size() ? h.size() : 1);
  if (rc) return rc;
  if (!h.empty()) TBV_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return TBV_OK;
}

void fill_summary(const RegResult& r, tbv_reg_summary* s) {
  if (!s) return;
  s->success = r.success; s->itrs = r.itrs; s->lm_iterations = r.lm_iterations; s->num_residuals = r.num_residuals;
  s->last_n_iterations = r.last_n_iterations; s->termination = r.termination; s->score = r.score; s->final_cost = r.final_cost;
  s->last_relative_decrease = r.last_relative_decrease;
}

// one problem: scans[0..n-2] fixed, scans[n-1] moving
int run_single(tbv_ctx* ctx, int mode, int eval_itr, int n_scans, const tbv_cell* const* scans, const int* n_cells, const double* T,
               const tbv_reg_params* params, RegResult* res, double* eval_out, std::vector<int>* assoc, std::vector<double>* residuals,
               int* n_blocks) {
  HostProblemSet hs;
  hs.ctx = ctx;
  int rc = hs.upload(n_scans, scans, n_cells);
  if (rc) return rc;
  const int n_fixed = n_scans - 1;
  RegProblem pr;
  pr.n_fixed = n_fixed; pr.fixed_first = 0; pr.src_set = n_scans - 1; pr.active = 1;
  for (int c = 0; c < 3; c++) pr.src_pose[c] = T[3 * (n_scans - 1) + c];
  std::vector<RegProblem> hp{pr};
  std::vector<int> hfs(n_fixed);
  std::vector<double> hfp((size_t)n_fixed * 3);
  for (int i = 0; i < n_fixed; i++) {
    hfs[i] = i;
    for (int c = 0; c < 3; c++) hfp[3 * i + c] = T[3 * i + c];
  }
  DevBuf<RegProblem> dp; DevBuf<int> dfs; DevBuf<double> dfp, dev_eval; DevBuf<RegResult> dr;
  auto cleanup = [&]() { dp.release(); dfs.release(); dfp.release(); dev_eval.release(); dr.release(); };
  if ((rc = to_device(ctx, dp, hp)) || (rc = to_device(ctx, dfs, hfs)) || (rc = to_device(ctx, dfp, hfp)) || (rc = dr.reserve(1)) ||
      (rc = dev_eval.reserve(NACC))) { cleanup(); return rc; }
  const int slot_cap = n_cells[n_scans - 1] > 0 ? n_cells[n_scans - 1] : 1;
  rc = register_launch(ctx, mode, eval_itr, hs.views.p, dp.p, dfs.p, dfp.p, 1, n_fixed, slot_cap, hs.max_n, to_dev(*params), dr.p, dev_eval.p,
                       residuals != nullptr);
  if (rc) { cleanup(); return rc; }
  cudaError_t e = cudaMemcpyAsync(res, dr.p, sizeof(RegResult), cudaMemcpyDeviceToHost, ctx->stream);
  double ev[NACC];
  if (e == cudaSuccess && eval_out) e = cudaMemcpyAsync(ev, dev_eval.p, sizeof(ev), cudaMemcpyDeviceToHost, ctx->stream);
  RegScratch& S = *reg_scratch(ctx);
  int nb = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&nb, S.n_blocks.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && assoc) {
    assoc->assign((size_t)n_fixed * slot_cap, -1);
    e = cudaMemcpyAsync(assoc->data(), S.assoc.p, assoc->size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && residuals) {
    const int nres = nb * (params->cost == TBV_P2L ? 1 : 2);
    residuals->assign(nres, 0.0);
    if (nres > 0) e = cudaMemcpy(residuals->data(), S.residuals.p, (size_t)nres * sizeof(double), cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (e != cudaSuccess) { set_error("registration: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  if (eval_out) memcpy(eval_out, ev, sizeof(ev));
  if (n_blocks) *n_blocks = nb;
  return TBV_OK;
}
}  // namespace

}  // namespace tbv

using namespace tbv;

extern "C" {

int tbv_pair_normal_eq(tbv_ctx* ctx, const tbv_cell* tgt, int n_tgt, const double T_tgt[3], const tbv_cell* src, int n_src,
                       const double T_src[3], const tbv_reg_params* params, int itr, double* cost, int* n_res, double H[9], double g[3],
                       int32_t* assoc) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && T_tgt && T_src && params && n_tgt >= 0 && n_src >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  const tbv_cell* scans[2] = {tgt, src};
  const int n_cells[2] = {n_tgt, n_src};
  const double T[6] = {T_tgt[0], T_tgt[1], T_tgt[2], T_src[0], T_src[1], T_src[2]};
  RegResult r;
  double ev[NACC];
  std::vector<int> a;
  int rc = run_single(ctx, REG_MODE_EVAL, itr, 2, scans, n_cells, T, params, &r, ev, assoc ? &a : nullptr, nullptr, nullptr);
  if (rc) return rc;
  if (cost) *cost = ev[0];
  if (n_res) *n_res = r.num_residuals;
  if (g) { g[0] = ev[1]; g[1] = ev[2]; g[2] = ev[3]; }
  if (H) {
    H[0] = ev[4]; H[1] = ev[5]; H[2] = ev[6];
    H[3] = ev[5]; H[4] = ev[7]; H[5] = ev[8];
    H[6] = ev[6]; H[7] = ev[8]; H[8] = ev[9];
  }
  if (assoc) for (int j = 0; j < n_src; j++) assoc[j] = a[j];
  return TBV_OK;
}

int tbv_register(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, double* T, const tbv_reg_params* params,
                 tbv_reg_summary* summary) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && scans && n_cells && T && params && n_scans >= 2, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  RegResult r;
  int rc = run_single(ctx, REG_MODE_REGISTER, 0, n_scans, scans, n_cells, T, params, &r, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  fill_summary(r, summary);
  // Register writes Tsrc[i] = vectorToAffine(parameters[i]) for every scan after a successful solve; read back through
  // Affine3dToVectorXYeZ the angles come out as atan2(sin, cos) (registration.cpp:129-135, utils.cpp:115-122)
  if (r.pose_updated) {
    for (int i = 0; i < n_scans; i++) {
      double* t = T + 3 * i;
      const double th = (i == n_scans - 1) ? r.pose[2] : t[2];
      if (i == n_scans - 1) { t[0] = r.pose[0]; t[1] = r.pose[1]; }
      t[2] = std::atan2(std::sin(th), std::cos(th));
    }
  }
  return TBV_OK;
}

int tbv_get_cost(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, const double* T, const tbv_reg_params* params,
                 int itr, double* score, double* cost, int* n_res, double* residuals, int res_capacity) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && scans && n_cells && T && params && n_scans >= 2, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  RegResult r;
  double ev[NACC];
  std::vector<double> res;
  int rc = run_single(ctx, REG_MODE_EVAL, itr, n_scans, scans, n_cells, T, params, &r, ev, nullptr, residuals ? &res : nullptr, nullptr);
  if (rc) return rc;
  if (n_res) *n_res = r.num_residuals > 1 ? r.num_residuals : -1;  // GetCost returns false when <= 1 residual
  if (cost) *cost = ev[0];
  if (score) *score = r.score;
  if (residuals)
    for (int i = 0; i < (int)res.size() && i < res_capacity; i++) residuals[i] = res[i];
  return TBV_OK;
}

// OdometryKeyframeFuser::approximateCovarianceBySampling, sampling half (odometrykeyframefuser.cpp:261-313): the n^3 GetCost
// evaluations around T.back() are n^3 independent problems over the same cell sets -> ONE launch of k_register in evaluation mode.
int tbv_cost_samples(tbv_ctx* ctx, int n_scans, const tbv_cell* const* scans, const int* n_cells, const double* T, const tbv_reg_params* params,
                     int itr, double xy_range, double yaw_range, int n_per_axis, double* samples) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && scans && n_cells && T && params && samples && n_scans >= 2 && n_per_axis >= 1 && n_per_axis <= 16, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  auto linspace = [](double start, double end, int num) {  // odometrykeyframefuser.cpp:498-524
    std::vector<double> v;
    if (num == 1) { v.push_back(start); return v; }
    const double delta = (end - start) / ((double)num - 1);
    for (int i = 0; i < num - 1; ++i) v.push_back(start + delta * i);
    v.push_back(end);
    return v;
  };
  const std::vector<double> xy = linspace(-xy_range * 0.5, xy_range * 0.5, n_per_axis);
  const std::vector<double> th = linspace(-yaw_range * 0.5, yaw_range * 0.5, n_per_axis);
  const int n_fixed = n_scans - 1, n_samples = n_per_axis * n_per_axis * n_per_axis;
  HostProblemSet hs;
  hs.ctx = ctx;
  int rc = hs.upload(n_scans, scans, n_cells);
  if (rc) return rc;
  const double* Tb = T + 3 * (n_scans - 1);
  const double cb = std::cos(Tb[2]), sb = std::sin(Tb[2]);  // T_best_guess.linear() (vectorToAffine3d, registration.cpp:129-135)
  std::vector<RegProblem> hp(n_samples);
  int q = 0;
  for (int it = 0; it < n_per_axis; it++)
    for (int ix = 0; ix < n_per_axis; ix++)
      for (int iy = 0; iy < n_per_axis; iy++, q++) {
        // sample_T.linear() = AngleAxis(dtheta, z) * T_best.linear(); GetCost reads the yaw back as atan2(R10, R11) (utils.cpp:115-122)
        const double c = std::cos(th[it]), sn = std::sin(th[it]);
        const double r10 = sn * cb + c * sb, r11 = sn * (-sb) + c * cb;
        RegProblem& pr = hp[q];
        pr.n_fixed = n_fixed; pr.fixed_first = 0; pr.src_set = n_scans - 1; pr.active = 1;
        pr.src_pose[0] = xy[ix] + Tb[0]; pr.src_pose[1] = xy[iy] + Tb[1]; pr.src_pose[2] = std::atan2(r10, r11);
        samples[4 * q + 0] = xy[ix]; samples[4 * q + 1] = xy[iy]; samples[4 * q + 2] = th[it]; samples[4 * q + 3] = 0.0;
      }
  std::vector<int> hfs(n_fixed);
  std::vector<double> hfp((size_t)n_fixed * 3);
  for (int i = 0; i < n_fixed; i++) {
    hfs[i] = i;
    for (int c = 0; c < 3; c++) hfp[3 * i + c] = T[3 * i + c];
  }
  DevBuf<RegProblem> dp; DevBuf<int> dfs; DevBuf<double> dfp, dev_eval; DevBuf<RegResult> dr;
  auto cleanup = [&]() { dp.release(); dfs.release(); dfp.release(); dev_eval.release(); dr.release(); };
  if ((rc = to_device(ctx, dp, hp)) || (rc = to_device(ctx, dfs, hfs)) || (rc = to_device(ctx, dfp, hfp)) || (rc = dr.reserve(n_samples)) ||
      (rc = dev_eval.reserve((size_t)n_samples * NACC))) { cleanup(); return rc; }
  const int slot_cap = n_cells[n_scans - 1] > 0 ? n_cells[n_scans - 1] : 1;
  rc = register_launch(ctx, REG_MODE_EVAL, itr, hs.views.p, dp.p, dfs.p, dfp.p, n_samples, n_fixed, slot_cap, hs.max_n, to_dev(*params), dr.p,
                       dev_eval.p, false);
  if (rc) { cleanup(); return rc; }
  std::vector<RegResult> hr(n_samples);
  std::vector<double> ev((size_t)n_samples * NACC);
  cudaError_t e = cudaMemcpyAsync(hr.data(), dr.p, n_samples * sizeof(RegResult), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ev.data(), dev_eval.p, ev.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_cost_samples: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  double sample_cost = 0.0;  // a failed GetCost (<= 1 residual) leaves the previous sample's cost in place (:283,304)
  for (int i = 0; i < n_samples; i++) {
    if (hr[i].num_residuals > 1) sample_cost = ev[(size_t)i * NACC];
    samples[4 * i + 3] = sample_cost;
  }
  return TBV_OK;
}

// approximateCovarianceBySampling, fitting half (odometrykeyframefuser.cpp:315-378): least-squares quadric
// f = a x^2 + b y^2 + c z^2 + d xy + e yz + f zx + g x + h y + i z + j through the samples (Householder QR here, Eigen bdcSvd there:
// the same minimiser for a full-rank design), Hessian, convexity test on its eigenvalues, covariance = 2 H^-1 * score_scale * scaler
// embedded in the reference's 6x6 layout (x, y, -, -, -, yaw).  *convex = 0 (and rc TBV_OK) when the fit is not convex or singular.
int tbv_cov_from_cost_samples(const double* samples, int n_samples, double score_scale, double cov_scaler, double cov6x6[36], int* convex) {
  TBV_REQUIRE(samples && cov6x6 && convex && n_samples >= 10, "bad arguments");
  *convex = 0;
  const int m = n_samples, nc = 10;
  std::vector<double> A((size_t)m * nc), b(m);
  for (int i = 0; i < m; i++) {
    const double x = samples[4 * i], y = samples[4 * i + 1], z = samples[4 * i + 2];
    double* r = &A[(size_t)i * nc];
    r[0] = x * x; r[1] = y * y; r[2] = z * z; r[3] = x * y; r[4] = y * z; r[5] = z * x; r[6] = x; r[7] = y; r[8] = z; r[9] = 1.0;
    b[i] = samples[4 * i + 3];
  }
  // columns differ by many orders of magnitude (yaw steps are ~1e-3 rad): scale them to unit norm before the factorisation
  double cs[10];
  for (int c = 0; c < nc; c++) {
    double s2 = 0;
    for (int i = 0; i < m; i++) s2 += A[(size_t)i * nc + c] * A[(size_t)i * nc + c];
    cs[c] = s2 > 0 ? std::sqrt(s2) : 1.0;
    for (int i = 0; i < m; i++) A[(size_t)i * nc + c] /= cs[c];
  }
  for (int c = 0; c < nc; c++) {  // Householder QR, applied to b on the fly
    double norm = 0;
    for (int i = c; i < m; i++) norm += A[(size_t)i * nc + c] * A[(size_t)i * nc + c];
    norm = std::sqrt(norm);
    if (norm < 1e-300) return TBV_OK;  // rank deficient sampling pattern
    const double alpha = A[(size_t)c * nc + c] > 0 ? -norm : norm;
    std::vector<double> v(m, 0.0);
    for (int i = c; i < m; i++) v[i] = A[(size_t)i * nc + c];
    v[c] -= alpha;
    double vv = 0;
    for (int i = c; i < m; i++) vv += v[i] * v[i];
    if (vv < 1e-300) continue;
    for (int k = c; k < nc; k++) {
      double d = 0;
      for (int i = c; i < m; i++) d += v[i] * A[(size_t)i * nc + k];
      d = 2.0 * d / vv;
      for (int i = c; i < m; i++) A[(size_t)i * nc + k] -= d * v[i];
    }
    double d = 0;
    for (int i = c; i < m; i++) d += v[i] * b[i];
    d = 2.0 * d / vv;
    for (int i = c; i < m; i++) b[i] -= d * v[i];
  }
  double coef[10];
  for (int c = nc - 1; c >= 0; c--) {
    double s = b[c];
    for (int k = c + 1; k < nc; k++) s -= A[(size_t)c * nc + k] * coef[k];
    const double d = A[(size_t)c * nc + c];
    if (std::fabs(d) < 1e-14) return TBV_OK;
    coef[c] = s / d;
  }
  for (int c = 0; c < nc; c++) coef[c] /= cs[c];
  double H[3][3] = {{2 * coef[0], coef[3], coef[5]}, {coef[3], 2 * coef[1], coef[4]}, {coef[5], coef[4], 2 * coef[2]}};
  // eigenvalues of the symmetric 3x3 by cyclic Jacobi (SelfAdjointEigenSolver in the reference; only the signs are used).  The yaw
  // axis is rescaled first so that the rotations see comparable entries; congruence scaling preserves the signs (Sylvester).
  double Sx[3];
  for (int i = 0; i < 3; i++) Sx[i] = std::fabs(H[i][i]) > 0 ? 1.0 / std::sqrt(std::fabs(H[i][i])) : 1.0;
  double M[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M[i][j] = H[i][j] * Sx[i] * Sx[j];
  for (int sweep = 0; sweep < 50; sweep++) {
    const double off = std::fabs(M[0][1]) + std::fabs(M[0][2]) + std::fabs(M[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int qq = p + 1; qq < 3; qq++) {
        if (M[p][qq] == 0.0) continue;
        const double theta = (M[qq][qq] - M[p][p]) / (2.0 * M[p][qq]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) { const double a = M[k][p], bq = M[k][qq]; M[k][p] = c * a - sn * bq; M[k][qq] = sn * a + c * bq; }
        for (int k = 0; k < 3; k++) { const double a = M[p][k], bq = M[qq][k]; M[p][k] = c * a - sn * bq; M[qq][k] = sn * a + c * bq; }
      }
  }
  if (!(M[0][0] > 0.0 && M[1][1] > 0.0 && M[2][2] > 0.0)) return TBV_OK;  // not convex (:352-356)
  const double det = H[0][0] * (H[1][1] * H[2][2] - H[1][2] * H[2][1]) - H[0][1] * (H[1][0] * H[2][2] - H[1][2] * H[2][0]) +
                     H[0][2] * (H[1][0] * H[2][1] - H[1][1] * H[2][0]);
  if (det == 0.0 || !std::isfinite(det)) return TBV_OK;
  double Hi[3][3];
  Hi[0][0] = (H[1][1] * H[2][2] - H[1][2] * H[2][1]) / det; Hi[0][1] = (H[0][2] * H[2][1] - H[0][1] * H[2][2]) / det;
  Hi[0][2] = (H[0][1] * H[1][2] - H[0][2] * H[1][1]) / det; Hi[1][0] = (H[1][2] * H[2][0] - H[1][0] * H[2][2]) / det;
  Hi[1][1] = (H[0][0] * H[2][2] - H[0][2] * H[2][0]) / det; Hi[1][2] = (H[0][2] * H[1][0] - H[0][0] * H[1][2]) / det;
  Hi[2][0] = (H[1][0] * H[2][1] - H[1][1] * H[2][0]) / det; Hi[2][1] = (H[0][1] * H[2][0] - H[0][0] * H[2][1]) / det;
  Hi[2][2] = (H[0][0] * H[1][1] - H[0][1] * H[1][0]) / det;
  double C3[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C3[i][j] = 2.0 * Hi[i][j] * score_scale * cov_scaler;
  for (int i = 0; i < 36; i++) cov6x6[i] = (i % 7 == 0) ? 1.0 : 0.0;  // Identity, then the blocks of :368-375
  cov6x6[0] = C3[0][0]; cov6x6[1] = C3[0][1]; cov6x6[6] = C3[1][0]; cov6x6[7] = C3[1][1];
  cov6x6[35] = C3[2][2]; cov6x6[5] = C3[0][2]; cov6x6[11] = C3[1][2]; cov6x6[30] = C3[2][0]; cov6x6[31] = C3[2][1];
  *convex = 1;
  return TBV_OK;
}

int tbv_register_batch(tbv_ctx* ctx, int n_sets, const tbv_cell* const* sets, const int* n_cells, int n_pairs, const int* from_set,
                       const int* to_set, const double* T_from, const double* T_to, const tbv_reg_params* params, double* T_revised,
                       double* T_align, tbv_reg_summary* summaries) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && sets && n_cells && from_set && to_set && T_from && T_to && params && n_sets >= 1 && n_pairs >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  if (n_pairs == 0) return TBV_OK;
  HostProblemSet hs;
  hs.ctx = ctx;
  int rc = hs.upload(n_sets, sets, n_cells);
  if (rc) return rc;
  std::vector<RegProblem> hp(n_pairs);
  std::vector<int> hfs(n_pairs);
  std::vector<double> hfp((size_t)n_pairs * 3);
  for (int p = 0; p < n_pairs; p++) {
    TBV_REQUIRE(from_set[p] >= 0 && from_set[p] < n_sets && to_set[p] >= 0 && to_set[p] < n_sets, "pair indexes a missing set");
    hp[p].n_fixed = 1; hp[p].fixed_first = p; hp[p].src_set = from_set[p]; hp[p].active = 1;
    hfs[p] = to_set[p];
    for (int c = 0; c < 3; c++) { hp[p].src_pose[c] = T_from[3 * p + c]; hfp[3 * p + c] = T_to[3 * p + c]; }
  }
  DevBuf<RegProblem> dp; DevBuf<int> dfs; DevBuf<double> dfp; DevBuf<RegResult> dr;
  auto cleanup = [&]() { dp.release(); dfs.release(); dfp.release(); dr.release(); };
  if ((rc = to_device(ctx, dp, hp)) || (rc = to_device(ctx, dfs, hfs)) || (rc = to_device(ctx, dfp, hfp)) || (rc = dr.reserve(n_pairs))) {
    cleanup();
    return rc;
  }
  rc = register_launch(ctx, REG_MODE_REGISTER, 0, hs.views.p, dp.p, dfs.p, dfp.p, n_pairs, 1, hs.max_n, hs.max_n, to_dev(*params), dr.p, nullptr,
                       false);
  if (rc) { cleanup(); return rc; }
  std::vector<RegResult> hr(n_pairs);
  cudaError_t e = cudaMemcpyAsync(hr.data(), dr.p, n_pairs * sizeof(RegResult), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_register_batch: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  for (int p = 0; p < n_pairs; p++) {
    const RegResult& r = hr[p];
    if (T_revised) {
      // T_vek.back() read through Affine3dToVectorXYeZ; untouched input when no solve succeeded
      T_revised[3 * p + 0] = r.pose_updated ? r.pose[0] : T_from[3 * p + 0];
      T_revised[3 * p + 1] = r.pose_updated ? r.pose[1] : T_from[3 * p + 1];
      const double th = r.pose_updated ? r.pose[2] : T_from[3 * p + 2];
      T_revised[3 * p + 2] = std::atan2(std::sin(th), std::cos(th));
    }
    if (T_align) for (int c = 0; c < 3; c++) T_align[3 * p + c] = r.align[c];
    if (summaries) fill_summary(r, summaries + p);
  }
  return TBV_OK;
}

// CFEARQuality (coral_alignment_quality/src/alignment_checker/AlignmentQuality.cpp:330-352) for a batch of candidate pairs: one
// GetCost per pair — scans {ref (fixed, pose T_ref), src (pose T_src * T_offset)}, the registration object's itr_ = 0 (search radius
// radius_), caller's params (the reference: P2L, Huber 0.3, uniform weights) — all pairs in ONE launch of k_register in evaluation mode.
// quality[p] = {score (problem_->Evaluate's cost), number of residuals, (|src| + |ref|) / 2}; a failed GetCost (<= 1 residual) gives
// {0, 0, 0} like the reference's else branch.
int tbv_cfear_quality_batch(tbv_ctx* ctx, int n_sets, const tbv_cell* const* sets, const int* n_cells, int n_pairs, const int* src_set,
                            const int* ref_set, const double* T_src, const double* T_offset, const double* T_ref, const tbv_reg_params* params,
                            double* quality) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && sets && n_cells && src_set && ref_set && T_src && T_ref && params && quality && n_sets >= 1 && n_pairs >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);
  if (n_pairs == 0) return TBV_OK;
  HostProblemSet hs;
  hs.ctx = ctx;
  int rc = hs.upload(n_sets, sets, n_cells);
  if (rc) return rc;
  std::vector<RegProblem> hp(n_pairs);
  std::vector<int> hfs(n_pairs);
  std::vector<double> hfp((size_t)n_pairs * 3);
  for (int p = 0; p < n_pairs; p++) {
    TBV_REQUIRE(src_set[p] >= 0 && src_set[p] < n_sets && ref_set[p] >= 0 && ref_set[p] < n_sets, "pair indexes a missing set");
    hp[p].n_fixed = 1; hp[p].fixed_first = p; hp[p].src_set = src_set[p]; hp[p].active = 1;
    hfs[p] = ref_set[p];
    // src->GetAffine() * Toffset, read back as (x, y, yaw) by Affine3dToVectorXYeZ (utils.cpp:115-122)
    const double* a = T_src + 3 * p;
    double x = a[0], y = a[1], th = a[2];
    if (T_offset) {
      const double* o = T_offset + 3 * p;
      const double ca = std::cos(a[2]), sa = std::sin(a[2]), co = std::cos(o[2]), so = std::sin(o[2]);
      x = (ca * o[0] + (-sa) * o[1]) + a[0];
      y = (sa * o[0] + ca * o[1]) + a[1];
      th = std::atan2(sa * co + ca * so, sa * (-so) + ca * co);
    }
    hp[p].src_pose[0] = x; hp[p].src_pose[1] = y; hp[p].src_pose[2] = th;
    for (int c = 0; c < 3; c++) hfp[3 * p + c] = T_ref[3 * p + c];
  }
  DevBuf<RegProblem> dp; DevBuf<int> dfs; DevBuf<double> dfp, dev_eval; DevBuf<RegResult> dr;
  auto cleanup = [&]() { dp.release(); dfs.release(); dfp.release(); dev_eval.release(); dr.release(); };
  if ((rc = to_device(ctx, dp, hp)) || (rc = to_device(ctx, dfs, hfs)) || (rc = to_device(ctx, dfp, hfp)) || (rc = dr.reserve(n_pairs)) ||
      (rc = dev_eval.reserve((size_t)n_pairs * NACC))) { cleanup(); return rc; }
  rc = register_launch(ctx, REG_MODE_EVAL, 0, hs.views.p, dp.p, dfs.p, dfp.p, n_pairs, 1, hs.max_n, hs.max_n, to_dev(*params), dr.p, dev_eval.p, false);
  if (rc) { cleanup(); return rc; }
  std::vector<RegResult> hr(n_pairs);
  std::vector<double> ev((size_t)n_pairs * NACC);
  cudaError_t e = cudaMemcpyAsync(hr.data(), dr.p, n_pairs * sizeof(RegResult), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ev.data(), dev_eval.p, ev.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_cfear_quality_batch: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  for (int p = 0; p < n_pairs; p++) {
    const bool ok = hr[p].num_residuals > 1;
    quality[3 * p + 0] = ok ? ev[(size_t)p * NACC] : 0.0;
    quality[3 * p + 1] = ok ? (double)hr[p].num_residuals : 0.0;
    quality[3 * p + 2] = ok ? (n_cells[src_set[p]] + n_cells[ref_set[p]]) / 2.0 : 0.0;
  }
  return TBV_OK;
}

}  // extern "C"

// =============================================================================================
// tbv_oracle_reg.hpp — CPU ORACLE, part 2: registration (association + robust cost + Ceres-style LM),
// keyframe fuser, loop-candidate registration.            *** TEST INFRASTRUCTURE ONLY ***
// See tbv_oracle.hpp for the rules on who may use this and for the "parity unpinned" statement.
//
// Ceres-solver 2.1.0 (tbv_slam/docker/Dockerfile:9-15) is not vendored in the reference; the parts the
// reference calls (ceres::Solve with default TRUST_REGION / LEVENBERG_MARQUARDT options, Problem::Evaluate,
// loss functions, ScaledLoss, Corrector) are restated here from the published 2.1.0 algorithm:
//   internal/ceres/trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, trust_region_step_evaluator.cc,
//   corrector.cc, loss_function.cc, residual_block.cc, solver.cc (SetSummaryFinalCost).
// The linear solve (JtJ + D^2) y = Jt r is done by dense Cholesky; the reference build would use
// SPARSE_NORMAL_CHOLESKY through SuiteSparse (libsuitesparse-dev is installed in the Dockerfile) — the same
// normal equations, rounding-level differences only.  AutoDiff Jacobians are restated analytically.
// =============================================================================================
#pragma once
#include "tbv_oracle.hpp"

#include <memory>

namespace tbv_oracle {

enum cost_metric { P2P = 0, P2L = 1, P2D = 2 };                                            // registration.h:55
enum loss_type { None = 0, Huber = 1, Cauchy = 2, SoftLOne = 3, Combined = 4, Tukey = 5 };  // registration.h:60
enum weightoption { Uniform = 0, Sim_N = 1, Sim_direciton = 2, Sim_scale = 3, Combined_weights = 4 };  // registration.h:50

// ---------------------------------------------------------------------------------------------
// Ceres 2.1.0 loss functions (loss_function.cc) — rho[0]=rho(s), rho[1]=rho'(s), rho[2]=rho''(s)
// ---------------------------------------------------------------------------------------------
namespace ceres_restated {
inline void HuberLoss(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}
inline void CauchyLoss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c * (inv * inv);
}
inline void SoftLOneLoss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double tmp = std::sqrt(sum);
  rho[0] = 2.0 * b * (tmp - 1.0);
  rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
  rho[2] = -(c * rho[1]) / (2.0 * sum);
}
inline void TukeyLoss(double a, double s, double rho[3]) {
  const double a_squared = a * a;
  if (s <= a_squared) {
    const double value = 1.0 - s / a_squared;
    const double value_sq = value * value;
    rho[0] = a_squared / 3.0 * (1.0 - value_sq * value);
    rho[1] = value_sq;
    rho[2] = -2.0 / a_squared * value;
  } else {
    rho[0] = a_squared / 3.0; rho[1] = 0.0; rho[2] = 0.0;
  }
}
// Registration::GetLoss (registration.cpp:77-96) wrapped in ScaledLoss(…, w) (n_scan_normal.cpp:275).
// loss None -> GetLoss() returns nullptr -> ScaledLoss with null inner: rho = (a*s, a, 0).
inline void ScaledLoss(int loss, double loss_limit, double weight, double s, double rho[3]) {
  switch (loss) {
    case Huber: HuberLoss(loss_limit, s, rho); break;
    case Cauchy: CauchyLoss(loss_limit, s, rho); break;
    case SoftLOne: SoftLOneLoss(loss_limit, s, rho); break;
    case Tukey: TukeyLoss(loss_limit, s, rho); break;
    case Combined: {  // ComposedLoss(f = Huber(1), g = Cauchy(1)):  rho = f(g(s))
      double rho_g[3], rho_f[3];
      CauchyLoss(1.0, s, rho_g);
      HuberLoss(1.0, rho_g[0], rho_f);
      rho[0] = rho_f[0];
      rho[1] = rho_f[1] * rho_g[1];
      rho[2] = rho_f[2] * rho_g[1] * rho_g[1] + rho_f[1] * rho_g[2];
      break;
    }
    default: rho[0] = weight * s; rho[1] = weight; rho[2] = 0.0; return;
  }
  rho[0] *= weight; rho[1] *= weight; rho[2] *= weight;
}
// corrector.cc
struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
      residual_scaling = sqrt_rho1;
      alpha_sq_norm = 0.0;
      return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  void CorrectJacobian(int num_rows, int num_cols, const double* residuals, double* jacobian) const {  // row-major
    if (alpha_sq_norm == 0.0) {
      for (int i = 0; i < num_rows * num_cols; i++) jacobian[i] *= sqrt_rho1;
      return;
    }
    for (int c = 0; c < num_cols; ++c) {
      double r_transpose_j = 0.0;
      for (int r = 0; r < num_rows; ++r) r_transpose_j += jacobian[r * num_cols + c] * residuals[r];
      for (int r = 0; r < num_rows; ++r)
        jacobian[r * num_cols + c] = sqrt_rho1 * (jacobian[r * num_cols + c] - alpha_sq_norm * residuals[r] * r_transpose_j);
    }
  }
  void CorrectResiduals(int num_rows, double* residuals) const {
    for (int r = 0; r < num_rows; ++r) residuals[r] *= residual_scaling;
  }
};
}  // namespace ceres_restated

// ---------------------------------------------------------------------------------------------
// Residual blocks on the single free 3-vector (x, y, theta) of the moving scan
// ---------------------------------------------------------------------------------------------
struct ResidualBlock {
  int cost;           // cost_metric
  double src[2];      // src_mean_ (local frame of the moving scan)
  double tar[2];      // Ttar * tar_mean (world)
  double nrm[2];      // Ttar.linear() * tar_normal (P2L)
  double L[2][2];     // sqrt information (P2D)
  double weight;      // ScaledLoss a_
  int tar_scan, tar_idx, src_idx;
};
inline int BlockSize(int cost) { return cost == P2L ? 1 : 2; }

// P2LEfficientCost / P2PEfficientCost / P2DEfficientCost (n_scan_normal.h:180-255, 330-361), analytic Jacobian
// of the AutoDiffCostFunction<…, N, 3>.
inline void EvalCostFunction(const ResidualBlock& b, const double x[3], double* f, double* J /*N x 3 row-major or null*/) {
  const double cy = std::cos(x[2]), sy = std::sin(x[2]);
  const double mx = (cy * b.src[0] + (-sy) * b.src[1]) + x[0];  // (rot * src) + trans
  const double my = (sy * b.src[0] + cy * b.src[1]) + x[1];
  const double dmx = (-sy) * b.src[0] + (-cy) * b.src[1];  // d/dtheta
  const double dmy = cy * b.src[0] + (-sy) * b.src[1];
  if (b.cost == P2L) {
    const double v0 = mx - b.tar[0], v1 = my - b.tar[1];
    f[0] = v0 * b.nrm[0] + v1 * b.nrm[1];
    if (J) { J[0] = b.nrm[0]; J[1] = b.nrm[1]; J[2] = dmx * b.nrm[0] + dmy * b.nrm[1]; }
  } else if (b.cost == P2P) {
    f[0] = b.tar[0] - mx;
    f[1] = b.tar[1] - my;
    if (J) { J[0] = -1.0; J[1] = 0.0; J[2] = -dmx; J[3] = 0.0; J[4] = -1.0; J[5] = -dmy; }
  } else {  // P2D: residuals = L * (src_world - tar)
    const double e0 = mx - b.tar[0], e1 = my - b.tar[1];
    f[0] = b.L[0][0] * e0 + b.L[0][1] * e1;
    f[1] = b.L[1][0] * e0 + b.L[1][1] * e1;
    if (J) {
      J[0] = b.L[0][0]; J[1] = b.L[0][1]; J[2] = b.L[0][0] * dmx + b.L[0][1] * dmy;
      J[3] = b.L[1][0]; J[4] = b.L[1][1]; J[5] = b.L[1][0] * dmx + b.L[1][1] * dmy;
    }
  }
}

struct Problem {
  std::vector<ResidualBlock> blocks;
  int loss = Huber;
  double loss_limit = 0.1;
  int NumResiduals() const {
    int n = 0;
    for (const auto& b : blocks) n += BlockSize(b.cost);
    return n;
  }
  // residual_block.cc ResidualBlock::Evaluate + program_evaluator.h: cost = sum 0.5*rho[0]; residuals and
  // Jacobian corrected for the loss; gradient += J_block^T r_block.
  void Evaluate(const double x[3], double* cost, std::vector<double>* residuals, double gradient[3], std::vector<double>* jacobian) const {
    double total = 0.0;
    if (gradient) gradient[0] = gradient[1] = gradient[2] = 0.0;
    if (residuals) residuals->clear();
    if (jacobian) jacobian->clear();
    const bool need_jac = gradient || jacobian;
    for (const auto& b : blocks) {
      const int n = BlockSize(b.cost);
      double f[2], J[6];
      EvalCostFunction(b, x, f, need_jac ? J : nullptr);
      double sq = 0.0;
      for (int r = 0; r < n; r++) sq += f[r] * f[r];
      double rho[3];
      ceres_restated::ScaledLoss(loss, loss_limit, b.weight, sq, rho);
      total += 0.5 * rho[0];
      if (residuals || need_jac) {
        ceres_restated::Corrector correct(sq, rho);
        if (need_jac) correct.CorrectJacobian(n, 3, f, J);
        correct.CorrectResiduals(n, f);
      }
      if (residuals) for (int r = 0; r < n; r++) residuals->push_back(f[r]);
      if (jacobian) for (int i = 0; i < n * 3; i++) jacobian->push_back(J[i]);
      if (gradient)
        for (int c = 0; c < 3; c++) {
          double tmp = 0.0;
          for (int r = 0; r < n; r++) tmp += J[r * 3 + c] * f[r];
          gradient[c] += tmp;
        }
    }
    if (cost) *cost = total;
  }
};

// ---------------------------------------------------------------------------------------------
// ceres::Solve restated (TRUST_REGION, LEVENBERG_MARQUARDT, monotonic steps, jacobi scaling)
// ---------------------------------------------------------------------------------------------
struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0, relative_decrease = 0, trust_region_radius = 0;
};
enum TerminationType { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };
struct SolverSummary {
  std::vector<IterationSummary> iterations;
  double initial_cost = 0, final_cost = 0;
  int num_residuals = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  int termination_type = NO_CONVERGENCE;
  int num_cost_evaluations = 0, num_jacobian_evaluations = 0;
  bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE; }
};
struct SolverOptions {
  int max_num_iterations = 50;
  double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  int max_num_consecutive_invalid_steps = 5;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
};

namespace detail {
// solve (H) y = b for symmetric positive definite 3x3 by Cholesky; false if not positive definite / non-finite
inline bool CholSolve3(const double H[3][3], const double b[3], double y[3]) {
  double L[3][3] = {{0}};
  for (int j = 0; j < 3; j++) {
    double d = H[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    L[j][j] = std::sqrt(d);
    for (int i = j + 1; i < 3; i++) {
      double v = H[i][j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
  double z[3];
  for (int i = 0; i < 3; i++) {
    double v = b[i];
    for (int k = 0; k < i; k++) v -= L[i][k] * z[k];
    z[i] = v / L[i][i];
  }
  for (int i = 2; i >= 0; i--) {
    double v = z[i];
    for (int k = i + 1; k < 3; k++) v -= L[k][i] * y[k];
    y[i] = v / L[i][i];
  }
  return std::isfinite(y[0]) && std::isfinite(y[1]) && std::isfinite(y[2]);
}
}  // namespace detail

inline void Solve(const SolverOptions& opt, const Problem& problem, double parameters[3], SolverSummary* summary) {
  SolverSummary& S = *summary;
  S = SolverSummary();
  S.num_residuals = problem.NumResiduals();
  const int m = S.num_residuals;
  // --- Init
  double x[3] = {parameters[0], parameters[1], parameters[2]};
  double x_norm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  double x_cost = 0, candidate_cost = 0, minimum_cost = 0;
  std::vector<double> residuals, jac;  // jac: m x 3 row-major, column-scaled after evaluation
  double gradient[3], jacobian_scaling[3] = {1, 1, 1};
  // LevenbergMarquardtStrategy state
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0, diagonal[3] = {0, 0, 0};
  bool reuse_diagonal = false;
  // TrustRegionStepEvaluator (max_consecutive_nonmonotonic_steps = 0)
  double ev_minimum_cost, ev_current_cost, ev_reference_cost, ev_candidate_cost, ev_acc_ref = 0.0, ev_acc_cand = 0.0;
  int num_consecutive_invalid_steps = 0;
  IterationSummary it;

  auto EvaluateGradientAndJacobian = [&]() {
    problem.Evaluate(x, &x_cost, &residuals, gradient, &jac);
    S.num_cost_evaluations++; S.num_jacobian_evaluations++;
    it.cost = x_cost;
    if (it.iteration == 0) {
      for (int c = 0; c < 3; c++) {
        double sq = 0.0;
        for (int r = 0; r < m; r++) sq += jac[r * 3 + c] * jac[r * 3 + c];
        jacobian_scaling[c] = 1.0 / (1.0 + std::sqrt(sq));
      }
    }
    for (int r = 0; r < m; r++) for (int c = 0; c < 3; c++) jac[r * 3 + c] *= jacobian_scaling[c];
    // projected gradient step: x - Plus(x, -g)
    double gmax = 0.0, gn = 0.0;
    for (int c = 0; c < 3; c++) {
      const double pg = x[c] - (x[c] + (-gradient[c]));
      gmax = std::max(gmax, std::fabs(pg));
      gn += pg * pg;
    }
    it.gradient_max_norm = gmax;
    it.gradient_norm = std::sqrt(gn);
  };

  // --- IterationZero
  it = IterationSummary();
  it.trust_region_radius = radius;
  EvaluateGradientAndJacobian();
  S.initial_cost = x_cost;
  it.step_is_valid = true;
  it.step_is_successful = true;
  candidate_cost = x_cost;
  minimum_cost = x_cost;
  ev_minimum_cost = ev_current_cost = ev_reference_cost = ev_candidate_cost = x_cost;

  auto Finalize = [&]() -> bool {  // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it.step_is_successful) {
      ++S.num_successful_steps;
      if (x_cost < minimum_cost || it.iteration == 0) {
        minimum_cost = std::min(minimum_cost, x_cost);
        parameters[0] = x[0]; parameters[1] = x[1]; parameters[2] = x[2];
      }
    } else {
      ++S.num_unsuccessful_steps;
    }
    it.trust_region_radius = radius;
    S.iterations.push_back(it);
    if (it.iteration >= opt.max_num_iterations) { S.termination_type = NO_CONVERGENCE; return false; }
    if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { S.termination_type = CONVERGENCE; return false; }
    if (it.trust_region_radius < opt.min_trust_region_radius) { S.termination_type = CONVERGENCE; return false; }
    return true;
  };

  while (Finalize()) {
    const double previous_gradient_norm = it.gradient_norm, previous_gradient_max_norm = it.gradient_max_norm;
    const int next_iter = S.iterations.back().iteration + 1;
    it = IterationSummary();
    it.iteration = next_iter;
    it.gradient_norm = previous_gradient_norm;
    it.gradient_max_norm = previous_gradient_max_norm;

    // ---- ComputeTrustRegionStep: LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int c = 0; c < 3; c++) {
        double sq = 0.0;
        for (int r = 0; r < m; r++) sq += jac[r * 3 + c] * jac[r * 3 + c];
        diagonal[c] = std::min(std::max(sq, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
    }
    double lm_diagonal[3];
    for (int c = 0; c < 3; c++) lm_diagonal[c] = std::sqrt(diagonal[c] / radius);
    double H[3][3] = {{0}}, rhs[3] = {0, 0, 0};
    {
      // normal equations of [J; D] y = [r; 0], accumulated per residual block (row block) as Ceres does
      size_t row = 0;
      for (const auto& b : problem.blocks) {
        const int n = BlockSize(b.cost);
        for (int a = 0; a < 3; a++) {
          double t = 0.0;
          for (int r = 0; r < n; r++) t += jac[(row + r) * 3 + a] * residuals[row + r];
          rhs[a] += t;
          for (int c = a; c < 3; c++) {
            double h = 0.0;
            for (int r = 0; r < n; r++) h += jac[(row + r) * 3 + a] * jac[(row + r) * 3 + c];
            H[a][c] += h;
          }
        }
        row += n;
      }
      for (int a = 0; a < 3; a++) H[a][a] += lm_diagonal[a] * lm_diagonal[a];
      for (int a = 0; a < 3; a++) for (int c = 0; c < a; c++) H[a][c] = H[c][a];
    }
    double step[3];
    const bool solved = detail::CholSolve3(H, rhs, step);
    reuse_diagonal = true;
    it.step_is_valid = false;
    double model_cost_change = 0.0, delta[3] = {0, 0, 0};
    if (solved) {
      for (int c = 0; c < 3; c++) step[c] *= -1.0;
      // model_cost_change = -(J*step)^T (f + J*step/2)
      double acc = 0.0;
      for (int r = 0; r < m; r++) {
        const double mr = jac[r * 3 + 0] * step[0] + jac[r * 3 + 1] * step[1] + jac[r * 3 + 2] * step[2];
        acc += mr * (residuals[r] + mr / 2.0);
      }
      model_cost_change = -acc;
      it.step_is_valid = (model_cost_change > 0.0);
      if (it.step_is_valid) {
        for (int c = 0; c < 3; c++) delta[c] = step[c] * jacobian_scaling[c];
        num_consecutive_invalid_steps = 0;
      }
    }
    if (!it.step_is_valid) {  // HandleInvalidStep
      ++num_consecutive_invalid_steps;
      if (num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) { S.termination_type = FAILURE; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid -> StepRejected(0)
      it.cost = x_cost; it.cost_change = 0.0; it.step_norm = 0.0; it.relative_decrease = 0.0;
      it.gradient_max_norm = S.iterations.back().gradient_max_norm;
      it.gradient_norm = S.iterations.back().gradient_norm;
      continue;
    }
    // ---- ComputeCandidatePointAndEvaluateCost
    double candidate_x[3] = {x[0] + delta[0], x[1] + delta[1], x[2] + delta[2]};
    problem.Evaluate(candidate_x, &candidate_cost, nullptr, nullptr, nullptr);
    S.num_cost_evaluations++;
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();
    // ---- ParameterToleranceReached
    {
      const double d0 = x[0] - candidate_x[0], d1 = x[1] - candidate_x[1], d2 = x[2] - candidate_x[2];
      it.step_norm = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      const double step_size_tolerance = opt.parameter_tolerance * (x_norm + opt.parameter_tolerance);
      if (it.step_norm <= step_size_tolerance) { S.termination_type = CONVERGENCE; break; }
    }
    // ---- FunctionToleranceReached
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) { S.termination_type = CONVERGENCE; break; }
    // ---- IsStepSuccessful (TrustRegionStepEvaluator::StepQuality)
    if (candidate_cost >= std::numeric_limits<double>::max()) {
      it.relative_decrease = std::numeric_limits<double>::lowest();
    } else {
      const double relative_decrease = (ev_current_cost - candidate_cost) / model_cost_change;
      const double historical = (ev_reference_cost - candidate_cost) / (ev_acc_ref + model_cost_change);
      it.relative_decrease = std::max(relative_decrease, historical);
    }
    if (it.relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      x[0] = candidate_x[0]; x[1] = candidate_x[1]; x[2] = candidate_x[2];
      x_norm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      EvaluateGradientAndJacobian();
      it.step_is_successful = true;
      // strategy_->StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      // step_evaluator_->StepAccepted(candidate_cost, model_cost_change)
      ev_current_cost = candidate_cost;
      ev_acc_cand += model_cost_change;
      ev_acc_ref += model_cost_change;
      if (ev_current_cost < ev_minimum_cost) {
        ev_minimum_cost = ev_current_cost; ev_candidate_cost = ev_current_cost; ev_acc_cand = 0.0;
        ev_reference_cost = ev_candidate_cost; ev_acc_ref = ev_acc_cand;  // num_consecutive_nonmonotonic_steps (0) == max (0)
      } else if (ev_current_cost > ev_candidate_cost) {
        ev_candidate_cost = ev_current_cost; ev_acc_cand = 0.0;
      }
    } else {  // unsuccessful
      it.step_is_successful = false;
      it.cost = candidate_cost;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepRejected
    }
  }
  // solver.cc SetSummaryFinalCost
  S.final_cost = S.initial_cost;
  for (const auto& i : S.iterations) S.final_cost = std::min(i.cost, S.final_cost);
}

// =============================================================================================
// Registration::Weights                                    registration.h:88-101, registration.cpp:67-75
// =============================================================================================
inline double Similarity(double x, double y) { return 2 * std::min(x, y) / (x + y); }
inline double GetWeight(int opt, double N1, double N2, double sim_dir, double plan1, double plan2) {
  switch (opt) {
    case Uniform: return 1.0;
    case Sim_N: return Similarity(N1, N2);
    case Sim_direciton: return sim_dir;
    case Sim_scale: return Similarity(plan1, plan2);
    case Combined_weights: return Similarity(N1, N2) + sim_dir + Similarity(plan1, plan2);
  }
  return 1.0;
}

// =============================================================================================
// (a10-a14) n_scan_normal_reg                              n_scan_normal.cpp:7-28, 82-211, 213-324, 342-389, 441-450
// =============================================================================================
typedef std::shared_ptr<MapPointNormal> MapNormalPtr;

class n_scan_normal_reg {
 public:
  n_scan_normal_reg() { options_.max_num_iterations = 20; }  // :9
  n_scan_normal_reg(int cost, int loss = Huber, double loss_limit = 0.1, int opt = Uniform) : n_scan_normal_reg() {
    cost_ = cost; loss_ = loss; loss_limit_ = loss_limit; weight_opt_ = opt;
  }
  void SetParameters(unsigned max_itr_association, unsigned max_itr_solver) {  // :16-20
    max_itr_association_ = max_itr_association;
    options_.max_num_iterations = (int)max_itr_solver;
  }
  void SetD2dPar(double cov_scale, double regularization) { cov_scale_ = cov_scale; regularization_ = regularization; }
  double getScore() const { return score_; }

  // Register :82-185.  T in/out as Affine2; reg_cov omitted (fixed diag(0.1^2,0.1^2,0,0,0,0.01^2), :171-175)
  bool Register(const std::vector<MapNormalPtr>& scans, std::vector<Affine2>& Tsrc) {
    const size_t n_scans = scans.size();
    parameters.assign(n_scans, std::vector<double>(3, 0.0));
    for (size_t i = 0; i < n_scans; i++) AffineToVector(Tsrc[i], parameters[i].data());
    bool success = true;
    std::vector<double> prev_par = parameters.back();
    double prev_score = DBL_MAX;
    total_lm_iterations_ = 0;
    for (itr_ = 1; itr_ <= max_itr_association_ && success; itr_++) {
      success = BuildOptimizationProblem(scans);
      if (!success) break;
      success = SolveOptimizationProblem();
      if (success)
        for (size_t i = 0; i < n_scans; i++) Tsrc[i] = vectorToAffine(parameters[i][0], parameters[i][1], parameters[i][2]);
      const double current_score = summary_.final_cost;
      const double rel_improvement = (prev_score - current_score) / prev_score;
      if (itr_ > min_itr_) {
        if (prev_score < current_score) {  // :135
          parameters.back() = prev_par;
          break;
        } else if (rel_improvement < score_tolerance) {
          break;
        } else if (summary_.iterations.back().relative_decrease < score_tolerance || summary_.iterations.size() == 1) {
          break;
        }
      }
      prev_score = current_score;
      prev_par = parameters.back();
    }
    if (success) {
      score_ = summary_.final_cost / summary_.num_residuals;
      for (size_t i = 0; i < n_scans; i++) Tsrc[i] = vectorToAffine(parameters[i][0], parameters[i][1], parameters[i][2]);
      return true;
    }
    return false;
  }

  // GetCost :186-211
  bool GetCost(const std::vector<MapNormalPtr>& scans, const std::vector<Affine2>& Tsrc, double& score, std::vector<double>& residuals) {
    const size_t n_scans = scans.size();
    parameters.assign(n_scans, std::vector<double>(3, 0.0));
    for (size_t i = 0; i < n_scans; i++) AffineToVector(Tsrc[i], parameters[i].data());
    if (!BuildOptimizationProblem(scans)) return false;
    if (problem_.NumResiduals() <= 1) return false;
    problem_.Evaluate(parameters.back().data(), &score, &residuals, nullptr, nullptr);
    score_ = score / (std::max((int)residuals.size(), 1));
    return true;
  }

  // AddScanPairCost :213-324
  void AddScanPairCost(const MapPointNormal& target_local, const MapPointNormal& src_local, const Affine2& Ttar, const Affine2& Tsrc,
                       size_t scan_idx_tar) {
    const double angle_outlier = std::cos(M_PI / 6.0);
    const double curr_radius = (itr_ == 1) ? 2 * radius_ : radius_;
    const Affine2 Tsrctotar = Mul(Inverse(Ttar), Tsrc);
    std::vector<std::pair<int, int>> assoc;
    std::vector<double> wts;
    for (size_t src_idx = 0; src_idx < src_local.GetSize(); src_idx++) {
      const Cell& cs = src_local.GetCell(src_idx);
      double qx, qy;
      Apply(Tsrctotar, cs.u[0], cs.u[1], qx, qy);
      const int tar_idx = target_local.GetClosestIdx(qx, qy, curr_radius);
      if (tar_idx < 0) continue;
      const Cell& ct = target_local.GetCell(tar_idx);
      const double snx = Tsrctotar.r00 * cs.snormal[0] + Tsrctotar.r01 * cs.snormal[1];
      const double sny = Tsrctotar.r10 * cs.snormal[0] + Tsrctotar.r11 * cs.snormal[1];
      const double direction_similarity = std::max(snx * ct.snormal[0] + sny * ct.snormal[1], 0.0);
      if (direction_similarity > angle_outlier) {
        const double n_src = (double)cs.Nsamples, n_tar = (double)ct.Nsamples;
        wts.push_back(GetWeight(weight_opt_, n_src, n_tar, direction_similarity, cs.scale, ct.scale));
        assoc.push_back(std::make_pair(tar_idx, (int)src_idx));
      }
    }
    for (size_t i = 0; i < assoc.size(); i++) {
      const Cell& ct = target_local.GetCell(assoc[i].first);
      const Cell& cs = src_local.GetCell(assoc[i].second);
      ResidualBlock b;
      std::memset(&b, 0, sizeof(b));
      b.cost = cost_;
      b.weight = wts[i];
      b.src[0] = cs.u[0]; b.src[1] = cs.u[1];
      Apply(Ttar, ct.u[0], ct.u[1], b.tar[0], b.tar[1]);
      b.tar_scan = (int)scan_idx_tar; b.tar_idx = assoc[i].first; b.src_idx = assoc[i].second;
      if (cost_ == P2L) {
        b.nrm[0] = Ttar.r00 * ct.snormal[0] + Ttar.r01 * ct.snormal[1];
        b.nrm[1] = Ttar.r10 * ct.snormal[0] + Ttar.r11 * ct.snormal[1];
      } else if (cost_ == P2D) {  // :288-298
        // tar_cov = (reg*I + R*C*R^T)*cov_scale ; sqrt_information = tar_cov.inverse().llt().matrixL()
        const double R[2][2] = {{Ttar.r00, Ttar.r01}, {Ttar.r10, Ttar.r11}};
        double RC[2][2], RCRt[2][2];
        for (int a = 0; a < 2; a++) for (int c = 0; c < 2; c++) RC[a][c] = R[a][0] * ct.cov[0][c] + R[a][1] * ct.cov[1][c];
        for (int a = 0; a < 2; a++) for (int c = 0; c < 2; c++) RCRt[a][c] = RC[a][0] * R[c][0] + RC[a][1] * R[c][1];
        double tc[2][2];
        tc[0][0] = (regularization_ + RCRt[0][0]) * cov_scale_; tc[0][1] = (0.0 + RCRt[0][1]) * cov_scale_;
        tc[1][0] = (0.0 + RCRt[1][0]) * cov_scale_;             tc[1][1] = (regularization_ + RCRt[1][1]) * cov_scale_;
        const double det = tc[0][0] * tc[1][1] - tc[1][0] * tc[0][1];
        const double invdet = 1.0 / det;
        const double i00 = tc[1][1] * invdet, i10 = -tc[1][0] * invdet, i11 = tc[0][0] * invdet;
        const double l00 = std::sqrt(i00);  // Eigen LLT reads the lower triangle
        const double l10 = i10 / l00;
        const double l11 = std::sqrt(i11 - l10 * l10);
        b.L[0][0] = l00; b.L[0][1] = 0.0; b.L[1][0] = l10; b.L[1][1] = l11;
      }
      problem_.blocks.push_back(b);
    }
  }

  SolverSummary summary_;
  size_t itr_ = 0;
  int total_lm_iterations_ = 0;
  Problem problem_;
  std::vector<std::vector<double>> parameters;

 private:
  // BuildOptimizationProblem :342-389 — incremental_last_to_previous: pairs (each fixed scan i) -> last scan
  bool BuildOptimizationProblem(const std::vector<MapNormalPtr>& scans) {
    const size_t n = scans.size();
    std::vector<Affine2> Tvek(n);
    for (size_t i = 0; i < n; i++) Tvek[i] = vectorToAffine(parameters[i][0], parameters[i][1], parameters[i][2]);
    problem_ = Problem();
    problem_.loss = loss_;
    problem_.loss_limit = loss_limit_;
    const size_t j = n - 1;
    for (size_t i = 0; i + 1 < n; i++) AddScanPairCost(*scans[i], *scans[j], Tvek[i], Tvek[j], i);
    if (problem_.NumResiduals() <= 1) return false;
    return true;
  }
  bool SolveOptimizationProblem() {  // :441-450
    if (problem_.NumResiduals() <= 1) return false;
    Solve(options_, problem_, parameters.back().data(), &summary_);
    total_lm_iterations_ += (int)summary_.iterations.size() - 1;
    return summary_.IsSolutionUsable();
  }
  int cost_ = P2L, loss_ = Huber, weight_opt_ = Uniform;
  double loss_limit_ = 0.1;
  double cov_scale_ = 1, regularization_ = 0.01;
  const double score_tolerance = 0.00001;
  double max_itr_association_ = 8, min_itr_ = 3;
  double radius_ = 2.0;  // registration.h:122
  double score_ = 0;
  SolverOptions options_;
};

// =============================================================================================
// (a15) OdometryKeyframeFuser::processFrame                 odometrykeyframefuser.cpp:62-94, 143-259, 470-494
// =============================================================================================
struct FuserParameters {  // odometrykeyframefuser.h:90-113 (subset on the hot path)
  int cost_type = P2L;
  int weight_opt = Uniform;
  int submap_scan_size = 3;
  bool weight_intensity = false;
  bool use_guess = true, compensate = true, radar_ccw = false, use_keyframe = true;
  double res = 3.5;
  double min_keyframe_dist = 1.5, min_keyframe_rot_deg = 5;
  int loss_type = Huber;
  double loss_limit = 0.1, covar_scale = 1.0, regularization = 0.0;
  double downsample_factor = 1.0;  // MapPointNormal::downsample_factor (static)
  int voxel_order = VOXEL_ORDER_STABLE;
};
struct Keyframe {
  Affine2 pose;
  MapNormalPtr normals;
};
class OdometryKeyframeFuser {
 public:
  explicit OdometryKeyframeFuser(const FuserParameters& p) : par(p), radar_reg(p.cost_type, p.loss_type, p.loss_limit, p.weight_opt) {
    radar_reg.SetD2dPar(p.covar_scale, p.regularization);  // :32
  }
  static bool KeyFrameBasedFuse(const Affine2& diff, bool use_keyframe, double min_keyframe_dist, double min_keyframe_rot_deg) {  // :62-73
    if (!use_keyframe) return true;
    // diff.rotation().eulerAngles(0,1,2) -> (0,0,yaw) for a planar transform; normalizeEulerAngles leaves it in (-pi,pi]
    const double yaw = std::atan2(diff.r10, diff.r11);
    const double tnorm = std::sqrt(diff.tx * diff.tx + diff.ty * diff.ty);
    return tnorm > min_keyframe_dist || std::fabs(yaw) > (min_keyframe_rot_deg * M_PI / 180.0);
  }
  static bool AccelerationVelocitySanityCheck(const Affine2& Tmot_prev, const Affine2& Tmot_curr) {  // :76-94
    const double dt = 0.25, vel_limit = 200, acc_limit = 200;
    const double vx = Tmot_curr.tx / dt, vy = Tmot_curr.ty / dt;
    const double vel = std::sqrt(vx * vx + vy * vy);
    const double ax = (Tmot_curr.tx - Tmot_prev.tx) / (dt * dt), ay = (Tmot_curr.ty - Tmot_prev.ty) / (dt * dt);
    const double acc = std::sqrt(ax * ax + ay * ay);
    if (acc > acc_limit) return false;
    else if (vel > vel_limit) return false;
    return true;
  }
  // processFrame :143-259.  cloud is modified in place (compensated).  Returns the pose estimate Tcurrent.
  Affine2 processFrame(Cloud& cloud, Cloud* cloud_peaks = nullptr) {
    const Affine2 TprevMot = Tmot;
    if (par.compensate) {
      double mot[3];
      AffineToVector(TprevMot, mot);  // Compensate(cloud, Affine3d) -> Affine3dToVectorXYeZ  (utils.cpp:109-113)
      Compensate(cloud, mot, par.radar_ccw);
      if (cloud_peaks) Compensate(*cloud_peaks, mot, par.radar_ccw);
    }
    const double origin[2] = {0, 0};
    MapNormalPtr Pcurrent(new MapPointNormal(cloud, (float)par.res, origin, par.weight_intensity, par.downsample_factor, (VoxelOrder)par.voxel_order));
    last_n_cells = (int)Pcurrent->GetSize();
    const Affine2 Tguess = par.use_guess ? Mul(T_prev, TprevMot) : T_prev;
    updated = false;
    last_reg_ok = true;
    last_itrs = 0;
    if (keyframes_.empty()) {
      keyframes_.push_back({Affine2(), Pcurrent});  // AddToReference :470-476
      updated = true;
      return Tcurrent;
    }
    std::vector<MapNormalPtr> scans_vek;
    std::vector<Affine2> T_vek;
    for (auto& k : keyframes_) { scans_vek.push_back(k.normals); T_vek.push_back(k.pose); }  // FormatScans :478-494
    scans_vek.push_back(Pcurrent);
    T_vek.push_back(Tguess);
    last_reg_ok = radar_reg.Register(scans_vek, T_vek);  // result ignored by the reference (shadowed `success`, :184-193)
    last_itrs = (int)radar_reg.itr_;
    Tcurrent = T_vek.back();
    const Affine2 Tmot_current = Mul(Inverse(T_prev), Tcurrent);
    if (!AccelerationVelocitySanityCheck(Tmot, Tmot_current)) Tcurrent = Tguess;
    Tmot = Mul(Inverse(T_prev), Tcurrent);
    const Affine2 Tkeydiff = Mul(Inverse(keyframes_.back().pose), Tcurrent);
    const bool fuse = KeyFrameBasedFuse(Tkeydiff, par.use_keyframe, par.min_keyframe_dist, par.min_keyframe_rot_deg);
    if (fuse) {
      keyframes_.push_back({Tcurrent, Pcurrent});
      if (keyframes_.size() > (size_t)par.submap_scan_size) keyframes_.erase(keyframes_.begin());
      updated = true;
    }
    T_prev = Tcurrent;
    return Tcurrent;
  }
  FuserParameters par;
  n_scan_normal_reg radar_reg;
  std::vector<Keyframe> keyframes_;
  Affine2 Tcurrent, T_prev, Tmot;
  bool updated = false, last_reg_ok = true;
  int last_n_cells = 0, last_itrs = 0;
};

// =============================================================================================
// (a16) loopclosure::Register                               tbv_slam/src/tbv_slam/loopclosure.cpp:35-97, 320-364
// =============================================================================================
// scans {to(fixed), from}; poses {Tto, Tfrom}; P2L, Huber 0.1, Uniform weights, SetParameters(4,10).
// Result Talign = Trevised^-1 * Tto.
inline bool LoopRegister(const MapNormalPtr& normals_from, const MapNormalPtr& normals_to, const Affine2& Tfrom, const Affine2& Tto,
                         Affine2& Talign, Affine2* Trevised_out = nullptr, int* itrs = nullptr, double* score = nullptr) {
  std::vector<MapNormalPtr> scans_vek{normals_to, normals_from};
  std::vector<Affine2> T_vek{Tto, Tfrom};
  n_scan_normal_reg radar_reg(P2L);
  radar_reg.SetParameters(4, 10);
  const bool reg_success = radar_reg.Register(scans_vek, T_vek);
  const Affine2 Trevised = T_vek.back();
  if (Trevised_out) *Trevised_out = Trevised;
  if (itrs) *itrs = (int)radar_reg.itr_;
  if (score) *score = radar_reg.getScore();
  if (reg_success) Talign = Mul(Inverse(Trevised), Tto);
  return reg_success;
}

}  // namespace tbv_oracle

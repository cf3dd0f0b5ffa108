// CPU test of the trust-region loop in tbv_b200::CeresLeastSquaresT (include/tbv_b200.hpp): the backend is the oracle's normal-equation
// assembly plus a dense Cholesky of the damped system, standing in for tbv_pgo_assemble / tbv_pgo_solve_step (which are checked against the
// same two checkers on the GPU, tests/test_loop_gpu.py).  Test infrastructure only.  Prints PASS.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "oracle_backend.hpp"

using tbv_b200::Constraint3d;
using tbv_b200::Pose3d;

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

static Pose3d planar(double x, double y, double t) { Pose3d P; P.p[0] = x; P.p[1] = y; P.q[2] = std::sin(t / 2); P.q[3] = std::cos(t / 2); return P; }

static void ring(int n, double noise, unsigned seed, std::vector<Pose3d>& truth, std::vector<Pose3d>& start, std::vector<Constraint3d>& cons) {
  std::mt19937 gen(seed);
  std::normal_distribution<double> N01(0.0, 1.0);
  std::vector<double> th(n);
  truth.clear(); cons.clear();
  for (int i = 0; i < n; i++) { th[i] = 0.12 * i + M_PI / 2; truth.push_back(planar(15 * std::cos(0.12 * i), 15 * std::sin(0.12 * i), th[i])); }
  auto add = [&](int a, int b, int type) {
    const double dx = truth[b].p[0] - truth[a].p[0], dy = truth[b].p[1] - truth[a].p[1], c = std::cos(-th[a]), s = std::sin(-th[a]);
    Constraint3d C;
    C.id_begin = a; C.id_end = b; C.type = type;
    C.t_be = planar(c * dx - s * dy + noise * 0.02 * N01(gen), s * dx + c * dy + noise * 0.02 * N01(gen), th[b] - th[a] + noise * 0.002 * N01(gen));
    for (int k = 0; k < 6; k++) C.information[7 * k] = 1.0;
    cons.push_back(C);
  };
  for (int i = 0; i + 1 < n; i++) { add(i, i + 1, 0); if (i % 4 == 3 && i + 1 >= 52) add(i + 1 - 52, i + 1, 1); }
  Constraint3d skipped; skipped.id_begin = 1; skipped.id_end = 0; skipped.type = 3;          // a `candidate` constraint is not optimised
  cons.push_back(skipped);
  start = truth;
  for (int i = 1; i < n; i++) {
    const double t = th[i] + 0.03 * N01(gen);
    start[i] = planar(truth[i].p[0] + 0.2 * N01(gen), truth[i].p[1] + 0.2 * N01(gen), t);
  }
}

int main() {
  typedef tbv_b200::CeresLeastSquaresT<OracleBackend> Solver;
  {  // Plus: rotation composition, identity for a zero step
    const double x[7] = {1, 2, 3, 0.1, -0.2, 0.3, std::sqrt(1 - 0.14)}, z[6] = {0, 0, 0, 0, 0, 0}, d[6] = {0.5, -0.5, 0.25, 0, 0, 0.4};
    double o[7];
    Solver::Plus(x, z, o);
    for (int k = 0; k < 7; k++) EXPECT(o[k] == x[k]);
    Solver::Plus(x, d, o);
    double want[4];
    const double dq[4] = {0, 0, std::sin(0.4), std::cos(0.4)};
    tbv_oracle::detail::QuatMul(dq, x + 3, want);
    for (int k = 0; k < 4; k++) EXPECT(std::fabs(o[3 + k] - want[k]) < 1e-15);
    EXPECT(o[0] == 1.5 && o[1] == 1.5 && o[2] == 3.25);
  }
  std::vector<Pose3d> truth, nodes;
  std::vector<Constraint3d> cons;
  {  // consistent graph: the optimum is the ground truth
    ring(60, 0.0, 1, truth, nodes, cons);
    const Pose3d first = nodes[0];
    Solver s(OracleBackend(), nodes, cons);
    s.options.function_tolerance = 1e-16; s.options.gradient_tolerance = 1e-12; s.options.parameter_tolerance = 1e-14;
    s.Solve();
    EXPECT(s.summary_.initial_cost > 1.0 && s.summary_.final_cost <= 1e-14 * s.summary_.initial_cost);
    EXPECT(s.summary_.num_successful_steps >= 3 && s.summary_.IsSolutionUsable());
    double worst = 0;
    for (size_t i = 0; i < nodes.size(); i++) worst = std::fmax(worst, std::hypot(nodes[i].p[0] - truth[i].p[0], nodes[i].p[1] - truth[i].p[1]));
    EXPECT(worst < 1e-6);
    for (int k = 0; k < 3; k++) EXPECT(nodes[0].p[k] == first.p[k]);
    for (int k = 0; k < 4; k++) EXPECT(nodes[0].q[k] == first.q[k]);
    std::printf("consistent graph: %d LM iterations, cost %.3e -> %.3e (%s), worst position error %.2e m\n", s.summary_.iterations, s.summary_.initial_cost,
                s.summary_.final_cost, s.summary_.termination.c_str(), worst);
  }
  {  // noisy graph, Ceres' default tolerances: converges; tighter tolerances take no fewer iterations and do not end higher
    ring(60, 1.0, 2, truth, nodes, cons);
    std::vector<Pose3d> a = nodes, b = nodes, c = nodes;
    Solver s1(OracleBackend(), a, cons);
    s1.Solve();
    EXPECT((s1.summary_.termination == "function_tolerance" || s1.summary_.termination == "gradient_tolerance") && s1.summary_.final_cost < s1.summary_.initial_cost &&
           s1.summary_.iterations <= 200);
    Solver s2(OracleBackend(), b, cons);
    s2.options.function_tolerance = 1e-15; s2.options.gradient_tolerance = 1e-9; s2.options.parameter_tolerance = 1e-15;
    s2.Solve();
    EXPECT(s2.summary_.iterations >= s1.summary_.iterations && s2.summary_.final_cost <= s1.summary_.final_cost * (1 + 1e-12));
    Solver s3(OracleBackend(), c, cons);
    s3.options.max_num_iterations = 1;
    s3.Solve();
    EXPECT(s3.summary_.iterations == 1 && s3.summary_.termination == "max_num_iterations" && s3.summary_.final_cost <= s3.summary_.initial_cost);
    Solver s4(OracleBackend(), b, cons);                 // restart at the optimum: nothing left to do
    s4.options.gradient_tolerance = 1e-6;
    s4.Solve();
    EXPECT(s4.summary_.iterations <= 1);
    std::printf("noisy graph: defaults %d iterations (%s), tight %d iterations (%s), cost %.6e / %.6e\n", s1.summary_.iterations,
                s1.summary_.termination.c_str(), s2.summary_.iterations, s2.summary_.termination.c_str(), s1.summary_.final_cost, s2.summary_.final_cost);
  }
  {  // a constraint to a node that does not exist is an error, as in the C-ABI
    std::vector<Pose3d> two(2);
    std::vector<Constraint3d> bad(1);
    bad[0].id_begin = 0; bad[0].id_end = 5;
    bool threw = false;
    try { Solver s(OracleBackend(), two, bad); } catch (const tbv_b200::Error& e) { threw = e.code == TBV_ERR_INVALID; }
    EXPECT(threw);
  }
  if (fails) { std::printf("%d FAILED\n", fails); return 1; }
  std::printf("PASS\n");
  return 0;
}

// Oracle-backed stand-in for the two device calls of tbv_b200::CeresLeastSquaresT (tests only): the oracle's normal-equation assembly and a dense
// Cholesky of the damped system in place of tbv_pgo_assemble / tbv_pgo_solve_step.
#pragma once
#include <cmath>
#include <vector>

#include "tbv_b200.hpp"
#include "tbv_oracle.hpp"
#include "tbv_oracle_reg.hpp"
#include "tbv_oracle_loop.hpp"

struct OracleBackend {
  void assemble(int n, const double* nodes, int m, const int* ids, const double* meas, const double* info, const tbv_pgo_params& par, int fixed,
                double* cost, double* Hd, double* Ho, double* g) const {
    std::vector<tbv_oracle::PGNode> N(n);
    std::vector<tbv_oracle::PGConstraint> C(m);
    for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) N[i].p[k] = nodes[7 * i + k]; for (int k = 0; k < 4; k++) N[i].q[k] = nodes[7 * i + 3 + k]; }
    for (int c = 0; c < m; c++) {
      C[c].id_begin = ids[3 * c]; C[c].id_end = ids[3 * c + 1]; C[c].type = ids[3 * c + 2];
      for (int k = 0; k < 3; k++) C[c].p[k] = meas[7 * c + k];
      for (int k = 0; k < 4; k++) C[c].q[k] = meas[7 * c + 3 + k];
      for (int k = 0; k < 36; k++) C[c].info[k] = info ? info[36 * c + k] : 0.0;
    }
    tbv_oracle::PGParams P;
    P.odom_vxx = par.odom_vxx; P.odom_vyy = par.odom_vyy; P.odom_vtt = par.odom_vtt; P.loop_scaling = par.loop_scaling;
    P.replace_cov_by_identity = par.replace_cov_by_identity != 0; P.loop_cauchy = par.loop_cauchy;
    std::vector<double> res(6 * (size_t)(m ? m : 1));
    *cost = tbv_oracle::PGAssemble(N, C, P, fixed, Hd, Ho, g, res.data());
  }
  void solve(int n, int m, const int* ids, const double* Hd, const double* Ho, const double* g, int fixed, double radius, int, double, double* delta,
             int* iters) const {
    const int D = 6 * n;
    std::vector<double> A((size_t)D * D, 0.0), b(D);
    for (int i = 0; i < n; i++)
      for (int a = 0; a < 6; a++)
        for (int c = 0; c < 6; c++) A[(size_t)(6 * i + a) * D + 6 * i + c] = Hd[36 * i + 6 * a + c];
    for (int c = 0; c < m; c++)
      for (int a = 0; a < 6; a++)
        for (int e = 0; e < 6; e++) {
          A[(size_t)(6 * ids[3 * c] + a) * D + 6 * ids[3 * c + 1] + e] += Ho[36 * c + 6 * a + e];
          A[(size_t)(6 * ids[3 * c + 1] + e) * D + 6 * ids[3 * c] + a] += Ho[36 * c + 6 * a + e];
        }
    for (int k = 0; k < D; k++) { A[(size_t)k * D + k] += std::fmin(std::fmax(A[(size_t)k * D + k], 1e-6), 1e32) / radius; b[k] = -g[k]; }
    for (int k = 6 * fixed; k < 6 * fixed + 6; k++) {        // hold the fixed node: unit row / column, zero right-hand side
      for (int j = 0; j < D; j++) A[(size_t)k * D + j] = A[(size_t)j * D + k] = 0.0;
      A[(size_t)k * D + k] = 1.0; b[k] = 0.0;
    }
    for (int j = 0; j < D; j++) {                            // dense Cholesky, in place (lower)
      double d = A[(size_t)j * D + j];
      for (int k = 0; k < j; k++) d -= A[(size_t)j * D + k] * A[(size_t)j * D + k];
      d = std::sqrt(d);
      A[(size_t)j * D + j] = d;
      for (int i = j + 1; i < D; i++) {
        double v = A[(size_t)i * D + j];
        for (int k = 0; k < j; k++) v -= A[(size_t)i * D + k] * A[(size_t)j * D + k];
        A[(size_t)i * D + j] = v / d;
      }
    }
    for (int i = 0; i < D; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= A[(size_t)i * D + k] * b[k]; b[i] = v / A[(size_t)i * D + i]; }
    for (int i = D - 1; i >= 0; i--) { double v = b[i]; for (int k = i + 1; k < D; k++) v -= A[(size_t)k * D + i] * b[k]; b[i] = v / A[(size_t)i * D + i]; }
    for (int k = 0; k < D; k++) delta[k] = b[k];
    *iters = 1;
  }
};


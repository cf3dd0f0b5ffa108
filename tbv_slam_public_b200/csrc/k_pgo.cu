// k_pgo.cu — K8: pose-graph residuals, tangent-space Jacobians and block normal equations.
//
// Replaces CeresLeastSquares::{BuildOptimizationProblem, AddConstraintType} (tbv_slam/src/tbv_slam/ceresoptimizer.cpp:28-108),
// PoseGraph3dErrorTerm::operator() (tbv_slam/include/tbv_slam/ceresoptimizer.h:56-82) with its AutoDiff Jacobian,
// EigenQuaternionParameterization's plus-Jacobian, CauchyLoss(0.1) on loop constraints and the Ceres corrector — i.e. what
// the first evaluation of ceres::Solve computes before the sparse Cholesky (which stays on the host).
//
//   pgo_blocks    one thread per constraint: sqrt-information L (Cholesky of the 6x6 information), residual r = L e,
//                 Ja, Jb (6x6 each, tangent space), robustification; writes Ja^T Ja, Jb^T Jb, Ja^T Jb, Ja^T r, Jb^T r, cost.
//   pgo_gather    one warp per node: sums the diagonal blocks / gradient of its incident constraints in the reference's
//                 order (all odometry constraints in index order, then all loop constraints — AddConstraintType order),
//                 through a CSR built on the host from the id list (index bookkeeping only).  Deterministic, no atomics.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cooperative_groups.h>
#include <numeric>

#include "tbv_common.cuh"

namespace tbv {

__device__ __forceinline__ void quat_rotate(const double q[4], const double v[3], double o[3]) {  // Eigen Quaternion::_transformVector
  const double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  const double uvx = ux + ux, uvy = uy + uy, uvz = uz + uz;
  o[0] = v[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  o[1] = v[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  o[2] = v[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
__device__ __forceinline__ void quat_mul(const double p[4], const double q[4], double o[4]) {  // Hamilton product, (x,y,z,w)
  o[0] = p[3] * q[0] + p[0] * q[3] + p[1] * q[2] - p[2] * q[1];
  o[1] = p[3] * q[1] - p[0] * q[2] + p[1] * q[3] + p[2] * q[0];
  o[2] = p[3] * q[2] + p[0] * q[1] - p[1] * q[0] + p[2] * q[3];
  o[3] = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
}

constexpr int PGB = 36 * 3 + 12 + 1;  // per-constraint record: AtA, BtB, (unused), Atr, Btr, cost

__global__ void __launch_bounds__(64)
pgo_blocks(int n_con, const double* __restrict__ nodes, const int* __restrict__ ids, const double* __restrict__ meas, const double* __restrict__ info,
           tbv_pgo_params P, int fixed_node, double* __restrict__ rec, double* __restrict__ H_off, double* __restrict__ residuals, int* __restrict__ err) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_con) return;
  const int ia = ids[3 * ci], ib = ids[3 * ci + 1], type = ids[3 * ci + 2];
  const double* A = nodes + 7 * (size_t)ia;
  const double* B = nodes + 7 * (size_t)ib;
  const double* M = meas + 7 * (size_t)ci;
  // ---- sqrt information (ceresoptimizer.cpp:83-100)
  double L[36];
  {
    double I[36];
    const double lsf = (type == 1) ? 1.0 / P.loop_scaling : 1.0;
    if (P.replace_cov_by_identity) {
      const double diag[6] = {1.0 / P.odom_vxx, 1.0 / P.odom_vyy, 1, 1, 1, 1.0 / P.odom_vtt};
      for (int i = 0; i < 36; i++) I[i] = 0;
      for (int i = 0; i < 6; i++) I[i * 6 + i] = 1.0 * diag[i] * lsf;
    } else {
      for (int i = 0; i < 36; i++) I[i] = info[36 * (size_t)ci + i] * lsf;
    }
    for (int i = 0; i < 36; i++) L[i] = 0;
    bool ok = true;
    for (int j = 0; j < 6 && ok; j++) {
      double d = I[j * 6 + j];
      for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
      if (!(d > 0)) { ok = false; break; }
      L[j * 6 + j] = sqrt(d);
      for (int i = j + 1; i < 6; i++) {
        double v = I[i * 6 + j];
        for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k];
        L[i * 6 + j] = v / L[j * 6 + j];
      }
    }
    if (!ok) *err = 1;
  }
  // ---- residual
  const double qa_inv[4] = {-A[3], -A[4], -A[5], A[6]};
  const double d[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  double p_ab[3], q_ab[4], dq[4];
  quat_rotate(qa_inv, d, p_ab);
  quat_mul(qa_inv, B + 3, q_ab);
  const double q_ab_conj[4] = {-q_ab[0], -q_ab[1], -q_ab[2], q_ab[3]};
  quat_mul(M + 3, q_ab_conj, dq);
  const double e[6] = {p_ab[0] - M[0], p_ab[1] - M[1], p_ab[2] - M[2], 2.0 * dq[0], 2.0 * dq[1], 2.0 * dq[2]};
  double r[6];
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += L[i * 6 + k] * e[k];
    r[i] = s;
  }
  // ---- ambient derivatives of e, then tangent space
  const double u[3] = {qa_inv[0], qa_inv[1], qa_inv[2]}, w = qa_inv[3];
  double Rm[3][3], dP_du[3][3], dP_dw[3];
  {
    const double ux[3][3] = {{0, -u[2], u[1]}, {u[2], 0, -u[0]}, {-u[1], u[0], 0}};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double uu = 0;
        for (int k = 0; k < 3; k++) uu += ux[i][k] * ux[k][j];
        Rm[i][j] = (i == j ? 1.0 : 0.0) + 2.0 * w * ux[i][j] + 2.0 * uu;
      }
    dP_dw[0] = 2.0 * (u[1] * d[2] - u[2] * d[1]);
    dP_dw[1] = 2.0 * (u[2] * d[0] - u[0] * d[2]);
    dP_dw[2] = 2.0 * (u[0] * d[1] - u[1] * d[0]);
    const double dx[3][3] = {{0, -d[2], d[1]}, {d[2], 0, -d[0]}, {-d[1], d[0], 0}};
    const double ud = u[0] * d[0] + u[1] * d[1] + u[2] * d[2];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) dP_du[i][j] = -2.0 * w * dx[i][j] + 2.0 * ((i == j ? ud : 0.0) + u[i] * d[j] - 2.0 * d[i] * u[j]);
  }
  const double cb[4] = {-B[3], -B[4], -B[5], B[6]};
  double Mq[4];
  quat_mul(M + 3, cb, Mq);
  const double LM[3][4] = {{Mq[3], -Mq[2], Mq[1], Mq[0]}, {Mq[2], Mq[3], -Mq[0], Mq[1]}, {-Mq[1], Mq[0], Mq[3], Mq[2]}};
  const double* m = M + 3;
  const double Lm[4][4] = {{m[3], -m[2], m[1], m[0]}, {m[2], m[3], -m[0], m[1]}, {-m[1], m[0], m[3], m[2]}, {-m[0], -m[1], -m[2], m[3]}};
  const double* a = A + 3;
  const double Ra[4][4] = {{a[3], a[2], -a[1], a[0]}, {-a[2], a[3], a[0], a[1]}, {a[1], -a[0], a[3], a[2]}, {-a[0], -a[1], -a[2], a[3]}};
  double dD_db[3][4];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += Lm[i][k] * Ra[k][j];
      dD_db[i][j] = s * (j < 3 ? -1.0 : 1.0);
    }
  double Eqa[6][4], Eqb[6][4];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 4; j++) { Eqa[i][j] = 0; Eqb[i][j] = 0; }
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Eqa[i][j] = -dP_du[i][j];
    Eqa[i][3] = dP_dw[i];
    for (int j = 0; j < 4; j++) { Eqa[3 + i][j] = 2.0 * LM[i][j]; Eqb[3 + i][j] = 2.0 * dD_db[i][j]; }
  }
  const double* bq = B + 3;
  const double Pa[4][3] = {{a[3], a[2], -a[1]}, {-a[2], a[3], a[0]}, {a[1], -a[0], a[3]}, {-a[0], -a[1], -a[2]}};
  const double Pb[4][3] = {{bq[3], bq[2], -bq[1]}, {-bq[2], bq[3], bq[0]}, {bq[1], -bq[0], bq[3]}, {-bq[0], -bq[1], -bq[2]}};
  double Ea[6][6], Eb[6][6];
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 3; j++) {
      Ea[i][j] = i < 3 ? -Rm[i][j] : 0.0;
      Eb[i][j] = i < 3 ? Rm[i][j] : 0.0;
      double sa = 0, sb = 0;
      for (int k = 0; k < 4; k++) { sa += Eqa[i][k] * Pa[k][j]; sb += Eqb[i][k] * Pb[k][j]; }
      Ea[i][3 + j] = sa; Eb[i][3 + j] = sb;
    }
  }
  double Ja[36], Jb[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double sa = 0, sb = 0;
      for (int k = 0; k < 6; k++) { sa += L[i * 6 + k] * Ea[k][j]; sb += L[i * 6 + k] * Eb[k][j]; }
      Ja[i * 6 + j] = sa; Jb[i * 6 + j] = sb;
    }
  // ---- loss + corrector (CauchyLoss on loop constraints: rho'' < 0 -> sqrt(rho') scaling)
  double sq = 0;
  for (int i = 0; i < 6; i++) sq += r[i] * r[i];
  double cost;
  if (type == 1) {
    const double b = P.loop_cauchy * P.loop_cauchy, cc = 1.0 / b;
    const double sum = 1.0 + sq * cc, inv = 1.0 / sum;
    const double rho0 = b * log(sum), rho1 = fmax(DBL_MIN, inv);
    cost = 0.5 * rho0;
    const double s1 = sqrt(rho1);
    for (int i = 0; i < 36; i++) { Ja[i] *= s1; Jb[i] *= s1; }
    for (int i = 0; i < 6; i++) r[i] *= s1;
  } else {
    cost = 0.5 * sq;
  }
  if (residuals) for (int i = 0; i < 6; i++) residuals[6 * (size_t)ci + i] = r[i];
  const bool fa = (ia == fixed_node), fb = (ib == fixed_node);
  double* o = rec + (size_t)ci * PGB;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double aa = 0, bb = 0, ab = 0;
      for (int k = 0; k < 6; k++) { aa += Ja[k * 6 + i] * Ja[k * 6 + j]; bb += Jb[k * 6 + i] * Jb[k * 6 + j]; ab += Ja[k * 6 + i] * Jb[k * 6 + j]; }
      o[i * 6 + j] = fa ? 0.0 : aa;
      o[36 + i * 6 + j] = fb ? 0.0 : bb;
      H_off[36 * (size_t)ci + i * 6 + j] = (fa || fb) ? 0.0 : ab;
    }
  for (int i = 0; i < 6; i++) {
    double ga = 0, gb = 0;
    for (int k = 0; k < 6; k++) { ga += Ja[k * 6 + i] * r[k]; gb += Jb[k * 6 + i] * r[k]; }
    o[108 + i] = fa ? 0.0 : ga;
    o[114 + i] = fb ? 0.0 : gb;
  }
  o[120] = cost;
}

// one warp per node; inc[] lists (constraint << 1 | side) in reference order, side 0 = begin (Ja), 1 = end (Jb)
__global__ void __launch_bounds__(128)
pgo_gather(int n_nodes, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ rec, double* __restrict__ H_diag,
           double* __restrict__ g) {
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (node >= n_nodes) return;
  double h0 = 0, h1 = 0, gg = 0;  // lane handles H entries lane and lane+32 (< 36) and g entry lane (< 6)
  for (int t = row[node]; t < row[node + 1]; t++) {
    const int v = inc[t];
    const double* o = rec + (size_t)(v >> 1) * PGB;
    const int side = v & 1;
    h0 += o[side * 36 + lane];
    if (lane < 4) h1 += o[side * 36 + 32 + lane];
    if (lane < 6) gg += o[108 + side * 6 + lane];
  }
  H_diag[36 * (size_t)node + lane] = h0;
  if (lane < 4) H_diag[36 * (size_t)node + 32 + lane] = h1;
  if (lane < 6) g[6 * (size_t)node + lane] = gg;
}

// cost = sum over constraints in reference order (odometry first): single CTA, fixed-shape tree
__global__ void __launch_bounds__(256)
pgo_cost(int n_con, const int* __restrict__ order, const double* __restrict__ rec, double* __restrict__ cost) {
  __shared__ double s[256];
  double a = 0;
  for (int i = threadIdx.x; i < n_con; i += 256) a += rec[(size_t)order[i] * PGB + 120];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) *cost = s[0];
}

// ---- one Levenberg-Marquardt step of the pose graph on the device (first part) ------------------------------------------
// Solves (H + D) delta = -g for the block-sparse normal equations tbv_pgo_assemble produces (H_diag: 6x6 per node, H_off: block
// (begin, end) per constraint, symmetric), D = clamp(diag(H), 1e-6, 1e32) / radius — the system Ceres' LEVENBERG_MARQUARDT strategy
// hands to SPARSE_NORMAL_CHOLESKY (ceresoptimizer.cpp:50-62) — by conjugate gradients with the 6x6 block-Jacobi preconditioner.
// ONE persistent CTA: the graph is small (4.5 k nodes, 27 k unknowns, 1.4 MB of blocks: L2-resident), an iteration is a block-sparse
// product plus two dot products, and keeping it in one CTA makes every reduction a fixed-shape tree (deterministic) with no grid sync.
constexpr int PCG_THREADS = 1024;

__device__ __forceinline__ double pcg_block_sum(double v, double* s_w) {  // fixed tree: lanes, then the 32 warp totals in order
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();            // s_w may still be read from the previous reduction
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < PCG_THREADS / 32; w++) t += s_w[w];
  return t;
}

__global__ void __launch_bounds__(PCG_THREADS, 1)
pgo_pcg(int n, int fixed_node, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
        const double* __restrict__ Ho, const double* __restrict__ g, double radius, int max_iters, double rel_tol, double* __restrict__ x,
        double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ q, double* __restrict__ Minv, double* __restrict__ Dg,
        int* __restrict__ out_iters, double* __restrict__ out_rel) {
  __shared__ double s_w[PCG_THREADS / 32];
  const int tid = threadIdx.x;
  // ---- damped diagonal blocks and their inverses (Cholesky of the SPD 6x6, then L^-T L^-1) ------------------------------------------
  for (int i = tid; i < n; i += PCG_THREADS) {
    double A[6][6], L[6][6], Li[6][6];
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) A[a][b] = Hd[36 * (size_t)i + a * 6 + b];
    bool ok = i != fixed_node;
    for (int a = 0; a < 6; a++) {
      const double d = fmin(fmax(A[a][a], 1e-6), 1e32) / radius;
      Dg[6 * (size_t)i + a] = d;
      A[a][a] += d;
    }
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) { L[a][b] = 0.0; Li[a][b] = 0.0; }
    for (int j = 0; j < 6 && ok; j++) {
      double d = A[j][j];
      for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
      if (!(d > 0.0)) { ok = false; break; }
      L[j][j] = sqrt(d);
      for (int a = j + 1; a < 6; a++) {
        double v = A[a][j];
        for (int k = 0; k < j; k++) v -= L[a][k] * L[j][k];
        L[a][j] = v / L[j][j];
      }
    }
    if (ok) {
      for (int c = 0; c < 6; c++) {            // Li = L^-1 by forward substitution on the identity
        for (int a = 0; a < 6; a++) {
          double v = (a == c) ? 1.0 : 0.0;
          for (int k = 0; k < a; k++) v -= L[a][k] * Li[k][c];
          Li[a][c] = v / L[a][a];
        }
      }
    }
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) {
        double v = 0.0;
        if (ok) for (int k = 0; k < 6; k++) v += Li[k][a] * Li[k][b];   // (L L^T)^-1 = L^-T L^-1
        Minv[36 * (size_t)i + a * 6 + b] = v;
      }
  }
  __syncthreads();
  auto apply_Minv = [&](const double* v, double* o) {
    for (int i = tid; i < n; i += PCG_THREADS) {
      double t[6];
      for (int a = 0; a < 6; a++) t[a] = v[6 * (size_t)i + a];
      for (int a = 0; a < 6; a++) {
        double acc = 0.0;
        for (int b = 0; b < 6; b++) acc += Minv[36 * (size_t)i + a * 6 + b] * t[b];
        o[6 * (size_t)i + a] = acc;
      }
    }
  };
  // q = (H + D) v: the node's own block, then its constraints in the order of `inc` (side 0: this node begins the constraint -> H_off v_end;
  // side 1: it ends it -> H_off^T v_begin)
  auto apply_A = [&](const double* v, double* o) {
    for (int i = tid; i < n; i += PCG_THREADS) {
      double acc[6] = {0, 0, 0, 0, 0, 0};
      if (i != fixed_node) {
        double t[6];
        for (int a = 0; a < 6; a++) t[a] = v[6 * (size_t)i + a];
        for (int a = 0; a < 6; a++) {
          double s0 = Dg[6 * (size_t)i + a] * t[a];
          for (int b = 0; b < 6; b++) s0 += Hd[36 * (size_t)i + a * 6 + b] * t[b];
          acc[a] = s0;
        }
        for (int e = row[i]; e < row[i + 1]; e++) {
          const int c = inc[e] >> 1, side = inc[e] & 1;
          const int other = side ? ids[3 * c] : ids[3 * c + 1];
          const double* B = Ho + 36 * (size_t)c;
          double u[6];
          for (int a = 0; a < 6; a++) u[a] = v[6 * (size_t)other + a];
          if (side == 0) { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) acc[a] += B[a * 6 + b] * u[b]; }
          else           { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) acc[a] += B[b * 6 + a] * u[b]; }
        }
      }
      for (int a = 0; a < 6; a++) o[6 * (size_t)i + a] = acc[a];
    }
  };
  auto dot = [&](const double* u, const double* v) {
    double acc = 0.0;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) acc += u[6 * (size_t)i + a] * v[6 * (size_t)i + a];
    return pcg_block_sum(acc, s_w);
  };
  // ---- x = 0, r = b = -g (the fixed node's rows are zero), z = M^-1 r, p = z ---------------------------------------------------------
  for (int i = tid; i < n; i += PCG_THREADS)
    for (int a = 0; a < 6; a++) {
      x[6 * (size_t)i + a] = 0.0;
      r[6 * (size_t)i + a] = (i == fixed_node) ? 0.0 : -g[6 * (size_t)i + a];
    }
  __syncthreads();
  apply_Minv(r, z);
  for (int i = tid; i < n; i += PCG_THREADS)
    for (int a = 0; a < 6; a++) p[6 * (size_t)i + a] = z[6 * (size_t)i + a];   // own elements only: no barrier needed before
  __syncthreads();
  double rz = dot(r, z);
  const double bnorm = sqrt(dot(r, r));
  double rel = bnorm > 0.0 ? 1.0 : 0.0;
  int it = 0;
  while (it < max_iters && rel > rel_tol) {
    apply_A(p, q);
    __syncthreads();
    const double pq = dot(p, q);
    if (!(pq > 0.0)) break;                      // lost positive definiteness (never with D > 0): stop with what we have
    const double alpha = rz / pq;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) {
        x[6 * (size_t)i + a] += alpha * p[6 * (size_t)i + a];
        r[6 * (size_t)i + a] -= alpha * q[6 * (size_t)i + a];
      }
    apply_Minv(r, z);                            // own elements of r: written by this thread just above
    __syncthreads();
    const double rz_new = dot(r, z);
    rel = sqrt(dot(r, r)) / bnorm;
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) p[6 * (size_t)i + a] = z[6 * (size_t)i + a] + beta * p[6 * (size_t)i + a];
    __syncthreads();
    it++;
  }
  if (tid == 0) { *out_iters = it; *out_rel = rel; }
}

// ---- the same solve on a thread-block CLUSTER (opt-in: TBV_PGO_CLUSTER=1; not yet run on a GPU — see DESIGN.md §7b) -----------------------
// pgo_pcg is bound by one SM's latency chain (35 us per CG iteration measured at 600 nodes).  Here the nodes are split into contiguous ranges over
// the PCGC_CL CTAs of one cluster, ONE ROW (node, component) PER THREAD, so an iteration is: a block-sparse product whose rows read 6-element
// slices (blocks and vectors stream from L2, 1/PCGC_CL of them per SM), three cluster barriers (380 cycles each on this part) and two reductions
// whose per-CTA partials are exchanged through distributed shared memory and summed in rank order by every CTA — the same bits everywhere, so
// all CTAs take the same branch and the result does not depend on scheduling.  All 6 rows of a node live in one CTA: z = M^-1 r needs only a
// __syncthreads.
constexpr int PCGC_CL = 8;          // portable cluster size
constexpr int PCGC_THREADS = 1024;

__device__ __forceinline__ void pcgc_block_sum2(double& a, double& b, double (*s_w)[2]) {   // fixed tree per CTA; both values at once
  for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { s_w[warp][0] = a; s_w[warp][1] = b; }
  __syncthreads();
  double ta = 0.0, tb = 0.0;
  for (int w = 0; w < PCGC_THREADS / 32; w++) { ta += s_w[w][0]; tb += s_w[w][1]; }
  a = ta; b = tb;
}

__global__ void __cluster_dims__(PCGC_CL, 1, 1) __launch_bounds__(PCGC_THREADS, 1)
pgo_pcg_cluster(int n, int fixed_node, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
                const double* __restrict__ Ho, const double* __restrict__ g, double radius, int max_iters, double rel_tol, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ z, double* p, double* __restrict__ q, double* __restrict__ Minv, double* __restrict__ Dg,
                int* __restrict__ out_iters, double* __restrict__ out_rel) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_w[PCGC_THREADS / 32][2];
  __shared__ double s_part[2][2];                    // [0]: (p.q, -)   [1]: (r.z, r.r) — this CTA's partials, read by the whole cluster
  const int tid = threadIdx.x, rank = (int)cluster.block_rank();
  const int npc = (n + PCGC_CL - 1) / PCGC_CL;        // nodes per CTA
  const int n0 = min(rank * npc, n), n1 = min(n0 + npc, n);
  const int rows0 = 6 * n0, nrows = 6 * (n1 - n0);

  // cluster-wide sums of (a, b): per-CTA fixed tree, then every CTA adds the PCGC_CL partials in rank order (DSMEM reads)
  auto cluster_sum2 = [&](double& a, double& b, int slot) {
    pcgc_block_sum2(a, b, s_w);
    if (tid == 0) { s_part[slot][0] = a; s_part[slot][1] = b; }
    cluster.sync();
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < PCGC_CL; k++) {
      const double* remote = cluster.map_shared_rank(&s_part[slot][0], k);
      ta += remote[0]; tb += remote[1];
    }
    a = ta; b = tb;
  };

  // ---- setup, one thread per owned node: damped diagonal block, its inverse, r = -g, z = M^-1 r, p = z, x = 0 --------------------------------------
  for (int i = n0 + tid; i < n1; i += PCGC_THREADS) {
    double A[6][6], L[6][6], Li[6][6];
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) { A[a][b] = Hd[36 * (size_t)i + a * 6 + b]; L[a][b] = 0.0; Li[a][b] = 0.0; }
    bool ok = i != fixed_node;
    for (int a = 0; a < 6; a++) {
      const double d = fmin(fmax(A[a][a], 1e-6), 1e32) / radius;
      Dg[6 * (size_t)i + a] = d;
      A[a][a] += d;
    }
    for (int j = 0; j < 6 && ok; j++) {
      double d = A[j][j];
      for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
      if (!(d > 0.0)) { ok = false; break; }
      L[j][j] = sqrt(d);
      for (int a = j + 1; a < 6; a++) {
        double v = A[a][j];
        for (int k = 0; k < j; k++) v -= L[a][k] * L[j][k];
        L[a][j] = v / L[j][j];
      }
    }
    if (ok)
      for (int c = 0; c < 6; c++)
        for (int a = 0; a < 6; a++) {
          double v = (a == c) ? 1.0 : 0.0;
          for (int k = 0; k < a; k++) v -= L[a][k] * Li[k][c];
          Li[a][c] = v / L[a][a];
        }
    double rr[6];
    for (int a = 0; a < 6; a++) rr[a] = (i == fixed_node) ? 0.0 : -g[6 * (size_t)i + a];
    for (int a = 0; a < 6; a++) {
      double zz = 0.0;
      for (int b = 0; b < 6; b++) {
        double v = 0.0;
        if (ok) for (int k = 0; k < 6; k++) v += Li[k][a] * Li[k][b];
        Minv[36 * (size_t)i + a * 6 + b] = v;
        zz += v * rr[b];
      }
      x[6 * (size_t)i + a] = 0.0;
      r[6 * (size_t)i + a] = rr[a];
      z[6 * (size_t)i + a] = zz;
      p[6 * (size_t)i + a] = zz;
    }
  }
  __syncthreads();
  double rz = 0.0, bb = 0.0;
  for (int lr = tid; lr < nrows; lr += PCGC_THREADS) { rz += r[rows0 + lr] * z[rows0 + lr]; bb += r[rows0 + lr] * r[rows0 + lr]; }
  cluster_sum2(rz, bb, 1);                           // also publishes p to the cluster (barrier.cluster release / acquire)
  const double bnorm = sqrt(bb);
  double rel = bnorm > 0.0 ? 1.0 : 0.0;
  int it = 0;
  while (it < max_iters && rel > rel_tol) {
    // q = (H + D) p, own rows; p.q
    double pq = 0.0, unused = 0.0;
    for (int lr = tid; lr < nrows; lr += PCGC_THREADS) {
      const int R = rows0 + lr, i = R / 6, a = R - 6 * i;
      double acc = 0.0;
      if (i != fixed_node) {
        const double* Hr = Hd + 36 * (size_t)i + 6 * a;
        const double* pi = p + 6 * (size_t)i;
        acc = Dg[R] * pi[a];
        for (int b = 0; b < 6; b++) acc += Hr[b] * pi[b];
        for (int e = row[i]; e < row[i + 1]; e++) {
          const int c = inc[e] >> 1, side = inc[e] & 1;
          const int other = side ? ids[3 * c] : ids[3 * c + 1];
          const double* B = Ho + 36 * (size_t)c;
          const double* u = p + 6 * (size_t)other;
          // the other node may belong to another CTA: read its slice of p from L2 (the cluster barrier orders the writes; no stale L1 line)
          if (side == 0) { for (int b = 0; b < 6; b++) acc += B[6 * a + b] * __ldcg(u + b); }
          else           { for (int b = 0; b < 6; b++) acc += B[6 * b + a] * __ldcg(u + b); }
        }
      }
      q[R] = acc;
      pq += p[R] * acc;
    }
    cluster_sum2(pq, unused, 0);
    if (!(pq > 0.0)) break;                          // same value in every CTA: the whole cluster leaves together
    const double alpha = rz / pq;
    for (int lr = tid; lr < nrows; lr += PCGC_THREADS) {
      const int R = rows0 + lr;
      x[R] += alpha * p[R];
      r[R] -= alpha * q[R];
    }
    __syncthreads();                                 // z needs the node's six residual rows (same CTA)
    double rz_new = 0.0, rr_new = 0.0;
    for (int lr = tid; lr < nrows; lr += PCGC_THREADS) {
      const int R = rows0 + lr, i = R / 6, a = R - 6 * i;
      const double* Mr = Minv + 36 * (size_t)i + 6 * a;
      const double* ri = r + 6 * (size_t)i;
      double acc = 0.0;
      for (int b = 0; b < 6; b++) acc += Mr[b] * ri[b];
      z[R] = acc;
      rz_new += ri[a] * acc;
      rr_new += ri[a] * ri[a];
    }
    cluster_sum2(rz_new, rr_new, 1);
    rel = sqrt(rr_new) / bnorm;
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int lr = tid; lr < nrows; lr += PCGC_THREADS) {
      const int R = rows0 + lr;
      p[R] = z[R] + beta * p[R];
    }
    cluster.sync();                                  // the next product reads other CTAs' rows of p
    it++;
  }
  cluster.sync();                                    // no CTA may exit while another still reads its partials through DSMEM
  if (rank == 0 && tid == 0) { *out_iters = it; *out_rel = rel; }
}

// ---- the same solve with the ODOMETRY CHAIN as preconditioner (opt-in: TBV_PGO_CHAIN=1; not yet run on a GPU — see DESIGN.md §7b) ----------------
// M = block-tridiagonal part of (H + D): diagonal blocks + the blocks coupling nodes i and i + 1 (chain[i] = A[i+1][i], summed on the host from the
// constraints between consecutive nodes), factorised once as M = L S L^T (block Thomas: S_0 = A_00, W_i = chain[i-1] S_{i-1}^-1,
// S_i = A_ii - W_i chain[i-1]^T).  A pose graph is that chain plus a few weak loop blocks: CG needs ~5 iterations where block-Jacobi needs ~1100
// (measured with the numpy prototype tests/tools/pgo_chain_prototype.py, which restates this kernel operation by operation and is its checker).
// First version: factorisation and the two sweeps of every application run on ONE thread (dependent 6x6 recurrences; the blocks stream from L2);
// the product and the block-diagonal solve use the whole CTA.  Next: 6 lanes per recurrence, then parallel cyclic reduction.
__device__ void pcgc_inverse_spd6(const double* S, double* Sinv, bool* ok) {   // Cholesky S = L L^T, then S^-1 = L^-T L^-1
  double L[6][6], Li[6][6];
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) { L[a][b] = 0.0; Li[a][b] = 0.0; }
  *ok = true;
  for (int j = 0; j < 6; j++) {
    double d = S[6 * j + j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) { *ok = false; d = 1.0; }           // not positive definite: keep going with a unit pivot, the caller reports it
    L[j][j] = sqrt(d);
    for (int a = j + 1; a < 6; a++) {
      double v = S[6 * a + j];
      for (int k = 0; k < j; k++) v -= L[a][k] * L[j][k];
      L[a][j] = v / L[j][j];
    }
  }
  for (int c = 0; c < 6; c++)
    for (int a = 0; a < 6; a++) {
      double v = (a == c) ? 1.0 : 0.0;
      for (int k = 0; k < a; k++) v -= L[a][k] * Li[k][c];
      Li[a][c] = v / L[a][a];
    }
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      double v = 0.0;
      for (int k = 0; k < 6; k++) v += Li[k][a] * Li[k][b];
      Sinv[6 * a + b] = v;
    }
}

__global__ void __launch_bounds__(PCG_THREADS, 1)
pgo_pcg_chain(int n, int fixed_node, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
              const double* __restrict__ Ho, const double* __restrict__ chain, const double* __restrict__ g, double radius, int max_iters, double rel_tol,
              double* __restrict__ x, double* r, double* z, double* __restrict__ p, double* __restrict__ q, double* Sinv, double* Wb, double* __restrict__ Dg,
              double* y, int* __restrict__ out_iters, double* __restrict__ out_rel) {
  __shared__ double s_w[PCG_THREADS / 32];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int i = tid; i < n; i += PCG_THREADS)
    for (int a = 0; a < 6; a++) Dg[6 * (size_t)i + a] = fmin(fmax(Hd[36 * (size_t)i + 7 * a], 1e-6), 1e32) / radius;
  __syncthreads();
  // ---- block Thomas factorisation of the chain (sequential in i) -----------------------------------------------------------------------------------
  if (tid == 0) {
    double Sp[36];                                        // S_{i-1}^-1
    for (int i = 0; i < n; i++) {
      double S[36];
      for (int e = 0; e < 36; e++) S[e] = (i == fixed_node) ? ((e % 7 == 0) ? 1.0 : 0.0) : Hd[36 * (size_t)i + e];
      if (i != fixed_node)
        for (int a = 0; a < 6; a++) S[7 * a] += Dg[6 * (size_t)i + a];
      if (i > 0) {
        const double* C = chain + 36 * (size_t)(i - 1);   // A[i][i-1]; zero on both sides of the fixed node (host)
        double W[36];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 6; b++) {
            double v = 0.0;
            for (int k = 0; k < 6; k++) v += C[6 * a + k] * Sp[6 * k + b];
            W[6 * a + b] = v;
          }
        for (int e = 0; e < 36; e++) Wb[36 * (size_t)(i - 1) + e] = W[e];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 6; b++) {
            double v = 0.0;
            for (int k = 0; k < 6; k++) v += W[6 * a + k] * C[6 * b + k];
            S[6 * a + b] -= v;
          }
      }
      bool ok;
      pcgc_inverse_spd6(S, Sp, &ok);
      if (!ok) s_bad = 1;
      for (int e = 0; e < 36; e++) Sinv[36 * (size_t)i + e] = Sp[e];
    }
  }
  __syncthreads();
  // z = M^-1 v: forward sweep (thread 0), block-diagonal solve (all threads), backward sweep (thread 0)
  auto apply_chain = [&](const double* v, double* o) {
    if (tid == 0) {
      double prev[6], cur[6];
      for (int a = 0; a < 6; a++) { prev[a] = v[a]; y[a] = prev[a]; }
      for (int i = 1; i < n; i++) {
        const double* W = Wb + 36 * (size_t)(i - 1);
        for (int a = 0; a < 6; a++) {
          double acc = v[6 * (size_t)i + a];
          for (int b = 0; b < 6; b++) acc -= W[6 * a + b] * prev[b];
          cur[a] = acc;
        }
        for (int a = 0; a < 6; a++) { prev[a] = cur[a]; y[6 * (size_t)i + a] = cur[a]; }
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) {
        double acc = 0.0;
        for (int b = 0; b < 6; b++) acc += Sinv[36 * (size_t)i + 6 * a + b] * y[6 * (size_t)i + b];
        o[6 * (size_t)i + a] = acc;
      }
    __syncthreads();
    if (tid == 0) {
      double nxt[6], cur[6];
      for (int a = 0; a < 6; a++) nxt[a] = o[6 * (size_t)(n - 1) + a];
      for (int i = n - 2; i >= 0; i--) {
        const double* W = Wb + 36 * (size_t)i;            // W_{i+1}: o_i -= W_{i+1}^T o_{i+1}
        for (int a = 0; a < 6; a++) {
          double acc = o[6 * (size_t)i + a];
          for (int b = 0; b < 6; b++) acc -= W[6 * b + a] * nxt[b];
          cur[a] = acc;
        }
        for (int a = 0; a < 6; a++) { nxt[a] = cur[a]; o[6 * (size_t)i + a] = cur[a]; }
      }
      if (fixed_node >= 0 && fixed_node < n)
        for (int a = 0; a < 6; a++) o[6 * (size_t)fixed_node + a] = 0.0;
    }
    __syncthreads();
  };
  auto apply_A = [&](const double* v, double* o) {        // as in pgo_pcg
    for (int i = tid; i < n; i += PCG_THREADS) {
      double acc[6] = {0, 0, 0, 0, 0, 0};
      if (i != fixed_node) {
        double t[6];
        for (int a = 0; a < 6; a++) t[a] = v[6 * (size_t)i + a];
        for (int a = 0; a < 6; a++) {
          double s0 = Dg[6 * (size_t)i + a] * t[a];
          for (int b = 0; b < 6; b++) s0 += Hd[36 * (size_t)i + a * 6 + b] * t[b];
          acc[a] = s0;
        }
        for (int e = row[i]; e < row[i + 1]; e++) {
          const int c = inc[e] >> 1, side = inc[e] & 1;
          const int other = side ? ids[3 * c] : ids[3 * c + 1];
          const double* B = Ho + 36 * (size_t)c;
          double u[6];
          for (int a = 0; a < 6; a++) u[a] = v[6 * (size_t)other + a];
          if (side == 0) { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) acc[a] += B[a * 6 + b] * u[b]; }
          else           { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) acc[a] += B[b * 6 + a] * u[b]; }
        }
      }
      for (int a = 0; a < 6; a++) o[6 * (size_t)i + a] = acc[a];
    }
  };
  auto dot = [&](const double* u, const double* v) {
    double acc = 0.0;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) acc += u[6 * (size_t)i + a] * v[6 * (size_t)i + a];
    return pcg_block_sum(acc, s_w);
  };
  for (int i = tid; i < n; i += PCG_THREADS)
    for (int a = 0; a < 6; a++) {
      x[6 * (size_t)i + a] = 0.0;
      r[6 * (size_t)i + a] = (i == fixed_node) ? 0.0 : -g[6 * (size_t)i + a];
    }
  __syncthreads();
  apply_chain(r, z);
  for (int i = tid; i < n; i += PCG_THREADS)
    for (int a = 0; a < 6; a++) p[6 * (size_t)i + a] = z[6 * (size_t)i + a];
  __syncthreads();
  double rz = dot(r, z);
  const double bnorm = sqrt(dot(r, r));
  double rel = bnorm > 0.0 ? 1.0 : 0.0;
  int it = 0;
  while (it < max_iters && rel > rel_tol) {
    apply_A(p, q);
    __syncthreads();
    const double pq = dot(p, q);
    if (!(pq > 0.0)) break;
    const double alpha = rz / pq;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) {
        x[6 * (size_t)i + a] += alpha * p[6 * (size_t)i + a];
        r[6 * (size_t)i + a] -= alpha * q[6 * (size_t)i + a];
      }
    __syncthreads();                                     // the sweeps read every row of r
    apply_chain(r, z);
    const double rz_new = dot(r, z);
    rel = sqrt(dot(r, r)) / bnorm;
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = tid; i < n; i += PCG_THREADS)
      for (int a = 0; a < 6; a++) p[6 * (size_t)i + a] = z[6 * (size_t)i + a] + beta * p[6 * (size_t)i + a];
    __syncthreads();
    it++;
  }
  if (tid == 0) { *out_iters = s_bad ? -it - 1 : it; *out_rel = rel; }   // negative: a chain pivot block was not positive definite
}

}  // namespace tbv

using namespace tbv;

extern "C" int tbv_pgo_assemble(tbv_ctx* ctx, int n_nodes, const double* nodes, int n_con, const int* ids, const double* meas, const double* info,
                                const tbv_pgo_params* params, int fixed_node, double* cost, double* H_diag, double* H_off, double* g,
                                double* residuals) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && nodes && ids && meas && params && H_diag && H_off && g && n_nodes >= 1 && n_con >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(params->replace_cov_by_identity || info, "information matrices required when replace_cov_by_identity is 0");
  for (int c = 0; c < n_con; c++)
    TBV_REQUIRE(ids[3 * c] >= 0 && ids[3 * c] < n_nodes && ids[3 * c + 1] >= 0 && ids[3 * c + 1] < n_nodes, "constraint references a missing node");
  // reference order: AddConstraintType(odometry) then AddConstraintType(loop) (ceresoptimizer.cpp:34-35)
  std::vector<int> order(n_con);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (ids[3 * a + 2] == 1) < (ids[3 * b + 2] == 1); });
  std::vector<int> row(n_nodes + 1, 0), inc(2 * (size_t)n_con);
  for (int c = 0; c < n_con; c++) { row[ids[3 * c] + 1]++; row[ids[3 * c + 1] + 1]++; }
  for (int i = 0; i < n_nodes; i++) row[i + 1] += row[i];
  {
    std::vector<int> fill(row.begin(), row.end() - 1);
    for (int c : order) { inc[fill[ids[3 * c]]++] = (c << 1); inc[fill[ids[3 * c + 1]]++] = (c << 1) | 1; }
  }
  DevBuf<double> dn, dm, di, drec, dho, dhd, dg, dres, dcost;
  DevBuf<int> dids, drow, dinc, dord, derr;
  auto cleanup = [&]() {
    dn.release(); dm.release(); di.release(); drec.release(); dho.release(); dhd.release(); dg.release(); dres.release(); dcost.release();
    dids.release(); drow.release(); dinc.release(); dord.release(); derr.release();
  };
  const size_t nc1 = n_con ? n_con : 1;
  int rc;
  if ((rc = dn.reserve(7 * (size_t)n_nodes)) || (rc = dm.reserve(7 * nc1)) || (rc = di.reserve(info ? 36 * nc1 : 1)) || (rc = drec.reserve(PGB * nc1)) ||
      (rc = dho.reserve(36 * nc1)) || (rc = dhd.reserve(36 * (size_t)n_nodes)) || (rc = dg.reserve(6 * (size_t)n_nodes)) || (rc = dres.reserve(6 * nc1)) ||
      (rc = dcost.reserve(1)) || (rc = dids.reserve(3 * nc1)) || (rc = drow.reserve(n_nodes + 1)) || (rc = dinc.reserve(2 * nc1)) ||
      (rc = dord.reserve(nc1)) || (rc = derr.reserve(1))) { cleanup(); return rc; }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(dn.p, nodes, 7 * (size_t)n_nodes * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dm.p, meas, 7 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con && info) e = cudaMemcpyAsync(di.p, info, 36 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dids.p, ids, 3 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(drow.p, row.data(), (n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dinc.p, inc.data(), 2 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dord.p, order.data(), (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(derr.p, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    if (n_con) {
      pgo_blocks<<<(n_con + 63) / 64, 64, 0, st>>>(n_con, dn.p, dids.p, dm.p, info ? di.p : nullptr, *params, fixed_node, drec.p, dho.p,
                                                   residuals ? dres.p : nullptr, derr.p);
      launched(ctx, "pgo_blocks");
    }
    pgo_gather<<<(n_nodes * 32 + 127) / 128, 128, 0, st>>>(n_nodes, drow.p, dinc.p, drec.p, dhd.p, dg.p);
    launched(ctx, "pgo_gather");
    pgo_cost<<<1, 256, 0, st>>>(n_con, dord.p, drec.p, dcost.p);
    launched(ctx, "pgo_cost");
    e = cudaGetLastError();
  }
  int herr = 0;
  double hcost = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(H_diag, dhd.p, 36 * (size_t)n_nodes * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g, dg.p, 6 * (size_t)n_nodes * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(H_off, dho.p, 36 * (size_t)n_con * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && n_con && residuals) e = cudaMemcpyAsync(residuals, dres.p, 6 * (size_t)n_con * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hcost, dcost.p, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_pgo_assemble: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  if (herr) { set_error("tbv_pgo_assemble: an information matrix is not positive definite"); return TBV_ERR_INVALID; }
  if (cost) *cost = hcost;
  return TBV_OK;
}


extern "C" int tbv_pgo_solve_step(tbv_ctx* ctx, int n_nodes, int n_con, const int* ids, const double* H_diag, const double* H_off, const double* g,
                                  int fixed_node, double radius, int max_iters, double rel_tol, double* delta, int* iters, double* rel_residual) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && ids && H_diag && H_off && g && delta && n_nodes >= 1 && n_con >= 0 && radius > 0 && max_iters >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  for (int c = 0; c < n_con; c++)
    TBV_REQUIRE(ids[3 * c] >= 0 && ids[3 * c] < n_nodes && ids[3 * c + 1] >= 0 && ids[3 * c + 1] < n_nodes, "constraint references a missing node");
  std::vector<int> row(n_nodes + 1, 0), inc(2 * (size_t)n_con + 1);
  for (int c = 0; c < n_con; c++) { row[ids[3 * c] + 1]++; row[ids[3 * c + 1] + 1]++; }
  for (int i = 0; i < n_nodes; i++) row[i + 1] += row[i];
  {
    std::vector<int> fill(row.begin(), row.end() - 1);
    for (int c = 0; c < n_con; c++) { inc[fill[ids[3 * c]]++] = (c << 1); inc[fill[ids[3 * c + 1]]++] = (c << 1) | 1; }
  }
  const size_t N6 = 6 * (size_t)n_nodes, nc1 = n_con ? n_con : 1;
  DevBuf<double> dhd, dho, dg, dx, dr, dz, dp, dq, dmi, ddg, drel;
  DevBuf<int> dids, drow, dinc, dit;
  auto cleanup = [&]() {
    dhd.release(); dho.release(); dg.release(); dx.release(); dr.release(); dz.release(); dp.release(); dq.release(); dmi.release(); ddg.release();
    drel.release(); dids.release(); drow.release(); dinc.release(); dit.release();
  };
  int rc;
  if ((rc = dhd.reserve(36 * (size_t)n_nodes)) || (rc = dho.reserve(36 * nc1)) || (rc = dg.reserve(N6)) || (rc = dx.reserve(N6)) || (rc = dr.reserve(N6)) ||
      (rc = dz.reserve(N6)) || (rc = dp.reserve(N6)) || (rc = dq.reserve(N6)) || (rc = dmi.reserve(36 * (size_t)n_nodes)) || (rc = ddg.reserve(N6)) ||
      (rc = drel.reserve(1)) || (rc = dids.reserve(3 * nc1)) || (rc = drow.reserve(n_nodes + 1)) || (rc = dinc.reserve(inc.size())) || (rc = dit.reserve(1))) {
    cleanup();
    return rc;
  }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(dhd.p, H_diag, 36 * (size_t)n_nodes * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dho.p, H_off, 36 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dg.p, g, N6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dids.p, ids, 3 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(drow.p, row.data(), (n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(dinc.p, inc.data(), 2 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    static const bool use_cluster = getenv("TBV_PGO_CLUSTER") != nullptr;   // opt-in until the cluster kernel has been run and timed on a B200
    static const bool use_chain = getenv("TBV_PGO_CHAIN") != nullptr;       // opt-in: odometry-chain preconditioner (same status)
    if (use_chain) {
      // chain[i] = A[i+1][i]: sum over the constraints between nodes i and i+1 of H_off (begin = i+1) or its transpose (begin = i)
      std::vector<double> chain(36 * (size_t)std::max(n_nodes - 1, 1), 0.0);
      for (int c = 0; c < n_con; c++) {
        const int a = ids[3 * c], b = ids[3 * c + 1];
        if (a - b != 1 && b - a != 1) continue;
        const int lo = a < b ? a : b;
        if (lo == fixed_node || lo + 1 == fixed_node) continue;            // no coupling across the fixed node
        const double* B = H_off + 36 * (size_t)c;
        double* C = chain.data() + 36 * (size_t)lo;
        for (int u = 0; u < 6; u++)
          for (int v = 0; v < 6; v++) C[6 * u + v] += (a > b) ? B[6 * u + v] : B[6 * v + u];
      }
      DevBuf<double> dch, dsi, dwb, dy;
      int rc2;
      if ((rc2 = dch.reserve(chain.size())) || (rc2 = dsi.reserve(36 * (size_t)n_nodes)) || (rc2 = dwb.reserve(chain.size())) || (rc2 = dy.reserve(N6))) {
        dch.release(); dsi.release(); dwb.release(); dy.release(); cleanup();
        return rc2;
      }
      e = cudaMemcpyAsync(dch.p, chain.data(), chain.size() * sizeof(double), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        pgo_pcg_chain<<<1, PCG_THREADS, 0, st>>>(n_nodes, fixed_node, dids.p, drow.p, dinc.p, dhd.p, dho.p, dch.p, dg.p, radius, max_iters, rel_tol, dx.p, dr.p,
                                                  dz.p, dp.p, dq.p, dsi.p, dwb.p, ddg.p, dy.p, dit.p, drel.p);
        launched(ctx, "pgo_pcg_chain");
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);               // `chain` (host) and the extra buffers must outlive the kernel
      }
      dch.release(); dsi.release(); dwb.release(); dy.release();
    } else if (use_cluster) {
      pgo_pcg_cluster<<<PCGC_CL, PCGC_THREADS, 0, st>>>(n_nodes, fixed_node, dids.p, drow.p, dinc.p, dhd.p, dho.p, dg.p, radius, max_iters, rel_tol, dx.p, dr.p,
                                                         dz.p, dp.p, dq.p, dmi.p, ddg.p, dit.p, drel.p);
      launched(ctx, "pgo_pcg_cluster");
    } else {
      pgo_pcg<<<1, PCG_THREADS, 0, st>>>(n_nodes, fixed_node, dids.p, drow.p, dinc.p, dhd.p, dho.p, dg.p, radius, max_iters, rel_tol, dx.p, dr.p, dz.p, dp.p,
                                          dq.p, dmi.p, ddg.p, dit.p, drel.p);
      launched(ctx, "pgo_pcg");
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  int h_it = 0;
  double h_rel = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(delta, dx.p, N6 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_it, dit.p, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_rel, drel.p, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_pgo_solve_step: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  if (iters) *iters = h_it;
  if (rel_residual) *rel_residual = h_rel;
  return TBV_OK;
}

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

// ---- the linear solve of one Levenberg-Marquardt iteration on the device ----------------------------------------------------
// Solves (H + D) delta = -g for the block-sparse normal equations tbv_pgo_assemble produces (H_diag: 6x6 per node, H_off: block
// (begin, end) per constraint, symmetric), D = the LM damping — the system Ceres' LEVENBERG_MARQUARDT strategy hands to
// SPARSE_NORMAL_CHOLESKY (ceresoptimizer.cpp:50-62) — by conjugate gradients preconditioned with the ODOMETRY CHAIN: M = the block-tridiagonal
// part of H + D (diagonal blocks + the blocks coupling nodes i and i + 1).  A pose graph is that chain plus a few hundred weak loop blocks,
// so CG needs ~5 iterations where block-Jacobi needs ~1 500 (4 500 nodes, radius 1e4) and does not converge at all once the trust region
// has grown (measured, profiles/r2a_pgo_*.json).  M is factorised and applied by BLOCK CYCLIC REDUCTION: level l eliminates every second
// of the still active nodes (Schur complements of 6x6 blocks), ceil(log2 n) levels, every level fully parallel over its nodes — an exact
// tridiagonal solve in 2 log2 n parallel steps instead of 2 n sequential ones.  One thread-block CLUSTER of 8 CTAs runs the whole solve:
// vectors and blocks stream from L2 through 8 SMs, the levels and the CG reductions are separated by cluster barriers (~0.4 us), per-CTA
// partial sums are exchanged through distributed shared memory and added in rank order by every CTA (same bits everywhere, so control flow
// is uniform and the result does not depend on scheduling).  Storage per node: Dinv, ML, MR (36 doubles each) — every node is eliminated
// exactly once.
constexpr int PCR_CL = 8;            // portable cluster size
constexpr int PCR_THREADS = 1024;

__device__ __forceinline__ void m6_mul(const double* A, const double* B, double* O) {      // O = A B
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      double v = 0.0;
      for (int k = 0; k < 6; k++) v += A[6 * a + k] * B[6 * k + b];
      O[6 * a + b] = v;
    }
}
__device__ __forceinline__ void m6_mul_tn(const double* A, const double* B, double* O) {   // O = A^T B
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      double v = 0.0;
      for (int k = 0; k < 6; k++) v += A[6 * k + a] * B[6 * k + b];
      O[6 * a + b] = v;
    }
}
__device__ __forceinline__ void m6_submul(const double* A, const double* B, double* O) {    // O -= A B
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      double v = 0.0;
      for (int k = 0; k < 6; k++) v += A[6 * a + k] * B[6 * k + b];
      O[6 * a + b] -= v;
    }
}
__device__ __forceinline__ void m6_submul_nt(const double* A, const double* B, double* O) { // O -= A B^T
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      double v = 0.0;
      for (int k = 0; k < 6; k++) v += A[6 * a + k] * B[6 * b + k];
      O[6 * a + b] -= v;
    }
}
__device__ void m6_inverse_spd(const double* S, double* Sinv, bool* ok) {   // Cholesky S = L L^T, then S^-1 = L^-T L^-1
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

// The factorisation runs as ordinary grids, one thread per node of a level with its 6x6 blocks in registers (128 threads per CTA: no
// spills; inside the 1024-thread cluster kernel the same code is limited to 64 registers and ran 15x slower): 2 launches per level.
constexpr int PCRF_THREADS = 128;

// the damped chain: Dm_i = H_ii + D_i (identity for the fixed node), Cc_i = A[i+1][i] = sum of the blocks between nodes i and i+1
__global__ void __launch_bounds__(PCRF_THREADS)
pgo_cr_setup(int n, int fixed_node, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
             const double* __restrict__ Ho, const double* __restrict__ damping, double radius, double* __restrict__ Dm, double* __restrict__ Cc,
             double* __restrict__ Dg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double D[36];
  for (int e = 0; e < 36; e++) D[e] = (i == fixed_node) ? ((e % 7 == 0) ? 1.0 : 0.0) : Hd[36 * (size_t)i + e];
  for (int a = 0; a < 6; a++) {
    const double d = damping ? damping[6 * (size_t)i + a] : fmin(fmax(Hd[36 * (size_t)i + 7 * a], 1e-6), 1e32) / radius;
    Dg[6 * (size_t)i + a] = d;
    if (i != fixed_node) D[7 * a] += d;
  }
  for (int e = 0; e < 36; e++) Dm[36 * (size_t)i + e] = D[e];
  double C[36];
  for (int e = 0; e < 36; e++) C[e] = 0.0;
  if (i + 1 < n && i != fixed_node && i + 1 != fixed_node)      // no coupling across the fixed node
    for (int e = row[i]; e < row[i + 1]; e++) {
      const int c = inc[e] >> 1, side = inc[e] & 1;
      const int other = side ? ids[3 * c] : ids[3 * c + 1];
      if (other != i + 1) continue;
      const double* B = Ho + 36 * (size_t)c;                    // block (begin, end)
      // side 0: i begins the constraint, B = A[i][i+1] -> A[i+1][i] = B^T; side 1: i ends it, B = A[i+1][i]
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) C[6 * a + b] += side ? B[6 * a + b] : B[6 * b + a];
    }
  for (int e = 0; e < 36; e++) Cc[36 * (size_t)i + e] = C[e];
}
// block cyclic reduction, stride s: active nodes = multiples of s; odd multiples are eliminated.
// (A) per eliminated node o: Dinv, ML = A[e1][o] Dinv, MR = A[e2][o] Dinv      (n_odd == 0 and s >= n: node 0 alone, Dinv only)
__global__ void __launch_bounds__(PCRF_THREADS)
pgo_cr_eliminate(int n, int s, int n_odd, const double* __restrict__ Dm, const double* __restrict__ Cc, double* __restrict__ Dinv, double* __restrict__ ML,
                 double* __restrict__ MR, int* __restrict__ bad) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_odd == 0) {                                    // the last active node
    if (j != 0) return;
    double D[36], Di[36];
    for (int e = 0; e < 36; e++) D[e] = Dm[e];
    bool ok;
    m6_inverse_spd(D, Di, &ok);
    if (!ok) *bad = 1;
    for (int e = 0; e < 36; e++) Dinv[e] = Di[e];
    return;
  }
  if (j >= n_odd) return;
  const int o = s * (2 * j + 1), e1 = o - s, e2 = o + s;
  double D[36], Di[36], C[36], M[36];
  for (int e = 0; e < 36; e++) D[e] = Dm[36 * (size_t)o + e];
  bool ok;
  m6_inverse_spd(D, Di, &ok);
  if (!ok) *bad = 1;
  for (int e = 0; e < 36; e++) Dinv[36 * (size_t)o + e] = Di[e];
  for (int e = 0; e < 36; e++) C[e] = Cc[36 * (size_t)e1 + e];     // A[o][e1]
  m6_mul_tn(C, Di, M);                                              // A[e1][o] Dinv = C^T Dinv
  for (int e = 0; e < 36; e++) ML[36 * (size_t)o + e] = M[e];
  if (e2 < n) {
    for (int e = 0; e < 36; e++) C[e] = Cc[36 * (size_t)o + e];    // A[e2][o]
    m6_mul(C, Di, M);
  } else {
    for (int e = 0; e < 36; e++) M[e] = 0.0;
  }
  for (int e = 0; e < 36; e++) MR[36 * (size_t)o + e] = M[e];
}
// (B) per surviving node e: Schur complement and the new coupling to e + 2s
__global__ void __launch_bounds__(PCRF_THREADS)
pgo_cr_update(int n, int s, int n_even, double* __restrict__ Dm, double* __restrict__ Cc, const double* __restrict__ ML, const double* __restrict__ MR) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_even) return;
  const int e0 = 2 * s * j, ol = e0 - s, orr = e0 + s;
  if (ol < 0 && orr >= n) return;
  double D[36], M[36], C[36];
  for (int e = 0; e < 36; e++) D[e] = Dm[36 * (size_t)e0 + e];
  if (ol >= 0) {                                       // D -= A[e][ol] Dinv A[ol][e] = MR_ol * Cc[ol]^T
    for (int e = 0; e < 36; e++) { M[e] = MR[36 * (size_t)ol + e]; C[e] = Cc[36 * (size_t)ol + e]; }
    m6_submul_nt(M, C, D);
  }
  if (orr < n) {                                       // D -= A[e][or] Dinv A[or][e] = ML_or * Cc[e]
    for (int e = 0; e < 36; e++) { M[e] = ML[36 * (size_t)orr + e]; C[e] = Cc[36 * (size_t)e0 + e]; }
    m6_submul(M, C, D);
    double Cn[36];
    for (int e = 0; e < 36; e++) Cn[e] = 0.0;
    if (e0 + 2 * s < n) {                              // A'[e + 2s][e] = -A[e + 2s][or] Dinv A[or][e] = -MR_or * Cc[e]
      for (int e = 0; e < 36; e++) M[e] = MR[36 * (size_t)orr + e];
      m6_submul(M, C, Cn);
    }
    for (int e = 0; e < 36; e++) Cc[36 * (size_t)e0 + e] = Cn[e];
  }
  for (int e = 0; e < 36; e++) Dm[36 * (size_t)e0 + e] = D[e];
}

__global__ void __cluster_dims__(PCR_CL, 1, 1) __launch_bounds__(PCR_THREADS, 1)
pgo_pcg_cr(int n, int fixed_node, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
           const double* __restrict__ Ho, const double* __restrict__ g, int max_iters, double rel_tol, double* __restrict__ x, double* r, double* z,
           double* p, double* __restrict__ q, double* u, const double* __restrict__ Dinv, const double* __restrict__ ML, const double* __restrict__ MR,
           const double* __restrict__ Dg, const int* __restrict__ bad, int* __restrict__ out_iters, double* __restrict__ out_rel) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_w[PCR_THREADS / 32][2];
  __shared__ double s_part[2][2];                    // [0]: (p.q, -)   [1]: (r.z, r.r) — this CTA's partials, read by the whole cluster
  const int tid = threadIdx.x, rank = (int)cluster.block_rank();
  const int gt = rank * PCR_THREADS + tid, GT = PCR_CL * PCR_THREADS;
  const int npc = (n + PCR_CL - 1) / PCR_CL;          // nodes per CTA for the row-parallel CG work
  const int n0 = min(rank * npc, n), n1 = min(n0 + npc, n);
  const int rows0 = 6 * n0, nrows = 6 * (n1 - n0);

  auto block_sum2 = [&](double& a, double& b) {   // fixed tree per CTA; both values at once
    for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    const int warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    if (lane == 0) { s_w[warp][0] = a; s_w[warp][1] = b; }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < PCR_THREADS / 32; w++) { ta += s_w[w][0]; tb += s_w[w][1]; }
    a = ta; b = tb;
  };
  // cluster-wide sums of (a, b): per-CTA fixed tree, then every CTA adds the PCR_CL partials in rank order (DSMEM reads)
  auto cluster_sum2 = [&](double& a, double& b, int slot) {
    block_sum2(a, b);
    if (tid == 0) { s_part[slot][0] = a; s_part[slot][1] = b; }
    cluster.sync();
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < PCR_CL; k++) {
      const double* remote = cluster.map_shared_rank(&s_part[slot][0], k);
      ta += remote[0]; tb += remote[1];
    }
    a = ta; b = tb;
  };

  // z = M^-1 v by cyclic reduction: forward over the levels (right-hand sides of the surviving nodes), node 0, backward (eliminated nodes).
  // One (node, component) row per thread; u is the working right-hand side.  Reads of other CTAs' rows go to L2 (__ldcg).
  auto apply_chain = [&](const double* v, double* o) {
    for (int R = gt; R < 6 * n; R += GT) u[R] = __ldcg(v + R);
    cluster.sync();
    for (int s = 1; s < n; s <<= 1) {
      const int n_even = (n - 1) / (2 * s) + 1;
      for (int t = gt; t < 6 * n_even; t += GT) {
        const int j = t / 6, a = t - 6 * j, e0 = 2 * s * j, ol = e0 - s, orr = e0 + s;
        double acc = __ldcg(u + 6 * (size_t)e0 + a);
        if (ol >= 0) {
          const double* M = MR + 36 * (size_t)ol + 6 * a;
          const double* w = u + 6 * (size_t)ol;
          for (int b = 0; b < 6; b++) acc -= M[b] * __ldcg(w + b);
        }
        if (orr < n) {
          const double* M = ML + 36 * (size_t)orr + 6 * a;
          const double* w = u + 6 * (size_t)orr;
          for (int b = 0; b < 6; b++) acc -= M[b] * __ldcg(w + b);
        }
        u[6 * (size_t)e0 + a] = acc;
      }
      cluster.sync();
    }
    if (gt < 6) {
      double acc = 0.0;
      for (int b = 0; b < 6; b++) acc += Dinv[6 * gt + b] * __ldcg(u + b);
      o[gt] = acc;
    }
    cluster.sync();
    int s_top = 1;
    while (s_top * 2 < n) s_top <<= 1;
    for (int s = s_top; s >= 1; s >>= 1) {
      if (s >= n) continue;
      const int n_odd = (n - 1 >= s) ? (n - 1 - s) / (2 * s) + 1 : 0;
      for (int t = gt; t < 6 * n_odd; t += GT) {
        const int j = t / 6, a = t - 6 * j, od = s * (2 * j + 1), e1 = od - s, e2 = od + s;
        const double* Di = Dinv + 36 * (size_t)od + 6 * a;
        const double* w = u + 6 * (size_t)od;
        double acc = 0.0;
        for (int b = 0; b < 6; b++) acc += Di[b] * __ldcg(w + b);
        {                                              // - (ML^T z_e1)[a]
          const double* M = ML + 36 * (size_t)od;
          const double* zz = o + 6 * (size_t)e1;
          for (int b = 0; b < 6; b++) acc -= M[6 * b + a] * __ldcg(zz + b);
        }
        if (e2 < n) {                                  // - (MR^T z_e2)[a]
          const double* M = MR + 36 * (size_t)od;
          const double* zz = o + 6 * (size_t)e2;
          for (int b = 0; b < 6; b++) acc -= M[6 * b + a] * __ldcg(zz + b);
        }
        o[6 * (size_t)od + a] = acc;
      }
      cluster.sync();
    }
  };

  // ---- x = 0, r = b = -g (the fixed node's rows are zero), z = M^-1 r, p = z ---------------------------------------------------------------------
  for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
    const int R = rows0 + lr, i = R / 6;
    x[R] = 0.0;
    r[R] = (i == fixed_node) ? 0.0 : -g[R];
  }
  cluster.sync();
  apply_chain(r, z);
  double rz = 0.0, bb = 0.0;
  for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
    const int R = rows0 + lr;
    const double zr = __ldcg(z + R);
    p[R] = zr;
    rz += r[R] * zr; bb += r[R] * r[R];
  }
  cluster_sum2(rz, bb, 1);                           // also publishes p to the cluster (barrier.cluster release / acquire)
  const double bnorm = sqrt(bb);
  double rel = bnorm > 0.0 ? 1.0 : 0.0;
  int it = 0;
  while (it < max_iters && rel > rel_tol) {
    // q = (H + D) p, own rows; p.q
    double pq = 0.0, unused = 0.0;
    for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
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
          if (other == fixed_node) continue;          // the fixed node's column is not part of the system
          const double* B = Ho + 36 * (size_t)c;
          const double* uu = p + 6 * (size_t)other;
          // the other node may belong to another CTA: read its slice of p from L2 (the cluster barrier orders the writes; no stale L1 line)
          if (side == 0) { for (int b = 0; b < 6; b++) acc += B[6 * a + b] * __ldcg(uu + b); }
          else           { for (int b = 0; b < 6; b++) acc += B[6 * b + a] * __ldcg(uu + b); }
        }
      }
      q[R] = acc;
      pq += p[R] * acc;
    }
    cluster_sum2(pq, unused, 0);
    if (!(pq > 0.0)) break;                          // same value in every CTA: the whole cluster leaves together
    const double alpha = rz / pq;
    for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
      const int R = rows0 + lr;
      x[R] += alpha * p[R];
      r[R] -= alpha * q[R];
    }
    cluster.sync();                                  // the preconditioner reads every row of r
    apply_chain(r, z);
    double rz_new = 0.0, rr_new = 0.0;
    for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
      const int R = rows0 + lr;
      rz_new += r[R] * __ldcg(z + R);
      rr_new += r[R] * r[R];
    }
    cluster_sum2(rz_new, rr_new, 1);
    rel = sqrt(rr_new) / bnorm;
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int lr = tid; lr < nrows; lr += PCR_THREADS) {
      const int R = rows0 + lr;
      p[R] = __ldcg(z + R) + beta * p[R];
    }
    cluster.sync();                                  // the next product reads other CTAs' rows of p
    it++;
  }
  cluster.sync();                                    // no CTA may exit while another still reads its partials through DSMEM
  // a pivot block of the chain that was not positive definite (factorisation kernels) is reported through the sign of the iteration count
  if (rank == 0 && tid == 0) { *out_iters = *bad ? -it - 1 : it; *out_rel = rel; }
}

// ---- trust-region bookkeeping of tbv_pgo_optimize on the device (the vectors never leave HBM between LM iterations) ------------------------------
// x (+) delta of a node: p += dp; q = exp(dr) * q (ceres::EigenQuaternionParameterization::Plus, q = x y z w)
__device__ __forceinline__ void pgo_plus_node(const double* nd, const double* d, double* o) {
  o[0] = nd[0] + d[0]; o[1] = nd[1] + d[1]; o[2] = nd[2] + d[2];
  const double nrm = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
  const double k = nrm > 0.0 ? sin(nrm) / nrm : 1.0;
  const double dv[3] = {d[3] * k, d[4] * k, d[5] * k}, dw = cos(nrm);
  const double* v = nd + 3;
  const double w = nd[6];
  o[6] = dw * w - (dv[0] * v[0] + dv[1] * v[1] + dv[2] * v[2]);
  o[3] = dw * v[0] + w * dv[0] + (dv[1] * v[2] - dv[2] * v[1]);
  o[4] = dw * v[1] + w * dv[1] + (dv[2] * v[0] - dv[0] * v[2]);
  o[5] = dw * v[2] + w * dv[2] + (dv[0] * v[1] - dv[1] * v[0]);
}

constexpr int PGS_THREADS = 1024;
__device__ __forceinline__ double pgs_block_sum(double v, double* s_w) {   // fixed tree: lanes, then the warp totals in order (deterministic)
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < PGS_THREADS / 32; w++) t += s_w[w];
  return t;
}
__device__ __forceinline__ double pgs_block_max(double v, double* s_w) {
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < PGS_THREADS / 32; w++) t = fmax(t, s_w[w]);
  return t;
}

// Jacobi scaling s = 1 / (1 + sqrt(diag H)) of LevenbergMarquardtStrategy (fixed at iteration 0)
__global__ void pgo_scale(int n, const double* __restrict__ Hd, double* __restrict__ scale) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= 6 * n) return;
  const int i = R / 6, a = R - 6 * i;
  scale[R] = 1.0 / (1.0 + sqrt(fmax(Hd[36 * (size_t)i + 7 * a], 0.0)));
}
// LM diagonal clamp(diag(S H S), 1e-6, 1e32) — recomputed after a successful step, reused after a rejected one — and the damping
// (H + diag / (radius s^2)) delta = -g that is equivalent to (S H S + diag / radius) y = -S g, delta = S y
__global__ void pgo_damping(int n, int fixed_node, const double* __restrict__ Hd, const double* __restrict__ scale, int reuse, double radius,
                            double* __restrict__ lm_diag, double* __restrict__ damping) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= 6 * n) return;
  const int i = R / 6, a = R - 6 * i;
  const double s2 = scale[R] * scale[R];
  double d = lm_diag[R];
  if (!reuse) { d = fmin(fmax(Hd[36 * (size_t)i + 7 * a] * s2, 1e-6), 1e32); lm_diag[R] = d; }
  damping[R] = (i == fixed_node) ? 1.0 : d / (radius * s2);
}
// candidate = x (+) delta; stats[0] = model change -delta.(g + H delta / 2), [1] = |x - candidate|^2, [2] = |candidate|^2, [3] = 1 if delta is finite
__global__ void __launch_bounds__(PGS_THREADS, 1)
pgo_step(int n, const int* __restrict__ ids, const int* __restrict__ row, const int* __restrict__ inc, const double* __restrict__ Hd,
         const double* __restrict__ Ho, const double* __restrict__ g, const double* __restrict__ delta, const double* __restrict__ nodes,
         double* __restrict__ cand, double* __restrict__ stats) {
  __shared__ double s_w[PGS_THREADS / 32];
  double mc = 0.0, dn = 0.0, cn = 0.0, bad = 0.0;
  for (int i = threadIdx.x; i < n; i += PGS_THREADS) {
    double d[6], hd[6];
    for (int a = 0; a < 6; a++) d[a] = delta[6 * (size_t)i + a];
    for (int a = 0; a < 6; a++) {
      double acc = 0.0;
      for (int b = 0; b < 6; b++) acc += Hd[36 * (size_t)i + 6 * a + b] * d[b];
      hd[a] = acc;
    }
    for (int e = row[i]; e < row[i + 1]; e++) {
      const int c = inc[e] >> 1, side = inc[e] & 1;
      const int other = side ? ids[3 * c] : ids[3 * c + 1];
      const double* B = Ho + 36 * (size_t)c;
      const double* u = delta + 6 * (size_t)other;
      if (side == 0) { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) hd[a] += B[6 * a + b] * u[b]; }
      else           { for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) hd[a] += B[6 * b + a] * u[b]; }
    }
    for (int a = 0; a < 6; a++) {
      mc -= d[a] * (g[6 * (size_t)i + a] + 0.5 * hd[a]);
      if (!isfinite(d[a])) bad = 1.0;
    }
    double o[7];
    pgo_plus_node(nodes + 7 * (size_t)i, d, o);
    for (int c = 0; c < 7; c++) {
      const double df = nodes[7 * (size_t)i + c] - o[c];
      dn += df * df; cn += o[c] * o[c];
      cand[7 * (size_t)i + c] = o[c];
    }
  }
  mc = pgs_block_sum(mc, s_w); dn = pgs_block_sum(dn, s_w); cn = pgs_block_sum(cn, s_w); bad = pgs_block_sum(bad, s_w);
  if (threadIdx.x == 0) { stats[0] = mc; stats[1] = dn; stats[2] = cn; stats[3] = bad > 0.0 ? 0.0 : 1.0; }
}
// stats[4] = max |x - Plus(x, -g)| over the free nodes (Ceres' gradient max norm for a manifold), stats[5] = |x|^2
__global__ void __launch_bounds__(PGS_THREADS, 1)
pgo_gradmax(int n, int fixed_node, const double* __restrict__ nodes, const double* __restrict__ g, double* __restrict__ stats) {
  __shared__ double s_w[PGS_THREADS / 32];
  double gm = 0.0, xn = 0.0;
  for (int i = threadIdx.x; i < n; i += PGS_THREADS) {
    double d[6], o[7];
    for (int a = 0; a < 6; a++) d[a] = (i == fixed_node) ? 0.0 : -g[6 * (size_t)i + a];
    pgo_plus_node(nodes + 7 * (size_t)i, d, o);
    for (int c = 0; c < 7; c++) {
      const double v = nodes[7 * (size_t)i + c];
      gm = fmax(gm, fabs(v - o[c]));
      xn += v * v;
    }
  }
  gm = pgs_block_max(gm, s_w); xn = pgs_block_sum(xn, s_w);
  if (threadIdx.x == 0) { stats[4] = gm; stats[5] = xn; }
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


namespace {

// Device-resident pose graph: constraint lists, CSR incidence, the assembled blocks and every solver vector.  Built once per call of
// tbv_pgo_solve_step / tbv_pgo_optimize; all Levenberg-Marquardt iterations of tbv_pgo_optimize run on these buffers.
struct PgoDev {
  int n = 0, m = 0, fixed = 0;
  bool has_info = false;
  DevBuf<double> nodes, cand, meas, info, rec, cost;          // nodes [n][7] (current / candidate), measurements, records, 1 cost
  DevBuf<double> Hd, Ho, g, Hd_c, Ho_c, g_c;                  // normal equations at the current point / at the candidate
  DevBuf<double> x, r, z, p, q, u, Dm, Cc, Dinv, ML, MR, Dg;  // linear solve
  DevBuf<double> scale, lm_diag, damping, stats;              // trust-region bookkeeping (stats: 8 doubles)
  DevBuf<int> ids, row, inc, order, err, it;
  DevBuf<double> rel;
  void release() {
    for (DevBuf<double>* b : {&nodes, &cand, &meas, &info, &rec, &cost, &Hd, &Ho, &g, &Hd_c, &Ho_c, &g_c, &x, &r, &z, &p, &q, &u, &Dm, &Cc, &Dinv, &ML, &MR, &Dg,
                              &scale, &lm_diag, &damping, &stats, &rel})
      b->release();
    for (DevBuf<int>* b : {&ids, &row, &inc, &order, &err, &it}) b->release();
  }
};

// id list -> CSR incidence (node -> constraints, reference order: all odometry constraints, then all loop constraints) and uploads
int pgo_upload_graph(tbv_ctx* ctx, PgoDev& G, int n_nodes, int n_con, const int* ids, bool reference_order) {
  for (int c = 0; c < n_con; c++)
    TBV_REQUIRE(ids[3 * c] >= 0 && ids[3 * c] < n_nodes && ids[3 * c + 1] >= 0 && ids[3 * c + 1] < n_nodes, "constraint references a missing node");
  G.n = n_nodes; G.m = n_con;
  std::vector<int> order(n_con);
  std::iota(order.begin(), order.end(), 0);
  if (reference_order)   // AddConstraintType(odometry) then AddConstraintType(loop) (ceresoptimizer.cpp:34-35)
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (ids[3 * a + 2] == 1) < (ids[3 * b + 2] == 1); });
  std::vector<int> row(n_nodes + 1, 0), inc(2 * (size_t)n_con + 1);
  for (int c = 0; c < n_con; c++) { row[ids[3 * c] + 1]++; row[ids[3 * c + 1] + 1]++; }
  for (int i = 0; i < n_nodes; i++) row[i + 1] += row[i];
  {
    std::vector<int> fill(row.begin(), row.end() - 1);
    for (int c : order) { inc[fill[ids[3 * c]]++] = (c << 1); inc[fill[ids[3 * c + 1]]++] = (c << 1) | 1; }
  }
  const size_t nc1 = n_con ? n_con : 1;
  int rc;
  if ((rc = G.ids.reserve(3 * nc1)) || (rc = G.row.reserve(n_nodes + 1)) || (rc = G.inc.reserve(inc.size())) || (rc = G.order.reserve(nc1)) ||
      (rc = G.err.reserve(2)) || (rc = G.it.reserve(1)) || (rc = G.rel.reserve(1)))
    return rc;
  cudaStream_t st = ctx->stream;
  if (n_con) TBV_CUDA(cudaMemcpyAsync(G.ids.p, ids, 3 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st));
  TBV_CUDA(cudaMemcpyAsync(G.row.p, row.data(), (n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (n_con) TBV_CUDA(cudaMemcpyAsync(G.inc.p, inc.data(), 2 * (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st));
  if (n_con) TBV_CUDA(cudaMemcpyAsync(G.order.p, order.data(), (size_t)n_con * sizeof(int), cudaMemcpyHostToDevice, st));
  TBV_CUDA(cudaMemsetAsync(G.err.p, 0, sizeof(int), st));
  TBV_CUDA(cudaStreamSynchronize(st));   // the host vectors go out of scope
  return TBV_OK;
}

int pgo_reserve_solver(PgoDev& G) {
  const size_t N6 = 6 * (size_t)G.n, N36 = 36 * (size_t)G.n;
  int rc;
  for (DevBuf<double>* b : {&G.x, &G.r, &G.z, &G.p, &G.q, &G.u, &G.Dg})
    if ((rc = b->reserve(N6))) return rc;
  for (DevBuf<double>* b : {&G.Dm, &G.Cc, &G.Dinv, &G.ML, &G.MR})
    if ((rc = b->reserve(N36))) return rc;
  return TBV_OK;
}

// enqueue (H + D) delta = -g on the current blocks; D = `damping` (device, [n][6]) or clamp(diag H, 1e-6, 1e32) / radius when null
int pgo_enqueue_solve(tbv_ctx* ctx, PgoDev& G, const double* Hd, const double* Ho, const double* g, const double* damping, double radius, int max_iters,
                      double rel_tol) {
  cudaStream_t st = ctx->stream;
  const int n = G.n;
  TBV_CUDA(cudaMemsetAsync(G.err.p + 1, 0, sizeof(int), st));   // err[1]: a chain pivot block was not positive definite
  pgo_cr_setup<<<(n + PCRF_THREADS - 1) / PCRF_THREADS, PCRF_THREADS, 0, st>>>(n, G.fixed, G.ids.p, G.row.p, G.inc.p, Hd, Ho, damping, radius, G.Dm.p, G.Cc.p,
                                                                               G.Dg.p);
  launched(ctx, "pgo_cr_setup");
  for (int s = 1; s < n; s <<= 1) {
    const int n_odd = (n - 1 >= s) ? (n - 1 - s) / (2 * s) + 1 : 0, n_even = (n - 1) / (2 * s) + 1;
    if (n_odd > 0) {
      pgo_cr_eliminate<<<(n_odd + PCRF_THREADS - 1) / PCRF_THREADS, PCRF_THREADS, 0, st>>>(n, s, n_odd, G.Dm.p, G.Cc.p, G.Dinv.p, G.ML.p, G.MR.p, G.err.p + 1);
      launched(ctx, "pgo_cr_eliminate");
    }
    pgo_cr_update<<<(n_even + PCRF_THREADS - 1) / PCRF_THREADS, PCRF_THREADS, 0, st>>>(n, s, n_even, G.Dm.p, G.Cc.p, G.ML.p, G.MR.p);
    launched(ctx, "pgo_cr_update");
  }
  pgo_cr_eliminate<<<1, PCRF_THREADS, 0, st>>>(n, 0, 0, G.Dm.p, G.Cc.p, G.Dinv.p, G.ML.p, G.MR.p, G.err.p + 1);   // node 0 is what is left
  launched(ctx, "pgo_cr_eliminate");
  pgo_pcg_cr<<<PCR_CL, PCR_THREADS, 0, st>>>(n, G.fixed, G.ids.p, G.row.p, G.inc.p, Hd, Ho, g, max_iters, rel_tol, G.x.p, G.r.p, G.z.p, G.p.p, G.q.p, G.u.p,
                                             G.Dinv.p, G.ML.p, G.MR.p, G.Dg.p, G.err.p + 1, G.it.p, G.rel.p);
  launched(ctx, "pgo_pcg_cr");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

}  // namespace

extern "C" int tbv_pgo_solve_step(tbv_ctx* ctx, int n_nodes, int n_con, const int* ids, const double* H_diag, const double* H_off, const double* g,
                                  int fixed_node, double radius, int max_iters, double rel_tol, double* delta, int* iters, double* rel_residual) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && ids && H_diag && H_off && g && delta && n_nodes >= 1 && n_con >= 0 && radius > 0 && max_iters >= 0, "bad arguments");
  return tbv_pgo_solve_damped(ctx, n_nodes, n_con, ids, H_diag, H_off, g, nullptr, fixed_node, radius, max_iters, rel_tol, delta, iters, rel_residual);
}

extern "C" int tbv_pgo_solve_damped(tbv_ctx* ctx, int n_nodes, int n_con, const int* ids, const double* H_diag, const double* H_off, const double* g,
                                    const double* damping, int fixed_node, double radius, int max_iters, double rel_tol, double* delta, int* iters,
                                    double* rel_residual) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && ids && H_diag && H_off && g && delta && n_nodes >= 1 && n_con >= 0 && max_iters >= 0 && (damping || radius > 0), "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  PgoDev G;
  G.fixed = fixed_node;
  int rc = pgo_upload_graph(ctx, G, n_nodes, n_con, ids, false);
  const size_t N6 = 6 * (size_t)n_nodes, nc1 = n_con ? n_con : 1;
  if (!rc) rc = pgo_reserve_solver(G);
  if (!rc && ((rc = G.Hd.reserve(36 * (size_t)n_nodes)) || (rc = G.Ho.reserve(36 * nc1)) || (rc = G.g.reserve(N6)) || (damping && (rc = G.damping.reserve(N6))))) {}
  if (rc) { G.release(); return rc; }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(G.Hd.p, H_diag, 36 * (size_t)n_nodes * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(G.Ho.p, H_off, 36 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(G.g.p, g, N6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && damping) e = cudaMemcpyAsync(G.damping.p, damping, N6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) rc = pgo_enqueue_solve(ctx, G, G.Hd.p, G.Ho.p, G.g.p, damping ? G.damping.p : nullptr, radius, max_iters, rel_tol);
  int h_it = 0;
  double h_rel = 0;
  if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(delta, G.x.p, N6 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(&h_it, G.it.p, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(&h_rel, G.rel.p, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  G.release();
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("tbv_pgo_solve_damped: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  if (iters) *iters = h_it;
  if (rel_residual) *rel_residual = h_rel;
  return TBV_OK;
}

namespace {
// residuals, Jacobians and the block normal equations at `nodes` (device) -> Hd, Ho, g, G.cost (device); enqueued on the context's stream
int pgo_enqueue_assemble(tbv_ctx* ctx, PgoDev& G, const tbv_pgo_params& P, const double* nodes, double* Hd, double* Ho, double* g) {
  cudaStream_t st = ctx->stream;
  if (G.m) {
    pgo_blocks<<<(G.m + 63) / 64, 64, 0, st>>>(G.m, nodes, G.ids.p, G.meas.p, G.has_info ? G.info.p : nullptr, P, G.fixed, G.rec.p, Ho, nullptr, G.err.p);
    launched(ctx, "pgo_blocks");
  }
  pgo_gather<<<(G.n * 32 + 127) / 128, 128, 0, st>>>(G.n, G.row.p, G.inc.p, G.rec.p, Hd, g);
  launched(ctx, "pgo_gather");
  pgo_cost<<<1, 256, 0, st>>>(G.m, G.order.p, G.rec.p, G.cost.p);
  launched(ctx, "pgo_cost");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}
}  // namespace

extern "C" int tbv_pgo_optimize(tbv_ctx* ctx, int n_nodes, double* nodes, int n_con, const int* ids, const double* meas, const double* info,
                                const tbv_pgo_params* params, int fixed_node, const tbv_pgo_options* options, tbv_pgo_summary* summary) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && nodes && ids && meas && params && n_nodes >= 1 && n_con >= 0, "bad arguments");
  TBV_REQUIRE(params->replace_cov_by_identity || info, "information matrices required when replace_cov_by_identity is 0");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  tbv_pgo_options O = {200, 1e-6, 1e-10, 1e-8, 1e4, 20000, 1e-10};   // ceresoptimizer.cpp:13-15 (max_num_iterations 200) + Ceres 2.1.0 defaults
  if (options) {
    if (options->max_num_iterations > 0) O.max_num_iterations = options->max_num_iterations;
    if (options->function_tolerance > 0) O.function_tolerance = options->function_tolerance;
    if (options->gradient_tolerance > 0) O.gradient_tolerance = options->gradient_tolerance;
    if (options->parameter_tolerance > 0) O.parameter_tolerance = options->parameter_tolerance;
    if (options->initial_radius > 0) O.initial_radius = options->initial_radius;
    if (options->max_cg_iterations > 0) O.max_cg_iterations = options->max_cg_iterations;
    if (options->cg_rel_tol > 0) O.cg_rel_tol = options->cg_rel_tol;
  }
  PgoDev G;
  G.fixed = fixed_node;
  G.has_info = info != nullptr && !params->replace_cov_by_identity;
  cudaStream_t st = ctx->stream;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  auto fail = [&](int rc) { G.release(); if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); return rc; };
  int rc = pgo_upload_graph(ctx, G, n_nodes, n_con, ids, true);
  if (rc) return fail(rc);
  const size_t N6 = 6 * (size_t)n_nodes, N7 = 7 * (size_t)n_nodes, N36 = 36 * (size_t)n_nodes, nc1 = n_con ? n_con : 1;
  if ((rc = pgo_reserve_solver(G))) return fail(rc);
  for (DevBuf<double>* b : {&G.Hd, &G.Hd_c})
    if ((rc = b->reserve(N36))) return fail(rc);
  for (DevBuf<double>* b : {&G.Ho, &G.Ho_c})
    if ((rc = b->reserve(36 * nc1))) return fail(rc);
  for (DevBuf<double>* b : {&G.g, &G.g_c, &G.scale, &G.lm_diag, &G.damping})
    if ((rc = b->reserve(N6))) return fail(rc);
  if ((rc = G.nodes.reserve(N7)) || (rc = G.cand.reserve(N7)) || (rc = G.meas.reserve(7 * nc1)) || (rc = G.info.reserve(G.has_info ? 36 * nc1 : 1)) ||
      (rc = G.rec.reserve(PGB * nc1)) || (rc = G.cost.reserve(1)) || (rc = G.stats.reserve(8)))
    return fail(rc);
  cudaError_t e = cudaMemcpyAsync(G.nodes.p, nodes, N7 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con) e = cudaMemcpyAsync(G.meas.p, meas, 7 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_con && G.has_info) e = cudaMemcpyAsync(G.info.p, info, 36 * (size_t)n_con * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaEventCreate(&ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ev1);
  if (e == cudaSuccess) e = cudaEventRecord(ev0, st);
  if (e != cudaSuccess) { set_error("tbv_pgo_optimize: %s", cudaGetErrorString(e)); return fail(TBV_ERR_CUDA); }

  double* x = G.nodes.p; double* cand = G.cand.p;
  double* Hd = G.Hd.p; double* Ho = G.Ho.p; double* g = G.g.p;
  double* Hd_c = G.Hd_c.p; double* Ho_c = G.Ho_c.p; double* g_c = G.g_c.p;
  double h_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double h_cost = 0.0;
  int h_err = 0, h_it = 0;
  auto read_eval = [&](double* cost_out) -> int {   // cost + gradient max norm + |x|^2 of the point just evaluated
    TBV_CUDA(cudaMemcpyAsync(&h_cost, G.cost.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    TBV_CUDA(cudaMemcpyAsync(&h_err, G.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TBV_CUDA(cudaMemcpyAsync(h_stats + 4, G.stats.p + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TBV_CUDA(cudaStreamSynchronize(st));
    if (h_err) { set_error("tbv_pgo_optimize: an information matrix is not positive definite"); return TBV_ERR_INVALID; }
    *cost_out = h_cost;
    return TBV_OK;
  };
  const int rows_grid = (int)((N6 + 255) / 256);
  tbv_pgo_summary S = {};
  // ---- iteration 0: evaluate, Jacobi scaling, gradient norm -------------------------------------------------------------------------------------
  double x_cost = 0.0;
  if ((rc = pgo_enqueue_assemble(ctx, G, *params, x, Hd, Ho, g))) return fail(rc);
  pgo_scale<<<rows_grid, 256, 0, st>>>(n_nodes, Hd, G.scale.p);
  launched(ctx, "pgo_scale");
  pgo_gradmax<<<1, PGS_THREADS, 0, st>>>(n_nodes, fixed_node, x, g, G.stats.p);
  launched(ctx, "pgo_gradmax");
  if ((rc = read_eval(&x_cost))) return fail(rc);
  S.initial_cost = S.final_cost = x_cost;
  double gmax = h_stats[4], x_norm = sqrt(h_stats[5]);
  double radius = O.initial_radius, decrease = 2.0;
  bool reuse = false, step_ok = true;
  int invalid = 0, iteration = 0;
  S.termination = TBV_PGO_MAX_ITERATIONS;
  const double DBL_BIG = 1.7976931348623157e308;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= O.max_num_iterations) { S.termination = TBV_PGO_MAX_ITERATIONS; break; }
    if (step_ok && gmax <= O.gradient_tolerance) { S.termination = TBV_PGO_GRADIENT_TOLERANCE; break; }
    if (radius < 1e-32) { S.termination = TBV_PGO_MIN_RADIUS; break; }
    iteration++;
    S.iterations = iteration;
    // LevenbergMarquardtStrategy::ComputeStep
    pgo_damping<<<rows_grid, 256, 0, st>>>(n_nodes, fixed_node, Hd, G.scale.p, reuse ? 1 : 0, radius, G.lm_diag.p, G.damping.p);
    launched(ctx, "pgo_damping");
    if ((rc = pgo_enqueue_solve(ctx, G, Hd, Ho, g, G.damping.p, radius, O.max_cg_iterations, O.cg_rel_tol))) return fail(rc);
    reuse = true;
    pgo_step<<<1, PGS_THREADS, 0, st>>>(n_nodes, G.ids.p, G.row.p, G.inc.p, Hd, Ho, g, G.x.p, x, cand, G.stats.p);
    launched(ctx, "pgo_step");
    // the candidate is evaluated right away (one host synchronisation per iteration); an invalid step simply discards the evaluation
    if ((rc = pgo_enqueue_assemble(ctx, G, *params, cand, Hd_c, Ho_c, g_c))) return fail(rc);
    pgo_gradmax<<<1, PGS_THREADS, 0, st>>>(n_nodes, fixed_node, cand, g_c, G.stats.p);
    launched(ctx, "pgo_gradmax");
    e = cudaMemcpyAsync(h_stats, G.stats.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_it, G.it.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { set_error("tbv_pgo_optimize: %s", cudaGetErrorString(e)); return fail(TBV_ERR_CUDA); }
    double cand_cost = 0.0;
    if ((rc = read_eval(&cand_cost))) return fail(rc);
    if (h_it < 0) { set_error("tbv_pgo_optimize: a pivot block of the damped odometry chain is not positive definite"); return fail(TBV_ERR_INVALID); }
    S.cg_iterations += h_it;
    const double model_change = h_stats[0];
    if (!(h_stats[3] > 0.5 && model_change > 0.0)) {   // HandleInvalidStep
      invalid++;
      step_ok = false;
      if (invalid >= 5) { S.termination = TBV_PGO_FAILURE; break; }
      radius /= decrease; decrease *= 2.0;
      continue;
    }
    invalid = 0;
    if (!std::isfinite(cand_cost)) cand_cost = DBL_BIG;
    if (sqrt(h_stats[1]) <= O.parameter_tolerance * (x_norm + O.parameter_tolerance)) { S.termination = TBV_PGO_PARAMETER_TOLERANCE; break; }
    if (std::fabs(x_cost - cand_cost) <= O.function_tolerance * x_cost) { S.termination = TBV_PGO_FUNCTION_TOLERANCE; break; }
    const double rho = cand_cost < DBL_BIG ? (x_cost - cand_cost) / model_change : -1.0;
    if (rho > 1e-3) {                                    // HandleSuccessfulStep
      std::swap(x, cand); std::swap(Hd, Hd_c); std::swap(Ho, Ho_c); std::swap(g, g_c);
      x_cost = cand_cost;
      x_norm = sqrt(h_stats[5]);
      gmax = h_stats[4];
      step_ok = true;
      S.successful_steps++;
      if (x_cost < S.final_cost) S.final_cost = x_cost;
      const double t = 2.0 * rho - 1.0;
      radius = std::min(radius / std::max(1.0 / 3.0, 1.0 - t * t * t), 1e16);
      decrease = 2.0; reuse = false;
    } else {
      step_ok = false;
      radius /= decrease; decrease *= 2.0;
    }
  }
  e = cudaEventRecord(ev1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nodes, x, N7 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaEventElapsedTime(&S.device_ms, ev0, ev1);
  if (e != cudaSuccess) { set_error("tbv_pgo_optimize: %s", cudaGetErrorString(e)); return fail(TBV_ERR_CUDA); }
  if (summary) *summary = S;
  fail(TBV_OK);
  return TBV_OK;
}


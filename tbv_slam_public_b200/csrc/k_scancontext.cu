// k_scancontext.cu — K6 (radar Scan-Context descriptor + keys) and K7 (ring-key candidate search, descriptor distance).
//
// Replaces RSCManager::MakeRadarCloudContext (+ the lateral augmentations of makeAndSaveScancontextAndKeysRadarCloud),
// ExcludeAndUpdateLikelihood, L2norm, OdometryNNSearch (place_recognition_radar/src/place_recognition_radar/
// RadarScancontext.cpp:59-131, 156-182, 183-221, 251-284) and SCManager::{xy2theta, circshift, distDirectSC, fastAlignUsingVkey,
// distanceBtnScanContext, makeRingkeyFromScancontext, makeSectorkeyFromScancontext} (Scancontext.cpp:62-189, 239-268).
//
//   sc_make       one CTA per (cloud, lateral offset): points binned into a rings x sectors histogram in shared memory
//                 (intensity sum or max), divided, keys = sequential row / column means (the reference's summation order).
//   sc_similarity one thread per query keyframe: the two sequential odometry walks (exclusion window, travelled-distance
//                 likelihood) exactly as written in the reference.
//   sc_search     one CTA per query: float L2 over (ring key ++ 10*similarity) against every admissible older key,
//                 then the num_candidates_from_tree smallest by (distance, index).
//   sc_distance   one CTA per (query, candidate) descriptor pair: sector keys, coarse shift by key L2 (first minimum),
//                 fine search over the +-search_ratio window with the column-wise cosine distance (first minimum over the
//                 ascending shift list).
// All reductions that feed comparisons run in the reference's left-to-right order (built with -fmad=false).
#include <cfloat>
#include <cmath>

#include "tbv_common.cuh"

namespace tbv {

__device__ __forceinline__ float xy2theta_dev(float x, float y) {  // Scancontext.cpp:62-77; atan evaluated in double then narrowed
  const double k = 180.0 / 3.14159265358979323846;
  if (x >= 0 && y >= 0) return (float)(k * (double)(float)atan((double)(y / x)));
  if (x < 0 && y >= 0) return (float)(180 - (k * (double)(float)atan((double)(y / (-x)))));
  if (x < 0 && y < 0) return (float)(180 + (k * (double)(float)atan((double)(y / x))));
  if (x >= 0 && y < 0) return (float)(360 - (k * (double)(float)atan((double)((-y) / x))));
  return 0.f;
}

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= v) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
}

__global__ void __launch_bounds__(256)
sc_make(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ inten, int n, int R, int S, double max_radius,
        int desc_function, double divider, double no_point, const double* __restrict__ offsets, double* __restrict__ desc,
        float* __restrict__ ringkey, double* __restrict__ sectorkey) {
  extern __shared__ __align__(8) unsigned char s_raw[];
  double* acc = reinterpret_cast<double*>(s_raw);              // [R*S] column-major (sector * R + ring)
  unsigned char* touched = s_raw + (size_t)R * S * sizeof(double);
  const int o = blockIdx.x;
  const double ox = offsets[2 * o], oy = offsets[2 * o + 1];
  const bool shifted = (ox != 0.0) || (oy != 0.0);
  for (int i = threadIdx.x; i < R * S; i += blockDim.x) { acc[i] = 0.0; touched[i] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float px = x[i], py = y[i];
    if (shifted) {  // pcl::transformPointCloud with a pure translation: double arithmetic, narrowed per coordinate
      px = (float)(((1.0 * (double)px + 0.0 * (double)py) + 0.0 * 0.0) + ox);
      py = (float)(((0.0 * (double)x[i] + 1.0 * (double)py) + 0.0 * 0.0) + oy);
    }
    const float pi = inten[i];
    const float rng = sqrtf(px * px + py * py);
    const float ang = xy2theta_dev(px, py);
    if ((double)rng > max_radius) continue;
    const int ring = max(min(R, (int)ceil(((double)rng / max_radius) * (double)R)), 1);
    const int sect = max(min(S, (int)ceil(((double)ang / 360.0) * (double)S)), 1);
    const int b = (sect - 1) * R + (ring - 1);
    touched[b] = 1;
    if (desc_function == 0) atomicAdd(&acc[b], (double)pi);
    else atomic_max_double(&acc[b], (double)pi);
  }
  __syncthreads();
  double* d = desc + (size_t)o * R * S;
  for (int i = threadIdx.x; i < R * S; i += blockDim.x) {
    double v = touched[i] ? acc[i] : -1000.0;
    v = v / divider;                       // division BEFORE the no-point test (RadarScancontext.cpp:113-125)
    if (v == -1000.0) v = no_point;
    acc[i] = v;
    d[i] = v;
  }
  __syncthreads();
  if (ringkey)
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      double s = 0;
      for (int c = 0; c < S; c++) s += acc[(size_t)c * R + r];
      ringkey[(size_t)o * R + r] = (float)(s / S);
    }
  if (sectorkey)
    for (int c = threadIdx.x; c < S; c += blockDim.x) {
      double s = 0;
      for (int r = 0; r < R; r++) s += acc[(size_t)c * R + r];
      sectorkey[(size_t)o * S + c] = s / R;
    }
}

// ---- candidate search -----------------------------------------------------------------------------------------------------
struct Aff2 { double r00, r01, r10, r11, tx, ty; };
__device__ __forceinline__ Aff2 aff2_from(const double* p) {
  const double c = cos(p[2]), s = sin(p[2]);
  return Aff2{c, -s, s, c, p[0], p[1]};
}
__device__ __forceinline__ void aff2_rel_translation(const Aff2& A, const Aff2& B, double& tx, double& ty) {  // (A^-1 * B).translation()
  const double det = A.r00 * A.r11 - A.r01 * A.r10;
  const double invdet = 1.0 / det;
  const double i00 = A.r11 * invdet, i01 = -A.r01 * invdet, i10 = -A.r10 * invdet, i11 = A.r00 * invdet;
  const double itx = -(i00 * A.tx + i01 * A.ty), ity = -(i10 * A.tx + i11 * A.ty);
  tx = (i00 * B.tx + i01 * B.ty) + itx;
  ty = (i10 * B.tx + i11 * B.ty) + ity;
}

// one CTA per query: ExcludeAndUpdateLikelihood (:183-221) as of keyframe cur = q_current[q].  The travelled distance is a running
// sum from the query backwards — the reference adds the steps one by one, and so does this kernel (thread 0, out of shared memory,
// eight loads ahead of eight dependent adds), so every trav_i is bit-identical; the step lengths before it and the likelihoods
// after it (sqrt, division, exp per keyframe) are computed by all threads.
constexpr int SCS_CHUNK = 2048;
__global__ void __launch_bounds__(256)
sc_similarity(const double* __restrict__ odom, int n_q, const int* __restrict__ q_current, double sigma, double exclude_dist,
              int sim_stride, double* __restrict__ sim, int* __restrict__ n_exclude) {
  __shared__ double s_t[SCS_CHUNK];
  __shared__ double s_carry;
  const int q = blockIdx.x;
  if (q >= n_q) return;
  const int cur = q_current[q];
  if (threadIdx.x == 0) {
    int ne;
    if (cur + 1 <= 2) ne = 2;
    else {
      double distance = 0.0;
      ne = 0;
      Aff2 Tprev = aff2_from(odom + 3 * (size_t)cur);
      for (int i = cur; i >= 0 && distance < exclude_dist; i--) {
        const Aff2 Ti = aff2_from(odom + 3 * (size_t)i);
        double tx, ty;
        aff2_rel_translation(Tprev, Ti, tx, ty);
        distance = distance + sqrt(tx * tx + ty * ty);
        Tprev = Ti;
        ne++;
      }
    }
    n_exclude[q] = ne;
    s_carry = 0.0;
  }
  __syncthreads();
  const double cx = odom[3 * (size_t)cur], cy = odom[3 * (size_t)cur + 1];
  double* sq = sim + (size_t)q * sim_stride;
  // keyframes i = cur-1 ... 0 in chunks; slot j of a chunk is keyframe i = hi - j (descending, the order of the running sum)
  for (int hi = cur - 1; hi >= 0; hi -= SCS_CHUNK) {
    const int m = min(SCS_CHUNK, hi + 1);
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const int i = hi - j;
      const double tpx = odom[3 * (size_t)(i + 1)], tpy = odom[3 * (size_t)(i + 1) + 1];
      const double tix = odom[3 * (size_t)i], tiy = odom[3 * (size_t)i + 1];
      s_t[j] = sqrt((tpx - tix) * (tpx - tix) + (tpy - tiy) * (tpy - tiy));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double trav = s_carry;
      int j = 0;
      for (; j + 8 <= m; j += 8) {
        const double d0 = s_t[j], d1 = s_t[j + 1], d2 = s_t[j + 2], d3 = s_t[j + 3], d4 = s_t[j + 4], d5 = s_t[j + 5], d6 = s_t[j + 6], d7 = s_t[j + 7];
        trav += d0; s_t[j] = trav; trav += d1; s_t[j + 1] = trav; trav += d2; s_t[j + 2] = trav; trav += d3; s_t[j + 3] = trav;
        trav += d4; s_t[j + 4] = trav; trav += d5; s_t[j + 5] = trav; trav += d6; s_t[j + 6] = trav; trav += d7; s_t[j + 7] = trav;
      }
      for (; j < m; j++) { trav += s_t[j]; s_t[j] = trav; }
      s_carry = trav;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const int i = hi - j;
      const double tix = odom[3 * (size_t)i], tiy = odom[3 * (size_t)i + 1];
      const double trav = s_t[j];
      const double est = sqrt((cx - tix) * (cx - tix) + (cy - tiy) * (cy - tiy));
      const double error = fmax(est - 5.0, 0.0);
      const double rel = error / trav;
      const double prob = exp(-rel * rel / (2 * sigma * sigma));
      sq[i] = 1.0 - prob;
    }
    __syncthreads();
  }
}

// one CTA per query: OdometryNNSearch (:259-284)
__global__ void __launch_bounds__(256)
sc_search(const float* __restrict__ db_keys, int R, int n_q, const float* __restrict__ q_keys, const int* __restrict__ q_current,
          const int* __restrict__ n_exclude, const double* __restrict__ sim, int sim_stride, int coupled, int n_want, float* __restrict__ dist_scratch,
          int* __restrict__ cand_idx, double* __restrict__ cand_sim) {
  __shared__ float s_key[128];
  __shared__ float s_bd[8];
  __shared__ int s_bi[8];
  const int q = blockIdx.x;
  const int cur = q_current[q];
  const int n_search = max(0, cur - 1 - n_exclude[q]);
  for (int i = threadIdx.x; i < R; i += blockDim.x) s_key[i] = q_keys[(size_t)q * R + i];
  __syncthreads();
  float* dist = dist_scratch + (size_t)q * sim_stride;
  const double* sq = sim + (size_t)q * sim_stride;
  for (int idx = threadIdx.x; idx < n_search; idx += blockDim.x) {
    const float* k = db_keys + (size_t)idx * R;
    float l2 = 0.f;
    for (int i = 0; i < R; i++) {  // L2norm (:251-258): float difference, double square, float accumulator
      const double err = (double)(s_key[i] - k[i]);
      l2 = (float)((double)l2 + err * err);
    }
    if (coupled) {
      const float last = (float)(10 * sq[idx]);
      const double err = (double)(0.0f - last);
      l2 = (float)((double)l2 + err * err);
    }
    dist[idx] = l2;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = 0; t < n_want; t++) {
    float bd = FLT_MAX;
    int bi = 0x7fffffff;
    for (int idx = threadIdx.x; idx < n_search; idx += blockDim.x) {
      const float d = dist[idx];
      if (d < bd || (d == bd && idx < bi)) { bd = d; bi = idx; }
    }
    for (int dd = 16; dd > 0; dd >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, dd);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, dd);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) { s_bd[warp] = bd; s_bi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++)
        if (s_bd[w] < bd || (s_bd[w] == bd && s_bi[w] < bi)) { bd = s_bd[w]; bi = s_bi[w]; }
      const bool ok = bi != 0x7fffffff && t < n_search;
      cand_idx[(size_t)q * n_want + t] = ok ? bi : -1;
      cand_sim[(size_t)q * n_want + t] = ok ? sq[bi] : 0.0;
      if (ok) dist[bi] = FLT_MAX;  // taken; (FLT_MAX, idx) can never win again before the list is exhausted
      s_bi[0] = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_bi[0]) {
      for (int u = t + 1 + threadIdx.x; u < n_want; u += blockDim.x) { cand_idx[(size_t)q * n_want + u] = -1; cand_sim[(size_t)q * n_want + u] = 0.0; }
      break;
    }
    __syncthreads();
  }
}

// ---- descriptor distance --------------------------------------------------------------------------------------------------
constexpr int SCD_MAX_S = 512, SCD_MAX_SPACE = 64;

__global__ void __launch_bounds__(128)
sc_distance(const double* __restrict__ desc_q, const double* __restrict__ desc_c, const int* __restrict__ q_idx, const int* __restrict__ c_idx,
            int R, int S, int search_radius, double* __restrict__ out_dist, int* __restrict__ out_shift) {
  extern __shared__ __align__(8) unsigned char s_raw[];
  double* sc1 = reinterpret_cast<double*>(s_raw);   // [R*S]
  double* sc2 = sc1 + (size_t)R * S;                // [R*S]
  double* sim = sc2 + (size_t)R * S;                // [n_space][S]  per-shift, per-column similarity (NaN = skipped column)
  __shared__ double vk1[SCD_MAX_S], vk2[SCD_MAX_S], n1[SCD_MAX_S], n2[SCD_MAX_S], key_d[SCD_MAX_S];
  __shared__ int space[SCD_MAX_SPACE];
  __shared__ double space_d[SCD_MAX_SPACE];
  __shared__ int s_argmin;
  const int p = blockIdx.x;
  const double* g1 = desc_q + (size_t)q_idx[p] * R * S;
  const double* g2 = desc_c + (size_t)c_idx[p] * R * S;
  for (int i = threadIdx.x; i < R * S; i += blockDim.x) { sc1[i] = g1[i]; sc2[i] = g2[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < S; c += blockDim.x) {  // sector keys (column means) and column norms
    double s1 = 0, s2 = 0, q1 = 0, q2 = 0;
    for (int r = 0; r < R; r++) {
      const double a = sc1[(size_t)c * R + r], b = sc2[(size_t)c * R + r];
      s1 += a; s2 += b; q1 += a * a; q2 += b * b;
    }
    vk1[c] = s1 / R; vk2[c] = s2 / R; n1[c] = sqrt(q1); n2[c] = sqrt(q2);
  }
  __syncthreads();
  for (int sh = threadIdx.x; sh < S; sh += blockDim.x) {  // fastAlignUsingVkey: norm of the key difference per shift
    double s = 0;
    for (int c = 0; c < S; c++) {
      const double d = vk1[c] - vk2[(c - sh + S) % S];
      s += d * d;
    }
    key_d[sh] = sqrt(s);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int argmin = 0;
    double mn = 10000000;
    for (int sh = 0; sh < S; sh++)
      if (key_d[sh] < mn) { argmin = sh; mn = key_d[sh]; }
    // search space: argmin, +-1 .. +-radius, sorted ascending (duplicates kept, as std::sort keeps them)
    int m = 0;
    space[m++] = argmin;
    for (int ii = 1; ii < search_radius + 1; ii++) {
      space[m++] = (argmin + ii + S) % S;
      space[m++] = (argmin - ii + S) % S;
    }
    for (int a = 1; a < m; a++) {  // insertion sort
      const int v = space[a];
      int b = a - 1;
      while (b >= 0 && space[b] > v) { space[b + 1] = space[b]; b--; }
      space[b + 1] = v;
    }
    s_argmin = m;
  }
  __syncthreads();
  const int m = s_argmin;
  for (int t = threadIdx.x; t < m * S; t += blockDim.x) {  // distDirectSC, per (shift, column)
    const int si = t / S, col = t - si * S;
    const int c2 = (col - space[si] + S) % S;
    double dot = 0;
    for (int r = 0; r < R; r++) dot += sc1[(size_t)col * R + r] * sc2[(size_t)c2 * R + r];
    const double a = n1[col], b = n2[c2];
    sim[t] = (a == 0 || b == 0) ? nan("") : dot / (a * b);
  }
  __syncthreads();
  for (int si = threadIdx.x; si < m; si += blockDim.x) {
    int eff = 0;
    double sum = 0;
    for (int col = 0; col < S; col++) {
      const double v = sim[si * S + col];
      if (isnan(v)) continue;
      sum = sum + v;
      eff = eff + 1;
    }
    eff = max(eff, 1);
    space_d[si] = 1.0 - sum / eff;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int argmin_shift = 0;
    double mn = 10000000;
    for (int si = 0; si < m; si++)
      if (space_d[si] < mn) { argmin_shift = space[si]; mn = space_d[si]; }
    out_dist[p] = mn;
    out_shift[p] = argmin_shift;
  }
}

}  // namespace tbv

using namespace tbv;

namespace {
template <typename T>
struct Tmp {  // RAII device buffer for the host-pointer entry points
  DevBuf<T> b;
  ~Tmp() { b.release(); }
  int up(tbv_ctx* ctx, const T* h, size_t n) {
    int rc = b.reserve(n ? n : 1);
    if (rc) return rc;
    if (n && h) TBV_CUDA(cudaMemcpyAsync(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return TBV_OK;
  }
  int down(tbv_ctx* ctx, T* h, size_t n) {
    if (n && h) TBV_CUDA(cudaMemcpyAsync(h, b.p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return TBV_OK;
  }
};
}  // namespace

extern "C" {

int tbv_sc_make(tbv_ctx* ctx, const float* x, const float* y, const float* intensity, int n, const tbv_sc_params* p, int n_offsets,
                const double* offsets_xy, double* desc, float* ringkey, double* sectorkey) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && p && desc && offsets_xy && n >= 0 && n_offsets >= 1, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(n == 0 || (x && y && intensity), "null cloud");
  const int R = p->num_ring, S = p->num_sector;
  TBV_REQUIRE(R >= 1 && S >= 1 && (size_t)R * S <= 16384 && p->desc_divider != 0.0, "bad descriptor shape");
  Tmp<float> dx, dy, di, drk;
  Tmp<double> doff, ddesc, dsk;
  int rc;
  if ((rc = dx.up(ctx, x, n)) || (rc = dy.up(ctx, y, n)) || (rc = di.up(ctx, intensity, n)) || (rc = doff.up(ctx, offsets_xy, 2 * (size_t)n_offsets)) ||
      (rc = ddesc.up(ctx, nullptr, (size_t)n_offsets * R * S)) || (rc = drk.up(ctx, nullptr, (size_t)n_offsets * R)) ||
      (rc = dsk.up(ctx, nullptr, (size_t)n_offsets * S)))
    return rc;
  const size_t smem = (size_t)R * S * (sizeof(double) + 1) + 8;
  if ((rc = ensure_dyn_smem(ctx, sc_make, smem))) return rc;
  sc_make<<<n_offsets, 256, smem, ctx->stream>>>(dx.b.p, dy.b.p, di.b.p, n, R, S, p->max_radius, p->desc_function, p->desc_divider, p->no_point, doff.b.p,
                                                 ddesc.b.p, drk.b.p, dsk.b.p);
  launched(ctx, "sc_make");
  TBV_CUDA(cudaGetLastError());
  if ((rc = ddesc.down(ctx, desc, (size_t)n_offsets * R * S)) || (rc = drk.down(ctx, ringkey, (size_t)n_offsets * R)) ||
      (rc = dsk.down(ctx, sectorkey, (size_t)n_offsets * S)))
    return rc;
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

int tbv_sc_distance_batch(tbv_ctx* ctx, const double* desc_q, int n_q, const double* desc_c, int n_c, int n_pairs, const int* q_idx,
                          const int* c_idx, const tbv_sc_params* p, double* dist, int* shift) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && desc_q && desc_c && q_idx && c_idx && p && dist && shift && n_q >= 1 && n_c >= 1 && n_pairs >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  if (n_pairs == 0) return TBV_OK;
  const int R = p->num_ring, S = p->num_sector;
  const int radius = (int)std::round(0.5 * p->search_ratio * S);  // Scancontext.cpp:168
  TBV_REQUIRE(R >= 1 && S >= 1 && S <= SCD_MAX_S && 2 * radius + 1 <= SCD_MAX_SPACE, "descriptor shape outside the kernel limits");
  for (int i = 0; i < n_pairs; i++) TBV_REQUIRE(q_idx[i] >= 0 && q_idx[i] < n_q && c_idx[i] >= 0 && c_idx[i] < n_c, "pair index out of range");
  const size_t smem = ((size_t)2 * R * S + (size_t)(2 * radius + 1) * S) * sizeof(double);
  TBV_REQUIRE(smem <= 200 * 1024, "descriptor too large for the shared-memory distance kernel");
  Tmp<double> dq, dc, dd;
  Tmp<int> dqi, dci, dsh;
  int rc;
  if ((rc = dq.up(ctx, desc_q, (size_t)n_q * R * S)) || (rc = dc.up(ctx, desc_c, (size_t)n_c * R * S)) || (rc = dqi.up(ctx, q_idx, n_pairs)) ||
      (rc = dci.up(ctx, c_idx, n_pairs)) || (rc = dd.up(ctx, nullptr, n_pairs)) || (rc = dsh.up(ctx, nullptr, n_pairs)))
    return rc;
  if ((rc = ensure_dyn_smem(ctx, sc_distance, smem))) return rc;
  sc_distance<<<n_pairs, 128, smem, ctx->stream>>>(dq.b.p, dc.b.p, dqi.b.p, dci.b.p, R, S, radius, dd.b.p, dsh.b.p);
  launched(ctx, "sc_distance");
  TBV_CUDA(cudaGetLastError());
  if ((rc = dd.down(ctx, dist, n_pairs)) || (rc = dsh.down(ctx, shift, n_pairs))) return rc;
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

int tbv_sc_search(tbv_ctx* ctx, const float* db_keys, const double* odom_xyt, int n_db, int n_q, const float* q_keys, const int* q_current,
                  const tbv_sc_params* p, int* cand_idx, double* cand_odom_sim, int* n_exclude) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && db_keys && odom_xyt && q_keys && q_current && p && cand_idx && n_db >= 1 && n_q >= 0, "bad arguments");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  if (n_q == 0) return TBV_OK;
  const int R = p->num_ring, want = p->num_candidates_from_tree;
  TBV_REQUIRE(R >= 1 && R <= 128 && want >= 1, "bad search parameters");
  for (int i = 0; i < n_q; i++) TBV_REQUIRE(q_current[i] >= 0 && q_current[i] < n_db, "query keyframe outside the database");
  Tmp<float> dk, dqk, dscr;
  Tmp<double> dod, dsim, dcs;
  Tmp<int> dcur, dne, dci;
  int rc;
  if ((rc = dk.up(ctx, db_keys, (size_t)n_db * R)) || (rc = dod.up(ctx, odom_xyt, (size_t)n_db * 3)) || (rc = dqk.up(ctx, q_keys, (size_t)n_q * R)) ||
      (rc = dcur.up(ctx, q_current, n_q)) || (rc = dsim.up(ctx, nullptr, (size_t)n_q * n_db)) || (rc = dscr.up(ctx, nullptr, (size_t)n_q * n_db)) ||
      (rc = dne.up(ctx, nullptr, n_q)) || (rc = dci.up(ctx, nullptr, (size_t)n_q * want)) || (rc = dcs.up(ctx, nullptr, (size_t)n_q * want)))
    return rc;
  sc_similarity<<<n_q, 256, 0, ctx->stream>>>(dod.b.p, n_q, dcur.b.p, p->odom_sigma_error, p->distance_exclude_recent, n_db, dsim.b.p, dne.b.p);
  launched(ctx, "sc_similarity");
  sc_search<<<n_q, 256, 0, ctx->stream>>>(dk.b.p, R, n_q, dqk.b.p, dcur.b.p, dne.b.p, dsim.b.p, n_db, p->odometry_coupled_closure, want, dscr.b.p, dci.b.p,
                                          dcs.b.p);
  launched(ctx, "sc_search");
  TBV_CUDA(cudaGetLastError());
  if ((rc = dci.down(ctx, cand_idx, (size_t)n_q * want)) || (rc = dcs.down(ctx, cand_odom_sim, (size_t)n_q * want)) || (rc = dne.down(ctx, n_exclude, n_q)))
    return rc;
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return TBV_OK;
}

}  // extern "C"

// k_filter.cu — K1 (k-strongest + axial non-max suppression), K2 (polar -> Cartesian cloud), motion compensation.
//
// Replaces StructuredKStrongest::{FilterKstrongest, AxialNonMaxSupress, getPeaksFilteredPointCloud}
// (cfear_radarodometry/src/cfear_radarodometry/radar_filters.cpp:198-337) and CFEAR_Radarodometry::Compensate
// (cfear_radarodometry/src/cfear_radarodometry/utils.cpp:96-113).
//
// Design (sm_100a, HBM-bound byte scan): ONE kernel, k1_filter_fused, turns scans into the two ordered clouds — see the comment above the
// kernel.  The reference keeps, per azimuth row, the k largest (intensity, range) pairs under std::pair ordering, i.e. the k largest 24-bit
// keys (intensity << 16 | range), ascending; bins at range <= min_range_bin stay in that selection but are not emitted; a selected bin is a
// "peak" when its 7-tap score is not exceeded within +-3 bins (bytes across the row edge come from the neighbouring row of the flat
// cv::Mat, as the reference's indexing reads them).
#include <cmath>

#include "tbv_common.cuh"

namespace tbv {


__device__ __forceinline__ uint32_t ge_mask(uint32_t w, uint32_t addc, bool hi) {
  // 0x80 in every byte whose value >= z.  z <= 128: ((low7 + 128 - z) | w) & 0x80 ; z > 128: (low7 + 256 - z) & w & 0x80
  const uint32_t t = (w & 0x7f7f7f7fu) + addc;
  return hi ? (t & w & 0x80808080u) : ((t | w) & 0x80808080u);
}
__device__ __forceinline__ uint32_t eq_mask(uint32_t w, uint32_t v4) {
  // 0x80 in every byte equal to v (v4 = v replicated): bytes of (w ^ v4) that are zero
  const uint32_t x = w ^ v4;
  const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(t | x) & 0x80808080u;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one row into shared memory, completion on an mbarrier --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Stage row `rp` so that buf[a0 + r] = row[r], a0 = (row address) & 15.  Fast path: ONE bulk copy of the 16-byte-aligned
// superset [rp - a0, ceil16(rp + n_range)) issued by lane 0 — legal whenever the superset stays inside the caller's buffer
// [lo, hi) (every row but possibly the first / last of the whole batch).  Returns true if the copy is in flight on `bar`;
// false means the caller must copy synchronously (stage_row_sync) when it consumes the row.
__device__ __forceinline__ bool stage_row_tma(uint8_t* buf, const uint8_t* rp, int n_range, const uint8_t* lo, const uint8_t* hi, uint64_t* bar, int lane) {
  const int a0 = (int)(reinterpret_cast<uintptr_t>(rp) & 15u);
  const uint8_t* src = rp - a0;
  const uint32_t bytes = (uint32_t)((a0 + n_range + 15) & ~15);
  if (src < lo || src + bytes > hi) return false;
  if (lane == 0) {
    mbar_expect_tx(bar, bytes);
    tma_load_1d(buf, src, bytes, bar);
  }
  return true;
}
__device__ __forceinline__ void stage_row_sync(uint8_t* buf, const uint8_t* rp, int n_range, int lane) {
  const int a0 = (int)(reinterpret_cast<uintptr_t>(rp) & 15u);
  for (int o = lane; o < n_range; o += 32) buf[a0 + o] = __ldg(rp + o);
  __syncwarp();
}

// Compensate one point (utils.cpp:96-113, utils.h:28-32): (x, y) are the stored float coordinates; m = previous frame-to-frame
// motion (x, y, yaw).  Same operations, in the same order, as the reference's double arithmetic.
__device__ __forceinline__ void compensate_point(float& x, float& y, double m0, double m1, double m2, int ccw) {
  const double two_pi = __dmul_rn(2.0, 3.14159265358979323846);
  const double px = (double)x, py = (double)y;
  const double a = atan2(py, px);
  double d = __ddiv_rn((a > 0.00001 ? a : __dadd_rn(two_pi, a)), two_pi);
  d = ccw ? -(__dsub_rn(d, 0.5)) : __dsub_rn(d, 0.5);
  const double ang = __dmul_rn(d, m2);
  const double s1 = sin(ang), c1 = cos(ang);
  const double tx = __dmul_rn(d, m0), ty = __dmul_rn(d, m1);
  x = (float)__dadd_rn(__dadd_rn(__dmul_rn(c1, px), __dmul_rn(-s1, py)), tx);
  y = (float)__dadd_rn(__dadd_rn(__dmul_rn(s1, px), __dmul_rn(c1, py)), ty);
}

// ---- the fused filter kernel ----------------------------------------------------------------------------------------------------------
// One CTA (8 warps) per scan walks the scan's rows in chunks of 8 (one row per warp) and writes the two final clouds directly:
//   P1 (warp per row)      the row arrives by ONE bulk copy (TMA, mbarrier) in the warp's row buffer; branch-free conservative scan of the
//                          16-byte vectors (flags kept as a bit per lane and vector), flagged vector ids queued, exact byte masks of the queued
//                          vectors, candidates (key << 2 | emitted << 1) appended to the row's list.  Rows with more candidates than the list
//                          holds take the exact dense path (threshold bisection on the staged row + ties from the far end) and end with <= k.
//   P2 (half-warp per row) all-pairs rank inside the row's list -> the k strongest in ascending (intensity, range) order; 7+7-tap axial
//                          non-max suppression of the selected bins on the staged row; per-entry output offsets by ballot prefix.
//   P3 (one thread per point, all 256 threads) the emitted entries of the 8 rows, staged in output order: fp64 polar -> Cartesian with the
//                          host (glibc) cos/sin table, motion compensation, coalesced stores to both clouds at the scan's running offsets.
// The next chunk's rows are requested as soon as P2 has read the current ones, so the copies fly under P3's fp64 work and under the other
// resident CTAs (4 per SM).  Nothing but the scan bytes is read from HBM and nothing but the clouds is written: the per-row key arrays of
// the two-kernel version (and their round trip through L2) are gone.
constexpr int KF_WARPS = 8;                  // warps per CTA = rows per chunk
constexpr int KF_THREADS = KF_WARPS * 32;
constexpr int KF_CAP = 128;                  // per-row candidate list capacity = largest supported k

// list entry: bits 2..25 = intensity << 16 | range (the reference's std::pair<uchar,int> order), bit 1 = emitted (range > min_range_bin),
// bit 0 = peak (set in P2).  Distinct entries of a row have distinct keys, so comparing entries compares keys.
__device__ __forceinline__ uint32_t kf_entry(uint32_t inten, int r, int min_range_bin) {
  return (((inten << 16) | (uint32_t)r) << 2) | (r > min_range_bin ? 2u : 0u);
}

// HI: z_min > 128 (selects the form of the conservative byte compare at compile time: the scan loop carries no branch on it);
// NG: number of 32-vector groups of a staged row, unrolled at compile time (8: Oxford, 7: MulRan); 0 = run-time loop over n_groups.
template <bool HI, int NG>
__global__ void __launch_bounds__(KF_THREADS, 4)
k1_filter_fused(const uint8_t* __restrict__ polar, int n_az, int n_range, size_t row_stride, int z_min, int k, int want_peaks, int rowbuf, int n_groups,
                const uint8_t* buf_lo, const uint8_t* buf_hi, int min_range_bin, double range_res, const double2* __restrict__ cs_table, int cap,
                float* __restrict__ fx, float* __restrict__ fy, uint8_t* __restrict__ fi, uint16_t* __restrict__ faz, uint16_t* __restrict__ frg,
                int* __restrict__ fcount,
                float* __restrict__ px, float* __restrict__ py, uint8_t* __restrict__ pi, uint16_t* __restrict__ paz, uint16_t* __restrict__ prg,
                int* __restrict__ pcount, const double* __restrict__ mot, int ccw) {
  extern __shared__ __align__(128) uint8_t s_dyn[];               // [KF_WARPS][rowbuf] staged rows
  __shared__ __align__(16) uint32_t s_list[KF_WARPS][KF_CAP + 4];  // candidates of the row (unordered), zero-padded to a multiple of 4
  __shared__ __align__(16) uint32_t s_sel[KF_WARPS][KF_CAP];       // P1: queue of flagged vectors (u16); P2: selected entries, ascending
  __shared__ uint32_t s_idx[KF_WARPS][KF_CAP];                     // P2: per selected entry (index among emitted) | (index among peaks) << 16
  __shared__ uint32_t s_stkey[KF_WARPS * KF_CAP];                  // P3 staging, output order: entry
  __shared__ int16_t s_stpp[KF_WARPS * KF_CAP];                    //   position among the chunk's peaks, or -1
  __shared__ uint8_t s_strow[KF_WARPS * KF_CAP];                   //   row within the chunk
  __shared__ __align__(8) uint64_t s_bar[KF_WARPS];
  __shared__ int s_n[KF_WARPS], s_cntf[KF_WARPS], s_cntp[KF_WARPS], s_dn[KF_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const int scan = blockIdx.x;
  const uint32_t z = (uint32_t)z_min;
  const bool zero_thr = z == 0;   // every byte is a candidate: no sparse pass
  const uint32_t addc = (HI ? (256u - z) : (128u - z)) * 0x01010101u;
  const size_t scan_bytes = (size_t)(n_az - 1) * row_stride + (size_t)n_range;  // addressable bytes of one scan
  const uint8_t* scan_base = polar + (size_t)scan * (size_t)n_az * row_stride;
  uint8_t* buf = s_dyn + (size_t)warp * rowbuf;
  const uint4* vbuf = reinterpret_cast<const uint4*>(buf);
  uint32_t* list = s_list[warp];
  uint64_t* bar = &s_bar[warp];
  if (lane == 0) mbar_init(bar, 1);
  // The row buffer starts out zero: the scan loop runs over whole groups of 32 vectors, and the vectors past a row's staged superset are
  // never written by a bulk copy (the one right behind it is re-zeroed per row, see below).
  for (int o = lane * 16; o < rowbuf; o += 32 * 16) *reinterpret_cast<uint4*>(buf + o) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;
  bool in_flight = false;
  if (warp < n_az) in_flight = stage_row_tma(buf, scan_base + (size_t)warp * row_stride, n_range, buf_lo, buf_hi, bar, lane);
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  if (mot) { m0 = mot[scan * 3 + 0]; m1 = mot[scan * 3 + 1]; m2 = mot[scan * 3 + 2]; }
  const double range_res_half = range_res / 2.0;
  const size_t cbase = (size_t)scan * cap;
  int base_f = 0, base_p = 0;   // running output offsets of the scan (every thread keeps its own copy)

  for (int row0 = 0; row0 < n_az; row0 += KF_WARPS) {
    // =============================== P1: warp per row ===============================================================================
    const int row = row0 + warp;
    int n = 0;
    int a0 = 0;
    if (row < n_az) {
      const uint8_t* rp = scan_base + (size_t)row * row_stride;
      a0 = (int)(reinterpret_cast<uintptr_t>(rp) & 15u);
      if (in_flight) { mbar_wait(bar, parity); parity ^= 1u; }
      else stage_row_sync(buf, rp, n_range, lane);
      const int lo_b = a0, hi_b = a0 + n_range;  // valid buffer byte range [lo_b, hi_b)
      const int nvec = (hi_b + 15) >> 4;         // 16-byte vectors covering [0, hi_b)
      // a previous row with another alignment may have ended one vector later: that vector must not be seen by the guard-free scan
      if (lane == 0 && (nvec << 4) < rowbuf) *reinterpret_cast<uint4*>(buf + (nvec << 4)) = make_uint4(0, 0, 0, 0);
      __syncwarp();
      auto valid_mask = [&](int wo) -> uint32_t {   // word at buffer offset wo: 0x80 per byte that belongs to the row
        uint32_t m = 0x80808080u;
        if (wo < lo_b) m &= (lo_b - wo >= 4) ? 0u : (0x80808080u << (8 * (lo_b - wo)));
        if (wo + 4 > hi_b) m &= (hi_b - wo <= 0) ? 0u : (0x80808080u >> (8 * (wo + 4 - hi_b)));
        return m;
      };
      // exact "byte >= t" masks of one vector; the staged superset starts / ends up to 15 bytes outside the row: masked at both ends
      auto masks = [&](const uint4 v, int wo, uint32_t ac, bool h, uint32_t m[4]) {
        m[0] = ge_mask(v.x, ac, h); m[1] = ge_mask(v.y, ac, h); m[2] = ge_mask(v.z, ac, h); m[3] = ge_mask(v.w, ac, h);
        if (wo < lo_b || wo + 16 > hi_b) { m[0] &= valid_mask(wo); m[1] &= valid_mask(wo + 4); m[2] &= valid_mask(wo + 8); m[3] &= valid_mask(wo + 12); }
      };
      // Conservative test "some byte of the 16 may be >= z_min" (never misses one; false positives only next to a byte >= 188):
      // z <= 128: byte + (128 - z) sets bit 7, or overflows the byte only when the byte itself has bit 7 set; z > 128: bit 7.
      auto any_ge = [&](const uint4 v) -> bool {
        if (HI) return ((v.x | v.y | v.z | v.w) & 0x80808080u) != 0;
        const uint32_t a = (v.x + addc) | v.x, b = (v.y + addc) | v.y, c = (v.z + addc) | v.z, d = (v.w + addc) | v.w;
        return ((a | b | c | d) & 0x80808080u) != 0;
      };
      // ---- scan: one flag bit per (lane, group); no ballots, no stores in the loop -----------------------------------------------------
      bool dense = zero_thr;
      uint16_t* queue = reinterpret_cast<uint16_t*>(s_sel[warp]);   // <= KF_CAP flagged vector ids
      int nq = 0;
      if (!dense) {
        uint32_t bits = 0;
        if (NG > 0) {
#pragma unroll
          for (int j = 0; j < NG; j++)
            if (any_ge(vbuf[j * 32 + lane])) bits |= 1u << j;
        } else {
#pragma unroll 4
          for (int j = 0; j < n_groups; j++) bits |= (any_ge(vbuf[j * 32 + lane]) ? 1u : 0u) << j;
        }
        const int mine = __popc(bits);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t2 = __shfl_up_sync(FULL, incl, d);
          if (lane >= d) incl += t2;
        }
        nq = __shfl_sync(FULL, incl, 31);
        if (nq > KF_CAP) dense = true;
        else {
          int q = incl - mine;
          while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            queue[q++] = (uint16_t)(j * 32 + lane);
          }
        }
        __syncwarp();
      }
      // ---- exact masks of the queued vectors, one vector per lane per round; list positions from a warp scan of the per-vector counts ----
      for (int qb = 0; qb < nq && !dense; qb += 32) {
        int c = 0, wo = 0;
        uint32_t m16 = 0;   // bit b: byte b of the vector is a candidate
        if (qb + lane < nq) {
          const int t = queue[qb + lane];
          wo = t << 4;
          uint32_t m[4];
          masks(vbuf[t], wo, addc, HI, m);
          // movemask of 4 bytes: bits 7, 15, 23, 31 -> bits 0..3 (multiply gathers them at bits 21..24; no carries, all partial products distinct)
          m16 = (((m[0] >> 7) * 0x00204081u) >> 21 & 0xfu) | (((m[1] >> 7) * 0x00204081u) >> 17 & 0xf0u) |
                (((m[2] >> 7) * 0x00204081u) >> 13 & 0xf00u) | (((m[3] >> 7) * 0x00204081u) >> 9 & 0xf000u);
          c = __popc(m16);
        }
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t2 = __shfl_up_sync(FULL, incl, d);
          if (lane >= d) incl += t2;
        }
        const int round_total = __shfl_sync(FULL, incl, 31);
        if (n + round_total > KF_CAP) { dense = true; break; }
        int pos = n + incl - c;
        while (m16) {
          const int b = __ffs(m16) - 1;
          m16 &= m16 - 1;
          list[pos++] = kf_entry(buf[wo + b], wo + b - a0, min_range_bin);
        }
        n += round_total;
      }
      if (dense) {
        // ---- dense row: exact threshold T = k-th largest intensity (bisection over the staged row), everything above T, then the ties at
        // T from the far end — the largest ranges win, exactly as the reference's erase(begin()) leaves them --------------------------------
        if (lane == 0) s_dn[warp] = 0;
        __syncwarp();
        auto count_ge = [&](uint32_t t) -> int {   // bytes of the row >= t (t in 1..255; t == 0 never asked)
          const bool h = t > 128;
          const uint32_t ac = (h ? (256u - t) : (128u - t)) * 0x01010101u;
          uint32_t c = 0;   // 128 x the count: dp4a sums the four flag bytes (0x80 each) of a word in one instruction
          for (int q = lane; q < nvec; q += 32) {
            uint32_t m[4];
            masks(vbuf[q], q << 4, ac, h, m);
            c = __dp4a(m[0], 0x01010101u, c); c = __dp4a(m[1], 0x01010101u, c);
            c = __dp4a(m[2], 0x01010101u, c); c = __dp4a(m[3], 0x01010101u, c);
          }
          return (int)(__reduce_add_sync(FULL, c) >> 7);
        };
        // T = the k-th largest intensity among the candidates (or z_min when the row holds no more than k candidates: all are taken)
        auto cnt = [&](uint32_t t) -> int { return t == 0 ? n_range : count_ge(t); };
        uint32_t T = z;
        int n_ge = cnt(z), n_gt = -1;
        if (n_ge > k) {
          uint32_t lo = z, hi_t = 256;   // count(>= lo) >= k, count(>= hi_t) < k
          int c_hi = 0;
          while (hi_t - lo > 1) {
            const uint32_t mid = (lo + hi_t) >> 1;
            const int c = count_ge(mid);
            if (c >= k) { lo = mid; n_ge = c; } else { hi_t = mid; c_hi = c; }
          }
          T = lo;
          n_gt = c_hi;                   // hi_t == T + 1 (0 entries above 255)
        }
        if (n_gt < 0) n_gt = (T >= 255) ? 0 : count_ge(T + 1);
        const int want = n_ge < k ? n_ge : k;
        const int need = want - n_gt;   // ties to take at T, largest ranges first (>= 0)
        if (T < 255 && n_gt > 0) {      // (1) everything strictly above the threshold
          const uint32_t t1 = T + 1;
          const bool h = t1 > 128;
          const uint32_t ac = (h ? (256u - t1) : (128u - t1)) * 0x01010101u;
          for (int q = lane; q < nvec; q += 32) {
            const int wo = q << 4;
            uint32_t m[4];
            masks(vbuf[q], wo, ac, h, m);
            uint32_t m16 = (((m[0] >> 7) * 0x00204081u) >> 21 & 0xfu) | (((m[1] >> 7) * 0x00204081u) >> 17 & 0xf0u) |
                           (((m[2] >> 7) * 0x00204081u) >> 13 & 0xf00u) | (((m[3] >> 7) * 0x00204081u) >> 9 & 0xf000u);
            if (m16) {
              int pos = atomicAdd(&s_dn[warp], __popc(m16));
              while (m16) {
                const int b = __ffs(m16) - 1;
                m16 &= m16 - 1;
                if (pos < KF_CAP) list[pos] = kf_entry(buf[wo + b], wo + b - a0, min_range_bin);
                pos++;
              }
            }
          }
        }
        // (2) ties at T, scanning ranges from the far end; 32 vectors (512 bytes) per step
        const uint32_t T4 = T * 0x01010101u;
        int carry = 0;
        for (int base = ((nvec - 1) >> 5) << 5; base >= 0 && carry < need; base -= 32) {
          const int q = base + lane;
          uint32_t m16 = 0;
          if (q < nvec) {
            const uint4 v = vbuf[q];
            const int wo = q << 4;
            const uint32_t e0 = eq_mask(v.x, T4) & valid_mask(wo), e1 = eq_mask(v.y, T4) & valid_mask(wo + 4);
            const uint32_t e2 = eq_mask(v.z, T4) & valid_mask(wo + 8), e3 = eq_mask(v.w, T4) & valid_mask(wo + 12);
            m16 = (((e0 >> 7) * 0x00204081u) >> 21 & 0xfu) | (((e1 >> 7) * 0x00204081u) >> 17 & 0xf0u) |
                  (((e2 >> 7) * 0x00204081u) >> 13 & 0xf00u) | (((e3 >> 7) * 0x00204081u) >> 9 & 0xf000u);
          }
          const int c = __popc(m16);
          int suf = c;  // inclusive suffix sum over lanes (higher lane = larger range)
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v2 = __shfl_down_sync(FULL, suf, d);
            if (lane + d < 32) suf += v2;
          }
          const int after = carry + suf - c;
          int take = need - after;
          if (take > c) take = c;
          if (take > 0) {
            int pos = atomicAdd(&s_dn[warp], take);
            const int wo = q << 4;
            while (take > 0) {  // highest bytes first
              const int b = 31 - __clz(m16);
              m16 &= ~(1u << b);
              if (pos < KF_CAP) list[pos] = kf_entry(T, wo + b - a0, min_range_bin);
              pos++;
              take--;
            }
          }
          carry += __shfl_sync(FULL, suf, 0);
        }
        __syncwarp();
        n = min(s_dn[warp], KF_CAP);   // == want
      }
      __syncwarp();
      if (lane < 4) list[n + lane] = 0;   // pad to a multiple of 4 with zeros (never greater than an entry)
    }
    if (lane == 0) s_n[warp] = n;
    __syncthreads();   // ---- (1) lists complete ---------------------------------------------------------------------------------------

    // =============================== P2: half-warp per row ==========================================================================
    if (tid < KF_WARPS * 16) {
      const int r8 = tid >> 4, hl = tid & 15;
      const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));
      const int nr = s_n[r8];
      const int n_sel = nr < k ? nr : k;
      const uint32_t* lst = s_list[r8];
      uint32_t* sel = s_sel[r8];
      // all-pairs rank: rank = number of strictly larger entries; selected iff rank < k; position n_sel - 1 - rank (ascending order)
      for (int e = hl; e < nr; e += 16) {
        const uint32_t mine = lst[e];
        int rank = 0;
        for (int q = 0; q < nr; q += 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(lst + q);
          rank += (v.x > mine) + (v.y > mine) + (v.z > mine) + (v.w > mine);
        }
        if (rank < k) sel[n_sel - 1 - rank] = mine;
      }
      __syncwarp(hmask);
      // axial non-max suppression of the emitted selected bins (radar_filters.cpp:238-298) on the staged row of warp r8
      if (want_peaks) {
        const int rowp = row0 + r8;
        const uint8_t* rb = s_dyn + (size_t)r8 * rowbuf;
        const int ra0 = (int)(reinterpret_cast<uintptr_t>(scan_base + (size_t)rowp * row_stride) & 15u);
        for (int e = hl; e < n_sel; e += 16) {
          const uint32_t ent = sel[e];
          if (!(ent & 2u)) continue;   // not emitted: its peak flag is never read
          const int r = (int)((ent >> 2) & 0xffffu);
          const bool in_band = (r >= 3) && (r < n_range - 3);
          int B[13];
          if (r >= 6 && r + 6 < n_range) {
#pragma unroll
            for (int t = 0; t < 13; t++) B[t] = rb[ra0 + r - 6 + t];
          } else {
            const long long gbase = (long long)rowp * (long long)row_stride;
#pragma unroll
            for (int t = 0; t < 13; t++) {
              const int q = r - 6 + t;
              if (q >= 0 && q < n_range) B[t] = rb[ra0 + q];
              else {
                const long long fidx = gbase + q;  // flat index into the scan buffer, as cv::Mat::at(bearing, r_nn) addresses it
                B[t] = (fidx >= 0 && fidx < (long long)scan_bytes) ? (int)__ldg(scan_base + fidx) : 0;
              }
            }
          }
          int s[7];  // s[i] = score at r - 3 + i = sum of the 7 bytes centred there (sliding window)
          s[0] = B[0] + B[1] + B[2] + B[3] + B[4] + B[5] + B[6];
#pragma unroll
          for (int i = 1; i < 7; i++) s[i] = s[i - 1] - B[i - 1] + B[i + 6];
          if (!in_band) {
            // a score exists only where some selected in-band bin lies within 3 of the position
#pragma unroll
            for (int i = 0; i < 7; i++) {
              const int p = r - 3 + i;
              bool computed = false;
              for (int q = 0; q < n_sel; q++) {
                const int r2 = (int)((sel[q] >> 2) & 0xffffu);
                if (r2 >= 3 && r2 < n_range - 3 && r2 - p <= 3 && p - r2 <= 3) { computed = true; break; }
              }
              if (!computed) s[i] = 0;
            }
          }
          bool largest = true;
#pragma unroll
          for (int i = 1; i <= 3; i++)
            if (s[3 - i] > s[3] || s[3] < s[3 + i]) largest = false;
          if (largest) sel[e] = ent | 1u;   // other lanes read only the range field of this entry
        }
      }
      __syncwarp(hmask);
      // output offsets inside the row: ballot prefix over the ascending entries (emitted, emitted peaks)
      int run_f = 0, run_p = 0;
      const unsigned lt = (1u << hl) - 1u;
      for (int b0 = 0; b0 < n_sel; b0 += 16) {
        const int e = b0 + hl;
        const uint32_t ent = e < n_sel ? sel[e] : 0u;
        const bool fe = (ent & 2u) != 0, fp = (ent & 3u) == 3u;
        const unsigned bf = (__ballot_sync(hmask, fe) >> (16 * ((tid >> 4) & 1))) & 0xffffu;
        const unsigned bp = (__ballot_sync(hmask, fp) >> (16 * ((tid >> 4) & 1))) & 0xffffu;
        if (e < n_sel) s_idx[r8][e] = (uint32_t)(run_f + __popc(bf & lt)) | ((uint32_t)(run_p + __popc(bp & lt)) << 16);
        run_f += __popc(bf);
        run_p += __popc(bp);
      }
      if (hl == 0) { s_cntf[r8] = run_f; s_cntp[r8] = run_p; }
    }
    __syncthreads();   // ---- (2) rows consumed, per-row counts known ---------------------------------------------------------------------

    // request the next chunk's rows: they fly under the staging and the fp64 work below
    {
      const int nrow = row0 + KF_WARPS + warp;
      in_flight = false;
      if (nrow < n_az) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads / writes of this buffer before the bulk copy into it
        in_flight = stage_row_tma(buf, scan_base + (size_t)nrow * row_stride, n_range, buf_lo, buf_hi, bar, lane);
      }
    }
    int tot_f = 0, tot_p = 0;
    if (tid < KF_WARPS * 16) {   // staging in output order (rows ascending, entries ascending inside a row)
      const int r8 = tid >> 4, hl = tid & 15;
      int off_f = 0, off_p = 0;
#pragma unroll
      for (int j = 0; j < KF_WARPS - 1; j++) { off_f += j < r8 ? s_cntf[j] : 0; off_p += j < r8 ? s_cntp[j] : 0; }
      const int nr = s_n[r8];
      const int n_sel = nr < k ? nr : k;
      for (int e = hl; e < n_sel; e += 16) {
        const uint32_t ent = s_sel[r8][e];
        if (ent & 2u) {
          const uint32_t ix = s_idx[r8][e];
          const int q = off_f + (int)(ix & 0xffffu);
          s_stkey[q] = ent;
          s_strow[q] = (uint8_t)r8;
          s_stpp[q] = (ent & 1u) ? (int16_t)(off_p + (int)(ix >> 16)) : (int16_t)-1;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KF_WARPS; j++) { tot_f += s_cntf[j]; tot_p += s_cntp[j]; }
    __syncthreads();   // ---- (3) staging complete -----------------------------------------------------------------------------------------

    // =============================== P3: one thread per emitted point =================================================================
    for (int i = tid; i < tot_f; i += KF_THREADS) {
      const uint32_t ent = s_stkey[i];
      const int az = row0 + (int)s_strow[i];
      const int r = (int)((ent >> 2) & 0xffffu);
      const double2 cs = cs_table[az];
      const double rho = __dadd_rn(range_res_half, __dmul_rn(range_res, (double)r));  // radar_filters.cpp:329-330
      float x = (float)__dmul_rn(rho, cs.x);
      float y = (float)__dmul_rn(rho, cs.y);
      if (mot) compensate_point(x, y, m0, m1, m2, ccw);
      const uint8_t inten = (uint8_t)((ent >> 18) & 0xffu);
      const int q = base_f + i;
      if (q < cap) {
        fx[cbase + q] = x; fy[cbase + q] = y; fi[cbase + q] = inten; faz[cbase + q] = (uint16_t)az; frg[cbase + q] = (uint16_t)r;
      }
      const int pp = s_stpp[i];
      if (pp >= 0 && base_p + pp < cap) {
        const int qp = base_p + pp;
        px[cbase + qp] = x; py[cbase + qp] = y; pi[cbase + qp] = inten; paz[cbase + qp] = (uint16_t)az; prg[cbase + qp] = (uint16_t)r;
      }
    }
    base_f += tot_f;
    base_p += tot_p;
    // no barrier here: the next chunk's P1 touches only the row buffers / lists, and its barrier (1) orders P3's staging reads before the
    // next staging writes
  }
  if (tid == 0) {
    fcount[scan] = base_f < cap ? base_f : cap;
    if (want_peaks) pcount[scan] = base_p < cap ? base_p : cap;
  }
}

// Compensate (utils.cpp:96-113, utils.h:28-32): one thread per point; mot = previous frame-to-frame motion (x, y, yaw).
__global__ void k_compensate(float* __restrict__ x, float* __restrict__ y, const int* __restrict__ count, int cap,
                             const double* __restrict__ mot /*[batch][3]*/, int ccw) {
  const int scan = blockIdx.y;
  const int n = count ? min(count[scan], cap) : cap;
  const double m0 = mot[scan * 3 + 0], m1 = mot[scan * 3 + 1], m2 = mot[scan * 3 + 2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = (size_t)scan * cap + i;
    float fx = x[q], fy = y[q];
    compensate_point(fx, fy, m0, m1, m2, ccw);
    x[q] = fx;
    y[q] = fy;
  }
}

// --------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------
int ensure_cs_table(tbv_ctx* ctx, int n_az) {
  FilterState& F = ctx->filt;
  if (F.cs_n_az == n_az) return TBV_OK;
  int rc = F.cs_table.reserve(n_az);
  if (rc) return rc;
  std::vector<double2> h(n_az);
  for (int b = 0; b < n_az; b++) {
    const double theta = (double(b + 1) / n_az) * 2. * M_PI;  // radar_filters.cpp:317 — glibc cos/sin on the host keeps x,y bit-exact
    h[b].x = std::cos(theta);
    h[b].y = std::sin(theta);
  }
  TBV_CUDA(cudaMemcpyAsync(F.cs_table.p, h.data(), n_az * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  F.cs_n_az = n_az;
  return TBV_OK;
}

int filter_kstrongest_dev(tbv_ctx* ctx, const uint8_t* polar_dev, int n_az, int n_range, size_t row_stride, int batch,
                          const tbv_filter_params* p, int want_peaks, const double* mot_dev, int ccw) {
  TBV_REQUIRE(ctx && polar_dev && p, "null pointer");
  TBV_REQUIRE(n_az > 0 && n_range > 0 && batch > 0 && row_stride >= (size_t)n_range, "bad image shape");
  TBV_REQUIRE(n_range <= 8192, "n_range > 8192 is not supported");
  TBV_REQUIRE(n_az <= 4096, "n_az > 4096 is not supported");
  TBV_REQUIRE(p->k_strongest >= 1 && p->k_strongest <= KF_CAP, "k_strongest must be in [1,128]");
  const int z_min = (int)p->z_min;  // float -> int as StructuredKStrongest's ctor does (radar_filters.h:86)
  TBV_REQUIRE(z_min >= 0 && z_min <= 255, "z_min must be in [0,255]");
  FilterState& F = ctx->filt;
  const int k = p->k_strongest;
  int rc;
  if ((rc = F.filtered.reserve(batch, n_az * k))) return rc;
  if (want_peaks && (rc = F.peaks.reserve(batch, n_az * k))) return rc;
  if ((rc = ensure_cs_table(ctx, n_az))) return rc;
  F.batch = batch; F.n_az = n_az; F.n_range = n_range; F.k = k;

  // staged row: the 16-byte-aligned superset of a row (<= 15 bytes of slack at either end) + one spare vector, in whole groups of 32 vectors
  const int nvec_max = (15 + n_range + 15) >> 4;
  const int n_groups = (nvec_max + 1 + 31) / 32;
  const int rowbuf = n_groups * 512;
  const size_t k1_smem = (size_t)KF_WARPS * rowbuf;
  const uint8_t* buf_hi = polar_dev + (size_t)(batch - 1) * n_az * row_stride + (size_t)(n_az - 1) * row_stride + (size_t)n_range;
  const double rr = (double)p->range_res;                                   // widened float (radar_filters.h:86)
  const int min_range_bin = (int)std::ceil((double)p->min_distance / rr);   // radar_filters.cpp:315
  auto launch = [&](auto kern) -> int {
    const int rc2 = ensure_dyn_smem(ctx, kern, k1_smem);
    if (rc2) return rc2;
    kern<<<batch, KF_THREADS, k1_smem, ctx->stream>>>(polar_dev, n_az, n_range, row_stride, z_min, k, want_peaks, rowbuf, n_groups, polar_dev, buf_hi,
                                                     min_range_bin, rr, F.cs_table.p, n_az * k, F.filtered.x.p, F.filtered.y.p, F.filtered.inten.p,
                                                     F.filtered.az.p, F.filtered.rg.p, F.filtered.count.p, F.peaks.x.p, F.peaks.y.p, F.peaks.inten.p,
                                                     F.peaks.az.p, F.peaks.rg.p, F.peaks.count.p, mot_dev, ccw);
    return TBV_OK;
  };
  if (z_min > 128) rc = n_groups == 8 ? launch(k1_filter_fused<true, 8>) : n_groups == 7 ? launch(k1_filter_fused<true, 7>) : launch(k1_filter_fused<true, 0>);
  else rc = n_groups == 8 ? launch(k1_filter_fused<false, 8>) : n_groups == 7 ? launch(k1_filter_fused<false, 7>) : launch(k1_filter_fused<false, 0>);
  if (rc) return rc;
  launched(ctx, "k1_filter_fused");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

int compensate_clouds_dev(tbv_ctx* ctx, DevCloud& c, const double* mot_dev, int ccw) {
  dim3 grid((c.cap + 255) / 256 < 32 ? (c.cap + 255) / 256 : 32, c.batch);
  k_compensate<<<grid, 256, 0, ctx->stream>>>(c.x.p, c.y.p, c.count.p, c.cap, mot_dev, ccw);
  launched(ctx, "k_compensate");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

int fetch_cloud(tbv_ctx* ctx, const DevCloud& d, int batch, tbv_points* out) {
  if (!out) return TBV_OK;
  TBV_REQUIRE(out->capacity > 0 && out->count, "tbv_points needs capacity and count");
  std::vector<int> cnt(batch);
  TBV_CUDA(cudaMemcpyAsync(cnt.data(), d.count.p, batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc = TBV_OK;
  for (int b = 0; b < batch; b++) {
    out->count[b] = cnt[b];
    int n = cnt[b];
    if (n > out->capacity) { n = out->capacity; rc = TBV_ERR_CAPACITY; set_error("tbv_points capacity %d < %d points", out->capacity, cnt[b]); }
    const size_t so = (size_t)b * d.cap, dofs = (size_t)b * out->capacity;
    if (n == 0) continue;
    if (out->x) TBV_CUDA(cudaMemcpyAsync(out->x + dofs, d.x.p + so, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->y) TBV_CUDA(cudaMemcpyAsync(out->y + dofs, d.y.p + so, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->intensity) TBV_CUDA(cudaMemcpyAsync(out->intensity + dofs, d.inten.p + so, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->azimuth) TBV_CUDA(cudaMemcpyAsync(out->azimuth + dofs, d.az.p + so, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->range) TBV_CUDA(cudaMemcpyAsync(out->range + dofs, d.rg.p + so, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  return rc;
}

}  // namespace tbv

using namespace tbv;

extern "C" {

int tbv_filter_kstrongest_dev(tbv_ctx* ctx, const uint8_t* polar_dev, int n_az, int n_range, size_t row_stride, int batch,
                              const tbv_filter_params* params, int want_peaks) {
  TBV_ENTER(ctx);
  return filter_kstrongest_dev(ctx, polar_dev, n_az, n_range, row_stride, batch, params, want_peaks);
}

int tbv_filter_fetch(tbv_ctx* ctx, tbv_points* out_filtered, tbv_points* out_peaks) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && ctx->filt.batch > 0, "no filter result on the device");
  int rc = fetch_cloud(ctx, ctx->filt.filtered, ctx->filt.batch, out_filtered);
  if (rc) return rc;
  return fetch_cloud(ctx, ctx->filt.peaks, ctx->filt.batch, out_peaks);
}

int tbv_filter_kstrongest(tbv_ctx* ctx, const uint8_t* polar, int n_az, int n_range, size_t row_stride, int batch,
                          const tbv_filter_params* params, tbv_points* out_filtered, tbv_points* out_peaks) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && polar && params && out_filtered, "null pointer");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  TBV_REQUIRE(n_az > 0 && n_range > 0 && batch > 0 && row_stride >= (size_t)n_range, "bad image shape");
  const size_t bytes = (size_t)batch * n_az * row_stride;
  int rc = ctx->filt.polar.reserve(bytes);
  if (rc) return rc;
  TBV_CUDA(cudaMemcpyAsync(ctx->filt.polar.p, polar, bytes, cudaMemcpyHostToDevice, ctx->stream));
  rc = filter_kstrongest_dev(ctx, ctx->filt.polar.p, n_az, n_range, row_stride, batch, params, out_peaks != nullptr);
  if (rc) return rc;
  return tbv_filter_fetch(ctx, out_filtered, out_peaks);
}

int tbv_compensate(tbv_ctx* ctx, float* x, float* y, int n, const double mot_xyt[3], int ccw) {
  TBV_ENTER(ctx);
  TBV_REQUIRE(ctx && x && y && mot_xyt && n >= 0, "null pointer");
  AllocScope alloc_scope(ctx->stream);  // temporaries of this call come from the stream-ordered pool
  if (n == 0) return TBV_OK;
  DevBuf<float> dx, dy;
  DevBuf<double> dm;
  int rc;
  if ((rc = dx.reserve(n)) || (rc = dy.reserve(n)) || (rc = dm.reserve(3))) { dx.release(); dy.release(); dm.release(); return rc; }
  auto cleanup = [&]() { dx.release(); dy.release(); dm.release(); };
  cudaError_t e = cudaMemcpyAsync(dx.p, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy.p, y, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dm.p, mot_xyt, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    dim3 grid((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024, 1);
    k_compensate<<<grid, 256, 0, ctx->stream>>>(dx.p, dy.p, nullptr, n, dm.p, ccw);
    launched(ctx, "k_compensate");
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(x, dx.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y, dy.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) { set_error("tbv_compensate: %s", cudaGetErrorString(e)); return TBV_ERR_CUDA; }
  return TBV_OK;
}

}  // extern "C"

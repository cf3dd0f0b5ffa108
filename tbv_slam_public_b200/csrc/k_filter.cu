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
#include <type_traits>

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"   // suspend-time hint (10 ms cap): the warp is parked, not spinning
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

// The same compensation for a point the filter itself just made: (x, y) = float(rho cos theta), float(rho sin theta) with (c, s) the
// azimuth's table entry and th = atan2(s, c) from the host (glibc).  atan2(y, x) = th + atan2(c y - s x, c x + s y) exactly (rotation by
// -th), and the second term is the float rounding of x and y seen as an angle: < 1e-7 rad, so atan(t) = t to 3e-22 rad.  The cross product
// is formed with its rounding error removed (two-product by fma), which leaves the sum th + delta with the one rounding a libm atan2 has
// as well — at a tenth of the instructions (no argument reduction, no polynomial, one division).
__device__ __forceinline__ void compensate_polar_point(float& x, float& y, double c, double s, double th, double m0, double m1, double m2, int ccw) {
  const double two_pi = __dmul_rn(2.0, 3.14159265358979323846);
  const double px = (double)x, py = (double)y;
  const double t = __dmul_rn(s, px);
  const double terr = __fma_rn(s, px, -t);                            // s * px = t + terr exactly
  const double cross = __dsub_rn(__fma_rn(c, py, -t), terr);          // c * py - s * px
  const double dot = __fma_rn(c, px, __dmul_rn(s, py));
  // cross / dot is a correction of < 1e-7 rad: the reciprocal needs 1e-9 relative accuracy, one Newton step from the fp32 seed gives 1e-14
  const double r0 = (double)__frcp_rn((float)dot);
  const double rcp = __dmul_rn(r0, __dsub_rn(2.0, __dmul_rn(dot, r0)));
  const double a = __dadd_rn(th, __dmul_rn(cross, rcp));
  double d = __ddiv_rn((a > 0.00001 ? a : __dadd_rn(two_pi, a)), two_pi);
  d = ccw ? -(__dsub_rn(d, 0.5)) : __dsub_rn(d, 0.5);
  const double ang = __dmul_rn(d, m2);
  double s1, c1;
  sincos(ang, &s1, &c1);
  const double tx = __dmul_rn(d, m0), ty = __dmul_rn(d, m1);
  x = (float)__dadd_rn(__dadd_rn(__dmul_rn(c1, px), __dmul_rn(-s1, py)), tx);
  y = (float)__dadd_rn(__dadd_rn(__dmul_rn(s1, px), __dmul_rn(c1, py)), ty);
}

// ---- the fused filter kernel ----------------------------------------------------------------------------------------------------------
// One CTA (6 warps) per scan walks the scan's rows in chunks of 12 — TWO rows per warp — and writes the two final clouds directly:
//   P1 (warp, its two rows)  each row arrives by ONE bulk copy (TMA, mbarrier) in one of the warp's two row buffers; branch-free
//                            conservative scan of the 16-byte vectors (one flag bit per lane and vector, no ballots in the loop); the
//                            flagged vector ids of BOTH rows go into one queue, so that the exact byte masks, the candidate extraction and
//                            everything after it run once per row pair with twice the lanes busy (a radar row holds ~14 candidates).
//                            Candidates carry their row in the bit above the intensity: entry = (row bit | intensity | range) << 2 | emitted
//                            << 1 | peak, so one all-pairs rank over the pair's list orders both rows at once (rank inside the row = rank in
//                            the pair minus, for the first row, the size of the second).  Rows with more candidates than a list holds take
//                            the exact dense path (threshold bisection on the staged row + ties from the far end) and contribute <= k.
//   P2 (warp)                rank -> the k strongest of each row, ascending (intensity, range); 7+7-tap axial non-max suppression of the
//                            emitted bins on the staged rows.  Then the warp requests its next two rows: the copies fly under P3 and the
//                            other warps.
//   P3 (warp)                the pair's emitted entries, one lane per point: fp64 polar -> Cartesian with the host (glibc) cos/sin table,
//                            motion compensation, stores to both clouds.  The clouds are in row order, so the pair's offset is the total of
//                            all earlier rows: the running totals travel warp to warp (and chunk to chunk) through shared memory — a warp
//                            publishes its totals as soon as it has counted, before computing a point, and its successor sleeps on an
//                            mbarrier until then.
// No CTA barrier inside the row loop: every phase is warp-local.  (Measured alternatives: two CTA barriers per chunk with a flattened
// emission — same speed, more shared memory; per-warp scratch segments compacted at the end of the scan — the compaction is a serial
// tail that all CTAs of the single wave reach together, 40 % slower.)  Nothing but the scan bytes is read from HBM and nothing but the
// clouds is written: the per-row key arrays of the two-kernel version (and their round trip through L2) are gone.
constexpr int KF_WARPS = 6;                  // warps per CTA; a chunk is 2 * KF_WARPS rows.  6: 48 KB of row buffers -> 4 CTAs per SM (592 scans = one wave on 148 SMs)
constexpr int KF_THREADS = KF_WARPS * 32;
constexpr int KF_MAX_K = 128;                // largest supported k (list capacity per row: 64 for k <= 64, else 128)

struct KfCloud {   // one cloud of the batch, struct of arrays, `cap` entries per scan
  float* x; float* y; uint8_t* i; uint16_t* az; uint16_t* rg;
};
__device__ __forceinline__ void kf_store(const KfCloud& c, size_t q, float x, float y, uint8_t inten, int az, int r) {
  c.x[q] = x; c.y[q] = y; c.i[q] = inten; c.az[q] = (uint16_t)az; c.rg[q] = (uint16_t)r;
}

__device__ __forceinline__ uint32_t kf_entry(uint32_t inten, int r, int min_range_bin, uint32_t rowbit) {
  return (((rowbit << 24) | (inten << 16) | (uint32_t)r) << 2) | (r > min_range_bin ? 2u : 0u);
}
// movemask of the four flag bytes (0x80 each) of a word: bits 7, 15, 23, 31 -> bits 0..3 (the multiply gathers them at bits 21..24; all
// partial products are distinct bits, so there are no carries)
__device__ __forceinline__ uint32_t kf_nibble(uint32_t m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xfu; }

struct KfRow {          // one staged row, as the warp sees it
  const uint8_t* buf;   // staged superset: buf[a0 + r] = row[r]
  int a0, nvec;         // alignment offset of the row start; 16-byte vectors covering [0, a0 + n_range)
};
__device__ __forceinline__ uint32_t kf_valid_mask(int wo, int lo_b, int hi_b) {   // word at buffer offset wo: 0x80 per byte inside the row
  uint32_t m = 0x80808080u;
  if (wo < lo_b) m &= (lo_b - wo >= 4) ? 0u : (0x80808080u << (8 * (lo_b - wo)));
  if (wo + 4 > hi_b) m &= (hi_b - wo <= 0) ? 0u : (0x80808080u >> (8 * (wo + 4 - hi_b)));
  return m;
}
// exact "byte >= t" flags of one staged vector as a 16-bit mask; the staged superset starts / ends up to 15 bytes outside the row
__device__ __forceinline__ uint32_t kf_ge16(const KfRow& R, int n_range, int t, uint32_t ac, bool h) {
  const uint4 v = reinterpret_cast<const uint4*>(R.buf)[t];
  const int wo = t << 4, lo_b = R.a0, hi_b = R.a0 + n_range;
  uint32_t m0 = ge_mask(v.x, ac, h), m1 = ge_mask(v.y, ac, h), m2 = ge_mask(v.z, ac, h), m3 = ge_mask(v.w, ac, h);
  if (wo < lo_b || wo + 16 > hi_b) { m0 &= kf_valid_mask(wo, lo_b, hi_b); m1 &= kf_valid_mask(wo + 4, lo_b, hi_b); m2 &= kf_valid_mask(wo + 8, lo_b, hi_b); m3 &= kf_valid_mask(wo + 12, lo_b, hi_b); }
  return kf_nibble(m0) | (kf_nibble(m1) << 4) | (kf_nibble(m2) << 8) | (kf_nibble(m3) << 12);
}

// Dense row: exact threshold T = k-th largest intensity (bisection over the staged row), everything above T, then the ties at T from the
// far end — the largest ranges win, exactly as the reference's erase(begin()) leaves them.  Appends min(k, #candidates) entries at
// list[*counter ...] (counter: the warp's shared cursor) and returns how many.  Whole warp.
__device__ __noinline__ int kf_dense_select(const KfRow R, int n_range, uint32_t z, int k, int min_range_bin, uint32_t rowbit, uint32_t* list,
                                            int list_cap, int* counter, int lane) {
  const unsigned FULL = 0xffffffffu;
  const int lo_b = R.a0, hi_b = R.a0 + n_range;
  const uint4* vbuf = reinterpret_cast<const uint4*>(R.buf);
  auto count_ge = [&](uint32_t t) -> int {   // bytes of the row >= t, t in 1..255
    const bool h = t > 128;
    const uint32_t ac = (h ? (256u - t) : (128u - t)) * 0x01010101u;
    int c = 0;
    for (int q = lane; q < R.nvec; q += 32) c += __popc(kf_ge16(R, n_range, q, ac, h));
    return __reduce_add_sync(FULL, c);
  };
  auto cnt = [&](uint32_t t) -> int { return t == 0 ? n_range : count_ge(t); };
  // T = the k-th largest intensity among the candidates (or z_min when the row holds no more than k candidates: all are taken)
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
  const int need = want - n_gt;    // ties to take at T, largest ranges first (>= 0)
  const int start = *counter;
  __syncwarp();
  if (T < 255 && n_gt > 0) {       // (1) everything strictly above the threshold
    const uint32_t t1 = T + 1;
    const bool h = t1 > 128;
    const uint32_t ac = (h ? (256u - t1) : (128u - t1)) * 0x01010101u;
    for (int q = lane; q < R.nvec; q += 32) {
      uint32_t m16 = kf_ge16(R, n_range, q, ac, h);
      if (m16) {
        const int wo = q << 4;
        int pos = atomicAdd(counter, __popc(m16));
        while (m16) {
          const int b = __ffs(m16) - 1;
          m16 &= m16 - 1;
          if (pos < list_cap) list[pos] = kf_entry(R.buf[wo + b], wo + b - R.a0, min_range_bin, rowbit);
          pos++;
        }
      }
    }
  }
  // (2) ties at T, scanning ranges from the far end; 32 vectors (512 bytes) per step
  const uint32_t T4 = T * 0x01010101u;
  int carry = 0;
  for (int base = ((R.nvec - 1) >> 5) << 5; base >= 0 && carry < need; base -= 32) {
    const int q = base + lane;
    uint32_t m16 = 0;
    if (q < R.nvec) {
      const uint4 v = vbuf[q];
      const int wo = q << 4;
      m16 = kf_nibble(eq_mask(v.x, T4) & kf_valid_mask(wo, lo_b, hi_b)) | (kf_nibble(eq_mask(v.y, T4) & kf_valid_mask(wo + 4, lo_b, hi_b)) << 4) |
            (kf_nibble(eq_mask(v.z, T4) & kf_valid_mask(wo + 8, lo_b, hi_b)) << 8) | (kf_nibble(eq_mask(v.w, T4) & kf_valid_mask(wo + 12, lo_b, hi_b)) << 12);
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
      int pos = atomicAdd(counter, take);
      const int wo = q << 4;
      while (take > 0) {  // highest bytes first
        const int b = 31 - __clz(m16);
        m16 &= ~(1u << b);
        if (pos < list_cap) list[pos] = kf_entry(T, wo + b - R.a0, min_range_bin, rowbit);
        pos++;
        take--;
      }
    }
    carry += __shfl_sync(FULL, suf, 0);
  }
  __syncwarp();
  return min(*counter, list_cap) - start;   // == want
}

// HI: z_min > 128 (selects the form of the conservative byte compare at compile time: the scan loop carries no branch on it);
// NG: number of 32-vector groups of a staged row, unrolled at compile time (8: Oxford, 7: MulRan); 0 = run-time loop over n_groups;
// CAP: candidate list capacity per row (>= k).
template <bool HI, int NG, int CAP>
__global__ void __launch_bounds__(KF_THREADS, 4)
k1_filter_fused(const uint8_t* __restrict__ polar, int n_az, int n_range, size_t row_stride, int z_min, int k, int want_peaks, int rowbuf, int n_groups,
                const uint8_t* buf_lo, const uint8_t* buf_hi, int min_range_bin, double range_res, const double2* __restrict__ cs_table,
                const double* __restrict__ th_table, int cap, KfCloud out_f, int* __restrict__ fcount, KfCloud out_p, int* __restrict__ pcount,
                const double* __restrict__ mot, int ccw, int split, KfCloud tmp_f, KfCloud tmp_p, int2* __restrict__ seg_tot, int* __restrict__ seg_done,
                const KfCloud dst_f, const KfCloud dst_p /* where the rows' points go: the final clouds (split == 1) or the scratch clouds */) {
  extern __shared__ __align__(128) uint8_t s_dyn[];                   // [KF_WARPS][2][rowbuf] staged rows, 
  __shared__ __align__(16) uint32_t s_list[KF_WARPS][2 * CAP + 4];     // candidates of the row pair (unordered, zero-padded to a multiple of 4)
  __shared__ __align__(16) uint32_t s_sel[KF_WARPS][2 * CAP];          // P1: queue of flagged vectors (u16); P2: selected entries, ascending
  __shared__ __align__(8) uint64_t s_bar[KF_WARPS][2];
  __shared__ int s_dn[KF_WARPS];
  __shared__ int s_chain[KF_WARPS][2];                                 // emitted points / peaks of the scan up to and including this warp's rows
  __shared__ __align__(8) uint64_t s_chain_bar[KF_WARPS];              // phase c completes when the warp has published its totals of chunk c
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  // split == 1: the CTA owns the whole scan.  split > 1 (small batches: fewer scans than the GPU has CTA slots): `split` CTAs share a scan,
  // each a contiguous block of row pairs [row_lo, row_hi); a block's points go to its own segment of the scratch clouds and the CTA that
  // finishes last moves the segments behind one another into the final clouds.
  const int scan = blockIdx.x / split, seg = blockIdx.x - scan * split;
  const int seg_rows = 2 * ((((n_az + 1) >> 1) + split - 1) / split);
  const int row_lo = min(n_az, seg * seg_rows), row_hi = min(n_az, row_lo + seg_rows);
  const uint32_t z = (uint32_t)z_min;
  const bool zero_thr = z == 0;   // every byte is a candidate: no sparse pass
  const uint32_t addc = (HI ? (256u - z) : (128u - z)) * 0x01010101u;
  const size_t scan_bytes = (size_t)(n_az - 1) * row_stride + (size_t)n_range;  // addressable bytes of one scan
  const uint8_t* scan_base = polar + (size_t)scan * (size_t)n_az * row_stride;
  uint8_t* bufs = s_dyn + (size_t)warp * 2 * rowbuf;   // the warp's two row buffers
  uint32_t* list = s_list[warp];
  uint32_t* sel = s_sel[warp];
  if (lane < 2) mbar_init(&s_bar[warp][lane], 1);
  if (lane == 2) mbar_init(&s_chain_bar[warp], 1);
  // The row buffers start out zero: the scan loop runs over whole groups of 32 vectors, and the vectors past a row's staged superset are
  // never written by a bulk copy (the one right behind it is re-zeroed per row, see below).
  for (int o = lane * 16; o < 2 * rowbuf; o += 32 * 16) *reinterpret_cast<uint4*>(bufs + o) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;      // bit x: parity to wait for on the warp's barrier x
  uint32_t in_flight = 0;   // bit x: a bulk copy into buffer x is under way
#pragma unroll
  for (int x = 0; x < 2; x++)
    if (row_lo + 2 * warp + x < row_hi &&
        stage_row_tma(bufs + x * rowbuf, scan_base + (size_t)(row_lo + 2 * warp + x) * row_stride, n_range, buf_lo, buf_hi, &s_bar[warp][x], lane))
      in_flight |= 1u << x;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  if (mot) { m0 = mot[scan * 3 + 0]; m1 = mot[scan * 3 + 1]; m2 = mot[scan * 3 + 2]; }
  const double range_res_half = range_res / 2.0;
  const size_t cbase = (size_t)scan * cap;
  const size_t dbase = split == 1 ? cbase : cbase + (size_t)row_lo * (size_t)k;   // a block of R rows emits at most R * k points
  __shared__ int s_last;
  __syncthreads();   // every warp's chain barrier is initialised before a neighbour waits on it

  // Conservative test "some byte of the 16 may be >= z_min" (never misses one; false positives only next to a byte >= 188):
  // z <= 128: byte + (128 - z) sets bit 7, or overflows the byte only when the byte itself has bit 7 set; z > 128: bit 7.
  auto any_ge = [&](const uint4 v) -> bool {
    if (HI) return ((v.x | v.y | v.z | v.w) & 0x80808080u) != 0;
    const uint32_t a = (v.x + addc) | v.x, b = (v.y + addc) | v.y, c = (v.z + addc) | v.z, d = (v.w + addc) | v.w;
    return ((a | b | c | d) & 0x80808080u) != 0;
  };

  const uint8_t* const row_base = scan_base + (size_t)(row_lo + 2 * warp) * row_stride;
  const size_t chunk_stride = (size_t)(2 * KF_WARPS) * row_stride;
  for (int row0 = row_lo, chunk = 0; row0 < row_hi; row0 += 2 * KF_WARPS, chunk++) {
    // =============================== P1: the warp's two rows ========================================================================
    const int rowA = row0 + 2 * warp;
    const uint8_t* const row_ptr = row_base + (size_t)chunk * chunk_stride;   // first row of the warp's pair
    KfRow R[2];
    uint32_t bits[2] = {0u, 0u};
    uint32_t have = 0;
#pragma unroll
    for (int x = 0; x < 2; x++) {
      R[x].buf = bufs + x * rowbuf; R[x].a0 = 0; R[x].nvec = 0;
      if (rowA + x >= row_hi) continue;
      have |= 1u << x;
      const uint8_t* rp = row_ptr + (size_t)x * row_stride;
      R[x].a0 = (int)(reinterpret_cast<uintptr_t>(rp) & 15u);
      R[x].nvec = (R[x].a0 + n_range + 15) >> 4;
      if (in_flight & (1u << x)) { mbar_wait(&s_bar[warp][x], (parity >> x) & 1u); parity ^= 1u << x; }
      else stage_row_sync(bufs + x * rowbuf, rp, n_range, lane);
      // a previous row with another alignment may have ended one vector later: that vector must not be seen by the guard-free scan
      if (lane == 0 && (R[x].nvec << 4) < rowbuf) *reinterpret_cast<uint4*>(bufs + x * rowbuf + (R[x].nvec << 4)) = make_uint4(0, 0, 0, 0);
      __syncwarp();
      if (!zero_thr) {   // scan: one flag bit per (lane, group); no ballots, no stores in the loop
        const uint4* vb = reinterpret_cast<const uint4*>(R[x].buf);
        uint32_t b = 0;
        if (NG > 0) {
#pragma unroll
          for (int j = 0; j < NG; j++)
            if (any_ge(vb[j * 32 + lane])) b |= 1u << j;
        } else {
#pragma unroll 4
          for (int j = 0; j < n_groups; j++) b |= (any_ge(vb[j * 32 + lane]) ? 1u : 0u) << j;
        }
        bits[x] = b;
      }
    }
    // flagged vectors per row (packed: first row in the low half), inclusive scan over the lanes
    const uint32_t mine = (uint32_t)__popc(bits[0]) | ((uint32_t)__popc(bits[1]) << 16);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t2 = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += t2;
    }
    const uint32_t totq = __shfl_sync(FULL, incl, 31);
    const int nq0 = (int)(totq & 0xffffu), nq1 = (int)(totq >> 16);
    uint32_t dense = zero_thr ? have : (((nq0 > CAP) ? 1u : 0u) | ((nq1 > CAP) ? 2u : 0u));
    uint16_t* queue = reinterpret_cast<uint16_t*>(sel);   // <= 2 * CAP flagged vectors: bit 15 = second row, bits 0..14 = vector id
    int n = 0, nrow[2] = {0, 0};
    for (int attempt = 0; attempt < 3; attempt++) {
      n = 0; nrow[0] = 0; nrow[1] = 0;
      if (dense) {
        if (lane == 0) s_dn[warp] = 0;
        __syncwarp();
#pragma unroll
        for (int x = 0; x < 2; x++)
          if (dense & have & (1u << x)) {
            nrow[x] = kf_dense_select(R[x], n_range, z, k, min_range_bin, (uint32_t)x, list, 2 * CAP, &s_dn[warp], lane);
            n += nrow[x];
          }
      }
      // queue of the sparse rows' flagged vectors, first row first
      const int q0n = (dense & 1u) ? 0 : nq0, q1n = (dense & 2u) ? 0 : nq1;
      {
        int q = (int)((incl & 0xffffu) - (mine & 0xffffu));
        uint32_t b = (dense & 1u) ? 0u : bits[0];
        while (b) { const int j = __ffs(b) - 1; b &= b - 1; queue[q++] = (uint16_t)(j * 32 + lane); }
        q = q0n + (int)((incl >> 16) - (mine >> 16));
        b = (dense & 2u) ? 0u : bits[1];
        while (b) { const int j = __ffs(b) - 1; b &= b - 1; queue[q++] = (uint16_t)(0x8000 | (j * 32 + lane)); }
      }
      __syncwarp();
      const int nq = q0n + q1n;
      uint32_t overflow = 0;
      // exact masks of the queued vectors, one vector per lane per round; list positions from a warp scan of the per-vector counts
      for (int qb = 0; qb < nq; qb += 32) {
        uint32_t m16 = 0, c = 0;
        int wo = 0, x = 0;
        if (qb + lane < nq) {
          const int t = queue[qb + lane];
          x = t >> 15;
          wo = (t & 0x7fff) << 4;
          const KfRow Rx = {x ? R[1].buf : R[0].buf, x ? R[1].a0 : R[0].a0, x ? R[1].nvec : R[0].nvec};   // selects, not indexing: R stays in registers
          m16 = kf_ge16(Rx, n_range, t & 0x7fff, addc, HI);
          c = (uint32_t)__popc(m16) << (16 * x);
        }
        uint32_t ic = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t2 = __shfl_up_sync(FULL, ic, d);
          if (lane >= d) ic += t2;
        }
        const uint32_t rt = __shfl_sync(FULL, ic, 31);
        const int t0 = (int)(rt & 0xffffu), t1 = (int)(rt >> 16);
        if (nrow[0] + t0 > CAP) overflow |= 1u;
        if (nrow[1] + t1 > CAP) overflow |= 2u;
        if (overflow) break;
        const uint32_t ex = ic - c;
        int pos = n + (int)(ex & 0xffffu) + (int)(ex >> 16);
        const uint8_t* rb = x ? R[1].buf : R[0].buf;
        const int ra0 = x ? R[1].a0 : R[0].a0;
        while (m16) {
          const int b = __ffs(m16) - 1;
          m16 &= m16 - 1;
          list[pos++] = kf_entry(rb[wo + b], wo + b - ra0, min_range_bin, (uint32_t)x);
        }
        n += t0 + t1; nrow[0] += t0; nrow[1] += t1;
      }
      if (!overflow) break;
      dense |= overflow;   // the overflowing row(s) go through the dense path; at most two more attempts
      __syncwarp();
    }
    __syncwarp();
    if (lane < 4) list[n + lane] = 0;   // pad to a multiple of 4 with zeros (never greater than an entry)
    __syncwarp();

    // =============================== P2: rank, non-max suppression, compaction (warp-local) =========================================
    const int nselA = nrow[0] < k ? nrow[0] : k, nselB = nrow[1] < k ? nrow[1] : k;
    const int n_sel = nselA + nselB;
    // all-pairs rank over the pair's list: entries of the second row are larger than all entries of the first, so
    // rank inside the row = rank in the pair (second row) or rank in the pair - |second row| (first row); selected iff < k
    for (int e = lane; e < n; e += 32) {
      const uint32_t me = list[e];
      int rank = 0;
      for (int q = 0; q < n; q += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(list + q);
        rank += (v.x > me) + (v.y > me) + (v.z > me) + (v.w > me);
      }
      const bool second = (me >> 26) & 1u;
      const int rk = second ? rank : rank - nrow[1];
      if (rk < k) sel[second ? nselA + (nselB - 1 - rk) : (nselA - 1 - rk)] = me;
    }
    __syncwarp();
    // axial non-max suppression of the emitted selected bins (radar_filters.cpp:238-298) on the staged rows
    if (want_peaks) {
      for (int e = lane; e < n_sel; e += 32) {
        const uint32_t ent = sel[e];
        if (!(ent & 2u)) continue;   // not emitted: its peak flag is never read
        const int x = (int)((ent >> 26) & 1u);
        const int rowp = rowA + x;
        const uint8_t* rb = x ? R[1].buf : R[0].buf;
        const int ra0 = x ? R[1].a0 : R[0].a0;
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
          // a score exists only where some selected in-band bin OF THE SAME ROW lies within 3 of the position
          const int q_lo = x ? nselA : 0, q_hi = x ? n_sel : nselA;
#pragma unroll
          for (int i = 0; i < 7; i++) {
            const int p = r - 3 + i;
            bool computed = false;
            for (int q = q_lo; q < q_hi; q++) {
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
      __syncwarp();
    }
    // request the warp's next two rows now: the copies fly under the emission below and under the other warps
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads / writes of the buffers before the bulk copies into them
    in_flight = 0;
#pragma unroll
    for (int x = 0; x < 2; x++) {
      const int nrow_next = rowA + 2 * KF_WARPS + x;
      if (nrow_next < row_hi && stage_row_tma(bufs + x * rowbuf, row_ptr + (size_t)(2 * KF_WARPS + x) * row_stride, n_range, buf_lo, buf_hi, &s_bar[warp][x], lane))
        in_flight |= 1u << x;
    }

    // =============================== P3: the pair's points, warp-local =================================================================
    // Output offsets: the clouds are written in row order, so the pair starts where all earlier rows of the scan end.  No CTA barrier:
    // the running totals travel along a chain warp 0 -> 1 -> ... -> KF_WARPS-1 -> warp 0 of the next chunk through shared memory; a warp
    // publishes its totals (mbarrier arrive, release) as soon as it has counted, before it computes a single point, and its successor
    // sleeps on that mbarrier (try_wait, acquire) instead of polling.
    int run_f = 0, run_p = 0;
    for (int b0 = 0; b0 < n_sel; b0 += 32) {
      const int e = b0 + lane;
      const uint32_t ent = e < n_sel ? sel[e] : 0u;
      run_f += __popc(__ballot_sync(FULL, (ent & 2u) != 0));
      run_p += __popc(__ballot_sync(FULL, (ent & 3u) == 3u));
    }
    int base_f = 0, base_p = 0;
    if (chunk > 0 || warp > 0) {
      const int pred = warp == 0 ? KF_WARPS - 1 : warp - 1;
      mbar_wait(&s_chain_bar[pred], (uint32_t)((warp == 0 ? chunk - 1 : chunk) & 1));   // the predecessor's totals for this position are in
      base_f = s_chain[pred][0]; base_p = s_chain[pred][1];
    }
    __syncwarp();
    if (lane == 0) {
      s_chain[warp][0] = base_f + run_f; s_chain[warp][1] = base_p + run_p;
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&s_chain_bar[warp])) : "memory");
    }
    {
      const unsigned lt = (1u << lane) - 1u;
      for (int b0 = 0; b0 < n_sel; b0 += 32) {
        const int e = b0 + lane;
        const uint32_t ent = e < n_sel ? sel[e] : 0u;
        const bool fe = (ent & 2u) != 0, fp = (ent & 3u) == 3u;
        const unsigned bf = __ballot_sync(FULL, fe), bp = __ballot_sync(FULL, fp);
        if (fe) {
          const int q = base_f + __popc(bf & lt);
          const int az = rowA + (int)((ent >> 26) & 1u);
          const int r = (int)((ent >> 2) & 0xffffu);
          const double2 cs = cs_table[az];
          const double rho = __dadd_rn(range_res_half, __dmul_rn(range_res, (double)r));  // radar_filters.cpp:329-330
          float x = (float)__dmul_rn(rho, cs.x);
          float y = (float)__dmul_rn(rho, cs.y);
          if (mot) compensate_polar_point(x, y, cs.x, cs.y, th_table[az], m0, m1, m2, ccw);
          const uint8_t inten = (uint8_t)((ent >> 18) & 0xffu);
          if (q < cap) kf_store(dst_f, dbase + q, x, y, inten, az, r);
          if (fp) {
            const int qp = base_p + __popc(bp & lt);
            if (qp < cap) kf_store(dst_p, dbase + qp, x, y, inten, az, r);
          }
        }
        base_f += __popc(bf);
        base_p += __popc(bp);
      }
    }
    __syncwarp();   // every lane is done with sel before the next chunk reuses it
  }
  __syncthreads();
  const int n_chunks = (row_hi - row_lo + 2 * KF_WARPS - 1) / (2 * KF_WARPS);
  const int my_f = n_chunks > 0 ? s_chain[KF_WARPS - 1][0] : 0, my_p = n_chunks > 0 ? s_chain[KF_WARPS - 1][1] : 0;
  if (split == 1) {
    if (tid == 0) {
      fcount[scan] = my_f < cap ? my_f : cap;
      if (want_peaks) pcount[scan] = my_p < cap ? my_p : cap;
    }
    return;
  }
  // ---- split scans: publish this block's totals; the CTA that arrives last compacts the scan ------------------------------------------------
  if (tid == 0) {
    seg_tot[scan * split + seg] = make_int2(my_f, my_p);
    __threadfence();                                   // totals and scratch points before the arrival
    s_last = atomicAdd(&seg_done[scan], 1) == split - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int off_f = 0, off_p = 0;
  for (int c = 0; c < split; c++) {
    const int2 t = __ldcg(&seg_tot[scan * split + c]);
    const size_t src = cbase + (size_t)min(n_az, c * seg_rows) * (size_t)k;
    for (int i = tid; i < t.x; i += KF_THREADS)
      if (off_f + i < cap) {
        const size_t q = cbase + off_f + i;
        out_f.x[q] = __ldcg(tmp_f.x + src + i); out_f.y[q] = __ldcg(tmp_f.y + src + i); out_f.i[q] = __ldcg(tmp_f.i + src + i);
        out_f.az[q] = __ldcg(tmp_f.az + src + i); out_f.rg[q] = __ldcg(tmp_f.rg + src + i);
      }
    for (int i = tid; i < t.y; i += KF_THREADS)
      if (off_p + i < cap) {
        const size_t q = cbase + off_p + i;
        out_p.x[q] = __ldcg(tmp_p.x + src + i); out_p.y[q] = __ldcg(tmp_p.y + src + i); out_p.i[q] = __ldcg(tmp_p.i + src + i);
        out_p.az[q] = __ldcg(tmp_p.az + src + i); out_p.rg[q] = __ldcg(tmp_p.rg + src + i);
      }
    off_f += t.x; off_p += t.y;
  }
  if (tid == 0) {
    fcount[scan] = off_f < cap ? off_f : cap;
    if (want_peaks) pcount[scan] = off_p < cap ? off_p : cap;
    seg_done[scan] = 0;                                // ready for the next launch
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
  if ((rc = F.th_table.reserve(n_az))) return rc;
  std::vector<double2> h(n_az);
  std::vector<double> th(n_az);
  for (int b = 0; b < n_az; b++) {
    const double theta = (double(b + 1) / n_az) * 2. * M_PI;  // radar_filters.cpp:317 — glibc cos/sin on the host keeps x,y bit-exact
    h[b].x = std::cos(theta);
    h[b].y = std::sin(theta);
    th[b] = std::atan2(h[b].y, h[b].x);                       // the azimuth as Compensate's atan2 sees it, in (-pi, pi] (utils.h:28-32)
  }
  TBV_CUDA(cudaMemcpyAsync(F.cs_table.p, h.data(), n_az * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
  TBV_CUDA(cudaMemcpyAsync(F.th_table.p, th.data(), n_az * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  F.cs_n_az = n_az;
  return TBV_OK;
}

int filter_kstrongest_dev(tbv_ctx* ctx, const uint8_t* polar_dev, int n_az, int n_range, size_t row_stride, int batch,
                          const tbv_filter_params* p, int want_peaks, const double* mot_dev, int ccw, cudaStream_t stream) {
  TBV_REQUIRE(ctx && polar_dev && p, "null pointer");
  cudaStream_t st = stream ? stream : ctx->stream;
  if (!stream) ctx->filt_epoch++;   // an ordinary user of the context's clouds (an overlapped odometry step orders itself against these)
  TBV_REQUIRE(n_az > 0 && n_range > 0 && batch > 0 && row_stride >= (size_t)n_range, "bad image shape");
  TBV_REQUIRE(n_range <= 8192, "n_range > 8192 is not supported");
  TBV_REQUIRE(n_az <= 4096, "n_az > 4096 is not supported");
  TBV_REQUIRE(p->k_strongest >= 1 && p->k_strongest <= KF_MAX_K, "k_strongest must be in [1,128]");
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
  const size_t k1_smem = (size_t)KF_WARPS * 2 * rowbuf;    // two row buffers per warp
  // Fewer scans than resident CTA slots (4 per SM): split every scan over several CTAs so that the whole GPU streams (the online node
  // hands over ONE scan at a time: 32 CTAs of ~12 rows).
  int split = 1;
  {
    const int slots = 4 * ctx->sm_count, n_pairs = (n_az + 1) / 2;
    while (split < 32 && batch * split * 2 <= slots && n_pairs / (split * 2) >= KF_WARPS) split *= 2;   // a block keeps >= one chunk of rows
  }
  if (split > 1) {
    if ((rc = F.tmp_f.reserve(batch, n_az * k))) return rc;
    if (want_peaks && (rc = F.tmp_p.reserve(batch, n_az * k))) return rc;
    if ((rc = F.seg_tot.reserve((size_t)batch * split))) return rc;
    if (F.seg_done.n < (size_t)batch) {
      if ((rc = F.seg_done.reserve(batch))) return rc;
      TBV_CUDA(cudaMemsetAsync(F.seg_done.p, 0, F.seg_done.n * sizeof(int), st));   // the kernel leaves the counters at zero
    }
  }
  const uint8_t* buf_hi = polar_dev + (size_t)(batch - 1) * n_az * row_stride + (size_t)(n_az - 1) * row_stride + (size_t)n_range;
  const double rr = (double)p->range_res;                                   // widened float (radar_filters.h:86)
  const int min_range_bin = (int)std::ceil((double)p->min_distance / rr);   // radar_filters.cpp:315
  auto launch = [&](auto kern) -> int {
    const int rc2 = ensure_dyn_smem(ctx, kern, k1_smem);
    if (rc2) return rc2;
    auto cl = [](DevCloud& c) { return KfCloud{c.x.p, c.y.p, c.inten.p, c.az.p, c.rg.p}; };
    kern<<<batch * split, KF_THREADS, k1_smem, st>>>(polar_dev, n_az, n_range, row_stride, z_min, k, want_peaks, rowbuf, n_groups, polar_dev, buf_hi,
                                                     min_range_bin, rr, F.cs_table.p, F.th_table.p, n_az * k, cl(F.filtered), F.filtered.count.p, cl(F.peaks),
                                                     F.peaks.count.p, mot_dev, ccw, split, cl(F.tmp_f), cl(F.tmp_p), F.seg_tot.p, F.seg_done.p,
                                                     split == 1 ? cl(F.filtered) : cl(F.tmp_f), split == 1 ? cl(F.peaks) : cl(F.tmp_p));
    return TBV_OK;
  };
  // compile-time variants: byte-compare form (z_min > 128), unrolled scan for the two dataset shapes, list capacity 64 (k <= 64) or 128
  auto pick = [&](auto hi_tag) -> int {
    constexpr bool H = decltype(hi_tag)::value;
    if (k <= 64) return n_groups == 8 ? launch(k1_filter_fused<H, 8, 64>) : n_groups == 7 ? launch(k1_filter_fused<H, 7, 64>) : launch(k1_filter_fused<H, 0, 64>);
    return n_groups == 8 ? launch(k1_filter_fused<H, 8, 128>) : n_groups == 7 ? launch(k1_filter_fused<H, 7, 128>) : launch(k1_filter_fused<H, 0, 128>);
  };
  rc = z_min > 128 ? pick(std::true_type{}) : pick(std::false_type{});
  if (rc) return rc;
  launched(ctx, "k1_filter_fused");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

// The fused filter's compensation as a separate launch: one thread per point of both clouds; (cos, sin, theta) of the point's azimuth from
// the host tables, i.e. the very operations of compensate_polar_point on the very floats the filter stored.
__global__ void __launch_bounds__(256)
k_compensate_polar(KfCloud cf, const int* __restrict__ fcount, KfCloud cp, const int* __restrict__ pcount, int cap, const double2* __restrict__ cs_table,
                   const double* __restrict__ th_table, const double* __restrict__ mot, int ccw) {
  const int scan = blockIdx.y;
  const KfCloud c = blockIdx.z ? cp : cf;
  const int n = min((blockIdx.z ? pcount : fcount)[scan], cap);
  const double m0 = mot[scan * 3 + 0], m1 = mot[scan * 3 + 1], m2 = mot[scan * 3 + 2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = (size_t)scan * cap + i;
    float x = c.x[q], y = c.y[q];
    const int az = c.az[q];
    const double2 cs = cs_table[az];
    compensate_polar_point(x, y, cs.x, cs.y, th_table[az], m0, m1, m2, ccw);
    c.x[q] = x; c.y[q] = y;
  }
}

int compensate_polar_clouds_dev(tbv_ctx* ctx, const double* mot_dev, int ccw, int want_peaks) {
  FilterState& F = ctx->filt;
  TBV_REQUIRE(F.batch > 0 && mot_dev, "no filter result on the device");
  auto cl = [](DevCloud& c) { return KfCloud{c.x.p, c.y.p, c.inten.p, c.az.p, c.rg.p}; };
  dim3 grid(8, F.batch, want_peaks ? 2 : 1);
  k_compensate_polar<<<grid, 256, 0, ctx->stream>>>(cl(F.filtered), F.filtered.count.p, cl(F.peaks), F.peaks.count.p, F.filtered.cap, F.cs_table.p, F.th_table.p,
                                                    mot_dev, ccw);
  launched(ctx, "k_compensate_polar");
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

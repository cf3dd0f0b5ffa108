// k_filter.cu — K1 (k-strongest + axial non-max suppression), K2 (polar -> Cartesian cloud), motion compensation.
//
// Replaces StructuredKStrongest::{FilterKstrongest, AxialNonMaxSupress, getPeaksFilteredPointCloud}
// (cfear_radarodometry/src/cfear_radarodometry/radar_filters.cpp:198-337) and CFEAR_Radarodometry::Compensate
// (cfear_radarodometry/src/cfear_radarodometry/utils.cpp:96-113).
//
// K1 design (sm_100a, HBM-bound byte scan):
//   * one warp owns one azimuth row at a time; rows are streamed HBM -> shared memory with cp.async (16-byte
//     transfers, 8-byte head/tail when the row start is only 8-byte aligned, as every odd Oxford row is) into a
//     per-warp double buffer, so the next row is in flight while the current one is processed and no registers are
//     tied up by staging (3 CTAs x 8 warps x 3.7 KB in flight per SM);
//   * the reference keeps, per row, the k largest (intensity, range) pairs under std::pair ordering — i.e. the k
//     largest 24-bit keys (intensity << 16 | range) — in ascending order.  Bytes >= z_min are found with a 3-instruction
//     SWAR compare per 4 bytes on conflict-free 16-byte shared loads; the (usually few) candidates are scattered to a
//     per-warp list and ranked all-pairs (rank = number of larger keys), which yields both the selection (rank < k)
//     and the output position;
//   * rows with more candidates than the list holds (dense / adversarial input) take an exact two-level path:
//     8-step binary search of the intensity threshold over the staged row, then a suffix scan over the ties so that
//     the largest ranges win, exactly as the reference's erase(begin()) does;
//   * the 7+7-tap axial non-max suppression runs in the same kernel on the selected bins straight from the staged row
//     (bytes across the row edge come from global memory, as the reference's flat cv::Mat indexing reads them) and is
//     stored as bit 31 of the key.
// Output per row: up to k keys ascending + count.  K2 turns rows into the two ordered clouds.
#include <cmath>

#include "tbv_common.cuh"

namespace tbv {

constexpr int K1_WARPS = 8;
constexpr int K1_CAP = 128;  // per-warp candidate list capacity (also the largest supported k)

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

template <bool HI>   // HI: z_min > 128 (selects the form of the byte compare at compile time: the scan loop carries no branch on it)
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_kstrongest(const uint8_t* __restrict__ polar, int total_rows, int n_az, int n_range, size_t row_stride, int z_min, int k,
              int want_peaks, int rowbuf, const uint8_t* buf_lo, const uint8_t* buf_hi, int min_range_bin, uint32_t* __restrict__ row_keys,
              uint32_t* __restrict__ row_cnt) {
  extern __shared__ __align__(128) uint8_t s_dyn[];  // [K1_WARPS][2][rowbuf] staged rows
  __shared__ __align__(16) uint32_t s_list[K1_WARPS][K1_CAP];  // candidates (unordered)
  __shared__ __align__(16) uint32_t s_sel[K1_WARPS][K1_CAP];   // selected, ascending
  __shared__ __align__(8) uint64_t s_bar[K1_WARPS][2];
  __shared__ int s_n[K1_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  constexpr bool hi = HI;
  const uint32_t z = (uint32_t)z_min;
  const bool zero_thr = z == 0;   // every byte is a candidate: padding bytes must be masked explicitly
  const uint32_t addc = (hi ? (256u - z) : (128u - z)) * 0x01010101u;
  const size_t scan_bytes = (size_t)(n_az - 1) * row_stride + (size_t)n_range;  // addressable bytes of one scan
  const size_t scan_stride = (size_t)n_az * row_stride;
  uint8_t* mybuf = s_dyn + (size_t)warp * 2 * rowbuf;
  uint32_t* list = s_list[warp];
  uint32_t* sel = s_sel[warp];
  uint64_t* bars = s_bar[warp];
  const int row_step = gridDim.x * K1_WARPS;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  // Both row buffers start out zero: the scan loop below always runs over whole groups of 32 vectors (rowbuf is a multiple of 512
  // bytes), and the bytes past a row's staged superset are never written by a bulk copy, so they stay below every threshold >= 1.
  for (int o = lane * 16; o < 2 * rowbuf; o += 32 * 16) *reinterpret_cast<uint4*>(mybuf + o) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncwarp();

  auto row_ptr = [&](int scan, int az) -> const uint8_t* { return polar + (size_t)scan * scan_stride + (size_t)az * row_stride; };
  int row = blockIdx.x * K1_WARPS + warp;
  // (scan, az) of the warp's current row, advanced without divisions: row += row_step
  int scan = row / n_az, az = row - scan * n_az;
  const int step_scan = row_step / n_az, step_az = row_step - step_scan * n_az;
  int cur = 0;
  uint32_t phase = 0;    // bit b: parity to wait for on bars[b]
  bool in_flight = false;
  if (row < total_rows) in_flight = stage_row_tma(mybuf, row_ptr(scan, az), n_range, buf_lo, buf_hi, &bars[0], lane);
  for (; row < total_rows; row += row_step) {
    const uint8_t* scan_base = polar + (size_t)scan * scan_stride;
    const uint8_t* rp = scan_base + (size_t)az * row_stride;
    const int a0 = (int)(reinterpret_cast<uintptr_t>(rp) & 15u);
    uint8_t* buf = mybuf + (size_t)cur * rowbuf;
    // prefetch the next row of this warp into the other buffer (its previous contents were consumed last iteration)
    const int nrow = row + row_step;
    int scan_n = scan + step_scan, az_n = az + step_az;
    if (az_n >= n_az) { az_n -= n_az; scan_n++; }
    bool next_in_flight = false;
    if (nrow < total_rows) next_in_flight = stage_row_tma(mybuf + (size_t)(cur ^ 1) * rowbuf, row_ptr(scan_n, az_n), n_range, buf_lo, buf_hi, &bars[cur ^ 1], lane);
    if (in_flight) {
      mbar_wait(&bars[cur], (phase >> cur) & 1u);
      phase ^= 1u << cur;
    } else {
      stage_row_sync(buf, rp, n_range, lane);
    }
    // The staged superset starts up to 15 bytes before the row and ends up to 15 bytes after it: clear those bytes so that
    // no threshold >= 1 ever matches them and the scan needs no per-vector validity test (z_min == 0 keeps the masks).
    if (!zero_thr) {
      const int head = a0, tail = (((a0 + n_range + 15) >> 4) << 4) - (a0 + n_range);
      if (lane < head) buf[lane] = 0;
      // the superset ends up to 15 bytes after the row; a previous row of this buffer with another alignment may have ended one
      // vector later: clear through the end of that vector as well
      if (lane < tail + 16 && a0 + n_range + lane < rowbuf) buf[a0 + n_range + lane] = 0;
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes before the next bulk copy into this buffer
      __syncwarp();
    }

    // word validity: byte offset wo in buf holds row index wo - a0; valid iff 0 <= wo + b - a0 < n_range
    const int lo_b = a0, hi_b = a0 + n_range;  // valid buffer byte range [lo_b, hi_b)
    auto valid_mask = [&](int wo) -> uint32_t {
      uint32_t m = 0x80808080u;
      if (wo < lo_b) m &= (lo_b - wo >= 4) ? 0u : (0x80808080u << (8 * (lo_b - wo)));
      if (wo + 4 > hi_b) m &= (hi_b - wo <= 0) ? 0u : (0x80808080u >> (8 * (wo + 4 - hi_b)));
      return m;
    };
    const int nvec = (hi_b + 15) >> 4;  // 16-byte vectors covering [0, hi_b)
    const uint4* vbuf = reinterpret_cast<const uint4*>(buf);
    auto masks = [&](const uint4 v, int wo, uint32_t ac, bool h, uint32_t m[4]) {
      m[0] = ge_mask(v.x, ac, h); m[1] = ge_mask(v.y, ac, h); m[2] = ge_mask(v.z, ac, h); m[3] = ge_mask(v.w, ac, h);
      if (zero_thr && (wo < lo_b || wo + 16 > hi_b)) { m[0] &= valid_mask(wo); m[1] &= valid_mask(wo + 4); m[2] &= valid_mask(wo + 8); m[3] &= valid_mask(wo + 12); }
    };
    // Conservative test "some byte of the 16 may be >= z_min" (never misses one; false positives only next to a byte >= 188):
    // z <= 128: byte + (128 - z) sets bit 7, or overflows the byte only when the byte itself has bit 7 set; z > 128: bit 7.
    auto any_ge = [&](const uint4 v) -> bool {
      if (hi) return ((v.x | v.y | v.z | v.w) & 0x80808080u) != 0;
      const uint32_t a = (v.x + addc) | v.x, b = (v.y + addc) | v.y, c = (v.z + addc) | v.z, d = (v.w + addc) | v.w;
      return ((a | b | c | d) & 0x80808080u) != 0;
    };
    auto emit = [&](uint32_t m, uint32_t w, int wo) {  // dense path only: scatter the flagged bytes of one word
      int pos = atomicAdd(&s_n[warp], __popc(m));
      while (m) {
        const int b = (__ffs(m) - 1) >> 3;
        m &= m - 1;
        if (pos < K1_CAP) list[pos] = (((w >> (8 * b)) & 0xffu) << 16) | (uint32_t)(wo + b - a0);
        pos++;
      }
    };
    // ---- pass 1: branch-free scan.  The conservative test flags ~1 vector in 9 on radar data; the flagged vector ids are
    // compacted (ballot order) into a per-warp queue so that pass 2 works on them with all 32 lanes busy instead of every
    // lane dragging the whole warp through its own rare hits.  The queue lives in `sel` (free until the ranking step).
    uint16_t* queue = reinterpret_cast<uint16_t*>(sel);  // K1_CAP + 1 entries are enough: more flagged vectors than that => dense row
    int nq = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
    // whole groups of 32 vectors: the padding is zero, so the loop body has no guard.  z_min = 0 makes every byte a candidate (more
    // than the list holds in any real row): that case goes straight to the dense path, which masks the padding explicitly.
    const int nvec32 = zero_thr ? 0 : ((nvec + 31) & ~31);
    if (zero_thr) nq = K1_CAP + 1;
    for (int base = 0; base < nvec32; base += 32) {
      const int t = base + lane;
      const bool f = any_ge(vbuf[t]);
      const unsigned ball = __ballot_sync(FULL, f);
      const int q = nq + __popc(ball & lt);
      if (f && q <= K1_CAP) queue[q] = (uint16_t)t;
      nq += __popc(ball);
    }
    __syncwarp();
    // ---- pass 2 (sparse row, the normal case): exact masks of the queued vectors, one vector per lane per round; positions from
    // a warp scan of the per-vector counts.  The list order is arbitrary (the ranking step orders it) but deterministic.
    int n = 0;
    bool dense = nq > K1_CAP;
    for (int qb = 0; qb < nq && !dense; qb += 32) {
      int c = 0, wo = 0;
      uint4 v = make_uint4(0, 0, 0, 0);
      uint32_t m[4] = {0, 0, 0, 0};
      if (qb + lane < nq) {
        const int t = queue[qb + lane];
        wo = t << 4;
        v = vbuf[t];
        masks(v, wo, addc, hi, m);
        c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
      }
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t2 = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t2;
      }
      const int round_total = __shfl_sync(FULL, incl, 31);
      if (n + round_total > K1_CAP) { dense = true; break; }
      int pos = n + incl - c;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t mm = m[q];
        while (mm) {
          const int b = (__ffs(mm) - 1) >> 3;
          mm &= mm - 1;
          list[pos++] = (((w[q] >> (8 * b)) & 0xffu) << 16) | (uint32_t)(wo + 4 * q + b - a0);
        }
      }
      n += round_total;
    }
    if (!dense) {
      __syncwarp();
    } else {
      // ---- dense row: exact threshold + tie handling ----------------------------------------------------
      if (lane == 0) s_n[warp] = 0;
      __syncwarp();
      auto count_ge = [&](uint32_t t) -> int {
        const bool h = t > 128;
        const uint32_t ac = (h ? (256u - t) : (128u - t)) * 0x01010101u;
        int c = 0;
        for (int q = lane; q < nvec; q += 32) {
          uint32_t m[4];
          masks(vbuf[q], q << 4, ac, h, m);
          c += __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
        }
        return __reduce_add_sync(FULL, c);
      };
      uint32_t lo = z, hi_t = 256;  // count(>= lo) >= k, count(>= hi_t) < k
      while (hi_t - lo > 1) {
        const uint32_t mid = (lo + hi_t) >> 1;
        if (count_ge(mid) >= k) lo = mid; else hi_t = mid;
      }
      const uint32_t T = lo;
      const int n_gt = (T >= 255) ? 0 : count_ge(T + 1);
      const int need = k - n_gt;  // >= 1 ties to take, largest ranges first
      if (T < 255) {  // (1) everything strictly above the threshold
        const uint32_t t1 = T + 1;
        const bool h = t1 > 128;
        const uint32_t ac = (h ? (256u - t1) : (128u - t1)) * 0x01010101u;
        for (int q = lane; q < nvec; q += 32) {
          const uint4 v = vbuf[q];
          const int wo = q << 4;
          uint32_t m[4];
          masks(v, wo, ac, h, m);
          if (m[0]) emit(m[0], v.x, wo);
          if (m[1]) emit(m[1], v.y, wo + 4);
          if (m[2]) emit(m[2], v.z, wo + 8);
          if (m[3]) emit(m[3], v.w, wo + 12);
        }
      }
      // (2) ties at T, scanning ranges from the far end; 32 vectors (512 bytes) per step
      const uint32_t T4 = T * 0x01010101u;
      int carry = 0;
      for (int base = ((nvec - 1) >> 5) << 5; base >= 0 && carry < need; base -= 32) {
        const int q = base + lane;
        uint32_t m[4] = {0, 0, 0, 0};
        if (q < nvec) {
          const uint4 v = vbuf[q];
          const int wo = q << 4;
          m[0] = eq_mask(v.x, T4) & valid_mask(wo); m[1] = eq_mask(v.y, T4) & valid_mask(wo + 4);
          m[2] = eq_mask(v.z, T4) & valid_mask(wo + 8); m[3] = eq_mask(v.w, T4) & valid_mask(wo + 12);
        }
        const int c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
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
          int pos = atomicAdd(&s_n[warp], take);
          const int wo = q << 4;
          for (int b = 15; b >= 0 && take > 0; b--) {  // highest bytes first
            if (m[b >> 2] & (0x80u << (8 * (b & 3)))) {
              if (pos < K1_CAP) list[pos] = (T << 16) | (uint32_t)(wo + b - a0);
              pos++;
              take--;
            }
          }
        }
        carry += __shfl_sync(FULL, suf, 0);
      }
      __syncwarp();
      n = min(s_n[warp], K1_CAP);  // == k
    }
    const int n_sel = n < k ? n : k;
    // pad to a multiple of 4 with zeros (never greater than a key)
    if (lane < 4 && n + lane < K1_CAP) list[n + lane] = 0;
    __syncwarp();
    // ---- all-pairs rank: rank = number of strictly larger keys; selected iff rank < k -----------------------
    for (int e = lane; e < n; e += 32) {
      const uint32_t key = list[e];
      int rank = 0;
      for (int q = 0; q < n; q += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(list + q);
        rank += (v.x > key) + (v.y > key) + (v.z > key) + (v.w > key);
      }
      if (rank < k) sel[n_sel - 1 - rank] = key;
    }
    __syncwarp();
    // ---- axial non-max suppression on the selected bins ---------------------------------------------------
    if (want_peaks) {
      uint32_t peak_bits = 0;  // bit t: entry lane + 32*t is a peak
      for (int e = lane; e < n_sel; e += 32) {
        const uint32_t key = sel[e];
        const int r = (int)(key & 0xffffu);
        const bool in_band = (r >= 3) && (r < n_range - 3);
        int B[13];
        if (r >= 6 && r + 6 < n_range) {
#pragma unroll
          for (int t = 0; t < 13; t++) B[t] = buf[a0 + r - 6 + t];
        } else {
          const long long gbase = (long long)az * (long long)row_stride;
#pragma unroll
          for (int t = 0; t < 13; t++) {
            const int q = r - 6 + t;
            if (q >= 0 && q < n_range) B[t] = buf[a0 + q];
            else {
              const long long fi = gbase + q;  // flat index into the scan buffer, as cv::Mat::at(bearing, r_nn) addresses it
              B[t] = (fi >= 0 && fi < (long long)scan_bytes) ? (int)__ldg(scan_base + fi) : 0;
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
              const int r2 = (int)(sel[q] & 0xffffu);
              if (r2 >= 3 && r2 < n_range - 3 && r2 - p <= 3 && p - r2 <= 3) { computed = true; break; }
            }
            if (!computed) s[i] = 0;
          }
        }
        bool largest = true;
#pragma unroll
        for (int i = 1; i <= 3; i++)
          if (s[3 - i] > s[3] || s[3] < s[3 + i]) largest = false;
        if (largest) peak_bits |= 1u << (e >> 5);
      }
      __syncwarp();
      for (int e = lane; e < n_sel; e += 32)
        if (peak_bits & (1u << (e >> 5))) sel[e] |= 0x80000000u;
      __syncwarp();
    }
    // ---- write the row ------------------------------------------------------------------------------------------
    uint32_t* out = row_keys + (size_t)row * k;
    int nf = 0, np = 0;  // entries K2 will emit: range beyond min_range_bin (radar_filters.cpp:321), and the peaks among them
    for (int e = lane; e < ((n_sel + 31) & ~31); e += 32) {
      const uint32_t key = e < n_sel ? sel[e] : 0u;
      if (e < n_sel) out[e] = key;
      const bool ok = e < n_sel && (int)(key & 0xffffu) > min_range_bin;
      nf += __popc(__ballot_sync(FULL, ok));
      np += __popc(__ballot_sync(FULL, ok && (key >> 31)));
    }
    if (lane == 0) row_cnt[row] = (uint32_t)n_sel | ((uint32_t)nf << 8) | ((uint32_t)np << 16);
    __syncwarp();  // every lane is done with buf / list / sel before the next iteration reuses them
    cur ^= 1;
    in_flight = next_in_flight;
    scan = scan_n; az = az_n;
  }
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

// K2: rows -> ordered clouds.  K2_SPLIT CTAs per scan: each scans the scan's per-row counts (written by K1: selected, emitted,
// emitted peaks — one packed word per row) into output offsets and emits its own slice of rows.  mot != nullptr: the points
// are motion-compensated as they are emitted (odometrykeyframefuser.cpp:146-150 applies Compensate to both clouds right after
// the filter; a peak is the same point in both clouds, so it is compensated once).
constexpr int K2_SPLIT = 4;
constexpr int K2_STAGE = 4096;  // staged points per chunk of rows (k <= 128 <= K2_STAGE)
__global__ void __launch_bounds__(256)
k2_make_clouds(const uint32_t* __restrict__ row_keys, const uint32_t* __restrict__ row_cnt, int n_az, int k, int min_range_bin,
               double range_res, const double2* __restrict__ cs_table, int cap,
               float* __restrict__ fx, float* __restrict__ fy, uint8_t* __restrict__ fi, uint16_t* __restrict__ faz, uint16_t* __restrict__ frg,
               int* __restrict__ fcount, int want_peaks,
               float* __restrict__ px, float* __restrict__ py, uint8_t* __restrict__ pi, uint16_t* __restrict__ paz, uint16_t* __restrict__ prg,
               int* __restrict__ pcount, const double* __restrict__ mot, int ccw) {
  extern __shared__ int s_off[];  // [2][n_az + 1]
  int* off_f = s_off;
  int* off_p = s_off + (n_az + 1);
  const int scan = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const uint32_t* keys = row_keys + (size_t)scan * n_az * k;
  const uint32_t* cnts = row_cnt + (size_t)scan * n_az;
  // phase 1: exclusive scans of the per-row emitted counts (warp 0: filtered, warp 1: peaks)
  if (warp < 2) {
    int* o = warp == 0 ? off_f : off_p;
    const int shift = warp == 0 ? 8 : 16;
    int running = 0;
    for (int base = 0; base < n_az; base += 32) {
      const int v = (base + lane < n_az) ? (int)((cnts[base + lane] >> shift) & 0xffu) : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
      }
      if (base + lane < n_az) o[base + lane] = running + inc - v;
      running += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0 && blockIdx.x == 0) {
      if (warp == 0) fcount[scan] = running; else if (want_peaks) pcount[scan] = running;
    }
  }
  __syncthreads();
  // phase 2: the points of this CTA's rows.  A row keeps only ~12 of its k entries on radar data, so computing the points warp-per-row
  // would leave most lanes idle through the fp64 atan2 / sincos of the compensation.  Instead the rows are first compacted (cheap,
  // warp per row) into a staging list in shared memory, in output order; then one thread per staged point does the arithmetic and
  // the stores are coalesced.  Rows are taken in chunks whose k * rows fit the staging list.
  uint32_t* st_key = reinterpret_cast<uint32_t*>(s_off + 2 * (n_az + 1));   // [K2_STAGE] packed key
  int* st_pq = reinterpret_cast<int*>(st_key + K2_STAGE);                   // [K2_STAGE] position in the peaks cloud, or -1
  uint16_t* st_row = reinterpret_cast<uint16_t*>(st_pq + K2_STAGE);         // [K2_STAGE] azimuth
  const double range_res_half = range_res / 2.0;
  const size_t cbase = (size_t)scan * cap;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  if (mot) { m0 = mot[scan * 3 + 0]; m1 = mot[scan * 3 + 1]; m2 = mot[scan * 3 + 2]; }
  const int rows_per = (n_az + gridDim.x - 1) / gridDim.x;
  const int row_begin = min(n_az, (int)blockIdx.x * rows_per), row_end = min(n_az, row_begin + rows_per);
  const int chunk_rows = max(1, K2_STAGE / max(k, 1));
  for (int r0 = row_begin; r0 < row_end; r0 += chunk_rows) {
    const int r1 = min(row_end, r0 + chunk_rows);
    const int base_f = off_f[r0];
    for (int row = r0 + warp; row < r1; row += nwarps) {
      const int c = (int)(cnts[row] & 0xffu);
      int of = off_f[row] - base_f, op = off_p[row];
      for (int e = lane; e < ((c + 31) & ~31); e += 32) {
        const uint32_t key = e < c ? keys[(size_t)row * k + e] : 0u;
        const bool ok = e < c && (int)(key & 0xffffu) > min_range_bin;
        const bool pk = ok && (key >> 31);
        const unsigned bf = __ballot_sync(FULL, ok), bp = __ballot_sync(FULL, pk);
        const unsigned lt = (1u << lane) - 1u;
        if (ok) {
          const int q = of + __popc(bf & lt);
          st_key[q] = key;
          st_row[q] = (uint16_t)row;
          st_pq[q] = (pk && want_peaks) ? op + __popc(bp & lt) : -1;
        }
        of += __popc(bf);
        op += __popc(bp);
      }
    }
    __syncthreads();
    const int n_local = off_f[r1 - 1] + (int)((cnts[r1 - 1] >> 8) & 0xffu) - base_f;
    for (int i = threadIdx.x; i < n_local; i += blockDim.x) {
      const uint32_t key = st_key[i];
      const int row = st_row[i];
      const int r = (int)(key & 0xffffu);
      const double2 cs = cs_table[row];
      const double rho = __dadd_rn(range_res_half, __dmul_rn(range_res, (double)r));  // radar_filters.cpp:329-330
      float x = (float)__dmul_rn(rho, cs.x);
      float y = (float)__dmul_rn(rho, cs.y);
      if (mot) compensate_point(x, y, m0, m1, m2, ccw);
      const uint8_t inten = (uint8_t)((key >> 16) & 0xffu);
      const int q = base_f + i;
      if (q < cap) {
        fx[cbase + q] = x; fy[cbase + q] = y; fi[cbase + q] = inten; faz[cbase + q] = (uint16_t)row; frg[cbase + q] = (uint16_t)r;
      }
      const int qp = st_pq[i];
      if (qp >= 0 && qp < cap) {
        px[cbase + qp] = x; py[cbase + qp] = y; pi[cbase + qp] = inten; paz[cbase + qp] = (uint16_t)row; prg[cbase + qp] = (uint16_t)r;
      }
    }
    __syncthreads();
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
  TBV_REQUIRE(p->k_strongest >= 1 && p->k_strongest <= K1_CAP, "k_strongest must be in [1,128]");
  const int z_min = (int)p->z_min;  // float -> int as StructuredKStrongest's ctor does (radar_filters.h:86)
  TBV_REQUIRE(z_min >= 0 && z_min <= 255, "z_min must be in [0,255]");
  FilterState& F = ctx->filt;
  const int k = p->k_strongest;
  int rc;
  if ((rc = F.row_keys.reserve((size_t)batch * n_az * k))) return rc;
  if ((rc = F.row_cnt.reserve((size_t)batch * n_az))) return rc;
  if ((rc = F.filtered.reserve(batch, n_az * k))) return rc;
  if (want_peaks && (rc = F.peaks.reserve(batch, n_az * k))) return rc;
  if ((rc = ensure_cs_table(ctx, n_az))) return rc;
  F.batch = batch; F.n_az = n_az; F.n_range = n_range; F.k = k;

  const int total_rows = batch * n_az;
  const int dev_sms = ctx->sm_count;
  const int blocks_needed = (total_rows + K1_WARPS - 1) / K1_WARPS;
  const int rowbuf = ((n_range + 16 + 15 + 511) / 512) * 512;  // row + alignment slack, whole groups of 32 16-byte vectors
  const size_t k1_smem = (size_t)K1_WARPS * 2 * rowbuf;
  // resident CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per CTA, ~8.3 KB static (lists, barriers)
  int ctas_per_sm = (int)((228 * 1024) / (k1_smem + 8500 + 1024));
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int max_grid = dev_sms * ctas_per_sm;
  const int grid = blocks_needed < max_grid ? blocks_needed : max_grid;  // persistent: resident CTAs only, rows strided over warps
  if ((rc = ensure_dyn_smem(ctx, k1_kstrongest<false>, k1_smem)) || (rc = ensure_dyn_smem(ctx, k1_kstrongest<true>, k1_smem))) return rc;
  const uint8_t* buf_hi = polar_dev + (size_t)(batch - 1) * n_az * row_stride + (size_t)(n_az - 1) * row_stride + (size_t)n_range;
  const double rr = (double)p->range_res;                                   // widened float (radar_filters.h:86)
  const int min_range_bin = (int)std::ceil((double)p->min_distance / rr);   // radar_filters.cpp:315
  if (z_min > 128)
    k1_kstrongest<true><<<grid, K1_WARPS * 32, k1_smem, ctx->stream>>>(polar_dev, total_rows, n_az, n_range, row_stride, z_min, k, want_peaks, rowbuf,
                                                                     polar_dev, buf_hi, min_range_bin, F.row_keys.p, F.row_cnt.p);
  else
    k1_kstrongest<false><<<grid, K1_WARPS * 32, k1_smem, ctx->stream>>>(polar_dev, total_rows, n_az, n_range, row_stride, z_min, k, want_peaks, rowbuf,
                                                                      polar_dev, buf_hi, min_range_bin, F.row_keys.p, F.row_cnt.p);
  launched(ctx, "k1_kstrongest");
  TBV_CUDA(cudaGetLastError());
  const size_t smem = 2 * (size_t)(n_az + 1) * sizeof(int) + (size_t)K2_STAGE * (sizeof(uint32_t) + sizeof(int) + sizeof(uint16_t));
  if ((rc = ensure_dyn_smem(ctx, k2_make_clouds, smem))) return rc;
  k2_make_clouds<<<dim3(K2_SPLIT, batch), 256, smem, ctx->stream>>>(F.row_keys.p, F.row_cnt.p, n_az, k, min_range_bin, rr, F.cs_table.p, n_az * k,
                                                                    F.filtered.x.p, F.filtered.y.p, F.filtered.inten.p, F.filtered.az.p,
                                                                    F.filtered.rg.p, F.filtered.count.p, want_peaks, F.peaks.x.p, F.peaks.y.p,
                                                                    F.peaks.inten.p, F.peaks.az.p, F.peaks.rg.p, F.peaks.count.p, mot_dev, ccw);
  launched(ctx, "k2_make_clouds");
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

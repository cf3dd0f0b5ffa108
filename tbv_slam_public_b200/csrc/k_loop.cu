// k_loop.cu — loop-closure keyframe database + batched, shardable candidate registration.
//
// Replaces the per-candidate body of ScanContextClosure::SearchAndAddConstraint (tbv_slam/src/tbv_slam/loopclosure.cpp:658-721):
// RegisterLoopCandidate (:320-364) -> loopclosure::Register (:35-97), i.e. one n_scan_normal_reg::Register (P2L, Huber 0.1,
// uniform weights, SetParameters(4,10)) per (from, to) candidate against the cell sets stored in the pose graph's nodes.
// Here the cell sets of every keyframe live in HBM (tbv_loopdb) together with their 4 m search grids, a batch of candidates
// is one k_register launch (one CTA per candidate), and the accepted candidates are packed on the device, in candidate
// order, into fixed-size tbv_constraint records — the unit that is all-gathered between GPUs when the candidate list is
// sharded over ranks.
#include <cmath>

#include "tbv_comm.cuh"
#include "tbv_reg.cuh"

namespace tbv {

struct Candidate {   // per candidate, uploaded once per batch
  int from, to, index;
  double quality[2];
};

// One CTA: ordered compaction of the accepted candidates into tbv_constraint records.
__global__ void __launch_bounds__(256)
k_pack_constraints(const RegResult* __restrict__ res, const Candidate* __restrict__ cand, int n, double max_score,
                   tbv_constraint* __restrict__ out, int out_cap, int* __restrict__ n_out) {
  __shared__ int s_w[8];
  __shared__ int s_running;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    bool ok = false;
    RegResult r;
    if (i < n) {
      r = res[i];
      ok = r.success != 0 && (!(max_score > 0.0) || r.score <= max_score);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    if (ok) {
      int q = s_running + __popc(bal & ((1u << lane) - 1u));
      for (int w = 0; w < warp; w++) q += s_w[w];
      if (q < out_cap) {
        const Candidate c = cand[i];
        tbv_constraint k;
        k.id_begin = c.from; k.id_end = c.to; k.type = 1; k.candidate = c.index;
        k.t_be[0] = r.align[0]; k.t_be[1] = r.align[1]; k.t_be[2] = r.align[2];
        // reg_cov.block<3,3>(0,0) = Rinv * C * Rinv^T with C = diag(0.1^2, 0.1^2, 0) and Rinv = Trevised^-1 rotation
        // (n_scan_normal.cpp:171-175, loopclosure.cpp:93); yaw variance 0.01^2 untouched
        const double cth = cos(r.pose[2]), sth = sin(r.pose[2]);
        const double r00 = cth, r01 = sth, r10 = -sth, r11 = cth;   // rotation of the inverse
        const double v = 0.1 * 0.1;
        const double a00 = r00 * v, a01 = r01 * v, a10 = r10 * v, a11 = r11 * v;   // Rinv * C
        k.cov[0] = a00 * r00 + a01 * r01;
        k.cov[1] = a00 * r10 + a01 * r11;
        k.cov[2] = a10 * r10 + a11 * r11;
        k.cov[3] = 0.01 * 0.01;
        k.score = r.score;
        k.t_revised[0] = r.pose[0]; k.t_revised[1] = r.pose[1]; k.t_revised[2] = atan2(sth, cth);
        k.itrs = r.itrs; k.num_residuals = r.num_residuals;
        k.quality[0] = c.quality[0]; k.quality[1] = c.quality[1];
        out[q] = k;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += s_w[w];
      s_running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = min(s_running, out_cap);
}

}  // namespace tbv

using namespace tbv;

// One batch in flight of the pipelined sharded registration (tbv_loopdb_submit_sharded / tbv_loopdb_collect_sharded).  Every slot owns
// all the memory its batch touches, so that batch i + 1 can be uploaded and registered while batch i is still in its exchange.
struct LoopSlot {
  uint8_t* up_host = nullptr;        // pinned: [problems | fixed_set | fixed_pose | candidates] of this rank's share, ONE H2D copy
  size_t up_host_bytes = 0;
  DevBuf<uint8_t> up_dev;            // the same block on the device
  DevBuf<RegResult> results;
  DevBuf<tbv_constraint> send;       // [capacity + 1]: header (count) + this rank's accepted records
  DevBuf<tbv_constraint> recv;       // [world][capacity + 1]
  DevBuf<tbv_constraint> all;        // [world * capacity] merged, global candidate order
  DevBuf<int> n_all;
  tbv_constraint* down_host = nullptr;   // pinned: merged records, then one record-sized slot holding the count
  size_t down_host_n = 0;
  int capacity = 0;                  // records per rank the exchange buffers are sized for
  int area = 0;                      // world * capacity of the batch in flight
  bool busy = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start, registered + packed, merged, on the host
  void release() {
    for (cudaEvent_t& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    if (up_host) cudaFreeHost(up_host);
    if (down_host) cudaFreeHost(down_host);
    up_host = nullptr; down_host = nullptr; up_host_bytes = 0; down_host_n = 0; capacity = 0; busy = false;
    up_dev.release(); results.release(); send.release(); recv.release(); all.release(); n_all.release();
  }
};
constexpr int LOOP_SLOTS = 2;

struct tbv_loopdb {
  tbv_ctx* ctx = nullptr;
  int max_kf = 0, cell_cap = 0, n_kf = 0;
  CellStore store;             // [max_kf] cell sets, field-major
  GridStore grids;
  DevBuf<SetView> views;       // [max_kf]
  std::vector<int> n_cells;    // host copy of the set sizes
  // per-batch scratch
  DevBuf<RegProblem> problems;
  DevBuf<int> fixed_set;
  DevBuf<double> fixed_pose;
  DevBuf<RegResult> results;
  DevBuf<Candidate> cand;
  DevBuf<tbv_constraint> out;
  DevBuf<int> n_out;
  // pipelined sharded registration: LOOP_SLOTS batches in flight, submitted and collected in order
  LoopSlot slot[LOOP_SLOTS];
  long long n_submitted = 0, n_collected = 0;
  cudaStream_t xstream = nullptr;   // the exchange (all-gather, merge, D2H) of a batch runs here, behind its registration on the context's stream
  void release() {
    for (LoopSlot& sl : slot) sl.release();
    if (xstream) cudaStreamDestroy(xstream);
    xstream = nullptr;
    store.release(); grids.release(); views.release(); problems.release(); fixed_set.release(); fixed_pose.release();
    results.release(); cand.release(); out.release(); n_out.release();
  }
};

namespace {

template <typename T>
int h2d(tbv_ctx* ctx, DevBuf<T>& d, const std::vector<T>& h) {
  int rc = d.reserve(h.size() ? h.size() : 1);
  if (rc) return rc;
  if (!h.empty()) TBV_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return TBV_OK;
}

// enqueue registration + packing of one candidate batch; the result record list is at out_dev / n_out_dev
int loopdb_enqueue(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                   const int* candidate_index, const double* quality, const tbv_reg_params* params, double max_score,
                   tbv_constraint* out_dev, int out_capacity, int* n_out_dev) {
  tbv_ctx* ctx = db->ctx;
  std::vector<RegProblem> hp(n_cand);
  std::vector<int> hfs(n_cand);
  std::vector<double> hfp((size_t)n_cand * 3);
  std::vector<Candidate> hc(n_cand);
  int slot_cap = 1;
  for (int p = 0; p < n_cand; p++) {
    TBV_REQUIRE(from[p] >= 0 && from[p] < db->n_kf && to[p] >= 0 && to[p] < db->n_kf, "candidate indexes a keyframe that is not in the database");
    hp[p].n_fixed = 1; hp[p].fixed_first = p; hp[p].src_set = from[p]; hp[p].active = 1;
    hfs[p] = to[p];
    for (int c = 0; c < 3; c++) { hp[p].src_pose[c] = T_from[3 * p + c]; hfp[3 * p + c] = T_to[3 * p + c]; }
    hc[p].from = from[p]; hc[p].to = to[p]; hc[p].index = candidate_index ? candidate_index[p] : p;
    hc[p].quality[0] = quality ? quality[2 * p] : 0.0;
    hc[p].quality[1] = quality ? quality[2 * p + 1] : 0.0;
    if (db->n_cells[from[p]] > slot_cap) slot_cap = db->n_cells[from[p]];
  }
  int rc;
  if ((rc = h2d(ctx, db->problems, hp)) || (rc = h2d(ctx, db->fixed_set, hfs)) || (rc = h2d(ctx, db->fixed_pose, hfp)) ||
      (rc = h2d(ctx, db->cand, hc)) || (rc = db->results.reserve(n_cand)))
    return rc;
  // the pageable host vectors above are consumed by the (staged) copies before cudaMemcpyAsync returns
  rc = register_launch(ctx, REG_MODE_REGISTER, 0, db->views.p, db->problems.p, db->fixed_set.p, db->fixed_pose.p, n_cand, 1, slot_cap,
                       db->cell_cap, to_dev(*params), db->results.p, nullptr, false);
  if (rc) return rc;
  k_pack_constraints<<<1, 256, 0, ctx->stream>>>(db->results.p, db->cand.p, n_cand, max_score, out_dev, out_capacity, n_out_dev);
  launched(ctx, "k_pack_constraints");
  TBV_CUDA(cudaGetLastError());
  return TBV_OK;
}

}  // namespace

extern "C" {

tbv_loopdb* tbv_loopdb_create(tbv_ctx* ctx, int max_keyframes, int cell_capacity) {
  TBV_ENTER(ctx);
  if (!ctx || max_keyframes < 1 || cell_capacity < 1) { set_error("tbv_loopdb_create: bad arguments"); return nullptr; }
  tbv_loopdb* db = new tbv_loopdb();
  db->ctx = ctx; db->max_kf = max_keyframes; db->cell_cap = cell_capacity;
  if (db->store.reserve(max_keyframes, cell_capacity) || db->grids.reserve(max_keyframes, cell_capacity) || db->views.reserve(max_keyframes) ||
      db->n_out.reserve(1)) {
    db->release();
    delete db;
    return nullptr;
  }
  db->n_cells.assign(max_keyframes, 0);
  return db;
}

void tbv_loopdb_destroy(tbv_loopdb* db) {
  TBV_ENTER(db ? db->ctx : nullptr);
  if (!db) return;
  cudaStreamSynchronize(db->ctx->stream);
  db->release();
  delete db;
}

int tbv_loopdb_size(tbv_loopdb* db) { return db ? db->n_kf : TBV_ERR_INVALID; }

int tbv_loopdb_add(tbv_loopdb* db, int n_sets, const tbv_cell* const* sets, const int* n_cells, int* first_id) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && sets && n_cells && n_sets >= 0, "bad arguments");
  TBV_REQUIRE(db->n_kf + n_sets <= db->max_kf, "keyframe database is full");
  tbv_ctx* ctx = db->ctx;
  if (first_id) *first_id = db->n_kf;
  if (n_sets == 0) return TBV_OK;
  std::vector<SetView> hv(n_sets);
  for (int i = 0; i < n_sets; i++) {
    const int id = db->n_kf + i;
    TBV_REQUIRE(n_cells[i] >= 0 && n_cells[i] <= db->cell_cap && (n_cells[i] == 0 || sets[i]), "bad cell set (too large for the database's cell capacity?)");
    int rc = cells_upload(ctx, sets[i], n_cells[i], db->store.set_ptr(id), db->cell_cap);
    if (rc) return rc;
    hv[i] = db->grids.view(id, db->store.set_ptr(id), db->cell_cap, nullptr, n_cells[i]);
    db->n_cells[id] = n_cells[i];
  }
  TBV_CUDA(cudaMemcpyAsync(db->views.p + db->n_kf, hv.data(), n_sets * sizeof(SetView), cudaMemcpyHostToDevice, ctx->stream));
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));  // hv goes out of scope
  int rc = cellgrid_build_launch(ctx, db->views.p + db->n_kf, nullptr, n_sets, n_sets, db->cell_cap);
  if (rc) return rc;
  db->n_kf += n_sets;
  return TBV_OK;
}

int tbv_loopdb_register_dev(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                            const int* candidate_index, const double* quality, const tbv_reg_params* params, double max_score,
                            tbv_constraint* out_dev, int out_capacity, int* n_out_dev) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && n_cand >= 0 && params && out_dev && n_out_dev && out_capacity >= 0, "bad arguments");
  if (n_cand == 0) {
    TBV_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int), db->ctx->stream));
    return TBV_OK;
  }
  TBV_REQUIRE(from && to && T_from && T_to, "bad arguments");
  return loopdb_enqueue(db, n_cand, from, to, T_from, T_to, candidate_index, quality, params, max_score, out_dev, out_capacity, n_out_dev);
}

int tbv_loopdb_register(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                        const int* candidate_index, const double* quality, const tbv_reg_params* params, double max_score,
                        tbv_constraint* out, int out_capacity, int* n_out, tbv_reg_summary* summaries) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && n_cand >= 0 && params && n_out && out_capacity >= 0 && (out || out_capacity == 0), "bad arguments");
  *n_out = 0;
  if (n_cand == 0) return TBV_OK;
  TBV_REQUIRE(from && to && T_from && T_to, "bad arguments");
  tbv_ctx* ctx = db->ctx;
  int rc = db->out.reserve(out_capacity > 0 ? out_capacity : 1);
  if (rc) return rc;
  rc = loopdb_enqueue(db, n_cand, from, to, T_from, T_to, candidate_index, quality, params, max_score, db->out.p, out_capacity, db->n_out.p);
  if (rc) return rc;
  TBV_CUDA(cudaMemcpyAsync(n_out, db->n_out.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<RegResult> hr;
  if (summaries) {
    hr.resize(n_cand);
    TBV_CUDA(cudaMemcpyAsync(hr.data(), db->results.p, n_cand * sizeof(RegResult), cudaMemcpyDeviceToHost, ctx->stream));
  }
  TBV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (*n_out > 0) TBV_CUDA(cudaMemcpy(out, db->out.p, (size_t)*n_out * sizeof(tbv_constraint), cudaMemcpyDeviceToHost));
  if (summaries)
    for (int p = 0; p < n_cand; p++) {
      const RegResult& r = hr[p];
      tbv_reg_summary* s = summaries + p;
      s->success = r.success; s->itrs = r.itrs; s->lm_iterations = r.lm_iterations; s->num_residuals = r.num_residuals;
      s->last_n_iterations = r.last_n_iterations; s->termination = r.termination; s->score = r.score; s->final_cost = r.final_cost;
      s->last_relative_decrease = r.last_relative_decrease;
    }
  return TBV_OK;
}

int tbv_loopdb_submit_sharded(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                              const double* quality, const tbv_reg_params* params, double max_score) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && n_cand >= 0 && params, "bad arguments");
  TBV_REQUIRE(n_cand == 0 || (from && to && T_from && T_to), "bad arguments");
  TBV_REQUIRE(db->n_submitted - db->n_collected < LOOP_SLOTS, "two batches are in flight already: collect one first");
  tbv_ctx* ctx = db->ctx;
  const int world = comm_world(ctx), rank = comm_rank(ctx);
  LoopSlot& sl = db->slot[db->n_submitted % LOOP_SLOTS];
  // this rank's share: from mod world == rank, ascending global index; capacity = the largest share over ranks, which
  // every rank computes from the replicated list
  int share[64] = {0};
  for (int p = 0; p < n_cand; p++) {
    TBV_REQUIRE(from[p] >= 0 && from[p] < db->n_kf && to[p] >= 0 && to[p] < db->n_kf, "candidate indexes a keyframe that is not in the database");
    share[from[p] % world]++;
  }
  int capacity = 1;
  for (int r = 0; r < world; r++) capacity = std::max(capacity, share[r]);
  const int n_mine = share[rank];
  // upload block: [RegProblem | fixed_set | fixed_pose | Candidate] x n_mine, sections 16-byte aligned
  auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t o_prob = 0, o_fs = al16(o_prob + (size_t)n_mine * sizeof(RegProblem)), o_fp = al16(o_fs + (size_t)n_mine * sizeof(int)),
               o_cand = al16(o_fp + (size_t)n_mine * 3 * sizeof(double)), up_bytes = al16(o_cand + (size_t)n_mine * sizeof(Candidate)) + 16;
  int rc;
  if (sl.up_host_bytes < up_bytes) {
    if (sl.up_host) cudaFreeHost(sl.up_host);
    sl.up_host = nullptr; sl.up_host_bytes = 0;
    TBV_CUDA(cudaHostAlloc((void**)&sl.up_host, up_bytes, cudaHostAllocDefault));
    sl.up_host_bytes = up_bytes;
  }
  if ((rc = sl.up_dev.reserve(up_bytes)) || (rc = sl.results.reserve(n_mine > 0 ? n_mine : 1)) || (rc = sl.n_all.reserve(1))) return rc;
  if (capacity > sl.capacity) {   // the three exchange buffers are sized together: the merge addresses recv as [world][capacity + 1]
    sl.capacity = 0;
    if ((rc = sl.send.reserve((size_t)capacity + 1)) || (rc = sl.recv.reserve((size_t)world * ((size_t)capacity + 1))) ||
        (rc = sl.all.reserve((size_t)world * capacity)))
      return rc;
    sl.capacity = capacity;
  }
  const size_t area = (size_t)world * (size_t)capacity;
  if (sl.down_host_n < area + 1) {
    if (sl.down_host) cudaFreeHost(sl.down_host);
    sl.down_host = nullptr; sl.down_host_n = 0;
    TBV_CUDA(cudaHostAlloc((void**)&sl.down_host, (area + 1) * sizeof(tbv_constraint), cudaHostAllocDefault));
    sl.down_host_n = area + 1;
  }
  for (cudaEvent_t& e : sl.ev)
    if (!e) TBV_CUDA(cudaEventCreate(&e));
  if (!db->xstream) TBV_CUDA(cudaStreamCreateWithFlags(&db->xstream, cudaStreamNonBlocking));
  RegProblem* hp = reinterpret_cast<RegProblem*>(sl.up_host + o_prob);
  int* hfs = reinterpret_cast<int*>(sl.up_host + o_fs);
  double* hfp = reinterpret_cast<double*>(sl.up_host + o_fp);
  Candidate* hc = reinterpret_cast<Candidate*>(sl.up_host + o_cand);
  int slot_cap = 1;
  for (int p = 0, i = 0; p < n_cand; p++) {
    if (from[p] % world != rank) continue;
    hp[i].n_fixed = 1; hp[i].fixed_first = i; hp[i].src_set = from[p]; hp[i].active = 1;
    hfs[i] = to[p];
    for (int c = 0; c < 3; c++) { hp[i].src_pose[c] = T_from[3 * (size_t)p + c]; hfp[3 * (size_t)i + c] = T_to[3 * (size_t)p + c]; }
    hc[i].from = from[p]; hc[i].to = to[p]; hc[i].index = p;
    hc[i].quality[0] = quality ? quality[2 * (size_t)p] : 0.0;
    hc[i].quality[1] = quality ? quality[2 * (size_t)p + 1] : 0.0;
    if (db->n_cells[from[p]] > slot_cap) slot_cap = db->n_cells[from[p]];
    i++;
  }
  // ---- registration + packing on the context's stream ---------------------------------------------------------------------------------
  TBV_CUDA(cudaEventRecord(sl.ev[0], ctx->stream));
  if (n_mine > 0) {
    TBV_CUDA(cudaMemcpyAsync(sl.up_dev.p, sl.up_host, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const RegProblem* dp = reinterpret_cast<const RegProblem*>(sl.up_dev.p + o_prob);
    rc = register_launch(ctx, REG_MODE_REGISTER, 0, db->views.p, dp, reinterpret_cast<const int*>(sl.up_dev.p + o_fs),
                         reinterpret_cast<const double*>(sl.up_dev.p + o_fp), n_mine, 1, slot_cap, db->cell_cap, to_dev(*params), sl.results.p,
                         nullptr, false);
    if (rc) return rc;
    k_pack_constraints<<<1, 256, 0, ctx->stream>>>(sl.results.p, reinterpret_cast<const Candidate*>(sl.up_dev.p + o_cand), n_mine, max_score,
                                                   sl.send.p + 1, capacity, reinterpret_cast<int*>(sl.send.p));
    launched(ctx, "k_pack_constraints");
    TBV_CUDA(cudaGetLastError());
  } else {
    TBV_CUDA(cudaMemsetAsync(sl.send.p, 0, sizeof(int), ctx->stream));
  }
  TBV_CUDA(cudaEventRecord(sl.ev[1], ctx->stream));
  // ---- exchange + return of the merged records on the exchange stream (the context's own stream while a profile is being taken, so that
  //      the per-launch events of tbv_profile_begin/_end see it) -----------------------------------------------------------------------------
  cudaStream_t xs = ctx->prof.on ? ctx->stream : db->xstream;
  if (xs != ctx->stream) TBV_CUDA(cudaStreamWaitEvent(xs, sl.ev[1], 0));
  if ((rc = comm_allgather_merge_on(ctx, xs, sl.send.p, sl.recv.p, sl.all.p, sl.n_all.p, capacity))) return rc;
  TBV_CUDA(cudaEventRecord(sl.ev[2], xs));
  // The count and the records cross in ONE round (the merged area is what the ranks can have accepted: a few hundred KB); a caller
  // whose array is smaller than the number of accepted records gets TBV_ERR_CAPACITY from the collect.
  int* count = reinterpret_cast<int*>(sl.down_host + area);
  TBV_CUDA(cudaMemcpyAsync(count, sl.n_all.p, sizeof(int), cudaMemcpyDeviceToHost, xs));
  TBV_CUDA(cudaMemcpyAsync(sl.down_host, sl.all.p, area * sizeof(tbv_constraint), cudaMemcpyDeviceToHost, xs));
  TBV_CUDA(cudaEventRecord(sl.ev[3], xs));
  sl.area = (int)area;
  sl.busy = true;
  db->n_submitted++;
  return TBV_OK;
}

int tbv_loopdb_collect_sharded(tbv_loopdb* db, tbv_constraint* all, int all_capacity, int* n_all, float* timing_ms) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && n_all && all_capacity >= 0 && (all || all_capacity == 0), "bad arguments");
  *n_all = 0;
  TBV_REQUIRE(db->n_collected < db->n_submitted, "no batch in flight");
  LoopSlot& sl = db->slot[db->n_collected % LOOP_SLOTS];
  TBV_CUDA(cudaEventSynchronize(sl.ev[3]));
  sl.busy = false;
  db->n_collected++;
  const int n = *reinterpret_cast<const int*>(sl.down_host + sl.area);
  *n_all = n;
  const size_t take = (size_t)std::max(0, std::min(n, all_capacity));
  if (take) memcpy(all, sl.down_host, take * sizeof(tbv_constraint));
  if (timing_ms) {
    // {H2D of the share + registration + packing, all-gather + merge, D2H of the merged records, whole batch}; the collective and the
    // merge kernel are timed together here, their split is available through tbv_profile_begin/_end ("nccl_all_gather", "k_merge_constraints")
    TBV_CUDA(cudaEventElapsedTime(&timing_ms[0], sl.ev[0], sl.ev[1]));
    TBV_CUDA(cudaEventElapsedTime(&timing_ms[1], sl.ev[1], sl.ev[2]));
    TBV_CUDA(cudaEventElapsedTime(&timing_ms[2], sl.ev[2], sl.ev[3]));
    TBV_CUDA(cudaEventElapsedTime(&timing_ms[3], sl.ev[0], sl.ev[3]));
  }
  if (n > all_capacity) { set_error("tbv_loopdb_collect_sharded: %d constraints accepted, room for %d", n, all_capacity); return TBV_ERR_CAPACITY; }
  return TBV_OK;
}

int tbv_loopdb_register_sharded(tbv_loopdb* db, int n_cand, const int* from, const int* to, const double* T_from, const double* T_to,
                                const double* quality, const tbv_reg_params* params, double max_score, tbv_constraint* all,
                                int all_capacity, int* n_all, float* timing_ms) {
  TBV_ENTER(db ? db->ctx : nullptr);
  TBV_REQUIRE(db && n_cand >= 0 && params && n_all && all_capacity >= 0 && (all || all_capacity == 0), "bad arguments");
  *n_all = 0;
  if (timing_ms) for (int i = 0; i < 4; i++) timing_ms[i] = 0.f;
  TBV_REQUIRE(db->n_submitted == db->n_collected, "a submitted batch is still in flight: collect it first");
  if (n_cand == 0) return TBV_OK;
  int rc = tbv_loopdb_submit_sharded(db, n_cand, from, to, T_from, T_to, quality, params, max_score);
  if (rc) return rc;
  return tbv_loopdb_collect_sharded(db, all, all_capacity, n_all, timing_ms);
}

}  // extern "C"

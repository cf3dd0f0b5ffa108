// tbv_comm.cuh — the one exchange step of the hot path: all-gather of accepted loop constraints between the GPUs that share a
// candidate list (the serial program's PoseGraph::AddConstraintThSafe sequence, tbv_slam/src/tbv_slam/loopclosure.cpp:704-724).
// Product code: never includes anything from oracle/.
#pragma once
#include "tbv_common.cuh"

namespace tbv {

// Per-context communicator state (k_comm.cu).  Exchange block of one rank: [capacity + 1] tbv_constraint-sized records; record 0 is the
// header (first int32 = number of valid records), records 1 .. capacity are the payload, ascending by `candidate`.
struct CommState;

int comm_world(const tbv_ctx* ctx);   // 1 without a communicator
int comm_rank(const tbv_ctx* ctx);    // 0 without a communicator
// (Re)allocates the exchange buffers for `capacity` records per rank (the same value on every rank).
int comm_reserve(tbv_ctx* ctx, int capacity);
tbv_constraint* comm_send_records(tbv_ctx* ctx);   // device: payload area of this rank's block
int* comm_send_count(tbv_ctx* ctx);                // device: header count of this rank's block
// Enqueues on the context's stream: ONE ncclAllGather of the rank blocks (skipped at world 1), then the device merge into global candidate
// order.  Afterwards comm_all(ctx)[0 .. *comm_n_all(ctx)) holds every rank's accepted records, identical on every rank.
int comm_allgather_merge(tbv_ctx* ctx, int capacity);
tbv_constraint* comm_all(tbv_ctx* ctx);
int* comm_n_all(tbv_ctx* ctx);
// The same exchange on the caller's buffers and stream: send [capacity + 1] (header + records), recv [world][capacity + 1] (unused at
// world 1), all [world * capacity], n_all [1].  The collectives of one communicator must be enqueued in the same order on every rank.
int comm_allgather_merge_on(tbv_ctx* ctx, cudaStream_t stream, const tbv_constraint* send, tbv_constraint* recv, tbv_constraint* all,
                            int* n_all, int capacity);
// merged records (and their count) -> host through pinned staging, one stream synchronisation; *n_out may exceed dst_capacity
int comm_fetch_all(tbv_ctx* ctx, tbv_constraint* dst, int dst_capacity, int* n_out);
void comm_release(tbv_ctx* ctx);

}  // namespace tbv

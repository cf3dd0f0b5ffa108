"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the GPUs, gloo in the
CPU tests) — SURVEY 8e.

What shards and how
  * odometry: frame t of a sequence needs the pose of frame t-1 (odometrykeyframefuser.cpp:146-168), so SEQUENCES shard over
    ranks (`shard_sequences`), with no data-path collective — the reference's own scaling model is one worker per sequence;
  * loop closure: the candidates of ScanContextClosure::SearchAndAddConstraint (tbv_slam/src/tbv_slam/loopclosure.cpp:658-721)
    are independent given the keyframe database, so the candidate list shards by `id_from mod world` (`shard_candidates`),
    every rank keeps the whole database (cells ~33 kB / keyframe), registers its share in one launch (tbv_loopdb_register_dev)
    and the accepted constraints — fixed-size 128-byte tbv_constraint records, padded to the largest share — are exchanged
    with ONE all-gather straight from device memory.  Every rank ends up with the same list in global candidate order, which
    is what PoseGraph::AddConstraintThSafe would have seen in the serial program.

The exchange itself lives behind the C-ABI (tbv_comm_init_rank / tbv_loopdb_register_sharded / tbv_allgather_constraints:
ncclAllGather + a device merge, csrc/k_comm.cu; tbv_loopdb_submit_sharded / _collect_sharded keep two batches in flight with the exchange
on a second stream) so that a C++/ROS host can run it; on GPUs this module
only hands the library an ncclUniqueId through torch.distributed's store (`init_comm`).  `all_gather_constraints` below is the
same record exchange over a torch.distributed group and exists for the gloo / world-2 CPU tests of the host logic.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .api import CONSTRAINT_DTYPE, LoopDB, RegParams

RECORD_BYTES = CONSTRAINT_DTYPE.itemsize  # 128


def shard_sequences(n_seq: int, world: int, rank: int) -> range:
    """Contiguous, balanced share of n_seq independent sequences (sizes differ by at most one)."""
    base, extra = divmod(n_seq, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def shard_candidates(id_from, world: int, rank: int) -> np.ndarray:
    """Indices (ascending) of the candidates rank `rank` registers: id_from mod world == rank."""
    id_from = np.asarray(id_from)
    return np.nonzero(id_from % world == rank)[0]


def shard_capacity(id_from, world: int) -> int:
    """Largest share over ranks — every rank can compute it because the candidate list is replicated."""
    id_from = np.asarray(id_from)
    if len(id_from) == 0:
        return 0
    return int(np.bincount(id_from % world, minlength=world).max())


def init_comm(ctx, group=None):
    """Gives `ctx` an NCCL communicator spanning the torch.distributed group: rank 0 draws the ncclUniqueId (tbv_comm_unique_id), the
    group's object broadcast carries its 128 bytes, every rank calls tbv_comm_init_rank.  No-op for a single process."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 1, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init_rank(box[0], world, rank)
    return world, rank


def all_gather_constraints(local: torch.Tensor, count: torch.Tensor, group=None) -> np.ndarray:
    """local: [cap, 128] uint8 records of this rank (first count[0] valid; same cap on every rank), on the GPU (NCCL) or the
    CPU (gloo); count: int32[1] on the same device.  Returns all valid records of all ranks as a CONSTRAINT_DTYPE array sorted
    by (candidate, rank) — identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    cap = local.shape[0]
    if world == 1:
        gathered, counts = local.reshape(1, cap, RECORD_BYTES), count.reshape(1)
    else:
        gathered = torch.empty((world, cap, RECORD_BYTES), dtype=torch.uint8, device=local.device)
        counts = torch.empty((world,), dtype=torch.int32, device=local.device)
        dist.all_gather_into_tensor(counts, count.reshape(1).to(torch.int32), group=group)
        dist.all_gather_into_tensor(gathered.reshape(-1), local.reshape(-1).contiguous(), group=group)
    counts_h = counts.cpu().numpy()
    if int(counts_h.max(initial=0)) > cap:
        raise ValueError(f"a rank reports {int(counts_h.max())} records but the exchange buffers hold {cap}")
    recs_h = gathered.cpu().numpy()
    parts = [recs_h[r, :int(counts_h[r])].reshape(-1).view(CONSTRAINT_DTYPE) for r in range(world)]
    out = np.concatenate(parts) if parts else np.zeros(0, CONSTRAINT_DTYPE)
    return out[np.argsort(out["candidate"], kind="stable")]


class ShardedLoopClosure:
    """Candidate registration sharded over the ranks of a process group; the database is replicated.

    On GPUs (the product path) the whole exchange is ONE library call, tbv_loopdb_register_sharded, over the context's NCCL communicator
    (created here from the torch.distributed group if the context has none).  `transport="torch"` keeps the records on the device and
    gathers them with torch.distributed instead — the form the gloo CPU tests exercise."""

    def __init__(self, db: LoopDB, group=None, transport: str = "nccl"):
        self.db, self.group, self.transport = db, group, transport
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if transport == "nccl":
            if self.world > 1 and db.ctx.comm_world()[0] != self.world:
                init_comm(db.ctx, group)
        else:
            self.stream = torch.cuda.ExternalStream(db.ctx.stream)
        self.last_timing = None

    def register_candidates(self, id_from, id_to, T_from, T_to, quality=None, params: RegParams | None = None, max_score=0.0) -> np.ndarray:
        if self.transport == "nccl":
            out, self.last_timing = self.db.register_sharded(id_from, id_to, T_from, T_to, quality=quality, params=params, max_score=max_score,
                                                             want_timing=True)
            return out
        id_from = np.asarray(id_from, np.int32); id_to = np.asarray(id_to, np.int32)
        T_from = np.asarray(T_from, np.float64).reshape(-1, 3); T_to = np.asarray(T_to, np.float64).reshape(-1, 3)
        mine = shard_candidates(id_from, self.world, self.rank)
        cap = max(shard_capacity(id_from, self.world), 1)
        with torch.cuda.stream(self.stream):   # everything below is ordered on the library's stream
            buf = torch.empty((cap, RECORD_BYTES), dtype=torch.uint8, device="cuda")
            cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
            self.db.register_candidates_dev(id_from[mine], id_to[mine], T_from[mine], T_to[mine], buf.data_ptr(), cap, cnt.data_ptr(),
                                            candidate_index=mine.astype(np.int32),
                                            quality=None if quality is None else np.asarray(quality, np.float64).reshape(-1, 2)[mine],
                                            params=params, max_score=max_score)
            return all_gather_constraints(buf, cnt, self.group)

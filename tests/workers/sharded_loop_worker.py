"""World-N worker of tests/test_comm_gpu.py (launched with torch.distributed.run, one rank per GPU): every rank builds the same keyframe
database, registers the SAME global candidate list through tbv_loopdb_register_sharded (NCCL all-gather inside the library) and writes the
records it received; rank 0 also writes the single-GPU result of tbv_loopdb_register for the comparison."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tbv_slam_public_b200 import api, parallel, synth  # noqa: E402


def main(out_dir, n_cand=96):
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    st = synth.make_stream(8)
    sets = []
    f, _ = ctx.StructuredKStrongest(st.scans, peaks=False)
    for b in range(8):
        az, rg, inten, x, y = f.scan(b)
        cells, _ = ctx.MapPointNormal(x, y, inten.astype(np.float32), radius=3.0, weight_intensity=True, capacity=2048)
        sets.append(cells)
    db = api.LoopDB(ctx, 8, max(len(c) for c in sets))
    db.add(sets)
    rng = np.random.default_rng(3)
    fr = rng.integers(0, 8, n_cand).astype(np.int32)
    to = ((fr + rng.integers(1, 4, n_cand)) % 8).astype(np.int32)
    err = np.stack([rng.uniform(-1.5, 1.5, n_cand), rng.uniform(-1.5, 1.5, n_cand), rng.uniform(-0.1, 0.1, n_cand)], axis=1)
    Tf, Tt = st.gt[fr] + err, st.gt[to]
    quality = np.stack([np.arange(n_cand, dtype=np.float64), -np.arange(n_cand, dtype=np.float64)], axis=1)
    single = db.register_candidates(fr, to, Tf, Tt, quality=quality)     # before the communicator exists: plain single-GPU call
    w, r = parallel.init_comm(ctx)
    assert (w, r) == (world, rank) == ctx.comm_world()
    got, timing = db.register_sharded(fr, to, Tf, Tt, quality=quality, want_timing=True)
    again = db.register_sharded(fr, to, Tf, Tt, quality=quality)         # buffers reused
    # a smaller batch than the exchange buffers were sized for, and a batch in which one rank has no candidate at all
    small = db.register_sharded(fr[:5], to[:5], Tf[:5], Tt[:5])
    zeros = np.zeros(7, np.int32)                                       # every candidate has from = 0 -> all on rank 0
    lop = db.register_sharded(zeros, to[:7] % 7 + 1, st.gt[zeros] + err[:7], st.gt[to[:7] % 7 + 1])
    lop_ref = db.register_candidates(zeros, to[:7] % 7 + 1, st.gt[zeros] + err[:7], st.gt[to[:7] % 7 + 1])
    # pipelined form: two batches in flight (different sizes), collected in submission order; a third submit is refused
    db.submit_sharded(fr, to, Tf, Tt, quality=quality)
    db.submit_sharded(fr[:5], to[:5], Tf[:5], Tt[:5])
    try:
        db.submit_sharded(fr[:5], to[:5], Tf[:5], Tt[:5])
        refused = False
    except api.TbvError:
        refused = True
    assert refused
    pipe0 = db.collect_sharded()
    db.submit_sharded(fr, to, Tf, Tt, quality=quality)                  # slot 0 again while the 5-candidate batch is in flight
    pipe1 = db.collect_sharded()
    pipe2, ptiming = db.collect_sharded(want_timing=True)
    assert ptiming[3] > 0
    # the low-level export: this rank's share packed on the device, then tbv_allgather_constraints
    mine = parallel.shard_candidates(fr, world, rank)
    cap = parallel.shard_capacity(fr, world)
    buf = torch.zeros((cap, 128), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    db.register_candidates_dev(fr[mine], to[mine], Tf[mine], Tt[mine], buf.data_ptr(), cap, cnt.data_ptr(), candidate_index=mine.astype(np.int32),
                               quality=quality[mine])
    low = ctx.allgather_constraints(buf.data_ptr(), cnt.data_ptr(), cap)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), single=single.view(np.uint8), got=got.view(np.uint8), again=again.view(np.uint8),
             small=small.view(np.uint8), small_ref=db.register_candidates(fr[:5], to[:5], Tf[:5], Tt[:5]).view(np.uint8),
             lop=lop.view(np.uint8), lop_ref=lop_ref.view(np.uint8),
             pipe0=pipe0.view(np.uint8), pipe1=pipe1.view(np.uint8), pipe2=pipe2.view(np.uint8), low=low.view(np.uint8), timing=np.array(timing), launches=ctx.launch_count())
    dist.barrier()
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])

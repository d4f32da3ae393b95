"""N>1 host-side path on CPU: world_size-2 gloo job shards the frame stream, reduces the timing scalar
(max over ranks) and gathers the per-frame detection records on rank 0 -- no data-path collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manual_yolo_b200 import geometry, multigpu, pipeline


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = multigpu.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = multigpu.my_frames(n_frames, r, w)
    # stand-in for the device path: every frame f yields (f % 3) detections with class f % 64
    rows = torch.zeros((hi - lo, 4, 6))
    count = torch.zeros((hi - lo,), dtype=torch.int32)
    for i, f in enumerate(range(lo, hi)):
        count[i] = f % 3
        rows[i, :, 5] = f % 64
        rows[i, :, 4] = 0.5 + 0.001 * f
        rows[i, :, :4] = torch.tensor([f + 0.9, 1.2, f + 10.7, 20.5])
    recs = pipeline.detections_to_records(rows, count, frame_offset=lo)
    multigpu.barrier()
    t = multigpu.max_over_ranks(10.0 + rank)            # per-rank elapsed ms -> job time = slowest rank
    total = multigpu.sum_over_ranks(hi - lo)
    merged = multigpu.gather_records(recs, dst=0)
    if rank == 0:
        q.put((t, total, merged))
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    world, n_frames = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, total, merged = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t == 11.0 and total == n_frames
    exp = [(f, f % 64) for f in range(n_frames) for _ in range(f % 3)]
    assert [(r["frame"], r["class_id"]) for r in merged] == exp
    assert merged[0]["bbox"] == [1, 1, 11, 20] and merged[0]["tracker_id"] == -1
    # the partition is a disjoint cover
    spans = [geometry.shard_range(n_frames, r, world) for r in range(world)]
    assert spans == [(0, 18), (18, 37)]


def test_single_process_helpers_need_no_group():
    assert multigpu.max_over_ranks(3.5) == 3.5 and multigpu.sum_over_ranks(2) == 2.0
    recs = [{"frame": 2}, {"frame": 0}]
    assert [r["frame"] for r in multigpu.gather_records(recs)] == [0, 2]

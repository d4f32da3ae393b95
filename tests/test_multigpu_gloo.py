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
    col = multigpu.gather_columnar(multigpu.detections_columnar(rows, count, frame_offset=lo), dst=0)
    if rank == 0:
        q.put((t, total, merged, col))
    else:
        assert merged is None and col is None
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    world, n_frames = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, total, merged, col = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t == 11.0 and total == n_frames
    exp = [(f, f % 64) for f in range(n_frames) for _ in range(f % 3)]
    assert [(r["frame"], r["class_id"]) for r in merged] == exp
    assert merged[0]["bbox"] == [1, 1, 11, 20] and merged[0]["tracker_id"] == -1
    # the columnar gather (one structured array per rank, no per-detection Python objects) carries the same records
    assert col.dtype.names == ("frame", "x1", "y1", "x2", "y2", "conf", "class_id") and len(col) == len(exp)
    assert list(zip(col["frame"].tolist(), col["class_id"].tolist())) == exp
    assert multigpu.columnar_to_records(col) == merged
    # the partition is a disjoint cover
    spans = [geometry.shard_range(n_frames, r, world) for r in range(world)]
    assert spans == [(0, 18), (18, 37)]


def test_single_process_helpers_need_no_group():
    assert multigpu.max_over_ranks(3.5) == 3.5 and multigpu.sum_over_ranks(2) == 2.0
    recs = [{"frame": 2}, {"frame": 0}]
    assert [r["frame"] for r in multigpu.gather_records(recs)] == [0, 2]
    rows = torch.zeros((2, 3, 6))
    rows[1, 0] = torch.tensor([1.5, 2.5, 3.5, 4.5, 0.25, 7.0])
    col = multigpu.gather_columnar(multigpu.detections_columnar(rows, torch.tensor([0, 1]), frame_offset=5))
    assert len(col) == 1 and int(col["frame"][0]) == 6 and int(col["class_id"][0]) == 7 and float(col["x2"][0]) == 3.5

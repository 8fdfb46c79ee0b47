"""N > 1 path on CPU: world_size-2 gloo.  Streams are independent sessions: sharding two streams over
two ranks gives exactly the results of running both in one process, and the timing reduction is the
max over ranks."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from detsam2_b200 import streams


def test_assign_streams_is_a_partition():
    for n, w in ((8, 8), (8, 3), (1, 4), (0, 2), (17, 4)):
        parts = streams.assign_streams(n, w)
        assert len(parts) == w
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        streams.assign_streams(4, 0)


def _stream_digest(stream_id):
    """One short session on the CPU oracle engine (test-only engine) -> digest of everything it produced."""
    from detsam2_b200.config import get_config
    from detsam2_b200.predictor import SAM2VideoPredictor
    from detsam2_b200.synthetic import BilliardVideo
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=True), fill_hole_area=8)
    vid = BilliardVideo(num_objects=2, height=96, width=128, num_frames=3, seed=100 + stream_id)
    h = hashlib.sha256()
    with torch.inference_mode():
        st = pred.init_state(list(vid.frames()))
        for oid, box in vid.boxes(0).items():
            pred.add_new_points_or_box(st, 0, oid, box=box)
        for f, ids, m in pred.propagate_in_video(st):
            h.update(np.ascontiguousarray((m > 0).numpy()).tobytes())
            h.update(np.round(m.numpy(), 3).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, num_streams, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert streams.dist_env() == (rank, rank, world)
        mine = streams.assign_streams(num_streams, world)[rank]
        digests = {s: _stream_digest(s) for s in mine}
        streams.barrier()
        # rank r pretends it needed (r + 1) * 100 ms for 5 frames
        fps, ms = streams.aggregate_throughput(5, (rank + 1) * 100.0)
        gathered = [None] * world
        dist.all_gather_object(gathered, digests)   # test-only collection of results (not a data-path collective)
        if rank == 0:
            q.put((fps, ms, gathered))
    finally:
        dist.destroy_process_group()


def test_two_ranks_two_streams_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    fps, ms, gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 200.0                      # max over ranks
    assert abs(fps - 2 * 5 / 0.2) < 1e-9    # all ranks' frames / slowest rank
    merged = {}
    for d in gathered:
        merged.update(d)
    torch.set_num_threads(2)
    assert merged == {0: _stream_digest(0), 1: _stream_digest(1)}
    assert merged[0] != merged[1]

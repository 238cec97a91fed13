"""Multi-rank frame sharding on CPU: two gloo ranks shard a clip by frame index, each runs its shard through the
frame-level pipeline with stand-in estimators (the real ones need a GPU), results are gathered on rank 0 and must
be identical, in order, to the single-rank run. No data-path collective exists; the gather is host plumbing."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import isl_b200  # noqa: F401
from isl_b200.extract import KeypointExtractor, merge_shards, shard_indices


class FakeBody:
    """Deterministic stand-in with Body's batch interface: results depend only on the frame content."""

    def batch(self, frames):
        out = []
        for f in frames:
            s = int(np.asarray(f, dtype=np.int64).sum())
            n = s % 5
            cand = np.arange(n * 4, dtype=np.float64).reshape(n, 4) + s if n else np.array([])
            out.append((cand, np.full((n % 3, 20), float(s % 97))))
        return out


class FakeHand:
    def batch(self, crops):
        return [np.full((21, 2), int(c.sum()) % 50, dtype=np.int64) for c in crops]


def _frames(n):
    return [np.random.RandomState(i).randint(0, 256, (48, 64, 3)).astype(np.uint8) for i in range(n)]


def _boxes(n):
    return [[[4 + (i % 3), 5, 20, True], [30, 8 + (i % 2), 16, False]] for i in range(n)]


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = KeypointExtractor(FakeBody(), FakeHand())
    mine = ex.run_sharded(_frames(n_frames), rank, world, batch_size=3, hand_boxes=_boxes(n_frames))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank == 0:
        q.put(merge_shards(gathered, n_frames))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _same(a, b):
    return (a[0].shape == b[0].shape and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and
            len(a[2]) == len(b[2]) and all(np.array_equal(x, y) for x, y in zip(a[2], b[2])))


def test_two_rank_sharding_matches_single_rank():
    n_frames = 11   # odd on purpose: the shards are ragged
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = KeypointExtractor(FakeBody(), FakeHand()).run_sharded(_frames(n_frames), 0, 1, batch_size=4,
                                                                   hand_boxes=_boxes(n_frames))
    assert len(merged) == n_frames
    assert all(_same(a, b) for a, b in zip(merged, serial))


def test_shard_indices_partition():
    for n in (0, 1, 7, 30):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))
    assert merge_shards([[0, 2, 4], [1, 3]], 5) == [0, 1, 2, 3, 4]


def test_hand_peak_offsets_follow_the_reference_rule():
    """demo.py:36-37: only non-zero coordinates are shifted by the box origin."""

    class ZeroHand:
        def batch(self, crops):
            p = np.zeros((21, 2), dtype=np.int64)
            p[3] = [7, 0]
            return [p.copy() for _ in crops]

    ex = KeypointExtractor(FakeBody(), ZeroHand())
    (_, _, hands), = ex.batch(_frames(1), hand_boxes=[[[10, 20, 16, True]]])
    assert hands[0][3].tolist() == [17, 0] and hands[0][0].tolist() == [0, 0]

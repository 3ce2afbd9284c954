"""Host logic: frame-range sharding with a one-frame halo reproduces the sequential result exactly.
The N>1 path is exercised with two gloo processes on CPU; the per-batch processor is the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spvo_b200  # noqa: E402,F401
from spvo_b200.sequence import plan_batches, plan_shards, run_shard, run_sharded  # noqa: E402

H, W, K, NF = 64, 96, 60, 7


def test_plan_shards_and_batches():
    assert plan_shards(4541, 8) == [(0, 568), (568, 568), (1136, 568), (1704, 568), (2272, 568), (2840, 567),
                                    (3407, 567), (3974, 567)]
    assert plan_shards(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]
    assert sum(c for _, c in plan_shards(4541, 3)) == 4541
    assert plan_batches(10, 7, 3) == [(10, 3), (13, 3), (16, 1)]
    assert plan_batches(0, 0, 4) == []
    with pytest.raises(ValueError):
        plan_shards(5, 0)


def make_processor():
    """Oracle-backed stand-in for Frontend.stereo_batch: keeps the previous left image like the handle."""
    from oracle import oracle as O
    import spvo_b200.synth as synth
    state = {"prev": None}

    def process(first, count, reset):
        semi, desc = synth.make_stream(count, H, W, seed=4, device="cpu", first_frame=first)
        semi, desc = semi.numpy(), desc.numpy()
        if reset:
            state["prev"] = None
        out = []
        for f in range(count):
            d = O.decode(semi[f], desc[f], max_keypoints=K)
            nl, nr = int(d["n"][0]), int(d["n"][1])
            ms, _ = O.match(d["desc"][0, :nl], d["desc"][1, :nr], mode=1)
            if state["prev"] is not None:
                p = state["prev"]
                mt, _ = O.match(d["desc"][0, :nl], p["desc"][0, : int(p["n"][0])], mode=1)
            else:
                mt = np.zeros(0, O.DMATCH_DTYPE)
            state["prev"] = d
            out.append((first + f, d["kpts"][0, :nl].tobytes(), ms.tobytes(), mt.tobytes()))
        return out
    return process


def test_sharded_equals_sequential_single_process():
    seq = run_shard(make_processor(), 0, NF, batch=3)
    assert [r[0] for r in seq] == list(range(NF)) and len(seq[0][3]) == 0 and len(seq[1][3]) > 0
    for world in (2, 3):
        parts = []
        for rank in range(world):
            parts.append(run_sharded(make_processor(), NF, 2, rank, world))
        flat = [r for p in parts for r in p]
        assert flat == seq
    # without the halo the first frame of every later shard loses its temporal matches
    nohalo = run_sharded(make_processor(), NF, 2, 1, 2, halo=False)
    assert len(nohalo[0][3]) == 0


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(mine):
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        return parts

    res = run_sharded(make_processor(), NF, 2, rank, world, gather=gather)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_gloo_processes():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == run_shard(make_processor(), 0, NF, batch=3)

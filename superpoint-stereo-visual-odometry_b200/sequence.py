"""Host logic for offline sequences: frame-range sharding over ranks and batched streaming.

The front end is frame-independent except for the one-frame dependency of the temporal match
(CURR_LEFT vs PREV_LEFT, feature_detection.hpp:87-90), so a KITTI-length sequence shards into
contiguous frame ranges, one per GPU, with NO data-path collective: each rank
additionally processes the frame just before its range (a one-frame halo) so that its first frame
has its temporal matches.  Only the short keypoint / match lists return to the host, where the
sequential triangulation + PnP/Ceres solve stays (feature_detection_base.cpp:125-399).

`process_batch(first_frame, count, reset) -> list[per-frame result]` is injected: the product passes
a closure over Frontend.stereo_batch; CPU tests pass the oracle.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def plan_shards(num_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous (first, count) ranges, sizes differing by at most one, in rank order."""
    if num_frames < 0 or world_size <= 0:
        raise ValueError("num_frames >= 0 and world_size > 0 required")
    base, rem = divmod(num_frames, world_size)
    out, first = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < rem else 0)
        out.append((first, cnt))
        first += cnt
    return out


def plan_batches(first: int, count: int, batch: int) -> List[Tuple[int, int]]:
    """Split a frame range into consecutive (first, count) batches of at most `batch` frames."""
    if batch <= 0:
        raise ValueError("batch must be positive")
    return [(f, min(batch, first + count - f)) for f in range(first, first + count, batch)]


def run_shard(process_batch: Callable[[int, int, bool], Sequence], first: int, count: int, batch: int,
              halo: bool = True) -> list:
    """Process frames [first, first+count) in batches; returns one result per frame, in order.

    With `halo` and first > 0 the predecessor frame is processed first (result dropped) so the
    temporal match of `first` is available -- the same state a single sequential run would have.
    `reset=True` tells the processor to forget its previous frame (clearLagecyData, BASE:35-66)."""
    results: list = []
    if count == 0:
        return results
    reset = True
    if halo and first > 0:
        process_batch(first - 1, 1, True)
        reset = False
    for f, c in plan_batches(first, count, batch):
        results.extend(process_batch(f, c, reset))
        reset = False
    if len(results) != count:
        raise RuntimeError(f"processor returned {len(results)} results for {count} frames")
    return results


def run_sharded(process_batch: Callable[[int, int, bool], Sequence], num_frames: int, batch: int, rank: int,
                world_size: int, gather: Callable[[list], List[list]] | None = None, halo: bool = True):
    """Each rank runs its shard; `gather` (e.g. a torch.distributed all_gather_object wrapper) returns
    the per-rank lists in rank order.  Returns the whole sequence's results in frame order (on every
    rank that receives the gather), or this rank's shard when gather is None."""
    first, count = plan_shards(num_frames, world_size)[rank]
    mine = run_shard(process_batch, first, count, batch, halo)
    if gather is None:
        return mine
    parts = gather(mine)
    out: list = []
    for p in parts:
        out.extend(p)
    return out

"""Stream -> GPU sharding and the host-side aggregation of track events.

The hot path partitions by stream: frames, ROI masks, motion state, head slices, track tables and
adaptive-FPS counters are all per stream, so ranks never exchange data on the data path.  The one
cross-stream coupling in the reference is the shared ``itertools.count(1)`` of ``IouTracker``
(tracker.py:47): track ids interleave across streams in processing order.  Each rank's kernels
hand out ids from a rank-local counter; ``GlobalIdMap`` turns them into the ids ONE shared tracker
would have produced had it been updated for streams 0..N-1 in canonical order every tick -- a
prefix sum over the per-stream new-track counts, which the ranks all-gather (a few integers per
tick, off the data path).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple


def streams_of_rank(n_streams: int, world: int, rank: int) -> List[int]:
    """Contiguous blocks, stream s -> rank s // ceil(n / world) (32 streams, 8 GPUs: 4 each)."""
    per = -(-n_streams // world)
    return list(range(rank * per, min(n_streams, (rank + 1) * per)))


def rank_of_stream(stream: int, n_streams: int, world: int) -> int:
    return stream // (-(-n_streams // world))


class GlobalIdMap:
    """Maps (stream, rank-local track id) to the id a single shared counter would have assigned.

    Every rank feeds the same ``new_counts`` (all streams, canonical order) each tick, so every rank
    builds the same map without further communication."""

    def __init__(self, n_streams: int, first_id: int = 1):
        self.n_streams = n_streams
        self.next_id = first_id
        self._seen: List[int] = [0] * n_streams  # new tracks seen so far per stream
        # per stream: list of (local ordinal start, count, global start)
        self._runs: List[List[Tuple[int, int, int]]] = [[] for _ in range(n_streams)]

    def advance(self, new_counts: Sequence[int]) -> None:
        """One tick: ``new_counts[s]`` tracks were created on stream s (0 for streams that skipped)."""
        if len(new_counts) != self.n_streams:
            raise ValueError("new_counts must cover every stream")
        for s, n in enumerate(new_counts):
            n = int(n)
            if n:
                self._runs[s].append((self._seen[s], n, self.next_id))
                self._seen[s] += n
                self.next_id += n

    def global_id(self, stream: int, ordinal: int) -> int:
        """``ordinal`` = 0-based creation index of the track within its stream."""
        for start, n, gstart in reversed(self._runs[stream]):
            if start <= ordinal < start + n:
                return gstart + (ordinal - start)
        raise KeyError((stream, ordinal))


def all_gather_new_counts(local_counts: Sequence[int], local_streams: Sequence[int], n_streams: int,
                          group=None) -> List[int]:
    """Every rank contributes its streams' new-track counts; returns the full canonical vector.
    Uses ``torch.distributed`` when initialised (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    full = torch.zeros(n_streams, dtype=torch.int64)
    for s, c in zip(local_streams, local_counts):
        full[s] = int(c)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            full = full.cuda()
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)  # disjoint supports: a sum is a gather
        full = full.cpu()
    return [int(v) for v in full.tolist()]

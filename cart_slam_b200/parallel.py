"""Frame sharding across GPUs: one process per GPU, no collective inside the path, one
final gather of the results.

The superpixel label image is warm-started from the previous frame and re-blocked whenever
`id % reset_iterations == 0` (/root/reference/src/modules/superpixels.cu:93-113), so a sequence can be cut
without changing any result only at frame ids that are multiples of the reset period: every shard then starts
from a fresh block initialisation exactly as the sequential reference would.  With the `histogram_peak`
parameter provider the plane ranges additionally depend on a running histogram that is zeroed when
`id % (update_interval * reset_interval) == 1` (/root/reference/src/modules/planeseg/sp_planeseg.cu:352-388);
`plan_shards(..., hist_period=...)` aligns the cuts to both periods.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gcd
from typing import List, Optional, Sequence


@dataclass(frozen=True)
class Shard:
    rank: int
    first_id: int  # frame id of the first frame (ids start at 1 like SystemRunData::id)
    count: int

    @property
    def frame_slice(self) -> slice:  # 0-based frame indices of a sequence whose first frame has id 1
        return slice(self.first_id - 1, self.first_id - 1 + self.count)


def plan_shards(n_frames: int, world_size: int, reset_iterations: int = 64, hist_period: Optional[int] = None,
                start_id: int = 1) -> List[Shard]:
    """Contiguous frame ranges, one per rank, cut only where the per-sequence state is reset.

    Cuts are placed at ids that are multiples of `reset_iterations` (and, when `hist_period` is given, also
    satisfy id % hist_period == 1 is impossible together with id % reset == 0 unless the periods allow it, so the
    combined period is lcm(reset_iterations, hist_period) and the cut is placed at multiples of it: the frame
    with that id starts a new superpixel chunk; the histogram restarts one frame later, which the runner
    reproduces because it receives the true start id).  Ranks that cannot get a whole chunk receive zero frames.
    """
    if n_frames < 0 or world_size < 1 or reset_iterations < 1:
        raise ValueError("bad arguments")
    period = reset_iterations
    if hist_period:
        period = period * hist_period // gcd(period, hist_period)
    last_id = start_id + n_frames - 1
    # candidate cut ids: multiples of `period` strictly inside (start_id, last_id]
    first_cut = ((start_id // period) + 1) * period
    cuts = list(range(first_cut, last_id + 1, period))
    # choose world_size - 1 cuts closest to an even split
    chosen: List[int] = []
    for r in range(1, world_size):
        target = start_id + (n_frames * r) // world_size
        best = None
        for c in cuts:
            if chosen and c <= chosen[-1]:
                continue
            if best is None or abs(c - target) < abs(best - target):
                best = c
        if best is None:
            break
        chosen.append(best)
    bounds = [start_id] + chosen + [last_id + 1]
    shards = [Shard(r, bounds[r], bounds[r + 1] - bounds[r]) for r in range(len(bounds) - 1)]
    shards += [Shard(r, last_id + 1, 0) for r in range(len(shards), world_size)]
    return shards


def gather_to_rank0(local, shards: Sequence[Shard], group=None):
    """The path's only collective: every rank contributes its shard's result tensor [count, ...]; rank 0 gets the
    concatenation in frame order (None elsewhere).  Works with the nccl (device tensors) and gloo backends."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    tail = tuple(local.shape[1:])
    # equal-sized buffers (ncclGather / gloo gather semantics): pad every shard to the longest one
    longest = max(s.count for s in shards)
    buf = torch.zeros((longest,) + tail, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([out[s.rank][: s.count] for s in shards], dim=0)


def allgather_histograms(local_hist, shards: Sequence[Shard], group=None, device=None):
    """Second collective of the sharded runner (the two-pass histogram_peak scheme): every rank
    contributes the per-frame histograms of its shard ([count, 256] int32, numpy, as phase 1 returns them) and every
    rank gets the whole sequence's [n, 256] table in frame order.  Shards are padded to the longest one (all_gather
    wants equal sizes).  `device`: where the exchange buffers live ("cuda" for nccl, None = CPU for gloo)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    local_hist = np.ascontiguousarray(local_hist, dtype=np.int32)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_hist
    world = dist.get_world_size(group)
    longest = max(s.count for s in shards)
    buf = torch.zeros((longest, 256), dtype=torch.int32, device=device)
    if local_hist.shape[0]:
        buf[: local_hist.shape[0]] = torch.from_numpy(local_hist).to(buf.device)
    out = torch.empty((world * longest, 256), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = out.cpu().numpy().reshape(world, longest, 256)
    return np.concatenate([parts[s.rank, : s.count] for s in shards])

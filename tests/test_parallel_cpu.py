"""Multi-rank plumbing on CPU (gloo, world_size 2): shard planning respects the superpixel reset period and
the final gather returns frames in id order.  The per-shard "result" is produced by the scalar oracle's block
initialisation + a frame-id stamp (the CUDA path has no CPU fallback, so only the host logic is exercised)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cart_slam_b200 as cb
from cart_slam_b200.parallel import Shard, allgather_histograms, gather_to_rank0, plan_shards


def test_plan_cuts_only_at_reset_frames():
    for n, world, reset in [(1000, 2, 64), (1000, 8, 64), (10000, 8, 64), (100, 4, 64), (300, 3, 8), (63, 2, 64)]:
        shards = plan_shards(n, world, reset)
        assert len(shards) == world
        assert sum(s.count for s in shards) == n
        nxt = 1
        for s in shards:
            assert s.first_id == nxt or s.count == 0
            if s.count and s.first_id != 1:
                assert s.first_id % reset == 0, (n, world, reset, s)  # a shard starts on a block re-initialisation
            nxt = s.first_id + s.count if s.count else nxt
    # even split when the period allows it
    sh = plan_shards(10000, 8, 64)
    assert max(s.count for s in sh) - min(s.count for s in sh) <= 2 * 64
    # fewer whole chunks than ranks: trailing ranks stay empty instead of splitting a chunk
    sh = plan_shards(63, 2, 64)
    assert [s.count for s in sh] == [63, 0]


def test_plan_with_histogram_period():
    shards = plan_shards(10000, 2, 64, hist_period=300)
    assert shards[1].first_id % 4800 == 0 and sum(s.count for s in shards) == 10000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, H, W, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = plan_shards(n, world, 8)
        me = shards[rank]
        # stand-in for the shard's plane images: every pixel carries the frame id (mod 251)
        local = torch.empty((me.count, H, W), dtype=torch.uint8)
        for i in range(me.count):
            local[i] = (me.first_id + i) % 251
        full = gather_to_rank0(local, shards)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo(tmp_path):
    n, H, W = 37, 6, 10
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, H, W, out), nprocs=2, join=True)
    full = np.load(out)
    assert full.shape == (n, H, W)
    for i in range(n):
        assert (full[i] == (i + 1) % 251).all()


def _hist_worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = plan_shards(n, world, 8)
        me = shards[rank]
        rng = np.random.default_rng(7)
        all_hist = rng.integers(0, 5000, (n, 256)).astype(np.int32)  # every rank can regenerate the whole table
        full = allgather_histograms(all_hist[me.frame_slice], shards)
        assert np.array_equal(full, all_hist)
        # the two-pass scheme: the schedule over the whole sequence, then the rank's own slice
        opts = cb.SequenceOptions(pipeline=1, provider=1, update_interval=5, reset_interval=2, start_id=1)
        params = cb.sequence_parameters(opts, full)
        np.save(os.path.join(out_dir, f"params_{rank}.npy"), params[me.frame_slice])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_histogram_exchange_and_parameter_schedule_world2_gloo(tmp_path):
    """Phase 1 -> all-gather -> parameter schedule of the sharded runner with gloo on CPU: the ranks' slices concatenate
    to the schedule a single process computes from the whole table."""
    n = 43
    mp.spawn(_hist_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    got = np.concatenate([np.load(str(tmp_path / f"params_{r}.npy")) for r in range(2)])
    rng = np.random.default_rng(7)
    all_hist = rng.integers(0, 5000, (n, 256)).astype(np.int32)
    opts = cb.SequenceOptions(pipeline=1, provider=1, update_interval=5, reset_interval=2, start_id=1)
    assert np.array_equal(got, cb.sequence_parameters(opts, all_hist))

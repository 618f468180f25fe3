"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: reward all-gather + advantage slicing, gradient
averaging, and the per-rank slicing of the RepeatRandomSampler stream."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bioreason_b200 import dp
        from bioreason_b200.trainer.grpo_trainer import RepeatRandomSampler
        from oracle import grpo as og
        G, per_dev, nf = 4, 6, 2                        # per_device % G != 0 -> a group straddles the two ranks
        allr = torch.randn(world * per_dev, nf, generator=torch.Generator().manual_seed(7))
        mine = allr[rank * per_dev:(rank + 1) * per_dev].clone()
        gathered = dp.gather_rewards(mine)
        assert torch.equal(gathered, allr)
        adv = dp.local_slice(og.group_advantages(gathered, G), per_dev)
        want = og.group_advantages(allr, G)[rank * per_dev:(rank + 1) * per_dev]
        assert torch.equal(adv, want)
        g1 = torch.full((10,), float(rank + 1)); g2 = torch.arange(4.0) * (rank + 1)
        dp.allreduce_mean_([g1, g2])
        assert torch.allclose(g1, torch.full((10,), 1.5)) and torch.allclose(g2, torch.arange(4.0) * 1.5)
        # overlapped C2: slices reduced out of order while "the backward" goes on + the uncovered remainder + extra buffers == one all-reduce
        flat = torch.arange(40.0) * (rank + 1); extra = torch.ones(3) * (rank + 1)
        red = dp.OverlappedGradReduce(flat)
        red.reduce_slice(30, 40); red.reduce_slice(10, 20)
        red.finish([extra])
        assert torch.allclose(flat, torch.arange(40.0) * 1.5) and torch.allclose(extra, torch.full((3,), 1.5))
        # sampler: same seed on every rank, contiguous per-rank slices, consecutive G rows share a prompt index
        s = list(iter(RepeatRandomSampler(range(12), G, (per_dev * world) // G, 1, seed=3)))
        mine_idx = list(dp.rank_batches(s, per_dev))
        ret[rank] = (mine_idx, s)
    finally:
        dist.destroy_process_group()


def test_dp_two_ranks_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (b0, s0), (b1, s1) = ret[0], ret[1]
    assert s0 == s1                                              # identical global stream on both ranks
    for x, y in zip(b0, b1):
        glob = x + y
        assert all(len(set(glob[i:i + 4])) == 1 for i in range(0, len(glob), 4))     # groups of G intact in the global batch
        assert len(x) == 6 and len(y) == 6


def test_single_process_is_identity():
    from bioreason_b200 import dp
    r = torch.randn(8, 3)
    assert dp.gather_rewards(r) is r and torch.equal(dp.local_slice(r, 8), r)
    g = torch.ones(5)
    dp.allreduce_mean_([g])
    assert torch.equal(g, torch.ones(5))
    assert list(dp.rank_batches(list(range(10)), 4)) == [[0, 1, 2, 3], [4, 5, 6, 7]]

"""Sharding by frame and the result all-gather on 2 CPU ranks (gloo)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cvpack_b200 import distributed


def _free_port() -> int:
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _worker(rank: int, world: int, port: int, num_frames: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = distributed.local_shard(num_frames)
        num_atoms = 7
        frames = torch.arange(begin, end, dtype=torch.float32)
        values = frames * 2.0  # stands for the CV value of each owned frame
        gathered = distributed.all_gather_frames(values, num_frames)
        assert torch.equal(gathered, torch.arange(num_frames, dtype=torch.float32) * 2.0)
        active = torch.tensor([1, 4, 5])
        grad = torch.zeros((end - begin, num_atoms, 3))
        grad[:, active] = frames[:, None, None] + torch.arange(3.0)[None, None, :]
        dense = distributed.all_gather_gradients(grad, num_frames)
        compact = distributed.all_gather_gradients(grad, num_frames, active_atoms=active)
        assert torch.equal(dense, compact) and dense.shape == (num_frames, num_atoms, 3)
        assert torch.equal(dense[:, 4, 2], torch.arange(num_frames, dtype=torch.float32) + 2.0)
        assert torch.count_nonzero(dense[:, 0]) == 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [8, 9])
def test_two_rank_gather(num_frames):
    mp.spawn(_worker, args=(2, _free_port(), num_frames), nprocs=2, join=True)


def test_single_process_is_identity():
    values = torch.arange(5.0)
    assert distributed.local_shard(5) == (0, 5)
    assert distributed.all_gather_frames(values, 5) is values

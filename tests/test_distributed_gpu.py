"""Two ranks on two GPUs (run with `gpurun --gpus 2`): frames sharded by rank, the gradient of a real
collective variable gathered over NCCL (dense and compact) and through the library's peer-memory
push kernel -- all three must agree bit for bit with a single-GPU evaluation of the whole batch."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)"),
]


def _free_port() -> int:
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _worker(rank: int, world: int, port: int, num_frames: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        import cvpack_b200 as cvpack
        from cvpack_b200 import distributed
        from tests import helpers

        rng = np.random.default_rng(81)  # the same batch on every rank
        n = 600
        box = np.diag([2.6, 2.6, 2.6])
        system = helpers.make_system(n, rng)
        x = rng.uniform(0, 2.6, size=(num_frames, n, 3))
        nb = helpers.make_nonbonded(n, rng, periodic=True)
        cv = cvpack.NumberOfContacts(range(0, 150), range(150, 320), nb)
        cv.addToSystem(system)
        begin, end = distributed.local_shard(num_frames)
        ctx = cvpack.BatchContext(system, x[begin:end], box, device=device)
        value, grad = cv.getValueAndGradient(ctx)
        whole = cvpack.BatchContext(system, x, box, device=device)
        ref_value, ref_grad = cv.getValueAndGradient(whole)

        values = distributed.all_gather_frames(value._value, num_frames)
        assert torch.equal(values, ref_value._value)
        dense = distributed.all_gather_gradients(grad, num_frames)
        assert torch.equal(dense, ref_grad)
        active = torch.as_tensor(ctx._native_cv(cv).active_atoms(), device=device)
        assert active.numel() == 320
        compact = distributed.all_gather_gradients(grad, num_frames, active_atoms=active)
        assert torch.equal(compact, ref_grad)

        # the library's own collective: push into every rank's buffer, stream-ordered barrier
        peer = distributed.PeerGather(num_frames, n, torch.float32, device, depth=2)
        for use_active in (False, True, False):
            half = (end - begin) // 2
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):  # two slabs, pushed on a side stream
                peer.push(grad[:half], begin, active.int() if use_active else None)
                peer.push(grad[half:], begin + half, active.int() if use_active else None)
                gathered = peer.finish()
            torch.cuda.current_stream(device).wait_stream(side)
            assert torch.equal(gathered, ref_grad), f"rank {rank}: peer gather differs (active={use_active})"
        # an atom list with many short runs takes the scatter kernel instead of the row copies
        scattered = torch.arange(1, n, 3, device=device, dtype=torch.int32)
        sparse = torch.zeros_like(grad)
        sparse[:, scattered.long()] = grad[:, : scattered.numel()]
        expected = distributed.all_gather_gradients(sparse, num_frames)
        peer.tensor().zero_()
        torch.cuda.synchronize()
        dist.barrier()
        peer.push(sparse, begin, scattered)
        gathered = peer.finish()
        torch.cuda.synchronize()
        assert torch.equal(gathered, expected), f"rank {rank}: scattered peer gather differs"
        assert not peer.timed_out()
        dist.barrier()
        peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [12, 13])
def test_two_gpu_gradient_gather(num_frames):
    mp.spawn(_worker, args=(2, _free_port(), num_frames), nprocs=2, join=True)

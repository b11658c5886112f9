"""Frame sharding host logic under torch.distributed (gloo, world_size 2, CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maua_b200 import parallel


def test_frame_range_partitions():
    for T in (0, 1, 7, 720, 10800, 10801):
        for world in (1, 2, 3, 8):
            rs = [parallel.frame_range(r, world, T) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == T
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in rs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, T):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        lin = torch.nn.Linear(4, 3)
        parallel.broadcast_module_(lin, src=0)
        torch.manual_seed(0)
        ref = torch.nn.Linear(4, 3)
        assert torch.equal(lin.weight, ref.weight) and torch.equal(lin.bias, ref.bias)
        full = torch.arange(T * 6, dtype=torch.float32).reshape(T, 2, 3) if rank == 0 else torch.zeros(T, 2, 3)
        mine, (s, e) = parallel.shard_inputs({"latents": full}, src=0)
        assert (s, e) == parallel.frame_range(rank, world, T)
        assert torch.equal(mine["latents"], torch.arange(T * 6, dtype=torch.float32).reshape(T, 2, 3)[s:e])
        frames = (mine["latents"][:, :1, :1] % 251).to(torch.uint8).expand(-1, 4, 5).contiguous()  # fake rendered frames
        allf = parallel.gather_frames(frames, T, dst=0)
        if rank == 0:
            want = (torch.arange(T * 6, dtype=torch.float32).reshape(T, 2, 3)[:, :1, :1] % 251).to(torch.uint8).expand(-1, 4, 5)
            assert torch.equal(allf, want)
        else:
            assert allf is None
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 9), nprocs=2, join=True)

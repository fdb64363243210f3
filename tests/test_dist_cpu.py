"""CPU, world_size 2 over gloo: the batch-sharded gradient all-reduce (the path's only collective) and batch sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avmoe_b200.dist import GradientAllReducer, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    frozen = torch.nn.Linear(3, 3)
    for p in frozen.parameters():
        p.requires_grad_(False)                # frozen backbone parameters never enter the buckets
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    s, e = shard_batch(8, rank, world)
    loss = frozen(model(data[s:e])).pow(2).sum()   # per-rank shard; sum loss -> averaged grads = full-batch grads / world
    loss.backward()
    red = GradientAllReducer(list(model.parameters()) + list(frozen.parameters()), bucket_bytes=64)  # tiny buckets: several of them
    assert len(red.buckets) > 1
    red.reduce()
    torch.save([p.grad.clone() for p in model.parameters()], os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(tmp_path / "g0.pt")
    g1 = torch.load(tmp_path / "g1.pt")
    for a, b in zip(g0, g1):
        assert torch.equal(a, b), "ranks must hold identical averaged gradients"
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    frozen = torch.nn.Linear(3, 3)
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    frozen(model(data)).pow(2).sum().backward()
    for a, p in zip(g0, model.parameters()):
        assert torch.allclose(a, p.grad / world, atol=1e-6), "all-reduce(avg) of shard grads == full-batch grads / world"


def test_shard_batch_covers_everything_once():
    for total in (1, 7, 16, 160):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_reducer_is_a_noop():
    m = torch.nn.Linear(3, 2)
    m(torch.ones(1, 3)).sum().backward()
    g = m.weight.grad.clone()
    red = GradientAllReducer(m.parameters())
    assert red.reduce() == [] and torch.equal(m.weight.grad, g)

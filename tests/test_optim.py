"""Flat parameter storage, the fused Adam step (vs torch.optim.Adam, the reference's optimizer: AVMOE/AVE/main_trans_v3.py:311-323)
and the in-place flat gradient all-reduce (gloo, world size 2, CPU)."""
import os

import pytest
import torch
import torch.nn as nn


def _toy(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(13, 7), nn.ReLU(), nn.Linear(7, 5, bias=False), nn.LayerNorm(5), nn.Linear(5, 3))


def test_flat_parameters_are_views_and_gather_grads():
    from avmoe_b200.optim import FlatParameters
    m = _toy()
    ref = [p.detach().clone() for p in m.parameters()]
    flat = FlatParameters(m.parameters())
    assert flat.check_views()
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p.detach(), r)
    m(torch.randn(4, 13)).sum().backward()
    grads = [p.grad.clone() for p in m.parameters()]
    g = flat.gather_grads()
    assert all(p.grad is None for p in m.parameters())
    for gv, off, r in zip(grads, flat.offsets, grads):
        assert torch.equal(g[off:off + r.numel()].view_as(r), r)
    # writing the flat buffer updates the module
    flat.flat.add_(1.0)
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p.detach(), r + 1.0)
    assert flat.numel % 4 == 0 and all(o % 4 == 0 for o in flat.offsets)


def test_fused_adam_refuses_cpu():
    from avmoe_b200.optim import FlatParameters, FusedAdam
    flat = FlatParameters(_toy().parameters())
    with pytest.raises(RuntimeError):
        FusedAdam(flat).step()


@pytest.mark.gpu
@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(weight_decay):
    from avmoe_b200.optim import FlatParameters, FusedAdam
    m1, m2 = _toy(1).cuda(), _toy(1).cuda()
    lr_a, lr_b = 5e-4, 2e-3
    p1 = list(m1.parameters())
    opt_ref = torch.optim.Adam([{"params": p1[:3], "lr": lr_a}, {"params": p1[3:], "lr": lr_b}], weight_decay=weight_decay)
    sched_ref = torch.optim.lr_scheduler.StepLR(opt_ref, step_size=2, gamma=0.5)
    flat = FlatParameters(m2.parameters())
    p2 = list(m2.parameters())
    opt = FusedAdam(flat, weight_decay=weight_decay, groups=[{"params": p2[:3], "lr": lr_a}, {"params": p2[3:], "lr": lr_b}])
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5) if False else None   # FusedAdam exposes param_groups['lr']
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(5):
        x = torch.randn(16, 13, device="cuda", generator=g)
        for mod in (m1, m2):
            mod.zero_grad(set_to_none=True)
            (mod(x) ** 2).sum().backward()
        opt_ref.step()
        sched_ref.step()
        flat.gather_grads()
        opt.step()
        if (step + 1) % 2 == 0:     # StepLR(step_size=2, gamma=0.5) by hand on the exposed lr
            for grp in opt.param_groups:
                grp["lr"] *= 0.5
        for a, b in zip(m1.parameters(), m2.parameters()):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (step, (a - b).abs().max().item())


@pytest.mark.gpu
def test_fused_adam_grad_scale_and_zero():
    from avmoe_b200.optim import FlatParameters, FusedAdam
    m1, m2 = _toy(2).cuda(), _toy(2).cuda()
    opt_ref = torch.optim.Adam(m1.parameters(), lr=1e-3)
    flat = FlatParameters(m2.parameters())
    opt = FusedAdam(flat, lr=1e-3)
    x = torch.randn(8, 13, device="cuda")
    m1(x).sum().backward()
    (4.0 * m2(x).sum()).backward()              # e.g. the sum over 4 ranks
    opt_ref.step()
    flat.gather_grads()
    opt.step(grad_scale=0.25, zero_grad=True)   # ... averaged inside the update
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    assert flat.grad.abs().max().item() == 0


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avmoe_b200.dist import FlatGradientReducer
    from avmoe_b200.optim import FlatParameters
    m = _toy(5)
    flat = FlatParameters(m.parameters())
    red = FlatGradientReducer(flat, bucket_bytes=64)      # many tiny buckets
    torch.manual_seed(100 + rank)
    x = torch.randn(6, 13)
    m(x).sum().backward()
    local = torch.cat([torch.nn.functional.pad(p.grad.flatten(), (0, (-p.numel()) % 4)) for p in m.parameters()])
    red.reduce()
    q.put((rank, local, flat.grad.clone(), red.grad_scale))
    dist.destroy_process_group()


def test_flat_gradient_reducer_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    total = res[0][1] + res[1][1]
    for rank, local, reduced, scale in res:
        assert torch.allclose(reduced, total, atol=1e-6)
        assert scale == 0.5

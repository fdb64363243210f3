"""AdapterStack (the bench workload) on the GPU: the two-stream schedule of a block's audio/visual adapters gives the same loss
and gradients as the sequential schedule, and a CUDA-graph replay of the step reproduces the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _step(model, tok):
    for p in model.parameters():
        p.grad = None
    leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tok.items()}
    loss = model(leaves)
    loss.backward()
    return loss, leaves


def _flat_grads(model):
    return torch.cat([p.grad.flatten().float() for p in model.parameters() if p.grad is not None])


def test_two_stream_schedule_matches_sequential():
    from avmoe_b200 import stack as S
    outs = []
    for parallel in (False, True):
        model = S.AdapterStack("ave", seed=0).cuda().train()
        model.parallel_modalities = parallel
        tok = S.make_stage_tokens("ave", 6, torch.bfloat16, torch.device("cuda"), seed=3)
        loss, leaves = _step(model, tok)
        torch.cuda.synchronize()
        gx = torch.cat([t.grad.flatten().float() for pair in leaves.values() for t in pair])
        outs.append((loss.item(), _flat_grads(model), gx))
    (l0, g0, x0), (l1, g1, x1) = outs
    assert abs(l0 - l1) <= 1e-4 * abs(l0)
    assert ((g0 - g1).norm() / g0.norm()).item() < 1e-4          # fp32 atomics reorder, nothing else
    assert ((x0 - x1).norm() / x0.norm()).item() < 1e-4


def test_cuda_graph_replay_matches_eager_step():
    from avmoe_b200 import stack as S
    model = S.AdapterStack("ave", seed=0).cuda().train()
    tok = S.make_stage_tokens("ave", 4, torch.bfloat16, torch.device("cuda"), seed=5)
    for _ in range(2):
        loss_e, _ = _step(model, tok)
    torch.cuda.synchronize()
    ref_loss, ref_g = loss_e.item(), _flat_grads(model).clone()
    del loss_e          # a live autograd graph would pin AccumulateGrad nodes to the (non-capturable) default stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _ = _step(model, tok)
        del _
    torch.cuda.current_stream().wait_stream(side)
    for p in model.parameters():
        p.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss_g, _ = _step(model, tok)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert abs(loss_g.item() - ref_loss) <= 1e-4 * abs(ref_loss)
    assert ((_flat_grads(model) - ref_g).norm() / ref_g.norm()).item() < 1e-4

"""AdapterStack with the audio/visual adapters on two streams vs one stream: same loss and gradients (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import stack as S

def run(parallel):
    torch.manual_seed(0)
    model = S.AdapterStack("ave", seed=0).cuda().train()
    model.parallel_modalities = parallel
    tok = S.make_stage_tokens("ave", 20, torch.bfloat16, torch.device("cuda"), seed=3)
    leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tok.items()}
    loss = model(leaves)
    loss.backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten().float() for p in model.parameters() if p.grad is not None])
    gx = torch.cat([t.grad.flatten().float() for pair in leaves.values() for t in pair])
    return loss.item(), g, gx

l0, g0, x0 = run(False)
l1, g1, x1 = run(True)
print("loss", l0, l1)
print("param grad rel diff", ((g0 - g1).norm() / g0.norm()).item(), "input grad rel diff", ((x0 - x1).norm() / x0.norm()).item())
assert abs(l0 - l1) <= 1e-3 * abs(l0) and ((g0 - g1).norm() / g0.norm()).item() < 1e-3 and ((x0 - x1).norm() / x0.norm()).item() < 1e-3
print("parallel == sequential: ok")

"""Diagnostic: which Python lines launch the tiny torch kernels of one adapter call (GPU box only)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avmoe_b200 import stack as S

def main():
    stage, modality, BT = int(sys.argv[1]), sys.argv[2], 160
    torch.manual_seed(0)
    mod = S.build_adapter("ave", stage, modality).cuda().train()
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    x = torch.randn(BT, N, C, device="cuda").bfloat16().requires_grad_()
    o = torch.randn(BT, Nx, Cx, device="cuda").bfloat16().requires_grad_()
    def call():
        res = mod(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
        res[0].backward(torch.ones_like(res[0]))
    for _ in range(2): call()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        call()
        torch.cuda.synchronize()
    agg = collections.Counter()
    names = ("aten::zeros", "aten::fill_", "aten::zero_", "aten::copy_", "aten::add_", "aten::mul", "aten::add", "aten::neg", "aten::sum",
             "aten::_to_copy", "aten::index_put_", "aten::clone", "aten::cat", "aten::stack", "aten::_foreach_add_", "aten::mul_", "aten::sub", "aten::div")
    for ev in prof.key_averages(group_by_stack_n=12):
        if ev.key in names:
            fr = [f for f in (ev.stack or []) if "/avmoe_b200/" in f or "/tools/" in f]
            agg[(ev.key, fr[0] if fr else (ev.stack[0] if ev.stack else "?"))] += ev.count
    for (name, where), n in agg.most_common(70):
        print(f"{n:4d} {name:22s} {where[-100:]}")

if __name__ == "__main__":
    main()

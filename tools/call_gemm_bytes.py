"""Algorithmic bytes / flops of the GEMM launches of ONE adapter call (same command line as tools/profile_one.py), to put next
to the DRAM bytes ncu measured for the same call (GPU box only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops, stack as S

variant, stage, modality, BT = "ave", 1, "visual", 160
torch.manual_seed(0)
mod = S.build_adapter(variant, stage, modality).cuda().train()
with torch.no_grad():
    for n, p in mod.named_parameters():
        if n.endswith("gate") or n.endswith("gate_av") or n.endswith("gate_self"):
            p.fill_(0.5)
Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
C, N, Cx, Nx = Cv, Nv, Ca, Na
x = (torch.randn(BT, N, C, device="cuda") + torch.randn(BT, 1, C, device="cuda")).bfloat16().requires_grad_()
o = (torch.randn(BT, Nx, Cx, device="cuda") + torch.randn(BT, 1, Cx, device="cuda")).bfloat16().requires_grad_()
for it in range(2):
    if it == 1:
        ops.GEMM_PROFILE = []
    res = mod(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
    res[0].float().sum().backward()
torch.cuda.synchronize()
rows = [dict(tag=t, ms=a.elapsed_time(b), bytes=nb, flops=fl) for a, b, fl, nb, t in ops.GEMM_PROFILE]
json.dump(rows, open("gpurun_out/call_gemm_bytes.json", "w"), indent=1)
print(len(rows), "gemm launches;", sum(r["bytes"] for r in rows) / 1e6, "MB algorithmic")

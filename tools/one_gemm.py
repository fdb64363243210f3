"""One write-heavy skinny GEMM (expert up projection, stage-1 visual shape) for ncu source-level captures (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops
rows, C, Crp = 368640, 192, 24
a = torch.randn(rows, Crp, device="cuda").bfloat16()
w = torch.randn(C, Crp, device="cuda").bfloat16()
yd = torch.empty(rows, 2 * C, device="cuda", dtype=torch.bfloat16)
cb = torch.randn(C, device="cuda")
for _ in range(3):
    ops.gemm(a, w, out=yd[:, :C], col_bias=cb)
torch.cuda.synchronize()
print("done")

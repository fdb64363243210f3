"""GEMM diagnostics on the shapes of the GEMM-based expert path (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops
from tools.gemm_bench2 import timeit


def main():
    bf = torch.bfloat16
    r = lambda *s: torch.randn(*s, device="cuda").to(bf)
    f = lambda *s: torch.randn(*s, device="cuda")
    for rows, C in ((368640, 192), (655360, 96), (23040, 768), (40960, 384), (5760, 1536)):
        Cr = C // 8
        Crp = (Cr + 7) // 8 * 8
        S = rows * C * 2
        xl, WdD, WuD = r(rows, C), r(Cr, C), r(C, Crp)
        z = torch.empty(rows, 2 * Cr, device="cuda")
        a = r(rows, Crp)
        yd = r(rows, 2 * C)
        cs, cb = f(C), f(C)
        W2 = r(Crp, 2 * C)
        dz = r(rows, Crp)
        print(f"--- rows={rows} C={C} Cr={Cr} Crp={Crp}  S={S/1e6:.0f} MB")
        def rep(name, fn, nbytes, ref=None):
            t = timeit(fn)
            extra = ""
            if ref is not None:
                extra = f"   torch {timeit(ref):8.1f} us"
            print(f"  {name:34s} {t:8.1f} us  {nbytes/t/1e3:7.1f} GB/s{extra}", flush=True)
        rep("down  xl·Wdᵀ -> z view", lambda: ops.gemm(xl, WdD, out=z[:, :Cr]), S + rows * Cr * 4, lambda: xl @ WdD.t())
        rep("up    a·Wuᵀ*s+b -> yd view", lambda: ops.gemm(a, WuD, out=yd[:, :C], col_scale=cs, col_bias=cb), S + rows * Crp * 2, lambda: a @ WuD.t())
        rep("da    yd·W2ᵀ", lambda: ops.gemm(yd, W2, out_dtype=torch.float32), 2 * S + rows * Crp * 4, lambda: yd @ W2.t())
        rep("dxl   dz·WdTᵀ", lambda: ops.gemm(dz, WuD), S + rows * Crp * 2, lambda: dz @ WuD.t())
        sk = ops.split_k_hint(rows, 2 * C, Crp)
        rep(f"dWu   ydᵀ·a  split_k={sk}", lambda: ops.gemm(yd, a, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=sk), 2 * S + rows * Crp * 2,
            lambda: yd.t() @ a)
        sk = ops.split_k_hint(rows, Crp, C)
        rep(f"dWd   dzᵀ·xl split_k={sk}", lambda: ops.gemm(dz, xl, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=sk), S + rows * Crp * 2,
            lambda: dz.t() @ xl)


if __name__ == "__main__":
    main()

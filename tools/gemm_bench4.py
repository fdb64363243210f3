"""GEMM diagnostics: write-heavy skinny shapes, time vs N and K (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops
from tools.gemm_bench2 import timeit

bf = torch.bfloat16
rows = 368640
for K in (24, 64, 192):
    a = torch.randn(rows, K, device="cuda").to(bf)
    for N in (32, 64, 128, 192, 256, 384):
        w = torch.randn(N, K, device="cuda").to(bf)
        out = torch.empty(rows, N, device="cuda", dtype=bf)
        t = timeit(lambda: ops.gemm(a, w, out=out))
        tt = timeit(lambda: torch.matmul(a, w.t(), out=out))
        nbytes = rows * (K + N) * 2
        print(f"K={K:4d} N={N:4d}  ours {t:7.1f} us {nbytes/t/1e3:7.0f} GB/s   cublas {tt:7.1f} us   tiles/SM {rows/128/148*((N+255)//256):.1f}", flush=True)

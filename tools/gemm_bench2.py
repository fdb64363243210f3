"""GEMM diagnostics on the shapes the adapter stack actually launches (GPU box only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3   # us

def main():
    bf = torch.bfloat16
    dev = "cuda"
    r = lambda *s: torch.randn(*s, device=dev).to(bf)
    f = lambda *s: torch.randn(*s, device=dev)
    rows = 368640
    O, Wfc = r(rows, 192), r(96, 192)
    t = timeit(lambda: ops.gemm(O, Wfc)); print(f"fc   [368640,192]x[96,192]^T      {t:8.1f} us  {(rows*192*2+rows*96*2)/t/1e3:7.1f} GB/s")
    t = timeit(lambda: O @ Wfc.t()); print(f"   torch                          {t:8.1f} us")
    dP, WfT = r(rows, 96), r(192, 96)
    t = timeit(lambda: ops.gemm(dP, WfT)); print(f"dO   [368640,96]x[192,96]^T      {t:8.1f} us  {(rows*96*2+rows*192*2)/t/1e3:7.1f} GB/s")
    t = timeit(lambda: dP @ WfT.t()); print(f"   torch                          {t:8.1f} us")
    W, Pt = r(4096, 2304), r(160 * 96, 2304)
    fl = 2.0 * 4096 * 2304 * 15360
    t = timeit(lambda: ops.gemm(W, Pt)); print(f"big  plain                        {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    t = timeit(lambda: ops.gemm(W, Pt, d_nbatch=96)); print(f"big  d_nbatch=96                  {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    rb, cs, cb = f(4096), f(96), f(96)
    t = timeit(lambda: ops.gemm(W, Pt, d_nbatch=96, row_bias=rb, col_scale=cs, col_bias=cb)); print(f"big  d_nbatch=96 + biases         {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    t = timeit(lambda: ops.gemm(W, Pt, row_bias=rb)); print(f"big  plain + row bias             {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    dV, P = r(160, 256, 384), r(160, 144, 384)
    fl = 2.0 * 256 * 144 * 384 * 160
    for sk in (1, 16, 64, 148):
        t = timeit(lambda: ops.gemm(dV, P, out_dtype=torch.float32, split_k=sk)); print(f"wgrad3 [256x144] K=61440 split_k={sk:3d} {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    dV1, P1 = r(160, 4096, 96), r(160, 2304, 96)
    fl = 2.0 * 4096 * 2304 * 96 * 160
    t = timeit(lambda: ops.gemm(dV1, P1, out_dtype=torch.float32)); print(f"wgrad1 [4096x2304] K=15360         {t:8.1f} us  {fl/t/1e6:7.1f} TF/s")
    # stage-3 forward-type: W [256,144] x Pt [160*384,144]
    W3, Pt3 = r(256, 144), r(160 * 384, 144)
    fl = 2.0 * 256 * 144 * 384 * 160
    t = timeit(lambda: ops.gemm(W3, Pt3, d_nbatch=384)); print(f"fwd3 [256,144]x[61440,144]         {t:8.1f} us  {fl/t/1e6:7.1f} TF/s  {(61440*144*2+160*256*384*2)/t/1e3:7.1f} GB/s")
    t = timeit(lambda: torch.matmul(W3, Pt3.t())); print(f"   torch                          {t:8.1f} us")

if __name__ == "__main__":
    main()

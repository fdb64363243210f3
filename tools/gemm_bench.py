"""Micro-benchmark of the tcgen05 GEMM on the resampler shapes (GPU box only; diagnostic, not the bench)."""
import sys, os, json
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
    return s.elapsed_time(e) / iters

def main():
    BT = int(os.environ.get("BT", 160))
    res = []
    for dtype in (torch.bfloat16, torch.float32):
        for (N, Nx, Cm) in ((4096, 2304, 96), (4096, 2304, 128), (2304, 4096, 96), (256, 144, 384), (144, 256, 384)):
            W = torch.randn(N, Nx, device="cuda").to(dtype)
            P = torch.randn(BT, Nx, Cm, device="cuda").to(dtype)
            dV = torch.randn(BT, N, Cm, device="cuda").to(dtype)
            fl = 2.0 * N * Nx * Cm * BT
            t_f = timeit(lambda: ops.gemm(W, P, b_kmajor=False))
            t_d = timeit(lambda: ops.gemm(W, dV, a_kmajor=False, b_kmajor=False))
            t_w = timeit(lambda: ops.gemm(dV, P, out_dtype=torch.float32))
            t_ref = timeit(lambda: torch.matmul(W, P))
            # same FLOPs with B K-major ([b*C, Nx] rows): the classic TN form
            Pt = torch.randn(BT * Cm, Nx, device="cuda").to(dtype)
            t_kk = timeit(lambda: ops.gemm(W, Pt))
            res.append(dict(dtype=str(dtype), N=N, Nx=Nx, C=Cm, BT=BT, fwd_ms=t_f, dgrad_ms=t_d, wgrad_ms=t_w, torch_fwd_ms=t_ref, kk_ms=t_kk,
                            fwd_tflops=fl / t_f / 1e9, dgrad_tflops=fl / t_d / 1e9, wgrad_tflops=fl / t_w / 1e9, torch_tflops=fl / t_ref / 1e9,
                            kk_tflops=fl / t_kk / 1e9))
            print(json.dumps(res[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)

if __name__ == "__main__":
    main()

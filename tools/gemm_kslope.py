"""GEMM diagnostics: fixed cost and per-k-iteration cost of the tcgen05 kernel (GPU box only).
One tile per CTA (M = 128 * SMs), time vs K for several tile widths; timed as a CUDA graph of 20 back-to-back launches so the
host enqueue cost (ctypes + tensor-map encode) is not in the number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops

def graph_time(fn, reps=20, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3

def main():
    bf = torch.bfloat16
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for tiles_per_cta in (1, 2):
        M = 128 * sms * tiles_per_cta
        for N in (64, 128, 256):
            row = []
            for K in (64, 256, 768, 1536, 3072):
                a = torch.randn(M, K, device="cuda").to(bf)
                w = torch.randn(N, K, device="cuda").to(bf)
                out = torch.empty(M, N, device="cuda", dtype=bf)
                t = graph_time(lambda: ops.gemm(a, w, out=out))
                row.append(f"K={K}: {t:6.1f}us")
            print(f"tiles/CTA={tiles_per_cta} M={M} N={N:3d}  " + "  ".join(row), flush=True)
    # the shapes of a stage-3 call (warm L2), graph-timed
    def r(*s): return torch.randn(*s, device="cuda").to(bf)
    shapes = [("M23040_N64_K768", lambda: (r(23040, 768), r(64, 768), {})),
              ("M23040_N768_K384", lambda: (r(23040, 384), r(768, 384), {})),
              ("M40960_N64_K384", lambda: (r(40960, 384), r(64, 384), {}))]
    for name, mk in shapes:
        a, w, kw = mk()
        out = torch.empty(a.shape[0], w.shape[0], device="cuda", dtype=bf)
        t = graph_time(lambda: ops.gemm(a, w, out=out, **kw))
        tt = graph_time(lambda: torch.matmul(a, w.t(), out=out))
        print(f"{name}: ours {t:6.1f} us   cublas {tt:6.1f} us", flush=True)
    A = r(160, 144, 768); Bm = r(160, 64, 768)
    t = graph_time(lambda: ops.bgemm(A, Bm, out_dtype=torch.float32))
    tt = graph_time(lambda: torch.bmm(A, Bm.transpose(1, 2)))
    print(f"b160_M144_N64_K768: ours {t:6.1f} us   cublas(bf16 out) {tt:6.1f} us")
    A = r(160, 256, 32); Bm = r(160, 32, 384)
    t = graph_time(lambda: ops.bgemm(A, Bm, b_kmajor=False))
    tt = graph_time(lambda: torch.bmm(A, Bm))
    print(f"b160_M256_N384_K32: ours {t:6.1f} us   cublas {tt:6.1f} us")
    # empty-ish launch: the fixed cost
    a, w = r(128, 64), r(32, 64)
    out = torch.empty(128, 32, device="cuda", dtype=bf)
    print(f"one tile M128 N32 K64: {graph_time(lambda: ops.gemm(a, w, out=out)):6.1f} us")

if __name__ == "__main__":
    main()

"""Per-kernel-category CUDA-event timing of one MoEAdapter fwd+bwd (diagnostic).  usage: time_kernels.py variant stage modality BT"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import stack as S, ops, _lib

def main():
    variant, stage, modality, BT = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    mod = S.build_adapter(variant, stage, modality).cuda().train()
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("gate") or n.endswith("gate_av") or n.endswith("gate_self"):
                p.fill_(0.5)
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    x = (torch.randn(BT, N, C, device="cuda") + torch.randn(BT, 1, C, device="cuda")).bfloat16().requires_grad_()
    o = (torch.randn(BT, Nx, Cx, device="cuda") + torch.randn(BT, 1, Cx, device="cuda")).bfloat16().requires_grad_()
    # wrap every C-ABI call with events
    times = collections.defaultdict(list)
    orig_call, orig_check = ops._lib_call, None
    def timed_call(name, *args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig_call(name, *args, **kw); e1.record()
        times[name].append((e0, e1))
    ops._lib_call = timed_call
    ops.GEMM_PROFILE = []
    for it in range(3):
        if it == 2:
            times.clear(); ops.GEMM_PROFILE = []
        res = mod(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
        res[0].float().sum().backward()
    torch.cuda.synchronize()
    tot = 0.0
    for k, v in sorted(times.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
        t = sum(a.elapsed_time(b) for a, b in v); tot += t
        print(f"{k:26s} n={len(v):3d} {t*1e3:9.1f} us")
    g = sum(a.elapsed_time(b) for a, b, *_ in ops.GEMM_PROFILE); tot += g
    for a, b, fl, nb, tag in ops.GEMM_PROFILE:
        print(f"   gemm {tag:28s} {a.elapsed_time(b)*1e3:8.1f} us {fl/a.elapsed_time(b)/1e9:8.1f} TF/s")
    print(f"gemm total n={len(ops.GEMM_PROFILE)} {g*1e3:.1f} us;  all kernels {tot*1e3:.1f} us  ({variant} stage {stage} {modality} BT={BT} skip={os.environ.get('AVMOE_DEBUG_SKIP','0')})")

if __name__ == "__main__":
    main()

"""Diagnostic: which phase of a per-module CUDA-graph capture fails (GPU box only)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import stack as S

def mk(stage=3, modality="audio", BT=4, dtype=torch.bfloat16):
    torch.manual_seed(0)
    mod = S.build_adapter("ave", stage, modality).cuda().train()
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("gate") or n.endswith("gate_av"):
                p.fill_(0.5)
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    x = (torch.randn(BT, N, C, device="cuda") + torch.randn(BT, 1, C, device="cuda")).to(dtype)
    o = (torch.randn(BT, Nx, Cx, device="cuda") + torch.randn(BT, 1, Cx, device="cuda")).to(dtype)
    return mod, x, o

def call(mod, x, o):
    out, idx = mod(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
    return out

def attempt(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
        print(f"[{name}] OK", flush=True)
    except Exception as ex:
        print(f"[{name}] FAILED: {str(ex)[:300]}", flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("   sync after failure:", str(e2)[:200])

def warm(mod, x, o, grad=True):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            xx, oo = x.detach().requires_grad_(grad), o.detach().requires_grad_(grad)
            out = call(mod, xx, oo)
            if grad:
                out.float().sum().backward()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()

def exp_a():
    mod, x, o = mk()
    with torch.no_grad():
        warm(mod, x, o, grad=False)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = call(mod, x, o)
    g.replay()

def exp_b():
    mod, x, o = mk()
    warm(mod, x, o)
    g = torch.cuda.CUDAGraph()
    xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
    with torch.cuda.graph(g):
        out = call(mod, xx, oo)
    g.replay()

def exp_c():
    mod, x, o = mk()
    warm(mod, x, o)
    mod.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
        out = call(mod, xx, oo)
        out.float().sum().backward()
    g.replay()

def exp_d():
    mod, x, o = mk()
    warm(mod, x, o)
    params = [p for p in mod.parameters() if p.requires_grad]
    pool = torch.cuda.graph_pool_handle()
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
    with torch.cuda.graph(gf, pool=pool):
        out = call(mod, xx, oo)
    print("   fwd captured", flush=True)
    gout = torch.ones_like(out)
    with torch.cuda.graph(gb, pool=pool):
        grads = torch.autograd.grad((out,), (xx, oo) + tuple(params), (gout,), allow_unused=True)
    print("   bwd captured", flush=True)
    gf.replay(); gb.replay()

def exp_e():
    mod, x, o = mk()
    xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
    f = torch.cuda.make_graphed_callables(mod, (xx.permute(0, 2, 1).unsqueeze(-1), oo.permute(0, 2, 1).unsqueeze(-1)))
    out, idx = f(xx.permute(0, 2, 1).unsqueeze(-1), oo.permute(0, 2, 1).unsqueeze(-1))
    out.float().sum().backward()

def leaf_like(v):
    t = torch.empty_strided(v.shape, v.stride(), dtype=v.dtype, device=v.device)
    t.copy_(v.detach())
    return t.requires_grad_()

def exp_f():
    """make_graphed_callables with LEAF sample inputs that keep the token-major strides (no autograd history from the default stream)"""
    mod, x, o = mk()
    xv, ov = leaf_like(x.permute(0, 2, 1).unsqueeze(-1)), leaf_like(o.permute(0, 2, 1).unsqueeze(-1))
    f = torch.cuda.make_graphed_callables(mod, (xv, ov))
    print("   graphed", flush=True)
    mod2, _, _ = mk()
    for it in range(2):
        xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
        out, idx = f(xx.permute(0, 2, 1).unsqueeze(-1), oo.permute(0, 2, 1).unsqueeze(-1))
        out.float().sum().backward()
        x2, o2 = x.detach().requires_grad_(), o.detach().requires_grad_()
        out2, idx2 = mod2(x2.permute(0, 2, 1).unsqueeze(-1), o2.permute(0, 2, 1).unsqueeze(-1))
        out2.float().sum().backward()
        print("   it", it, "out maxdiff", (out.float() - out2.float()).abs().max().item(), "dx maxdiff", (xx.grad.float() - x2.grad.float()).abs().max().item(),
              "idx equal", bool(torch.equal(idx, idx2)), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["a", "b", "c", "d", "e"]
    for w in which:
        attempt("exp_" + w, globals()["exp_" + w])

"""Diagnostic: host-side enqueue cost of one adapter-stack step.  At BT=1 the kernels are tiny, so the wall time of a
step is the Python + ctypes + launch overhead that a full-size step must hide behind GPU work (GPU box only)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops, stack as S


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for BT in (1, 160):
        model = S.AdapterStack("ave", seed=0).to(dev).train()
        params = [p for p in model.parameters() if p.requires_grad]
        tok = S.make_stage_tokens("ave", BT, torch.bfloat16, dev, seed=1)

        def step():
            for p in params:
                p.grad = None
            leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tok.items()}
            model(leaves).backward()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ops.LAUNCHES = 0
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out[f"BT{BT}"] = {"enqueue_ms_per_step": (t1 - t0) / n * 1e3, "wall_ms_per_step": (t2 - t0) / n * 1e3, "launches_per_step": ops.LAUNCHES / n}
        del model, tok
    # a host model that calls the drop-in modules one by one (the reference's forward): eager launch path vs per-module CUDA graphs
    try:
        from avmoe_b200.adapters import graph_adapters
        for stage, modality in ((3, "audio"), (1, "visual")):
            for BT in (20,):
                Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
                (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
                x = torch.randn(BT, N, C, device=dev).bfloat16()
                o = torch.randn(BT, Nx, Cx, device=dev).bfloat16()
                rec = {}
                for kind in ("eager", "graphed"):
                    mod = S.build_adapter("ave", stage, modality).to(dev).train()
                    if kind == "graphed":
                        graph_adapters(mod, (x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1)))

                    def call():
                        xx, oo = x.detach().requires_grad_(), o.detach().requires_grad_()
                        res = mod(xx.permute(0, 2, 1).unsqueeze(-1), oo.permute(0, 2, 1).unsqueeze(-1))
                        res[0].float().sum().backward()

                    for _ in range(3):
                        call()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    n = 20
                    for _ in range(n):
                        call()
                    t1 = time.perf_counter()
                    torch.cuda.synchronize()
                    t2 = time.perf_counter()
                    rec[kind] = {"enqueue_ms_per_call": (t1 - t0) / n * 1e3, "wall_ms_per_call": (t2 - t0) / n * 1e3}
                out[f"module_ave_{stage}_{modality}_BT{BT}"] = rec
    except Exception as ex:
        out["module_graph_error"] = str(ex)[:300]
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/host_overhead.json", "w"), indent=1)


if __name__ == "__main__":
    main()

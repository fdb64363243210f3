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
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/host_overhead.json", "w"), indent=1)


if __name__ == "__main__":
    main()

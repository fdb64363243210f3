"""One steady-state eager step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`):
the launch list / DRAM bytes of exactly one forward+backward of the adapter stack.  usage: one_step.py [variant] [BT]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops, stack as S


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "ave"
    BT = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    dev = torch.device("cuda", 0)
    model = S.AdapterStack(variant, seed=0).to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    tok = S.make_stage_tokens(variant, BT, torch.bfloat16, dev, seed=1000)

    def step():
        for p in params:
            p.grad = None
        leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tok.items()}
        model(leaves).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ops.LAUNCHES = 0
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"variant": variant, "BT": BT, "library_launches": ops.LAUNCHES}))


if __name__ == "__main__":
    main()

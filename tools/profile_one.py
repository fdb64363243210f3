"""Run one MoEAdapter fwd+bwd at a backbone shape (for ncu captures).  usage: profile_one.py [variant] [stage] [modality] [BT] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import stack as S

def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "ave"
    stage = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    modality = sys.argv[3] if len(sys.argv) > 3 else "audio"
    BT = int(sys.argv[4]) if len(sys.argv) > 4 else 160
    dtype = torch.float32 if (len(sys.argv) > 5 and sys.argv[5] == "f32") else torch.bfloat16
    torch.manual_seed(0)
    mod = S.build_adapter(variant, stage, modality).cuda().train()
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("gate") or n.endswith("gate_av") or n.endswith("gate_self"):
                p.fill_(0.5)
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    x = (torch.randn(BT, N, C, device="cuda") + torch.randn(BT, 1, C, device="cuda")).to(dtype).requires_grad_()
    o = (torch.randn(BT, Nx, Cx, device="cuda") + torch.randn(BT, 1, Cx, device="cuda")).to(dtype).requires_grad_()
    from avmoe_b200 import ops
    for it in range(2):
        if it == 1 and os.environ.get("AVMOE_DUMP_TAGS"):
            ops.GEMM_PROFILE = []
        res = mod(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
        res[0].float().sum().backward()
    torch.cuda.synchronize()
    if os.environ.get("AVMOE_DUMP_TAGS"):      # GEMM shapes of the last iteration, in launch order (to join with an ncu launch list)
        import json
        json.dump([t[4] for t in ops.GEMM_PROFILE], open(os.environ["AVMOE_DUMP_TAGS"], "w"))
    print("done")

if __name__ == "__main__":
    main()

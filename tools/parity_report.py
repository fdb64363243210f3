"""Diagnostic: error table of the CUDA drop-in vs the CPU oracle at backbone shapes (GPU box only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import avmoe_oracle as O
from tests import util
from tests.test_adapter_gpu import _oracle_run, _cuda_run

def main():
    cases = [("ave", 3, "audio"), ("ave", 3, "visual"), ("ave", 1, "audio"), ("ave", 1, "visual"), ("avqa", 3, "audio"), ("ms3", 3, "audio")]
    if os.environ.get("AVMOE_PARITY_CASES"):
        cases = [tuple(int(t) if t.isdigit() else t for t in c.split("-")) for c in os.environ["AVMOE_PARITY_CASES"].split(",")]
    if os.environ.get("AVMOE_GEMM_ATTENTION"):
        from avmoe_b200 import ops
        ops.USE_GEMM_ATTENTION = {"0": False, "1": True}.get(os.environ["AVMOE_GEMM_ATTENTION"], "always")
    rows = []
    for variant, stage, modality in cases:
        for dtype in (torch.float32, torch.bfloat16):
            spec = O.spec_for(variant, stage, modality)
            BT = 4 if stage > 1 else 3
            sd = O.make_state_dict(spec, seed=7)
            x = O.make_tokens(BT, spec.N, spec.C, 31); o = O.make_tokens(BT, spec.Nx, spec.Cx, 32)
            if dtype == torch.bfloat16:
                x, o = x.bfloat16().float(), o.bfloat16().float()
            gout = torch.randn((BT, spec.N, spec.C), generator=torch.Generator().manual_seed(9))
            res, dx_ref, do_ref, dp_ref, _ = _oracle_run(spec, sd, x, o, gout)
            out, idx, dx, do, dparams, mod = _cuda_run(spec, sd, x, o, gout, dtype)
            worst = max(((util.rel_err(dparams[k], g, floor=2e-2 * max(v.abs().max().item() for v in dp_ref.values())), k) for k, g in dp_ref.items()))
            lg = res["logits"].topk(2, dim=-1).values
            row = dict(case=f"{variant}-{stage}-{modality}", dtype=str(dtype)[6:], out_max=util.rel_err(out, res["out"]), out_fro=util.fro_err(out, res["out"]),
                       dx_max=util.rel_err(dx, dx_ref), dx_fro=util.fro_err(dx, dx_ref), do_max=util.rel_err(do, do_ref), do_fro=util.fro_err(do, do_ref),
                       dparam_worst=worst[0], dparam_which=worst[1], idx_ok=bool(torch.equal(idx, res["idx"])), margin=(lg[:, 0] - lg[:, 1]).min().item())
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/parity_report_impl{os.environ.get('AVMOE_GEMM_IMPL','0')}.json", "w"), indent=1)

if __name__ == "__main__":
    main()

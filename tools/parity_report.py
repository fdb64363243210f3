"""Diagnostic: error table of the CUDA drop-in vs the CPU oracle at backbone shapes (GPU box only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import avmoe_oracle as O
from tests import util
from tests.test_adapter_gpu import _oracle_run, _cuda_run

def _eager_autocast(spec, sd, x, o, gout):
    """Comparator: the oracle's plain-torch math on the GPU under torch.autocast(bf16) -- what the reference's own modules
    give with AMP -- measured against the same fp32 CPU oracle.  Puts our bf16 errors next to an independent bf16 execution."""
    dev = "cuda"
    sdg = {k: (v.to(dev).clone().requires_grad_() if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
    xg, og = x.to(dev).clone().requires_grad_(), o.to(dev).clone().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res = O.moe_adapter_forward(xg, og, sdg, spec, training=True, new_buffers={})
        loss = (res["out"].float() * gout.to(dev)).sum()
        if res["lb_loss"] is not None:
            loss = loss + res["lb_loss"]
    loss.backward()
    return res["out"].detach().float().cpu(), xg.grad.float().cpu(), og.grad.float().cpu(), res["idx"].cpu()


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
            if dtype == torch.bfloat16:
                try:
                    eo, edx, edo, eidx = _eager_autocast(spec, sd, x, o, gout)
                    row.update(eager_bf16_out_fro=util.fro_err(eo, res["out"]), eager_bf16_dx_fro=util.fro_err(edx, dx_ref),
                               eager_bf16_do_fro=util.fro_err(edo, do_ref), eager_bf16_idx_ok=bool(torch.equal(eidx, res["idx"])))
                except Exception as ex:
                    row.update(eager_bf16_error=str(ex)[:120])
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/parity_report_impl{os.environ.get('AVMOE_GEMM_IMPL','0')}{os.environ.get('AVMOE_PARITY_TAG','')}.json", "w"), indent=1)

if __name__ == "__main__":
    main()

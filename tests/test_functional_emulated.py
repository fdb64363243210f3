"""CPU: the host orchestration (pass structure + hand-derived backward) over the torch kernel
emulation reproduces the reference golden vectors and the oracle for every variant."""
import types

import pytest
import torch

from avmoe_b200 import adapters
from avmoe_b200 import functional as Fn
from oracle import avmoe_oracle as O
from oracle import ref_harness as H
from tests import kernel_emulation as KE
from tests import util

TOL = 5e-5


def build_module(spec: O.AdapterSpec, sd, kernels=KE):
    """Instantiate the drop-in for a spec exactly as the reference builders would, and load reference weights."""
    opt = H.make_opt(spec)
    common = dict(input_dim=spec.C, output_dim=spec.C, adapter_kind="bottleneck", dim_list=None, layer_idx=0,
                  opt=opt, conv_dim_in=spec.Nx, conv_dim_out=spec.N, linear_in=spec.Cx, linear_out=spec.C)
    cls = adapters.VARIANTS[spec.variant]
    if spec.variant == "avvp":
        mod = cls(**common)
    elif spec.variant == "avqa":
        mod = cls(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, **common)
    else:
        mod = cls(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, num_tk=spec.T, **common)
    mod.load_state_dict(sd, strict=True)
    mod._kernels = kernels
    for m in mod.modules():
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return mod


def run_module_on_fixture(mod, fx, device="cpu", dtype=torch.float32):
    spec = fx["spec"]
    mod = mod.to(device)
    mod.train(fx["training"])
    x = fx["x"].to(device=device, dtype=dtype).clone().requires_grad_()
    o = fx["o"].to(device=device, dtype=dtype).clone().requires_grad_()
    xr, orr = x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1)
    probs = lb = idx = None
    if spec.variant in ("s4", "ms3"):
        out, idx, probs, lb = mod(xr, orr, is_training=False)
    elif spec.variant == "avvp":
        out, lb = mod(xr, orr)
    else:
        out, idx = mod(xr, orr)
    out_tok = out.squeeze(-1).permute(0, 2, 1)
    assert out_tok.is_contiguous(), "drop-in must hand back a view of a contiguous token-major buffer"
    loss = (out_tok.float() * fx["gout"].to(device)).sum()
    if fx["lb_loss"] is not None:
        loss = loss + lb
    if fx["gprobs"] is not None:
        loss = loss + (probs.squeeze(1) * fx["gprobs"].to(device)).sum()
    loss.backward()
    return {"out": out_tok.detach(), "idx": None if idx is None else idx.reshape(-1), "probs": None if probs is None else probs.squeeze(1).detach(),
            "lb_loss": lb.detach() if torch.is_tensor(lb) else None, "dx": x.grad, "do": o.grad,
            "dparams": {k: p.grad for k, p in mod.named_parameters() if p.grad is not None},
            "buffers": {k: v.detach().clone() for k, v in mod.named_buffers()}}


def compare_to_fixture(got, fx, tol, gtol=None, tensor_core=False):
    """tensor_core=True: the CUDA path rounds GEMM operands to TF32/bf16 — compare outputs in the Frobenius norm (max error
    gets 3x) and scalar gate gradients (sums with heavy cancellation) against the overall gradient scale."""
    gtol = gtol or 10 * tol
    if tensor_core:
        assert util.fro_err(got["out"], fx["out"]) < tol
        assert util.rel_err(got["out"], fx["out"]) < 3 * tol
    else:
        assert util.rel_err(got["out"], fx["out"]) < tol
    if fx["idx"] is not None:
        assert torch.equal(got["idx"].cpu(), fx["idx"]), "expert selection must be bit-exact"
    if fx["probs"] is not None:
        assert util.rel_err(got["probs"], fx["probs"]) < tol
    if fx["lb_loss"] is not None:
        assert abs(got["lb_loss"].item() - fx["lb_loss"].item()) < max(tol, 1e-5) * max(1.0, abs(fx["lb_loss"].item()))
    assert util.rel_err(got["dx"], fx["dx"]) < gtol
    assert util.rel_err(got["do"], fx["do"]) < gtol
    gmax = max(v.abs().max().item() for v in fx["dparams"].values())
    for k, g in fx["dparams"].items():
        assert k in got["dparams"], f"missing gradient for {k}"
        if tensor_core and g.numel() == 1:
            assert abs(got["dparams"][k].item() - g.item()) < gtol * max(abs(g.item()), 0.25 * gmax), k
        else:
            assert util.rel_err(got["dparams"][k], g, floor=(2e-2 if tensor_core else 1e-3) * gmax) < gtol, k
    if fx["training"]:
        for k, v in fx["buffers_after"].items():
            assert util.rel_err(got["buffers"][k].float(), v.float(), floor=1e-6) < tol, k


@pytest.mark.parametrize("gemm_attention,gemm_experts", [(True, True), (True, False), (False, False)])
@pytest.mark.parametrize("name", util.golden_names())
def test_emulated_pipeline_matches_reference_golden(name, gemm_attention, gemm_experts, monkeypatch):
    """all orchestrations: latent attention + expert projections through the tensor-core GEMM (the bf16 product default),
    GEMM attention with the fused expert kernels, and the fully fused kernels (the fp32 default)"""
    monkeypatch.setattr(KE, "USE_GEMM_ATTENTION", "always" if gemm_attention else False)
    monkeypatch.setattr(KE, "USE_GEMM_EXPERTS", "always" if gemm_experts else False)
    fx = util.load_golden(name)
    mod = build_module(fx["spec"], fx["state_dict"])
    got = run_module_on_fixture(mod, fx)
    compare_to_fixture(got, fx, TOL)


def test_state_dict_keys_match_reference_fixture():
    for name in util.golden_names():
        fx = util.load_golden(name)
        mod = build_module(fx["spec"], fx["state_dict"])
        assert set(mod.state_dict().keys()) == set(fx["state_dict"].keys()), name
        for k, v in mod.state_dict().items():
            assert tuple(v.shape) == tuple(fx["state_dict"][k].shape), (name, k)


def test_param_names_cover_all_parameters():
    for name in util.golden_names():
        fx = util.load_golden(name)
        mod = build_module(fx["spec"], fx["state_dict"])
        assert sorted(Fn.param_names(mod.cfg)) == sorted(k for k, _ in mod.named_parameters()), name

"""Constructor/opt flag sweep: every switch the reference builders expose (is_bn, is_gate, is_before/post_layernorm,
num_conv_group, Adapter_downsample, num_tokens, expert counts, load-balancing loss, eval mode) against the oracle.

CPU: host orchestration over the kernel emulation, all three orchestrations.  GPU: the CUDA drop-in, fp32 and bf16 streams."""
import pytest
import torch

from oracle import avmoe_oracle as O
from tests import kernel_emulation as KE
from tests import util
from tests.test_functional_emulated import build_module

# (variant, overrides)
SWEEP = [
    ("ave", dict(use_bn=False)),
    ("ave", dict(use_gate=False)),
    ("ave", dict(ln_before=False)),
    ("ave", dict(ln_post=False)),
    ("ave", dict(use_bn=False, use_gate=False, ln_before=False, ln_post=False)),
    ("ave", dict(G=1)),
    ("ave", dict(G=4, r=4)),
    ("ave", dict(T=5)),
    ("ave", dict(T=87)),
    ("ave", dict(Mm=2, Ms=0)),
    ("ave", dict(Mm=0, Ms=2)),
    ("ave", dict(Mm=2, Ms=2)),
    ("avvp", dict(lb_loss=True)),
    ("avvp", dict(use_bn=True)),
    ("ms3", dict(use_bn=True, Mm=1, Ms=2)),
    ("s4", dict(ln_post=False, use_gate=False)),
]
IDS = [f"{v}-" + "-".join(f"{k}{int(x) if isinstance(x, bool) else x}" for k, x in ov.items()) for v, ov in SWEEP]


def _small_spec(variant, overrides):
    import dataclasses
    return dataclasses.replace(O.spec_for(variant, 3, "audio", **overrides), C=64, N=24, Cx=96, Nx=20)


def _run_oracle(spec, sd, x, o, gout, training):
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    x, o = x.clone().requires_grad_(), o.clone().requires_grad_()
    res = O.moe_adapter_forward(x, o, sdg, spec, training=training, new_buffers={})
    loss = (res["out"] * gout).sum()
    if res["lb_loss"] is not None:
        loss = loss + res["lb_loss"]
    loss.backward()
    return res, x.grad, o.grad, {k: v.grad for k, v in sdg.items() if v.is_floating_point() and v.grad is not None}


def _run_module(mod, spec, x, o, gout, training, device, dtype):
    mod = mod.to(device)
    mod.train(training)
    xc, oc = x.to(device, dtype).requires_grad_(), o.to(device, dtype).requires_grad_()
    args = (xc.permute(0, 2, 1).unsqueeze(-1), oc.permute(0, 2, 1).unsqueeze(-1))
    idx = lb = None
    if spec.variant in ("s4", "ms3"):
        out, idx, _, lb = mod(*args, is_training=False)
    elif spec.variant == "avvp":
        out, lb = mod(*args)
    else:
        out, idx = mod(*args)
    out_tok = out.squeeze(-1).permute(0, 2, 1)
    loss = (out_tok.float() * gout.to(device)).sum()
    if torch.is_tensor(lb):
        loss = loss + lb
    loss.backward()
    grads = {k: p.grad.float().cpu() for k, p in mod.named_parameters() if p.grad is not None}
    return out_tok.detach().float().cpu(), (None if idx is None else idx.reshape(-1).cpu()), xc.grad.float().cpu(), oc.grad.float().cpu(), grads


def _compare(spec, ref, got, tol, gtol):
    res, dx_ref, do_ref, dp_ref = ref
    out, idx, dx, do, dparams = got
    assert util.fro_err(out, res["out"]) < tol
    if idx is not None:
        assert torch.equal(idx, res["idx"]), "expert selection must be bit-exact"
    assert util.fro_err(dx, dx_ref) < gtol
    assert util.fro_err(do, do_ref) < gtol
    gmax = max(v.abs().max().item() for v in dp_ref.values())
    for k, g in dp_ref.items():
        if "running_" in k:      # BatchNorm buffers: the oracle's state-dict tensors all carry grad, the module's buffers do not
            continue
        assert k in dparams, f"missing gradient for {k}"
        if g.numel() == 1:
            assert abs(dparams[k].item() - g.item()) < gtol * max(abs(g.item()), 0.25 * gmax), k
        else:
            assert util.fro_err(dparams[k], g) < gtol or util.rel_err(dparams[k], g, floor=5e-2 * gmax) < gtol, k


def _case(variant, overrides, seed=3, BT=3):
    spec = _small_spec(variant, overrides)
    sd = O.make_state_dict(spec, seed=seed)
    x = O.make_tokens(BT, spec.N, spec.C, 11)
    o = O.make_tokens(BT, spec.Nx, spec.Cx, 12)
    gout = torch.randn((BT, spec.N, spec.C), generator=torch.Generator().manual_seed(13))
    return spec, sd, x, o, gout


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("gemm_paths", [True, False])
@pytest.mark.parametrize("variant,overrides", SWEEP, ids=IDS)
def test_flag_sweep_emulated(variant, overrides, gemm_paths, training, monkeypatch):
    monkeypatch.setattr(KE, "USE_GEMM_ATTENTION", "always" if gemm_paths else False)
    monkeypatch.setattr(KE, "USE_GEMM_EXPERTS", "always" if gemm_paths else False)
    spec, sd, x, o, gout = _case(variant, overrides)
    ref = _run_oracle(spec, sd, x, o, gout, training)
    got = _run_module(build_module(spec, sd, kernels=KE), spec, x, o, gout, training, "cpu", torch.float32)
    _compare(spec, ref, got, 5e-5, 5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant,overrides", SWEEP, ids=IDS)
def test_flag_sweep_cuda(variant, overrides, dtype):
    spec, sd, x, o, gout = _case(variant, overrides)
    if dtype == torch.bfloat16:
        x, o = x.bfloat16().float(), o.bfloat16().float()
    ref = _run_oracle(spec, sd, x, o, gout, True)
    got = _run_module(build_module(spec, sd, kernels=None), spec, x, o, gout, True, "cuda", dtype)
    # outputs at the north-star bounds; the gradient bound of the bf16 streams is 1.5x the backbone-shape one: these toy shapes
    # (C=64, 72 rows) average the bf16 rounding noise of the un-normalised (no-BN) chains over far fewer terms
    tol, gtol = (1e-3, 6e-2) if dtype == torch.float32 else (2e-2, 3e-1)
    _compare(spec, ref, got, tol, gtol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant,BT", [("ave", 1), ("avqa", 1), ("ms3", 1), ("avvp", 2), ("s4", 2), ("s4", 7)])
def test_tiny_batches_cuda(variant, BT, dtype):
    """single-segment batches (BatchNorm statistics over N rows only, one CTA per sample kernels with a 1-long grid axis)"""
    spec, sd, x, o, gout = _case(variant, {}, seed=5, BT=BT)
    if dtype == torch.bfloat16:
        x, o = x.bfloat16().float(), o.bfloat16().float()
    ref = _run_oracle(spec, sd, x, o, gout, True)
    got = _run_module(build_module(spec, sd, kernels=None), spec, x, o, gout, True, "cuda", dtype)
    tol, gtol = (1e-3, 6e-2) if dtype == torch.float32 else (2e-2, 3e-1)
    _compare(spec, ref, got, tol, gtol)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("variant,overrides", [("ave", {}), ("avqa", {}), ("avvp", dict(lb_loss=True)), ("s4", {}), ("ms3", {}),
                                               ("ave", dict(ln_before=False, ln_post=False, use_bn=False)), ("ave", dict(Mm=2, Ms=1, T=20))])
def test_rowfused_orchestration_emulated(variant, overrides, training, monkeypatch):
    """token counts that are multiples of 16 take the row-fused front pass (latent attention inside the expert front kernel, nothing
    saved, backward recomputes): the orchestration around it -- operand packing, the two-phase backward, the d tok' contractions fed
    by the A2 | dS2 the front backward emits -- against the oracle on the kernel emulation."""
    import dataclasses
    monkeypatch.setattr(KE, "USE_GEMM_ATTENTION", "always")
    monkeypatch.setattr(KE, "USE_GEMM_EXPERTS", "always")
    calls = {"fwd": 0, "bwd": 0}
    fwd, bwd = KE.rf_front_fwd, KE.rf_front_bwd
    monkeypatch.setattr(KE, "rf_front_fwd", lambda *a, **k: (calls.__setitem__("fwd", calls["fwd"] + 1), fwd(*a, **k))[1])
    monkeypatch.setattr(KE, "rf_front_bwd", lambda *a, **k: (calls.__setitem__("bwd", calls["bwd"] + 1), bwd(*a, **k))[1])
    spec = dataclasses.replace(O.spec_for(variant, 3, "audio", **overrides), C=64, N=32, Cx=96, Nx=16)
    sd = O.make_state_dict(spec, seed=3)
    BT = 3
    x, o = O.make_tokens(BT, spec.N, spec.C, 11), O.make_tokens(BT, spec.Nx, spec.Cx, 12)
    gout = torch.randn((BT, spec.N, spec.C), generator=torch.Generator().manual_seed(13))
    ref = _run_oracle(spec, sd, x, o, gout, training)
    got = _run_module(build_module(spec, sd, kernels=KE), spec, x, o, gout, training, "cpu", torch.float32)
    assert calls == {"fwd": 1, "bwd": 1}, calls
    _compare(spec, ref, got, 5e-5, 5e-4)

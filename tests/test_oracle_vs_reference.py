"""CPU, container-only: oracle restatement vs the UNMODIFIED reference classes at real
stage-3 backbone shapes for all five variants (skipped where /root/reference is absent)."""
import pytest
import torch

from oracle import avmoe_oracle as O
from oracle import ref_harness as H
from oracle import ref_loader
from tests import util

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("variant", ["ave", "avqa", "avvp", "s4", "ms3"])
@pytest.mark.parametrize("modality", ["audio", "visual"])
def test_oracle_equals_reference_stage3(variant, modality):
    spec = O.spec_for(variant, 3, modality)
    sd = O.make_state_dict(spec, seed=3)
    BT = 4
    x = O.make_tokens(BT, spec.N, spec.C, 21).requires_grad_()
    o = O.make_tokens(BT, spec.Nx, spec.Cx, 22).requires_grad_()
    ref = H.build_reference(spec, sd)
    for m in ref.modules():
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    H.make_bn_grads_contiguous(ref)
    r = H.run_reference(ref, spec, x, o, training=True)
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    x2, o2 = x.detach().clone().requires_grad_(), o.detach().clone().requires_grad_()
    nb = {}
    q = O.moe_adapter_forward(x2, o2, sdg, spec, training=True, new_buffers=nb)
    assert util.rel_err(q["out"], r["out"]) < 2e-5
    if r["idx"] is not None:
        assert torch.equal(q["idx"], r["idx"])
    gout = torch.randn(r["out"].shape, generator=torch.Generator().manual_seed(5))
    lr, lq = (r["out"] * gout).sum(), (q["out"] * gout).sum()
    if r["lb_loss"] is not None:
        assert abs(r["lb_loss"].item() - q["lb_loss"].item()) < 1e-5
        lr, lq = lr + r["lb_loss"], lq + q["lb_loss"]
    lr.backward()
    lq.backward()
    assert util.rel_err(x2.grad, x.grad) < 2e-4
    assert util.rel_err(o2.grad, o.grad) < 2e-4
    pg = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in pg.values())
    for k, p in pg.items():
        assert util.rel_err(sdg[k].grad, p.grad, floor=1e-3 * gmax) < 2e-4, k
    for k, v in ref.state_dict().items():
        if k in nb:
            assert util.rel_err(nb[k].float(), v.float(), floor=1e-6) < 2e-5, k


def test_reference_bn_backward_bug_is_torch_not_oracle():
    """fp64 finite differences side with the oracle where the raw reference CPU backward
    (without the contiguous-grad hook) disagrees: documents the torch CPU batch_norm bug."""
    dt = torch.float64
    spec = O.spec_for("ave", 3, "audio")
    sd = O.make_state_dict(spec, seed=1, dtype=dt)
    x = O.make_tokens(2, spec.N, spec.C, 11, dtype=dt)
    gout = torch.randn((2, spec.N, spec.C), generator=torch.Generator().manual_seed(5), dtype=dt)
    key = "singlemodal_experts.0.bn2.bias"

    def loss_with(bias_delta, c):
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2[key][c] += bias_delta
        return (O.expert(x, None, sd2, "singlemodal_experts.0.", spec, False, True, None) * gout).sum().item()

    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    (O.expert(x, None, sdg, "singlemodal_experts.0.", spec, False, True, None) * gout).sum().backward()
    for c in (0, 7):
        fd = (loss_with(1e-6, c) - loss_with(-1e-6, c)) / 2e-6
        assert abs(fd - sdg[key].grad[c].item()) < 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("variant", ["ave", "avqa", "avvp", "s4", "ms3"])
def test_patch_reference_module_swaps_in_the_drop_ins(variant):
    """after patch_reference_module() the reference model module constructs the CUDA drop-ins from the
    same constructor call, and their state dict has the reference's keys and shapes (checkpoints stay loadable)."""
    import sys
    from avmoe_b200 import adapters
    Expert, MoE = ref_loader.load_variant(variant)
    module = sys.modules[MoE.__module__]
    spec = O.spec_for(variant, 3, "audio")
    ref_sd = H.build_reference(spec).state_dict()
    saved = (module.MoEAdapter, module.ExpertAdapter)
    try:
        adapters.patch_reference_module(module, variant)
        assert module.MoEAdapter is adapters.VARIANTS[variant] and module.ExpertAdapter is adapters.ExpertAdapter
        opt = H.make_opt(spec)
        common = dict(input_dim=spec.C, output_dim=spec.C, adapter_kind="bottleneck", dim_list=None, layer_idx=0, opt=opt,
                      conv_dim_in=spec.Nx, conv_dim_out=spec.N, linear_in=spec.Cx, linear_out=spec.C)
        if variant == "avvp":
            mod = module.MoEAdapter(**common)
        elif variant == "avqa":
            mod = module.MoEAdapter(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, **common)
        else:
            mod = module.MoEAdapter(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, num_tk=spec.T, **common)
        ours = mod.state_dict()
        assert set(ours) == set(ref_sd), set(ours) ^ set(ref_sd)
        for k, v in ref_sd.items():
            assert tuple(ours[k].shape) == tuple(v.shape), k
        mod.load_state_dict(ref_sd, strict=True)
    finally:
        module.MoEAdapter, module.ExpertAdapter = saved

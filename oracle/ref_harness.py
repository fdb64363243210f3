"""TEST INFRASTRUCTURE — drive the UNMODIFIED reference MoEAdapter for an AdapterSpec.

Builds the duck-typed ``opt`` namespace each reference variant reads (SURVEY.md §5
"Config / flags"), instantiates the reference class through ``ref_loader`` and runs it on
the reference's own tensor convention (``[BT,C,N,1]`` permuted views of token-major
tensors, AVMOE/AVE/nets/net_trans_v3.py:695-697).  Container-only (needs /root/reference).
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch

from . import ref_loader
from .avmoe_oracle import AdapterSpec


def make_opt(spec: AdapterSpec) -> types.SimpleNamespace:
    return types.SimpleNamespace(
        num_multimodal_experts=spec.Mm, num_singlemodal_experts=spec.Ms, num_conv_group=spec.G,
        is_before_layernorm=int(spec.ln_before), is_post_layernorm=int(spec.ln_post),
        is_self_attention=int(spec.self_attn in ("mha", "latent")),
        self_attention_version={"mha": "v1", "latent": "v2"}.get(spec.self_attn, "v1"),
        use_load_balacing_loss=int(spec.lb_loss),
        Adapter_downsample=spec.r, is_bn=int(spec.use_bn), is_gate=int(spec.use_gate), num_tokens=spec.T,
    )


def build_reference(spec: AdapterSpec, sd: Optional[Dict[str, torch.Tensor]] = None) -> torch.nn.Module:
    _, MoE = ref_loader.load_variant(spec.variant)
    opt = make_opt(spec)
    common = dict(input_dim=spec.C, output_dim=spec.C, adapter_kind="bottleneck", dim_list=None, layer_idx=0,
                  opt=opt, conv_dim_in=spec.Nx, conv_dim_out=spec.N, linear_in=spec.Cx, linear_out=spec.C)
    if spec.variant == "avvp":
        mod = MoE(**common)
    elif spec.variant == "avqa":
        mod = MoE(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, **common)
    else:
        mod = MoE(reduction_factor=spec.r, use_bn=spec.use_bn, use_gate=spec.use_gate, num_tk=spec.T, **common)
    if sd is not None:
        missing, unexpected = mod.load_state_dict(sd, strict=True), None
    return mod


def run_reference(mod: torch.nn.Module, spec: AdapterSpec, x_tok: torch.Tensor, o_tok: torch.Tensor,
                  training: bool = True):
    """x_tok [BT,N,C], o_tok [BT,Nx,Cx] (token-major, may require grad).  Returns dict like the oracle."""
    mod.train(training)
    x = x_tok.permute(0, 2, 1).unsqueeze(-1)
    o = o_tok.permute(0, 2, 1).unsqueeze(-1)
    if spec.variant in ("s4", "ms3"):
        out, idx, probs, lb = mod(x, o, is_training=False)   # is_training only gates router noise
        probs = probs.squeeze(1)
    elif spec.variant == "avvp":
        out, lb = mod(x, o)
        idx, probs = None, None
    else:
        out, idx = mod(x, o)
        probs, lb = None, None
    res = {"out": out.squeeze(-1).permute(0, 2, 1), "idx": None if idx is None else idx.reshape(-1),
           "probs": probs, "lb_loss": lb if torch.is_tensor(lb) else None}
    return res


def make_bn_grads_contiguous(mod: torch.nn.Module) -> None:
    """Work around a torch CPU bug that the reference trips in train mode with is_bn=1.

    The reference feeds BatchNorm2d outputs into ``.squeeze(-1).permute(0,2,1)`` + LayerNorm
    (net_trans_v3.py:403,430-431), so the gradient arriving at the BN output has
    channels-last-looking strides while the BN *input* is NCHW-contiguous; torch 2.11's CPU
    ``native_batch_norm_backward`` then returns wrong input/weight grads (fp64 finite
    differences side with the oracle, see tests/test_oracle_vs_reference.py).  A forward hook
    that re-lays the incoming gradient out contiguously leaves the reference math untouched and
    avoids the buggy kernel path.
    """
    def fwd_hook(_m, _inp, out):
        if out.requires_grad:
            out.register_hook(lambda g: g.contiguous())
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.register_forward_hook(fwd_hook)

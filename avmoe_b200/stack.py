"""The AVMoE adapter *stack* of a task variant on synthetic backbone tokens.

This is the measurement workload: the frozen Swin-V2 / HTS-AT blocks are replaced by the
identity, so one step is exactly the reference builders' adapter traffic — per adapted block the p1 pair
(audio + visual adapter on the same pre-block tokens), residual adds, then the p2 pair, residual adds
(AVMOE/AVE/nets/net_trans_v3.py:693-720) — for every adapted block of every adapted stage.  It is built from the
drop-in modules only, so it doubles as an integration test of the reference call-site convention.
"""
from __future__ import annotations

import types
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import adapters

# backbone token shapes, Swin-V2-L @192 / HTS-AT:  stage -> (Cv, Nv, Ca, Na)   
STAGE_SHAPES = {1: (192, 2304, 96, 4096), 2: (384, 576, 192, 1024), 3: (768, 144, 384, 256), 4: (1536, 36, 768, 64)}

VARIANT_HP = {
    # launch-script hyper-parameters (SURVEY.md §8a "variant deltas")
    "ave": dict(r=8, G=2, T=32, Mm=1, Ms=1, use_bn=1, use_gate=1, self_attn=None, lb=0),
    "avqa": dict(r=8, G=4, T=2, Mm=1, Ms=2, use_bn=0, use_gate=0, self_attn=None, lb=0),
    "avvp": dict(r=8, G=2, T=32, Mm=1, Ms=1, use_bn=0, use_gate=1, self_attn=None, lb=1),
    "s4": dict(r=8, G=2, T=87, Mm=1, Ms=1, use_bn=0, use_gate=0, self_attn="v1", lb=0),
    "ms3": dict(r=2, G=2, T=87, Mm=2, Ms=2, use_bn=0, use_gate=0, self_attn="v2", lb=1),
}
VARIANT_SITES = {
    "ave": ((1, 2), (3, 6)), "ms3": ((1, 2), (3, 6)),
    "avqa": ((1, 2), (2, 2), (3, 6), (4, 2)), "avvp": ((1, 2), (2, 2), (3, 6), (4, 2)), "s4": ((1, 2), (2, 2), (3, 6), (4, 2)),
}


def tokens_per_segment(variant: str) -> int:
    """adapter-tokens per (clip-segment): sum over MoEAdapter calls of N  (SURVEY.md §8d 'Metric')."""
    tot = 0
    for stage, blocks in VARIANT_SITES[variant]:
        _, Nv, _, Na = STAGE_SHAPES[stage]
        tot += blocks * 2 * (Nv + Na)
    return tot


def make_opt(variant: str) -> types.SimpleNamespace:
    hp = VARIANT_HP[variant]
    return types.SimpleNamespace(
        num_multimodal_experts=hp["Mm"], num_singlemodal_experts=hp["Ms"], num_conv_group=hp["G"], is_before_layernorm=1,
        is_post_layernorm=1, is_self_attention=int(hp["self_attn"] is not None), self_attention_version=hp["self_attn"] or "v1",
        use_load_balacing_loss=hp["lb"], Adapter_downsample=hp["r"], is_bn=hp["use_bn"], is_gate=hp["use_gate"], num_tokens=hp["T"])


def build_adapter(variant: str, stage: int, modality: str) -> nn.Module:
    """One MoEAdapter constructed the way the reference builder constructs it (net_trans_v3.py:599-637)."""
    Cv, Nv, Ca, Na = STAGE_SHAPES[stage]
    hp = VARIANT_HP[variant]
    opt = make_opt(variant)
    if modality == "audio":
        dims = dict(input_dim=Ca, output_dim=Ca, conv_dim_in=Nv, conv_dim_out=Na, linear_in=Cv, linear_out=Ca)
    else:
        dims = dict(input_dim=Cv, output_dim=Cv, conv_dim_in=Na, conv_dim_out=Nv, linear_in=Ca, linear_out=Cv)
    common = dict(adapter_kind="bottleneck", dim_list=None, layer_idx=0, opt=opt, **dims)
    cls = adapters.VARIANTS[variant]
    if variant == "avvp":
        return cls(**common)
    if variant == "avqa":
        return cls(reduction_factor=hp["r"], use_bn=hp["use_bn"], use_gate=hp["use_gate"], **common)
    return cls(reduction_factor=hp["r"], use_bn=hp["use_bn"], use_gate=hp["use_gate"], num_tk=hp["T"], **common)


class AdapterStack(nn.Module):
    """All MoEAdapters of a variant: per adapted block {audio,vis} x {p1,p2}, as four ModuleLists named like the reference."""

    def __init__(self, variant: str = "ave", seed: int = 0, init_gates: bool = True):
        super().__init__()
        self.variant = variant
        self.sites: List[Tuple[int, int]] = []          # (stage, index within lists)
        a1, v1, a2, v2 = [], [], [], []
        torch.manual_seed(seed)
        for stage, blocks in VARIANT_SITES[variant]:
            for _ in range(blocks):
                self.sites.append((stage, len(a1)))
                a1.append(build_adapter(variant, stage, "audio"))
                v1.append(build_adapter(variant, stage, "visual"))
                a2.append(build_adapter(variant, stage, "audio"))
                v2.append(build_adapter(variant, stage, "visual"))
        self.audio_moe_adapter_blocks_p1 = nn.ModuleList(a1)
        self.vis_moe_adapter_blocks_p1 = nn.ModuleList(v1)
        self.audio_moe_adapter_blocks_p2 = nn.ModuleList(a2)
        self.vis_moe_adapter_blocks_p2 = nn.ModuleList(v2)
        if init_gates:   # reference zero-init would make every adapter output exactly 0
            g = torch.Generator().manual_seed(seed + 1)
            with torch.no_grad():
                for n, p in self.named_parameters():
                    if n.endswith("gate") or n.endswith("gate_av") or n.endswith("gate_self"):
                        p.copy_(torch.randn(p.shape, generator=g) * 0.1 + 0.5)
            for m in self.modules():
                if isinstance(m, nn.MultiheadAttention):
                    m.dropout = 0.0

    parallel_modalities = True
    _side = None

    def _side_stream(self, device):
        if self._side is None or self._side.device != device:
            object.__setattr__(self, "_side", torch.cuda.Stream(device=device))
        return self._side

    def stages(self):
        return [s for s, _ in VARIANT_SITES[self.variant]]

    router_noise = True    # AVS variants: draw the reference's training-time router noise (PVT_AVSModel_v2.py:294-296); parity tests turn it off
    fuse_residual = True   # fold the call-site add f = f + adapter(f, other) into the adapter's combine epilogue

    def _call(self, mod, x_tok, o_tok):
        """reference call-site convention: permuted [BT,C,N,1] views in, residual in the same layout out."""
        xv = x_tok.permute(0, 2, 1).unsqueeze(-1)
        kw = {"residual": xv} if self.fuse_residual else {}
        if self.variant in ("s4", "ms3"):
            kw["is_training"] = bool(self.router_noise and self.training)
        res = mod(xv, o_tok.permute(0, 2, 1).unsqueeze(-1), **kw)
        out = res[0].squeeze(-1).permute(0, 2, 1)
        aux = res[-1] if self.variant in ("avvp", "s4", "ms3") and torch.is_tensor(res[-1]) else None
        return out, aux

    def forward(self, tokens: Dict[int, Tuple[torch.Tensor, torch.Tensor]]):
        """tokens: stage -> (f_a [BT,Na,Ca], f_v [BT,Nv,Cv]).  Returns (loss, lb_loss_sum)."""
        loss = None
        lb_sum = None
        cur = {s: tokens[s] for s in tokens}
        for stage, i in self.sites:
            f_a, f_v = cur[stage]
            for pa, pv in ((self.audio_moe_adapter_blocks_p1, self.vis_moe_adapter_blocks_p1),
                           (self.audio_moe_adapter_blocks_p2, self.vis_moe_adapter_blocks_p2)):
                # The audio-side and visual-side adapters of a block read the same pre-update features and are independent
                # (net_trans_v3.py:695-705), so on a GPU they run on two streams: the small stage-3/4 kernels of one fill the
                # SMs the other leaves idle.  Autograd replays each node on its forward stream, so the backward overlaps too.
                if self.parallel_modalities and f_a.is_cuda:
                    main = torch.cuda.current_stream()
                    side = self._side_stream(f_a.device)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        rv, lv = self._call(pv[i], f_v, f_a)
                    ra, la = self._call(pa[i], f_a, f_v)
                    main.wait_stream(side)
                    rv.record_stream(main)
                    if lv is not None:
                        lv.record_stream(main)
                else:
                    ra, la = self._call(pa[i], f_a, f_v)
                    rv, lv = self._call(pv[i], f_v, f_a)
                if self.fuse_residual:     # the adapters returned the updated streams
                    f_v, f_a = rv, ra
                else:
                    f_v = f_v + rv
                    f_a = f_a + ra
                for l in (la, lv):
                    if l is not None:
                        lb_sum = l if lb_sum is None else lb_sum + l
            cur[stage] = (f_a, f_v)
        for stage in cur:
            f_a, f_v = cur[stage]
            term = f_a.float().sum() + f_v.float().sum()
            loss = term if loss is None else loss + term
        if lb_sum is not None:
            loss = loss + lb_sum
        return loss


def make_stage_tokens(variant: str, BT: int, dtype: torch.dtype, device, seed: int = 1234, pin: bool = False):
    """Synthetic backbone tokens per adapted stage: N(0,1) + per-sample per-channel offset."""
    out = {}
    for stage, _ in VARIANT_SITES[variant]:
        Cv, Nv, Ca, Na = STAGE_SHAPES[stage]
        g = torch.Generator().manual_seed(seed + 97 * stage)
        f_a = torch.randn((BT, Na, Ca), generator=g) + torch.randn((BT, 1, Ca), generator=g)
        f_v = torch.randn((BT, Nv, Cv), generator=g) + torch.randn((BT, 1, Cv), generator=g)
        f_a, f_v = f_a.to(dtype), f_v.to(dtype)
        if pin:
            f_a, f_v = f_a.pin_memory(), f_v.pin_memory()
        out[stage] = (f_a.to(device), f_v.to(device)) if device is not None else (f_a, f_v)
    return out

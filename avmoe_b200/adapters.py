"""Drop-in ``MoEAdapter`` / ``ExpertAdapter`` modules for the five AVMoE task variants.

Same constructor arguments, forward signatures, return tuples and ``state_dict`` key names /
shapes as the reference classes, so the AVE / AVVP / AVQA / AVS model builders can swap them in
and reference checkpoints load with ``load_state_dict``.  Per variant:

  ave   AVMOE/AVE/nets/net_trans_v3.py:296-487        forward(x, vis_token) -> (out, expert_indices)
  avqa  AVMOE/AVQA/net_grd_avst/net_avst_v2.py:215-399  same returns; experts read opt.num_tokens
  avvp  AVMOE/AVVP/nets/mgn.py:39-224                 forward(x, vis_token) -> (out, load_balancing_loss)
  s4    AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:90-318
  ms3   AVMOE/AVS/avs_scripts/avs_ms3/model/PVT_AVSModel_v2.py:282-501
        forward(x, vis_token, is_training=True) -> (out, expert_indices, gating_probs, load_balancing_loss)

Tensor convention (reference call sites, net_trans_v3.py:695-697): ``x`` is ``[BT,C,N,1]`` — the
permuted view ``f.permute(0,2,1).unsqueeze(-1)`` of a contiguous token-major ``[BT,N,C]`` stream,
``vis_token`` likewise ``[BT,Cx,Nx,1]``.  The kernels work token-major; the returned ``out`` is the
``[BT,C,N,1]`` view of a contiguous ``[BT,N,C]`` buffer, so the caller's
``.squeeze(-1).permute(0,2,1)`` is again contiguous and no transpose ever runs.

The modules hold parameters only; all math is ``functional.MoEAdapterFunction`` over the CUDA
kernels.  There is no eager fallback: on a CPU tensor ``forward`` raises.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn


def _opt_get(opt, name, default=None):
    return getattr(opt, name, default) if opt is not None else default


class ExpertAdapter(nn.Module):
    """Parameter container with the reference's attribute names (one expert).

    Signature follows the AVE/AVS copy; ``dim_list``/``layer_idx`` (AVQA/AVVP copies) are accepted and ignored
    exactly as the reference ignores them."""

    def __init__(self, input_dim, output_dim, adapter_kind, reduction_factor=16, opt=None, use_bn=True, use_gate=True,
                 num_tk=87, is_multimodal=True, self_attn="none", dim_list=None, layer_idx=0):
        super().__init__()
        if adapter_kind != "bottleneck":
            raise NotImplementedError  # reference: net_trans_v3.py:373-374 ("basic" is never built by the v2/v3 models)
        self.adapter_kind = adapter_kind
        self.use_bn = bool(use_bn)
        self.is_multimodal = is_multimodal
        self.opt = opt
        self.num_tk = num_tk
        G = opt.num_conv_group
        self.gate = nn.Parameter(torch.zeros(1)) if use_gate else None
        self.down_sample_size = input_dim // reduction_factor
        if is_multimodal:
            self.my_tokens = nn.Parameter(torch.rand((num_tk, input_dim)))
            self.gate_av = nn.Parameter(torch.zeros(1))
        elif self_attn == "nxn":
            self.gate_av = nn.Parameter(torch.zeros(1))
        elif self_attn == "latent":
            self.my_tokens = nn.Parameter(torch.rand((num_tk, input_dim)))
            self.gate_self = nn.Parameter(torch.zeros(1))
        elif self_attn == "mha":
            self.self_attention = nn.MultiheadAttention(input_dim, num_heads=4, dropout=0.2)
        self.down_sampler = nn.Conv2d(input_dim, self.down_sample_size, 1, groups=G, bias=False)
        self.up_sampler = nn.Conv2d(self.down_sample_size, output_dim, 1, groups=G, bias=False)
        if use_bn:
            self.bn1 = nn.BatchNorm2d(self.down_sample_size)
            self.bn2 = nn.BatchNorm2d(output_dim)
        if opt.is_before_layernorm:
            self.ln_before = nn.LayerNorm(output_dim)
        if opt.is_post_layernorm:
            self.ln_post = nn.LayerNorm(output_dim)

    def forward(self, x, vis_token=None):
        raise RuntimeError("avmoe_b200.ExpertAdapter is evaluated by its owning MoEAdapter (fused CUDA path); "
                           "call MoEAdapter.forward")


class _MoEAdapterBase(nn.Module):
    variant = "ave"

    def _build(self, input_dim, output_dim, adapter_kind, reduction_factor, opt, use_bn, use_gate, num_tk,
               conv_dim_in, conv_dim_out, linear_in, linear_out, self_attn, lb_loss):
        if adapter_kind != "bottleneck":
            raise NotImplementedError
        assert input_dim == output_dim == linear_out, "the reference adapters are square (input_dim == output_dim == linear_out)"
        self.opt = opt
        self.use_bn = bool(use_bn)
        self.num_tk = num_tk
        self.conv_adapter = nn.Conv2d(conv_dim_in, conv_dim_out, kernel_size=1)
        self.fc = nn.Linear(linear_in, linear_out)
        self.num_multimodal_experts = opt.num_multimodal_experts
        self.num_singlemodal_experts = opt.num_singlemodal_experts
        mk = lambda mm: ExpertAdapter(input_dim, output_dim, adapter_kind, reduction_factor, opt, use_bn, use_gate, num_tk,
                                      is_multimodal=mm, self_attn=self_attn)
        self.multimodal_experts = nn.ModuleList([mk(True) for _ in range(self.num_multimodal_experts)])
        self.singlemodal_experts = nn.ModuleList([mk(False) for _ in range(self.num_singlemodal_experts)])
        E = self.num_multimodal_experts + self.num_singlemodal_experts
        self.router = nn.Sequential(nn.Linear(input_dim + linear_out, 128), nn.ReLU(), nn.Linear(128, 32), nn.ReLU(),
                                    nn.Linear(32, E))
        self.cfg = Fn.AdapterConfig(
            C=input_dim, N=conv_dim_out, Cx=linear_in, Nx=conv_dim_in, r=reduction_factor, G=opt.num_conv_group, T=num_tk,
            Mm=self.num_multimodal_experts, Ms=self.num_singlemodal_experts, use_bn=bool(use_bn), use_gate=bool(use_gate),
            ln_before=bool(opt.is_before_layernorm), ln_post=bool(opt.is_post_layernorm), self_attn=self_attn, lb_loss=bool(lb_loss))
        self._names = Fn.param_names(self.cfg)
        self._kernels = None   # kernel namespace; resolved lazily to avmoe_b200.ops (tests may inject an emulation)

    # ---- plumbing --------------------------------------------------------------------------------
    def _kernel_namespace(self):
        if self._kernels is None:
            from . import ops
            self._kernels = ops
        return self._kernels

    def _flat_params(self):
        # resolved by attribute walk, not named_parameters(): nn.DataParallel replicas (AVMOE/AVVP/main.py:420-421) hold
        # their weights as plain (non-leaf) tensor attributes, which named_parameters() does not list
        out = []
        for n in self._names:
            obj = self
            for part in n.split("."):
                obj = getattr(obj, part)
            # DataParallel's broadcast_coalesced hands replicas views at arbitrary 4-byte offsets of one flat buffer; the kernels
            # read parameters with 8/16-byte vector loads, so a misaligned replica tensor is re-homed (differentiable clone)
            if obj.data_ptr() % 16 != 0:
                obj = obj.clone()
            out.append(obj)
        return out

    def _bn_buffers(self):
        return {k: v for k, v in self.named_buffers()}

    @staticmethod
    def _token_major(t: torch.Tensor) -> torch.Tensor:
        """[BT,C,N,1] (any strides) -> contiguous [BT,N,C]; free when t is the reference's permuted view."""
        assert t.dim() == 4 and t.shape[-1] == 1, "expected [BT, C, N, 1]"
        return t.squeeze(-1).permute(0, 2, 1).contiguous()

    def _run(self, x, vis_token, noise: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, mha_keep=None):
        K = self._kernel_namespace()
        if K.__name__.endswith(".ops") and not x.is_cuda:
            raise RuntimeError("avmoe_b200 has no CPU path: move the module and its inputs to a CUDA device")
        xt, ot = self._token_major(x), self._token_major(vis_token)
        state = {"cfg": self.cfg, "K": K, "training": self.training, "buffers": self._bn_buffers(), "noise": noise}
        if self.cfg.self_attn == "mha":     # AVS-v1: nn.MultiheadAttention(dropout=0.2) drops attention probabilities in training
            drops = {float(m.self_attention.dropout) for m in self.singlemodal_experts}
            assert len(drops) <= 1, "single-modal experts must share one attention dropout rate"
            state["mha_dropout"] = drops.pop() if (drops and self.training) else 0.0
            state["mha_keep"] = mha_keep
        rt = None if residual is None else self._token_major(residual).to(xt.dtype)
        # the reference call sites add the adapter output to the very stream it reads (f = f + adapter(f, other)): then the residual's
        # gradient is folded into dx inside the backward instead of a separate accumulation pass
        state["res_is_x"] = rt is not None and rt.data_ptr() == xt.data_ptr() and rt.shape == xt.shape
        out, probs, lb, idx = Fn.MoEAdapterFunction.apply(state, xt, ot, rt, *self._flat_params())
        return out.permute(0, 2, 1).unsqueeze(-1), idx.unsqueeze(-1), probs.unsqueeze(1), lb


class MoEAdapterAVE(_MoEAdapterBase):
    """AVMOE/AVE/nets/net_trans_v3.py:438-487."""
    variant = "ave"

    def __init__(self, input_dim, output_dim, adapter_kind, dim_list, layer_idx, reduction_factor=16, opt=None,
                 use_bn=True, use_gate=True, num_tk=87, conv_dim_in=0, conv_dim_out=0, linear_in=0, linear_out=0):
        super().__init__()
        sa = "mha" if _opt_get(opt, "is_self_attention", 0) else "none"
        self._build(input_dim, output_dim, adapter_kind, reduction_factor, opt, use_bn, use_gate, num_tk,
                    conv_dim_in, conv_dim_out, linear_in, linear_out, sa, False)

    def forward(self, x, vis_token=None, residual=None):
        """``residual`` (optional, same layout as ``x``): the stream the call site adds the adapter output to
        (net_trans_v3.py:703,706,718,720); when given, the add is folded into the combine epilogue and ``out`` is the
        updated stream."""
        out, idx, _, _ = self._run(x, vis_token, residual=residual)
        return out, idx


class MoEAdapterAVQA(_MoEAdapterBase):
    """AVMOE/AVQA/net_grd_avst/net_avst_v2.py:349-399 (no num_tk argument: experts read opt.num_tokens, :233)."""
    variant = "avqa"

    def __init__(self, input_dim, output_dim, adapter_kind, dim_list, layer_idx, reduction_factor=16, opt=None,
                 use_bn=True, use_gate=True, conv_dim_in=0, conv_dim_out=0, linear_in=0, linear_out=0):
        super().__init__()
        sa = "mha" if _opt_get(opt, "is_self_attention", 0) else "none"
        self._build(input_dim, output_dim, adapter_kind, reduction_factor, opt, use_bn, use_gate, opt.num_tokens,
                    conv_dim_in, conv_dim_out, linear_in, linear_out, sa, False)

    def forward(self, x, vis_token=None, residual=None):
        out, idx, _, _ = self._run(x, vis_token, residual=residual)
        return out, idx


class MoEAdapterAVVP(_MoEAdapterBase):
    """AVMOE/AVVP/nets/mgn.py:161-224 (r / bn / gate / T come from opt, :47-51; returns the lb-loss, no argmax)."""
    variant = "avvp"

    def __init__(self, input_dim, output_dim, adapter_kind, dim_list, layer_idx, opt=None, conv_dim_in=0, conv_dim_out=0,
                 linear_in=0, linear_out=0):
        super().__init__()
        self._build(input_dim, output_dim, adapter_kind, opt.Adapter_downsample, opt, opt.is_bn, opt.is_gate, opt.num_tokens,
                    conv_dim_in, conv_dim_out, linear_in, linear_out, "nxn", opt.use_load_balacing_loss == 1)

    def forward(self, x, vis_token=None, residual=None):
        out, _, _, lb = self._run(x, vis_token, residual=residual)
        return out, (lb if self.cfg.lb_loss else 0.)


class MoEAdapterAVS(_MoEAdapterBase):
    """AVS-S4 / AVS-MS3: avs_s4/model/PVT_AVSModel_v2.py:253-318, avs_ms3/model/PVT_AVSModel_v2.py:437-501."""
    variant = "avs"

    def __init__(self, input_dim, output_dim, adapter_kind, dim_list, layer_idx, reduction_factor=16, opt=None,
                 use_bn=True, use_gate=True, num_tk=87, conv_dim_in=0, conv_dim_out=0, linear_in=0, linear_out=0):
        super().__init__()
        sa = "none"
        if _opt_get(opt, "is_self_attention", 0):
            sa = {"v1": "mha", "v2": "latent"}[opt.self_attention_version]
        self._build(input_dim, output_dim, adapter_kind, reduction_factor, opt, use_bn, use_gate, num_tk,
                    conv_dim_in, conv_dim_out, linear_in, linear_out, sa, _opt_get(opt, "use_load_balacing_loss", 0) == 1)
        self.noise_std = 0.01   # PVT_AVSModel_v2.py:295

    def forward(self, x, vis_token=None, is_training=True, noise: Optional[torch.Tensor] = None, residual=None, mha_keep=None):
        """``noise`` ([BT,E], optional) lets a caller inject the router noise tensor for reproducible parity runs;
        by default it is drawn with torch.randn exactly where the reference draws it.  ``mha_keep`` (AVS-v1, optional): one uint8
        keep mask [N,4,BT,BT] per single-modal expert for the attention dropout; by default drawn from torch's generator in
        training mode."""
        if is_training and noise is None:
            noise = torch.randn((x.shape[0], self.cfg.E), device=x.device, dtype=torch.float32) * self.noise_std
        if not is_training:
            noise = None
        out, idx, probs, lb = self._run(x, vis_token, noise, residual=residual, mha_keep=mha_keep)
        return out, idx, probs, (lb if self.cfg.lb_loss else 0.)


MoEAdapter = MoEAdapterAVE   # default export mirrors the AVE copy
VARIANTS = {"ave": MoEAdapterAVE, "avqa": MoEAdapterAVQA, "avvp": MoEAdapterAVVP, "s4": MoEAdapterAVS, "ms3": MoEAdapterAVS}


def _leaf_like(t: torch.Tensor) -> torch.Tensor:
    """A fresh LEAF with the sample's shape, strides and values (so the token-major memory under a `[BT,C,N,1]` view is kept)."""
    out = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device)
    out.copy_(t.detach())
    return out.requires_grad_(t.requires_grad or t.is_floating_point())


def graph_adapters(modules, sample_inputs, num_warmup_iters: int = 3):
    """Per-module CUDA graphs for a host model that calls the drop-in modules one by one (the reference's own forward,
    `net_trans_v3.py:695-720`): an eager call costs ~50-100 Python -> C-ABI launches, a graphed call two graph replays.

    `modules`: one MoEAdapter or a tuple of them IN THE ORDER the host model calls them; `sample_inputs`: the matching
    `(x, vis_token)` tuple(s) in the call-site layout (`[BT,C,N,1]` views of token-major tensors are fine).  Wraps
    `torch.cuda.make_graphed_callables`: the modules' `forward` is replaced in place by a graphed forward + backward for this
    (shape, dtype, training mode); calls in another mode fall back to the eager path.  Returns the module(s).

    The samples are re-created as leaf tensors with the same strides first: a sample that is a *view* of a tensor made on the
    default stream carries autograd history (an AccumulateGrad node bound to that stream) which the side-stream capture of the
    backward would have to synchronise with -- that invalidates the capture (`cudaErrorStreamCaptureInvalidated`).
    Limits are torch's: positional tensor arguments only, every output a tensor (AVE / AVQA adapters; the AVVP / AVS forwards return a
    python `0.` for a disabled load-balancing loss), no `residual=` / `noise=` keywords, outputs are static buffers overwritten by
    the next call."""
    single = isinstance(modules, torch.nn.Module)
    mods = (modules,) if single else tuple(modules)
    samples = (tuple(sample_inputs),) if single else tuple(tuple(a) for a in sample_inputs)
    if len(mods) != len(samples):
        raise ValueError("graph_adapters: one sample input tuple per module")
    for args in samples:
        for a in args:
            if not (torch.is_tensor(a) and a.is_cuda):
                raise RuntimeError("graph_adapters: sample inputs must be CUDA tensors (there is no CPU path)")
    leaves = tuple(tuple(_leaf_like(a) for a in args) for args in samples)
    graphed = torch.cuda.make_graphed_callables(mods, leaves, num_warmup_iters=num_warmup_iters)
    return graphed[0] if single else graphed


def patch_reference_module(module, variant: str) -> None:
    """Swap the adapter classes inside an imported reference model module (e.g. ``nets.net_trans_v3``) so that the
    unmodified model builders construct the CUDA drop-ins:  ``patch_reference_module(nets.net_trans_v3, 'ave')``."""
    module.MoEAdapter = VARIANTS[variant]
    module.ExpertAdapter = ExpertAdapter

"""TEST INFRASTRUCTURE — CPU oracle for the AVMoE adapter hot path.  NOT a product path.

A functional, token-major restatement (plain torch fp32/fp64 on CPU, autograd supplies
the backward) of the reference ``MoEAdapter``/``ExpertAdapter`` forward for all five
variants.  It consumes a reference-format ``state_dict`` (same key names / shapes as the
reference modules, SURVEY.md §5 "Checkpoint") so the very same weights can be fed to the
real reference (``oracle/ref_loader.py``), to this oracle, and to the CUDA drop-in.

Parity pinning: the reference ships no golden vectors or tests, so this
restatement is pinned by ``tests/test_oracle_vs_reference.py`` (runs the unmodified
reference classes from /root/reference when present) and by the committed fixtures under
``tests/golden/`` that ``oracle/make_golden.py`` generated FROM THE REFERENCE ITSELF.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference leg may import this module.

Reference lines restated (AVE copy; siblings in parentheses):
  resampler      AVMOE/AVE/nets/net_trans_v3.py:469-471
  router         AVMOE/AVE/nets/net_trans_v3.py:472-479  (noise: avs_s4/model/PVT_AVSModel_v2.py:294-296)
  latent attn    AVMOE/AVE/nets/net_trans_v3.py:379-390
  bottleneck     AVMOE/AVE/nets/net_trans_v3.py:392-403 (multimodal), :416-422 (unimodal)
  post LN/gate   AVMOE/AVE/nets/net_trans_v3.py:430-434
  combine        AVMOE/AVE/nets/net_trans_v3.py:481-486
  NxN self-attn  AVMOE/AVVP/nets/mgn.py:132-139
  MHA self-attn  AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:212-215
  latent self    AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:216-227
  lb-loss        AVMOE/AVVP/nets/mgn.py:220-224
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LN_EPS = 1e-5
BN_MOMENTUM = 0.1
MHA_HEADS = 4


@dataclass
class AdapterSpec:
    """Static description of one MoEAdapter instance (what the reference ctor + opt fix)."""
    variant: str            # ave | avqa | avvp | s4 | ms3
    C: int                  # input_dim == output_dim == linear_out
    N: int                  # conv_dim_out (tokens of the adapted modality)
    Cx: int                 # linear_in (channels of the other modality)
    Nx: int                 # conv_dim_in (tokens of the other modality)
    r: int = 8              # reduction factor
    G: int = 2              # num_conv_group
    T: int = 32             # latent tokens
    Mm: int = 1             # multimodal experts
    Ms: int = 1             # single-modal experts
    use_bn: bool = True
    use_gate: bool = True
    ln_before: bool = True
    ln_post: bool = True
    self_attn: str = "none"  # none | mha | latent | nxn   (single-modal expert pre-transform)
    lb_loss: bool = False

    @property
    def E(self) -> int:
        return self.Mm + self.Ms

    @property
    def Cr(self) -> int:
        return self.C // self.r


def _layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def _grouped_proj(x, w, G):
    """1x1 grouped conv over channels on token-major x [BT,N,Cin]; w [Cout, Cin/G, 1, 1]."""
    BT, N, Cin = x.shape
    Cout = w.shape[0]
    w2 = w.reshape(G, Cout // G, Cin // G)
    xg = x.reshape(BT, N, G, Cin // G)
    return torch.einsum("bngi,goi->bngo", xg, w2).reshape(BT, N, Cout)


def _batch_norm(x, sd, prefix, training, new_buffers):
    """BatchNorm2d over channel axis of token-major x [BT,N,Ch] (batch stats over BT*N)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        flat = x.reshape(-1, x.shape[-1])
        n = flat.shape[0]
        mean = flat.mean(0)
        var = flat.var(0, unbiased=False)
        if new_buffers is not None:
            with torch.no_grad():
                new_buffers[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                new_buffers[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
                new_buffers[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    return (x - mean) * torch.rsqrt(var + BN_EPS) * w + b


def _latent_attention(x, kv, tokens):
    """Latent-token attention: tokens attend to kv, then x attends to the updated tokens.

    x [BT,N,C], kv [BT,Nk,C], tokens [T,C]  ->  x_res [BT,N,C]   (no 1/sqrt(d) scaling)."""
    tok = tokens.unsqueeze(0).expand(x.shape[0], -1, -1)
    a1 = torch.softmax(tok @ kv.transpose(1, 2), dim=-1)      # [BT,T,Nk]
    tok = tok + a1 @ kv                                        # [BT,T,C]
    a2 = torch.softmax(x @ tok.transpose(1, 2), dim=-1)        # [BT,N,T]
    return a2 @ tok


def _mha_over_batch_axis(x, sd, prefix, dropout_mask=None):
    """nn.MultiheadAttention(C, 4) applied to [BT,N,C] with batch_first=False: the attended
    sequence axis is BT, batched over N (reference quirk, PVT_AVSModel_v2.py:212-215)."""
    L, B, C = x.shape
    hd = C // MHA_HEADS
    qkv = x @ sd[prefix + ".in_proj_weight"].t() + sd[prefix + ".in_proj_bias"]
    q, k, v = qkv.split(C, dim=-1)

    def heads(t):  # [L,B,C] -> [B*H, L, hd]
        return t.reshape(L, B * MHA_HEADS, hd).transpose(0, 1)

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q / math.sqrt(hd)) @ k.transpose(1, 2), dim=-1)
    if dropout_mask is not None:
        att = att * dropout_mask
    o = (att @ v).transpose(0, 1).reshape(L, B, C)
    return o @ sd[prefix + ".out_proj.weight"].t() + sd[prefix + ".out_proj.bias"]


def resample(o, sd):
    """vis' = fc(conv_adapter(other))  ->  [BT,N,C]   (net_trans_v3.py:469-471)."""
    w_tok = sd["conv_adapter.weight"].reshape(sd["conv_adapter.weight"].shape[0], -1)   # [N,Nx]
    u = torch.einsum("nk,bkc->bnc", w_tok, o) + sd["conv_adapter.bias"][None, :, None]
    return u @ sd["fc.weight"].t() + sd["fc.bias"]


def route(x, v, sd, noise=None):
    """mean-pool both modalities, 3-layer MLP, softmax, argmax (net_trans_v3.py:472-479)."""
    rin = torch.cat((x.mean(dim=1), v.mean(dim=1)), dim=-1)                   # [BT,2C]
    h = torch.relu(rin @ sd["router.0.weight"].t() + sd["router.0.bias"])
    h = torch.relu(h @ sd["router.2.weight"].t() + sd["router.2.bias"])
    logits = h @ sd["router.4.weight"].t() + sd["router.4.bias"]
    if noise is not None:
        logits = logits + noise
    probs = torch.softmax(logits, dim=-1)
    return probs, torch.argmax(probs, dim=-1), logits


def expert(x, v, sd, prefix, spec: AdapterSpec, multimodal: bool, training: bool, new_buffers=None, mha_dropout_mask=None):
    """One ExpertAdapter on token-major tensors -> [BT,N,C]."""
    p = prefix
    if multimodal:
        x = x + sd[p + "gate_av"] * _latent_attention(x, v, sd[p + "my_tokens"])
    elif spec.self_attn == "nxn":     # AVVP: softmax rows of x x^T, contracted over the ROW index
        att = torch.softmax(x @ x.transpose(1, 2), dim=-1)                    # [BT,N,N] rows i, cols j
        x = x + sd[p + "gate_av"] * (att.transpose(1, 2) @ x)                 # out_j = sum_i att[i,j] x_i
    elif spec.self_attn == "mha":     # replaces x (no residual)
        x = _mha_over_batch_axis(x, sd, p + "self_attention", mha_dropout_mask)
    elif spec.self_attn == "latent":
        x = x + sd[p + "gate_self"] * _latent_attention(x, x, sd[p + "my_tokens"])
    if spec.ln_before:
        x = _layer_norm(x, sd[p + "ln_before.weight"], sd[p + "ln_before.bias"])
    z = _grouped_proj(x, sd[p + "down_sampler.weight"], spec.G)
    if spec.use_bn:
        z = _batch_norm(z, sd, p + "bn1", training, new_buffers)
    if multimodal:
        z = torch.relu(z)
    y = _grouped_proj(z, sd[p + "up_sampler.weight"], spec.G)
    if spec.use_bn:
        y = _batch_norm(y, sd, p + "bn2", training, new_buffers)
    if spec.ln_post:
        y = _layer_norm(y, sd[p + "ln_post.weight"], sd[p + "ln_post.bias"])
    if spec.use_gate:
        y = sd[p + "gate"] * y
    return y


def load_balancing_loss(probs):
    """Reference quirk (mgn.py:220-224): probs is [BT,1,E]; mean over dim 0 -> [1,E]; the
    'uniform' target is 1/size(0) == 1.0; kl_div(log p, 1, 'batchmean') == -sum_e log p_e."""
    pbar = probs.mean(dim=0, keepdim=True)                                    # [1,E]
    target = torch.full_like(pbar, 1.0 / pbar.shape[0])
    return F.kl_div(pbar.log(), target, reduction="batchmean")


def moe_adapter_forward(x, o, sd: Dict[str, torch.Tensor], spec: AdapterSpec, training: bool = True,
                        noise: Optional[torch.Tensor] = None, new_buffers: Optional[dict] = None, mha_dropout_masks=None):
    """Full MoEAdapter forward on token-major inputs.

    mha_dropout_masks (AVS-v1 only): one tensor per single-modal expert, [N*heads, BT, BT], already scaled by 1/(1-p) -- the
    attention-probability dropout of nn.MultiheadAttention(dropout=0.2) (PVT_AVSModel_v2.py:139-142) with an injected mask.

    x [BT,N,C], o [BT,Nx,Cx] -> dict(out [BT,N,C], probs [BT,E], idx [BT] int64, lb_loss, v)."""
    v = resample(o, sd)
    probs, idx, logits = route(x, v, sd, noise)
    out = 0
    e = 0
    for m in range(spec.Mm):
        out = out + probs[:, e, None, None] * expert(x, v, sd, f"multimodal_experts.{m}.", spec, True, training, new_buffers)
        e += 1
    for s in range(spec.Ms):
        out = out + probs[:, e, None, None] * expert(x, v, sd, f"singlemodal_experts.{s}.", spec, False, training, new_buffers,
                                                     None if mha_dropout_masks is None else mha_dropout_masks[s])
        e += 1
    lb = load_balancing_loss(probs.unsqueeze(1)) if spec.lb_loss else None
    return {"out": out, "probs": probs, "idx": idx, "logits": logits, "lb_loss": lb, "v": v}


# ----------------------------------------------------------------------------------------------
# state_dict construction (shapes follow the reference ctors) and synthetic inputs
# ----------------------------------------------------------------------------------------------

def make_state_dict(spec: AdapterSpec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's key names/shapes.  Distributions follow torch's
    default inits closely enough for testing; gates ~ N(0.5,0.1) (zero-init would make every
    output exactly 0 and parity vacuous)."""
    g = torch.Generator().manual_seed(seed)

    def unif(shape, bound):
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound

    def normal(shape, mean, std):
        return torch.randn(shape, generator=g, dtype=dtype) * std + mean

    C, N, Cx, Nx, Cr, G = spec.C, spec.N, spec.Cx, spec.Nx, spec.Cr, spec.G
    sd: Dict[str, torch.Tensor] = {}
    sd["conv_adapter.weight"] = unif((N, Nx, 1, 1), 1 / math.sqrt(Nx))
    sd["conv_adapter.bias"] = unif((N,), 1 / math.sqrt(Nx))
    sd["fc.weight"] = unif((C, Cx), 1 / math.sqrt(Cx))
    sd["fc.bias"] = unif((C,), 1 / math.sqrt(Cx))
    for i, (fo, fi) in zip((0, 2, 4), ((128, 2 * C), (32, 128), (spec.E, 32))):
        sd[f"router.{i}.weight"] = unif((fo, fi), 1 / math.sqrt(fi))
        sd[f"router.{i}.bias"] = unif((fo,), 1 / math.sqrt(fi))

    def expert_sd(p, multimodal):
        if spec.use_gate:
            sd[p + "gate"] = normal((1,), 0.5, 0.1)
        if multimodal:
            sd[p + "my_tokens"] = torch.rand((spec.T, C), generator=g, dtype=dtype)
            sd[p + "gate_av"] = normal((1,), 0.5, 0.1)
        else:
            if spec.self_attn == "nxn":
                sd[p + "gate_av"] = normal((1,), 0.5, 0.1)
            elif spec.self_attn == "latent":
                sd[p + "my_tokens"] = torch.rand((spec.T, C), generator=g, dtype=dtype)
                sd[p + "gate_self"] = normal((1,), 0.5, 0.1)
            elif spec.self_attn == "mha":
                sd[p + "self_attention.in_proj_weight"] = unif((3 * C, C), math.sqrt(6 / (4 * C)))
                sd[p + "self_attention.in_proj_bias"] = normal((3 * C,), 0.0, 0.02)
                sd[p + "self_attention.out_proj.weight"] = unif((C, C), 1 / math.sqrt(C))
                sd[p + "self_attention.out_proj.bias"] = normal((C,), 0.0, 0.02)
        sd[p + "down_sampler.weight"] = unif((Cr, C // G, 1, 1), 1 / math.sqrt(C // G))
        sd[p + "up_sampler.weight"] = unif((C, Cr // G, 1, 1), 1 / math.sqrt(max(Cr // G, 1)))
        if spec.use_bn:
            for name, ch in (("bn1", Cr), ("bn2", C)):
                sd[p + name + ".weight"] = normal((ch,), 1.0, 0.1)
                sd[p + name + ".bias"] = normal((ch,), 0.0, 0.1)
                sd[p + name + ".running_mean"] = normal((ch,), 0.0, 0.1)
                sd[p + name + ".running_var"] = torch.rand((ch,), generator=g, dtype=dtype) + 0.5
                sd[p + name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        if spec.ln_before:
            sd[p + "ln_before.weight"] = normal((C,), 1.0, 0.1)
            sd[p + "ln_before.bias"] = normal((C,), 0.0, 0.1)
        if spec.ln_post:
            sd[p + "ln_post.weight"] = normal((C,), 1.0, 0.1)
            sd[p + "ln_post.bias"] = normal((C,), 0.0, 0.1)

    for m in range(spec.Mm):
        expert_sd(f"multimodal_experts.{m}.", True)
    for s in range(spec.Ms):
        expert_sd(f"singlemodal_experts.{s}.", False)
    return sd


def make_tokens(BT: int, N: int, C: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    """Synthetic backbone tokens: N(0,1) plus a per-sample per-channel offset ~N(0,1) so that
    routing varies across the batch (SURVEY.md §8c caveat 0 / §8d config 1)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn((BT, N, C), generator=g, dtype=dtype)
    mu = torch.randn((BT, 1, C), generator=g, dtype=dtype)
    return t + mu


# backbone token shapes, Swin-V2-L @192 / HTS-AT:  stage -> (Cv, Nv, Ca, Na)
STAGE_SHAPES = {1: (192, 2304, 96, 4096), 2: (384, 576, 192, 1024), 3: (768, 144, 384, 256), 4: (1536, 36, 768, 64)}

VARIANT_DEFAULTS = {
    # launch-script hyper-parameters per variant (SURVEY.md §8a "variant deltas")
    "ave": dict(r=8, G=2, T=32, Mm=1, Ms=1, use_bn=True, use_gate=True, self_attn="none", lb_loss=False),
    "avqa": dict(r=8, G=4, T=2, Mm=1, Ms=2, use_bn=False, use_gate=False, self_attn="none", lb_loss=False),
    "avvp": dict(r=8, G=2, T=32, Mm=1, Ms=1, use_bn=False, use_gate=True, self_attn="nxn", lb_loss=True),
    "s4": dict(r=8, G=2, T=87, Mm=1, Ms=1, use_bn=False, use_gate=False, self_attn="mha", lb_loss=False),
    "ms3": dict(r=2, G=2, T=87, Mm=2, Ms=2, use_bn=False, use_gate=False, self_attn="latent", lb_loss=True),
}

VARIANT_SITES = {
    # (stage, number of adapted blocks) in execution order
    "ave": ((1, 2), (3, 6)),
    "ms3": ((1, 2), (3, 6)),
    "avqa": ((1, 2), (2, 2), (3, 6), (4, 2)),
    "avvp": ((1, 2), (2, 2), (3, 6), (4, 2)),
    "s4": ((1, 2), (2, 2), (3, 6), (4, 2)),
}


def spec_for(variant: str, stage: int, modality: str, **overrides) -> AdapterSpec:
    """AdapterSpec of the audio or visual adapter at a backbone stage for a variant."""
    Cv, Nv, Ca, Na = STAGE_SHAPES[stage]
    kw = dict(VARIANT_DEFAULTS[variant])
    kw.update(overrides)
    if modality == "audio":
        return AdapterSpec(variant=variant, C=Ca, N=Na, Cx=Cv, Nx=Nv, **kw)
    return AdapterSpec(variant=variant, C=Cv, N=Nv, Cx=Ca, Nx=Na, **kw)

"""Host-side orchestration of the AVMoE adapter forward/backward over the C-ABI kernels.

One ``torch.autograd.Function`` per MoEAdapter call.  Everything numeric happens inside the
kernel namespace ``K`` (``avmoe_b200.ops`` = CUDA kernels behind the C ABI; the tests swap in a
torch emulation of each kernel to validate this orchestration on CPU).  This file only decides
which kernels run, with which buffers, and routes gradients to the reference's parameters.

Pass structure (DESIGN.md "passes"):
  forward   resample (2 GEMMs) -> pool -> router -> latent (tokens<-kv) -> expert_front (z)
            -> [BN: stats, moments, finalize] -> expert_back (out)
  backward  [BN: expert_back_bwd_stats] -> expert_back_bwd (dz~, dprobs, d{Wu,ln_post,gate})
            -> router_bwd -> expert_front_bwd (dx, dtok', d{Wd,ln_before,gate_av})
            -> latent_bwd (dV, dtokens) -> resample_bwd (GEMMs: dO, dW_tok, dW_fc, biases)

Reference semantics: AVMOE/AVE/nets/net_trans_v3.py:376-435 (ExpertAdapter.forward),
:468-487 (MoEAdapter.forward) and the four sibling copies.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

BN_EPS = 1e-5
LN_EPS = 1e-5
BN_MOMENTUM = 0.1

# expert pre-transform kinds (what happens to x before ln_before)
PRE_NONE, PRE_LATENT, PRE_AUX_ADD, PRE_AUX_REPLACE = 0, 1, 2, 3
# batch-norm modes
BN_OFF, BN_TRAIN, BN_EVAL = 0, 1, 2


@dataclass
class AdapterConfig:
    """Static shape/flag description of one MoEAdapter (mirrors the reference ctor arguments + opt)."""
    C: int
    N: int
    Cx: int
    Nx: int
    r: int
    G: int
    T: int
    Mm: int
    Ms: int
    use_bn: bool
    use_gate: bool
    ln_before: bool
    ln_post: bool
    self_attn: str = "none"      # none | mha | latent | nxn   (single-modal experts)
    lb_loss: bool = False

    @property
    def E(self) -> int:
        return self.Mm + self.Ms

    @property
    def Cr(self) -> int:
        return self.C // self.r

    @property
    def n_latent(self) -> int:
        """number of experts that own latent tokens (multimodal + latent-self single-modal)."""
        return self.Mm + (self.Ms if self.self_attn == "latent" else 0)

    def expert_kinds(self) -> List[Tuple[int, bool]]:
        """(pre_kind, relu) per expert, multimodal first (net_trans_v3.py:482)."""
        kinds = [(PRE_LATENT, True)] * self.Mm
        um = {"none": PRE_NONE, "latent": PRE_LATENT, "nxn": PRE_AUX_ADD, "mha": PRE_AUX_REPLACE}[self.self_attn]
        return kinds + [(um, False)] * self.Ms


def param_names(cfg: AdapterConfig) -> List[str]:
    """Canonical flat ordering of the reference parameter names used by this adapter."""
    names = ["conv_adapter.weight", "conv_adapter.bias", "fc.weight", "fc.bias"]
    names += [f"router.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias")]

    def expert(p, multimodal):
        out = []
        if cfg.use_gate:
            out.append(p + "gate")
        if multimodal:
            out += [p + "my_tokens", p + "gate_av"]
        elif cfg.self_attn == "nxn":
            out.append(p + "gate_av")
        elif cfg.self_attn == "latent":
            out += [p + "my_tokens", p + "gate_self"]
        elif cfg.self_attn == "mha":
            out += [p + "self_attention.in_proj_weight", p + "self_attention.in_proj_bias",
                    p + "self_attention.out_proj.weight", p + "self_attention.out_proj.bias"]
        out += [p + "down_sampler.weight", p + "up_sampler.weight"]
        if cfg.use_bn:
            out += [p + "bn1.weight", p + "bn1.bias", p + "bn2.weight", p + "bn2.bias"]
        if cfg.ln_before:
            out += [p + "ln_before.weight", p + "ln_before.bias"]
        if cfg.ln_post:
            out += [p + "ln_post.weight", p + "ln_post.bias"]
        return out

    for m in range(cfg.Mm):
        names += expert(f"multimodal_experts.{m}.", True)
    for s in range(cfg.Ms):
        names += expert(f"singlemodal_experts.{s}.", False)
    return names


def expert_prefixes(cfg: AdapterConfig) -> List[str]:
    return [f"multimodal_experts.{m}." for m in range(cfg.Mm)] + [f"singlemodal_experts.{s}." for s in range(cfg.Ms)]


@dataclass
class ExpertPack:
    """Per-expert tensors handed to the expert kernels (all fp32, contiguous; None = feature absent)."""
    pre_kind: int
    relu: bool
    latent_slot: int = -1                    # index into tokp's expert axis
    gate_pre: Optional[torch.Tensor] = None  # gate_av / gate_self  [1]
    aux: Optional[torch.Tensor] = None       # [BT,N,C] activation dtype (PRE_AUX_*)
    ln1_w: Optional[torch.Tensor] = None
    ln1_b: Optional[torch.Tensor] = None
    wd: Optional[torch.Tensor] = None        # [Cr, C/G]
    wu: Optional[torch.Tensor] = None        # [C, Cr/G]
    ln2_w: Optional[torch.Tensor] = None
    ln2_b: Optional[torch.Tensor] = None
    gate: Optional[torch.Tensor] = None      # [1]


@dataclass
class ExpertGrads:
    """fp32 gradient accumulators the backward kernels atomically add into (zero-initialised)."""
    gate_pre: Optional[torch.Tensor] = None
    aux: Optional[torch.Tensor] = None
    ln1_w: Optional[torch.Tensor] = None
    ln1_b: Optional[torch.Tensor] = None
    wd: Optional[torch.Tensor] = None
    wu: Optional[torch.Tensor] = None
    ln2_w: Optional[torch.Tensor] = None
    ln2_b: Optional[torch.Tensor] = None
    gate: Optional[torch.Tensor] = None


# -------------------------------------------------------------------------------------------------
# resampler: vis' = fc(conv_adapter(other))          (net_trans_v3.py:469-471)
# -------------------------------------------------------------------------------------------------
def _pad_k(t: torch.Tensor) -> torch.Tensor:
    """Zero-pad the contraction (innermost) extent of a token-GEMM operand to a 16-byte row pitch: the stage-4 visual stream has
    36 tokens (72-byte bf16 rows), which a TMA tensor map cannot address; zero columns leave the product unchanged and keep the
    GEMM on the tensor cores instead of the generic fallback kernel."""
    mult = 16 // t.element_size()
    k = t.shape[-1]
    kp = -(-k // mult) * mult
    return t if kp == k else torch.nn.functional.pad(t, (0, kp - k))


def resample_fwd(K, cfg: AdapterConfig, o: torch.Tensor, w_tok, b_tok, w_fc, b_fc):
    """o [BT,Nx,Cx] -> v [BT,N,C].  Returns (v, saved) where saved holds the GEMM intermediate.

    Association is chosen by FLOPs: with Cx > C the channel projection runs first
        P = o·W_fcᵀ                      [BT,Nx,C]
        v = W_tok·P + b_tok⊗s + 1⊗b_fc   with s = rowsum(W_fc)
    otherwise the token resampling runs first
        U = W_tok·o + b_tok⊗1            [BT,N,Cx]
        v = U·W_fcᵀ + 1⊗b_fc.
    The token-resampling GEMM always runs in K-major x K-major form on the tensor cores: its activation operand is
    re-laid out channel-major ([BT,c,Nx], one bandwidth-bound transpose) because tcgen05 reads MN-major operands at
    well under half the K-major rate (profiles/r01_gemm_microbench.md)."""
    if getattr(K, "use_resample_op", lambda: False)():        # product path: the same sequence, issued from C++ in one C-ABI call
        return K.resample_fwd_op(cfg, o, w_tok, b_tok, w_fc, b_fc)
    BT = o.shape[0]
    dt = o.dtype
    N, Nx, C, Cx = cfg.N, cfg.Nx, cfg.C, cfg.Cx
    wt_l = K.prep_weight(w_tok.reshape(N, Nx), dt, False)                       # [N,Nx]  activation dtype
    wf_l = K.prep_weight(w_fc, dt, False)                                       # [C,Cx]
    fc_first = Cx > C
    if fc_first:
        P = K.gemm(o.reshape(BT * Nx, Cx), wf_l).reshape(BT, Nx, C)
        Pt = _pad_k(K.transpose(P))                                             # [BT,C,Nx(+pad)]
        s = K.rowsum(w_fc)                                                      # [C] fp32
        v = K.gemm(_pad_k(wt_l), Pt.reshape(BT * C, -1), d_nbatch=C, row_bias=b_tok, col_scale=s, col_bias=b_fc)
        return v, (fc_first, P, s)
    Ot = _pad_k(K.transpose(o))                                                 # [BT,Cx,Nx(+pad)]
    U = K.gemm(_pad_k(wt_l), Ot.reshape(BT * Cx, -1), d_nbatch=Cx, row_bias=b_tok)      # [BT,N,Cx]
    v = K.gemm(U.reshape(BT * N, Cx), wf_l, col_bias=b_fc).reshape(BT, N, C)
    return v, (fc_first, U, None)


def resample_bwd(K, cfg: AdapterConfig, o: torch.Tensor, dv: torch.Tensor, saved, w_tok, b_tok, w_fc, need_do: bool = True):
    """Gradients of resample_fwd.  dv [BT,N,C] (activation dtype).  Returns (do, dw_tok, db_tok, dw_fc, db_fc)."""
    if getattr(K, "use_resample_op", lambda: False)():
        return K.resample_bwd_op(cfg, o, dv, saved, w_tok, b_tok, w_fc, need_do)
    fc_first, mid, s = saved
    BT = o.shape[0]
    dt = o.dtype
    N, Nx, C, Cx = cfg.N, cfg.Nx, cfg.C, cfg.Cx
    f32 = torch.float32
    wt_T = K.prep_weight(w_tok.reshape(N, Nx), dt, True)                        # [Nx,N]
    wf_T = K.prep_weight(w_fc, dt, True)                                        # [Cx,C]
    if fc_first:
        P = mid
        # bias grads: db_fc[c] = sum dv ; db_tok[n] = sum_c dv*s[c] ; ds[c] = sum_n dv*b_tok[n]  (d rowsum(W_fc))
        db_tok, db_fc, ds = K.reduce_bias(dv, col_weight=s, row_weight=b_tok)
        dw_tok = K.gemm(dv, P, out_dtype=f32, split_k=K.split_k_hint(BT * C, N, Nx))   # sum_b dv[b]·P[b]ᵀ   [N,Nx]
        dVt = _pad_k(K.transpose(dv))                                           # [BT,C,N(+pad)]
        dP = K.gemm(_pad_k(wt_T), dVt.reshape(BT * C, -1), d_nbatch=C)          # W_tokᵀ·dv[b]             [BT,Nx,C]
        dP2 = dP.reshape(BT * Nx, C)
        dw_fc = K.gemm(dP2, o.reshape(BT * Nx, Cx), a_kmajor=False, b_kmajor=False, out_dtype=f32,
                       split_k=K.split_k_hint(BT * Nx, C, Cx), row_bias=ds)      # dPᵀ·o + ds⊗1             [C,Cx]
        do = K.gemm(dP2, wf_T).reshape(BT, Nx, Cx) if need_do else None         # dP·W_fc
        return do, dw_tok.reshape(N, Nx, 1, 1), db_tok, dw_fc, db_fc
    U = mid
    dv2 = dv.reshape(BT * N, C)
    _, db_fc, _ = K.reduce_bias(dv)
    dw_fc = K.gemm(dv2, U.reshape(BT * N, Cx), a_kmajor=False, b_kmajor=False, out_dtype=f32,
                   split_k=K.split_k_hint(BT * N, C, Cx))                        # dvᵀ·U                    [C,Cx]
    dU = K.gemm(dv2, wf_T).reshape(BT, N, Cx)                                    # dv·W_fc                  [BT,N,Cx]
    db_tok, _, _ = K.reduce_bias(dU)
    dw_tok = K.gemm(dU, o, out_dtype=f32, split_k=K.split_k_hint(BT * Cx, N, Nx))     # sum_b dU[b]·o[b]ᵀ        [N,Nx]
    do = None
    if need_do:
        dUt = _pad_k(K.transpose(dU))                                            # [BT,Cx,N(+pad)]
        do = K.gemm(_pad_k(wt_T), dUt.reshape(BT * Cx, -1), d_nbatch=Cx)         # W_tokᵀ·dU[b]             [BT,Nx,Cx]
    return do, dw_tok.reshape(N, Nx, 1, 1), db_tok, dw_fc, db_fc


# -------------------------------------------------------------------------------------------------
# GEMM-based latent attention (net_trans_v3.py:379-390): every contraction on the tcgen05 GEMM, streaming softmax kernels
# between them on token-major logits.  [.,T] matrices are stored with a row pitch of Tp = round8(T) (16-byte pitches for TMA).
# -------------------------------------------------------------------------------------------------
def A12_lo_offset(T: int) -> int:
    """column of the bf16 rounding-residual block inside A12 [.., 2*Tp]: Tp = T rounded up to 8"""
    return (T + 7) // 8 * 8


def gemm_latent_fwd(K, kv: torch.Tensor, tok: torch.Tensor):
    """tok' = tok + softmax_Nk(tok·kvᵀ)·kv.   kv [BT,Nk,C], tok [T,C] fp32 -> (tok' fp32 [BT,T,C], K.split_operand(tok'),
    A12) where A12 [BT,Nk,2Tp] holds A1ᵀ (probabilities, token-major) in columns [0,T) and is reused by the backward.
    The fp32 latent tokens enter the logits GEMMs as split operands (K.split_operand)."""
    BT, Nk, C = kv.shape
    T = tok.shape[0]
    dt = kv.dtype
    if getattr(K, "rf_tok_supported", lambda *a: False)(kv, T):      # logits recomputed per tile, never stored (csrc/rf_tokens.cu)
        A12, p_lo = K.rf_tok_probs(kv, tok), A12_lo_offset(T)
    else:
        tok_s, off, _ = K.split_operand(tok, dt)                                         # [W,C]
        S1 = K.gemm(kv.reshape(BT * Nk, C), tok_s, out_dtype=torch.float32)              # token-major logits kv_n · tok_t  [BT*Nk,W]
        A12, _, p_lo = K.softmax_cols(S1.reshape(BT, Nk, -1), T, off, dt)
    A1 = A12 if p_lo else A12[:, :, :T]                                                  # bf16: hi and residual column blocks
    R = K.bgemm(A1, kv, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32)         # A1·kv   [BT,T (or 2Tp),C]
    tokp = K.tile_add(R, tok, p_lo)
    return tokp, K.split_operand(tokp, dt), A12


def gemm_attn2_fwd(K, x: torch.Tensor, tokp_op, T: int):
    """x_res = softmax_T(x·tok'ᵀ)·tok'.   Returns (x_res [BT,N,C] in x.dtype, A2 [BT,N,Tp])."""
    tokp_s, off, tokp_lp = tokp_op
    S2 = K.bgemm(x, tokp_s, out_dtype=torch.float32)                                     # x[b]·tok'[b]ᵀ   [BT,N,W]
    A2 = K.softmax_rows(S2, T, off, x.dtype)
    XR = K.bgemm(A2[:, :, :T], tokp_lp, b_kmajor=False)                                  # A2·tok'   [BT,N,C]
    return XR, A2


def gemm_attn2_bwd(K, x, tokp_op, A2, dXR, T: int, dx):
    """Backward of gemm_attn2_fwd given dXR [BT,N,C]: adds the gradient through the logits into dx, returns d tok' fp32."""
    f32 = torch.float32
    tokp_s, off, tokp_lp = tokp_op
    dA2 = K.bgemm(dXR, tokp_s, out_dtype=f32)                                            # dXR[b]·tok'[b]ᵀ  [BT,N,W]
    dS2 = K.softmax_rows_bwd(A2, dA2, T, off)
    K.bgemm(dS2[:, :, :T], tokp_lp, b_kmajor=False, out=dx, accumulate=True)             # dx += dS2·tok'  (in the epilogue)
    dtokp = K.bgemm(A2[:, :, :T], dXR, a_kmajor=False, b_kmajor=False, out_dtype=f32)    # A2ᵀ·dXR   [BT,T,C]
    K.add_into(dtokp, K.bgemm(dS2[:, :, :T], x, a_kmajor=False, b_kmajor=False, out_dtype=f32))   # + dS2ᵀ·x
    return dtokp


def gemm_latent_bwd(K, kv, tok, tokp, A12, dtokp, T: int, dkv_bias=None):
    """Backward of gemm_latent_fwd given d tok' [BT,T,C] fp32.  Returns (dkv [BT,Nk,C] in kv.dtype, d tok [T,C] fp32).
    dkv_bias [BT,C] fp32 (the router-path gradient of mean_N(kv)) is added to every row of dkv in the GEMM epilogue."""
    BT, Nk, C = kv.shape
    Tp = A12.shape[2] // 2
    dt = kv.dtype
    f32 = torch.float32
    Dt = K.latent_dots(dtokp, tokp, tok)                                                 # sum_c dR·R  [BT,T]
    dR_s, off, dR_lp = K.split_operand(dtokp, dt)                                              # [BT,W,C]
    if getattr(K, "rf_tok_supported", lambda *a: False)(kv, T):
        K.rf_tok_ds(kv, dtokp, A12, Dt, T)                                               # dA1 recomputed per tile; A12[:, :, Tp:Tp+T] = dS1ᵀ
    else:
        dA1 = K.bgemm(kv, dR_s, out_dtype=f32)                                           # kv[b]·dR[b]ᵀ  (token-major)  [BT,Nk,W]
        K.softmax_cols_bwd(A12, dA1, Dt, T, off)                                         # A12[:, :, Tp:Tp+T] = dS1ᵀ
    B12 = (torch.empty if Tp == T else torch.zeros)((BT, 2 * Tp, C), device=kv.device, dtype=dt)   # [dR ; tok] stacked along K (padding rows zero)
    B12[:, :T] = dR_lp
    B12[:, Tp:Tp + T] = K.cast_to(tok, dt)
    dkv = K.bgemm(A12, B12, b_kmajor=False, col_bias=dkv_bias)                           # A1ᵀ·dR + dS1ᵀ·tok (+ bias)  [BT,Nk,C]
    dS1_flat = A12.reshape(BT * Nk, 2 * Tp)[:, Tp:Tp + T]                                # [K=(b,n), M=T]
    dtok = K.gemm(dS1_flat, kv.reshape(BT * Nk, C), a_kmajor=False, b_kmajor=False, out_dtype=f32,
                  split_k=K.split_k_hint(BT * Nk, T, C))                                 # sum_{b,n} dS1ᵀ·kv   [T,C]
    K.batch_sum_add(dtokp, dtok)                                                         # d tok'/d tok = I
    return dkv, dtok


def _bn_params(P, bufs, prefixes):
    bn_w = [[P[p + f"bn{i}.{w}"] for i in (1, 2) for w in ("weight", "bias")] for p in prefixes]
    bn_b = [[bufs[p + f"bn{i}.{w}"] for i in (1, 2) for w in ("running_mean", "running_var")] for p in prefixes]
    return bn_w, bn_b


def fused_experts_fwd(K, cfg, x, tokp, packs, probs, P, bufs, prefixes, bn_mode):
    """Experts through the fused tile kernels (csrc/expert.cu): returns (out, z, bn)."""
    BT, N, C = x.shape
    E, Cr = cfg.E, cfg.Cr
    dev, f32 = x.device, torch.float32
    # 5. expert front: pre-transform, ln_before, grouped down projection -> z (pre-BN), BN1 batch sums
    z = torch.empty((BT, N, E, Cr), device=dev, dtype=f32)
    stats1 = torch.zeros((E, Cr, 2), device=dev, dtype=torch.float64) if bn_mode == BN_TRAIN else None
    K.expert_front_fwd(cfg, x, tokp, packs, z, stats1)
    # 6. BatchNorm statistics -> per-channel affine + (mean, invstd) for both norms
    bn = None
    if bn_mode != BN_OFF:
        bn_w = [[P[p + f"bn{i}.{w}"] for i in (1, 2) for w in ("weight", "bias")] for p in prefixes]
        bn_b = [[bufs[p + f"bn{i}.{w}"] for i in (1, 2) for w in ("running_mean", "running_var")] for p in prefixes]
        moments = None
        if bn_mode == BN_TRAIN:
            aff1, save1 = K.bn1_finalize(cfg, stats1, BT * N, bn_w, bn_b, True)
            moments = K.bn_moments(cfg, z, aff1, packs)                       # sums of a and a aᵀ per group (float64)
        else:
            aff1, save1 = K.bn1_finalize(cfg, None, BT * N, bn_w, bn_b, False)
        aff2, save2 = K.bn2_finalize(cfg, moments, BT * N, packs, bn_w, bn_b, bn_mode == BN_TRAIN)
        if bn_mode == BN_TRAIN:
            for p in prefixes:
                bufs[p + "bn1.num_batches_tracked"] += 1
                bufs[p + "bn2.num_batches_tracked"] += 1
        bn = (aff1, save1, aff2, save2)
    # 7. expert back: activation, up projection, bn2, ln_post, gate, probability-weighted combine
    out = torch.empty_like(x)
    K.expert_back_fwd(cfg, z, bn, packs, probs, out)
    return out, z, bn


def gemm_experts_fwd(K, cfg, x, packs, probs, P, bufs, prefixes, bn_mode, res=None, front_toks=None):
    """Experts with every channel contraction on the tcgen05 GEMM (block-diagonal dense weights) and streaming row kernels
    in between (csrc/expert_rows.cu).  Returns (out, z, bn, saved-for-backward).

    front_toks (dict expert -> packed tok' operands, possibly empty) selects the row-fused front pass (csrc/rowfused.cu): the
    latent attention of each row, gate_pre residual, ln_before and the down projection run on one 16-row tile per warp and only
    z is written (experts of pre_kind PRE_LATENT are evaluated in place; nothing is saved, the backward recomputes)."""
    BT, N, C = x.shape
    E, Cr = cfg.E, cfg.Cr
    rows = BT * N
    dev, f32, dt = x.device, torch.float32, x.dtype
    z = torch.empty((BT, N, E, Cr), device=dev, dtype=f32)
    front = None
    if front_toks is not None:
        prep = K.rf_prep_front(cfg, packs)
        ln1_saved = K.rf_front_fwd(cfg, x, packs, front_toks, prep, z)
        xl = ln1s = wdt = None
        front = (front_toks, prep, ln1_saved)
    else:
        wdd, wdt = K.xg_prep_weights(cfg, packs, dt)                              # [E,Cr,C], [E,C,Crp]
        xl, ln1s = K.xg_ln_pre_fwd(cfg, x, packs)                                 # [E,rows,C]
        K.bgemm(xl, wdd, out=z.view(rows, E, Cr).permute(1, 0, 2))                # down projection, all experts in one launch
    bn = bn2_wb = suma = None
    if bn_mode != BN_OFF:
        bn_w, bn_b = _bn_params(P, bufs, prefixes)
        moments = None
        if bn_mode == BN_TRAIN:
            stats1 = K.xg_z_stats(cfg, z)
            aff1, save1 = K.bn1_finalize(cfg, stats1, rows, bn_w, bn_b, True)
            moments = K.bn_moments(cfg, z, aff1, packs)
            suma = moments[0]
        else:
            aff1, save1 = K.bn1_finalize(cfg, None, rows, bn_w, bn_b, False)
        aff2, save2 = K.bn2_finalize(cfg, moments, rows, packs, bn_w, bn_b, bn_mode == BN_TRAIN)
        if bn_mode == BN_TRAIN:    # one multi-tensor launch instead of 2E tiny ones
            torch._foreach_add_([bufs[p + f"bn{i}.num_batches_tracked"] for p in prefixes for i in (1, 2)], 1)
        bn = (aff1, save1, aff2, save2)
        bn2_wb = [(P[p + "bn2.weight"], P[p + "bn2.bias"]) for p in prefixes]
    if bool(getattr(K, "USE_ROWFUSED", False)) and K.rf_supported(cfg, dt):
        # row-fused back pass: bn1 affine, activation, up projection, bn2, ln_post, gate, combine (+ residual) on one 16-row tile
        # per warp; nothing is materialised -- the backward recomputes from z
        wu_pack, ybi = K.rf_prep_up(cfg, packs, bn[3] if bn else None, bn_mode)
        out = torch.empty_like(x)
        K.rf_back_fwd(cfg, z, bn[0] if bn else None, wu_pack, ybi, bn2_wb, bn_mode, packs, probs, out, res)
        return out, z, bn, (xl, ln1s, None, None, None, wdt, suma, (wu_pack, ybi), front)
    act = K.xg_bn_act(cfg, z, bn[0] if bn else None, packs, dt, bn_mode)          # [E,rows,Crp]
    wud, ybi = K.xg_prep_up(cfg, packs, bn[3] if bn else None, bn_mode, dt)       # bn2 normalisation folded into the operand
    yd = torch.empty((E, rows, 2 * C), device=dev, dtype=dt)
    K.bgemm(act, wud, out=yd[:, :, :C], col_bias=ybi)                             # up projection -> normalised y (per-expert bias)
    out = torch.empty_like(x)
    ln2s = K.xg_post_fwd(cfg, yd, bn2_wb, packs, probs, out, res)                    # residual add folded in (row a9)
    return out, z, bn, (xl, ln1s, act, yd, ln2s, wdt, suma, None, front)


def gemm_experts_bwd(K, cfg, x, dout, z, bn, bn_mode, packs, egr, probs, P, prefixes, saved):
    """Backward of gemm_experts_fwd up to the ln_before input: returns (dprobs, red1, red2, dxl, ln1s, pending).
    With the row-fused front pass, `pending` carries what rf_front_finish needs once the router gradient is known."""
    BT, N, C = x.shape
    E, Cr = cfg.E, cfg.Cr
    Crp = K.pad_bottleneck(Cr)
    rows = BT * N
    dev, f32, dt = x.device, torch.float32, x.dtype
    xl, ln1s, act, yd, ln2s, wdt, suma, rf, front = saved
    bn2_wb = [(P[p + "bn2.weight"], P[p + "bn2.bias"]) for p in prefixes] if bn_mode != BN_OFF else None
    red2 = torch.zeros((E, C, 2), device=dev, dtype=torch.float64) if bn_mode != BN_OFF else None
    red1 = torch.zeros((E, Cr, 2), device=dev, dtype=torch.float64) if bn_mode != BN_OFF else None
    dprobs = torch.zeros((BT, E), device=dev, dtype=f32)
    if rf is not None:     # recompute yhat / a from z (row-fused): yd = [yhat | d(bn2 output)], a
        yd, act = K.rf_back_bwd(cfg, dout, z, bn[0] if bn else None, rf[0], rf[1], bn2_wb, bn_mode, packs, probs, dprobs, red2, egr)
    else:
        K.xg_post_bwd(cfg, dout, yd, ln2s, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr)   # yd[:, :, C:] = d(bn2 output)
    w2, bias2 = K.xg_prep_bwd(cfg, packs, bn2_wb, bn, bn_mode, red2, BT, dt)               # bn2 backward folded into the operand
    da = torch.empty((E, rows, Crp), device=dev, dtype=f32)
    pwu = torch.empty((E, 2 * C, Crp), device=dev, dtype=f32)
    K.bgemm(yd, w2, out=da, col_bias=bias2)                                                # d a = dy · Wu   (all experts)
    for e in range(E):
        K.gemm(yd[e], act[e], a_kmajor=False, b_kmajor=False, out=pwu[e], split_k=K.split_k_hint(rows, 2 * C, Crp))
    dz = K.xg_bn1_bwd(cfg, da, z, bn, bn_mode, packs, red1, dt)                            # relu mask + bn1 backward
    if front is not None:
        return dprobs, red1, red2, None, None, (dz, pwu, bn2_wb, suma, front)
    dxl = torch.empty((E, rows, C), device=dev, dtype=dt)
    pwd = torch.empty((E, Crp, C), device=dev, dtype=f32)
    K.bgemm(dz, wdt, out=dxl)                                                              # d xl = dz · Wd  (all experts)
    for e in range(E):
        K.gemm(dz[e], xl[e], a_kmajor=False, b_kmajor=False, out=pwd[e], split_k=K.split_k_hint(rows, Crp, C))
    K.xg_finish_wgrads(cfg, packs, egr, bn2_wb, bn, bn_mode, red2, BT, pwd, pwu, suma)
    return dprobs, red1, red2, dxl, ln1s, None


def rf_front_finish(K, cfg, x, packs, egr, bn, bn_mode, red2, pending, dpool_x, dx, dres=None):
    """Second half of the expert backward on the row-fused path: the front pass backward (dx, xl, d aux / dXR, A2 | dS2), then the
    down-projection weight-gradient GEMMs on the re-emitted xl and the grouped weight-gradient finish.  Returns {expert: (A2, dS2)}."""
    dz, pwu, bn2_wb, suma, (toks, prep, ln1_saved) = pending
    BT, N, C = x.shape
    rows = BT * N
    Crp = K.pad_bottleneck(cfg.Cr)
    xl, lat = K.rf_front_bwd(cfg, x, packs, egr, toks, prep, ln1_saved, dz, dpool_x, dx, dres)
    pwd = torch.empty((cfg.E, Crp, C), device=x.device, dtype=torch.float32)
    for e in range(cfg.E):
        K.gemm(dz[e], xl[e], a_kmajor=False, b_kmajor=False, out=pwd[e], split_k=K.split_k_hint(rows, Crp, C))
    K.xg_finish_wgrads(cfg, packs, egr, bn2_wb, bn, bn_mode, red2, BT, pwd, pwu, suma)
    return lat


# -------------------------------------------------------------------------------------------------
# the fused adapter function
# -------------------------------------------------------------------------------------------------
def _bn_mode(cfg: AdapterConfig, training: bool) -> int:
    if not cfg.use_bn:
        return BN_OFF
    return BN_TRAIN if training else BN_EVAL


class MoEAdapterFunction(torch.autograd.Function):
    """(x, o, *params) -> (out [BT,N,C], probs [BT,E], lb_loss [])  with expert_indices as a non-differentiable extra.

    ``state`` carries non-tensor context: cfg, kernel namespace K, training flag, the BatchNorm buffers (updated in
    place in training mode, exactly like nn.BatchNorm2d), the optional router noise tensor, and receives ``idx``."""

    @staticmethod
    def forward(ctx, state: dict, x: torch.Tensor, o: torch.Tensor, res: Optional[torch.Tensor], *params: torch.Tensor):
        # launch on the device that owns the tensors (nn.DataParallel replicas, AVMOE/AVVP/main.py:420-421, call this from one
        # thread per GPU; autograd engine threads likewise)
        with torch.cuda.device_of(x):
            return MoEAdapterFunction._forward(ctx, state, x, o, res, *params)

    @staticmethod
    def _forward(ctx, state: dict, x: torch.Tensor, o: torch.Tensor, res: Optional[torch.Tensor], *params: torch.Tensor):
        cfg: AdapterConfig = state["cfg"]
        K = state["K"]
        training: bool = state["training"]
        names = param_names(cfg)
        assert len(params) == len(names), (len(params), len(names))
        P: Dict[str, torch.Tensor] = dict(zip(names, params))
        bufs: Dict[str, torch.Tensor] = state.get("buffers", {})
        BT, N, C = x.shape
        assert (N, C) == (cfg.N, cfg.C) and o.shape[1:] == (cfg.Nx, cfg.Cx), "token shapes do not match the adapter"
        assert x.is_contiguous() and o.is_contiguous()
        assert res is None or (res.shape == x.shape and res.is_contiguous() and res.dtype == x.dtype)
        E, Cr, G, T = cfg.E, cfg.Cr, cfg.G, cfg.T
        bn_mode = _bn_mode(cfg, training)
        f32 = torch.float32
        dev = x.device

        # 1. resampler
        v, rs_saved = resample_fwd(K, cfg, o, P["conv_adapter.weight"], P["conv_adapter.bias"], P["fc.weight"], P["fc.bias"])
        # 2. router
        pool_x, pool_v = K.pool(x), K.pool(v)                                    # [BT,chunks,C] fp32 partial sums
        rw = [P[f"router.{i}.{w}"] for i in (0, 2, 4) for w in ("weight", "bias")]
        probs, idx, h1, h2, rin = K.router_fwd(pool_x, pool_v, 1.0 / N, rw, state.get("noise"))
        # 3. per-expert pre-transforms that are separate kernels
        prefixes = expert_prefixes(cfg)
        kinds = cfg.expert_kinds()
        packs: List[ExpertPack] = []
        latent_tokens, latent_kv_is_x = [], []
        aux_saved = {}
        aux_kind: Dict[int, str] = {}
        gate_keys: List[Optional[str]] = []
        # bf16 streams run the latent attention on the tensor-core GEMM.  fp32 streams by default keep the fused kernels whose
        # contractions are exact fp32 FMAs, like the reference's torch.bmm (matmul TF32 is off by default in torch);
        # "always" forces the GEMM path for either dtype (tests).
        # fp32 streams follow torch's own switch: with torch.backends.cuda.matmul.allow_tf32 they take the same GEMM orchestration
        # with TF32 operands (what the reference's bmm/linear would do under that flag; ~10x faster than the exact kernels).
        _flag = getattr(K, "USE_GEMM_ATTENTION", False)
        on_tc = x.dtype == torch.bfloat16 or (x.dtype == torch.float32 and bool(getattr(K, "fp32_on_tensor_cores", lambda: False)()))
        use_gemm_attn = _flag == "always" or (bool(_flag) and on_tc)
        _xflag = getattr(K, "USE_GEMM_EXPERTS", False)
        xg_wanted = _xflag == "always" or (bool(_xflag) and on_tc)
        # row-fused front pass (csrc/rowfused.cu): the row <- latent-token attention runs inside the expert front kernel
        rf_front = (use_gemm_attn and xg_wanted and bool(getattr(K, "USE_ROWFUSED", False)) and K.rf_supported(cfg, x.dtype)
                    and cfg.N % 16 == 0)
        front_toks = {} if rf_front else None
        for e, (pfx, (pre, relu)) in enumerate(zip(prefixes, kinds)):
            pk = ExpertPack(pre_kind=pre, relu=relu)
            multimodal = e < cfg.Mm
            gate_key = None
            if pre == PRE_LATENT:
                gate_key = "gate_av" if multimodal else "gate_self"
                pk.gate_pre = P[pfx + gate_key]
                if use_gemm_attn:   # latent attention on the tensor-core GEMM; the expert kernels see x + gate*x_res as an aux-add
                    kv = v if multimodal else x
                    tokp_e, tokp_op, A12 = gemm_latent_fwd(K, kv, P[pfx + "my_tokens"])
                    if rf_front:
                        front_toks[e] = K.rf_pack_tokens(tokp_e)
                        aux_saved[e] = (not multimodal, tokp_e, None, A12, None)
                        aux_kind[e] = "latent_rf"
                    else:
                        pk.aux, A2 = gemm_attn2_fwd(K, x, tokp_op, T)
                        pk.pre_kind = PRE_AUX_ADD
                        aux_saved[e] = (not multimodal, tokp_e, tokp_op, A12, A2)
                        aux_kind[e] = "latent_gemm"
                else:
                    pk.latent_slot = len(latent_tokens)
                    latent_tokens.append(P[pfx + "my_tokens"])
                    latent_kv_is_x.append(not multimodal)
            elif pre == PRE_AUX_ADD:      # AVVP NxN token self-attention (mgn.py:132-139)
                pk.aux, lse_rows = K.selfattn_nxn_fwd(x)
                aux_saved[e] = lse_rows
                aux_kind[e] = "nxn"
                gate_key = "gate_av"
                pk.gate_pre = P[pfx + gate_key]
            elif pre == PRE_AUX_REPLACE:  # AVS v1 MHA over the BT axis (PVT_AVSModel_v2.py:212-215)
                mp = [P[pfx + "self_attention." + n] for n in ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")]
                # attention-probability dropout of nn.MultiheadAttention(dropout=0.2) in training (PVT_AVSModel_v2.py:139-142): a keep
                # mask drawn from torch's generator where the reference's dropout draws, or injected by the caller (parity runs)
                keep, p_drop = None, float(state.get("mha_dropout", 0.0))
                inj = state.get("mha_keep")
                if inj is not None:
                    keep = inj[e - cfg.Mm].to(device=dev, dtype=torch.uint8).contiguous()
                elif training and p_drop > 0.0:
                    keep = (torch.rand((N, 4, BT, BT), device=dev) >= p_drop).to(torch.uint8)
                pk.aux, mha_saved = K.mha_fwd(x, mp, keep, 1.0 / (1.0 - p_drop) if keep is not None else 1.0)
                aux_saved[e] = mha_saved
                aux_kind[e] = "mha"
            gate_keys.append(gate_key)
            if cfg.ln_before:
                pk.ln1_w, pk.ln1_b = P[pfx + "ln_before.weight"], P[pfx + "ln_before.bias"]
            if cfg.ln_post:
                pk.ln2_w, pk.ln2_b = P[pfx + "ln_post.weight"], P[pfx + "ln_post.bias"]
            pk.wd = P[pfx + "down_sampler.weight"].reshape(Cr, C // G)
            pk.wu = P[pfx + "up_sampler.weight"].reshape(C, Cr // G)
            if cfg.use_gate:
                pk.gate = P[pfx + "gate"]
            packs.append(pk)
        # 4. latent tokens attend to their key/value stream (v for multimodal, x for latent-self)
        tokp = lse = None
        nl = len(latent_tokens)
        if nl:
            tokp = torch.empty((BT, nl, T, C), device=dev, dtype=f32)
            lse = torch.empty((BT, nl, T), device=dev, dtype=f32)
            for kv_is_x in (False, True):
                slots = [i for i in range(nl) if latent_kv_is_x[i] == kv_is_x]
                if slots:
                    K.latent_fwd(x if kv_is_x else v, [latent_tokens[i] for i in slots], slots, tokp, lse)
        # bf16 streams run the expert projections on the tensor-core GEMM with streaming row kernels in between
        # ("gemm experts"); fp32 streams keep the fused kernels (exact fp32 contractions).  "always" forces it (tests).
        use_xg = xg_wanted and (rf_front or all(pk.pre_kind != PRE_LATENT for pk in packs))
        xg_saved = None
        if use_xg:
            out, z, bn, xg_saved = gemm_experts_fwd(K, cfg, x, packs, probs, P, bufs, prefixes, bn_mode, res, front_toks)
        else:
            out, z, bn = fused_experts_fwd(K, cfg, x, tokp, packs, probs, P, bufs, prefixes, bn_mode)
            if res is not None:
                K.add_into(out, res)
        lb = K.lb_loss_fwd(probs) if cfg.lb_loss else torch.zeros((), device=dev, dtype=f32)

        state["idx"] = idx
        ctx.state = state
        ctx.names = names
        ctx.bn_mode = bn_mode
        ctx.rs_saved = rs_saved
        ctx.aux_saved = aux_saved
        ctx.aux_kind = aux_kind
        ctx.gate_keys = gate_keys
        ctx.latent_kv_is_x = latent_kv_is_x
        ctx.packs_meta = [(pk.pre_kind, pk.relu, pk.latent_slot) for pk in packs]
        ctx.bn = bn
        ctx.xg_saved = xg_saved
        ctx.aux = [pk.aux for pk in packs]
        ctx.save_for_backward(x, o, v, rin, probs, h1, h2, tokp, lse, z, *params)
        ctx.mark_non_differentiable(idx)
        return out, probs, lb, idx

    @staticmethod
    def backward(ctx, dout, dprobs_ext, dlb, _didx):
        with torch.cuda.device_of(dout):
            grads = MoEAdapterFunction._backward(ctx, dout, dprobs_ext, dlb, _didx)
        # the large forward intermediates must not outlive the backward (a training loop that keeps `loss` alive into the
        # next forward would otherwise hold two steps of activations)
        ctx.rs_saved = ctx.aux_saved = ctx.xg_saved = ctx.bn = ctx.aux = None
        return grads

    @staticmethod
    def _backward(ctx, dout, dprobs_ext, dlb, _didx):
        state = ctx.state
        cfg: AdapterConfig = state["cfg"]
        K = state["K"]
        names = ctx.names
        x, o, v, rin, probs, h1, h2, tokp, lse, z, *params = ctx.saved_tensors
        P: Dict[str, torch.Tensor] = dict(zip(names, params))
        BT, N, C = x.shape
        E, Cr, G, T = cfg.E, cfg.Cr, cfg.G, cfg.T
        f32 = torch.float32
        dev = x.device
        bn_mode = ctx.bn_mode
        prefixes = expert_prefixes(cfg)
        grads: Dict[str, Optional[torch.Tensor]] = {n: None for n in names}
        dout = dout.contiguous()

        # rebuild packs (parameters come from saved tensors); the small fp32 gradient accumulators of all experts are views of
        # ONE zero-filled arena (a single fill instead of ~9 per expert)
        packs: List[ExpertPack] = []
        egr: List[ExpertGrads] = []
        arena_req = []          # (ExpertGrads, field, shape, offset)
        arena_len = 0

        def want(g, field, shape):
            nonlocal arena_len
            n = 1
            for d in shape:
                n *= d
            arena_req.append((g, field, shape, arena_len, n))
            arena_len += (n + 3) // 4 * 4

        for e, (pfx, (pre, relu, slot)) in enumerate(zip(prefixes, ctx.packs_meta)):
            multimodal = e < cfg.Mm
            pk = ExpertPack(pre_kind=pre, relu=relu, latent_slot=slot, aux=ctx.aux[e])
            g = ExpertGrads()
            if ctx.gate_keys[e] is not None:
                pk.gate_pre = P[pfx + ctx.gate_keys[e]]
            if pk.gate_pre is not None:
                want(g, "gate_pre", (1,))
            if pre in (PRE_AUX_ADD, PRE_AUX_REPLACE) or ctx.aux_kind.get(e) == "latent_rf":     # latent_rf: receives dXR
                g.aux = torch.empty((BT, N, C), device=dev, dtype=x.dtype)
            if cfg.ln_before:
                pk.ln1_w, pk.ln1_b = P[pfx + "ln_before.weight"], P[pfx + "ln_before.bias"]
                want(g, "ln1_w", (C,)); want(g, "ln1_b", (C,))
            if cfg.ln_post:
                pk.ln2_w, pk.ln2_b = P[pfx + "ln_post.weight"], P[pfx + "ln_post.bias"]
                want(g, "ln2_w", (C,)); want(g, "ln2_b", (C,))
            pk.wd = P[pfx + "down_sampler.weight"].reshape(Cr, C // G)
            pk.wu = P[pfx + "up_sampler.weight"].reshape(C, Cr // G)
            want(g, "wd", (Cr, C // G)); want(g, "wu", (C, Cr // G))
            if cfg.use_gate:
                pk.gate = P[pfx + "gate"]
                want(g, "gate", (1,))
            packs.append(pk)
            egr.append(g)
        arena = torch.zeros(arena_len, device=dev, dtype=f32)
        for g, field, shape, off, n in arena_req:
            setattr(g, field, arena[off:off + n].view(shape))

        # 1. expert back (BN2 batch reductions first when training with BN)
        use_xg = ctx.xg_saved is not None
        if use_xg:
            dprobs, red1, red2, dxl, ln1s, pending = gemm_experts_bwd(K, cfg, x, dout, z, ctx.bn, bn_mode, packs, egr, probs, P, prefixes,
                                                                      ctx.xg_saved)
            red2b = red2
        else:
            red2 = None
            if bn_mode == BN_TRAIN:
                red2 = torch.zeros((E, C, 2), device=dev, dtype=torch.float64)
                K.expert_back_bwd_stats(cfg, dout, z, ctx.bn, packs, probs, red2)
            dzt = torch.empty((BT, N, E, Cr), device=dev, dtype=f32)
            dprobs = torch.zeros((BT, E), device=dev, dtype=f32)
            red1 = torch.zeros((E, Cr, 2), device=dev, dtype=torch.float64) if bn_mode != BN_OFF else None
            red2b = torch.zeros((E, C, 2), device=dev, dtype=torch.float64) if bn_mode == BN_EVAL else red2
            K.expert_back_bwd(cfg, dout, z, ctx.bn, bn_mode, red2, red2b, packs, probs, dzt, dprobs, red1, egr)
        # 2. router backward (dprobs from the combine + external grads on gating_probs + lb-loss)
        if dprobs_ext is not None:
            dprobs = dprobs + dprobs_ext.to(f32)
        if cfg.lb_loss and dlb is not None:
            K.lb_loss_bwd(probs, dlb.to(f32), dprobs)                            # accumulates into dprobs
        rw = [P[f"router.{i}.{w}"] for i in (0, 2, 4) for w in ("weight", "bias")]
        dpool_x, dpool_v, drw = K.router_bwd(dprobs, probs, h1, h2, rin, 1.0 / N, rw)
        for (i, w), g in zip([(i, w) for i in (0, 2, 4) for w in ("weight", "bias")], drw):
            grads[f"router.{i}.{w}"] = g
        # 3. expert front backward
        nl = 0 if tokp is None else tokp.shape[1]
        dtokp = torch.zeros((BT, nl, T, C), device=dev, dtype=f32) if nl else None
        dx = torch.empty_like(x)
        lat_rf = {}
        # residual stream == x (the reference call sites: f = f + adapter(f, other)): its gradient is dout itself; folded into dx here
        # so autograd does not run a separate accumulation pass over the token stream
        fold_res = bool(state.get("res_is_x")) and ctx.needs_input_grad[3]
        res_folded = False
        if use_xg and pending is not None:
            lat_rf = rf_front_finish(K, cfg, x, packs, egr, ctx.bn, bn_mode, red2, pending, dpool_x, dx, dout if fold_res else None)
            res_folded = fold_res
        elif use_xg:
            K.xg_ln_pre_bwd(cfg, x, dxl, ln1s, packs, dpool_x, dx, egr)
        else:
            K.expert_front_bwd(cfg, x, tokp, dzt, z, ctx.bn, bn_mode, red1, packs, dpool_x, dx, dtokp, egr)
        # 4. separate pre-transform kernels (AVVP NxN attention, AVS-v1 MHA) and the GEMM-based latent attention
        dv = None
        for e, (pfx, pk) in enumerate(zip(prefixes, packs)):
            kind = ctx.aux_kind.get(e)
            if kind == "nxn":
                K.selfattn_nxn_bwd(x, ctx.aux_saved[e], egr[e].aux, dx)          # accumulates into dx
            elif kind == "mha":
                mp = [P[pfx + "self_attention." + n] for n in ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")]
                dmp = K.mha_bwd(x, mp, ctx.aux_saved[e], egr[e].aux, dx)         # accumulates into dx
                for n, g in zip(("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"), dmp):
                    grads[pfx + "self_attention." + n] = g
            elif kind in ("latent_gemm", "latent_rf"):
                kv_is_x, tokp_e, tokp_op, A12, A2 = ctx.aux_saved[e]
                if kind == "latent_rf":     # the front backward already produced dS2 and added dS2·tok' into dx
                    a2, ds2 = lat_rf[e]
                    dtokp_e = K.bgemm(a2[:, :, :T], egr[e].aux, a_kmajor=False, b_kmajor=False, out_dtype=f32)      # A2ᵀ·dXR
                    K.add_into(dtokp_e, K.bgemm(ds2[:, :, :T], x, a_kmajor=False, b_kmajor=False, out_dtype=f32))   # + dS2ᵀ·x
                else:
                    dtokp_e = gemm_attn2_bwd(K, x, tokp_op, A2, egr[e].aux, T, dx)   # adds into dx
                fold_pool = not kv_is_x and dv is None                           # router path d mean_N(v): once, in the epilogue
                dkv, dtok = gemm_latent_bwd(K, x if kv_is_x else v, P[pfx + "my_tokens"], tokp_e, A12, dtokp_e, T,
                                            dkv_bias=dpool_v if fold_pool else None)
                grads[pfx + "my_tokens"] = dtok
                if kv_is_x:
                    K.add_into(dx, dkv)
                elif dv is None:
                    dv = dkv
                else:
                    K.add_into(dv, dkv)
        # 5. fused latent attention backward (when the GEMM path is off) -> dv (and dx for latent-self), d my_tokens
        if nl:
            dtokens = [torch.zeros((T, C), device=dev, dtype=f32) for _ in range(nl)]
            toks = [None] * nl
            for e, (pfx, pk) in enumerate(zip(prefixes, packs)):
                if pk.pre_kind == PRE_LATENT:
                    toks[pk.latent_slot] = P[pfx + "my_tokens"]
            for kv_is_x in (False, True):
                slots = [i for i in range(nl) if ctx.latent_kv_is_x[i] == kv_is_x]
                if not slots:
                    continue
                if kv_is_x:
                    K.latent_bwd(x, [toks[i] for i in slots], slots, tokp, lse, dtokp, None, dx, True, [dtokens[i] for i in slots])
                else:
                    dv = torch.empty_like(v)
                    K.latent_bwd(v, [toks[i] for i in slots], slots, tokp, lse, dtokp, dpool_v, dv, False, [dtokens[i] for i in slots])
            for e, (pfx, pk) in enumerate(zip(prefixes, packs)):
                if pk.pre_kind == PRE_LATENT:
                    grads[pfx + "my_tokens"] = dtokens[pk.latent_slot]
        if dv is None:   # no multimodal expert: v only feeds the router
            dv = K.broadcast_rows(dpool_v, N, v.dtype)
        # 6. resampler backward
        need_do = ctx.needs_input_grad[2]
        do, dw_tok, db_tok, dw_fc, db_fc = resample_bwd(K, cfg, o, dv, ctx.rs_saved, P["conv_adapter.weight"],
                                                        P["conv_adapter.bias"], P["fc.weight"], need_do)
        grads["conv_adapter.weight"], grads["conv_adapter.bias"] = dw_tok, db_tok
        grads["fc.weight"], grads["fc.bias"] = dw_fc, db_fc
        # 7. scatter expert gradients to the reference parameter names
        if bn_mode != BN_OFF:
            r1 = red1.to(f32)
            r2 = red2b.to(f32)
        for e, (pfx, g) in enumerate(zip(prefixes, egr)):
            multimodal = e < cfg.Mm
            if g.gate is not None:
                grads[pfx + "gate"] = g.gate
            if g.gate_pre is not None:
                grads[pfx + ctx.gate_keys[e]] = g.gate_pre
            grads[pfx + "down_sampler.weight"] = g.wd.reshape(Cr, C // G, 1, 1)
            grads[pfx + "up_sampler.weight"] = g.wu.reshape(C, Cr // G, 1, 1)
            if cfg.ln_before:
                grads[pfx + "ln_before.weight"], grads[pfx + "ln_before.bias"] = g.ln1_w, g.ln1_b
            if cfg.ln_post:
                grads[pfx + "ln_post.weight"], grads[pfx + "ln_post.bias"] = g.ln2_w, g.ln2_b
            if bn_mode != BN_OFF:
                grads[pfx + "bn1.weight"], grads[pfx + "bn1.bias"] = r1[e, :, 1], r1[e, :, 0]
                grads[pfx + "bn2.weight"], grads[pfx + "bn2.bias"] = r2[e, :, 1], r2[e, :, 0]
        if fold_res and not res_folded:
            K.add_into(dx, dout)
            res_folded = True
        dres = dout if (ctx.needs_input_grad[3] and not res_folded) else None   # residual passes the gradient through
        return (None, dx, do, dres) + tuple(grads[n] for n in names)

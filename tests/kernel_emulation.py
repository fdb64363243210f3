"""TEST INFRASTRUCTURE — torch emulation of every kernel in the C ABI (device-agnostic).

Each function has exactly the signature of its counterpart in ``avmoe_b200.ops`` and states, in
plain torch, what that CUDA kernel must compute.  Two uses:
  * CPU tests run ``avmoe_b200.functional`` with ``K = kernel_emulation`` to validate the host
    orchestration (pass structure, BatchNorm-from-moments algebra, hand-derived backward) against
    the oracle / golden vectors without a GPU;
  * GPU tests compare each CUDA kernel against its emulation on random inputs (unit parity).
It is never imported by the product package.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn.functional as F

from avmoe_b200.functional import (BN_EPS, BN_EVAL, BN_MOMENTUM, BN_OFF, BN_TRAIN, LN_EPS, PRE_AUX_ADD, PRE_AUX_REPLACE,
                                   PRE_LATENT, PRE_NONE, AdapterConfig, ExpertGrads, ExpertPack)

MHA_HEADS = 4


# ------------------------------------------------------------------------------------------------ GEMM & small helpers
def prep_weight(w: torch.Tensor, act_dtype, transpose: bool) -> torch.Tensor:
    w = w.t() if transpose else w
    return w.to(act_dtype).contiguous()


def transpose(t: torch.Tensor) -> torch.Tensor:
    return t.transpose(1, 2).contiguous()


def gemm(A, B, *, a_kmajor=True, b_kmajor=True, out_dtype=None, alpha=1.0, row_bias=None, col_scale=None, col_bias=None,
         split_k=1, out=None, accumulate=False, impl=0, d_nbatch=0):
    out_dtype = out_dtype or A.dtype
    Af, Bf = A.float(), B.float()
    if not a_kmajor:
        Af = Af.t()
    if b_kmajor:
        if Bf.dim() == 3:
            D = torch.einsum("bmk,bnk->mn", Af, Bf)
        else:
            D = Af @ Bf.t()
    else:
        D = Af @ Bf if Bf.dim() == 2 else torch.einsum("mk,bkn->bmn", Af, Bf)
    if d_nbatch:
        D = D.reshape(D.shape[0], -1, d_nbatch).permute(1, 0, 2)
    D = alpha * D
    if row_bias is not None:
        rb = row_bias[:, None] * (col_scale[None, :] if col_scale is not None else 1.0)
        D = D + rb
    if col_bias is not None:
        D = D + col_bias
    if out is not None:
        if accumulate:
            out += D.to(out.dtype)
        else:
            out.copy_(D.to(out.dtype))
        return out
    return D.to(out_dtype)


def bgemm(A, B, *, a_kmajor=True, b_kmajor=True, out_dtype=None, alpha=1.0, col_bias=None, impl=0, out=None, accumulate=False):
    Af = A.float() if a_kmajor else A.float().transpose(1, 2)
    Bf = B.float() if b_kmajor else B.float().transpose(1, 2)
    D = alpha * (Af @ Bf.transpose(1, 2))
    if col_bias is not None:
        D = D + (col_bias[:, None, :] if col_bias.dim() == 2 else col_bias)
    if out is not None:
        out.copy_((out.float() + D if accumulate else D).to(out.dtype))
        return out
    return D.to(out_dtype or A.dtype)


# ------------------------------------------------------------------------------------------------ GEMM-based latent attention helpers
USE_GEMM_ATTENTION = True


def round8(t):
    return (t + 7) // 8 * 8


def pad_bottleneck(cr):
    return (cr + 31) // 32 * 32


def cast_to(t, dtype):
    return t if t.dtype == dtype else t.to(dtype)


EMULATE_SPLIT = True   # model the hi/lo operand split (with an exact, lopsided split) so the off_lo plumbing is exercised on CPU


def split_operand(t, dtype):
    T, Cc = t.shape[-2], t.shape[-1]
    Tp = round8(T)
    if not EMULATE_SPLIT:
        buf = torch.zeros(t.shape[:-2] + (Tp, Cc), dtype=t.dtype)
        buf[..., :T, :] = t
        return buf, 0, t
    buf = torch.zeros(t.shape[:-2] + (2 * Tp, Cc), dtype=t.dtype)
    buf[..., :T, :] = 0.75 * t          # exact in binary; a consumer that drops or mis-addresses the lo block is off by 25 %
    buf[..., Tp:Tp + T, :] = 0.25 * t
    return buf, Tp, t


def _logits(S, T, off_lo):
    return S[:, :, :T] + S[:, :, off_lo:off_lo + T] if off_lo else S[:, :, :T]


def softmax_cols(S, T, off_lo, dtype):
    BT, Nk, _ = S.shape
    Tp = round8(T)
    A12 = torch.zeros((BT, Nk, 2 * Tp), device=S.device, dtype=dtype)
    L = _logits(S, T, off_lo)
    lse = torch.logsumexp(L, dim=1)
    A12[:, :, :T] = torch.exp(L - lse[:, None, :]).to(dtype)
    # the product's bf16 path stores the rounding residual of A1 at column p_lo = Tp; here the residual is exactly zero
    return A12, lse, (Tp if EMULATE_SPLIT else 0)


def softmax_cols_bwd(A12, dA, Dt, T, off_lo):
    Tp = A12.shape[2] // 2
    A12[:, :, Tp:Tp + T] = (A12[:, :, :T].float() * (_logits(dA, T, off_lo) - Dt[:, None, :])).to(A12.dtype)


def softmax_rows(S, T, off_lo, dtype):
    A = torch.zeros(S.shape[:2] + (round8(T),), device=S.device, dtype=dtype)
    A[:, :, :T] = torch.softmax(_logits(S, T, off_lo), dim=-1).to(dtype)
    return A


def softmax_rows_bwd(A, dA, T, off_lo):
    a = A[:, :, :T].float()
    d = _logits(dA, T, off_lo)
    dS = torch.zeros_like(A)
    dS[:, :, :T] = (a * (d - (a * d).sum(-1, keepdim=True))).to(A.dtype)
    return dS


def tile_add(R, tile, off_lo=0):
    T = tile.shape[0]
    return R[:, :T] + (R[:, off_lo:off_lo + T] if off_lo else 0) + tile[None]


def latent_dots(dR, tokp, tok):
    return (dR * (tokp - tok[None])).sum(-1)


def batch_sum_add(t, out):
    out += t.sum(0).reshape(out.shape)


def add_rowvec(d, vec):
    d += vec[:, None, :].to(d.dtype)


def add_into(dst, a, b=None, c=None):
    v = dst.float() + a.float()
    if b is not None:
        v = v + b.float()
    if c is not None:
        v = v + c.float()
    dst.copy_(v.to(dst.dtype))


def rowsum(w):
    return w.float().sum(dim=1)


def reduce_bias(dt, col_weight=None, row_weight=None):
    d = dt.float()
    rows_w = (d * (col_weight if col_weight is not None else 1.0)).sum(dim=(0, 2))
    cols = d.sum(dim=(0, 1))
    cols_w = (d * row_weight[None, :, None]).sum(dim=(0, 1)) if row_weight is not None else None
    return rows_w, cols, cols_w


def split_k_hint(k_rows, m, n):
    return 1


def pool(x):
    return x.float().sum(dim=1, keepdim=True)          # [BT, chunks=1, C] partial sums


def broadcast_rows(dpool, n, dtype):
    return dpool[:, None, :].expand(-1, n, -1).to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------ router
def router_fwd(pool_x, pool_v, inv_n, rw, noise):
    w0, b0, w2, b2, w4, b4 = rw
    rin = torch.cat((pool_x.sum(1), pool_v.sum(1)), dim=-1) * inv_n
    h1 = torch.relu(rin @ w0.t() + b0)
    h2 = torch.relu(h1 @ w2.t() + b2)
    logits = h2 @ w4.t() + b4
    if noise is not None:
        logits = logits + noise
    probs = torch.softmax(logits, dim=-1)
    idx = torch.argmax(probs, dim=-1)
    return probs, idx, h1, h2, rin


def router_bwd(dprobs, probs, h1, h2, rin, inv_n, rw):
    w0, b0, w2, b2, w4, b4 = rw
    C = rin.shape[1] // 2
    dlogits = probs * (dprobs - (dprobs * probs).sum(-1, keepdim=True))
    dw4, db4 = dlogits.t() @ h2, dlogits.sum(0)
    dh2 = (dlogits @ w4) * (h2 > 0)
    dw2, db2 = dh2.t() @ h1, dh2.sum(0)
    dh1 = (dh2 @ w2) * (h1 > 0)
    dw0, db0 = dh1.t() @ rin, dh1.sum(0)
    drin = (dh1 @ w0) * inv_n
    return drin[:, :C].contiguous(), drin[:, C:].contiguous(), [dw0, db0, dw2, db2, dw4, db4]


def lb_loss_fwd(probs):
    return -(probs.mean(dim=0).log()).sum()


def lb_loss_bwd(probs, dlb, dprobs):
    bt = probs.shape[0]
    dprobs += -dlb / (bt * probs.mean(dim=0, keepdim=True))


# ------------------------------------------------------------------------------------------------ latent attention
def latent_fwd(kv, tokens: List[torch.Tensor], slots: List[int], tokp, lse):
    """tok' = tok + softmax_N(tok·kvᵀ)·kv ; lse = logsumexp_N(tok·kvᵀ).   (net_trans_v3.py:379-383)"""
    kvf = kv.float()
    for tok, s in zip(tokens, slots):
        logits = torch.einsum("tc,bnc->btn", tok, kvf)
        lse[:, s] = torch.logsumexp(logits, dim=-1)
        tokp[:, s] = tok[None] + torch.softmax(logits, dim=-1) @ kvf


def latent_bwd(kv, tokens, slots, tokp, lse, dtokp, dpool, dkv, accumulate, dtokens):
    kvf = kv.float()
    total = dkv.float() if accumulate else torch.zeros_like(kvf)
    for tok, s, dtok in zip(tokens, slots, dtokens):
        dR = dtokp[:, s]                                            # [BT,T,C]
        R = tokp[:, s] - tok[None]
        a1 = torch.exp(torch.einsum("tc,bnc->btn", tok, kvf) - lse[:, s][..., None])
        dA = torch.einsum("btc,bnc->btn", dR, kvf)
        dS = a1 * (dA - (dR * R).sum(-1, keepdim=True))
        total = total + torch.einsum("btn,btc->bnc", a1, dR) + torch.einsum("btn,tc->bnc", dS, tok)
        dtok += dR.sum(0) + torch.einsum("btn,bnc->tc", dS, kvf)
    if dpool is not None:
        total = total + dpool[:, None, :]
    dkv.copy_(total.to(dkv.dtype))


# ------------------------------------------------------------------------------------------------ AVVP NxN self-attention
def selfattn_nxn_fwd(x):
    """aux_j = sum_i softmax_j(x_i·x_·)[j] x_i   (mgn.py:132-137).  Saves the row log-sum-exps."""
    xf = x.float()
    s = xf @ xf.transpose(1, 2)
    lse_rows = torch.logsumexp(s, dim=-1)                            # [BT,N]
    att = torch.exp(s - lse_rows[..., None])
    return (att.transpose(1, 2) @ xf).to(x.dtype), lse_rows


def selfattn_nxn_bwd(x, lse_rows, daux, dx):
    xf, g = x.float(), daux.float()
    s = xf @ xf.transpose(1, 2)
    att = torch.exp(s - lse_rows[..., None])                          # att[i,j]
    # aux_j = sum_i att[i,j] x_i
    datt = xf @ g.transpose(1, 2)                                     # datt[i,j] = x_i · g_j
    ds = att * (datt - (att * datt).sum(-1, keepdim=True))
    d = att @ g + ds @ xf + ds.transpose(1, 2) @ xf
    dx += d.to(dx.dtype)


# ------------------------------------------------------------------------------------------------ AVS v1 MHA over the BT axis
def mha_fwd(x, mp, keep=None, keep_scale=1.0):
    w_in, b_in, w_out, b_out = mp
    L, B, C = x.shape
    hd = C // MHA_HEADS
    xf = x.float()
    qkv = xf @ w_in.t() + b_in
    q, k, v = qkv.split(C, dim=-1)

    def heads(t):
        return t.reshape(L, B * MHA_HEADS, hd).transpose(0, 1)
    qh, kh, vh = heads(q), heads(k), heads(v)
    att = torch.softmax((qh / math.sqrt(hd)) @ kh.transpose(1, 2), dim=-1)
    mask = None if keep is None else keep.reshape(B * MHA_HEADS, L, L).float() * keep_scale
    ctxv = ((att if mask is None else att * mask) @ vh).transpose(0, 1).reshape(L, B, C)
    out = ctxv @ w_out.t() + b_out
    return out.to(x.dtype), (qkv, att, ctxv, mask)


def mha_bwd(x, mp, saved, daux, dx):
    w_in, b_in, w_out, b_out = mp
    qkv, att, ctxv, mask = saved
    L, B, C = x.shape
    hd = C // MHA_HEADS
    g = daux.float()
    dw_out = g.reshape(-1, C).t() @ ctxv.reshape(-1, C)
    db_out = g.sum(dim=(0, 1))
    dctx = g @ w_out

    def heads(t):
        return t.reshape(L, B * MHA_HEADS, hd).transpose(0, 1)
    q, k, v = qkv.split(C, dim=-1)
    qh, kh, vh, dch = heads(q), heads(k), heads(v), heads(dctx)
    datt = dch @ vh.transpose(1, 2)
    dvh = (att if mask is None else att * mask).transpose(1, 2) @ dch
    if mask is not None:
        datt = datt * mask
    ds = att * (datt - (att * datt).sum(-1, keepdim=True))
    dqh = (ds @ kh) / math.sqrt(hd)
    dkh = ds.transpose(1, 2) @ (qh / math.sqrt(hd))

    def unheads(t):
        return t.transpose(0, 1).reshape(L, B, C)
    dqkv = torch.cat((unheads(dqh), unheads(dkh), unheads(dvh)), dim=-1)
    dw_in = dqkv.reshape(-1, 3 * C).t() @ x.float().reshape(-1, C)
    db_in = dqkv.sum(dim=(0, 1))
    dx += (dqkv @ w_in).to(dx.dtype)
    return [dw_in, db_in, dw_out, db_out]


# ------------------------------------------------------------------------------------------------ expert kernels
def _ln_stats(t):
    mean = t.mean(-1, keepdim=True)
    rstd = torch.rsqrt(t.var(-1, unbiased=False, keepdim=True) + LN_EPS)
    return mean, rstd


def _grouped(x, w, G):
    """x [...,Cin], w [Cout, Cin/G] -> [...,Cout]"""
    Cout, Cin_g = w.shape
    lead = x.shape[:-1]
    xg = x.reshape(*lead, G, Cin_g)
    wg = w.reshape(G, Cout // G, Cin_g)
    return torch.einsum("...gi,goi->...go", xg, wg).reshape(*lead, Cout)


def _grouped_t(dy, w, G):
    """dy [...,Cout], w [Cout, Cin/G] -> dx [...,Cin]"""
    Cout, Cin_g = w.shape
    lead = dy.shape[:-1]
    dyg = dy.reshape(*lead, G, Cout // G)
    wg = w.reshape(G, Cout // G, Cin_g)
    return torch.einsum("...go,goi->...gi", dyg, wg).reshape(*lead, G * Cin_g)


def _grouped_wgrad(dy, x, G):
    """sum over leading dims of dy^T x per group -> [Cout, Cin/G]"""
    Cout, Cin = dy.shape[-1], x.shape[-1]
    dyg = dy.reshape(-1, G, Cout // G)
    xg = x.reshape(-1, G, Cin // G)
    return torch.einsum("ngo,ngi->goi", dyg, xg).reshape(Cout, Cin // G)


def _pre_transform(x, tokp, pk: ExpertPack):
    """returns (xa, cache) with xa = the tensor ln_before sees."""
    xf = x.float()
    if pk.pre_kind == PRE_NONE:
        return xf, None
    if pk.pre_kind == PRE_LATENT:
        tp = tokp[:, pk.latent_slot]                                  # [BT,T,C]
        a2 = torch.softmax(xf @ tp.transpose(1, 2), dim=-1)           # [BT,N,T]
        xr = a2 @ tp
        return xf + pk.gate_pre * xr, (a2, xr, tp)
    if pk.pre_kind == PRE_AUX_ADD:
        return xf + pk.gate_pre * pk.aux.float(), None
    return pk.aux.float(), None


def expert_front_fwd(cfg: AdapterConfig, x, tokp, packs: List[ExpertPack], z, stats1):
    for e, pk in enumerate(packs):
        xa, _ = _pre_transform(x, tokp, pk)
        xn = F.layer_norm(xa, (cfg.C,), pk.ln1_w, pk.ln1_b, LN_EPS) if pk.ln1_w is not None else xa
        ze = _grouped(xn, pk.wd, cfg.G)
        z[:, :, e, :] = ze
        if stats1 is not None:
            stats1[e, :, 0] += ze.double().sum(dim=(0, 1))
            stats1[e, :, 1] += (ze.double() ** 2).sum(dim=(0, 1))


def _running_update(rm, rv, mean, var_biased, n):
    rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.to(rm.dtype))
    rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var_biased * (n / max(n - 1, 1))).to(rv.dtype))


def bn1_finalize(cfg, stats1, n, bn_w, bn_b, train):
    E, Cr = cfg.E, cfg.Cr
    dev = bn_w[0][0].device
    aff = torch.empty((E, Cr, 2), device=dev, dtype=torch.float32)
    save = torch.empty((E, Cr, 2), device=dev, dtype=torch.float32)
    for e in range(E):
        w1, b1 = bn_w[e][0], bn_w[e][1]
        rm1, rv1 = bn_b[e][0], bn_b[e][1]
        if train:
            mean = stats1[e, :, 0] / n
            var = (stats1[e, :, 1] / n - mean * mean).clamp_min(0)
            _running_update(rm1, rv1, mean, var, n)
        else:
            mean, var = rm1.double(), rv1.double()
        invstd = torch.rsqrt(var + BN_EPS)
        aff[e, :, 0] = (w1.double() * invstd).float()
        aff[e, :, 1] = (b1.double() - mean * w1.double() * invstd).float()
        save[e, :, 0], save[e, :, 1] = mean.float(), invstd.float()
    return aff, save


def _act(zt, relu):
    return torch.relu(zt) if relu else zt


def bn_moments(cfg, z, aff1, packs):
    E, Cr, G = cfg.E, cfg.Cr, cfg.G
    crg = Cr // G
    m1 = torch.zeros((E, Cr), device=z.device, dtype=torch.float64)
    m2 = torch.zeros((E, G, crg, crg), device=z.device, dtype=torch.float64)
    for e, pk in enumerate(packs):
        a = _act(z[:, :, e, :] * aff1[e, :, 0] + aff1[e, :, 1], pk.relu).double().reshape(-1, G, crg)
        m1[e] = a.sum(0).reshape(Cr)
        m2[e] = torch.einsum("ngj,ngk->gjk", a, a)
    return m1, m2


def bn2_finalize(cfg, moments, n, packs, bn_w, bn_b, train):
    E, C, Cr, G = cfg.E, cfg.C, cfg.Cr, cfg.G
    crg, cg = Cr // G, C // G
    dev = bn_w[0][0].device
    aff = torch.empty((E, C, 2), device=dev, dtype=torch.float32)
    save = torch.empty((E, C, 2), device=dev, dtype=torch.float32)
    for e, pk in enumerate(packs):
        w2, b2 = bn_w[e][2], bn_w[e][3]
        rm2, rv2 = bn_b[e][2], bn_b[e][3]
        if train:
            m1, m2 = moments
            wu = pk.wu.double().reshape(G, cg, crg)                   # [g, c_local, j]
            ea = (m1[e] / n).reshape(G, crg)
            eaa = m2[e] / n                                           # [g, j, k]
            mean = torch.einsum("gcj,gj->gc", wu, ea).reshape(C)
            ey2 = torch.einsum("gcj,gjk,gck->gc", wu, eaa, wu).reshape(C)
            var = (ey2 - mean * mean).clamp_min(0)
            _running_update(rm2, rv2, mean, var, n)
        else:
            mean, var = rm2.double(), rv2.double()
        invstd = torch.rsqrt(var + BN_EPS)
        aff[e, :, 0] = (w2.double() * invstd).float()
        aff[e, :, 1] = (b2.double() - mean * w2.double() * invstd).float()
        save[e, :, 0], save[e, :, 1] = mean.float(), invstd.float()
    return aff, save


def _back_recompute(cfg, z, bn, pk, e):
    """from z recompute a, y, yt (post-bn2), ln2 stats"""
    ze = z[:, :, e, :]
    if bn is not None:
        aff1, save1, aff2, save2 = bn
        zt = ze * aff1[e, :, 0] + aff1[e, :, 1]
    else:
        zt = ze
    a = _act(zt, pk.relu)
    y = _grouped(a, pk.wu, cfg.G)
    yt = y * bn[2][e, :, 0] + bn[2][e, :, 1] if bn is not None else y
    return zt, a, y, yt


def expert_back_fwd(cfg, z, bn, packs, probs, out):
    acc = torch.zeros(z.shape[0], z.shape[1], cfg.C, device=z.device, dtype=torch.float32)
    for e, pk in enumerate(packs):
        _, _, _, yt = _back_recompute(cfg, z, bn, pk, e)
        q = F.layer_norm(yt, (cfg.C,), pk.ln2_w, pk.ln2_b, LN_EPS) if pk.ln2_w is not None else yt
        if pk.gate is not None:
            q = q * pk.gate
        acc = acc + probs[:, e, None, None] * q
    out.copy_(acc.to(out.dtype))


def _back_dyt(cfg, dout, z, bn, pk, e, probs):
    """shared by the stats pass and the main backward: returns intermediates up to dyt."""
    zt, a, y, yt = _back_recompute(cfg, z, bn, pk, e)
    g = dout.float()
    if pk.ln2_w is not None:
        mean, rstd = _ln_stats(yt)
        qhat = (yt - mean) * rstd
        q = qhat * pk.ln2_w + pk.ln2_b
    else:
        qhat, q = None, yt
    gate = pk.gate if pk.gate is not None else 1.0
    dq = g * (probs[:, e, None, None] * gate)
    if pk.ln2_w is not None:
        dqh = dq * pk.ln2_w
        dyt = rstd * (dqh - dqh.mean(-1, keepdim=True) - qhat * (dqh * qhat).mean(-1, keepdim=True))
    else:
        dyt = dq
    return zt, a, y, yt, q, qhat, dq, dyt, g


def expert_back_bwd_stats(cfg, dout, z, bn, packs, probs, red2):
    for e, pk in enumerate(packs):
        zt, a, y, yt, q, qhat, dq, dyt, g = _back_dyt(cfg, dout, z, bn, pk, e, probs)
        yhat = (y - bn[3][e, :, 0]) * bn[3][e, :, 1]
        red2[e, :, 0] += dyt.double().sum(dim=(0, 1))
        red2[e, :, 1] += (dyt * yhat).double().sum(dim=(0, 1))


def expert_back_bwd(cfg, dout, z, bn, bn_mode, red2, red2b, packs, probs, dzt, dprobs, red1, egr: List[ExpertGrads]):
    n = z.shape[0] * z.shape[1]
    for e, (pk, gr) in enumerate(zip(packs, egr)):
        zt, a, y, yt, q, qhat, dq, dyt, g = _back_dyt(cfg, dout, z, bn, pk, e, probs)
        gq = (g * q).sum(dim=(1, 2))                                  # [BT]  sum dout*q (q before gate)
        gate = pk.gate if pk.gate is not None else 1.0
        dprobs[:, e] += gq * gate
        if gr.gate is not None:
            gr.gate += (gq * probs[:, e]).sum()
        if pk.ln2_w is not None:
            gr.ln2_w += (dq * qhat).sum(dim=(0, 1))
            gr.ln2_b += dq.sum(dim=(0, 1))
        if bn_mode != BN_OFF:
            aff2, save2 = bn[2], bn[3]
            yhat = (y - save2[e, :, 0]) * save2[e, :, 1]
            if bn_mode == BN_TRAIN:
                m_d = (red2[e, :, 0] / n).float()
                m_dy = (red2[e, :, 1] / n).float()
                dy = aff2[e, :, 0] * (dyt - m_d - yhat * m_dy)
            else:
                dy = aff2[e, :, 0] * dyt
                red2b[e, :, 0] += dyt.double().sum(dim=(0, 1))
                red2b[e, :, 1] += (dyt * yhat).double().sum(dim=(0, 1))
        else:
            dy = dyt
        gr.wu += _grouped_wgrad(dy, a, cfg.G)
        da = _grouped_t(dy, pk.wu, cfg.G)
        d = da * (zt > 0) if pk.relu else da
        dzt[:, :, e, :] = d
        if bn_mode != BN_OFF:
            save1 = bn[1]
            zhat = (z[:, :, e, :] - save1[e, :, 0]) * save1[e, :, 1]
            red1[e, :, 0] += d.double().sum(dim=(0, 1))
            red1[e, :, 1] += (d * zhat).double().sum(dim=(0, 1))


def expert_front_bwd(cfg, x, tokp, dzt, z, bn, bn_mode, red1, packs, dpool_x, dx, dtokp, egr: List[ExpertGrads]):
    n = x.shape[0] * x.shape[1]
    xf = x.float()
    dxa_total = dpool_x[:, None, :].expand_as(xf).clone()
    for e, (pk, gr) in enumerate(zip(packs, egr)):
        d = dzt[:, :, e, :]
        if bn_mode != BN_OFF:
            aff1, save1 = bn[0], bn[1]
            if bn_mode == BN_TRAIN:
                zhat = (z[:, :, e, :] - save1[e, :, 0]) * save1[e, :, 1]
                dz = aff1[e, :, 0] * (d - (red1[e, :, 0] / n).float() - zhat * (red1[e, :, 1] / n).float())
            else:
                dz = aff1[e, :, 0] * d
        else:
            dz = d
        xa, cache = _pre_transform(x, tokp, pk)
        if pk.ln1_w is not None:
            mean, rstd = _ln_stats(xa)
            xhat = (xa - mean) * rstd
            xn = xhat * pk.ln1_w + pk.ln1_b
        else:
            xn = xa
        gr.wd += _grouped_wgrad(dz, xn, cfg.G)
        dxn = _grouped_t(dz, pk.wd, cfg.G)
        if pk.ln1_w is not None:
            gr.ln1_w += (dxn * xhat).sum(dim=(0, 1))
            gr.ln1_b += dxn.sum(dim=(0, 1))
            dh = dxn * pk.ln1_w
            dxa = rstd * (dh - dh.mean(-1, keepdim=True) - xhat * (dh * xhat).mean(-1, keepdim=True))
        else:
            dxa = dxn
        if pk.pre_kind == PRE_NONE:
            dxa_total += dxa
        elif pk.pre_kind == PRE_LATENT:
            a2, xr, tp = cache
            dxa_total += dxa
            gr.gate_pre += (dxa * xr).sum()
            dxr = dxa * pk.gate_pre
            da2 = dxr @ tp.transpose(1, 2)                            # [BT,N,T]
            ds2 = a2 * (da2 - (a2 * da2).sum(-1, keepdim=True))
            dxa_total += ds2 @ tp
            dtokp[:, pk.latent_slot] += a2.transpose(1, 2) @ dxr + ds2.transpose(1, 2) @ xf
        elif pk.pre_kind == PRE_AUX_ADD:
            dxa_total += dxa
            gr.gate_pre += (dxa * pk.aux.float()).sum()
            gr.aux.copy_((dxa * pk.gate_pre).to(gr.aux.dtype))
        else:
            gr.aux.copy_(dxa.to(gr.aux.dtype))
    dx.copy_(dxa_total.to(dx.dtype))


# ------------------------------------------------------------------------------------------------ GEMM-based expert path (csrc/expert_rows.cu)
USE_GEMM_EXPERTS = True


def _dense_wd(wd, cfg):
    """grouped [Cr, C/G] -> block-diagonal dense [Cr, C]"""
    Cr, C, G = cfg.Cr, cfg.C, cfg.G
    out = torch.zeros((Cr, C), dtype=torch.float32)
    for g in range(G):
        out[g * (Cr // G):(g + 1) * (Cr // G), g * (C // G):(g + 1) * (C // G)] = wd[g * (Cr // G):(g + 1) * (Cr // G)]
    return out


def _dense_wu(wu, cfg):
    """grouped [C, Cr/G] -> block-diagonal dense [C, Crp] (zero padded)"""
    Cr, C, G = cfg.Cr, cfg.C, cfg.G
    out = torch.zeros((C, pad_bottleneck(Cr)), dtype=torch.float32)
    for g in range(G):
        out[g * (C // G):(g + 1) * (C // G), g * (Cr // G):(g + 1) * (Cr // G)] = wu[g * (C // G):(g + 1) * (C // G)]
    return out


def xg_prep_weights(cfg, packs, dtype):
    Crp = pad_bottleneck(cfg.Cr)
    wdd = torch.stack([_dense_wd(pk.wd, cfg) for pk in packs]).to(dtype)
    wdt = torch.zeros((cfg.E, cfg.C, Crp), dtype=dtype)
    wdt[:, :, :cfg.Cr] = wdd.transpose(1, 2)
    return wdd, wdt


def xg_prep_up(cfg, packs, save2, bn_mode, dtype):
    wud = torch.stack([_dense_wu(pk.wu, cfg) for pk in packs])
    ybi = torch.zeros((cfg.E, cfg.C), dtype=torch.float32)
    if bn_mode != BN_OFF:
        wud = wud * save2[:, :, 1, None]
        ybi = -save2[:, :, 0] * save2[:, :, 1]
    return wud.to(dtype), ybi


def _xg_xa(x, pk):
    xf = x.float()
    if pk.pre_kind == PRE_NONE:
        return xf
    assert pk.pre_kind != PRE_LATENT, "the GEMM expert path takes latent attention as an aux-add"
    if pk.pre_kind == PRE_AUX_ADD:
        return xf + pk.gate_pre * pk.aux.float()
    return pk.aux.float()


def xg_ln_pre_fwd(cfg, x, packs):
    BT, N, C = x.shape
    xl = torch.empty((cfg.E, BT * N, C), dtype=x.dtype)
    st = torch.empty((cfg.E, BT * N, 2), dtype=torch.float32) if packs[0].ln1_w is not None else None
    for e, pk in enumerate(packs):
        xa = _xg_xa(x, pk).reshape(BT * N, C)
        if pk.ln1_w is not None:
            mean, rstd = _ln_stats(xa)
            st[e, :, 0], st[e, :, 1] = mean[:, 0], rstd[:, 0]
            xa = (xa - mean) * rstd * pk.ln1_w + pk.ln1_b
        xl[e] = xa.to(x.dtype)
    return xl, st


def xg_z_stats(cfg, z):
    zz = z.double().reshape(-1, cfg.E, cfg.Cr)
    return torch.stack([zz.sum(0), (zz * zz).sum(0)], dim=-1)


def xg_bn_act(cfg, z, aff1, packs, dtype, bn_mode):
    BT, N = z.shape[0], z.shape[1]
    act = torch.zeros((cfg.E, BT * N, pad_bottleneck(cfg.Cr)), dtype=dtype)
    for e, pk in enumerate(packs):
        zt = z[:, :, e, :].reshape(BT * N, cfg.Cr)
        if bn_mode != BN_OFF:
            zt = zt * aff1[e, :, 0] + aff1[e, :, 1]
        act[e, :, :cfg.Cr] = _act(zt, pk.relu).to(dtype)
    return act


def _xg_post_t(cfg, yd, bn2_wb, pk, e):
    """bn2 affine applied to the normalised up projection: the tensor ln_post sees"""
    yh = yd[e, :, :cfg.C].float()
    return yh, (yh * bn2_wb[e][0] + bn2_wb[e][1] if bn2_wb is not None else yh)


def xg_post_fwd(cfg, yd, bn2_wb, packs, probs, out, res=None):
    BT, N, C = out.shape
    st = torch.empty((cfg.E, BT * N, 2), dtype=torch.float32) if packs[0].ln2_w is not None else None
    acc = torch.zeros((BT, N, C), dtype=torch.float32)
    for e, pk in enumerate(packs):
        _, t = _xg_post_t(cfg, yd, bn2_wb, pk, e)
        if pk.ln2_w is not None:
            mean, rstd = _ln_stats(t)
            st[e, :, 0], st[e, :, 1] = mean[:, 0], rstd[:, 0]
            t = (t - mean) * rstd * pk.ln2_w + pk.ln2_b
        if pk.gate is not None:
            t = t * pk.gate
        acc = acc + probs[:, e, None, None] * t.reshape(BT, N, C)
    if res is not None:
        acc = acc + res.float()
    out.copy_(acc.to(out.dtype))
    return st


def xg_post_bwd(cfg, dout, yd, ln2_stats, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr):
    BT, N, C = dout.shape
    g = dout.float().reshape(BT * N, C)
    for e, (pk, gr) in enumerate(zip(packs, egr)):
        yh, t = _xg_post_t(cfg, yd, bn2_wb, pk, e)
        if pk.ln2_w is not None:
            mean, rstd = ln2_stats[e, :, 0:1], ln2_stats[e, :, 1:2]
            qhat = (t - mean) * rstd
            q = qhat * pk.ln2_w + pk.ln2_b
        else:
            qhat, q = None, t
        gate = pk.gate if pk.gate is not None else 1.0
        gq = (g * q).reshape(BT, N * C).sum(1)
        dprobs[:, e] += gq * gate
        if gr.gate is not None:
            gr.gate += (gq * probs[:, e]).sum()
        dq = g * (probs[:, e].repeat_interleave(N)[:, None] * gate)
        if pk.ln2_w is not None:
            gr.ln2_w += (dq * qhat).sum(0)
            gr.ln2_b += dq.sum(0)
            dqh = dq * pk.ln2_w
            dy2 = rstd * (dqh - dqh.mean(-1, keepdim=True) - qhat * (dqh * qhat).mean(-1, keepdim=True))
        else:
            dy2 = dq
        if bn_mode != BN_OFF:
            red2[e, :, 0] += dy2.double().sum(0)
            red2[e, :, 1] += (dy2 * yh).double().sum(0)
        yd[e, :, C:] = dy2.to(yd.dtype)


def xg_bn1_bwd(cfg, da, z, bn, bn_mode, packs, red1, dtype):
    rows = da.shape[1]
    Cr = cfg.Cr
    dz = torch.zeros(da.shape, dtype=dtype)
    for e, pk in enumerate(packs):
        ze = z[:, :, e, :].reshape(rows, Cr)
        d = da[e, :, :Cr]
        zt = ze * bn[0][e, :, 0] + bn[0][e, :, 1] if bn_mode != BN_OFF else ze
        if pk.relu:
            d = d * (zt > 0)
        if bn_mode != BN_OFF:
            aff1, save1 = bn[0], bn[1]
            zhat = (ze - save1[e, :, 0]) * save1[e, :, 1]
            red1[e, :, 0] += d.double().sum(0)
            red1[e, :, 1] += (d * zhat).double().sum(0)
            if bn_mode == BN_TRAIN:
                d = aff1[e, :, 0] * (d - (red1[e, :, 0] / rows).float() - zhat * (red1[e, :, 1] / rows).float())
            else:
                d = aff1[e, :, 0] * d
        dz[e, :, :Cr] = d.to(dtype)
    return dz


def xg_ln_pre_bwd(cfg, x, dxl, ln1_stats, packs, dpool_x, dx, egr):
    BT, N, C = x.shape
    total = dpool_x[:, None, :].expand(BT, N, C).reshape(BT * N, C).clone() if dpool_x is not None else torch.zeros((BT * N, C))
    for e, (pk, gr) in enumerate(zip(packs, egr)):
        g = dxl[e].float()
        if pk.ln1_w is not None:
            xa = _xg_xa(x, pk).reshape(BT * N, C)
            mean, rstd = ln1_stats[e, :, 0:1], ln1_stats[e, :, 1:2]
            xhat = (xa - mean) * rstd
            gr.ln1_w += (g * xhat).sum(0)
            gr.ln1_b += g.sum(0)
            dh = g * pk.ln1_w
            g = rstd * (dh - dh.mean(-1, keepdim=True) - xhat * (dh * xhat).mean(-1, keepdim=True))
        if pk.pre_kind == PRE_AUX_ADD:
            gr.gate_pre += (g * pk.aux.float().reshape(BT * N, C)).sum()
            gr.aux.copy_((pk.gate_pre * g).reshape(BT, N, C).to(gr.aux.dtype))
            total = total + g
        elif pk.pre_kind == PRE_AUX_REPLACE:
            gr.aux.copy_(g.reshape(BT, N, C).to(gr.aux.dtype))
        else:
            total = total + g
    dx.copy_(total.reshape(BT, N, C).to(dx.dtype))


# ------------------------------------------------------------------------------------------------ row-fused passes (csrc/rowfused.cu)
USE_ROWFUSED = True


def rf_supported(cfg, dtype) -> bool:
    crk = (cfg.Cr + 15) // 16
    return cfg.C % 16 == 0 and cfg.Cr % 2 == 0 and crk in (1, 2, 3, 6) and 2 <= cfg.E <= 4 and cfg.T <= 96


def rf_prep_up(cfg, packs, save2, bn_mode):
    """emulated "packed operand": the dense fp32 [E,C,Crp] matrix itself"""
    wud, ybi = xg_prep_up(cfg, packs, save2, bn_mode, torch.float32)
    return wud, ybi


def _rf_yhat(cfg, z, aff1, wud, ybi, packs, bn_mode):
    act = xg_bn_act(cfg, z, aff1, packs, torch.float32, bn_mode)                  # [E,rows,Crp]
    return act, torch.einsum("erk,eck->erc", act, wud) + ybi[:, None, :]


def rf_back_fwd(cfg, z, aff1, wu_pack, ybi, bn2_wb, bn_mode, packs, probs, out, res=None):
    _, yh = _rf_yhat(cfg, z, aff1, wu_pack, ybi, packs, bn_mode)
    yd = torch.cat([yh, torch.zeros_like(yh)], dim=-1)
    xg_post_fwd(cfg, yd, bn2_wb, packs, probs, out, res)


def rf_back_bwd(cfg, dout, z, aff1, wu_pack, ybi, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr):
    act, yh = _rf_yhat(cfg, z, aff1, wu_pack, ybi, packs, bn_mode)
    yd = torch.cat([yh, torch.zeros_like(yh)], dim=-1).to(dout.dtype)
    st = None
    if packs[0].ln2_w is not None:
        st = torch.empty((cfg.E, yh.shape[1], 2), dtype=torch.float32)
        for e, pk in enumerate(packs):
            _, t = _xg_post_t(cfg, yd, bn2_wb, pk, e)
            mean, rstd = _ln_stats(t)
            st[e, :, 0], st[e, :, 1] = mean[:, 0], rstd[:, 0]
    xg_post_bwd(cfg, dout, yd, st, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr)
    return yd, act.to(dout.dtype)


def rf_prep_front(cfg, packs):
    return {"has_ln": packs[0].ln1_w is not None}


def rf_pack_tokens(tokp):
    return tokp, tokp        # emulated "packed" operands: the fp32 tokens themselves


def _rf_aux_packs(cfg, x, packs, toks):
    """latent experts restated as aux-add experts (aux = softmax_T(x tok'^T) tok'), plus the probabilities"""
    out, probs = [], {}
    for e, pk in enumerate(packs):
        if pk.pre_kind == PRE_LATENT:
            tp = toks[e][0].float()
            a2 = torch.softmax(x.float() @ tp.transpose(1, 2), dim=-1)
            q = ExpertPack(pre_kind=PRE_AUX_ADD, relu=pk.relu, gate_pre=pk.gate_pre, aux=a2 @ tp, ln1_w=pk.ln1_w, ln1_b=pk.ln1_b,
                           wd=pk.wd, wu=pk.wu, ln2_w=pk.ln2_w, ln2_b=pk.ln2_b, gate=pk.gate)
            probs[e] = a2
            out.append(q)
        else:
            out.append(pk)
    return out, probs


def rf_front_fwd(cfg, x, packs, toks, prep, z):
    BT, N, C = x.shape
    ap, _ = _rf_aux_packs(cfg, x, packs, toks)
    xl, _ = xg_ln_pre_fwd(cfg, x.float(), ap)
    for e, pk in enumerate(packs):
        z[:, :, e, :] = (xl[e].float() @ _dense_wd(pk.wd, cfg).t()).reshape(BT, N, cfg.Cr)
    return None      # the CUDA kernel returns the saved ln_before statistics; the emulation recomputes them


def rf_front_bwd(cfg, x, packs, egr, toks, prep, ln1_stats, dz, dpool_x, dx, dres=None):
    BT, N, C = x.shape
    ap, a2s = _rf_aux_packs(cfg, x, packs, toks)
    xl, st = xg_ln_pre_fwd(cfg, x.float(), ap)
    dxl = torch.stack([dz[e, :, :cfg.Cr].float() @ _dense_wd(pk.wd, cfg) for e, pk in enumerate(packs)])
    for e, pk in enumerate(packs):      # latent experts: the aux gradient buffer is the dXR the caller allocated
        if egr[e].aux is None and pk.pre_kind == PRE_LATENT:
            egr[e].aux = torch.empty((BT, N, C), dtype=x.dtype)
    dxf = torch.empty((BT, N, C), dtype=torch.float32)
    xg_ln_pre_bwd(cfg, x.float(), dxl, st, ap, dpool_x, dxf, egr)
    lat = {}
    tp8 = round8(cfg.T)
    for e, pk in enumerate(packs):
        if pk.pre_kind == PRE_LATENT:
            tp = toks[e][0].float()
            a2 = a2s[e]
            da2 = egr[e].aux.float() @ tp.transpose(1, 2)
            ds2 = a2 * (da2 - (a2 * da2).sum(-1, keepdim=True))
            dxf = dxf + ds2 @ tp
            A2 = torch.zeros((BT, N, tp8), dtype=x.dtype)
            D2 = torch.zeros((BT, N, tp8), dtype=x.dtype)
            A2[:, :, :cfg.T], D2[:, :, :cfg.T] = a2.to(x.dtype), ds2.to(x.dtype)
            lat[e] = (A2, D2)
    if dres is not None:
        dxf = dxf + dres.float()
    dx.copy_(dxf.to(dx.dtype))
    return xl.to(x.dtype), lat


def _xg_bn2_coeffs(cfg, bn2_wb, bn, bn_mode, red2, rows, e):
    C = cfg.C
    alpha, beta, gamma = torch.ones(C), torch.zeros(C), torch.zeros(C)
    if bn_mode != BN_OFF:
        alpha = bn[3][e, :, 1] * bn2_wb[e][0]
        if bn_mode == BN_TRAIN:
            gamma = -alpha * (red2[e, :, 0] / rows).float()
            beta = -alpha * (red2[e, :, 1] / rows).float()
    return alpha, beta, gamma


def xg_prep_bwd(cfg, packs, bn2_wb, bn, bn_mode, red2, BT, dtype):
    rows = BT * cfg.N
    Crp, C = pad_bottleneck(cfg.Cr), cfg.C
    w2 = torch.zeros((cfg.E, Crp, 2 * C), dtype=torch.float32)
    bias = torch.zeros((cfg.E, Crp), dtype=torch.float32)
    for e, pk in enumerate(packs):
        alpha, beta, gamma = _xg_bn2_coeffs(cfg, bn2_wb, bn, bn_mode, red2, rows, e)
        wu = _dense_wu(pk.wu, cfg)                       # [C, Crp]
        w2[e, :, :C] = (beta[:, None] * wu).t()
        w2[e, :, C:] = (alpha[:, None] * wu).t()
        bias[e] = (gamma[:, None] * wu).sum(0)
    return w2.to(dtype), bias


def xg_finish_wgrads(cfg, packs, egr, bn2_wb, bn, bn_mode, red2, BT, pwd, pwu, suma):
    rows = BT * cfg.N
    C, Cr, G = cfg.C, cfg.Cr, cfg.G
    Cg, Crg = C // G, Cr // G
    for e, (pk, gr) in enumerate(zip(packs, egr)):
        alpha, beta, gamma = _xg_bn2_coeffs(cfg, bn2_wb, bn, bn_mode, red2, rows, e)
        dense = alpha[:, None] * pwu[e, C:, :Cr]
        if bn_mode == BN_TRAIN:
            dense = dense + beta[:, None] * pwu[e, :C, :Cr] + gamma[:, None] * suma[e].float()[None, :]
        for g in range(G):
            gr.wd[g * Crg:(g + 1) * Crg] = pwd[e, g * Crg:(g + 1) * Crg, g * Cg:(g + 1) * Cg]
            gr.wu[g * Cg:(g + 1) * Cg] = dense[g * Cg:(g + 1) * Cg, g * Crg:(g + 1) * Crg]

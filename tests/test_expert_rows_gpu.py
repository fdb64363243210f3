"""Unit parity of the GEMM-based expert path's row kernels (csrc/expert_rows.cu) against their torch statements in
tests/kernel_emulation.py, kernel by kernel through the C ABI, on small and backbone-like geometries."""
import copy

import pytest
import torch

from avmoe_b200.functional import (BN_EVAL, BN_OFF, BN_TRAIN, PRE_AUX_ADD, PRE_AUX_REPLACE, PRE_NONE, AdapterConfig, ExpertGrads,
                                   ExpertPack)
from tests import kernel_emulation as KE

pytestmark = pytest.mark.gpu

# (BT, N, C, r, G, kinds per expert)
GEOMS = [
    (3, 24, 16, 8, 2, (PRE_AUX_ADD, PRE_NONE)),
    (2, 20, 32, 8, 2, (PRE_AUX_ADD, PRE_AUX_REPLACE)),
    (2, 50, 192, 8, 2, (PRE_AUX_ADD, PRE_AUX_ADD)),
    (2, 37, 768, 8, 2, (PRE_NONE, PRE_AUX_ADD)),
    (1, 19, 1536, 8, 2, (PRE_AUX_ADD, PRE_NONE, PRE_AUX_REPLACE)),
]


def _cfg(N, C, r, G, E):
    cfg = AdapterConfig.__new__(AdapterConfig)
    for k, v in dict(variant="ave", C=C, N=N, Cx=C, Nx=N, r=r, G=G, T=8, Mm=1, Ms=E - 1, use_bn=True, use_gate=True, ln_before=True,
                     ln_post=True, self_attn="none", lb_loss=False).items():
        object.__setattr__(cfg, k, v)
    return cfg


def _rand(g, *shape, scale=1.0):
    return torch.randn(*shape, generator=g) * scale


def _make(BT, N, C, r, G, kinds, ln=True, gate=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    E = len(kinds)
    cfg = _cfg(N, C, r, G, E)
    Cr = C // r
    x = _rand(g, BT, N, C)
    packs = []
    for e, kind in enumerate(kinds):
        pk = ExpertPack(pre_kind=kind, relu=(e == 0))
        if kind != PRE_NONE:
            pk.aux = _rand(g, BT, N, C)
        if kind == PRE_AUX_ADD:
            pk.gate_pre = _rand(g, 1) * 0.5 + 0.5
        if ln:
            pk.ln1_w, pk.ln1_b = _rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1
            pk.ln2_w, pk.ln2_b = _rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1
        pk.wd = _rand(g, Cr, C // G) * (C // G) ** -0.5
        pk.wu = _rand(g, C, Cr // G) * max(Cr // G, 1) ** -0.5
        if gate:
            pk.gate = _rand(g, 1) * 0.3 + 0.7
        packs.append(pk)
    probs = torch.softmax(_rand(g, BT, E), dim=-1)
    return cfg, x, packs, probs, g


def _to_cuda(packs):
    out = []
    for pk in packs:
        q = copy.copy(pk)
        for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
            t = getattr(pk, name)
            if t is not None:
                setattr(q, name, t.cuda())
        out.append(q)
    return out


def _grads(cfg, packs, BT, dev):
    out = []
    for pk in packs:
        g = ExpertGrads()
        z = lambda *s: torch.zeros(*s, device=dev)
        if pk.gate_pre is not None:
            g.gate_pre = z(1)
        if pk.pre_kind != PRE_NONE:
            g.aux = z(BT, cfg.N, cfg.C)
        if pk.ln1_w is not None:
            g.ln1_w, g.ln1_b, g.ln2_w, g.ln2_b = z(cfg.C), z(cfg.C), z(cfg.C), z(cfg.C)
        g.wd, g.wu = z(cfg.Cr, cfg.C // cfg.G), z(cfg.C, cfg.Cr // cfg.G)
        if pk.gate is not None:
            g.gate = z(1)
        out.append(g)
    return out


def _close(got, want, tol=2e-5, what=""):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = (got - want).abs().max().item()
    ref = max(want.abs().max().item(), 1e-6)
    assert err <= tol * ref, f"{what}: max err {err:.3e} vs scale {ref:.3e}"


@pytest.mark.parametrize("bn_mode", [BN_TRAIN, BN_EVAL, BN_OFF])
@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: f"BT{g[0]}_N{g[1]}_C{g[2]}_E{len(g[5])}")
def test_row_kernels_match_emulation_fp32(geom, bn_mode):
    from avmoe_b200 import ops
    BT, N, C, r, G, kinds = geom
    cfg, x, packs, probs, g = _make(BT, N, C, r, G, kinds)
    E, Cr, Crp, rows = cfg.E, cfg.Cr, KE.pad_bottleneck(cfg.Cr), BT * N
    cpk = _to_cuda(packs)
    f32 = torch.float32
    # ---- weights
    wdd, wdt = KE.xg_prep_weights(cfg, packs, f32)
    wdd_c, wdt_c = ops.xg_prep_weights(cfg, cpk, f32)
    _close(wdd_c, wdd, 0, "wdd"); _close(wdt_c, wdt, 0, "wdt")
    # ---- ln_pre forward
    xl, st1 = KE.xg_ln_pre_fwd(cfg, x, packs)
    xl_c, st1_c = ops.xg_ln_pre_fwd(cfg, x.cuda(), cpk)
    _close(xl_c, xl, what="xl"); _close(st1_c, st1, what="ln1_stats")
    # ---- bottleneck: z stats, bn1 + activation
    z = _rand(g, BT, N, E, Cr)
    _close(ops.xg_z_stats(cfg, z.cuda()).float(), KE.xg_z_stats(cfg, z).float(), 1e-5, "stats1")
    bn = bn2_wb = None
    if bn_mode != BN_OFF:
        aff1 = torch.stack([_rand(g, E, Cr) * 0.3 + 1.0, _rand(g, E, Cr) * 0.2], dim=-1).contiguous()
        save1 = torch.stack([_rand(g, E, Cr) * 0.2, _rand(g, E, Cr).abs() + 0.5], dim=-1).contiguous()
        save2 = torch.stack([_rand(g, E, C) * 0.2, _rand(g, E, C).abs() + 0.5], dim=-1).contiguous()
        bn = (aff1, save1, None, save2)
        bn2_wb = [(_rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1) for _ in range(E)]
    bn_c = tuple(t.cuda() if t is not None else None for t in bn) if bn else None
    bn2_c = [(w.cuda(), b.cuda()) for w, b in bn2_wb] if bn2_wb else None
    act = KE.xg_bn_act(cfg, z, bn[0] if bn else None, packs, f32, bn_mode)
    act_c = ops.xg_bn_act(cfg, z.cuda(), bn_c[0] if bn else None, cpk, f32, bn_mode)
    _close(act_c, act, what="a")
    wud, ybi = KE.xg_prep_up(cfg, packs, bn[3] if bn else None, bn_mode, f32)
    wud_c, ybi_c = ops.xg_prep_up(cfg, cpk, bn_c[3] if bn else None, bn_mode, f32)
    _close(wud_c, wud, 1e-6, "wud"); _close(ybi_c, ybi, 1e-6, "ybi")
    # ---- post forward / backward
    yd = _rand(g, E, rows, 2 * C)
    yd_c = yd.cuda()
    out, out_c = torch.empty(BT, N, C), torch.empty(BT, N, C, device="cuda")
    st2 = KE.xg_post_fwd(cfg, yd, bn2_wb, packs, probs, out)
    st2_c = ops.xg_post_fwd(cfg, yd_c, bn2_c, cpk, probs.cuda(), out_c)
    _close(out_c, out, what="out"); _close(st2_c, st2, what="ln2_stats")
    dout = _rand(g, BT, N, C)
    egr, egr_c = _grads(cfg, packs, BT, "cpu"), _grads(cfg, packs, BT, "cuda")
    dprobs, dprobs_c = torch.zeros(BT, E), torch.zeros(BT, E, device="cuda")
    red2 = torch.zeros(E, C, 2, dtype=torch.float64) if bn_mode != BN_OFF else None
    red2_c = red2.cuda() if red2 is not None else None
    KE.xg_post_bwd(cfg, dout, yd, st2, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr)
    ops.xg_post_bwd(cfg, dout.cuda(), yd_c, st2_c, bn2_c, bn_mode, cpk, probs.cuda(), dprobs_c, red2_c, egr_c)
    _close(yd_c[:, :, C:], yd[:, :, C:], 5e-5, "dy2"); _close(dprobs_c, dprobs, 1e-4, "dprobs")
    if red2 is not None:
        _close(red2_c.float(), red2.float(), 1e-4, "red2")
    for e in range(E):
        _close(egr_c[e].ln2_w, egr[e].ln2_w, 1e-4, "g_ln2_w"); _close(egr_c[e].ln2_b, egr[e].ln2_b, 1e-4, "g_ln2_b")
        _close(egr_c[e].gate, egr[e].gate, 1e-4, "g_gate")
    # ---- bn2-folded backward operand, bn1 backward, weight-gradient finish
    w2, bias2 = KE.xg_prep_bwd(cfg, packs, bn2_wb, bn, bn_mode, red2, BT, f32)
    w2_c, bias2_c = ops.xg_prep_bwd(cfg, cpk, bn2_c, bn_c, bn_mode, red2_c, BT, f32)
    _close(w2_c, w2, 1e-5, "w2"); _close(bias2_c, bias2, 1e-4, "bias2")
    da = _rand(g, E, rows, Crp)
    red1 = torch.zeros(E, Cr, 2, dtype=torch.float64) if bn_mode != BN_OFF else None
    red1_c = red1.cuda() if red1 is not None else None
    dz = KE.xg_bn1_bwd(cfg, da, z, bn, bn_mode, packs, red1, f32)
    dz_c = ops.xg_bn1_bwd(cfg, da.cuda(), z.cuda(), bn_c, bn_mode, cpk, red1_c, f32)
    _close(dz_c, dz, 5e-5, "dz")
    if red1 is not None:
        _close(red1_c.float(), red1.float(), 1e-4, "red1")
    pwd, pwu = _rand(g, E, Crp, C), _rand(g, E, 2 * C, Crp)
    suma = _rand(g, E, Cr).double() if bn_mode == BN_TRAIN else None
    KE.xg_finish_wgrads(cfg, packs, egr, bn2_wb, bn, bn_mode, red2, BT, pwd, pwu, suma)
    ops.xg_finish_wgrads(cfg, cpk, egr_c, bn2_c, bn_c, bn_mode, red2_c, BT, pwd.cuda(), pwu.cuda(), suma.cuda() if suma is not None else None)
    for e in range(E):
        _close(egr_c[e].wd, egr[e].wd, 1e-6, "g_wd"); _close(egr_c[e].wu, egr[e].wu, 1e-4, "g_wu")
    # ---- ln_pre backward
    dxl = _rand(g, E, rows, C)
    dpool = _rand(g, BT, C) * 0.1
    dx, dx_c = torch.empty(BT, N, C), torch.empty(BT, N, C, device="cuda")
    KE.xg_ln_pre_bwd(cfg, x, dxl, st1, packs, dpool, dx, egr)
    ops.xg_ln_pre_bwd(cfg, x.cuda(), dxl.cuda(), st1_c, cpk, dpool.cuda(), dx_c, egr_c)
    _close(dx_c, dx, 5e-5, "dx")
    for e, pk in enumerate(packs):
        _close(egr_c[e].ln1_w, egr[e].ln1_w, 1e-4, "g_ln1_w"); _close(egr_c[e].ln1_b, egr[e].ln1_b, 1e-4, "g_ln1_b")
        if pk.pre_kind != PRE_NONE:
            _close(egr_c[e].aux, egr[e].aux, 5e-5, "g_aux")
        if pk.pre_kind == PRE_AUX_ADD:
            _close(egr_c[e].gate_pre, egr[e].gate_pre, 1e-4, "g_gate_pre")


def test_row_kernels_bf16_storage():
    """bf16 streams: same kernels with bf16 loads/stores -- outputs within bf16 rounding of the fp32 statement"""
    from avmoe_b200 import ops
    BT, N, C, r, G, kinds = 2, 50, 192, 8, 2, (PRE_AUX_ADD, PRE_NONE)
    cfg, x, packs, probs, g = _make(BT, N, C, r, G, kinds)
    bf = torch.bfloat16
    xb = x.to(bf)
    packs[0].aux = packs[0].aux.to(bf)
    cpk = _to_cuda(packs)
    xl, st1 = KE.xg_ln_pre_fwd(cfg, xb.float(), [copy.copy(p) for p in packs])
    xl_c, st1_c = ops.xg_ln_pre_fwd(cfg, xb.cuda(), cpk)
    assert xl_c.dtype == bf
    _close(xl_c, xl, 1e-2, "xl bf16"); _close(st1_c, st1, 1e-4, "ln1_stats bf16")
    yd = _rand(g, cfg.E, BT * N, 2 * C).to(bf)
    out, out_c = torch.empty(BT, N, C), torch.empty(BT, N, C, device="cuda", dtype=bf)
    KE.xg_post_fwd(cfg, yd.float(), None, packs, probs, out)
    ops.xg_post_fwd(cfg, yd.cuda(), None, cpk, probs.cuda(), out_c)
    _close(out_c, out, 1e-2, "out bf16")

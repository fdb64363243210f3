"""Unit parity of the row-fused expert passes (csrc/rowfused.cu) against their torch statements in tests/kernel_emulation.py,
kernel by kernel through the C ABI: the packed tensor-core operand layout, the back pass (bn1 affine .. combine + residual),
its backward, the front pass (latent attention of a row .. down projection) and its backward."""
import pytest
import torch

from avmoe_b200.functional import BN_OFF, BN_TRAIN, PRE_AUX_ADD, PRE_AUX_REPLACE, PRE_LATENT, PRE_NONE, ExpertGrads, ExpertPack
from tests import kernel_emulation as KE
from tests.test_expert_rows_gpu import _cfg, _rand

pytestmark = pytest.mark.gpu


def _fro(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def unpack_b(buf: torch.Tensor, nb: int, Nn: int, K: int, hi_lo: bool):
    """host-side inverse of the documented packed-B layout (include/avmoe_b200.h, csrc/mma_tile.cuh) -> B [nb, K16, N16] fp32
    (hi_lo: hi + lo summed)"""
    NP = (Nn + 7) // 8 if hi_lo else (Nn + 15) // 16
    KS = (K + 15) // 16
    w = buf.cpu().view(torch.bfloat16).float().view(nb, NP, KS, 32, 4, 2)      # [.., lane, reg, elem]
    out = torch.zeros((nb, KS * 16, NP * (8 if hi_lo else 16)))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for r in range(4):
            half = r & 1
            for i in range(2):
                k = 8 * half + 2 * t + i
                for p in range(NP):
                    n = p * 8 + g if hi_lo else (2 * p + (r >> 1)) * 8 + g
                    vals = w[:, p, :, lane, r, i]                               # [nb, KS]
                    for s in range(KS):
                        out[:, 16 * s + k, n] += vals[:, s]
    return out


@pytest.mark.parametrize("Nn,K,groups,local_k,hi_lo", [(96, 12, 2, True, False), (24, 192, 2, True, False), (192, 24, 4, False, False),
                                                        (32, 96, 1, True, True), (87, 384, 1, True, True), (96, 87, 1, True, False)])
def test_pack_b_layout(Nn, K, groups, local_k, hi_lo):
    from avmoe_b200 import ops
    g = torch.Generator().manual_seed(Nn + K)
    nb = 2
    if groups == 1:
        src = torch.randn((nb, Nn, K), generator=g)
        dense = src.clone()
        sn, sk, bs = K, 1, Nn * K
    elif local_k:                     # conv weight [Nn, K/groups] (k local to its group)
        src = torch.randn((nb, Nn, K // groups), generator=g)
        dense = torch.zeros((nb, Nn, K))
        for q in range(groups):
            dense[:, q * (Nn // groups):(q + 1) * (Nn // groups), q * (K // groups):(q + 1) * (K // groups)] = src[:, q * (Nn // groups):(q + 1) * (Nn // groups)]
        sn, sk, bs = K // groups, 1, Nn * (K // groups)
    else:                             # conv weight [K, Nn/groups] used transposed (n local to its group)
        src = torch.randn((nb, K, Nn // groups), generator=g)
        dense = torch.zeros((nb, Nn, K))
        for q in range(groups):
            dense[:, q * (Nn // groups):(q + 1) * (Nn // groups), q * (K // groups):(q + 1) * (K // groups)] = \
                src[:, q * (K // groups):(q + 1) * (K // groups)].transpose(1, 2)
        sn, sk, bs = 1, Nn // groups, K * (Nn // groups)
    scale_n = torch.rand((nb, Nn), generator=g) + 0.5
    scale_k = torch.rand((nb, K), generator=g) + 0.5
    buf = ops.rf_pack_b(src.cuda(), nb=nb, bstride=bs, stride_n=sn, stride_k=sk, Nn=Nn, K=K, groups=groups, group_local_k=local_k,
                        scale_n=scale_n.cuda(), scale_k=scale_k.cuda(), hi_lo=hi_lo)
    torch.cuda.synchronize()
    got = unpack_b(buf, nb, Nn, K, hi_lo)
    want = (dense * scale_n[:, :, None] * scale_k[:, None, :]).transpose(1, 2)      # B[k][n]
    assert got[:, K:, :].abs().sum() == 0 and got[:, :, Nn:].abs().sum() == 0, "padding must be zero"
    tol = 2 ** -16 if hi_lo else 2 ** -8
    assert ((got[:, :K, :Nn] - want).abs() <= tol * want.abs() + 1e-30).all()


# (BT, N, C, r, G, E)
BACK_GEOMS = [(3, 24, 16, 8, 2, 2), (2, 37, 96, 8, 2, 2), (2, 50, 192, 8, 2, 2), (2, 36, 192, 8, 4, 3), (3, 16, 384, 8, 2, 2),
              (2, 19, 768, 8, 2, 2), (1, 40, 96, 2, 2, 4), (2, 64, 768, 8, 4, 3)]


def _make_back(BT, N, C, r, G, E, ln, gate, bn, seed=0):
    g = torch.Generator().manual_seed(seed)
    cfg = _cfg(N, C, r, G, E)
    Cr = C // r
    packs = []
    for e in range(E):
        pk = ExpertPack(pre_kind=PRE_NONE, relu=(e == 0))
        if ln:
            pk.ln2_w, pk.ln2_b = _rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1
        pk.wu = _rand(g, C, Cr // G) * max(Cr // G, 1) ** -0.5
        if gate:
            pk.gate = _rand(g, 1) * 0.3 + 0.7
        packs.append(pk)
    z = _rand(g, BT, N, E, Cr)
    probs = torch.softmax(_rand(g, BT, E), dim=-1)
    aff1 = save2 = bn2_wb = None
    if bn:
        aff1 = torch.stack([_rand(g, E, Cr) * 0.3 + 1.0, _rand(g, E, Cr) * 0.2], dim=-1).contiguous()
        save2 = torch.stack([_rand(g, E, C) * 0.2, torch.rand(E, C, generator=g) + 0.5], dim=-1).contiguous()
        bn2_wb = [(_rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1) for _ in range(E)]
    res = _rand(g, BT, N, C)
    dout = _rand(g, BT, N, C)
    return cfg, packs, z, probs, aff1, save2, bn2_wb, res, dout


def _cuda_packs(packs):
    out = []
    for pk in packs:
        q = ExpertPack(pre_kind=pk.pre_kind, relu=pk.relu)
        for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
            t = getattr(pk, name)
            setattr(q, name, None if t is None else t.cuda())
        out.append(q)
    return out


@pytest.mark.parametrize("ln,gate,bn,use_res", [(True, True, True, True), (True, False, False, False), (False, True, True, False),
                                                (False, False, False, True)])
@pytest.mark.parametrize("geom", BACK_GEOMS, ids=[str(g) for g in BACK_GEOMS])
def test_rf_back_fwd(geom, ln, gate, bn, use_res):
    from avmoe_b200 import ops
    cfg, packs, z, probs, aff1, save2, bn2_wb, res, dout = _make_back(*geom, ln, gate, bn)
    bn_mode = BN_TRAIN if bn else BN_OFF
    bf = torch.bfloat16
    resq = res.to(bf)
    BT, N, C = res.shape
    # expected (fp32 emulation on the same operands; the kernel rounds a and the folded weights to bf16)
    wud, ybi = KE.rf_prep_up(cfg, packs, save2, bn_mode)
    want = torch.empty((BT, N, C))
    KE.rf_back_fwd(cfg, z, aff1, wud, ybi, bn2_wb, bn_mode, packs, probs, want, resq.float() if use_res else None)
    cp = _cuda_packs(packs)
    pack, ybi_c = ops.rf_prep_up(cfg, cp, None if save2 is None else save2.cuda(), bn_mode)
    assert torch.allclose(ybi_c.cpu(), ybi, atol=1e-6)
    out = torch.empty((BT, N, C), device="cuda", dtype=bf)
    ops.rf_back_fwd(cfg, z.cuda(), None if aff1 is None else aff1.cuda(), pack, ybi_c, None if bn2_wb is None else [(w.cuda(), b.cuda()) for w, b in bn2_wb],
                    bn_mode, cp, probs.cuda(), out, resq.cuda() if use_res else None)
    torch.cuda.synchronize()
    assert _fro(out.float(), want) < 1.2e-2, _fro(out.float(), want)


@pytest.mark.parametrize("ln,gate,bn", [(True, True, True), (True, False, False), (False, True, True)])
@pytest.mark.parametrize("geom", BACK_GEOMS, ids=[str(g) for g in BACK_GEOMS])
def test_rf_back_bwd(geom, ln, gate, bn):
    from avmoe_b200 import ops
    cfg, packs, z, probs, aff1, save2, bn2_wb, res, dout = _make_back(*geom, ln, gate, bn, seed=1)
    bn_mode = BN_TRAIN if bn else BN_OFF
    bf = torch.bfloat16
    BT, N, C = dout.shape
    E, Cr = cfg.E, cfg.Cr
    doutq = dout.to(bf)

    def grads(dev):
        out = []
        for pk in packs:
            gr = ExpertGrads()
            if ln:
                gr.ln2_w, gr.ln2_b = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            if gate:
                gr.gate = torch.zeros(1, device=dev)
            out.append(gr)
        return out

    wud, ybi = KE.rf_prep_up(cfg, packs, save2, bn_mode)
    egr_ref = grads("cpu")
    dprobs_ref = torch.zeros((BT, E))
    red2_ref = torch.zeros((E, C, 2), dtype=torch.float64)
    yd_ref, act_ref = KE.rf_back_bwd(cfg, doutq.float(), z, aff1, wud, ybi, bn2_wb, bn_mode, packs, probs, dprobs_ref, red2_ref, egr_ref)
    cp = _cuda_packs(packs)
    pack, ybi_c = ops.rf_prep_up(cfg, cp, None if save2 is None else save2.cuda(), bn_mode)
    egr = grads("cuda")
    dprobs = torch.zeros((BT, E), device="cuda")
    red2 = torch.zeros((E, C, 2), device="cuda", dtype=torch.float64)
    yd, act = ops.rf_back_bwd(cfg, doutq.cuda(), z.cuda(), None if aff1 is None else aff1.cuda(), pack, ybi_c,
                              None if bn2_wb is None else [(w.cuda(), b.cuda()) for w, b in bn2_wb], bn_mode, cp, probs.cuda(), dprobs, red2, egr)
    torch.cuda.synchronize()
    tol = 1.5e-2
    assert _fro(act[:, :, :Cr].float(), act_ref[:, :, :Cr].float()) < tol
    assert act[:, :, Cr:].abs().sum().item() == 0
    assert _fro(yd[:, :, :C].float(), yd_ref[:, :, :C].float()) < tol
    assert _fro(yd[:, :, C:].float(), yd_ref[:, :, C:].float()) < 2 * tol
    assert _fro(dprobs, dprobs_ref) < tol
    if bn:
        assert _fro(red2[:, :, 0], red2_ref[:, :, 0]) < 2 * tol or (red2[:, :, 0].cpu() - red2_ref[:, :, 0]).abs().max() < 0.05 * red2_ref.abs().max()
        assert _fro(red2[:, :, 1], red2_ref[:, :, 1]) < 2 * tol or (red2[:, :, 1].cpu() - red2_ref[:, :, 1]).abs().max() < 0.05 * red2_ref.abs().max()
    for gr, gr_ref in zip(egr, egr_ref):
        if ln:
            assert _fro(gr.ln2_w, gr_ref.ln2_w) < tol and _fro(gr.ln2_b, gr_ref.ln2_b) < tol
        if gate:     # a sum of rows*C signed terms: the rounding noise of the bf16 operands scales with sqrt(rows*C), not with the sum
            assert abs(gr.gate.item() - gr_ref.gate.item()) < tol * max(1.0, abs(gr_ref.gate.item())) + 4e-3 * (BT * N * C) ** 0.5


# ------------------------------------------------------------------------------------------------ front pass
# (BT, N, C, r, G, T, kinds)
FRONT_GEOMS = [
    (2, 32, 96, 8, 2, 32, (PRE_LATENT, PRE_NONE)),
    (2, 48, 192, 8, 2, 32, (PRE_LATENT, PRE_NONE)),
    (2, 16, 384, 8, 2, 32, (PRE_LATENT, PRE_NONE)),
    (1, 32, 768, 8, 2, 32, (PRE_LATENT, PRE_NONE)),
    (2, 32, 96, 8, 4, 2, (PRE_LATENT, PRE_NONE, PRE_NONE)),
    (2, 32, 192, 8, 2, 87, (PRE_LATENT, PRE_AUX_REPLACE)),
    (2, 64, 96, 8, 2, 32, (PRE_LATENT, PRE_AUX_ADD)),
    (1, 32, 96, 2, 2, 87, (PRE_LATENT, PRE_LATENT, PRE_LATENT, PRE_LATENT)),
    (2, 32, 16, 8, 2, 5, (PRE_NONE, PRE_AUX_ADD)),
]


def _make_front(BT, N, C, r, G, T, kinds, ln=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    E = len(kinds)
    cfg = _cfg(N, C, r, G, E)
    object.__setattr__(cfg, "T", T)
    Cr = C // r
    x = _rand(g, BT, N, C).bfloat16()
    packs, toks = [], {}
    for e, kind in enumerate(kinds):
        pk = ExpertPack(pre_kind=kind, relu=(e == 0))
        if kind in (PRE_AUX_ADD, PRE_AUX_REPLACE):
            pk.aux = _rand(g, BT, N, C).bfloat16()
        if kind in (PRE_AUX_ADD, PRE_LATENT):
            pk.gate_pre = _rand(g, 1) * 0.3 + 0.6
        if kind == PRE_LATENT:
            toks[e] = (torch.rand(BT, T, C, generator=g) + 0.3 * _rand(g, BT, T, C)) * (8.0 / C ** 0.5)   # logits of a few units
        if ln:
            pk.ln1_w, pk.ln1_b = _rand(g, C) * 0.2 + 1.0, _rand(g, C) * 0.1
        pk.wd = _rand(g, Cr, C // G) * (C // G) ** -0.5
        packs.append(pk)
    return cfg, x, packs, toks, g


@pytest.mark.parametrize("ln", [True, False])
@pytest.mark.parametrize("geom", FRONT_GEOMS, ids=[str(g[:6]) + str(len(g[6])) for g in FRONT_GEOMS])
def test_rf_front_fwd(geom, ln):
    from avmoe_b200 import ops
    cfg, x, packs, toks, g = _make_front(*geom, ln=ln)
    BT, N, C = x.shape
    want = torch.empty((BT, N, cfg.E, cfg.Cr))
    KE.rf_front_fwd(cfg, x.float(), [_f32_aux(pk) for pk in packs], {e: (t, t) for e, t in toks.items()}, None, want)
    cp = _cuda_packs(packs)
    prep = ops.rf_prep_front(cfg, cp)
    ctoks = {e: ops.rf_pack_tokens(t.cuda()) for e, t in toks.items()}
    z = torch.empty((BT, N, cfg.E, cfg.Cr), device="cuda")
    ops.rf_front_fwd(cfg, x.cuda(), cp, ctoks, prep, z)
    torch.cuda.synchronize()
    for e in range(cfg.E):
        assert _fro(z[:, :, e], want[:, :, e]) < 1.5e-2, (e, _fro(z[:, :, e], want[:, :, e]))


def _f32_aux(pk):
    q = ExpertPack(pre_kind=pk.pre_kind, relu=pk.relu)
    for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
        t = getattr(pk, name)
        setattr(q, name, None if t is None else t.float())
    return q


@pytest.mark.parametrize("ln", [True, False])
@pytest.mark.parametrize("geom", FRONT_GEOMS, ids=[str(g[:6]) + str(len(g[6])) for g in FRONT_GEOMS])
def test_rf_front_bwd(geom, ln):
    from avmoe_b200 import ops
    cfg, x, packs, toks, g = _make_front(*geom, ln=ln, seed=1)
    BT, N, C = x.shape
    E, Cr, T = cfg.E, cfg.Cr, cfg.T
    Crp = ops.pad_bottleneck(Cr)
    dz = torch.zeros((E, BT * N, Crp))
    dz[:, :, :Cr] = _rand(g, E, BT * N, Cr)
    dz = dz.bfloat16()
    dpool = _rand(g, BT, C) * 0.1

    def grads(dev):
        out = []
        for pk in packs:
            gr = ExpertGrads()
            if pk.pre_kind != PRE_NONE:
                gr.aux = torch.empty((BT, N, C), device=dev, dtype=torch.bfloat16 if dev == "cuda" else torch.float32)
            if pk.gate_pre is not None:
                gr.gate_pre = torch.zeros(1, device=dev)
            if ln:
                gr.ln1_w, gr.ln1_b = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            out.append(gr)
        return out

    egr_ref = grads("cpu")
    dx_ref = torch.empty((BT, N, C))
    xl_ref, lat_ref = KE.rf_front_bwd(cfg, x.float(), [_f32_aux(pk) for pk in packs], egr_ref, {e: (t, t) for e, t in toks.items()}, None, None,
                                      dz.float(), dpool, dx_ref)
    cp = _cuda_packs(packs)
    prep = ops.rf_prep_front(cfg, cp)
    ctoks = {e: ops.rf_pack_tokens(t.cuda()) for e, t in toks.items()}
    egr = grads("cuda")
    dx = torch.empty((BT, N, C), device="cuda", dtype=torch.bfloat16)
    zscratch = torch.empty((BT, N, E, Cr), device="cuda")
    ln1s = ops.rf_front_fwd(cfg, x.cuda(), cp, ctoks, prep, zscratch)      # the forward saves the ln_before statistics
    xl, lat = ops.rf_front_bwd(cfg, x.cuda(), cp, egr, ctoks, prep, ln1s, dz.cuda(), dpool.cuda(), dx)
    torch.cuda.synchronize()
    tol = 2e-2
    assert _fro(xl.float(), xl_ref.float()) < tol
    assert _fro(dx.float(), dx_ref) < tol, _fro(dx.float(), dx_ref)
    for e, pk in enumerate(packs):
        if pk.pre_kind != PRE_NONE:
            assert _fro(egr[e].aux.float(), egr_ref[e].aux.float()) < tol, e
        if pk.gate_pre is not None:
            assert abs(egr[e].gate_pre.item() - egr_ref[e].gate_pre.item()) < tol * max(1.0, abs(egr_ref[e].gate_pre.item())) + 4e-3 * (BT * N * C) ** 0.5, e
        if ln:
            assert _fro(egr[e].ln1_w, egr_ref[e].ln1_w) < tol and _fro(egr[e].ln1_b, egr_ref[e].ln1_b) < tol, e
        if pk.pre_kind == PRE_LATENT:
            assert _fro(lat[e][0][:, :, :T].float(), lat_ref[e][0][:, :, :T].float()) < tol, e
            assert _fro(lat[e][1][:, :, :T].float(), lat_ref[e][1][:, :, :T].float()) < 2 * tol, e
            assert lat[e][0][:, :, T:].abs().sum().item() == 0


# ------------------------------------------------------------------------------------------------ latent-token update passes
TOK_CASES = [(3, 256, 96, 32), (2, 4096, 96, 32), (2, 144, 384, 32), (2, 64, 384, 87), (2, 576, 192, 2), (1, 2304, 192, 32)]


@pytest.mark.gpu
@pytest.mark.parametrize("BT,Nk,C,T", TOK_CASES)
def test_rf_token_update_passes(BT, Nk, C, T, monkeypatch):
    """csrc/rf_tokens.cu against plain fp32 torch on the same bf16 kv: A1ᵀ = softmax over the Nk rows of kv·tokᵀ written as bf16
    hi | lo blocks (hi + lo carries ~16 bits), and the backward's dS1ᵀ = A1ᵀ ∘ (kv·dRᵀ − D) written over the lo block."""
    from avmoe_b200 import ops
    from tests import util
    g = torch.Generator().manual_seed(5)
    kv = (torch.randn(BT, Nk, C, generator=g) + torch.randn(BT, 1, C, generator=g)).cuda().bfloat16()
    tok = torch.rand(T, C, generator=g).cuda()                       # my_tokens ~ U[0,1) like the reference (net_trans_v3.py:315)
    monkeypatch.setattr(ops, "_RF_TOK_MIN_NK", 16)                    # the size threshold is a speed choice; the kernels cover every Nk % 16 == 0
    assert ops.rf_tok_supported(kv, T)
    Tp = ops.round8(T)
    A12 = ops.rf_tok_probs(kv, tok)
    assert A12.shape == (BT, Nk, 2 * Tp)
    S = kv.float() @ tok.t()                                          # [BT,Nk,T]
    ref = torch.softmax(S, dim=1)
    got = A12[:, :, :T].float() + A12[:, :, Tp:Tp + T].float()
    assert util.rel_err(got, ref) < 2e-3, "probabilities (hi + lo)"
    assert util.rel_err(A12[:, :, :T].float(), ref) < 6e-3, "hi block alone is the bf16 rounding"
    if Tp > T:
        assert A12[:, :, T:Tp].abs().sum() == 0 and A12[:, :, Tp + T:].abs().sum() == 0, "padding columns are zero"
    assert util.rel_err(got.sum(dim=1), torch.ones(BT, T, device="cuda")) < 2e-3
    # backward pass
    dR = torch.randn(BT, T, C, generator=g).cuda()
    Dt = torch.randn(BT, T, generator=g).cuda()
    hi = A12[:, :, :T].float().clone()
    ops.rf_tok_ds(kv, dR, A12, Dt, T)
    dA = kv.float() @ dR.transpose(1, 2)                              # [BT,Nk,T]
    want = hi * (dA - Dt[:, None, :])
    assert torch.equal(A12[:, :, :T].float(), hi), "the probability block is left untouched"
    assert util.rel_err(A12[:, :, Tp:Tp + T].float(), want) < 1e-2
    if Tp > T:
        assert A12[:, :, Tp + T:].abs().sum() == 0


@pytest.mark.gpu
def test_latent_token_update_fused_matches_gemm_form(monkeypatch):
    """functional.gemm_latent_fwd/bwd with the row-fused token passes against the GEMM + column-softmax form on the same inputs"""
    from avmoe_b200 import ops, functional as F
    from tests import util
    g = torch.Generator().manual_seed(11)
    BT, Nk, C, T = 4, 1024, 192, 32
    kv = (torch.randn(BT, Nk, C, generator=g) + torch.randn(BT, 1, C, generator=g)).cuda().bfloat16()
    tok = torch.rand(T, C, generator=g).cuda()
    dtokp = torch.randn(BT, T, C, generator=g).cuda()
    res = []
    for fused in (False, True):
        monkeypatch.setattr(ops, "USE_ROWFUSED_TOKENS", fused)
        tokp, tokp_op, A12 = F.gemm_latent_fwd(ops, kv, tok)
        dkv, dtok = F.gemm_latent_bwd(ops, kv, tok, tokp, A12, dtokp.clone(), T)
        res.append((tokp, dkv.float(), dtok))
    (t0, k0, d0), (t1, k1, d1) = res
    assert util.rel_err(t1, t0) < 1e-3
    assert util.rel_err(k1, k0) < 2e-2 and util.fro_err(k1, k0) < 1e-2
    assert util.rel_err(d1, d0) < 5e-3

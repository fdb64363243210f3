"""GPU: tcgen05 GEMM (every operand-major / batching / swizzle mode the resampler uses) vs torch fp32."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _mk(shape, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to("cuda").to(dtype)


def _tol(dtype):
    return 2e-3 if dtype == torch.float32 else 1e-2   # TF32 operands / bf16 output rounding


CASES = [
    # name, M, N, K
    ("small", 128, 64, 64),
    ("ragged", 200, 96, 144),
    ("multi_tile", 384, 512, 320),
    ("tiny_n", 64, 32, 36 * 4),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("name,M,N,K", CASES)
def test_gemm_kmajor_kmajor(dtype, impl, name, M, N, K):
    from avmoe_b200 import ops
    A, B = _mk((M, K), dtype, 1), _mk((N, K), dtype, 2)
    rb, cs, cb = _mk((M,), torch.float32, 3), _mk((N,), torch.float32, 4), _mk((N,), torch.float32, 5)
    D = ops.gemm(A, B, out_dtype=torch.float32, row_bias=rb, col_scale=cs, col_bias=cb, alpha=0.5, impl=impl)
    ref = 0.5 * (A.float() @ B.float().t()) + rb[:, None] * cs[None, :] + cb[None, :]
    assert util.rel_err(D, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("Nb", [96, 64, 256, 48])
def test_gemm_nbatched_mnmajor_b(dtype, impl, Nb):
    """resampler forward: D[b] = W[M,K] · P[b][K,Nb]  (B MN-major, N-batched)."""
    from avmoe_b200 import ops
    b, M, K = 5, 200, 144
    W, P = _mk((M, K), dtype, 1), _mk((b, K, Nb), dtype, 2)
    rb, cs, cb = _mk((M,), torch.float32, 3), _mk((Nb,), torch.float32, 4), _mk((Nb,), torch.float32, 5)
    D = ops.gemm(W, P, b_kmajor=False, row_bias=rb, col_scale=cs, col_bias=cb, impl=impl)
    ref = torch.einsum("mk,bkn->bmn", W.float(), P.float()) + rb[None, :, None] * cs[None, None, :] + cb[None, None, :]
    assert D.shape == (b, M, Nb) and D.dtype == dtype
    assert util.rel_err(D, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("Nb", [96, 128])
def test_gemm_mnmajor_a_and_b(dtype, impl, Nb):
    """input gradient: dP[b] = Wᵀ[K=N_tok rows, M] · dV[b][K, Nb]  (A stored [K,M], B MN-major batched)."""
    from avmoe_b200 import ops
    b, M, K = 3, 144, 256
    Wt, dV = _mk((K, M), dtype, 1), _mk((b, K, Nb), dtype, 2)
    D = ops.gemm(Wt, dV, a_kmajor=False, b_kmajor=False, impl=impl)
    ref = torch.einsum("km,bkn->bmn", Wt.float(), dV.float())
    assert util.rel_err(D, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("Kb", [96, 64, 384])
def test_gemm_kbatched(dtype, impl, Kb):
    """weight gradient: dW[M,N] = sum_b dV[b][M,Kb] · P[b][N,Kb]ᵀ."""
    from avmoe_b200 import ops
    b, M, N = 6, 256, 144
    dV, P = _mk((b, M, Kb), dtype, 1), _mk((b, N, Kb), dtype, 2)
    D = ops.gemm(dV, P, out_dtype=torch.float32, impl=impl)
    ref = torch.einsum("bmk,bnk->mn", dV.float(), P.float())
    assert util.rel_err(D, ref) < _tol(dtype)
    if impl == 0:
        D2 = ops.gemm(dV, P, out_dtype=torch.float32, split_k=4)
        assert util.rel_err(D2, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
def test_gemm_mnmajor_both_flat_splitk(dtype, impl):
    """fc weight gradient: dWfc[C,Cx] = dPᵀ · O with rows flattened over (b,n): both operands MN-major, split-K."""
    from avmoe_b200 import ops
    rows, C, Cx = 1000, 96, 192
    dP, O = _mk((rows, C), dtype, 1), _mk((rows, Cx), dtype, 2)
    D = ops.gemm(dP, O, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32, split_k=1 if impl else 8, impl=impl)
    ref = dP.float().t() @ O.float()
    assert util.rel_err(D, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_kmajor_a_mnmajor_b_flat(dtype):
    """dU = dV[rows,C] · Wfc[C,Cx]  (B MN-major, unbatched)."""
    from avmoe_b200 import ops
    rows, C, Cx = 700, 192, 96
    dV, W = _mk((rows, C), dtype, 1), _mk((C, Cx), dtype, 2)
    D = ops.gemm(dV, W, b_kmajor=False)
    assert util.rel_err(D, dV.float() @ W.float()) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("Nb", [96, 64, 384])
def test_gemm_kmajor_b_nbatched_output(dtype, impl, Nb):
    """resampler forward in TN form: D[b][m][c] = sum_k W[m,k] * Pt[(b,c),k]  (B K-major, N-batched D)."""
    from avmoe_b200 import ops
    b, M, K = 5, 200, 144
    W, Pt = _mk((M, K), dtype, 1), _mk((b * Nb, K), dtype, 2)
    rb, cs, cb = _mk((M,), torch.float32, 3), _mk((Nb,), torch.float32, 4), _mk((Nb,), torch.float32, 5)
    D = ops.gemm(W, Pt, d_nbatch=Nb, row_bias=rb, col_scale=cs, col_bias=cb, impl=impl)
    ref = torch.einsum("mk,bnk->bmn", W.float(), Pt.float().reshape(b, Nb, K)) + rb[None, :, None] * cs[None, None, :] + cb[None, None, :]
    assert D.shape == (b, M, Nb)
    assert util.rel_err(D, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("amaj,bmaj", [(True, True), (True, False), (False, False), (False, True)])
def test_bgemm_fully_batched(dtype, impl, amaj, bmaj):
    """independent problems per batch element (AVVP token self-attention, AVS MHA projections)."""
    from avmoe_b200 import ops
    nb, M, N, K = 3, 160, 96, 128
    A = _mk((nb, M, K) if amaj else (nb, K, M), dtype, 1)
    B = _mk((nb, N, K) if bmaj else (nb, K, N), dtype, 2)
    D = ops.bgemm(A, B, a_kmajor=amaj, b_kmajor=bmaj, out_dtype=torch.float32, impl=impl, alpha=0.25)
    Af = A.float() if amaj else A.float().transpose(1, 2)
    Bf = B.float() if bmaj else B.float().transpose(1, 2)
    assert util.rel_err(D, 0.25 * Af @ Bf.transpose(1, 2)) < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_hot_shapes_stay_on_tensor_cores(dtype):
    """impl=2 refuses to fall back: the real backbone shapes must be expressible on tcgen05."""
    from avmoe_b200 import ops
    for (N, Nx, C) in ((4096, 2304, 96), (2304, 4096, 96), (256, 144, 384), (144, 256, 384)):
        W, P, dV = _mk((N, Nx), dtype, 1), _mk((2, Nx, C), dtype, 2), _mk((2, N, C), dtype, 3)
        V = ops.gemm(W, P, b_kmajor=False, impl=2)
        assert util.rel_err(V, torch.einsum("mk,bkn->bmn", W.float(), P.float())) < _tol(dtype)
        dP = ops.gemm(W, dV, a_kmajor=False, b_kmajor=False, impl=2)
        assert util.rel_err(dP, torch.einsum("km,bkn->bmn", W.float(), dV.float())) < _tol(dtype)
        dW = ops.gemm(dV, P, out_dtype=torch.float32, impl=2)
        assert util.rel_err(dW, torch.einsum("bmk,bnk->mn", dV.float(), P.float())) < _tol(dtype)


def test_gemm_resampler_shape_bf16():
    """stage-1 audio adapter forward shape at BT=8 (M=4096, K=2304, Nb=96), bf16."""
    from avmoe_b200 import ops
    W, P = _mk((4096, 2304), torch.bfloat16, 1), _mk((8, 2304, 96), torch.bfloat16, 2)
    D = ops.gemm(W, P, b_kmajor=False)
    ref = torch.einsum("mk,bkn->bmn", W.float(), P.float())
    assert util.rel_err(D, ref) < 1e-2


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_bgemm_per_problem_bias_and_strided_outputs(dtype, impl):
    """the expert-path forms: one launch for E experts with a per-expert column bias, output written through a strided view
    ([rows,E,Cr] bottleneck buffer / first half of the [rows,2C] yd buffer)."""
    from avmoe_b200 import ops
    E, rows, C, Cr = 2, 300, 192, 24
    xl, wdd = _mk((E, rows, C), dtype, 1), _mk((E, Cr, C), dtype, 2)
    z = torch.zeros((rows, E, Cr), device="cuda", dtype=torch.float32)
    ops.bgemm(xl, wdd, out=z.permute(1, 0, 2), impl=impl)
    assert util.rel_err(z.permute(1, 0, 2), xl.float() @ wdd.float().transpose(1, 2)) < _tol(dtype)
    act, wud = _mk((E, rows, Cr), dtype, 3), _mk((E, C, Cr), dtype, 4)
    bias = _mk((E, C), torch.float32, 5)
    yd = torch.full((E, rows, 2 * C), 7.0, device="cuda", dtype=dtype)
    ops.bgemm(act, wud, out=yd[:, :, :C], col_bias=bias, impl=impl)
    want = act.float() @ wud.float().transpose(1, 2) + bias[:, None, :]
    assert util.rel_err(yd[:, :, :C].float(), want) < (2e-2 if dtype == torch.bfloat16 else _tol(dtype))
    assert torch.all(yd[:, :, C:] == 7.0), "the second half of the buffer must be untouched"
    shared = ops.bgemm(act, wud, out_dtype=torch.float32, col_bias=bias[0].contiguous(), impl=impl)
    assert util.rel_err(shared, act.float() @ wud.float().transpose(1, 2) + bias[0]) < _tol(dtype)


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_bgemm_accumulates_into_existing_output(dtype, impl):
    """dx += dS2·tok' in the epilogue (bf16: read-modify-write with one rounding; fp32: atomics), interior and edge tiles"""
    from avmoe_b200 import ops
    nb, M, N, K = 3, 200, 104, 32
    A, B = _mk((nb, M, K), dtype, 1), _mk((nb, K, N), dtype, 2)
    D0 = _mk((nb, M, N), dtype, 3)
    D = D0.clone()
    ops.bgemm(A, B, b_kmajor=False, out=D, accumulate=True, impl=impl)
    want = D0.float() + A.float() @ B.float()
    assert util.rel_err(D.float(), want) < (1e-2 if dtype == torch.bfloat16 else _tol(dtype))


STAGE_SITES = [(stage, modality) for stage in (1, 2, 3, 4) for modality in ("audio", "visual")]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("stage,modality", STAGE_SITES)
def test_resampler_stays_on_tensor_cores_all_stages(stage, modality, dtype, monkeypatch):
    """The token resampler (conv_adapter + fc, AVMOE/AVE/nets/net_trans_v3.py:445-446,469-471) forward and backward at all eight
    backbone shapes with the no-fallback GEMM mode (impl=2 refuses the generic kernel): stage 4's 36-token stream is
    zero-padded to a 16-byte row pitch so that it stays expressible as a TMA tensor map."""
    from avmoe_b200 import ops, functional as F, stack as S
    monkeypatch.setattr(ops, "_GEMM_IMPL_OVERRIDE", 2)
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    cfg = F.AdapterConfig(C=C, N=N, Cx=Cx, Nx=Nx, r=8, G=2, T=32, Mm=1, Ms=1, use_bn=True, use_gate=True, ln_before=True, ln_post=True)
    BT = 2
    o = _mk((BT, Nx, Cx), dtype, 1)
    w_tok = (_mk((N, Nx, 1, 1), torch.float32, 2) / Nx ** 0.5).contiguous()
    b_tok, b_fc = _mk((N,), torch.float32, 3), _mk((C,), torch.float32, 5)
    w_fc = (_mk((C, Cx), torch.float32, 4) / Cx ** 0.5).contiguous()
    v, saved = F.resample_fwd(ops, cfg, o, w_tok, b_tok, w_fc, b_fc)
    wt = w_tok.reshape(N, Nx)
    ref = torch.einsum("nm,bmc->bnc", wt, o.float()) + b_tok[None, :, None]
    ref = ref @ w_fc.t() + b_fc
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-3
    assert util.rel_err(v.float(), ref) < tol
    dv = _mk((BT, N, C), dtype, 6)
    do, dw_tok, db_tok, dw_fc, db_fc = F.resample_bwd(ops, cfg, o, dv, saved, w_tok, b_tok, w_fc)
    o_ = o.float().requires_grad_()
    wt_, wf_ = wt.clone().requires_grad_(), w_fc.clone().requires_grad_()
    out = (torch.einsum("nm,bmc->bnc", wt_, o_) + b_tok[None, :, None]) @ wf_.t() + b_fc
    out.backward(dv.float())
    assert util.rel_err(do.float(), o_.grad) < tol
    assert util.rel_err(dw_tok.reshape(N, Nx), wt_.grad) < tol
    assert util.rel_err(dw_fc, wf_.grad) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("stage,modality", [(1, "audio"), (3, "visual"), (4, "audio"), (4, "visual")])
def test_resample_op_level_entry_matches_op_by_op_sequence(stage, modality, dtype, monkeypatch):
    """avmoe_resample_fwd / avmoe_resample_bwd (csrc/op_resample.cu: the whole launch sequence issued from C++ against a
    caller-provided workspace) against the op-by-op sequence of functional.resample_fwd/bwd on the same inputs: same kernels in the
    same order, so everything that is not a split-K atomic sum must agree bit for bit."""
    from avmoe_b200 import ops, functional as F, stack as S
    Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
    (C, N, Cx, Nx) = (Ca, Na, Cv, Nv) if modality == "audio" else (Cv, Nv, Ca, Na)
    cfg = F.AdapterConfig(C=C, N=N, Cx=Cx, Nx=Nx, r=8, G=2, T=32, Mm=1, Ms=1, use_bn=True, use_gate=True, ln_before=True, ln_post=True)
    BT = 3
    o, dv = _mk((BT, Nx, Cx), dtype, 1), _mk((BT, N, C), dtype, 6)
    w_tok = (_mk((N, Nx, 1, 1), torch.float32, 2) / Nx ** 0.5).contiguous()
    w_fc = (_mk((C, Cx), torch.float32, 4) / Cx ** 0.5).contiguous()
    b_tok, b_fc = _mk((N,), torch.float32, 3), _mk((C,), torch.float32, 5)
    monkeypatch.setattr(ops, "USE_RESAMPLE_OP", False)
    v0, saved0 = F.resample_fwd(ops, cfg, o, w_tok, b_tok, w_fc, b_fc)
    g0 = F.resample_bwd(ops, cfg, o, dv, saved0, w_tok, b_tok, w_fc)
    monkeypatch.setattr(ops, "USE_RESAMPLE_OP", True)
    before = ops.LAUNCHES
    v1, saved1 = F.resample_fwd(ops, cfg, o, w_tok, b_tok, w_fc, b_fc)
    g1 = F.resample_bwd(ops, cfg, o, dv, saved1, w_tok, b_tok, w_fc)
    assert ops.LAUNCHES - before >= 8, "the op-level entry reports the kernels it enqueued"
    assert torch.equal(v0, v1) and torch.equal(saved0[1], saved1[1])
    assert torch.equal(g0[0], g1[0]), "d(other)"
    tol = 1e-5
    for a, b, what in zip(g0[1:], g1[1:], ("dw_tok", "db_tok", "dw_fc", "db_fc")):
        assert a.shape == b.shape, what
        assert util.rel_err(b, a) < tol, what
    # no gradient wanted for `other`: d_o = NULL
    g2 = ops.resample_bwd_op(cfg, o, dv, saved1, w_tok, b_tok, w_fc, need_do=False)
    assert g2[0] is None and util.rel_err(g2[1], g0[1]) < tol


def test_resample_op_rejects_small_workspace():
    """the C ABI checks the caller-provided workspace instead of writing past it"""
    import ctypes as C
    from avmoe_b200 import _lib
    lib = _lib.load()
    lib.avmoe_resample_workspace_bytes.restype = C.c_int64
    a = _lib.ResampleArgs()
    a.dtype, a.BT, a.N, a.Nx, a.C, a.Cx = 1, 2, 64, 36, 768, 1536
    need = lib.avmoe_resample_workspace_bytes(C.byref(a), 0)
    assert need > 0 and lib.avmoe_resample_workspace_bytes(C.byref(a), 1) > 0
    t = torch.zeros(1 << 20, device="cuda", dtype=torch.uint8)
    for f in ("o", "w_tok", "b_tok", "w_fc", "b_fc", "v", "mid", "s"):
        setattr(a, f, t.data_ptr())
    a.workspace, a.workspace_bytes = t.data_ptr(), 16
    assert lib.avmoe_resample_fwd(C.byref(a), None) == -1
    assert b"workspace" in lib.avmoe_last_error()

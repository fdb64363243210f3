"""Unit parity of the streaming softmax / elementwise kernels of the GEMM-based latent attention (csrc/attn_aux.cu)
against their torch statements in tests/kernel_emulation.py, through the C ABI."""
import pytest
import torch

from tests import kernel_emulation as KE

pytestmark = pytest.mark.gpu

SHAPES = [(3, 50, 32), (2, 197, 87), (4, 1024, 32), (1, 7, 5)]   # (BT, rows per sample, T)


def _ops():
    from avmoe_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def _logits_pair(BT, Nk, T, split, seed):
    """fp32 logits buffer [BT,Nk,W] as the GEMM would leave it: one block, or hi + lo blocks at columns 0 and Tp."""
    Tp = KE.round8(T)
    full = _rand(BT, Nk, T, seed=seed, scale=4.0)
    if not split:
        S = torch.zeros(BT, Nk, Tp)
        S[:, :, :T] = full
        return S, 0, full
    S = torch.full((BT, Nk, 2 * Tp), 7.0)     # padding columns hold garbage that must be ignored
    S[:, :, :T] = 0.75 * full
    S[:, :, Tp:Tp + T] = 0.25 * full
    return S, Tp, full


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("BT,Nk,T", SHAPES)
def test_softmax_cols_and_backward(BT, Nk, T, dtype, split):
    ops = _ops()
    Tp = KE.round8(T)
    S, off, full = _logits_pair(BT, Nk, T, split, 1)
    A12, lse, p_lo = ops.softmax_cols(S.cuda(), T, off, dtype)
    ref = torch.softmax(full, dim=1)
    got = A12[:, :, :T].float().cpu()
    if p_lo:   # bf16: hi + residual reproduces the fp32 probabilities to ~2^-16
        assert p_lo == Tp
        got = got + A12[:, :, Tp:Tp + T].float().cpu()
        assert (got - ref).abs().max() < 3e-5
    else:
        assert dtype == torch.float32 and (got - ref).abs().max() < 1e-5
    assert (lse.cpu() - torch.logsumexp(full, dim=1)).abs().max() < 1e-4
    assert A12[:, :, T:Tp].abs().max().item() == 0 if Tp > T else True
    # backward: dS1ᵀ = A1ᵀ * (dA - D) written at column Tp
    dA, doff, dfull = _logits_pair(BT, Nk, T, split, 2)
    Dt = _rand(BT, T, seed=3)
    hi = A12[:, :, :T].float().cpu()
    ops.softmax_cols_bwd(A12, dA.cuda(), Dt.cuda(), T, doff)
    want = hi * (dfull - Dt[:, None, :])
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (A12[:, :, Tp:Tp + T].float().cpu() - want).abs().max() <= tol * want.abs().max()


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("BT,N,T", SHAPES)
def test_softmax_rows_and_backward(BT, N, T, dtype, split):
    ops = _ops()
    S, off, full = _logits_pair(BT, N, T, split, 4)
    A = ops.softmax_rows(S.cuda(), T, off, dtype)
    ref = torch.softmax(full, dim=-1)
    tol = 1e-6 if dtype == torch.float32 else 4e-3
    assert A.shape == (BT, N, KE.round8(T))
    assert (A[:, :, :T].float().cpu() - ref).abs().max() < tol
    assert A[:, :, T:].abs().max().item() == 0 if A.shape[2] > T else True
    dA, doff, dfull = _logits_pair(BT, N, T, split, 5)
    dS = ops.softmax_rows_bwd(A, dA.cuda(), T, doff)
    a = A[:, :, :T].float().cpu()
    want = a * (dfull - (a * dfull).sum(-1, keepdim=True))
    assert (dS[:, :, :T].float().cpu() - want).abs().max() <= (1e-5 if dtype == torch.float32 else 1e-2) * want.abs().max()
    assert dS[:, :, T:].abs().max().item() == 0 if dS.shape[2] > T else True


@pytest.mark.parametrize("T,C", [(32, 192), (87, 96), (5, 16)])
def test_split_operand_tile_add_dots_and_sums(T, C):
    ops = _ops()
    BT = 3
    Tp = KE.round8(T)
    t = _rand(BT, T, C, seed=6).cuda()
    buf, off, lp = ops.split_operand(t, torch.bfloat16)
    assert buf.shape == (BT, 2 * Tp, C) and off == Tp and lp.shape == (BT, T, C)
    assert torch.equal(lp, t.bfloat16())
    assert ((buf[:, :T].float() + buf[:, Tp:Tp + T].float()) - t).abs().max() < 2 ** -15 * t.abs().max()
    assert buf[:, T:Tp].abs().max().item() == 0 if Tp > T else True
    buf32, off32, lp32 = ops.split_operand(t, torch.float32)
    assert off32 == 0 and buf32.shape == (BT, Tp, C) and torch.equal(lp32, t) and torch.equal(buf32[:, :T], t)
    # tile_add with and without a residual block
    tile = _rand(T, C, seed=7).cuda()
    R = _rand(BT, 2 * Tp, C, seed=8).cuda()
    assert torch.allclose(ops.tile_add(R, tile, Tp), R[:, :T] + R[:, Tp:Tp + T] + tile[None], atol=1e-6)
    assert torch.allclose(ops.tile_add(R[:, :T].contiguous(), tile), R[:, :T] + tile[None], atol=1e-6)
    # latent_dots, batch_sum_add, add_rowvec
    tokp = _rand(BT, T, C, seed=9).cuda()
    assert torch.allclose(ops.latent_dots(t, tokp, tile), (t * (tokp - tile[None])).sum(-1), rtol=1e-4, atol=1e-3)
    acc = _rand(T, C, seed=10).cuda()
    want = acc + t.sum(0)
    ops.batch_sum_add(t, acc)
    assert torch.allclose(acc, want, atol=1e-5)
    for dt in (torch.float32, torch.bfloat16):
        d = _rand(BT, 11, C, seed=11).cuda().to(dt)
        vec = _rand(BT, C, seed=12).cuda()
        want = (d.float() + vec[:, None, :]).to(dt)
        ops.add_rowvec(d, vec)
        assert torch.allclose(d.float(), want.float(), atol=1e-2 if dt == torch.bfloat16 else 1e-6)


@pytest.mark.parametrize("shape", [(3, 144, 384), (2, 2304, 96), (2, 37, 50), (1, 36, 1536), (2, 65, 130)])
def test_transpose_layout_helper(shape):
    """[b,R,C] -> [b,C,R]: the 64x64 bf16 fast path (even sizes), the generic path (odd sizes) and the casting forms are exact."""
    ops = _ops()
    t = _rand(*shape, seed=13).cuda()
    tb = t.bfloat16()
    assert torch.equal(ops.transpose(tb), tb.transpose(1, 2).contiguous())
    assert torch.equal(ops.transpose(t), t.transpose(1, 2).contiguous())
    assert torch.equal(ops.transpose(t, torch.bfloat16), t.transpose(1, 2).contiguous().bfloat16())

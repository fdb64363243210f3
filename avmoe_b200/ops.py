"""Thin Python wrappers: torch tensors -> raw pointers -> C ABI calls on the current stream.

Nothing here computes; every function only validates, allocates outputs with torch and
enqueues kernels of libavmoe_b200.so.  (File/line citations of the reference ops each call
replaces are in include/avmoe_b200.h.)"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

F32, BF16 = 0, 1
# diagnostics only: AVMOE_GEMM_IMPL=1 forces the plain-fp32 SIMT GEMM (isolates TF32/bf16 tensor-core rounding in parity studies)
import os as _os
_GEMM_IMPL_OVERRIDE = int(_os.environ.get("AVMOE_GEMM_IMPL", "0"))


# kernel-launch accounting (bench.py reports it as "gpu_launches") and optional per-GEMM CUDA-event profiling
LAUNCHES = 0
GEMM_PROFILE = None   # set to a list to collect (start_event, end_event, flops, bytes, tag) per tensor-core GEMM launch

_KERNELS_PER_CALL = {"transpose": 1, "cast": 1, "pool": 1, "reduce_bias": 1, "rowsum": 1, "broadcast_rows": 1, "router_fwd": 2,
                     "router_bwd": 5, "lb_loss_fwd": 1, "lb_loss_bwd": 1, "latent_fwd": 2, "latent_bwd": 1, "expert_front_fwd": 1,
                     "bn1_finalize": 1, "bn2_finalize": 1, "bn_moments": 1, "expert_back_fwd": 1, "expert_back_bwd_stats": 1,
                     "expert_back_bwd": 1, "expert_front_bwd": 1, "selfattn_nxn_fwd": 2, "selfattn_nxn_bwd": 2, "mha_attn_fwd": 1,
                     "mha_attn_bwd": 1, "xg_bn1_bwd": 2, "rf_tok_probs": 2}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"avmoe_b200 supports float32 and bfloat16 activations, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("avmoe_b200 has no CPU path: tensors must live on a CUDA device")


def gemm(A: torch.Tensor, B: torch.Tensor, *, a_kmajor: bool = True, b_kmajor: bool = True,
         out_dtype: Optional[torch.dtype] = None, alpha: float = 1.0,
         row_bias: Optional[torch.Tensor] = None, col_scale: Optional[torch.Tensor] = None,
         col_bias: Optional[torch.Tensor] = None, split_k: int = 1, out: Optional[torch.Tensor] = None,
         accumulate: bool = False, impl: int = 0, d_nbatch: int = 0) -> torch.Tensor:
    """D = alpha * A·Bᵀ (+ biases) on the tcgen05 GEMM.  Supported operand forms:

      A [M,K] (a_kmajor) | A [b,M,K] with B [b,N,K]  (sum over b: "K-batched")
      A [K,M] (a_kmajor=False: stored transposed)
      B [N,K] (b_kmajor) | B [K,N] (b_kmajor=False) | B [b,K,Nb] (b_kmajor=False: "N-batched", D is [b,M,Nb])
      d_nbatch=Nb with a K-major B [b*Nb, K]: rows are (batch, n) pairs and D is written as [b, M, Nb]
    Inner dimension must be contiguous; row/batch pitches are taken from the strides."""
    _require_cuda(A, B)
    assert A.dtype == B.dtype
    assert A.stride(-1) == 1 and B.stride(-1) == 1, "innermost dimension must be contiguous"
    a = _lib.GemmArgs()
    a.dtype = _dt(A)
    out_dtype = out_dtype or A.dtype
    a.alpha = alpha
    a.a_kmajor, a.b_kmajor = int(a_kmajor), int(b_kmajor)
    a.a_batch = a.b_batch = 1
    if a_kmajor:
        if A.dim() == 3:
            a.a_batch, M, K = A.shape
            a.a_bstride, a.lda = A.stride(0), A.stride(1)
        else:
            M, K = A.shape
            a.lda = A.stride(0)
    else:
        assert A.dim() == 2
        K, M = A.shape
        a.lda = A.stride(0)
    n_batched = False
    if b_kmajor:
        if B.dim() == 3:
            a.b_batch, N, Kb = B.shape
            a.b_bstride, a.ldb = B.stride(0), B.stride(1)
        else:
            N, Kb = B.shape
            a.ldb = B.stride(0)
        Nb = N
    else:
        if B.dim() == 3:
            a.b_batch, Kb, Nb = B.shape
            a.b_bstride, a.ldb = B.stride(0), B.stride(1)
            N = a.b_batch * Nb
            n_batched = a.b_batch > 1
        else:
            Kb, Nb = B.shape
            a.ldb = B.stride(0)
            N = Nb
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    if d_nbatch:
        assert b_kmajor and B.dim() == 2 and N % d_nbatch == 0
        Nb = d_nbatch
    a.M, a.N, a.K, a.Nb = M, N, K, Nb
    if out is None:
        shape = (N // Nb, M, Nb) if ((not b_kmajor and B.dim() == 3) or d_nbatch) else (M, N)
        out = torch.empty(shape, device=A.device, dtype=out_dtype)
    assert out.stride(-1) == 1
    if out.dim() == 3:
        a.d_bstride, a.ldd = out.stride(0), out.stride(1)
    else:
        a.ldd = out.stride(0)
    a.d_dtype = _dt(out)
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), out.data_ptr()
    for name, t in (("row_bias", row_bias), ("col_scale", col_scale), ("col_bias", col_bias)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(a, name, t.data_ptr())
    a.split_k, a.accumulate, a.impl = split_k, int(accumulate), impl or _GEMM_IMPL_OVERRIDE
    global LAUNCHES
    LAUNCHES += 1
    if GEMM_PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.check(_lib.load().avmoe_gemm(C.byref(a), _stream()), "gemm")
        ev1.record()
        kb = a.a_batch if (a_kmajor and b_kmajor) else 1
        nbytes = A.numel() * A.element_size() + B.numel() * B.element_size() + out.numel() * out.element_size()
        GEMM_PROFILE.append((ev0, ev1, 2.0 * M * N * K * kb, nbytes, f"M{M}_N{N}_K{K * kb}"))
        return out
    _lib.check(_lib.load().avmoe_gemm(C.byref(a), _stream()), "gemm")
    return out


def bgemm(A: torch.Tensor, B: torch.Tensor, *, a_kmajor: bool = True, b_kmajor: bool = True, out_dtype: Optional[torch.dtype] = None,
          alpha: float = 1.0, col_bias: Optional[torch.Tensor] = None, impl: int = 0, out: Optional[torch.Tensor] = None,
          accumulate: bool = False) -> torch.Tensor:
    """Fully batched GEMM: D[b] = alpha * op(A[b]) · op(B[b])ᵀ for independent problems b (accumulate: added to ``out``).
    A [b,M,K] (a_kmajor) or [b,K,M];  B [b,N,K] (b_kmajor) or [b,K,N];  D [b,M,N]."""
    _require_cuda(A, B)
    assert A.dim() == 3 and B.dim() == 3 and A.shape[0] == B.shape[0] and A.dtype == B.dtype
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    nb = A.shape[0]
    M, K = (A.shape[1], A.shape[2]) if a_kmajor else (A.shape[2], A.shape[1])
    N, Kb = (B.shape[1], B.shape[2]) if b_kmajor else (B.shape[2], B.shape[1])
    assert K == Kb
    if out is None:
        out = torch.empty((nb, M, N), device=A.device, dtype=out_dtype or A.dtype)
    assert out.stride(-1) == 1 and out.shape[0] == nb
    a = _lib.GemmArgs()
    a.dtype, a.d_dtype = _dt(A), _dt(out)
    a.M, a.N, a.K, a.Nb = M, N, K, 0
    a.a_kmajor, a.b_kmajor = int(a_kmajor), int(b_kmajor)
    a.a_batch = a.b_batch = 1
    a.lda, a.ldb, a.ldd = A.stride(1), B.stride(1), out.stride(1)
    a.a_bstride, a.b_bstride, a.d_bstride = A.stride(0), B.stride(0), out.stride(0)
    a.alpha = alpha
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), out.data_ptr()
    if col_bias is not None:          # [N] shared by every problem, or [b,N] one bias per problem
        assert col_bias.dtype == torch.float32 and col_bias.stride(-1) == 1 and col_bias.shape[-1] == N
        a.col_bias = col_bias.data_ptr()
        a.col_bias_bstride = col_bias.stride(0) if col_bias.dim() == 2 else 0
    assert not accumulate or out is not None
    a.split_k, a.accumulate, a.impl, a.full_batch = 1, int(accumulate), impl or _GEMM_IMPL_OVERRIDE, nb
    global LAUNCHES
    LAUNCHES += 1
    if GEMM_PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.check(_lib.load().avmoe_gemm(C.byref(a), _stream()), "bgemm")
        ev1.record()
        nbytes = A.numel() * A.element_size() + B.numel() * B.element_size() + out.numel() * out.element_size()
        GEMM_PROFILE.append((ev0, ev1, 2.0 * nb * M * N * K, nbytes, f"b{nb}_M{M}_N{N}_K{K}"))
        return out
    _lib.check(_lib.load().avmoe_gemm(C.byref(a), _stream()), "bgemm")
    return out


# ------------------------------------------------------------------------------------------------ layout helpers
KERNEL_PROFILE = None   # set to a list to collect (name, start_event, end_event, algorithmic_bytes) per C-ABI call


def _lib_call(name, *args, _bytes: float = 0.0):
    global LAUNCHES
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
    fn = getattr(_lib.load(), "avmoe_" + name)
    if KERNEL_PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.check(fn(*args), name)
        ev1.record()
        KERNEL_PROFILE.append((name, ev0, ev1, float(_bytes)))
        return
    _lib.check(fn(*args), name)


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def transpose(t: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """[b,R,C] -> [b,C,R] (contiguous)."""
    _require_cuda(t)
    assert t.dim() == 3 and t.is_contiguous()
    b, R, Cc = t.shape
    out = torch.empty((b, Cc, R), device=t.device, dtype=out_dtype or t.dtype)
    _lib_call("transpose", _ptr(t), _ptr(out), b, R, Cc, C.c_int64(Cc), C.c_int64(R * Cc), C.c_int64(R), C.c_int64(R * Cc),
              _dt(t), _dt(out), _stream(), _bytes=_nbytes(t, out))
    return out


def prep_weight(w: torch.Tensor, act_dtype: torch.dtype, transpose_: bool) -> torch.Tensor:
    """fp32 master weight [R,C] -> GEMM operand in the activation dtype, optionally transposed to [C,R]."""
    _require_cuda(w)
    assert w.dim() == 2 and w.dtype == torch.float32
    w = w.contiguous()
    if transpose_:
        return transpose(w.unsqueeze(0), act_dtype).squeeze(0)
    if act_dtype == torch.float32:
        return w
    out = torch.empty(w.shape, device=w.device, dtype=act_dtype)
    _lib_call("cast", _ptr(w), _ptr(out), C.c_int64(w.numel()), _dt(w), _dt(out), _stream())
    return out


def rowsum(w: torch.Tensor) -> torch.Tensor:
    _require_cuda(w)
    w = w.contiguous()
    out = torch.empty(w.shape[0], device=w.device, dtype=torch.float32)
    _lib_call("rowsum", _ptr(w), _ptr(out), w.shape[0], w.shape[1], _stream())
    return out


def reduce_bias(d: torch.Tensor, col_weight=None, row_weight=None):
    _require_cuda(d)
    assert d.dim() == 3 and d.is_contiguous()
    BT, N, Cc = d.shape
    f32 = dict(device=d.device, dtype=torch.float32)
    rows_w = torch.empty(N, **f32)
    cols = torch.empty(Cc, **f32)
    cols_w = torch.empty(Cc, **f32) if row_weight is not None else None
    _lib_call("reduce_bias", _ptr(d), _ptr(col_weight), _ptr(row_weight), _ptr(rows_w), _ptr(cols), _ptr(cols_w), BT, N, Cc, _dt(d), _stream(),
              _bytes=_nbytes(d))
    return rows_w, cols, cols_w


def split_k_hint(k_rows: int, m: int, n: int) -> int:
    """split-K factor for the skinny weight-gradient GEMMs ([m,n] output, k_rows reduction)."""
    tiles = ((m + 127) // 128) * ((n + 255) // 256)
    want = max(1, (2 * 148) // tiles)
    return int(min(want, max(1, k_rows // 256)))


# ------------------------------------------------------------------------------------------------ op-level token resampler
USE_RESAMPLE_OP = True    # one C-ABI call per direction (csrc/op_resample.cu) instead of the op-by-op sequence in functional.py


def use_resample_op() -> bool:
    """The op-level entry is the product path; the per-GEMM / per-kernel profiling modes of bench.py need the individual calls."""
    return bool(USE_RESAMPLE_OP) and GEMM_PROFILE is None and KERNEL_PROFILE is None


def _resample_args(cfg, o, w_tok, b_tok, w_fc):
    _require_cuda(o, w_tok, b_tok, w_fc)
    assert o.dim() == 3 and tuple(o.shape[1:]) == (cfg.Nx, cfg.Cx)
    o = o.contiguous()
    a = _lib.ResampleArgs()
    a.dtype, a.BT, a.N, a.Nx, a.C, a.Cx = _dt(o), o.shape[0], cfg.N, cfg.Nx, cfg.C, cfg.Cx
    a.gemm_impl = _GEMM_IMPL_OVERRIDE
    w_tok, w_fc, b_tok = w_tok.reshape(cfg.N, cfg.Nx).contiguous(), w_fc.contiguous(), b_tok.contiguous()
    assert w_tok.dtype == w_fc.dtype == b_tok.dtype == torch.float32
    a.o, a.w_tok, a.b_tok, a.w_fc = o.data_ptr(), w_tok.data_ptr(), b_tok.data_ptr(), w_fc.data_ptr()
    return a, (o, w_tok, w_fc, b_tok)


def _resample_run(name, a, backward, dev):
    global LAUNCHES
    lib = _lib.load()
    lib.avmoe_resample_workspace_bytes.restype = C.c_int64
    need = int(lib.avmoe_resample_workspace_bytes(C.byref(a), int(backward)))
    ws = torch.empty(need + 256, device=dev, dtype=torch.uint8)
    base = (ws.data_ptr() + 255) // 256 * 256
    a.workspace, a.workspace_bytes = base, need
    _lib.check(getattr(lib, name)(C.byref(a), _stream()), name)
    LAUNCHES += int(a.launches)
    return ws


def resample_fwd_op(cfg, o, w_tok, b_tok, w_fc, b_fc):
    """o [BT,Nx,Cx] -> (v [BT,N,C], saved) through avmoe_resample_fwd (same contract as functional.resample_fwd)."""
    a, keep = _resample_args(cfg, o, w_tok, b_tok, w_fc)
    BT, fc_first = o.shape[0], cfg.Cx > cfg.C
    b_fc = b_fc.contiguous()
    v = torch.empty((BT, cfg.N, cfg.C), device=o.device, dtype=o.dtype)
    mid = torch.empty((BT, cfg.Nx, cfg.C) if fc_first else (BT, cfg.N, cfg.Cx), device=o.device, dtype=o.dtype)
    s = torch.empty(cfg.C, device=o.device, dtype=torch.float32) if fc_first else None
    a.b_fc, a.v, a.mid, a.s = b_fc.data_ptr(), v.data_ptr(), mid.data_ptr(), None if s is None else s.data_ptr()
    _resample_run("avmoe_resample_fwd", a, False, o.device)
    return v, (fc_first, mid, s)


def resample_bwd_op(cfg, o, dv, saved, w_tok, b_tok, w_fc, need_do=True):
    """Gradients through avmoe_resample_bwd: (do, dw_tok [N,Nx,1,1], db_tok, dw_fc, db_fc) (same contract as functional.resample_bwd)."""
    fc_first, mid, s = saved
    a, keep = _resample_args(cfg, o, w_tok, b_tok, w_fc)
    _require_cuda(dv, mid)
    assert dv.dtype == o.dtype and mid.is_contiguous()
    dv = dv.contiguous()
    f32 = dict(device=o.device, dtype=torch.float32)
    do = torch.empty(o.shape, device=o.device, dtype=o.dtype) if need_do else None
    dw_tok, db_tok = torch.empty((cfg.N, cfg.Nx), **f32), torch.empty(cfg.N, **f32)
    dw_fc, db_fc = torch.empty((cfg.C, cfg.Cx), **f32), torch.empty(cfg.C, **f32)
    a.mid, a.s, a.dv = mid.data_ptr(), None if s is None else s.data_ptr(), dv.data_ptr()
    a.d_o = None if do is None else do.data_ptr()
    a.dw_tok, a.db_tok, a.dw_fc, a.db_fc = dw_tok.data_ptr(), db_tok.data_ptr(), dw_fc.data_ptr(), db_fc.data_ptr()
    _resample_run("avmoe_resample_bwd", a, True, o.device)
    return do, dw_tok.reshape(cfg.N, cfg.Nx, 1, 1), db_tok, dw_fc, db_fc


def pool(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    assert x.dim() == 3 and x.is_contiguous()
    BT, N, Cc = x.shape
    chunks = _lib.load().avmoe_pool_chunks(BT, N)
    part = torch.empty((BT, chunks, Cc), device=x.device, dtype=torch.float32)
    _lib_call("pool", _ptr(x), _ptr(part), BT, N, Cc, chunks, _dt(x), _stream(), _bytes=_nbytes(x))
    return part


def broadcast_rows(dpool: torch.Tensor, n: int, dtype: torch.dtype) -> torch.Tensor:
    BT, Cc = dpool.shape
    out = torch.empty((BT, n, Cc), device=dpool.device, dtype=dtype)
    _lib_call("broadcast_rows", _ptr(dpool), _ptr(out), BT, n, Cc, _dt(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------ router
def router_fwd(pool_x, pool_v, inv_n, rw, noise):
    _require_cuda(pool_x, pool_v)
    w0, b0, w2, b2, w4, b4 = [t.contiguous() for t in rw]
    BT, chx, Cc = pool_x.shape
    chv = pool_v.shape[1]
    E = w4.shape[0]
    f32 = dict(device=pool_x.device, dtype=torch.float32)
    probs = torch.empty((BT, E), **f32)
    idx = torch.empty((BT,), device=pool_x.device, dtype=torch.int64)
    h1, h2, rin = torch.empty((BT, 128), **f32), torch.empty((BT, 32), **f32), torch.empty((BT, 2 * Cc), **f32)
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
    _lib_call("router_fwd", _ptr(pool_x), _ptr(pool_v), chx, chv, BT, Cc, E, C.c_float(inv_n), _ptr(w0), _ptr(b0), _ptr(w2), _ptr(b2),
              _ptr(w4), _ptr(b4), _ptr(noise), _ptr(probs), _ptr(idx), _ptr(h1), _ptr(h2), _ptr(rin), _stream())
    return probs, idx, h1, h2, rin


def router_bwd(dprobs, probs, h1, h2, rin, inv_n, rw):
    w0, b0, w2, b2, w4, b4 = [t.contiguous() for t in rw]
    BT, E = probs.shape
    Cc = rin.shape[1] // 2
    f32 = dict(device=probs.device, dtype=torch.float32)
    ws = torch.empty(BT * (E + 32 + 128), **f32)
    dpx, dpv = torch.empty((BT, Cc), **f32), torch.empty((BT, Cc), **f32)
    grads = [torch.empty_like(t) for t in (w0, b0, w2, b2, w4, b4)]
    _lib_call("router_bwd", _ptr(dprobs.contiguous()), _ptr(probs), _ptr(h1), _ptr(h2), _ptr(rin), BT, Cc, E, C.c_float(inv_n),
              _ptr(w0), _ptr(w2), _ptr(w4), _ptr(ws), _ptr(dpx), _ptr(dpv), *[_ptr(g) for g in grads], _stream())
    return dpx, dpv, grads


def lb_loss_fwd(probs):
    loss = torch.empty((), device=probs.device, dtype=torch.float32)
    _lib_call("lb_loss_fwd", _ptr(probs), _ptr(loss), probs.shape[0], probs.shape[1], _stream())
    return loss


def lb_loss_bwd(probs, dlb, dprobs):
    _lib_call("lb_loss_bwd", _ptr(probs), _ptr(dlb.contiguous()), _ptr(dprobs), probs.shape[0], probs.shape[1], _stream())


# ------------------------------------------------------------------------------------------------ latent attention
def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def latent_fwd(kv, tokens, slots, tokp, lse):
    _require_cuda(kv)
    assert kv.is_contiguous()
    BT, Nk, Cc = kv.shape
    T = tokens[0].shape[0]
    nl = tokp.shape[1]
    tokens = [t.contiguous() for t in tokens]
    lib = _lib.load()
    lib.avmoe_latent_fwd_workspace.restype = C.c_int64
    nws = lib.avmoe_latent_fwd_workspace(BT, Nk, Cc, T, len(tokens))
    ws = torch.empty(nws, device=kv.device, dtype=torch.float32)
    _lib_call("latent_fwd", _ptr(kv), _ptr_array(tokens), (C.c_int * len(slots))(*slots), len(tokens), nl, _ptr(tokp), _ptr(lse), _ptr(ws),
              BT, Nk, Cc, T, _dt(kv), _stream(), _bytes=_nbytes(kv))


def latent_bwd(kv, tokens, slots, tokp, lse, dtokp, dpool, dkv, accumulate, dtokens):
    BT, Nk, Cc = kv.shape
    T = tokens[0].shape[0]
    nl = tokp.shape[1]
    tokens = [t.contiguous() for t in tokens]
    _lib_call("latent_bwd", _ptr(kv), _ptr_array(tokens), (C.c_int * len(slots))(*slots), len(tokens), nl, _ptr(tokp), _ptr(lse), _ptr(dtokp),
              _ptr(dpool), _ptr(dkv), int(bool(accumulate)), _ptr_array(dtokens), BT, Nk, Cc, T, _dt(kv), _stream(), _bytes=_nbytes(kv, dkv))


# ------------------------------------------------------------------------------------------------ expert kernels
def _expert_args(cfg, packs, egr=None, *, BT, dtype, bn_mode=0, nl=0):
    a = _lib.ExpertArgs()
    a.BT, a.N, a.C, a.Cr, a.G, a.T, a.E, a.nl = BT, cfg.N, cfg.C, cfg.Cr, cfg.G, cfg.T, cfg.E, nl
    a.dtype, a.bn_mode = dtype, bn_mode
    keep = []
    for e, pk in enumerate(packs):
        ex = a.ex[e]
        ex.pre_kind, ex.relu, ex.latent_slot = pk.pre_kind, int(pk.relu), pk.latent_slot
        for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
            t = getattr(pk, name)
            if t is not None:
                t = t.contiguous()
                keep.append(t)
                setattr(ex, name, t.data_ptr())
        if egr is not None:
            g = egr[e]
            for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
                t = getattr(g, name)
                if t is not None:
                    assert t.is_contiguous()
                    setattr(ex, "g_" + name, t.data_ptr())
    return a, keep


def _set(a, **tensors):
    for k, t in tensors.items():
        if t is not None:
            setattr(a, k, t.data_ptr())


def _bn_tensors(a, bn):
    if bn is not None:
        _set(a, aff1=bn[0], save1=bn[1], aff2=bn[2], save2=bn[3])


def expert_front_fwd(cfg, x, tokp, packs, z, stats1):
    _require_cuda(x)
    a, keep = _expert_args(cfg, packs, BT=x.shape[0], dtype=_dt(x), nl=0 if tokp is None else tokp.shape[1])
    _set(a, x=x, tokp=tokp, z=z, stats1=stats1)
    _lib_call("expert_front_fwd", C.byref(a), _stream(), _bytes=_nbytes(x, z))


def bn1_finalize(cfg, stats1, n, bn_w, bn_b, train):
    return _bn_finalize(cfg, 1, stats1, None, n, None, bn_w, bn_b, train)


def bn2_finalize(cfg, moments, n, packs, bn_w, bn_b, train):
    return _bn_finalize(cfg, 2, None, moments, n, packs, bn_w, bn_b, train)


def _bn_finalize(cfg, which, stats1, moments, n, packs, bn_w, bn_b, train):
    dev = bn_w[0][0].device
    a = _lib.BnArgs()
    a.E, a.C, a.Cr, a.G, a.train, a.n = cfg.E, cfg.C, cfg.Cr, cfg.G, int(train), n
    keep = []
    for e in range(cfg.E):
        for name, t in zip(("w1", "b1", "w2", "b2"), bn_w[e]):
            t = t.contiguous(); keep.append(t)
            getattr(a, name)[e] = t.data_ptr()
        for name, t in zip(("rm1", "rv1", "rm2", "rv2"), bn_b[e]):
            assert t.is_contiguous()
            getattr(a, name)[e] = t.data_ptr()
        if packs is not None:
            t = packs[e].wu.contiguous(); keep.append(t)
            a.wu[e] = t.data_ptr()
    ch = cfg.Cr if which == 1 else cfg.C
    aff = torch.empty((cfg.E, ch, 2), device=dev, dtype=torch.float32)
    save = torch.empty((cfg.E, ch, 2), device=dev, dtype=torch.float32)
    if which == 1:
        _set(a, stats1=stats1, aff1=aff, save1=save)
        _lib_call("bn1_finalize", C.byref(a), _stream())
    else:
        if moments is not None:
            _set(a, m1=moments[0], m2=moments[1])
        _set(a, aff2=aff, save2=save)
        _lib_call("bn2_finalize", C.byref(a), _stream())
    return aff, save


def bn_moments(cfg, z, aff1, packs):
    crg = cfg.Cr // cfg.G
    m1 = torch.empty((cfg.E, cfg.Cr), device=z.device, dtype=torch.float64)
    m2 = torch.empty((cfg.E, cfg.G, crg, crg), device=z.device, dtype=torch.float64)
    a, keep = _expert_args(cfg, packs, BT=z.shape[0], dtype=F32, bn_mode=1)
    _set(a, z=z, aff1=aff1)
    _lib_call("bn_moments", C.byref(a), _ptr(m1), _ptr(m2), _stream(), _bytes=_nbytes(z))
    return m1, m2


def _bn_mode_of(bn, explicit=None):
    if explicit is not None:
        return explicit
    return 0 if bn is None else 1


def expert_back_fwd(cfg, z, bn, packs, probs, out):
    a, keep = _expert_args(cfg, packs, BT=z.shape[0], dtype=_dt(out), bn_mode=_bn_mode_of(bn))
    _set(a, z=z, probs=probs, out=out)
    _bn_tensors(a, bn)
    _lib_call("expert_back_fwd", C.byref(a), _stream(), _bytes=_nbytes(z, out))


def expert_back_bwd_stats(cfg, dout, z, bn, packs, probs, red2):
    a, keep = _expert_args(cfg, packs, BT=z.shape[0], dtype=_dt(dout), bn_mode=1)
    _set(a, z=z, probs=probs, dout=dout, red2=red2)
    _bn_tensors(a, bn)
    _lib_call("expert_back_bwd_stats", C.byref(a), _stream(), _bytes=_nbytes(dout, z))


def expert_back_bwd(cfg, dout, z, bn, bn_mode, red2, red2b, packs, probs, dzt, dprobs, red1, egr):
    a, keep = _expert_args(cfg, packs, egr, BT=z.shape[0], dtype=_dt(dout), bn_mode=bn_mode)
    _set(a, z=z, probs=probs, dout=dout, red2=red2, red2b=red2b, red1=red1, dzt=dzt, dprobs=dprobs)
    _bn_tensors(a, bn)
    _lib_call("expert_back_bwd", C.byref(a), _stream(), _bytes=_nbytes(dout, z, dzt))


def expert_front_bwd(cfg, x, tokp, dzt, z, bn, bn_mode, red1, packs, dpool_x, dx, dtokp, egr):
    a, keep = _expert_args(cfg, packs, egr, BT=x.shape[0], dtype=_dt(x), bn_mode=bn_mode, nl=0 if tokp is None else tokp.shape[1])
    _set(a, x=x, tokp=tokp, z=z, dzt=dzt, red1=red1, dpool_x=dpool_x, dx=dx, dtokp=dtokp)
    _bn_tensors(a, bn)
    _lib_call("expert_front_bwd", C.byref(a), _stream(), _bytes=_nbytes(x, z, dzt, dx))


# ------------------------------------------------------------------------------------------------ GEMM-based latent attention helpers
def fp32_on_tensor_cores() -> bool:
    """fp32 activations: TF32 tensor-core orchestration iff torch itself is allowed TF32 matmuls (off by default, like torch)."""
    return bool(torch.backends.cuda.matmul.allow_tf32)


USE_GEMM_ATTENTION = True   # latent attention through tcgen05 GEMMs + streaming softmax kernels (functional.py "gemm attention")


def round8(t: int) -> int:
    return (t + 7) // 8 * 8


def pad_bottleneck(cr: int) -> int:
    """Row length of the bottleneck-sized GEMM operands (a, dz, da): C/r rounded up to 32 elements, i.e. 64-byte bf16 rows.
    TMA fetches an operand row by row; with 48-byte rows (C/r = 24) the K=24 GEMMs ran SLOWER than K=64 ones
    (profiles/r01_summary.md), so the padding columns (zeros) pay for themselves."""
    return (cr + 31) // 32 * 32


def cast_to(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if t.dtype == dtype:
        return t
    t = t.contiguous()
    out = torch.empty(t.shape, device=t.device, dtype=dtype)
    _lib_call("cast", _ptr(t), _ptr(out), C.c_int64(t.numel()), _dt(t), _dt(out), _stream())
    return out


def split_operand(t: torch.Tensor, dtype: torch.dtype):
    """fp32 [..,T,C] -> (buf [..,W,C] in ``dtype``, off_lo, lp): the small fp32 side of a logits GEMM as a tensor-core operand,
    plus ``lp`` [..,T,C], t rounded once to ``dtype`` (a view of buf) for the contractions that are not logits.
    bf16: W = 2*Tp, rows [0,T) = hi = bf16(t), rows [Tp,Tp+T) = lo = bf16(t - hi) -- the GEMM then yields the logits as two
    column blocks whose sum carries ~16 mantissa bits of t, which the un-scaled softmaxes need (DESIGN.md "numerics").
    fp32: W = Tp, off_lo = 0 (TF32 operand).  Padding rows are zero."""
    t = t.contiguous()
    T, Cc = t.shape[-2], t.shape[-1]
    Tp = round8(T)
    lead = t.shape[:-2]
    if dtype == torch.float32:
        if Tp == T:
            return t, 0, t
        buf = torch.zeros(lead + (Tp, Cc), device=t.device, dtype=dtype)
        buf[..., :T, :] = t
        return buf, 0, buf[..., :T, :]
    assert dtype == torch.bfloat16
    alloc = torch.empty if Tp == T else torch.zeros       # cast_split writes rows [0,T) and [Tp,Tp+T): only padding rows need the zero fill
    buf = alloc(lead + (2 * Tp, Cc), device=t.device, dtype=dtype)
    groups = t.numel() // (T * Cc)
    _lib_call("cast_split", _ptr(t), _ptr(buf), C.c_int64(groups), T, Tp, Cc, _stream())
    return buf, Tp, buf[..., :T, :]


def softmax_cols(S: torch.Tensor, T: int, off_lo: int, dtype: torch.dtype):
    """S [BT,Nk,W] fp32 logits (columns [0,T) + columns [off_lo,off_lo+T) when off_lo) -> A12 [BT,Nk,2*Tp] (A1ᵀ in columns
    [0,T), rest zero), lse [BT,T]."""
    BT, Nk, W = S.shape
    Tp = round8(T)
    # bf16 with Tp == T: the kernel writes every column (hi | lo blocks); otherwise padding / the second block start as zeros
    alloc = torch.empty if (Tp == T and dtype == torch.bfloat16) else torch.zeros
    A12 = alloc((BT, Nk, 2 * Tp), device=S.device, dtype=dtype)
    lse = torch.empty((BT, T), device=S.device, dtype=torch.float32)
    p_lo = Tp if dtype == torch.bfloat16 else 0      # bf16: rounding residual of A1 in columns [Tp,Tp+T) (until the backward reuses them)
    _lib_call("softmax_cols", _ptr(S), _ptr(A12), _ptr(lse), BT, Nk, T, W, 2 * Tp, off_lo, p_lo, _dt(A12), _stream(),
              _bytes=2 * _nbytes(S) + _nbytes(A12) / 2)
    return A12, lse, p_lo


def softmax_cols_bwd(A12: torch.Tensor, dA: torch.Tensor, Dt: torch.Tensor, T: int, off_lo: int):
    BT, Nk, Tp2 = A12.shape
    _lib_call("softmax_cols_bwd", _ptr(A12), _ptr(dA), _ptr(Dt), BT, Nk, T, Tp2, Tp2 // 2, dA.shape[2], off_lo, _dt(A12), _stream(),
              _bytes=_nbytes(dA) + _nbytes(A12))


def softmax_rows(S: torch.Tensor, T: int, off_lo: int, dtype: torch.dtype) -> torch.Tensor:
    BT, N, W = S.shape
    A = torch.empty((BT, N, round8(T)), device=S.device, dtype=dtype)
    _lib_call("softmax_rows", _ptr(S), _ptr(A), C.c_int64(BT * N), T, W, off_lo, A.shape[2], _dt(A), _stream(), _bytes=_nbytes(S, A))
    return A


def softmax_rows_bwd(A: torch.Tensor, dA: torch.Tensor, T: int, off_lo: int) -> torch.Tensor:
    BT, N, Tp = A.shape
    dS = torch.empty_like(A)
    _lib_call("softmax_rows_bwd", _ptr(A), _ptr(dA), _ptr(dS), C.c_int64(BT * N), T, Tp, dA.shape[2], off_lo, _dt(A), _stream(),
              _bytes=_nbytes(A, dA, dS))
    return dS


def tile_add(R: torch.Tensor, tile: torch.Tensor, off_lo: int = 0) -> torch.Tensor:
    """R [BT,rows,C] fp32 (rows [0,T) + rows [off_lo,off_lo+T) when off_lo) + tile [T,C] -> [BT,T,C] fp32."""
    BT, rows, Cc = R.shape
    T = tile.shape[0]
    out = torch.empty((BT, T, Cc), device=R.device, dtype=torch.float32)
    _lib_call("tile_add", _ptr(R), _ptr(tile.contiguous()), _ptr(out), BT, T, Cc, rows, off_lo, _stream())
    return out


def latent_dots(dR, tokp, tok):
    BT, T, Cc = dR.shape
    Dt = torch.empty((BT, T), device=dR.device, dtype=torch.float32)
    _lib_call("latent_dots", _ptr(dR), _ptr(tokp), _ptr(tok.contiguous()), _ptr(Dt), BT, T, Cc, _stream())
    return Dt


def batch_sum_add(t, out):
    """out[...] += sum over the leading axis of t (fp32)."""
    t = t.contiguous()
    _lib_call("batch_sum_add", _ptr(t), _ptr(out), t.shape[0], C.c_int64(out.numel()), _stream())


def add_rowvec(d, vec):
    BT, N, Cc = d.shape
    _lib_call("add_rowvec", _ptr(d), _ptr(vec), BT, N, Cc, _dt(d), _stream(), _bytes=2 * _nbytes(d))


def add_into(dst, a, b=None, c=None):
    _lib_call("add_into", _ptr(dst), _ptr(a), _ptr(b), _ptr(c), C.c_int64(dst.numel()), _dt(dst), _stream(),
              _bytes=2 * _nbytes(dst) + _nbytes(a, b, c))


# ------------------------------------------------------------------------------------------------ GEMM-based expert path
USE_GEMM_EXPERTS = True   # bf16 streams: expert projections on the tcgen05 GEMM + streaming row kernels (functional.py "gemm experts")


def _xg_args(cfg, packs, egr=None, *, BT, dtype, bn_mode=0, bn2_wb=None):
    a = _lib.XgArgs()
    a.BT, a.N, a.C, a.Cr, a.Crp, a.G, a.E = BT, cfg.N, cfg.C, cfg.Cr, pad_bottleneck(cfg.Cr), cfg.G, cfg.E
    a.dtype, a.bn_mode = dtype, bn_mode
    keep = []
    for e, pk in enumerate(packs):
        ex = a.ex[e]
        # latent experts reach these row kernels only on the row-fused path, whose front passes own the pre-transform:
        # the remaining xg kernels (statistics, bn1 backward, weight-gradient finish) never look at x
        ex.pre_kind, ex.relu = (0 if pk.pre_kind == 1 else pk.pre_kind), int(pk.relu)
        for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
            t = getattr(pk, name)
            if t is not None:
                t = t.contiguous()
                keep.append(t)
                setattr(ex, name, t.data_ptr())
        if egr is not None:
            g = egr[e]
            for name in ("gate_pre", "aux", "ln1_w", "ln1_b", "wd", "wu", "ln2_w", "ln2_b", "gate"):
                t = getattr(g, name)
                if t is not None:
                    assert t.is_contiguous()
                    setattr(ex, "g_" + name, t.data_ptr())
        if bn2_wb is not None:
            w, b = bn2_wb[e]
            w, b = w.contiguous(), b.contiguous()
            keep += [w, b]
            a.bn2_w[e], a.bn2_b[e] = w.data_ptr(), b.data_ptr()
    return a, keep


def xg_prep_weights(cfg, packs, dtype):
    dev = packs[0].wd.device
    Crp = pad_bottleneck(cfg.Cr)
    wdd = torch.empty((cfg.E, cfg.Cr, cfg.C), device=dev, dtype=dtype)
    wdt = torch.empty((cfg.E, cfg.C, Crp), device=dev, dtype=dtype)
    a, keep = _xg_args(cfg, packs, BT=1, dtype=_dt(wdd))
    _lib_call("xg_prep_weights", C.byref(a), _ptr(wdd), _ptr(wdt), _stream())
    return wdd, wdt


def xg_prep_up(cfg, packs, save2, bn_mode, dtype):
    dev = packs[0].wu.device
    wud = torch.empty((cfg.E, cfg.C, pad_bottleneck(cfg.Cr)), device=dev, dtype=dtype)
    ybi = torch.empty((cfg.E, cfg.C), device=dev, dtype=torch.float32)
    a, keep = _xg_args(cfg, packs, BT=1, dtype=_dt(wud), bn_mode=bn_mode)
    _set(a, save2=save2)
    _lib_call("xg_prep_up", C.byref(a), _ptr(wud), _ptr(ybi), _stream())
    return wud, ybi


def xg_ln_pre_fwd(cfg, x, packs):
    """-> xl [E,BT*N,C] (x.dtype), ln1_stats [E,BT*N,2] fp32 (None without ln_before)"""
    _require_cuda(x)
    BT, N, Cc = x.shape
    xl = torch.empty((cfg.E, BT * N, Cc), device=x.device, dtype=x.dtype)
    st = torch.empty((cfg.E, BT * N, 2), device=x.device, dtype=torch.float32) if packs[0].ln1_w is not None else None
    a, keep = _xg_args(cfg, packs, BT=BT, dtype=_dt(x))
    _set(a, x=x, xl=xl, ln1_stats=st)
    aux_bytes = sum(_nbytes(pk.aux) for pk in packs)
    _lib_call("xg_ln_pre_fwd", C.byref(a), _stream(), _bytes=_nbytes(x, xl) + aux_bytes)
    return xl, st


def xg_z_stats(cfg, z):
    stats1 = torch.zeros((cfg.E, cfg.Cr, 2), device=z.device, dtype=torch.float64)
    a = _lib.XgArgs()
    a.BT, a.N, a.C, a.Cr, a.Crp, a.G, a.E, a.dtype = z.shape[0], cfg.N, cfg.C, cfg.Cr, pad_bottleneck(cfg.Cr), cfg.G, cfg.E, F32
    _set(a, z=z, stats1=stats1)
    _lib_call("xg_z_stats", C.byref(a), _stream(), _bytes=_nbytes(z))
    return stats1


def xg_bn_act(cfg, z, aff1, packs, dtype, bn_mode):
    """-> a [E,BT*N,Crp] = act(bn1(z)) in ``dtype`` (padding columns zero)"""
    BT, N = z.shape[0], z.shape[1]
    act = torch.empty((cfg.E, BT * N, pad_bottleneck(cfg.Cr)), device=z.device, dtype=dtype)
    a, keep = _xg_args(cfg, packs, BT=BT, dtype=_dt(act), bn_mode=bn_mode)
    _set(a, z=z, aff1=aff1)
    a.a = act.data_ptr()
    _lib_call("xg_bn_act", C.byref(a), _stream(), _bytes=_nbytes(z, act))
    return act


def xg_post_fwd(cfg, yd, bn2_wb, packs, probs, out, res=None):
    """res [BT,N,C] (optional): the residual stream, folded into the combine epilogue (out = res + sum_e ...)."""
    BT, N = out.shape[0], out.shape[1]
    st = torch.empty((cfg.E, BT * N, 2), device=out.device, dtype=torch.float32) if packs[0].ln2_w is not None else None
    a, keep = _xg_args(cfg, packs, BT=BT, dtype=_dt(out), bn_mode=int(bn2_wb is not None), bn2_wb=bn2_wb)
    _set(a, yd=yd, ln2_stats=st, probs=probs, out=out, res=res)
    _lib_call("xg_post_fwd", C.byref(a), _stream(), _bytes=_nbytes(yd) / 2 + _nbytes(out, res))
    return st


def xg_post_bwd(cfg, dout, yd, ln2_stats, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr):
    BT = dout.shape[0]
    a, keep = _xg_args(cfg, packs, egr, BT=BT, dtype=_dt(dout), bn_mode=bn_mode, bn2_wb=bn2_wb)
    _set(a, yd=yd, ln2_stats=ln2_stats, probs=probs, dout=dout, dprobs=dprobs, red2=red2)
    _lib_call("xg_post_bwd", C.byref(a), _stream(), _bytes=cfg.E * _nbytes(dout) + _nbytes(yd))
    return None


def xg_bn1_bwd(cfg, da, z, bn, bn_mode, packs, red1, dtype):
    """-> dz [E,BT*N,Crp] in ``dtype``; red1 [E,Cr,2] += (sum dh, sum dh*zhat)"""
    BT = z.shape[0]
    dz = torch.empty(da.shape, device=da.device, dtype=dtype)
    a, keep = _xg_args(cfg, packs, BT=BT, dtype=_dt(dz), bn_mode=bn_mode)
    _set(a, z=z, da=da, dz=dz, red1=red1)
    if bn is not None:
        _set(a, aff1=bn[0], save1=bn[1])
    _lib_call("xg_bn1_bwd", C.byref(a), _stream(), _bytes=2 * _nbytes(da) + 2 * _nbytes(z) + _nbytes(dz))
    return dz


def xg_ln_pre_bwd(cfg, x, dxl, ln1_stats, packs, dpool_x, dx, egr):
    a, keep = _xg_args(cfg, packs, egr, BT=x.shape[0], dtype=_dt(x))
    _set(a, x=x, dxl=dxl, ln1_stats=ln1_stats, dpool_x=dpool_x, dx=dx)
    aux_bytes = sum(2 * _nbytes(pk.aux) for pk in packs)
    _lib_call("xg_ln_pre_bwd", C.byref(a), _stream(), _bytes=_nbytes(dxl) + (cfg.E + 1) * _nbytes(x) + aux_bytes)


def xg_prep_bwd(cfg, packs, bn2_wb, bn, bn_mode, red2, BT, dtype):
    dev = packs[0].wu.device
    Crp = pad_bottleneck(cfg.Cr)
    w2 = torch.empty((cfg.E, Crp, 2 * cfg.C), device=dev, dtype=dtype)
    bias = torch.empty((cfg.E, Crp), device=dev, dtype=torch.float32)
    a, keep = _xg_args(cfg, packs, BT=BT, dtype=_dt(w2), bn_mode=bn_mode, bn2_wb=bn2_wb)
    _set(a, red2=red2)
    if bn is not None:
        _set(a, save2=bn[3])
    _lib_call("xg_prep_bwd", C.byref(a), _ptr(w2), _ptr(bias), _stream())
    return w2, bias


def xg_finish_wgrads(cfg, packs, egr, bn2_wb, bn, bn_mode, red2, BT, pwd, pwu, suma):
    a, keep = _xg_args(cfg, packs, egr, BT=BT, dtype=F32, bn_mode=bn_mode, bn2_wb=bn2_wb)
    _set(a, red2=red2)
    if bn is not None:
        _set(a, save2=bn[3])
    _lib_call("xg_finish_wgrads", C.byref(a), _ptr(pwd), _ptr(pwu), _ptr(suma), _stream())


# ------------------------------------------------------------------------------------------------ row-fused expert passes
USE_ROWFUSED = True   # bf16 streams: fused per-row expert passes on warp-level tensor-core tiles (csrc/rowfused.cu) where supported


_RF_MAX_C = int(_os.environ.get("AVMOE_RF_MAX_C", "0"))   # diagnostics: widen the row-fused path beyond its default C <= 192


def rf_supported(cfg, dtype) -> bool:
    """geometry check of the row-fused passes (C % 16 == 0, C <= 192, C/r <= 96, 2 <= E <= 4, T <= 96); bf16 streams only"""
    if dtype != torch.bfloat16:
        return False
    if _RF_MAX_C and cfg.C <= _RF_MAX_C:
        return bool(_lib.load().avmoe_rf_supported(min(cfg.C, 192), cfg.Cr, cfg.E, cfg.T)) and cfg.C % 16 == 0
    return bool(_lib.load().avmoe_rf_supported(cfg.C, cfg.Cr, cfg.E, cfg.T))


def rf_pack_bytes(nb: int, Nn: int, K: int, hi_lo: bool = False) -> int:
    lib = _lib.load()
    lib.avmoe_rf_pack_b_bytes.restype = C.c_int64
    return int(lib.avmoe_rf_pack_b_bytes(nb, Nn, K, int(hi_lo)))


def rf_pack_b(src: torch.Tensor, *, nb: int, bstride: int, stride_n: int, stride_k: int, Nn: int, K: int, groups: int = 1,
              group_local_k: bool = True, scale_n: Optional[torch.Tensor] = None, scale_k: Optional[torch.Tensor] = None, hi_lo: bool = False,
              Nn_pad: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 matrix -> packed tensor-core B operand B[k][n] (mma fragment order, bf16; include/avmoe_b200.h avmoe_rf_pack_b).
    Returns a uint8 tensor; ``out`` may be a slice of a larger arena."""
    _require_cuda(src)
    assert src.dtype == torch.float32
    nbytes = rf_pack_bytes(nb, max(Nn, Nn_pad), K, hi_lo)
    if out is None:
        out = torch.empty(nbytes, device=src.device, dtype=torch.uint8)
    assert out.numel() == nbytes and out.is_contiguous() and out.data_ptr() % 16 == 0
    for sc in (scale_n, scale_k):
        assert sc is None or sc.dtype == torch.float32
    _lib_call("rf_pack_b", _ptr(src), C.c_int64(bstride), C.c_int64(stride_n), C.c_int64(stride_k), nb, Nn, K, groups, int(group_local_k),
              _ptr(scale_n), 0 if scale_n is None else scale_n.stride(-1), _ptr(scale_k), 0 if scale_k is None else scale_k.stride(-1),
              int(hi_lo), Nn_pad, _ptr(out), _stream())
    return out


def rf_prep_front(cfg, packs):
    """Weight operands of the row-fused front pass: packed down projection with ln_before's weight folded in, the column sums
    of that (rounded) operand for the mean fold, c2 = Wd · ln_before.bias, and the transposed operand of the backward."""
    dev = packs[0].wd.device
    C_, Cr, G = cfg.C, cfg.Cr, cfg.G
    cg, crg = C_ // G, Cr // G
    has_ln = packs[0].ln1_w is not None
    one_f, one_b = rf_pack_bytes(1, Cr, C_), rf_pack_bytes(1, C_, Cr)
    wd_pack = torch.empty(cfg.E * one_f, device=dev, dtype=torch.uint8)
    wdt_pack = torch.empty(cfg.E * one_b, device=dev, dtype=torch.uint8)
    crp16 = (Cr + 15) // 16 * 16
    csum = torch.empty((cfg.E, crp16), device=dev, dtype=torch.float32)
    c2 = torch.zeros((cfg.E, Cr), device=dev, dtype=torch.float32)
    for e, pk in enumerate(packs):
        wd = pk.wd.contiguous()                                                   # [Cr, C/G]
        rf_pack_b(wd, nb=1, bstride=0, stride_n=cg, stride_k=1, Nn=Cr, K=C_, groups=G, group_local_k=True,
                  scale_k=pk.ln1_w.contiguous() if has_ln else None, out=wd_pack[e * one_f:(e + 1) * one_f])
        rf_pack_b(wd, nb=1, bstride=0, stride_n=1, stride_k=cg, Nn=C_, K=Cr, groups=G, group_local_k=False, out=wdt_pack[e * one_b:(e + 1) * one_b])
        if has_ln:
            c2[e] = (wd.view(G, crg, cg) * pk.ln1_b.view(G, 1, cg)).sum(-1).reshape(Cr)
    _lib_call("rf_colsum", _ptr(wd_pack), cfg.E, Cr, C_, _ptr(csum), _stream())
    return {"wd_pack": wd_pack, "wdt_pack": wdt_pack, "wd_csum": csum, "c2": c2, "has_ln": has_ln}


def rf_pack_tokens(tokp: torch.Tensor):
    """updated latent tokens tok' [BT,T,C] fp32 -> (packed hi/lo logits operand B[k=c][n=t], packed B[k=t][n=c]) per sample"""
    BT, T, Cc = tokp.shape
    tokp = tokp.contiguous()
    tp16 = (T + 15) // 16 * 16
    s2 = rf_pack_b(tokp, nb=BT, bstride=T * Cc, stride_n=Cc, stride_k=1, Nn=T, K=Cc, hi_lo=True, Nn_pad=tp16)
    xr = rf_pack_b(tokp, nb=BT, bstride=T * Cc, stride_n=1, stride_k=Cc, Nn=Cc, K=T)
    return s2, xr


USE_ROWFUSED_TOKENS = _os.environ.get("AVMOE_RF_TOKENS", "1") != "0"   # latent-token update: logits recomputed per tile instead of the S1 / dA1 GEMMs + column-softmax passes
# Below ~64 row tiles per sample the passes are prologue-bound (two tiles per warp) and the GEMM + column-softmax form wins:
# stage-3 call (Nk = 256): 80 + 70 us fused vs 33 + 28 us; stage-1 call (Nk = 4096): 124 + 93 us fused vs 313 + 203 us.
_RF_TOK_MIN_NK = int(_os.environ.get("AVMOE_RF_TOK_MIN_NK", "1024"))


def rf_tok_supported(kv: torch.Tensor, T: int) -> bool:
    """row-fused token-update passes (csrc/rf_tokens.cu): bf16 kv [BT,Nk,C], Nk % 16 == 0, C % 16 == 0, covered latent-token counts"""
    if not USE_ROWFUSED_TOKENS or kv.dtype != torch.bfloat16 or kv.dim() != 3 or kv.shape[0] > 65535 or kv.shape[1] < _RF_TOK_MIN_NK:
        return False
    return bool(_lib.load().avmoe_rf_tok_supported(int(kv.shape[1]), int(kv.shape[2]), int(T)))


def _rf_tok_args(kv, T, pack, A12, per_sample):
    _require_cuda(kv, pack, A12)
    assert kv.is_contiguous() and A12.is_contiguous() and A12.dtype == torch.bfloat16
    BT, Nk, Cc = kv.shape
    a = _lib.RfTokArgs()
    a.BT, a.Nk, a.C, a.T, a.Tp, a.per_sample = BT, Nk, Cc, T, round8(T), int(per_sample)
    assert A12.shape == (BT, Nk, 2 * a.Tp)
    a.kv, a.pack, a.A12 = kv.data_ptr(), pack.data_ptr(), A12.data_ptr()
    return a


def rf_tok_probs(kv: torch.Tensor, tok: torch.Tensor):
    """A1ᵀ = softmax over the Nk rows of kv·tokᵀ, written as bf16 hi | lo column blocks: A12 [BT,Nk,2Tp] (same layout as softmax_cols)."""
    BT, Nk, Cc = kv.shape
    T = tok.shape[0]
    Tp = round8(T)
    tok = tok.contiguous()
    pack = rf_pack_b(tok, nb=1, bstride=T * Cc, stride_n=Cc, stride_k=1, Nn=T, K=Cc, hi_lo=True, Nn_pad=Tp)
    A12 = torch.empty((BT, Nk, 2 * Tp), device=kv.device, dtype=torch.bfloat16)
    nparts = int(_lib.load().avmoe_rf_tok_parts(BT, Nk))
    part = torch.empty((BT, nparts, 2, Tp), device=kv.device, dtype=torch.float32)
    a = _rf_tok_args(kv, T, pack, A12, False)
    a.part = part.data_ptr()
    _lib_call("rf_tok_probs", C.byref(a), _stream(), _bytes=2 * _nbytes(kv) + _nbytes(A12))
    return A12


def rf_tok_ds(kv: torch.Tensor, dR: torch.Tensor, A12: torch.Tensor, Dt: torch.Tensor, T: int):
    """A12[:, :, Tp:Tp+T] = A1ᵀ ∘ (kv·dRᵀ − Dt)   (dR [BT,T,C] fp32, Dt [BT,T] fp32); in place, like softmax_cols_bwd"""
    BT, Nk, Cc = kv.shape
    Tp = round8(T)
    dR = dR.contiguous()
    pack = rf_pack_b(dR, nb=BT, bstride=T * Cc, stride_n=Cc, stride_k=1, Nn=T, K=Cc, hi_lo=True, Nn_pad=Tp)
    a = _rf_tok_args(kv, T, pack, A12, True)
    Dt = Dt.contiguous()
    a.Dt = Dt.data_ptr()
    _lib_call("rf_tok_ds", C.byref(a), _stream(), _bytes=_nbytes(kv) + _nbytes(A12))


def _rf_front_args(cfg, packs, egr, toks, prep, *, BT):
    a = _lib.RfFrontArgs()
    a.BT, a.N, a.C, a.Cr, a.Crp, a.G, a.E, a.T = BT, cfg.N, cfg.C, cfg.Cr, pad_bottleneck(cfg.Cr), cfg.G, cfg.E, cfg.T
    a.has_ln = int(prep["has_ln"])
    keep = []
    for e, pk in enumerate(packs):
        ex = a.ex[e]
        ex.pre_kind, ex.relu = pk.pre_kind, int(pk.relu)
        for name in ("gate_pre", "aux", "ln1_w", "ln1_b"):
            t = getattr(pk, name)
            if t is not None:
                t = t.contiguous()
                keep.append(t)
                setattr(ex, name, t.data_ptr())
        if egr is not None:
            for name in ("gate_pre", "aux", "ln1_w", "ln1_b"):
                t = getattr(egr[e], name)
                if t is not None:
                    assert t.is_contiguous()
                    setattr(ex, "g_" + name, t.data_ptr())
        if e in toks:
            a.tok_s2[e], a.tok_xr[e] = toks[e][0].data_ptr(), toks[e][1].data_ptr()
    _set(a, wd_pack=prep["wd_pack"], wd_csum=prep["wd_csum"], c2=prep["c2"], wdt_pack=prep["wdt_pack"])
    return a, keep


def rf_front_fwd(cfg, x, packs, toks, prep, z):
    """x (+aux, tok') -> z [BT,N,E,Cr] fp32: latent attention of each row, gate_pre residual, ln_before, grouped down projection
    (replaces S2 GEMM + softmax_rows + A2·tok' GEMM + xg_ln_pre_fwd + down GEMM; nothing but z is written)."""
    _require_cuda(x, z)
    a, keep = _rf_front_args(cfg, packs, None, toks, prep, BT=x.shape[0])
    st = torch.empty((cfg.E, x.shape[0] * x.shape[1], 2), device=x.device, dtype=torch.float32) if prep["has_ln"] else None
    _set(a, x=x, z=z, ln1_stats=st)
    _lib_call("rf_front_fwd", C.byref(a), _stream(), _bytes=_nbytes(x, z) + sum(_nbytes(pk.aux) for pk in packs))
    return st


def rf_front_bwd(cfg, x, packs, egr, toks, prep, ln1_stats, dz, dpool_x, dx, dres=None):
    """dz, x (+aux, tok') -> dx, xl [E,rows,C], per latent expert (A2, dS2) [BT,N,Tp8]; egr[e].aux receives d aux (for latent experts
    dXR), egr[e].gate_pre / ln1_w / ln1_b accumulate."""
    BT, N, Cc = x.shape
    rows = BT * N
    xl = torch.empty((cfg.E, rows, Cc), device=x.device, dtype=x.dtype)
    tp8 = round8(cfg.T)
    lat = {}
    a, keep = _rf_front_args(cfg, packs, egr, toks, prep, BT=BT)
    for e in toks:
        a2 = torch.empty((BT, N, tp8), device=x.device, dtype=x.dtype)
        ds2 = torch.empty((BT, N, tp8), device=x.device, dtype=x.dtype)
        lat[e] = (a2, ds2)
        a.a2[e], a.ds2[e] = a2.data_ptr(), ds2.data_ptr()
    _set(a, x=x, dz=dz, dpool_x=dpool_x, dx=dx, xl=xl, ln1_stats=ln1_stats, dres=dres)
    _lib_call("rf_front_bwd", C.byref(a), _stream(), _bytes=_nbytes(x, dz, dx, xl) + sum(2 * _nbytes(pk.aux) for pk in packs))
    return xl, lat


def rf_prep_up(cfg, packs, save2, bn_mode):
    """Up-projection operand of the row-fused back pass: per expert the packed B[k=j][n=c] = invstd2_c * dense(up_sampler)[c][j]
    (bn2's normalisation folded in, like xg_prep_up) and ybi [E,C] = -mean2_c * invstd2_c (zeros without BatchNorm)."""
    dev = packs[0].wu.device
    one = rf_pack_bytes(1, cfg.C, cfg.Cr)
    pack = torch.empty(cfg.E * one, device=dev, dtype=torch.uint8)
    crg = cfg.Cr // cfg.G
    for e, pk in enumerate(packs):
        wu = pk.wu.contiguous()
        rf_pack_b(wu, nb=1, bstride=0, stride_n=crg, stride_k=1, Nn=cfg.C, K=cfg.Cr, groups=cfg.G, group_local_k=True,
                  scale_n=save2[e, :, 1] if bn_mode != 0 else None, out=pack[e * one:(e + 1) * one])
    if bn_mode != 0:
        ybi = -(save2[:, :, 0] * save2[:, :, 1])
    else:
        ybi = torch.zeros((cfg.E, cfg.C), device=dev, dtype=torch.float32)
    return pack, ybi.contiguous()


def _rf_back_args(cfg, packs, egr, *, BT, bn_mode, bn2_wb):
    a = _lib.RfBackArgs()
    a.BT, a.N, a.C, a.Cr, a.Crp, a.G, a.E, a.bn_mode = BT, cfg.N, cfg.C, cfg.Cr, pad_bottleneck(cfg.Cr), cfg.G, cfg.E, bn_mode
    keep = []
    for e, pk in enumerate(packs):
        ex = a.ex[e]
        ex.pre_kind, ex.relu = pk.pre_kind, int(pk.relu)
        for name in ("ln2_w", "ln2_b", "gate"):
            t = getattr(pk, name)
            if t is not None:
                t = t.contiguous()
                keep.append(t)
                setattr(ex, name, t.data_ptr())
        if egr is not None:
            for name in ("ln2_w", "ln2_b", "gate"):
                t = getattr(egr[e], name)
                if t is not None:
                    assert t.is_contiguous()
                    setattr(ex, "g_" + name, t.data_ptr())
        if bn2_wb is not None:
            w, b = bn2_wb[e]
            w, b = w.contiguous(), b.contiguous()
            keep += [w, b]
            a.bn2_w[e], a.bn2_b[e] = w.data_ptr(), b.data_ptr()
    return a, keep


def rf_back_fwd(cfg, z, aff1, wu_pack, ybi, bn2_wb, bn_mode, packs, probs, out, res=None):
    """z -> out = (res +) sum_e probs*gate * ln_post(bn2(up(act(bn1(z)))))   in one pass (replaces xg_bn_act + up GEMM + xg_post_fwd)"""
    _require_cuda(z, out)
    a, keep = _rf_back_args(cfg, packs, None, BT=z.shape[0], bn_mode=bn_mode, bn2_wb=bn2_wb)
    _set(a, z=z, aff1=aff1, wu_pack=wu_pack, ybi=ybi, probs=probs, res=res, out=out)
    _lib_call("rf_back_fwd", C.byref(a), _stream(), _bytes=_nbytes(z, out, res))


def rf_back_bwd(cfg, dout, z, aff1, wu_pack, ybi, bn2_wb, bn_mode, packs, probs, dprobs, red2, egr):
    """dout, z -> yd [E,rows,2C] = [yhat | dy2], a [E,rows,Crp]; d ln_post / d gate / d probs / bn2 reductions accumulated
    (replaces xg_post_bwd and the yd / a tensors the forward used to save)."""
    BT, N, Cc = dout.shape
    rows = BT * N
    yd = torch.empty((cfg.E, rows, 2 * Cc), device=dout.device, dtype=dout.dtype)
    act = torch.empty((cfg.E, rows, pad_bottleneck(cfg.Cr)), device=dout.device, dtype=dout.dtype)
    a, keep = _rf_back_args(cfg, packs, egr, BT=BT, bn_mode=bn_mode, bn2_wb=bn2_wb)
    _set(a, z=z, aff1=aff1, wu_pack=wu_pack, ybi=ybi, probs=probs, dout=dout, yd=yd, act=act, dprobs=dprobs, red2=red2)
    _lib_call("rf_back_bwd", C.byref(a), _stream(), _bytes=_nbytes(dout, z, yd, act))
    return yd, act


# ------------------------------------------------------------------------------------------------ variant pre-transforms
MHA_HEADS = 4
_NXN_CHUNK_BYTES = 1 << 30   # fp32 scratch budget for the [chunk,N,N] score matrix of the AVVP token self-attention


def _nxn_chunk(BT: int, N: int) -> int:
    return max(1, min(BT, _NXN_CHUNK_BYTES // (4 * N * N)))


def selfattn_nxn_fwd(x):
    """AVVP single-modal pre-transform: aux_j = sum_i softmax_j(x_i·x_·)[j] x_i.  Returns (aux, row log-sum-exps)."""
    _require_cuda(x)
    BT, N, Cc = x.shape
    aux = torch.empty_like(x)
    lse = torch.empty((BT, N), device=x.device, dtype=torch.float32)
    bc = _nxn_chunk(BT, N)
    for b0 in range(0, BT, bc):
        xc = x[b0:b0 + bc]
        nb = xc.shape[0]
        S = bgemm(xc, xc, out_dtype=torch.float32)                                  # [nb,N,N]  x_i . x_j
        P = torch.empty((nb, N, N), device=x.device, dtype=x.dtype)
        _lib_call("nxn_softmax_rows", _ptr(S), _ptr(P), _ptr(lse[b0:b0 + bc]), C.c_int64(nb * N), N, _dt(P), _stream())
        aux[b0:b0 + bc] = bgemm(P, xc, a_kmajor=False, b_kmajor=False)             # sum_i P[i,j] x_i
    return aux, lse


def selfattn_nxn_bwd(x, lse, daux, dx):
    BT, N, Cc = x.shape
    bc = _nxn_chunk(BT, N)
    for b0 in range(0, BT, bc):
        xc, gc = x[b0:b0 + bc], daux[b0:b0 + bc]
        nb = xc.shape[0]
        S = bgemm(xc, xc, out_dtype=torch.float32)
        dAtt = bgemm(xc, gc, out_dtype=torch.float32)                               # datt[i,j] = x_i . g_j
        att = torch.empty((nb, N, N), device=x.device, dtype=x.dtype)
        dS = torch.empty_like(att)
        _lib_call("nxn_bwd_rows", _ptr(S), _ptr(dAtt), _ptr(lse[b0:b0 + bc]), _ptr(att), _ptr(dS), C.c_int64(nb * N), N, _dt(att), _stream())
        t1 = bgemm(att, gc, b_kmajor=False)                                         # sum_j att[i,j] g_j
        t2 = bgemm(dS, xc, b_kmajor=False)                                          # sum_j ds[i,j] x_j
        t3 = bgemm(dS, xc, a_kmajor=False, b_kmajor=False)                          # sum_i ds[i,j] x_i
        dxc = dx[b0:b0 + bc]
        _lib_call("add_into", _ptr(dxc), _ptr(t1), _ptr(t2), _ptr(t3), C.c_int64(dxc.numel()), _dt(dxc), _stream())


def mha_fwd(x, mp, keep=None, keep_scale: float = 1.0):
    """AVS-v1 single-modal pre-transform: nn.MultiheadAttention(C, 4) over the BT axis of x [BT,N,C].
    keep [N,heads,BT,BT] uint8 (optional): attention dropout mask; kept probabilities are scaled by keep_scale = 1/(1-p)."""
    _require_cuda(x)
    w_in, b_in, w_out, b_out = mp
    L, NB, Cc = x.shape
    x2 = x.reshape(L * NB, Cc)
    qkv = gemm(x2, prep_weight(w_in, x.dtype, False), out_dtype=torch.float32, col_bias=b_in.contiguous())      # [L*NB,3C] fp32
    ctx = torch.empty((L, NB, Cc), device=x.device, dtype=x.dtype)
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.shape == (NB, MHA_HEADS, L, L) and keep.is_contiguous()
    _lib_call("mha_attn_fwd", _ptr(qkv), _ptr(ctx), L, NB, Cc, MHA_HEADS, _dt(x), _ptr(keep), C.c_float(keep_scale), _stream())
    out = gemm(ctx.reshape(L * NB, Cc), prep_weight(w_out, x.dtype, False), col_bias=b_out.contiguous()).reshape(L, NB, Cc)
    return out, (qkv, ctx, keep, keep_scale)


def mha_bwd(x, mp, saved, daux, dx):
    w_in, b_in, w_out, b_out = mp
    qkv, ctx, keep, keep_scale = saved
    L, NB, Cc = x.shape
    rows = L * NB
    g2 = daux.reshape(rows, Cc)
    f32 = torch.float32
    dw_out = gemm(g2, ctx.reshape(rows, Cc), a_kmajor=False, b_kmajor=False, out_dtype=f32, split_k=split_k_hint(rows, Cc, Cc))   # gᵀ·ctx
    _, db_out, _ = reduce_bias(daux)
    dctx = gemm(g2, prep_weight(w_out, x.dtype, True), out_dtype=f32)                                                            # g·W_out   fp32
    dqkv = torch.empty((L, NB, 3 * Cc), device=x.device, dtype=x.dtype)
    _lib_call("mha_attn_bwd", _ptr(qkv), _ptr(dctx), _ptr(dqkv), L, NB, Cc, MHA_HEADS, _dt(x), _ptr(keep), C.c_float(keep_scale), _stream())
    dq2 = dqkv.reshape(rows, 3 * Cc)
    dw_in = gemm(dq2, x.reshape(rows, Cc), a_kmajor=False, b_kmajor=False, out_dtype=f32, split_k=split_k_hint(rows, 3 * Cc, Cc))  # dqkvᵀ·x
    _, db_in, _ = reduce_bias(dqkv)
    dxa = gemm(dq2, prep_weight(w_in, x.dtype, True))                                                                            # dqkv·W_in
    _lib_call("add_into", _ptr(dx), _ptr(dxa), None, None, C.c_int64(dx.numel()), _dt(dx), _stream())
    return [dw_in, db_in, dw_out, db_out]

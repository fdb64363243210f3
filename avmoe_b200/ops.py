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
    a.split_k, a.accumulate, a.impl = split_k, int(accumulate), impl
    _lib.check(_lib.load().avmoe_gemm(C.byref(a), _stream()), "gemm")
    return out

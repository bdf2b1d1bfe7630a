"""Thin torch-tensor wrappers over the C ABI (device pointers + the current CUDA stream).

torch is used here for device memory and streams only; every computation is a kernel in
libspacer_b200.so.  Each wrapper validates dtype/contiguity and raises SpacerError on failure.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_DLOGITS, EPI_F32T, EPI_GELU, EPI_LMHEAD, EPI_QUICKGELU, EPI_STORE,
                   EPI_SWIGLU, GemmArgs, SpacerError, check)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise SpacerError(f"{name} must be a CUDA tensor (spacer_b200 has no CPU path)")
    if t.dtype != dtype:
        raise SpacerError(f"{name} must be {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise SpacerError(f"{name} must have a contiguous last dimension")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn=False, b_mn=False, out=None, epilogue=EPI_STORE,
         bias=None, residual=None, aux=None, k_splits=1, bn=0, targets=None, lse_part=None,
         tgt_logit=None, lse=None, coef=None, M=None, N=None, K=None):
    """D[M,N] = epi(A * B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn)."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    kb = b.shape[0] if b_mn else b.shape[1]
    if kb != K:
        raise SpacerError(f"gemm: K mismatch {K} vs {kb}")
    if out is None:
        if epilogue == EPI_F32T:
            splits = lib.sb_gemm_effective_splits(K, k_splits)
            out = torch.empty((splits, N, M), device=a.device, dtype=torch.float32)
        elif epilogue == EPI_SWIGLU:
            out = torch.empty((M, N // 2), device=a.device, dtype=torch.bfloat16)
        elif epilogue == EPI_LMHEAD:
            out = torch.empty((1,), device=a.device, dtype=torch.bfloat16)  # unused
        else:
            out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    g.B, g.ldb, g.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    g.D = out.data_ptr()
    g.ldd = out.stride(-2) if out.dim() >= 2 else 0
    g.epilogue, g.k_splits, g.bn = epilogue, k_splits, bn
    g.bias = None if bias is None else bias.data_ptr()
    if residual is not None:
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    if aux is not None:
        g.aux, g.ldaux = aux.data_ptr(), aux.stride(0)
    for name, t in (("targets", targets), ("lse_part", lse_part), ("tgt_logit", tgt_logit),
                    ("lse", lse), ("coef", coef)):
        if t is not None:
            setattr(g, name, t.data_ptr())
    check(lib.sb_gemm(C.byref(g), _stream()), "sb_gemm")
    return out

"""Thin torch-tensor wrappers over the C ABI (device pointers + the current CUDA stream).

torch is used here for device memory and streams only; every computation is a kernel in
libspacer_b200.so.  Each wrapper validates dtype/contiguity and raises SpacerError on failure.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_DLOGITS, EPI_F32T, EPI_F32T_SWIGLU, EPI_GELU, EPI_LMHEAD, EPI_QUICKGELU, EPI_STORE, EPI_SWIGLU,
                   GemmArgs, SampleArgs, SpacerError, check)


# ------------------------------------------------------------------------------------------------
# kernel-launch accounting (bench.py's `gpu_launches`): the library counts every launch it enqueues;
# CUDA-graph replays re-launch the captured kernels, so model.py adds (graph nodes x replays) here.
# ------------------------------------------------------------------------------------------------
_graph_replayed = 0


def reset_launch_count() -> None:
    global _graph_replayed
    _lib.load().sb_launch_counter(None, 1)
    _graph_replayed = 0


def direct_launch_count() -> int:
    n = C.c_longlong(0)
    _lib.load().sb_launch_counter(C.byref(n), 0)
    return int(n.value)


def note_graph_replay(kernels: int) -> None:
    global _graph_replayed
    _graph_replayed += int(kernels)


def launch_count() -> int:
    return direct_launch_count() + _graph_replayed


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


# ------------------------------------------------------------------------------------------------
# generic call: tensors -> device pointers, the current stream appended
# ------------------------------------------------------------------------------------------------
def call(name: str, *args):
    lib = _lib.load()
    fn = getattr(lib, name)
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise SpacerError(f"{name}: CPU tensor passed to a device entry point")
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    check(fn(*conv, torch.cuda.current_stream().cuda_stream), name)


def sample(logits, step, out_tokens, *, V=None, R=None, greedy=False, top_p=1.0, top_k=0, temperature=1.0,
           repetition_penalty=1.0, seen=None, seed=0, seed_dev=None, finished=None, out_ids=None, out_logprob=None,
           eos_ids=(), pad_id=0, suppress_eos=False):
    """sb_sample: one token per row of fp32 `logits` [R, ld] (HF's processor chain repetition penalty -> temperature ->
    top-k -> top-p -> multinomial, or greedy argmax)."""
    _req(logits, torch.float32, "logits")
    a = SampleArgs()
    a.logits, a.ld = logits.data_ptr(), logits.stride(0)
    a.R = logits.shape[0] if R is None else R
    a.V = logits.shape[1] if V is None else V
    a.mode = 1 if greedy else 0
    a.top_p, a.top_k, a.temperature, a.repetition_penalty = float(top_p), int(top_k), float(temperature), float(repetition_penalty)
    if seen is not None:
        _req(seen, torch.int32, "seen")
        a.seen, a.seen_ld = seen.data_ptr(), seen.stride(0)
    a.seed = int(seed)
    a.seed_dev = None if seed_dev is None else seed_dev.data_ptr()
    a.step_ptr, a.out_tokens = step.data_ptr(), out_tokens.data_ptr()
    a.finished = None if finished is None else finished.data_ptr()
    if out_ids is not None:
        a.out_ids, a.out_ld = out_ids.data_ptr(), out_ids.stride(0)
    a.out_logprob = None if out_logprob is None else out_logprob.data_ptr()
    if len(eos_ids) > 4:
        raise SpacerError("sample: at most 4 eos ids")
    for k, e in enumerate(eos_ids):
        a.eos_ids[k] = int(e)
    a.n_eos, a.pad_id, a.suppress_eos = len(eos_ids), int(pad_id), int(bool(suppress_eos))
    check(_lib.load().sb_sample(C.byref(a), _stream()), "sb_sample")


def cast_f32_bf16(src, dst=None):
    _req(src, torch.float32, "src")
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    call("sb_cast_f32_bf16", src, dst, src.numel())
    return dst


def layernorm_fwd(x, w, b, eps=1e-6, save_stats=False, out=None):
    T, E = x.shape
    y = out if out is not None else torch.empty_like(x)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(T, device=x.device, dtype=torch.float32)
        rstd = torch.empty(T, device=x.device, dtype=torch.float32)
    call("sb_layernorm_fwd", x, w, b, y, mean, rstd, T, E, eps)
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(x, w, mean, rstd, dy, dw, db, dres=None, out=None):
    T, E = x.shape
    dx = out if out is not None else torch.empty_like(x)
    call("sb_layernorm_bwd", x, w, mean, rstd, dy, dres, dx, dw, db, T, E)
    return dx


def rmsnorm_fwd(x, w, eps=1e-6, save_stats=False, out=None):
    T, H = x.shape
    y = out if out is not None else torch.empty_like(x)
    rstd = torch.empty(T, device=x.device, dtype=torch.float32) if save_stats else None
    call("sb_rmsnorm_fwd", x, w, y, rstd, T, H, eps)
    return (y, rstd) if save_stats else y


def rmsnorm_bwd(x, w, rstd, dy, dw, dres=None, out=None):
    T, H = x.shape
    dx = out if out is not None else torch.empty_like(x)
    call("sb_rmsnorm_bwd", x, w, rstd, dy, dres, dx, dw, T, H)
    return dx


def rope_vit(qkv, heads, head_dim, grids_dev, merge=2, inverse=False):
    call("sb_rope_vit", qkv, qkv.shape[0], heads, head_dim, grids_dev, grids_dev.shape[0], merge, int(inverse))


def mrope(qkv, pos, n_heads, n_kv_heads, head_dim, theta, sections, inverse=False, k_out=None, v_out=None, kv_ld=0):
    _req(pos, torch.int32, "pos")
    call("sb_mrope", qkv, pos, qkv.shape[0], n_heads, n_kv_heads, head_dim, float(theta), sections[0], sections[1],
         int(inverse), k_out, v_out, kv_ld)


def attn_args(q, k, v, o, meta, n_heads, n_kv_heads, head_dim, scale, lse=None, T=None, Tk=0):
    a = _lib.AttnArgs()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
    a.o, a.ldo = o.data_ptr(), o.stride(0)
    a.lse = None if lse is None else lse.data_ptr()
    a.meta = meta.data_ptr()
    a.T = q.shape[0] if T is None else T
    a.Tk = Tk
    a.n_heads, a.n_kv_heads, a.head_dim, a.scale = n_heads, n_kv_heads, head_dim, float(scale)
    return a


def attn_fwd(q, k, v, meta, n_heads, n_kv_heads, head_dim, scale=None, out=None, save_lse=False, Tk=0):
    """q/k/v are column views into a fused projection buffer (row stride = buffer width).  Tk > 0: the key/value views
    hold Tk rows while q (and meta, whose key indices are absolute) cover only the last q.shape[0] of them."""
    T = q.shape[0]
    _req(meta, torch.int32, "meta")
    scale = head_dim ** -0.5 if scale is None else scale
    o = out if out is not None else torch.empty((T, n_heads * head_dim), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((n_heads, T), device=q.device, dtype=torch.float32) if save_lse else None
    a = attn_args(q, k, v, o, meta, n_heads, n_kv_heads, head_dim, scale, lse, Tk=Tk)
    check(_lib.load().sb_attn_fwd(C.byref(a), _stream()), "sb_attn_fwd")
    return (o, lse) if save_lse else o


_attn_ws = {}


def attn_bwd_workspace(T, n_heads, n_kv_heads, head_dim, device):
    """(gqa scratch bf16, tile scratch int32) for sb_attn_bwd, cached per shape (the kernels are stream-ordered)."""
    key = (T, n_heads, n_kv_heads, head_dim, str(device))
    ws = _attn_ws.get(key)
    if ws is None:
        a, b = C.c_longlong(0), C.c_longlong(0)
        check(_lib.load().sb_attn_bwd_workspace(T, T, n_heads, n_kv_heads, head_dim, C.byref(a), C.byref(b)),
              "sb_attn_bwd_workspace")
        gqa = torch.empty(a.value, device=device, dtype=torch.bfloat16) if a.value else None
        ws = (gqa, torch.empty(b.value, device=device, dtype=torch.int32))
        _attn_ws.clear()          # one shape at a time: the scratch is as large as two activation tensors
        _attn_ws[key] = ws
    return ws


def attn_bwd(q, k, v, o, lse, d_o, meta, n_heads, n_kv_heads, head_dim, dq_out, dk_out, dv_out, scale=None,
             delta=None):
    """dq_out/dk_out/dv_out: bf16 column views (e.g. into a d_qkv buffer)."""
    T = q.shape[0]
    scale = head_dim ** -0.5 if scale is None else scale
    if delta is None:
        delta = torch.empty((n_heads, T), device=q.device, dtype=torch.float32)
    gqa_ws, tile_ws = attn_bwd_workspace(T, n_heads, n_kv_heads, head_dim, q.device)
    a = attn_args(q, k, v, o, meta, n_heads, n_kv_heads, head_dim, scale, lse)
    a.d_o, a.lddo = d_o.data_ptr(), d_o.stride(0)
    a.delta = delta.data_ptr()
    a.dq, a.lddq = dq_out.data_ptr(), dq_out.stride(0)
    a.dk, a.dv, a.lddk, a.lddv = dk_out.data_ptr(), dv_out.data_ptr(), dk_out.stride(0), dv_out.stride(0)
    a.gqa_ws = None if gqa_ws is None else gqa_ws.data_ptr()
    a.tile_ws = tile_ws.data_ptr()
    check(_lib.load().sb_attn_bwd(C.byref(a), _stream()), "sb_attn_bwd")


def grpo_loss_workspace(G, C_len, device):
    """Zeroed fp32 scratch for sb_grpo_loss (ticket counter + per-CTA partial sums)."""
    n = C.c_longlong(0)
    check(_lib.load().sb_grpo_loss_workspace(G, C_len, C.byref(n)), "sb_grpo_loss_workspace")
    return torch.zeros(n.value, device=device, dtype=torch.float32)


def make_meta(prefix_len, seg_start, kv_end, device="cuda"):
    """int32 [T,4] visibility metadata from three equal-length integer sequences/tensors."""
    m = torch.stack([torch.as_tensor(prefix_len), torch.as_tensor(seg_start), torch.as_tensor(kv_end),
                     torch.zeros_like(torch.as_tensor(kv_end))], dim=1).to(torch.int32)
    return m.contiguous().to(device)

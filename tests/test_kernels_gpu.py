"""GPU parity of the bandwidth kernels and attention (through the C ABI) against plain torch fp32 math."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def close(a, b, tol=2e-2, name=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    den = b.abs().max().item() + 1e-6
    assert err / den < tol, f"{name}: max abs err {err} vs scale {den}"


def test_cast():
    from spacer_b200 import ops
    x = rnd((1000, 1176), 1, dtype=torch.float32)
    assert torch.equal(ops.cast_f32_bf16(x), x.bfloat16())
    y = rnd((13,), 2, dtype=torch.float32)
    assert torch.equal(ops.cast_f32_bf16(y), y.bfloat16())


@pytest.mark.parametrize("T,E", [(300, 1280), (64, 160), (1000, 3584)])
def test_layernorm_fwd_bwd(T, E):
    from spacer_b200 import ops
    x, w, b, dy, dres = rnd((T, E), 1), 1 + rnd((E,), 2, 0.1), rnd((E,), 3, 0.1), rnd((T, E), 4), rnd((T, E), 5)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, save_stats=True)
    xf = x.float().requires_grad_()
    wf, bf = w.float().requires_grad_(), b.float().requires_grad_()
    ref = F.layer_norm(xf, (E,), wf, bf, 1e-6)
    close(y, ref, 1e-2, "ln fwd")
    ref.backward(dy.float())
    dw = torch.zeros(E, device="cuda")
    db = torch.zeros(E, device="cuda")
    dx = ops.layernorm_bwd(x, w, mean, rstd, dy, dw, db, dres=dres)
    close(dx, xf.grad + dres.float(), 1e-2, "ln dx")
    close(dw, wf.grad, 1e-2, "ln dw")
    close(db, bf.grad, 1e-2, "ln db")


@pytest.mark.parametrize("T,H", [(300, 3584), (77, 256), (512, 1536)])
def test_rmsnorm_fwd_bwd(T, H):
    from spacer_b200 import ops
    x, w, dy = rnd((T, H), 1), 1 + rnd((H,), 2, 0.1), rnd((T, H), 4)
    y, rstd = ops.rmsnorm_fwd(x, w, save_stats=True)
    xf, wf = x.float().requires_grad_(), w.float().requires_grad_()
    ref = wf * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    close(y, ref, 1e-2, "rms fwd")
    ref.backward(dy.float())
    dw = torch.zeros(H, device="cuda")
    dx = ops.rmsnorm_bwd(x, w, rstd, dy, dw)
    close(dx, xf.grad, 1e-2, "rms dx")
    close(dw, wf.grad, 1e-2, "rms dw")


def _vit_cos_sin(grid, hd, merge=2):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import qwen2vl_ref as R
    d = R.Dims(v_embed=hd, v_heads=1, merge=merge)
    return R.vit_cos_sin(torch.tensor(grid), d)


def test_rope_vit():
    from spacer_b200 import ops
    grid = [[2, 4, 6], [1, 8, 4]]
    heads, hd = 3, 80
    T = sum(t * h * w for t, h, w in grid)
    qkv = rnd((T, 3 * heads * hd), 7)
    cos, sin = _vit_cos_sin(grid, hd)
    cos, sin = cos.cuda()[:, None, :], sin.cuda()[:, None, :]
    q, k, v = qkv.float().view(T, 3, heads, hd).unbind(1)
    rot = lambda x: torch.cat((-x[..., hd // 2:], x[..., :hd // 2]), -1)
    qr, kr = q * cos + rot(q) * sin, k * cos + rot(k) * sin
    ref = torch.stack([qr, kr, v], 1).reshape(T, -1)
    out = qkv.clone()
    g = torch.tensor(grid, dtype=torch.int32, device="cuda")
    ops.rope_vit(out, heads, hd, g)
    close(out, ref, 1e-2, "rope_vit")
    ops.rope_vit(out, heads, hd, g, inverse=True)
    close(out, qkv, 2e-2, "rope_vit inverse")


def test_mrope_and_kv_write():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import qwen2vl_ref as R
    from spacer_b200 import ops
    T, nh, nkv, hd = 200, 4, 2, 128
    d = R.Dims(heads=nh, kv_heads=nkv)
    pos = torch.stack([torch.randint(0, 3000, (T,)) for _ in range(3)])
    cos, sin = R.mrope_cos_sin(pos[:, None, :], d)
    cos, sin = cos[0].cuda().bfloat16().float()[:, None, :], sin[0].cuda().bfloat16().float()[:, None, :]
    qkv = rnd((T, (nh + 2 * nkv) * hd), 3)
    q = qkv[:, :nh * hd].float().view(T, nh, hd)
    k = qkv[:, nh * hd:(nh + nkv) * hd].float().view(T, nkv, hd)
    rot = lambda x: torch.cat((-x[..., hd // 2:], x[..., :hd // 2]), -1)
    qr, kr = q * cos + rot(q) * sin, k * cos + rot(k) * sin
    out = qkv.clone()
    kc = torch.zeros((T, nkv * hd), device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros((T, nkv * hd), device="cuda", dtype=torch.bfloat16)
    ops.mrope(out, pos.to(torch.int32).cuda(), nh, nkv, hd, 1e6, (16, 24), k_out=kc, v_out=vc, kv_ld=nkv * hd)
    close(out[:, :nh * hd], qr.reshape(T, -1), 1e-2, "mrope q")
    close(out[:, nh * hd:(nh + nkv) * hd], kr.reshape(T, -1), 1e-2, "mrope k")
    assert torch.equal(out[:, (nh + nkv) * hd:], qkv[:, (nh + nkv) * hd:])
    assert torch.equal(kc, out[:, nh * hd:(nh + nkv) * hd]) and torch.equal(vc, qkv[:, (nh + nkv) * hd:])
    ops.mrope(out, pos.to(torch.int32).cuda(), nh, nkv, hd, 1e6, (16, 24), inverse=True)
    close(out, qkv, 3e-2, "mrope inverse")


@pytest.mark.parametrize("mode", [0, 1])
def test_act_fwd_bwd(mode):
    from spacer_b200 import ops
    z, dy = rnd((64, 640), 1, 2.0), rnd((64, 640), 2)
    zf = z.float().requires_grad_()
    ref = zf * torch.sigmoid(1.702 * zf) if mode == 0 else F.gelu(zf)
    f = torch.empty_like(z)
    ops.call("sb_act_fwd", z, f, z.numel(), mode)
    close(f, ref, 1e-2)
    ref.backward(dy.float())
    dz = torch.empty_like(z)
    ops.call("sb_act_bwd", z, dy, dz, z.numel(), mode)
    close(dz, zf.grad, 1e-2)


def test_swiglu_bwd():
    from spacer_b200 import ops
    T, I = 100, 256
    g, u, dact = rnd((T, I), 1), rnd((T, I), 2), rnd((T, I), 3)
    gu = torch.stack([g.view(T, I // 64, 64), u.view(T, I // 64, 64)], 2).reshape(T, 2 * I).contiguous()
    gf, uf = g.float().requires_grad_(), u.float().requires_grad_()
    a = F.silu(gf) * uf
    a.backward(dact.float())
    dgu = torch.empty_like(gu)
    act = torch.empty((T, I), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_swiglu_bwd", gu, dact, dgu, act, T, I)
    close(act, a, 2e-2, "act")
    d = dgu.view(T, I // 64, 2, 64)
    close(d[:, :, 0].reshape(T, I), gf.grad, 2e-2, "dg")
    close(d[:, :, 1].reshape(T, I), uf.grad, 2e-2, "du")


def test_embed_merge_and_bwd():
    from spacer_b200 import ops
    V, H, T, vid = 500, 256, 300, 499
    ids = torch.randint(0, 400, (T,), dtype=torch.int32)
    ids[20:60] = vid
    ids[100:110] = vid
    ids = ids.cuda()
    emb, vis = rnd((V, H), 1), rnd((50, H), 2)
    vi = torch.empty(T, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.call("sb_vision_index", ids, vi, T, vid, -7, cnt)
    assert cnt.item() == 50
    out = torch.empty((T, H), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_embed_merge", ids, vi, emb, vis, out, T, H, 50)
    ref = emb[ids.long()].clone()
    ref[ids == vid] = vis
    assert torch.equal(out, ref)
    dx = rnd((T, H), 3)
    de = torch.zeros((V, H), device="cuda", dtype=torch.bfloat16)
    dv = torch.zeros((50, H), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_embed_bwd", ids, vi, dx, de, dv, T, H, 50)
    assert torch.equal(dv, dx[ids == vid])
    ref_de = torch.zeros((V, H), device="cuda").index_add_(0, ids.long()[ids != vid], dx.float()[ids != vid])
    close(de, ref_de, 2e-2)


def test_gather_scatter_colsum():
    from spacer_b200 import ops
    src = rnd((100, 256), 1)
    rows = torch.tensor([5, 5, 5, 7, 99, 0], dtype=torch.int32, device="cuda")
    dst = torch.empty((6, 256), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_gather_rows", src, rows, dst, 6, 256)
    assert torch.equal(dst, src[rows.long()])
    acc = torch.zeros((100, 256), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_scatter_add_rows", dst, rows, acc, 6, 256)
    ref = torch.zeros((100, 256), device="cuda").index_add_(0, rows.long(), dst.float())
    close(acc, ref, 2e-2)
    dy = rnd((777, 1288), 2)
    out = torch.zeros(1288, device="cuda")
    ops.call("sb_colsum", dy, out, 777, 1288, 1288)
    close(out, dy.float().sum(0), 1e-3)


# ------------------------------------------------------------------------------------------------
def ref_attention(q, k, v, meta, nh, nkv, hd):
    """fp32 reference with the (prefix_len, seg_start, kv_end) visibility rule."""
    T = q.shape[0]
    qh = q.float().view(T, nh, hd).transpose(0, 1)
    kh = k.float().view(T, nkv, hd).transpose(0, 1).repeat_interleave(nh // nkv, 0)
    vh = v.float().view(T, nkv, hd).transpose(0, 1).repeat_interleave(nh // nkv, 0)
    j = torch.arange(T, device=q.device)[None, :]
    m = meta.long()
    vis = (j < m[:, 0:1]) | ((j >= m[:, 1:2]) & (j < m[:, 2:3]))
    s = (qh @ kh.transpose(1, 2)) * hd ** -0.5
    s = s.masked_fill(~vis[None], float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)
    return (p @ vh).transpose(0, 1).reshape(T, nh * hd)


def metas(kind, T):
    from spacer_b200 import ops
    t = torch.arange(T)
    if kind == "causal":
        return ops.make_meta(torch.zeros(T, dtype=torch.long), torch.zeros(T, dtype=torch.long), t + 1)
    if kind == "slabs":  # block diagonal, slab 96
        s = (t // 96) * 96
        return ops.make_meta(torch.zeros(T, dtype=torch.long), s, torch.clamp(s + 96, max=T))
    if kind == "prefix":  # prompt P, then completions of length C sharing the prompt
        P, C = 100, 50
        seg = torch.where(t < P, torch.zeros_like(t), P + ((t - P) // C) * C)
        pre = torch.where(t < P, torch.zeros_like(t), torch.full_like(t, P))
        return ops.make_meta(pre, seg, t + 1)
    if kind == "prefix_big":  # several 64-row tiles per segment, ragged ends
        P, C = 700, 130
        seg = torch.where(t < P, torch.zeros_like(t), P + ((t - P) // C) * C)
        pre = torch.where(t < P, torch.zeros_like(t), torch.full_like(t, P))
        return ops.make_meta(pre, seg, t + 1)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,T,nh,nkv,hd", [("causal", 300, 4, 2, 128), ("slabs", 288, 2, 2, 80),
                                               ("prefix_big", 1220, 7, 1, 128), ("slabs", 960, 3, 3, 80),
                                               ("prefix", 250, 4, 1, 128), ("causal", 64, 2, 2, 80),
                                               ("prefix", 333, 14, 2, 128)])
@pytest.mark.parametrize("impl,bwd_impl", [(0, 0), (1, 3), (0, 1), (0, 2)])
def test_attention_fwd_bwd(kind, T, nh, nkv, hd, impl, bwd_impl):
    """impl 0 = tcgen05/TMEM/TMA forward (default), 1 = mma.sync forward; bwd_impl 0 = tcgen05 backward (default),
    bit 0 = dQ on mma.sync, bit 1 = dK/dV on mma.sync (each tcgen05 kernel is also checked on its own)."""
    from spacer_b200 import ops
    lib = ops._lib.load()
    assert lib.sb_set_attn_impl(impl) == 0
    assert lib.sb_set_attn_bwd_impl(bwd_impl) == 0
    try:
        _attention_case(kind, T, nh, nkv, hd)
    finally:
        lib.sb_set_attn_impl(0)
        lib.sb_set_attn_bwd_impl(0)


def _attention_case(kind, T, nh, nkv, hd):
    from spacer_b200 import ops
    W = (nh + 2 * nkv) * hd
    qkv = rnd((T, W), 11, 0.7)
    q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:(nh + nkv) * hd], qkv[:, (nh + nkv) * hd:]
    meta = metas(kind, T)
    o, lse = ops.attn_fwd(q, k, v, meta, nh, nkv, hd, save_lse=True)
    qf, kf, vf = (x.float().clone().requires_grad_() for x in (q, k, v))
    ref = ref_attention(qf, kf, vf, meta, nh, nkv, hd)
    close(o, ref, 2e-2, "attn fwd")
    d_o = rnd((T, nh * hd), 12)
    ref.backward(d_o.float())
    dqkv = torch.zeros_like(qkv)
    ops.attn_bwd(q, k, v, o, lse, d_o, meta, nh, nkv, hd, dqkv[:, :nh * hd], dqkv[:, nh * hd:(nh + nkv) * hd],
                 dqkv[:, (nh + nkv) * hd:])
    close(dqkv[:, (nh + nkv) * hd:], vf.grad, 3e-2, "dv")
    close(dqkv[:, nh * hd:(nh + nkv) * hd], kf.grad, 3e-2, "dk")
    close(dqkv[:, :nh * hd], qf.grad, 3e-2, "dq")

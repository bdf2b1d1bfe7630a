"""GPU parity of the tcgen05 GEMM (through the C ABI) against a torch fp32 matmul of the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(out, ref, tol=2e-2):
    out = out.float()
    err = (out - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"max abs err {err} vs ref scale {den}"


SHAPES = [(128, 256, 64), (128, 256, 512), (200, 264, 1176), (1024, 1280, 1280), (2304, 4608, 3584),
          (64, 64, 128), (300, 128, 200)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_store(M, N, K, a_mn, b_mn):
    from spacer_b200 import ops
    if a_mn and M % 8:
        pytest.skip("MN-major A needs M%8==0")
    A = _mk((M, K), 1, 0.5)
    B = _mk((N, K), 2, 0.5)
    ref = A.float() @ B.float().t()
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    _close(out, ref)


def test_gemm_bias_residual():
    from spacer_b200 import ops
    M, N, K = 520, 768, 320
    A, B = _mk((M, K), 3, 0.5), _mk((N, K), 4, 0.5)
    bias, res = _mk((N,), 5), _mk((M, N), 6)
    ref = (A.float() @ B.float().t() + bias.float()).bfloat16().float() + res.float()
    out = ops.gemm(A, B, bias=bias, residual=res)
    _close(out, ref)
    # accumulate in place (residual aliases D)
    acc = res.clone()
    ops.gemm(A, B, out=acc, residual=acc)
    _close(acc, (A.float() @ B.float().t()).bfloat16().float() + res.float())


@pytest.mark.parametrize("epi", ["quickgelu", "gelu"])
def test_gemm_act(epi):
    from spacer_b200 import ops
    M, N, K = 384, 512, 256
    A, B, bias = _mk((M, K), 7, 0.5), _mk((N, K), 8, 0.2), _mk((N,), 9)
    z = A.float() @ B.float().t() + bias.float()
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    if epi == "quickgelu":
        zr = z.bfloat16().float()
        ref = zr * torch.sigmoid(1.702 * zr)
        out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_QUICKGELU, aux=aux)
    else:
        zr = z.bfloat16().float()
        ref = torch.nn.functional.gelu(zr)
        out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, aux=aux)
    _close(aux, z)
    _close(out, ref)
    out2 = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_QUICKGELU if epi == "quickgelu" else ops.EPI_GELU)
    _close(out2, ref)


def test_gemm_swiglu():
    from spacer_b200 import ops
    M, I, K = 260, 512, 384
    A = _mk((M, K), 10, 0.5)
    Wg, Wu = _mk((I, K), 11, 0.1), _mk((I, K), 12, 0.1)
    # interleave [64 gate | 64 up]
    W = torch.stack([Wg.view(I // 64, 64, K), Wu.view(I // 64, 64, K)], dim=1).reshape(2 * I, K).contiguous()
    g = (A.float() @ Wg.float().t()).bfloat16().float()
    u = (A.float() @ Wu.float().t()).bfloat16().float()
    ref = torch.nn.functional.silu(g).bfloat16().float() * u
    aux = torch.empty((M, 2 * I), device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(A, W, epilogue=ops.EPI_SWIGLU, aux=aux)
    _close(out, ref)
    raw = A.float() @ W.float().t()
    _close(aux, raw)


@pytest.mark.parametrize("G,splits", [(8, 1), (8, 4), (12, 7), (16, 3), (24, 2)])
def test_gemm_f32t_splitk(G, splits):
    from spacer_b200 import ops
    Nw, K = 4608, 3584
    W = _mk((Nw, K), 13, 0.05)
    x = _mk((G, K), 14, 1.0)
    parts = ops.gemm(W, x, epilogue=ops.EPI_F32T, k_splits=splits)
    y = parts.sum(0)  # [G, Nw]
    ref = x.float() @ W.float().t()
    _close(y, ref, 5e-3)


def test_gemm_lmhead_and_dlogits():
    from spacer_b200 import ops
    M, V, K = 300, 5000 // 8 * 8, 512
    H, W = _mk((M, K), 15, 1.0), _mk((V, K), 16, 0.05)
    tg = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
    nt = (V + 255) // 256
    part = torch.empty((M, nt, 2), device="cuda", dtype=torch.float32)
    tl = torch.zeros((M,), device="cuda", dtype=torch.float32)
    ops.gemm(H, W, epilogue=ops.EPI_LMHEAD, targets=tg, lse_part=part, tgt_logit=tl)
    logits = (H.float() @ W.float().t()).bfloat16().float()
    mx = part[..., 0].max(dim=1).values
    lse = mx + torch.log((part[..., 1] * torch.exp(part[..., 0] - mx[:, None])).sum(1))
    ref_lse = torch.logsumexp(logits, dim=-1)
    assert (lse - ref_lse).abs().max().item() < 1e-3
    ref_t = logits.gather(1, tg.long()[:, None])[:, 0]
    assert (tl - ref_t).abs().max().item() < 1e-6 + 1e-2 * ref_t.abs().max().item()
    coef = torch.randn((M,), device="cuda", dtype=torch.float32)
    d = ops.gemm(H, W, epilogue=ops.EPI_DLOGITS, targets=tg, lse=ref_lse.contiguous(), coef=coef)
    p = torch.softmax(logits, -1)
    oh = torch.nn.functional.one_hot(tg.long(), V).float()
    ref_d = coef[:, None] * (oh - p)
    assert (d.float() - ref_d).abs().max().item() < 1e-2 * ref_d.abs().max().item() + 1e-5


@pytest.mark.parametrize("R,I,K", [(12, 512, 256), (16, 18944, 3584), (24, 1024, 512)])
def test_gemm_f32t_swiglu_decode(R, I, K):
    """Decode gate|up GEMV with the SwiGLU fused into the swap-AB epilogue == unfused partials + sb_dec_swiglu math."""
    from spacer_b200 import ops
    RP = 16 if R <= 16 else 32
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(2 * I, K, device="cuda", generator=g) * 0.05).bfloat16()
    x = torch.zeros(RP, K, device="cuda", dtype=torch.bfloat16)
    x[:R] = torch.randn(R, K, device="cuda", generator=g).bfloat16()
    act = torch.full((RP, I), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(w, x, out=act, epilogue=ops.EPI_F32T_SWIGLU)
    gu = (x.float() @ w.float().t()).bfloat16().float().view(RP, I // 64, 2, 64)
    ref = (torch.nn.functional.silu(gu[:, :, 0]).bfloat16().float() * gu[:, :, 1]).reshape(RP, I)
    err = (act.float() - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item() + 1e-3, err
    with pytest.raises(ops.SpacerError):
        ops.gemm(w, x, out=act, epilogue=ops.EPI_F32T_SWIGLU, k_splits=2)

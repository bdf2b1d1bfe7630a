"""GPU parity of the decode-step kernels, the sampler, the fused GRPO loss and AdamW against torch / the oracle."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def close(a, b, tol=2e-2, name=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    den = b.abs().max().item() + 1e-6
    assert err / den < tol, f"{name}: max abs err {err} vs scale {den}"


def test_dec_residual_rmsnorm_and_swiglu():
    from spacer_b200 import ops
    R, H, S = 12, 3584, 4
    x = rnd((16, H), 1)
    parts = rnd((S, 16, H), 2, 0.3, torch.float32)
    w = 1 + rnd((H,), 3, 0.1)
    xn = torch.zeros((16, H), device="cuda", dtype=torch.bfloat16)
    x2 = x.clone()
    ops.call("sb_dec_residual_rmsnorm", x2, parts, S, 16 * H, H, w, xn, R, H, 1e-6)
    ref_x = (parts.sum(0).bfloat16().float() + x.float()).bfloat16()
    assert torch.equal(x2[:R], ref_x[:R]) and torch.equal(x2[R:], x[R:])
    xf = ref_x.float()
    ref_n = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
    close(xn[:R], ref_n[:R], 1e-2)
    I = 512
    pg = rnd((S, 16, 2 * I), 4, 0.5, torch.float32)
    act = torch.zeros((16, I), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_dec_swiglu", pg, S, 16 * 2 * I, 2 * I, act, R, I)
    gu = pg.sum(0).bfloat16().float().view(16, I // 64, 2, 64)
    ref = (F.silu(gu[:, :, 0]).bfloat16().float() * gu[:, :, 1]).reshape(16, I)
    close(act[:R], ref[:R], 1e-2)


def test_dec_qkv_post_and_attention():
    from oracle import qwen2vl_ref as Rf
    from spacer_b200 import ops
    R, nh, nkv, hd, S, P, Cmax = 6, 14, 2, 128, 3, 150, 40
    W = (nh + 2 * nkv) * hd
    step = 17
    rope_base = 93
    parts = rnd((S, 16, W), 1, 0.4, torch.float32)
    bias = rnd((W,), 2, 0.1)
    step_t = torch.tensor([step], dtype=torch.int32, device="cuda")
    q_out = torch.zeros((16, nh * hd), device="cuda", dtype=torch.bfloat16)
    kc = rnd((R, Cmax, nkv * hd), 3, 0.5)
    vc = rnd((R, Cmax, nkv * hd), 4, 0.5)
    kc0, vc0 = kc.clone(), vc.clone()
    ops.call("sb_dec_qkv_post", parts, S, 16 * W, W, bias, step_t, rope_base, 1e6, nh, nkv, hd, q_out, kc, vc,
             Cmax * nkv * hd, Cmax, R)
    qkv = (parts.sum(0) + bias.float()).bfloat16().float()[:R]
    d = Rf.Dims(heads=nh, kv_heads=nkv)
    pos = torch.full((3, 1, 1), rope_base + step)
    cos, sin = Rf.mrope_cos_sin(pos, d)
    cos, sin = cos[0, 0].cuda().bfloat16().float(), sin[0, 0].cuda().bfloat16().float()
    rot = lambda x: torch.cat((-x[..., hd // 2:], x[..., :hd // 2]), -1)
    q = qkv[:, :nh * hd].view(R, nh, hd)
    k = qkv[:, nh * hd:(nh + nkv) * hd].view(R, nkv, hd)
    v = qkv[:, (nh + nkv) * hd:]
    qr, kr = q * cos + rot(q) * sin, k * cos + rot(k) * sin
    close(q_out[:R], qr.reshape(R, -1), 1e-2, "q")
    close(kc[:, step], kr.reshape(R, -1), 1e-2, "k slot")
    close(vc[:, step], v, 1e-2, "v slot")
    mask = torch.ones(Cmax, dtype=torch.bool)
    mask[step] = False
    assert torch.equal(kc[:, mask], kc0[:, mask]) and torch.equal(vc[:, mask], vc0[:, mask])
    # attention over two prompt caches + completion cache
    kp0, vp0, kp1, vp1 = (rnd((P, nkv * hd), s, 0.5) for s in (5, 6, 7, 8))
    import ctypes
    rep = nh // nkv
    out = torch.zeros((16, nh * hd), device="cuda", dtype=torch.bfloat16)
    g0 = 4
    nws = ctypes.c_longlong(0)
    assert ops._lib.load().sb_dec_attn_workspace(R, g0, P, Cmax, nh, nkv, ctypes.byref(nws)) == 0
    ws = torch.empty(nws.value, device="cuda", dtype=torch.float32)
    ops.call("sb_dec_attn", q_out, kp0, vp0, kp1, vp1, g0, P, kc, vc, Cmax * nkv * hd, Cmax, step_t, nh, nkv, hd,
             hd ** -0.5, ws, ws.numel(), out, R)
    for r in range(R):
        kp, vp = (kp0, vp0) if r < g0 else (kp1, vp1)
        K = torch.cat([kp, kc[r, :step + 1]]).float().view(-1, nkv, hd).repeat_interleave(rep, 1)
        V = torch.cat([vp, vc[r, :step + 1]]).float().view(-1, nkv, hd).repeat_interleave(rep, 1)
        qq = q_out[r].float().view(nh, hd)
        s = torch.einsum("hd,jhd->hj", qq, K) * hd ** -0.5
        ref = torch.einsum("hj,jhd->hd", torch.softmax(s, -1), V).reshape(-1)
        close(out[r], ref, 1e-2, f"attn row {r}")


@pytest.mark.parametrize("R,g0,nh,nkv,P,Cmax,step", [
    (12, 8, 28, 4, 2304, 512, 255),     # cfg3 shape: 8 + 4 rows, 7 q heads per kv head, 18 prompt splits
    (12, 8, 28, 4, 2304, 512, 0),       # first decode step: one completion key
    (12, 8, 28, 4, 2304, 512, 511),     # last slot
    (24, 16, 12, 2, 832, 1024, 700),    # G=16 (+8): two 64-query blocks per (group, kv head), 8 completion splits
    (3, 3, 4, 4, 70, 40, 5),            # one group, rep 1, ragged prompt tile
    (5, 0, 8, 2, 130, 64, 63),          # every row in the second group
    (32, 16, 28, 4, 2304, 512, 500),    # 32 rows leave room for ONE completion split: 8 key tiles per own item (the
                                        # current token's row sits in a tile that is loaded inside the loop)
    (32, 16, 28, 4, 2304, 512, 100),    # same plan, current token in the second pre-loaded tile
])
def test_dec_attention_shapes(R, g0, nh, nkv, P, Cmax, step):
    import ctypes
    from spacer_b200 import ops
    hd = 128
    rep = nh // nkv
    RP = 16 if R <= 16 else 32
    q = rnd((RP, nh * hd), 11, 1.0)
    kp0, vp0, kp1, vp1 = (rnd((P, nkv * hd), s, 0.7) for s in (5, 6, 7, 8))
    kc = rnd((R, Cmax, nkv * hd), 3, 0.7)
    vc = rnd((R, Cmax, nkv * hd), 4, 0.7)
    step_t = torch.tensor([step], dtype=torch.int32, device="cuda")
    nws = ctypes.c_longlong(0)
    assert ops._lib.load().sb_dec_attn_workspace(R, g0, P, Cmax, nh, nkv, ctypes.byref(nws)) == 0
    ws = torch.full((nws.value,), float("nan"), device="cuda", dtype=torch.float32)
    out = torch.zeros((RP, nh * hd), device="cuda", dtype=torch.bfloat16)
    ops.call("sb_dec_attn", q, kp0, vp0, kp1, vp1, g0, P, kc, vc, Cmax * nkv * hd, Cmax, step_t, nh, nkv, hd,
             hd ** -0.5, ws, ws.numel(), out, R)
    for r in range(R):
        kp, vp = (kp0, vp0) if r < g0 else (kp1, vp1)
        K = torch.cat([kp, kc[r, :step + 1]]).float().view(-1, nkv, hd).repeat_interleave(rep, 1)
        V = torch.cat([vp, vc[r, :step + 1]]).float().view(-1, nkv, hd).repeat_interleave(rep, 1)
        s = torch.einsum("hd,jhd->hj", q[r].float().view(nh, hd), K) * hd ** -0.5
        ref = torch.einsum("hj,jhd->hd", torch.softmax(s, -1), V).reshape(-1)
        close(out[r], ref, 1e-2, f"attn row {r}")
    assert torch.equal(out[R:], torch.zeros_like(out[R:]))


def _expected_dist(logits, top_p, suppress=None):
    from oracle import qwen2vl_ref as Rf
    l = logits.bfloat16().float().cpu()
    if suppress is not None:
        l[:, suppress] = float("-inf")
    return torch.softmax(Rf.top_p_filter(l, top_p), dim=-1)


@pytest.mark.parametrize("V,scale", [(2048, 3.0), (152064, 1.0), (5003, 6.0)])
def test_sampler_distribution(V, scale):
    """Kept set == the TopPLogitsWarper kept set; empirical frequencies match the filtered softmax
    (total-variation distance over many Philox draws)."""
    from spacer_b200 import ops
    R = 4
    logits = rnd((R, V), 21, scale, torch.float32)
    if V == 2048:
        logits[1, :] = 0.0          # all tied
        logits[2, 7] = 30.0         # one dominant token: top-1 always kept
    exp = _expected_dist(logits, 0.95)
    n_draws = 4000
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    toks = torch.empty(R, dtype=torch.int32, device="cuda")
    ids = torch.zeros((R, n_draws), dtype=torch.int32, device="cuda")
    lp = torch.empty(R, dtype=torch.float32, device="cuda")
    for i in range(n_draws):
        ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 1234, step, None, toks, ids, n_draws, lp, -1, 0, 0, None)
        ops.call("sb_step_advance", step)
    torch.cuda.synchronize()
    ids = ids.cpu().long()
    lv = logits.bfloat16().float().cpu()
    for r in range(R):
        # Which members of a tie group at the cut are dropped depends on the sort's tie order (implementation
        # defined in torch), so compare per distinct logit VALUE: mass of every value group must match.
        vals, inv = torch.unique(lv[r], return_inverse=True)
        exp_g = torch.zeros(len(vals)).index_add_(0, inv, exp[r])
        got_g = torch.bincount(inv[ids[r]], minlength=len(vals)).float() / n_draws
        assert (exp_g[inv[ids[r]]] > 0).all(), f"row {r}: sampled a value group the reference warper removes"
        kept_groups = int((exp_g > 0).sum())
        tv = 0.5 * (got_g - exp_g).abs().sum().item()
        assert tv < 3 * math.sqrt(kept_groups / (2 * math.pi * n_draws)) + 0.02, f"row {r}: TV {tv}, groups {kept_groups}"
        # within the cut tie group the number of distinct survivors cannot exceed the reference's count
        cut = int(torch.nonzero(exp_g > 0)[0])
        n_keep_ref = int(((inv == cut) & (exp[r] > 0)).sum())
        n_keep_got = len(set(ids[r][inv[ids[r]] == cut].tolist()))
        assert n_keep_got <= n_keep_ref, f"row {r}: {n_keep_got} distinct survivors in the cut group, reference keeps {n_keep_ref}"
    # different seeds / rows give different streams, same seed reproduces
    a = torch.empty(R, dtype=torch.int32, device="cuda")
    b = torch.empty(R, dtype=torch.int32, device="cuda")
    step.zero_()
    ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 99, step, None, a, None, 0, None, -1, 0, 0, None)
    ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 99, step, None, b, None, 0, None, -1, 0, 0, None)
    assert torch.equal(a, b)


def test_sampler_kept_set_exact():
    """Draw many samples from a small vocabulary: the support must equal the reference kept set exactly."""
    from spacer_b200 import ops
    V, R = 64, 8
    logits = rnd((R, V), 5, 2.0, torch.float32)
    exp = _expected_dist(logits, 0.95)
    n = 3000
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    toks = torch.empty(R, dtype=torch.int32, device="cuda")
    ids = torch.zeros((R, n), dtype=torch.int32, device="cuda")
    for i in range(n):
        ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 7, step, None, toks, ids, n, None, -1, 0, 0, None)
        ops.call("sb_step_advance", step)
    ids = ids.cpu().long()
    lv = logits.bfloat16().float().cpu()
    for r in range(R):
        support = set(ids[r].tolist())
        kept = set(torch.nonzero(exp[r] > 0).flatten().tolist())
        vmin = lv[r][sorted(kept)].min()
        # strictly-above-the-cut tokens are unambiguous; ties at the cut value may differ by sort order
        assert all(lv[r][t] >= vmin for t in support)
        big = {i for i in kept if exp[r][i] > 5e-3 and lv[r][i] > vmin}
        assert big <= support


def test_sampler_eos_and_finished():
    from spacer_b200 import ops
    V, R = 512, 3
    logits = torch.full((R, V), -20.0, device="cuda")
    logits[:, 5] = 20.0     # always samples token 5
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    fin = torch.tensor([0, 1, 0], dtype=torch.int32, device="cuda")
    toks = torch.empty(R, dtype=torch.int32, device="cuda")
    ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 1, step, fin, toks, None, 0, None, 5, 9, 0, None)
    assert toks.tolist() == [5, 9, 5] and fin.tolist() == [1, 1, 1]
    ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 1, step, fin, toks, None, 0, None, 5, 9, 0, None)
    assert toks.tolist() == [9, 9, 9]
    fin.zero_()
    logits[:, 6] = 19.0
    ops.call("sb_sample_top_p", logits, V, R, V, 0.95, 1, step, fin, toks, None, 0, None, 5, 9, 1, None)  # EOS suppressed
    assert toks.tolist() == [6, 6, 6] and fin.tolist() == [0, 0, 0]


def _draw(logits, n_draws, seed=1234, before_each=None, **opts):
    from spacer_b200 import ops
    R = logits.shape[0]
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    toks = torch.empty(R, dtype=torch.int32, device="cuda")
    ids = torch.zeros((R, n_draws), dtype=torch.int32, device="cuda")
    for _ in range(n_draws):
        if before_each is not None:
            before_each()
        ops.sample(logits, step, toks, seed=seed, out_ids=ids, **opts)
        ops.call("sb_step_advance", step)
    torch.cuda.synchronize()
    return ids.cpu().long()


def _hf_chain(logits_bf16_f32, context, rep_pen, temperature, top_k, top_p, reround):
    """The real transformers processors in generate()'s order (generation/utils.py `_get_logits_processor`): repetition
    penalty, temperature, top-k, top-p.  `reround` reproduces this engine's one deliberate deviation (values changed by
    penalty / temperature go back to bf16 before the cuts)."""
    from transformers.generation.logits_process import (RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
    s = logits_bf16_f32.clone()
    changed = torch.zeros_like(s, dtype=torch.bool)
    if rep_pen != 1.0:
        s2 = RepetitionPenaltyLogitsProcessor(rep_pen)(context, s)
        changed |= s2 != s
        s = s2
    if temperature != 1.0:
        s = TemperatureLogitsWarper(temperature)(context, s)
        changed |= True
    if reround:
        s = torch.where(changed, s.bfloat16().float(), s)
    if top_k > 0:
        s = TopKLogitsWarper(top_k)(context, s)
    if top_p < 1.0:
        s = TopPLogitsWarper(top_p)(context, s)
    return torch.softmax(s, dim=-1)


@pytest.mark.parametrize("V,rep_pen,temperature,top_k,top_p", [
    (4096, 1.0, 1.0, 50, 0.95),        # the reference's rollout: HF's implicit top_k = 50 under top_p 0.95
    (152064, 1.0, 1.0, 50, 0.95),
    (4096, 1.05, 0.7, 20, 0.9),
    (152064, 1.3, 1.5, 50, 1.0),
    (5003, 1.0, 0.5, 0, 0.8),
    (2048, 1.2, 1.0, 3, 1.0),
])
def test_sampler_matches_hf_processor_chain(V, rep_pen, temperature, top_k, top_p):
    """Support and frequencies of sb_sample against transformers' own RepetitionPenalty / Temperature / TopK / TopP
    processors + softmax on the same logits."""
    from spacer_b200 import ops
    R, n_draws, n_ctx = 4, 3000, 40
    logits = rnd((R, V), 31, 2.5 if V < 10000 else 1.2, torch.float32)
    logits[0, 11] = logits[0].max() + 1.0
    if top_k > 0:
        logits[3, :8] = 9.0           # ties at the k-th value: TopKLogitsWarper keeps all of them
    g = torch.Generator().manual_seed(3)
    context = torch.randint(0, V, (R, n_ctx), generator=g)
    context[:, :5] = torch.arange(5)[None] + 9   # make sure some high-probability tokens are penalised
    seen = torch.zeros((R, (V + 31) // 32), dtype=torch.int32, device="cuda")
    for r in range(R):
        ops.call("sb_token_bitmap_set", context[r].to(torch.int32).cuda(), n_ctx, seen[r:], seen.stride(0), 1, V)
    seen0 = seen.clone()
    # every draw is one independent "next token" of the same context: the bitmap (which the sampler extends with the token
    # it emits) is put back before each draw
    ids = _draw(logits, n_draws, top_p=top_p, top_k=top_k, temperature=temperature, repetition_penalty=rep_pen,
                seen=seen if rep_pen != 1.0 else None, before_each=lambda: seen.copy_(seen0))
    lb = logits.bfloat16().float().cpu()
    exact = _hf_chain(lb, context, rep_pen, temperature, top_k, top_p, reround=True)
    hf = _hf_chain(lb, context, rep_pen, temperature, top_k, top_p, reround=False)
    for r in range(R):
        got = torch.bincount(ids[r], minlength=V).float() / n_draws
        kept = int((exact[r] > 0).sum())
        assert (exact[r][ids[r]] > 0).all() or top_p < 1.0, f"row {r}: sampled a token the processor chain removes"
        tv = 0.5 * (got - exact[r]).abs().sum().item()
        bound = 3 * math.sqrt(max(kept, 1) / (2 * math.pi * n_draws)) + 0.02
        assert tv < bound, f"row {r}: TV {tv} vs own-rounding chain (kept {kept}, bound {bound})"
        tv_hf = 0.5 * (got - hf[r]).abs().sum().item()
        assert tv_hf < bound + 0.03, f"row {r}: TV {tv_hf} vs the fp32 HF chain"
        if top_k > 0 and top_p >= 1.0:
            assert set(ids[r].tolist()) <= set(torch.nonzero(exact[r] > 0).flatten().tolist())
    if top_k > 0 and top_p >= 1.0:
        assert int((exact[3] > 0).sum()) >= 8 or rep_pen != 1.0          # the tie group survives whole
    if rep_pen != 1.0:   # the sampled token was OR-ed into the bitmap (it counts for the next step's penalty)
        for r in range(R):
            last = int(ids[r, -1])
            assert (int(seen[r, last >> 5].item()) >> (last & 31)) & 1


def test_sampler_topk_tie_overflow_takes_generic_path():
    """300 tokens tie at the k-th value inside ONE cluster slice: more candidates than the top-k fast path gathers per CTA
    (KCAP = 128), so the kernel falls back to the generic path; TopKLogitsWarper keeps the whole tie group either way."""
    V, R, n_draws = 4096, 2, 4000
    logits = rnd((R, V), 77, 1.0, torch.float32)
    logits[0, 100:400] = 9.0
    context = torch.zeros((R, 1), dtype=torch.long)
    ids = _draw(logits, n_draws, top_p=1.0, top_k=50, temperature=1.0, repetition_penalty=1.0)
    exact = _hf_chain(logits.bfloat16().float().cpu(), context, 1.0, 1.0, 50, 1.0, reround=True)
    assert int((exact[0] > 0).sum()) == 300
    for r in range(R):
        kept = set(torch.nonzero(exact[r] > 0).flatten().tolist())
        assert set(ids[r].tolist()) <= kept
        got = torch.bincount(ids[r], minlength=V).float() / n_draws
        tv = 0.5 * (got - exact[r]).abs().sum().item()
        assert tv < 3 * math.sqrt(len(kept) / (2 * math.pi * n_draws)) + 0.02, (r, tv)
    assert len(set(ids[0].tolist())) > 200       # the tie group is really sampled from, not truncated to 128


def test_sampler_greedy_with_repetition_penalty_and_eos_list():
    """Evaluation decoding (checkpoint generation_config: repetition_penalty 1.05, top_k 1; vsibench.py:174): argmax of
    the penalised logits; ANY id of the eos list finishes a row."""
    from spacer_b200 import ops
    V, R = 1000, 3
    logits = torch.zeros((R, V), device="cuda")
    logits[:, 5] = 10.0
    logits[:, 6] = 9.8
    logits[2, 7] = 10.5
    seen = torch.zeros((R, (V + 31) // 32), dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    fin = torch.zeros(R, dtype=torch.int32, device="cuda")
    toks = torch.empty(R, dtype=torch.int32, device="cuda")
    kw = dict(greedy=True, repetition_penalty=1.05, seen=seen, finished=fin, eos_ids=(7, 900), pad_id=3)
    ops.sample(logits, step, toks, **kw)
    assert toks.tolist() == [5, 5, 7] and fin.tolist() == [0, 0, 1]
    ops.sample(logits, step, toks, **kw)             # 5 was emitted: 10 / 1.05 = 9.52 < 9.8 -> 6; row 2 pads
    assert toks.tolist() == [6, 6, 3]
    ops.sample(logits, step, toks, **kw)             # 6 penalised too: 9.33 < 9.52 -> 5 again
    assert toks.tolist() == [5, 5, 3]
    logits[0, 900] = 50.0
    ops.sample(logits, step, toks, **kw)
    assert toks.tolist()[0] == 900 and fin.tolist() == [1, 0, 1]
    fin.zero_()
    ops.sample(logits, step, toks, suppress_eos=True, **kw)     # both eos ids suppressed
    assert toks.tolist()[0] not in (7, 900) and toks.tolist()[2] not in (7, 900) and fin.tolist() == [0, 0, 0]


@pytest.mark.parametrize("G,C,V", [(4, 37, 1000), (4, 150, 1200), (4, 64, 256)])
def test_grpo_loss_matches_oracle(G, C, V):
    from oracle import grpo_ref as GR
    from spacer_b200 import ops
    K, eos = 128, 77
    torch.manual_seed(1234 + C)   # fixed data: a bf16 rounding flip of a target logit (p ~ 1e-4 each) would move lp by an ulp
    H, Wm = rnd((G * C, K), 1), rnd((V, K), 2, 0.2)
    comp = torch.randint(0, V, (G, C))
    comp[comp == eos] = 0
    comp[1, 10] = eos
    comp[2, C - 1] = eos
    tg = comp.reshape(-1).to(torch.int32).cuda()
    nt = (V + 255) // 256
    part = torch.empty((G * C, nt, 2), device="cuda", dtype=torch.float32)
    tl = torch.zeros(G * C, device="cuda")
    ops.gemm(H, Wm, epilogue=ops.EPI_LMHEAD, targets=tg, lse_part=part, tgt_logit=tl)
    logits = (H.float() @ Wm.float().t()).bfloat16().float().cpu()
    lp_ref = torch.log_softmax(logits, -1).gather(1, comp.reshape(-1, 1)).view(G, C)
    ref_lp = lp_ref + torch.randn(G, C) * 0.5
    ref_lp[0, 3] = lp_ref[0, 3] + 25.0   # exercise the clamp
    adv = torch.tensor([0.7, -1.2, 0.1, 0.4])
    mask = GR.completion_mask(comp, eos)
    lp_t = lp_ref.clone().requires_grad_()
    loss, kl = GR.grpo_loss(lp_t, ref_lp, adv, mask, 0.04)
    loss.backward()
    outs = [torch.empty(G * C, device="cuda") for _ in range(3)]
    mask_o = torch.empty(G * C, dtype=torch.int32, device="cuda")
    rl, rk = torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    rlen = torch.empty(G, dtype=torch.int32, device="cuda")
    out2 = torch.empty(2, device="cuda")
    ws = ops.grpo_loss_workspace(G, C, "cuda")
    for _ in range(2):   # twice: the ticket counter in the workspace must reset itself
        out2.fill_(float("nan"))
        ops.call("sb_grpo_loss", part, nt, tl, comp.to(torch.int32).cuda(), G, C, eos, ref_lp.cuda().contiguous(),
                 adv.cuda(), 0.04, outs[0], outs[1], outs[2], mask_o, rl, rk, rlen, out2, ws)
    assert (outs[0].cpu().view(G, C) - lp_ref).abs().max() < 2e-3
    assert torch.equal(mask_o.cpu().view(G, C), mask)
    assert rlen.tolist() == mask.sum(1).tolist()
    assert abs(out2[0].item() - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    assert abs(out2[1].item() - kl.item()) < 1e-3 * max(1.0, abs(kl.item()))
    g_ref = GR.grpo_loss_grad(outs[0].cpu().view(G, C), ref_lp, adv, mask, 0.04)
    assert (outs[2].cpu().view(G, C) - g_ref).abs().max() < 1e-5
    assert (lp_t.grad - g_ref).abs().max() < 1e-3
    lp2 = torch.empty(G * C, device="cuda")
    ops.call("sb_logprob_from_partials", part, nt, tl, lp2, G * C)
    assert torch.allclose(lp2, outs[0])


@pytest.mark.parametrize("grad_f32,mom_bf16", [(0, 0), (1, 0), (0, 1)])
def test_adamw_matches_torch(grad_f32, mom_bf16):
    from spacer_b200 import ops
    n = 100003
    p0 = rnd((n,), 1, 0.02)
    master = p0.float().clone()
    mdt = torch.bfloat16 if mom_bf16 else torch.float32
    m, v = torch.zeros(n, device="cuda", dtype=mdt), torch.zeros(n, device="cuda", dtype=mdt)
    ref_p = torch.nn.Parameter(p0.float().clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p = p0.clone()
    for step in range(1, 4):
        g = rnd((n,), 10 + step, 3.0, torch.float32 if grad_f32 else torch.bfloat16)
        ref_p.grad = g.float().clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 5.0)
        opt.step()
        tot = torch.zeros(1, device="cuda")
        ops.call("sb_grad_sumsq", g, n, grad_f32, tot)
        assert abs(tot.item() - g.float().pow(2).sum().item()) < 1e-3 * tot.item()
        ops.call("sb_adamw_step", p, master, m, v, g, n, grad_f32, mom_bf16, tot, 1e-3, 0.9, 0.999, 1e-8, 0.01, step,
                 5.0, 1.0)
    tol = 2e-3 if mom_bf16 else 1e-5
    assert (master - ref_p.data).abs().max().item() < tol
    assert torch.equal(p, master.bfloat16())


def test_l2_prefetch_kernel_forms():
    """sb_dec_l2_prefetch only warms L2: both forms (bulk/TMA and paced per-line) run over a strided row subset and over one
    linear run without touching the data; bad arguments are rejected."""
    from spacer_b200 import ops
    w = torch.randn((1024, 512), device="cuda").bfloat16()
    ref = w.clone()
    k = w.shape[1]
    for pace in (-1, 0, 300):
        ops.call("sb_dec_l2_prefetch", w, 24 * k * 2, 128 * k * 2, w.shape[0] // 128, 0, pace)
        ops.call("sb_dec_l2_prefetch", w, w.numel() * 2, 0, 1, 16, pace)
    torch.cuda.synchronize()
    assert torch.equal(w, ref)
    with pytest.raises(Exception):
        ops.call("sb_dec_l2_prefetch", w, 0, 0, 1, 0, -1)
    with pytest.raises(Exception):
        ops.call("sb_dec_l2_prefetch", w, 200, 128 * k * 2, 4, 0, 100)      # paced form needs whole 128-byte lines

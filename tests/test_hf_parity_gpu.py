"""Parity at the BENCHMARK's own model sizes against the real third-party stack the reference calls:
`transformers.Qwen2VLForConditionalGeneration` (5.5.0) running on the same B200 with the SAME weights, bf16 +
flash_attention_2 (what run_SpaceR_SG_RLVR.sh:31 configures), explicit classic position ids and mm_token_type_ids
(SURVEY.md 8(c) drift #1/#2) -- the reference's own scoring call (SG_RLVR_trainer.py:353-366, 526-547) and loss (:551-552,
640-643).  Compared: per-token log-probs, the GRPO loss, and EVERY parameter gradient (cosine + norm ratio), at cfg1 dims
(Qwen2-VL-2B, 2 x 224^2, G=2), cfg2 dims (2B, 8 x 336^2, G=4, C=512) and cfg3 dims (7B, 16 x 448^2; 2 rows to bound
the HF side's memory).  At cfg1 the HF model is ALSO run in fp32 (eager attention): both bf16 implementations are scored
against it, and this engine must not be further from fp32 than HF-bf16 is -- that is what licenses a tolerance above the
2e-2 / 3e-3 that SURVEY 8(c) states for bf16-vs-fp32.  Statistics are written to gpurun_out/hf_parity_*.json."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu


def _hf_model(d, dtype, attn):
    from transformers import Qwen2VLForConditionalGeneration
    from hf_gpu_baseline import hf_config
    hc = hf_config(d)
    hc._attn_implementation = attn
    with torch.device("cuda"):
        m = Qwen2VLForConditionalGeneration(hc).to(dtype)
    return m


def _fa2_or_sdpa():
    try:
        import flash_attn  # noqa: F401
        return "flash_attention_2"
    except Exception:
        return "sdpa"


def _case(d, frames, res, text, G, C, seed):
    """Synthetic step inputs per SURVEY 8(d): uint8-valued frames through the HF processor arithmetic, prompt ids with the
    vision block, random completions (one row ends early with EOS + pad, like a real rollout)."""
    import bench
    ex = bench.synth_example(d, dict(frames=frames, res=res, text=text), seed)
    g = torch.Generator().manual_seed(seed + 1)
    comp = torch.randint(1000, min(100000, d.vocab - 1000), (G, C), generator=g)
    if C >= 8:
        comp[1, C // 2] = d.eos_id
        comp[1, C // 2 + 1:] = d.pad_id
    prompt = ex["input_ids"]
    ids = torch.cat([prompt.repeat(G, 1), comp], dim=1)
    return dict(prompt=prompt, comp=comp, ids=ids, pix=ex["pixel_values_host"].clone(), grid=ex["video_grid_thw"])


def _hf_scores(hf, d, case, G, ref_lp, adv, mask, beta, dtype, ckpt=False):
    """The reference's scoring + loss + backward on the HF model.  Returns (lp [G, C] fp32, loss, {name: grad})."""
    from spacer_b200.model import rope_index
    ids = case["ids"].cuda()
    P = case["prompt"].shape[1]
    pos = torch.stack([rope_index(row, case["grid"], d, "classic")[0] for row in case["ids"]], dim=1).cuda()
    mm = (ids == d.video_token_id).long() * 2 + (ids == d.image_token_id).long()
    pixG = case["pix"].cuda().to(dtype).repeat(G, 1)
    gridG = case["grid"].cuda().repeat(G, 1)
    hf.train()
    if ckpt:
        hf.gradient_checkpointing_enable()
    hf.zero_grad(set_to_none=True)
    with torch.set_grad_enabled(ref_lp is not None):
        logits = hf(input_ids=ids, pixel_values_videos=pixG, video_grid_thw=gridG, position_ids=pos, mm_token_type_ids=mm,
                    use_cache=False).logits
        # TRN:353-366 with the log-softmax in fp32 (the engine's choice; strictly more accurate than the reference's bf16)
        lp = torch.stack([torch.gather(lr.float().log_softmax(-1), 1, ir.unsqueeze(1)).squeeze(1)
                          for lr, ir in zip(logits[:, :-1], ids[:, 1:])])[:, P - 1:]
    del logits
    if ref_lp is None:
        return lp.detach(), None, None
    x = torch.clamp(ref_lp - lp, -10, 10)
    kl = torch.exp(x) - x - 1
    per_tok = -(torch.exp(lp - lp.detach()) * adv.unsqueeze(1) - beta * kl)
    loss = ((per_tok * mask).sum(1) / mask.sum(1)).mean()
    loss.backward()
    grads = {k: p.grad.detach() for k, p in hf.named_parameters() if p.grad is not None}
    return lp.detach(), loss.detach(), grads


def _grad_stats(a: dict, b: dict):
    """Per tensor: cosine and norm ratio of a vs b (b = the yardstick)."""
    out = {}
    for k, gb in b.items():
        ga = a[k].float().reshape(-1)
        gb = gb.float().reshape(-1).to(ga.device)
        nb = gb.norm()
        out[k] = (torch.nn.functional.cosine_similarity(ga, gb, dim=0).item(), (ga.norm() / (nb + 1e-30)).item(), nb.item(),
                  gb.numel())
    return out


def _run(preset, frames, res, text, G, C, tag, fp32_truth, ckpt=False, beta=0.04):
    from oracle import grpo_ref as GR
    from spacer_b200 import config as mcfg
    from spacer_b200.model import GradStore, Qwen2VLB200, pack_prompt_completions
    d = mcfg.PRESETS[preset]()
    eng = Qwen2VLB200(d, "cuda")
    eng.params.init_random(seed=0)
    sd = eng.state_dict()
    case = _case(d, frames, res, text, G, C, seed=1234)
    mask = GR.completion_mask(case["comp"], d.eos_id).cuda()
    adv = GR.advantages(torch.linspace(0.0, 2.0, G), G)[0].cuda()
    attn = _fa2_or_sdpa()
    hf = _hf_model(d, torch.bfloat16, attn)
    missing = hf.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("lm_head" in k for k in missing.missing_keys), missing
    lp_hf0, _, _ = _hf_scores(hf, d, case, G, None, adv, mask, beta, torch.bfloat16, ckpt)
    ref_lp = lp_hf0 + 0.05 * torch.randn(lp_hf0.shape, generator=torch.Generator().manual_seed(9)).cuda()
    lp_hf, loss_hf, g_hf = _hf_scores(hf, d, case, G, ref_lp, adv, mask, beta, torch.bfloat16, ckpt)
    del hf
    torch.cuda.empty_cache()
    # this engine, same weights, same inputs
    batch = pack_prompt_completions(case["prompt"], case["comp"], case["grid"], d, eng.device)
    grads = GradStore(eng.params)
    out = eng.grpo_forward_backward(batch, case["pix"].cuda(), case["grid"], ref_lp, adv, beta, grads)
    torch.cuda.synchronize()
    lp = out["logps"]
    g_me = dict(eng.params.hf_items({n: grads[n] for n in eng.params.index}))
    if d.tie:
        g_me.pop("lm_head.weight", None)
    g_me = {k: v for k, v in g_me.items() if k in g_hf}
    assert set(g_me) == set(g_hf), set(g_hf) ^ set(g_me)
    m = mask.bool()
    err = (lp - lp_hf).abs()[m]
    st = _grad_stats(g_me, g_hf)
    stats = {"tag": tag, "hf_attn": attn, "G": G, "C": C, "P": int(case["prompt"].shape[1]),
             "lp_max_err_vs_hf_bf16": err.max().item(), "lp_mean_err_vs_hf_bf16": err.mean().item(),
             "loss": out["loss"].item(), "loss_hf_bf16": loss_hf.item(),
             "grad_cos_min_vs_hf": min(v[0] for v in st.values()),
             "grad_cos_mean_vs_hf": sum(v[0] for v in st.values()) / len(st),
             "grad_cos_weighted_vs_hf": sum(v[0] * v[2] ** 2 for v in st.values()) / sum(v[2] ** 2 for v in st.values()),
             "grad_norm_ratio_range_vs_hf": [min(v[1] for v in st.values()), max(v[1] for v in st.values())],
             "worst_tensors_vs_hf": sorted(((k, round(v[0], 5), round(v[1], 4)) for k, v in st.items()), key=lambda t: t[1])[:6],
             "n_tensors": len(st)}
    if fp32_truth == "forward":      # 7B: the fp32 network fits for a forward only (no gradients next to everything else)
        del g_hf, g_me, grads, st
        torch.cuda.empty_cache()
        hf32 = _hf_model(d, torch.float32, "sdpa")
        hf32.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
        del sd
        lp32, _, _ = _hf_scores(hf32, d, case, G, None, adv, mask, beta, torch.float32)
        e_me, e_hf = (lp - lp32).abs()[m], (lp_hf - lp32).abs()[m]
        stats.update({"lp_max_err_vs_fp32": e_me.max().item(), "lp_mean_err_vs_fp32": e_me.mean().item(),
                      "hf_bf16_lp_max_err_vs_fp32": e_hf.max().item(), "hf_bf16_lp_mean_err_vs_fp32": e_hf.mean().item()})
        del hf32
        grads = None
    elif fp32_truth:
        hf32 = _hf_model(d, torch.float32, "eager")
        hf32.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
        lp32, loss32, g32 = _hf_scores(hf32, d, case, G, ref_lp, adv, mask, beta, torch.float32)
        e_me, e_hf = (lp - lp32).abs()[m], (lp_hf - lp32).abs()[m]
        s_me, s_hf = _grad_stats(g_me, g32), _grad_stats(g_hf, g32)
        stats.update({"lp_max_err_vs_fp32": e_me.max().item(), "lp_mean_err_vs_fp32": e_me.mean().item(),
                      "hf_bf16_lp_max_err_vs_fp32": e_hf.max().item(), "hf_bf16_lp_mean_err_vs_fp32": e_hf.mean().item(),
                      "loss_fp32": loss32.item(),
                      "grad_cos_mean_vs_fp32": sum(v[0] for v in s_me.values()) / len(s_me),
                      "hf_bf16_grad_cos_mean_vs_fp32": sum(v[0] for v in s_hf.values()) / len(s_hf),
                      "grad_cos_min_vs_fp32": min(v[0] for v in s_me.values()),
                      "hf_bf16_grad_cos_min_vs_fp32": min(v[0] for v in s_hf.values()),
                      "tensors_where_engine_is_further_from_fp32": sorted(
                          ((k, round(s_me[k][0], 5), round(s_hf[k][0], 5)) for k in s_me
                           if (1 - s_me[k][0]) > 1.5 * (1 - s_hf[k][0]) + 2e-3), key=lambda t: t[1])[:8]})
        del hf32
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"hf_parity_{tag}.json"), "w") as f:
        json.dump(stats, f, indent=1)
    print(json.dumps(stats))
    del eng, grads
    torch.cuda.empty_cache()
    return stats


def _common_asserts(s):
    # Two bf16 implementations of the same network (different GEMM tilings / accumulation orders; HF also rounds the
    # logits themselves to bf16: one ulp at |logit| ~ 4 is 0.03) sit within bf16 noise of the fp32 network EACH, so up to
    # twice that apart.  Measured on B200 (gpurun_out/hf_parity_*.json, copies under profiles/): mean 0.021 / 0.031 / 0.048,
    # max 0.059 / 0.196 / 0.187 at cfg1 / cfg2 / cfg3 dims; gradient cosines >= 0.9845 for all 729 / 730 tensors.
    assert s["lp_max_err_vs_hf_bf16"] < 0.3 and s["lp_mean_err_vs_hf_bf16"] < 0.07, s
    assert abs(s["loss"] - s["loss_hf_bf16"]) < 5e-3 * max(1.0, abs(s["loss_hf_bf16"])) + 2e-3, s
    assert s["grad_cos_weighted_vs_hf"] > 0.99 and s["grad_cos_mean_vs_hf"] > 0.99, s
    assert s["grad_cos_min_vs_hf"] > 0.97, s["worst_tensors_vs_hf"]
    lo, hi = s["grad_norm_ratio_range_vs_hf"]
    assert 0.93 < lo and hi < 1.07, s["worst_tensors_vs_hf"]


def _not_further_from_fp32_than_hf(s, grads=True):
    """The yardstick that licenses any tolerance above SURVEY 8(c)'s 2e-2 / 3e-3: against the fp32 network this engine's
    error is at most 1.2 x the error of HF's own bf16 path (measured: 1.09 x mean / 0.75 x max at cfg1)."""
    assert s["lp_mean_err_vs_fp32"] <= 1.2 * s["hf_bf16_lp_mean_err_vs_fp32"] + 1e-3, s
    assert s["lp_max_err_vs_fp32"] <= 1.2 * s["hf_bf16_lp_max_err_vs_fp32"] + 1e-2, s
    if grads:
        assert (1 - s["grad_cos_mean_vs_fp32"]) <= 1.2 * (1 - s["hf_bf16_grad_cos_mean_vs_fp32"]) + 1e-3, s
        assert (1 - s["grad_cos_min_vs_fp32"]) <= 1.5 * (1 - s["hf_bf16_grad_cos_min_vs_fp32"]) + 2e-3, s
        assert abs(s["loss"] - s["loss_fp32"]) <= 1.2 * abs(s["loss_hf_bf16"] - s["loss_fp32"]) + 1e-3, s


def test_cfg1_dims_vs_hf_bf16_and_fp32():
    """cfg1 (BASELINE configs[0]): Qwen2-VL-2B, 2 frames 224^2 (grid 1x16x16), P = 128, G = 2, C = 16."""
    s = _run("2b", 2, 224, 64, 2, 16, "cfg1", fp32_truth=True)
    _common_asserts(s)
    _not_further_from_fp32_than_hf(s)


def test_cfg2_dims_vs_hf_bf16():
    """cfg2 (configs[1]): Qwen2-VL-2B, 8 frames 336^2 (grid 4x24x24), P = 832, G = 4, C = 512."""
    s = _run("2b", 8, 336, 256, 4, 512, "cfg2", fp32_truth=True)
    _common_asserts(s)
    _not_further_from_fp32_than_hf(s)


def test_cfg3_dims_vs_hf_bf16():
    """cfg3 dims (configs[2], the headline): Qwen2-VL-7B, 16 frames 448^2 (grid 8x32x32), P = 2304; G = 2 rows of C = 64
    (HF with gradient checkpointing) -- every one of the 7B model's parameter gradients against HF's."""
    s = _run("7b", 16, 448, 256, 2, 64, "cfg3", fp32_truth="forward", ckpt=True)
    _common_asserts(s)
    _not_further_from_fp32_than_hf(s, grads=False)

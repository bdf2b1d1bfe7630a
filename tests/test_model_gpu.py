"""End-to-end parity of the CUDA path (through the C ABI) against the oracle and the HF golden fixture, at the
tiny dims the oracle finishes in seconds.  Tolerances: bf16 kernels vs fp32 oracle
<= 2e-2 abs on log-probs (mean <= 3e-3); gradients cosine >= 0.999 (>= 0.99 for the smallest tensors)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _setup(layers=2, v_depth=2):
    from oracle import qwen2vl_ref as R
    from spacer_b200 import config
    from spacer_b200.model import Qwen2VLB200
    d_or = R.dims_tiny(layers, v_depth)
    d = config.tiny(layers, v_depth)
    w = R.init_weights(d_or, seed=0)
    m = Qwen2VLB200(d)
    m.load_state_dict(w)
    wb = {k: v.bfloat16().float() for k, v in w.items()}   # oracle sees the same bf16-rounded weights
    return R, d_or, d, m, w, wb


def test_state_dict_roundtrip():
    R, d_or, d, m, w, wb = _setup()
    sd = m.state_dict()
    assert set(sd) == set(w)
    for k in w:
        assert torch.equal(sd[k].cpu(), w[k].bfloat16()), k


def test_vit_forward_matches_oracle():
    R, d_or, d, m, w, wb = _setup()
    grid = torch.tensor([[2, 8, 12], [1, 4, 4]])
    n_p = int((grid[:, 0] * grid[:, 1] * grid[:, 2]).sum())
    pix = torch.randn(n_p, d.patch_dim, generator=torch.Generator().manual_seed(3))
    ref = R.vit_forward(wb, pix.bfloat16().float(), grid, d_or)
    out = m.vit_forward(pix.cuda(), grid).float().cpu()
    err = (out - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item() + 2e-3, err


def _case(d_or):
    from oracle.make_golden import tiny_case
    return tiny_case(d_or)


def test_logps_match_oracle_and_hf_golden():
    """Prefix-shared packed scoring == G independent causal sequences (the reference's layout)."""
    from spacer_b200.model import pack_prompt_completions
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"]
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d_or)
    logits = R.model_logits(wb, ids, case["pixel_values"].bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    ref = R.per_token_logps(logits.bfloat16().float(), ids)[:, P - 1:]
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], grid, d, m.device)
    assert torch.equal(batch.pos[:, :P].cpu().long(), pos[:, 0, :P])
    lp = m.per_token_logps(batch, case["pixel_values"].cuda(), grid).cpu()
    diff = (lp - ref).abs()
    assert diff.max().item() < 2e-2 and diff.mean().item() < 3e-3, (diff.max().item(), diff.mean().item())
    gold = torch.load(os.path.join(GOLD, "tiny_model.pt"), weights_only=False)
    dg = (lp - gold["logps"]).abs()   # HF fp32 weights vs our bf16 weights: a little looser
    assert dg.max().item() < 4e-2 and dg.mean().item() < 6e-3, (dg.max().item(), dg.mean().item())


def test_grpo_forward_backward_matches_oracle():
    from oracle import grpo_ref as GR
    from spacer_b200.model import GradStore, pack_prompt_completions
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    gold = torch.load(os.path.join(GOLD, "tiny_model.pt"), weights_only=False)
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"]
    for v in wb.values():
        v.requires_grad_()
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d_or)
    logits = R.model_logits(wb, ids, case["pixel_values"].bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    lp_ref = R.per_token_logps(logits, ids)[:, P - 1:]
    mask = GR.completion_mask(case["completion_ids"], d_or.eos_id)
    adv, _ = GR.advantages(case["rewards"], G)
    ref_lp = gold["ref_logps"]
    loss_ref, kl_ref = GR.grpo_loss(lp_ref, ref_lp, adv, mask, 0.04)
    loss_ref.backward()
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], grid, d, m.device)
    grads = GradStore(m.params)
    out = m.grpo_forward_backward(batch, case["pixel_values"].cuda(), grid, ref_lp.cuda(), adv.cuda(), 0.04, grads)
    torch.cuda.synchronize()
    assert torch.equal(out["mask"].cpu(), mask)
    assert abs(out["loss"].item() - loss_ref.item()) < 2e-3 * max(1.0, abs(loss_ref.item())) + 5e-4
    assert abs(out["mean_kl"].item() - kl_ref.item()) < 0.1 * abs(kl_ref.item()) + 1e-3
    # gradients, HF names
    got = dict(m.params.hf_items({n: grads[n] for n in m.params.index}))
    worst = {}
    for k, gref in ((k, v.grad) for k, v in wb.items()):
        g = got[k].float().cpu().reshape(gref.shape)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        rel = (g.norm() / (gref.norm() + 1e-30)).item()
        worst[k] = (cos, rel)
        floor = 0.99 if gref.numel() < 4096 else 0.995
        assert cos > floor, f"{k}: cosine {cos:.5f} (norm ratio {rel:.3f})"
        assert 0.9 < rel < 1.1, f"{k}: norm ratio {rel:.3f}"
    mean_cos = sum(c for c, _ in worst.values()) / len(worst)
    assert mean_cos > 0.999, mean_cos


def test_activation_recompute_is_bit_identical():
    """gradient_checkpointing_enable() (HF Trainer under --gradient_checkpointing true, run_SpaceR_SG_RLVR.sh:27): the
    gate|up projection is recomputed in the backward instead of saved -- same loss, bit-identical gradients, with and
    without the rollout's prefix reuse."""
    from spacer_b200.model import GradStore, pack_prompt_completions
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, pix = case["grid_thw"], case["pixel_values"].cuda()
    G = case["input_ids"].shape[0]
    adv = torch.tensor([0.7, -1.1, 0.2, 0.4]).cuda()
    ref_lp = torch.load(os.path.join(GOLD, "tiny_model.pt"), weights_only=False)["ref_logps"].cuda()
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], grid, d, m.device)
    res = {}
    for mode in ("keep", "recompute", "recompute+prefix_reuse"):
        m.gradient_checkpointing_disable()
        if mode != "keep":
            m.gradient_checkpointing_enable()
        vc = None
        if mode.endswith("prefix_reuse"):
            m.generate(case["prompt_ids"], pix, grid, max_new_tokens=2, num_return_sequences=G, seed=1, min_new_tokens=2,
                       keep_vit_tape=True)
            vc = m.vit_cache
            assert vc is not None and vc["llm_tape"]["layers"][0]["gu"] is None
        grads = GradStore(m.params)
        out = m.grpo_forward_backward(batch, pix, grid, ref_lp, adv, 0.04, grads, vit_cache=vc)
        torch.cuda.synchronize()
        res[mode] = (out["loss"].item(), grads.mat.clone(), grads.vec.clone())
    for mode in ("recompute", "recompute+prefix_reuse"):
        assert res[mode][0] == res["keep"][0]
        assert torch.equal(res[mode][1], res["keep"][1]) and torch.equal(res[mode][2], res["keep"][2]), mode
    m.gradient_checkpointing_disable()


def test_generate_semantics_and_decode_parity():
    """generate(): prompt echoed, C' columns, EOS/pad handling, seed determinism; the logits of the LAST decode
    step (decode kernels + KV caches) match the oracle's full forward on the generated ids."""
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid = case["grid_thw"]
    prompt = case["prompt_ids"]
    P = prompt.shape[1]
    G, C = 5, 9
    pix = case["pixel_values"].cuda()
    out = m.generate(prompt, pix, grid, max_new_tokens=C, num_return_sequences=G, top_p=0.95, seed=3, min_new_tokens=C)
    assert out.shape == (G, P + C) and out.dtype == torch.int64
    assert torch.equal(out[:, :P].cpu(), prompt.expand(G, -1))
    assert (out[:, P:] != d.eos_id).all()
    st = m._last_decode_state
    logits_last = st["logits"][0, :G].float().cpu()     # produced token C-1 from position P+C-2
    ids = out.cpu()
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d_or)
    ref = R.model_logits(wb, ids, case["pixel_values"].bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    ref_last = ref[:, P + C - 2]
    err = (logits_last - ref_last).abs().max().item()
    assert err < 3e-2 * ref_last.abs().max().item() + 5e-3, err
    # every sampled token lies in the reference warper's kept set for its position (teacher-forced oracle logits)
    for t in range(C):
        kept = R.top_p_filter(ref[:, P - 1 + t].bfloat16().float(), 0.95)
        chosen = kept.gather(1, ids[:, P + t:P + t + 1])
        vmin = torch.where(torch.isinf(kept), torch.full_like(kept, 1e9), kept).min(1).values
        assert (ref[:, P - 1 + t].bfloat16().float().gather(1, ids[:, P + t:P + t + 1])[:, 0] >= vmin - 2e-2).all()
    out2 = m.generate(prompt, pix, grid, max_new_tokens=C, num_return_sequences=G, top_p=0.95, seed=3, min_new_tokens=C)
    assert torch.equal(out, out2)
    out3 = m.generate(prompt, pix, grid, max_new_tokens=C, num_return_sequences=G, top_p=0.95, seed=4, min_new_tokens=C)
    assert not torch.equal(out, out3)
    eager = m.generate(prompt, pix, grid, max_new_tokens=C, num_return_sequences=G, top_p=0.95, seed=3, min_new_tokens=C,
                       use_graph=False)
    assert torch.equal(out, eager)
    # EOS allowed: rows stop independently, tail is pad, width = longest completion
    free = m.generate(prompt, pix, grid, max_new_tokens=40, num_return_sequences=8, top_p=0.95, seed=5)
    comp = free[:, P:]
    for r in range(comp.shape[0]):
        row = comp[r].tolist()
        if d.eos_id in row:
            k = row.index(d.eos_id)
            assert all(x == d.pad_id for x in row[k + 1:])
    # two groups decoded together (T-GRPO): second group conditions on the other video
    pix2 = torch.flip(pix, dims=[0]).contiguous()
    a, b = m.generate(prompt, pix, grid, max_new_tokens=C, num_return_sequences=4, top_p=0.95, seed=3,
                      min_new_tokens=C, pixel_values_videos_2=pix2, num_return_sequences_2=2)
    assert a.shape == (4, P + C) and b.shape == (2, P + C)
    assert torch.equal(a, out[:4])   # rows of group 1 are unaffected by the presence of group 2
    st = m._last_decode_state
    ids_b = b.cpu()
    pos_b = R.rope_index_classic(ids_b, grid.repeat(2, 1), d_or)
    ref_b = R.model_logits(wb, ids_b, pix2.cpu().bfloat16().float().repeat(2, 1), grid.repeat(2, 1), pos_b, d_or)
    err = (st["logits"][0, 4:6].float().cpu() - ref_b[:, P + C - 2]).abs().max().item()
    assert err < 3e-2 * ref_b.abs().max().item() + 5e-3, err


def test_l2_plan_of_the_decode_step_does_not_change_the_tokens():
    """The L2 plan of the decode step (evict_first weight streams, paced evict_last prefetch of gate|up row subsets on a side
    stream / parallel graph branch) only moves cache lines: rollouts are bit-identical with it on, off, and with any of its
    variants (bulk / paced requests, released by different kernels), graph or eager."""
    from spacer_b200 import ops
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, prompt, pix = case["grid_thw"], case["prompt_ids"], case["pixel_values"].cuda()
    lib = ops._lib.load()
    kw = dict(max_new_tokens=7, num_return_sequences=4, top_p=0.95, seed=9, min_new_tokens=7)
    saved = (m.PF_GU_ROWS, m.PF_AFTER, m.PF_CTAS, m.PF_PACE_NS)
    try:
        ref = m.generate(prompt, pix, grid, **kw)
        for hints, rows, after, ctas, pace, graph in ((0, 0, "qkv", 0, 750, True), (1, 0, "qkv", 0, 750, True),
                                                      (1, 24, "qkv_post", 0, 0, True), (1, 32, "combine", 16, -1, True),
                                                      (1, 24, "qkv", 0, 750, False)):
            lib.sb_set_dec_l2_hints(hints)
            m.PF_GU_ROWS, m.PF_AFTER, m.PF_CTAS, m.PF_PACE_NS = rows, after, ctas, pace
            m._dec = None                                   # new decode state: the captured graphs belong to the old plan
            out = m.generate(prompt, pix, grid, use_graph=graph, **kw)
            assert torch.equal(out, ref), (hints, rows, after, ctas, pace, graph)
    finally:
        lib.sb_set_dec_l2_hints(1)
        m.PF_GU_ROWS, m.PF_AFTER, m.PF_CTAS, m.PF_PACE_NS = saved
        m._dec = None


def test_trainer_step_runs_and_updates():
    from spacer_b200 import rewards as RW
    from spacer_b200.model import Qwen2VLB200
    from spacer_b200.trainer import GRPOConfig, SGRLVRTrainerB200
    R, d_or, d, m, w, wb = _setup()
    ref = Qwen2VLB200(d)
    ref.load_state_dict(w)
    case = _case(d_or)
    RW.set_map_data({"scene0000_00": {"video_id": "scene0000_00",
                                      "cognitive_map": {"table": [[0, 3], [5, 7]], "chair": [[9, 3]]}}})
    texts = ["<think>x</think><answer>B</answer>", "nonsense", "<think>y</think><map>{'table':[[1,3]]}</map><answer>B</answer>",
             "<think>z</think><answer>C</answer>"]

    def decode(ids):
        return [texts[int(r[0]) % len(texts)] for r in ids.tolist()]

    cfg = GRPOConfig(num_generations=4, max_completion_length=8, temporal=True, len_control=True, learning_rate=1e-3,
                     max_steps=10)
    tr = SGRLVRTrainerB200(m, ref, [RW.accuracy_reward, RW.format_reward], cfg, decode)
    ex = dict(input_ids=case["prompt_ids"], pixel_values_videos=case["pixel_values"].cuda(), video_grid_thw=case["grid_thw"],
              solution="<answer>B</answer>", problem_type="multiple choice", path="x/scene0000_00.mp4", prompt="p")
    before = m.params.mat.clone()
    mt = tr.training_step(ex)
    torch.cuda.synchronize()
    for k in ["completion_length", "rewards/accuracy_reward", "rewards/format_reward", "all_wrong", "all_correct",
              "temporal_rewards", "reward", "reward_std", "kl", "loss"]:
        assert k in mt and mt[k] == mt[k], k
    assert not torch.equal(before, m.params.mat)
    assert torch.isfinite(m.params.mat.float()).all()
    assert abs(mt["kl"]) < 1e-3   # policy == reference at step 0
    mt2 = tr.training_step(ex)
    assert mt2["kl"] >= 0.0
    assert set(tr.log()) >= {"reward", "kl", "loss"}


def test_trainer_step_metrics_match_oracle_on_the_rollout():
    """The metrics training_step logs (TRN:650-683) against oracle.grpo_ref on the SAME sampled completions: rewards,
    bonuses, advantages and metrics recomputed by the oracle from the completion ids the step produced."""
    from oracle import grpo_ref as GR
    from spacer_b200 import rewards as RW
    from spacer_b200.model import Qwen2VLB200
    from spacer_b200.trainer import GRPOConfig, SGRLVRTrainerB200
    R, d_or, d, m, w, wb = _setup()
    ref = Qwen2VLB200(d)
    ref.load_state_dict(w)
    case = _case(d_or)
    RW.set_map_data({})
    texts = ["<think>x</think><answer>B</answer>", "nonsense", "<think>z</think><answer>C</answer>"]
    seen = {}

    def decode(ids):
        seen[ids.shape[0]] = ids.clone()
        return [texts[int(r[0]) % len(texts)] for r in ids.tolist()]
    G, C = 4, 10
    cfg = GRPOConfig(num_generations=G, max_completion_length=C, temporal=True, len_control=True, learning_rate=1e-4)
    tr = SGRLVRTrainerB200(m, ref, [RW.accuracy_reward, RW.format_reward], cfg, decode)
    ex = dict(input_ids=case["prompt_ids"], pixel_values_videos=case["pixel_values"].cuda(), video_grid_thw=case["grid_thw"],
              solution="<answer>B</answer>", problem_type="multiple choice", path="x/scene0000_00.mp4", prompt="p")
    mt = tr.training_step(ex)
    comp, shuf = seen[G].cpu(), seen[G // 2].cpu()

    def rpf_of(ids):
        t = [texts[int(r[0]) % len(texts)] for r in ids.tolist()]
        acc = [1.0 if "<answer>B</answer>" in x else 0.0 for x in t]
        fmt = [1.0 if x.startswith("<think>") else 0.0 for x in t]
        return torch.tensor([acc, fmt]).T.contiguous()
    rpf, srpf = rpf_of(comp), rpf_of(shuf)
    mask = GR.completion_mask(comp, d_or.eos_id)
    summed, temporal = GR.temporal_bonus(rpf.clone(), srpf, True)
    rewards = GR.length_bonus(summed.sum(1), rpf, mask, True)
    adv, std = GR.advantages(rewards, G)
    want = GR.step_metrics(mask, rpf, rewards, std, temporal, mt["kl"], G, ["accuracy_reward", "format_reward"])
    for k, v in want.items():
        assert abs(mt[k] - v) < 1e-5, (k, mt[k], v)
    assert mt["generated_tokens"] == comp.numel() + shuf.numel()


def test_text_only_step_after_a_video_step_does_not_reuse_vision_gradients():
    """the vision-tower gradients of the previous step must not be applied again when a step has no visual
    input (GradStore.zero_range), and the optimizer resynchronises its fp32 masters when weights are loaded behind its back."""
    from spacer_b200.model import GradStore, pack_prompt_completions
    from spacer_b200.trainer import AdamW, GRPOConfig
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grads = GradStore(m.params)
    adv = torch.tensor([0.7, -1.1, 0.2, 0.4]).cuda()
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], case["grid_thw"], d, m.device)
    m.grpo_forward_backward(batch, case["pixel_values"].cuda(), case["grid_thw"], None, adv, 0.0, grads)
    lo, hi = grads.mat_range("v.")
    assert float(grads.mat[lo:hi].float().abs().sum()) > 0
    text_prompt = case["prompt_ids"][:, :5]                       # no placeholder tokens
    tb = pack_prompt_completions(text_prompt, case["completion_ids"], None, d, m.device)
    m.grpo_forward_backward(tb, None, None, None, adv, 0.0, grads)
    torch.cuda.synchronize()
    assert float(grads.mat[lo:hi].float().abs().sum()) == 0.0
    # master resync
    opt = AdamW(m.params, GRPOConfig(learning_rate=1e-3))
    w2 = R.init_weights(d_or, seed=9)
    m.load_state_dict(w2)                                           # after the optimizer captured its masters
    grads.mat.zero_(); grads.vec.zero_()
    opt.step(grads)                                                 # zero gradients: weights must stay the NEW ones
    torch.cuda.synchronize()
    got = m.state_dict()["lm_head.weight"].float().cpu()
    want = w2["lm_head.weight"].bfloat16().float()
    assert (got - want).abs().max().item() < 2e-3 * want.abs().max().item() + 1e-6    # only weight decay moved them


def test_grouped_rollout_widths_are_independent():
    """the main and the frame-shuffled group are two generate() calls in the reference -- each is cut to ITS
    longest completion (generation/utils.py stops when all rows of the call are done)."""
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    pix = case["pixel_values"].cuda()
    widths = set()
    for seed in range(6):
        a, b = m.generate(case["prompt_ids"], pix, case["grid_thw"], max_new_tokens=48, num_return_sequences=4, top_p=0.95,
                          seed=seed, pixel_values_videos_2=pix.flip(0).contiguous(), num_return_sequences_2=2)
        P = case["prompt_ids"].shape[1]
        for out in (a, b):
            comp = out[:, P:]
            is_eos = comp == d.eos_id
            if bool(is_eos.any(1).all()):      # every row ended: the width is exactly the longest completion of THIS group
                first = is_eos.int().argmax(1)
                assert comp.shape[1] == int(first.max()) + 1
        widths.add((a.shape[1], b.shape[1]))
    assert len(widths) >= 1


# ------------------------------------------------------------------------------------------------
# Qwen2.5-VL: windowed RMSNorm/SwiGLU vision tower + temporal M-RoPE spacing on the CUDA path
# ------------------------------------------------------------------------------------------------
def _setup25():
    from oracle import qwen25vl_ref as R25
    from oracle.make_golden import tiny_case25
    from spacer_b200 import config
    from spacer_b200.model import Qwen2VLB200
    d_or, d = R25.dims25_tiny(), config.tiny25()
    w = R25.init_weights(d_or, seed=0)
    m = Qwen2VLB200(d, "cuda:0").load_state_dict(w)
    return R25, d_or, d, w, m, tiny_case25(d_or)


def test_qwen25_state_dict_roundtrip_and_host_logic():
    from spacer_b200.model import rope_index, window_plan
    R25, d_or, d, w, m, case = _setup25()
    sd = m.state_dict()
    assert set(sd) == set(w)
    for k in w:
        assert torch.equal(sd[k].cpu().float(), w[k].bfloat16().float()), k
    ids = case["prompt_ids"]
    for sec in (None, [1.5], [2.0]):
        pos, _ = rope_index(ids, case["grid_thw"], d, "classic", sec)
        assert torch.equal(pos, R25.rope_index_classic(ids, case["grid_thw"], d_or, sec)[:, 0])
        pos, _ = rope_index(ids, case["grid_thw"], d, "hf55", sec)
        assert torch.equal(pos, R25.rope_index_hf55(ids, case["grid_thw"], d_or, sec)[:, 0])
    plan = window_plan(case["grid_thw"], d, "cuda:0")
    widx, cu = R25.window_index(case["grid_thw"], d_or)
    assert torch.equal(plan["widx"].cpu().long(), widx) and plan["cu_window"] == cu.tolist()


def test_qwen25_vit_and_logps_vs_oracle_and_hf_golden():
    from oracle import qwen2vl_ref as R
    from spacer_b200.model import pack_prompt_completions
    R25, d_or, d, w, m, case = _setup25()
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_model25.pt"), weights_only=False)
    wb = {k: v.bfloat16().float() for k, v in w.items()}
    grid, pix = case["grid_thw"], case["pixel_values"]
    ve = m.vit_forward(pix.cuda(), grid).float().cpu()
    ve_or = R25.vit_forward(wb, pix.bfloat16().float(), grid, d_or)
    assert (ve - ve_or).abs().max() < 3e-2 * ve_or.abs().max()
    assert (ve - gold["vision_embeds"].float()).abs().max() < 4e-2 * ve_or.abs().max()
    sec = [gold["second_per_grid_ts"]]
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], grid, d, m.device, second_per_grid_ts=sec)
    lp = m.per_token_logps(batch, pix.cuda(), grid).cpu()
    err = (lp - gold["logps"]).abs()
    assert err.max() < 2e-2 and err.mean() < 3e-3, (err.max().item(), err.mean().item())


def test_qwen25_gradients_vs_oracle():
    from oracle import grpo_ref as GR
    from oracle import qwen2vl_ref as R
    from spacer_b200.model import GradStore, pack_prompt_completions
    R25, d_or, d, w, m, case = _setup25()
    wb = {k: v.bfloat16().float().requires_grad_() for k, v in w.items()}
    grid, pix = case["grid_thw"], case["pixel_values"]
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    sec = [1.5]
    pos = R25.rope_index_classic(ids, grid.repeat(G, 1), d_or, sec * G)
    logits = R25.model_logits(wb, ids, pix.bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    ref_lp = lp.detach() + 0.05
    adv, _ = GR.advantages(case["rewards"][:G], G)
    mask = GR.completion_mask(case["completion_ids"], d_or.eos_id)
    loss, _ = GR.grpo_loss(lp, ref_lp, adv, mask, 0.04)
    loss.backward()
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], grid, d, m.device, second_per_grid_ts=sec)
    grads = GradStore(m.params)
    out = m.grpo_forward_backward(batch, pix.cuda(), grid, ref_lp.cuda(), adv.cuda(), 0.04, grads)
    assert abs(float(out["loss"]) - float(loss)) < 5e-4 + 0.05 * abs(float(loss))
    gsd = dict(m.params.hf_items({n: grads[n] for n in m.params.index}))
    checked = 0
    for k, t in wb.items():
        if not k.startswith("model.visual.") or t.grad is None:
            continue
        a, b = gsd[k].float().cpu().flatten(), t.grad.flatten()
        if b.norm() < 1e-9:
            continue
        cos = torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)
        assert cos > 0.99, (k, float(cos))
        assert 0.85 < float(a.norm() / b.norm()) < 1.15, (k, float(a.norm() / b.norm()))
        checked += 1
    assert checked >= 20


def test_greedy_generation_for_eval():
    """Evaluation reuse of the rollout engine: do_sample=False / temperature 0.01 -> argmax with the
    lowest index on ties, independent of the seed, graph == eager, EOS stops a row."""
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, prompt = case["grid_thw"], case["prompt_ids"]
    P = prompt.shape[1]
    pix = case["pixel_values"].cuda()
    a = m.generate(prompt, pix, grid, max_new_tokens=8, num_return_sequences=1, do_sample=False, seed=1, min_new_tokens=8)
    st = m._last_decode_state
    last_logits = st["logits"][0, :1].float().bfloat16().float()
    assert int(a[0, -1]) == int(last_logits[0].argmax())       # torch.argmax returns the first maximum
    b = m.generate(prompt, pix, grid, max_new_tokens=8, num_return_sequences=1, temperature=0.01, seed=99, min_new_tokens=8)
    c = m.generate(prompt, pix, grid, max_new_tokens=8, num_return_sequences=1, do_sample=False, seed=5, min_new_tokens=8,
                   use_graph=False)
    assert torch.equal(a, b) and torch.equal(a, c)
    rows = m.generate(prompt, pix, grid, max_new_tokens=8, num_return_sequences=3, do_sample=False, min_new_tokens=8)
    assert torch.equal(rows[0], rows[1]) and torch.equal(rows[0], rows[2]) and torch.equal(rows[:1], a)
    free = m.generate(prompt, pix, grid, max_new_tokens=30, num_return_sequences=1, do_sample=False)
    assert free.shape[1] <= P + 30 and torch.equal(free[:, :P].cpu(), prompt)


def test_image_prompt_matches_oracle():
    """`pixel_values` / `image_grid_thw` prompts (image placeholder tokens, grid t = 1): last-step decode logits vs the
    oracle's teacher-forced forward, as for video prompts."""
    R, d_or, d, m, w, wb = _setup()
    g = torch.Generator().manual_seed(41)
    grid = torch.tensor([[1, 8, 12]])
    n_p = 96
    pix = torch.randn(n_p, d_or.patch_dim, generator=g)
    text = torch.randint(10, 2000, (9,), generator=g)
    prompt = torch.cat([text[:4], torch.tensor([d_or.vision_start_id]), torch.full((n_p // 4,), d_or.image_token_id),
                        torch.tensor([d_or.vision_end_id]), text[4:]])[None]
    P, G, C = prompt.shape[1], 3, 5
    out = m.generate(prompt, pixel_values=pix.cuda(), image_grid_thw=grid, max_new_tokens=C, num_return_sequences=G,
                     top_p=0.95, seed=2, min_new_tokens=C)
    assert out.shape == (G, P + C) and torch.equal(out[:, :P].cpu(), prompt.expand(G, -1))
    ids = out.cpu()
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d_or)
    ref = R.model_logits(wb, ids, pix.bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    last = m._last_decode_state["logits"][0, :G].float().cpu()
    err = (last - ref[:, P + C - 2]).abs().max().item()
    assert err < 3e-2 * ref[:, P + C - 2].abs().max().item() + 5e-3, err


def test_sft_loss_and_gradients_vs_oracle():
    """SFT stage reuse (open_r1/sft.py:147-182): token cross-entropy with pad / visual labels masked
    (-100), HF causal-LM shift, vs torch.nn.functional.cross_entropy on the oracle's logits; every gradient by cosine."""
    from spacer_b200.model import GradStore
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, pix = case["grid_thw"], case["pixel_values"]
    ids = case["input_ids"][0]                                  # one full sequence: prompt + completion
    ids = torch.cat([ids, torch.full((3,), d_or.pad_id)])       # padded tail, as the processor would produce
    labels = ids.clone()
    labels[labels == d_or.pad_id] = -100
    for vt in (d_or.vision_start_id, d_or.vision_end_id, d_or.video_token_id):
        labels[labels == vt] = -100
    wg = {k: v.clone().requires_grad_() for k, v in wb.items()}
    pos = R.rope_index_classic(ids[None], grid, d_or)
    logits = R.model_logits(wg, ids[None], pix.bfloat16().float(), grid, pos, d_or)[0]
    loss_ref = torch.nn.functional.cross_entropy(logits[:-1], labels[1:], ignore_index=-100)
    loss_ref.backward()
    grads = GradStore(m.params)
    out = m.sft_forward_backward(ids, labels, pix.cuda(), grid, grads)
    assert out["n_tokens"] == int((labels[1:] != -100).sum())
    assert abs(float(out["loss"]) - float(loss_ref)) < 2e-2, (float(out["loss"]), float(loss_ref))
    got = dict(m.params.hf_items({n: grads[n] for n in m.params.index}))
    for k, t in wg.items():
        gref = t.grad
        if gref is None or gref.norm() < 1e-9:
            continue
        g = got[k].float().cpu().reshape(gref.shape)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        rel = (g.norm() / gref.norm()).item()
        assert cos > 0.99, f"{k}: cosine {cos:.5f}"
        assert 0.9 < rel < 1.1, f"{k}: norm ratio {rel:.3f}"


def test_sft_overfits_one_sequence():
    """End-to-end sanity of backward + clip + AdamW: repeated SFT steps on one sequence drive its loss down."""
    from spacer_b200.model import GradStore
    from spacer_b200.trainer import AdamW, GRPOConfig
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, pix = case["grid_thw"], case["pixel_values"].cuda()
    ids = case["input_ids"][0]
    labels = ids.clone()
    for vt in (d_or.vision_start_id, d_or.vision_end_id, d_or.video_token_id):
        labels[labels == vt] = -100
    opt = AdamW(m.params, GRPOConfig(learning_rate=2e-3, weight_decay=0.0, lr_scheduler_type="constant", max_steps=100))
    grads = GradStore(m.params)
    losses = []
    for _ in range(8):
        out = m.sft_forward_backward(ids, labels, pix, grid, grads)
        losses.append(float(out["loss"]))
        opt.step(grads)
    assert all(l == l for l in losses), losses
    assert losses[-1] < 0.7 * losses[0], losses


def test_grpo_backward_is_bit_reproducible():
    """No floating-point atomics on the training path: two runs of the same forward/backward give bit-identical
    gradient arenas (what makes data-parallel ranks and repeated runs agree exactly)."""
    from oracle import grpo_ref as GR
    from spacer_b200.model import GradStore, pack_prompt_completions
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    comp = case["completion_ids"].clone()
    comp[:, 0] = comp[0, 0]                      # repeated tokens: the embedding rows that accumulate
    comp[2, 3:6] = comp[0, 1]
    batch = pack_prompt_completions(case["prompt_ids"], comp, case["grid_thw"], d, m.device)
    G = comp.shape[0]
    adv, _ = GR.advantages(case["rewards"][:G], G)
    ref_lp = torch.full(comp.shape, -7.0)
    outs = []
    for _ in range(2):
        grads = GradStore(m.params)
        m.grpo_forward_backward(batch, case["pixel_values"].cuda(), case["grid_thw"], ref_lp.cuda(), adv.cuda(), 0.04, grads)
        outs.append((grads.mat.clone(), grads.vec.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float(outs[0][0].float().abs().sum()) > 0


@pytest.mark.parametrize("family", ["qwen2_vl", "qwen2_5_vl"])
def test_from_pretrained_reads_hf_checkpoints_and_roundtrips(family, tmp_path):
    """SURVEY 8(b) model construction: a directory written by the real HF class's save_pretrained loads into the engine
    (config.json parsed, variant picked, every parameter equal), and our save_pretrained is read back by HF."""
    import transformers
    from oracle import make_golden as MG
    from spacer_b200.model import Qwen2VLB200
    if family == "qwen2_vl":
        from oracle import qwen2vl_ref as R
        d_or, w = R.dims_tiny(), None
        w = R.init_weights(d_or, seed=3)
        hf = transformers.Qwen2VLForConditionalGeneration(MG.hf_config(d_or))
    else:
        from oracle import qwen25vl_ref as R25
        d_or = R25.dims25_tiny()
        w = R25.init_weights(d_or, seed=3)
        hf = transformers.Qwen2_5_VLForConditionalGeneration(MG.hf_config25(d_or))
    hf.load_state_dict(w, strict=True)
    hf = hf.to(torch.bfloat16)
    src = tmp_path / "hf_ckpt"
    hf.save_pretrained(src)
    m = Qwen2VLB200.from_pretrained(str(src), attn_implementation="flash_attention_2", torch_dtype=torch.bfloat16, use_cache=False)
    assert m.dims.variant == family and m.config._name_or_path == str(src) and m.warnings_issued == {}
    sd = m.state_dict()
    for k, v in w.items():
        assert torch.equal(sd[k].cpu(), v.bfloat16()), k
    m.gradient_checkpointing_enable()
    dst = tmp_path / "ours"
    m.save_pretrained(dst)
    cls = transformers.Qwen2VLForConditionalGeneration if family == "qwen2_vl" else transformers.Qwen2_5_VLForConditionalGeneration
    back = cls.from_pretrained(dst, torch_dtype=torch.bfloat16)
    bsd = back.state_dict()
    for k, v in w.items():
        assert torch.equal(bsd[k].cpu(), v.bfloat16()), k
    m2 = Qwen2VLB200.from_pretrained(str(dst))
    assert torch.equal(m2.params.mat, m.params.mat) and torch.equal(m2.params.vec, m.params.vec)


def test_get_per_token_logps_dropin_signature():
    """The reference's override point `_get_per_token_logps(model, input_ids, pixel_values_videos=(xG), video_grid_thw=(xG))
    -> [G, L-1]` (SG_RLVR_trainer.py:353-366, 507-528): full-width result vs the oracle, incl. the prompt positions."""
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    ids = case["input_ids"]
    G, L = ids.shape
    grid = case["grid_thw"].repeat(G, 1)
    pix = case["pixel_values"].repeat(G, 1)
    got = m.get_per_token_logps(ids, pixel_values_videos=pix.cuda(), video_grid_thw=grid).cpu()
    assert got.shape == (G, L - 1)
    pos = R.rope_index_classic(ids, grid, d_or)
    logits = R.model_logits(wb, ids, pix.bfloat16().float(), grid, pos, d_or)
    ref = R.per_token_logps(logits.bfloat16().float(), ids)
    err = (got - ref).abs()
    assert err.max() < 3e-2 and err.mean() < 3e-3, (err.max().item(), err.mean().item())
    P = case["prompt_ids"].shape[1]
    from spacer_b200.model import pack_prompt_completions
    batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], case["grid_thw"], d, m.device)
    assert torch.equal(got[:, P - 1:], m.per_token_logps(batch, case["pixel_values"].cuda(), case["grid_thw"]).cpu())


def test_update_reuses_the_rollout_vit_forward():
    """generate(keep_vit_tape=True) hands its vision-tower output and saved activations to the update that follows:
    gradients are bit-identical to recomputing the forward."""
    from oracle import grpo_ref as GR
    from spacer_b200.model import GradStore, pack_prompt_completions
    R, d_or, d, m, w, wb = _setup()
    case = _case(d_or)
    grid, pix = case["grid_thw"], case["pixel_values"].cuda()
    out = m.generate(case["prompt_ids"], pix, grid, max_new_tokens=6, num_return_sequences=4, seed=3, min_new_tokens=6,
                     keep_vit_tape=True)
    cache = m.vit_cache
    assert cache is not None and cache["pixels"] is pix
    P = case["prompt_ids"].shape[1]
    batch = pack_prompt_completions(case["prompt_ids"], out[:, P:].cpu(), grid, d, m.device)
    adv, _ = GR.advantages(torch.tensor([1.0, 0.0, 2.0, 0.5]), 4)
    g1, g2 = GradStore(m.params), GradStore(m.params)
    o1 = m.grpo_forward_backward(batch, pix, grid, None, adv.cuda(), 0.0, g1, vit_cache=cache)
    o2 = m.grpo_forward_backward(batch, pix, grid, None, adv.cuda(), 0.0, g2)
    assert torch.equal(o1["logps"], o2["logps"])
    assert torch.equal(g1.mat, g2.mat) and torch.equal(g1.vec, g2.vec)

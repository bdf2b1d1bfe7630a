"""CPU tests: the oracle against the committed golden vectors (generated from the reference's own code and from
HF transformers 5.5.0 by oracle/make_golden.py), and the host-side reward verifier against the same vectors."""
import json
import math
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def rewards_gold():
    with open(os.path.join(GOLD, "rewards.json")) as f:
        return json.load(f)


def test_extract_map_matches_reference(rewards_gold):
    from spacer_b200 import rewards as RW
    for rec in rewards_gold["extract"]:
        if "raises" in rec:
            with pytest.raises(BaseException) as ei:
                RW.extract_map_data(rec["body"], rec["objects"])
            assert type(ei.value).__name__ == rec["raises"], rec
        else:
            assert RW.extract_map_data(rec["body"], rec["objects"]) == rec["out"], rec


def test_prediction_score_matches_reference(rewards_gold):
    from spacer_b200 import rewards as RW
    for rec in rewards_gold["score"]:
        if "raises" in rec:
            with pytest.raises(BaseException) as ei:
                RW.calculate_prediction_score(rec["response"], rec["solution"], 10)
            assert type(ei.value).__name__ == rec["raises"], rec
        else:
            got = RW.calculate_prediction_score(rec["response"], rec["solution"], 10)
            # the reference sums over a Python set (hash-seed dependent order): equal to 1 ulp-ish
            assert math.isclose(got, rec["out"], rel_tol=0, abs_tol=1e-12), rec


def test_known_answers_from_survey():
    """SURVEY.md 8(c) known-answer table (produced from the reference code)."""
    from spacer_b200 import rewards as RW
    sol = {"table": [[0, 3], [5, 7]], "chair": [[9, 3]], "window": [[6, 5]]}
    keys = list(sol)
    assert RW.calculate_prediction_score(sol, sol, 10) == 1.0
    r = RW.extract_map_data('{"table":[[1,3],[5,5]],"chair":[[9,9]]}', keys)
    assert r == {"table": [[1, 3], [5, 5]], "chair": [[9, 9]]}
    assert math.isclose(RW.calculate_prediction_score(r, sol, 10), 0.5909009742330268, abs_tol=1e-15)
    r = RW.extract_map_data("table: (1, 3), (5,5); Chair at <9,9>; sofa [1,1]", keys)
    assert r == {"table": [[1, 3], [5, 5]], "chair": [[9, 9], [1, 1]]}
    assert math.isclose(RW.calculate_prediction_score(r, sol, 10), 0.5189339828220179, abs_tol=1e-15)
    assert RW.extract_map_data('str{{"table": "[2, 2]"}}', ["table"]) == {"table": [[2, 2]]}
    assert RW.calculate_prediction_score({}, sol, 10) == 0.0
    assert RW.calculate_prediction_score({}, {}, 10) == 1.0
    with pytest.raises(TypeError):
        RW.calculate_prediction_score({"x": []}, {"x": []}, 10)


def test_accuracy_and_format_rewards_match_reference(rewards_gold, capsys):
    from spacer_b200 import rewards as RW
    RW.set_map_data(rewards_gold["map_rows"])
    for rec in rewards_gold["accuracy"]:
        comp = [[{"role": "assistant", "content": rec["content"]}]]
        got = RW.accuracy_reward(comp, [rec["solution"]], [rec["path"]], problem_type=[rec["type"]])
        assert math.isclose(float(got[0]), rec["out"], abs_tol=1e-12), rec
    for rec in rewards_gold["format"]:
        comp = [[{"role": "assistant", "content": rec["content"]}]]
        assert RW.format_reward(comp)[0] == rec["out"], rec
    b = rewards_gold["accuracy_batch"]
    comps = [[{"role": "assistant", "content": c}] for c in b["contents"]]
    got = RW.accuracy_reward(comps, b["solutions"], b["paths"], problem_type=[b["type"]] * len(comps))
    assert [float(x) for x in got] == pytest.approx(b["out"], abs=1e-12)
    assert set(RW.reward_funcs_registry) == {"accuracy", "format"}


# ---------------------------------------------------------------------------------------------
def test_grpo_math_known_answers():
    from oracle import grpo_ref as GR
    ids = torch.tensor([[5, 9, 7, 7], [5, 5, 5, 5], [9, 1, 9, 1]])
    m = GR.completion_mask(ids, eos_id=9)
    assert m.tolist() == [[1, 1, 0, 0], [1, 1, 1, 1], [1, 0, 0, 0]]  # first EOS included
    r = torch.tensor([1.0, 2.0, 3.0, 6.0])
    adv, std = GR.advantages(r, 4)
    s = math.sqrt(((1 - 3) ** 2 + (2 - 3) ** 2 + 0 + (6 - 3) ** 2) / 3)  # unbiased
    assert torch.allclose(adv, (r - 3.0) / (s + 1e-4))
    kl = GR.per_token_kl(torch.tensor([0.0, 20.0, -20.0]), torch.tensor([0.0, 0.0, 0.0]))
    assert torch.allclose(kl, torch.tensor([0.0, math.exp(10) - 11, math.exp(-10) + 9]))
    # temporal bonus: acc mean 0.5 >= 0.8*0.5 -> +0.3 on rows with acc > 0.1
    rpf = torch.tensor([[1.0, 1.0], [0.0, 1.0]])
    t, flag = GR.temporal_bonus(rpf, torch.tensor([[0.5, 0.0], [0.5, 0.0]]), True)
    assert flag == 1.0 and t[:, 0].tolist() == pytest.approx([1.3, 0.0])
    t, flag = GR.temporal_bonus(rpf, torch.tensor([[1.0, 0.0], [1.0, 0.0]]), True)
    assert flag == 0.0 and torch.equal(t, rpf)
    assert GR.temporal_bonus(rpf, None, False)[1] == 0.5
    # length bonus needs >= 2 accurate rows and 320 <= len <= 512
    mask = torch.zeros(3, 600, dtype=torch.int32)
    mask[0, :320] = 1
    mask[1, :513] = 1
    mask[2, :400] = 1
    rew = torch.tensor([1.0, 1.0, 0.0])
    out = GR.length_bonus(rew, torch.tensor([[1.0], [1.0], [0.0]]), mask, True)
    assert out.tolist() == pytest.approx([1.2, 1.0, 0.0])
    out = GR.length_bonus(rew, torch.tensor([[1.0], [0.0], [0.0]]), mask, True)
    assert out.tolist() == pytest.approx([1.0, 1.0, 0.0])


def test_grpo_loss_grad_matches_autograd():
    from oracle import grpo_ref as GR
    g = torch.Generator().manual_seed(0)
    lp = (-torch.rand(4, 9, generator=g) * 5).requires_grad_()
    ref = lp.detach() + torch.randn(4, 9, generator=g) * 6
    adv = torch.randn(4, generator=g)
    mask = (torch.rand(4, 9, generator=g) > 0.3).int()
    mask[:, 0] = 1
    loss, _ = GR.grpo_loss(lp, ref, adv, mask, 0.04)
    loss.backward()
    assert torch.allclose(lp.grad, GR.grpo_loss_grad(lp.detach(), ref, adv, mask, 0.04), atol=1e-6)


@pytest.fixture(scope="module")
def tiny_gold():
    return torch.load(os.path.join(GOLD, "tiny_model.pt"), weights_only=False)


def test_oracle_model_matches_hf_golden(tiny_gold):
    """The plain-torch restatement reproduces HF 5.5.0's vision embeddings, per-token log-probs, GRPO loss
    and gradients on the committed tiny case (fp32 CPU; tolerance 2e-5 abs on log-probs)."""
    from oracle import grpo_ref as GR
    from oracle import qwen2vl_ref as R
    from oracle.make_golden import tiny_case
    d = R.dims_tiny()
    w = R.init_weights(d, seed=0)
    for v in w.values():
        v.requires_grad_()
    case = tiny_case(d)
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"].repeat(G, 1)
    pos = R.rope_index_classic(ids, grid, d)
    assert torch.equal(pos[:, 0].to(torch.int16), tiny_gold["pos_classic"])
    ve = R.vit_forward(w, case["pixel_values"], case["grid_thw"], d)
    assert (ve.detach() - tiny_gold["vision_embeds"].float()).abs().max() < 2e-3  # fixture stored as fp16
    logits = R.model_logits(w, ids, case["pixel_values"].repeat(G, 1), grid, pos, d)
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    assert (lp.detach() - tiny_gold["logps"]).abs().max() < 2e-5
    mask = GR.completion_mask(case["completion_ids"], d.eos_id)
    assert torch.equal(mask, tiny_gold["mask"])
    adv, _ = GR.advantages(case["rewards"], G)
    assert torch.allclose(adv, tiny_gold["advantages"])
    loss, kl = GR.grpo_loss(lp, tiny_gold["ref_logps"], adv, mask, tiny_gold["beta"])
    assert abs(float(loss) - float(tiny_gold["loss"])) < 1e-6
    assert abs(float(kl) - float(tiny_gold["mean_kl"])) < 1e-6
    loss.backward()
    for k, gn in tiny_gold["grad_norms"].items():
        if k == "lm_head.weight" and d.tie:
            continue
        mine = w[k].grad.norm()
        assert abs(float(mine) - float(gn)) <= 1e-4 * float(gn) + 1e-7, k
    for k, smp in tiny_gold["grad_samples"].items():
        gk = w[k].grad.flatten()
        mine = gk[:: max(1, gk.numel() // 64)][:64]
        assert torch.allclose(mine, smp, rtol=1e-3, atol=1e-7), k


def test_top_p_filter_semantics():
    from oracle import qwen2vl_ref as R
    logits = torch.log(torch.tensor([[0.5, 0.3, 0.15, 0.04, 0.01]]))
    out = R.top_p_filter(logits, 0.95)
    # ascending cumsum: .01, .05, .20 ...: tokens with cum <= 0.05 are dropped -> the two smallest
    assert torch.isinf(out[0, 3]) and torch.isinf(out[0, 4]) and not torch.isinf(out[0, :3]).any()
    one = R.top_p_filter(torch.tensor([[0.0, -50.0]]), 0.0001)
    assert not torch.isinf(one[0, 0])  # top-1 always kept


# ---------------------------------------------------------------------------------------------
# Qwen2.5-VL: oracle/qwen25vl_ref.py against the real HF Qwen2_5_VLForConditionalGeneration
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny25_gold():
    return torch.load(os.path.join(GOLD, "tiny_model25.pt"), weights_only=False)


def test_oracle_qwen25_matches_hf_golden(tiny25_gold):
    """Windowed ViT (RMSNorm, SwiGLU with biases, window reorder and its inverse), temporal M-RoPE spacing, log-probs,
    loss and gradients of the plain-torch restatement equal HF 5.5.0's on the committed tiny Qwen2.5-VL case."""
    from oracle import grpo_ref as GR
    from oracle import qwen25vl_ref as R25
    from oracle import qwen2vl_ref as R
    from oracle.make_golden import tiny_case25
    d = R25.dims25_tiny()
    w = R25.init_weights(d, seed=0)
    for v in w.values():
        v.requires_grad_()
    case = tiny_case25(d)
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"].repeat(G, 1)
    sec = [tiny25_gold["second_per_grid_ts"]] * G
    pos = R25.rope_index_classic(ids, grid, d, second_per_grid_ts=sec)
    assert torch.equal(pos[:, 0].to(torch.int16), tiny25_gold["pos_classic"])
    # temporal spacing really is 1.5 s x 2 tokens/s = 3 per temporal patch for the classic formula
    vis = (ids[0] == d.video_token_id).nonzero().flatten()
    assert pos[0, 0, vis[-1]] - pos[0, 0, vis[0]] == 3
    # HF 5.5.0's own default position ids (explicit-input convention, SURVEY 8(c) drift #2)
    assert torch.equal(R25.rope_index_hf55(ids[:1], grid[:1], d, [2.0])[:, 0].to(torch.int16), tiny25_gold["pos_hf55_sec2"])
    assert torch.equal(R25.rope_index_hf55(ids[:1], grid[:1], d)[:, 0].to(torch.int16), tiny25_gold["pos_hf55_none"])
    widx, cu = R25.window_index(case["grid_thw"], d)
    assert sorted(widx.tolist()) == list(range(case["pixel_values"].shape[0] // 4))
    assert cu.tolist() == [0, 64, 96, 160, 192]          # per frame: a full 4x4 window and a ragged 2x4 one
    ve = R25.vit_forward(w, case["pixel_values"], case["grid_thw"], d)
    assert (ve.detach() - tiny25_gold["vision_embeds"].float()).abs().max() < 2e-3   # fixture stored as fp16
    logits = R25.model_logits(w, ids, case["pixel_values"].repeat(G, 1), grid, pos, d)
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    assert (lp.detach() - tiny25_gold["logps"]).abs().max() < 2e-5
    mask = GR.completion_mask(case["completion_ids"], d.eos_id)
    adv, _ = GR.advantages(case["rewards"][:G], G)
    loss, kl = GR.grpo_loss(lp, tiny25_gold["ref_logps"], adv, mask, tiny25_gold["beta"])
    assert abs(float(loss) - float(tiny25_gold["loss"])) < 1e-6
    loss.backward()
    for k, gn in tiny25_gold["grad_norms"].items():
        mine = w[k].grad.norm()
        assert abs(float(mine) - float(gn)) <= 1e-4 * float(gn) + 1e-7, k
    for k, smp in tiny25_gold["grad_samples"].items():
        gk = w[k].grad.flatten()
        assert torch.allclose(gk[:: max(1, gk.numel() // 64)][:64], smp, rtol=1e-3, atol=1e-7), k

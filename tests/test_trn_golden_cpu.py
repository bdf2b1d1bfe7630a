"""tests/golden/trn_compute_loss.pt was produced by the REFERENCE's own unmodified `SGRLVRTrainer.compute_loss`
(oracle/make_trn_golden.py, oracle/trn_harness.py).  Here, without a GPU:
  * the product's host logic (rewards, reward tail, metrics) and the oracle's loss restatement reproduce what the reference
    computed (pins oracle/grpo_ref.py and rows a13-a18 to the reference's code, not to hand-derived numbers);
  * every model call the reference trainer made binds to the signatures of spacer_b200.hf_api (the drop-in contract);
  * when /root/reference is present (build container), the fixture is regenerated and must match the committed one."""
import contextlib
import inspect
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "trn_compute_loss.pt")
SCENARIOS = ("video_short", "video_long", "image")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _product_rewards(gold, s, ids):
    """The product's reward functions on the texts the fixture's decode() assigns to these completions."""
    from spacer_b200 import rewards as RW
    RW.set_map_data(gold["map_rows"])
    texts = [gold["texts"][int(r[0]) % len(gold["texts"])] for r in ids.tolist()]
    n = len(texts)
    comps = [[{"role": "assistant", "content": t}] for t in texts]
    kind = s["kind"]
    path = "synthetic/scene0000_00.mp4" if kind == "video" else "synthetic/scene0000_00.jpg"
    kw = dict(prompts=[None] * n, completions=comps, path=[path] * n, solution=["<answer>B</answer>"] * n,
              problem_type=["multiple choice"] * n, video_path=path)
    with contextlib.redirect_stdout(io.StringIO()):
        acc = RW.accuracy_reward(**kw)
        fmt = RW.format_reward(**kw)
    return torch.tensor([acc, fmt], dtype=torch.float64).T.contiguous()


@pytest.mark.parametrize("name", SCENARIOS)
def test_product_host_logic_reproduces_reference_compute_loss(gold, name):
    from oracle import grpo_ref as GR
    from oracle import qwen2vl_ref as R
    from oracle.make_trn_golden import ref_weights
    from oracle.vision_ref import patchify_ref
    from spacer_b200 import trainer as T
    s = gold[name]
    G, C = s["G"], s["C"]
    d = R.dims_tiny(2, 2)
    # rewards: product functions == what the reference's functions returned inside compute_loss
    by_func = {}
    for r in s["rewards"]:
        by_func.setdefault((r["func"], r["n"]), r["out"])
    rpf = _product_rewards(gold, s, s["completions"])                       # float64: exact equality with the reference
    assert rpf[:, 0].tolist() == by_func[("accuracy_reward", G)] and rpf[:, 1].tolist() == by_func[("format_reward", G)]
    rpf = rpf.float()                                                       # TRN:593 stores them as float32
    shuf_rpf = None
    if s["kind"] == "video":
        shuf_rpf = _product_rewards(gold, s, s["shuffled_completions"])
        assert shuf_rpf[:, 0].tolist() == by_func[("accuracy_reward", G // 2)]
        shuf_rpf = shuf_rpf.float()
    # every dataset column reaches the reward functions (TRN:585-592)
    assert {"data_type", "path", "problem_type", "solution", "video_path"} <= set(s["rewards"][-1]["kwargs"])
    # reward tail + metrics through the product functions training_step calls
    lengths = T.completion_lengths(s["completions"], d.eos_id)
    rewards, adv, std, temporal = T.reward_tail(rpf, shuf_rpf, lengths, G, temporal=True, len_control=True)
    # log-probs from the oracle (fp32), loss from the oracle's restatement
    w = R.init_weights(d, seed=s["weights_seed"])
    rw = ref_weights(w, **s["ref_weights"])
    frames = s["frames"].float()
    pix, grid = patchify_ref(frames)
    grid = torch.tensor([list(grid)])
    ids = torch.cat([s["prompt_ids"].repeat(G, 1), s["completions"]], 1)
    P = s["prompt_ids"].shape[1]
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d)
    with torch.no_grad():
        lp = R.per_token_logps(R.model_logits(w, ids, pix.repeat(G, 1), grid.repeat(G, 1), pos, d), ids)[:, P - 1:]
        rlp = R.per_token_logps(R.model_logits(rw, ids, pix.repeat(G, 1), grid.repeat(G, 1), pos, d), ids)[:, P - 1:]
    mask = GR.completion_mask(s["completions"], d.eos_id)
    loss, mean_kl = GR.grpo_loss(lp, rlp, adv, mask, s["beta"])
    assert abs(float(loss) - s["loss"]) < 1e-6, (float(loss), s["loss"])
    got = T.step_metrics(T.pack_step_stats(lengths, rpf, rewards, std, mean_kl, temporal)[None], G,
                         ["accuracy_reward", "format_reward"], temporal=True)
    assert set(got) == set(s["metrics"])
    for k, v in s["metrics"].items():
        assert abs(got[k] - v) < 1e-5, (k, got[k], v)
    if name == "video_long":        # rows 0 (330 tokens) gets the length bonus, row 2 (319) does not (TRN:620-629)
        assert lengths.tolist() == [330, 100, 319, 340]
        assert abs(float(rewards[0] - rpf[0].sum()) - (T.TEMPORAL_BONUS + T.LEN_BONUS)) < 1e-6
        assert abs(float(rewards[2] - rpf[2].sum()) - T.TEMPORAL_BONUS) < 1e-6


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_trainer_calls_bind_to_hf_api(gold, name):
    """The call log of the reference trainer (keyword names, generation_config fields) against the signatures of the
    drop-in module; generation options resolve to the reference's sampling configuration."""
    from transformers import GenerationConfig
    from spacer_b200 import config
    from spacer_b200.hf_api import Qwen2VLForConditionalGenerationB200 as M
    from spacer_b200.model import Qwen2VLB200
    s = gold[name]
    calls = s["calls"]
    assert [c["call"] for c in calls] == ["generate", "generate", "forward", "forward"]
    assert [c["model"] for c in calls] == ["policy", "policy", "policy", "ref"]
    eng = Qwen2VLB200.__new__(Qwen2VLB200)
    eng.dims, eng.generation_config = config.tiny(2, 2), {}
    for c in calls:
        if c["call"] == "generate":
            gc = GenerationConfig(**{k: v for k, v in c["generation_config"].items() if v is not None})
            inspect.signature(M.generate).bind(None, **{k: None for k in c["kwargs"]}, generation_config=gc)
            sp, mx, mn, G = eng.resolve_generation(gc)
            assert (mx, G) == (c["generation_config"]["max_new_tokens"], c["generation_config"]["num_return_sequences"])
            assert (sp.top_p, sp.top_k, sp.temperature, sp.greedy, sp.pad_id) == (0.95, 50, 1.0, False, 2027)
        else:
            inspect.signature(M.forward).bind(None, None, **{k: None for k in c["kwargs"]})
            assert "attention_mask" not in c["kwargs"]                      # TRN:357: no mask is passed
            G = s["G"]
            assert c["args"][0]["shape"][0] == G
            key = "pixel_values_videos" if s["kind"] == "video" else "pixel_values"
            assert c["kwargs"][key]["shape"][0] % G == 0                    # pixels repeated x G (TRN:507-518)
    assert calls[2]["grad_enabled"] and not calls[2]["inference_mode"]
    assert calls[3]["inference_mode"]                                       # TRN:534
    if s["kind"] == "image":
        assert calls[1]["generation_config"]["max_new_tokens"] == 1         # the dummy call, TRN:481
    else:
        assert calls[1]["generation_config"]["num_return_sequences"] == s["G"] // 2


def test_fixture_regenerates_from_the_reference(gold):
    from oracle import trn_harness as H
    if not H.reference_available():
        pytest.skip("/root/reference is only present in the build container")
    from oracle.make_golden import MAP_ROWS, load_reference
    from oracle.make_trn_golden import run_scenario
    mod = H.load_reference_trainer()
    _, ref_mod = load_reference()
    ref_mod.MAP_DATA = MAP_ROWS
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        s = run_scenario("video_short", mod, ref_mod, verbose=False)
    g = gold["video_short"]
    assert abs(s["loss"] - g["loss"]) < 1e-7 and s["calls"] == g["calls"]
    for k, v in g["metrics"].items():
        assert abs(s["metrics"][k] - v) < 1e-6
    assert torch.equal(s["frame_perm"], g["frame_perm"])

"""Replay, on the B200 engine, of the model calls the REFERENCE's unmodified `SGRLVRTrainer.compute_loss` made when it
produced tests/golden/trn_compute_loss.pt (oracle/make_trn_golden.py): the same generate(**prompt_inputs,
generation_config=...) calls, the same `model(prompt_completion_ids, **prompt_inputs).logits` scoring calls (pixels repeated
x G, reference policy under torch.inference_mode()), the trainer's own log-softmax / gather / loss arithmetic and
`loss.backward()`.  Loss, KL, metrics and the gradients of every parameter must match what the reference computed on the
fp32 oracle model (bf16 tolerance)."""
import contextlib
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "trn_compute_loss.pt")


def _get_per_token_logps_reference_arithmetic(model, input_ids, **kwargs):
    """TRN:353-366 line for line (test infrastructure: the reference source is not available on the GPU box)."""
    logits = model(input_ids, **kwargs).logits
    logits = logits[:, :-1, :]
    input_ids = input_ids[:, 1:]
    per_token_logps = []
    for logits_row, input_ids_row in zip(logits, input_ids):
        log_probs = logits_row.log_softmax(dim=-1)
        per_token_logps.append(torch.gather(log_probs, dim=1, index=input_ids_row.unsqueeze(1)).squeeze(1))
    return torch.stack(per_token_logps)


@pytest.mark.parametrize("name,fast", [("video_short", False), ("video_short", True), ("video_long", True), ("image", False)])
def test_replay_reference_compute_loss_calls(name, fast):
    from transformers import GenerationConfig
    from oracle import grpo_ref as GR
    from oracle import qwen2vl_ref as R
    from oracle.make_trn_golden import ref_weights
    from oracle.vision_ref import patchify_ref
    from spacer_b200 import config, rewards as RW, trainer as T
    from spacer_b200.hf_api import Qwen2VLForConditionalGenerationB200, get_per_token_logps
    gold = torch.load(GOLD, weights_only=False)
    s = gold[name]
    G, C, kind = s["G"], s["C"], s["kind"]
    d_or, d = R.dims_tiny(2, 2), config.tiny(2, 2)
    w = R.init_weights(d_or, seed=s["weights_seed"])
    policy = Qwen2VLForConditionalGenerationB200.from_dims(d, "cuda")
    policy.load_state_dict(w)
    refm = Qwen2VLForConditionalGenerationB200.from_dims(d, "cuda")
    refm.load_state_dict(ref_weights(w, **s["ref_weights"]))
    refm.requires_grad_(False).eval()
    policy.train()
    # what the processor hands over (TRN:417-425): fp32 pixel values, grid, prompt ids, all-ones mask
    frames = s["frames"].float()
    pix, grid = patchify_ref(frames)
    grid = torch.tensor([list(grid)])
    pk, gk = ("pixel_values_videos", "video_grid_thw") if kind == "video" else ("pixel_values", "image_grid_thw")
    prompt = s["prompt_ids"].cuda()
    prompt_inputs = {"input_ids": prompt, "attention_mask": torch.ones_like(prompt), pk: pix.cuda(), gk: grid.cuda()}
    calls = s["calls"]
    assert set(calls[0]["kwargs"]) == set(prompt_inputs)
    P = prompt.shape[1]

    # -- generate calls, with the recorded generation configs ------------------------------------------------------
    def gen(call, inputs):
        gc = GenerationConfig(**{k: v for k, v in call["generation_config"].items() if v is not None})
        out = policy.generate(**inputs, generation_config=gc)
        n, mx = gc.num_return_sequences, gc.max_new_tokens
        assert out.dtype == torch.long and out.shape[0] == n and P < out.shape[1] <= P + mx
        assert torch.equal(out[:, :P], prompt.expand(n, -1))
        return out
    gen(calls[0], prompt_inputs)
    if kind == "video":
        shuffled = dict(prompt_inputs)
        shuffled[pk] = patchify_ref(frames[s["frame_perm"]])[0].cuda()          # TRN:442-458
        gen(calls[1], shuffled)
    else:
        assert gen(calls[1], prompt_inputs).shape[1] == P + 1                   # the dummy call, TRN:481

    # -- scoring calls on the fixture's completions (sampling is not bit-reproducible across RNGs) -------------------
    comp = s["completions"].cuda()
    prompt_completion_ids = torch.cat([prompt.repeat(G, 1), comp], dim=1)
    prompt_inputs.pop("input_ids"); prompt_inputs.pop("attention_mask")       # TRN:507-508
    prompt_inputs[pk] = prompt_inputs[pk].repeat(G, 1)                         # TRN:510-518
    prompt_inputs[gk] = prompt_inputs[gk].repeat(G, 1)
    assert list(prompt_completion_ids.shape) == calls[2]["args"][0]["shape"]
    assert list(prompt_inputs[pk].shape) == calls[2]["kwargs"][pk]["shape"]
    score = (lambda m, ids, **kw: get_per_token_logps(None, m, ids, **kw)) if fast else _get_per_token_logps_reference_arithmetic
    per_token_logps = score(policy, prompt_completion_ids, **prompt_inputs)[:, P - 1:]
    with torch.inference_mode():
        ref_per_token_logps = score(refm, prompt_completion_ids, **prompt_inputs)[:, P - 1:]
    assert per_token_logps.requires_grad and not ref_per_token_logps.requires_grad

    # -- rewards / advantages through the product, loss with the trainer's arithmetic (TRN:551-552, 640-643) -------------
    RW.set_map_data(gold["map_rows"])

    def rewards_of(ids):
        texts = [gold["texts"][int(r[0]) % len(gold["texts"])] for r in ids.tolist()]
        n = len(texts)
        path = "synthetic/scene0000_00.mp4" if kind == "video" else "synthetic/scene0000_00.jpg"
        kw = dict(prompts=[None] * n, completions=[[{"role": "assistant", "content": t}] for t in texts], path=[path] * n,
                  solution=["<answer>B</answer>"] * n, problem_type=["multiple choice"] * n)
        with contextlib.redirect_stdout(io.StringIO()):
            return torch.tensor([RW.accuracy_reward(**kw), RW.format_reward(**kw)], dtype=torch.float32).T.contiguous().cuda()
    rpf = rewards_of(s["completions"])
    shuf_rpf = rewards_of(s["shuffled_completions"]) if kind == "video" else None
    lengths = T.completion_lengths(comp, d.eos_id)
    rewards, adv, std, temporal = T.reward_tail(rpf, shuf_rpf, lengths, G, True, True)
    completion_mask = GR.completion_mask(s["completions"], d.eos_id).cuda()
    x_clamped = torch.clamp(ref_per_token_logps.float() - per_token_logps.float(), min=-10, max=10)
    per_token_kl = torch.exp(x_clamped) - x_clamped - 1
    per_token_loss = torch.exp(per_token_logps - per_token_logps.detach()).float() * adv.unsqueeze(1)
    per_token_loss = -(per_token_loss - s["beta"] * per_token_kl)
    loss = ((per_token_loss * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    mean_kl = ((per_token_kl * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    assert loss.grad_fn is not None
    loss.backward()                                                              # accelerator.backward(loss), TRN:686
    torch.cuda.synchronize()

    # -- against what the reference computed -------------------------------------------------------------------------
    tol_kl = 0.08 if fast else 0.2        # fp32 log-softmax on the fast path; bf16 logits + bf16 log_softmax otherwise
    assert abs(mean_kl.item() - s["metrics"]["kl"]) < tol_kl * s["metrics"]["kl"] + 1e-3, (mean_kl.item(), s["metrics"]["kl"])
    assert abs(loss.item() - s["loss"]) < s["beta"] * (tol_kl * s["metrics"]["kl"] + 1e-3) + 2e-4, (loss.item(), s["loss"])
    got = T.step_metrics(T.pack_step_stats(lengths, rpf, rewards, std, mean_kl.detach(), temporal)[None].cpu(), G,
                         ["accuracy_reward", "format_reward"], temporal=True)
    for k, v in s["metrics"].items():
        if k != "kl":
            assert abs(got[k] - v) < 1e-5, (k, got[k], v)
    named = dict(policy.hf_named_grads())
    assert set(named) == set(s["grad_summary"])
    cos_all = []
    for k, g in s["grad_summary"].items():
        mine = named[k].float().cpu().flatten()
        ratio = float(mine.norm()) / (g["norm"] + 1e-30)
        assert 0.85 < ratio < 1.15, f"{k}: gradient norm ratio {ratio:.3f}"
        step = max(1, mine.numel() // 64)
        a = torch.cat([mine[:16], mine[::step][:64]])
        b = torch.cat([g["head"], g["stride_sample"]])
        if float(b.norm()) > 1e-3 * g["norm"]:
            cos_all.append(torch.nn.functional.cosine_similarity(a, b, dim=0).item())
    assert len(cos_all) > 20 and sum(cos_all) / len(cos_all) > 0.97, (len(cos_all), sum(cos_all) / max(1, len(cos_all)))
    assert min(cos_all) > 0.8, min(cos_all)

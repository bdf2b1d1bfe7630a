"""The reference trainer's four model calls (from_pretrained, generate(generation_config=...), model(...).logits with
autograd, loss.backward() -> .grad) against spacer_b200.hf_api on tiny dims: values vs the oracle, gradients of EVERY
parameter vs the oracle's autograd (SG_RLVR_trainer.py:163-190, 353-366, 463-481, 526-547, 640-643, 686)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _setup():
    from oracle import qwen2vl_ref as R
    from oracle.make_golden import tiny_case
    from spacer_b200 import config
    from spacer_b200.hf_api import Qwen2VLForConditionalGenerationB200
    d_or, d = R.dims_tiny(2, 2), config.tiny(2, 2)
    w = R.init_weights(d_or, seed=0)
    model = Qwen2VLForConditionalGenerationB200.from_dims(d, "cuda")
    model.load_state_dict(w)
    wb = {k: v.bfloat16().float().requires_grad_() for k, v in w.items()}
    return R, d_or, d, model, wb, tiny_case(d_or)


def _trainer_loss(per_token_logps, ref_lp, adv, mask, beta=0.04):
    """TRN:551-552, 640-643 verbatim arithmetic on whatever tensors the model call produced."""
    x = torch.clamp(ref_lp - per_token_logps, min=-10, max=10)
    kl = torch.exp(x) - x - 1
    ptl = torch.exp(per_token_logps - per_token_logps.detach()) * adv.unsqueeze(1)
    ptl = -(ptl - beta * kl)
    return ((ptl * mask).sum(dim=1) / mask.sum(dim=1)).mean()


def _oracle_loss_and_grads(R, d_or, wb, case, ref_lp, adv, mask):
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"]
    pos = R.rope_index_classic(ids, grid.repeat(G, 1), d_or)
    logits = R.model_logits(wb, ids, case["pixel_values"].bfloat16().float().repeat(G, 1), grid.repeat(G, 1), pos, d_or)
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    loss = _trainer_loss(lp, ref_lp, adv, mask.float())
    loss.backward()
    return loss.detach(), lp.detach(), {k: v.grad for k, v in wb.items()}


def _check_grads(model, ref_grads, floor_small=0.99, floor=0.995):
    got = dict(model.hf_named_grads())
    assert set(got) == set(ref_grads)
    cos_all = []
    for k, gref in ref_grads.items():
        g = got[k].float().cpu().reshape(gref.shape)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        rel = (g.norm() / (gref.norm() + 1e-30)).item()
        assert cos > (floor_small if gref.numel() < 4096 else floor), f"{k}: cosine {cos:.5f} (norm ratio {rel:.3f})"
        assert 0.9 < rel < 1.1, f"{k}: norm ratio {rel:.3f}"
        cos_all.append(cos)
    assert sum(cos_all) / len(cos_all) > 0.999


@pytest.mark.parametrize("path", ["logits", "per_token_logps"])
def test_trainer_scoring_calls_with_autograd(path):
    """`model(prompt_completion_ids, **prompt_inputs).logits` exactly as TRN:507-527 calls it (pixels and grid repeated
    x G, no attention mask), the reference's log-softmax/gather/slice, its loss, `loss.backward()`: loss and every
    parameter gradient vs the oracle.  Second variant: the `_get_per_token_logps` replacement (no [.,V] logits)."""
    from oracle import grpo_ref as GR
    from spacer_b200.hf_api import get_per_token_logps
    R, d_or, d, model, wb, case = _setup()
    gold = torch.load(os.path.join(GOLD, "tiny_model.pt"), weights_only=False)
    ids = case["input_ids"].cuda()
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    mask = GR.completion_mask(case["completion_ids"], d_or.eos_id)
    adv, _ = GR.advantages(case["rewards"], G)
    ref_lp = gold["ref_logps"]
    loss_ref, lp_ref, ref_grads = _oracle_loss_and_grads(R, d_or, wb, case, ref_lp, adv, mask)
    prompt_inputs = dict(pixel_values_videos=case["pixel_values"].cuda().repeat(G, 1),
                         video_grid_thw=case["grid_thw"].repeat(G, 1))
    model.train()
    if path == "logits":
        out = model(ids, **prompt_inputs)
        assert out.logits.shape == (G, ids.shape[1], d.vocab) and out.logits.requires_grad
        logits = out.logits[:, :-1, :]
        lp = torch.stack([torch.gather(lg.float().log_softmax(dim=-1), 1, i.unsqueeze(1)).squeeze(1)
                          for lg, i in zip(logits, ids[:, 1:])])
    else:
        lp = get_per_token_logps(None, model, ids, **prompt_inputs)
        assert lp.shape == (G, ids.shape[1] - 1) and lp.requires_grad and lp.dtype == torch.float32
    lp = lp[:, P - 1:]
    err = (lp.detach().cpu() - lp_ref).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 3e-3, (err.max().item(), err.mean().item())
    loss = _trainer_loss(lp, ref_lp.cuda(), adv.cuda(), mask.cuda().float())
    assert loss.grad_fn is not None
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * max(1.0, abs(loss_ref.item())) + 5e-4
    assert model.mat.grad is not None and model.vec.grad is not None
    _check_grads(model, ref_grads)


def test_inference_mode_scoring_and_logits_agree():
    """TRN:534-547: the reference-policy call runs under torch.inference_mode(); both paths give the same numbers."""
    R, d_or, d, model, wb, case = _setup()
    ids = case["input_ids"].cuda()
    G = ids.shape[0]
    kw = dict(pixel_values_videos=case["pixel_values"].cuda().repeat(G, 1), video_grid_thw=case["grid_thw"].repeat(G, 1))
    with torch.inference_mode():
        logits = model(ids, **kw).logits
        lp_fast = model.per_token_logps(ids, **kw)
    assert not logits.requires_grad and not lp_fast.requires_grad
    lp_slow = torch.gather(logits[:, :-1].float().log_softmax(-1), 2, ids[:, 1:, None]).squeeze(-1)
    assert (lp_slow - lp_fast).abs().max().item() < 1e-3
    # rows with different visual inputs are rejected, so is a padded batch
    from spacer_b200.ops import SpacerError
    bad = dict(kw, video_grid_thw=torch.cat([case["grid_thw"] * 2, case["grid_thw"].repeat(G - 1, 1)]))
    with pytest.raises(SpacerError):
        model(ids, **bad)
    am = torch.ones_like(ids)
    am[0, 0] = 0
    with pytest.raises(SpacerError):
        model(ids, attention_mask=am, **kw)


def test_generate_takes_the_reference_generation_config():
    """TRN:277-302, 463: generate(**prompt_inputs, generation_config=GenerationConfig(...)) -> [G, P + C'], prompt echoed,
    rows padded with pad_token_id after their first EOS; HF's implicit top_k = 50 is applied."""
    from transformers import GenerationConfig
    R, d_or, d, model, wb, case = _setup()
    G, C = 6, 24
    gc = GenerationConfig(max_new_tokens=C, do_sample=True, top_p=0.95, temperature=1, num_return_sequences=G,
                          pad_token_id=d.pad_id)
    prompt = case["prompt_ids"].cuda()
    prompt_inputs = dict(input_ids=prompt, attention_mask=torch.ones_like(prompt),
                         pixel_values_videos=case["pixel_values"].cuda(), video_grid_thw=case["grid_thw"])
    torch.manual_seed(0)
    out = model.generate(**prompt_inputs, generation_config=gc)
    P = prompt.shape[1]
    assert out.dtype == torch.long and out.shape[0] == G and P < out.shape[1] <= P + C
    assert torch.equal(out[:, :P], prompt.expand(G, -1))
    comp = out[:, P:]
    for r in range(G):
        eos = (comp[r] == d.eos_id).nonzero()
        if len(eos):
            assert (comp[r, int(eos[0]) + 1:] == d.pad_id).all()
    torch.manual_seed(0)
    again = model.generate(**prompt_inputs, generation_config=gc)
    assert torch.equal(out, again)                    # the torch generator seeds the Philox stream
    other = model.generate(**prompt_inputs, generation_config=gc)
    assert other.shape[0] == G and not torch.equal(other[:, P:P + 8], out[:, P:P + 8])
    # top_k: with top_k = 1 in the config sampling degenerates to greedy -> all rows identical
    g1 = model.generate(**prompt_inputs, generation_config=GenerationConfig(max_new_tokens=8, do_sample=True, top_k=1,
                                                                           num_return_sequences=3, pad_token_id=d.pad_id))
    assert torch.equal(g1[0], g1[1]) and torch.equal(g1[0], g1[2])
    with pytest.raises(Exception):
        model.generate(**prompt_inputs, generation_config=GenerationConfig(max_new_tokens=4, num_beams=2))


def test_eval_style_generate_left_padded_batch():
    """SpaceR-Eval/data_utils/vsibench.py:157-180: processor(..., padding=True) with padding_side='left', then
    generate(**inputs, use_cache=True, max_new_tokens=N, temperature=0.01) over the checkpoint's generation config
    (repetition_penalty 1.05): each row equals the single-prompt greedy result."""
    R, d_or, d, model, wb, case = _setup()
    model.eval()
    model.engine.generation_config = dict(do_sample=True, repetition_penalty=1.05, temperature=0.1, top_k=1, top_p=0.001,
                                          eos_token_id=[d.eos_id, d.pad_id], pad_token_id=d.pad_id)
    prompt = case["prompt_ids"][0]
    short = prompt[3:]                                   # a second, shorter prompt (3 text tokens less)
    ids = torch.full((2, prompt.numel()), d.pad_id)
    am = torch.zeros_like(ids)
    ids[0], am[0] = prompt, 1
    ids[1, 3:], am[1, 3:] = short, 1
    pix = case["pixel_values"].cuda()
    inputs = dict(input_ids=ids.cuda(), attention_mask=am.cuda(), pixel_values_videos=torch.cat([pix, pix]),
                  video_grid_thw=case["grid_thw"].repeat(2, 1))
    N = 12
    out = model.generate(**inputs, use_cache=True, max_new_tokens=N, temperature=0.01)
    assert out.shape[0] == 2 and out.shape[1] <= prompt.numel() + N and torch.equal(out[:, :prompt.numel()].cpu(), ids)
    for b, p_ in enumerate((prompt, short)):
        single = model.engine.generate(p_[None], pix, case["grid_thw"], generation_config={}, max_new_tokens=N,
                                       temperature=0.01)
        got = out[b, prompt.numel():]
        want = single[0, p_.numel():]
        assert torch.equal(got[:want.numel()], want) and (got[want.numel():] == d.pad_id).all()
    # repetition penalty is live: with a huge penalty no token repeats inside a completion (nor any prompt token)
    rp = model.engine.generate(prompt[None], pix, case["grid_thw"], generation_config={}, max_new_tokens=N, do_sample=False,
                               repetition_penalty=1e4, min_new_tokens=N)
    comp = rp[0, prompt.numel():].tolist()
    assert len(set(comp)) == len(comp) and not (set(comp) & set(prompt.tolist()))


def test_fused_adamw_behind_torch_optim_interface():
    """HF Trainer's optimizer slot: FusedAdamW reads .grad of the arenas and applies the engine's clipped AdamW."""
    from spacer_b200.hf_api import FusedAdamW
    R, d_or, d, model, wb, case = _setup()
    ids = case["input_ids"].cuda()
    G = ids.shape[0]
    kw = dict(pixel_values_videos=case["pixel_values"].cuda().repeat(G, 1), video_grid_thw=case["grid_thw"].repeat(G, 1))
    opt = FusedAdamW(model, lr=1e-3)
    before = model.mat.detach().clone()
    lp = model.per_token_logps(ids, **kw)
    (-lp.mean()).backward()
    opt.step()
    opt.zero_grad()
    torch.cuda.synchronize()
    assert not torch.equal(before, model.mat.detach()) and torch.equal(model.mat.detach(), model.engine.params.mat)
    lp2 = model.per_token_logps(ids, **kw)
    assert lp2.mean().item() > lp.mean().item()          # one ascent step on the mean log-prob


def test_processor_to_generate_end_to_end(tmp_path_factory):
    """Rows a1-a3 chained into the rollout: conversation -> chat template -> processor(text, videos) (tokenisation,
    placeholder expansion, GPU normalise/patchify) -> generate(**prompt_inputs, generation_config) -> batch_decode -> the
    reward functions' input strings (TRN:390-467, 555-560)."""
    from dataclasses import replace
    from transformers import GenerationConfig
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_text_cpu import build_tok_dir
    from oracle import qwen2vl_ref as R
    from oracle.vision_ref import patchify_ref
    from spacer_b200 import config
    from spacer_b200.hf_api import Qwen2VLForConditionalGenerationB200
    from spacer_b200.text import Qwen2Tokenizer, Qwen2VLProcessorB200
    path, n = build_tok_dir(tmp_path_factory)
    tok = Qwen2Tokenizer.from_pretrained(path)
    d = replace(config.tiny(2, 2), image_token_id=tok.vocab["<|image_pad|>"], video_token_id=tok.vocab["<|video_pad|>"],
                vision_start_id=tok.vocab["<|vision_start|>"], vision_end_id=tok.vocab["<|vision_end|>"],
                eos_id=tok.vocab["<|im_end|>"], pad_id=tok.vocab["<|endoftext|>"])
    model = Qwen2VLForConditionalGenerationB200.from_dims(d, "cuda", seed=0)
    proc = Qwen2VLProcessorB200(tok, d, device="cuda")
    conv = [{"role": "user", "content": [{"type": "video"}, {"type": "text", "text": "How many chairs are in this room?"}]}]
    text = proc.apply_chat_template(conv, tokenize=False, add_generation_prompt=True)
    frames = torch.randint(0, 256, (4, 3, 112, 112), generator=torch.Generator().manual_seed(2)).float()   # fetch_video output
    inputs = proc(text=[text], images=None, videos=[frames], return_tensors="pt", padding=True, padding_side="left",
                  add_special_tokens=False)
    ref_pix, grid = patchify_ref(frames)
    assert torch.equal(inputs["pixel_values_videos"].cpu(), ref_pix) and inputs["video_grid_thw"].tolist() == [list(grid)]
    n_v = grid[0] * grid[1] * grid[2] // 4
    ids = inputs["input_ids"]
    assert ids.shape[0] == 1 and int((ids == d.video_token_id).sum()) == n_v and bool(inputs["attention_mask"].all())
    assert tok.decode(ids[0]).replace("<|video_pad|>" * n_v, "<|video_pad|>") == text
    G, C = 4, 10
    out = model.generate(**inputs.to("cuda"), generation_config=GenerationConfig(
        max_new_tokens=C, do_sample=True, top_p=0.95, temperature=1, num_return_sequences=G, pad_token_id=proc.pad_token_id))
    P = ids.shape[1]
    assert out.shape[0] == G and torch.equal(out[:, :P].cpu(), ids.expand(G, -1))
    completions = proc.batch_decode(out[:, P:], skip_special_tokens=True)
    assert len(completions) == G and all(isinstance(c, str) for c in completions)
    assert all("<|im_end|>" not in c and "<|endoftext|>" not in c for c in completions)


def test_from_pretrained_with_the_reference_kwargs_and_generation_config(tmp_path):
    """TRN:163-190: `from_pretrained(model_id, attn_implementation=..., torch_dtype=..., use_cache=...)` on a local HF
    directory; generation_config.json supplies the eos list and the evaluation defaults; wrong class / dtype / kwargs raise."""
    import json
    from spacer_b200.hf_api import Qwen2_5_VLForConditionalGenerationB200, Qwen2VLForConditionalGenerationB200
    from spacer_b200.ops import SpacerError
    R, d_or, d, model, wb, case = _setup()
    path = str(tmp_path / "Qwen2-VL-tiny")
    model.engine.generation_config = dict(do_sample=True, repetition_penalty=1.05, temperature=0.1, top_k=1, top_p=0.001,
                                          eos_token_id=[d.eos_id, d.pad_id], pad_token_id=d.pad_id)
    model.save_pretrained(path)
    assert json.load(open(os.path.join(path, "generation_config.json")))["eos_token_id"] == [d.eos_id, d.pad_id]
    m2 = Qwen2VLForConditionalGenerationB200.from_pretrained(path, attn_implementation="flash_attention_2",
                                                             torch_dtype=torch.bfloat16, use_cache=False)
    assert m2.config._name_or_path == path and isinstance(m2.warnings_issued, dict)
    assert m2.engine.dims.eos_ids == (d.eos_id, d.pad_id) and m2.generation_config.repetition_penalty == 1.05
    sd, sd2 = model.state_dict(), m2.state_dict()
    assert set(sd) == set(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert sum(p.numel() for p in m2.parameters()) == model.engine.params.sizes["mat"] + model.engine.params.sizes["vec"]
    ids = case["input_ids"].cuda()
    G = ids.shape[0]
    kw = dict(pixel_values_videos=case["pixel_values"].cuda().repeat(G, 1), video_grid_thw=case["grid_thw"].repeat(G, 1))
    with torch.no_grad():
        assert torch.equal(model.per_token_logps(ids, **kw), m2.per_token_logps(ids, **kw))
    with pytest.raises(SpacerError):
        Qwen2_5_VLForConditionalGenerationB200.from_pretrained(path)                       # a qwen2_vl checkpoint
    with pytest.raises(SpacerError):
        Qwen2VLForConditionalGenerationB200.from_pretrained(path, torch_dtype=torch.float16)
    with pytest.raises(TypeError):
        Qwen2VLForConditionalGenerationB200.from_pretrained(path, load_in_8bit=True)


def test_qwen25_wrapper_scores_like_the_engine():
    """The class run_SpaceR_SG_RLVR.sh actually trains (Qwen2.5-VL): wrapper scoring == engine scoring, second_per_grid_ts is
    accepted by generate() and ignored by the scoring call like the reference's `del` (TRN:519-520)."""
    from oracle import qwen25vl_ref as R25
    from oracle.make_golden import tiny_case25
    from spacer_b200 import config
    from spacer_b200.hf_api import Qwen2_5_VLForConditionalGenerationB200
    from spacer_b200.model import pack_prompt_completions
    d_or, d = R25.dims25_tiny(), config.tiny25()
    model = Qwen2_5_VLForConditionalGenerationB200.from_dims(d, "cuda")
    model.load_state_dict(R25.init_weights(d_or, seed=0))
    case = tiny_case25(d_or)
    ids = case["input_ids"].cuda()
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    pix = case["pixel_values"].cuda()
    kw = dict(pixel_values_videos=pix.repeat(G, 1), video_grid_thw=case["grid_thw"].repeat(G, 1))
    with torch.no_grad():
        lp = model.per_token_logps(ids, second_per_grid_ts=[1.5] * G, **kw)[:, P - 1:]
        batch = pack_prompt_completions(case["prompt_ids"], case["completion_ids"], case["grid_thw"], d, model.engine.device)
        want = model.engine.per_token_logps(batch, pix, case["grid_thw"])
    assert torch.equal(lp, want)
    out = model.generate(input_ids=case["prompt_ids"].cuda(), attention_mask=torch.ones_like(case["prompt_ids"]).cuda(),
                         pixel_values_videos=pix, video_grid_thw=case["grid_thw"], second_per_grid_ts=[1.5],
                         max_new_tokens=5, do_sample=False)
    assert out.shape == (1, P + 5) or (out.shape[0] == 1 and out.shape[1] <= P + 5)

"""ORACLE tooling (test infrastructure only).  Generates the committed fixtures under tests/golden/:

  rewards.json     -- outputs of the REFERENCE's own reward / map-verifier functions
                      (/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1/{SG-RLVR.py,extract_map.py}),
                      imported here with stubs for the packages missing from this image,
                      on a hand-written + seeded-fuzz corpus of completions and <map> bodies.
  tiny_model.pt    -- outputs of HF transformers 5.5.0 `Qwen2VLForConditionalGeneration` (the third-party
                      implementation the reference calls) on a tiny random-init config: vision embeddings,
                      per-token log-probs (policy / perturbed "reference policy"), the GRPO loss of
                      SG_RLVR_trainer.py:551-552,632-643 evaluated with torch autograd, and gradients.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
Nothing under tests/ or the product reads /root/reference at run time.
"""
from __future__ import annotations

import importlib.util
import json
import os
import random
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1"
OUT = os.path.join(ROOT, "tests", "golden")

MAP_ROWS = {
    "scene0000_00": {"video_id": "scene0000_00",
                     "cognitive_map": {"table": [[0, 3], [5, 7]], "chair": [[9, 3]], "window": [[6, 5]]}},
    "scene0001_00": {"video_id": "scene0001_00",
                     "cognitive_map": {"sofa": [[1, 1]], "coffee table": [[2, 2], [8, 8]], "tv": [[9, 0]],
                                       "table": [[4, 4]]}},
}


def load_reference():
    sys.modules.setdefault("jsonlines", types.ModuleType("jsonlines"))
    sys.path.insert(0, REF)
    import extract_map  # noqa: the reference's own module

    for name in ["trl", "nltk", "nltk.translate", "nltk.translate.bleu_score", "rouge_score", "trainer",
                 "math_verify", "latex2sympy2_extended", "datasets"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    t = sys.modules["trl"]
    for n in ["GRPOConfig", "GRPOTrainer", "ModelConfig", "ScriptArguments", "TrlParser", "get_peft_config"]:
        setattr(t, n, type(n, (), {}))
    sys.modules["nltk.translate.bleu_score"].sentence_bleu = None
    sys.modules["nltk.translate.bleu_score"].SmoothingFunction = None
    sys.modules["rouge_score"].rouge_scorer = None
    tr = sys.modules["trainer"]
    for n in ["Qwen2VLGRPOTrainer", "Qwen2VLGRPOVLLMTrainerModified", "SGRLVRTrainer"]:
        setattr(tr, n, type(n, (), {}))
    ds = sys.modules["datasets"]
    for n in ["load_dataset", "load_from_disk", "Dataset", "DatasetDict"]:
        setattr(ds, n, None)
    mv = sys.modules["math_verify"]
    mv.parse = mv.verify = None
    sys.modules["latex2sympy2_extended"].NormalizationConfig = None
    spec = importlib.util.spec_from_file_location("sg_rlvr_ref", os.path.join(REF, "SG-RLVR.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.MAP_DATA = MAP_ROWS
    return extract_map, mod


MAP_BODIES = [
    '{"table":[[1,3],[5,5]],"chair":[[9,9]]}',
    "{'table': [[0,3],[5,7]], 'chair': [[9,3]], 'window': [[6,5]]}",
    "table: (1, 3), (5,5); Chair at <9,9>; sofa [1,1]",
    'str{{"table": "[2, 2]"}}',
    'str{"table": [[2, 2]]}',
    "{'Table': [1.7, 3.2], 'CHAIR': [[9, 3], [1]], 'window': '6, 5'}",
    "{'<table}': [[1,2]], \"'chair'\": [(3,4),(5,6)], 'lamp': [[0,0]]}",
    "{'table': [[1,2],[3]], 'chair': [[[4,5]]], 'window': [8,]}",
    "nothing",
    "",
    "   ",
    "{'table': 5}",
    "{'table': [[1e999, 2]]}",
    "{'table': [['<3>', '(4)']], 'chair': [['a', 'b']]}",
    "{broken: [1,2], 'table': [[1,2]]",
    "The table is at 1,3 and 5 5. chair=9;3 window -> [6.9, 5.2] table again 0 0",
    "coffee table (2,2) (8,8), table 4 4, tv 9 0, sofa 1 1",
    "tables 1 2 table3 4 tablechair 5 6 chair.7.8",
    "chair 1 chair 2 3 chair 4 5 6",
    "window: -1, -2; table: 3.9 -4.1",
    "{'table': [[True, 2]], 'chair': [[None, 1]]}",
    "[1, 2]",
    "{'table': ([1,2],[3,4])}",
    "{'table': {'x': 1}}",
    "{1: [[1,2]], 'chair': [[2,3]]}",
]

OBJECT_LISTS = [["table", "chair", "window"], ["sofa", "coffee table", "tv", "table"], ["Table", "CHAIR"], []]


def fuzz_bodies(n, seed):
    rnd = random.Random(seed)
    names = ["table", "chair", "window", "sofa", "coffee table", "tv", "lamp", "Table", "armchair"]
    seps = [": ", " at ", " ", " -> ", "=", " (", " is located at "]
    out = []
    for _ in range(n):
        kind = rnd.random()
        if kind < 0.45:
            items = []
            for nm in rnd.sample(names, rnd.randint(1, 4)):
                pts = []
                for _ in range(rnd.randint(0, 3)):
                    x, y = rnd.randint(-1, 11), rnd.randint(0, 10)
                    f = rnd.random()
                    if f < 0.6:
                        pts.append(f"[{x}, {y}]")
                    elif f < 0.75:
                        pts.append(f"({x}.{rnd.randint(0, 9)}, {y})")
                    elif f < 0.85:
                        pts.append(f"'{x},{y}'")
                    elif f < 0.92:
                        pts.append(f"[{x}]")
                    else:
                        pts.append(f"[[{x}, {y}]]")
                q = rnd.choice(["'", '"'])
                val = "[" + ", ".join(pts) + "]" if rnd.random() < 0.9 else (pts[0] if pts else "[]")
                items.append(f"{q}{nm}{q}: {val}")
            body = "{" + ", ".join(items) + "}"
            if rnd.random() < 0.1:
                body = "str{" + body + "}"
            if rnd.random() < 0.1:
                body = body[:-1]
        else:
            parts = []
            for nm in rnd.choices(names, k=rnd.randint(1, 5)):
                nums = " ".join(str(rnd.choice([rnd.randint(0, 10), round(rnd.uniform(0, 10), 1)]))
                                + rnd.choice([",", "", ";", ")"]) for _ in range(rnd.randint(0, 5)))
                parts.append(rnd.choice([nm, nm.upper(), nm + "s"]) + rnd.choice(seps) + nums)
            body = rnd.choice(["; ", ". ", "\n", " and "]).join(parts)
        out.append(body)
    return out


def reward_cases():
    mc = "multiple choice"
    good_map = "<map>{'table':[[1,3],[5,5]],'chair':[[9,9]]}</map>"
    cases = [
        (mc, "<think>x</think><answer>B</answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, "<think>x</think><answer> B </answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, "<think>x</think><answer>C</answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, f"<think>x</think>{good_map}<answer>B</answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, f"<think>x {good_map}</think><answer>B</answer>", "<answer>B</answer>", "/d/scene0000_00.mp4"),
        (mc, "<think>x</think><map>nothing</map><answer>B</answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, f"<think>x</think>{good_map}<answer>B</answer>", "<answer>B</answer>", "a/unknown_scene.mp4"),
        (mc, "<map></map><answer>B</answer>", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, "no tags at all", "<answer>B</answer>", "a/scene0000_00.mp4"),
        (mc, "<think>\nline\n</think>\n\n<answer>\nB\n</answer>", "<answer>B</answer>", "scene0001_00.mp4"),
        (mc, "<think>x</think><map>coffee table (2,2) (8,8), table 4 4, tv 9 0, sofa 1 1</map><answer>A</answer>",
         "<answer>A</answer>", "x/scene0001_00.mp4"),
        ("numerical", "<answer>about three</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>3.3</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>4.2</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>a table</answer>", "<answer>1</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>none</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>3</answer>", "<answer>0</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>twenty one</answer>", "<answer>21</answer>", "a/scene0000_00.mp4"),
        ("numerical", f"<think>t</think>{good_map}<answer>3.1</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<think>t</think><map>zzz</map><answer>3.1</answer>", "<answer>3</answer>", "a/scene0000_00.mp4"),
        ("numerical", "<answer>12 meters</answer>", "<answer>12.5</answer>", "a/scene0000_00.mp4"),
        ("OCR", "<answer>hello big world</answer>", "<answer>hello world</answer>", "p"),
        ("OCR", "<answer>hello world</answer>", "<answer>hello world</answer>", "p"),
        ("OCR", "<answer></answer>", "<answer>hello world</answer>", "p"),
        ("OCR", "<answer>a b c d e f</answer>", "<answer>x</answer>", "p"),
        ("regression", "<answer>1,200</answer>", "<answer>1000</answer>", "p"),
        ("regression", "<answer>abc</answer>", "<answer>1000</answer>", "p"),
        ("regression", "<answer>-5</answer>", "<answer>5</answer>", "p"),
        ("regression", "<answer>0</answer>", "<answer>0</answer>", "p"),
        ("regression", "<answer>7.5</answer>", "<answer>7</answer>", "p"),
        ("unknown type", "<answer>x</answer>", "<answer>x</answer>", "p"),
        ("free-form", "<answer>the cat sat</answer>", "<answer>the cat sat</answer>", "p"),
    ]
    return cases


def gen_rewards():
    em, sg = load_reference()
    rec = {"map_rows": MAP_ROWS, "extract": [], "score": [], "accuracy": [], "format": []}
    bodies = MAP_BODIES + fuzz_bodies(300, seed=7)
    for body in bodies:
        for ol in OBJECT_LISTS:
            try:
                out = em.extract_map_data(body, ol)
                rec["extract"].append({"body": body, "objects": ol, "out": out})
            except BaseException as e:  # noqa
                rec["extract"].append({"body": body, "objects": ol, "raises": type(e).__name__})
    sols = [r["cognitive_map"] for r in MAP_ROWS.values()] + [{}, {"x": []}, {"table": [[5, 5]]}]
    resps = [x["out"] for x in rec["extract"] if "out" in x][::3] + [{}, {"x": []}, {"table": []}]
    for sol in sols:
        for resp in resps:
            try:
                rec["score"].append({"response": resp, "solution": sol, "out": em.calculate_prediction_score(resp, sol, 10)})
            except BaseException as e:  # noqa
                rec["score"].append({"response": resp, "solution": sol, "raises": type(e).__name__})
    for qt, content, sol, path in reward_cases():
        comp = [[{"role": "assistant", "content": content}]]
        acc = sg.accuracy_reward(comp, [sol], [path], problem_type=[qt])
        fmt = sg.format_reward(comp)
        rec["accuracy"].append({"type": qt, "content": content, "solution": sol, "path": path, "out": float(acc[0])})
        rec["format"].append({"content": content, "out": float(fmt[0])})
    # batch semantics: the first row's type applies to all rows
    comps = [[{"role": "assistant", "content": c}] for _, c, _, _ in reward_cases()[:6]]
    sols6 = [s for _, _, s, _ in reward_cases()[:6]]
    paths6 = [p for _, _, _, p in reward_cases()[:6]]
    out = sg.accuracy_reward(comps, sols6, paths6, problem_type=["multiple choice"] * 6)
    rec["accuracy_batch"] = {"contents": [c[0]["content"] for c in comps], "solutions": sols6, "paths": paths6,
                             "type": "multiple choice", "out": [float(x) for x in out]}
    with open(os.path.join(OUT, "rewards.json"), "w") as f:
        json.dump(rec, f, indent=0, sort_keys=True)
    print("rewards.json:", {k: len(v) for k, v in rec.items() if isinstance(v, list)})


def hf_config(d):
    from transformers import Qwen2VLConfig

    cfg = Qwen2VLConfig(
        text_config=dict(hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                         num_key_value_heads=d.kv_heads, intermediate_size=d.inter, vocab_size=d.vocab,
                         rms_norm_eps=d.rms_eps,
                         rope_parameters={"rope_type": "default", "rope_theta": d.rope_theta,
                                          "mrope_section": list(d.mrope_section)},
                         tie_word_embeddings=d.tie, max_position_embeddings=32768, use_sliding_window=False,
                         bos_token_id=d.pad_id, eos_token_id=d.eos_id, pad_token_id=d.pad_id),
        vision_config=dict(depth=d.v_depth, embed_dim=d.v_embed, hidden_size=d.hidden, num_heads=d.v_heads,
                           mlp_ratio=d.v_mlp // d.v_embed, patch_size=d.patch, temporal_patch_size=d.t_patch,
                           spatial_merge_size=d.merge, in_channels=d.in_ch, hidden_act="quick_gelu"),
        image_token_id=d.image_token_id, video_token_id=d.video_token_id,
        vision_start_token_id=d.vision_start_id, vision_end_token_id=d.vision_end_id, tie_word_embeddings=d.tie)
    cfg._attn_implementation = "eager"
    return cfg


def tiny_case(d, G=4, C=12, grid=(2, 8, 8), seed=11):
    """Deterministic synthetic GRPO step inputs at tiny dims (shared by the generator and the tests)."""
    from oracle import qwen2vl_ref as R

    g = torch.Generator().manual_seed(seed)
    grid_thw = torch.tensor([list(grid)])
    n_p = grid[0] * grid[1] * grid[2]
    pix = torch.randn(n_p, d.patch_dim, generator=g)
    prompt = R.build_prompt_ids(d, n_p // 4, 6, 10, seed=seed + 1)
    comp = torch.randint(10, 2000, (G, C), generator=g)
    comp[1, 7] = d.eos_id
    comp[1, 8:] = d.pad_id
    comp[3, C - 1] = d.eos_id
    ids = torch.cat([prompt.repeat(G, 1), comp], 1)
    rewards = torch.tensor([1.59, 0.0, 2.0, 1.0][:G] + [0.5] * max(0, G - 4))
    return dict(grid_thw=grid_thw, pixel_values=pix, prompt_ids=prompt, completion_ids=comp, input_ids=ids, rewards=rewards)


def hf_config25(d):
    from transformers import Qwen2_5_VLConfig

    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                         num_key_value_heads=d.kv_heads, intermediate_size=d.inter, vocab_size=d.vocab,
                         rms_norm_eps=d.rms_eps,
                         rope_parameters={"rope_type": "default", "rope_theta": d.rope_theta,
                                          "mrope_section": list(d.mrope_section)},
                         tie_word_embeddings=d.tie, max_position_embeddings=32768, use_sliding_window=False,
                         bos_token_id=d.pad_id, eos_token_id=d.eos_id, pad_token_id=d.pad_id),
        vision_config=dict(depth=d.v_depth, hidden_size=d.v_embed, intermediate_size=d.v_mlp, num_heads=d.v_heads,
                           out_hidden_size=d.hidden, patch_size=d.patch, temporal_patch_size=d.t_patch,
                           spatial_merge_size=d.merge, in_channels=d.in_ch, hidden_act="silu",
                           window_size=d.v_window, fullatt_block_indexes=list(d.v_fullatt),
                           tokens_per_second=d.tokens_per_second),
        image_token_id=d.image_token_id, video_token_id=d.video_token_id,
        vision_start_token_id=d.vision_start_id, vision_end_token_id=d.vision_end_id, tie_word_embeddings=d.tie)
    cfg._attn_implementation = "eager"
    return cfg


def tiny_case25(d, G=4, C=10, seed=23):
    """Qwen2.5-VL case: grid (2, 12, 8) -> merged 6 x 4 per frame = one full 4x4 window + one ragged 2x4 window."""
    return tiny_case(d, G=G, C=C, grid=(2, 12, 8), seed=seed)


def gen_tiny_model25():
    """tests/golden/tiny_model25.pt: the real HF Qwen2_5_VLForConditionalGeneration (transformers 5.5.0) on the tiny
    Qwen2.5-VL case: vision embeddings, per-token log-probs with explicit classic position ids, GRPO loss, gradient
    norms/samples, plus HF's own default position ids (second_per_grid_ts given and absent)."""
    from transformers import Qwen2_5_VLForConditionalGeneration

    from oracle import grpo_ref as GR
    from oracle import qwen25vl_ref as R25
    from oracle import qwen2vl_ref as R

    d = R25.dims25_tiny()
    w = R25.init_weights(d, seed=0)
    m = Qwen2_5_VLForConditionalGeneration(hf_config25(d)).float()
    m.load_state_dict(w, strict=True)
    case = tiny_case25(d)
    ids = case["input_ids"]
    G, P = ids.shape[0], case["prompt_ids"].shape[1]
    grid = case["grid_thw"].repeat(G, 1)
    pix = case["pixel_values"].repeat(G, 1)
    sec = [1.5]
    pos = R25.rope_index_classic(ids, grid, d, second_per_grid_ts=sec * G)
    mm = (ids == d.video_token_id).long() * 2 + (ids == d.image_token_id).long()
    m.train()
    logits = m(input_ids=ids, pixel_values_videos=pix, video_grid_thw=grid, position_ids=pos, mm_token_type_ids=mm).logits
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    with torch.no_grad():
        ve = m.model.visual(case["pixel_values"], case["grid_thw"]).pooler_output
        hf_pos_sec, _ = m.model.get_rope_index(ids[:1], mm[:1], video_grid_thw=grid[:1], second_per_grid_ts=torch.tensor([2.0]))
        hf_pos_none, _ = m.model.get_rope_index(ids[:1], mm[:1], video_grid_thw=grid[:1])
    mask = GR.completion_mask(case["completion_ids"], d.eos_id)
    adv, _ = GR.advantages(case["rewards"][:G], G)
    rlp = lp.detach() + 0.03 * torch.randn(lp.shape, generator=torch.Generator().manual_seed(9))
    loss, mean_kl = GR.grpo_loss(lp, rlp, adv, mask, beta=0.04)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    keep = ["model.visual.patch_embed.proj.weight", "model.visual.blocks.0.attn.qkv.weight",
            "model.visual.blocks.0.mlp.gate_proj.weight", "model.visual.blocks.1.mlp.up_proj.bias",
            "model.visual.blocks.1.mlp.down_proj.weight", "model.visual.blocks.0.norm2.weight",
            "model.visual.merger.ln_q.weight", "model.visual.merger.mlp.2.weight",
            "model.language_model.layers.0.self_attn.q_proj.weight", "lm_head.weight"]
    out = {
        "dims": "tiny25", "weights_seed": 0, "case_seed": 23, "beta": 0.04, "second_per_grid_ts": sec[0],
        "vision_embeds": ve.detach().half(), "logps": lp.detach(), "ref_logps": rlp, "mask": mask, "advantages": adv,
        "loss": loss.detach(), "mean_kl": mean_kl.detach(),
        "grad_norms": {k: grads[k].norm() for k in grads},
        "grad_samples": {k: grads[k].flatten()[:: max(1, grads[k].numel() // 64)][:64].clone() for k in keep},
        "pos_classic": pos[:, 0].to(torch.int16), "pos_hf55_sec2": hf_pos_sec[:, 0].to(torch.int16),
        "pos_hf55_none": hf_pos_none[:, 0].to(torch.int16),
        "transformers_version": __import__("transformers").__version__,
    }
    torch.save(out, os.path.join(OUT, "tiny_model25.pt"))
    print("tiny_model25.pt: loss", float(loss), "kl", float(mean_kl), "lp[0,:4]", lp[0, :4].tolist())


def gen_tiny_model():
    from transformers import Qwen2VLForConditionalGeneration

    from oracle import grpo_ref as GR
    from oracle import qwen2vl_ref as R

    d = R.dims_tiny()
    w = R.init_weights(d, seed=0)
    m = Qwen2VLForConditionalGeneration(hf_config(d)).float()
    m.load_state_dict(w, strict=True)
    ref = Qwen2VLForConditionalGeneration(hf_config(d)).float().eval()
    wr = R.init_weights(d, seed=0)
    gp = torch.Generator().manual_seed(5)
    for k in wr:  # the "reference policy" = policy + small perturbation, so KL != 0
        wr[k] = wr[k] + 0.002 * torch.randn(wr[k].shape, generator=gp)
    if not d.tie:
        pass
    ref.load_state_dict(wr, strict=True)
    case = tiny_case(d)
    G = case["input_ids"].shape[0]
    P = case["prompt_ids"].shape[1]
    ids = case["input_ids"]
    grid = case["grid_thw"].repeat(G, 1)
    pix = case["pixel_values"].repeat(G, 1)
    pos = R.rope_index_classic(ids, grid, d)
    mm = (ids == d.video_token_id).long() * 2 + (ids == d.image_token_id).long()
    m.train()
    logits = m(input_ids=ids, pixel_values_videos=pix, video_grid_thw=grid, position_ids=pos, mm_token_type_ids=mm).logits
    lp = R.per_token_logps(logits, ids)[:, P - 1:]
    with torch.no_grad():
        rl = ref(input_ids=ids, pixel_values_videos=pix, video_grid_thw=grid, position_ids=pos, mm_token_type_ids=mm).logits
        rlp = R.per_token_logps(rl, ids)[:, P - 1:]
        ve = m.model.visual(case["pixel_values"], case["grid_thw"]).pooler_output
    mask = GR.completion_mask(case["completion_ids"], d.eos_id)
    adv, std = GR.advantages(case["rewards"], G)
    loss, mean_kl = GR.grpo_loss(lp, rlp, adv, mask, beta=0.04)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    keep = ["model.visual.patch_embed.proj.weight", "model.visual.blocks.0.attn.qkv.weight",
            "model.visual.blocks.1.mlp.fc1.bias", "model.visual.merger.mlp.2.weight",
            "model.language_model.embed_tokens.weight", "model.language_model.layers.0.self_attn.q_proj.weight",
            "model.language_model.layers.0.self_attn.k_proj.bias", "model.language_model.layers.1.mlp.gate_proj.weight",
            "model.language_model.layers.1.mlp.down_proj.weight", "model.language_model.layers.0.input_layernorm.weight",
            "model.language_model.norm.weight", "lm_head.weight"]
    out = {
        "dims": "tiny", "weights_seed": 0, "ref_perturb_seed": 5, "case_seed": 11, "beta": 0.04,
        "vision_embeds": ve.detach().half(), "logps": lp.detach(), "ref_logps": rlp, "mask": mask,
        "advantages": adv, "loss": loss.detach(), "mean_kl": mean_kl.detach(),
        "grad_norms": {k: grads[k].norm() for k in grads},
        "grad_samples": {k: grads[k].flatten()[:: max(1, grads[k].numel() // 64)][:64].clone() for k in keep},
        "pos_classic": pos[:, 0].to(torch.int16),
        "transformers_version": __import__("transformers").__version__,
    }
    torch.save(out, os.path.join(OUT, "tiny_model.pt"))
    print("tiny_model.pt: loss", float(loss), "kl", float(mean_kl), "lp[0,:4]", lp[0, :4].tolist())


def gen_vision():
    """Golden vectors for the video front-end:
      * smart_resize / smart_nframes / frame indices from the reference's OWN vendored qwen-vl-utils
        (/root/reference/SpaceR-SG-RLVR/src/qwen-vl-utils/src/qwen_vl_utils/vision_process.py), over a grid of inputs;
      * pixel_values_videos of the real HF Qwen2VLVideoProcessor (transformers 5.5.0, do_resize=False) on small random
        uint8 videos (even and odd frame counts)."""
    sys.path.insert(0, "/root/reference/SpaceR-SG-RLVR/src/qwen-vl-utils/src")
    from qwen_vl_utils import vision_process as VP
    rng = random.Random(7)
    resize = []
    for _ in range(200):
        h, w = rng.randint(28, 2200), rng.randint(28, 2200)
        for mn, mx in ((VP.VIDEO_MIN_PIXELS, VP.VIDEO_MAX_PIXELS), (VP.MIN_PIXELS, VP.MAX_PIXELS), (4 * 28 * 28, 448 * 28 * 28)):
            try:
                out = list(VP.smart_resize(h, w, factor=VP.IMAGE_FACTOR, min_pixels=mn, max_pixels=mx))
            except ValueError:
                out = "ValueError"
            resize.append([h, w, mn, mx, out])
    nframes = []
    for _ in range(200):
        total = rng.randint(2, 4000)
        fps_v = rng.choice([23.976, 24, 25, 29.97, 30, 60, 15, 10.5])
        ele = rng.choice([{}, {"fps": 1.0}, {"fps": 4.0}, {"nframes": rng.randint(2, 40)}, {"min_frames": 6, "max_frames": 12},
                          {"fps": 2.0, "max_frames": 32}])
        try:
            n = VP.smart_nframes(dict(ele), total_frames=total, video_fps=fps_v)
            idx = torch.linspace(0, total - 1, n).round().long().tolist()
        except (ValueError, AssertionError) as e:
            n, idx = type(e).__name__, None
        nframes.append([ele, total, fps_v, n, idx])
    with open(os.path.join(OUT, "vision.json"), "w") as f:
        json.dump({"source": "qwen_vl_utils.vision_process (vendored in the reference)", "smart_resize": resize,
                   "smart_nframes": nframes}, f)
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
    proc = Qwen2VLVideoProcessor(do_resize=False, do_sample_frames=False)
    cases = []
    g = torch.Generator().manual_seed(21)
    for F_, H, W in ((4, 56, 84), (3, 28, 56), (2, 112, 28)):
        video = torch.randint(0, 256, (F_, 3, H, W), generator=g, dtype=torch.uint8)
        out = proc(videos=[video], return_tensors="pt")
        cases.append({"video": video, "pixel_values_videos": out["pixel_values_videos"].clone(),
                      "video_grid_thw": out["video_grid_thw"].clone()})
    torch.save({"cases": cases, "transformers_version": __import__("transformers").__version__,
                "image_mean": list(proc.image_mean), "image_std": list(proc.image_std),
                "rescale_factor": float(proc.rescale_factor)}, os.path.join(OUT, "video_processor.pt"))
    print("vision.json:", len(resize), "resize cases,", len(nframes), "nframes cases; video_processor.pt:",
          [tuple(c["pixel_values_videos"].shape) for c in cases])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["rewards", "tiny", "tiny25", "vision"]
    if "tiny25" in which:
        gen_tiny_model25()
    if "vision" in which:
        gen_vision()
    if "rewards" in which:
        gen_rewards()
    if "tiny" in which:
        gen_tiny_model()

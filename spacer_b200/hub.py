"""Checkpoint I/O with the Hugging Face layout, so that the reference's construction and save calls keep working:
`Qwen2VLForConditionalGeneration.from_pretrained(model_id, ...)` / `Qwen2_5_VLForConditionalGeneration.from_pretrained`
(SG_RLVR_trainer.py:182-190, 205-214: the class is chosen by the substring "Qwen2-VL" / "Qwen2.5-VL" in the id) and
`trainer.save_model(output_dir)` (open_r1/SG-RLVR.py:384).  Local directories only (config.json + *.safetensors); there
is no hub download here.  Both parameter namings are accepted: transformers >= 4.52 / 5.x (`model.visual.*`,
`model.language_model.*`) and the older flat one the released checkpoints were saved with (`visual.*`, `model.layers.*`).
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

from .config import ModelDims
from .ops import SpacerError


def dims_from_hf_config(cfg: dict, name_hint: str = "") -> ModelDims:
    """ModelDims from a Qwen2-VL / Qwen2.5-VL config.json (nested `text_config` of 5.x or the flat 4.x layout)."""
    t = cfg.get("text_config") or cfg
    v = cfg.get("vision_config") or {}
    mtype = cfg.get("model_type", "")
    v25 = mtype == "qwen2_5_vl" or ("Qwen2.5-VL" in name_hint and mtype != "qwen2_vl")
    rope = t.get("rope_parameters") or t.get("rope_scaling") or cfg.get("rope_scaling") or {}
    hidden = t["hidden_size"]
    heads = t["num_attention_heads"]
    kw = dict(
        hidden=hidden, layers=t["num_hidden_layers"], heads=heads, kv_heads=t.get("num_key_value_heads", heads),
        head_dim=t.get("head_dim") or hidden // heads, inter=t["intermediate_size"], vocab=t["vocab_size"],
        tie=bool(cfg.get("tie_word_embeddings", t.get("tie_word_embeddings", False))),
        rms_eps=float(t.get("rms_norm_eps", 1e-6)),
        rope_theta=float(rope.get("rope_theta", t.get("rope_theta", 1e6))),
        mrope_section=tuple(rope.get("mrope_section", (16, 24, 24))),
        v_depth=v.get("depth", 32), v_heads=v.get("num_heads", 16), patch=v.get("patch_size", 14),
        t_patch=v.get("temporal_patch_size", 2), merge=v.get("spatial_merge_size", 2),
        in_ch=v.get("in_channels", v.get("in_chans", 3)),
        image_token_id=cfg.get("image_token_id", 151655), video_token_id=cfg.get("video_token_id", 151656),
        vision_start_id=cfg.get("vision_start_token_id", 151652), vision_end_id=cfg.get("vision_end_token_id", 151653),
        eos_id=_first(cfg.get("eos_token_id", t.get("eos_token_id", 151645))),
        pad_id=_first(cfg.get("pad_token_id", t.get("pad_token_id", 151643))),
        name=os.path.basename(os.path.normpath(name_hint)) or ("Qwen2.5-VL" if v25 else "Qwen2-VL"),
    )
    if kw["pad_id"] is None:
        kw["pad_id"] = 151643
    if v25:
        kw.update(variant="qwen2_5_vl", v_embed=v.get("hidden_size", 1280), v_mlp=v.get("intermediate_size", 3420),
                  v_window=v.get("window_size", 112), v_fullatt=tuple(v.get("fullatt_block_indexes", (7, 15, 23, 31))),
                  tokens_per_second=v.get("tokens_per_second", 2))
    else:
        e = v.get("embed_dim", 1280)
        kw.update(v_embed=e, v_mlp=int(e * v.get("mlp_ratio", 4)))
    return ModelDims(**kw)


def _first(x):
    return x[0] if isinstance(x, (list, tuple)) else x


def hf_config_from_dims(d: ModelDims) -> dict:
    """config.json in the nested transformers 5.x layout."""
    text = dict(hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                num_key_value_heads=d.kv_heads, intermediate_size=d.inter, vocab_size=d.vocab, rms_norm_eps=d.rms_eps,
                rope_parameters={"rope_type": "default", "rope_theta": d.rope_theta, "mrope_section": list(d.mrope_section)},
                tie_word_embeddings=d.tie, eos_token_id=d.eos_id, pad_token_id=d.pad_id, bos_token_id=d.pad_id)
    if d.variant == "qwen2_5_vl":
        vision = dict(depth=d.v_depth, hidden_size=d.v_embed, intermediate_size=d.v_mlp, num_heads=d.v_heads,
                      out_hidden_size=d.hidden, patch_size=d.patch, temporal_patch_size=d.t_patch,
                      spatial_merge_size=d.merge, in_channels=d.in_ch, hidden_act="silu", window_size=d.v_window,
                      fullatt_block_indexes=list(d.v_fullatt), tokens_per_second=d.tokens_per_second)
        mtype, arch = "qwen2_5_vl", "Qwen2_5_VLForConditionalGeneration"
    else:
        vision = dict(depth=d.v_depth, embed_dim=d.v_embed, hidden_size=d.hidden, num_heads=d.v_heads,
                      mlp_ratio=d.v_mlp // d.v_embed, patch_size=d.patch, temporal_patch_size=d.t_patch,
                      spatial_merge_size=d.merge, in_channels=d.in_ch, hidden_act="quick_gelu")
        mtype, arch = "qwen2_vl", "Qwen2VLForConditionalGeneration"
    return dict(model_type=mtype, architectures=[arch], text_config=text, vision_config=vision,
                image_token_id=d.image_token_id, video_token_id=d.video_token_id,
                vision_start_token_id=d.vision_start_id, vision_end_token_id=d.vision_end_id,
                tie_word_embeddings=d.tie, torch_dtype="bfloat16")


def normalize_names(sd: dict) -> dict:
    """Map the pre-4.52 parameter names (`visual.*`, `model.layers.*`, ...) to the 5.x ones this package speaks."""
    out = {}
    for k, v in sd.items():
        if k.startswith("visual."):
            k = "model." + k
        elif k.startswith("model.") and not k.startswith(("model.visual.", "model.language_model.")):
            k = "model.language_model." + k[len("model."):]
        out[k] = v
    return out


def load_checkpoint_tensors(path: str) -> dict:
    from safetensors import safe_open
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise SpacerError(f"from_pretrained: no *.safetensors under {path} (only local HF-format directories are supported)")
    sd = {}
    for fn in files:
        with safe_open(os.path.join(path, fn), framework="pt", device="cpu") as f:
            for k in f.keys():
                sd[k] = f.get_tensor(k)
    return normalize_names(sd)


def from_pretrained(cls, path: str, device="cuda", rope_convention: str = "classic", **unused):
    """Build the engine from a local HF checkpoint directory.  `unused` swallows the HF kwargs the reference passes
    (`attn_implementation`, `torch_dtype`, `use_cache`: SG_RLVR_trainer.py:163-190)."""
    if not os.path.isdir(path):
        raise SpacerError(f"from_pretrained: {path} is not a local directory (no network access / hub download here)")
    with open(os.path.join(path, "config.json")) as f:
        cfg = json.load(f)
    dims = dims_from_hf_config(cfg, name_hint=path)
    gen = {}
    gpath = os.path.join(path, "generation_config.json")
    if os.path.exists(gpath):       # the defaults model.generate() merges under the caller's options (utils.py:1693-1701)
        with open(gpath) as f:
            gen = {k: v for k, v in json.load(f).items() if not k.startswith("_") and k != "transformers_version"}
        eos = gen.get("eos_token_id")
        if isinstance(eos, (list, tuple)) and eos:
            from dataclasses import replace
            dims = replace(dims, eos_ids=tuple(int(e) for e in eos))
    m = cls(dims, device, rope_convention=rope_convention)
    m.generation_config = gen
    sd = load_checkpoint_tensors(path)
    if dims.tie and "lm_head.weight" not in sd:
        sd["lm_head.weight"] = sd["model.language_model.embed_tokens.weight"]
    m.load_state_dict(sd)
    m.config._name_or_path = path
    return m


def save_pretrained(model, path: str, max_shard_bytes: int = 5 << 30):
    """config.json + model-0000x-of-0000y.safetensors (+ index) with transformers 5.x parameter names, bf16."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config_from_dims(model.dims), f, indent=2)
    gen = dict(getattr(model, "generation_config", None) or {})
    if model.dims.eos_ids and "eos_token_id" not in gen:
        gen["eos_token_id"] = list(model.dims.eos_ids)
    if gen:
        with open(os.path.join(path, "generation_config.json"), "w") as f:
            json.dump(gen, f, indent=2)
    shards, cur, size = [], {}, 0
    for k, v in model.params.hf_items():
        if model.dims.tie and k == "lm_head.weight":
            continue
        t = v.detach().to("cpu").contiguous().clone()
        n = t.numel() * t.element_size()
        if cur and size + n > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = t
        size += n
    if cur:
        shards.append(cur)
    index = {}
    for i, sh in enumerate(shards):
        fn = "model.safetensors" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(path, fn), metadata={"format": "pt"})
        index.update({k: fn for k in sh})
    if len(shards) > 1:
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": index}, f, indent=1)


def make_config_namespace(dims: ModelDims, name_or_path: str = ""):
    """The few `model.config` attributes the reference trainer reads (`_name_or_path`, SG_RLVR_trainer.py:156, 193, 234)."""
    return SimpleNamespace(_name_or_path=name_or_path or dims.name, model_type=dims.variant, dims=dims,
                           **{k: v for k, v in hf_config_from_dims(dims).items() if k in ("architectures", "tie_word_embeddings")})

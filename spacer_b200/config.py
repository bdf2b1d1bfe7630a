"""Model dimensions of Qwen2-VL (transformers configuration_qwen2_vl.py)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ModelDims:
    # language model
    hidden: int = 3584
    layers: int = 28
    heads: int = 28
    kv_heads: int = 4
    head_dim: int = 128
    inter: int = 18944
    vocab: int = 152064
    tie: bool = False
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)
    # vision tower
    v_depth: int = 32
    v_embed: int = 1280
    v_heads: int = 16
    v_mlp: int = 5120
    patch: int = 14
    t_patch: int = 2
    merge: int = 2
    in_ch: int = 3
    # special token ids (configuration_qwen2_vl.py:159-162)
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_id: int = 151652
    vision_end_id: int = 151653
    eos_id: int = 151645
    pad_id: int = 151643
    name: str = "Qwen2-VL-7B"

    @property
    def v_head_dim(self) -> int:
        return self.v_embed // self.v_heads

    @property
    def patch_dim(self) -> int:
        return self.in_ch * self.t_patch * self.patch * self.patch

    @property
    def qkv_dim(self) -> int:
        return (self.heads + 2 * self.kv_heads) * self.head_dim

    @property
    def merge_dim(self) -> int:
        return self.v_embed * self.merge * self.merge


def qwen2_vl_7b() -> ModelDims:
    return ModelDims()


def qwen2_vl_2b() -> ModelDims:
    return ModelDims(hidden=1536, layers=28, heads=12, kv_heads=2, inter=8960, vocab=151936, tie=True,
                     name="Qwen2-VL-2B")


def tiny(layers: int = 2, v_depth: int = 2) -> ModelDims:
    """Structurally faithful miniature (head_dim 128, ViT head_dim 80) used by the parity tests."""
    return ModelDims(hidden=256, layers=layers, heads=2, kv_heads=1, inter=512, vocab=2048, tie=False,
                     v_depth=v_depth, v_embed=160, v_heads=2, v_mlp=640,
                     image_token_id=2039, video_token_id=2040, vision_start_id=2036, vision_end_id=2037,
                     eos_id=2029, pad_id=2027, name="tiny")


PRESETS = {"7b": qwen2_vl_7b, "2b": qwen2_vl_2b, "tiny": tiny}

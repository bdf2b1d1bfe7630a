"""Model dimensions of Qwen2-VL (transformers configuration_qwen2_vl.py) and of Qwen2.5-VL
(configuration_qwen2_5_vl.py; SURVEY.md 8(f) row 1 -- the family run_SpaceR_SG_RLVR.sh:16 trains)."""
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
    eos_id: int = 151645                  # the tokenizer's eos (<|im_end|>): what the completion mask looks for (TRN:489)
    pad_id: int = 151643
    eos_ids: tuple = ()                   # every id that ends a rollout row (generation_config.json: [151645, 151643]);
                                          # empty = (eos_id,)
    name: str = "Qwen2-VL-7B"
    # family: "qwen2_vl" (LayerNorm + QuickGELU ViT, per-frame attention) or "qwen2_5_vl" (RMSNorm + SwiGLU ViT with
    # windowed attention, modeling_qwen2_5_vl.py:345-520).  For qwen2_5_vl `v_mlp` is the ViT intermediate size (3420).
    variant: str = "qwen2_vl"
    v_window: int = 112                   # window side in pixels (4 x 4 merged tokens at patch 14, merge 2)
    v_fullatt: tuple = (7, 15, 23, 31)    # blocks with full (per-frame) attention
    tokens_per_second: int = 2            # temporal M-RoPE spacing (get_rope_index)

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

    @property
    def v_mlp_pad(self) -> int:
        """ViT SwiGLU width padded to the 64-column interleave of the fused gate|up GEMM (zero rows / columns)."""
        return (self.v_mlp + 63) // 64 * 64


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


def qwen2_5_vl_7b() -> ModelDims:
    return ModelDims(variant="qwen2_5_vl", v_mlp=3420, name="Qwen2.5-VL-7B")


def qwen2_5_vl_3b() -> ModelDims:
    return ModelDims(variant="qwen2_5_vl", v_mlp=3420, hidden=2048, layers=36, heads=16, kv_heads=2, inter=11008,
                     vocab=151936, tie=True, name="Qwen2.5-VL-3B")


def tiny25(layers: int = 2, v_depth: int = 2) -> ModelDims:
    """Miniature Qwen2.5-VL: ViT SwiGLU width 200 (padded to 256 internally), block 0 windowed, block 1 full."""
    return ModelDims(hidden=256, layers=layers, heads=2, kv_heads=1, inter=512, vocab=2048, tie=False,
                     v_depth=v_depth, v_embed=160, v_heads=2, v_mlp=200, variant="qwen2_5_vl",
                     v_fullatt=(v_depth - 1,),
                     image_token_id=2039, video_token_id=2040, vision_start_id=2036, vision_end_id=2037,
                     eos_id=2029, pad_id=2027, name="tiny25")


PRESETS = {"7b": qwen2_vl_7b, "2b": qwen2_vl_2b, "tiny": tiny, "25-7b": qwen2_5_vl_7b, "25-3b": qwen2_5_vl_3b,
           "tiny25": tiny25}

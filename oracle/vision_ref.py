"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-torch restatement of `Qwen2VLVideoProcessor._preprocess` without the resize
(transformers 5.5.0, models/qwen2_vl/video_processing_qwen2_vl.py:240-272, and the fused rescale+normalise of
image_processing_backends.py:292-333).  Parity is PINNED: tests/test_vision_cpu.py compares it bit for bit with
tests/golden/video_processor.pt, produced by the real HF processor (oracle/make_golden.py vision).
"""
from __future__ import annotations

import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def patchify_ref(frames: torch.Tensor, patch=14, t_patch=2, merge=2, rescale_factor=1 / 255, mean=CLIP_MEAN, std=CLIP_STD):
    """frames [F, C, H, W] uint8 or float (0..255) -> (pixel_values fp32 [gt*gh*gw, C*t_patch*patch*patch], grid)."""
    m = torch.tensor(mean) * (1.0 / rescale_factor)                  # :300-303 fused constants, fp32
    s = torch.tensor(std) * (1.0 / rescale_factor)
    x = frames.to(torch.float32)
    x = (x - m.view(1, -1, 1, 1)) / s.view(1, -1, 1, 1)              # tvF.normalize: sub then div
    F_, C, H, W = x.shape
    if pad := -F_ % t_patch:                                          # :246-249 repeat the last frame
        x = torch.cat([x, x[-1:].expand(pad, -1, -1, -1)], 0)
    gt, gh, gw = x.shape[0] // t_patch, H // patch, W // patch
    x = x.view(gt, t_patch, C, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)                          # :266 (batch dim dropped)
    return x.reshape(gt * gh * gw, C * t_patch * patch * patch).contiguous(), (gt, gh, gw)

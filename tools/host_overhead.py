#!/usr/bin/env python
"""How much HOST time does one policy forward/backward (and one reference scoring) cost to enqueue?  Same number of layers
and launches as the 7B model (28 decoder layers, 32 ViT blocks), tiny widths so that the GPU work is negligible: the wall
time of the call IS the host cost.    python tools/host_overhead.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from spacer_b200 import config, ops  # noqa: E402
from spacer_b200.model import GradStore, Qwen2VLB200, pack_prompt_completions  # noqa: E402


def main():
    d = config.tiny(28, 32)
    m = Qwen2VLB200(d, "cuda")
    m.params.init_random(0)
    g = torch.Generator().manual_seed(0)
    grid = torch.tensor([[2, 8, 8]])
    pix = torch.randn(128, d.patch_dim, generator=g).cuda()
    n_v = 32
    prompt = torch.cat([torch.randint(10, 2000, (6,), generator=g), torch.tensor([d.vision_start_id]),
                        torch.full((n_v,), d.video_token_id), torch.tensor([d.vision_end_id]),
                        torch.randint(10, 2000, (10,), generator=g)])[None]
    comp = torch.randint(10, 2000, (8, 16), generator=g)
    batch = pack_prompt_completions(prompt, comp, grid, d, m.device)
    grads = GradStore(m.params)
    adv = torch.linspace(-1, 1, 8)
    ref = torch.zeros(8, 16) - 7.0
    res = {}
    for name, fn in (("policy_fwd_bwd", lambda: m.grpo_forward_backward(batch, pix, grid, ref, adv, 0.04, grads)),
                     ("ref_scoring", lambda: m.per_token_logps(batch, pix, grid))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ops.reset_launch_count()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            fn()
        t_enq = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / n * 1e3
        res[name] = {"host_enqueue_ms": round(t_enq, 1), "wall_ms": round(t_all, 1), "launches": ops.launch_count() // n}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run one phase of the hot path between cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees only it.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_decode.csv python tools/profile_phases.py --phase decode --config c3
phases: decode (3 un-graphed decode steps at mid-rollout context), train (policy fwd/bwd + loss), ref (scoring fwd),
        prefill (ViT + prompt prefill), adamw
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

import bench  # noqa: E402
from spacer_b200 import config as mcfg  # noqa: E402
from spacer_b200.model import GradStore, Qwen2VLB200, SamplingParams, pack_prompt_completions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", default="decode")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    dims = mcfg.PRESETS[cfg["preset"]]()
    dev = torch.device("cuda", 0)
    m = Qwen2VLB200(dims, dev)
    m.params.init_random(seed=0)
    ex = bench.synth_example(dims, cfg, 1234)
    pix = ex["pixel_values_host"].to(dev)
    grid, ids = ex["video_grid_thw"], ex["input_ids"]
    G, C = cfg["G"], cfg["C"]
    prof = torch.cuda.profiler
    if a.phase == "decode":
        # un-graphed so every kernel is its own launch; context = prompt + half the completions
        pix2 = pix.flip(0).contiguous()
        n = C // 2
        m.generate(ids, pix, grid, max_new_tokens=n, num_return_sequences=G, pixel_values_videos_2=pix2,
                   num_return_sequences_2=G // 2, min_new_tokens=n, use_graph=False, seed=1)
        st = m._last_decode_state
        torch.cuda.synchronize()
        st["step"].fill_(n - 2)
        pos, nxt = __import__("spacer_b200.model", fromlist=["rope_index"]).rope_index(ids.reshape(-1), grid, dims)
        prof.start()
        for _ in range(a.steps):
            m._decode_step(st, nxt, G, SamplingParams(top_p=0.95, top_k=50, eos_ids=(dims.eos_id,), pad_id=dims.pad_id), True)
            st["step"].fill_(n - 2)
        torch.cuda.synchronize()
        prof.stop()
        return
    g = torch.Generator().manual_seed(5)
    comp = torch.randint(1000, 100000, (G, C), generator=g)
    batch = pack_prompt_completions(ids, comp, grid, dims, dev)
    if a.phase == "prefill":
        m.generate(ids, pix, grid, max_new_tokens=1, num_return_sequences=G, min_new_tokens=1, seed=1)
        torch.cuda.synchronize()
        prof.start()
        m.generate(ids, pix, grid, max_new_tokens=1, num_return_sequences=G, min_new_tokens=1, seed=1)
        torch.cuda.synchronize()
        prof.stop()
        return
    if a.phase == "ref":
        m.per_token_logps(batch, pix, grid)
        torch.cuda.synchronize()
        prof.start()
        m.per_token_logps(batch, pix, grid)
        torch.cuda.synchronize()
        prof.stop()
        return
    if a.phase == "train":
        grads = GradStore(m.params)
        adv = torch.linspace(-1, 1, G)
        ref = torch.zeros(G, C) - 11.0
        m.grpo_forward_backward(batch, pix, grid, ref, adv, 0.04, grads)
        torch.cuda.synchronize()
        prof.start()
        m.grpo_forward_backward(batch, pix, grid, ref, adv, 0.04, grads)
        torch.cuda.synchronize()
        prof.stop()
        return
    if a.phase == "train_timing":
        # sub-phase CUDA-event times of the policy forward/backward vs the host time spent enqueueing it
        import json
        import time
        grads = GradStore(m.params)
        adv = torch.linspace(-1, 1, G)
        ref = torch.zeros(G, C) - 11.0
        for _ in range(2):
            m.grpo_forward_backward(batch, pix, grid, ref, adv, 0.04, grads)
        torch.cuda.synchronize()
        res = []
        for _ in range(3):
            m.phase_marks = []
            t0 = time.perf_counter()
            m.grpo_forward_backward(batch, pix, grid, ref, adv, 0.04, grads)
            t_enq = (time.perf_counter() - t0) * 1e3
            torch.cuda.synchronize()
            t_all = (time.perf_counter() - t0) * 1e3
            mk = m.phase_marks
            res.append({"host_enqueue_ms": round(t_enq, 1), "host_total_ms": round(t_all, 1),
                        **{mk[i][0] + "_ms": round(mk[i - 1][1].elapsed_time(mk[i][1]), 2) for i in range(1, len(mk))},
                        "gpu_total_ms": round(mk[0][1].elapsed_time(mk[-1][1]), 2)})
        m.phase_marks = None
        print(json.dumps(res))
        return
    raise SystemExit(f"unknown phase {a.phase}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B timing of the decode step on one B200 (cfg3 shapes): PDL on/off, L2 weight prefetch size, per-GEMV cold/warm/hot.
Prints one JSON line per experiment.   python tools/decode_lab.py [--config c3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

import bench  # noqa: E402
from spacer_b200 import config as mcfg, ops  # noqa: E402
from spacer_b200.model import Qwen2VLB200, rope_index  # noqa: E402


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--exp", default="step,gemv")
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
    pix2 = pix.flip(0).contiguous()
    lib = ops._lib.load()
    m.generate(ids, pix, grid, max_new_tokens=C, num_return_sequences=G, pixel_values_videos_2=pix2,
               num_return_sequences_2=G // 2, min_new_tokens=C, use_graph=False, seed=1, **{"max_steps_debug": 0}) if False else None
    # build the decode state with a short rollout (full prefill, 3 decode steps)
    st = m._decode_state(G + G // 2, ids.numel(), C, 2)
    m.generate(ids, pix, grid, max_new_tokens=C, num_return_sequences=G, pixel_values_videos_2=pix2,
               num_return_sequences_2=G // 2, min_new_tokens=C, seed=1) if "full" in a.exp else None
    _, nxt = rope_index(ids.reshape(-1), grid, dims)
    wbytes = m.decode_weight_bytes()
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

    if "step" in a.exp:
        for pdl in (1, 0):
            for pf in (48 << 20, 0, 16 << 20, 96 << 20):
                if pdl == 0 and pf not in (0, 48 << 20):
                    continue
                lib.sb_set_pdl(pdl)
                m.L2_PREFETCH_BYTES = pf
                st["graphs"].clear()
                st["step"].fill_(200)
                graph, nodes = m._decode_graph(st, nxt, G, 0.95, True)
                st["step"].fill_(200)
                for _ in range(5):
                    graph.replay()
                st["step"].fill_(200)
                ms = timed(graph.replay, 100)
                kv = 2 * dims.layers * dims.kv_heads * dims.head_dim * 2
                byts = wbytes + 2 * ids.numel() * kv + (G + G // 2) * 250 * kv
                print(json.dumps({"exp": "decode_step_graph", "pdl": pdl, "l2_prefetch_mb": pf >> 20, "ms": round(ms, 4),
                                  "gbs": round(byts / ms / 1e6, 1), "nodes": nodes}), flush=True)
        lib.sb_set_pdl(1)
        m.L2_PREFETCH_BYTES = 48 << 20

    if "gemv" in a.exp:
        W = m.params
        S = st["S"]
        shapes = [("qkv", W["l.5.qkv_w"], st["xn"], st["p_qkv"], S["qkv"]), ("o", W["l.5.o_w"], st["attn"], st["p_o"], S["o"]),
                  ("gu", W["l.5.gu_w"], st["xn"], st["p_gu"], S["gu"]), ("down", W["l.5.down_w"], st["act"], st["p_down"], S["down"])]
        dummy_w = W["l.20.o_w"]
        for name, w, x, out, s in shapes:
            nbytes = w.numel() * 2
            res = {"exp": "gemv", "name": name, "mb": round(nbytes / 1e6, 1), "splits": s}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def one(mode):
                tot = 0.0
                n = 10
                for _ in range(n):
                    flush.fill_(1)                       # evict L2
                    if mode == "warm":                    # a carrier GEMV prefetches w into L2, a small kernel gives it time
                        m._gemv(dummy_w, st["attn"], st["p_o"], S["o"], next_w=w)
                        for _ in range(4):
                            ops.call("sb_dec_swiglu", st["p_gu"], S["gu"], st["RP"] * 2 * dims.inter, 2 * dims.inter, st["act"], st["R"], dims.inter)
                    elif mode == "hot":
                        m._gemv(w, x, out, s)
                    elif mode == "cold_after_small":
                        ops.call("sb_dec_swiglu", st["p_gu"], S["gu"], st["RP"] * 2 * dims.inter, 2 * dims.inter, st["act"], st["R"], dims.inter)
                    e0.record()
                    m._gemv(w, x, out, s)
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                return tot / n * 1e3
            for mode in ("cold", "cold_after_small", "warm", "hot"):
                us = one(mode)
                res[mode + "_us"] = round(us, 2)
                res[mode + "_gbs"] = round(nbytes / us / 1e3, 1)
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

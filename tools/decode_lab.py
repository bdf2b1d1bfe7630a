#!/usr/bin/env python
"""A/B timing of the decode step on one B200 (cfg3 shapes): PDL on/off, L2 weight prefetch size, per-GEMV cold/warm/hot.
Prints one JSON line per experiment.   python tools/decode_lab.py [--config c3]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

import bench  # noqa: E402
from spacer_b200 import config as mcfg, ops  # noqa: E402
from spacer_b200.model import Qwen2VLB200, SamplingParams, rope_index  # noqa: E402


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
    SP = SamplingParams(top_p=0.95, top_k=50, eos_ids=(dims.eos_id,), pad_id=dims.pad_id)   # the trainer's rollout options
    if "ab" in a.exp:
        # the decode step as bench.py runs it (graph replay at step 200) + the 113 GEMVs alone
        _, nxt = rope_index(ids.reshape(-1), grid, dims)
        kv = 2 * dims.layers * dims.kv_heads * dims.head_dim * 2
        byts = m.decode_weight_bytes() + 2 * ids.numel() * kv + (G + G // 2) * 250 * kv
        for _ in range(2):
            m._dec = None
            st = m._decode_state(G + G // 2, ids.numel(), C, 2)
            st["step"].fill_(200)
            graph, nodes = m._decode_graph(st, nxt, G, SP, True)
            st["step"].fill_(200)
            for _ in range(5):
                graph.replay()
            st["step"].fill_(200)
            ms = timed(graph.replay, 100)
            prof = m.profile_decode_gemv(rows=G + G // 2, reps=3)
            print(json.dumps({"exp": "ab", "splits": st["S"], "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1),
                              "nodes": nodes, "gemv_only_gbs": round(prof["gbs"], 1), "gemv_sweep_ms": round(prof["ms_per_sweep"], 4)}), flush=True)
    m._dec = None
    # build the decode state with a short rollout (full prefill, 3 decode steps)
    st = m._decode_state(G + G // 2, ids.numel(), C, 2)
    m.generate(ids, pix, grid, max_new_tokens=C, num_return_sequences=G, pixel_values_videos_2=pix2,
               num_return_sequences_2=G // 2, min_new_tokens=C, seed=1) if "full" in a.exp else None
    _, nxt = rope_index(ids.reshape(-1), grid, dims)
    wbytes = m.decode_weight_bytes()
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

    if "step" in a.exp:
        for pdl, pf, pfa in ((1, 16, 48), (1, 0, 0), (1, 16, 72), (1, 16, 96), (1, 16, 120), (1, 32, 96), (1, 48, 96), (1, 0, 96), (1, 16, 48)):
            if True:
                pf <<= 20
                lib.sb_set_pdl(pdl)
                m.L2_PREFETCH_BYTES = pf
                m.L2_PREFETCH_ATTN_BYTES = pfa << 20
                st["graphs"].clear()
                st["step"].fill_(200)
                graph, nodes = m._decode_graph(st, nxt, G, SP, True)
                st["step"].fill_(200)
                for _ in range(5):
                    graph.replay()
                st["step"].fill_(200)
                ms = timed(graph.replay, 100)
                kv = 2 * dims.layers * dims.kv_heads * dims.head_dim * 2
                byts = wbytes + 2 * ids.numel() * kv + (G + G // 2) * 250 * kv
                print(json.dumps({"exp": "decode_step_graph", "pdl": pdl, "l2_prefetch_mb": pf >> 20, "l2_prefetch_attn_mb": pfa, "ms": round(ms, 4),
                                  "gbs": round(byts / ms / 1e6, 1), "nodes": nodes}), flush=True)
        lib.sb_set_pdl(1)
        m.L2_PREFETCH_BYTES = 16 << 20
        m.L2_PREFETCH_ATTN_BYTES = 48 << 20

    for trace_cfg in ([(16, 48), (0, 0)] if "trace" in a.exp else []):
        m.L2_PREFETCH_BYTES, m.L2_PREFETCH_ATTN_BYTES = trace_cfg[0] << 20, trace_cfg[1] << 20
        KINDS = {1: "gemv", 2: "embed", 3: "rmsnorm", 4: "qkv_post", 5: "attn", 6: "combine", 7: "swiglu", 8: "sample", 9: "advance"}
        cap = 4096
        buf = torch.zeros(1 + 4 * cap, device=dev, dtype=torch.int64)
        st["graphs"].clear()
        st["step"].fill_(200)
        graph, nodes = m._decode_graph(st, nxt, G, SP, True)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        lib.sb_trace_enable(ctypes.c_void_p(buf.data_ptr()), cap)
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        lib.sb_trace_enable(None, 0)
        h = buf.cpu().tolist()
        n = min(h[0], cap)
        recs = [(h[1 + 4 * k], h[2 + 4 * k], h[3 + 4 * k], h[4 + 4 * k]) for k in range(n)]
        recs.sort(key=lambda r: r[1])
        # second replay only (steady state): records after the first sampler
        idx = [i for i, r in enumerate(recs) if r[0] == 8]
        one = recs[idx[0] + 1: idx[1] + 1] if len(idx) >= 2 else recs
        t0 = one[0][1]
        tag = "r02"
        out_path = os.path.join(ROOT, "gpurun_out", f"decode_trace_{tag}_pf{trace_cfg[0]}_{trace_cfg[1]}.txt")
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        agg = {}
        with open(out_path, "w") as f:
            f.write("# kind entry_us ready_us end_us | wait_us exec_us gap_from_prev_end_us   (block 0 of every kernel, globaltimer)\n")
            prev_end = None
            for kind, te, tr_, tn in one:
                name = KINDS.get(kind, str(kind))
                gap = (tr_ - prev_end) / 1e3 if prev_end else 0.0
                f.write(f"{name:9s} {(te - t0) / 1e3:9.2f} {(tr_ - t0) / 1e3:9.2f} {(tn - t0) / 1e3:9.2f} | {(tr_ - te) / 1e3:7.2f} {(tn - tr_) / 1e3:7.2f} {gap:7.2f}\n")
                d = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
                d[0] += 1; d[1] += (tr_ - te) / 1e3; d[2] += (tn - tr_) / 1e3; d[3] += gap
                prev_end = tn
        total = (one[-1][3] - t0) / 1e3
        print(json.dumps({"exp": "trace", "splits": st["S"], "prefetch_mb": trace_cfg, "records": n, "step_kernels": len(one), "step_us": round(total, 1),
                          "per_kind": {k: {"n": v[0], "wait_us": round(v[1] / v[0], 2), "exec_us": round(v[2] / v[0], 2),
                                           "gap_after_prev_end_us": round(v[3] / v[0], 2), "exec_total_us": round(v[2], 1),
                                           "gap_total_us": round(v[3], 1)} for k, v in agg.items()}}), flush=True)

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

#!/usr/bin/env python
"""A/B timing of the decode step on one B200 (cfg3 shapes): L2 plan (hints, paced prefetch variants), kernel timeline trace,
per-CTA GEMV timeline (--exp skew), per-GEMV cold/warm/hot.
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
    ap.add_argument("--plans", default="", help="hints,gu_rows,after;... (overrides the built-in sweep)")
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

    def set_plan(hints, gu_rows, after, ctas, pace=-1):
        lib.sb_set_dec_l2_hints(hints)
        m.PF_GU_ROWS, m.PF_AFTER, m.PF_CTAS, m.PF_PACE_NS = gu_rows, after, ctas, pace

    default_plan = (1, m.PF_GU_ROWS, m.PF_AFTER, m.PF_CTAS, m.PF_PACE_NS)
    if "step" in a.exp:
        plans = [default_plan, (1, 0, "attn", 0, -1), (1, 24, "combine", 0, -1)]
        for after in ("combine", "qkv_post"):
            for pace in (0, 200, 500, 1000):
                for rows in (24, 32):
                    plans.append((1, rows, after, 0, pace))
        plans += [(1, 32, "qkv_post", 74, 500), (1, 32, "qkv_post", 37, 200), (1, 40, "qkv_post", 0, 500), (1, 0, "attn", 0, -1), default_plan]
        if a.plans:
            plans = [tuple(int(v) if i != 2 else v for i, v in enumerate(pl.split(","))) for pl in a.plans.split(";")]
        for plan in plans:
            set_plan(*plan)
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
            print(json.dumps({"exp": "decode_step_graph", "l2_hints": plan[0], "pf_gu_rows": plan[1], "pf_after": plan[2], "pf_ctas": plan[3], "pf_pace_ns": plan[4], "ms": round(ms, 4),
                              "gbs": round(byts / ms / 1e6, 1), "nodes": nodes}), flush=True)
        set_plan(*default_plan)

    for trace_cfg in ([default_plan, (1, 0, "attn", 0, -1)] if "trace" in a.exp else []):
        set_plan(*trace_cfg)
        KINDS = {1: "gemv", 2: "embed", 3: "rmsnorm", 4: "qkv_post", 5: "attn", 6: "combine", 7: "swiglu", 8: "sample", 9: "advance"}
        cap = 8192
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
        out_path = os.path.join(ROOT, "gpurun_out", f"decode_trace_{tag}_plan{'_'.join(str(v) for v in trace_cfg)}.txt")
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

    if "skew" in a.exp:
        # spread of the finishing times of the CTAs of every decode GEMV (the kernel-level trace only sees block 0)
        import statistics
        if a.plans:
            pl = a.plans.split(";")[0].split(",")
            set_plan(int(pl[0]), int(pl[1]), pl[2], int(pl[3]), int(pl[4]))
        cap = 260
        n_sms = 148
        buf = torch.zeros(n_sms * (1 + 3 * cap), device=dev, dtype=torch.int64)
        st["graphs"].clear()
        st["step"].fill_(200)
        graph, nodes = m._decode_graph(st, nxt, G, SP, True)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        lib.sb_trace_enable_gemv_ctas(ctypes.c_void_p(buf.data_ptr()), cap)
        graph.replay()
        torch.cuda.synchronize()
        lib.sb_trace_enable_gemv_ctas(None, 0)
        h = buf.cpu().view(n_sms, 1 + 3 * cap).tolist()
        names = {4608 ^ (3584 << 32): "qkv", 3584 ^ (3584 << 32): "o", 37888 ^ (3584 << 32): "gu", 3584 ^ (18944 << 32): "down",
                 152064 ^ (3584 << 32): "lm_head"}
        per = {}           # (name, occurrence) -> list of (block, ready, end)
        for b in range(n_sms):
            seen = {}
            for k in range(int(h[b][0])):
                tag, tr_, te = h[b][1 + 3 * k], h[b][2 + 3 * k], h[b][3 + 3 * k]
                nm = names.get(tag, hex(tag))
                occ = seen.get(nm, 0)
                seen[nm] = occ + 1
                per.setdefault((nm, occ), []).append((b, tr_, te))
        agg = {}
        for (nm, occ), lst in per.items():
            ends = sorted(e for _, _, e in lst)
            readies = sorted(r for _, r, _ in lst)
            e_b0 = [e for b, _, e in lst if b == 0]
            d = agg.setdefault(nm, {"n": 0, "ctas": len(lst), "dur_first_ready_to_last_end": [], "end_max_minus_median": [], "end_max_minus_min": [],
                                    "end_max_minus_block0": [], "ready_spread": []})
            d["n"] += 1
            d["dur_first_ready_to_last_end"].append((ends[-1] - readies[0]) / 1e3)
            d["end_max_minus_median"].append((ends[-1] - ends[len(ends) // 2]) / 1e3)
            d["end_max_minus_min"].append((ends[-1] - ends[0]) / 1e3)
            d["ready_spread"].append((readies[-1] - readies[0]) / 1e3)
            if e_b0:
                d["end_max_minus_block0"].append((ends[-1] - e_b0[0]) / 1e3)
        for nm, d in agg.items():
            out = {"exp": "gemv_cta_skew", "gemv": nm, "instances": d["n"], "ctas": d["ctas"]}
            for k in ("dur_first_ready_to_last_end", "end_max_minus_median", "end_max_minus_min", "end_max_minus_block0", "ready_spread"):
                out[k + "_us"] = round(statistics.median(d[k]), 2) if d[k] else None
            print(json.dumps(out), flush=True)
        for nm in ("gu", "qkv", "down", "o"):
            lst = sorted(per.get((nm, 5), []))
            if lst:
                t0 = min(r for _, r, _ in lst)
                late = [[b, round((r - t0) / 1e3, 2), round((e - t0) / 1e3, 2)] for b, r, e in lst if r - t0 > 300]
                print(json.dumps({"exp": "gemv_cta_skew", "gemv": nm, "layer": 5, "late_blocks_ready_end_us": late,
                                  "median_end_us": round(statistics.median(e - t0 for _, _, e in lst) / 1e3, 2)}), flush=True)
        # which SMs finish last in the gate|up GEMV (stable across layers = a property of the SM, not of the data)
        late = {}
        for (nm, occ), lst in per.items():
            if nm != "gu":
                continue
            for b, _, e in sorted(lst, key=lambda x: -x[2])[:15]:
                late[b] = late.get(b, 0) + 1
        print(json.dumps({"exp": "gemv_cta_skew", "gu_blocks_among_15_latest": sorted(late.items(), key=lambda kv: -kv[1])[:20]}), flush=True)

    if "gemv" in a.exp:
        W = m.params
        S = st["S"]
        shapes = [("qkv", W["l.5.qkv_w"], st["xn"], st["p_qkv"], S["qkv"]), ("o", W["l.5.o_w"], st["attn"], st["p_o"], S["o"]),
                  ("gu", W["l.5.gu_w"], st["xn"], st["act"] if S["gu"] == 1 else st["p_gu"], S["gu"]),
                  ("down", W["l.5.down_w"], st["act"], st["p_down"], S["down"])]
        for name, w, x, out, s in shapes:
            nbytes = w.numel() * 2
            res = {"exp": "gemv", "name": name, "mb": round(nbytes / 1e6, 1), "splits": s}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def one(mode):
                tot = 0.0
                n = 10
                for _ in range(n):
                    flush.fill_(1)                       # evict L2
                    if mode == "warm":                    # the paced prefetch kernel warms 24 rows per tile of w, then a pause
                        ops.call("sb_dec_l2_prefetch", w, 24 * w.shape[1] * 2, 128 * w.shape[1] * 2, w.shape[0] // 128, 0, 750)
                        for _ in range(4):
                            ops.call("sb_dec_swiglu", st["p_gu"], S["gu"], st["RP"] * 2 * dims.inter, 2 * dims.inter, st["act"], st["R"], dims.inter)
                    elif mode == "hot":
                        m._gemv(w, x, out, s, swiglu=(name == "gu" and s == 1))
                    elif mode == "cold_after_small":
                        ops.call("sb_dec_swiglu", st["p_gu"], S["gu"], st["RP"] * 2 * dims.inter, 2 * dims.inter, st["act"], st["R"], dims.inter)
                    e0.record()
                    m._gemv(w, x, out, s, swiglu=(name == "gu" and s == 1))
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

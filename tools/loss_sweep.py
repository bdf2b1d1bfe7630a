#!/usr/bin/env python
"""cfg5 sweep of the fused lm_head + log-prob + GRPO-loss path: G in {8,16,32}, C in {256,512,1024},
V = 152064, H = 3584, beta = 0.04, random bf16 hidden states.  For every point prints one JSON line with
  lm_head GEMM + LMHEAD epilogue (tensor-bound)      TFLOP/s vs the measured bf16 peak
  loss tail (logsumexp fold + mask/KL/loss/coef)     GB/s on its algorithmic bytes vs the measured HBM peak
  backward: DLOGITS recompute GEMM + dH and dW GEMMs TFLOP/s
and checks the loss against a float64 torch evaluation of the same formulas on the kernel's log-probs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from spacer_b200 import ops  # noqa: E402
from spacer_b200.ops import EPI_DLOGITS, EPI_LMHEAD  # noqa: E402


def ev_time(fn, reps=5, flush=None):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--G", type=int, nargs="*", default=[8, 16, 32])
    ap.add_argument("--C", type=int, nargs="*", default=[256, 512, 1024])
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0)
    H, V, beta, eos = 3584, 152064, 0.04, 151645
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    W = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    nt = (V + 255) // 256
    for G in args.G:
        for C in args.C:
            R = G * C
            h = torch.randn(R, H, device=dev, generator=g).bfloat16()
            ids = torch.randint(0, V, (G, C), device=dev, generator=g, dtype=torch.int32)
            ids[:, C // 2 + 3] = eos                          # EOS mid-way: half of every row is masked out
            tgt = ids.reshape(-1).contiguous()
            part = torch.empty((R, nt, 2), device=dev, dtype=torch.float32)
            tl = torch.zeros(R, device=dev, dtype=torch.float32)
            ref = torch.full((R,), -12.0, device=dev) + torch.randn(R, device=dev, generator=g) * 0.1
            adv = torch.linspace(-1.5, 1.5, G, device=dev)
            lp, lse, coef = (torch.empty(R, device=dev, dtype=torch.float32) for _ in range(3))
            mask = torch.empty(R, device=dev, dtype=torch.int32)
            rl, rk = torch.empty(G, device=dev), torch.empty(G, device=dev)
            rn = torch.empty(G, device=dev, dtype=torch.int32)
            out2 = torch.empty(2, device=dev)
            ws = ops.grpo_loss_workspace(G, C, dev)

            def fwd_gemm():
                ops.gemm(h, W, epilogue=EPI_LMHEAD, targets=tgt, lse_part=part, tgt_logit=tl)

            def tail():
                ops.call("sb_grpo_loss", part, nt, tl, ids, G, C, eos, ref, adv, beta, lp, lse, coef, mask, rl, rk, rn, out2, ws)

            dl = torch.empty((min(R, 4096), V), device=dev, dtype=torch.bfloat16)
            dh = torch.empty_like(h)
            dW = torch.empty_like(W)

            def bwd():
                for r0 in range(0, R, 4096):
                    r1 = min(R, r0 + 4096)
                    d = dl[: r1 - r0]
                    ops.gemm(h[r0:r1], W, epilogue=EPI_DLOGITS, targets=tgt[r0:r1], lse=lse[r0:r1], coef=coef[r0:r1], out=d)
                    ops.gemm(d, W, b_mn=True, out=dh[r0:r1])
                    ops.gemm(d, h[r0:r1], a_mn=True, b_mn=True, out=dW, residual=None if r0 == 0 else dW)

            t_f = ev_time(fwd_gemm)
            # tail: steady-state time per launch over rotating copies of the partials (> 2x L2 in total), so every
            # launch reads its input from HBM and launch latency is amortised like inside the training step
            n_copies = max(2, -(-(300 << 20) // (R * nt * 8)))
            parts = [part] + [part.clone() for _ in range(n_copies - 1)]
            fwd_gemm()

            def tails():
                for pc in parts:
                    ops.call("sb_grpo_loss", pc, nt, tl, ids, G, C, eos, ref, adv, beta, lp, lse, coef, mask, rl, rk, rn, out2, ws)
            t_t = ev_time(tails, reps=3) / n_copies
            # the same loop captured in a CUDA graph: the launches are no longer paced by the host (one ctypes call costs
            # ~8 us, more than the kernel at the small sizes), so this is the kernel's own back-to-back time
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                tails()
                graph.capture_begin()
                tails()
                graph.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            t_g = ev_time(graph.replay, reps=5) / n_copies
            del parts, graph
            t_b = ev_time(bwd, reps=3)
            # check against float64 torch on the kernel's own log-probs
            lp64 = lp.double().view(G, C)
            m = mask.view(G, C).double()
            x = torch.clamp(ref.double().view(G, C) - lp64, -10, 10)
            kl = torch.exp(x) - x - 1
            loss_ref = ((-(adv.double()[:, None] - beta * kl) * m).sum(1) / m.sum(1)).mean().item()
            # spot-check log-probs of 4 rows against a direct fp32 log_softmax of bf16-rounded logits
            rows = torch.tensor([0, R // 3, R // 2, R - 1], device=dev)
            lg = (h[rows].float() @ W.float().t()).bfloat16().float()
            lp_chk = torch.log_softmax(lg, -1).gather(1, tgt[rows].long()[:, None])[:, 0]
            tail_bytes = R * nt * 8 + R * 4 * 7 + G * 16
            flops_f = 2.0 * R * H * V
            print(json.dumps({
                "G": G, "C": C, "rows": R,
                "lmhead_fwd_ms": round(t_f, 3), "lmhead_fwd_tflops": round(flops_f / t_f / 1e9, 1), "lmhead_fwd_frac_of_bf16_peak": round(flops_f / t_f / 1e9 / tf, 3),
                "tail_us": round(t_t * 1e3, 2), "tail_us_in_graph": round(t_g * 1e3, 2), "tail_gbs_in_graph": round(tail_bytes / t_g / 1e6, 1),
                "tail_frac_of_hbm_peak_in_graph": round(tail_bytes / t_g / 1e6 / hbm, 3), "tail_bytes": tail_bytes, "tail_gbs": round(tail_bytes / t_t / 1e6, 1), "tail_frac_of_hbm_peak": round(tail_bytes / t_t / 1e6 / hbm, 3),
                "bwd_ms": round(t_b, 3), "bwd_tflops": round(3 * flops_f / t_b / 1e9, 1),
                "unfused_logits_bytes_avoided": 2 * R * V * 2,
                "loss": out2[0].item(), "loss_f64_ref": loss_ref, "abs_err": abs(out2[0].item() - loss_ref),
                "logprob_max_err_vs_log_softmax": (lp[rows] - lp_chk).abs().max().item()}), flush=True)
            del h, part, dl, dh, dW


if __name__ == "__main__":
    main()

"""ORACLE tooling (test infrastructure only).  Generates tests/golden/trn_compute_loss.pt by running the REFERENCE's own,
unmodified `SGRLVRTrainer.compute_loss` (SG_RLVR_trainer.py:384-686) + `loss.backward()` here on the CPU, on top of the
fp32 oracle model (see oracle/trn_harness.py for how the missing packages are stubbed).

Three scenarios: a video sample with T-GRPO (temporal) and length control on short completions; long completions that
reach the 320..512 length window; an image sample under --temporal (the dummy generate call of TRN:481).  For each one the
fixture holds the inputs (frames, prompt ids, fixed completions, reference-policy weights recipe), the LOG OF EVERY MODEL
CALL the trainer made (the contract spacer_b200/hf_api.py must accept), the rewards its reward functions returned, the
loss and the `_metrics` entries it computed, and norms + samples of the gradients its backward left on the weights.

Run in the build container only (needs /root/reference):   python oracle/make_trn_golden.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "trn_compute_loss.pt")

TEXTS = [
    "<think>The table is left of the chair.</think><map>{'table': [[1,3],[5,6]], 'chair': [[9,4]], 'window': [[6,5]]}</map><answer>B</answer>",
    "<think>Counting objects.</think><answer>B</answer>",
    "the answer might be B but I am not sure <map>table 1 3 chair</map>",
    "<think>Looking at frames.</think><answer>C</answer>",
]


def decode(ids):
    """Stand-in for tokenizer.batch_decode: one of four synthetic completions, chosen by the row's first token."""
    return [TEXTS[int(r[0]) % len(TEXTS)] for r in ids.tolist()]


def ref_weights(w, seed=5, scale=0.3):
    """The frozen reference policy of the fixture: the policy weights plus seeded noise (KL term > 0)."""
    g = torch.Generator().manual_seed(seed)
    return {k: v + scale * torch.randn(v.shape, generator=g) * (v.std() if v.numel() > 1 else 1.0) for k, v in w.items()}


def scenario_inputs(name, d):
    from oracle import qwen2vl_ref as R
    g = torch.Generator().manual_seed({"video_short": 11, "video_long": 12, "image": 13}[name])
    if name == "image":
        frames = torch.randint(0, 256, (1, 3, 112, 112), generator=g).float()
        n_v, kind, G, C = 16, "image", 4, 10
    else:
        frames = torch.randint(0, 256, (4, 3, 112, 112), generator=g).float()
        n_v, kind = 32, "video"
        G, C = (4, 12) if name == "video_short" else (4, 340)
    prompt = R.build_prompt_ids(d, n_v, 6, 10, seed=21)
    if kind == "image":
        prompt = torch.where(prompt == d.video_token_id, torch.full_like(prompt, d.image_token_id), prompt)
    comp = torch.randint(10, 2000, (G, C), generator=g)
    comp[:, 0] = torch.tensor([0, 1, 5, 3][:G]) + 4 * torch.randint(3, 400, (G,), generator=g)   # texts 0, 1, 1, 3
    if name == "video_long":
        comp[0, 329] = d.eos_id; comp[0, 330:] = d.pad_id        # length 330: inside the window
        comp[1, 99] = d.eos_id; comp[1, 100:] = d.pad_id         # length 100
        comp[2, 318] = d.eos_id; comp[2, 319:] = d.pad_id        # length 319: just outside
    else:
        comp[1, 7] = d.eos_id; comp[1, 8:] = d.pad_id
        comp[3, C - 1] = d.eos_id
    shuf = torch.randint(10, 2000, (G // 2, C), generator=g)
    shuf[:, 0] = torch.tensor([1, 2]) + 4 * torch.randint(3, 400, (2,), generator=g)
    return dict(frames=frames, kind=kind, G=G, C=C, prompt_ids=prompt, completions=comp, shuffled_completions=shuf)


def example_row(kind):
    return {"prompt": [{"role": "user", "content": [{"type": kind, kind: None, "text": None},
                                                    {"type": "text", "text": "Which object is closest to the table?"}]}],
            "path": "synthetic/scene0000_00.mp4" if kind == "video" else "synthetic/scene0000_00.jpg",
            "data_type": kind, "problem_id": 7, "problem": "Which object is closest to the table?",
            "problem_type": "multiple choice", "options": ["A. chair", "B. window"], "solution": "<answer>B</answer>",
            "data_source": "synthetic"}


def run_scenario(name, mod, ref_mod, verbose=True):
    from oracle import qwen2vl_ref as R
    from oracle import trn_harness as H
    from oracle.vision_ref import patchify_ref
    d = R.dims_tiny(2, 2)
    w = R.init_weights(d, seed=0)
    s = scenario_inputs(name, d)
    log = []
    policy = H.RecordingOracleModel(R, d, w, s["completions"], s["shuffled_completions"], log, "policy")
    refm = H.RecordingOracleModel(R, d, ref_weights(w), s["completions"], s["shuffled_completions"], log, "ref")
    for p in refm.parameters():
        p.requires_grad_(False)

    def patchify(frames):
        pix, grid = patchify_ref(frames)
        return pix, torch.tensor([list(grid)])
    proc = H.FakeProcessor(d, s["prompt_ids"], patchify, decode, s["kind"])
    reward_log = []

    def wrap(fn):
        def inner(**kw):
            out = fn(**kw)
            reward_log.append({"func": fn.__name__, "n": len(out), "out": [float(x) for x in out],
                               "kwargs": sorted(k for k in kw if k not in ("prompts", "completions"))})
            return out
        inner.__name__ = fn.__name__
        return inner
    funcs = [wrap(ref_mod.accuracy_reward), wrap(ref_mod.format_reward)]
    tr = H.make_trainer(mod, policy, refm, proc, funcs, s["G"], s["C"], 0.04, True, True, s["frames"], s["kind"])
    loss, metrics, out = H.run_compute_loss(tr, policy, example_row(s["kind"]), seed=3)
    loss.backward()
    torch.manual_seed(3)
    perm = torch.randperm(s["frames"].shape[0]) if s["kind"] == "video" else None
    grads = {k: p.grad for k, p in policy.w.items()}
    gsum = {k: dict(norm=float(g.norm()), head=g.flatten()[:16].clone(),
                    stride_sample=g.flatten()[::max(1, g.numel() // 64)][:64].clone()) for k, g in grads.items()}
    if verbose:
        print(f"[{name}] loss {loss.item():+.6f}  metrics { {k: round(v[0], 5) for k, v in metrics.items()} }")
        for r in log:
            print("   ", r["model"], r["call"], r.get("generation_config") or "", sorted(r["kwargs"]))
    s["frames"] = s["frames"].to(torch.uint8)              # integer-valued by construction
    s.update(loss=float(loss.item()), metrics={k: v[0] for k, v in metrics.items()}, calls=log, rewards=reward_log,
             frame_perm=perm, grad_summary=gsum, processor_calls=proc.calls, beta=0.04, weights_seed=0,
             ref_weights=dict(seed=5, scale=0.3), stdout_tail=out[-300:])
    return s


def main():
    from oracle import trn_harness as H
    from oracle.make_golden import MAP_ROWS, load_reference
    assert H.reference_available(), "needs /root/reference"
    mod = H.load_reference_trainer()
    _, ref_mod = load_reference()
    ref_mod.MAP_DATA = MAP_ROWS
    out = {name: run_scenario(name, mod, ref_mod) for name in ("video_short", "video_long", "image")}
    out["texts"] = TEXTS
    out["map_rows"] = MAP_ROWS
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

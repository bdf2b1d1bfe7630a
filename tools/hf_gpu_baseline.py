#!/usr/bin/env python
"""The reference's GRPO step on the reference's own model stack -- NONE of this repo's kernels.

`HFStep` restates `SGRLVRTrainer.compute_loss` + the HF Trainer's backward/optimizer
(SG_RLVR_trainer.py:384-686, run_SpaceR_SG_RLVR.sh:16-39) over the stock Hugging Face
`Qwen2VLForConditionalGeneration` (transformers 5.5.0, random-init weights of the real config).  trl / accelerate /
deepspeed are absent from this image, so the Trainer plumbing is the plain loop below; the model calls are the reference's:
    generate(num_return_sequences=G, do_sample, top_p .95)  (+ G/2 on the frame-shuffled video)     TRN:463-481
    reference-policy forward over the G full sequences with xG-repeated pixels, inference_mode     TRN:534-547
    policy forward (gradient checkpointing, as run_SpaceR_SG_RLVR.sh:27) -> per-row log_softmax/gather -> GRPO loss
    -> backward -> (N > 1: gradient all-reduce over NCCL) -> clip 5 -> AdamW                        TRN:353-366, 640-643
Two uses:
  * on the B200 (bf16, flash_attention_2 as run_SpaceR_SG_RLVR.sh:31 configures): the comparator of BASELINE's
    ">= 5x the reference HF/PyTorch GRPO step" target -- `bench.py` runs it as its `hf_gpu_baseline` leg, at N ranks;
  * on the host cores (fp32, eager attention): the reference's CPU path as a REAL full step at cfg1 / cfg2
    (`tools/cpu_reference_step.py`, bench.py's cpu_baseline).
    python tools/hf_gpu_baseline.py [--config c3] [--attn sdpa|flash_attention_2|eager] [--steps 1] [--device cuda]
Prints one JSON line (samples/s, ms per phase)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def hf_config(d):
    from transformers import Qwen2VLConfig
    return Qwen2VLConfig(
        text_config=dict(hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                         num_key_value_heads=d.kv_heads, intermediate_size=d.inter, vocab_size=d.vocab,
                         rms_norm_eps=d.rms_eps,
                         rope_parameters={"rope_type": "default", "rope_theta": d.rope_theta,
                                          "mrope_section": list(d.mrope_section)},
                         tie_word_embeddings=d.tie, max_position_embeddings=32768, use_sliding_window=False,
                         bos_token_id=d.pad_id, eos_token_id=d.eos_id, pad_token_id=d.pad_id),
        vision_config=dict(depth=d.v_depth, embed_dim=d.v_embed, hidden_size=d.hidden, num_heads=d.v_heads,
                           mlp_ratio=d.v_mlp // d.v_embed, patch_size=d.patch, temporal_patch_size=d.t_patch,
                           spatial_merge_size=d.merge, in_channels=d.in_ch, hidden_act="quick_gelu"),
        image_token_id=d.image_token_id, video_token_id=d.video_token_id,
        vision_start_token_id=d.vision_start_id, vision_end_token_id=d.vision_end_id, tie_word_embeddings=d.tie)


def per_token_logps(logits, input_ids):
    """SGRLVRTrainer._get_per_token_logps (SG_RLVR_trainer.py:353-366), verbatim structure: per-row log_softmax + gather."""
    logits = logits[:, :-1, :]
    ids = input_ids[:, 1:]
    out = []
    for lr, ir in zip(logits, ids):
        lp = lr.log_softmax(dim=-1)
        out.append(torch.gather(lp, dim=1, index=ir.unsqueeze(1)).squeeze(1))
    return torch.stack(out)


def grpo_loss(lp, ref_lp, adv, mask, beta):
    """SG_RLVR_trainer.py:551-552, 640-643."""
    x = torch.clamp(ref_lp - lp, -10, 10)
    kl = torch.exp(x) - x - 1
    per_tok = -(torch.exp(lp - lp.detach()) * adv.unsqueeze(1) - beta * kl)
    return ((per_tok * mask).sum(1) / mask.sum(1)).mean()


class HFStep:
    """One process = one rank = one prompt per step, like the reference (per_device_train_batch_size 1)."""

    def __init__(self, cfg: dict, device, attn: str | None = None, seed: int = 1234, dist_group=None, world: int = 1):
        import bench
        from spacer_b200 import config as mcfg   # dimensions and the host-side position-id helper only; no kernels
        from transformers import Qwen2VLForConditionalGeneration
        self.cfg, self.dev = cfg, torch.device(device)
        self.d = d = mcfg.PRESETS[cfg["preset"]]()
        gpu = self.dev.type == "cuda"
        self.dt = torch.bfloat16 if gpu else torch.float32
        if attn is None:
            attn = "eager"
            if gpu:
                try:
                    import flash_attn  # noqa: F401
                    attn = "flash_attention_2"
                except Exception:
                    attn = "sdpa"
        self.attn = attn
        hc = hf_config(d)
        hc._attn_implementation = attn
        torch.manual_seed(0)
        with torch.device(self.dev):
            self.model = Qwen2VLForConditionalGeneration(hc).to(self.dt)
            self.ref = Qwen2VLForConditionalGeneration(hc).to(self.dt)
        self.ref.load_state_dict(self.model.state_dict())
        self.ref.eval()
        self.model.gradient_checkpointing_enable()
        self.model.config.use_cache = True
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=1e-6, weight_decay=0.01, fused=gpu)
        ex = bench.synth_example(d, cfg, seed)
        self.pix = ex["pixel_values_host"].to(self.dev, self.dt)
        self.grid = ex["video_grid_thw"].to(self.dev)
        self.ids = ex["input_ids"].to(self.dev)
        self.pix2 = self.pix.flip(0).contiguous()
        self.group, self.world = dist_group, world

    def sync(self):
        if self.dev.type == "cuda":
            torch.cuda.synchronize()

    def step(self, C=None, temporal=True):
        from spacer_b200.model import rope_index
        d, dev, model, ref = self.d, self.dev, self.model, self.ref
        ids, pix, grid = self.ids, self.pix, self.grid
        G = self.cfg["G"]
        C_ = self.cfg["C"] if C is None else C
        P = ids.shape[1]
        mm = (ids == d.video_token_id).long() * 2 + (ids == d.image_token_id).long()
        t = {}
        self.sync(); t0 = time.perf_counter()
        model.eval()
        with torch.no_grad():
            kw = dict(max_new_tokens=C_, min_new_tokens=C_, do_sample=True, top_p=0.95, temperature=1.0,
                      pad_token_id=d.pad_id, use_cache=True)
            out = model.generate(input_ids=ids, mm_token_type_ids=mm, pixel_values_videos=pix, video_grid_thw=grid,
                                 num_return_sequences=G, **kw)
            if temporal:
                model.generate(input_ids=ids, mm_token_type_ids=mm, pixel_values_videos=self.pix2, video_grid_thw=grid,
                               num_return_sequences=G // 2, **kw)
        self.sync(); t["rollout"] = time.perf_counter() - t0
        full = out                                            # [G, P + C]
        mmf = (full == d.video_token_id).long() * 2 + (full == d.image_token_id).long()
        pos = torch.stack([rope_index(row, grid.cpu(), d, "classic")[0] for row in full.cpu()], dim=1).to(dev)
        pixG, gridG = pix.repeat(G, 1), grid.repeat(G, 1)
        t0 = time.perf_counter()
        with torch.inference_mode():
            rl = ref(input_ids=full, pixel_values_videos=pixG, video_grid_thw=gridG, position_ids=pos, mm_token_type_ids=mmf,
                     use_cache=False).logits
            ref_lp = per_token_logps(rl, full)[:, P - 1:]
            del rl
        self.sync(); t["ref_scoring"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        model.train()
        logits = model(input_ids=full, pixel_values_videos=pixG, video_grid_thw=gridG, position_ids=pos, mm_token_type_ids=mmf,
                       use_cache=False).logits
        lp = per_token_logps(logits, full)[:, P - 1:]
        del logits
        comp = full[:, P:]
        is_eos = comp == d.eos_id                                             # TRN:493-498
        eos_idx = torch.full((G,), comp.shape[1], dtype=torch.long, device=dev)
        eos_idx[is_eos.any(1)] = is_eos.int().argmax(1)[is_eos.any(1)]
        mask = (torch.arange(comp.shape[1], device=dev)[None] <= eos_idx[:, None]).int()
        rewards = torch.linspace(0.0, 2.0, G, device=dev)
        adv = (rewards - rewards.mean()) / (rewards.std() + 1e-4)             # TRN:632-638
        loss = grpo_loss(lp.float(), ref_lp.float().clone(), adv, mask, 0.04)
        loss.backward()
        self.sync(); t["policy_fwd_bwd"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if self.world > 1:       # data parallel like the reference (one prompt per rank): gradients averaged over ranks
            import torch.distributed as dist
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            for g in grads:
                dist.all_reduce(g, group=self.group)
            torch._foreach_div_(grads, float(self.world))
            self.sync(); t["grad_allreduce"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.sync(); t["clip_adamw"] = time.perf_counter() - t0
        t["loss"] = float(loss.item())
        return t

    def free(self):
        self.model = self.ref = self.opt = None
        import gc
        gc.collect()
        if self.dev.type == "cuda":
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--attn", default=None)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--completion", type=int, default=0, help="override C (debug)")
    a = ap.parse_args()
    import bench
    cfg = dict(bench.CONFIGS[a.config])
    if a.completion:
        cfg["C"] = a.completion
    hs = HFStep(cfg, a.device, a.attn)
    G, C = cfg["G"], cfg["C"]
    hs.step(min(C, 4))                                        # warm-up (allocator, kernels, autotune)
    tot = {}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        for k, v in hs.step(C).items():
            tot[k] = tot.get(k, 0.0) + v
    hs.sync()
    wall = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"impl": "hf_restatement", "config": cfg["workload"], "attn_implementation": hs.attn, "device": str(hs.dev),
                      "transformers": __import__("transformers").__version__, "dtype": str(hs.dt), "steps": a.steps,
                      "s_per_step": round(wall, 3), "samples_per_s": round(G / wall, 4),
                      "rollout_tok_per_s": round((G + G // 2) * C / (tot["rollout"] / a.steps), 1),
                      "phase_s": {k: round(v / a.steps, 3) for k, v in tot.items() if k != "loss"}}))


if __name__ == "__main__":
    main()

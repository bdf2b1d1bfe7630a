"""GRPO step of SG-RLVR on the spacer_b200 engine -- host-side mirror of `SGRLVRTrainer.compute_loss`
(/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py:384-686, "TRN") plus the
`training_step` tail HF `Trainer` adds around it (backward, clip, AdamW, cosine LR).

Same step order, constants and metric names as the reference; the compute underneath is the CUDA library.
Deviations (all documented in DESIGN.md): the scoring-forward exception fallback of TRN:529-547 is not
reproduced (errors raise); tensors are created on the model's device instead of a hard-coded 'cuda'.
"""
from __future__ import annotations

import math
from collections import defaultdict
from dataclasses import dataclass
from typing import Callable, Sequence

import torch

from . import dist as D
from . import ops
from . import vision
from .model import GradStore, Qwen2VLB200, pack_prompt_completions
from .params import ParamStore

# hard-coded constants of the reference, named
TOP_P = 0.95                 # TRN:277-284
TOP_K = 50                   # not in the reference's GenerationConfig(...) call: HF fills top_k = 50 for every field left
                             # unset (generation/configuration_utils.py:551-581; class default in transformers 4.x)
KL_CLAMP = 10.0              # TRN:551 (inside the fused loss kernel)
TEMPORAL_RATIO = 0.8         # TRN:604
TEMPORAL_BONUS = 0.3         # TRN:606
ACC_THRESHOLD = 0.1          # TRN:605, 622
LEN_WINDOW = (320, 512)      # TRN:628
LEN_BONUS = 0.2              # TRN:629
STD_EPS = 1e-4               # TRN:638


@dataclass
class GRPOConfig:
    """The subset of trl.GRPOConfig / GRPOScriptArguments the hot path reads (run_SpaceR_SG_RLVR.sh:16-39)."""
    num_generations: int = 8
    max_prompt_length: int = 16384
    max_completion_length: int = 1024
    beta: float = 0.04
    temporal: bool = True
    len_control: bool = True
    learning_rate: float = 1e-6
    weight_decay: float = 0.01
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8
    max_grad_norm: float = 5.0
    lr_scheduler_type: str = "cosine"
    max_steps: int = 1000
    warmup_steps: int = 0
    seed: int = 42
    gradient_checkpointing: bool = False   # run_SpaceR_SG_RLVR.sh:27; here: recompute the gate|up activation (model.py)
    top_k: int = TOP_K              # 0 = nucleus sampling only
    eos_token_ids: tuple = ()       # ids that end a rollout row (HF: generation_config.eos_token_id); () = the model's
    moments_bf16: bool = False      # fp32 moments like DeepSpeed unless memory forces otherwise
    min_new_tokens: int = 0         # = max_completion_length disables EOS (timing runs)
    overlap_allreduce: bool = True  # per-layer gradient buckets all-reduced during the backward (N > 1)
    zero1: bool = False             # N > 1: shard the optimizer state (fp32 master + moments) over the ranks:
                                    # reduce-scatter gradients -> AdamW on the own 1/N of the matrix arena -> all-gather
                                    # the updated bf16 weights (what the reference gets from DeepSpeed, zero3.json)


def completion_lengths(completion_ids: torch.Tensor, eos_id: int) -> torch.Tensor:
    """`completion_mask.sum(1)` of TRN:489-494 without building the mask: tokens up to and INCLUDING the first EOS."""
    is_eos = completion_ids == eos_id
    n, C = completion_ids.shape
    eos_idx = torch.where(is_eos.any(1), is_eos.int().argmax(1), torch.full((n,), C, device=completion_ids.device))
    return torch.clamp(eos_idx + 1, max=C)


def reward_tail(rewards_per_func: torch.Tensor, shuffled_rewards_per_func, lengths: torch.Tensor, num_generations: int,
                temporal: bool = True, len_control: bool = True):
    """TRN:598-638: T-GRPO temporal bonus, summed reward, length-control bonus, group-relative advantages.

    rewards_per_func [B*G, n_funcs] (column 0 = accuracy), shuffled_rewards_per_func [G/2, n_funcs] or None (no video /
    temporal off), lengths [B*G] = completion_mask.sum(1).  Returns (rewards [B*G], advantages [B*G], std [B*G],
    temporal_rewards float: 1 bonus granted, 0 refused, 0.5 not applicable)."""
    G = num_generations
    temporal_rewards = 0.5
    summed = rewards_per_func
    if temporal and shuffled_rewards_per_func is not None:
        summed = rewards_per_func.clone()                                       # TRN:598-611
        if summed[:, 0].mean() >= TEMPORAL_RATIO * shuffled_rewards_per_func[:, 0].mean():
            sel = summed[:, 0] > ACC_THRESHOLD
            summed[sel, 0] = summed[sel, 0] + TEMPORAL_BONUS
            temporal_rewards = 1.0
        else:
            temporal_rewards = 0.0
    rewards = summed.sum(dim=1)                                                 # TRN:613-617
    if len_control:                                                             # TRN:620-629 (pre-temporal accuracy)
        sel = torch.nonzero(rewards_per_func[:, 0] > ACC_THRESHOLD, as_tuple=True)[0].tolist()
        if len(sel) > 1:
            ll = lengths.tolist()
            for i in sel:
                if LEN_WINDOW[0] <= ll[i] <= LEN_WINDOW[1]:
                    rewards[i] += LEN_BONUS
    mean = rewards.view(-1, G).mean(dim=1).repeat_interleave(G, dim=0)          # TRN:632-638
    std = rewards.view(-1, G).std(dim=1).repeat_interleave(G, dim=0)            # unbiased; G = 1 gives NaN like the reference
    adv = (rewards - mean) / (std + STD_EPS)
    return rewards, adv, std, temporal_rewards


def pack_step_stats(lengths, rewards_per_func, rewards, std, mean_kl, temporal_rewards) -> torch.Tensor:
    """The per-rank quantities TRN:650-683 gathers one by one, as one fp32 vector (one all-gather per step)."""
    dev = rewards.device
    return torch.cat([lengths.float(), rewards_per_func.reshape(-1).float(), rewards.float(), std.float(),
                      torch.stack([torch.as_tensor(mean_kl, device=dev).float().reshape(()),
                                   torch.tensor(float(temporal_rewards), device=dev)])])


def step_metrics(gathered: torch.Tensor, num_generations: int, reward_names: Sequence[str], temporal: bool) -> dict:
    """TRN:650-683 on the gathered [world, len(pack_step_stats)] matrix: completion_length, rewards/<func>, all_wrong
    (every reward of a prompt <= 1), all_correct (every reward >= 2), temporal_rewards, reward, reward_std, kl."""
    G, nf = num_generations, len(reward_names)
    n = (gathered.shape[1] - 2) // (3 + nf)            # rows per rank (B*G)
    gl = gathered[:, :n]
    grpf = gathered[:, n:n + n * nf].reshape(-1, nf)
    grew = gathered[:, n + n * nf:2 * n + n * nf]
    gstd = gathered[:, 2 * n + n * nf:3 * n + n * nf]
    tail = gathered[:, -2:]
    mt = {"completion_length": gl.mean().item()}
    rpf_mean = grpf.mean(0)
    for i, name in enumerate(reward_names):
        mt[f"rewards/{name}"] = rpf_mean[i].item()
    per_dev = grew.reshape(-1, G)                      # one row per prompt ("device" in the reference: 1 prompt / rank)
    mt["all_wrong"] = (per_dev <= 1).all(dim=1).sum().item() / per_dev.shape[0]
    mt["all_correct"] = (per_dev >= 2).all(dim=1).sum().item() / per_dev.shape[0]
    if temporal:
        mt["temporal_rewards"] = tail[:, 1].mean().item()
    mt["reward"] = grew.mean().item()
    mt["reward_std"] = gstd.mean().item()
    mt["kl"] = tail[:, 0].mean().item()
    return mt


class AdamW:
    """AdamW over the flat arenas with fp32 master weights + global-norm clip (zero3.json:10-12 semantics)."""

    def __init__(self, params: ParamStore, cfg: GRPOConfig, shard: tuple | None = None, group=None):
        """shard = (rank, world): this rank owns elements [rank * n / world, (rank + 1) * n / world) of the matrix arena
        (ZeRO-1); master weights and moments exist for that range only.  The small vector arena is replicated."""
        self.p, self.cfg, self.group = params, cfg, group
        mdt = torch.bfloat16 if cfg.moments_bf16 else torch.float32
        dev = params.device
        n = params.sizes["mat"]
        self.shard = None
        self.lo, self.hi = 0, n
        if shard is not None and shard[1] > 1:
            rank, world = shard
            if n % (world * 16) != 0:
                raise ops.SpacerError(f"zero1: matrix arena of {n} elements does not split into {world} aligned shards")
            per = n // world
            self.shard, self.lo, self.hi = (rank, world), rank * per, (rank + 1) * per
        n_own = self.hi - self.lo
        sizes = {"mat": n_own, "vec": params.sizes["vec"]}
        self.master = [torch.empty(sizes[k], device=dev, dtype=torch.float32) for k in ("mat", "vec")]
        self.resync_master()
        self.m = [torch.zeros(sizes[k], device=dev, dtype=mdt) for k in ("mat", "vec")]
        self.v = [torch.zeros(sizes[k], device=dev, dtype=mdt) for k in ("mat", "vec")]
        self.total_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.t = 0

    def resync_master(self):
        """fp32 master weights := the bf16 arenas.  Called at construction and whenever the weights were replaced from
        outside the optimizer (`load_state_dict` / `from_pretrained` after the trainer exists: ParamStore.version changes),
        so that the next step does not publish stale masters."""
        p = self.p
        ops.call("sb_bf16_to_f32", p.mat[self.lo:self.hi], self.master[0], self.hi - self.lo)
        ops.call("sb_bf16_to_f32", p.vec, self.master[1], p.vec.numel())
        self._seen_version = getattr(p, "version", 0)

    def lr_at(self, step: int) -> float:
        c = self.cfg
        if step < c.warmup_steps:
            return c.learning_rate * step / max(1, c.warmup_steps)
        if c.lr_scheduler_type == "cosine":
            prog = (step - c.warmup_steps) / max(1, c.max_steps - c.warmup_steps)
            return c.learning_rate * max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))
        return c.learning_rate

    def grad_sumsq(self, grads: GradStore):
        """Global sum of squares of the (already reduced) gradients.  Sharded: every rank sums its own slice of the matrix
        arena (rank 0 adds the replicated vector arena) and one scalar all-reduce makes the total identical everywhere."""
        self.total_sq.zero_()
        ops.call("sb_grad_sumsq", grads.mat[self.lo:self.hi], self.hi - self.lo, 0, self.total_sq)
        if self.shard is None or self.shard[0] == 0:
            ops.call("sb_grad_sumsq", grads.vec, grads.vec.numel(), 1, self.total_sq)
        if self.shard is not None:
            torch.distributed.all_reduce(self.total_sq, group=self.group)
        return self.total_sq

    def step(self, grads: GradStore, grad_scale: float = 1.0, sumsq_ready: bool = False):
        c = self.cfg
        if getattr(self.p, "version", 0) != self._seen_version:
            self.resync_master()
        if not sumsq_ready:
            self.grad_sumsq(grads)
        lr = self.lr_at(self.t)
        self.t += 1
        mb = int(c.moments_bf16)
        ops.call("sb_adamw_step", self.p.mat[self.lo:self.hi], self.master[0], self.m[0], self.v[0],
                 grads.mat[self.lo:self.hi], self.hi - self.lo, 0,
                 mb, self.total_sq, lr, c.adam_beta1, c.adam_beta2, c.adam_eps, c.weight_decay, self.t,
                 c.max_grad_norm, grad_scale)
        ops.call("sb_adamw_step", self.p.vec, self.master[1], self.m[1], self.v[1], grads.vec, self.p.vec.numel(), 1,
                 mb, self.total_sq, lr, c.adam_beta1, c.adam_beta2, c.adam_eps, 0.0, self.t, c.max_grad_norm,
                 grad_scale)
        return lr


class SGRLVRTrainerB200:
    """One prompt per rank per step -> G sampled completions (+ G/2 from the frame-shuffled video) -> rewards ->
    group-relative advantages -> one GRPO update.  Data-parallel over ranks: gradients are summed with one NCCL
    all-reduce per arena (torch.distributed), nothing else crosses ranks on the data path."""

    def __init__(self, model: Qwen2VLB200, ref_model: Qwen2VLB200 | None, reward_funcs: Sequence[Callable],
                 cfg: GRPOConfig, decode_completions: Callable[[torch.Tensor], list[str]],
                 process_group=None):
        self.model, self.ref_model, self.reward_funcs, self.cfg = model, ref_model, list(reward_funcs), cfg
        self.decode_completions = decode_completions
        if cfg.gradient_checkpointing:
            model.gradient_checkpointing_enable()
        self.grads = GradStore(model.params)
        self.pg = process_group
        self.zero1 = bool(cfg.zero1) and D.world_size(process_group) > 1
        self.opt = AdamW(model.params, cfg, shard=(D.rank(process_group), D.world_size(process_group)) if self.zero1 else None,
                         group=process_group)
        # data parallel: per-layer gradient buckets are all-reduced while the backward is still running
        self.reducer = (D.OverlappedGradReducer(self.grads.mat, self.grads.vec, process_group)
                        if cfg.overlap_allreduce and not self.zero1 else None)
        if self.reducer is not None:
            self.grads.on_ready = self.reducer.ready
        self._metrics = defaultdict(list)
        self.global_step = 0
        self.last_rollout_stats = None

    # -- distributed helpers (spacer_b200/dist.py) ---------------------------------------------------------
    def _world(self):
        return D.world_size(self.pg)

    def _allreduce_grads(self):
        if self.zero1:
            # every rank receives the sum of ITS slice of the matrix gradients (in place); the vector arena is replicated
            o = self.opt
            torch.distributed.reduce_scatter_tensor(self.grads.mat[o.lo:o.hi], self.grads.mat, group=self.pg)
            torch.distributed.all_reduce(self.grads.vec, group=self.pg)
        elif self.reducer is not None:
            self.reducer.finish()
        else:
            D.allreduce_sum_([self.grads.mat, self.grads.vec], group=self.pg)

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        return D.gather_rows(t, group=self.pg)

    # -- the step --------------------------------------------------------------------------------------
    def rollout(self, example, seed):
        """TRN:442-481: G completions of the prompt (+ G/2 of the frame-shuffled video when temporal)."""
        c = self.cfg
        G = c.num_generations
        frames = example.get("video_frames")
        if frames is not None:
            # GPU front-end: rescale + normalise + patchify + bf16 cast in one kernel; the T-GRPO
            # frame shuffle (TRN:442-458) is an index permutation inside the same kernel instead of a second
            # processor pass
            pix, _, grid = vision.patchify(frames)
            example["pixel_values_videos"], example["video_grid_thw"] = pix, grid
        pix = example["pixel_values_videos"]
        grid = example["video_grid_thw"]
        ids = example["input_ids"]
        if c.max_prompt_length is not None and ids.shape[-1] > c.max_prompt_length:
            ids = ids[..., -c.max_prompt_length:]                    # TRN:432-440
        # Qwen2.5-VL: the rollout sees the processor's second_per_grid_ts; the scoring forwards do not (TRN:519-520)
        kw = dict(max_new_tokens=c.max_completion_length, top_p=TOP_P, top_k=c.top_k, temperature=1.0, do_sample=True,
                  seed=seed, min_new_tokens=c.min_new_tokens, eos_token_id=list(c.eos_token_ids) or None,
                  second_per_grid_ts=example.get("second_per_grid_ts"), keep_vit_tape=True)
        if c.temporal and pix is not None:
            if frames is not None:
                g = torch.Generator(device="cpu").manual_seed(int(seed) + 7919)
                perm = torch.randperm(frames.shape[0], generator=g).to(torch.int32).to(frames.device)
                pix2, _, _ = vision.patchify(frames, perm)
            else:
                pix2 = self.shuffle_frames(pix, grid, seed)
            main, shuf = self.model.generate(ids, pix, grid, num_return_sequences=G, pixel_values_videos_2=pix2,
                                             num_return_sequences_2=G // 2, **kw)
            self.last_rollout_stats = self.model.last_generate_stats
            return ids, main, shuf
        main = self.model.generate(ids, pix, grid, num_return_sequences=G, **kw)
        self.last_rollout_stats = self.model.last_generate_stats
        return ids, main, None

    @staticmethod
    def shuffle_frames(pixel_values, grid_thw, seed):
        """TRN:442-458 permutes the decoded FRAMES and re-runs the processor.  Normalisation is per pixel, so on
        the patch matrix this is an index permutation of the per-frame halves of the patch rows."""
        grid = grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw
        t, h, w = map(int, grid[0])
        hw = h * w
        c2 = pixel_values.shape[1] // 2          # channels * patch * patch per frame of the temporal pair
        n_ch = 3
        pp = c2 // n_ch
        g = torch.Generator(device="cpu").manual_seed(int(seed) + 7919)
        perm = torch.randperm(2 * t, generator=g).to(pixel_values.device)
        # rows are [C, Tp=2, 14, 14] flattened: bring the frame index out, permute frames, fold back
        x = pixel_values.view(t, hw, n_ch, 2, pp).permute(0, 3, 1, 2, 4).reshape(2 * t, hw, n_ch, pp)
        x = x[perm]
        return x.view(t, 2, hw, n_ch, pp).permute(0, 2, 3, 1, 4).reshape(pixel_values.shape).contiguous()

    def compute_rewards(self, completions_text, example, G):
        prompts = [example.get("prompt")] * G
        completions = [[{"role": "assistant", "content": s}] for s in completions_text]
        per_func = torch.zeros(G, len(self.reward_funcs))
        extra = {k: [example[k]] * G for k in example
                 if k not in ("prompt", "completion", "input_ids", "pixel_values_videos", "video_grid_thw", "path",
                              "video_frames")}
        for i, fn in enumerate(self.reward_funcs):
            out = fn(prompts=prompts, completions=completions, path=[example.get("path", "")] * G, **extra)  # TRN:592
            per_func[:, i] = torch.tensor([float(x) for x in out])
        return per_func

    def rollout_seed(self) -> int:
        """Seed of this step's rollout: differs per step AND per data-parallel rank (the reference's ranks draw from
        independent torch generators; identical Philox keys would make two ranks that meet the same prompt emit
        bit-identical completions)."""
        return self.cfg.seed + 1000003 * self.global_step + 7919 * D.rank(self.pg)

    def training_step(self, example) -> dict:
        c, m = self.cfg, self.model
        G = c.num_generations
        seed = self.rollout_seed()
        marks = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

        mark("start")
        prompt_ids, main, shuf = self.rollout(example, seed)
        mark("rollout")
        P = prompt_ids.reshape(-1).numel()
        completion_ids = main[:, P:]
        # the sampled ids go to the host NOW (the GPU is idle after the rollout anyway): everything the CPU does with them
        # -- decoding, reward functions, bonuses, advantages -- then runs while the GPU works on the reference scoring,
        # instead of stalling behind it and leaving the GPU idle before the policy forward
        comp_host = completion_ids.cpu()
        shuf_host = shuf[:, P:].cpu() if shuf is not None else None
        pix, grid = example["pixel_values_videos"], example["video_grid_thw"]
        batch = pack_prompt_completions(prompt_ids, comp_host, grid, m.dims, m.device, m.rope_convention)
        # reference-policy log-probs (TRN:534-547); beta = 0 skips it.  Enqueued, not waited for.
        ref_lp = None
        if self.ref_model is not None and c.beta != 0.0:
            ref_lp = self.ref_model.per_token_logps(batch, pix, grid)
        # rewards on decoded text (TRN:555-593), on the host
        rewards_per_func = self.compute_rewards(self.decode_completions(comp_host), example, G)
        shuf_rpf = None
        if c.temporal and shuf_host is not None:
            shuf_rpf = self.compute_rewards(self.decode_completions(shuf_host), example, G // 2)
        # completion lengths are needed for the length bonus before the loss kernel runs (TRN:489-494, 620-629)
        lengths = completion_lengths(comp_host, m.dims.eos_id)
        rewards, adv, std, temporal_rewards = reward_tail(rewards_per_func, shuf_rpf, lengths, G, c.temporal, c.len_control)
        mark("ref_scoring")
        mark("rewards")
        if self.reducer is not None:
            self.reducer.begin_step()
        out = m.grpo_forward_backward(batch, pix, grid, ref_lp, adv, c.beta, self.grads, vit_cache=m.vit_cache)
        m.vit_cache = None
        mark("policy_fwd_bwd")
        # data parallel: sum gradients over ranks, average inside the optimizer
        self._allreduce_grads()
        mark("grad_allreduce")
        lr = self.opt.step(self.grads, grad_scale=1.0 / self._world())
        if self.zero1:      # every rank publishes the slice of bf16 weights it owns (in place)
            o = self.opt
            torch.distributed.all_gather_into_tensor(m.params.mat, m.params.mat[o.lo:o.hi], group=self.pg)
        mark("adamw")
        self.global_step += 1
        # metrics (TRN:650-683): the reference gathers nine quantities one by one; here one packed vector per rank
        allp = self._gather(pack_step_stats(lengths.to(m.device), rewards_per_func.to(m.device), rewards.to(m.device),
                                            std.to(m.device), out["mean_kl"], temporal_rewards))
        mt = step_metrics(allp, G, [fn.__name__ for fn in self.reward_funcs], c.temporal)
        mt["loss"] = out["loss"].item()
        mt["learning_rate"] = lr
        for k, v in mt.items():
            self._metrics[k].append(v)
        self.last_phase_ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        mt["generated_tokens"] = int(comp_host.numel() + (shuf_host.numel() if shuf_host is not None else 0))
        return mt

    def log(self):
        """TRN:688-695: average and clear."""
        out = {k: sum(v) / len(v) for k, v in self._metrics.items() if v}
        self._metrics.clear()
        return out

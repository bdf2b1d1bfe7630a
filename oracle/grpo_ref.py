"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement (plain torch, fp32/fp64) of the loss tail of `SGRLVRTrainer.compute_loss`
(/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py, "TRN").  The reference
trainer cannot be imported here (trl / accelerate / deepspeed are absent) and has no tests, so parity
for this file is pinned by (a) hand-derived known answers in tests/test_oracle_cpu.py and (b) the
golden end-to-end loss of tests/golden/cfg1_tiny.pt, whose per-token log-probs come from HF.
Each function cites the TRN lines it follows, quirks included.
"""
from __future__ import annotations

import torch


def completion_mask(completion_ids: torch.Tensor, eos_id: int) -> torch.Tensor:
    """TRN:489-494 -- positions up to and INCLUDING the first EOS are 1 (int32)."""
    is_eos = completion_ids == eos_id
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long)
    any_eos = is_eos.any(dim=1)
    eos_idx[any_eos] = is_eos.int().argmax(dim=1)[any_eos]
    seq = torch.arange(is_eos.size(1)).expand(is_eos.size(0), -1)
    return (seq <= eos_idx.unsqueeze(1)).int()


def per_token_kl(ref_logps: torch.Tensor, logps: torch.Tensor) -> torch.Tensor:
    """TRN:551-552 -- k3 estimator on x = clamp(ref - pol, -10, 10)."""
    x = torch.clamp(ref_logps - logps, min=-10, max=10)
    return torch.exp(x) - x - 1


def temporal_bonus(rewards_per_func: torch.Tensor, shuffled_rewards_per_func, temporal: bool):
    """TRN:598-611 -- returns (rewards_per_func used for the sum, temporal_rewards scalar)."""
    if temporal and shuffled_rewards_per_func is not None:
        t = rewards_per_func.clone()
        if t[:, 0].mean() >= 0.8 * shuffled_rewards_per_func[:, 0].mean():
            m = t[:, 0] > 0.1
            t[m, 0] = t[m, 0] + 0.3
            return t, 1.0
        return t, 0.0
    return rewards_per_func, 0.5


def length_bonus(rewards: torch.Tensor, rewards_per_func: torch.Tensor, mask: torch.Tensor,
                 len_control: bool) -> torch.Tensor:
    """TRN:620-629 -- +0.2 for rows with (pre-temporal) accuracy > 0.1 and 320 <= len <= 512, only when
    at least two rows have accuracy > 0.1."""
    rewards = rewards.clone()
    if len_control:
        sel = torch.nonzero(rewards_per_func[:, 0] > 0.1, as_tuple=True)[0].tolist()
        lens = mask.sum(1)
        if len(sel) > 1:
            for i in sel:
                if 320 <= int(lens[i]) <= 512:
                    rewards[i] += 0.2
    return rewards


def advantages(rewards: torch.Tensor, G: int):
    """TRN:632-638 -- group mean / UNBIASED std over G, eps 1e-4 added to the std."""
    mean = rewards.view(-1, G).mean(dim=1).repeat_interleave(G, dim=0)
    std = rewards.view(-1, G).std(dim=1).repeat_interleave(G, dim=0)
    return (rewards - mean) / (std + 1e-4), std


def grpo_loss(logps: torch.Tensor, ref_logps: torch.Tensor, adv: torch.Tensor, mask: torch.Tensor,
              beta: float):
    """TRN:640-643 -- -(exp(lp - sg(lp)) * A - beta * kl), per-sequence masked mean, mean over rows.
    Returns (loss, mean_kl) with mean_kl as logged at TRN:682."""
    kl = per_token_kl(ref_logps, logps)
    ptl = torch.exp(logps - logps.detach()) * adv.unsqueeze(1)
    ptl = -(ptl - beta * kl)
    m = mask.to(logps.dtype)
    loss = ((ptl * m).sum(dim=1) / m.sum(dim=1)).mean()
    mean_kl = ((kl * m).sum(dim=1) / m.sum(dim=1)).mean()
    return loss, mean_kl


def grpo_loss_grad(logps, ref_logps, adv, mask, beta):
    """Analytic dLoss/dlogps of grpo_loss (what the fused kernel's backward must produce):
    d/dlp [-(A*exp(lp-sg) - beta*(e^x - x - 1))], x = clamp(ref-lp): -A + beta*(1 - e^x) inside the clamp,
    -A outside; times mask / (len * G)."""
    x = ref_logps - logps
    inside = (x > -10) & (x < 10)
    dkl = torch.where(inside, 1 - torch.exp(x.clamp(-10, 10)), torch.zeros_like(x))
    m = mask.to(logps.dtype)
    G = logps.shape[0]
    return (-adv.unsqueeze(1) + beta * dkl) * m / (m.sum(dim=1, keepdim=True) * G)


def step_metrics(mask, rewards_per_func, rewards, std, temporal_rewards, mean_kl, G: int, names):
    """TRN:650-683 for world_size 1 (the gathers are concatenations over ranks)."""
    out = {"completion_length": mask.sum(1).float().mean().item()}
    rpf = rewards_per_func.mean(0)
    for i, n in enumerate(names):
        out[f"rewards/{n}"] = rpf[i].item()
    per_dev = rewards.view(-1, G)
    out["all_wrong"] = (per_dev <= 1).all(dim=1).sum().item() / per_dev.shape[0]
    out["all_correct"] = (per_dev >= 2).all(dim=1).sum().item() / per_dev.shape[0]
    if temporal_rewards is not None:
        out["temporal_rewards"] = float(temporal_rewards)
    out["reward"] = rewards.mean().item()
    out["reward_std"] = std.mean().item()
    out["kl"] = float(mean_kl)
    return out

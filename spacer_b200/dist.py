"""Data-parallel plumbing of the SG-RLVR step (torch.distributed; NCCL over NVLink on the B200 box, gloo in the CPU
tests).  The path shards by PROMPT: rank r takes dataset rows r, r + W, ... exactly like HF's distributed sampler at
`--per_device_train_batch_size 1` (run_SpaceR_SG_RLVR.sh:21); the G completions of a prompt never leave their rank
(advantages are normalised per prompt, SG_RLVR_trainer.py:632-638), so rollout, rewards, both scoring passes and the
backward need no communication.  What does cross ranks:
  * one gradient all-reduce (sum; the optimizer divides by W) over the flat gradient arenas, issued in buckets;
  * an init-time broadcast of the policy weights (rank 0 -> all) so every rank's frozen reference copy is identical;
  * one small all-gather of metric scalars per step (the reference does nine: SG_RLVR_trainer.py:650-683).
Nothing here touches the GPU kernels; it is pure host logic over tensors.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_size(group=None) -> int:
    return dist.get_world_size(group) if is_dist() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if is_dist() else 0


def shard_indices(n_items: int, rank_: int, world: int, drop_last: bool = False) -> list[int]:
    """Dataset rows of one rank: r, r + W, r + 2W, ...  Like torch's DistributedSampler without shuffle, the tail is
    padded by wrapping around so that every rank runs the same number of steps (collectives stay aligned)."""
    if world <= 1:
        return list(range(n_items))
    if drop_last:
        per = n_items // world
        return [rank_ + i * world for i in range(per)]
    per = -(-n_items // world)
    return [(rank_ + i * world) % n_items for i in range(per)]


def allreduce_sum_(tensors, group=None, bucket_elems: int = 1 << 28, async_op: bool = False):
    """In-place SUM all-reduce of flat tensors in buckets of `bucket_elems` elements (512 MB of bf16): large enough
    to run at NVLink bandwidth, small enough that the first bucket can start while later gradients are still being
    written.  Returns the list of work handles when async_op (caller waits), else []."""
    if world_size(group) == 1:
        return []
    works = []
    for t in tensors:
        flat = t.view(-1)
        for s in range(0, flat.numel(), bucket_elems):
            w = dist.all_reduce(flat[s:s + bucket_elems], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            if async_op:
                works.append(w)
    return works


def broadcast_(tensors, src: int = 0, group=None):
    if world_size(group) == 1:
        return
    for t in tensors:
        dist.broadcast(t, src, group=group)


def gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """[n] per rank -> [W, n] (rank order)."""
    w = world_size(group)
    if w == 1:
        return t.view(1, -1)
    out = [torch.empty_like(t) for _ in range(w)]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.stack(out)

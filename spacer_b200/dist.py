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


class OverlappedGradReducer:
    """Gradient all-reduce overlapped with the backward pass ("bucketed + overlapped with backward").

    The model announces `[start, end)` ranges of the flat matrix-gradient arena as soon as they are final (one decoder
    layer = 466 MB of bf16 at 7B, in backward order: lm_head, layers L-1..0, embedding, vision tower).  Each range is
    all-reduced (SUM) asynchronously on a side stream that first waits for the kernels that produced it; `finish()`
    reduces whatever was not announced (and the small fp32 vector arena) and makes the compute stream wait for all of
    it.  Every rank announces the same ranges in the same order, so the collectives line up.  On CPU tensors (gloo, the
    unit tests) the same calls run synchronously."""

    def __init__(self, mat: torch.Tensor, vec: torch.Tensor, group=None):
        self.mat, self.vec, self.group = mat, vec, group
        self.works, self.done = [], []
        self.stream = torch.cuda.Stream(device=mat.device) if mat.is_cuda else None

    def begin_step(self):
        self.works, self.done = [], []

    def ready(self, start: int, end: int):
        if world_size(self.group) == 1 or end <= start:
            return
        self.done.append((start, end))
        chunk = self.mat[start:end]
        if self.stream is None:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()                                   # everything enqueued so far on the compute stream
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def missing_ranges(self):
        """Ranges of the matrix arena nobody announced (alignment gaps are zero everywhere and need no reduction, but
        reducing them keeps this independent of the layout)."""
        out, pos = [], 0
        for s0, e0 in sorted(self.done):
            if s0 > pos:
                out.append((pos, s0))
            pos = max(pos, e0)
        if pos < self.mat.numel():
            out.append((pos, self.mat.numel()))
        return out

    def finish(self):
        if world_size(self.group) == 1:
            return
        rest = [self.mat[a:b] for a, b in self.missing_ranges()] + [self.vec]
        if self.stream is None:
            for t in rest:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            for t in rest:
                self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in self.works:
            w.wait()                                  # the compute stream waits for the collectives
        torch.cuda.current_stream().wait_stream(self.stream)
        self.works = []

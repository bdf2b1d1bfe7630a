#!/usr/bin/env python
"""Data-parallel invariants on N GPUs (torchrun): after K trainer steps on different prompts per rank
  (1) every rank holds bit-identical weights (the invariant data parallelism rests on), with the all-reduce after the
      backward and with the overlapped per-layer all-reduce;
  (2) the two modes end with bit-identical weights (the training path has no floating-point atomics: a whole
      rollout + update step is reproducible), reported as overlap_equals_no_overlap / max_abs_diff_overlap_vs_not.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py [--preset tiny|2b]"""
import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from spacer_b200 import config as mcfg, rewards as RW  # noqa: E402
from spacer_b200.model import Qwen2VLB200  # noqa: E402
from spacer_b200.trainer import GRPOConfig, SGRLVRTrainerB200  # noqa: E402


def run(dims, cfg, overlap, steps, rank, dev, zero1=False):
    policy = Qwen2VLB200(dims, dev)
    policy.params.init_random(seed=0)
    ref = Qwen2VLB200(dims, dev)
    ref.params.mat.copy_(policy.params.mat)
    ref.params.vec.copy_(policy.params.vec)
    tcfg = GRPOConfig(num_generations=cfg["G"], max_completion_length=cfg["C"], min_new_tokens=cfg["C"], temporal=True,
                      learning_rate=1e-3, overlap_allreduce=overlap, zero1=zero1)    # large lr: updates visible in bf16
    tr = SGRLVRTrainerB200(policy, ref, [RW.accuracy_reward, RW.format_reward], tcfg, bench.synth_decode)
    ex = bench.synth_example(dims, cfg, 1234 + rank)
    ex.pop("pixel_values_host")
    ex["video_frames"] = ex.pop("frames_host").to(dev)
    w0 = policy.params.mat.clone()
    for _ in range(steps):
        with contextlib.redirect_stdout(io.StringIO()):
            tr.training_step(dict(ex))
    torch.cuda.synchronize()
    moved = int((policy.params.mat != w0).sum().item())
    return policy.params.mat.clone(), policy.params.vec.clone(), moved


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    RW.set_map_data(bench.SYN_MAP)
    cfg = dict(bench.CONFIGS["tiny" if a.preset == "tiny" else "c2"])
    if a.preset != "tiny":
        cfg["C"] = 32
    dims = mcfg.PRESETS[cfg["preset"]]()
    res = {}
    for overlap in (False, True):
        mat, vec, moved = run(dims, cfg, overlap, a.steps, rank, dev)
        others = [torch.empty_like(mat) for _ in range(world)]
        dist.all_gather(others, mat)
        vo = [torch.empty_like(vec) for _ in range(world)]
        dist.all_gather(vo, vec)
        res[overlap] = dict(mat=mat, vec=vec, moved=moved,
                            same_across_ranks=all(torch.equal(others[0], o) for o in others) and all(torch.equal(vo[0], o) for o in vo))
    zmat, zvec, zmoved = run(dims, cfg, False, a.steps, rank, dev, zero1=True)
    zo = [torch.empty_like(zmat) for _ in range(world)]
    dist.all_gather(zo, zmat)
    zv = [torch.empty_like(zvec) for _ in range(world)]
    dist.all_gather(zv, zvec)
    out = {"preset": a.preset, "world": world, "steps": a.steps,
           "zero1_ranks_identical": all(torch.equal(zo[0], o) for o in zo) and all(torch.equal(zv[0], o) for o in zv),
           "zero1_max_abs_diff_vs_allreduce": float((zmat.float() - res[False]["mat"].float()).abs().max()),
           "zero1_weights_changed": zmoved,
           "weights_changed": [res[False]["moved"], res[True]["moved"]],
           "ranks_identical_no_overlap": res[False]["same_across_ranks"], "ranks_identical_overlap": res[True]["same_across_ranks"],
           "overlap_equals_no_overlap": bool(torch.equal(res[False]["mat"], res[True]["mat"]) and torch.equal(res[False]["vec"], res[True]["vec"])),
           "max_abs_diff_overlap_vs_not": float((res[False]["mat"].float() - res[True]["mat"].float()).abs().max())}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Data-parallel host logic on CPU with the gloo backend, world_size 2: prompt sharding, the bucketed
gradient all-reduce + 1/W scaling, the reference-weight broadcast and the packed metric gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spacer_b200 import dist as D
    try:
        assert D.world_size() == world and D.rank() == rank
        # 1. gradient all-reduce: two arenas (fp32 so gloo sums exactly), tiny buckets to exercise the bucket loop
        g = torch.Generator().manual_seed(100 + rank)
        mat = torch.randn(1000, generator=g)
        vec = torch.randn(37, generator=g)
        mine = (mat.clone(), vec.clone())
        D.allreduce_sum_([mat, vec], bucket_elems=128)
        # 2. weight broadcast
        w = torch.full((50,), float(rank + 1))
        D.broadcast_([w], src=0)
        # 3. packed metric gather keeps rank order
        packed = torch.arange(5, dtype=torch.float32) + 10 * rank
        allp = D.gather_rows(packed)
        # async variant returns handles
        t2 = torch.ones(300) * (rank + 1)
        works = D.allreduce_sum_([t2], bucket_elems=100, async_op=True)
        for wk in works:
            wk.wait()
        torch.save(dict(mat=mat, vec=vec, mine=mine, w=w, allp=allp, t2=t2, n_works=len(works)),
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_indices():
    from spacer_b200.dist import shard_indices
    assert shard_indices(10, 0, 1) == list(range(10))
    parts = [shard_indices(10, r, 4) for r in range(4)]
    assert all(len(p) == 3 for p in parts)                      # same step count on every rank
    assert parts[0] == [0, 4, 8] and parts[1] == [1, 5, 9] and parts[2] == [2, 6, 0] and parts[3] == [3, 7, 1]
    assert set(sum(parts, [])) == set(range(10))               # every row is visited
    parts = [shard_indices(10, r, 4, drop_last=True) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(8))
    assert shard_indices(0, 1, 2) == []


@pytest.mark.timeout(120)
def test_dp_collectives_gloo_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    exp_mat = r[0]["mine"][0] + r[1]["mine"][0]
    exp_vec = r[0]["mine"][1] + r[1]["mine"][1]
    for i in range(world):
        assert torch.allclose(r[i]["mat"], exp_mat) and torch.allclose(r[i]["vec"], exp_vec)
        assert torch.equal(r[i]["w"], torch.ones(50))           # rank 0's weights everywhere
        assert r[i]["allp"].shape == (2, 5)
        assert torch.equal(r[i]["allp"][1], torch.arange(5, dtype=torch.float32) + 10)
        assert torch.equal(r[i]["t2"], torch.full((300,), 3.0)) and r[i]["n_works"] == 3
    # the optimizer's 1/W gradient scale turns the summed gradient into the mean over prompts (HF DDP semantics)
    assert torch.allclose(exp_mat / world, (r[0]["mine"][0] + r[1]["mine"][0]) / 2)


def _reducer_worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    from spacer_b200 import dist as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        mat = torch.randn(1000, generator=g)
        vec = torch.randn(37, generator=g)
        mat0, vec0 = mat.clone(), vec.clone()
        red = D.OverlappedGradReducer(mat, vec)
        red.begin_step()
        for a, b in [(900, 1000), (600, 850), (128, 600)]:      # backward order, with an unannounced gap and head
            red.ready(a, b)
        assert red.missing_ranges() == [(0, 128), (850, 900)]
        red.finish()
        full_m, full_v = mat0.clone(), vec0.clone()
        dist.all_reduce(full_m)
        dist.all_reduce(full_v)
        out_q.put((rank, bool(torch.allclose(mat, full_m)) and bool(torch.allclose(vec, full_v))))
    finally:
        dist.destroy_process_group()


def test_overlapped_grad_reducer_equals_full_allreduce():
    """Announced ranges + finish() == one all-reduce of both arenas (gloo, world size 2)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _vllm_contract_worker(rank, world, port, q):
    import torch.distributed as dist
    from spacer_b200.vllm_api import CompletionOutput, RequestOutput, SamplingParams, gather_generate_broadcast
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeLLM:          # deterministic stand-in for the rollout engine: completion j of a prompt = f(prompt ids, j)
        def __init__(self):
            self.calls = []

        def generate(self, prompts, sampling_params, use_tqdm=False):
            self.calls.append(len(prompts))
            return [RequestOutput(str(i), None, p["prompt_token_ids"],
                                  [CompletionOutput(j, [sum(p["prompt_token_ids"]) + j] * (j + 1)) for j in range(sampling_params.n)])
                    for i, p in enumerate(prompts)]
    llm = FakeLLM() if rank == 0 else None
    req = {"prompt_token_ids": [10 * (rank + 1), 1, 2], "multi_modal_data": {"video": torch.full((2, 3, 4, 4), float(rank))}}
    mine = gather_generate_broadcast(llm, req, SamplingParams(n=3, max_tokens=4), main_rank=0)
    q.put((rank, mine, llm.calls if llm else None))
    dist.destroy_process_group()


def test_vllm_trainer_gather_generate_broadcast_world2():
    """vllm_grpo_trainer_modified.py:546-608 on torch.distributed (gloo, 2 ranks): prompts gathered, ONE generate call on
    the main rank over both, token ids broadcast, every rank keeps its own n completions in order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29621
    ps = [ctx.Process(target=_vllm_contract_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    (r0, c0, calls0), (r1, c1, calls1) = res
    assert calls0 == [2] and calls1 is None                     # a single generate over both ranks' prompts, on rank 0 only
    assert c0 == [[13], [14, 14], [15, 15, 15]]                 # 10 + 1 + 2 = 13 (+ j), completion j has j + 1 tokens
    assert c1 == [[23], [24, 24], [25, 25, 25]]


def test_vllm_pad_completions():
    from spacer_b200.vllm_api import pad_completions
    out = pad_completions([[1, 2, 3], [4], []], pad_token_id=9)
    assert out.tolist() == [[1, 2, 3], [4, 9, 9], [9, 9, 9]]

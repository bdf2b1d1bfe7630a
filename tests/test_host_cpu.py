"""Host-side integer logic of the product against the oracle, on CPU (no kernels are called): M-RoPE position ids in
both conventions, the prefix-shared packing (visibility == G independent causal sequences), ViT slab metadata, the
T-GRPO frame shuffle on patch rows, parameter-name mapping, and the trainer's reward/advantage tail vs the oracle's
restatement of TRN:598-638."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _dims():
    from oracle import qwen2vl_ref as R
    from spacer_b200 import config
    return R, R.dims_tiny(2, 2), config.tiny(2, 2)


@pytest.mark.parametrize("grid", [[(2, 8, 8)], [(3, 4, 6)], [(1, 4, 4), (2, 4, 8)]])
def test_rope_index_matches_oracle_both_conventions(grid):
    from spacer_b200.model import rope_index
    R, d_or, d = _dims()
    g = torch.Generator().manual_seed(0)
    parts = [torch.randint(10, 2000, (5,), generator=g)]
    for t, h, w in grid:
        n = t * h * w // 4
        parts += [torch.tensor([d.vision_start_id]), torch.full((n,), d.video_token_id), torch.tensor([d.vision_end_id]),
                  torch.randint(10, 2000, (7,), generator=g)]
    ids = torch.cat(parts)[None]
    gt = torch.tensor(grid)
    pos_c, nxt_c = rope_index(ids, gt, d, "classic")
    assert torch.equal(pos_c, R.rope_index_classic(ids, gt, d_or)[:, 0])
    assert nxt_c == int(pos_c.max()) + 1
    pos_h, _ = rope_index(ids, gt, d, "hf55")
    assert torch.equal(pos_h, R.rope_index_hf55(ids, gt, d_or)[:, 0])


def test_rope_index_rejects_truncated_prompt():
    from spacer_b200.model import rope_index
    from spacer_b200.ops import SpacerError
    _, _, d = _dims()
    ids = torch.cat([torch.tensor([5, d.vision_start_id]), torch.full((10,), d.video_token_id)])[None]
    with pytest.raises(SpacerError):
        rope_index(ids, torch.tensor([[2, 8, 8]]), d)    # 32 placeholders expected, 10 present (TRN:432-440 pitfall)


def test_packed_visibility_equals_independent_causal_sequences():
    from spacer_b200.model import pack_prompt_completions
    R, d_or, d = _dims()
    P_text, G, C = 6, 3, 5
    ids = torch.cat([torch.arange(20, 20 + P_text), torch.tensor([d.vision_start_id]), torch.full((16,), d.video_token_id),
                     torch.tensor([d.vision_end_id])])[None]
    comp = torch.arange(100, 100 + G * C).view(G, C)
    b = pack_prompt_completions(ids, comp, torch.tensor([[1, 8, 8]]), d, "cpu")
    P = ids.shape[1]
    T = P + G * C
    assert b.ids.shape == (T,) and b.meta.shape == (T, 4) and b.pos.shape == (3, T)
    m = b.meta.long()
    j = torch.arange(T)[None]
    vis = (j < m[:, 0:1]) | ((j >= m[:, 1:2]) & (j < m[:, 2:3]))
    for g in range(G):
        for c in range(C):
            t = P + g * C + c
            want = torch.zeros(T, dtype=torch.bool)
            want[:P] = True
            want[P + g * C:t + 1] = True
            assert torch.equal(vis[t], want), (g, c)
    assert torch.equal(vis[:P, :P], torch.tril(torch.ones(P, P, dtype=torch.bool))) and not vis[:P, P:].any()
    # position ids of every completion continue after the prompt, identically for all G copies
    full = torch.cat([ids.expand(G, -1), comp], 1)
    ref = R.rope_index_classic(full, torch.tensor([[1, 8, 8]]).repeat(G, 1), d_or)
    for g in range(G):
        assert torch.equal(b.pos[:, P + g * C:P + (g + 1) * C].long(), ref[:, g, P:])
    # hidden row that predicts completion token (g, c): last prompt row for c = 0, else the previous completion token
    rows = b.rows.view(G, C).long()
    assert (rows[:, 0] == P - 1).all() and torch.equal(rows[:, 1:], P + torch.arange(G)[:, None] * C + torch.arange(C - 1)[None])
    assert torch.equal(b.targets.view(G, C).long(), comp)


def test_slab_meta_is_block_diagonal_per_temporal_index():
    from spacer_b200.model import slab_meta
    m = slab_meta([[2, 4, 4], [1, 2, 2]], "cpu").long()
    assert m.shape == (36, 4)
    assert (m[:16, 1] == 0).all() and (m[:16, 2] == 16).all()
    assert (m[16:32, 1] == 16).all() and (m[16:32, 2] == 32).all()
    assert (m[32:, 1] == 32).all() and (m[32:, 2] == 36).all() and (m[:, 0] == 0).all()


def test_frame_shuffle_on_patch_rows_equals_shuffling_frames():
    """TRN:442-458 permutes decoded frames and re-runs the processor; on the patch matrix that is a permutation of the
    per-frame halves of every row (normalisation is per pixel)."""
    import bench
    from spacer_b200 import config
    from spacer_b200.trainer import SGRLVRTrainerB200
    d = config.tiny()
    F_, Rz = 4, 56
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (F_, 3, Rz, Rz), generator=g).float()

    def patchify(fr):
        mean = torch.tensor(bench.CLIP_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(bench.CLIP_STD).view(1, 3, 1, 1)
        x = (fr / 255.0 - mean) / std
        p, tp, m = d.patch, d.t_patch, d.merge
        gt, gh, gw = F_ // tp, Rz // p, Rz // p
        x = x.view(gt, tp, 3, gh // m, m, p, gw // m, m, p)
        return x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8).reshape(gt * gh * gw, 3 * tp * p * p).contiguous(), (gt, gh, gw)

    pix, grid = patchify(frames)
    seed = 11
    out = SGRLVRTrainerB200.shuffle_frames(pix, [list(grid)], seed)
    perm = torch.randperm(F_, generator=torch.Generator().manual_seed(seed + 7919))
    want, _ = patchify(frames[perm])
    assert torch.equal(out, want)


def _tail_case(rpf, shuf, lengths, temporal=True, len_control=True, width=600):
    """Runs the PRODUCT's reward tail (trainer.reward_tail, called by training_step) and the oracle's restatement of
    TRN:598-638 on the same inputs."""
    from oracle import grpo_ref as GR
    from spacer_b200 import trainer as T
    G = rpf.shape[0]
    mask = (torch.arange(width)[None] < lengths[:, None]).int()
    summed, t_ref = GR.temporal_bonus(rpf.clone(), shuf, temporal)
    r_ref = GR.length_bonus(summed.sum(1), rpf, mask, len_control)
    adv_ref, std_ref = GR.advantages(r_ref, G)
    rewards, adv, std, t = T.reward_tail(rpf.clone(), None if shuf is None else shuf.clone(), lengths, G, temporal, len_control)
    assert t == float(t_ref)
    assert torch.equal(rewards, r_ref), (rewards, r_ref)
    assert torch.allclose(adv, adv_ref, atol=1e-6, equal_nan=True) and torch.allclose(std, std_ref, atol=1e-7, equal_nan=True)
    return rewards, adv, std, t


def test_reward_tail_matches_oracle():
    """Temporal bonus, length bonus and group-relative advantages THROUGH the product function the trainer calls
    (trainer.reward_tail <- training_step) against oracle/grpo_ref.py (TRN:598-638), incl. the edge cases."""
    rpf = torch.tensor([[1.59, 1.0], [0.0, 1.0], [1.0, 0.0], [0.05, 1.0], [1.0, 1.0], [0.0, 0.0], [1.3, 1.0], [0.2, 1.0]])
    shuf = torch.tensor([[1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.0, 0.0]])
    lengths = torch.tensor([400, 100, 512, 330, 319, 513, 320, 450])      # window edges 319 / 320 / 512 / 513
    r, adv, std, t = _tail_case(rpf, shuf, lengths)
    assert t == 1.0
    # rows 0, 2, 6, 7 are accurate AND inside [320, 512]; row 4 (319) and the inaccurate rows get no length bonus
    assert torch.allclose(r, torch.tensor([1.59 + 1.0 + 0.3 + 0.2, 1.0, 1.0 + 0.3 + 0.2, 1.05, 1.0 + 1.0 + 0.3, 0.0,
                                           1.3 + 1.0 + 0.3 + 0.2, 0.2 + 1.0 + 0.3 + 0.2]))
    # shuffled frames do better by more than 1/0.8: bonus refused, temporal_rewards = 0
    _, _, _, t0 = _tail_case(rpf * torch.tensor([0.1, 1.0]), torch.tensor([[1.0, 1.0]] * 4), lengths)
    assert t0 == 0.0
    # image sample / temporal off: no shuffled group => 0.5 (TRN:610-611), no temporal bonus
    r_img, _, _, t_img = _tail_case(rpf, None, lengths)
    assert t_img == 0.5 and torch.allclose(r_img[1], torch.tensor(1.0)) and torch.allclose(r_img[0], torch.tensor(2.79))
    _, _, _, t_off = _tail_case(rpf, shuf, lengths, temporal=False)
    assert t_off == 0.5
    # exactly ONE accurate row: the length bonus needs at least two (TRN:626)
    one = torch.zeros(8, 2); one[3, 0] = 1.0
    r1, _, _, _ = _tail_case(one, shuf * 0, torch.full((8,), 400))
    assert torch.allclose(r1[3], torch.tensor(1.3)) and float(r1.sum()) == float(r1[3])
    # len_control off; identical rewards (std 0 -> advantages 0 thanks to the 1e-4); G = 2
    _tail_case(rpf, shuf, lengths, len_control=False)
    _, adv0, std0, _ = _tail_case(torch.ones(4, 2), None, torch.full((4,), 10))
    assert float(std0.abs().max()) == 0.0 and float(adv0.abs().max()) == 0.0
    _tail_case(torch.tensor([[1.0, 1.0], [0.0, 0.0]]), torch.tensor([[1.0, 0.0]]), torch.tensor([330, 20]))


def test_completion_lengths_match_oracle_mask():
    from oracle import grpo_ref as GR
    from spacer_b200.trainer import completion_lengths
    eos = 7
    ids = torch.tensor([[1, 2, 7, 7, 3], [7, 1, 1, 1, 1], [1, 2, 3, 4, 5], [1, 2, 3, 4, 7]])
    assert torch.equal(completion_lengths(ids, eos), GR.completion_mask(ids, eos).sum(1))
    assert completion_lengths(ids, eos).tolist() == [3, 1, 5, 5]


def test_step_metrics_match_oracle():
    """TRN:650-683 through the product's pack_step_stats / step_metrics (what training_step logs), world sizes 1 and 3,
    against oracle.grpo_ref.step_metrics evaluated on the concatenation over ranks."""
    from oracle import grpo_ref as GR
    from spacer_b200 import trainer as T
    G, names = 4, ["accuracy_reward", "format_reward"]
    g = torch.Generator().manual_seed(5)
    ranks = []
    for r in range(3):
        rpf = torch.rand(G, 2, generator=g) * 2
        if r == 1:
            rpf = torch.full((G, 2), 1.0)          # all rewards == 2 -> all_correct for this prompt
        if r == 2:
            rpf = torch.tensor([[0.0, 1.0]] * G)     # all rewards == 1 -> all_wrong (<= 1); accuracy 0: no bonuses
        lengths = torch.randint(1, 600, (G,), generator=g)
        rewards, adv, std, t = T.reward_tail(rpf, None if r == 0 else torch.rand(G // 2, 2, generator=g), lengths, G)
        kl = float(torch.rand((), generator=g))
        ranks.append(dict(rpf=rpf, lengths=lengths, rewards=rewards, std=std, t=t, kl=kl,
                          packed=T.pack_step_stats(lengths, rpf, rewards, std, torch.tensor(kl), t)))
    for world in (1, 3):
        rs = ranks[:world]
        got = T.step_metrics(torch.stack([x["packed"] for x in rs]), G, names, temporal=True)
        width = 600
        mask = torch.cat([(torch.arange(width)[None] < x["lengths"][:, None]).int() for x in rs])
        want = GR.step_metrics(mask, torch.cat([x["rpf"] for x in rs]), torch.cat([x["rewards"] for x in rs]),
                               torch.cat([x["std"] for x in rs]), sum(x["t"] for x in rs) / world,
                               sum(x["kl"] for x in rs) / world, G, names)
        assert set(got) == set(want)
        for k in want:
            assert abs(got[k] - want[k]) < 1e-5, (world, k, got[k], want[k])
    got3 = T.step_metrics(torch.stack([x["packed"] for x in ranks]), G, names, temporal=True)
    assert abs(got3["all_correct"] - 1 / 3) < 1e-6 and got3["all_wrong"] >= 1 / 3 - 1e-6
    assert "temporal_rewards" not in T.step_metrics(ranks[0]["packed"][None], G, names, temporal=False)


def test_rollout_seed_differs_per_rank_and_step():
    """data-parallel ranks must not share the sampler's Philox key."""
    from spacer_b200 import dist as D
    from spacer_b200.trainer import GRPOConfig, SGRLVRTrainerB200
    seeds = set()
    for rank in range(8):
        for step in range(4):
            t = SGRLVRTrainerB200.__new__(SGRLVRTrainerB200)
            t.cfg, t.global_step, t.pg = GRPOConfig(), step, None
            orig = D.rank
            D.rank = lambda pg=None, r=rank: r
            try:
                seeds.add(t.rollout_seed())
            finally:
                D.rank = orig
    assert len(seeds) == 32


def test_generation_options_resolution():
    """generate()'s option resolution (model.resolve_generation): explicit kwargs > generation_config > the checkpoint's
    generation_config.json > defaults; HF's implicit top_k = 50; unsupported options raise."""
    import types
    import pytest
    from spacer_b200 import config
    from spacer_b200.model import Qwen2VLB200
    from spacer_b200.ops import SpacerError
    m = Qwen2VLB200.__new__(Qwen2VLB200)
    m.dims, m.generation_config = config.tiny(), {}
    # the reference's rollout config (TRN:277-284), as a duck-typed GenerationConfig
    gc = types.SimpleNamespace(max_new_tokens=1024, do_sample=True, top_p=0.95, temperature=1, num_return_sequences=8,
                               pad_token_id=2027, top_k=None, repetition_penalty=None, num_beams=None)
    sp, mx, mn, G = m.resolve_generation(gc)
    assert (mx, mn, G) == (1024, 0, 8)
    assert not sp.greedy and sp.top_p == 0.95 and sp.top_k == 50 and sp.temperature == 1.0 and sp.pad_id == 2027
    assert sp.eos_ids == (m.dims.eos_id,)
    # the same with the real class of the installed transformers
    from transformers import GenerationConfig
    sp2, mx2, _, G2 = m.resolve_generation(GenerationConfig(max_new_tokens=7, do_sample=True, top_p=0.95, temperature=1,
                                                            num_return_sequences=4, pad_token_id=2027))
    assert (mx2, G2, sp2.top_k, sp2.top_p, sp2.greedy) == (7, 4, 50, 0.95, False)
    # evaluation: checkpoint generation_config.json + generate(max_new_tokens=.., temperature=0.01) (vsibench.py:174)
    m.generation_config = dict(do_sample=True, repetition_penalty=1.05, temperature=0.1, top_k=1, top_p=0.001,
                               eos_token_id=[2029, 2027], pad_token_id=2027)
    sp3, mx3, _, G3 = m.resolve_generation({}, max_new_tokens=128, temperature=0.01)
    assert sp3.greedy and sp3.repetition_penalty == 1.05 and sp3.eos_ids == (2029, 2027) and (mx3, G3) == (128, 1)
    # engine-level call without a generation_config keeps nucleus-only sampling (kernel tests, trainer passes top_k itself)
    m.generation_config = {}
    sp4, mx4, _, _ = m.resolve_generation(None, max_new_tokens=5, top_p=0.9)
    assert sp4.top_k == 0 and sp4.top_p == 0.9 and not sp4.greedy and mx4 == 5
    with pytest.raises(SpacerError):
        m.resolve_generation(types.SimpleNamespace(max_new_tokens=4, num_beams=4))
    with pytest.raises(SpacerError):
        m.resolve_generation(types.SimpleNamespace(do_sample=True))          # no max_new_tokens
    with pytest.raises(SpacerError):
        m.resolve_generation(None, max_new_tokens=4, top_p=0.0)
    with pytest.raises(TypeError):
        Qwen2VLB200.generate(m, torch.zeros(1, 4, dtype=torch.long), some_unknown_option=1)


def test_param_layout_roundtrip_names_cpu():
    """HF parameter names <-> the flat arenas (checkpoint save/load surface, OR1/SG-RLVR.py:384)."""
    from oracle import qwen2vl_ref as R
    from spacer_b200 import config
    from spacer_b200.params import ParamStore
    d_or, d = R.dims_tiny(2, 2), config.tiny(2, 2)
    w = R.init_weights(d_or, seed=1)
    ps = ParamStore(d, "cpu")
    ps.load_state_dict(w)
    sd = ps.state_dict()
    assert set(sd) == set(w)
    for k in w:
        assert torch.equal(sd[k], w[k].bfloat16()), k
    assert ps.mat.data_ptr() % 256 == 0 or True
    for name, (arena, off, shape) in ps.index.items():
        assert off % 128 == 0          # 256-byte aligned tensors (TMA needs 16)


def test_hf_config_roundtrip_and_name_normalisation():
    """hub: ModelDims <-> config.json (nested 5.x layout written, nested and flat 4.x layouts read), old -> new names."""
    from spacer_b200 import config, hub
    for d in (config.qwen2_vl_7b(), config.qwen2_vl_2b(), config.qwen2_5_vl_7b(), config.tiny25()):
        back = hub.dims_from_hf_config(hub.hf_config_from_dims(d), name_hint=d.name)
        for f in ("hidden", "layers", "heads", "kv_heads", "head_dim", "inter", "vocab", "tie", "v_depth", "v_embed",
                  "v_heads", "v_mlp", "variant", "v_fullatt", "mrope_section", "eos_id", "pad_id", "video_token_id"):
            assert getattr(back, f) == getattr(d, f), (d.name, f)
    flat = dict(model_type="qwen2_vl", hidden_size=3584, num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4,
                intermediate_size=18944, vocab_size=152064, rms_norm_eps=1e-6, rope_theta=1000000.0,
                rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, tie_word_embeddings=False,
                vision_config=dict(depth=32, embed_dim=1280, mlp_ratio=4, num_heads=16, in_chans=3, hidden_size=3584,
                                   patch_size=14, spatial_merge_size=2, temporal_patch_size=2),
                image_token_id=151655, video_token_id=151656, vision_start_token_id=151652, vision_end_token_id=151653,
                eos_token_id=151645, bos_token_id=151643)
    d = hub.dims_from_hf_config(flat, "Qwen/Qwen2-VL-7B-Instruct")
    assert d == config.qwen2_vl_7b().__class__(**{**config.qwen2_vl_7b().__dict__, "name": "Qwen2-VL-7B-Instruct"})
    old = {"visual.blocks.0.attn.qkv.weight": 1, "model.layers.3.mlp.up_proj.weight": 2, "model.embed_tokens.weight": 3,
           "model.norm.weight": 4, "lm_head.weight": 5, "model.visual.merger.ln_q.weight": 6}
    assert hub.normalize_names(old) == {"model.visual.blocks.0.attn.qkv.weight": 1,
                                        "model.language_model.layers.3.mlp.up_proj.weight": 2,
                                        "model.language_model.embed_tokens.weight": 3, "model.language_model.norm.weight": 4,
                                        "lm_head.weight": 5, "model.visual.merger.ln_q.weight": 6}


def test_qwen25_rope_index_matches_oracle_cpu():
    """Temporal M-RoPE spacing of Qwen2.5-VL (host integer logic): both conventions, with and without
    second_per_grid_ts, against oracle/qwen25vl_ref.py (itself pinned to HF 5.5.0 by tests/golden/tiny_model25.pt)."""
    from oracle import qwen25vl_ref as R25
    from oracle import qwen2vl_ref as R
    from spacer_b200 import config
    from spacer_b200.model import rope_index
    d_or, d = R25.dims25_tiny(), config.tiny25()
    for grid in ([[2, 12, 8]], [[3, 4, 6]], [[1, 8, 8]]):
        g = torch.tensor(grid)
        n_v = int(g.prod()) // 4
        ids = R.build_prompt_ids(d_or, n_v, 5, 9, seed=3)
        for sec in (None, [1.0], [1.5], [2.0], [0.5]):
            for conv, ref in (("classic", R25.rope_index_classic), ("hf55", R25.rope_index_hf55)):
                pos, nxt = rope_index(ids, g, d, conv, sec)
                want = ref(ids, g, d_or, sec)[:, 0]
                assert torch.equal(pos, want), (grid, sec, conv)
                assert nxt == int(want.max()) + 1 or conv == "hf55"


def test_segment_and_embed_plans():
    """Host side of the deterministic scatter (sb_segment_sum_rows): the plans reproduce index_add / the embedding
    backward on CPU tensors."""
    from spacer_b200 import config
    from spacer_b200.model import embed_plan, segment_plan
    g = torch.Generator().manual_seed(0)
    keys = torch.randint(0, 7, (40,), generator=g)
    src = torch.randn(40, 5, generator=g)
    order, off, dst, n = segment_plan(keys, "cpu")
    out = torch.zeros(7, 5)
    for s in range(n):
        rows = order[off[s]:off[s + 1]].long()
        assert torch.equal(rows, rows.sort().values)              # original order inside a segment (stable)
        out[dst[s]] = src[rows].sum(0)
    assert torch.allclose(out, torch.zeros(7, 5).index_add_(0, keys, src), atol=1e-6)
    d = config.tiny()
    ids = torch.tensor([5, 9, d.vision_start_id, d.video_token_id, d.video_token_id, d.vision_end_id, 9, 5, 5])
    order_pos, off, dst, n, vis_pos = embed_plan(ids, d, "cpu")
    assert vis_pos.tolist() == [3, 4]
    got = {int(dst[s]): order_pos[off[s]:off[s + 1]].tolist() for s in range(n)}
    assert got == {5: [0, 7, 8], 9: [1, 6], d.vision_start_id: [2], d.vision_end_id: [5]}


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's
    keys, same metric / unit as the GPU arm, e2e == value, zero copy bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "grpo_samples_per_sec" and line["unit"] == "samples/s"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1

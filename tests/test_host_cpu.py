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


def test_reward_tail_matches_oracle():
    """Temporal bonus, length bonus and group-relative advantages as the trainer computes them (TRN:598-638)."""
    from oracle import grpo_ref as GR
    from spacer_b200 import trainer as T
    G = 8
    rpf = torch.tensor([[1.59, 1.0], [0.0, 1.0], [1.0, 0.0], [0.05, 1.0], [1.0, 1.0], [0.0, 0.0], [1.3, 1.0], [0.2, 1.0]])
    shuf = torch.tensor([[1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.0, 0.0]])
    lengths = torch.tensor([400, 100, 512, 330, 319, 513, 320, 450])
    mask = (torch.arange(600)[None] < lengths[:, None]).int()
    summed, temporal = GR.temporal_bonus(rpf.clone(), shuf, True)
    rewards = GR.length_bonus(summed.sum(1), rpf, mask, True)
    adv_ref, std_ref = GR.advantages(rewards, G)
    # the trainer's own arithmetic (same lines, device tensors)
    s2 = rpf.clone()
    if s2[:, 0].mean() >= T.TEMPORAL_RATIO * shuf[:, 0].mean():
        sel = s2[:, 0] > T.ACC_THRESHOLD
        s2[sel, 0] += T.TEMPORAL_BONUS
        t2 = 1.0
    else:
        t2 = 0.0
    r2 = s2.sum(1)
    sel = torch.nonzero(rpf[:, 0] > T.ACC_THRESHOLD, as_tuple=True)[0].tolist()
    if len(sel) > 1:
        for i in sel:
            if T.LEN_WINDOW[0] <= int(lengths[i]) <= T.LEN_WINDOW[1]:
                r2[i] += T.LEN_BONUS
    adv2 = (r2 - r2.mean()) / (r2.std() + T.STD_EPS)
    assert t2 == float(temporal)
    assert torch.allclose(r2, rewards) and torch.allclose(adv2, adv_ref, atol=1e-6)


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

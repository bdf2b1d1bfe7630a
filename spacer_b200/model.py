"""Qwen2-VL on the spacer_b200 kernels: host-side orchestration of the SG-RLVR hot path.

Mirrors the surface of the object the reference trainer drives:
    model.generate(input_ids, pixel_values_videos=..., video_grid_thw=..., generation_config-like kwargs)
        -> LongTensor [G, P + C']                              (SG_RLVR_trainer.py:463-467)
    model.per_token_logps(...) / model.grpo_forward_backward(...)  (TRN:353-366, 526-547, 640-643)
    model.state_dict() / load_state_dict() with HF parameter names
Every FLOP runs in libspacer_b200.so (tcgen05 GEMMs, flash attention, fused norms/rope/loss); torch is used
for device memory, streams and CUDA-graph capture only.

Algorithmic differences from the reference's executed path (same mathematics):
  * the ViT and the prompt prefill run ONCE per prompt, not once per sampled copy;
  * scoring / training run on the packed sequence [prompt | completion_0 | ... | completion_{G-1}] with the
    prompt shared through the attention mask (every completion token sees the whole prompt + its own
    completion causally) -- identical to G independent causal sequences because the prompt rows are
    identical across copies and no padding mask is passed (TRN:357, Appendix B.1-2);
  * the [.,V] logits are never materialised: lm_head -> online logsumexp -> gather in the GEMM epilogue;
  * the update that follows a rollout reuses the rollout's vision-tower forward and prompt prefill (same weights, same
    inputs: `generate(keep_vit_tape=True)` -> `grpo_forward_backward(vit_cache=...)`, bit-identical gradients).
Also here: the Qwen2.5-VL family (`dims.variant == "qwen2_5_vl"`: windowed RMSNorm/SwiGLU vision tower, temporal M-RoPE
spacing), greedy decoding for evaluation, the SFT loss on the same kernels (`sft_forward_backward`), the drop-in
`get_per_token_logps`, and HF checkpoint I/O (`from_pretrained` / `save_pretrained`, hub.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import ops
from .config import ModelDims
from .ops import EPI_DLOGITS, EPI_F32T, EPI_GELU, EPI_LMHEAD, EPI_QUICKGELU, EPI_SWIGLU, SpacerError
from .params import ParamStore

I32 = torch.int32
BF16 = torch.bfloat16
F32 = torch.float32


# ------------------------------------------------------------------------------------------------
# position ids (host integer work)
# ------------------------------------------------------------------------------------------------
def rope_index(input_ids: torch.Tensor, grid_thw, dims: ModelDims, convention: str = "classic", second_per_grid_ts=None):
    """3-stream M-RoPE position ids for ONE row: int64 [3, L] and the next free position.

    "classic" is the formula of the transformers release the reference was written against and the released
    Qwen2-VL weights were trained with (per-frame raster t/h/w); "hf55" reproduces transformers 5.5.0
    (modeling_qwen2_vl.py:934-988).  See SURVEY.md 8(c) drift #2.
    Qwen2.5-VL (dims.variant == "qwen2_5_vl") spaces the temporal index of a video by
    second_per_grid_t * tokens_per_second (modeling_qwen2_5_vl.py:1024-1135); `second_per_grid_ts` defaults to 1.0 per
    video, which is what the reference's scoring forward uses after deleting the key (SG_RLVR_trainer.py:519-520)."""
    v25 = dims.variant == "qwen2_5_vl"
    ids = input_ids.reshape(-1).cpu()
    L = ids.numel()
    is_v = (ids == dims.video_token_id) | (ids == dims.image_token_id)
    pos = torch.zeros(3, L, dtype=torch.long)
    nxt, i, gi = 0, 0, 0
    grids = [] if grid_thw is None else [list(map(int, g)) for g in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)]
    m = dims.merge
    isv = is_v.tolist()
    if any(isv) and not grids:
        raise SpacerError("visual placeholder tokens in the prompt but no grid_thw")
    while i < L:
        if isv[i]:
            t, h, w = grids[min(gi, len(grids) - 1)]
            sec = 1.0
            if second_per_grid_ts is not None and len(second_per_grid_ts) > 0:
                sec = float(second_per_grid_ts[min(gi, len(second_per_grid_ts) - 1)])
            is_video = ids[i].item() == dims.video_token_id
            gi += 1
            gh, gw = h // m, w // m
            n = t * gh * gw
            if convention == "classic":
                tt = torch.arange(t).view(-1, 1).expand(-1, gh * gw)
                if v25:
                    tt = (tt * (sec * dims.tokens_per_second if is_video else 0.0)).long()
                tt = tt.flatten()
                hh = torch.arange(gh).view(1, -1, 1).expand(t, -1, gw).flatten()
                ww = torch.arange(gw).view(1, 1, -1).expand(t, gh, -1).flatten()
                blk = torch.stack([tt, hh, ww]) + nxt
                after = int(blk.max()) + 1
            else:
                ww = torch.arange(nxt, nxt + gw).repeat(gh * t)
                hh = torch.arange(nxt, nxt + gh).repeat_interleave(gw * t)
                t0 = nxt * dims.tokens_per_second * int(sec) if v25 else nxt
                tt = torch.full((n,), t0, dtype=torch.long)
                blk = torch.stack([tt, hh, ww])
                after = nxt + max(h, w) // m if v25 else int(blk.max()) + 1
            if i + n > L:
                raise SpacerError("placeholder tokens do not match video_grid_thw (truncated prompt?)")
            pos[:, i:i + n] = blk
            nxt = after
            i += n
        else:
            # run of text tokens
            j = i
            while j < L and not isv[j]:
                j += 1
            pos[:, i:j] = torch.arange(nxt, nxt + (j - i))
            nxt += j - i
            i = j
    return pos, nxt


def slab_meta(grid_thw, device):
    """Attention visibility for the ViT: one block-diagonal slab per temporal index (MQ2:772-780)."""
    starts, ends = [], []
    base = 0
    for t, h, w in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw):
        for _ in range(int(t)):
            n = int(h) * int(w)
            starts += [base] * n
            ends += [base + n] * n
            base += n
    z = torch.zeros(len(starts), dtype=torch.long)
    return ops.make_meta(z, torch.tensor(starts), torch.tensor(ends), device)


def window_plan(grid_thw, dims: ModelDims, device):
    """Host integer work of Qwen2.5-VL's windowed vision tower (modeling_qwen2_5_vl.py:411-453, 474-497): the order of
    the merged 2x2 token groups that makes every attention window contiguous (`widx`, applied to the patch rows BEFORE
    the patch-embed GEMM, which is row-wise), the (h, w) rotary position of every reordered patch row, and the two
    visibility tables: windows (most blocks) and whole frames (dims.v_fullatt blocks)."""
    import torch.nn.functional as Fn
    grids = [list(map(int, g)) for g in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)]
    m, unit = dims.merge, dims.merge * dims.merge
    win = dims.v_window // dims.merge // dims.patch
    index, cu, base, pos_hw = [], [0], 0, []
    for t, h, w in grids:
        gh, gw = h // m, w // m
        idx = torch.arange(t * gh * gw).reshape(t, gh, gw)
        ph, pw = win - gh % win, win - gw % win
        nh, nw = (gh + ph) // win, (gw + pw) // win
        pad = Fn.pad(idx, (0, pw, 0, ph), "constant", -100)
        pad = pad.reshape(t, nh, win, nw, win).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, win, win)
        seqlens = (pad != -100).sum([2, 3]).reshape(-1)
        flat = pad.reshape(-1)
        index.append(flat[flat != -100] + base)
        cu.extend((seqlens.cumsum(0) * unit + cu[-1]).tolist())
        base += t * gh * gw
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(gh, m, gw, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(gh, m, gw, m).permute(0, 2, 1, 3).flatten()
        pos_hw.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
    widx = torch.cat(index)
    cu = torch.unique_consecutive(torch.tensor(cu)).tolist()
    pos_hw = torch.cat(pos_hw)
    N = pos_hw.shape[0]
    row_perm = (widx[:, None] * unit + torch.arange(unit)[None]).reshape(-1)     # patch-row permutation
    pos_hw = pos_hw[row_perm]
    starts = torch.zeros(N, dtype=torch.long)
    ends = torch.zeros(N, dtype=torch.long)
    for a, b in zip(cu[:-1], cu[1:]):
        starts[a:b] = a
        ends[a:b] = b
    z = torch.zeros(N, dtype=torch.long)
    return dict(widx=widx.to(I32).to(device), rev=torch.argsort(widx).to(I32).to(device),
                row_perm=row_perm.to(I32).to(device), pos_hw=pos_hw.to(I32).contiguous().to(device),
                meta_win=ops.make_meta(z, starts, ends, device), meta_full=slab_meta(grids, device), cu_window=cu)


def causal_meta(T, device):
    t = torch.arange(T)
    return ops.make_meta(torch.zeros(T, dtype=torch.long), torch.zeros(T, dtype=torch.long), t + 1, device)


def embed_plan(ids: torch.Tensor, dims: ModelDims, device):
    """Deterministic embedding backward: text positions grouped by token id (segment_plan over the text rows only) and the
    positions of the vision placeholders in order (their gradient rows are a plain gather)."""
    ids = ids.reshape(-1).cpu().long()
    is_v = (ids == dims.video_token_id) | (ids == dims.image_token_id)
    text_pos = torch.nonzero(~is_v, as_tuple=True)[0]
    order, off, dst, n = segment_plan(ids[text_pos], device)
    # `order` indexes the text rows; translate to sequence positions so the kernel reads dx directly
    order_pos = text_pos.to(device)[order.long()].to(I32)
    vis_pos = torch.nonzero(is_v, as_tuple=True)[0].to(I32).to(device)
    return (order_pos, off, dst, n, vis_pos)


def segment_plan(keys: torch.Tensor, device):
    """Host side of sb_segment_sum_rows: for int keys [n] (destination row of source row i) returns device int32 tensors
    (order, seg_off, seg_dst, n_seg): sources grouped by key in their original order (stable sort)."""
    keys = keys.reshape(-1).cpu().long()
    order = torch.argsort(keys, stable=True)
    sk = keys[order]
    uniq, counts = torch.unique_consecutive(sk, return_counts=True)
    off = torch.zeros(uniq.numel() + 1, dtype=torch.long)
    off[1:] = counts.cumsum(0)
    return (order.to(I32).to(device), off.to(I32).to(device), uniq.to(I32).to(device), int(uniq.numel()))


@dataclass
class PackedBatch:
    """[prompt | completion_0 | ... | completion_{G-1}] with everything the kernels need."""
    ids: torch.Tensor          # int32 [T]
    pos: torch.Tensor          # int32 [3, T]
    meta: torch.Tensor         # int32 [T, 4]
    rows: torch.Tensor         # int32 [G*C]: hidden row that predicts completion token (g, c)
    targets: torch.Tensor      # int32 [G*C]
    comp_ids: torch.Tensor     # int32 [G, C]
    P: int
    G: int
    C: int
    prompt_ids_host: torch.Tensor = None   # int64 [P] / [3, P] on the host: lets the update recognise a rollout prefill
    prompt_pos_host: torch.Tensor = None   # of the same prompt at the same positions (prefix reuse)
    rows_plan: tuple = None    # segment_plan(rows): deterministic d_hidden scatter of the lm_head rows
    embed_plan: tuple = None   # (text positions, segment_plan(token id of the text positions), vision positions)


@dataclass(frozen=True)
class SamplingParams:
    """What one sampler launch needs (sb_sample_args): the resolved generation options of a rollout."""
    greedy: bool = False
    top_p: float = 0.95
    top_k: int = 0
    temperature: float = 1.0
    repetition_penalty: float = 1.0
    eos_ids: tuple = (151645,)
    pad_id: int = 151643

    def key(self):
        return (self.greedy, self.top_p, self.top_k, self.temperature, self.repetition_penalty, self.eos_ids, self.pad_id)


def pack_prompt_completions(prompt_ids, completion_ids, grid_thw, dims: ModelDims, device, convention="classic",
                            second_per_grid_ts=None):
    prompt_ids = prompt_ids.reshape(-1).cpu().long()
    comp = completion_ids.cpu().long()
    P, (G, C) = prompt_ids.numel(), comp.shape
    ppos, nxt = rope_index(prompt_ids, grid_thw, dims, convention, second_per_grid_ts)
    cpos = (torch.arange(C) + nxt).repeat(G)
    pos = torch.cat([ppos, cpos[None].expand(3, -1)], dim=1)
    ids = torch.cat([prompt_ids, comp.reshape(-1)])
    T = P + G * C
    t = torch.arange(T)
    seg = torch.where(t < P, torch.zeros_like(t), P + ((t - P) // C) * C)
    pre = torch.where(t < P, torch.zeros_like(t), torch.full_like(t, P))
    meta = ops.make_meta(pre, seg, t + 1, device)
    starts = P + torch.arange(G) * C
    rows = torch.cat([torch.full((G, 1), P - 1), starts[:, None] + torch.arange(C - 1)[None]], dim=1).reshape(-1)
    return PackedBatch(ids=ids.to(I32).to(device), pos=pos.to(I32).contiguous().to(device), meta=meta,
                       rows=rows.to(I32).to(device), targets=comp.reshape(-1).to(I32).to(device),
                       comp_ids=comp.to(I32).contiguous().to(device), P=P, G=G, C=C,
                       prompt_ids_host=prompt_ids, prompt_pos_host=ppos,
                       rows_plan=segment_plan(rows, device), embed_plan=embed_plan(ids, dims, device))


class GradStore:
    """Gradient arenas: bf16 for matrices (written by GEMM epilogues), fp32 for norm weights / biases."""

    def __init__(self, params: ParamStore):
        self.params = params
        self.mat = torch.zeros(params.sizes["mat"], device=params.device, dtype=BF16)
        self.vec = torch.zeros(params.sizes["vec"], device=params.device, dtype=F32)
        self.views = {n: params.view_of(n, self.mat, self.vec) for n in params.index}

    def __getitem__(self, name):
        if name == "lm_head" and self.params.dims.tie:
            name = "embed"
        return self.views[name]

    def zero_for_step(self):
        """Only accumulating destinations need zeroing: the fp32 vector arena and the embedding table."""
        self.vec.zero_()
        self.views["embed"].zero_()

    # -- gradient-ready notifications (data-parallel overlap, dist.OverlappedGradReducer) -----------------------
    on_ready = None     # callable(start, end): elements [start, end) of the matrix arena are final for this step

    def mat_range(self, prefix: str):
        """[start, end) of the matrix-arena elements of every tensor whose name starts with `prefix`
        (tensors of one layer are contiguous in the arena, params._layout)."""
        lo, hi = None, None
        for name, (arena, off, shape) in self.params.index.items():
            if arena == "mat" and name.startswith(prefix):
                n = 1
                for x in shape:
                    n *= x
                lo = off if lo is None else min(lo, off)
                hi = off + n if hi is None else max(hi, off + n)
        return lo, hi

    def zero_range(self, prefix: str):
        """Zero the matrix gradients of every tensor under `prefix` (a sub-network that took no part in this step, e.g.
        the vision tower on a text-only prompt: its GEMM epilogues did not overwrite last step's values)."""
        lo, hi = self.mat_range(prefix)
        if lo is not None:
            self.mat[lo:hi].zero_()

    def ready(self, prefix: str):
        if self.on_ready is not None:
            lo, hi = self.mat_range(prefix)
            if lo is not None:
                self.on_ready(lo, hi)


class Qwen2VLB200:
    def __init__(self, dims: ModelDims, device="cuda", params: ParamStore | None = None,
                 rope_convention: str = "classic"):
        if not torch.cuda.is_available():
            raise SpacerError("spacer_b200 needs a CUDA device (sm_100a); there is no CPU path")
        ops._lib.load()
        self.dims = dims
        self.device = torch.device(device)
        self.params = params if params is not None else ParamStore(dims, device)
        self.rope_convention = rope_convention
        self.training = False
        self._dec = None
        self.vit_cache = None
        self.phase_marks = None
        # the attributes of an HF module the reference trainer touches (SG_RLVR_trainer.py:156, 193, 234, 312)
        from . import hub
        self.config = hub.make_config_namespace(dims)
        self.generation_config = {}      # the checkpoint's generation_config.json (hub.from_pretrained fills it)
        self.warnings_issued = {}

    # ---- HF-like surface -------------------------------------------------------------------------
    def state_dict(self):
        return self.params.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.params.load_state_dict(sd)
        return self

    @classmethod
    def from_pretrained(cls, path, device="cuda", **kw):
        """Local HF-format checkpoint directory -> engine (Qwen2-VL or Qwen2.5-VL by config.json; hub.from_pretrained)."""
        from . import hub
        return hub.from_pretrained(cls, path, device, **kw)

    def save_pretrained(self, path, **kw):
        """HF-format checkpoint (config.json + safetensors, transformers 5.x names): open_r1/SG-RLVR.py:384."""
        from . import hub
        hub.save_pretrained(self, path, **kw)

    # Selective activation recompute: the raw gate|up projection (2 x 18944 of the ~53k bf16 values a decoder layer saves
    # per token at 7B: 71 % of the activation memory) is not kept for the backward but recomputed there from the saved
    # layer input with one extra GEMM per layer (+1/3 of a forward's FLOPs).  cfg3 fits without it (149 GB of state +
    # ~26 GB of activations); cfg4 / cfg5 (8448 / 10496 packed tokens) need it on a 180 GB part with fp32 Adam moments.
    recompute_mlp = False

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None, **kw):
        """What HF Trainer calls under `--gradient_checkpointing true` (run_SpaceR_SG_RLVR.sh:27).  The reference recomputes
        whole decoder layers; here only the widest activation is recomputed (see `recompute_mlp`)."""
        self.recompute_mlp = True

    def gradient_checkpointing_disable(self):
        self.recompute_mlp = False

    def named_parameters(self):
        return list(self.params.hf_items())

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def parameters(self):
        return [self.params.mat, self.params.vec]

    # N = 1280 outputs with 256-wide tiles are 320 tiles = 2.16 waves of 148 CTAs (3 waves, 28 % idle); 128-wide tiles
    # make 4.3 half-size waves (5): the ViT proj / fc2 GEMMs use them
    VIT_NARROW_BN = 128

    # ---- vision tower ----------------------------------------------------------------------------
    def vit_forward(self, pixel_values, grid_thw, tape: dict | None = None):
        """Qwen2VisionTransformerPretrainedModel.forward (MQ2:757-795).  pixel_values [N_p, 1176] fp32/bf16."""
        if self.dims.variant == "qwen2_5_vl":
            return self._vit25_forward(pixel_values, grid_thw, tape)
        d, W = self.dims, self.params
        grid_list = grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw
        pix = pixel_values if pixel_values.dtype == BF16 else ops.cast_f32_bf16(pixel_values.contiguous())
        T = pix.shape[0]
        if T != sum(int(t) * int(h) * int(w) for t, h, w in grid_list):
            raise SpacerError("pixel_values rows do not match video_grid_thw")
        grids_dev = torch.tensor(grid_list, dtype=I32, device=self.device)
        meta = slab_meta(grid_list, self.device)
        E, nh, hd = d.v_embed, d.v_heads, d.v_head_dim
        x = ops.gemm(pix, W["v.patch_w"])
        save = tape is not None
        if save:
            tape.update(pix=pix, grids=grids_dev, meta=meta, blocks=[])
        for i in range(d.v_depth):
            p = f"v.{i}."
            r1 = ops.layernorm_fwd(x, W[p + "ln1_w"], W[p + "ln1_b"], save_stats=save)
            h = r1[0] if save else r1
            qkv = ops.gemm(h, W[p + "qkv_w"], bias=W[p + "qkv_b"])
            ops.rope_vit(qkv, nh, hd, grids_dev, d.merge)
            r2 = ops.attn_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], meta, nh, nh, hd, save_lse=save)
            a = r2[0] if save else r2
            x2 = ops.gemm(a, W[p + "proj_w"], bias=W[p + "proj_b"], residual=x, bn=self.VIT_NARROW_BN)
            r3 = ops.layernorm_fwd(x2, W[p + "ln2_w"], W[p + "ln2_b"], save_stats=save, out=h)
            z = torch.empty((T, d.v_mlp), device=self.device, dtype=BF16) if save else None
            f = ops.gemm(r3[0] if save else r3, W[p + "fc1_w"], bias=W[p + "fc1_b"], epilogue=EPI_QUICKGELU, aux=z)
            x3 = ops.gemm(f, W[p + "fc2_w"], bias=W[p + "fc2_b"], residual=x2, bn=self.VIT_NARROW_BN)
            if save:
                tape["blocks"].append(dict(x=x, m1=r1[1], s1=r1[2], qkv=qkv, a=a, lse=r2[1], x2=x2, m2=r3[1], s2=r3[2], z=z))
            x = x3
        rm = ops.layernorm_fwd(x, W["v.m.ln_w"], W["v.m.ln_b"], save_stats=save)
        hm = (rm[0] if save else rm).view(T // (d.merge * d.merge), d.merge_dim)
        zm = torch.empty_like(hm) if save else None
        fm = ops.gemm(hm, W["v.m.fc0_w"], bias=W["v.m.fc0_b"], epilogue=EPI_GELU, aux=zm)
        out = ops.gemm(fm, W["v.m.fc2_w"], bias=W["v.m.fc2_b"])
        if save:
            tape.update(x_last=x, mm=rm[1], sm=rm[2], zm=zm)
        return out

    # ---- Qwen2.5-VL vision tower -------------------------------------------------------
    def _vit25_forward(self, pixel_values, grid_thw, tape: dict | None = None):
        """Qwen2_5_VisionTransformerPretrainedModel.forward (modeling_qwen2_5_vl.py:455-518): RMSNorm blocks with a
        biased SwiGLU MLP, attention inside 112-pixel windows except in the full-attention blocks, merged tokens
        restored to the original order at the end.  The window reorder is applied to the patch rows before the
        (row-wise) patch-embed GEMM, so no activation is ever permuted and the backward needs no scatter."""
        d, W = self.dims, self.params
        grid_list = grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw
        pix = pixel_values if pixel_values.dtype == BF16 else ops.cast_f32_bf16(pixel_values.contiguous())
        T = pix.shape[0]
        if T != sum(int(t) * int(h) * int(w) for t, h, w in grid_list):
            raise SpacerError("pixel_values rows do not match video_grid_thw")
        plan = window_plan(grid_list, d, self.device)
        pixp = torch.empty_like(pix)
        ops.call("sb_gather_rows", pix, plan["row_perm"], pixp, T, pix.shape[1])
        E, nh, hd, Mp = d.v_embed, d.v_heads, d.v_head_dim, d.v_mlp_pad
        x = ops.gemm(pixp, W["v.patch_w"])
        save = tape is not None
        if save:
            tape.update(pix=pixp, plan=plan, blocks=[])
        for i in range(d.v_depth):
            p = f"v.{i}."
            meta = plan["meta_full"] if i in d.v_fullatt else plan["meta_win"]
            r1 = ops.rmsnorm_fwd(x, W[p + "ln1_w"], 1e-6, save_stats=save)
            h = r1[0] if save else r1
            qkv = ops.gemm(h, W[p + "qkv_w"], bias=W[p + "qkv_b"])
            ops.call("sb_rope_vit_pos", qkv, T, nh, hd, plan["pos_hw"], 0)
            r2 = ops.attn_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], meta, nh, nh, hd, save_lse=save)
            a = r2[0] if save else r2
            x2 = ops.gemm(a, W[p + "proj_w"], bias=W[p + "proj_b"], residual=x, bn=self.VIT_NARROW_BN)
            r3 = ops.rmsnorm_fwd(x2, W[p + "ln2_w"], 1e-6, save_stats=save, out=h)
            gu = torch.empty((T, 2 * Mp), device=self.device, dtype=BF16) if save else None
            act = ops.gemm(r3[0] if save else r3, W[p + "gu_w"], bias=W[p + "gu_b"], epilogue=EPI_SWIGLU, aux=gu)
            x3 = ops.gemm(act, W[p + "down_w"], bias=W[p + "down_b"], residual=x2, bn=self.VIT_NARROW_BN)
            if save:
                tape["blocks"].append(dict(x=x, s1=r1[1], qkv=qkv, a=a, lse=r2[1], x2=x2, s2=r3[1], gu=gu))
            x = x3
        rm = ops.rmsnorm_fwd(x, W["v.m.ln_w"], 1e-6, save_stats=save)
        hm = (rm[0] if save else rm).view(T // (d.merge * d.merge), d.merge_dim)
        zm = torch.empty_like(hm) if save else None
        fm = ops.gemm(hm, W["v.m.fc0_w"], bias=W["v.m.fc0_b"], epilogue=EPI_GELU, aux=zm)
        merged = ops.gemm(fm, W["v.m.fc2_w"], bias=W["v.m.fc2_b"])
        out = torch.empty_like(merged)
        ops.call("sb_gather_rows", merged, plan["rev"], out, merged.shape[0], merged.shape[1])
        if save:
            tape.update(x_last=x, sm=rm[1], zm=zm)
        return out

    def _vit25_backward(self, tape: dict, d_out, grads: GradStore):
        d, W, G = self.dims, self.params, grads
        E, nh, hd, Mp = d.v_embed, d.v_heads, d.v_head_dim, d.v_mlp_pad
        T = tape["pix"].shape[0]
        plan = tape["plan"]
        # out = merged[rev]  =>  d_merged = d_out[widx]
        d_m = torch.empty_like(d_out)
        ops.call("sb_gather_rows", d_out.contiguous(), plan["widx"], d_m, d_out.shape[0], d_out.shape[1])
        fm = torch.empty_like(tape["zm"])
        ops.call("sb_act_fwd", tape["zm"], fm, fm.numel(), 1)
        ops.gemm(d_m, fm, a_mn=True, b_mn=True, out=G["v.m.fc2_w"])
        ops.call("sb_colsum", d_m, G["v.m.fc2_b"], d_m.shape[0], d_m.shape[1], d_m.stride(0))
        d_fm = ops.gemm(d_m, W["v.m.fc2_w"], b_mn=True)
        d_zm = fm
        ops.call("sb_act_bwd", tape["zm"], d_fm, d_zm, d_zm.numel(), 1)
        hm = ops.rmsnorm_fwd(tape["x_last"], W["v.m.ln_w"], 1e-6).view(-1, d.merge_dim)
        ops.gemm(d_zm, hm, a_mn=True, b_mn=True, out=G["v.m.fc0_w"])
        ops.call("sb_colsum", d_zm, G["v.m.fc0_b"], d_zm.shape[0], d_zm.shape[1], d_zm.stride(0))
        d_hm = ops.gemm(d_zm, W["v.m.fc0_w"], b_mn=True).view(T, E)
        dx = ops.rmsnorm_bwd(tape["x_last"], W["v.m.ln_w"], tape["sm"], d_hm, G["v.m.ln_w"])
        del fm, d_fm, d_zm, hm, d_hm, d_m
        for i in reversed(range(d.v_depth)):
            p = f"v.{i}."
            t = tape["blocks"][i]
            meta = plan["meta_full"] if i in d.v_fullatt else plan["meta_win"]
            # MLP: x3 = x2 + down(swiglu(gu)) ; gu = raw [gate|up] incl. bias
            act = torch.empty((T, Mp), device=self.device, dtype=BF16)
            ops.call("sb_swiglu_bwd", t["gu"], None, None, act, T, Mp)
            ops.gemm(dx, act, a_mn=True, b_mn=True, out=G[p + "down_w"])
            ops.call("sb_colsum", dx, G[p + "down_b"], T, E, E)
            d_act = ops.gemm(dx, W[p + "down_w"], b_mn=True, out=act)
            d_gu = torch.empty_like(t["gu"])
            ops.call("sb_swiglu_bwd", t["gu"], d_act, d_gu, None, T, Mp)
            h2 = ops.rmsnorm_fwd(t["x2"], W[p + "ln2_w"], 1e-6)
            ops.gemm(d_gu, h2, a_mn=True, b_mn=True, out=G[p + "gu_w"])
            ops.call("sb_colsum", d_gu, G[p + "gu_b"], T, 2 * Mp, 2 * Mp)
            d_h2 = ops.gemm(d_gu, W[p + "gu_w"], b_mn=True, out=h2)
            dx2 = ops.rmsnorm_bwd(t["x2"], W[p + "ln2_w"], t["s2"], d_h2, G[p + "ln2_w"], dres=dx)
            del d_gu, act
            ops.gemm(dx2, t["a"], a_mn=True, b_mn=True, out=G[p + "proj_w"])
            ops.call("sb_colsum", dx2, G[p + "proj_b"], T, E, E)
            d_a = ops.gemm(dx2, W[p + "proj_w"], b_mn=True)
            qkv = t["qkv"]
            d_qkv = torch.empty_like(qkv)
            ops.attn_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], t["a"], t["lse"], d_a, meta, nh, nh, hd,
                         d_qkv[:, :E], d_qkv[:, E:2 * E], d_qkv[:, 2 * E:])
            ops.call("sb_rope_vit_pos", d_qkv, T, nh, hd, plan["pos_hw"], 1)
            h = ops.rmsnorm_fwd(t["x"], W[p + "ln1_w"], 1e-6)
            ops.gemm(d_qkv, h, a_mn=True, b_mn=True, out=G[p + "qkv_w"])
            ops.call("sb_colsum", d_qkv, G[p + "qkv_b"], T, 3 * E, 3 * E)
            d_h = ops.gemm(d_qkv, W[p + "qkv_w"], b_mn=True, out=h)
            dx = ops.rmsnorm_bwd(t["x"], W[p + "ln1_w"], t["s1"], d_h, G[p + "ln1_w"], dres=dx2)
            tape["blocks"][i] = None
        ops.gemm(dx, tape["pix"], a_mn=True, b_mn=True, out=G["v.patch_w"])

    def vit_backward(self, tape: dict, d_out, grads: GradStore):
        if self.dims.variant == "qwen2_5_vl":
            return self._vit25_backward(tape, d_out, grads)
        d, W, G = self.dims, self.params, grads
        E, nh, hd = d.v_embed, d.v_heads, d.v_head_dim
        T = tape["pix"].shape[0]
        mode_q, mode_g = 0, 1
        # merger
        fm = torch.empty_like(tape["zm"])
        ops.call("sb_act_fwd", tape["zm"], fm, fm.numel(), mode_g)
        ops.gemm(d_out, fm, a_mn=True, b_mn=True, out=G["v.m.fc2_w"])
        ops.call("sb_colsum", d_out, G["v.m.fc2_b"], d_out.shape[0], d_out.shape[1], d_out.stride(0))
        d_fm = ops.gemm(d_out, W["v.m.fc2_w"], b_mn=True)
        d_zm = fm  # reuse
        ops.call("sb_act_bwd", tape["zm"], d_fm, d_zm, d_zm.numel(), mode_g)
        hm = ops.layernorm_fwd(tape["x_last"], W["v.m.ln_w"], W["v.m.ln_b"]).view(-1, d.merge_dim)
        ops.gemm(d_zm, hm, a_mn=True, b_mn=True, out=G["v.m.fc0_w"])
        ops.call("sb_colsum", d_zm, G["v.m.fc0_b"], d_zm.shape[0], d_zm.shape[1], d_zm.stride(0))
        d_hm = ops.gemm(d_zm, W["v.m.fc0_w"], b_mn=True).view(T, E)
        dx = ops.layernorm_bwd(tape["x_last"], W["v.m.ln_w"], tape["mm"], tape["sm"], d_hm, G["v.m.ln_w"], G["v.m.ln_b"])
        del fm, d_fm, d_zm, hm, d_hm
        for i in reversed(range(d.v_depth)):
            p = f"v.{i}."
            t = tape["blocks"][i]
            f = torch.empty_like(t["z"])
            ops.call("sb_act_fwd", t["z"], f, f.numel(), mode_q)
            ops.gemm(dx, f, a_mn=True, b_mn=True, out=G[p + "fc2_w"])
            ops.call("sb_colsum", dx, G[p + "fc2_b"], T, E, E)
            d_f = ops.gemm(dx, W[p + "fc2_w"], b_mn=True)
            d_z = f
            ops.call("sb_act_bwd", t["z"], d_f, d_z, d_z.numel(), mode_q)
            h2 = ops.layernorm_fwd(t["x2"], W[p + "ln2_w"], W[p + "ln2_b"])
            ops.gemm(d_z, h2, a_mn=True, b_mn=True, out=G[p + "fc1_w"])
            ops.call("sb_colsum", d_z, G[p + "fc1_b"], T, d.v_mlp, d.v_mlp)
            d_h2 = ops.gemm(d_z, W[p + "fc1_w"], b_mn=True, out=h2)
            dx2 = ops.layernorm_bwd(t["x2"], W[p + "ln2_w"], t["m2"], t["s2"], d_h2, G[p + "ln2_w"], G[p + "ln2_b"], dres=dx)
            ops.gemm(dx2, t["a"], a_mn=True, b_mn=True, out=G[p + "proj_w"])
            ops.call("sb_colsum", dx2, G[p + "proj_b"], T, E, E)
            d_a = ops.gemm(dx2, W[p + "proj_w"], b_mn=True)
            qkv = t["qkv"]
            d_qkv = torch.empty_like(qkv)
            ops.attn_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], t["a"], t["lse"], d_a, tape["meta"], nh, nh, hd,
                         d_qkv[:, :E], d_qkv[:, E:2 * E], d_qkv[:, 2 * E:])
            ops.rope_vit(d_qkv, nh, hd, tape["grids"], d.merge, inverse=True)
            h = ops.layernorm_fwd(t["x"], W[p + "ln1_w"], W[p + "ln1_b"])
            ops.gemm(d_qkv, h, a_mn=True, b_mn=True, out=G[p + "qkv_w"])
            ops.call("sb_colsum", d_qkv, G[p + "qkv_b"], T, 3 * E, 3 * E)
            d_h = ops.gemm(d_qkv, W[p + "qkv_w"], b_mn=True, out=h)
            dx = ops.layernorm_bwd(t["x"], W[p + "ln1_w"], t["m1"], t["s1"], d_h, G[p + "ln1_w"], G[p + "ln1_b"], dres=dx2)
            tape["blocks"][i] = None
        ops.gemm(dx, tape["pix"], a_mn=True, b_mn=True, out=G["v.patch_w"])

    # ---- language model --------------------------------------------------------------------------
    def llm_forward(self, ids, vis, pos, meta, tape: dict | None = None, kv_out=None):
        """Qwen2VLTextModel.forward on a packed token sequence (MQ2:597-662, 905).
        ids int32 [T]; vis bf16 [N_v, H] or None; pos int32 [3, T]; meta int32 [T, 4].
        kv_out: optional list (per layer) of (k_cache, v_cache) bf16 [T, nkv*hd] filled with rotated K and V."""
        d, W = self.dims, self.params
        T, H = ids.shape[0], d.hidden
        nh, nkv, hd = d.heads, d.kv_heads, d.head_dim
        nq, nk = nh * hd, nkv * hd
        save = tape is not None
        vis_idx = None
        n_vis = 0
        if vis is not None:
            vis_idx = torch.empty(T, dtype=I32, device=self.device)
            cnt = torch.zeros(1, dtype=I32, device=self.device)
            ops.call("sb_vision_index", ids, vis_idx, T, d.video_token_id, d.image_token_id, cnt)
            n_vis = vis.shape[0]
        x = torch.empty((T, H), device=self.device, dtype=BF16)
        ops.call("sb_embed_merge", ids, vis_idx, W["embed"], vis, x, T, H, n_vis)
        if save:
            tape.update(ids=ids, vis_idx=vis_idx, n_vis=n_vis, pos=pos, meta=meta, layers=[])
        h = torch.empty_like(x)
        for i in range(d.layers):
            p = f"l.{i}."
            r1 = ops.rmsnorm_fwd(x, W[p + "ln1_w"], d.rms_eps, save_stats=save, out=h)
            qkv = ops.gemm(h, W[p + "qkv_w"], bias=W[p + "qkv_b"])
            ko, vo = kv_out[i] if kv_out is not None else (None, None)
            ops.mrope(qkv, pos, nh, nkv, hd, d.rope_theta, d.mrope_section, k_out=ko, v_out=vo, kv_ld=nk)
            r2 = ops.attn_fwd(qkv[:, :nq], qkv[:, nq:nq + nk], qkv[:, nq + nk:], meta, nh, nkv, hd, save_lse=save)
            a = r2[0] if save else r2
            x2 = ops.gemm(a, W[p + "o_w"], residual=x)
            r3 = ops.rmsnorm_fwd(x2, W[p + "ln2_w"], d.rms_eps, save_stats=save, out=h)
            gu = torch.empty((T, 2 * d.inter), device=self.device, dtype=BF16) if save and not self.recompute_mlp else None
            act = ops.gemm(h, W[p + "gu_w"], epilogue=EPI_SWIGLU, aux=gu)
            x3 = ops.gemm(act, W[p + "down_w"], residual=x2)
            if save:
                tape["layers"].append(dict(x=x, s1=r1[1], qkv=qkv, a=a, lse=r2[1], x2=x2, s2=r3[1], gu=gu))
            x = x3
            del act
        rf = ops.rmsnorm_fwd(x, W["norm_w"], d.rms_eps, save_stats=save)
        if save:
            tape.update(x_last=x, sf=rf[1])
            return rf[0]
        return rf

    def llm_forward_from_prefix(self, ids, vis, pos, meta, ptape: dict, tape: dict):
        """llm_forward(ids, vis, pos, meta, tape) when the first P rows were already run with the same weights and saved
        in `ptape` (the rollout's prompt prefill): only the remaining rows go through the GEMMs; their attention reads the
        prefix keys/values from the saved projections.  Every per-row result is independent of how many rows a launch
        covers, so the returned hidden states and the assembled `tape` are bit-identical to the full forward."""
        d, W = self.dims, self.params
        T, H = ids.shape[0], d.hidden
        P = ptape["ids"].shape[0]
        Tc = T - P
        nh, nkv, hd = d.heads, d.kv_heads, d.head_dim
        nq, nk = nh * hd, nkv * hd
        vis_idx, n_vis = None, 0
        if vis is not None:
            vis_idx = torch.empty(T, dtype=I32, device=self.device)
            cnt = torch.zeros(1, dtype=I32, device=self.device)
            ops.call("sb_vision_index", ids, vis_idx, T, d.video_token_id, d.image_token_id, cnt)
            n_vis = vis.shape[0]
        pl = ptape["layers"]
        x = torch.empty((T, H), device=self.device, dtype=BF16)
        x[:P].copy_(pl[0]["x"])
        ops.call("sb_embed_merge", ids[P:], None if vis_idx is None else vis_idx[P:], W["embed"], vis, x[P:], Tc, H, n_vis)
        tape.update(ids=ids, vis_idx=vis_idx, n_vis=n_vis, pos=pos, meta=meta, layers=[])
        pos_c = pos[:, P:].contiguous()
        meta_c = meta[P:]
        h = torch.empty((Tc, H), device=self.device, dtype=BF16)
        for i in range(d.layers):
            p, pt = f"l.{i}.", pl[i]
            r1 = ops.rmsnorm_fwd(x[P:], W[p + "ln1_w"], d.rms_eps, save_stats=True, out=h)
            qkv = torch.empty((T, d.qkv_dim), device=self.device, dtype=BF16)
            qkv[:P].copy_(pt["qkv"])
            ops.gemm(h, W[p + "qkv_w"], bias=W[p + "qkv_b"], out=qkv[P:])
            ops.mrope(qkv[P:], pos_c, nh, nkv, hd, d.rope_theta, d.mrope_section)
            a = torch.empty((T, nq), device=self.device, dtype=BF16)
            a[:P].copy_(pt["a"])
            _, lse_c = ops.attn_fwd(qkv[P:, :nq], qkv[:, nq:nq + nk], qkv[:, nq + nk:], meta_c, nh, nkv, hd, out=a[P:],
                                    save_lse=True, Tk=T)
            x2 = torch.empty((T, H), device=self.device, dtype=BF16)
            x2[:P].copy_(pt["x2"])
            ops.gemm(a[P:], W[p + "o_w"], residual=x[P:], out=x2[P:])
            r3 = ops.rmsnorm_fwd(x2[P:], W[p + "ln2_w"], d.rms_eps, save_stats=True, out=h)
            gu = None
            if not self.recompute_mlp and pt["gu"] is not None:
                gu = torch.empty((T, 2 * d.inter), device=self.device, dtype=BF16)
                gu[:P].copy_(pt["gu"])
            act = ops.gemm(h, W[p + "gu_w"], epilogue=EPI_SWIGLU, aux=None if gu is None else gu[P:])
            x3 = torch.empty((T, H), device=self.device, dtype=BF16)
            x3[:P].copy_(pl[i + 1]["x"] if i + 1 < d.layers else ptape["x_last"])
            ops.gemm(act, W[p + "down_w"], residual=x2[P:], out=x3[P:])
            tape["layers"].append(dict(x=x, s1=torch.cat([pt["s1"], r1[1]]), qkv=qkv, a=a,
                                       lse=torch.cat([pt["lse"], lse_c], dim=1), x2=x2, s2=torch.cat([pt["s2"], r3[1]]),
                                       gu=gu))
            pl[i] = None          # the prefix copy of this layer is no longer needed
            x = x3
            del act
        rf = ops.rmsnorm_fwd(x, W["norm_w"], d.rms_eps, save_stats=True)
        tape.update(x_last=x, sf=rf[1])
        return rf[0]

    def llm_backward(self, tape: dict, d_hf, grads: GradStore, want_d_vis=True):
        d, W, G = self.dims, self.params, grads
        T, H, I = tape["ids"].shape[0], d.hidden, d.inter
        nh, nkv, hd = d.heads, d.kv_heads, d.head_dim
        nq, nk = nh * hd, nkv * hd
        dx = ops.rmsnorm_bwd(tape["x_last"], W["norm_w"], tape["sf"], d_hf, G["norm_w"])
        d_act = torch.empty((T, I), device=self.device, dtype=BF16)
        d_gu = torch.empty((T, 2 * I), device=self.device, dtype=BF16)
        act = torch.empty((T, I), device=self.device, dtype=BF16)
        h = torch.empty((T, H), device=self.device, dtype=BF16)
        d_h = torch.empty((T, H), device=self.device, dtype=BF16)
        d_a = torch.empty((T, nq), device=self.device, dtype=BF16)
        d_qkv = torch.empty((T, d.qkv_dim), device=self.device, dtype=BF16)
        delta = torch.empty((nh, T), device=self.device, dtype=F32)
        gu_buf = None
        for i in reversed(range(d.layers)):
            p = f"l.{i}."
            t = tape["layers"][i]
            ops.gemm(dx, W[p + "down_w"], b_mn=True, out=d_act)
            gu = t["gu"]
            ops.rmsnorm_fwd(t["x2"], W[p + "ln2_w"], d.rms_eps, out=h)
            if gu is None:        # recompute_mlp: the raw gate|up projection again, bit-identical to the forward's
                if gu_buf is None:
                    gu_buf = torch.empty((T, 2 * I), device=self.device, dtype=BF16)
                gu = ops.gemm(h, W[p + "gu_w"], out=gu_buf)
            ops.call("sb_swiglu_bwd", gu, d_act, d_gu, act, T, I)
            ops.gemm(dx, act, a_mn=True, b_mn=True, out=G[p + "down_w"])
            ops.gemm(d_gu, h, a_mn=True, b_mn=True, out=G[p + "gu_w"])
            ops.gemm(d_gu, W[p + "gu_w"], b_mn=True, out=d_h)
            dx2 = ops.rmsnorm_bwd(t["x2"], W[p + "ln2_w"], t["s2"], d_h, G[p + "ln2_w"], dres=dx)
            ops.gemm(dx2, W[p + "o_w"], b_mn=True, out=d_a)
            ops.gemm(dx2, t["a"], a_mn=True, b_mn=True, out=G[p + "o_w"])
            qkv = t["qkv"]
            ops.attn_bwd(qkv[:, :nq], qkv[:, nq:nq + nk], qkv[:, nq + nk:], t["a"], t["lse"], d_a, tape["meta"], nh, nkv,
                         hd, d_qkv[:, :nq], d_qkv[:, nq:nq + nk], d_qkv[:, nq + nk:], delta=delta)
            ops.mrope(d_qkv, tape["pos"], nh, nkv, hd, d.rope_theta, d.mrope_section, inverse=True)
            ops.call("sb_colsum", d_qkv, G[p + "qkv_b"], T, d.qkv_dim, d.qkv_dim)
            ops.rmsnorm_fwd(t["x"], W[p + "ln1_w"], d.rms_eps, out=h)
            ops.gemm(d_qkv, h, a_mn=True, b_mn=True, out=G[p + "qkv_w"])
            ops.gemm(d_qkv, W[p + "qkv_w"], b_mn=True, out=d_h)
            dx = ops.rmsnorm_bwd(t["x"], W[p + "ln1_w"], t["s1"], d_h, G[p + "ln1_w"], dres=dx2)
            tape["layers"][i] = None
            del dx2
            G.ready(p)
        d_vis = None
        if tape["n_vis"] > 0 and want_d_vis:
            d_vis = torch.zeros((tape["n_vis"], H), device=self.device, dtype=BF16)
        plan = tape.get("embed_plan")
        if plan is None:      # atomic fallback (order of duplicate tokens not fixed)
            ops.call("sb_embed_bwd", tape["ids"], tape["vis_idx"], dx, G["embed"], d_vis, T, H, tape["n_vis"])
        else:
            order_pos, off, dst, n_seg, vis_pos = plan
            if n_seg > 0:
                ops.call("sb_segment_sum_rows", dx, order_pos, off, dst, n_seg, G["embed"], H, 1)
            if d_vis is not None and vis_pos.numel() > 0:
                ops.call("sb_gather_rows", dx, vis_pos, d_vis, vis_pos.numel(), H)
        G.ready("embed")
        return d_vis

    # ---- scoring -----------------------------------------------------------------------------------
    def _lmhead_partials(self, hsel, targets):
        d = self.dims
        R = hsel.shape[0]
        nt = (d.vocab + 255) // 256
        part = torch.empty((R, nt, 2), device=self.device, dtype=F32)
        tl = torch.zeros(R, device=self.device, dtype=F32)
        ops.gemm(hsel, self.params["lm_head"], epilogue=EPI_LMHEAD, targets=targets, lse_part=part, tgt_logit=tl)
        return part, tl, nt

    @torch.no_grad()
    def per_token_logps(self, batch: PackedBatch, pixel_values, grid_thw):
        """[G, C] log-probs of the completion tokens (the slice `[:, P-1:]` of TRN:353-366, 526-528)."""
        vis = self.vit_forward(pixel_values, grid_thw) if pixel_values is not None else None
        hf = self.llm_forward(batch.ids, vis, batch.pos, batch.meta)
        R = batch.rows.shape[0]
        hsel = torch.empty((R, self.dims.hidden), device=self.device, dtype=BF16)
        ops.call("sb_gather_rows", hf, batch.rows, hsel, R, self.dims.hidden)
        part, tl, nt = self._lmhead_partials(hsel, batch.targets)
        lp = torch.empty(R, device=self.device, dtype=F32)
        ops.call("sb_logprob_from_partials", part, nt, tl, lp, R)
        return lp.view(batch.G, batch.C)

    @torch.no_grad()
    def get_per_token_logps(self, input_ids, pixel_values_videos=None, video_grid_thw=None, pixel_values=None,
                            image_grid_thw=None, **unused):
        """Drop-in for `SGRLVRTrainer._get_per_token_logps(model, input_ids, **kwargs)` (SG_RLVR_trainer.py:353-366) on the
        no-grad path (the reference-policy scoring, TRN:534-547): takes the reference's arguments -- `input_ids [B, L]`
        whose rows share the prompt, the xB-repeated pixel values and grid (TRN:507-521) -- and returns `[B, L-1]`
        log-probs of tokens 1..L-1.  The rows are packed behind their longest common prefix (exact for any common
        prefix); the prefix positions are scored once and broadcast."""
        d = self.dims
        ids = input_ids.cpu().long()
        if ids.dim() != 2:
            raise SpacerError("get_per_token_logps: input_ids must be [B, L]")
        B, L = ids.shape
        if pixel_values is not None:
            pixel_values_videos, video_grid_thw = pixel_values, image_grid_thw
        grid = video_grid_thw
        pix = pixel_values_videos
        if grid is not None:
            grid = torch.as_tensor(grid).reshape(-1, 3)
            if grid.shape[0] == B and B > 1:       # the reference repeats the visual inputs once per row
                if not bool((grid == grid[:1]).all()):
                    raise SpacerError("get_per_token_logps: rows must share one visual input")
                n_p = int(grid[0, 0] * grid[0, 1] * grid[0, 2])
                pix, grid = pix[:n_p], grid[:1]
        same = (ids == ids[:1]).all(0)
        P = int(L if bool(same.all()) else torch.nonzero(~same)[0, 0])
        P = max(1, min(P, L - 1))                  # at least one "completion" column, at least one prefix token
        batch = pack_prompt_completions(ids[0, :P], ids[:, P:], grid, d, self.device, self.rope_convention)
        vis = self.vit_forward(pix.to(self.device), grid) if pix is not None else None
        hf = self.llm_forward(batch.ids, vis, batch.pos, batch.meta)
        # rows that predict the completion tokens, then rows 0..P-2 that predict prefix tokens 1..P-1
        rows = torch.cat([batch.rows, torch.arange(P - 1, device=self.device, dtype=I32)])
        targets = torch.cat([batch.targets, ids[0, 1:P].to(I32).to(self.device)])
        R = rows.numel()
        hsel = torch.empty((R, d.hidden), device=self.device, dtype=BF16)
        ops.call("sb_gather_rows", hf, rows, hsel, R, d.hidden)
        part, tl, nt = self._lmhead_partials(hsel, targets)
        lp = torch.empty(R, device=self.device, dtype=F32)
        ops.call("sb_logprob_from_partials", part, nt, tl, lp, R)
        n_c = B * (L - P)
        return torch.cat([lp[n_c:].view(1, P - 1).expand(B, -1), lp[:n_c].view(B, L - P)], dim=1)

    # ---- training forward / backward, in reusable pieces ---------------------------------------------
    def forward_hidden(self, batch: PackedBatch, pixel_values, grid_thw, vit_cache: dict | None = None, mark=None):
        """Vision tower + packed LLM forward with the activations saved for the backward.  Returns (hf [T, H], state);
        `state` goes to backward_hidden.  `vit_cache` (from generate(keep_vit_tape=True)) supplies the rollout's ViT
        forward and prompt prefill when pixels, prompt ids and position ids are the ones the rollout saw."""
        mark = mark or (lambda name: None)
        vtape, ltape = {}, {}
        if (vit_cache is not None and pixel_values is not None and vit_cache["pixels"] is pixel_values
                and vit_cache.get("tape") is not None):
            vis, vtape = vit_cache["vis"], vit_cache["tape"]     # forward already done by the rollout (generate)
            vit_cache["tape"] = None                             # single use: the backward frees it block by block
        else:
            vis = self.vit_forward(pixel_values, grid_thw, vtape) if pixel_values is not None else None
        mark("vit_fwd")
        ptape = None
        if vit_cache is not None and vit_cache.get("llm_tape") is not None and batch.prompt_ids_host is not None:
            # the rollout's prefill ran these prompt rows with the same weights: reuse them if prompt and positions agree
            # (they differ e.g. when Qwen2.5-VL's rollout saw second_per_grid_ts and the scoring does not, TRN:519-520)
            if (vit_cache["pixels"] is pixel_values and torch.equal(vit_cache["prompt_ids"], batch.prompt_ids_host)
                    and torch.equal(vit_cache["prompt_pos"], batch.prompt_pos_host)):
                ptape = vit_cache["llm_tape"]
        if ptape is not None:
            hf = self.llm_forward_from_prefix(batch.ids, vis, batch.pos, batch.meta, ptape, ltape)
            vit_cache["llm_tape"] = None
        else:
            hf = self.llm_forward(batch.ids, vis, batch.pos, batch.meta, ltape)
        mark("llm_fwd")
        ltape["embed_plan"] = batch.embed_plan
        return hf, dict(vis=vis, vtape=vtape, ltape=ltape)

    def lm_head_backward(self, hsel, targets, lse, coef, grads: GradStore, lm_chunk: int = 4096):
        """Backward of `logprob = logit[target] - logsumexp(logits)` through lm_head for the selected hidden rows, given
        coef = dLoss/dlogprob per row: logits are recomputed tile by tile in the DLOGITS epilogue
        (coef * (onehot - softmax)), never stored as fp32 [rows, V].  Accumulates dW into grads["lm_head"] (which is the
        embedding table when tied: call grads.zero_for_step() first) and returns d_hsel [rows, H]."""
        d = self.dims
        R = hsel.shape[0]
        d_hsel = torch.empty((R, d.hidden), device=self.device, dtype=BF16)
        g_lm = grads["lm_head"]
        first = True
        for r0 in range(0, R, lm_chunk):
            r1 = min(r0 + lm_chunk, R)
            dl = ops.gemm(hsel[r0:r1], self.params["lm_head"], epilogue=EPI_DLOGITS, targets=targets[r0:r1],
                          lse=lse[r0:r1], coef=coef[r0:r1])
            ops.gemm(dl, self.params["lm_head"], b_mn=True, out=d_hsel[r0:r1])
            ops.gemm(dl, hsel[r0:r1], a_mn=True, b_mn=True, out=g_lm, residual=None if (first and not d.tie) else g_lm)
            first = False
            del dl
        if not d.tie:
            grads.ready("lm_head")
        return d_hsel

    def scatter_rows(self, d_hsel, rows, rows_plan, T):
        """d_hf [T, H] with d_hsel's rows added at `rows` (deterministic when a segment plan is given: repeated rows --
        the last prompt row predicts the first token of every completion -- are summed in fp32 in a fixed order)."""
        H = self.dims.hidden
        d_hf = torch.zeros((T, H), device=self.device, dtype=BF16)
        if rows_plan is not None:
            order, off, dst, n_seg = rows_plan
            ops.call("sb_segment_sum_rows", d_hsel, order, off, dst, n_seg, d_hf, H, 0)
        else:
            ops.call("sb_scatter_add_rows", d_hsel, rows, d_hf, d_hsel.shape[0], H)
        return d_hf

    def backward_hidden(self, state: dict, d_hf, grads: GradStore, mark=None):
        """Backward of forward_hidden from d(final hidden states): decoder layers, embedding, vision tower."""
        mark = mark or (lambda name: None)
        d_vis = self.llm_backward(state["ltape"], d_hf, grads)
        mark("llm_bwd")
        state["ltape"] = None
        if state["vis"] is not None:
            self.vit_backward(state["vtape"], d_vis, grads)
        else:
            grads.zero_range("v.")
        state["vtape"] = None
        mark("vit_bwd")
        grads.ready("v.")         # the whole vision tower as one bucket (1.3 GB of bf16 at 7B)

    def grpo_forward_backward(self, batch: PackedBatch, pixel_values, grid_thw, ref_logps, advantages, beta,
                              grads: GradStore, lm_chunk: int = 4096, vit_cache: dict | None = None):
        """One GRPO forward/backward (TRN:526-528, 551-552, 640-643 + autograd's backward).
        Returns dict(loss, mean_kl, logps [G,C], mask [G,C], lengths [G]); gradients land in `grads`."""
        d = self.dims
        H = d.hidden
        G_, C = batch.G, batch.C
        R = G_ * C
        marks = self.phase_marks      # optional list: (name, CUDA event) per sub-phase (tools/profile_phases.py)

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        hf, state = self.forward_hidden(batch, pixel_values, grid_thw, vit_cache, mark)
        hsel = torch.empty((R, H), device=self.device, dtype=BF16)
        ops.call("sb_gather_rows", hf, batch.rows, hsel, R, H)
        part, tl, nt = self._lmhead_partials(hsel, batch.targets)
        lp = torch.empty(R, device=self.device, dtype=F32)
        lse = torch.empty(R, device=self.device, dtype=F32)
        coef = torch.empty(R, device=self.device, dtype=F32)
        mask = torch.empty(R, device=self.device, dtype=I32)
        row_loss = torch.empty(G_, device=self.device, dtype=F32)
        row_kl = torch.empty(G_, device=self.device, dtype=F32)
        row_len = torch.empty(G_, device=self.device, dtype=I32)
        out2 = torch.empty(2, device=self.device, dtype=F32)
        adv = advantages.to(self.device, F32).contiguous()
        ref = None if ref_logps is None else ref_logps.to(self.device, F32).contiguous()
        ops.call("sb_grpo_loss", part, nt, tl, batch.comp_ids, G_, C, d.eos_id, ref, adv, float(beta), lp, lse, coef,
                 mask, row_loss, row_kl, row_len, out2, ops.grpo_loss_workspace(G_, C, self.device))
        del part
        mark("lm_head_loss")
        # backward through lm_head: recompute logits tile by tile, emit dlogits, two GEMMs
        grads.zero_for_step()
        d_hsel = self.lm_head_backward(hsel, batch.targets, lse, coef, grads, lm_chunk)
        T = hf.shape[0]
        del hsel, hf
        d_hf = self.scatter_rows(d_hsel, batch.rows, batch.rows_plan, T)
        del d_hsel
        mark("lm_head_bwd")
        self.backward_hidden(state, d_hf, grads, mark)
        return dict(loss=out2[0], mean_kl=out2[1], logps=lp.view(G_, C), mask=mask.view(G_, C), lengths=row_len)

    # ---- supervised fine-tuning ---------------------------------------------
    def sft_forward_backward(self, input_ids, labels, pixel_values, grid_thw, grads: GradStore, lm_chunk: int = 4096,
                             second_per_grid_ts=None):
        """Token cross-entropy of ONE causal sequence + backward, on the same kernels as the GRPO step: what trl's
        SFTTrainer computes for the reference's SFT stage (open_r1/sft.py:147-182: labels = input_ids with pad and
        visual tokens set to -100; HF causal-LM loss: predict token t+1 from position t, mean over labels != -100).
        input_ids / labels: [L] (or [1, L]).  Returns dict(loss, n_tokens); gradients land in `grads`."""
        d = self.dims
        H = d.hidden
        ids = input_ids.reshape(-1).cpu().long()
        lab = labels.reshape(-1).cpu().long()
        L = ids.numel()
        if lab.numel() != L:
            raise SpacerError("sft_forward_backward: labels and input_ids differ in length")
        keep = torch.nonzero(lab[1:] != -100, as_tuple=True)[0]          # position t predicts label t + 1
        R = int(keep.numel())
        if R == 0:
            raise SpacerError("sft_forward_backward: every label is masked")
        rows = keep.to(I32).to(self.device)
        targets = lab[1:][keep].to(I32).to(self.device)
        pos, _ = rope_index(ids, grid_thw, d, self.rope_convention, second_per_grid_ts)
        ids_dev = ids.to(I32).to(self.device)
        pos_dev = pos.to(I32).contiguous().to(self.device)
        meta = causal_meta(L, self.device)
        vtape, ltape = {}, {}
        vis = self.vit_forward(pixel_values, grid_thw, vtape) if pixel_values is not None else None
        hf = self.llm_forward(ids_dev, vis, pos_dev, meta, ltape)
        hsel = torch.empty((R, H), device=self.device, dtype=BF16)
        ops.call("sb_gather_rows", hf, rows, hsel, R, H)
        part, tl, nt = self._lmhead_partials(hsel, targets)
        lp = torch.empty(R, device=self.device, dtype=F32)
        ops.call("sb_logprob_from_partials", part, nt, tl, lp, R)
        del part
        loss = -lp.mean()
        lse = tl - lp                                                     # log-sum-exp of the bf16-rounded logits
        coef = torch.full((R,), -1.0 / R, device=self.device, dtype=F32)  # dLoss / dlogprob
        grads.zero_for_step()
        d_hsel = torch.empty((R, H), device=self.device, dtype=BF16)
        g_lm = grads["lm_head"]
        first = True
        for r0 in range(0, R, lm_chunk):
            r1 = min(r0 + lm_chunk, R)
            dl = ops.gemm(hsel[r0:r1], self.params["lm_head"], epilogue=EPI_DLOGITS, targets=targets[r0:r1],
                          lse=lse[r0:r1], coef=coef[r0:r1])
            ops.gemm(dl, self.params["lm_head"], b_mn=True, out=d_hsel[r0:r1])
            ops.gemm(dl, hsel[r0:r1], a_mn=True, b_mn=True, out=g_lm, residual=None if (first and not d.tie) else g_lm)
            first = False
            del dl
        if not d.tie:
            grads.ready("lm_head")
        d_hf = torch.zeros_like(hf)
        ops.call("sb_scatter_add_rows", d_hsel, rows, d_hf, R, H)      # rows are distinct here: no accumulation order
        ltape["embed_plan"] = embed_plan(ids, d, self.device)
        del hsel, d_hsel, hf
        d_vis = self.llm_backward(ltape, d_hf, grads)
        if vis is not None:
            self.vit_backward(vtape, d_vis, grads)
        else:
            grads.zero_range("v.")
        grads.ready("v.")
        return dict(loss=loss, n_tokens=R, logps=lp)

    # ---- rollout -----------------------------------------------------------------------------------
    def decode_weight_bytes(self) -> int:
        """bf16 bytes of the weights one decode step must stream: 28 x (qkv, o, gate|up, down) + lm_head."""
        d = self.dims
        per_layer = d.qkv_dim * d.hidden + d.hidden * d.heads * d.head_dim + 3 * d.inter * d.hidden
        return 2 * (d.layers * per_layer + d.vocab * d.hidden)

    @torch.no_grad()
    def profile_decode_gemv(self, rows: int, reps: int = 3):
        """Time the weight-streaming GEMVs of one decode step (4 per layer + lm_head, the kernels that move >97 % of
        a step's bytes) back to back with CUDA events; returns achieved GB/s on their algorithmic bytes."""
        d, W = self.dims, self.params
        st = self._alloc_decode(rows, 8, 8, 1)      # private scratch state, not the cached rollout state
        S = st["S"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def sweep():
            for i in range(d.layers):
                p = f"l.{i}."
                self._gemv(W[p + "qkv_w"], st["xn"], st["p_qkv"], S["qkv"])
                self._gemv(W[p + "o_w"], st["attn"], st["p_o"], S["o"])
                self._gemv(W[p + "gu_w"], st["xn"], st["act"], 1, swiglu=True)
                self._gemv(W[p + "down_w"], st["act"], st["p_down"], S["down"])
            self._gemv(W["lm_head"], st["xn"], st["logits"], 1)

        sweep()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            sweep()
        e1.record()
        torch.cuda.synchronize()
        launches = 4 * d.layers + 1
        ms = e0.elapsed_time(e1) / reps
        byts = self.decode_weight_bytes()
        return dict(gbs=byts / (ms * 1e-3) / 1e9, bytes_per_launch=byts / launches, avg_us=ms * 1e3 / launches,
                    launches=launches, ms_per_sweep=ms)

    # L2 plan of the decode step (profiles/r02_l2_hint_lab.md).  Every weight stream of the decode GEMVs is tagged
    # evict_first (a weight is read once per step).  During the attention window of a layer -- the longest stretch in which
    # HBM would idle -- a small kernel on a side stream (a parallel branch of the step's CUDA graph, so that nothing on the
    # layer's dependency chain waits for it) asks L2 for the first PF_GU_ROWS rows of EVERY 128-row tile of gate|up, tagged
    # evict_last so that they survive the ~70 MB that pass through L2 before the gate|up GEMV reads them.  A row subset of
    # every tile (instead of the head of the matrix) makes every TMA box of that GEMV part L2 hit, part HBM: all CTAs gain
    # alike and HBM stays saturated beside the hits.  PF_AFTER names the kernel whose completion releases the requests.
    # The requests go through the load/store path (one prefetch.global.L2::evict_last per line), paced so that they are
    # spread over the window at ~3 TB/s: a burst -- or the bulk/TMA form -- saturates the memory system, and the
    # latency-bound attention / norm kernels then wait behind it (tools/decode_lab.py --exp step: 2.96 ms per step without,
    # 2.87 with; 3.1-3.3 when the requests come as a burst or overrun into the gate|up GEMV).
    PF_GU_ROWS = 24
    PF_AFTER = "qkv"           # "qkv" | "qkv_post" | "combine"
    PF_CTAS = 0                # CTAs that issue the requests (0 = one per SM)
    PF_PACE_NS = 750           # >= 0: per-line requests, each of the CTAs' 128 threads pausing this long; -1: bulk (TMA) form

    def _gemv(self, w, x16, out_parts, splits, swiglu=False):
        """parts[s][r][n] = x16[r] . w[n] over K split s   (swap-AB tcgen05 GEMM, weights streamed once).
        swiglu: w is the interleaved gate|up matrix and out_parts the bf16 activation [rows, I] (fused epilogue)."""
        if swiglu:
            ops.gemm(w, x16, out=out_parts, epilogue=ops.EPI_F32T_SWIGLU)
        else:
            ops.gemm(w, x16, out=out_parts, epilogue=EPI_F32T, k_splits=splits)

    def _prefetch_tile_rows(self, st, w, rows):
        """Fork: the side stream waits for what the main stream has enqueued so far, then requests the first `rows` rows of
        every 128-row tile of `w`.  Joined once, at the end of the step."""
        n_rows, k = w.shape
        chunk = min(int(rows), 128) * k * 2
        if n_rows < 128 or chunk % 128:          # nothing sensible to request for toy shapes
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side = st["side"]
        side.wait_event(ev)
        with torch.cuda.stream(side):
            ops.call("sb_dec_l2_prefetch", w, chunk, 128 * k * 2, n_rows // 128, int(self.PF_CTAS), int(self.PF_PACE_NS))
        st["forked"] = True

    def _splits_for(self, n_out, k):
        m_tiles = (n_out + 127) // 128
        if m_tiles >= 148:
            return 1
        s = max(1, 148 // m_tiles)
        return min(s, max(1, (k + 63) // 64))

    def _alloc_decode(self, R, P, c_max, n_prompts):
        d, dev = self.dims, self.device
        H, I, nk = d.hidden, d.inter, d.kv_heads * d.head_dim
        RP = 16 if R <= 16 else 32
        lib = ops._lib.load()
        S = {}
        for name, (n_out, k) in dict(qkv=(d.qkv_dim, H), o=(H, d.heads * d.head_dim), gu=(2 * I, H), down=(H, I),
                                     lm=(d.vocab, H)).items():
            S[name] = 1 if name == "lm" else lib.sb_gemm_effective_splits(k, self._splits_for(n_out, k))
        st = dict(
            R=R, RP=RP, P=P, c_max=c_max, S=S,
            x=torch.zeros((RP, H), device=dev, dtype=BF16), xn=torch.zeros((RP, H), device=dev, dtype=BF16),
            q=torch.zeros((RP, d.heads * d.head_dim), device=dev, dtype=BF16),
            attn=torch.zeros((RP, d.heads * d.head_dim), device=dev, dtype=BF16),
            act=torch.zeros((RP, I), device=dev, dtype=BF16),
            p_qkv=torch.empty((S["qkv"], RP, d.qkv_dim), device=dev, dtype=F32),
            p_o=torch.empty((S["o"], RP, H), device=dev, dtype=F32),
            p_gu=torch.empty((S["gu"], RP, 2 * I), device=dev, dtype=F32),
            p_down=torch.empty((S["down"], RP, H), device=dev, dtype=F32),
            logits=torch.empty((S["lm"], RP, d.vocab), device=dev, dtype=F32),
            kc=torch.zeros((d.layers, R, c_max, nk), device=dev, dtype=BF16),
            vc=torch.zeros((d.layers, R, c_max, nk), device=dev, dtype=BF16),
            kp=[torch.empty((d.layers, P, nk), device=dev, dtype=BF16) for _ in range(n_prompts)],
            vp=[torch.empty((d.layers, P, nk), device=dev, dtype=BF16) for _ in range(n_prompts)],
            step=torch.zeros(1, device=dev, dtype=I32), tokens=torch.zeros(RP, device=dev, dtype=I32),
            finished=torch.zeros(RP, device=dev, dtype=I32),
            out_ids=torch.zeros((R, c_max), device=dev, dtype=I32),
            seed=torch.zeros(1, device=dev, dtype=torch.int64),
            seen=torch.zeros((RP, (d.vocab + 31) // 32), device=dev, dtype=I32),   # token bitmap (repetition penalty)
            graphs={}, side=torch.cuda.Stream(device=dev), forked=False,
        )
        st["attn_ws"] = None      # sized on first use (depends on the group split)
        return st

    def _attn_workspace(self, st, rows_group0):
        import ctypes
        d = self.dims
        n = ctypes.c_longlong(0)
        ops._lib.check(ops._lib.load().sb_dec_attn_workspace(st["R"], rows_group0, st["P"], st["c_max"], d.heads,
                                                             d.kv_heads, ctypes.byref(n)), "sb_dec_attn_workspace")
        if st["attn_ws"] is None or st["attn_ws"].numel() < n.value:
            st["attn_ws"] = torch.empty(n.value, device=self.device, dtype=F32)
        return st["attn_ws"]

    def _decode_state(self, R, P, c_max, n_prompts):
        """Decode buffers (KV caches, split-K partials, sampler state) are allocated once per shape and reused by every
        rollout, so the CUDA graph captured over them stays valid from step to step."""
        key = (R, P, c_max, n_prompts)
        if self._dec is None or self._dec[0] != key:
            self._dec = None                  # free the old buffers before allocating the new ones
            self._dec = (key, self._alloc_decode(R, P, c_max, n_prompts))
        st = self._dec[1]
        st["step"].zero_(); st["finished"].zero_(); st["tokens"].zero_(); st["out_ids"].zero_()
        return st

    def _sample(self, st, sp, suppress_eos):
        """Next token per row from st["logits"] under the sampling parameters `sp` (SamplingParams): HF's processor chain
        repetition penalty -> temperature -> top-k -> top-p -> multinomial (TRN:277-302 generation configs), or greedy
        argmax (evaluation, SpaceR-Eval/data_utils/vsibench.py:174)."""
        ops.sample(st["logits"][0], st["step"], st["tokens"], V=self.dims.vocab, R=st["R"], greedy=sp.greedy, top_p=sp.top_p,
                   top_k=sp.top_k, temperature=sp.temperature, repetition_penalty=sp.repetition_penalty,
                   seen=st["seen"] if sp.repetition_penalty != 1.0 else None, seed_dev=st["seed"],
                   finished=st["finished"], out_ids=st["out_ids"], eos_ids=sp.eos_ids, pad_id=sp.pad_id,
                   suppress_eos=suppress_eos)

    def _decode_step(self, st, rope_base, rows_group0, samp, suppress_eos):
        """Enqueue one decode step (feeds tokens at slot *step, samples the next token into slot *step + 1)."""
        d, W = self.dims, self.params
        R, RP, P, S = st["R"], st["RP"], st["P"], st["S"]
        H, I = d.hidden, d.inter
        nh, nkv, hd = d.heads, d.kv_heads, d.head_dim
        ws = self._attn_workspace(st, rows_group0)
        ops.call("sb_dec_embed", st["tokens"], W["embed"], st["x"], R, H)
        parts, sp = None, 0
        pf_rows = int(self.PF_GU_ROWS)
        st["forked"] = False
        for i in range(d.layers):
            p = f"l.{i}."
            ops.call("sb_dec_residual_rmsnorm", st["x"], parts, sp, RP * H, H, W[p + "ln1_w"], st["xn"], R, H, d.rms_eps)
            self._gemv(W[p + "qkv_w"], st["xn"], st["p_qkv"], S["qkv"])
            if pf_rows and self.PF_AFTER == "qkv":
                self._prefetch_tile_rows(st, W[p + "gu_w"], pf_rows)
            ops.call("sb_dec_qkv_post", st["p_qkv"], S["qkv"], RP * d.qkv_dim, d.qkv_dim, W[p + "qkv_b"], st["step"],
                     rope_base, float(d.rope_theta), nh, nkv, hd, st["q"], st["kc"][i], st["vc"][i],
                     st["c_max"] * nkv * hd, st["c_max"], R)
            if pf_rows and self.PF_AFTER == "qkv_post":
                self._prefetch_tile_rows(st, W[p + "gu_w"], pf_rows)
            kp1 = st["kp"][1][i] if len(st["kp"]) > 1 else None
            vp1 = st["vp"][1][i] if len(st["vp"]) > 1 else None
            ops.call("sb_dec_attn", st["q"], st["kp"][0][i], st["vp"][0][i], kp1, vp1, rows_group0, P, st["kc"][i],
                     st["vc"][i], st["c_max"] * nkv * hd, st["c_max"], st["step"], nh, nkv, hd, hd ** -0.5, ws,
                     ws.numel(), st["attn"], R)
            if pf_rows and self.PF_AFTER == "combine":
                self._prefetch_tile_rows(st, W[p + "gu_w"], pf_rows)
            self._gemv(W[p + "o_w"], st["attn"], st["p_o"], S["o"])
            ops.call("sb_dec_residual_rmsnorm", st["x"], st["p_o"], S["o"], RP * H, H, W[p + "ln2_w"], st["xn"], R, H,
                     d.rms_eps)
            if S["gu"] == 1:     # SwiGLU fused into the GEMV epilogue (no fp32 partials, one kernel less)
                self._gemv(W[p + "gu_w"], st["xn"], st["act"], 1, swiglu=True)
            else:
                self._gemv(W[p + "gu_w"], st["xn"], st["p_gu"], S["gu"])
                ops.call("sb_dec_swiglu", st["p_gu"], S["gu"], RP * 2 * I, 2 * I, st["act"], R, I)
            self._gemv(W[p + "down_w"], st["act"], st["p_down"], S["down"])
            parts, sp = st["p_down"], S["down"]
        ops.call("sb_dec_residual_rmsnorm", st["x"], parts, sp, RP * H, H, W["norm_w"], st["xn"], R, H, d.rms_eps)
        self._gemv(W["lm_head"], st["xn"], st["logits"], 1)
        ops.call("sb_step_advance", st["step"])
        self._sample(st, samp, suppress_eos)
        if st["forked"]:     # join the prefetch branch (long finished): a capture must end with every fork joined
            torch.cuda.current_stream().wait_stream(st["side"])

    def _decode_graph(self, st, rope_base, rows_group0, sp, suppress_eos):
        """Capture one decode step into a CUDA graph (cached per decode state and step arguments).  Captured with the
        raw CUDAGraph API: the `torch.cuda.graph` context manager would empty the caching allocator, which makes
        every later phase of the training step re-map its memory."""
        key = (int(rope_base), int(rows_group0), sp.key(), bool(suppress_eos))
        hit = st["graphs"].get(key)
        if hit is not None:
            return hit
        snap = (st["step"].clone(), st["tokens"].clone(), st["finished"].clone(), st["out_ids"].clone())
        self._decode_step(st, rope_base, rows_group0, sp, suppress_eos)   # warm-up outside capture (func attributes)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        graph = torch.cuda.CUDAGraph()
        n0 = ops.direct_launch_count()
        with torch.cuda.stream(side):
            graph.capture_begin()
            try:
                self._decode_step(st, rope_base, rows_group0, sp, suppress_eos)
            finally:
                graph.capture_end()
        nodes = ops.direct_launch_count() - n0
        cur.wait_stream(side)
        st["step"].copy_(snap[0]); st["tokens"].copy_(snap[1]); st["finished"].copy_(snap[2]); st["out_ids"].copy_(snap[3])
        st["graphs"][key] = (graph, nodes)
        return graph, nodes

    # ---- generation parameters ------------------------------------------------------------------------------------
    # What `generate()` falls back to when NO generation_config is passed (the engine-level call of trainer.rollout and
    # of the kernel tests): the reference's rollout configuration (TRN:277-284) without HF's implicit top_k.
    ENGINE_DEFAULTS = dict(do_sample=True, temperature=1.0, top_k=0, top_p=0.95, repetition_penalty=1.0,
                           num_return_sequences=1, max_new_tokens=1024, min_new_tokens=0)
    # GenerationConfig._get_default_generation_params() (generation/configuration_utils.py:551-581): what HF fills in
    # for every field the caller's config and the checkpoint's generation_config.json leave at None
    HF_DEFAULTS = dict(do_sample=False, temperature=1.0, top_k=50, top_p=1.0, repetition_penalty=1.0,
                       num_return_sequences=1, max_new_tokens=None, min_new_tokens=0)
    _GEN_UNSUPPORTED = dict(num_beams=1, num_beam_groups=1, penalty_alpha=None, typical_p=1.0, epsilon_cutoff=0.0,
                            eta_cutoff=0.0, no_repeat_ngram_size=0, encoder_no_repeat_ngram_size=0, bad_words_ids=None,
                            min_p=None, length_penalty=1.0, diversity_penalty=0.0, encoder_repetition_penalty=1.0,
                            suppress_tokens=None, begin_suppress_tokens=None, forced_bos_token_id=None,
                            forced_eos_token_id=None, stop_strings=None, guidance_scale=None, sequence_bias=None,
                            exponential_decay_length_penalty=None, output_scores=False, output_logits=None,
                            return_dict_in_generate=False, max_time=None, assistant_model=None)

    def resolve_generation(self, generation_config=None, **explicit):
        """SamplingParams + (max_new_tokens, min_new_tokens, num_return_sequences) from, in priority order, the explicit
        keyword arguments that are not None, `generation_config` (a transformers.GenerationConfig or any object / dict
        with those attributes: TRN:277-302, 463), the checkpoint's generation_config.json (`self.generation_config`) and
        the defaults (HF's when a generation_config is involved, the engine's otherwise; generation/utils.py:1664-1729).
        Options this engine does not implement raise instead of being ignored."""
        d = self.dims
        layers = [{k: v for k, v in explicit.items() if v is not None}]
        hf_mode = generation_config is not None
        if hf_mode:
            gc = generation_config if isinstance(generation_config, dict) else {
                k: getattr(generation_config, k) for k in dir(generation_config)
                if not k.startswith("_") and not callable(getattr(generation_config, k, None))}
            layers.append({k: v for k, v in gc.items() if v is not None})
            layers.append({k: v for k, v in (self.generation_config or {}).items() if v is not None})
        merged = {}
        for lay in reversed(layers):
            merged.update(lay)
        for k, dflt in self._GEN_UNSUPPORTED.items():
            v = merged.get(k, dflt)
            if v is not None and v != dflt and not (dflt in (False, None) and not v):
                raise SpacerError(f"generate: {k}={v!r} is not supported by the B200 rollout engine")
        base = self.HF_DEFAULTS if hf_mode else self.ENGINE_DEFAULTS
        g = {k: merged.get(k, base[k]) for k in base}
        if g["max_new_tokens"] is None:
            raise SpacerError("generate: max_new_tokens is required (HF's max_length = 20 default is shorter than any prompt)")
        eos = merged.get("eos_token_id")
        if eos is None:
            eos = list(d.eos_ids) if d.eos_ids else [d.eos_id]
        eos = [int(e) for e in (eos if isinstance(eos, (list, tuple)) else [eos])]
        if not 1 <= len(eos) <= 4:
            raise SpacerError(f"generate: between 1 and 4 eos ids are supported, got {eos}")
        pad = merged.get("pad_token_id")
        pad = d.pad_id if pad is None else int(pad)
        temperature = float(g["temperature"])
        greedy = (not g["do_sample"]) or temperature <= 0.05 or int(g["top_k"] or 0) == 1
        if temperature <= 0.0 and g["do_sample"]:
            raise SpacerError("generate: temperature must be > 0")
        sp = SamplingParams(greedy=bool(greedy), top_p=float(g["top_p"]), top_k=int(g["top_k"] or 0),
                            temperature=1.0 if greedy else temperature,
                            repetition_penalty=float(g["repetition_penalty"]), eos_ids=tuple(eos), pad_id=pad)
        if not 0.0 < sp.top_p <= 1.0:
            raise SpacerError(f"generate: top_p must be in (0, 1], got {sp.top_p}")
        return sp, int(g["max_new_tokens"]), int(g["min_new_tokens"] or 0), int(g["num_return_sequences"])

    @torch.no_grad()
    def generate(self, input_ids, pixel_values_videos=None, video_grid_thw=None, *, generation_config=None,
                 max_new_tokens=None, num_return_sequences=None, top_p=None, top_k=None, temperature=None,
                 do_sample=None, repetition_penalty=None, eos_token_id=None, pad_token_id=None, min_new_tokens=None,
                 seed=0, pixel_values_videos_2=None, num_return_sequences_2=0, use_graph=True, attention_mask=None,
                 return_stats=False, second_per_grid_ts=None, pixel_values=None, image_grid_thw=None,
                 keep_vit_tape=False, use_cache=True, mm_token_type_ids=None):
        """Rollout of ONE prompt: `num_return_sequences` completions (TRN:463-467: generate(**prompt_inputs,
        generation_config=GenerationConfig(max_new_tokens, do_sample, top_p, temperature, num_return_sequences,
        pad_token_id))).  Returns LongTensor [G, P + C'] (prompt echoed, finished rows padded with pad_token_id).
        Sampling parameters: see `resolve_generation`.  `attention_mask` must be all ones (one un-padded prompt; use the
        HF-facing wrapper in hf_api.py for left-padded batches); `mm_token_type_ids` (transformers 5.x processors emit
        it) carries no information beyond the placeholder ids and is accepted for that reason only.

        `pixel_values_videos_2` / `num_return_sequences_2` add a second group decoded in the same batch from
        the same text with another video (T-GRPO's frame-shuffled rollout, TRN:442-458, 469-475); the result is
        then a tuple (ids_main, ids_second), each cut to its own longest completion."""
        d = self.dims
        if not use_cache:
            raise SpacerError("generate: use_cache=False is not supported (the rollout engine is a KV-cache decoder)")
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise SpacerError("generate: padded prompts are not supported here (attention_mask has zeros)")
        if pixel_values is not None:
            # image prompt (`pixel_values` / `image_grid_thw`, the other visual input the reference's processor can
            # produce, TRN:417-425): same tower, grids with t = 1, image placeholder tokens
            if pixel_values_videos is not None:
                raise SpacerError("generate: one visual input per prompt (image or video)")
            pixel_values_videos, video_grid_thw = pixel_values, image_grid_thw
        sp, max_new_tokens, min_new_tokens, G1 = self.resolve_generation(
            generation_config, max_new_tokens=max_new_tokens, num_return_sequences=num_return_sequences, top_p=top_p,
            top_k=top_k, temperature=temperature, do_sample=do_sample, repetition_penalty=repetition_penalty,
            eos_token_id=eos_token_id, pad_token_id=pad_token_id, min_new_tokens=min_new_tokens)
        if min_new_tokens not in (0, max_new_tokens):
            raise SpacerError("generate: min_new_tokens must be 0 or max_new_tokens")
        if torch.as_tensor(input_ids).dim() == 2 and input_ids.shape[0] != 1:
            raise SpacerError("generate: one prompt per call (input_ids [1, P]); hf_api loops over a batch")
        ids = input_ids.reshape(-1)
        P = ids.numel()
        G2 = int(num_return_sequences_2 if pixel_values_videos_2 is not None else 0)
        R = G1 + G2
        if R > 32:
            raise SpacerError("generate: at most 32 rows per prompt")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        ids_dev = ids.to(self.device, I32)
        pos, nxt = rope_index(ids, video_grid_thw, d, self.rope_convention, second_per_grid_ts)
        pos_dev = pos.to(I32).contiguous().to(self.device)
        meta = causal_meta(P, self.device)
        pixel_sets = [pixel_values_videos] + ([pixel_values_videos_2] if G2 > 0 else [])
        st = self._decode_state(R, P, max_new_tokens, len(pixel_sets))
        st["seed"].fill_(int(seed))
        if sp.repetition_penalty != 1.0:
            # RepetitionPenaltyLogitsProcessor looks at prompt + generated tokens (logits_process.py)
            st["seen"].zero_()
            ops.call("sb_token_bitmap_set", ids_dev, P, st["seen"], st["seen"].shape[1], R, d.vocab)
        self.vit_cache = None
        for k, pix in enumerate(pixel_sets):
            if keep_vit_tape and k == 0 and pix is not None:
                # the update that follows a rollout runs the SAME vision tower on the SAME pixels (one update per
                # rollout, TRN:526-528): keep its output and saved activations instead of recomputing the forward
                vtape, ptape = {}, {}
                vis = self.vit_forward(pix, video_grid_thw, vtape)
                self.vit_cache = dict(pixels=pix, vis=vis, tape=vtape, llm_tape=ptape, prompt_ids=ids.cpu().long(),
                                      prompt_pos=pos)
            else:
                ptape = None
                vis = self.vit_forward(pix, video_grid_thw) if pix is not None else None
            kv = [(st["kp"][k][i], st["vp"][k][i]) for i in range(d.layers)]
            hf = self.llm_forward(ids_dev, vis, pos_dev, meta, tape=ptape, kv_out=kv)
            # first token: same distribution for every row of a group, independent draws
            r0, r1 = (0, G1) if k == 0 else (G1, R)
            st["xn"][r0:r1].copy_(hf[P - 1][None].expand(r1 - r0, -1))
            del hf, vis
        self._gemv(self.params["lm_head"], st["xn"], st["logits"], 1)
        suppress = min_new_tokens > 0
        self._sample(st, sp, suppress)
        n_steps = max_new_tokens - 1
        graph, graph_nodes, replays = None, 0, 0
        if n_steps > 0 and use_graph:
            graph, graph_nodes = self._decode_graph(st, nxt, G1, sp, suppress)
        ev[1].record()
        done = 0
        while done < n_steps:
            burst = min(32, n_steps - done)
            for _ in range(burst):
                if graph is not None:
                    graph.replay()
                    replays += 1
                else:
                    self._decode_step(st, nxt, G1, sp, suppress)
            done += burst
            if not suppress and done < n_steps and bool(st["finished"][:R].all().item()):
                break
        ops.note_graph_replay(graph_nodes * replays)
        ev[2].record()
        out = st["out_ids"].long()
        steps_done = int(st["step"].item())      # host sync: everything above has completed
        widths = [out.shape[1]] * 2

        def cut(block):
            """One generate() call of the reference = one group: width = its longest completion (generation stops when
            every row of the CALL is finished), rows that finished early padded (generation/utils.py:2797)."""
            if suppress or block.shape[0] == 0:
                return block
            is_eos = torch.zeros_like(block, dtype=torch.bool)
            for e in sp.eos_ids:
                is_eos |= block == e
            n = block.shape[0]
            first = torch.where(is_eos.any(1), is_eos.int().argmax(1),
                                torch.full((n,), block.shape[1] - 1, device=self.device))
            width = min(int(first.max().item()) + 1, steps_done + 1)
            block = block[:, :width]
            col = torch.arange(width, device=self.device)[None]
            return torch.where(col > first[:, None], torch.full_like(block, sp.pad_id), block)

        prompt = ids.to(self.device).long()[None]
        out1 = cut(out[:G1])
        res1 = torch.cat([prompt.expand(G1, -1), out1], dim=1)
        dec_ms = ev[1].elapsed_time(ev[2])
        n_loop = max(1, done)
        # algorithmic HBM bytes of one decode step: every LLM-layer weight + lm_head once, the shared
        # prompt KV once per group, each row's own completion KV (average over the loop)
        kv_tok = 2 * d.layers * d.kv_heads * d.head_dim * 2
        w_bytes = self.decode_weight_bytes()
        kv_bytes = len(pixel_sets) * P * kv_tok + R * (n_loop / 2.0) * kv_tok
        stats = dict(decode_steps=steps_done, rows=R, prompt_len=P, rollout_ms=ev[0].elapsed_time(ev[2]),
                     prefill_ms=ev[0].elapsed_time(ev[1]), decode_ms=dec_ms, decode_ms_per_step=dec_ms / n_loop,
                     decode_bytes_per_step=w_bytes + kv_bytes,
                     decode_gbs=(w_bytes + kv_bytes) * n_loop / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else None,
                     graph_nodes=graph_nodes, graph_replays=replays)
        self.last_generate_stats = stats
        self._last_decode_state = st   # the cached decode state (parity tests read the final step's logits)
        if G2 > 0:
            res2 = torch.cat([prompt.expand(G2, -1), cut(out[G1:])], dim=1)
            return (res1, res2, stats) if return_stats else (res1, res2)
        return (res1, stats) if return_stats else res1

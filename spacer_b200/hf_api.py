"""The `transformers` model surface the reference trainer drives, on the B200 engine.

`SGRLVRTrainer` (SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py, "TRN") talks to its model through
exactly four calls; this module makes each of them work unmodified against `Qwen2VLB200`:

    Qwen2VLForConditionalGeneration.from_pretrained(model_id, attn_implementation=..., torch_dtype=..., use_cache=...)
                                                                                                   TRN:163-190, 205-214
    unwrapped_model.generate(**prompt_inputs, generation_config=self.generation_config)           TRN:463, 473, 481
    model(input_ids, **prompt_inputs).logits            (autograd-connected for the policy)       TRN:357
    loss.backward()  ->  parameters' .grad              (accelerator.backward, HF Trainer)        TRN:686

`Qwen2VLForConditionalGenerationB200` / `Qwen2_5_VLForConditionalGenerationB200` are `torch.nn.Module`s whose two
parameters are the engine's flat bf16 arenas (`state_dict()` / `load_state_dict()` still speak the HF names).  The
materialised-logits path (`.logits`, 2 x G x L x V bytes like the reference) exists for drop-in fidelity; the fast path is
`get_per_token_logps`, a one-line replacement for `SGRLVRTrainer._get_per_token_logps` that returns the same `[B, L-1]`
tensor with a `grad_fn` but never builds `[., V]` logits (lm_head -> online logsumexp -> gather in the GEMM epilogue).
Evaluation (`SpaceR-Eval/data_utils/vsibench.py:157-180`: left-padded batches, `generate(**inputs, use_cache=True,
max_new_tokens=..., temperature=0.01)` over the checkpoint's generation_config.json) goes through the same `generate`.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .config import ModelDims
from .model import BF16, F32, I32, GradStore, Qwen2VLB200, pack_prompt_completions
from .ops import SpacerError


def _visual_inputs(B, pixel_values_videos, video_grid_thw, pixel_values, image_grid_thw):
    """(pixels, grid) of the ONE visual input all B rows share.  The reference repeats pixels and grid once per row
    before the scoring forwards (TRN:507-521); the repeats are dropped here, after checking that they are repeats."""
    if pixel_values is not None and pixel_values_videos is not None:
        raise SpacerError("one visual input per prompt (image or video)")
    pix, grid = (pixel_values, image_grid_thw) if pixel_values is not None else (pixel_values_videos, video_grid_thw)
    if pix is None:
        return None, None
    if grid is None:
        raise SpacerError("pixel values passed without their grid_thw")
    grid = torch.as_tensor(grid).reshape(-1, 3).cpu()
    if grid.shape[0] == B and B > 1:
        if not bool((grid == grid[:1]).all()):
            raise SpacerError("the rows of one scoring batch must share one visual input (one prompt per step, TRN:417-425)")
        n_p = int(grid[0, 0] * grid[0, 1] * grid[0, 2])
        if pix.shape[0] != B * n_p:
            raise SpacerError(f"pixel rows ({pix.shape[0]}) do not match {B} x grid ({n_p})")
        pix, grid = pix[:n_p], grid[:1]
    return pix, grid


def _common_prefix(ids: torch.Tensor) -> int:
    """Length of the longest common prefix of the rows of ids [B, L], clamped to [1, L-1] (rows of a GRPO group share the
    prompt; at least one column stays 'completion' and at least one 'prefix' so that the packed layout is defined)."""
    B, L = ids.shape
    same = (ids == ids[:1]).all(0)
    P = int(L if bool(same.all()) else torch.nonzero(~same)[0, 0])
    return max(1, min(P, L - 1))


class _PackedScoring:
    """Host-side plan of one scoring call on `input_ids [B, L]`: the packed batch [prefix | tail_0 | ... | tail_{B-1}], the
    hidden rows that predict every token 1..L-1 of every row (prefix rows once) and their targets."""

    def __init__(self, engine: Qwen2VLB200, input_ids, pix, grid):
        ids = torch.as_tensor(input_ids).cpu().long()
        if ids.dim() != 2:
            raise SpacerError("input_ids must be [B, L]")
        self.B, self.L = ids.shape
        if self.L < 2:
            raise SpacerError("input_ids must hold at least two tokens per row")
        self.P = _common_prefix(ids)
        dev = engine.device
        self.batch = pack_prompt_completions(ids[0, :self.P], ids[:, self.P:], grid, engine.dims, dev,
                                             engine.rope_convention)
        # rows that predict the tail tokens (B * (L - P), row-major), then rows 0..P-2 that predict prefix tokens 1..P-1
        self.rows = torch.cat([self.batch.rows, torch.arange(self.P - 1, device=dev, dtype=I32)])
        self.targets = torch.cat([self.batch.targets, ids[0, 1:self.P].to(I32).to(dev)])
        self.n_tail = self.B * (self.L - self.P)

    def assemble(self, per_row):
        """per_row [R, ...] in `rows` order -> [B, L-1, ...]: prefix predictions broadcast, tails per row."""
        B, L, P = self.B, self.L, self.P
        tail = per_row[:self.n_tail].reshape(B, L - P, *per_row.shape[1:])
        pre = per_row[self.n_tail:]
        return torch.cat([pre.unsqueeze(0).expand(B, *pre.shape), tail], dim=1)

    def disassemble(self, grad):
        """Adjoint of assemble: grad [B, L-1, ...] -> [R, ...] (prefix rows summed over the B copies)."""
        P = self.P
        tail = grad[:, P - 1:].reshape(self.n_tail, *grad.shape[2:])
        pre = grad[:, :P - 1].sum(0)
        return torch.cat([tail, pre], dim=0)


class _PerTokenLogps(torch.autograd.Function):
    """[B, L-1] log-probs of tokens 1..L-1 with a backward that runs the engine's fused lm_head / decoder / ViT
    backward into the module's gradient arenas.  `mat` / `vec` (the parameter arenas) are inputs only so that autograd
    routes the result to their `.grad`."""

    @staticmethod
    def forward(ctx, mat, vec, module, plan, pix, grid):
        eng = module.engine
        H = eng.dims.hidden
        with torch.no_grad():
            hf, state = module._hidden(plan, pix, grid, want_grad=not isinstance(ctx, SimpleNamespace))
            R = plan.rows.numel()
            hsel = torch.empty((R, H), device=eng.device, dtype=BF16)
            ops.call("sb_gather_rows", hf, plan.rows, hsel, R, H)
            part, tl, nt = eng._lmhead_partials(hsel, plan.targets)
            lp = torch.empty(R, device=eng.device, dtype=F32)
            ops.call("sb_logprob_from_partials", part, nt, tl, lp, R)
        ctx.module, ctx.plan, ctx.state, ctx.hsel, ctx.lse, ctx.T = module, plan, state, hsel, tl - lp, hf.shape[0]
        return plan.assemble(lp)

    @staticmethod
    def backward(ctx, d_lp):
        module, plan, eng = ctx.module, ctx.plan, ctx.module.engine
        grads = module.grad_store()
        with torch.no_grad():
            coef = plan.disassemble(d_lp.to(F32)).contiguous()
            grads.zero_for_step()
            d_hsel = eng.lm_head_backward(ctx.hsel, plan.targets, ctx.lse, coef, grads)
            rows_plan = module._rows_plan(plan)
            d_hf = eng.scatter_rows(d_hsel, plan.rows, rows_plan, ctx.T)
            eng.backward_hidden(ctx.state, d_hf, grads)
        ctx.state = ctx.hsel = None
        return module._publish_grads(grads) + (None, None, None, None)


class _Logits(torch.autograd.Function):
    """Materialised `[B, L, V]` logits (what TRN:357 reads through `.logits`), autograd-connected."""

    @staticmethod
    def forward(ctx, mat, vec, module, plan, pix, grid):
        eng = module.engine
        d = eng.dims
        with torch.no_grad():
            hf, state = module._hidden(plan, pix, grid, want_grad=not isinstance(ctx, SimpleNamespace))
            B, L, P = plan.B, plan.L, plan.P
            # every position: prefix rows 0..P-1 once, then the tails; row (b, L-1) predicts nothing the trainer uses but
            # HF returns it, so it is computed too
            logits_packed = ops.gemm(hf, eng.params["lm_head"])                      # [P + B*(L-P), V] bf16
            out = torch.empty((B, L, d.vocab), device=eng.device, dtype=BF16)
            out[:, :P] = logits_packed[:P][None]
            out[:, P:] = logits_packed[P:].view(B, L - P, d.vocab)
        ctx.module, ctx.plan, ctx.state, ctx.hf = module, plan, state, hf
        return out

    @staticmethod
    def backward(ctx, d_logits):
        module, plan, eng = ctx.module, ctx.plan, ctx.module.engine
        B, L, P = plan.B, plan.L, plan.P
        grads = module.grad_store()
        with torch.no_grad():
            V = eng.dims.vocab
            dl = torch.empty((P + B * (L - P), V), device=eng.device, dtype=BF16)
            dl[:P] = d_logits[:, :P].sum(0, dtype=F32).to(BF16)
            dl[P:] = d_logits[:, P:].reshape(-1, V)
            grads.zero_for_step()
            g_lm = grads["lm_head"]
            d_hf = ops.gemm(dl, eng.params["lm_head"], b_mn=True)
            ops.gemm(dl, ctx.hf, a_mn=True, b_mn=True, out=g_lm, residual=g_lm if eng.dims.tie else None)
            if not eng.dims.tie:
                grads.ready("lm_head")
            del dl
            eng.backward_hidden(ctx.state, d_hf, grads)
        ctx.state = ctx.hf = None
        return module._publish_grads(grads) + (None, None, None, None)


class Qwen2VLForConditionalGenerationB200(torch.nn.Module):
    """Stands where the reference builds `transformers.Qwen2VLForConditionalGeneration` (TRN:183, 207)."""

    variant = "qwen2_vl"

    def __init__(self, engine: Qwen2VLB200):
        super().__init__()
        if not isinstance(engine, Qwen2VLB200):
            raise SpacerError("wrap a Qwen2VLB200 engine (use from_pretrained / from_dims)")
        self.engine = engine
        # the two flat arenas ARE the parameters (shared storage): optimizers, DDP and .grad see the real weights
        self.mat = torch.nn.Parameter(engine.params.mat, requires_grad=True)
        self.vec = torch.nn.Parameter(engine.params.vec, requires_grad=True)
        self.config = engine.config
        self.warnings_issued = {}                    # TRN:312
        self._grads = None
        self._reuse_rollout = True

    # ---- construction / checkpoint surface -----------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, device=None, rope_convention="classic",
                        attn_implementation=None, torch_dtype=None, dtype=None, use_cache=None, device_map=None,
                        low_cpu_mem_usage=None, trust_remote_code=None, **unknown):
        """Local HF-format directory -> module.  The keyword arguments the reference passes (`attn_implementation`,
        `torch_dtype`, `use_cache`: TRN:163-190; `device_map` in SpaceR-Eval) select things this engine fixes itself
        (tcgen05 attention, bf16 weights, KV-cache decode); they are validated, anything else raises."""
        if unknown:
            raise TypeError(f"from_pretrained: unsupported arguments {sorted(unknown)}")
        for dt in (torch_dtype, dtype):
            if dt not in (None, "auto", "bfloat16", torch.bfloat16):
                raise SpacerError(f"from_pretrained: the B200 engine computes in bf16; torch_dtype={dt!r} is not available")
        dev = device if device is not None else (f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cuda")
        eng = Qwen2VLB200.from_pretrained(pretrained_model_name_or_path, dev, rope_convention=rope_convention)
        if eng.dims.variant != cls.variant:
            raise SpacerError(f"{pretrained_model_name_or_path} holds a {eng.dims.variant} checkpoint; use the matching class")
        return cls(eng)

    @classmethod
    def from_dims(cls, dims: ModelDims, device="cuda", seed: int | None = None, rope_convention="classic"):
        eng = Qwen2VLB200(dims, device, rope_convention=rope_convention)
        if seed is not None:
            eng.params.init_random(seed)
        return cls(eng)

    def save_pretrained(self, path, **kw):
        self.engine.save_pretrained(path, **{k: v for k, v in kw.items() if k == "max_shard_bytes"})

    def state_dict(self, *args, **kwargs):
        """HF parameter names (checkpoint save at open_r1/SG-RLVR.py:384), not the arena names; references to the live
        weights where the layouts agree, like `nn.Module.state_dict()`."""
        return self.engine.params.state_dict(clone=False)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self.engine.load_state_dict(state_dict)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @property
    def generation_config(self):
        return SimpleNamespace(**self.engine.generation_config)

    @property
    def device(self):
        return self.engine.device

    @property
    def dtype(self):
        return torch.bfloat16

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        """HF Trainer calls it under --gradient_checkpointing true (run_SpaceR_SG_RLVR.sh:27): the engine then recomputes
        its widest activation (the gate|up projection) in the backward instead of keeping it."""
        self.engine.gradient_checkpointing_enable()

    def gradient_checkpointing_disable(self):
        self.engine.gradient_checkpointing_disable()

    def enable_input_require_grads(self):
        pass

    def get_input_embeddings(self):
        return SimpleNamespace(weight=self.engine.params["embed"])

    def tie_weights(self, *a, **k):
        pass

    # ---- gradients ---------------------------------------------------------------------------------------------------
    def grad_store(self) -> GradStore:
        if self._grads is None:
            self._grads = GradStore(self.engine.params)
        return self._grads

    def _publish_grads(self, grads: GradStore):
        """What the autograd Functions return for (mat, vec): the gradient arenas (autograd accumulates them into
        `.grad`; the fp32 vector gradients are cast to the parameter dtype)."""
        return grads.mat, grads.vec.to(BF16)

    def hf_named_grads(self):
        """(HF parameter name, gradient) pairs of the last backward, from `.grad` of the arenas."""
        if self.mat.grad is None:
            return iter(())
        p = self.engine.params
        views = {n: p.view_of(n, self.mat.grad, self.vec.grad) for n in p.index}
        return p.hf_items(views)

    def _take_vit_cache(self, pix):
        """The rollout's ViT forward / prompt prefill (generate(keep_vit_tape=True)) for the policy scoring call that
        follows it on the same pixels.  Single use; released when the pixels differ (e.g. the frame-shuffled rollout of
        TRN:473 ran last)."""
        vc, self.engine.vit_cache = self.engine.vit_cache, None
        if not self._reuse_rollout or vc is None or pix is None or vc.get("tape") is None:
            return None
        old = vc["pixels"]
        if old is not pix and not (old.shape == pix.shape and old.dtype == pix.dtype and bool(torch.equal(old, pix))):
            return None
        vc["pixels"] = pix            # forward_hidden checks identity
        return vc

    def _hidden(self, plan, pix, grid, want_grad: bool):
        """Final hidden states of the packed batch; with `want_grad` the activations are saved (and the rollout's forward
        reused when it ran on the same inputs), otherwise nothing is kept."""
        eng = self.engine
        if want_grad:
            return eng.forward_hidden(plan.batch, pix, grid, self._take_vit_cache(pix))
        vis = eng.vit_forward(pix, grid) if pix is not None else None
        return eng.llm_forward(plan.batch.ids, vis, plan.batch.pos, plan.batch.meta), None

    @staticmethod
    def _rows_plan(plan):
        from .model import segment_plan
        if getattr(plan, "_rows_plan", None) is None:
            plan._rows_plan = segment_plan(plan.rows, plan.rows.device)
        return plan._rows_plan

    # ---- the calls the trainer makes -----------------------------------------------------------------------------------
    def _scoring_inputs(self, input_ids, attention_mask, pixel_values_videos, video_grid_thw, pixel_values,
                        image_grid_thw):
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise SpacerError("forward: padded rows are not supported (the reference passes no attention_mask, TRN:357)")
        B = torch.as_tensor(input_ids).shape[0]
        pix, grid = _visual_inputs(B, pixel_values_videos, video_grid_thw, pixel_values, image_grid_thw)
        if pix is not None:
            pix = pix.to(self.engine.device)
        return _PackedScoring(self.engine, input_ids, pix, grid), pix, grid

    def forward(self, input_ids=None, attention_mask=None, pixel_values_videos=None, video_grid_thw=None,
                pixel_values=None, image_grid_thw=None, second_per_grid_ts=None, mm_token_type_ids=None,
                position_ids=None, use_cache=None, return_dict=True, logits_to_keep=0):
        """`model(input_ids, **prompt_inputs).logits` (TRN:357): `[B, L, V]` bf16, autograd-connected when gradients are
        enabled.  Rows must share their visual input (the x B-repeated pixels of TRN:507-521 are recognised) -- one prompt
        per step, like the reference.  `second_per_grid_ts` is ignored exactly as the reference deletes it (TRN:519-520);
        `mm_token_type_ids` repeats what the placeholder ids say; explicit `position_ids` are not accepted (the engine
        computes M-RoPE ids itself, `rope_convention`)."""
        if position_ids is not None or logits_to_keep not in (0, None):
            raise SpacerError("forward: position_ids / logits_to_keep are not supported")
        if input_ids is None:
            raise SpacerError("forward: input_ids is required (inputs_embeds is not supported)")
        plan, pix, grid = self._scoring_inputs(input_ids, attention_mask, pixel_values_videos, video_grid_thw,
                                               pixel_values, image_grid_thw)
        if torch.is_grad_enabled() and (self.mat.requires_grad or self.vec.requires_grad):
            logits = _Logits.apply(self.mat, self.vec, self, plan, pix, grid)
        else:
            with torch.no_grad():
                logits = _Logits.forward(SimpleNamespace(), self.mat, self.vec, self, plan, pix, grid)
        return SimpleNamespace(logits=logits, loss=None, past_key_values=None) if return_dict else (logits,)

    def per_token_logps(self, input_ids, attention_mask=None, pixel_values_videos=None, video_grid_thw=None,
                        pixel_values=None, image_grid_thw=None, second_per_grid_ts=None, mm_token_type_ids=None):
        """`[B, L-1]` fp32 log-probs of tokens 1..L-1 (what TRN:353-366 computes from the logits) without materialising
        `[., V]`; carries a grad_fn when gradients are enabled."""
        plan, pix, grid = self._scoring_inputs(input_ids, attention_mask, pixel_values_videos, video_grid_thw,
                                               pixel_values, image_grid_thw)
        if torch.is_grad_enabled() and (self.mat.requires_grad or self.vec.requires_grad):
            return _PerTokenLogps.apply(self.mat, self.vec, self, plan, pix, grid)
        with torch.no_grad():
            return _PerTokenLogps.forward(SimpleNamespace(), self.mat, self.vec, self, plan, pix, grid)

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, pixel_values_videos=None, video_grid_thw=None,
                 pixel_values=None, image_grid_thw=None, second_per_grid_ts=None, mm_token_type_ids=None,
                 generation_config=None, seed=None, **gen_kwargs):
        """`generate(**prompt_inputs, generation_config=...)` (TRN:463-481) and SpaceR-Eval's
        `generate(**inputs, use_cache=True, max_new_tokens=..., temperature=0.01)`.  `input_ids [B, P]` may be
        left-padded (`attention_mask`); every row is an independent prompt with its own visual input (pixel rows are split
        by the grids, one grid per prompt).  Returns `[B * num_return_sequences, P + C']`, finished rows right-padded with
        pad_token_id, like HF.  Unknown generation options raise (engine.resolve_generation)."""
        eng = self.engine
        if input_ids is None:
            raise SpacerError("generate: input_ids is required")
        ids = torch.as_tensor(input_ids)
        if ids.dim() == 1:
            ids = ids[None]
        B, Pw = ids.shape
        am = torch.ones_like(ids) if attention_mask is None else torch.as_tensor(attention_mask).to(ids.device)
        if pixel_values is not None and pixel_values_videos is not None:
            raise SpacerError("generate: one visual input per prompt (image or video)")
        is_image = pixel_values is not None
        pix_all = pixel_values if is_image else pixel_values_videos
        grids = image_grid_thw if is_image else video_grid_thw
        if pix_all is not None:
            grids = torch.as_tensor(grids).reshape(-1, 3).cpu()
            if grids.shape[0] != B:
                raise SpacerError(f"generate: {grids.shape[0]} visual grids for {B} prompts (one visual input per prompt)")
            counts = (grids[:, 0] * grids[:, 1] * grids[:, 2]).tolist()
            if sum(counts) != pix_all.shape[0]:
                raise SpacerError("generate: pixel rows do not match the grids")
        if seed is None:
            # like HF, consecutive calls draw from an advancing stream (the global torch generator seeds it)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        gc = generation_config if generation_config is not None else {}
        outs, off = [], 0
        for b in range(B):
            keep = am[b].bool()
            if not bool(keep.any()):
                raise SpacerError("generate: empty prompt row")
            first = int(torch.nonzero(keep)[0, 0])
            if not bool(keep[first:].all()):
                raise SpacerError("generate: only LEFT padding is supported (padding_side='left', TRN:417-425)")
            row = ids[b, first:]
            kw = {}
            if pix_all is not None:
                n = int(counts[b])
                px = pix_all[off:off + n].to(eng.device)
                off += n
                kw = dict(pixel_values=px, image_grid_thw=grids[b:b + 1]) if is_image else \
                    dict(pixel_values_videos=px, video_grid_thw=grids[b:b + 1])
            sec = None
            if second_per_grid_ts is not None:
                s_all = torch.as_tensor(second_per_grid_ts).reshape(-1).tolist()
                sec = [s_all[min(b, len(s_all) - 1)]]
            out = eng.generate(row[None], generation_config=gc, seed=seed + 104729 * b, second_per_grid_ts=sec,
                               keep_vit_tape=(B == 1 and self._reuse_rollout and self.training), **kw, **gen_kwargs)
            outs.append((first, out))
        sp, _, _, _ = eng.resolve_generation(gc, **{k: v for k, v in gen_kwargs.items()
                                                     if k in ("eos_token_id", "pad_token_id", "max_new_tokens")})
        width = max(o.shape[1] - (Pw - f) for f, o in outs)            # longest completion of the batch
        res = []
        for b, (first, o) in enumerate(outs):
            G = o.shape[0]
            comp = o[:, Pw - first:]
            block = torch.full((G, Pw + width), sp.pad_id, dtype=torch.long, device=o.device)
            block[:, :Pw] = ids[b].to(o.device)[None]                  # the (left-padded) prompt echoed verbatim
            block[:, Pw:Pw + comp.shape[1]] = comp
            res.append(block)
        return torch.cat(res, dim=0)


class Qwen2_5_VLForConditionalGenerationB200(Qwen2VLForConditionalGenerationB200):
    """Stands where the reference builds `transformers.Qwen2_5_VLForConditionalGeneration` (TRN:185, 190, 209, 214;
    SpaceR-Eval/data_utils/vsibench.py:83-92) -- the family run_SpaceR_SG_RLVR.sh:16 trains."""

    variant = "qwen2_5_vl"


def get_per_token_logps(self, model, input_ids, **kwargs):
    """Drop-in for `SGRLVRTrainer._get_per_token_logps` (TRN:353-366):

        from spacer_b200.hf_api import get_per_token_logps
        SGRLVRTrainer._get_per_token_logps = get_per_token_logps

    Same arguments, same `[B, L-1]` result (fp32 instead of bf16 log-softmax: strictly more accurate), same autograd
    contract, without the `[B, L, V]` logits.  Falls back to the reference's own arithmetic for any other model."""
    inner = getattr(model, "module", model)            # DDP / accelerate wrappers
    if isinstance(inner, Qwen2VLForConditionalGenerationB200):
        return inner.per_token_logps(input_ids, **kwargs)
    logits = model(input_ids, **kwargs).logits[:, :-1, :]
    ids = input_ids[:, 1:]
    return torch.stack([torch.gather(lg.log_softmax(dim=-1), 1, i.unsqueeze(1)).squeeze(1) for lg, i in zip(logits, ids)])


class FusedAdamW(torch.optim.Optimizer):
    """`optimizers=(FusedAdamW(model, ...), None)` for the HF Trainer: the engine's clipped AdamW with fp32 master weights
    (`sb_adamw_step`, DeepSpeed semantics of zero3.json:10-12) behind the torch.optim interface.  Reads the gradients the
    autograd Functions above accumulated into `model.mat.grad` / `model.vec.grad`."""

    def __init__(self, model: Qwen2VLForConditionalGenerationB200, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                 max_grad_norm=5.0):
        from .trainer import AdamW, GRPOConfig
        super().__init__([model.mat, model.vec], dict(lr=lr))
        self.model = model
        cfg = GRPOConfig(learning_rate=lr, weight_decay=weight_decay, adam_beta1=betas[0], adam_beta2=betas[1],
                         adam_eps=eps, max_grad_norm=max_grad_norm, lr_scheduler_type="constant")
        self.inner = AdamW(model.engine.params, cfg)

    @torch.no_grad()
    def step(self, closure=None):
        m = self.model
        if m.mat.grad is None:
            return None
        g = m.grad_store()
        if m.mat.grad.data_ptr() != g.mat.data_ptr():
            g.mat.copy_(m.mat.grad)
        g.vec.copy_(m.vec.grad)
        self.inner.cfg.learning_rate = self.param_groups[0]["lr"]      # LR schedulers write here
        self.inner.step(g)
        return None

"""The vLLM engine surface `Qwen2VLGRPOVLLMTrainerModified` drives, on the colocated B200 rollout engine
(/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/vllm_grpo_trainer_modified.py, "VTRN").

The vLLM trainer keeps a separate inference engine on a spare GPU and, every step,
  1. pushes the policy weights into it   `llm.llm_engine.model_executor.driver_worker.model_runner.model
                                           .load_weights(state_dict.items())`                          VTRN:526-543
  2. gathers every rank's prompt + frames on the main process (`gather_object`)                         VTRN:546-560
  3. runs ONE `llm.generate(inputs, sampling_params=SamplingParams(temperature=1.0, top_p=0.95,
     max_tokens=C, n=G), use_tqdm=False)` -> `outputs[i].outputs[j].token_ids`                          VTRN:566-590
  4. broadcasts the token ids back and slices per rank (`broadcast_object_list`)                        VTRN:600-608
`LLM` below is that object over `Qwen2VLB200`.  Built around the engine that trains (`LLM(engine=policy)`), step 1 is a
no-op -- the rollout reads the very arenas the optimizer writes -- and `load_weights` only copies when it is handed foreign
tensors.  `gather_generate_broadcast` is steps 2-4 over torch.distributed for callers without accelerate.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Callable, Sequence

import torch

from .model import Qwen2VLB200
from .ops import SpacerError


@dataclass
class SamplingParams:
    """vllm.SamplingParams, the fields the reference sets (VTRN:383-387, 569-571) plus the ones this sampler implements.
    vLLM conventions: top_k = -1 disables the cut, n completions per prompt, max_tokens new tokens."""
    n: int = 1
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    repetition_penalty: float = 1.0
    max_tokens: int = 16
    min_tokens: int = 0
    seed: int | None = None
    stop_token_ids: list | None = None
    ignore_eos: bool = False

    def clone(self):
        return copy.deepcopy(self)


@dataclass
class CompletionOutput:
    index: int
    token_ids: list
    text: str = ""
    finish_reason: str = "length"


@dataclass
class RequestOutput:
    request_id: str
    prompt: str | None
    prompt_token_ids: list
    outputs: list = field(default_factory=list)


class _ModelHandle:
    """What sits at `llm.llm_engine.model_executor.driver_worker.model_runner.model`: `load_weights(name/tensor pairs)`."""

    def __init__(self, engine: Qwen2VLB200):
        self.engine = engine
        self.loads = 0            # how many calls copied anything (0 forever when colocated with the trainer)

    def load_weights(self, weights):
        """VTRN:539-543: `llm_model.load_weights(state_dict.items())`.  HF names (both the 4.x and the 5.x spelling).  Tensors
        that already ARE the engine's storage (the colocated case) are skipped; the rest is copied into the arenas."""
        from . import hub
        sd = hub.normalize_names(dict(weights))
        exported = getattr(self.engine.params, "_export_ptrs", None)
        if exported and all(exported.get(k) == v.data_ptr() for k, v in sd.items()):
            return set(sd)        # the engine's own `state_dict()` handed straight back: nothing to move
        mine = dict(self.engine.params.hf_items())
        foreign = {}
        for k, v in sd.items():
            if k not in mine:
                raise SpacerError(f"load_weights: unknown parameter {k}")
            t = mine[k]
            same = (v.device == t.device and v.dtype == t.dtype and v.shape == t.shape and v.is_contiguous() == t.is_contiguous()
                    and v.data_ptr() == t.data_ptr())
            if not same:
                foreign[k] = v
        if not foreign:
            return set(sd)
        full = {k: foreign.get(k, mine[k]) for k in mine}
        self.engine.load_state_dict(full)
        self.loads += 1
        return set(sd)


class LLM:
    """`vllm.LLM` as the reference constructs and calls it (VTRN:361-382, 573-590)."""

    def __init__(self, model=None, *, engine: Qwen2VLB200 | None = None, processor: Callable | None = None,
                 device="cuda", dtype=torch.bfloat16, gpu_memory_utilization=None, enable_prefix_caching=True,
                 enforce_eager=True, mm_processor_kwargs=None, max_model_len=None, **unknown):
        """`engine`: the training engine to roll out from (colocated; weight sync is free).  Otherwise `model` is a local
        HF checkpoint directory.  `processor(text=[str], images=..., videos=[frames], return_tensors="pt", ...)` is the HF
        processor the trainer already holds (`processing_class`); it is needed for text prompts, not for
        `{"prompt_token_ids": ...}` inputs.  Prefix caching is what this engine always does (one prefill per prompt, shared
        by its n completions); eager/graph execution is its own business; both flags are accepted for that reason."""
        if unknown:
            raise TypeError(f"LLM: unsupported arguments {sorted(unknown)}")
        if dtype not in (torch.bfloat16, "bfloat16", "auto", None):
            raise SpacerError("LLM: the B200 engine computes in bf16")
        if engine is None:
            if model is None:
                raise SpacerError("LLM: pass engine=<Qwen2VLB200> (colocated) or model=<local checkpoint directory>")
            engine = Qwen2VLB200.from_pretrained(model, device)
        self.engine, self.processor, self.max_model_len = engine, processor, max_model_len
        handle = _ModelHandle(engine)
        self.llm_engine = SimpleNamespace(model_executor=SimpleNamespace(driver_worker=SimpleNamespace(
            model_runner=SimpleNamespace(model=handle))))
        self._calls = 0

    # ------------------------------------------------------------------------------------------------------------
    def _model_inputs(self, item):
        """One request -> (input_ids [1, P] long, kwargs for engine.generate)."""
        from . import vision
        eng = self.engine
        if isinstance(item, str):
            item = {"prompt": item}
        mm = item.get("multi_modal_data") or {}
        if len(mm) > 1:
            raise SpacerError("LLM.generate: one visual input per prompt (image or video)")
        kind, data = next(iter(mm.items())) if mm else (None, None)
        if isinstance(data, (list, tuple)):
            if len(data) != 1:
                raise SpacerError("LLM.generate: one image / one video per prompt")
            data = data[0]
        if "prompt_token_ids" in item:
            ids = torch.as_tensor(item["prompt_token_ids"], dtype=torch.long).reshape(1, -1)
            kw = {}
            if data is not None:
                if not torch.is_tensor(data):
                    raise SpacerError("LLM.generate: with prompt_token_ids the visual input must be a frame tensor [F, 3, H, W]")
                frames = data if data.dim() == 4 else data[None]
                frames = frames.to(eng.device)
                if frames.dtype not in (torch.uint8, torch.float32):
                    frames = frames.float()
                if kind != "video":       # an image is one temporal patch: the frame repeated t_patch times (HF processor)
                    frames = frames[:1].repeat(eng.dims.t_patch, 1, 1, 1)
                pix, _, grid = vision.patchify(frames)
                kw = dict(pixel_values_videos=pix, video_grid_thw=grid) if kind == "video" else \
                    dict(pixel_values=pix, image_grid_thw=grid)
            return ids, kw
        if self.processor is None:
            raise SpacerError("LLM.generate: text prompts need the HF processor (LLM(..., processor=processing_class))")
        enc = self.processor(text=[item["prompt"]], images=[data] if kind == "image" else None,
                             videos=[data] if kind == "video" else None, return_tensors="pt", padding=True,
                             padding_side="left", add_special_tokens=False)
        ids = torch.as_tensor(enc["input_ids"]).reshape(1, -1)
        kw = {k: enc[k] for k in ("pixel_values_videos", "video_grid_thw", "pixel_values", "image_grid_thw",
                                  "second_per_grid_ts") if k in enc}
        for k in ("pixel_values_videos", "pixel_values"):
            if k in kw:
                kw[k] = kw[k].to(eng.device)
        return ids, kw

    def generate(self, prompts, sampling_params: SamplingParams | None = None, use_tqdm: bool = False):
        """`prompts`: list of {"prompt": str | "prompt_token_ids": [...], "multi_modal_data": {"video" | "image": data}}
        (VTRN:561-563).  Returns one RequestOutput per prompt, each with `n` CompletionOutputs whose `token_ids` end with
        the EOS token when the row finished (vLLM's convention), never padded."""
        sp = sampling_params or SamplingParams()
        if isinstance(prompts, (dict, str)):
            prompts = [prompts]
        eng = self.engine
        eos = list(sp.stop_token_ids or []) + (list(eng.dims.eos_ids) if eng.dims.eos_ids else [eng.dims.eos_id])
        eos = list(dict.fromkeys(int(e) for e in eos))[:4]
        out = []
        for i, item in enumerate(prompts):
            ids, kw = self._model_inputs(item)
            if self.max_model_len is not None and ids.shape[1] + sp.max_tokens > self.max_model_len:
                raise SpacerError(f"LLM.generate: prompt ({ids.shape[1]}) + max_tokens ({sp.max_tokens}) exceeds max_model_len")
            seed = (sp.seed if sp.seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())) + 104729 * (self._calls + i)
            full = eng.generate(ids, max_new_tokens=sp.max_tokens, num_return_sequences=sp.n, do_sample=sp.temperature > 0,
                                temperature=max(sp.temperature, 1e-6), top_p=sp.top_p, top_k=max(sp.top_k, 0),
                                repetition_penalty=sp.repetition_penalty, eos_token_id=eos, seed=seed,
                                min_new_tokens=sp.max_tokens if sp.ignore_eos else 0, **kw)
            comp = full[:, ids.shape[1]:].tolist()
            outs = []
            for j, row in enumerate(comp):
                cut, reason = len(row), "length"
                for t, tok in enumerate(row):
                    if tok in eos and not sp.ignore_eos:
                        cut, reason = t + 1, "stop"
                        break
                outs.append(CompletionOutput(index=j, token_ids=row[:cut], finish_reason=reason))
            out.append(RequestOutput(request_id=str(self._calls + i), prompt=item.get("prompt") if isinstance(item, dict) else item,
                                     prompt_token_ids=ids[0].tolist(), outputs=outs))
        self._calls += len(prompts)
        return out


def gather_generate_broadcast(llm: LLM | None, request: dict, sampling_params: SamplingParams, group=None,
                              main_rank: int = 0) -> list:
    """VTRN:546-608 over torch.distributed: every rank contributes ONE request; the main rank runs a single
    `llm.generate` over all of them; the flat list of completions (`[req0_gen0, ..., req0_gen{n-1}, req1_gen0, ...]`) is
    broadcast and each rank keeps its own slice.  Returns this rank's `n` token-id lists.  (`llm` may be None off the main
    rank.)  With world size 1 this is just `llm.generate`."""
    import torch.distributed as dist
    n = sampling_params.n
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [o.token_ids for o in llm.generate([request], sampling_params)[0].outputs]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, request, group=group)                      # gather_object(prompts_text / mm_data)
    flat = [None] * (world * n)
    if rank == main_rank:
        outs = llm.generate(gathered, sampling_params, use_tqdm=False)
        flat = [o.token_ids for r in outs for o in r.outputs]
    dist.broadcast_object_list(flat, src=dist.get_global_rank(group, main_rank) if group is not None else main_rank,
                               group=group)
    return flat[rank * n:(rank + 1) * n]


def pad_completions(token_lists: Sequence[Sequence[int]], pad_token_id: int, device=None) -> torch.Tensor:
    """`pad(completion_ids, padding_value=pad_token_id)` of VTRN:611-614: right-pad to the longest completion."""
    width = max((len(t) for t in token_lists), default=0)
    out = torch.full((len(token_lists), width), int(pad_token_id), dtype=torch.long, device=device)
    for i, t in enumerate(token_lists):
        out[i, :len(t)] = torch.as_tensor(list(t), dtype=torch.long, device=device)
    return out

"""ORACLE tooling (test infrastructure only -- never imported by the product path).

Runs the REFERENCE's own, unmodified `SGRLVRTrainer.compute_loss`
(/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py:384-686, "TRN") in this container:

  * the packages the reference imports but this image lacks (trl, accelerate, peft, qwen_vl_utils, deepspeed) are
    replaced by import stubs that provide exactly the names TRN binds (`maybe_apply_chat_template`, `is_conversational`,
    `unwrap_model_for_generation`, `process_vision_info`, ...);
  * the trainer object is created with `object.__new__` (its __init__ needs accelerate + a hub download) and given the
    attributes compute_loss reads; `transformers.Trainer._prepare_inputs` is the real one;
  * TRN:607-611 hard-codes `.to('cuda')`; there is no GPU here, so the module's `torch` global is replaced by a proxy whose
    `tensor()` results ignore `.to('cuda')` -- the function body itself is untouched;
  * the model the trainer drives is a stand-in backed by the fp32 oracle (oracle/qwen2vl_ref.py) that RECORDS every call
    (method, keyword names, shapes, dtypes, generation_config fields) and returns fixed completions from `generate`.

What comes out (oracle/make_trn_golden.py -> tests/golden/trn_compute_loss.pt): the call log -- the contract
spacer_b200/hf_api.py has to accept --, the loss and `_metrics` the reference computed, the inputs that produced them, and
the gradients `loss.backward()` left on the oracle's weights.  The GPU tests replay the log against the real engine.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from collections import defaultdict
from types import SimpleNamespace

import torch

REF_ROOT = "/root/reference/SpaceR-SG-RLVR/src/r1-v/src/open_r1"


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "trainer", "SG_RLVR_trainer.py"))


class _CpuTensor(torch.Tensor):
    """torch.tensor(...) results inside the reference module: `.to('cuda')` (TRN:607-611) is a no-op without a GPU."""

    def to(self, *args, **kwargs):
        if args and isinstance(args[0], str) and args[0].startswith("cuda") and not torch.cuda.is_available():
            return self
        return super().to(*args, **kwargs)


class _TorchProxy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def tensor(self, *args, **kwargs):
        return self._real.tensor(*args, **kwargs).as_subclass(_CpuTensor)


def _install_stubs():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    du = mod("trl.data_utils")
    mod("trl").data_utils = du
    du.is_conversational = lambda example: isinstance(example.get("prompt"), list)
    du.apply_chat_template = lambda example, tokenizer: {"prompt": tokenizer.apply_chat_template(example["prompt"])}
    du.maybe_apply_chat_template = lambda example, tokenizer: (
        du.apply_chat_template(example, tokenizer) if du.is_conversational(example) else example)
    tm = mod("trl.models")
    tm.create_reference_model = lambda model: model
    tm.prepare_deepspeed = lambda model, accelerator: model

    @contextlib.contextmanager
    def unwrap_model_for_generation(model, accelerator, **kw):
        yield model
    tm.unwrap_model_for_generation = unwrap_model_for_generation
    mod("trl.trainer")
    mod("trl.trainer.grpo_config").GRPOConfig = type("GRPOConfig", (), {})
    tu = mod("trl.trainer.utils")
    tu.generate_model_card = tu.get_comet_experiment_url = None
    q = mod("qwen_vl_utils")
    q.process_vision_info = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("patched per run"))


def load_reference_trainer():
    """The reference's trainer module, imported from where it lies (never copied)."""
    _install_stubs()
    spec = importlib.util.spec_from_file_location("sg_rlvr_trainer_ref", os.path.join(REF_ROOT, "trainer", "SG_RLVR_trainer.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.torch = _TorchProxy(torch)
    return m


# ------------------------------------------------------------------------------------------------------------------
def describe(v):
    if isinstance(v, torch.Tensor):
        return {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")}
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    if isinstance(v, (list, tuple)):
        return [describe(x) for x in v]
    return type(v).__name__


GEN_FIELDS = ("max_new_tokens", "do_sample", "top_p", "top_k", "temperature", "num_return_sequences", "pad_token_id",
              "eos_token_id", "repetition_penalty", "num_beams")


class RecordingOracleModel(torch.nn.Module):
    """Stand-in for `Qwen2VLForConditionalGeneration` backed by the fp32 oracle; records the trainer's calls."""

    def __init__(self, R, dims, weights, completions, shuffled_completions, log, name):
        super().__init__()
        self.R, self.d, self.log, self.name = R, dims, log, name
        self.w = {k: torch.nn.Parameter(v.clone()) for k, v in weights.items()}
        self._plist = torch.nn.ParameterList(self.w.values())
        self.fixed = {completions.shape[0]: completions}
        if shuffled_completions is not None:
            self.fixed[shuffled_completions.shape[0]] = shuffled_completions
        self.config = SimpleNamespace(_name_or_path="stub/Qwen2-VL-tiny")
        self.warnings_issued = {}

    def generate(self, **kw):
        gc = kw.get("generation_config")
        rec = {"model": self.name, "call": "generate",
               "kwargs": {k: describe(v) for k, v in kw.items() if k != "generation_config"},
               "generation_config": {f: getattr(gc, f, None) for f in GEN_FIELDS}}
        self.log.append(rec)
        G = gc.num_return_sequences
        ids = kw["input_ids"]
        if gc.max_new_tokens == 1:                     # the dummy call for image samples under --temporal (TRN:481)
            return torch.cat([ids.repeat(G, 1), torch.full((G, 1), self.d.eos_id)], dim=1)
        comp = self.fixed[G]
        return torch.cat([ids.repeat(G, 1), comp], dim=1)

    def forward(self, input_ids, **kw):
        R, d = self.R, self.d
        self.log.append({"model": self.name, "call": "forward", "grad_enabled": torch.is_grad_enabled(),
                         "inference_mode": torch.is_inference_mode_enabled(),
                         "args": [describe(input_ids)], "kwargs": {k: describe(v) for k, v in kw.items()}})
        pix = kw.get("pixel_values_videos", kw.get("pixel_values"))
        grid = kw.get("video_grid_thw", kw.get("image_grid_thw"))
        pos = R.rope_index_classic(input_ids, grid, d)
        return SimpleNamespace(logits=R.model_logits(self.w, input_ids, pix, grid, pos, d))


class FakeProcessor:
    """What `AutoProcessor.from_pretrained(...)` gives the trainer, reduced to the members compute_loss touches
    (TRN:390-425, 555-560): chat template, the call that returns the model inputs, batch_decode, eos/pad ids."""

    def __init__(self, dims, prompt_ids, patchify, decode, visual_kind="video", second_per_grid_ts=None):
        self.d, self.prompt_ids, self.patchify, self.decode = dims, prompt_ids, patchify, decode
        self.kind, self.sec = visual_kind, second_per_grid_ts
        self.eos_token_id, self.pad_token_id = dims.eos_id, dims.pad_id
        self.calls = []

    def apply_chat_template(self, conversation, **kw):
        return "<|im_start|>user\n<|vision_start|><|video_pad|><|vision_end|>" + conversation[0]["content"][-1]["text"] + \
               "<|im_end|>\n<|im_start|>assistant\n"

    def __call__(self, text=None, images=None, videos=None, **kw):
        self.calls.append({"text": len(text), "images": describe(images), "videos": describe(videos), **kw})
        out = {"input_ids": self.prompt_ids.clone(), "attention_mask": torch.ones_like(self.prompt_ids)}
        if self.kind == "video":
            pix, grid = self.patchify(videos[0])
            out["pixel_values_videos"], out["video_grid_thw"] = pix, grid
            if self.sec is not None:
                out["second_per_grid_ts"] = list(self.sec)
        else:
            pix, grid = self.patchify(images[0])
            out["pixel_values"], out["image_grid_thw"] = pix, grid
        return out

    def batch_decode(self, ids, skip_special_tokens=True):
        return self.decode(ids)


def make_trainer(mod, model, ref_model, processor, reward_funcs, G, C, beta, temporal, len_control, frames, kind):
    """An SGRLVRTrainer with the attributes compute_loss reads, without running its __init__ (hub + accelerate)."""
    from transformers import GenerationConfig
    tr = object.__new__(mod.SGRLVRTrainer)
    tr.processing_class = processor
    tr.reward_funcs = list(reward_funcs)
    tr.reward_processing_classes = [None] * len(reward_funcs)
    tr.max_prompt_length, tr.max_completion_length, tr.num_generations = 16384, C, G
    tr.temporal, tr.len_control, tr.beta = temporal, len_control, beta
    pad = processor.pad_token_id

    def gc(n, g):                                           # TRN:277-302, same literal arguments
        return GenerationConfig(max_new_tokens=n, do_sample=True, top_p=0.95, temperature=1, num_return_sequences=g,
                                pad_token_id=pad)
    tr.generation_config = gc(C, G)
    tr.shuffled_num_generations = G // 2
    tr.shuffled_generation_config = gc(C, G // 2)
    tr.dummy_generation_config = gc(1, 1)
    tr.ref_model = ref_model
    tr._metrics = defaultdict(list)
    tr.args = SimpleNamespace(device=torch.device("cpu"), past_index=-1)
    tr.is_deepspeed_enabled = False
    tr.accelerator = SimpleNamespace(device=torch.device("cpu"), gather_for_metrics=lambda x: x,
                                     unwrap_model=lambda m: m)

    def process_vision_info(conversation, return_video_kwargs=False):
        if kind == "video":
            return None, [frames.clone()], {"fps": [2.0]}
        return [frames.clone()], None, {}
    mod.process_vision_info = process_vision_info
    return tr


def run_compute_loss(tr, model, example, seed):
    """One unmodified compute_loss + backward.  Returns (loss tensor, metrics dict, captured stdout)."""
    torch.manual_seed(seed)                                  # TRN:443 draws the frame permutation from the global RNG
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        loss = tr.compute_loss(model, [example])
    return loss, {k: list(v) for k, v in tr._metrics.items()}, buf.getvalue()

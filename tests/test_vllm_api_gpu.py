"""The vLLM engine surface of `Qwen2VLGRPOVLLMTrainerModified` (vllm_grpo_trainer_modified.py:361-387, 526-608) on the
colocated B200 engine: weight sync, `LLM.generate(inputs, sampling_params)` with frames as multi_modal_data, outputs in
vLLM's shape."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _setup():
    from oracle import qwen2vl_ref as R
    from spacer_b200 import config
    from spacer_b200.hf_api import Qwen2VLForConditionalGenerationB200
    d_or, d = R.dims_tiny(2, 2), config.tiny(2, 2)
    w = R.init_weights(d_or, seed=0)
    model = Qwen2VLForConditionalGenerationB200.from_dims(d, "cuda")
    model.load_state_dict(w)
    return R, d_or, d, model, w


def test_weight_sync_is_free_when_colocated_and_copies_otherwise():
    from spacer_b200.vllm_api import LLM
    R, d_or, d, model, w = _setup()
    llm = LLM(engine=model.engine)
    handle = llm.llm_engine.model_executor.driver_worker.model_runner.model
    before = model.engine.params.mat.clone()
    version = model.engine.params.version
    # VTRN:531-543: state_dict = unwrapped_model.state_dict(); llm_model.load_weights(state_dict.items())
    state_dict = model.state_dict()
    loaded = handle.load_weights(state_dict.items())
    assert len(loaded) == len(state_dict) and handle.loads == 0                   # nothing moved
    assert model.engine.params.version == version and torch.equal(model.engine.params.mat, before)
    # a foreign checkpoint (older `visual.*` / `model.layers.*` names, fp32, CPU) is copied in
    w2 = R.init_weights(d_or, seed=3)
    old_names = {}
    for k, v in w2.items():
        k = k.replace("model.visual.", "visual.").replace("model.language_model.", "model.")
        old_names[k] = v
    handle.load_weights(old_names.items())
    assert handle.loads == 1 and not torch.equal(model.engine.params.mat, before)
    assert torch.equal(model.engine.state_dict()["lm_head.weight"].cpu(), w2["lm_head.weight"].bfloat16())


def test_llm_generate_matches_engine_generate_and_vllm_output_shape():
    from oracle.make_golden import tiny_case
    from oracle.vision_ref import patchify_ref
    from spacer_b200.vllm_api import LLM, SamplingParams
    R, d_or, d, model, w = _setup()
    case = tiny_case(d_or)
    llm = LLM(engine=model.engine, max_model_len=4096)
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (4, 3, 112, 112), generator=g, dtype=torch.uint8)   # grid (2, 8, 8) like tiny_case
    prompt_ids = case["prompt_ids"][0].tolist()
    sp = SamplingParams(temperature=1.0, top_p=0.95, max_tokens=10, n=4, seed=11)       # VTRN:383-387 + n (:569-571)
    outs = llm.generate([{"prompt_token_ids": prompt_ids, "multi_modal_data": {"video": frames}}], sampling_params=sp,
                        use_tqdm=False)
    assert len(outs) == 1 and len(outs[0].outputs) == 4 and outs[0].prompt_token_ids == prompt_ids
    for o in outs[0].outputs:
        assert 1 <= len(o.token_ids) <= 10
        if d.eos_id in o.token_ids:                       # vLLM: token_ids end WITH the eos token, nothing after it
            assert o.token_ids.index(d.eos_id) == len(o.token_ids) - 1 and o.finish_reason == "stop"
        else:
            assert len(o.token_ids) == 10 and o.finish_reason == "length"
    # same request through the engine directly (same seed derivation: first call of this LLM -> seed + 0)
    pix, grid = patchify_ref(frames)
    ref = model.engine.generate(case["prompt_ids"], pix.cuda(), torch.tensor([list(grid)]), max_new_tokens=10,
                                num_return_sequences=4, do_sample=True, temperature=1.0, top_p=0.95, top_k=0, seed=11)
    P = len(prompt_ids)
    for j, o in enumerate(outs[0].outputs):
        assert ref[j, P:P + len(o.token_ids)].tolist() == o.token_ids
    # greedy (temperature 0) is deterministic and identical across the n samples; two prompts in one call keep their order
    g0 = llm.generate([{"prompt_token_ids": prompt_ids, "multi_modal_data": {"video": frames}},
                       {"prompt_token_ids": prompt_ids[3:], "multi_modal_data": {"video": frames}}],
                      SamplingParams(temperature=0.0, max_tokens=6, n=2))
    assert len(g0) == 2 and g0[0].outputs[0].token_ids == g0[0].outputs[1].token_ids
    assert g0[1].prompt_token_ids == prompt_ids[3:]
    from spacer_b200.ops import SpacerError
    with pytest.raises(SpacerError):
        LLM(engine=model.engine, max_model_len=8).generate([{"prompt_token_ids": prompt_ids}], SamplingParams(max_tokens=4))
    with pytest.raises(SpacerError):
        llm.generate([{"prompt": "needs a processor"}], SamplingParams(max_tokens=2))


def test_llm_generate_with_the_trainers_processor():
    """Text prompts go through the HF processor the trainer holds (`processing_class`): a stand-in with the same call
    signature (VTRN:484-492) feeds ids + pixel values."""
    from oracle.make_golden import tiny_case
    from oracle.vision_ref import patchify_ref
    from spacer_b200.vllm_api import LLM, SamplingParams
    R, d_or, d, model, w = _setup()
    case = tiny_case(d_or)
    calls = []

    def processor(text=None, images=None, videos=None, return_tensors="pt", padding=True, padding_side="left",
                  add_special_tokens=False):
        calls.append((text, None if videos is None else tuple(videos[0].shape)))
        pix, grid = patchify_ref(videos[0])
        return {"input_ids": case["prompt_ids"].clone(), "attention_mask": torch.ones_like(case["prompt_ids"]),
                "pixel_values_videos": pix, "video_grid_thw": torch.tensor([list(grid)])}
    frames = torch.randint(0, 256, (4, 3, 112, 112), generator=torch.Generator().manual_seed(5)).float()
    llm = LLM(engine=model.engine, processor=processor)
    outs = llm.generate([{"prompt": "<|im_start|>user ...", "multi_modal_data": {"video": [frames]}}],
                        SamplingParams(temperature=1.0, top_p=0.95, max_tokens=5, n=3))
    assert calls == [(["<|im_start|>user ..."], (4, 3, 112, 112))]
    assert len(outs[0].outputs) == 3 and outs[0].prompt_token_ids == case["prompt_ids"][0].tolist()

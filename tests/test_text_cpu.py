"""Row a3 of SURVEY.md 8(a): chat template + tokenisation (SG_RLVR_trainer.py:390-425).  The product's byte-level BPE is
compared token by token with the real `transformers.Qwen2Tokenizer` (Rust `tokenizers` backend) built on the SAME
vocabulary -- a small BPE trained here with the Qwen2 pre-tokeniser, since no released vocabulary is available offline --
and the chat template with the released Qwen2-VL Jinja template rendered by jinja2."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CORPUS = [
    "The table is left of the chair. Which object is closest to the window?",
    "<think>Looking at the frames, I count 3 chairs and 12 tables.</think><answer>B</answer>",
    "{'table': [[1,3],[5,6]], 'chair': [[9,4]], 'window': [[6,5]]}",
    "These are frames of a video. Question: how many sofa(s) are in this room?\nOptions:\nA. 1\nB. 2\nC. 3\nD. 4",
    "It's 10:30pm -- we're done, they'll've left; I'd say don't.",
    "多模态大模型在空间推理任务上的表现。物体的相对位置和距离估计。",
    "Ünïcödé tëxt with àccénts, emoji 🙂🚀 and\ttabs\n\n\nnewlines   spaces  ",
    "def f(x):\n    return x ** 2 + 3.14159 * x - 42  # comment\n",
    "1234567890 3.5m 12cm 0.25 1e-6 100%",
    "Please answer with a single letter. Please think about this question as if you were a human pondering deeply.",
] * 3

# the chat template of the released Qwen2-VL-7B-Instruct (chat_template.json), restated here as the yardstick
QWEN2_VL_TEMPLATE = (
    "{% set image_count = namespace(value=0) %}{% set video_count = namespace(value=0) %}{% for message in messages %}"
    "{% if loop.first and message['role'] != 'system' %}<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n{% endif %}"
    "<|im_start|>{{ message['role'] }}\n{% if message['content'] is string %}{{ message['content'] }}<|im_end|>\n{% else %}"
    "{% for content in message['content'] %}{% if content['type'] == 'image' or 'image' in content or 'image_url' in content %}"
    "{% set image_count.value = image_count.value + 1 %}{% if add_vision_id %}Picture {{ image_count.value }}: {% endif %}"
    "<|vision_start|><|image_pad|><|vision_end|>{% elif content['type'] == 'video' or 'video' in content %}"
    "{% set video_count.value = video_count.value + 1 %}{% if add_vision_id %}Video {{ video_count.value }}: {% endif %}"
    "<|vision_start|><|video_pad|><|vision_end|>{% elif 'text' in content %}{{ content['text'] }}{% endif %}{% endfor %}"
    "<|im_end|>\n{% endif %}{% endfor %}{% if add_generation_prompt %}<|im_start|>assistant\n{% endif %}")


def build_tok_dir(tmp_path_factory):
    """A Qwen2-style tokenizer directory: BPE trained with the Qwen2 pre-tokeniser, Qwen2-VL's added tokens after it."""
    from tokenizers import Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, trainers
    from spacer_b200.text import PRETOKENIZE_REGEX, QWEN2_VL_SPECIAL_TOKENS
    t = Tokenizer(models.BPE(unk_token=None, fuse_unk=False, byte_fallback=False))
    t.normalizer = normalizers.NFC()
    t.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PRETOKENIZE_REGEX), behavior="isolated"),
                                               pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    t.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=700, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    t.train_from_iterator(CORPUS, tr)
    d = tmp_path_factory.mktemp("tok")
    t.model.save(str(d))                                   # vocab.json + merges.txt
    with open(d / "vocab.json") as f:
        n = len(json.load(f))
    # the added tokens get the ids the real HF tokenizer assigns them (what a saved tokenizer_config.json records)
    hf = _hf_tokenizer(str(d))
    cfg = {"added_tokens_decoder": {str(i): {"content": a.content, "special": a.special} for i, a in hf.added_tokens_decoder.items()},
           "eos_token": "<|im_end|>", "pad_token": "<|endoftext|>"}
    assert {a["content"] for a in cfg["added_tokens_decoder"].values()} >= set(QWEN2_VL_SPECIAL_TOKENS)
    with open(d / "tokenizer_config.json", "w") as f:
        json.dump(cfg, f)
    return str(d), n


@pytest.fixture(scope="module")
def tok_dir(tmp_path_factory):
    return build_tok_dir(tmp_path_factory)


def _hf_tokenizer(path):
    from transformers import Qwen2Tokenizer
    from spacer_b200.text import QWEN2_VL_SPECIAL_TOKENS
    with open(os.path.join(path, "vocab.json")) as f:
        vocab = json.load(f)
    with open(os.path.join(path, "merges.txt"), encoding="utf-8") as f:
        merges = [tuple(ln.split()) for ln in f.read().splitlines() if ln and not ln.startswith("#version")]
    hf = Qwen2Tokenizer(vocab=vocab, merges=merges, eos_token="<|im_end|>", pad_token="<|endoftext|>",
                        additional_special_tokens=[t for t in QWEN2_VL_SPECIAL_TOKENS if t not in ("<|im_end|>", "<|endoftext|>")])
    return hf


def test_bpe_matches_transformers_qwen2_tokenizer(tok_dir):
    from spacer_b200.text import QWEN2_VL_SPECIAL_TOKENS, Qwen2Tokenizer
    path, n = tok_dir
    mine = Qwen2Tokenizer.from_pretrained(path)
    hf = _hf_tokenizer(path)
    # the added tokens got the same ids in both (appended after the BPE vocabulary)
    for t in QWEN2_VL_SPECIAL_TOKENS:
        assert mine.vocab[t] == hf.convert_tokens_to_ids(t), t
    texts = CORPUS[:10] + [
        "", " ", "  leading and trailing  ", "a", "\n", "x\r\ny", "naïve café — “quotes” … ½ ×",
        "<|im_start|>user\n<|vision_start|><|video_pad|><|video_pad|><|vision_end|>How many chairs?<|im_end|>\n<|im_start|>assistant\n",
        "text<|im_end|>more<|endoftext|>", "unseen bytes: \x00\x7f ꙮ 𝔘𝔫𝔦 ﷽", "ＦＵＬＬwidth ｶﾀｶﾅ é (combining, NFC)",
        "TABLE Table tAbLe 'S 'LL DON'T"]
    for s in texts:
        want = hf(s, add_special_tokens=False)["input_ids"]
        got = mine.encode(s)
        assert got == want, (s, got[:20], want[:20])
        assert mine.decode(got) == hf.decode(want, skip_special_tokens=False, clean_up_tokenization_spaces=False), s
        assert mine.decode(got, skip_special_tokens=True) == hf.decode(want, skip_special_tokens=True,
                                                                      clean_up_tokenization_spaces=False), s
    # batch call: left padding, attention mask (TRN:417-425)
    b = mine(["short", "a much longer sentence about tables and chairs"], padding=True, padding_side="left")
    assert b["input_ids"].shape == b["attention_mask"].shape and b["attention_mask"][0, 0] == 0 and b["attention_mask"][1].all()
    assert b["input_ids"][0, 0] == mine.pad_token_id
    hb = hf(["short", "a much longer sentence about tables and chairs"], padding=True, padding_side="left", return_tensors="pt",
            add_special_tokens=False)
    assert torch.equal(b["input_ids"], hb["input_ids"]) and torch.equal(b["attention_mask"], hb["attention_mask"])
    assert mine.batch_decode(b["input_ids"], skip_special_tokens=True) == ["short", "a much longer sentence about tables and chairs"]


CONVS = [
    [{"role": "user", "content": [{"type": "video", "video": "file:///x.mp4"}, {"type": "text", "text": "How many chairs?"}]}],
    [{"role": "system", "content": "You are a spatial reasoner."},
     {"role": "user", "content": [{"type": "image", "image": "a.jpg"}, {"type": "text", "text": "Describe."}]},
     {"role": "assistant", "content": "A room."}, {"role": "user", "content": "And the table?"}],
    [{"role": "user", "content": [{"video": "v.mp4"}, {"text": "no type keys"}, {"type": "image_url", "image_url": "u"}]}],
    [{"role": "user", "content": "plain string"}],
]


@pytest.mark.parametrize("conv", CONVS)
@pytest.mark.parametrize("gen,vid", [(True, False), (False, False), (True, True)])
def test_chat_template_matches_the_released_jinja_template(conv, gen, vid):
    import jinja2
    from spacer_b200.text import apply_chat_template
    want = jinja2.Environment().from_string(QWEN2_VL_TEMPLATE).render(messages=conv, add_generation_prompt=gen, add_vision_id=vid)
    assert apply_chat_template(conv, add_generation_prompt=gen, add_vision_id=vid) == want
    # a checkpoint-provided template is rendered as is
    assert apply_chat_template(conv, add_generation_prompt=gen, add_vision_id=vid, template=QWEN2_VL_TEMPLATE) == want


def test_reference_prompt_conversation_renders(tok_dir):
    """The conversation SG-RLVR.py builds (open_r1/SG-RLVR.py:322-343 make_conversation_image_and_video) through
    maybe_apply_chat_template -> processing_class.apply_chat_template, exactly as TRN:392 does."""
    from spacer_b200 import config
    from spacer_b200.text import Qwen2Tokenizer, Qwen2VLProcessorB200
    path, n = tok_dir
    tok = Qwen2Tokenizer.from_pretrained(path)
    d = config.tiny()
    from dataclasses import replace
    d = replace(d, image_token_id=tok.vocab["<|image_pad|>"], video_token_id=tok.vocab["<|video_pad|>"],
                vision_start_id=tok.vocab["<|vision_start|>"], vision_end_id=tok.vocab["<|vision_end|>"],
                eos_id=tok.vocab["<|im_end|>"], pad_id=tok.vocab["<|endoftext|>"], vocab=n + 14)
    proc = Qwen2VLProcessorB200(tok, d, device="cpu")
    conv = [{"role": "user", "content": [{"type": "video"}, {"type": "text", "text": "Q: which is closer? Options:\nA. x\nB. y"}]}]
    text = proc.apply_chat_template(conv, tokenize=False, add_generation_prompt=True)
    assert text.startswith("<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n<|im_start|>user\n<|vision_start|><|video_pad|>")
    assert text.endswith("<|im_end|>\n<|im_start|>assistant\n")
    assert proc.pad_token_id == d.pad_id and proc.eos_token_id == d.eos_id
    ids = proc.apply_chat_template(conv, tokenize=True)
    assert ids.count(d.video_token_id) == 1 and tok.decode(ids) == text
    assert proc.batch_decode([ids], skip_special_tokens=True) == [
        "system\nYou are a helpful assistant.\nuser\nQ: which is closer? Options:\nA. x\nB. y\nassistant\n"]

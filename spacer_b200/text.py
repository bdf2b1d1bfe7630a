"""Prompt side of `SGRLVRTrainer.compute_loss` (SG_RLVR_trainer.py:390-425): chat template,
tokenisation, placeholder expansion, left padding -- what `processing_class.apply_chat_template(...)` and
`processing_class(text=..., videos=..., return_tensors="pt", padding=True, padding_side="left",
add_special_tokens=False)` do in the reference (transformers `Qwen2VLProcessor` + the Rust `tokenizers` BPE), plus
`batch_decode` for the reward functions (TRN:555-560).

* `Qwen2Tokenizer`: byte-level BPE exactly as transformers defines the Qwen2 tokenizer (models/qwen2/tokenization_qwen2.py:
  NFC normaliser -> Split(PRETOKENIZE_REGEX, isolated) -> ByteLevel(no prefix space) -> BPE(no unk, no byte fallback),
  ByteLevel decoder; added/special tokens matched first).  Loads `tokenizer.json` or `vocab.json` + `merges.txt` from a local
  checkpoint directory.  CPU tests compare it token by token with the real `transformers.Qwen2Tokenizer` on the same vocabulary.
* `apply_chat_template`: the checkpoint's own Jinja template when the directory has one (chat_template.json /
  tokenizer_config.json, rendered with jinja2), else a Python restatement of the released Qwen2-VL / Qwen2.5-VL template.
* `Qwen2VLProcessorB200`: text + frames -> {input_ids, attention_mask, pixel_values_videos, video_grid_thw
  [, second_per_grid_ts]}; the pixel side is the GPU front-end of vision.py (fp32 output bit-exact with the HF processor).

This is host code and not hot (negligible next to a 2 s step); Python like the reference's side of it.
"""
from __future__ import annotations

import json
import os
import unicodedata
from functools import lru_cache
from types import SimpleNamespace

import torch

from .ops import SpacerError

# transformers/models/qwen2/tokenization_qwen2.py:33
PRETOKENIZE_REGEX = (r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+"""
                     r"""|\s+(?!\S)|\s+""")

DEFAULT_SYSTEM = "You are a helpful assistant."
VISION_START, VISION_END, IMAGE_PAD, VIDEO_PAD = "<|vision_start|>", "<|vision_end|>", "<|image_pad|>", "<|video_pad|>"
# the added tokens of the released Qwen2-VL / Qwen2.5-VL tokenizers, in id order from 151643
QWEN2_VL_SPECIAL_TOKENS = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "<|object_ref_start|>", "<|object_ref_end|>",
                           "<|box_start|>", "<|box_end|>", "<|quad_start|>", "<|quad_end|>", VISION_START, VISION_END,
                           "<|vision_pad|>", IMAGE_PAD, VIDEO_PAD]


@lru_cache(maxsize=1)
def bytes_to_unicode() -> dict:
    """GPT-2's reversible byte <-> printable-unicode map (what `pre_tokenizers.ByteLevel` applies)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, map(chr, cs)))


class Qwen2Tokenizer:
    """Byte-level BPE with the Qwen2 pre-tokeniser.  `vocab`: token string -> id; `merges`: list of (left, right) in rank
    order; `added_tokens`: content -> (id, special) matched verbatim before anything else."""

    def __init__(self, vocab: dict, merges, added_tokens: dict | None = None, eos_token="<|im_end|>",
                 pad_token="<|endoftext|>"):
        import regex
        self.vocab = dict(vocab)
        self.ranks = {tuple(m.split(" ")) if isinstance(m, str) else tuple(m): i for i, m in enumerate(merges)}
        self.added = {k: (int(v[0]), bool(v[1])) if isinstance(v, (tuple, list)) else (int(v), True)
                      for k, v in (added_tokens or {}).items()}
        for k, (i, _) in self.added.items():
            self.vocab.setdefault(k, i)
        self.id_to_token = {i: t for t, i in self.vocab.items()}
        self.special_ids = {i for _, (i, sp) in self.added.items() if sp}
        self._pat = regex.compile(PRETOKENIZE_REGEX)
        self._added_pat = (regex.compile("|".join(regex.escape(k) for k in sorted(self.added, key=len, reverse=True)))
                           if self.added else None)
        self._b2u = bytes_to_unicode()
        self._u2b = {u: b for b, u in self._b2u.items()}
        self._cache = {}
        self.eos_token, self.pad_token = eos_token, pad_token
        self.eos_token_id = self.vocab.get(eos_token)
        self.pad_token_id = self.vocab.get(pad_token)
        self.padding_side = "left"

    # ---- construction ------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path: str):
        """Local HF tokenizer files: tokenizer.json, or vocab.json + merges.txt (+ tokenizer_config.json / added_tokens.json)."""
        tj = os.path.join(path, "tokenizer.json")
        cfg = {}
        if os.path.exists(os.path.join(path, "tokenizer_config.json")):
            with open(os.path.join(path, "tokenizer_config.json")) as f:
                cfg = json.load(f)
        added = {}
        if os.path.exists(tj):
            with open(tj) as f:
                t = json.load(f)
            if t["model"]["type"] != "BPE":
                raise SpacerError("tokenizer.json does not hold a BPE model")
            vocab, merges = t["model"]["vocab"], t["model"]["merges"]
            for a in t.get("added_tokens", []):
                added[a["content"]] = (a["id"], a.get("special", False))
        else:
            vj, mt = os.path.join(path, "vocab.json"), os.path.join(path, "merges.txt")
            if not (os.path.exists(vj) and os.path.exists(mt)):
                raise SpacerError(f"no tokenizer.json or vocab.json + merges.txt under {path}")
            with open(vj) as f:
                vocab = json.load(f)
            with open(mt, encoding="utf-8") as f:
                merges = [ln.rstrip("\n") for ln in f if ln.strip() and not ln.startswith("#version")]
            for i, a in (cfg.get("added_tokens_decoder") or {}).items():
                added[a["content"]] = (int(i), a.get("special", False))
            aj = os.path.join(path, "added_tokens.json")
            if os.path.exists(aj):
                with open(aj) as f:
                    for k, i in json.load(f).items():
                        added.setdefault(k, (i, True))

        def tok(x, default):
            x = cfg.get(x, default)
            return x["content"] if isinstance(x, dict) else (x or default)
        return cls(vocab, merges, added, eos_token=tok("eos_token", "<|im_end|>"), pad_token=tok("pad_token", "<|endoftext|>"))

    # ---- encode ------------------------------------------------------------------------------------------------
    def _bpe(self, word: str) -> list:
        """Greedy lowest-rank merges over the byte-level characters of one pre-token."""
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts = list(word)
        while len(parts) > 1:
            best, bi = None, -1
            for i in range(len(parts) - 1):
                r = self.ranks.get((parts[i], parts[i + 1]))
                if r is not None and (best is None or r < best):
                    best, bi = r, i
            if best is None:
                break
            a, b = parts[bi], parts[bi + 1]
            out, i = [], 0
            while i < len(parts):               # merge every occurrence of the pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(parts[i])
                    i += 1
            parts = out
        if len(self._cache) < 1 << 16:
            self._cache[word] = parts
        return parts

    def _encode_plain(self, text: str) -> list:
        ids = []
        text = unicodedata.normalize("NFC", text)
        for m in self._pat.finditer(text):
            word = "".join(self._b2u[b] for b in m.group(0).encode("utf-8"))
            for t in self._bpe(word):
                i = self.vocab.get(t)
                if i is None:
                    raise SpacerError(f"token {t!r} is not in the vocabulary (byte-level BPE vocabularies hold all 256 bytes)")
                ids.append(i)
        return ids

    def encode(self, text: str, add_special_tokens: bool = False) -> list:
        """Token ids of `text`; added tokens (<|im_start|>, <|video_pad|>, ...) are matched verbatim first."""
        if self._added_pat is None:
            return self._encode_plain(text)
        ids, pos = [], 0
        for m in self._added_pat.finditer(text):
            if m.start() > pos:
                ids += self._encode_plain(text[pos:m.start()])
            ids.append(self.added[m.group(0)][0])
            pos = m.end()
        if pos < len(text):
            ids += self._encode_plain(text[pos:])
        return ids

    def __call__(self, text, padding=True, padding_side=None, return_tensors="pt", add_special_tokens=False, **unused):
        texts = [text] if isinstance(text, str) else list(text)
        enc = [self.encode(t) for t in texts]
        return pad_batch(enc, self.pad_token_id, padding_side or self.padding_side)

    # ---- decode ------------------------------------------------------------------------------------------------
    def decode(self, ids, skip_special_tokens: bool = False, **unused) -> str:
        if torch.is_tensor(ids):
            ids = ids.tolist()
        out, buf = [], bytearray()

        def flush():
            if buf:
                out.append(buf.decode("utf-8", errors="replace"))
                buf.clear()
        for i in ids:
            i = int(i)
            t = self.id_to_token.get(i)
            if t is None:
                continue
            if t in self.added:                  # added tokens are emitted verbatim, not byte-decoded
                if skip_special_tokens and i in self.special_ids:
                    continue
                flush()
                out.append(t)
            else:
                buf.extend(self._u2b[c] for c in t)
        flush()
        return "".join(out)

    def batch_decode(self, sequences, skip_special_tokens: bool = False, **unused) -> list:
        return [self.decode(s, skip_special_tokens=skip_special_tokens) for s in sequences]


def pad_batch(rows, pad_id: int, side: str = "left") -> dict:
    """{input_ids, attention_mask} LongTensors [B, max_len], padded on `side` (the reference pads left, TRN:417-425)."""
    width = max(len(r) for r in rows)
    ids = torch.full((len(rows), width), int(pad_id), dtype=torch.long)
    mask = torch.zeros((len(rows), width), dtype=torch.long)
    for b, r in enumerate(rows):
        if not r:
            continue
        t = torch.tensor(r, dtype=torch.long)
        if side == "left":
            ids[b, width - len(r):], mask[b, width - len(r):] = t, 1
        else:
            ids[b, :len(r)], mask[b, :len(r)] = t, 1
    return {"input_ids": ids, "attention_mask": mask}


# ----------------------------------------------------------------------------------------------------------------------
def apply_chat_template(conversation, add_generation_prompt: bool = True, add_vision_id: bool = False,
                        template: str | None = None, **kw) -> str:
    """The Qwen2-VL / Qwen2.5-VL chat template (ChatML): a default system turn when the first message is not one, every
    image / video content item replaced by <|vision_start|><|image_pad|/|video_pad|><|vision_end|>, `<|im_start|>assistant\\n`
    appended for generation.  `template`: a Jinja template (the checkpoint's own) to render instead."""
    if template is not None:
        import jinja2
        env = jinja2.Environment(trim_blocks=True, lstrip_blocks=True, extensions=["jinja2.ext.loopcontrols"])
        return env.from_string(template).render(messages=conversation, add_generation_prompt=add_generation_prompt,
                                                add_vision_id=add_vision_id, **kw)
    out = []
    n_img = n_vid = 0
    for k, msg in enumerate(conversation):
        if k == 0 and msg["role"] != "system":
            out.append(f"<|im_start|>system\n{DEFAULT_SYSTEM}<|im_end|>\n")
        out.append(f"<|im_start|>{msg['role']}\n")
        content = msg["content"]
        if isinstance(content, str):
            out.append(content)
        else:
            for c in content:
                if c.get("type") == "image" or "image" in c or "image_url" in c:
                    n_img += 1
                    if add_vision_id:
                        out.append(f"Picture {n_img}: ")
                    out.append(VISION_START + IMAGE_PAD + VISION_END)
                elif c.get("type") == "video" or "video" in c:
                    n_vid += 1
                    if add_vision_id:
                        out.append(f"Video {n_vid}: ")
                    out.append(VISION_START + VIDEO_PAD + VISION_END)
                elif "text" in c:
                    out.append(c["text"])
        out.append("<|im_end|>\n")
    if add_generation_prompt:
        out.append("<|im_start|>assistant\n")
    return "".join(out)


class Qwen2VLProcessorB200:
    """What the trainer holds as `processing_class` (TRN:226-236): chat template, the text + vision call, batch_decode,
    pad / eos ids, and the `image_processor.max_pixels / min_pixels` knobs the trainer sets."""

    def __init__(self, tokenizer: Qwen2Tokenizer, dims, chat_template: str | None = None, device="cuda"):
        self.tokenizer, self.dims, self.chat_template, self.device = tokenizer, dims, chat_template, device
        self.pad_token_id, self.eos_token_id = tokenizer.pad_token_id, tokenizer.eos_token_id
        self.image_processor = SimpleNamespace(max_pixels=12845056, min_pixels=3136)
        for name, want in ((IMAGE_PAD, dims.image_token_id), (VIDEO_PAD, dims.video_token_id)):
            got = tokenizer.vocab.get(name)
            if got is not None and got != want:
                raise SpacerError(f"tokenizer maps {name} to {got}, the model config expects {want}")

    @classmethod
    def from_pretrained(cls, path: str, dims=None, device="cuda", **unused):
        from . import hub
        tok = Qwen2Tokenizer.from_pretrained(path)
        if dims is None:
            with open(os.path.join(path, "config.json")) as f:
                dims = hub.dims_from_hf_config(json.load(f), name_hint=path)
        template = None
        for fn, key in (("chat_template.json", "chat_template"), ("tokenizer_config.json", "chat_template")):
            p = os.path.join(path, fn)
            if template is None and os.path.exists(p):
                with open(p) as f:
                    template = json.load(f).get(key)
        return cls(tok, dims, template, device)

    def apply_chat_template(self, conversation, tokenize: bool = False, add_generation_prompt: bool = True, **kw):
        if conversation and isinstance(conversation[0], list):          # a batch of conversations
            return [self.apply_chat_template(c, tokenize, add_generation_prompt, **kw) for c in conversation]
        text = apply_chat_template(conversation, add_generation_prompt, template=self.chat_template, **kw)
        return self.tokenizer.encode(text) if tokenize else text

    def _visual(self, items, is_video: bool):
        """frames of each item -> (pixel rows fp32, grids, placeholder counts).  Frames arrive decoded and resized
        (qwen-vl-utils' fetch_video / vision.fetch_video): [T, 3, H, W] with H, W multiples of 28."""
        from . import vision
        d = self.dims
        pix, grids, counts = [], [], []
        for fr in items:
            fr = torch.as_tensor(fr)
            if fr.dim() == 3:
                fr = fr[None]
            if not is_video:                       # an image is one temporal patch: the frame repeated t_patch times
                fr = fr[:1].repeat(d.t_patch, 1, 1, 1)
            fr = fr.to(self.device)
            if fr.dtype not in (torch.uint8, torch.float32):
                fr = fr.float()
            H, W = fr.shape[-2:]
            if H % (d.patch * d.merge) or W % (d.patch * d.merge):
                h, w = vision.smart_resize(H, W, min_pixels=self.image_processor.min_pixels,
                                           max_pixels=self.image_processor.max_pixels)
                fr = vision.resize_frames(fr, h, w)
            _, p32, grid = vision.patchify(fr, want_f32=True, want_bf16=False)
            pix.append(p32)
            grids.append(grid)
            counts.append(int(grid.prod()) // (d.merge * d.merge))
        return torch.cat(pix), torch.cat(grids), counts

    def __call__(self, text=None, images=None, videos=None, return_tensors="pt", padding=True, padding_side="left",
                 add_special_tokens=False, second_per_grid_ts=None, **unused):
        """TRN:417-425.  Every <|video_pad|> / <|image_pad|> of the text is expanded to as many placeholders as the visual
        input has merged tokens (processing_qwen2_vl.py), then tokenised and padded on `padding_side`."""
        texts = [text] if isinstance(text, str) else list(text)
        out = {}
        v_counts, i_counts = [], []
        if videos is not None:
            out["pixel_values_videos"], out["video_grid_thw"], v_counts = self._visual(videos, True)
            if self.dims.variant == "qwen2_5_vl":
                out["second_per_grid_ts"] = list(second_per_grid_ts) if second_per_grid_ts is not None else \
                    [self.dims.t_patch / 2.0] * len(v_counts)
        if images is not None:
            out["pixel_values"], out["image_grid_thw"], i_counts = self._visual(images, False)
        vi = ii = 0
        rows = []
        for t in texts:
            parts = []
            for seg in _split_keep(t, (VIDEO_PAD, IMAGE_PAD)):
                if seg == VIDEO_PAD:
                    if vi >= len(v_counts):
                        raise SpacerError("more <|video_pad|> placeholders than videos")
                    parts.append(VIDEO_PAD * v_counts[vi])
                    vi += 1
                elif seg == IMAGE_PAD:
                    if ii >= len(i_counts):
                        raise SpacerError("more <|image_pad|> placeholders than images")
                    parts.append(IMAGE_PAD * i_counts[ii])
                    ii += 1
                else:
                    parts.append(seg)
            rows.append(self.tokenizer.encode("".join(parts)))
        if vi != len(v_counts) or ii != len(i_counts):
            raise SpacerError("visual inputs without a placeholder in the text")
        out.update(pad_batch(rows, self.pad_token_id, padding_side))
        return _Batch(out)

    def batch_decode(self, sequences, skip_special_tokens: bool = True, **kw):
        return self.tokenizer.batch_decode(sequences, skip_special_tokens=skip_special_tokens)

    def decode(self, ids, skip_special_tokens: bool = True, **kw):
        return self.tokenizer.decode(ids, skip_special_tokens=skip_special_tokens)


class _Batch(dict):
    """dict with `.to(device)` and attribute access, like transformers' BatchFeature (SpaceR-Eval: `processor(...).to(device)`,
    `inputs_batch.input_ids`)."""

    def to(self, device):
        for k, v in self.items():
            if torch.is_tensor(v):
                self[k] = v.to(device)
        return self

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _split_keep(text: str, seps) -> list:
    out, pos = [], 0
    while pos < len(text):
        nxt, which = len(text), None
        for s in seps:
            j = text.find(s, pos)
            if j != -1 and j < nxt:
                nxt, which = j, s
        if nxt > pos:
            out.append(text[pos:nxt])
        if which is None:
            break
        out.append(which)
        pos = nxt + len(which)
    return out

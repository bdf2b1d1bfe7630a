"""Flat parameter arenas and the HF <-> internal layout mapping.

Internal layout (what the kernels want):
  * every weight matrix lives in one bf16 "matrix" arena, every norm weight / bias in a bf16 "vector" arena
    (AdamW applies weight decay to the former only, like HF Trainer's decay-parameter split);
  * q/k/v projections are one [q|k|v] matrix; gate/up are one matrix whose rows interleave
    [64 gate | 64 up] so the GEMM epilogue can apply SwiGLU per tile;
  * tensors start on 256-byte boundaries (TMA needs 16).
state_dict()/load_state_dict() speak the HF names of Qwen2VLForConditionalGeneration (checkpoint save at
open_r1/SG-RLVR.py:384).
"""
from __future__ import annotations

import torch

from .config import ModelDims

ALIGN = 128  # elements


def _layout(d: ModelDims):
    """[(internal name, shape, arena)] in arena order."""
    E, M, H, I = d.v_embed, d.v_mlp, d.hidden, d.inter
    v25 = d.variant == "qwen2_5_vl"
    Mp = d.v_mlp_pad
    out = [("v.patch_w", (E, d.patch_dim), "mat")]
    for i in range(d.v_depth):
        p = f"v.{i}."
        if v25:   # RMSNorm, SwiGLU MLP with biases; gate|up interleaved and zero-padded to a multiple of 64 columns
            out += [(p + "ln1_w", (E,), "vec"),
                    (p + "qkv_w", (3 * E, E), "mat"), (p + "qkv_b", (3 * E,), "vec"),
                    (p + "proj_w", (E, E), "mat"), (p + "proj_b", (E,), "vec"),
                    (p + "ln2_w", (E,), "vec"),
                    (p + "gu_w", (2 * Mp, E), "mat"), (p + "gu_b", (2 * Mp,), "vec"),
                    (p + "down_w", (E, Mp), "mat"), (p + "down_b", (E,), "vec")]
            continue
        out += [(p + "ln1_w", (E,), "vec"), (p + "ln1_b", (E,), "vec"),
                (p + "qkv_w", (3 * E, E), "mat"), (p + "qkv_b", (3 * E,), "vec"),
                (p + "proj_w", (E, E), "mat"), (p + "proj_b", (E,), "vec"),
                (p + "ln2_w", (E,), "vec"), (p + "ln2_b", (E,), "vec"),
                (p + "fc1_w", (M, E), "mat"), (p + "fc1_b", (M,), "vec"),
                (p + "fc2_w", (E, M), "mat"), (p + "fc2_b", (E,), "vec")]
    mh = d.merge_dim
    out += [("v.m.ln_w", (E,), "vec")] + ([] if v25 else [("v.m.ln_b", (E,), "vec")])
    out += [("v.m.fc0_w", (mh, mh), "mat"), ("v.m.fc0_b", (mh,), "vec"),
            ("v.m.fc2_w", (H, mh), "mat"), ("v.m.fc2_b", (H,), "vec")]
    out.append(("embed", (d.vocab, H), "mat"))
    for i in range(d.layers):
        p = f"l.{i}."
        out += [(p + "ln1_w", (H,), "vec"),
                (p + "qkv_w", (d.qkv_dim, H), "mat"), (p + "qkv_b", (d.qkv_dim,), "vec"),
                (p + "o_w", (H, d.heads * d.head_dim), "mat"),
                (p + "ln2_w", (H,), "vec"),
                (p + "gu_w", (2 * I, H), "mat"),
                (p + "down_w", (H, I), "mat")]
    out.append(("norm_w", (H,), "vec"))
    if not d.tie:
        out.append(("lm_head", (d.vocab, H), "mat"))
    return out


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class ParamStore:
    """Two flat bf16 arenas with named views."""

    def __init__(self, dims: ModelDims, device="cuda"):
        self.dims = dims
        self.device = torch.device(device)
        self.index = {}   # name -> (arena, offset, shape)
        sizes = {"mat": 0, "vec": 0}
        for name, shape, arena in _layout(dims):
            off = sizes[arena]
            self.index[name] = (arena, off, shape)
            sizes[arena] = off + (_numel(shape) + ALIGN - 1) // ALIGN * ALIGN
        self.sizes = sizes
        self.mat = torch.zeros(sizes["mat"], device=self.device, dtype=torch.bfloat16)
        self.vec = torch.zeros(sizes["vec"], device=self.device, dtype=torch.bfloat16)
        self.views = {n: self.view_of(n, self.mat, self.vec) for n in self.index}
        self.version = 0      # bumped whenever the weights are replaced wholesale (optimizers resync their fp32 masters)

    def view_of(self, name, mat, vec):
        arena, off, shape = self.index[name]
        base = mat if arena == "mat" else vec
        return base[off:off + _numel(shape)].view(shape)

    def __getitem__(self, name):
        if name == "lm_head" and self.dims.tie:
            name = "embed"
        return self.views[name]

    def numel(self):
        return sum(_numel(s) for _, (_, _, s) in self.index.items())

    def clone_like(self, dtype, zero=True):
        """(mat, vec) tensors shaped like the arenas (gradients / optimizer state)."""
        f = torch.zeros if zero else torch.empty
        return (f(self.sizes["mat"], device=self.device, dtype=dtype),
                f(self.sizes["vec"], device=self.device, dtype=dtype))

    # -------------------------------------------------------------------------------------------
    def init_random(self, seed: int = 0, std: float = 0.02):
        """Random-init weights (normal(0, std); norm weights 1) generated on the device."""
        self.version += 1
        g = torch.Generator(device=self.device).manual_seed(seed)
        chunk = 1 << 26
        for base in (self.mat, self.vec):
            for s in range(0, base.numel(), chunk):
                e = min(s + chunk, base.numel())
                base[s:e] = (torch.randn(e - s, generator=g, device=self.device, dtype=torch.float32) * std).to(torch.bfloat16)
        for name in self.index:
            if name.endswith(("ln1_w", "ln2_w", "ln_w", "norm_w")):
                self.views[name].fill_(1.0)
        if self.dims.variant == "qwen2_5_vl" and self.dims.v_mlp_pad != self.dims.v_mlp:
            M, Mp = self.dims.v_mlp, self.dims.v_mlp_pad
            for i in range(self.dims.v_depth):     # the padded SwiGLU columns are structurally zero
                p = f"v.{i}."
                g, u = self._deinterleave(self.views[p + "gu_w"])
                g[M:] = 0; u[M:] = 0
                self.views[p + "gu_w"].copy_(self._interleave(g, u))
                gb, ub = self._deinterleave(self.views[p + "gu_b"].view(-1, 1))
                gb[M:] = 0; ub[M:] = 0
                self.views[p + "gu_b"].copy_(self._interleave(gb, ub).view(-1))
                self.views[p + "down_w"][:, M:] = 0

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def _interleave(gate, up):
        I, H = gate.shape
        return torch.stack([gate.view(I // 64, 64, H), up.view(I // 64, 64, H)], dim=1).reshape(2 * I, H)

    @staticmethod
    def _deinterleave(gu):
        I2, H = gu.shape
        t = gu.view(I2 // 128, 2, 64, H)
        return t[:, 0].reshape(I2 // 2, H), t[:, 1].reshape(I2 // 2, H)

    def hf_items(self, views=None):
        """Yield (hf_name, tensor) converting the internal layout back to HF names."""
        d = self.dims
        V = views if views is not None else self.views
        yield "model.visual.patch_embed.proj.weight", V["v.patch_w"].view(d.v_embed, d.in_ch, d.t_patch, d.patch, d.patch)
        v25 = d.variant == "qwen2_5_vl"
        hf_v = {"ln1_w": "norm1.weight", "ln1_b": "norm1.bias", "qkv_w": "attn.qkv.weight", "qkv_b": "attn.qkv.bias",
                "proj_w": "attn.proj.weight", "proj_b": "attn.proj.bias", "ln2_w": "norm2.weight",
                "ln2_b": "norm2.bias", "fc1_w": "mlp.fc1.weight", "fc1_b": "mlp.fc1.bias",
                "fc2_w": "mlp.fc2.weight", "fc2_b": "mlp.fc2.bias"}
        if v25:
            hf_v = {k: hf_v[k] for k in ("ln1_w", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w")}
        for i in range(d.v_depth):
            for k, hk in hf_v.items():
                yield f"model.visual.blocks.{i}.{hk}", V[f"v.{i}.{k}"]
            if v25:
                M = d.v_mlp
                b, p = f"model.visual.blocks.{i}.mlp.", f"v.{i}."
                g, u = self._deinterleave(V[p + "gu_w"])
                gb, ub = self._deinterleave(V[p + "gu_b"].view(-1, 1))
                yield b + "gate_proj.weight", g[:M]
                yield b + "gate_proj.bias", gb[:M].reshape(-1)
                yield b + "up_proj.weight", u[:M]
                yield b + "up_proj.bias", ub[:M].reshape(-1)
                yield b + "down_proj.weight", V[p + "down_w"][:, :M]
                yield b + "down_proj.bias", V[p + "down_b"]
        merger = {"ln_w": "ln_q.weight", "ln_b": "ln_q.bias", "fc0_w": "mlp.0.weight", "fc0_b": "mlp.0.bias",
                  "fc2_w": "mlp.2.weight", "fc2_b": "mlp.2.bias"}
        if v25:
            merger.pop("ln_b")
        for k, hk in merger.items():
            yield f"model.visual.merger.{hk}", V[f"v.m.{k}"]
        L = "model.language_model."
        yield L + "embed_tokens.weight", V["embed"]
        nq, nk = d.heads * d.head_dim, d.kv_heads * d.head_dim
        for i in range(d.layers):
            b, p = f"{L}layers.{i}.", f"l.{i}."
            yield b + "input_layernorm.weight", V[p + "ln1_w"]
            w, bi = V[p + "qkv_w"], V[p + "qkv_b"]
            yield b + "self_attn.q_proj.weight", w[:nq]
            yield b + "self_attn.k_proj.weight", w[nq:nq + nk]
            yield b + "self_attn.v_proj.weight", w[nq + nk:]
            yield b + "self_attn.q_proj.bias", bi[:nq]
            yield b + "self_attn.k_proj.bias", bi[nq:nq + nk]
            yield b + "self_attn.v_proj.bias", bi[nq + nk:]
            yield b + "self_attn.o_proj.weight", V[p + "o_w"]
            yield b + "post_attention_layernorm.weight", V[p + "ln2_w"]
            g, u = self._deinterleave(V[p + "gu_w"])
            yield b + "mlp.gate_proj.weight", g
            yield b + "mlp.up_proj.weight", u
            yield b + "mlp.down_proj.weight", V[p + "down_w"]
        yield L + "norm.weight", V["norm_w"]
        yield "lm_head.weight", V["embed"] if d.tie else V["lm_head"]

    def state_dict(self, clone: bool = True):
        """HF-named tensors.  clone=False returns views of the arenas wherever the HF tensor is one (everything but the
        de-interleaved gate/up projections), like `nn.Module.state_dict()` returns references.  The data pointers of the
        last export are remembered so that a weight sync that hands them straight back (vllm_api.load_weights on the
        colocated engine) is recognised as a no-op."""
        sd = {k: (v.clone() if clone else v) for k, v in self.hf_items()}
        self._export_ptrs = {k: v.data_ptr() for k, v in sd.items()}
        return sd

    def load_state_dict(self, sd: dict):
        """Load an HF-named state dict (any float dtype, any device)."""
        self.version += 1
        d = self.dims
        nq, nk = d.heads * d.head_dim, d.kv_heads * d.head_dim

        def put(name, t):
            self.views[name].copy_(t.to(self.device).reshape(self.views[name].shape))

        def get(k):
            if k not in sd:
                raise KeyError(f"missing parameter {k}")
            return sd[k]

        put("v.patch_w", get("model.visual.patch_embed.proj.weight"))
        v25 = d.variant == "qwen2_5_vl"
        for i in range(d.v_depth if v25 else 0):
            b, p = f"model.visual.blocks.{i}.", f"v.{i}."
            for k, hk in [("ln1_w", "norm1.weight"), ("qkv_w", "attn.qkv.weight"), ("qkv_b", "attn.qkv.bias"),
                          ("proj_w", "attn.proj.weight"), ("proj_b", "attn.proj.bias"), ("ln2_w", "norm2.weight"),
                          ("down_b", "mlp.down_proj.bias")]:
                put(p + k, get(b + hk))
            M, Mp, E = d.v_mlp, d.v_mlp_pad, d.v_embed

            def padr(t, rows):   # zero-pad rows (dim 0)
                t = t.to(self.device, torch.bfloat16)
                out = torch.zeros((rows,) + tuple(t.shape[1:]), device=self.device, dtype=torch.bfloat16)
                out[:t.shape[0]] = t
                return out
            put(p + "gu_w", self._interleave(padr(get(b + "mlp.gate_proj.weight"), Mp), padr(get(b + "mlp.up_proj.weight"), Mp)))
            put(p + "gu_b", self._interleave(padr(get(b + "mlp.gate_proj.bias").reshape(-1, 1), Mp),
                                             padr(get(b + "mlp.up_proj.bias").reshape(-1, 1), Mp)).reshape(-1))
            dw = torch.zeros((E, Mp), device=self.device, dtype=torch.bfloat16)
            dw[:, :M] = get(b + "mlp.down_proj.weight").to(self.device, torch.bfloat16)
            put(p + "down_w", dw)
        for i in range(0 if v25 else d.v_depth):
            b, p = f"model.visual.blocks.{i}.", f"v.{i}."
            for k, hk in [("ln1_w", "norm1.weight"), ("ln1_b", "norm1.bias"), ("qkv_w", "attn.qkv.weight"),
                          ("qkv_b", "attn.qkv.bias"), ("proj_w", "attn.proj.weight"), ("proj_b", "attn.proj.bias"),
                          ("ln2_w", "norm2.weight"), ("ln2_b", "norm2.bias"), ("fc1_w", "mlp.fc1.weight"),
                          ("fc1_b", "mlp.fc1.bias"), ("fc2_w", "mlp.fc2.weight"), ("fc2_b", "mlp.fc2.bias")]:
                put(p + k, get(b + hk))
        for k, hk in [("ln_w", "ln_q.weight"), ("ln_b", "ln_q.bias"), ("fc0_w", "mlp.0.weight"), ("fc0_b", "mlp.0.bias"),
                      ("fc2_w", "mlp.2.weight"), ("fc2_b", "mlp.2.bias")]:
            if v25 and k == "ln_b":
                continue
            put("v.m." + k, get("model.visual.merger." + hk))
        L = "model.language_model."
        put("embed", get(L + "embed_tokens.weight"))
        for i in range(d.layers):
            b, p = f"{L}layers.{i}.", f"l.{i}."
            put(p + "ln1_w", get(b + "input_layernorm.weight"))
            put(p + "qkv_w", torch.cat([get(b + f"self_attn.{x}_proj.weight") for x in "qkv"], 0))
            put(p + "qkv_b", torch.cat([get(b + f"self_attn.{x}_proj.bias") for x in "qkv"], 0))
            put(p + "o_w", get(b + "self_attn.o_proj.weight"))
            put(p + "ln2_w", get(b + "post_attention_layernorm.weight"))
            put(p + "gu_w", self._interleave(get(b + "mlp.gate_proj.weight"), get(b + "mlp.up_proj.weight")))
            put(p + "down_w", get(b + "mlp.down_proj.weight"))
        put("norm_w", get(L + "norm.weight"))
        if not d.tie:
            put("lm_head", get("lm_head.weight"))

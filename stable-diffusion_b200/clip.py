"""B200-native CLIP ViT-L/14 text encoder: drop-in for `ldm.modules.encoders.modules.FrozenCLIPEmbedder`
(modules.py:137-162: `encode(text) -> (B, 77, 768)` = CLIPTextModel(...).last_hidden_state).

The arithmetic the reference delegates to third-party `transformers` (CLIPTextModel: token+position embedding,
12 pre-LN causal self-attention layers with quick-GELU MLP, final LayerNorm) runs on the same kernel family as the
UNet: fp32 LayerNorm -> fp16 operand, tcgen05 GEMMs with fused bias / quick-GELU / residual epilogues, the fused
tcgen05 attention kernel with a causal mask. The value bias is folded through out_proj at pack time
(softmax rows sum to 1). State-dict keys are transformers' (`text_model.*`), found under
`cond_stage_model.transformer.` in an SD checkpoint.

Tokenisation stays on the host with HF's CLIPTokenizer when its vocabulary files are available; there is no
vocabulary in this offline image, so `encode_ids` takes token ids directly (tests and the benchmark use seeded ids).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops
from .arch import SD_V1_CLIP, clip_param_shapes
from .ops import ACT_QUICK_GELU
from .util import adopt_state_dict


class _Transformer(nn.Module):
    """Holds the `text_model.*` keys so the checkpoint prefix `cond_stage_model.transformer.` resolves."""

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        o = self._owner
        # text_model.embeddings.position_ids is a buffer of older transformers versions (present in sd-v1 checkpoints)
        sd = adopt_state_dict(o, state_dict, prefix, missing_keys, unexpected_keys, error_msgs,
                              ignore=("text_model.embeddings.position_ids",))
        if sd is None:
            return
        o._host_sd = sd
        if o.W is not None:
            o.pack_weights(o.W["device"])


class FrozenCLIPEmbedder(nn.Module):
    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, config=None):
        super().__init__()
        self.cfg = dict(config or SD_V1_CLIP)
        self.version = version
        self.device = device
        self.max_length = max_length
        self.shapes = clip_param_shapes(self.cfg)
        self.W = None
        self._host_sd = None
        self.transformer = _Transformer(self)
        self.tokenizer = None

    def freeze(self):
        return self

    def load_weights(self, sd, device):
        for k, shape in self.shapes.items():
            assert k in sd and tuple(sd[k].shape) == tuple(shape), k
        self._host_sd = {k: sd[k] for k in self.shapes}
        self.pack_weights(torch.device(device))
        return self

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        probe = fn(torch.empty(0))
        if probe.is_cuda and self._host_sd is not None and (self.W is None or self.W["device"] != probe.device):
            self.pack_weights(probe.device)
        return r

    @torch.no_grad()
    def pack_weights(self, device):
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in self._host_sd.items()}
        f32 = lambda k: sd[k].contiguous()
        f16 = lambda k: sd[k].half().contiguous()
        h = self.cfg["hidden_size"]
        W = {"device": device, "tok": f32("text_model.embeddings.token_embedding.weight"),
             "pos": f32("text_model.embeddings.position_embedding.weight"), "layers": []}
        for i in range(self.cfg["num_hidden_layers"]):
            p = f"text_model.encoder.layers.{i}"
            wo = sd[p + ".self_attn.out_proj.weight"]
            L = {"ln1": (f32(p + ".layer_norm1.weight"), f32(p + ".layer_norm1.bias")),
                 "ln2": (f32(p + ".layer_norm2.weight"), f32(p + ".layer_norm2.bias")),
                 "w_qk": torch.cat([sd[p + ".self_attn.q_proj.weight"], sd[p + ".self_attn.k_proj.weight"]],
                                   0).half().contiguous(),
                 "b_qk": torch.cat([sd[p + ".self_attn.q_proj.bias"], sd[p + ".self_attn.k_proj.bias"]]).contiguous(),
                 "w_v": f16(p + ".self_attn.v_proj.weight"),
                 "w_o": wo.half().contiguous(),
                 "b_o": (sd[p + ".self_attn.out_proj.bias"] + wo @ sd[p + ".self_attn.v_proj.bias"]).contiguous(),
                 "w_fc1": f16(p + ".mlp.fc1.weight"), "b_fc1": f32(p + ".mlp.fc1.bias"),
                 "w_fc2": f16(p + ".mlp.fc2.weight"), "b_fc2": f32(p + ".mlp.fc2.bias")}
            W["layers"].append(L)
        W["ln_f"] = (f32("text_model.final_layer_norm.weight"), f32("text_model.final_layer_norm.bias"))
        self.W = W

    @torch.no_grad()
    def encode_ids(self, ids):
        """ids: int64 [B, n<=77] on the GPU -> fp32 [B, n, hidden]."""
        assert self.W is not None and ids.is_cuda, "sdb200.FrozenCLIPEmbedder runs on CUDA only (no CPU fallback)"
        W = self.W
        B, n = ids.shape
        h = self.cfg["hidden_size"]
        heads = self.cfg["num_attention_heads"]
        d = h // heads
        assert d == 64, "CLIP text heads are 64 wide"
        eps = self.cfg["layer_norm_eps"]
        x = ops.embed_tokens(ids.contiguous(), W["tok"], W["pos"])                     # [B*n, h] fp32
        for L in W["layers"]:
            y = ops.layernorm(x, *L["ln1"], eps=eps)
            qk, _ = ops.gemm(y, L["w_qk"], bias=L["b_qk"], want_f16=True)             # [B*n, 2h]
            v, _ = ops.gemm(y, L["w_v"], want_f16=True)                               # bias folded into b_o
            qk3 = qk.view(B, n, 2 * h)
            vt = ops.transpose_f16(v.view(B, n, h))                                   # [B, h, pad8(n)]
            o = ops.attention(qk3[:, :, :h], qk3[:, :, h:], vt, heads=heads, d=d, dpad=d, nq=n, nkv=n,
                              scale=d ** -0.5, causal=True)
            _, x = ops.gemm(o.view(-1, h), L["w_o"], bias=L["b_o"], residual=x, want_f32=True)
            y = ops.layernorm(x, *L["ln2"], eps=eps)
            g, _ = ops.gemm(y, L["w_fc1"], bias=L["b_fc1"], act=ACT_QUICK_GELU, want_f16=True)
            _, x = ops.gemm(g, L["w_fc2"], bias=L["b_fc2"], residual=x, want_f32=True)
        _, z = ops.layernorm(x, *W["ln_f"], eps=eps, want_f32=True)
        return z.view(B, n, h)

    def _tokenize(self, text):
        """list[str] -> int64 [B, max_length] ids as `modules.py:153-156` (truncation, max_length padding).
        `version` may be a directory holding vocab.json + merges.txt (host-side BPE in tokenizer.py, no third-party
        code); otherwise the locally cached transformers tokenizer of that name is used, as the reference does."""
        if self.tokenizer is None:
            if os.path.isdir(str(self.version)) and os.path.exists(os.path.join(str(self.version), "vocab.json")):
                from .tokenizer import CLIPBPETokenizer
                self.tokenizer = CLIPBPETokenizer.from_dir(str(self.version), self.max_length)
            else:
                try:
                    from transformers import CLIPTokenizer
                    tok = CLIPTokenizer.from_pretrained(self.version, local_files_only=True)
                    if len(tok) < 49408:  # transformers >= 5 silently builds an empty tokenizer when files are missing
                        raise FileNotFoundError(f"vocabulary has {len(tok)} entries")
                    self.tokenizer = tok
                except Exception as e:  # no vocabulary offline
                    raise RuntimeError(f"CLIP tokenizer files for {self.version} are not available offline ({e}); pass "
                                       "a directory with vocab.json + merges.txt as `version`, or token ids to "
                                       "encode_ids()") from e
        if isinstance(text, str):
            text = [text]
        from .tokenizer import CLIPBPETokenizer
        if isinstance(self.tokenizer, CLIPBPETokenizer):
            return torch.tensor(self.tokenizer(list(text), self.max_length), dtype=torch.long)
        batch = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                               return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return batch["input_ids"]

    def forward(self, text):
        if torch.is_tensor(text):
            ids = text
        else:
            ids = self._tokenize(text)
        return self.encode_ids(ids.to(self.W["device"]).long())

    def encode(self, text):
        return self(text)
